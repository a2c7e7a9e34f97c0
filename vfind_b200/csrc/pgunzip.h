// Parallel decoding of one gzip stream on host threads (see pgunzip.cu).
#pragma once
#include <stdint.h>

#include <cstdio>
#include <string>

namespace vfb {

class ParallelGunzip {
  public:
    ParallelGunzip();
    ~ParallelGunzip();
    ParallelGunzip(const ParallelGunzip &) = delete;
    ParallelGunzip &operator=(const ParallelGunzip &) = delete;
    // `f` is positioned at a gzip member header (or at end of file); MultiGzDecoder semantics.
    void init(FILE *f, int threads);
    // Fills out[0..cap) with text; returns the bytes produced (0 at the end of the stream), -1 on error.
    long long read(uint8_t *out, size_t cap, std::string *err);
    // The same, with the copy split over the decoder's threads and the '\n' bytes of what was copied counted on the way.
    long long read_counting(uint8_t *out, size_t cap, size_t *newlines, std::string *err);
    void stats(uint64_t *segments, uint64_t *workers_used, uint64_t *workers_dropped) const;

  private:
    struct Impl;
    Impl *impl_;
};

}  // namespace vfb
