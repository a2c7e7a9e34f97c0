// Single-stream gzip decoded on the device (see gunzip_gpu.cu / kernels_inflate.cu).
#pragma once
#include <stdint.h>

#include <cstdio>
#include <string>

namespace vfb {

class GpuGunzip {
  public:
    GpuGunzip();
    ~GpuGunzip();
    GpuGunzip(const GpuGunzip &) = delete;
    GpuGunzip &operator=(const GpuGunzip &) = delete;
    // `f` is positioned at a gzip member header.  false: the decoder cannot be used here (the caller keeps the host path).
    bool init(FILE *f, int device, std::string *err);
    // Fills out[0..cap) with the member's text (out should be pinned); the '\n' bytes of what was produced are counted.
    // Returns the bytes produced, -1 on error.  Less than cap: the member has ended — see handover().
    long long read_counting(uint8_t *out, size_t cap, size_t *newlines, std::string *err);
    // After the member's trailer has been checked: the file offset right behind it (further members, or the end of the file).
    bool handover(long *file_off) const;
    void stats(uint64_t *segments, uint64_t *chunks, uint64_t *visited, uint64_t *host_bytes) const;

    struct Impl;

  private:
    Impl *impl_;
};

}  // namespace vfb
