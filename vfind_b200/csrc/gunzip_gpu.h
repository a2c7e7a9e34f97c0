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
    // Returns the bytes produced, -1 on error.  Less than cap: the member has ended — see handover() — or, with
    // stop_at_segment, the current device segment has (member_ended() tells which).
    long long read_counting(uint8_t *out, size_t cap, size_t *newlines, std::string *err, bool stop_at_segment = false);
    bool member_ended() const;
    // The same text WITHOUT leaving the device: a range of whole 64 KiB pieces of the current segment (at most max_len bytes;
    // the segment's last keep_tail bytes are left to read_counting, which also handles everything near the stream's end).
    // peek_device does not move the read position — commit_device does, release_device gives the range back once the
    // caller's copy out of it has completed.  false: nothing to hand out this way right now (call read_counting).
    struct DeviceRange {
        const uint8_t *d_ptr = nullptr;
        size_t len = 0, newlines = 0;
        int device = -1;
        uint64_t seq = 0;          // the segment it belongs to
    };
    bool peek_device(size_t max_len, size_t keep_tail, DeviceRange *out);
    bool range_tail(const DeviceRange &r, uint8_t *out_host, size_t n, std::string *err);      // its last n bytes
    void commit_device(const DeviceRange &r);
    void release_device(const DeviceRange &r);
    // After the member's trailer has been checked: the file offset right behind it (further members, or the end of the file).
    bool handover(long *file_off) const;
    void stats(uint64_t *segments, uint64_t *chunks, uint64_t *visited, uint64_t *host_bytes) const;

    struct Impl;

  private:
    Impl *impl_;
};

}  // namespace vfb
