// K1 — exact adapter scan, and the DP worklist builder.
//
// Replaces the two `memmem::find(seq, adapter)` calls per read of
// /root/reference/src/lib.rs:148 (called from :278-286): leftmost byte-exact, case-sensitive
// occurrence of each adapter over the whole read.  Output is the region boundary the exact
// hit implies (:151-152): start = pos + A for the prefix, end = pos for the suffix, VFB_NONE
// when there is no exact hit (the DP kernel may fill it in later).
#include "vfb_internal.cuh"

namespace vfb {

struct ScanArgs {
    ScanJob job;
    AdapterBytes prefix, suffix;
};

#define SCAN_THREADS 256

// One warp per read; lanes test consecutive start positions and vote.
__device__ __forceinline__ uint32_t scan_one(const uint8_t *seq, uint32_t L, const uint8_t *ad,
                                            uint32_t A, int lane)
{
    if (A == 0) return 0;              // an empty needle matches at 0 (memchr convention)
    if (A > L) return VFB_NONE;
    const uint32_t last = L - A;       // last candidate start
    const uint8_t a0 = ad[0];
    for (uint32_t base = 0; base <= last; base += 32) {
        const uint32_t p = base + lane;
        bool hit = false;
        if (p <= last && __ldg(seq + p) == a0) {
            hit = true;
            for (uint32_t t = 1; t < A; ++t) {
                if (__ldg(seq + p + t) != ad[t]) { hit = false; break; }
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (m) return base + (uint32_t)(__ffs((int)m) - 1);
    }
    return VFB_NONE;
}

__global__ void __launch_bounds__(SCAN_THREADS)
k1_scan(const __grid_constant__ ScanArgs args)
{
    __shared__ uint8_t s_pre[VFB_MAX_SCAN_ADAPTER], s_suf[VFB_MAX_SCAN_ADAPTER];
    for (uint32_t i = threadIdx.x; i < args.prefix.len; i += blockDim.x) s_pre[i] = args.prefix.b[i];
    for (uint32_t i = threadIdx.x; i < args.suffix.len; i += blockDim.x) s_suf[i] = args.suffix.b[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t warps_total = gridDim.x * (SCAN_THREADS / 32);
    for (uint32_t r = blockIdx.x * (SCAN_THREADS / 32) + (threadIdx.x >> 5); r < args.job.n_reads;
         r += warps_total) {
        const vfb_span sp = args.job.spans[r];
        const uint8_t *seq = args.job.text + sp.off;
        const uint32_t pp = scan_one(seq, sp.len, s_pre, args.prefix.len, lane);
        const uint32_t ps = scan_one(seq, sp.len, s_suf, args.suffix.len, lane);
        if (lane == 0) {
            args.job.start[r] = pp == VFB_NONE ? VFB_NONE : pp + args.prefix.len;
            args.job.end[r] = ps;
        }
    }
}

int launch_scan(const ScanJob &job, const AdapterBytes &prefix, const AdapterBytes &suffix,
                int sm_count, cudaStream_t st)
{
    if (job.n_reads == 0) return VFB_OK;
    ScanArgs a;
    a.job = job;
    a.prefix = prefix;
    a.suffix = suffix;
    uint32_t blocks = (job.n_reads + (SCAN_THREADS / 32) - 1) / (SCAN_THREADS / 32);
    uint32_t cap = (uint32_t)sm_count * 8u;
    if (blocks > cap) blocks = cap;
    k1_scan<<<blocks, SCAN_THREADS, 0, st>>>(a);
    ++g_launches;
    VFB_CUDA(cudaGetLastError());
    return VFB_OK;
}

// Worklist: warp-aggregated append of the reads that still need an alignment.
__global__ void __launch_bounds__(256)
k_worklist(const uint32_t *__restrict__ bound, const uint32_t *__restrict__ require,
           const vfb_span *__restrict__ spans, uint32_t n, uint32_t *__restrict__ list,
           uint32_t *__restrict__ count)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += stride) {
        const uint32_t i = base + lane;
        bool need = false;
        if (i < n) {
            need = bound[i] == VFB_NONE && spans[i].len > 0;
            if (need && require) need = require[i] != VFB_NONE;
        }
        const unsigned m = __ballot_sync(0xffffffffu, need);
        if (m) {
            uint32_t pos = 0;
            if (lane == 0) pos = atomicAdd(count, (uint32_t)__popc(m));
            pos = __shfl_sync(0xffffffffu, pos, 0);
            if (need) list[pos + __popc(m & ((1u << lane) - 1))] = i;
        }
    }
}

int launch_worklist(const uint32_t *bound, const uint32_t *require, const vfb_span *spans,
                    uint32_t n_reads, uint32_t *list, uint32_t *count, cudaStream_t st)
{
    if (n_reads == 0) return VFB_OK;
    uint32_t blocks = (n_reads + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_worklist<<<blocks, 256, 0, st>>>(bound, require, spans, n_reads, list, count);
    ++g_launches;
    VFB_CUDA(cudaGetLastError());
    return VFB_OK;
}

}  // namespace vfb
