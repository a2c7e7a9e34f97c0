// Device-side synthetic read generator and the integer-pipe peak microbenchmark.
#include "synth.h"
#include "vfb_internal.cuh"

namespace vfb {

struct SynthArgs {
    vfb_synth_cfg cfg;
    uint64_t first, n;
    uint8_t *text;
    vfb_span *spans;
    uint8_t prefix[64], suffix[64];
};

// One thread per read, staged through shared memory so that global stores are coalesced.
#define SYNTH_THREADS 64
__global__ void __launch_bounds__(SYNTH_THREADS)
k_synth(const __grid_constant__ SynthArgs a)
{
    extern __shared__ uint8_t stage[];   // SYNTH_THREADS * L
    const uint32_t L = a.cfg.read_len;
    const uint64_t blocks_total = (a.n + SYNTH_THREADS - 1) / SYNTH_THREADS;
    for (uint64_t blk = blockIdx.x; blk < blocks_total; blk += gridDim.x) {
        const uint64_t i0 = blk * SYNTH_THREADS;
        const uint64_t i = i0 + threadIdx.x;
        __syncthreads();
        if (i < a.n) {
            vfs_read(&a.cfg, a.first + i, a.prefix, a.suffix, stage + (size_t)threadIdx.x * L);
            a.spans[i] = vfb_span{(uint32_t)(i * L), L};
        }
        __syncthreads();
        const uint64_t cnt = (a.n - i0 < SYNTH_THREADS ? a.n - i0 : SYNTH_THREADS) * (uint64_t)L;
        uint8_t *dst = a.text + i0 * L;
        for (uint64_t b = threadIdx.x; b < cnt; b += SYNTH_THREADS) dst[b] = stage[b];
    }
}

int launch_synth(const vfb_synth_cfg &cfg, uint64_t first, uint64_t n, uint8_t *d_text,
                 vfb_span *d_spans, cudaStream_t st)
{
    if (n == 0) return VFB_OK;
    if (cfg.adapter_len > 61 || cfg.adapter_len == 0 || cfg.read_len == 0 || cfg.read_len > 700) {
        set_error("synth: adapter_len must be 1..61 and read_len 1..700");
        return VFB_ERR_ARG;
    }
    if (n * (uint64_t)cfg.read_len > 0xFFFFFFFFull) {
        set_error("synth: one call generates at most 4 GiB of text");
        return VFB_ERR_ARG;
    }
    SynthArgs a;
    a.cfg = cfg; a.first = first; a.n = n; a.text = d_text; a.spans = d_spans;
    vfs_adapter(cfg.seed, 0, cfg.adapter_len, a.prefix);
    vfs_adapter(cfg.seed, 1, cfg.adapter_len, a.suffix);
    uint64_t blocks = (n + SYNTH_THREADS - 1) / SYNTH_THREADS;
    if (blocks > 148 * 16) blocks = 148 * 16;
    size_t smem = (size_t)SYNTH_THREADS * cfg.read_len;
    {
        // (per device: set every time, it costs nothing)
        VFB_CUDA(cudaFuncSetAttribute(k_synth, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    }
    k_synth<<<(uint32_t)blocks, SYNTH_THREADS, smem, st>>>(a);
    ++g_launches;
    VFB_CUDA(cudaGetLastError());
    return VFB_OK;
}

// ------------------------------------------------------------------------------------
// vfb_submit_device: the caller's spans are validated and measured on the device.
//   out[0] += sum of the span lengths (bounds the key bytes of the batch: spans may overlap or repeat)
//   out[1] += spans that do not lie inside the text buffer
__global__ void __launch_bounds__(256)
k_span_check(const vfb_span *__restrict__ spans, uint64_t n, uint64_t text_bytes, unsigned long long *out)
{
    unsigned long long sum = 0, bad = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const vfb_span s = spans[i];
        sum += s.len;
        bad += ((uint64_t)s.off + s.len > text_bytes) ? 1ull : 0ull;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        bad += __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (sum) atomicAdd(out, sum);
        if (bad) atomicAdd(out + 1, bad);
    }
}

int launch_span_check(const vfb_span *d_spans, uint64_t n, uint64_t text_bytes, unsigned long long *d_out, cudaStream_t st)
{
    if (n == 0) return VFB_OK;
    uint64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_span_check<<<(uint32_t)blocks, 256, 0, st>>>(d_spans, n, text_bytes, d_out);
    ++g_launches;
    VFB_CUDA(cudaGetLastError());
    return VFB_OK;
}

// ------------------------------------------------------------------------------------
// Integer peak: 8 independent chains per thread.
//   MODE 0: VIADDMNMX + IADD3 only (ALU pipe)      -> alu lane-ops/s
//   MODE 1: alternating IADD3/VIADDMNMX and IMAD   -> ALU + FMA pipe dual issue
template <int MODE>
__global__ void __launch_bounds__(256)
k_int_peak(int *out, int iters, int one, int k)
{
    int v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = threadIdx.x + c;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if (MODE == 0) {
                    v[c] = __viaddmax_s32(v[c], k, v[(c + 1) & 7]);
                } else {
                    if (c & 1) v[c] = v[c] * one + k;                       // IMAD (FMA pipe)
                    else v[c] = __viaddmax_s32(v[c], k, v[(c + 2) & 7]);   // ALU pipe
                }
            }
        }
    }
    int s = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) s ^= v[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int measure_int_peak(int device, double *alu_gops, double *dual_gops)
{
    if (device >= 0) VFB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    int dev = 0;
    VFB_CUDA(cudaGetDevice(&dev));
    VFB_CUDA(cudaGetDeviceProperties(&prop, dev));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    int *out = nullptr;
    VFB_CUDA(cudaMalloc(&out, (size_t)blocks * threads * sizeof(int)));
    cudaEvent_t e0, e1;
    VFB_CUDA(cudaEventCreate(&e0));
    VFB_CUDA(cudaEventCreate(&e1));
    double res[2] = {0, 0};
    for (int mode = 0; mode < 2; ++mode) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            VFB_CUDA(cudaEventRecord(e0));
            if (mode == 0) k_int_peak<0><<<blocks, threads>>>(out, iters, 1, 3);
            else k_int_peak<1><<<blocks, threads>>>(out, iters, 1, 3);
            ++g_launches;
            VFB_CUDA(cudaEventRecord(e1));
            VFB_CUDA(cudaEventSynchronize(e1));
            float ms = 0;
            VFB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        const double ops = (double)blocks * threads * (double)iters * 64.0;
        res[mode] = ops / (best * 1e-3) / 1e9;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (alu_gops) *alu_gops = res[0];
    if (dual_gops) *dual_gops = res[1];
    return VFB_OK;
}

}  // namespace vfb
