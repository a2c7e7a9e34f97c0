// FASTQ record parser on the GPU (ingest).
//
// Replaces what seq_io's fastq::Reader does for /root/reference/src/lib.rs:234 and `record.seq()`
// (:277): the host only inflates and cuts the text at a 4-line boundary; the device finds every
// newline, frames records, trims "\r", checks the '@' / '+' markers and that sequence and
// quality have the same length, and emits one span per record's sequence line.
//   pass 1  k_parse_count : newlines per 4 KiB tile
//   scan    k_parse_scan  : exclusive scan of the tile counts (one block)
//   pass 2  k_parse_index : position of newline #k -> line_end[k]
//   pass 3  k_parse_spans : record r = lines 4r..4r+3 -> span + validation
#include "vfb_internal.cuh"

namespace vfb {

#define PARSE_THREADS 256
#define PARSE_TILE (PARSE_THREADS * 16)

__device__ __forceinline__ uint32_t nl_mask16(const uint4 v, uint32_t valid)
{
    // bit b set iff byte b of the 16 is '\n' (and b < valid)
    uint32_t m = 0;
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t x = w[k] ^ 0x0A0A0A0Au;                       // zero byte where '\n'
        const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);   // 0x80 per zero byte, exact
        m |= (((z >> 7) & 1u) | ((z >> 14) & 2u) | ((z >> 21) & 4u) | ((z >> 28) & 8u)) << (4 * k);
    }
    return valid >= 16 ? m : (m & ((1u << valid) - 1u));
}

__global__ void __launch_bounds__(PARSE_THREADS)
k_parse_count(const uint8_t *__restrict__ text, uint32_t n, uint32_t skip, unsigned long long *__restrict__ tile_counts)
{
    __shared__ uint32_t s_part[PARSE_THREADS / 32];
    const uint32_t n_tiles = (n + PARSE_TILE - 1) / PARSE_TILE;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t p = tile * PARSE_TILE + threadIdx.x * 16;
        uint32_t c = 0;
        if (p < n) {
            uint32_t m = nl_mask16(__ldg(reinterpret_cast<const uint4 *>(text + p)), n - p);
            if (p == 0) m &= ~((1u << skip) - 1u);          // the first `skip` (< 16) bytes are not text
            c = __popc(m);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = c;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
            for (int w = 0; w < PARSE_THREADS / 32; ++w) t += s_part[w];
            tile_counts[tile] = t;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024)
k_parse_scan(unsigned long long *v, uint32_t n, unsigned long long *total)
{
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const unsigned long long x = i < n ? v[i] : 0ull;
        unsigned long long incl = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = s_warp[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long u = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += u;
            }
            s_warp[lane] = wi - w;
        }
        __syncthreads();
        const unsigned long long carry = s_carry;
        if (i < n) v[i] = carry + s_warp[warp] + (incl - x);
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + s_warp[warp] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_carry;
}

__global__ void __launch_bounds__(PARSE_THREADS)
k_parse_index(const uint8_t *__restrict__ text, uint32_t n, uint32_t skip, const unsigned long long *__restrict__ tile_off,
              uint32_t *__restrict__ line_end, uint32_t max_lines)
{
    __shared__ uint32_t s_warp[PARSE_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_tiles = (n + PARSE_TILE - 1) / PARSE_TILE;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t p = tile * PARSE_TILE + threadIdx.x * 16;
        uint32_t m = 0;
        if (p < n) m = nl_mask16(__ldg(reinterpret_cast<const uint4 *>(text + p)), n - p);
        if (p == 0) m &= ~((1u << skip) - 1u);
        const uint32_t c = __popc(m);
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t wbase = 0;
        for (int w = 0; w < warp; ++w) wbase += s_warp[w];
        unsigned long long k = tile_off[tile] + wbase + (incl - c);
        while (m) {
            const int b = __ffs((int)m) - 1;
            m &= m - 1;
            if (k < max_lines) line_end[k] = p + b;
            ++k;
        }
        __syncthreads();
    }
}

// err[0] = smallest bad record index in this chunk (0xFFFFFFFF = none)
__global__ void __launch_bounds__(256)
k_parse_spans(const uint8_t *__restrict__ text, uint32_t skip, const uint32_t *__restrict__ line_end, uint32_t n_records,
              vfb_span *__restrict__ spans, uint32_t *__restrict__ err)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_records) return;
    uint32_t s[4], e[4];
    uint32_t prev = r ? line_end[4 * r - 1] + 1 : skip;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        const uint32_t nl = line_end[4 * r + l];
        s[l] = prev;
        e[l] = nl;
        if (e[l] > s[l] && text[e[l] - 1] == '\r') --e[l];          // "\r\n"
        prev = nl + 1;
    }
    const bool ok = e[0] > s[0] && text[s[0]] == '@' && e[2] > s[2] && text[s[2]] == '+' &&
                    (e[1] - s[1]) == (e[3] - s[3]);
    if (!ok) atomicMin(err, r);
    spans[r] = vfb_span{s[1], e[1] - s[1]};
}

int launch_parse(const uint8_t *d_text, uint32_t n_bytes, uint32_t n_lines, uint32_t n_records,
                 unsigned long long *tile_scratch, uint32_t *line_end, vfb_span *spans, uint32_t *err,
                 cudaStream_t st)
{
    if (n_records == 0) return VFB_OK;
    const uint32_t n_tiles = (n_bytes + PARSE_TILE - 1) / PARSE_TILE;
    uint32_t blocks = n_tiles < 148u * 8u ? n_tiles : 148u * 8u;
    k_parse_count<<<blocks, PARSE_THREADS, 0, st>>>(d_text, n_bytes, 0u, tile_scratch);
    k_parse_scan<<<1, 1024, 0, st>>>(tile_scratch, n_tiles, tile_scratch + n_tiles);
    k_parse_index<<<blocks, PARSE_THREADS, 0, st>>>(d_text, n_bytes, 0u, tile_scratch, line_end, n_lines);
    k_parse_spans<<<(n_records + 255) / 256, 256, 0, st>>>(d_text, 0u, line_end, n_records, spans, err);
    g_launches += 4;
    VFB_CUDA(cudaGetLastError());
    return VFB_OK;
}

uint64_t parse_tile_words(uint32_t n_bytes) { return (uint64_t)(n_bytes + PARSE_TILE - 1) / PARSE_TILE + 1; }

// ---- the same passes, split for text that was inflated on the device: the host does not know
// ---- the line count (it never sees the text), so it reads it back between the passes.
int launch_parse_count(const uint8_t *d_text, uint32_t n_bytes, uint32_t skip, unsigned long long *tile_scratch, cudaStream_t st)
{
    if (n_bytes == 0) return VFB_OK;
    const uint32_t n_tiles = (n_bytes + PARSE_TILE - 1) / PARSE_TILE;
    uint32_t blocks = n_tiles < 148u * 8u ? n_tiles : 148u * 8u;
    k_parse_count<<<blocks, PARSE_THREADS, 0, st>>>(d_text, n_bytes, skip, tile_scratch);
    k_parse_scan<<<1, 1024, 0, st>>>(tile_scratch, n_tiles, tile_scratch + n_tiles);   // total at [n_tiles]
    g_launches += 2;
    VFB_CUDA(cudaGetLastError());
    return VFB_OK;
}

// info[0] = cut (bytes of complete records), info[1] = tail length; tail receives up to tail_cap
// bytes of the text after the last complete record.
__global__ void __launch_bounds__(256)
k_parse_tail(const uint8_t *__restrict__ text, uint32_t n_bytes, uint32_t skip, const uint32_t *__restrict__ line_end,
             uint32_t n_records, uint8_t *__restrict__ tail, uint32_t tail_cap, uint32_t *__restrict__ info)
{
    const uint32_t cut = n_records ? line_end[4 * n_records - 1] + 1 : skip;
    const uint32_t tl = n_bytes - cut;
    if (threadIdx.x == 0) { info[0] = cut; info[1] = tl; }
    const uint32_t n = tl < tail_cap ? tl : tail_cap;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) tail[i] = text[cut + i];
}

int launch_parse_index(const uint8_t *d_text, uint32_t n_bytes, uint32_t skip, uint32_t n_lines, uint32_t n_records,
                       const unsigned long long *tile_scratch, uint32_t *line_end, vfb_span *spans, uint32_t *err,
                       uint8_t *tail, uint32_t tail_cap, uint32_t *info, cudaStream_t st)
{
    const uint32_t n_tiles = (n_bytes + PARSE_TILE - 1) / PARSE_TILE;
    uint32_t blocks = n_tiles < 148u * 8u ? n_tiles : 148u * 8u;
    if (n_lines) {
        k_parse_index<<<blocks, PARSE_THREADS, 0, st>>>(d_text, n_bytes, skip, tile_scratch, line_end, n_lines);
        ++g_launches;
    }
    if (n_records) {
        k_parse_spans<<<(n_records + 255) / 256, 256, 0, st>>>(d_text, skip, line_end, n_records, spans, err);
        ++g_launches;
    }
    k_parse_tail<<<1, 256, 0, st>>>(d_text, n_bytes, skip, line_end, n_records, tail, tail_cap, info);
    ++g_launches;
    VFB_CUDA(cudaGetLastError());
    return VFB_OK;
}

}  // namespace vfb
