// K3 — variable-region extraction + codon translation + key hashing, and
// K4 — exact on-device counting in an open-addressing table, plus the table utilities the
// multi-GPU merge needs.
//
// K3 replaces the slice of /root/reference/src/lib.rs:288-290 and `translate` (:16-44 with
// the tables at :52-95); K4 replaces `*variants.entry(k).or_insert(0) += 1` (:296, :301) on
// HashMap<String,u64> (:263) and the unzip to columns (:312).
#include "hash.h"
#include "vfb_internal.cuh"

namespace vfb {

// ------------------------------------------------------------------------------------ K3
// AA_TABLE_CANONICAL (src/lib.rs:52-77) flattened as c1*16 + c2*4 + c3, A=0 C=1 G=2 T/U=3.
__constant__ char c_aa[65] =
    "KNKN" "TTTT" "RSRS" "IIMI"
    "QHQH" "PPPP" "RRRR" "LLLL"
    "EDED" "AAAA" "GGGG" "VVVV"
    "*Y*Y" "SSSS" "*CWC" "LFLF";

// ASCII_TO_INDEX (src/lib.rs:86-95); bytes >= 128 are handled before the lookup (:24-29).
__device__ __forceinline__ uint32_t tr_index(uint8_t b)
{
    switch (b) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': case 'U': case 'u': return 3;
    default: return 4;
    }
}

// String::from_utf8 validity (src/lib.rs:295), sequential; only runs when a region holds
// a byte >= 0x80.
__device__ bool utf8_valid(const uint8_t *s, uint32_t n)
{
    uint32_t i = 0;
    while (i < n) {
        const uint8_t b = s[i];
        if (b < 0x80) { ++i; continue; }
        if (b >= 0xC2 && b <= 0xDF) {
            if (i + 1 >= n || (s[i + 1] & 0xC0) != 0x80) return false;
            i += 2;
        } else if (b >= 0xE0 && b <= 0xEF) {
            if (i + 2 >= n) return false;
            const uint8_t b1 = s[i + 1], b2 = s[i + 2];
            if ((b1 & 0xC0) != 0x80 || (b2 & 0xC0) != 0x80) return false;
            if (b == 0xE0 && b1 < 0xA0) return false;
            if (b == 0xED && b1 > 0x9F) return false;
            i += 3;
        } else if (b >= 0xF0 && b <= 0xF4) {
            if (i + 3 >= n) return false;
            const uint8_t b1 = s[i + 1], b2 = s[i + 2], b3 = s[i + 3];
            if ((b1 & 0xC0) != 0x80 || (b2 & 0xC0) != 0x80 || (b3 & 0xC0) != 0x80) return false;
            if (b == 0xF0 && b1 < 0x90) return false;
            if (b == 0xF4 && b1 > 0x8F) return false;
            i += 4;
        } else {
            return false;
        }
    }
    return true;
}

#define KEY_THREADS 256

__global__ void __launch_bounds__(KEY_THREADS)
k3_keys(const __grid_constant__ KeyJob job)
{
    __shared__ uint8_t s_lut[256];
    __shared__ char s_aa[64];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = (uint8_t)(i >= 128 ? 4 : tr_index((uint8_t)i));
    if (threadIdx.x < 64) s_aa[threadIdx.x] = c_aa[threadIdx.x];
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const uint32_t warps_total = gridDim.x * (KEY_THREADS / 32);
    for (uint32_t r = blockIdx.x * (KEY_THREADS / 32) + (threadIdx.x >> 5); r < job.n_reads; r += warps_total) {
        const uint32_t s = job.start[r], e = job.end[r];
        const vfb_span sp = job.spans[r];
        uint32_t klen = 0;
        // src/lib.rs:288: both located and start < end (strict).  end <= len always holds for
        // located suffix boundaries; a prefix boundary beyond it fails start < end.
        const bool located = s != VFB_NONE && e != VFB_NONE && s < e && e <= sp.len;
        const uint8_t *var = job.text + sp.off + s;
        const uint32_t V = located ? e - s : 0;
        if (located) {
            if (!job.skip_translation) {
                if (V % 3 == 0) klen = V / 3;                       // :17-19 partial codon -> None
            } else {
                bool high = false;
                for (uint32_t c = lane; c < V; c += 32) high |= __ldg(var + c) >= 0x80;
                klen = V;
                if (__any_sync(0xffffffffu, high)) {
                    int ok = 1;
                    if (lane == 0) ok = utf8_valid(var, V) ? 1 : 0;
                    ok = __shfl_sync(0xffffffffu, ok, 0);
                    if (!ok) klen = 0;                              // :295 from_utf8 Err -> dropped
                }
            }
        }
        if (klen) {
            const uint32_t padded = (klen + 15u) & ~15u;
            unsigned long long off = 0;
            if (lane == 0) off = atomicAdd(job.key_cursor, (unsigned long long)padded);
            off = __shfl_sync(0xffffffffu, off, 0);
            uint8_t *key = job.keys + off;
            if (!job.skip_translation) {
                for (uint32_t c = lane; c < klen; c += 32) {
                    const uint8_t b0 = __ldg(var + 3 * c), b1 = __ldg(var + 3 * c + 1), b2 = __ldg(var + 3 * c + 2);
                    const uint32_t i0 = s_lut[b0], i1 = s_lut[b1], i2 = s_lut[b2];
                    key[c] = (i0 | i1 | i2) & 4 ? (uint8_t)'X' : (uint8_t)s_aa[i0 * 16 + i1 * 4 + i2];
                }
            } else {
                for (uint32_t c = lane; c < klen; c += 32) key[c] = __ldg(var + c);
            }
            for (uint32_t c = klen + lane; c < padded; c += 32) key[c] = 0;
            __syncwarp();
            uint64_t acc = 0;
            const uint32_t nw = (klen + 3) / 4;
            const uint32_t *kw = reinterpret_cast<const uint32_t *>(key);
            for (uint32_t i = lane; i < nw; i += 32) acc += vfb_hash_term(kw[i], i);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) {
                uint64_t h = vfb_hash_finish(acc, klen);
                if (job.hash_bits > 0 && job.hash_bits < 64) h &= (1ull << job.hash_bits) - 1;
                job.khash[r] = h;
                job.koff[r] = off;
            }
        }
        if (lane == 0) job.klen[r] = klen;
        __syncwarp();
    }
}

int launch_keys(const KeyJob &job, cudaStream_t st)
{
    if (job.n_reads == 0) return VFB_OK;
    uint32_t blocks = (job.n_reads + (KEY_THREADS / 32) - 1) / (KEY_THREADS / 32);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k3_keys<<<blocks, KEY_THREADS, 0, st>>>(job);
    ++g_launches;
    VFB_CUDA(cudaGetLastError());
    return VFB_OK;
}

// ------------------------------------------------------------------------------------ K4
#define REF_BATCH 0x80000000u

struct InsertArgs {
    DevTable t;
    InsertJob job;
};

__device__ __forceinline__ const uint8_t *job_key(const InsertJob &j, uint32_t i)
{
    return j.koff ? j.keys + j.koff[i] : j.keys + (size_t)i * j.key_stride;
}

__device__ __forceinline__ bool keys_equal16(const uint8_t *a, const uint8_t *b, uint32_t len)
{
    const uint4 *x = reinterpret_cast<const uint4 *>(a), *y = reinterpret_cast<const uint4 *>(b);
    const uint32_t n = (len + 15) / 16;
    for (uint32_t i = 0; i < n; ++i) {
        const uint4 u = x[i], v = y[i];
        if (u.x != v.x || u.y != v.y || u.z != v.z || u.w != v.w) return false;
    }
    return true;
}

// One thread per key: probe, claim or match (full key compare), count.
__global__ void __launch_bounds__(256)
k4_insert(const __grid_constant__ InsertArgs a)
{
    const InsertJob &j = a.job;
    const DevTable &t = a.t;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t klen = 0;
    if (i < j.n_keys) klen = j.klen[i];
    const unsigned active = __ballot_sync(0xffffffffu, klen != 0);
    if ((threadIdx.x & 31) == 0 && active && !j.kcount)
        atomicAdd(&t.counters[2], (unsigned long long)__popc(active));
    if (i >= j.n_keys) return;
    uint32_t owner = VFB_NONE;
    if (klen) {
        const uint64_t h = j.khash[i];
        const unsigned long long cnt = j.kcount ? j.kcount[i] : 1ull;
        if (j.kcount && cnt) atomicAdd(&t.counters[2], cnt);
        const uint32_t tag = (uint32_t)(h >> 32);
        const uint64_t mask = t.capacity - 1;
        uint64_t slot = h & mask;
        const unsigned long long mine = ((unsigned long long)tag << 32) | REF_BATCH | i;
        const uint8_t *mykey = job_key(j, i);
        for (;;) {
            unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(&t.slots[slot]);
            if (cur == 0ull) {
                cur = atomicCAS(&t.slots[slot], 0ull, mine);
                if (cur == 0ull) {
                    owner = (uint32_t)slot;
                    atomicAdd(&t.counts[slot], cnt);
                    break;
                }
            }
            if ((uint32_t)(cur >> 32) == tag) {
                const uint32_t ref = (uint32_t)cur;
                const uint8_t *okey;
                uint32_t olen;
                if (ref & REF_BATCH) {
                    const uint32_t o = ref & ~REF_BATCH;
                    okey = job_key(j, o);
                    olen = j.klen[o];
                } else {
                    const uint32_t row = ref - 1;
                    okey = t.arena + t.row_off[row];
                    olen = t.row_len[row];
                }
                if (olen == klen && keys_equal16(okey, mykey, klen)) {
                    atomicAdd(&t.counts[slot], cnt);
                    break;
                }
            }
            slot = (slot + 1) & mask;
        }
    }
    j.owner_slot[i] = owner;
}

// Owners of freshly claimed slots move their key into the arena and turn the slot's
// batch reference into a row reference.
__global__ void __launch_bounds__(256)
k4_publish(const __grid_constant__ InsertArgs a)
{
    const InsertJob &j = a.job;
    const DevTable &t = a.t;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= j.n_keys) return;
    const uint32_t slot = j.owner_slot[i];
    if (slot == VFB_NONE) return;
    const uint32_t klen = j.klen[i];
    const uint32_t padded = (klen + 15u) & ~15u;
    const unsigned long long row = atomicAdd(&t.counters[0], 1ull);
    const unsigned long long off = atomicAdd(&t.counters[1], (unsigned long long)padded);
    const uint4 *src = reinterpret_cast<const uint4 *>(job_key(j, i));
    uint4 *dst = reinterpret_cast<uint4 *>(t.arena + off);
    for (uint32_t c = 0; c < padded / 16; ++c) dst[c] = src[c];
    const uint64_t h = j.khash[i];
    t.row_hash[row] = h;
    t.row_off[row] = off;
    t.row_len[row] = klen;
    t.slots[slot] = ((unsigned long long)(uint32_t)(h >> 32) << 32) | (unsigned long long)(row + 1);
}

int launch_insert(const DevTable &t, const InsertJob &job, cudaStream_t st)
{
    if (job.n_keys == 0) return VFB_OK;
    InsertArgs a;
    a.t = t;
    a.job = job;
    const uint32_t blocks = (job.n_keys + 255) / 256;
    k4_insert<<<blocks, 256, 0, st>>>(a);
    k4_publish<<<blocks, 256, 0, st>>>(a);
    g_launches += 2;
    VFB_CUDA(cudaGetLastError());
    return VFB_OK;
}

__global__ void __launch_bounds__(256)
k4_rehash(const DevTable o, const DevTable n)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t mask = n.capacity - 1;
    for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < o.capacity; s += stride) {
        const unsigned long long w = o.slots[s];
        if (!w) continue;
        const uint32_t row = (uint32_t)w - 1;
        uint64_t slot = o.row_hash[row] & mask;
        for (;;) {
            if (atomicCAS(&n.slots[slot], 0ull, w) == 0ull) break;
            slot = (slot + 1) & mask;
        }
        n.counts[slot] = o.counts[s];
    }
}

int launch_rehash(const DevTable &old_t, const DevTable &new_t, cudaStream_t st)
{
    uint64_t blocks = (old_t.capacity + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k4_rehash<<<(uint32_t)blocks, 256, 0, st>>>(old_t, new_t);
    ++g_launches;
    VFB_CUDA(cudaGetLastError());
    return VFB_OK;
}

__global__ void __launch_bounds__(256)
k4_export_counts(const DevTable t, unsigned long long *row_count)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < t.capacity; s += stride) {
        const unsigned long long w = t.slots[s];
        if (w) row_count[(uint32_t)w - 1] = t.counts[s];
    }
}

int launch_export_counts(const DevTable &t, uint64_t rows, unsigned long long *row_count, cudaStream_t st)
{
    if (rows == 0) return VFB_OK;
    uint64_t blocks = (t.capacity + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k4_export_counts<<<(uint32_t)blocks, 256, 0, st>>>(t, row_count);
    ++g_launches;
    VFB_CUDA(cudaGetLastError());
    return VFB_OK;
}

// ------------------------------------------------------------------------------------ merge
__global__ void __launch_bounds__(256)
k5_partition_count(const DevTable t, uint64_t rows, uint32_t n_parts,
                   unsigned long long *part_rows, unsigned long long *part_keybytes)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += stride) {
        const uint32_t p = vfb_hash_owner(t.row_hash[r], n_parts);
        atomicAdd(&part_rows[p], 1ull);
        atomicAdd(&part_keybytes[p], (unsigned long long)((t.row_len[r] + 15u) & ~15u));
    }
}

int launch_partition_count(const DevTable &t, uint64_t rows, uint32_t n_parts,
                           unsigned long long *part_rows, unsigned long long *part_keybytes,
                           cudaStream_t st)
{
    if (rows == 0) return VFB_OK;
    uint64_t blocks = (rows + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k5_partition_count<<<(uint32_t)blocks, 256, 0, st>>>(t, rows, n_parts, part_rows, part_keybytes);
    ++g_launches;
    VFB_CUDA(cudaGetLastError());
    return VFB_OK;
}

__global__ void __launch_bounds__(256)
k5_partition_fill(const DevTable t, uint64_t rows, uint32_t n_parts,
                  const unsigned long long *row_count, uint8_t *buf, const uint64_t *chunk_off,
                  const uint64_t *part_rows, const uint64_t * /*part_keybytes*/,
                  unsigned long long *cursors)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += stride) {
        const uint64_t h = t.row_hash[r];
        const uint32_t p = vfb_hash_owner(h, n_parts);
        const uint32_t len = t.row_len[r];
        const uint32_t padded = (len + 15u) & ~15u;
        const unsigned long long idx = atomicAdd(&cursors[2 * p], 1ull);
        const unsigned long long koff = atomicAdd(&cursors[2 * p + 1], (unsigned long long)padded);
        const uint64_t n = part_rows[p];
        uint8_t *c = buf + chunk_off[p];
        uint64_t *c_hash = reinterpret_cast<uint64_t *>(c + sizeof(ChunkHeader));
        uint64_t *c_count = reinterpret_cast<uint64_t *>(c + sizeof(ChunkHeader) + vfb_align16(n * 8));
        uint64_t *c_koff = reinterpret_cast<uint64_t *>(c + sizeof(ChunkHeader) + vfb_align16(n * 8) * 2);
        uint32_t *c_klen = reinterpret_cast<uint32_t *>(c + sizeof(ChunkHeader) + vfb_align16(n * 8) * 3);
        uint8_t *c_keys = c + sizeof(ChunkHeader) + vfb_align16(n * 8) * 3 + vfb_align16(n * 4);
        c_hash[idx] = h;
        c_count[idx] = row_count[r];
        c_koff[idx] = koff;
        c_klen[idx] = len;
        const uint4 *src = reinterpret_cast<const uint4 *>(t.arena + t.row_off[r]);
        uint4 *dst = reinterpret_cast<uint4 *>(c_keys + koff);
        for (uint32_t q = 0; q < padded / 16; ++q) dst[q] = src[q];
    }
}

int launch_partition_fill(const DevTable &t, uint64_t rows, uint32_t n_parts,
                          const unsigned long long *row_count, uint8_t *buf,
                          const uint64_t *d_chunk_off, const uint64_t *d_part_rows,
                          const uint64_t *d_part_keybytes, unsigned long long *cursors,
                          cudaStream_t st)
{
    if (rows == 0) return VFB_OK;
    uint64_t blocks = (rows + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k5_partition_fill<<<(uint32_t)blocks, 256, 0, st>>>(t, rows, n_parts, row_count, buf, d_chunk_off,
                                                        d_part_rows, d_part_keybytes, cursors);
    ++g_launches;
    VFB_CUDA(cudaGetLastError());
    return VFB_OK;
}

}  // namespace vfb
