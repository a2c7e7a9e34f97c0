// Key hashing shared by host and device code.
//
// The count table (HashMap<String,u64>, /root/reference/src/lib.rs:263) is keyed by the
// variant bytes.  On the device a key is hashed once, by the kernel that produces it, as a
// multilinear form over its zero-padded little-endian 32-bit words followed by a finaliser:
//     h = fmix64( sum_i (w_i ^ VFB_HASH_K) * M(i)  ^  len * VFB_HASH_L )
// M(i) is an odd 64-bit multiplier derived from i alone, so lanes can hash disjoint words
// and add.  The hash only places keys; equality is always decided on the full key bytes.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define VFB_HD __host__ __device__ __forceinline__
#else
#define VFB_HD static inline
#endif

#define VFB_HASH_K 0x9E3779B9u
#define VFB_HASH_L 0xD6E8FEB86659FD93ull

VFB_HD uint64_t vfb_fmix64(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

VFB_HD uint64_t vfb_hash_mult(uint32_t i)
{
    uint64_t z = 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z | 1ull;
}

VFB_HD uint64_t vfb_hash_term(uint32_t word, uint32_t i)
{
    return (uint64_t)(word ^ VFB_HASH_K) * vfb_hash_mult(i);
}

VFB_HD uint64_t vfb_hash_finish(uint64_t acc, uint32_t len)
{
    return vfb_fmix64(acc ^ ((uint64_t)len * VFB_HASH_L));
}

// Reference implementation over a byte string (host side of merges and tests).
VFB_HD uint64_t vfb_hash_bytes(const uint8_t *key, uint32_t len)
{
    uint64_t acc = 0;
    uint32_t nw = (len + 3) / 4;
    for (uint32_t i = 0; i < nw; ++i) {
        uint32_t w = 0;
        for (uint32_t b = 0; b < 4; ++b) {
            uint32_t p = i * 4 + b;
            if (p < len) w |= (uint32_t)key[p] << (8 * b);
        }
        acc += vfb_hash_term(w, i);
    }
    return vfb_hash_finish(acc, len);
}

// Owner rank of a key in an n-way partition (multi-GPU merge).
VFB_HD uint32_t vfb_hash_owner(uint64_t h, uint32_t n_parts)
{
    return (uint32_t)(((h >> 32) * (uint64_t)n_parts) >> 32);
}
