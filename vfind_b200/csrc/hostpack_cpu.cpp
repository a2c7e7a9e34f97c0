// Host side of the packed host->device path (hostpack.cu): 32-byte groups of read text that
// consist of upper-case A/C/G/T only are turned into 64 bits (2 bits per base, base k of the
// group in bits 2k..2k+1, code = (byte >> 1) & 3: A 0, C 1, T 2, G 3); every other group is
// flagged in a bit map and kept verbatim.  The device expands the codes again (k_unpack), so
// the text the kernels see is byte-identical to the caller's -- this only moves fewer bytes
// over PCIe.  Plain C++ (g++), AVX2 when the CPU has it, chosen at run time.
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {

inline bool group_scalar(const uint8_t *s, uint64_t *code)
{
    uint64_t c = 0;
    bool ok = true;
    for (int k = 0; k < 32; ++k) {
        const uint8_t b = s[k];
        const uint64_t q = (b >> 1) & 3u;
        ok &= (b == "ACTG"[q]);
        c |= q << (2 * k);
    }
    *code = c;
    return ok;
}

size_t pack_scalar(const uint8_t *src, size_t n_groups, uint64_t *codes, uint32_t *rawmap, uint8_t *raw, size_t raw_cap)
{
    size_t n_raw = 0;
    for (size_t g = 0; g < n_groups; ++g) {
        if (!group_scalar(src + 32 * g, codes + g)) {
            if (n_raw == raw_cap) return SIZE_MAX;
            rawmap[g >> 5] |= 1u << (g & 31);
            memcpy(raw + 32 * n_raw, src + 32 * g, 32);
            ++n_raw;
        }
    }
    return n_raw;
}

#if defined(__x86_64__)
__attribute__((target("avx2")))
size_t pack_avx2(const uint8_t *src, size_t n_groups, uint64_t *codes, uint32_t *rawmap, uint8_t *raw, size_t raw_cap)
{
    const __m256i three = _mm256_set1_epi8(3);
    const __m256i lut = _mm256_setr_epi8('A', 'C', 'T', 'G', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                         'A', 'C', 'T', 'G', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i w14 = _mm256_set1_epi16(0x0401);         // c0 + 4 c1 per byte pair
    const __m256i w116 = _mm256_set1_epi32(0x00100001);    // + 16 (c2 + 4 c3) per 4 bytes
    const __m256i gather = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                            0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    size_t n_raw = 0;
    for (size_t g = 0; g < n_groups; ++g) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + 32 * g));
        const __m256i c = _mm256_and_si256(_mm256_srli_epi16(v, 1), three);
        const __m256i canon = _mm256_shuffle_epi8(lut, c);
        const int ok = _mm256_movemask_epi8(_mm256_cmpeq_epi8(canon, v));
        const __m256i p32 = _mm256_madd_epi16(_mm256_maddubs_epi16(c, w14), w116);
        const __m256i pk = _mm256_shuffle_epi8(p32, gather);
        const uint64_t lo = (uint32_t)_mm256_extract_epi32(pk, 0), hi = (uint32_t)_mm256_extract_epi32(pk, 4);
        codes[g] = lo | (hi << 32);
        if (ok != -1) {
            if (n_raw == raw_cap) return SIZE_MAX;
            rawmap[g >> 5] |= 1u << (g & 31);
            memcpy(raw + 32 * n_raw, src + 32 * g, 32);
            ++n_raw;
        }
    }
    return n_raw;
}
#endif

}  // namespace

// Packs n_groups 32-byte groups.  rawmap (ceil(n_groups / 32) words) must be zeroed by the caller.
// Returns the number of verbatim groups written to `raw`, or SIZE_MAX when more than raw_cap of
// them turn up (the caller then sends the block as it is).
extern "C" size_t vfb_pack_groups(const uint8_t *src, size_t n_groups, uint64_t *codes, uint32_t *rawmap, uint8_t *raw,
                                  size_t raw_cap)
{
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) return pack_avx2(src, n_groups, codes, rawmap, raw, raw_cap);
#endif
    return pack_scalar(src, n_groups, codes, rawmap, raw, raw_cap);
}

// Reference expansion (tests): the inverse of vfb_pack_groups.
extern "C" void vfb_unpack_groups(const uint64_t *codes, const uint32_t *rawmap, const uint8_t *raw, size_t n_groups,
                                  uint8_t *dst)
{
    size_t n_raw = 0;
    for (size_t g = 0; g < n_groups; ++g) {
        if (rawmap[g >> 5] >> (g & 31) & 1u) {
            memcpy(dst + 32 * g, raw + 32 * n_raw, 32);
            ++n_raw;
        } else {
            for (int k = 0; k < 32; ++k) dst[32 * g + k] = (uint8_t)"ACTG"[(codes[g] >> (2 * k)) & 3u];
        }
    }
}
