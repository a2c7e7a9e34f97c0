// Counter-based synthetic read generator, identical on host and device (integer only).
//
// Shapes follow BASELINE.json's configs as made concrete in SURVEY.md §8(d):
//   read = lead | P' | VR | S' | tail, total length exactly L (tail padded / truncated).
//   P, S      : the two adapters, uniform ACGT of length A derived from the seed.
//   P', S'    : the adapter, mutated with probability p_err by k edits (k = 1/2/3 w.p. .6/.3/.1)
//               at uniform positions; an edit is an indel w.p. `indel` (half ins, half del),
//               else a substitution by a different base.
//   VR        : entry v of a library of U uniform-ACGT strings of length V (a `frameshift`
//               fraction has V-1 or V+1 to exercise the partial-codon drop); v is uniform or
//               octave-Zipf (a uniform octave, then uniform inside it: P(v) ~ 1/v);
//               per-base substitution noise and rare 'N'.
//   lead      : uniform length in [0, L - 2A - V - 4] when that is positive, uniform ACGT.
// Every draw is splitmix64(seed, read index or variant index, field), so any shard of the
// stream can be generated independently and reproducibly.
#pragma once
#include <stdint.h>

#include "../../include/vfind_b200.h"

#if defined(__CUDACC__)
#define VFS_HD __host__ __device__ __forceinline__
#else
#define VFS_HD static inline
#endif

VFS_HD uint64_t vfs_mix(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

VFS_HD uint64_t vfs_draw(uint64_t seed, uint64_t a, uint64_t b)
{
    return vfs_mix(vfs_mix(seed ^ (a * 0xD1342543DE82EF95ull)) + b * 0x2545F4914F6CDD1Dull);
}

VFS_HD uint8_t vfs_base(uint64_t r) { return (uint8_t)("ACGT"[r & 3]); }

// uniform integer in [0, n) from the high bits of a 64-bit draw
VFS_HD uint32_t vfs_below(uint64_t r, uint32_t n)
{
    return (uint32_t)(((r >> 32) * (uint64_t)n) >> 32);
}

VFS_HD void vfs_adapter(uint64_t seed, int which, uint32_t A, uint8_t *out)
{
    for (uint32_t i = 0; i < A; ++i)
        out[i] = vfs_base(vfs_draw(seed, 0xADA0ull + (uint64_t)which, i) >> 17);
}

// Length of library entry v.
VFS_HD uint32_t vfs_variant_len(const vfb_synth_cfg *c, uint32_t v)
{
    uint64_t r = vfs_draw(c->seed, 0x11B0000000ull + v, 0);
    if (vfs_below(r, 1000000u) < c->frameshift_ppm) return (r & 1) ? c->region_len + 1 : c->region_len - 1;
    return c->region_len;
}

VFS_HD uint8_t vfs_variant_base(const vfb_synth_cfg *c, uint32_t v, uint32_t pos)
{
    return vfs_base(vfs_draw(c->seed, 0x11B0000000ull + v, 1 + pos) >> 23);
}

VFS_HD uint32_t vfs_pick_variant(const vfb_synth_cfg *c, uint64_t read)
{
    uint64_t r = vfs_draw(c->seed, read, 1);
    uint32_t U = c->n_variants ? c->n_variants : 1;
    if (!c->zipf) return vfs_below(r, U);
    // octave Zipf: octave k uniform in [0, bits), then uniform in [2^k - 1, 2^(k+1) - 1) clipped to U
    uint32_t bits = 0;
    while ((1ull << bits) <= U) ++bits;            // values 0 .. U-1 live in octaves 0 .. bits-1
    uint32_t k = vfs_below(r, bits);
    uint64_t lo = (1ull << k) - 1, hi = (2ull << k) - 1;
    if (hi > U) hi = U;
    if (lo >= hi) lo = hi - 1;
    uint64_t r2 = vfs_draw(c->seed, read, 2);
    return (uint32_t)(lo + vfs_below(r2, (uint32_t)(hi - lo)));
}

// Write the (possibly mutated) adapter instance; returns its length (<= A + 3).
VFS_HD uint32_t vfs_adapter_instance(const vfb_synth_cfg *c, uint64_t read, int which,
                                     const uint8_t *adapter, uint8_t *out)
{
    uint32_t A = c->adapter_len, n = A;
    for (uint32_t i = 0; i < A; ++i) out[i] = adapter[i];
    uint64_t r = vfs_draw(c->seed, read, 0x100 + (uint64_t)which);
    if (vfs_below(r, 1000000u) >= c->p_err_ppm) return n;
    uint32_t kk = vfs_below(vfs_draw(c->seed, read, 0x110 + (uint64_t)which), 10);
    uint32_t k = kk < 6 ? 1 : (kk < 9 ? 2 : 3);
    for (uint32_t e = 0; e < k; ++e) {
        uint64_t re = vfs_draw(c->seed, read, 0x120 + (uint64_t)which * 16 + e);
        uint64_t rp = vfs_draw(c->seed, read, 0x160 + (uint64_t)which * 16 + e);
        int indel = vfs_below(re, 1000000u) < c->indel_ppm || (c->force_indel && e == 0);
        if (n == 0) break;
        uint32_t pos = vfs_below(rp, n);
        if (!indel) {
            uint8_t nb = vfs_base(rp);
            if (nb == out[pos]) nb = vfs_base(rp + 1);
            out[pos] = nb;
        } else if (re & 1) {                      // insertion before pos
            for (uint32_t i = n; i > pos; --i) out[i] = out[i - 1];
            out[pos] = vfs_base(rp >> 7);
            ++n;
        } else if (n > 1) {                       // deletion
            for (uint32_t i = pos; i + 1 < n; ++i) out[i] = out[i + 1];
            --n;
        }
    }
    return n;
}

// Generate read `read` into out[0 .. L).
VFS_HD void vfs_read(const vfb_synth_cfg *c, uint64_t read, const uint8_t *prefix,
                     const uint8_t *suffix, uint8_t *out)
{
    const uint32_t L = c->read_len, A = c->adapter_len;
    uint32_t v = vfs_pick_variant(c, read);
    uint32_t V = vfs_variant_len(c, v);
    int32_t slack = (int32_t)L - (int32_t)(2 * A + V) - 4;
    uint32_t lead = slack > 0 ? vfs_below(vfs_draw(c->seed, read, 3), (uint32_t)slack + 1) : 0;
    uint8_t inst[96];
    uint32_t p = 0;
    for (uint32_t i = 0; i < lead && p < L; ++i) out[p++] = vfs_base(vfs_draw(c->seed, read, 0x1000 + i) >> 11);
    uint32_t n = vfs_adapter_instance(c, read, 0, prefix, inst);
    for (uint32_t i = 0; i < n && p < L; ++i) out[p++] = inst[i];
    for (uint32_t i = 0; i < V && p < L; ++i) {
        uint8_t b = vfs_variant_base(c, v, i);
        uint64_t rn = vfs_draw(c->seed, read, 0x2000 + i);
        uint32_t u = vfs_below(rn, 1000000u);
        if (u < c->noise_ppm) {
            uint8_t nb = vfs_base(rn);
            if (nb == b) nb = vfs_base(rn + 1);
            b = nb;
        } else if (u < c->noise_ppm + c->n_ppm) {
            b = 'N';
        }
        out[p++] = b;
    }
    n = vfs_adapter_instance(c, read, 1, suffix, inst);
    for (uint32_t i = 0; i < n && p < L; ++i) out[p++] = inst[i];
    for (uint32_t i = 0; p < L; ++i) out[p++] = vfs_base(vfs_draw(c->seed, read, 0x3000 + i) >> 11);
}
