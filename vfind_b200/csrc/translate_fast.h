// Arithmetic fast path of `translate` (/root/reference/src/lib.rs:16-44 with the tables at :52-95) for twelve bases
// at a time, shared by the device kernel (kernels_count.cu) and a host build of the same code that the CPU tests
// drive through vfb_debug_translate12 (the kernel itself cannot run without a GPU; its arithmetic can).
//
// A canonical base byte (A C G T U, either case) is identified by its low three bits alone — A 1, C 3, G 7, T 4,
// U 5 — so a codon is a 12-bit index n0 | n1 << 4 | n2 << 8 into a 4096-entry amino-acid table, and whether the
// twelve bytes ARE canonical is checked by rebuilding the upper-case letter from those three bits (one byte
// permute per word through an 8-entry letter table) and comparing it with the byte, case bit cleared.  Anything
// else (N, '-', non-ASCII, ...) makes the check fail and the caller takes the table-per-byte path, which turns such
// codons into 'X' as the reference does.
#pragma once
#include <stdint.h>

#include "hash.h"

// PRMT with a host stand-in; only the low 16 selector bits are used
VFB_HD uint32_t tf_prmt(uint32_t a, uint32_t b, uint32_t s)
{
#if defined(__CUDA_ARCH__)
    // prmt.b32 in its default mode (selector bit 3 = replicate the sign); __byte_perm would mask the selector first
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(s));
    return r;
#else
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t n = (s >> (4 * i)) & 0xFu;
        uint32_t byte = (uint32_t)(v >> (8 * (n & 7u))) & 0xFFu;
        if (n & 8u) byte = (byte & 0x80u) ? 0xFFu : 0u;
        r |= byte << (8 * i);
    }
    return r;
#endif
}

// letter table by low three bits: 1 'A', 3 'C', 4 'T', 5 'U', 7 'G'; 0xFF (never equal to a byte whose bit 5 is
// cleared) elsewhere
#define TF_LETTERS_LO 0x43FF41FFu
#define TF_LETTERS_HI 0x47FF5554u

// base index of AA_TABLE_CANONICAL's ordering (A 0, C 1, G 2, T/U 3) by low three bits; 4 = not a base
VFB_HD uint32_t tf_base_of_low3(uint32_t n)
{
    return n == 1 ? 0u : n == 3 ? 1u : n == 7 ? 2u : (n == 4 || n == 5) ? 3u : 4u;
}

// Entry i of the 4096-entry codon table; aa64 = the 64 amino acids as c0 * 16 + c1 * 4 + c2.
VFB_HD uint8_t tf_codon_entry(uint32_t i, const char *aa64)
{
    const uint32_t c0 = tf_base_of_low3(i & 7u), c1 = tf_base_of_low3((i >> 4) & 7u), c2 = tf_base_of_low3((i >> 8) & 7u);
    if ((i & 0x888u) || c0 == 4 || c1 == 4 || c2 == 4) return (uint8_t)'X';
    return (uint8_t)aa64[c0 * 16 + c1 * 4 + c2];
}

// Twelve bases x0 x1 x2 (little-endian words, base k in byte k & 3 of word k >> 2) -> the four codon indices;
// returns false when a byte is not a canonical base (the indices are then meaningless).
VFB_HD bool tf_codons12(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t &i0, uint32_t &i1, uint32_t &i2, uint32_t &i3)
{
    const uint32_t s0 = x0 & 0x07070707u, s1 = x1 & 0x07070707u, s2 = x2 & 0x07070707u;
    // byte 0 = n0 | n1 << 4, byte 2 = n2 | n3 << 4
    const uint32_t t0 = s0 | (s0 >> 4), t1 = s1 | (s1 >> 4), t2 = s2 | (s2 >> 4);
    // eight nibbles each: the three-bit codes of words (0, 1) and (1, 2)
    const uint32_t n01 = tf_prmt(t0, t1, 0x6420u), n12 = tf_prmt(t1, t2, 0x6420u);
    const uint32_t e0 = tf_prmt(TF_LETTERS_LO, TF_LETTERS_HI, n01);
    const uint32_t e1 = tf_prmt(TF_LETTERS_LO, TF_LETTERS_HI, n12);
    const uint32_t e2 = tf_prmt(TF_LETTERS_LO, TF_LETTERS_HI, n12 >> 16);
    const uint32_t bad = ((x0 & 0xDFDFDFDFu) ^ e0) | ((x1 & 0xDFDFDFDFu) ^ e1) | ((x2 & 0xDFDFDFDFu) ^ e2);
    i0 = n01 & 0xFFFu;             // nibbles 0..2
    i1 = (n01 >> 12) & 0xFFFu;     // nibbles 3..5
    i2 = (n12 >> 8) & 0xFFFu;      // nibbles 6..8 = nibbles 2..4 of (1, 2)
    i3 = n12 >> 20;                // nibbles 9..11
    return bad == 0u;
}
