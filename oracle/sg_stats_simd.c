/*
 * sg_stats_simd.c — the oracle's semi-global alignment (vfo_sg_stats) for SIXTEEN reads at a time.
 *
 * TEST INFRASTRUCTURE ONLY (see vfind_oracle.h).  This is the CPU baseline's fast leg: the reference aligns with
 * parasail's SIMD kernels (`sg_stats_scan_profile_sat`, src/lib.rs:128-135, :155), which vectorise WITHIN one
 * alignment; parasail cannot be built here, so the baseline vectorises ACROSS alignments instead — one AVX2 lane
 * (16 bit) per read, all lanes against the same adapter — which for 20..40-row adapters is at least as good a use of
 * the vector unit.  The recurrence, the tie rules and the end-cell rule are those of vfo_sg_stats, cell for cell; the
 * scalar function stays the checker (tests/test_oracle.py compares the two under every rule switch).
 *
 * Column-major: for read position j (outer) and adapter row i (inner) the lanes hold cell (i, j) of sixteen matrices.
 * State per row: H, HL of the previous column and E, EL (the horizontal gap); F, FL run down the column.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "vfind_oracle.h"

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define VFO_HAVE_X86 1
#else
#define VFO_HAVE_X86 0
#endif

#define LANES 16

int vfo_simd_available(void)
{
#if VFO_HAVE_X86
    return __builtin_cpu_supports("avx2") ? 1 : 0;
#else
    return 0;
#endif
}

#if VFO_HAVE_X86

/* base_code of vfind_oracle.c as a table: A/a 0, T/t 1, C/c 2, G/g 3, anything else 4 */
static const uint8_t CODE[256] = {
#define R4 4, 4, 4, 4
#define R16 R4, R4, R4, R4
    R16, R16, R16, R16,                                                                 /* 0x00..0x3F */
    4, 0, 4, 2, 4, 4, 4, 3, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 1, 4, 4, 4, R4, R4,       /* '@' A..G, T   */
    4, 0, 4, 2, 4, 4, 4, 3, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 1, 4, 4, 4, R4, R4,       /* '`' a..g, t   */
    R16, R16, R16, R16, R16, R16, R16, R16                                              /* 0x80..0xFF */
#undef R4
#undef R16
};
static inline int code_of(uint8_t b) { return CODE[b]; }

typedef struct {
    int16_t *codes;       /* [Lmax][LANES] base codes, 4 beyond a lane's read                    */
    __m256i *state;       /* [A+1][4]: H, HL (previous column), E, EL                             */
    int16_t *col;         /* [A+1][2][LANES]: H, HL of each lane's LAST column                    */
    size_t codes_cap, state_cap, col_cap;
} simd_scratch;

static __thread simd_scratch tls;

static int scratch_fit(int A, int Lmax)
{
    size_t nc = (size_t)Lmax * LANES * sizeof(int16_t), ns = ((size_t)A + 1) * 4 * sizeof(__m256i),
           nl = ((size_t)A + 1) * 2 * LANES * sizeof(int16_t);
    if (nc > tls.codes_cap) {
        free(tls.codes);
        if (posix_memalign((void **)&tls.codes, 32, nc)) { tls.codes = NULL; tls.codes_cap = 0; return -1; }
        tls.codes_cap = nc;
    }
    if (ns > tls.state_cap) {
        free(tls.state);
        if (posix_memalign((void **)&tls.state, 32, ns)) { tls.state = NULL; tls.state_cap = 0; return -1; }
        tls.state_cap = ns;
    }
    if (nl > tls.col_cap) {
        free(tls.col);
        if (posix_memalign((void **)&tls.col, 32, nl)) { tls.col = NULL; tls.col_cap = 0; return -1; }
        tls.col_cap = nl;
    }
    return 0;
}

/* the calling thread's scratch (a worker calls this before it ends) */
void vfo_simd_thread_release(void)
{
    free(tls.codes); free(tls.state); free(tls.col);
    memset(&tls, 0, sizeof tls);
}

/* The matrices.  TIE_OPEN / HPRI / ROW_GE are compile-time constants in the four instantiations below. */
__attribute__((target("avx2"), always_inline)) static inline void
fill16(const uint8_t *acode, int A, int Lmax, const int16_t *lens16, int match, int mismatch, int open, int extend,
       int wild, const int TIE_OPEN, const int HPRI, const int ROW_GE,
       int16_t *row_best, int16_t *row_len, int16_t *row_j)
{
    const __m256i vopen = _mm256_set1_epi16((short)open), vext = _mm256_set1_epi16((short)extend);
    const __m256i vmat = _mm256_set1_epi16((short)match), vmis = _mm256_set1_epi16((short)mismatch);
    const __m256i vwild = _mm256_set1_epi16((short)wild), one = _mm256_set1_epi16(1), zero = _mm256_setzero_si256();
    const __m256i neg = _mm256_set1_epi16(INT16_MIN), four = _mm256_set1_epi16(4);
    const __m256i vlen = _mm256_loadu_si256((const __m256i *)lens16);
    __m256i *st = tls.state;
    for (int i = 0; i <= A; ++i) {
        st[4 * i + 0] = zero; st[4 * i + 1] = zero;      /* H[i][0] = 0, length 0 (adapter begin free) */
        st[4 * i + 2] = neg;  st[4 * i + 3] = zero;      /* E = -inf                                   */
    }
    __m256i best = neg, bestl = zero, bestj = zero;
    __m256i S[5];
    S[4] = vwild;
    for (int j = 1; j <= Lmax; ++j) {
        const __m256i rc = _mm256_load_si256((const __m256i *)(tls.codes + (size_t)(j - 1) * LANES));
        const __m256i isw = _mm256_cmpeq_epi16(rc, four);
        for (int a = 0; a < 4; ++a) {
            const __m256i eq = _mm256_cmpeq_epi16(rc, _mm256_set1_epi16((short)a));
            S[a] = _mm256_blendv_epi8(_mm256_blendv_epi8(vmis, vmat, eq), vwild, isw);
        }
        __m256i dH = zero, dL = zero;                    /* H[i-1][j-1]: row 0 is all zero (read begin free) */
        __m256i uH = zero, uL = zero;                    /* H[i-1][j]                                        */
        __m256i F = neg, FL = zero;
        for (int i = 1; i <= A; ++i) {
            __m256i *s = st + 4 * i;
            const __m256i lH = s[0], lL = s[1];          /* H[i][j-1] */
            /* F: vertical gap */
            const __m256i Fo = _mm256_subs_epi16(uH, vopen), Fe = _mm256_subs_epi16(F, vext);
            const __m256i fm = TIE_OPEN ? _mm256_cmpeq_epi16(_mm256_cmpgt_epi16(Fe, Fo), zero) : _mm256_cmpgt_epi16(Fo, Fe);
            F = _mm256_blendv_epi8(Fe, Fo, fm);
            FL = _mm256_add_epi16(_mm256_blendv_epi8(FL, uL, fm), one);
            /* E: horizontal gap */
            const __m256i Eo = _mm256_subs_epi16(lH, vopen), Ee = _mm256_subs_epi16(s[2], vext);
            const __m256i em = TIE_OPEN ? _mm256_cmpeq_epi16(_mm256_cmpgt_epi16(Ee, Eo), zero) : _mm256_cmpgt_epi16(Eo, Ee);
            const __m256i E = _mm256_blendv_epi8(Ee, Eo, em);
            const __m256i EL = _mm256_add_epi16(_mm256_blendv_epi8(s[3], lL, em), one);
            /* diagonal */
            const __m256i D = _mm256_adds_epi16(dH, S[acode[i - 1]]);
            const __m256i DL = _mm256_add_epi16(dL, one);
            const __m256i notD = _mm256_or_si256(_mm256_cmpgt_epi16(E, D), _mm256_cmpgt_epi16(F, D));
            const __m256i takeE = HPRI ? _mm256_cmpeq_epi16(_mm256_cmpgt_epi16(F, E), zero) : _mm256_cmpgt_epi16(E, F);
            const __m256i G = _mm256_blendv_epi8(F, E, takeE), GL = _mm256_blendv_epi8(FL, EL, takeE);
            const __m256i W = _mm256_blendv_epi8(D, G, notD), WL = _mm256_blendv_epi8(DL, GL, notD);
            dH = lH; dL = lL;
            uH = W; uL = WL;
            s[0] = W; s[1] = WL; s[2] = E; s[3] = EL;
        }
        /* last row: the running best of each lane, over its own columns only */
        const __m256i vj = _mm256_set1_epi16((short)j);
        const __m256i live = _mm256_cmpeq_epi16(_mm256_cmpgt_epi16(vj, vlen), zero);       /* j <= len */
        __m256i better = ROW_GE ? _mm256_cmpeq_epi16(_mm256_cmpgt_epi16(best, uH), zero) : _mm256_cmpgt_epi16(uH, best);
        better = _mm256_and_si256(better, live);
        best = _mm256_blendv_epi8(best, uH, better);
        bestl = _mm256_blendv_epi8(bestl, uL, better);
        bestj = _mm256_blendv_epi8(bestj, vj, better);
        /* lanes whose read ends here keep this column */
        const __m256i last = _mm256_cmpeq_epi16(vj, vlen);
        if (_mm256_movemask_epi8(last)) {
            for (int i = 1; i <= A; ++i) {
                __m256i *c = (__m256i *)(tls.col + (size_t)i * 2 * LANES);
                c[0] = _mm256_blendv_epi8(c[0], st[4 * i + 0], last);
                c[1] = _mm256_blendv_epi8(c[1], st[4 * i + 1], last);
            }
        }
    }
    _mm256_storeu_si256((__m256i *)row_best, best);
    _mm256_storeu_si256((__m256i *)row_len, bestl);
    _mm256_storeu_si256((__m256i *)row_j, bestj);
}

#define FILL_VARIANT(name, T, H, G)                                                                                   \
    __attribute__((target("avx2"), noinline)) static void name(const uint8_t *ac, int A, int Lmax, const int16_t *l,  \
                                                               int m, int x, int o, int e, int w, int16_t *rb,        \
                                                               int16_t *rl, int16_t *rj)                              \
    { fill16(ac, A, Lmax, l, m, x, o, e, w, T, H, G, rb, rl, rj); }
FILL_VARIANT(fill_000, 0, 0, 0) FILL_VARIANT(fill_001, 0, 0, 1) FILL_VARIANT(fill_010, 0, 1, 0) FILL_VARIANT(fill_011, 0, 1, 1)
FILL_VARIANT(fill_100, 1, 0, 0) FILL_VARIANT(fill_101, 1, 0, 1) FILL_VARIANT(fill_110, 1, 1, 0) FILL_VARIANT(fill_111, 1, 1, 1)

static inline int iabs(int v) { return v < 0 ? -v : v; }

#endif /* VFO_HAVE_X86 */

#if !VFO_HAVE_X86
void vfo_simd_thread_release(void) {}
#endif

/* Up to 16 reads against one adapter.  Returns 0 when every output has been written, 1 when the caller has to use
 * vfo_sg_stats (no AVX2, values that do not fit 16-bit lanes, empty inputs).  The outputs equal vfo_sg_stats's. */
int vfo_sg_stats_x16(const uint8_t *adapter, int A, const uint8_t *const *reads, const int *lens, int n,
                     int match, int mismatch, int open, int extend, const vfo_dp_rules *rules,
                     int *score, int *length, int *end_i, int *end_j)
{
#if !VFO_HAVE_X86
    (void)adapter; (void)A; (void)reads; (void)lens; (void)n; (void)match; (void)mismatch; (void)open; (void)extend;
    (void)rules; (void)score; (void)length; (void)end_i; (void)end_j;
    return 1;
#else
    vfo_dp_rules dr;
    if (!rules) { vfo_default_rules(&dr); rules = &dr; }
    if (!vfo_simd_available() || n < 1 || n > LANES || A < 1) return 1;
    int Lmax = 0;
    for (int k = 0; k < n; ++k) {
        if (lens[k] < 1) return 1;
        if (lens[k] > Lmax) Lmax = lens[k];
    }
    /* 16-bit lanes: every H, E, F lies in [-(2*open + (A+1)*extend + maxabs), A*maxabs], every length in [0, A+Lmax] */
    if (open < 0 || extend < 0 || open > 10000 || extend > 10000 || iabs(match) > 10000 || iabs(mismatch) > 10000) return 1;
    {
        const long maxabs = iabs(match) > iabs(mismatch) ? iabs(match) : iabs(mismatch);
        if ((long)A * maxabs + 2L * open + ((long)A + 2) * extend + 2 * maxabs > 30000) return 1;
        if ((long)A + Lmax > 30000) return 1;
    }
    if (scratch_fit(A, Lmax) != 0) return 1;
    uint8_t acode_small[256], *acode = A <= 256 ? acode_small : (uint8_t *)malloc((size_t)A);
    if (!acode) return 1;
    for (int i = 0; i < A; ++i) acode[i] = (uint8_t)code_of(adapter[i]);
    int16_t lens16[LANES];
    for (int k = 0; k < LANES; ++k) lens16[k] = k < n ? (int16_t)lens[k] : 0;
    for (int j = 0; j < Lmax; ++j) {
        int16_t *c = tls.codes + (size_t)j * LANES;
        for (int k = 0; k < LANES; ++k) c[k] = (k < n && j < lens[k]) ? (int16_t)code_of(reads[k][j]) : 4;
    }
    memset(tls.col, 0, ((size_t)A + 1) * 2 * LANES * sizeof(int16_t));
    int16_t rb[LANES], rl[LANES], rj[LANES];
    const int wild = rules->wildcard_zero ? 0 : mismatch;
    const int v = (rules->gap_tie_open ? 4 : 0) | (rules->h_priority ? 2 : 0) | (rules->end_rule == 1 ? 1 : 0);
    switch (v) {
    case 0: fill_000(acode, A, Lmax, lens16, match, mismatch, open, extend, wild, rb, rl, rj); break;
    case 1: fill_001(acode, A, Lmax, lens16, match, mismatch, open, extend, wild, rb, rl, rj); break;
    case 2: fill_010(acode, A, Lmax, lens16, match, mismatch, open, extend, wild, rb, rl, rj); break;
    case 3: fill_011(acode, A, Lmax, lens16, match, mismatch, open, extend, wild, rb, rl, rj); break;
    case 4: fill_100(acode, A, Lmax, lens16, match, mismatch, open, extend, wild, rb, rl, rj); break;
    case 5: fill_101(acode, A, Lmax, lens16, match, mismatch, open, extend, wild, rb, rl, rj); break;
    case 6: fill_110(acode, A, Lmax, lens16, match, mismatch, open, extend, wild, rb, rl, rj); break;
    default: fill_111(acode, A, Lmax, lens16, match, mismatch, open, extend, wild, rb, rl, rj); break;
    }
    if (acode != acode_small) free(acode);
    /* end-cell choice per lane: the same order of candidates as vfo_sg_stats */
    const int NEG = INT32_MIN / 2;
    for (int k = 0; k < n; ++k) {
        const int L = lens[k];
        int sc = NEG, ln = 0, ei = A, ej = 0;
#define COLH(i) ((int)tls.col[(size_t)(i) * 2 * LANES + k])
#define COLL(i) ((int)tls.col[(size_t)(i) * 2 * LANES + LANES + k])
        if (rules->end_rule == 3)
            for (int i = 1; i < A; ++i)
                if (COLH(i) > sc) { sc = COLH(i); ln = COLL(i); ei = i; ej = L; }
        /* the row scan of vfo_sg_stats ends on the leftmost (end_rule 1: rightmost) best cell of the row, and only if
           that cell beats (end_rule 1: reaches) what the scan started from */
        if (rules->end_rule == 1 ? (int)rb[k] >= sc : (int)rb[k] > sc) { sc = rb[k]; ln = rl[k]; ei = A; ej = rj[k]; }
        if (rules->end_rule != 2 && rules->end_rule != 3) {
            int cbest = NEG, ci = 0;
            for (int i = 1; i <= A; ++i)
                if (COLH(i) > cbest) { cbest = COLH(i); ci = i; }
            if (cbest > sc || (cbest == sc && ej == L)) { sc = cbest; ln = COLL(ci); ei = ci; ej = L; }
        }
#undef COLH
#undef COLL
        score[k] = sc; length[k] = ln;
        if (end_i) end_i[k] = ei;
        if (end_j) end_j[k] = ej;
    }
    return 0;
#endif
}
