/*
 * vfind_oracle.c — CPU restatement of vFind's per-read variant-recovery path.
 * TEST INFRASTRUCTURE ONLY (see vfind_oracle.h for the parity status: the DP's
 * gapped/tie behaviour is "parity unpinned" because parasail is not buildable here).
 *
 * Citations are file:line into /root/reference/.
 */
#define _GNU_SOURCE
#include "vfind_oracle.h"

#include <limits.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <zlib.h>

/* ------------------------------------------------------------------ rules */

void vfo_default_rules(vfo_dp_rules *r)
{
    r->gap_tie_open = 0;
    r->h_priority = 0;
    r->end_rule = 0;
    r->wildcard_zero = 1;
}

/* ------------------------------------------------------------------ exact search */

/* memchr 2.7.4 memmem::find as called at src/lib.rs:148: leftmost, byte-exact,
 * case-sensitive.  An empty needle matches at 0 (Rust and memchr convention). */
int64_t vfo_memmem(const uint8_t *hay, size_t n, const uint8_t *needle, size_t m)
{
    if (m == 0) return 0;
    if (m > n) return VFO_NONE;
    for (size_t p = 0; p + m <= n; ++p) {
        if (hay[p] == needle[0] && memcmp(hay + p, needle, m) == 0) return (int64_t)p;
    }
    return VFO_NONE;
}

/* ------------------------------------------------------------------ thresholds */

/* src/lib.rs:100-110.  NaN fails both comparisons and falls through to the error. */
int vfo_threshold_preflight(double thr, int *skip)
{
    if (thr > 0. && thr < 1.) {
        *skip = 0;
        return 0;
    } else if (thr == 1.) {
        *skip = 1;
        return 0;
    }
    return -1;
}

/* src/lib.rs:260-261: `accept * match_score as f64 * adapter.len() as f64`
 * (left-associative: (thr*match)*len). */
double vfo_min_score(double thr, int32_t match_score, size_t adapter_len)
{
    volatile double a = thr * (double)match_score;
    volatile double b = a * (double)adapter_len;
    return b;
}

/* ------------------------------------------------------------------ DP */

#define NEG_INF (INT32_MIN / 2)

/* parasail_matrix_create(alphabet="ATCG", match, mismatch) (src/lib.rs:236): square
 * match/mismatch matrix over the alphabet, case-insensitive mapper, plus a wildcard
 * row/column that every other byte maps to.  [RECALLED] the wildcard scores 0. */
static inline int base_code(uint8_t b)
{
    switch (b) {
    case 'A': case 'a': return 0;
    case 'T': case 't': return 1;
    case 'C': case 'c': return 2;
    case 'G': case 'g': return 3;
    default: return 4;
    }
}

static inline int subst(int a, int b, int match, int mismatch, int wildcard_zero)
{
    if (a == 4 || b == 4) return wildcard_zero ? 0 : mismatch;
    return a == b ? match : mismatch;
}

/* Scalar sg_stats: the result parasail's test-suite forces every vectorised variant
 * (incl. sg_stats_scan_profile_sat, the one src/lib.rs:128-135 selects) to reproduce.
 * Row-major over the adapter (s1, i) with the read (s2, j) as the inner loop. */
int vfo_sg_stats(const uint8_t *adapter, int A, const uint8_t *read, int L,
                 int match, int mismatch, int open, int extend,
                 const vfo_dp_rules *rules,
                 int *score_out, int *length_out, int *end_i_out, int *end_j_out)
{
    vfo_dp_rules dr;
    if (!rules) {
        vfo_default_rules(&dr);
        rules = &dr;
    }
    if (A <= 0 || L <= 0) return -1;

    size_t cols = (size_t)L + 1;
    int *Hrow = (int *)malloc(sizeof(int) * cols * 4);
    int *lastcolH = (int *)malloc(sizeof(int) * ((size_t)A + 1) * 2);
    uint8_t *rc = (uint8_t *)malloc((size_t)L);
    if (!Hrow || !lastcolH || !rc) {
        free(Hrow); free(lastcolH); free(rc);
        return -1;
    }
    int *H = Hrow, *HL = Hrow + cols, *F = Hrow + 2 * cols, *FL = Hrow + 3 * cols;
    int *colH = lastcolH, *colHL = lastcolH + (A + 1);
    for (int j = 0; j < L; ++j) rc[j] = (uint8_t)base_code(read[j]);

    /* first row: H = 0 (s2 begin free), stats 0, F = -inf */
    for (int j = 0; j <= L; ++j) {
        H[j] = 0; HL[j] = 0; F[j] = NEG_INF; FL[j] = 0;
    }
    for (int i = 1; i <= A; ++i) {
        int ac = base_code(adapter[i - 1]);
        int NH = H[0], NHL = HL[0];    /* H[i-1][0] */
        int WH = 0, WHL = 0;           /* H[i][0] = 0 (s1 begin free) */
        int E = NEG_INF, EL = 0;
        H[0] = WH; HL[0] = WHL;
        for (int j = 1; j <= L; ++j) {
            int NWH = NH, NWL = NHL;   /* H[i-1][j-1] */
            NH = H[j]; NHL = HL[j];    /* H[i-1][j]   */
            /* F: vertical gap (consumes an adapter base) */
            int F_opn = NH - open, F_ext = F[j] - extend;
            int f_take_open = rules->gap_tie_open ? (F_opn >= F_ext) : (F_opn > F_ext);
            if (f_take_open) { F[j] = F_opn; FL[j] = NHL + 1; }
            else             { F[j] = F_ext; FL[j] = FL[j] + 1; }
            /* E: horizontal gap (consumes a read base) */
            int E_opn = WH - open, E_ext = E - extend;
            int e_take_open = rules->gap_tie_open ? (E_opn >= E_ext) : (E_opn > E_ext);
            if (e_take_open) { E = E_opn; EL = WHL + 1; }
            else             { E = E_ext; EL = EL + 1; }
            int D = NWH + subst(ac, rc[j - 1], match, mismatch, rules->wildcard_zero);
            int Fv = F[j];
            if (D >= E && D >= Fv) { WH = D; WHL = NWL + 1; }
            else if (rules->h_priority == 0) {
                if (Fv >= E) { WH = Fv; WHL = FL[j]; } else { WH = E; WHL = EL; }
            } else {
                if (E >= Fv) { WH = E; WHL = EL; } else { WH = Fv; WHL = FL[j]; }
            }
            H[j] = WH; HL[j] = WHL;
        }
        colH[i] = WH; colHL[i] = WHL;  /* last column value of this row */
    }

    /* end-cell choice (sg: both s1 end and s2 end free) */
    int score = NEG_INF, length = 0, ei = A, ej = 0;
    if (rules->end_rule == 3) {
        /* candidate order "column inside the row loop": the last-column cells of rows 1..A-1 are seen first
           (strict >), the last row afterwards (strict >, the corner included), so a last-column cell that ties
           the best last-row cell wins */
        for (int i = 1; i < A; ++i)
            if (colH[i] > score) { score = colH[i]; length = colHL[i]; ei = i; ej = L; }
    }
    for (int j = 1; j <= L; ++j) {
        int better = rules->end_rule == 1 ? (H[j] >= score) : (H[j] > score);
        if (better) { score = H[j]; length = HL[j]; ei = A; ej = j; }
    }
    if (rules->end_rule != 2 && rules->end_rule != 3) {
        int cbest = NEG_INF, ci = 0;
        for (int i = 1; i <= A; ++i)
            if (colH[i] > cbest) { cbest = colH[i]; ci = i; }
        if (cbest > score || (cbest == score && ej == L)) {
            score = cbest; length = colHL[ci]; ei = ci; ej = L;
        }
    }
    *score_out = score;
    *length_out = length;
    if (end_i_out) *end_i_out = ei;
    if (end_j_out) *end_j_out = ej;
    free(Hrow); free(lastcolH); free(rc);
    return 0;
}

/* ------------------------------------------------------------------ find_adapter_match */

/* src/lib.rs:141-166 */
int64_t vfo_find_adapter_match(const uint8_t *seq, size_t n, const uint8_t *adapter, size_t m,
                               const vfo_params *p, int align_enabled, double min_score,
                               int is_prefix, int32_t *exact_pos, int32_t *score_o, int32_t *len_o)
{
    int64_t pos = vfo_memmem(seq, n, adapter, m);                     /* :148 */
    if (exact_pos) *exact_pos = (int32_t)pos;
    if (score_o) *score_o = INT32_MIN;
    if (len_o) *len_o = -1;
    if (pos != VFO_NONE) {
        return is_prefix ? pos + (int64_t)m : pos;                     /* :151-152 */
    }
    if (!align_enabled) return VFO_NONE;                               /* `aligner?` :155 */
    int score, length;
    if (vfo_sg_stats(adapter, (int)m, seq, (int)n, p->match_score, p->mismatch_score,
                     p->gap_open_penalty, p->gap_extend_penalty, &p->rules,
                     &score, &length, NULL, NULL) != 0)
        return VFO_NONE;   /* empty read/adapter: the reference would panic (Q10) */
    if (score_o) *score_o = score;
    if (len_o) *len_o = length;
    if ((double)score > min_score) {                                   /* :157 */
        if (is_prefix) return (int64_t)length;                         /* :159 */
        if ((size_t)length > n) return VFO_NONE;                       /* :160 underflow (Q8) -> no region */
        return (int64_t)n - (int64_t)length;
    }
    return VFO_NONE;
}

/* ------------------------------------------------------------------ translate */

/* AA_TABLE_CANONICAL (src/lib.rs:52-77) flattened: index = c1*16 + c2*4 + c3 with
 * A=0, C=1, G=2, T/U=3. */
static const char AA_TABLE[65] =
    "KNKN" "TTTT" "RSRS" "IIMI"
    "QHQH" "PPPP" "RRRR" "LLLL"
    "EDED" "AAAA" "GGGG" "VVVV"
    "*Y*Y" "SSSS" "*CWC" "LFLF";

static uint8_t ASCII_TO_INDEX[128];
static pthread_once_t tables_once = PTHREAD_ONCE_INIT;

/* ASCII_TO_INDEX (src/lib.rs:86-95) */
static void init_tables(void)
{
    memset(ASCII_TO_INDEX, 4, sizeof ASCII_TO_INDEX);
    ASCII_TO_INDEX['A'] = ASCII_TO_INDEX['a'] = 0;
    ASCII_TO_INDEX['C'] = ASCII_TO_INDEX['c'] = 1;
    ASCII_TO_INDEX['G'] = ASCII_TO_INDEX['g'] = 2;
    ASCII_TO_INDEX['T'] = ASCII_TO_INDEX['t'] = 3;
    ASCII_TO_INDEX['U'] = ASCII_TO_INDEX['u'] = 3;
}

const char *vfo_aa_table(void) { return AA_TABLE; }
const uint8_t *vfo_ascii_to_index(void)
{
    pthread_once(&tables_once, init_tables);
    return ASCII_TO_INDEX;
}

/* src/lib.rs:16-44 */
int64_t vfo_translate(const uint8_t *seq, size_t n, uint8_t *out)
{
    pthread_once(&tables_once, init_tables);
    if (n % 3 != 0) return -1;                                         /* :17-19 */
    size_t k = 0;
    for (size_t i = 0; i < n; i += 3) {
        const uint8_t *t = seq + i;
        if ((t[0] | t[1] | t[2]) & 0x80) {                             /* :24-29 non-ASCII */
            out[k++] = 'X';
            continue;
        }
        unsigned c1 = ASCII_TO_INDEX[t[0]], c2 = ASCII_TO_INDEX[t[1]], c3 = ASCII_TO_INDEX[t[2]];
        if (c1 == 4 || c2 == 4 || c3 == 4) out[k++] = 'X';             /* :35-36 */
        else out[k++] = (uint8_t)AA_TABLE[c1 * 16 + c2 * 4 + c3];      /* :38 */
    }
    return (int64_t)k;
}

/* Rust's String::from_utf8 (src/lib.rs:295): well-formed UTF-8 only — no overlongs,
 * no surrogates, nothing above U+10FFFF. */
int vfo_is_utf8(const uint8_t *s, size_t n)
{
    size_t i = 0;
    while (i < n) {
        uint8_t b = s[i];
        if (b < 0x80) { i++; continue; }
        if (b >= 0xC2 && b <= 0xDF) {
            if (i + 1 >= n || (s[i + 1] & 0xC0) != 0x80) return 0;
            i += 2;
        } else if (b >= 0xE0 && b <= 0xEF) {
            if (i + 2 >= n) return 0;
            uint8_t b1 = s[i + 1], b2 = s[i + 2];
            if ((b1 & 0xC0) != 0x80 || (b2 & 0xC0) != 0x80) return 0;
            if (b == 0xE0 && b1 < 0xA0) return 0;
            if (b == 0xED && b1 > 0x9F) return 0;
            i += 3;
        } else if (b >= 0xF0 && b <= 0xF4) {
            if (i + 3 >= n) return 0;
            uint8_t b1 = s[i + 1], b2 = s[i + 2], b3 = s[i + 3];
            if ((b1 & 0xC0) != 0x80 || (b2 & 0xC0) != 0x80 || (b3 & 0xC0) != 0x80) return 0;
            if (b == 0xF0 && b1 < 0x90) return 0;
            if (b == 0xF4 && b1 > 0x8F) return 0;
            i += 4;
        } else {
            return 0;
        }
    }
    return 1;
}

/* ------------------------------------------------------------------ count table */

typedef struct {
    uint64_t hash;
    uint64_t off;     /* into arena */
    uint32_t len;
    uint32_t used;
    uint64_t count;
} vfo_slot;

struct vfo_table {
    vfo_slot *slots;
    uint64_t cap;     /* power of two */
    uint64_t rows;
    uint8_t *arena;
    uint64_t arena_len, arena_cap;
};

static uint64_t fnv1a(const uint8_t *s, size_t n)
{
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) { h ^= s[i]; h *= 1099511628211ull; }
    h ^= h >> 29; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 32;
    return h;
}

vfo_table *vfo_table_new(void)
{
    vfo_table *t = (vfo_table *)calloc(1, sizeof *t);
    t->cap = 1024;
    t->slots = (vfo_slot *)calloc(t->cap, sizeof(vfo_slot));
    t->arena_cap = 1 << 16;
    t->arena = (uint8_t *)malloc(t->arena_cap);
    return t;
}

void vfo_table_free(vfo_table *t)
{
    if (!t) return;
    free(t->slots);
    free(t->arena);
    free(t);
}

static void table_grow(vfo_table *t)
{
    uint64_t ncap = t->cap * 2;
    vfo_slot *ns = (vfo_slot *)calloc(ncap, sizeof(vfo_slot));
    for (uint64_t i = 0; i < t->cap; ++i) {
        if (!t->slots[i].used) continue;
        uint64_t p = t->slots[i].hash & (ncap - 1);
        while (ns[p].used) p = (p + 1) & (ncap - 1);
        ns[p] = t->slots[i];
    }
    free(t->slots);
    t->slots = ns;
    t->cap = ncap;
}

/* `*variants.entry(variant).or_insert(0) += 1` (src/lib.rs:296, :301) */
void vfo_table_add(vfo_table *t, const uint8_t *key, size_t n, uint64_t count)
{
    if ((t->rows + 1) * 2 > t->cap) table_grow(t);
    uint64_t h = fnv1a(key, n);
    uint64_t p = h & (t->cap - 1);
    for (;;) {
        vfo_slot *s = &t->slots[p];
        if (!s->used) {
            if (t->arena_len + n > t->arena_cap) {
                while (t->arena_len + n > t->arena_cap) t->arena_cap *= 2;
                t->arena = (uint8_t *)realloc(t->arena, t->arena_cap);
            }
            memcpy(t->arena + t->arena_len, key, n);
            s->hash = h; s->off = t->arena_len; s->len = (uint32_t)n; s->used = 1; s->count = count;
            t->arena_len += n;
            t->rows++;
            return;
        }
        if (s->hash == h && s->len == n && memcmp(t->arena + s->off, key, n) == 0) {
            s->count += count;
            return;
        }
        p = (p + 1) & (t->cap - 1);
    }
}

uint64_t vfo_table_rows(const vfo_table *t) { return t->rows; }
uint64_t vfo_table_key_bytes(const vfo_table *t) { return t->arena_len; }

typedef struct { const uint8_t *arena; const vfo_slot *s; } sort_ent;

static int cmp_ent(const void *a, const void *b)
{
    const sort_ent *x = (const sort_ent *)a, *y = (const sort_ent *)b;
    uint32_t m = x->s->len < y->s->len ? x->s->len : y->s->len;
    int c = memcmp(x->arena + x->s->off, y->arena + y->s->off, m);
    if (c) return c;
    return (x->s->len > y->s->len) - (x->s->len < y->s->len);
}

void vfo_table_export(const vfo_table *t, uint64_t *offsets, uint8_t *data, uint64_t *counts)
{
    vfo_table_export_ex(t, offsets, data, counts, 1);
}

void vfo_table_export_ex(const vfo_table *t, uint64_t *offsets, uint8_t *data, uint64_t *counts, int sorted)
{
    sort_ent *e = (sort_ent *)malloc(sizeof(sort_ent) * (t->rows ? t->rows : 1));
    uint64_t k = 0;
    for (uint64_t i = 0; i < t->cap; ++i)
        if (t->slots[i].used) { e[k].arena = t->arena; e[k].s = &t->slots[i]; k++; }
    if (sorted) qsort(e, k, sizeof(sort_ent), cmp_ent);
    uint64_t o = 0;
    for (uint64_t i = 0; i < k; ++i) {
        offsets[i] = o;
        memcpy(data + o, t->arena + e[i].s->off, e[i].s->len);
        o += e[i].s->len;
        counts[i] = e[i].s->count;
    }
    offsets[k] = o;
    free(e);
}

/* ------------------------------------------------------------------ per-read path */

typedef struct {
    const vfo_params *p;
    const uint8_t *text;
    const uint32_t *off, *len;
    uint64_t lo, hi;
    int prefix_align, suffix_align;
    double min_prefix, min_suffix;
    vfo_table *table;
    vfo_read_diag *diag;
    uint64_t cells;
    int simd;
} work_t;

/* worker closure src/lib.rs:275-291 followed by the reducer closure :292-306 for one read.
 * The per-read output slot is None unless set here (the intended semantics; see Q7 in
 * SURVEY §8(a) for the reference's stale-slot hazard, which is NOT reproduced). */
static void do_read(work_t *w, uint64_t r, uint8_t *scratch)
{
    const vfo_params *p = w->p;
    const uint8_t *seq = w->text + w->off[r];
    size_t n = w->len[r];
    vfo_read_diag d;
    int64_t start = vfo_find_adapter_match(seq, n, p->prefix, p->prefix_len, p, w->prefix_align,
                                           w->min_prefix, 1, &d.exact_prefix, &d.score_prefix,
                                           &d.len_prefix);                         /* :278-279 */
    int64_t end = vfo_find_adapter_match(seq, n, p->suffix, p->suffix_len, p, w->suffix_align,
                                         w->min_suffix, 0, &d.exact_suffix, &d.score_suffix,
                                         &d.len_suffix);                           /* :280-286 */
    if (d.score_prefix != INT32_MIN) w->cells += (uint64_t)p->prefix_len * n;
    if (d.score_suffix != INT32_MIN) w->cells += (uint64_t)p->suffix_len * n;
    d.start = (int32_t)start;
    d.end = (int32_t)end;
    if (w->diag) w->diag[r] = d;
    if (start != VFO_NONE && end != VFO_NONE && start < end) {                     /* :288 */
        /* a prefix boundary from the length statistic can exceed the read (gapped
         * alignment on a short read): the reference's slice would panic; no region. */
        if ((size_t)end > n) return;
        const uint8_t *var = seq + start;
        size_t vn = (size_t)(end - start);
        if (p->skip_translation) {
            if (vfo_is_utf8(var, vn)) vfo_table_add(w->table, var, vn, 1);          /* :294-297 */
        } else {
            int64_t k = vfo_translate(var, vn, scratch);                           /* :300 */
            if (k >= 0) vfo_table_add(w->table, scratch, (size_t)k, 1);            /* :301 */
        }
    }
}

/* The same closures with the alignments of a block of reads gathered and run sixteen at a time (sg_stats_simd.c).
 * Per read nothing changes: exact search, alignment on a miss, accept test, region, reducer — only the ORDER in which
 * the alignments of a block are evaluated does, and they are independent. */
#define SIMD_BLOCK 1024
static void run_alignments(work_t *w, const uint32_t *list, int n_list, int is_prefix, vfo_read_diag *bd, uint64_t b0)
{
    const vfo_params *p = w->p;
    const uint8_t *ad = is_prefix ? p->prefix : p->suffix;
    const int A = (int)(is_prefix ? p->prefix_len : p->suffix_len);
    for (int g = 0; g < n_list; g += 16) {
        const int m = n_list - g < 16 ? n_list - g : 16;
        const uint8_t *rd[16];
        int ln[16], sc[16], al[16];
        for (int k = 0; k < m; ++k) {
            const uint64_t r = b0 + list[g + k];
            rd[k] = w->text + w->off[r];
            ln[k] = (int)w->len[r];
        }
        if (vfo_sg_stats_x16(ad, A, rd, ln, m, p->match_score, p->mismatch_score, p->gap_open_penalty,
                             p->gap_extend_penalty, &p->rules, sc, al, NULL, NULL) != 0) {
            for (int k = 0; k < m; ++k)       /* out of the 16-bit lanes' range: one at a time */
                if (vfo_sg_stats(ad, A, rd[k], ln[k], p->match_score, p->mismatch_score, p->gap_open_penalty,
                                 p->gap_extend_penalty, &p->rules, &sc[k], &al[k], NULL, NULL) != 0) {
                    sc[k] = INT32_MIN; al[k] = -1;
                }
        }
        for (int k = 0; k < m; ++k) {
            vfo_read_diag *d = &bd[list[g + k]];
            if (is_prefix) { d->score_prefix = sc[k]; d->len_prefix = al[k]; }
            else           { d->score_suffix = sc[k]; d->len_suffix = al[k]; }
        }
    }
}

/* the tail of vfo_find_adapter_match (src/lib.rs:149-165) from its parts */
static int64_t boundary_of(int64_t pos, size_t m, size_t n, int is_prefix, int32_t score, int32_t length, double min_score)
{
    if (pos != VFO_NONE) return is_prefix ? pos + (int64_t)m : pos;    /* :151-152 */
    if (score == INT32_MIN) return VFO_NONE;                           /* no alignment ran */
    if ((double)score > min_score) {                                   /* :157 */
        if (is_prefix) return (int64_t)length;                         /* :159 */
        if ((size_t)length > n) return VFO_NONE;                       /* :160 underflow (Q8) -> no region */
        return (int64_t)n - (int64_t)length;
    }
    return VFO_NONE;
}

/* the C library's vectorised search (the reference's memchr::memmem is a SIMD search too); same answer as vfo_memmem */
static inline int64_t fast_memmem(const uint8_t *hay, size_t n, const uint8_t *needle, size_t m)
{
    if (m == 0) return 0;
    if (m > n) return VFO_NONE;
    const uint8_t *q = (const uint8_t *)memmem(hay, n, needle, m);
    return q ? (int64_t)(q - hay) : VFO_NONE;
}

static void worker_blocks_simd(work_t *w, uint8_t *scratch)
{
    const vfo_params *p = w->p;
    vfo_read_diag bd[SIMD_BLOCK];
    uint32_t lp[SIMD_BLOCK], ls[SIMD_BLOCK];
    for (uint64_t b0 = w->lo; b0 < w->hi; b0 += SIMD_BLOCK) {
        const int nb = (int)(w->hi - b0 < SIMD_BLOCK ? w->hi - b0 : SIMD_BLOCK);
        int np = 0, ns = 0;
        for (int k = 0; k < nb; ++k) {
            const uint8_t *seq = w->text + w->off[b0 + k];
            const size_t n = w->len[b0 + k];
            vfo_read_diag *d = &bd[k];
            d->exact_prefix = (int32_t)fast_memmem(seq, n, p->prefix, p->prefix_len);  /* :148 */
            d->exact_suffix = (int32_t)fast_memmem(seq, n, p->suffix, p->suffix_len);
            d->score_prefix = d->score_suffix = INT32_MIN;
            d->len_prefix = d->len_suffix = -1;
            /* `aligner?` :155; an empty read or adapter never reaches the matrices (vfo_sg_stats refuses it) */
            if (d->exact_prefix == VFO_NONE && w->prefix_align && n > 0 && p->prefix_len > 0) lp[np++] = (uint32_t)k;
            if (d->exact_suffix == VFO_NONE && w->suffix_align && n > 0 && p->suffix_len > 0) ls[ns++] = (uint32_t)k;
        }
        run_alignments(w, lp, np, 1, bd, b0);
        run_alignments(w, ls, ns, 0, bd, b0);
        for (int k = 0; k < nb; ++k) {
            const uint64_t r = b0 + k;
            const uint8_t *seq = w->text + w->off[r];
            const size_t n = w->len[r];
            vfo_read_diag *d = &bd[k];
            const int64_t start = boundary_of(d->exact_prefix, p->prefix_len, n, 1, d->score_prefix, d->len_prefix, w->min_prefix);
            const int64_t end = boundary_of(d->exact_suffix, p->suffix_len, n, 0, d->score_suffix, d->len_suffix, w->min_suffix);
            if (d->score_prefix != INT32_MIN) w->cells += (uint64_t)p->prefix_len * n;
            if (d->score_suffix != INT32_MIN) w->cells += (uint64_t)p->suffix_len * n;
            d->start = (int32_t)start;
            d->end = (int32_t)end;
            if (w->diag) w->diag[r] = *d;
            if (start != VFO_NONE && end != VFO_NONE && start < end) {                 /* :288 */
                if ((size_t)end > n) continue;
                const uint8_t *var = seq + start;
                const size_t vn = (size_t)(end - start);
                if (p->skip_translation) {
                    if (vfo_is_utf8(var, vn)) vfo_table_add(w->table, var, vn, 1);      /* :294-297 */
                } else {
                    const int64_t kk = vfo_translate(var, vn, scratch);                /* :300 */
                    if (kk >= 0) vfo_table_add(w->table, scratch, (size_t)kk, 1);      /* :301 */
                }
            }
        }
    }
}

static void *worker_main(void *arg)
{
    work_t *w = (work_t *)arg;
    uint32_t maxlen = 0;
    for (uint64_t r = w->lo; r < w->hi; ++r)
        if (w->len[r] > maxlen) maxlen = w->len[r];
    uint8_t *scratch = (uint8_t *)malloc((size_t)maxlen / 3 + 8);
    if (w->simd) { worker_blocks_simd(w, scratch); vfo_simd_thread_release(); }
    else for (uint64_t r = w->lo; r < w->hi; ++r) do_read(w, r, scratch);
    free(scratch);
    return NULL;
}

static int setup_thresholds(const vfo_params *p, int *pa, int *sa, double *minp, double *mins)
{
    int skip;
    if (vfo_threshold_preflight(p->accept_prefix_alignment, &skip) != 0) return -1;  /* :239 */
    *pa = !skip;
    if (vfo_threshold_preflight(p->accept_suffix_alignment, &skip) != 0) return -1;  /* :249 */
    *sa = !skip;
    *minp = vfo_min_score(p->accept_prefix_alignment, p->match_score, p->prefix_len); /* :260 */
    *mins = vfo_min_score(p->accept_suffix_alignment, p->match_score, p->suffix_len); /* :261 */
    return 0;
}

int vfo_process_reads(const vfo_params *p, const uint8_t *text,
                      const uint32_t *off, const uint32_t *len, uint64_t n,
                      int n_threads, vfo_table *table, vfo_read_diag *diag, uint64_t *dp_cells)
{
    return vfo_process_reads_ex(p, text, off, len, n, n_threads, table, diag, dp_cells, 0);
}

int vfo_process_reads_ex(const vfo_params *p, const uint8_t *text,
                         const uint32_t *off, const uint32_t *len, uint64_t n,
                         int n_threads, vfo_table *table, vfo_read_diag *diag, uint64_t *dp_cells, unsigned flags)
{
    int pa, sa;
    double minp, mins;
    if (setup_thresholds(p, &pa, &sa, &minp, &mins) != 0) return -1;
    if (n_threads < 1) n_threads = 1;
    if ((uint64_t)n_threads > n) n_threads = n ? (int)n : 1;
    work_t *w = (work_t *)calloc((size_t)n_threads, sizeof(work_t));
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int t = 0; t < n_threads; ++t) {
        w[t].p = p; w[t].text = text; w[t].off = off; w[t].len = len;
        w[t].lo = n * (uint64_t)t / (uint64_t)n_threads;
        w[t].hi = n * (uint64_t)(t + 1) / (uint64_t)n_threads;
        w[t].prefix_align = pa; w[t].suffix_align = sa;
        w[t].min_prefix = minp; w[t].min_suffix = mins;
        w[t].diag = diag;
        w[t].simd = (flags & VFO_FLAG_SIMD) && vfo_simd_available();
        w[t].table = (t == 0) ? table : vfo_table_new();
    }
    if (n_threads == 1) {
        worker_main(&w[0]);
    } else {
        for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, worker_main, &w[t]);
        for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    }
    uint64_t cells = 0;
    for (int t = 0; t < n_threads; ++t) {
        cells += w[t].cells;
        if (t == 0) continue;
        vfo_table *s = w[t].table;
        for (uint64_t i = 0; i < s->cap; ++i)
            if (s->slots[i].used)
                vfo_table_add(table, s->arena + s->slots[i].off, s->slots[i].len, s->slots[i].count);
        vfo_table_free(s);
    }
    if (dp_cells) *dp_cells = cells;
    free(w);
    free(th);
    return 0;
}

/* ------------------------------------------------------------------ gz FASTQ ingest */

/* flate2 MultiGzDecoder (src/lib.rs:233): concatenated gzip members decode as one
 * stream; anything that is not a gzip member is an error (plain text is rejected).
 * An empty file decodes to nothing.  Returns malloc'd text or NULL. */
static uint8_t *inflate_all(const char *path, uint64_t *out_len, int *status)
{
    FILE *f = fopen(path, "rb");
    if (!f) { *status = -2; return NULL; }
    fseek(f, 0, SEEK_END);
    long fsz = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t *in = (uint8_t *)malloc(fsz > 0 ? (size_t)fsz : 1);
    if (fsz > 0 && fread(in, 1, (size_t)fsz, f) != (size_t)fsz) {
        fclose(f); free(in); *status = -2; return NULL;
    }
    fclose(f);
    uint64_t cap = 1 << 22, len = 0;
    uint8_t *buf = (uint8_t *)malloc(cap);
    uint64_t ip = 0;
    while (ip < (uint64_t)fsz) {
        z_stream z;
        memset(&z, 0, sizeof z);
        if (inflateInit2(&z, 15 + 16) != Z_OK) { free(in); free(buf); *status = -3; return NULL; }
        int rc = Z_OK;
        while (rc != Z_STREAM_END) {
            if (cap - len < (1u << 20)) { cap *= 2; buf = (uint8_t *)realloc(buf, cap); }
            uint64_t ain = (uint64_t)fsz - ip, aout = cap - len;
            z.next_in = in + ip;
            z.avail_in = (uInt)(ain > (1u << 30) ? (1u << 30) : ain);
            z.next_out = buf + len;
            z.avail_out = (uInt)(aout > (1u << 30) ? (1u << 30) : aout);
            uInt in0 = z.avail_in, out0 = z.avail_out;
            rc = inflate(&z, Z_NO_FLUSH);
            ip += in0 - z.avail_in;
            len += out0 - z.avail_out;
            if (rc == Z_STREAM_END) break;
            if (rc != Z_OK || (in0 == z.avail_in && out0 == z.avail_out)) {
                inflateEnd(&z); free(in); free(buf); *status = -3; return NULL;
            }
        }
        inflateEnd(&z);
    }
    free(in);
    *out_len = len;
    *status = 0;
    return buf;
}

/* seq_io 0.3.4 fastq::Reader grammar [RECALLED, SURVEY Q11]: four lines per record —
 * '@' header, one sequence line, '+' separator, quality of equal length; "\r\n" trimmed;
 * a missing final newline is tolerated; trailing blank lines are tolerated. */
static int parse_fastq(const uint8_t *t, uint64_t n, uint32_t **off_o, uint32_t **len_o,
                       uint64_t *count, char *err, size_t errlen)
{
    uint64_t cap = 1024, k = 0;
    uint32_t *off = (uint32_t *)malloc(cap * 4), *len = (uint32_t *)malloc(cap * 4);
    uint64_t p = 0;
    while (p < n) {
        /* tolerate trailing newlines / CRs at the end of input */
        uint64_t q = p;
        while (q < n && (t[q] == '\n' || t[q] == '\r')) q++;
        if (q == n) break;
        uint64_t ls[4], ll[4];
        for (int l = 0; l < 4; ++l) {
            if (p > n || (p == n && l < 3)) {
                snprintf(err, errlen, "truncated FASTQ record %llu", (unsigned long long)k);
                free(off); free(len);
                return -3;
            }
            const uint8_t *nl = p < n ? (const uint8_t *)memchr(t + p, '\n', n - p) : NULL;
            uint64_t e = nl ? (uint64_t)(nl - t) : n;
            if (!nl && l < 3) {
                snprintf(err, errlen, "truncated FASTQ record %llu", (unsigned long long)k);
                free(off); free(len);
                return -3;
            }
            uint64_t ee = e;
            if (ee > p && t[ee - 1] == '\r') ee--;
            ls[l] = p; ll[l] = ee - p;
            p = nl ? e + 1 : n + 1;
        }
        if (ll[0] == 0 || t[ls[0]] != '@') {
            snprintf(err, errlen, "FASTQ record %llu: expected '@'", (unsigned long long)k);
            free(off); free(len);
            return -3;
        }
        if (ll[2] == 0 || t[ls[2]] != '+') {
            snprintf(err, errlen, "FASTQ record %llu: expected '+'", (unsigned long long)k);
            free(off); free(len);
            return -3;
        }
        if (ll[1] != ll[3]) {
            snprintf(err, errlen, "FASTQ record %llu: sequence and quality lengths differ",
                     (unsigned long long)k);
            free(off); free(len);
            return -3;
        }
        if (ls[1] > UINT32_MAX) {
            snprintf(err, errlen, "oracle: input larger than 4 GiB is not supported");
            free(off); free(len);
            return -3;
        }
        if (k == cap) {
            cap *= 2;
            off = (uint32_t *)realloc(off, cap * 4);
            len = (uint32_t *)realloc(len, cap * 4);
        }
        off[k] = (uint32_t)ls[1];
        len[k] = (uint32_t)ll[1];
        k++;
    }
    *off_o = off; *len_o = len; *count = k;
    return 0;
}

int vfo_find_variants_file(const char *path, const vfo_params *p, int n_threads,
                           vfo_table *table, uint64_t *n_reads, char *err, size_t errlen)
{
    return vfo_find_variants_file_ex(path, p, n_threads, table, n_reads, err, errlen, 0, NULL);
}

static double now_s(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

int vfo_find_variants_file_ex(const char *path, const vfo_params *p, int n_threads,
                              vfo_table *table, uint64_t *n_reads, char *err, size_t errlen,
                              unsigned flags, double *phase_seconds)
{
    char dummy[8];
    double ph_dummy[3];
    if (!phase_seconds) phase_seconds = ph_dummy;
    phase_seconds[0] = phase_seconds[1] = phase_seconds[2] = 0.;
    if (!err) { err = dummy; errlen = sizeof dummy; }
    err[0] = 0;
    /* the reference opens the file first (:233) and validates thresholds after (:239) */
    FILE *f = fopen(path, "rb");
    if (!f) { snprintf(err, errlen, "cannot open %s", path); return -2; }
    fclose(f);
    int pa, sa;
    double a, b;
    if (setup_thresholds(p, &pa, &sa, &a, &b) != 0) {
        snprintf(err, errlen, "Accept alignment threshold must be between 0 and 1.");
        return -1;
    }
    int st = 0;
    uint64_t tlen = 0;
    double t0 = now_s();
    uint8_t *text = inflate_all(path, &tlen, &st);
    phase_seconds[0] = now_s() - t0;
    if (!text) {
        snprintf(err, errlen, st == -2 ? "cannot open %s" : "malformed gzip: %s", path);
        return st;
    }
    uint32_t *off = NULL, *len = NULL;
    uint64_t n = 0;
    t0 = now_s();
    st = parse_fastq(text, tlen, &off, &len, &n, err, errlen);
    phase_seconds[1] = now_s() - t0;
    if (st != 0) { free(text); return st; }
    t0 = now_s();
    st = vfo_process_reads_ex(p, text, off, len, n, n_threads, table, NULL, NULL, flags);
    phase_seconds[2] = now_s() - t0;
    if (n_reads) *n_reads = n;
    free(off); free(len); free(text);
    return st;
}
