// GPU inflate of block-gzip members (ingest, SURVEY §8(f) next-1: "GPU inflate for multi-member
// (BGZF-style) input").
//
// Replaces flate2's MultiGzDecoder (/root/reference/src/lib.rs:233) for input whose members
// carry their own compressed size (BGZF "BC" extra field) and are therefore independent streams
// of at most 64 KiB: one thread inflates one member, thousands of members per launch.  The
// decoder is a plain RFC 1951 decoder (stored / fixed / dynamic blocks, canonical Huffman codes
// decoded length by length); per-thread tables live in local memory, the four base/extra
// tables in shared memory.  Each member's CRC-32 and ISIZE (RFC 1952 trailer) are checked, as
// flate2 does; the first bad member is reported.
#include <cstdlib>

#include "vfb_internal.cuh"

namespace vfb {

#define INF_MAXBITS 15
#define INF_MAXL 288
#define INF_MAXD 30

struct Huff {
    short count[INF_MAXBITS + 1];
    short symbol[INF_MAXL];
};

struct InfTables {
    unsigned short lbase[29], lext[29], dbase[30], dext[30];
    unsigned char order[19];
    uint32_t crc[256];
};

struct BitReader {
    const uint8_t *in, *end;
    uint64_t buf;
    int cnt;
    bool over;     // ran past the end of the member
};

__device__ __forceinline__ void br_refill(BitReader &b)
{
    while (b.cnt <= 56) {
        if (b.in < b.end) b.buf |= (uint64_t)(*b.in++) << b.cnt;
        else if (b.in >= b.end + 8) { b.over = true; }
        else ++b.in;                                   // zero padding past the end, bounded
        b.cnt += 8;
    }
}

__device__ __forceinline__ uint32_t br_bits(BitReader &b, int n)
{
    if (b.cnt < n) br_refill(b);
    const uint32_t v = (uint32_t)(b.buf & ((1ull << n) - 1ull));
    b.buf >>= n;
    b.cnt -= n;
    return v;
}

// Canonical Huffman decode, one bit at a time (codes are stored MSB first).
__device__ __forceinline__ int huff_decode(BitReader &b, const Huff &h)
{
    if (b.cnt < INF_MAXBITS) br_refill(b);
    int code = 0, first = 0, index = 0;
    uint64_t bits = b.buf;
    for (int len = 1; len <= INF_MAXBITS; ++len) {
        code |= (int)(bits & 1u);
        bits >>= 1;
        const int count = h.count[len];
        if (code - count < first) {
            b.buf >>= len;
            b.cnt -= len;
            return h.symbol[index + (code - first)];
        }
        index += count;
        first += count;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

// Build a canonical code from code lengths.  Returns <0 if over-subscribed, >0 if incomplete.
__device__ int huff_construct(Huff &h, const short *length, int n)
{
    short offs[INF_MAXBITS + 1];
    for (int len = 0; len <= INF_MAXBITS; ++len) h.count[len] = 0;
    for (int s = 0; s < n; ++s) h.count[length[s]]++;
    if (h.count[0] == n) return 0;
    int left = 1;
    for (int len = 1; len <= INF_MAXBITS; ++len) {
        left <<= 1;
        left -= h.count[len];
        if (left < 0) return left;
    }
    offs[1] = 0;
    for (int len = 1; len < INF_MAXBITS; ++len) offs[len + 1] = offs[len] + h.count[len];
    for (int s = 0; s < n; ++s)
        if (length[s] != 0) h.symbol[offs[length[s]]++] = (short)s;
    return left;
}

struct InflateArgs {
    const uint8_t *z;            // compressed members, back to back as read from the file
    const vfb_member *members;   // per member: offset/size in z, offset in out, ISIZE
    uint32_t n_members;
    uint8_t *out;
    uint32_t *first_bad;         // atomicMin of the first member that failed
    uint32_t lanes;              // lanes of each warp that decode (1 = lane 0 only)
};

// error codes are only used to tell "ok" from "bad"
__device__ int inflate_member(const InfTables *T, const uint8_t *zin, uint32_t zlen, uint8_t *out, uint32_t isize)
{
    // gzip header (RFC 1952): fixed 10 bytes, optional FEXTRA / FNAME / FCOMMENT / FHCRC
    if (zlen < 18 || zin[0] != 0x1f || zin[1] != 0x8b || zin[2] != 8) return -1;
    const uint32_t flg = zin[3];
    uint32_t p = 10;
    if (flg & 4) { if (p + 2 > zlen) return -1; p += 2 + (zin[p] | (zin[p + 1] << 8)); }
    if (flg & 8) { while (p < zlen && zin[p]) ++p; ++p; }
    if (flg & 16) { while (p < zlen && zin[p]) ++p; ++p; }
    if (flg & 2) p += 2;
    if (p + 8 > zlen) return -1;
    const uint8_t *tr = zin + zlen - 8;
    const uint32_t want_crc = tr[0] | (tr[1] << 8) | (tr[2] << 16) | ((uint32_t)tr[3] << 24);
    const uint32_t want_len = tr[4] | (tr[5] << 8) | (tr[6] << 16) | ((uint32_t)tr[7] << 24);
    if (want_len != isize) return -2;

    BitReader b;
    b.in = zin + p; b.end = zin + zlen - 8; b.buf = 0; b.cnt = 0; b.over = false;
    Huff lencode, distcode;
    short lengths[INF_MAXL + INF_MAXD];
    uint32_t n_out = 0, crc = 0xFFFFFFFFu;
    int last;
    do {
        last = (int)br_bits(b, 1);
        const uint32_t type = br_bits(b, 2);
        if (type == 0) {
            // stored: skip to a byte boundary, LEN, NLEN, bytes
            const int drop = b.cnt & 7;
            b.buf >>= drop; b.cnt -= drop;
            const uint32_t len = br_bits(b, 16), nlen = br_bits(b, 16);
            if ((len ^ 0xFFFFu) != nlen) return -3;
            if (n_out + len > isize) return -4;
            for (uint32_t i = 0; i < len; ++i) {
                const uint32_t c = br_bits(b, 8);
                out[n_out++] = (uint8_t)c;
                crc = T->crc[(crc ^ c) & 0xFFu] ^ (crc >> 8);
            }
        } else if (type == 1 || type == 2) {
            if (type == 1) {
                int s = 0;
                for (; s < 144; ++s) lengths[s] = 8;
                for (; s < 256; ++s) lengths[s] = 9;
                for (; s < 280; ++s) lengths[s] = 7;
                for (; s < 288; ++s) lengths[s] = 8;
                huff_construct(lencode, lengths, 288);
                for (s = 0; s < 30; ++s) lengths[s] = 5;
                huff_construct(distcode, lengths, 30);
            } else {
                const int nlen = (int)br_bits(b, 5) + 257, ndist = (int)br_bits(b, 5) + 1, ncode = (int)br_bits(b, 4) + 4;
                if (nlen > 286 || ndist > 30) return -5;
                int idx = 0;
                for (; idx < ncode; ++idx) lengths[T->order[idx]] = (short)br_bits(b, 3);
                for (; idx < 19; ++idx) lengths[T->order[idx]] = 0;
                if (huff_construct(lencode, lengths, 19) != 0) return -6;     // must be complete
                idx = 0;
                while (idx < nlen + ndist) {
                    int sym = huff_decode(b, lencode);
                    if (sym < 0) return -7;
                    if (sym < 16) {
                        lengths[idx++] = (short)sym;
                    } else {
                        int len = 0, rep;
                        if (sym == 16) {
                            if (idx == 0) return -8;
                            len = lengths[idx - 1];
                            rep = 3 + (int)br_bits(b, 2);
                        } else if (sym == 17) rep = 3 + (int)br_bits(b, 3);
                        else rep = 11 + (int)br_bits(b, 7);
                        if (idx + rep > nlen + ndist) return -9;
                        while (rep--) lengths[idx++] = (short)len;
                    }
                }
                if (lengths[256] == 0) return -10;
                int err = huff_construct(lencode, lengths, nlen);
                if (err < 0 || (err > 0 && nlen - lencode.count[0] != 1)) return -11;
                err = huff_construct(distcode, lengths + nlen, ndist);
                if (err < 0 || (err > 0 && ndist - distcode.count[0] != 1)) return -12;
            }
            // decode literals / length-distance pairs until end-of-block
            for (;;) {
                int sym = huff_decode(b, lencode);
                if (sym < 0) return -13;
                if (sym < 256) {
                    if (n_out >= isize) return -14;
                    out[n_out++] = (uint8_t)sym;
                    crc = T->crc[(crc ^ (uint32_t)sym) & 0xFFu] ^ (crc >> 8);
                } else if (sym == 256) {
                    break;
                } else {
                    sym -= 257;
                    if (sym >= 29) return -15;
                    const uint32_t len = T->lbase[sym] + br_bits(b, T->lext[sym]);
                    const int ds = huff_decode(b, distcode);
                    if (ds < 0 || ds >= 30) return -16;
                    const uint32_t dist = T->dbase[ds] + br_bits(b, T->dext[ds]);
                    if (dist > n_out) return -17;
                    if (n_out + len > isize) return -18;
                    for (uint32_t i = 0; i < len; ++i) {
                        const uint8_t c = out[n_out - dist];
                        out[n_out++] = c;
                        crc = T->crc[(crc ^ c) & 0xFFu] ^ (crc >> 8);
                    }
                }
                if (b.over) return -19;
            }
        } else {
            return -20;
        }
        if (b.over) return -19;
    } while (!last);
    if (n_out != isize) return -21;
    if ((crc ^ 0xFFFFFFFFu) != want_crc) return -22;
    return 0;
}

#define INF_THREADS 128

__global__ void __launch_bounds__(INF_THREADS)
k_inflate_members(const __grid_constant__ InflateArgs a)
{
    __shared__ InfTables T;
    if (threadIdx.x == 0) {
        const unsigned short lb[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
        const unsigned short le[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
        const unsigned short db[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
        const unsigned short de[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
        const unsigned char od[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        for (int i = 0; i < 29; ++i) { T.lbase[i] = lb[i]; T.lext[i] = le[i]; }
        for (int i = 0; i < 30; ++i) { T.dbase[i] = db[i]; T.dext[i] = de[i]; }
        for (int i = 0; i < 19; ++i) T.order[i] = od[i];
    }
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        uint32_t c = (uint32_t)i;
        for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        T.crc[i] = c;
    }
    __syncthreads();
    // Decoding is branchy and every stream takes its own path: lanes of a warp that decode
    // different members serialise each other.  So only `lanes` lanes per warp decode (default 1:
    // a warp is one independent decoder and the SM interleaves warps instead of lanes).
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (lane >= a.lanes) return;
    const uint32_t i = warp * a.lanes + lane;
    if (i >= a.n_members) return;
    const vfb_member m = a.members[i];
    int rc = 0;
    if (m.isize || m.z_len) rc = inflate_member(&T, a.z + m.z_off, m.z_len, a.out + m.out_off, m.isize);
    if (rc != 0) atomicMin(a.first_bad, i);
}

// =====================================================================================
// v2: one WARP per member, table-driven.
//
// All 32 lanes run the same decode (same bit buffer, same tables: no divergence, no broadcasts);
// the lanes differ only where a warp helps: the compressed input arrives through a 512-byte
// shared-memory ring that the lanes top up with coalesced 128-byte loads three lines ahead of
// the bit reader, Huffman decode is one shared-memory lookup per symbol (10-bit table for
// literal/length codes, 8-bit for distances, built by all lanes from the canonical code; longer
// codes fall back to the bit-by-bit canonical decode), a match is copied by all lanes at once
// (period `dist` when it overlaps itself), and the CRC-32 is computed after the decode, each
// lane over 1/32 of the output, the partial CRCs combined with x^(8n) mod P multiplications.
#define INFW_WARPS 4
#define INFW_LBITS 9
#define INFW_DBITS 8

struct InfWarp {
    uint32_t ltab[1 << INFW_LBITS];   // nbits | kind << 4 | value << 8 | extra << 24 ; 0 = long / unused code
    uint32_t dtab[1 << INFW_DBITS];   // nbits | extra << 4 | base << 8
    uint16_t lcount[16], dcount[16];  // canonical code: symbols per length
    uint16_t lsym[INF_MAXL], dsym[32];
    uint8_t lens[32 + INF_MAXL + 32];   // code lengths being read (staged 32 bytes up while the code-length code is in use)
    uint32_t ring[128 + 4];           // 4 lines of 32 compressed words; ring[128] mirrors ring[0] (the symbol loop reads
                                      // word pairs without wrapping the second index)
};

struct InfShared {
    InfTables T;
    uint32_t x2n[32];                 // x^(2^n) mod P (reflected CRC-32 polynomial)
    InfWarp w[INFW_WARPS];
};

__device__ __forceinline__ uint32_t crc_multmodp(uint32_t a, uint32_t b)
{
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1)) == 0) break;
        }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ 0xEDB88320u : b >> 1;
    }
    return p;
}

// x^(n * 2^k) mod P
__device__ uint32_t crc_x2nmodp(const uint32_t *x2n, uint32_t n, uint32_t k)
{
    uint32_t p = 1u << 31;
    while (n) {
        if (n & 1u) p = crc_multmodp(x2n[k & 31], p);
        n >>= 1;
        ++k;
    }
    return p;
}

struct WBits {
    uint64_t buf;
    int cnt;
    uint32_t rw;          // next word (relative to `words`) to enter the bit buffer
    uint32_t next_line;   // next 32-word line to load into the ring
    const uint32_t *words;
    uint32_t nwords;
};

__device__ __forceinline__ void wb_refill(WBits &b, InfWarp *W, int lane)
{
    if (b.cnt > 32) return;
    // (two lines ahead of the word that enters next: the ring's four lines then still hold the line before it, which
    // the symbol loop — it works from a bit position, not from this buffer — may have to read again)
    while (b.next_line <= (b.rw >> 5) + 2) {
        const uint32_t w = b.next_line * 32u + (uint32_t)lane;
        const uint32_t v = w < b.nwords ? __ldg(b.words + w) : 0u;
        W->ring[w & 127u] = v;
        if ((w & 127u) == 0u) W->ring[128] = v;
        ++b.next_line;
        __syncwarp();
    }
    b.buf |= (uint64_t)W->ring[b.rw & 127u] << b.cnt;
    b.cnt += 32;
    ++b.rw;
}

// Start reading at byte `bytepos` (relative to `words`).
__device__ __forceinline__ void wb_seek(WBits &b, InfWarp *W, int lane, uint32_t bytepos)
{
    __syncwarp();
    b.rw = bytepos >> 2;
    b.next_line = b.rw >> 5;
    b.buf = 0;
    b.cnt = 0;
    wb_refill(b, W, lane);
    const int skip = (int)(bytepos & 3u) * 8;
    b.buf >>= skip;
    b.cnt -= skip;
}

__device__ __forceinline__ uint32_t wb_bits(WBits &b, InfWarp *W, int lane, int n)   // n <= 16
{
    wb_refill(b, W, lane);
    const uint32_t v = (uint32_t)b.buf & ((1u << n) - 1u);
    b.buf >>= n;
    b.cnt -= n;
    return v;
}

// bytes consumed so far, rounded up to whole bytes
__device__ __forceinline__ uint32_t wb_bytepos(const WBits &b) { return b.rw * 4u - (uint32_t)(b.cnt >> 3); }

// bit-by-bit canonical decode from the low bits of `bits` (codes are MSB first); returns the
// symbol and its length, or -1
__device__ __forceinline__ int canon_decode(uint32_t bits, const uint16_t *count, const uint16_t *symbol, int maxlen, int *len_out)
{
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= maxlen; ++len) {
        code |= (int)(bits & 1u);
        bits >>= 1;
        const int c = count[len];
        if (code - c < first) {
            *len_out = len;
            return symbol[index + (code - first)];
        }
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

// canonical code from code lengths (one lane); <0 over-subscribed, >0 incomplete, 0 complete
__device__ int canon_construct(uint16_t *count, uint16_t *symbol, const uint8_t *length, int n)
{
    uint16_t offs[INF_MAXBITS + 1];
    for (int len = 0; len <= INF_MAXBITS; ++len) count[len] = 0;
    for (int s = 0; s < n; ++s) count[length[s]]++;
    if (count[0] == n) return 0;
    int left = 1;
    for (int len = 1; len <= INF_MAXBITS; ++len) {
        left <<= 1;
        left -= count[len];
        if (left < 0) return left;
    }
    offs[1] = 0;
    for (int len = 1; len < INF_MAXBITS; ++len) offs[len + 1] = offs[len] + count[len];
    for (int s = 0; s < n; ++s)
        if (length[s] != 0) symbol[offs[length[s]]++] = (uint16_t)s;
    return left;
}

__device__ __forceinline__ uint32_t lit_entry(const InfTables *T, int sym, int nbits)
{
    if (sym < 256) return (uint32_t)nbits | ((uint32_t)sym << 8);
    if (sym == 256) return (uint32_t)nbits | (2u << 4);
    if (sym > 285) return (uint32_t)nbits | (3u << 4);                         // invalid length symbol
    return (uint32_t)nbits | (1u << 4) | ((uint32_t)T->lbase[sym - 257] << 8) | ((uint32_t)T->lext[sym - 257] << 24);
}

__device__ __forceinline__ uint32_t dist_entry(const InfTables *T, int sym, int nbits)
{
    if (sym > 29) return (uint32_t)nbits | (15u << 4);                         // invalid distance symbol (extra 15 never occurs)
    return (uint32_t)nbits | ((uint32_t)T->dext[sym] << 4) | ((uint32_t)T->dbase[sym] << 8);
}

// lengths of the literal/length code in W->lens[0..nlen), of the distance code behind them
__device__ int infw_build(const InfTables *T, InfWarp *W, int nlen, int ndist, int lane, bool fixed)
{
    int e1 = 0, e2 = 0;
    if (lane == 0) {
        e1 = canon_construct(W->lcount, W->lsym, W->lens, nlen);
        if (!(e1 < 0 || (e1 > 0 && nlen - W->lcount[0] != 1))) e1 = 0; else e1 = 1;
        e2 = canon_construct(W->dcount, W->dsym, W->lens + nlen, ndist);
        if (!(e2 < 0 || (e2 > 0 && ndist - W->dcount[0] != 1))) e2 = 0; else e2 = 1;
        if (fixed) e1 = e2 = 0;            // the fixed distance code leaves two codes unused (RFC 1951 3.2.6)
    }
    e1 = __shfl_sync(0xffffffffu, e1, 0);
    e2 = __shfl_sync(0xffffffffu, e2, 0);
    if (e1) return -11;
    if (e2) return -12;
    __syncwarp();
    for (int e = lane; e < (1 << INFW_LBITS); e += 32) {
        int len = 0;
        const int sym = canon_decode((uint32_t)e, W->lcount, W->lsym, INFW_LBITS, &len);
        W->ltab[e] = sym < 0 ? 0u : lit_entry(T, sym, len);
    }
    for (int e = lane; e < (1 << INFW_DBITS); e += 32) {
        int len = 0;
        const int sym = canon_decode((uint32_t)e, W->dcount, W->dsym, INFW_DBITS, &len);
        W->dtab[e] = sym < 0 ? 0u : dist_entry(T, sym, len);
    }
    __syncwarp();
    // Literal runs: where the index bits behind a literal's code hold one or two more complete literal
    // codes, the entry decodes them all at once -- bytes in bits 8-15 / 16-23 / 24-31, (count - 1) in bits
    // 6-7, total code length in bits 0-3.  Read text is mostly literals with 2-3 bit codes (the bases), so
    // most look-ups then yield three bytes.  Every lane first derives its entries from the single-symbol
    // table, then all write.
    uint32_t fused[(1 << INFW_LBITS) / 32];
#pragma unroll
    for (int k = 0; k < (1 << INFW_LBITS) / 32; ++k) {
        const uint32_t e = (uint32_t)lane + 32u * (uint32_t)k;
        uint32_t a = W->ltab[e];
        if (a != 0 && ((a >> 4) & 3u) == 0) {
            uint32_t used = a & 15u, cnt = 1, bytes = (a >> 8) & 0xFFu;
            while (cnt < 3 && used < INFW_LBITS) {
                const uint32_t nx = W->ltab[e >> used];             // the bits above `used`, zero-extended
                const uint32_t nb = nx & 15u;
                if (nx == 0 || ((nx >> 4) & 3u) != 0 || used + nb > INFW_LBITS) break;   // not a literal fully inside the index
                bytes |= ((nx >> 8) & 0xFFu) << (8 * cnt);
                used += nb;
                ++cnt;
            }
            a = used | ((cnt - 1) << 6) | (bytes << 8);
        }
        fused[k] = a;
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < (1 << INFW_LBITS) / 32; ++k) W->ltab[lane + 32 * k] = fused[k];
    __syncwarp();
    return 0;
}

__device__ int inflate_member_warp(const InfShared *S, InfWarp *W, const uint8_t *zin, uint32_t zlen, uint8_t *out,
                                   uint32_t isize, int lane)
{
    const InfTables *T = &S->T;
    if (zlen < 18 || zin[0] != 0x1f || zin[1] != 0x8b || zin[2] != 8) return -1;
    const uint32_t flg = zin[3];
    uint32_t p = 10;
    if (flg & 4) { if (p + 2 > zlen) return -1; p += 2 + (zin[p] | (zin[p + 1] << 8)); }
    if (flg & 8) { while (p < zlen && zin[p]) ++p; ++p; }
    if (flg & 16) { while (p < zlen && zin[p]) ++p; ++p; }
    if (flg & 2) p += 2;
    if (p + 8 > zlen) return -1;
    const uint8_t *tr = zin + zlen - 8;
    const uint32_t want_crc = tr[0] | (tr[1] << 8) | (tr[2] << 16) | ((uint32_t)tr[3] << 24);
    const uint32_t want_len = tr[4] | (tr[5] << 8) | (tr[6] << 16) | ((uint32_t)tr[7] << 24);
    if (want_len != isize) return -2;

    WBits b;
    const uintptr_t a0 = reinterpret_cast<uintptr_t>(zin + p);
    b.words = reinterpret_cast<const uint32_t *>(a0 & ~(uintptr_t)3);
    const uint32_t lead = (uint32_t)(a0 & 3u);
    const uint32_t data_len = zlen - 8 - p;                  // deflate stream bytes
    b.nwords = (lead + data_len + 8 + 3) >> 2;               // (the trailer may be looked at, never beyond)
    wb_seek(b, W, lane, lead);
    uint32_t n_out = 0;
    int last;
    do {
        last = (int)wb_bits(b, W, lane, 1);
        const uint32_t type = wb_bits(b, W, lane, 2);
        if (type == 0) {
            const int drop = b.cnt & 7;
            b.buf >>= drop; b.cnt -= drop;
            const uint32_t len = wb_bits(b, W, lane, 16), nlen = wb_bits(b, W, lane, 16);
            if ((len ^ 0xFFFFu) != nlen) return -3;
            if (n_out + len > isize) return -4;
            const uint32_t src = wb_bytepos(b);              // byte aligned here
            if (src + len > lead + data_len) return -19;
            const uint8_t *sp = reinterpret_cast<const uint8_t *>(b.words) + src;
            for (uint32_t i = lane; i < len; i += 32) out[n_out + i] = __ldg(sp + i);
            n_out += len;
            wb_seek(b, W, lane, src + len);
        } else if (type == 1 || type == 2) {
            int nl, nd;
            if (type == 1) {
                nl = 288; nd = 30;
                for (int s = lane; s < 288; s += 32) W->lens[s] = (uint8_t)(s < 144 ? 8 : (s < 256 ? 9 : (s < 280 ? 7 : 8)));
                if (lane < 30) W->lens[288 + lane] = 5;
                __syncwarp();
            } else {
                nl = (int)wb_bits(b, W, lane, 5) + 257;
                nd = (int)wb_bits(b, W, lane, 5) + 1;
                const int ncode = (int)wb_bits(b, W, lane, 4) + 4;
                if (nl > 286 || nd > 30) return -5;
                __syncwarp();
                if (lane < 19) W->lens[lane] = 0;
                __syncwarp();
                for (int idx = 0; idx < ncode; ++idx) {
                    const uint32_t v = wb_bits(b, W, lane, 3);
                    if (lane == 0) W->lens[T->order[idx]] = (uint8_t)v;
                }
                __syncwarp();
                int e0 = 0;
                if (lane == 0) e0 = canon_construct(W->lcount, W->lsym, W->lens, 19);
                e0 = __shfl_sync(0xffffffffu, e0, 0);
                if (e0 != 0) return -6;                          // must be complete
                __syncwarp();
                int idx = 0;
                while (idx < nl + nd) {
                    wb_refill(b, W, lane);
                    int cl = 0;
                    const int sym = canon_decode((uint32_t)b.buf, W->lcount, W->lsym, 7, &cl);
                    if (sym < 0) return -7;
                    b.buf >>= cl; b.cnt -= cl;
                    if (sym < 16) {
                        if (lane == 0) W->lens[32 + idx] = (uint8_t)sym;      // (kept clear of the 19 code-length lengths)
                        ++idx;
                    } else {
                        int len = 0, rep;
                        if (sym == 16) {
                            if (idx == 0) return -8;
                            __syncwarp();
                            len = W->lens[32 + idx - 1];
                            rep = 3 + (int)wb_bits(b, W, lane, 2);
                        } else if (sym == 17) rep = 3 + (int)wb_bits(b, W, lane, 3);
                        else rep = 11 + (int)wb_bits(b, W, lane, 7);
                        if (idx + rep > nl + nd) return -9;
                        if (lane < rep) W->lens[32 + idx + lane] = (uint8_t)len;
                        if (lane + 32 < rep) W->lens[32 + idx + lane + 32] = (uint8_t)len;
                        if (lane + 64 < rep) W->lens[32 + idx + lane + 64] = (uint8_t)len;
                        if (lane + 96 < rep) W->lens[32 + idx + lane + 96] = (uint8_t)len;
                        if (lane + 128 < rep) W->lens[32 + idx + lane + 128] = (uint8_t)len;
                        idx += rep;
                    }
                    __syncwarp();
                }
                // move the lengths down to lens[0..nl+nd)
                uint8_t t[10];
#pragma unroll
                for (int k = 0; k < 10; ++k) t[k] = lane + 32 * k < nl + nd ? W->lens[32 + lane + 32 * k] : (uint8_t)0;
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 10; ++k) if (lane + 32 * k < nl + nd) W->lens[lane + 32 * k] = t[k];
                __syncwarp();
                if (W->lens[256] == 0) return -10;
            }
            const int brc = infw_build(T, W, nl, nd, lane, type == 1);
            if (brc) return brc;
            // ---- the symbol loop.  The bit buffer of the header code gives way to a bit POSITION `bp` (relative to
            // b.words): a symbol starts by fetching the 32 bits at bp — two ring words and a funnel shift, no buffer
            // to shift and count down — and a literal/length code with its extra bits (<= 20 bits) or a distance
            // code with its extra bits (<= 28 bits) is taken out of that one window.  `thr` is the position from which
            // the ring needs its next line (lines up to two ahead of the reader are kept loaded, the line behind it stays).
            {
                uint32_t bp = b.rw * 32u - (uint32_t)b.cnt;
                uint32_t thr = (b.next_line - 2u) << 10;
                const uint32_t *ring = W->ring;
                const uint32_t lane_sh = 8u + 8u * (uint32_t)(lane < 2 ? lane : 2);
                uint8_t *out_lane = out + lane;
                // One way out of the loop (`status`: 1 = end of block, anything else an error code): early returns from
                // inside it cost several reconvergence instructions per symbol.
                int status = 0;
                do {
                    while (bp >= thr) {                            // (rarely taken: once per 128 bytes of input)
                        const uint32_t w = b.next_line * 32u + (uint32_t)lane;
                        const uint32_t v = w < b.nwords ? __ldg(b.words + w) : 0u;
                        W->ring[w & 127u] = v;
                        if ((w & 127u) == 0u) W->ring[128] = v;
                        ++b.next_line;
                        thr += 1024u;
                        __syncwarp();
                    }
                    const uint32_t i0 = (bp >> 5) & 127u;
                    const uint32_t win = __funnelshift_r(ring[i0], ring[i0 + 1], bp);
                    uint32_t e = W->ltab[win & ((1u << INFW_LBITS) - 1u)];
                    if (e == 0) {
                        int cl = 0;
                        const int sym = canon_decode(win, W->lcount, W->lsym, INF_MAXBITS, &cl);
                        e = sym < 0 ? (1u | (3u << 4)) : lit_entry(T, sym, cl);      // no such code: an invalid entry
                    }
                    const uint32_t nb = e & 15u, kind = e & 0x30u;
                    bp += nb;
                    if (kind == 0) {
                        // one to three literals per look-up: bytes in bits 8.., (count - 1) in bits 6-7
                        const uint32_t c1 = (e >> 6) & 3u;
                        if (n_out + c1 < isize) {
                            if ((uint32_t)lane <= c1) out_lane[n_out] = (uint8_t)(e >> lane_sh);
                            n_out += c1 + 1u;
                        } else {
                            status = -14;
                        }
                    } else if (kind == 0x10u) {
                        const uint32_t xb = e >> 24;
                        const uint32_t len = ((e >> 8) & 0xFFFFu) + ((win >> nb) & ~(0xFFFFFFFFu << xb));
                        bp += xb;
                        const uint32_t j0 = (bp >> 5) & 127u;
                        const uint32_t win2 = __funnelshift_r(ring[j0], ring[j0 + 1], bp);
                        uint32_t d = W->dtab[win2 & ((1u << INFW_DBITS) - 1u)];
                        if (d == 0) {
                            int cl = 0;
                            const int ds = canon_decode(win2, W->dcount, W->dsym, INF_MAXBITS, &cl);
                            d = ds < 0 ? (1u | (15u << 4)) : dist_entry(T, ds, cl);  // no such code: an invalid entry
                        }
                        const uint32_t db = d & 15u, dx = (d >> 4) & 15u;
                        // (dx = 15 marks an invalid distance symbol: its 15 "extra bits" only yield a distance that is checked
                        // like any other and then rejected by the flag)
                        const uint32_t dist = (d >> 8) + ((win2 >> db) & ~(0xFFFFFFFFu << dx));
                        bp += db + dx;
                        if (dx != 15u && dist <= n_out && n_out + len <= isize) {
                            __syncwarp();                             // earlier stores of all lanes are visible
                            const uint8_t *from = out_lane + (n_out - dist);
                            if (dist >= len) {
                                // (most matches in read text are a few bases long: one predicated load / store)
                                if (len <= 32) {
                                    if ((uint32_t)lane < len) out_lane[n_out] = __ldcg(from);
                                } else {
                                    for (uint32_t i = 0; i + lane < len; i += 32) out_lane[n_out + i] = __ldcg(from + i);
                                }
                            } else {
                                for (uint32_t i = lane; i < len; i += 32) out[n_out + i] = __ldcg(out + (n_out - dist) + i % dist);
                            }
                            n_out += len;
                        } else {
                            status = dx == 15u ? -16 : (dist > n_out ? -17 : -18);
                        }
                    } else {
                        status = kind == 0x20u ? 1 : (nb == 1u && e == (1u | (3u << 4)) ? -13 : -15);
                    }
                } while (status == 0);
                if (status != 1) return status;
                // back to the bit buffer for the next block header / the trailer check
                b.rw = bp >> 5;
                b.buf = (uint64_t)(W->ring[b.rw & 127u] >> (bp & 31u));
                b.cnt = 32 - (int)(bp & 31u);
                ++b.rw;
            }
        } else {
            return -20;
        }
    } while (!last);
    if (n_out != isize) return -21;
    // the gzip trailer follows the deflate stream at the next byte boundary (RFC 1952)
    if (wb_bytepos(b) != lead + data_len) return -19;
    // CRC-32 of the output: one slice per lane, then combine
    __syncwarp();
    const uint32_t chunk = (isize + 31u) / 32u;
    const uint32_t lo = min(isize, (uint32_t)lane * chunk), hi = min(isize, lo + chunk);
    uint32_t crc = 0xFFFFFFFFu;
    {
        // bytes up to the first 16-byte boundary, then 16 bytes per load (two loads in flight), then the rest
        uint32_t i = lo;
        while (i < hi && ((reinterpret_cast<uintptr_t>(out) + i) & 15u)) { crc = T->crc[(crc ^ __ldcg(out + i)) & 0xFFu] ^ (crc >> 8); ++i; }
        auto word = [&](uint32_t w) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { crc = T->crc[(crc ^ w) & 0xFFu] ^ (crc >> 8); w >>= 8; }
        };
        for (; i + 32 <= hi; i += 32) {
            const uint4 a = __ldcg(reinterpret_cast<const uint4 *>(out + i)), c = __ldcg(reinterpret_cast<const uint4 *>(out + i + 16));
            word(a.x); word(a.y); word(a.z); word(a.w); word(c.x); word(c.y); word(c.z); word(c.w);
        }
        for (; i < hi; ++i) crc = T->crc[(crc ^ __ldcg(out + i)) & 0xFFu] ^ (crc >> 8);
    }
    crc ^= 0xFFFFFFFFu;
    uint32_t total = __shfl_sync(0xffffffffu, crc, 0);
    if (chunk) {
        const uint32_t q_full = crc_x2nmodp(S->x2n, chunk, 3);
        for (int l = 1; l < 32; ++l) {
            const uint32_t c_l = __shfl_sync(0xffffffffu, crc, l);
            const uint32_t l_lo = min(isize, (uint32_t)l * chunk), l_len = min(isize, l_lo + chunk) - l_lo;
            if (l_len == 0) break;
            const uint32_t q = l_len == chunk ? q_full : crc_x2nmodp(S->x2n, l_len, 3);
            total = crc_multmodp(q, total) ^ c_l;
        }
    }
    if (total != want_crc) return -22;
    return 0;
}

// Occupancy: a member is one long serial chain (a launch of 2048 members takes 4.9 ms, one of 4096 members 5.5 ms), so
// throughput is members in flight per SM: the tables are sized (9-bit literal/length look-up, 20 KB per block) and the
// registers capped (40) for 11 blocks = 44 members per SM; at 6500 members per launch 55.6 GB/s of text (32 per SM: 48.3).
__global__ void __launch_bounds__(INFW_WARPS * 32, 11)
k_inflate_warp(const __grid_constant__ InflateArgs a)
{
    __shared__ InfShared S;
    if (threadIdx.x == 0) {
        const unsigned short lb[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
        const unsigned short le[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
        const unsigned short db[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
        const unsigned short de[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
        const unsigned char od[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        for (int i = 0; i < 29; ++i) { S.T.lbase[i] = lb[i]; S.T.lext[i] = le[i]; }
        for (int i = 0; i < 30; ++i) { S.T.dbase[i] = db[i]; S.T.dext[i] = de[i]; }
        for (int i = 0; i < 19; ++i) S.T.order[i] = od[i];
        uint32_t p = 1u << 30;                     // x^1
        S.x2n[0] = p;
        for (int n = 1; n < 32; ++n) S.x2n[n] = p = crc_multmodp(p, p);
    }
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        uint32_t c = (uint32_t)i;
        for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        S.T.crc[i] = c;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t warps_total = gridDim.x * INFW_WARPS;
    for (uint32_t i = blockIdx.x * INFW_WARPS + warp; i < a.n_members; i += warps_total) {
        const vfb_member m = a.members[i];
        int rc = 0;
        if (m.isize || m.z_len) rc = inflate_member_warp(&S, &S.w[warp], a.z + m.z_off, m.z_len, a.out + m.out_off, m.isize, lane);
        if (rc != 0 && lane == 0) atomicMin(a.first_bad, i);
        __syncwarp();
    }
}

int launch_inflate(const uint8_t *d_z, const vfb_member *d_members, uint32_t n_members, uint8_t *d_out,
                   uint32_t *d_first_bad, cudaStream_t st)
{
    if (n_members == 0) return VFB_OK;
    static int lanes = 0;
    if (!lanes) {
        const char *e = getenv("VFB_INFLATE_LANES");
        lanes = e ? atoi(e) : 1;
        if (lanes < 1 || lanes > 32) lanes = 1;
    }
    InflateArgs a{d_z, d_members, n_members, d_out, d_first_bad, (uint32_t)lanes};
    static int v1 = -1;
    if (v1 < 0) v1 = getenv("VFB_INFLATE_V1") ? 1 : 0;
    if (!v1) {
        static int bps = 0, sms = 0;
        if (!bps) {
            int dev = 0;
            VFB_CUDA(cudaGetDevice(&dev));
            VFB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            VFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_inflate_warp, INFW_WARPS * 32, 0));
            if (bps < 1) bps = 1;
        }
        uint32_t blocks = (n_members + INFW_WARPS - 1) / INFW_WARPS;
        if (blocks > (uint32_t)(sms * bps)) blocks = (uint32_t)(sms * bps);
        k_inflate_warp<<<blocks, INFW_WARPS * 32, 0, st>>>(a);
        ++g_launches;
        VFB_CUDA(cudaGetLastError());
        return VFB_OK;
    }
    const uint32_t warps = (n_members + lanes - 1) / lanes;
    k_inflate_members<<<(warps + INF_THREADS / 32 - 1) / (INF_THREADS / 32), INF_THREADS, 0, st>>>(a);
    ++g_launches;
    VFB_CUDA(cudaGetLastError());
    return VFB_OK;
}

}  // namespace vfb
