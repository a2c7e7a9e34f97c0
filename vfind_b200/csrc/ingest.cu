// Ingest: gzip (multi-member) FASTQ file -> inflate -> pinned chunks -> async H2D -> GPU parse
// -> the hot loop.
//
// Replaces `File::open(fq_path).map(MultiGzDecoder::new)` + `seq_io::fastq::Reader` +
// `parallel_fastq` of /root/reference/src/lib.rs:233-234, :271-308 (flate2 1.1.1 /
// seq_io 0.3.4).  Grammar: gzip only; strict 4-line records — '@' header, one
// sequence line, '+' line, quality of the same length; "\r\n" trimmed; a missing final
// newline and trailing blank lines are tolerated.
//
// Pipeline: a producer thread fills one of three pinned chunk buffers with inflated text,
// counts newlines as it goes and cuts the chunk at the last 4-line boundary (the remainder is
// carried into the next chunk).  The calling thread queues each chunk on the GPU: H2D copy on
// the copy stream, record framing / validation / span extraction on the device
// (kernels_parse.cu), then K1..K4.  The host never looks at a record.
//
// Inflate: a plain gzip stream is serial (one zlib stream, ~0.3 GB/s).  Block-gzip files
// (BGZF: every member carries its compressed size in a "BC" extra field, as written by
// bgzip / htslib and most sequencer pipelines) are inflated member-parallel on `n_threads`
// host threads — the reference's n_threads knob (src/lib.rs:228) keeps its meaning of
// "worker threads".
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "pgunzip.h"
#include "vfb_internal.cuh"

using namespace vfb;

namespace {

size_t count_nl(const uint8_t *p, size_t n)
{
    // eight bytes at a time: x = word ^ 0x0A.. has a zero byte where the text has '\n'; ((x & 0x7F..) + 0x7F..) | x
    // sets the high bit of every NON-zero byte without carries between bytes, so the cleared high bits count
    size_t c = 0, i = 0;
    const uint64_t k = 0x0A0A0A0A0A0A0A0Aull, lo7 = 0x7F7F7F7F7F7F7F7Full;
    for (; i + 8 <= n; i += 8) {
        uint64_t w;
        memcpy(&w, p + i, 8);
        const uint64_t x = w ^ k;
        // per byte: high bit set iff the byte of x is non-zero (no cross-byte carries)
        const uint64_t nz = ((x & lo7) + lo7) | x;
        c += (size_t)__builtin_popcountll(~nz & ~lo7);
    }
    for (; i < n; ++i) c += p[i] == '\n';
    return c;
}

// Serial multi-member gzip stream (flate2 MultiGzDecoder semantics).
struct Inflater {
    FILE *f = nullptr;
    z_stream z;
    bool z_open = false;
    bool eof = false;          // no more compressed input
    bool member_done = true;   // between members
    std::vector<uint8_t> in;
    Inflater() : in(1 << 20) { memset(&z, 0, sizeof z); }
    ~Inflater()
    {
        if (z_open) inflateEnd(&z);
    }
    // Fill out[0..cap) with inflated bytes; returns bytes produced (0 at end) or -1 on error.
    long long read(uint8_t *out, size_t cap, std::string *err)
    {
        size_t produced = 0;
        while (produced < cap) {
            if (z.avail_in == 0 && !eof) {
                size_t got = fread(in.data(), 1, in.size(), f);
                if (got == 0) eof = true;
                z.next_in = in.data();
                z.avail_in = (uInt)got;
            }
            if (member_done) {
                if (z.avail_in == 0 && eof) break;     // clean end between members
                if (z_open) inflateEnd(&z);
                Bytef *ni = z.next_in;
                uInt ai = z.avail_in;
                memset(&z, 0, sizeof z);
                z.next_in = ni;
                z.avail_in = ai;
                if (inflateInit2(&z, 15 + 16) != Z_OK) { *err = "zlib init failed"; return -1; }
                z_open = true;
                member_done = false;
            }
            if (z.avail_in == 0 && eof) { *err = "truncated gzip stream"; return -1; }
            z.next_out = out + produced;
            size_t room = cap - produced;
            z.avail_out = (uInt)(room > (1u << 30) ? (1u << 30) : room);
            const uInt out0 = z.avail_out;
            int rc = inflate(&z, Z_NO_FLUSH);
            produced += out0 - z.avail_out;
            if (rc == Z_STREAM_END) { member_done = true; continue; }
            if (rc != Z_OK && rc != Z_BUF_ERROR) { *err = std::string("invalid gzip data: ") + (z.msg ? z.msg : "?"); return -1; }
            if (rc == Z_BUF_ERROR && z.avail_in == 0 && eof) { *err = "truncated gzip stream"; return -1; }
        }
        return (long long)produced;
    }
};

// One BGZF member read from the file but not yet inflated.
struct Member {
    size_t in_off = 0, in_len = 0;   // whole member inside the compressed staging buffer
    size_t out_off = 0;
    uint32_t isize = 0;
    size_t lines = 0;
};

// pread of a large range split over `threads` threads (page-cache copies into pinned memory run
// at a few GB/s per core).  Returns the bytes read from `off` on (short only at end of file), -1 on error.
ssize_t pread_parallel(int fd, uint8_t *buf, size_t want, off_t off, int threads)
{
    const size_t min_part = (size_t)4 << 20;
    int parts = threads < 1 ? 1 : threads;
    if ((size_t)parts > want / min_part) parts = (int)(want / min_part);
    if (parts <= 1) {
        size_t done = 0;
        while (done < want) {
            const ssize_t got = pread(fd, buf + done, want - done, off + (off_t)done);
            if (got < 0) { if (errno == EINTR) continue; return -1; }
            if (got == 0) break;
            done += (size_t)got;
        }
        return (ssize_t)done;
    }
    const size_t part = ((want / (size_t)parts) + 4095) & ~(size_t)4095;
    std::vector<ssize_t> got((size_t)parts, 0);
    std::vector<std::thread> pool;
    auto run = [&](int i) {
        const size_t lo = (size_t)i * part, hi = std::min(want, lo + part);
        got[(size_t)i] = lo < hi ? pread_parallel(fd, buf + lo, hi - lo, off + (off_t)lo, 1) : 0;
    };
    for (int i = 1; i < parts; ++i) pool.emplace_back(run, i);
    run(0);
    for (auto &t : pool) t.join();
    size_t total = 0;
    for (int i = 0; i < parts; ++i) {
        if (got[(size_t)i] < 0) return -1;
        const size_t lo = (size_t)i * part, hi = std::min(want, lo + part);
        total += (size_t)got[(size_t)i];
        if ((size_t)got[(size_t)i] < hi - lo) break;       // end of file inside this part
    }
    return (ssize_t)total;
}

// Reads the next gzip member header at the current file position.  Returns 1 and the total
// member size if it is a BGZF member, 0 if it is some other gzip member (position restored),
// -1 at a clean end of file.
int peek_bgzf(FILE *f, size_t *member_size)
{
    const long pos = ftell(f);
    uint8_t h[12];
    const size_t got = fread(h, 1, 12, f);
    if (got == 0) return -1;
    int is_bgzf = 0;
    if (got == 12 && h[0] == 0x1f && h[1] == 0x8b && h[2] == 8 && (h[3] & 4)) {
        const uint32_t xlen = h[10] | (h[11] << 8);
        std::vector<uint8_t> x(xlen);
        if (xlen >= 6 && fread(x.data(), 1, xlen, f) == xlen) {
            for (uint32_t p = 0; p + 4 <= xlen;) {
                const uint32_t slen = x[p + 2] | (x[p + 3] << 8);
                if (x[p] == 'B' && x[p + 1] == 'C' && slen == 2 && p + 6 <= xlen) {
                    *member_size = (size_t)(x[p + 4] | (x[p + 5] << 8)) + 1;
                    is_bgzf = 1;
                    break;
                }
                p += 4 + slen;
            }
        }
    }
    fseek(f, pos, SEEK_SET);
    return is_bgzf;
}

// Fills chunk buffers with inflated text and cuts them at record boundaries.
struct ChunkProducer {
    FILE *f = nullptr;
    ParallelGunzip pgz;         // plain gzip streams on `threads` > 1 host threads
    int pgz_state = 0;          // 0 = not decided, 1 = in use, -1 = zlib stream
    uint64_t file_size = 0;
    std::atomic<uint64_t> consumed{0};   // compressed bytes read so far (progress only)
    Inflater serial;
    bool bgzf = false;          // still reading BGZF members
    bool plain = false;         // uncompressed FASTQ text (vfb_run_file_ex with VFB_INPUT_ALLOW_TEXT only)
    int threads = 1;
    std::vector<uint8_t> carry;
    std::vector<uint8_t> zbuf;  // compressed members of the current chunk
    bool at_end = false;
    std::string err;

    ~ChunkProducer()
    {
        if (f) fclose(f);
    }

    bool open(const char *path, int n_threads, std::string *e, bool allow_text = false)
    {
        f = fopen(path, "rb");
        if (!f) { *e = std::string("cannot open ") + path + ": " + strerror(errno); return false; }
        serial.f = f;
        fseek(f, 0, SEEK_END);
        file_size = (uint64_t)ftell(f);
        fseek(f, 0, SEEK_SET);
        threads = n_threads < 1 ? 1 : n_threads;
        size_t ms = 0;
        bgzf = peek_bgzf(f, &ms) == 1;
        if (allow_text && !bgzf) {
            uint8_t m[2] = {0, 0};
            const size_t got = fread(m, 1, 2, f);
            fseek(f, 0, SEEK_SET);
            plain = !(got == 2 && m[0] == 0x1f && m[1] == 0x8b);
        }
        return true;
    }

    // Member-parallel fill of buf[used..cap).  Returns false on error.
    bool fill_bgzf(uint8_t *buf, size_t cap, size_t &used, size_t &lines)
    {
        std::vector<Member> ms;
        zbuf.clear();
        while (used < cap) {
            size_t msize = 0;
            const int k = peek_bgzf(f, &msize);
            if (k < 0) { at_end = true; break; }
            if (k == 0) { bgzf = false; break; }          // the rest is plain gzip: serial from here
            if (msize < 26) { err = "invalid BGZF member"; return false; }
            const size_t zo = zbuf.size();
            zbuf.resize(zo + msize);
            const long pos = ftell(f);
            if (fread(zbuf.data() + zo, 1, msize, f) != msize) { err = "truncated gzip stream"; return false; }
            const uint8_t *t = zbuf.data() + zo + msize - 4;
            const uint32_t isize = t[0] | (t[1] << 8) | (t[2] << 16) | ((uint32_t)t[3] << 24);
            if (isize > cap - used) {
                fseek(f, pos, SEEK_SET);                   // does not fit: first member of the next chunk
                zbuf.resize(zo);
                if (ms.empty()) { err = "ingest chunk too small for a gzip member plus a carried record"; return false; }
                break;
            }
            Member m;
            m.in_off = zo; m.in_len = msize; m.out_off = used; m.isize = isize;
            ms.push_back(m);
            used += isize;
        }
        if (ms.empty()) return true;
        std::atomic<size_t> next{0};
        std::atomic<bool> bad{false};
        auto work = [&]() {
            z_stream z;
            for (;;) {
                const size_t i = next.fetch_add(1);
                if (i >= ms.size() || bad.load()) return;
                Member &m = ms[i];
                memset(&z, 0, sizeof z);
                if (inflateInit2(&z, 15 + 16) != Z_OK) { bad = true; return; }
                z.next_in = zbuf.data() + m.in_off;
                z.avail_in = (uInt)m.in_len;
                z.next_out = buf + m.out_off;
                z.avail_out = m.isize;
                const int rc = inflate(&z, Z_FINISH);
                const bool ok = rc == Z_STREAM_END && z.avail_out == 0 && z.avail_in == 0;
                inflateEnd(&z);
                if (!ok) { bad = true; return; }
                m.lines = count_nl(buf + m.out_off, m.isize);
            }
        };
        const int nt = (int)std::min<size_t>((size_t)threads, ms.size());
        std::vector<std::thread> pool;
        for (int t = 1; t < nt; ++t) pool.emplace_back(work);
        work();
        for (auto &t : pool) t.join();
        if (bad.load()) { err = "invalid gzip data in a BGZF member"; return false; }
        for (const Member &m : ms) lines += m.lines;
        return true;
    }

    // Produce the next chunk into buf (capacity cap, plus 64 spare bytes).  On success sets
    // cut (bytes to submit) and keep_lines (complete lines in [0,cut)); *last when the input is
    // exhausted.  Returns VFB_OK or an error code with `err` set.
    int next(uint8_t *buf, size_t cap, size_t *cut_out, size_t *lines_out, bool *last)
    {
        if (carry.size() >= cap) { err = "a FASTQ record is larger than the ingest chunk"; return VFB_ERR_FORMAT; }
        size_t used = carry.size();
        if (used) memcpy(buf, carry.data(), used);
        size_t lines = count_nl(buf, used);
        while (used < cap && !at_end) {
            if (bgzf) {
                if (!fill_bgzf(buf, cap, used, lines)) return VFB_ERR_FORMAT;
                if (bgzf && !at_end) break;                   // as full as whole members allow
            } else {
                if (!plain && pgz_state == 0) {
                    const char *e = getenv("VFB_PGUNZIP");
                    pgz_state = threads > 1 && !(e && e[0] == '0') ? 1 : -1;
                    if (pgz_state == 1) pgz.init(f, threads);
                }
                if (!plain && pgz_state == 1) {
                    // already decoded in the background: take all the chunk has room for, copied and counted in parallel
                    const size_t want = cap - used;
                    size_t nl = 0;
                    const long long got = pgz.read_counting(buf + used, want, &nl, &err);
                    if (got < 0) return VFB_ERR_FORMAT;
                    if ((size_t)got < want) at_end = true;
                    lines += nl;
                    used += (size_t)got;
                    continue;
                }
                const size_t want = cap - used < ((size_t)4 << 20) ? cap - used : ((size_t)4 << 20);
                const long long got = plain ? (long long)fread(buf + used, 1, want, f) : serial.read(buf + used, want, &err);
                if (got < 0) return VFB_ERR_FORMAT;
                if ((size_t)got < want) at_end = true;
                lines += count_nl(buf + used, (size_t)got);
                used += (size_t)got;
            }
        }
        size_t cut = used, keep = lines;
        if (at_end) {
            // trailing blank lines are tolerated, and so is a missing final newline
            while (used && (buf[used - 1] == '\n' || buf[used - 1] == '\r')) --used;
            if (used) buf[used++] = '\n';
            lines = count_nl(buf, used);
            if (lines % 4) { err = "truncated FASTQ record at the end of the input"; return VFB_ERR_FORMAT; }
            cut = used;
            keep = lines;
            carry.clear();
        } else {
            const size_t rem = lines % 4;
            keep = lines - rem;
            // cut just after newline number `keep`: walk back over the last `rem` newlines
            size_t end = used;
            for (size_t s = 0; s <= rem && keep; ++s) {
                const void *q = end ? memrchr(buf, '\n', end) : nullptr;
                end = q ? (size_t)((const uint8_t *)q - buf) : 0;
            }
            cut = keep ? end + 1 : 0;
            carry.assign(buf + cut, buf + used);
        }
        *cut_out = cut;
        *lines_out = keep;
        *last = at_end;
        consumed = (uint64_t)ftell(f);
        return VFB_OK;
    }
};

size_t pick_chunk(FILE *f)
{
    fseek(f, 0, SEEK_END);
    const long fsz = ftell(f);
    fseek(f, 0, SEEK_SET);
    size_t cap = (size_t)128 << 20;
    if (const char *e = getenv("VFB_INGEST_CHUNK")) cap = (size_t)strtoull(e, nullptr, 10);
    else if (fsz >= 0 && (size_t)fsz * 12 + 65536 < cap) cap = (size_t)fsz * 12 + 65536;
    if (cap < 256) cap = 256;
    return cap;
}

struct Chunk {
    uint8_t *buf = nullptr;
    size_t buf_cap = 0;
    size_t cut = 0;        // bytes to submit (ends at a record boundary)
    size_t lines = 0;      // complete lines in [0, cut)
    cudaEvent_t copied = nullptr;
};

struct Pipe {
    std::mutex mu;
    std::condition_variable cv;
    std::deque<int> free_q, ready_q;
    bool done = false, failed = false, abort = false;
    std::string err;
    int err_code = VFB_OK;
};

// ---- GPU inflate phase: block-gzip members are shipped compressed and inflated on the device.
struct ZSegment {
    uint8_t *z = nullptr;          // pinned: whole members back to back
    vfb_member *members = nullptr; // pinned
    size_t z_cap = 0, m_cap = 0;   // pinned capacities (for the pool)
    size_t z_bytes = 0, text_bytes = 0;
    uint32_t n = 0;
    bool last = false;             // nothing more for the GPU phase after this segment
    double read_ms = 0;            // what the reader thread spent filling it (trace)
};

// Consumes BGZF members from prod.f until the end of the file or the first member that is not
// BGZF.  On return prod.carry holds the text after the last complete record, prod.at_end /
// prod.bgzf say what is left for the host path (which also applies the end-of-input rules).
int run_bgzf_gpu(vfb_ctx *ctx, ChunkProducer &prod, size_t text_target, uint64_t *n_total, bool trace)
{
    // pinned staging for the compressed members of one segment: sized for a 3.2x ratio (a segment
    // closes early when the buffer fills first) and never beyond what is left of the file
    size_t zcap = text_target / 16 * 5 + (1u << 20);
    {
        const long here = ftell(prod.f);
        fseek(prod.f, 0, SEEK_END);
        const long fsz = ftell(prod.f);
        fseek(prod.f, here, SEEK_SET);
        if (fsz > here && (size_t)(fsz - here) + 65536 < zcap) zcap = (size_t)(fsz - here) + 65536;
    }
    const size_t mcap = text_target / 512 + 4096;
    constexpr int NSEG = 2;
    ZSegment seg[NSEG];
    int rc = VFB_OK;
    for (auto &g : seg) {
        void *a = pinned_acquire(zcap, &g.z_cap), *b = pinned_acquire(mcap * sizeof(vfb_member), &g.m_cap);
        if (!a || !b) {
            set_error("cannot allocate pinned ingest buffers");
            rc = VFB_ERR_NOMEM;
        }
        g.z = (uint8_t *)a;
        g.members = (vfb_member *)b;
    }
    vfb::trace("ingest: pinned segment buffers ready (2 x %zu MB)", zcap >> 20);
    Pipe pp;
    for (int i = 0; i < NSEG; ++i) pp.free_q.push_back(i);
    const int fd = fileno(prod.f);
    size_t pos = (size_t)ftell(prod.f);
    size_t z_estimate = std::min(zcap, text_target / 4 + (1u << 20));     // compressed bytes a segment is expected to need
    std::thread reader;
    if (rc == VFB_OK) reader = std::thread([&]() {
        for (;;) {
            int k;
            {
                std::unique_lock<std::mutex> lk(pp.mu);
                pp.cv.wait(lk, [&] { return !pp.free_q.empty() || pp.abort; });
                if (pp.abort) return;
                k = pp.free_q.front();
                pp.free_q.pop_front();
            }
            ZSegment &g = seg[k];
            g.z_bytes = g.text_bytes = 0; g.n = 0; g.last = false;
            const auto r0 = std::chrono::steady_clock::now();
            std::string err;
            // pread straight into the pinned buffer — about as much as the last segment needed,
            // topped up while whole members are still missing — then walk the member headers
            size_t avail = 0, off = 0;
            bool eof = false, stop = false, full = false;
            size_t step = z_estimate;
            while (err.empty() && !stop && !full) {
                if (!eof && avail < zcap) {
                    const size_t want = std::min(zcap - avail, step);
                    const ssize_t got = pread_parallel(fd, g.z + avail, want, (off_t)(pos + avail), prod.threads);
                    if (got < 0) { err = std::string("read error: ") + strerror(errno); break; }
                    if ((size_t)got < want) eof = true;
                    avail += (size_t)got;
                    step = std::max<size_t>((size_t)4 << 20, z_estimate / 4);
                }
                bool need_more = false;
                while (off < avail && g.n < mcap) {
                    const uint8_t *h = g.z + off;
                    size_t msize = 0;
                    bool is_bgzf = false, cut = false;
                    if (avail - off >= 18 && h[0] == 0x1f && h[1] == 0x8b && h[2] == 8 && (h[3] & 4)) {
                        const uint32_t xlen = h[10] | (h[11] << 8);
                        if (avail - off >= 12 + (size_t)xlen) {
                            for (uint32_t x = 0; x + 4 <= xlen;) {
                                const uint8_t *sf = h + 12 + x;
                                const uint32_t slen = sf[2] | (sf[3] << 8);
                                if (sf[0] == 'B' && sf[1] == 'C' && slen == 2 && x + 6 <= xlen) { msize = (size_t)(sf[4] | (sf[5] << 8)) + 1; is_bgzf = true; break; }
                                x += 4 + slen;
                            }
                        } else cut = true;                          // header cut by what has been read so far
                    } else if (avail - off < 18) cut = true;
                    if (cut && !eof) { need_more = true; break; }
                    if (!is_bgzf) {
                        if (cut) { err = "truncated gzip stream"; break; }     // end of file inside a header
                        prod.bgzf = false; g.last = true; stop = true;       // plain gzip from here: host path
                        break;
                    }
                    if (msize < 26) { err = "invalid BGZF member"; break; }
                    if (off + msize > avail) {
                        if (eof) err = "truncated gzip stream";              // end of file inside a member
                        else need_more = true;
                        break;
                    }
                    const uint8_t *t = h + msize - 4;
                    const uint32_t isize = t[0] | (t[1] << 8) | (t[2] << 16) | ((uint32_t)t[3] << 24);
                    if (g.n && g.text_bytes + isize > text_target) { full = true; break; }
                    g.members[g.n++] = vfb_member{(uint32_t)off, (uint32_t)msize, (uint32_t)g.text_bytes, isize};
                    off += msize;
                    g.text_bytes += isize;
                }
                if (!err.empty() || stop || full) break;
                if (g.n >= mcap) break;
                if (avail >= zcap) {
                    if (need_more && off == 0) err = "a gzip member does not fit the ingest buffer";
                    break;                                           // the buffer is full: next segment
                }
                if (eof) break;                                      // everything read and consumed
            }
            if (off) z_estimate = off + off / 16 + 65536;
            vfb::trace("ingest reader: segment of %u members, %zu compressed bytes read", g.n, off);
            g.z_bytes = off;
            g.read_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - r0).count();
            pos += off;
            if (err.empty() && !stop && eof && off == avail) { prod.at_end = true; g.last = true; }
            std::lock_guard<std::mutex> lk(pp.mu);
            if (!err.empty()) {
                pp.failed = true; pp.err = err; pp.err_code = VFB_ERR_FORMAT; pp.done = true;
                pp.cv.notify_all();
                return;
            }
            pp.ready_q.push_back(k);
            if (g.last) pp.done = true;
            pp.cv.notify_all();
            if (g.last) return;
        }
    });
    std::vector<uint8_t> tail(VFB_TAIL_CAP);
    uint64_t seg_bytes_done = pos;
    while (rc == VFB_OK) {
        int k = -1;
        const auto w0 = std::chrono::steady_clock::now();
        {
            std::unique_lock<std::mutex> lk(pp.mu);
            pp.cv.wait(lk, [&] { return !pp.ready_q.empty() || pp.done; });
            if (!pp.ready_q.empty()) { k = pp.ready_q.front(); pp.ready_q.pop_front(); }
            else if (pp.failed) { set_error(pp.err); rc = pp.err_code; break; }
            else break;
        }
        ZSegment &g = seg[k];
        const auto t0 = std::chrono::steady_clock::now();
        uint64_t n_rec = 0, tail_len = 0;
        uint32_t bad = 0xFFFFFFFFu;
        rc = vfb_internal_submit_bgzf(ctx, g.z, g.z_bytes, g.members, g.n, g.text_bytes, prod.carry.data(),
                                      prod.carry.size(), *n_total, &n_rec, tail.data(), &tail_len, &bad);
        if (rc == VFB_OK && bad != 0xFFFFFFFFu) {
            set_error("invalid gzip data in a BGZF member (deflate stream, CRC-32 or size mismatch)");
            rc = VFB_ERR_FORMAT;
        }
        if (rc == VFB_OK) {
            prod.carry.assign(tail.data(), tail.data() + tail_len);
            *n_total += n_rec;
            seg_bytes_done += g.z_bytes;
            vfb_internal_progress(ctx, *n_total, seg_bytes_done, prod.file_size, false);
            if (trace) fprintf(stderr, "[vfb ingest] gpu segment: %u members, %zu -> %zu bytes, %llu records, %.1f ms (waited %.1f ms for the reader, which took %.1f ms)\n", g.n,
                               g.z_bytes, g.text_bytes, (unsigned long long)n_rec,
                               std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(),
                               std::chrono::duration<double, std::milli>(t0 - w0).count(), g.read_ms);
        }
        std::lock_guard<std::mutex> lk(pp.mu);
        pp.free_q.push_back(k);
        pp.cv.notify_all();
    }
    {
        std::lock_guard<std::mutex> lk(pp.mu);
        pp.abort = true;
        pp.cv.notify_all();
    }
    if (reader.joinable()) reader.join();
    fseek(prod.f, (long)pos, SEEK_SET);      // the host path continues where the GPU phase stopped
    if (rc == VFB_OK && pp.failed) { set_error(pp.err); rc = pp.err_code; }
    const std::string keep = rc ? std::string(vfb_last_error()) : std::string();
    vfb_sync(ctx);                 // the pinned buffers are about to go away
    if (rc) set_error(keep);
    for (auto &g : seg) {
        pinned_release(g.z, g.z_cap);
        pinned_release(g.members, g.m_cap);
    }
    return rc;
}

}  // namespace

// Host-only view of the ingest front half (no GPU): inflates `path` chunk by chunk exactly as
// vfb_run_file does and returns the concatenated submitted text, the number of complete lines
// and the number of chunks.  For CPU tests of the inflate / cut / carry logic.
extern "C" int vfb_debug_inflate_file(const char *path, uint32_t n_threads, uint8_t *out, uint64_t out_cap,
                                      uint64_t *n_bytes, uint64_t *n_lines, uint64_t *n_chunks)
{
    if (!path || !n_bytes || !n_lines) { set_error("null argument"); return VFB_ERR_ARG; }
    ChunkProducer prod;
    std::string e;
    if (!prod.open(path, (int)n_threads, &e)) { set_error(e); return VFB_ERR_IO; }
    const size_t cap = pick_chunk(prod.f);
    std::vector<uint8_t> buf(cap + 64);
    uint64_t total = 0, lines = 0, chunks = 0;
    for (;;) {
        size_t cut = 0, kl = 0;
        bool last = false;
        const int rc = prod.next(buf.data(), cap, &cut, &kl, &last);
        if (rc) { set_error(prod.err); return rc; }
        if (out) {
            if (total + cut > out_cap) { set_error("output buffer too small"); return VFB_ERR_ARG; }
            memcpy(out + total, buf.data(), cut);
        }
        total += cut;
        lines += kl;
        ++chunks;
        if (last) break;
    }
    *n_bytes = total;
    *n_lines = lines;
    if (n_chunks) *n_chunks = chunks;
    return VFB_OK;
}

extern "C" int vfb_run_file(vfb_ctx *ctx, const char *path, uint64_t *n_reads_out)
{
    return vfb_run_file_ex(ctx, path, 0, n_reads_out);
}

extern "C" int vfb_run_file_ex(vfb_ctx *ctx, const char *path, uint32_t flags, uint64_t *n_reads_out)
{
    if (!ctx || !path) { set_error("null argument"); return VFB_ERR_ARG; }
    ChunkProducer prod;
    {
        std::string e;
        if (!prod.open(path, vfb_internal_ingest_threads(ctx), &e, (flags & VFB_INPUT_ALLOW_TEXT) != 0)) { set_error(e); return VFB_ERR_IO; }
    }
    size_t cap = pick_chunk(prod.f);
    const bool trace = getenv("VFB_INGEST_TRACE") != nullptr;
    const auto t_start = std::chrono::steady_clock::now();
    uint64_t n_total = 0;
    int rc = VFB_OK;
    // Block-gzip input is inflated on the GPU (VFB_GPU_INFLATE=0 keeps it on the host threads);
    // whatever follows — plain gzip members, the end-of-input rules — goes through the host path.
    const char *gi = getenv("VFB_GPU_INFLATE");
    if (prod.bgzf && !(gi && gi[0] == '0')) {
        if (trace) fprintf(stderr, "[vfb ingest] %s: block gzip, inflating on the GPU\n", path);
        rc = run_bgzf_gpu(ctx, prod, getenv("VFB_INGEST_CHUNK") ? cap : ((size_t)256 << 20), &n_total, trace);
        if (rc) return rc;
        // only the text after the last complete record may be left: no need for big chunks
        if (prod.at_end) cap = prod.carry.size() * 2 + 65536;
    }

    constexpr int NCH = 3;
    Chunk ch[NCH];
    Pipe pp;
    for (int i = 0; i < NCH; ++i) {
        void *p = pinned_acquire(cap + 64, &ch[i].buf_cap);
        if (!p || cudaEventCreateWithFlags(&ch[i].copied, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            set_error("cannot allocate pinned ingest buffers");
            rc = VFB_ERR_NOMEM;
        }
        ch[i].buf = (uint8_t *)p;
        pp.free_q.push_back(i);
    }

    auto ms_since = [&](std::chrono::steady_clock::time_point t) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count();
    };
    if (trace) fprintf(stderr, "[vfb ingest] %s: %s, chunk %zu bytes, %d inflate threads\n", path,
                       prod.plain ? "plain text" : (prod.bgzf ? "block gzip (member-parallel)" : "gzip stream (serial)"), cap, prod.threads);
    std::thread producer;
    if (rc == VFB_OK) producer = std::thread([&]() {
        for (;;) {
            int k;
            {
                std::unique_lock<std::mutex> lk(pp.mu);
                pp.cv.wait(lk, [&] { return !pp.free_q.empty() || pp.abort; });
                if (pp.abort) return;
                k = pp.free_q.front();
                pp.free_q.pop_front();
            }
            bool last = false;
            const auto t0 = std::chrono::steady_clock::now();
            const int prc = prod.next(ch[k].buf, cap, &ch[k].cut, &ch[k].lines, &last);
            if (trace) fprintf(stderr, "[vfb ingest] chunk %zu bytes %zu lines inflated in %.1f ms\n", ch[k].cut, ch[k].lines,
                               std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
            std::lock_guard<std::mutex> lk(pp.mu);
            if (prc) {
                pp.failed = true; pp.err = prod.err; pp.err_code = prc; pp.done = true;
                pp.cv.notify_all();
                return;
            }
            pp.ready_q.push_back(k);
            if (last) pp.done = true;
            pp.cv.notify_all();
            if (last) return;
        }
    });

    std::deque<int> in_flight;
    while (rc == VFB_OK) {
        int k = -1;
        {
            std::unique_lock<std::mutex> lk(pp.mu);
            pp.cv.wait(lk, [&] { return !pp.ready_q.empty() || pp.done; });
            if (!pp.ready_q.empty()) { k = pp.ready_q.front(); pp.ready_q.pop_front(); }
            else if (pp.failed) { set_error(pp.err); rc = pp.err_code; break; }
            else break;     // done and drained
        }
        Chunk &c = ch[k];
        if (c.lines) {
            const auto t0 = std::chrono::steady_clock::now();
            rc = vfb_internal_submit_fastq(ctx, c.buf, c.cut, c.lines, n_total, c.copied);
            if (trace) fprintf(stderr, "[vfb ingest] submit at %.1f ms took %.1f ms\n", ms_since(t_start), ms_since(t0));
            n_total += c.lines / 4;
            vfb_internal_progress(ctx, n_total, prod.consumed.load(), prod.file_size, false);
        } else if (cudaEventRecord(c.copied, 0) != cudaSuccess) {
            cudaGetLastError();
        }
        in_flight.push_back(k);
        // hand buffers back to the producer once their H2D copy has finished; the newest copy
        // stays outstanding so that inflating the next chunk overlaps it
        while (rc == VFB_OK && in_flight.size() > 1) {
            const int o = in_flight.front();
            if (cudaEventSynchronize(ch[o].copied) != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), "ingest event", __FILE__, __LINE__); break; }
            in_flight.pop_front();
            std::lock_guard<std::mutex> lk(pp.mu);
            pp.free_q.push_back(o);
            pp.cv.notify_all();
        }
    }
    {
        std::lock_guard<std::mutex> lk(pp.mu);
        pp.abort = true;
        pp.cv.notify_all();
    }
    if (producer.joinable()) producer.join();
    if (rc == VFB_OK && pp.failed) { set_error(pp.err); rc = pp.err_code; }
    // the malformed-record flag comes back from the device
    const std::string keep = rc ? std::string(vfb_last_error()) : std::string();
    if (trace) fprintf(stderr, "[vfb ingest] all chunks queued at %.1f ms\n", ms_since(t_start));
    const int src = vfb_sync(ctx);
    if (trace) fprintf(stderr, "[vfb ingest] device drained at %.1f ms\n", ms_since(t_start));
    if (rc == VFB_OK) rc = src; else set_error(keep);
    if (rc == VFB_OK) {
        uint64_t bad = UINT64_MAX;
        rc = vfb_internal_parse_error(ctx, &bad);
        if (rc == VFB_OK && bad != UINT64_MAX) {
            set_error("FASTQ record " + std::to_string(bad) +
                      ": malformed (expected '@' header, '+' separator, sequence and quality of equal length)");
            rc = VFB_ERR_FORMAT;
        }
    }
    for (int i = 0; i < NCH; ++i) {
        pinned_release(ch[i].buf, ch[i].buf_cap);
        if (ch[i].copied) cudaEventDestroy(ch[i].copied);
    }
    if (n_reads_out) *n_reads_out = n_total;
    if (rc == VFB_OK) vfb_internal_progress(ctx, n_total, prod.file_size, prod.file_size, true);
    if (trace) fprintf(stderr, "[vfb ingest] %llu records in %.1f ms\n", (unsigned long long)n_total,
                       std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count());
    return rc;
}
