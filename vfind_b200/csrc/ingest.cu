// Ingest: gzip (multi-member) FASTQ file -> inflate -> pinned chunks -> async H2D -> GPU parse
// -> the hot loop.
//
// Replaces `File::open(fq_path).map(MultiGzDecoder::new)` + `seq_io::fastq::Reader` +
// `parallel_fastq` of /root/reference/src/lib.rs:233-234, :271-308 (flate2 1.1.1 /
// seq_io 0.3.4).  Grammar: gzip only; strict 4-line records — '@' header, one
// sequence line, '+' line, quality of the same length; "\r\n" trimmed; a missing final
// newline and trailing blank lines are tolerated.
//
// Pipeline: a producer thread inflates straight into one of three pinned chunk buffers,
// counts newlines as it goes and cuts the chunk at the last 4-line boundary (the remainder is
// carried into the next chunk).  The calling thread queues each chunk on the GPU: H2D copy on
// the copy stream, record framing / validation / span extraction on the device
// (kernels_parse.cu), then K1..K4.  The host never looks at a record.
#include <zlib.h>

#include <cerrno>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "vfb_internal.cuh"

using namespace vfb;

namespace {

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
        if (f) fclose(f);
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

size_t count_nl(const uint8_t *p, size_t n)
{
    size_t c = 0;
    const uint8_t *e = p + n;
    while (p < e) {
        const uint8_t *q = (const uint8_t *)memchr(p, '\n', (size_t)(e - p));
        if (!q) break;
        ++c;
        p = q + 1;
    }
    return c;
}

struct Chunk {
    uint8_t *buf = nullptr;
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

}  // namespace

extern "C" int vfb_run_file(vfb_ctx *ctx, const char *path, uint64_t *n_reads_out)
{
    if (!ctx || !path) { set_error("null argument"); return VFB_ERR_ARG; }
    Inflater inf;
    inf.f = fopen(path, "rb");
    if (!inf.f) {
        set_error(std::string("cannot open ") + path + ": " + strerror(errno));
        return VFB_ERR_IO;
    }
    fseek(inf.f, 0, SEEK_END);
    const long fsz = ftell(inf.f);
    fseek(inf.f, 0, SEEK_SET);
    size_t cap = (size_t)128 << 20;
    if (const char *e = getenv("VFB_INGEST_CHUNK")) cap = (size_t)strtoull(e, nullptr, 10);
    else if (fsz >= 0 && (size_t)fsz * 12 + 65536 < cap) cap = (size_t)fsz * 12 + 65536;
    if (cap < 256) cap = 256;

    constexpr int NCH = 3;
    Chunk ch[NCH];
    Pipe pp;
    int rc = VFB_OK;
    for (int i = 0; i < NCH; ++i) {
        void *p = nullptr;
        if (cudaMallocHost(&p, cap + 64) != cudaSuccess || cudaEventCreateWithFlags(&ch[i].copied, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            set_error("cannot allocate pinned ingest buffers");
            rc = VFB_ERR_NOMEM;
        }
        ch[i].buf = (uint8_t *)p;
        pp.free_q.push_back(i);
    }

    std::thread producer;
    if (rc == VFB_OK) producer = std::thread([&]() {
        std::vector<uint8_t> carry;
        bool at_end = false;
        auto fail = [&](int code, const std::string &m) {
            std::lock_guard<std::mutex> lk(pp.mu);
            pp.failed = true; pp.err = m; pp.err_code = code; pp.done = true;
            pp.cv.notify_all();
        };
        while (!at_end) {
            int k;
            {
                std::unique_lock<std::mutex> lk(pp.mu);
                pp.cv.wait(lk, [&] { return !pp.free_q.empty() || pp.abort; });
                if (pp.abort) return;
                k = pp.free_q.front();
                pp.free_q.pop_front();
            }
            Chunk &c = ch[k];
            if (carry.size() >= cap) { fail(VFB_ERR_FORMAT, "a FASTQ record is larger than the ingest chunk"); return; }
            size_t used = carry.size();
            if (used) memcpy(c.buf, carry.data(), used);
            size_t lines = count_nl(c.buf, used);
            std::string err;
            while (used < cap && !at_end) {
                const size_t want = cap - used < ((size_t)4 << 20) ? cap - used : ((size_t)4 << 20);
                const long long got = inf.read(c.buf + used, want, &err);
                if (got < 0) { fail(VFB_ERR_FORMAT, err); return; }
                if ((size_t)got < want) at_end = true;
                lines += count_nl(c.buf + used, (size_t)got);
                used += (size_t)got;
            }
            size_t cut = used, keep_lines = lines;
            if (at_end) {
                // trailing blank lines are tolerated, and so is a missing final newline
                while (used && (c.buf[used - 1] == '\n' || c.buf[used - 1] == '\r')) --used;
                if (used) c.buf[used++] = '\n';            // cap + 64 bytes were allocated
                lines = count_nl(c.buf, used);
                if (lines % 4) { fail(VFB_ERR_FORMAT, "truncated FASTQ record at the end of the input"); return; }
                cut = used;
                keep_lines = lines;
                carry.clear();
            } else {
                const size_t rem = lines % 4;
                keep_lines = lines - rem;
                // cut just after newline number keep_lines: walk back over the last `rem` newlines
                size_t end = used;
                for (size_t s = 0; s <= rem && keep_lines; ++s) {
                    const void *q = end ? memrchr(c.buf, '\n', end) : nullptr;
                    end = q ? (size_t)((const uint8_t *)q - c.buf) : 0;
                }
                cut = keep_lines ? end + 1 : 0;
                carry.assign(c.buf + cut, c.buf + used);
            }
            c.cut = cut;
            c.lines = keep_lines;
            {
                std::lock_guard<std::mutex> lk(pp.mu);
                pp.ready_q.push_back(k);
                if (at_end) pp.done = true;
                pp.cv.notify_all();
            }
        }
    });

    uint64_t n_total = 0;
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
            rc = vfb_internal_submit_fastq(ctx, c.buf, c.cut, c.lines, n_total, c.copied);
            n_total += c.lines / 4;
        } else if (cudaEventRecord(c.copied, 0) != cudaSuccess) {
            cudaGetLastError();
        }
        in_flight.push_back(k);
        // hand buffers back to the producer once their H2D copy has finished; keep at most one
        // copy outstanding beyond the newest so the producer always has a buffer to fill
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
    int src = vfb_sync(ctx);
    if (rc == VFB_OK) rc = src;
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
        if (ch[i].buf) cudaFreeHost(ch[i].buf);
        if (ch[i].copied) cudaEventDestroy(ch[i].copied);
    }
    if (n_reads_out) *n_reads_out = n_total;
    return rc;
}
