// Ingest: gzip (multi-member) FASTQ file -> inflate -> record parse -> vfb_submit_host.
//
// Replaces `File::open(fq_path).map(MultiGzDecoder::new)` + `seq_io::fastq::Reader` +
// `parallel_fastq` of /root/reference/src/lib.rs:233-234, :271-308 (flate2 1.1.1 /
// seq_io 0.3.4).  Grammar: gzip only; strict 4-line records — '@' header, one
// sequence line, '+' line, quality of the same length; "\r\n" trimmed; a missing final
// newline and trailing blank lines are tolerated.  Only the sequence line reaches the GPU:
// the inflated text is uploaded as is and reads are (offset, length) spans into it.
#include <zlib.h>

#include <cstdio>
#include <cstring>
#include <string>
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
    // chunk size: large enough to amortise launches, small enough for toy files
    fseek(inf.f, 0, SEEK_END);
    long fsz = ftell(inf.f);
    fseek(inf.f, 0, SEEK_SET);
    size_t chunk = (size_t)1 << 28;
    if (fsz >= 0 && (size_t)fsz * 8 + 4096 < chunk) chunk = (size_t)fsz * 8 + 4096;
    std::vector<uint8_t> buf(chunk);
    std::vector<vfb_span> spans;
    size_t carry = 0;            // bytes of an incomplete record kept at the front of buf
    uint64_t n_total = 0;
    bool at_end = false;
    std::string err;
    while (!at_end) {
        long long got = inf.read(buf.data() + carry, buf.size() - carry, &err);
        if (got < 0) { set_error(err); return VFB_ERR_FORMAT; }
        size_t have = carry + (size_t)got;
        if ((size_t)got < buf.size() - carry) at_end = true;
        // parse complete records
        spans.clear();
        size_t p = 0;
        while (p < have) {
            size_t ls[4], ll[4];
            size_t q = p;
            bool complete = true;
            for (int l = 0; l < 4; ++l) {
                if (q >= have && !(at_end && l == 3 && q == have)) { complete = false; break; }
                const uint8_t *nl = q < have ? (const uint8_t *)memchr(buf.data() + q, '\n', have - q) : nullptr;
                size_t e;
                if (nl) e = (size_t)(nl - buf.data());
                else if (at_end && l == 3) e = have;      // final newline missing
                else { complete = false; break; }
                size_t ee = e;
                if (ee > q && buf[ee - 1] == '\r') --ee;
                ls[l] = q; ll[l] = ee - q;
                q = nl ? e + 1 : have;
            }
            if (!complete) {
                if (at_end) {
                    // only blank lines may remain
                    bool blank = true;
                    for (size_t k = p; k < have; ++k) if (buf[k] != '\n' && buf[k] != '\r') { blank = false; break; }
                    if (!blank) { set_error("truncated FASTQ record " + std::to_string(n_total + spans.size())); return VFB_ERR_FORMAT; }
                    p = have;
                }
                break;
            }
            if (at_end) {
                // a tail made only of newlines is not a record
                bool blank = true;
                for (size_t k = p; k < have; ++k) if (buf[k] != '\n' && buf[k] != '\r') { blank = false; break; }
                if (blank) { p = have; break; }
            }
            const uint64_t rec = n_total + spans.size();
            if (ll[0] == 0 || buf[ls[0]] != '@') { set_error("FASTQ record " + std::to_string(rec) + ": expected '@'"); return VFB_ERR_FORMAT; }
            if (ll[2] == 0 || buf[ls[2]] != '+') { set_error("FASTQ record " + std::to_string(rec) + ": expected '+'"); return VFB_ERR_FORMAT; }
            if (ll[1] != ll[3]) { set_error("FASTQ record " + std::to_string(rec) + ": sequence and quality lengths differ"); return VFB_ERR_FORMAT; }
            spans.push_back(vfb_span{(uint32_t)ls[1], (uint32_t)ll[1]});
            p = q;
        }
        if (!spans.empty()) {
            int rc = vfb_submit_host(ctx, buf.data(), p, spans.data(), spans.size());
            if (rc) return rc;
            // the submit staged the bytes into pinned memory; buf may be reused
            n_total += spans.size();
        }
        carry = have - p;
        if (carry == buf.size()) {
            // one record larger than the chunk: grow
            buf.resize(buf.size() * 2);
        }
        if (carry && p) memmove(buf.data(), buf.data() + p, carry);
        if (at_end && carry) { set_error("truncated FASTQ record " + std::to_string(n_total)); return VFB_ERR_FORMAT; }
    }
    if (n_reads_out) *n_reads_out = n_total;
    return VFB_OK;
}
