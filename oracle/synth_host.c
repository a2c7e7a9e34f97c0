/*
 * Standalone host generator of the synthetic read stream — TEST INFRASTRUCTURE (bench.py's reference arm and
 * ingest legs, tests).  It restates nothing of the reference: it produces INPUTS.  The stream is defined by the
 * counter-based generator header the CUDA library also compiles (vfind_b200/csrc/synth.h, integer only, identical
 * on host and device); compiling it here gives bench.py --impl reference its inputs without mapping
 * libvfind_b200.so into the process (tests/test_oracle.py asserts byte equality with vfb_synth_host).
 *
 * Also writes the stream as FASTQ: plain text, or block-gzip (BGZF members of <= 65280 text bytes, zlib level 1,
 * like bgzip / sequencer pipelines), compressed on `threads` threads.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include "../vfind_b200/csrc/synth.h"

int vfo_synth_adapters(const vfb_synth_cfg *cfg, uint8_t *prefix, uint8_t *suffix)
{
    if (!cfg || cfg->adapter_len == 0 || cfg->adapter_len > 61) return -1;
    vfs_adapter(cfg->seed, 0, cfg->adapter_len, prefix);
    vfs_adapter(cfg->seed, 1, cfg->adapter_len, suffix);
    return 0;
}

typedef struct {
    const vfb_synth_cfg *cfg;
    uint64_t first, lo, hi;
    uint8_t *text;
    uint32_t *off, *len;
    const uint8_t *pre, *suf;
} synth_job;

static void *synth_run(void *arg)
{
    synth_job *j = (synth_job *)arg;
    const uint32_t L = j->cfg->read_len;
    for (uint64_t i = j->lo; i < j->hi; ++i) {
        vfs_read(j->cfg, j->first + i, j->pre, j->suf, j->text + i * L);
        if (j->off) j->off[i] = (uint32_t)(i * L);
        if (j->len) j->len[i] = L;
    }
    return NULL;
}

/* Reads [first, first+n) as fixed-stride records: text[i*L ..), off[i] = i*L, len[i] = L (off / len may be NULL). */
int vfo_synth_reads(const vfb_synth_cfg *cfg, uint64_t first, uint64_t n, uint8_t *text, uint32_t *off, uint32_t *len,
                    int threads)
{
    if (!cfg || cfg->adapter_len == 0 || cfg->adapter_len > 61 || cfg->read_len == 0 || cfg->read_len > 700) return -1;
    if (n * (uint64_t)cfg->read_len > 0xFFFFFFFFull && off) return -1;
    uint8_t pre[64], suf[64];
    vfs_adapter(cfg->seed, 0, cfg->adapter_len, pre);
    vfs_adapter(cfg->seed, 1, cfg->adapter_len, suf);
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256];
    synth_job jobs[256];
    const uint64_t per = (n + (uint64_t)threads - 1) / (uint64_t)threads;
    int started = 0;
    for (int t = 0; t < threads; ++t) {
        synth_job *j = &jobs[t];
        j->cfg = cfg; j->first = first; j->lo = (uint64_t)t * per; j->hi = j->lo + per < n ? j->lo + per : n;
        j->text = text; j->off = off; j->len = len; j->pre = pre; j->suf = suf;
        if (j->lo >= j->hi) break;
        if (pthread_create(&th[t], NULL, synth_run, j) != 0) { synth_run(j); th[t] = 0; }
        ++started;
    }
    for (int t = 0; t < started; ++t) if (th[t]) pthread_join(th[t], NULL);
    return 0;
}

/* ---- FASTQ writers ---- */
#define REC_HDR 13            /* "@r0000000000\n" */
static size_t rec_bytes(uint32_t L) { return REC_HDR + (size_t)L + 3 + (size_t)L + 1; }

static void format_records(const vfb_synth_cfg *cfg, uint64_t first, uint64_t n, const uint8_t *pre, const uint8_t *suf,
                           uint8_t *out)
{
    const uint32_t L = cfg->read_len;
    const size_t rb = rec_bytes(L);
    for (uint64_t i = 0; i < n; ++i) {
        uint8_t *r = out + i * rb;
        uint64_t idx = first + i;
        r[0] = '@'; r[1] = 'r';
        for (int d = 11; d >= 2; --d) { r[d] = (uint8_t)('0' + idx % 10); idx /= 10; }
        r[12] = '\n';
        vfs_read(cfg, first + i, pre, suf, r + REC_HDR);
        r[REC_HDR + L] = '\n'; r[REC_HDR + L + 1] = '+'; r[REC_HDR + L + 2] = '\n';
        memset(r + REC_HDR + L + 3, 'F', L);
        r[rb - 1] = '\n';
    }
}

/* one BGZF member from `n` text bytes; returns its size (out must hold 65536 + 64) */
static size_t bgzf_member(const uint8_t *text, size_t n, int level, uint8_t *out)
{
    static const uint8_t head[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
    memcpy(out, head, 16);
    z_stream z;
    memset(&z, 0, sizeof z);
    deflateInit2(&z, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
    z.next_in = (Bytef *)text; z.avail_in = (uInt)n;
    z.next_out = out + 18; z.avail_out = 65536 + 64 - 18 - 8;
    deflate(&z, Z_FINISH);
    const size_t body = z.total_out;
    deflateEnd(&z);
    const size_t total = 18 + body + 8;
    out[16] = (uint8_t)((total - 1) & 0xff); out[17] = (uint8_t)((total - 1) >> 8);
    const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), text, (uInt)n), isz = (uint32_t)n;
    uint8_t *t = out + 18 + body;
    t[0] = crc & 0xff; t[1] = (crc >> 8) & 0xff; t[2] = (crc >> 16) & 0xff; t[3] = crc >> 24;
    t[4] = isz & 0xff; t[5] = (isz >> 8) & 0xff; t[6] = (isz >> 16) & 0xff; t[7] = isz >> 24;
    return total;
}

typedef struct {
    const vfb_synth_cfg *cfg;
    const uint8_t *pre, *suf;
    uint64_t first;               /* index of the stream's first record */
    size_t text_from, text_to;    /* the slab's members cover text bytes [text_from, text_to) of the whole stream */
    uint8_t *buf, *zout;
    size_t zlen;
    int level;
} bgzf_job;

static void *bgzf_run(void *arg)
{
    bgzf_job *j = (bgzf_job *)arg;
    const size_t rb = rec_bytes(j->cfg->read_len);
    /* format the records that overlap [text_from, text_to) */
    const uint64_t r_lo = j->text_from / rb, r_hi = (j->text_to + rb - 1) / rb;
    format_records(j->cfg, j->first + r_lo, r_hi - r_lo, j->pre, j->suf, j->buf);
    const uint8_t *text = j->buf + (j->text_from - r_lo * rb);
    size_t left = j->text_to - j->text_from, zo = 0;
    while (left) {
        const size_t take = left < 65280 ? left : 65280;
        zo += bgzf_member(text, take, j->level, j->zout + zo);
        text += take; left -= take;
    }
    j->zlen = zo;
    return NULL;
}

/* Reads [first, first+n) as a FASTQ file: '@r<10-digit index>' headers, quality 'F'.  bgzf = 0: plain text;
 * bgzf = 1: BGZF members of 65280 text bytes (the last one shorter) + the empty end-of-file member.
 * Returns the text bytes written (before compression), 0 on failure; *file_bytes = size of the file. */
uint64_t vfo_write_fastq(const vfb_synth_cfg *cfg, uint64_t first, uint64_t n, const char *path, int bgzf, int level,
                         int threads, int append, uint64_t *file_bytes)
{
    if (!cfg || !path || cfg->adapter_len == 0 || cfg->adapter_len > 61 || cfg->read_len == 0 || cfg->read_len > 700) return 0;
    uint8_t pre[64], suf[64];
    vfs_adapter(cfg->seed, 0, cfg->adapter_len, pre);
    vfs_adapter(cfg->seed, 1, cfg->adapter_len, suf);
    FILE *f = fopen(path, append ? "ab" : "wb");
    if (!f) return 0;
    const size_t rb = rec_bytes(cfg->read_len);
    const size_t total = (size_t)n * rb;
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    uint64_t written = 0;
    int ok = 1;
    if (!bgzf) {
        const uint64_t slab = 1 << 16;
        uint8_t *buf = (uint8_t *)malloc(slab * rb);
        for (uint64_t r = 0; r < n && ok; r += slab) {
            const uint64_t m = n - r < slab ? n - r : slab;
            format_records(cfg, first + r, m, pre, suf, buf);
            ok = fwrite(buf, 1, m * rb, f) == m * rb;
            written += m * rb;
        }
        free(buf);
    } else {
        /* slabs of 256 members each, `threads` slabs per round, written in order */
        const size_t slab_text = (size_t)65280 * 256;
        bgzf_job *jobs = (bgzf_job *)calloc((size_t)threads, sizeof *jobs);
        pthread_t th[256];
        for (int t = 0; t < threads; ++t) {
            jobs[t].buf = (uint8_t *)malloc(slab_text + 2 * rb + 64);
            jobs[t].zout = (uint8_t *)malloc((size_t)256 * (65536 + 64));
            if (!jobs[t].buf || !jobs[t].zout) ok = 0;
        }
        size_t pos = 0;
        while (pos < total && ok) {
            int k = 0;
            for (; k < threads && pos < total; ++k) {
                bgzf_job *j = &jobs[k];
                j->cfg = cfg; j->pre = pre; j->suf = suf; j->first = first; j->level = level;
                j->text_from = pos; j->text_to = pos + slab_text < total ? pos + slab_text : total;
                pos = j->text_to;
                if (pthread_create(&th[k], NULL, bgzf_run, j) != 0) { bgzf_run(j); th[k] = 0; }
            }
            for (int t = 0; t < k; ++t) {
                if (th[t]) pthread_join(th[t], NULL);
                if (ok) ok = fwrite(jobs[t].zout, 1, jobs[t].zlen, f) == jobs[t].zlen;
            }
        }
        /* (with append the caller writes the end-of-file member last: n = 0 does only that) */
        if (ok) {
            uint8_t eofm[64];
            const size_t e = bgzf_member((const uint8_t *)"", 0, level, eofm);
            if (n == 0 || !append) ok = fwrite(eofm, 1, e, f) == e;
        }
        for (int t = 0; t < threads; ++t) { free(jobs[t].buf); free(jobs[t].zout); }
        free(jobs);
        written = total;
    }
    if (fflush(f) != 0) ok = 0;
    if (file_bytes) *file_bytes = (uint64_t)ftell(f);
    fclose(f);
    return ok ? written : 0;
}
