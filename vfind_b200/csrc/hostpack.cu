// Packed host->device path of vfb_submit_host.
//
// The reference hands reads to its workers in host memory (seq_io record sets,
// /root/reference/src/lib.rs:271-277); here they have to cross PCIe first, and at 258 bytes per read
// the link (~55 GB/s) is an order of magnitude slower than the kernels.  Read text is almost
// entirely upper-case A/C/G/T, i.e. 2 bits of information per byte, so the host threads the
// context owns turn 4 MiB blocks of the caller's text into 2-bit codes (hostpack_cpu.cpp; groups
// holding anything else stay verbatim) while the copy engine moves other blocks as they are:
//
//   * workers take blocks from the BACK of the batch, pack them into pinned staging buffers and
//     queue them; the calling thread sends each packed block (a quarter of the bytes) and
//     launches k_unpack, which rebuilds the exact bytes in their place in the device text;
//   * the calling thread meanwhile keeps a few raw block copies in flight from the FRONT.
// The two meet wherever the host's packing rate and the link's rate put them.  The device text is
// byte-identical to the caller's; nothing downstream knows.  Opt-in (VFB_HOST_PACK=<threads>|auto):
// it helps where host DRAM bandwidth is well above the link's; on the boxes this was measured on both
// the copy engine and the packing threads are fed by the same ~48 GB/s of host DRAM reads, and the
// step takes as long with half the bytes on the link (profiles/README.md).
#include "vfb_internal.cuh"

#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

extern "C" size_t vfb_pack_groups(const uint8_t *src, size_t n_groups, uint64_t *codes, uint32_t *rawmap, uint8_t *raw,
                                  size_t raw_cap);

namespace vfb {

#define HP_BLOCK (4u << 20)
#define HP_GROUPS (HP_BLOCK / 32)
#define HP_RAW_CAP (HP_GROUPS / 8)        // verbatim groups a packed block may hold; beyond that it travels raw
#define HP_RING 8                         // device staging slots
#define HP_RAW_INFLIGHT 12

struct HpLayout {
    size_t off_prefix, off_map, off_raw;
};
__host__ __device__ __forceinline__ HpLayout hp_layout(uint32_t groups)
{
    const size_t words = (groups + 31) / 32;
    HpLayout l;
    l.off_prefix = ((size_t)groups * 8 + 15) & ~(size_t)15;
    l.off_map = l.off_prefix + ((words * 4 + 15) & ~(size_t)15);
    l.off_raw = l.off_map + ((words * 4 + 15) & ~(size_t)15);
    return l;
}

// One thread per 32-byte group: 64 bits of codes -> 32 letters, or a verbatim group.
__global__ void __launch_bounds__(256)
k_unpack(const uint8_t *__restrict__ stage, uint32_t groups, uint8_t *__restrict__ dst)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups) return;
    const HpLayout l = hp_layout(groups);
    const uint32_t mw = reinterpret_cast<const uint32_t *>(stage + l.off_map)[g >> 5];
    uint4 *out = reinterpret_cast<uint4 *>(dst + (size_t)g * 32);
    if (mw >> (g & 31) & 1u) {
        const uint32_t idx = reinterpret_cast<const uint32_t *>(stage + l.off_prefix)[g >> 5] + __popc(mw & ((1u << (g & 31)) - 1u));
        const uint4 *src = reinterpret_cast<const uint4 *>(stage + l.off_raw + (size_t)idx * 32);
        out[0] = src[0];
        out[1] = src[1];
        return;
    }
    const unsigned long long code = reinterpret_cast<const unsigned long long *>(stage)[g];
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t x = (uint32_t)(code >> (8 * j)) & 0xFFu;          // four 2-bit codes
        uint32_t s = (x | (x << 4)) & 0x0F0Fu;
        s = (s | (s << 2)) & 0x3333u;                                    // one nibble each: a PRMT selector
        w[j] = __byte_perm(0x47544341u, 0u, s);                          // 0 A, 1 C, 2 T, 3 G
    }
    out[0] = make_uint4(w[0], w[1], w[2], w[3]);
    out[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

struct HpItem {
    uint32_t block;
    int hbuf;
    uint32_t groups;
    size_t bytes;
};

struct HostPacker {
    int n_threads = 0;
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv_job, cv_buf;
    bool quit = false;
    uint64_t gen = 0;
    // current job
    const uint8_t *text = nullptr;
    uint64_t bytes = 0;
    uint32_t front = 0, back = 0;
    std::deque<HpItem> doneq;
    std::deque<uint32_t> rawq;
    // staging
    size_t stage_cap = 0;
    int n_hbufs = 0;
    uint8_t *h_stage = nullptr;
    size_t h_cap = 0;
    std::vector<int> free_bufs;
    uint8_t *d_ring = nullptr;
    cudaEvent_t ev_h2d[HP_RING] = {}, ev_free[HP_RING] = {}, ev_raw[HP_RAW_INFLIGHT] = {}, ev_join = nullptr;
    bool ring_used[HP_RING] = {};
    std::vector<int> pending[HP_RING];      // host buffers whose copy was queued behind ev_h2d[r]
    uint64_t ring_next = 0;
    cudaStream_t st_unpack = nullptr;

    void worker()
    {
        std::unique_lock<std::mutex> lk(mu);
        uint64_t my_gen = 0;
        for (;;) {
            cv_job.wait(lk, [&] { return quit || gen != my_gen; });
            if (quit) return;
            my_gen = gen;
            for (;;) {
                cv_buf.wait(lk, [&] { return quit || !free_bufs.empty() || front >= back; });
                if (quit) return;
                if (front >= back) break;
                const int hb = free_bufs.back();
                free_bufs.pop_back();
                const uint32_t b = --back;
                const uint8_t *src = text + (uint64_t)b * HP_BLOCK;
                const uint64_t len = std::min<uint64_t>(HP_BLOCK, bytes - (uint64_t)b * HP_BLOCK);
                lk.unlock();
                const uint32_t groups = (uint32_t)(len / 32);
                uint8_t *st = h_stage + (size_t)hb * stage_cap;
                const HpLayout l = hp_layout(groups);
                uint32_t *prefix = reinterpret_cast<uint32_t *>(st + l.off_prefix);
                uint32_t *map = reinterpret_cast<uint32_t *>(st + l.off_map);
                const uint32_t words = (groups + 31) / 32;
                memset(map, 0, (size_t)words * 4);
                const size_t n_raw = groups ? vfb_pack_groups(src, groups, reinterpret_cast<uint64_t *>(st), map, st + l.off_raw, HP_RAW_CAP)
                                            : SIZE_MAX;
                if (n_raw != SIZE_MAX) {
                    uint32_t run = 0;
                    for (uint32_t w = 0; w < words; ++w) { prefix[w] = run; run += (uint32_t)__builtin_popcount(map[w]); }
                }
                lk.lock();
                if (n_raw == SIZE_MAX) {
                    free_bufs.push_back(hb);
                    rawq.push_back(b);
                } else {
                    doneq.push_back(HpItem{b, hb, groups, l.off_raw + n_raw * 32});
                }
            }
        }
    }
};

HostPacker *hostpack_create(int threads)
{
    if (threads < 1) return nullptr;
    HostPacker *hp = new HostPacker;
    hp->n_threads = threads;
    hp->stage_cap = (hp_layout(HP_GROUPS).off_raw + (size_t)HP_RAW_CAP * 32 + 255) & ~(size_t)255;
    hp->n_hbufs = threads + 4;
    hp->h_stage = static_cast<uint8_t *>(pinned_acquire(hp->stage_cap * hp->n_hbufs, &hp->h_cap));
    bool ok = hp->h_stage != nullptr;
    ok = ok && cudaMalloc(&hp->d_ring, hp->stage_cap * HP_RING) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&hp->st_unpack, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; ok && i < HP_RING; ++i)
        ok = cudaEventCreateWithFlags(&hp->ev_h2d[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&hp->ev_free[i], cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; ok && i < HP_RAW_INFLIGHT; ++i) ok = cudaEventCreateWithFlags(&hp->ev_raw[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&hp->ev_join, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        cudaGetLastError();
        hostpack_destroy(hp);
        return nullptr;
    }
    for (int i = 0; i < hp->n_hbufs; ++i) hp->free_bufs.push_back(i);
    for (int i = 0; i < threads; ++i) hp->th.emplace_back([hp] { hp->worker(); });
    return hp;
}

void hostpack_destroy(HostPacker *hp)
{
    if (!hp) return;
    {
        std::lock_guard<std::mutex> lk(hp->mu);
        hp->quit = true;
    }
    hp->cv_job.notify_all();
    hp->cv_buf.notify_all();
    for (auto &t : hp->th) t.join();
    if (hp->st_unpack) { cudaStreamSynchronize(hp->st_unpack); cudaStreamDestroy(hp->st_unpack); }
    for (auto e : hp->ev_h2d) if (e) cudaEventDestroy(e);
    for (auto e : hp->ev_free) if (e) cudaEventDestroy(e);
    for (auto e : hp->ev_raw) if (e) cudaEventDestroy(e);
    if (hp->ev_join) cudaEventDestroy(hp->ev_join);
    if (hp->d_ring) cudaFree(hp->d_ring);
    if (hp->h_stage) pinned_release(hp->h_stage, hp->h_cap);
    delete hp;
}

int hostpack_threads(const HostPacker *hp) { return hp ? hp->n_threads : 0; }

// text[0, bytes) (pinned host memory) -> d_text (16-byte aligned), through st_copy.  Returns when every
// block has been queued; st_copy then waits for the last expansion, so an event recorded on st_copy
// afterwards covers the whole text.  *link_bytes receives the bytes that actually crossed the link.
int hostpack_copy(HostPacker *hp, const uint8_t *text, uint64_t bytes, uint8_t *d_text, cudaStream_t st_copy,
                  uint64_t *link_bytes, uint64_t *packed_blocks)
{
    const uint32_t n_blocks = (uint32_t)((bytes + HP_BLOCK - 1) / HP_BLOCK);
    {
        std::lock_guard<std::mutex> lk(hp->mu);
        hp->text = text;
        hp->bytes = bytes;
        hp->front = 0;
        hp->back = n_blocks;
        hp->doneq.clear();
        hp->rawq.clear();
        ++hp->gen;
    }
    hp->cv_job.notify_all();
    hp->cv_buf.notify_all();
    auto fail = [&](int rc) {
        {   // stop handing out blocks, let the workers run dry
            std::lock_guard<std::mutex> lk(hp->mu);
            hp->back = hp->front;
        }
        hp->cv_buf.notify_all();
        return rc;
    };
#define HP_CUDA(call)                                                                  \
    do {                                                                               \
        cudaError_t e__ = (call);                                                      \
        if (e__ != cudaSuccess) return fail(cuda_fail(e__, #call, __FILE__, __LINE__)); \
    } while (0)
    auto block_len = [&](uint32_t b) { return std::min<uint64_t>(HP_BLOCK, bytes - (uint64_t)b * HP_BLOCK); };
    bool raw_busy[HP_RAW_INFLIGHT] = {};
    std::deque<uint32_t> my_raw;
    std::deque<HpItem> my_done;
    uint32_t scheduled = 0;
    uint64_t sent = 0, n_packed = 0;
    while (scheduled < n_blocks) {
        bool progress = false;
        {
            std::lock_guard<std::mutex> lk(hp->mu);
            my_done.swap(hp->doneq);
            for (uint32_t b : hp->rawq) my_raw.push_back(b);
            hp->rawq.clear();
        }
        for (const HpItem &it : my_done) {
            const int r = (int)(hp->ring_next++ % HP_RING);
            uint8_t *d_stage = hp->d_ring + (size_t)r * hp->stage_cap;
            if (hp->ring_used[r]) HP_CUDA(cudaStreamWaitEvent(st_copy, hp->ev_free[r], 0));
            HP_CUDA(cudaMemcpyAsync(d_stage, hp->h_stage + (size_t)it.hbuf * hp->stage_cap, it.bytes, cudaMemcpyHostToDevice, st_copy));
            HP_CUDA(cudaEventRecord(hp->ev_h2d[r], st_copy));
            hp->pending[r].push_back(it.hbuf);
            HP_CUDA(cudaStreamWaitEvent(hp->st_unpack, hp->ev_h2d[r], 0));
            uint8_t *dst = d_text + (uint64_t)it.block * HP_BLOCK;
            k_unpack<<<(it.groups + 255) / 256, 256, 0, hp->st_unpack>>>(d_stage, it.groups, dst);
            ++g_launches;
            HP_CUDA(cudaEventRecord(hp->ev_free[r], hp->st_unpack));
            hp->ring_used[r] = true;
            const uint64_t len = block_len(it.block), tail = len - (uint64_t)it.groups * 32;
            if (tail) {
                const uint64_t off = (uint64_t)it.block * HP_BLOCK + (uint64_t)it.groups * 32;
                HP_CUDA(cudaMemcpyAsync(d_text + off, text + off, tail, cudaMemcpyHostToDevice, st_copy));
            }
            sent += it.bytes + tail;
            ++n_packed;
            ++scheduled;
            progress = true;
        }
        my_done.clear();
        // staging buffers whose copy has left the host
        for (int r = 0; r < HP_RING; ++r) {
            if (hp->pending[r].empty() || cudaEventQuery(hp->ev_h2d[r]) != cudaSuccess) continue;
            {
                std::lock_guard<std::mutex> lk(hp->mu);
                for (int hb : hp->pending[r]) hp->free_bufs.push_back(hb);
            }
            hp->pending[r].clear();
            hp->cv_buf.notify_all();
        }
        cudaGetLastError();      // cudaErrorNotReady of the queries above is not an error
        // raw blocks: a few in flight, from the front (or handed back by a worker)
        for (int sl = 0; sl < HP_RAW_INFLIGHT; ++sl) {
            if (raw_busy[sl]) {
                if (cudaEventQuery(hp->ev_raw[sl]) != cudaSuccess) { cudaGetLastError(); continue; }
                raw_busy[sl] = false;
            }
            int64_t b = -1;
            if (!my_raw.empty()) {
                b = my_raw.front();
                my_raw.pop_front();
            } else {
                std::lock_guard<std::mutex> lk(hp->mu);
                if (hp->front < hp->back) b = hp->front++;
            }
            if (b < 0) break;
            const uint64_t off = (uint64_t)b * HP_BLOCK, len = block_len((uint32_t)b);
            HP_CUDA(cudaMemcpyAsync(d_text + off, text + off, len, cudaMemcpyHostToDevice, st_copy));
            HP_CUDA(cudaEventRecord(hp->ev_raw[sl], st_copy));
            raw_busy[sl] = true;
            sent += len;
            ++scheduled;
            progress = true;
        }
        if (!progress) std::this_thread::yield();
    }
    hp->cv_buf.notify_all();     // front >= back now: workers waiting for a buffer go back to sleep on cv_job
    HP_CUDA(cudaEventRecord(hp->ev_join, hp->st_unpack));
    HP_CUDA(cudaStreamWaitEvent(st_copy, hp->ev_join, 0));
#undef HP_CUDA
    if (link_bytes) *link_bytes = sent;
    if (packed_blocks) *packed_blocks = n_packed;
    return VFB_OK;
}

}  // namespace vfb
