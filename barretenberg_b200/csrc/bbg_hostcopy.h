// Host <-> device copies for caller-owned buffers.
//
// The reference's callers hand us plain aligned_alloc (pageable) memory.  cudaMemcpyAsync on pageable memory is
// staged by the driver through one bounce buffer with a single-threaded copy (~5-6 GB/s measured on the B200 box,
// against ~55 GB/s for pinned memory), which made the PLONK prover's NTT calls copy-bound.  Here pageable buffers
// go through a small ring of pinned chunks filled / drained by several host threads while the previous chunk is on
// the wire; buffers that are already pinned (bench.py, cudaHostRegister'ed memory) are copied directly.
#pragma once
#include "bbg_rt.h"

#include <string.h>

#ifndef BBG_EMULATE
#include <thread>
#include <vector>

namespace bbg_hostcopy
{
constexpr size_t CHUNK = (size_t)8 << 20;
constexpr int RING = 4;
constexpr size_t SMALL = (size_t)1 << 20; // below this the plain path is as good

struct Ring
{
    void* buf[RING] = {};
    cudaEvent_t ev[RING] = {};
    bool ready = false;
    int threads = 1;
    unsigned next = 0; // slots rotate across calls, so a short copy never waits for the previous call's chunk
    int init()
    {
        if (ready) return 0;
        for (int i = 0; i < RING; ++i)
        {
            cudaError_t e = cudaHostAlloc(&buf[i], CHUNK, cudaHostAllocDefault);
            if (e != cudaSuccess) return (int)e;
            e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
            if (e != cudaSuccess) return (int)e;
        }
        unsigned hc = std::thread::hardware_concurrency();
        threads = hc >= 16 ? 8 : (hc >= 8 ? 4 : (hc >= 4 ? 2 : 1));
        ready = true;
        return 0;
    }
    void release()
    {
        for (int i = 0; i < RING; ++i)
        {
            if (buf[i]) cudaFreeHost(buf[i]);
            if (ev[i]) cudaEventDestroy(ev[i]);
            buf[i] = nullptr;
            ev[i] = nullptr;
        }
        ready = false;
    }
};
inline Ring& ring()
{
    static Ring r;
    return r;
}

inline bool is_pinned(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess)
    {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// memcpy split over a few threads (a single core does not saturate even one PCIe direction)
inline void parallel_memcpy(void* dst, const void* src, size_t bytes, int threads)
{
    if (threads <= 1 || bytes < ((size_t)1 << 20))
    {
        memcpy(dst, src, bytes);
        return;
    }
    std::vector<std::thread> pool;
    const size_t part = ((bytes / threads) + 4095) & ~(size_t)4095;
    for (int t = 1; t < threads; ++t)
    {
        const size_t off = (size_t)t * part;
        if (off >= bytes) break;
        const size_t len = bytes - off < part ? bytes - off : part;
        pool.emplace_back([=]() { memcpy((char*)dst + off, (const char*)src + off, len); });
    }
    memcpy(dst, src, part < bytes ? part : bytes);
    for (auto& th : pool) th.join();
}

// `r`: the pinned staging ring to use (one per host thread that copies; the default ring belongs to the caller's thread)
inline int h2d_ring(Ring& r, void* d, const void* h, size_t bytes, cudaStream_t st)
{
    if (bytes < SMALL || is_pinned(h)) return (int)cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st);
    BBG_CHECK(r.init());
    size_t off = 0;
    while (off < bytes)
    {
        const int i = (int)(r.next++ % RING);
        const size_t len = bytes - off < CHUNK ? bytes - off : CHUNK;
        BBG_CHECK(cudaEventSynchronize(r.ev[i])); // chunk i is off the wire (an unrecorded event is complete)
        parallel_memcpy(r.buf[i], (const char*)h + off, len, r.threads);
        BBG_CHECK(cudaMemcpyAsync((char*)d + off, r.buf[i], len, cudaMemcpyHostToDevice, st));
        BBG_CHECK(cudaEventRecord(r.ev[i], st));
        off += len;
    }
    return 0;
}
inline int h2d(void* d, const void* h, size_t bytes, cudaStream_t st) { return h2d_ring(ring(), d, h, bytes, st); }

// returns with the data in h (synchronous for pageable destinations, like cudaMemcpy)
inline int d2h(void* h, const void* d, size_t bytes, cudaStream_t st)
{
    if (bytes < SMALL || is_pinned(h)) return (int)cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, st);
    Ring& r = ring();
    BBG_CHECK(r.init());
    const size_t chunks = (bytes + CHUNK - 1) / CHUNK;
    for (size_t k = 0; k < chunks + (RING - 1); ++k)
    {
        if (k < chunks)
        {
            const int i = (int)(k % RING);
            const size_t off = k * CHUNK, len = bytes - off < CHUNK ? bytes - off : CHUNK;
            BBG_CHECK(cudaMemcpyAsync(r.buf[i], (const char*)d + off, len, cudaMemcpyDeviceToHost, st));
            BBG_CHECK(cudaEventRecord(r.ev[i], st));
        }
        if (k >= (size_t)(RING - 1))
        {
            const size_t j = k - (RING - 1);
            const int i = (int)(j % RING);
            const size_t off = j * CHUNK, len = bytes - off < CHUNK ? bytes - off : CHUNK;
            BBG_CHECK(cudaEventSynchronize(r.ev[i]));
            parallel_memcpy((char*)h + off, r.buf[i], len, r.threads);
        }
    }
    return 0;
}
} // namespace bbg_hostcopy
#else
namespace bbg_hostcopy
{
inline int h2d(void* d, const void* h, size_t bytes, cudaStream_t st) { return bbg_rt::h2d(d, h, bytes, st); }
inline int d2h(void* h, const void* d, size_t bytes, cudaStream_t st) { return bbg_rt::d2h(h, d, bytes, st); }
} // namespace bbg_hostcopy
#endif
