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
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace bbg_hostcopy
{
constexpr size_t CHUNK = (size_t)8 << 20;
constexpr int RING = 4;
constexpr size_t SMALL = (size_t)1 << 20; // below this the plain path is as good

// memcpy split over a few threads (a single core does not saturate even one PCIe direction).  The workers are
// persistent: creating and joining threads per 8 MiB chunk cost about as much as the copy itself.
class CopyPool
{
  public:
    explicit CopyPool(int workers) : stop_(false), generation_(0), pending_(0)
    {
        for (int i = 0; i < workers; ++i) threads_.emplace_back([this, i]() { run(i); });
    }
    ~CopyPool()
    {
        {
            std::lock_guard<std::mutex> lock(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : threads_) t.join();
    }
    int workers() const { return (int)threads_.size(); }
    // copies [0, bytes) with the calling thread taking one slice; returns when every slice is done
    void copy(void* dst, const void* src, size_t bytes)
    {
        const int parts = workers() + 1;
        const size_t part = (((bytes + parts - 1) / parts) + 4095) & ~(size_t)4095;
        {
            std::lock_guard<std::mutex> lock(m_);
            dst_ = (char*)dst;
            src_ = (const char*)src;
            bytes_ = bytes;
            part_ = part;
            pending_ = workers();
            ++generation_;
        }
        cv_.notify_all();
        slice(workers()); // the caller's share: the last slice
        std::unique_lock<std::mutex> lock(m_);
        done_.wait(lock, [this]() { return pending_ == 0; });
    }

  private:
    void slice(int idx)
    {
        const size_t off = (size_t)idx * part_;
        if (off >= bytes_) return;
        const size_t len = bytes_ - off < part_ ? bytes_ - off : part_;
        memcpy(dst_ + off, src_ + off, len);
    }
    void run(int idx)
    {
        unsigned long seen = 0;
        for (;;)
        {
            {
                std::unique_lock<std::mutex> lock(m_);
                cv_.wait(lock, [&]() { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
            }
            slice(idx);
            {
                std::lock_guard<std::mutex> lock(m_);
                if (--pending_ == 0) done_.notify_all();
            }
        }
    }
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    bool stop_;
    unsigned long generation_;
    int pending_;
    char* dst_ = nullptr;
    const char* src_ = nullptr;
    size_t bytes_ = 0, part_ = 0;
};

struct Ring
{
    void* buf[RING] = {};
    cudaEvent_t ev[RING] = {};
    bool ready = false;
    int threads = 1;
    unsigned next = 0; // slots rotate across calls, so a short copy never waits for the previous call's chunk
    std::unique_ptr<CopyPool> pool; // one pool per ring: each ring belongs to one copying thread
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
        if (threads > 1) pool.reset(new CopyPool(threads - 1));
        ready = true;
        return 0;
    }
    void parallel_memcpy(void* dst, const void* src, size_t bytes)
    {
        if (!pool || bytes < ((size_t)1 << 20)) memcpy(dst, src, bytes);
        else pool->copy(dst, src, bytes);
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
        pool.reset();
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
        r.parallel_memcpy(r.buf[i], (const char*)h + off, len);
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
            r.parallel_memcpy((char*)h + off, r.buf[i], len);
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
