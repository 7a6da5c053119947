// Host <-> device copies for caller-owned buffers.
//
// The reference's callers hand us plain aligned_alloc (pageable) memory.  cudaMemcpyAsync on pageable memory is
// staged by the driver through one bounce buffer with a single-threaded copy (~5-6 GB/s measured on the B200 box,
// against ~55 GB/s for pinned memory), which made the PLONK prover's NTT calls copy-bound.  Here pageable buffers
// go through a small ring of pinned chunks filled / drained by several host threads while the previous chunk is on
// the wire; buffers that are already pinned (bench.py, cudaHostRegister'ed memory) are copied directly.
//
// Registration cache (off by default, bbg_set_host_register_cache; the shims switch it on): the reference's callers
// reuse the same long-lived buffers call after call (barretenberg::polynomial members of the prover and the proving key,
// ReferenceString::monomials), so a pageable buffer of >= 1 MiB that is seen a SECOND time behind the same address is
// page-locked in place with cudaHostRegister and from then on copied at the pinned rate without the staging memcpy.
// The caller's side of the contract: a registered buffer must be handed to bbg_host_buffer_forget() before it is freed
// (the prover link wraps free() to do that, shim/host_buffer_free_wrap.cpp) — the driver keeps DMA mappings of the
// physical pages, and an address range that was unmapped and mapped again would be read / written through the old ones.
#pragma once
#include "bbg_rt.h"

#include <string.h>

#ifndef BBG_EMULATE
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <stdlib.h>
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

// ---- cudaHostRegister cache for caller-owned pageable buffers -----------------------------------------------------------
// Only the WHOLE pages inside a buffer are page-locked: heap blocks share their first and last page with their neighbours,
// and a page can be registered once.  A copy of a registered buffer is therefore three copies — a head and a tail of less
// than a page each through the driver's own pageable path, the page-locked middle directly.
struct PinnedSpan
{
    size_t head = 0; // bytes before the page-locked middle
    size_t mid = 0;  // page-locked bytes (0: nothing usable)
    bool whole = false; // the caller pinned the whole buffer itself (cudaHostAlloc / cudaHostRegister)
    bool mixed = false; // touches more than one registration: only a host memcpy through the staging ring is safe
};
class RegCache
{
  public:
    void enable(bool on)
    {
        {
            std::lock_guard<std::mutex> lock(m_);
            enabled_ = on;
            if (const char* e = getenv("BBG_HOST_REGISTER_MAX_MB"))
            {
                const long v = atol(e);
                if (v >= 0) max_bytes_ = (size_t)v << 20;
            }
            if (const char* e = getenv("BBG_HOST_REGISTER_AFTER"))
            {
                const long v = atol(e);
                if (v >= 1) register_after_ = (unsigned)v;
            }
            if (!on) drop_all_locked();
        }
        flush_retired();
    }
    bool enabled() const { return enabled_; }
    // threads that must never wait for the library to go idle (the per-device MSM workers: they ARE the library's work)
    // only look registrations up
    static bool& lookup_only_thread()
    {
        static thread_local bool v = false;
        return v;
    }
    // Called for a buffer about to be copied: which part of [p, p + bytes) is page-locked.  Counts sightings (one per
    // copy: an in-place transform is two); the register_after_-th sighting of the same address and size page-locks the
    // buffer's whole pages in place.  Page-locking costs about as much as ten staged copies of the same buffer (measured on
    // the B200 box: cudaHostRegister runs at 1-2 GB/s, the staging ring at ~12 GB/s, a pinned copy at ~50 GB/s), so only
    // buffers that keep coming back are worth it; temporaries freed after one proof never get there.
    PinnedSpan classify(const void* p, size_t bytes)
    {
        if (lookup_only_thread()) return peek(p, bytes);
        PinnedSpan r;
        const char* lo = (const char*)p;
        const char* hi = lo + bytes;
        if (!enabled_)
        {
            r.whole = caller_pinned(lo, hi);
            return r;
        }
        bool retired = false;
        {
            std::lock_guard<std::mutex> lock(m_);
            ++clock_;
            if (touching_locked(lo, hi, &r)) return r;
            if (caller_pinned(lo, hi))
            {
                r.whole = true;
                return r;
            }
            Entry* e = nullptr;
            for (Entry& x : entries_)
                if (x.base == lo && x.bytes == bytes) e = &x;
            if (e == nullptr)
            {
                if (entries_.size() >= MAX_ENTRIES) retired = evict_locked(/*registered_only=*/false) || retired;
                Entry n;
                n.base = lo;
                n.bytes = bytes;
                n.used = clock_;
                n.sightings = 1;
                entries_.push_back(n);
                count_.store(entries_.size(), std::memory_order_release);
            }
            else
            {
                e->used = clock_;
                if (++e->sightings >= register_after_ && !e->failed)
                {
                    // make room first: an older registration behind overlapping addresses is gone (freed without a forget,
                    // or the same allocation seen with another size), and the page-locked total has a cap.  Retired
                    // registrations are released outside the lock (below); this buffer registers at its next sighting.
                    if (make_room_locked(*e)) retired = true;
                    else if (register_locked(*e)) r = span_of(*e, lo, hi);
                }
            }
        }
        if (retired) flush_retired();
        return r;
    }
    // the same answer without counting a sighting or registering anything
    PinnedSpan peek(const void* p, size_t bytes)
    {
        PinnedSpan r;
        const char* lo = (const char*)p;
        const char* hi = lo + bytes;
        if (enabled_ && count_.load(std::memory_order_acquire) != 0)
        {
            std::lock_guard<std::mutex> lock(m_);
            if (touching_locked(lo, hi, &r)) return r;
        }
        r.whole = caller_pinned(lo, hi);
        return r;
    }
    // the caller is about to free (or has remapped) the buffer that starts at / contains p
    void forget(const void* p)
    {
        if (count_.load(std::memory_order_acquire) == 0) return; // fast path: free() wrappers call this for every block
        bool retired = false;
        {
            std::lock_guard<std::mutex> lock(m_);
            for (size_t i = 0; i < entries_.size();)
            {
                Entry& e = entries_[i];
                if ((const char*)p >= e.base && (const char*)p < e.base + e.bytes)
                {
                    retired = retire_locked(e) || retired;
                    entries_.erase(entries_.begin() + (long)i);
                }
                else
                    ++i;
            }
            count_.store(entries_.size(), std::memory_order_release);
        }
        if (retired) flush_retired();
    }
    void release()
    {
        {
            std::lock_guard<std::mutex> lock(m_);
            drop_all_locked();
        }
        flush_retired();
    }
    size_t registered_bytes() const { return total_; }
    void stats(double* ms, unsigned long long* bytes, unsigned long long* count) const
    {
        *ms = stat_ms_;
        *bytes = stat_bytes_;
        *count = stat_count_;
    }
    // hook for "nothing of this library is still in flight on any device" (set by the C-ABI layer; default: this device)
    void set_quiesce(void (*fn)()) { quiesce_ = fn; }

  private:
    struct Entry
    {
        const char* base = nullptr;
        size_t bytes = 0;
        unsigned long used = 0;
        unsigned sightings = 0;
        bool registered = false;
        bool failed = false;
        char* reg_base = nullptr; // the whole pages inside [base, base + bytes)
        size_t reg_bytes = 0;
    };
    static constexpr size_t MAX_ENTRIES = 96;
    static constexpr size_t PAGE = 4096;
    static bool caller_pinned(const char* lo, const char* hi) { return hi > lo && is_pinned(lo) && is_pinned(hi - 1); }
    static PinnedSpan span_of(const Entry& e, const char* lo, const char* hi)
    {
        PinnedSpan r;
        const char* a = lo > e.reg_base ? lo : e.reg_base;
        const char* b = hi < e.reg_base + e.reg_bytes ? hi : e.reg_base + e.reg_bytes;
        if (b > a)
        {
            r.head = (size_t)(a - lo);
            r.mid = (size_t)(b - a);
        }
        return r;
    }
    // Any range that touches page-locked pages of ours is copied in pieces: one cudaMemcpyAsync must never span
    // page-locked and pageable memory, and a sub-range of a registered buffer (a device's share of the scalars, the low n
    // coefficients of a 4n polynomial) usually starts in the buffer's unregistered first page.
    bool touching_locked(const char* lo, const char* hi, PinnedSpan* out)
    {
        int touching = 0;
        Entry* hit = nullptr;
        for (Entry& e : entries_)
        {
            if (e.registered && lo < e.reg_base + e.reg_bytes && hi > e.reg_base)
            {
                ++touching;
                hit = &e;
            }
        }
        if (touching == 0) return false;
        if (touching == 1)
        {
            hit->used = clock_;
            *out = span_of(*hit, lo, hi);
        }
        else
            out->mixed = true;
        return true;
    }
    static void inner_pages(const Entry& e, char** lo, char** hi)
    {
        *lo = (char*)(((uintptr_t)e.base + PAGE - 1) & ~(uintptr_t)(PAGE - 1));
        *hi = (char*)(((uintptr_t)e.base + e.bytes) & ~(uintptr_t)(PAGE - 1));
    }
    // true when something had to be retired (then nothing is registered now: the caller flushes and tries again later)
    bool make_room_locked(Entry& e)
    {
        char *lo, *hi;
        inner_pages(e, &lo, &hi);
        if (hi <= lo + 16 * PAGE)
        {
            e.failed = true;
            return false;
        }
        const size_t len = (size_t)(hi - lo);
        bool retired = false;
        for (Entry& o : entries_)
            if (&o != &e && o.registered && lo < o.reg_base + o.reg_bytes && hi > o.reg_base) retired = retire_locked(o) || retired;
        while (total_ + len > max_bytes_)
        {
            Entry* oldest = nullptr;
            for (Entry& o : entries_)
                if (&o != &e && o.registered && (oldest == nullptr || o.used < oldest->used)) oldest = &o;
            if (oldest == nullptr)
            {
                e.failed = true; // larger than the cap on its own
                return retired;
            }
            retired = retire_locked(*oldest) || retired;
        }
        return retired || !retired_.empty();
    }
    bool register_locked(Entry& e)
    {
        char *lo, *hi;
        inner_pages(e, &lo, &hi);
        const size_t len = (size_t)(hi - lo);
        const auto t0 = std::chrono::steady_clock::now();
        const cudaError_t err = cudaHostRegister(lo, len, cudaHostRegisterPortable);
        if (err != cudaSuccess)
        {
            cudaGetLastError();
            // (pages still held by a registration that is being retired on another thread: try again next time)
            if (err != cudaErrorHostMemoryAlreadyRegistered) e.failed = true;
            return false;
        }
        stat_ms_ += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        stat_bytes_ += len;
        stat_count_ += 1;
        e.registered = true;
        e.reg_base = lo;
        e.reg_bytes = len;
        total_ += len;
        return true;
    }
    // no new copy will use e's pages from here on; the pages themselves are released by flush_retired(), outside the lock
    bool retire_locked(Entry& e)
    {
        if (!e.registered) return false;
        retired_.push_back(e.reg_base);
        total_ -= e.reg_bytes;
        e.registered = false;
        e.sightings = 0;
        return true;
    }
    void flush_retired()
    {
        std::vector<char*> list;
        {
            std::lock_guard<std::mutex> lock(m_);
            list.swap(retired_);
        }
        if (list.empty()) return;
        if (quiesce_) quiesce_(); // nothing may still be on the wire from / to these pages
        else cudaDeviceSynchronize();
        for (char* p : list) cudaHostUnregister(p);
        cudaGetLastError();
    }
    bool evict_locked(bool registered_only)
    {
        long best = -1;
        for (size_t i = 0; i < entries_.size(); ++i)
        {
            if (registered_only && !entries_[i].registered) continue;
            if (best < 0 || entries_[i].used < entries_[(size_t)best].used) best = (long)i;
        }
        if (best < 0) return false;
        const bool retired = retire_locked(entries_[(size_t)best]);
        entries_.erase(entries_.begin() + best);
        count_.store(entries_.size(), std::memory_order_release);
        return retired;
    }
    void drop_all_locked()
    {
        for (Entry& e : entries_) retire_locked(e);
        entries_.clear();
        count_.store(0, std::memory_order_release);
    }
    std::mutex m_;
    std::vector<Entry> entries_;
    std::vector<char*> retired_;
    std::atomic<size_t> count_{ 0 };
    bool enabled_ = false;
    unsigned long clock_ = 0;
    size_t total_ = 0;
    size_t max_bytes_ = (size_t)16 << 30;
    unsigned register_after_ = 6;
    double stat_ms_ = 0;
    unsigned long long stat_bytes_ = 0, stat_count_ = 0;
    void (*quiesce_)() = nullptr;
};
inline RegCache& reg_cache()
{
    static RegCache c;
    return c;
}

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


// one buffer, up to three copies: pageable head (< 1 page), page-locked middle, pageable tail
inline int copy_spans(void* d, const void* h, size_t bytes, const PinnedSpan& sp, bool to_device, cudaStream_t st)
{
    const cudaMemcpyKind kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    auto part = [&](size_t off, size_t len) -> int {
        if (len == 0) return 0;
        return to_device ? (int)cudaMemcpyAsync((char*)d + off, (const char*)h + off, len, kind, st)
                         : (int)cudaMemcpyAsync((char*)const_cast<void*>(h) + off, (const char*)d + off, len, kind, st);
    };
    if (to_device)
    {
        // the two fragments first (the driver stages them and returns), then the asynchronous middle
        BBG_CHECK(part(0, sp.head));
        BBG_CHECK(part(sp.head + sp.mid, bytes - sp.head - sp.mid));
        return part(sp.head, sp.mid);
    }
    BBG_CHECK(part(sp.head, sp.mid));
    BBG_CHECK(part(0, sp.head)); // pageable destination: returns once everything queued before it, and itself, has landed
    return part(sp.head + sp.mid, bytes - sp.head - sp.mid);
}

// `r`: the pinned staging ring to use (one per host thread that copies; the default ring belongs to the caller's thread)
inline int h2d_ring(Ring& r, void* d, const void* h, size_t bytes, cudaStream_t st)
{
    const PinnedSpan sp = bytes < SMALL ? reg_cache().peek(h, bytes) : reg_cache().classify(h, bytes);
    if (sp.mid > 0) return copy_spans(d, h, bytes, sp, true, st);
    if (sp.whole || (bytes < SMALL && !sp.mixed)) return (int)cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st);
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
// true when a copy from h queued by h2d may still be reading h after h2d returned (the caller pinned it, or the
// registration cache did): callers that promise "the buffer is yours again on return" wait for the copy in that case
inline bool copy_is_async(const void* h, size_t bytes)
{
    const PinnedSpan sp = reg_cache().peek(h, bytes);
    return sp.whole || sp.mid > 0;
}
inline bool caller_pinned_whole(const void* h, size_t bytes) { return reg_cache().peek(h, bytes).whole; }

// returns with the data in h (synchronous for pageable destinations, like cudaMemcpy)
inline int d2h_ring(Ring& r, void* h, const void* d, size_t bytes, cudaStream_t st)
{
    const PinnedSpan sp = bytes < SMALL ? reg_cache().peek(h, bytes) : reg_cache().classify(h, bytes);
    if (sp.mid > 0) return copy_spans(const_cast<void*>(d), h, bytes, sp, false, st);
    if (sp.whole || (bytes < SMALL && !sp.mixed)) return (int)cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, st);
    BBG_CHECK(r.init());
    const size_t chunks = (bytes + CHUNK - 1) / CHUNK;
    for (size_t k = 0; k < chunks + (RING - 1); ++k)
    {
        if (k < chunks)
        {
            const int i = (int)(k % RING);
            const size_t off = k * CHUNK, len = bytes - off < CHUNK ? bytes - off : CHUNK;
            // the slot may still hold an upload queued by an earlier h2d on another stream (an MSM launched beside the
            // transforms): wait until it is off the wire before overwriting it
            if (k < (size_t)RING) BBG_CHECK(cudaEventSynchronize(r.ev[i]));
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
inline int d2h(void* h, const void* d, size_t bytes, cudaStream_t st) { return d2h_ring(ring(), h, d, bytes, st); }
} // namespace bbg_hostcopy
#else
namespace bbg_hostcopy
{
inline int h2d(void* d, const void* h, size_t bytes, cudaStream_t st) { return bbg_rt::h2d(d, h, bytes, st); }
inline int d2h(void* h, const void* d, size_t bytes, cudaStream_t st) { return bbg_rt::d2h(h, d, bytes, st); }
struct RegCache
{
    void enable(bool) {}
    void forget(const void*) {}
    void release() {}
    void set_quiesce(void (*)()) {}
    void stats(double* ms, unsigned long long* bytes, unsigned long long* count) const { *ms = 0; *bytes = 0; *count = 0; }
};
inline RegCache& reg_cache()
{
    static RegCache c;
    return c;
}
} // namespace bbg_hostcopy
#endif
