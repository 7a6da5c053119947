// Optional call statistics for the shims: set BBG_SHIM_STATS=1 and a table of calls / wall milliseconds per
// replaced entry point is printed to stderr at exit (how much of a prove is spent behind the boundary).
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "bbgpu.h"

namespace bbg_shim
{
struct Stats
{
    static constexpr int MAX = 40;
    const char* name[MAX];
    double ms[MAX];
    double min_ms[MAX];
    unsigned long calls[MAX];
    unsigned long units[MAX];
    int count = 0;
    bool enabled = false;
    Stats()
    {
        const char* e = getenv("BBG_SHIM_STATS");
        enabled = e != nullptr && e[0] == '1';
    }
    // Printed on request only (bbg_shim_report(), called from main): the CUDA runtime may already be torn down
    // when static destructors run, so nothing here happens at process exit.
    void report()
    {
        if (!enabled) return;
        double total = 0;
        for (int i = 0; i < count; ++i) total += ms[i];
        fprintf(stderr, "bbgpu shim stats: %.1f ms behind the boundary\n", total);
        for (int i = 0; i < count; ++i)
            fprintf(stderr, "  %-36s calls %5lu  units %5lu  total %10.2f ms  best call %9.3f ms\n", name[i], calls[i], units[i], ms[i], min_ms[i]);
        fprintf(stderr, "bbgpu kernel profile (CUDA events):\n");
        for (int i = 0; i < bbg_profile_count(); ++i)
        {
            double t = 0;
            uint64_t n = 0;
            if (bbg_profile_read(i, &t, &n) == 0 && n > 0) fprintf(stderr, "  %-32s launches %5lu  %10.2f ms\n", bbg_profile_name(i), (unsigned long)n, t);
        }
    }
    // call once the library is initialised: switches the per-kernel stopwatch on when statistics are requested
    void after_init()
    {
        if (enabled) bbg_profile_enable(1);
    }
    void add(const char* n, double t, unsigned long u)
    {
        for (int i = 0; i < count; ++i)
        {
            if (strcmp(name[i], n) == 0)
            {
                ms[i] += t;
                if (t < min_ms[i]) min_ms[i] = t;
                calls[i] += 1;
                units[i] += u;
                return;
            }
        }
        if (count < MAX)
        {
            name[count] = n;
            ms[count] = t;
            min_ms[count] = t;
            calls[count] = 1;
            units[count] = u;
            ++count;
        }
    }
};
inline Stats& stats()
{
    static Stats s;
    return s;
}
// One place where every shim brings the library up (0 on success, the bbgpu error otherwise; idempotent):
//   BBG_DEVICE=d          primary CUDA device (default 0)
//   BBG_NUM_GPUS=g        drive g GPUs of the box, devices d .. d+g-1: MSMs over the SRS are cut into point ranges, one per
//                         device (bbg_init_multi; the reference's callers get all cores the same way)
//   BBG_SRS_PRECOMPUTE=0  plain Pippenger windows instead of the fixed-base tables built per ReferenceString
//   BBG_HOST_REGISTER=0   do not page-lock long-lived caller buffers in place (on by default in the shims; a prover
//                         linked with -Wl,--wrap=free gets host_buffer_free_wrap.cpp's release hook)
inline int ensure_library()
{
    static int state = -1; // -1 not tried
    if (state == 0) return 0;
    const char* dev = getenv("BBG_DEVICE");
    const char* gpus = getenv("BBG_NUM_GPUS");
    const int first = dev ? atoi(dev) : 0;
    int count = gpus ? atoi(gpus) : 1;
    if (count < 1) count = 1;
    if (count > 16) count = 16;
    int list[16];
    for (int i = 0; i < count; ++i) list[i] = first + i;
    int e = count > 1 ? bbg_init_multi(list, count) : bbg_init(first);
    if (e == 0) e = bbg_set_auto_srs_cache(1);
    const char* pre = getenv("BBG_SRS_PRECOMPUTE");
    if (e == 0 && !(pre != nullptr && pre[0] == '0')) e = bbg_set_srs_precompute(1);
    const char* reg = getenv("BBG_HOST_REGISTER");
    if (e == 0 && !(reg != nullptr && reg[0] == '0')) e = bbg_set_host_register_cache(1);
    if (e == 0) stats().after_init();
    state = e;
    return e;
}
struct Timer
{
    const char* name_;
    unsigned long units_;
    std::chrono::steady_clock::time_point t0_;
    Timer(const char* name, unsigned long units = 1) : name_(name), units_(units), t0_(std::chrono::steady_clock::now()) { stats(); }
    ~Timer()
    {
        Stats& s = stats();
        if (s.enabled) s.add(name_, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0_).count(), units_);
    }
};
} // namespace bbg_shim
