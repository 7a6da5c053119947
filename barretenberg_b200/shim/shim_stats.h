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
