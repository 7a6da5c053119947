// Thin runtime seam between the C-ABI driver code and CUDA.
//
// Product build (nvcc): real CUDA runtime calls and <<<>>> launches.
// Emulation build (g++ -DBBG_EMULATE, tests/emul only): the same driver code runs its kernels on CPU
// threads through tests/emul/cuda_emul.h so index logic can be checked without a GPU.  The emulation
// is test infrastructure; libbbgpu.so is never built with BBG_EMULATE.
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef BBG_EMULATE
#include "cuda_emul.h"
#include <stdio.h>

#define BBG_LAUNCH(kernel, grid, block, smem, stream, ...) emul::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })
#define BBG_LAUNCH_NOSYNC(kernel, grid, block, stream, ...) emul::launch_nosync((grid), (block), [&]() { kernel(__VA_ARGS__); })
#define BBG_DYN_SMEM(name) unsigned char* name = emul::t_dyn_smem
#define BBG_CONSTANT static const

namespace bbg_rt
{
inline int dev_alloc(void** p, size_t bytes) { *p = aligned_alloc(256, ((bytes + 255) / 256 + 1) * 256); return *p ? 0 : 2; }
inline int dev_free(void* p) { free(p); return 0; }
inline int h2d(void* d, const void* h, size_t bytes, cudaStream_t) { memcpy(d, h, bytes); return 0; }
inline int d2h(void* h, const void* d, size_t bytes, cudaStream_t) { memcpy(h, d, bytes); return 0; }
inline int d2d(void* d, const void* s, size_t bytes, cudaStream_t) { memmove(d, s, bytes); return 0; }
inline int dev_memset(void* d, int v, size_t bytes, cudaStream_t) { memset(d, v, bytes); return 0; }
inline int sync(cudaStream_t) { return 0; }
inline int last_error() { return 0; }
inline int num_sms() { return 4; }
inline int set_smem_limit(const void*, size_t) { return 0; }
} // namespace bbg_rt

#else
#include <cuda_runtime.h>
#include <vector>

#define BBG_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define BBG_LAUNCH_NOSYNC(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#define BBG_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define BBG_CONSTANT __device__ __constant__

namespace bbg_rt
{
inline int dev_alloc(void** p, size_t bytes) { return (int)cudaMalloc(p, bytes ? bytes : 256); }
inline int dev_free(void* p) { return (int)cudaFree(p); }
inline int h2d(void* d, const void* h, size_t bytes, cudaStream_t s) { return (int)cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s); }
inline int d2h(void* h, const void* d, size_t bytes, cudaStream_t s) { return (int)cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s); }
inline int d2d(void* d, const void* s_, size_t bytes, cudaStream_t s) { return (int)cudaMemcpyAsync(d, s_, bytes, cudaMemcpyDeviceToDevice, s); }
inline int dev_memset(void* d, int v, size_t bytes, cudaStream_t s) { return (int)cudaMemsetAsync(d, v, bytes, s); }
inline int sync(cudaStream_t s) { return (int)cudaStreamSynchronize(s); }
inline int last_error() { return (int)cudaGetLastError(); }
inline int num_sms()
{
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
}
inline int set_smem_limit(const void* fn, size_t bytes)
{
    return (int)cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}
} // namespace bbg_rt
#endif

// ---------------------------------------------------------------------------------------------------
// Optional per-kernel stopwatch (CUDA events on the launching stream), off by default.  bench.py switches it
// on for a separate, untimed pass to attribute the step time to kernels (roofline.achieved).
// ---------------------------------------------------------------------------------------------------
namespace bbg_prof
{
enum id
{
    NTT_PASS_A = 0,
    NTT_PASS_B,
    NTT_SMALL,
    NTT_TABLES,
    MSM_DIGITS,
    MSM_SCAN,
    MSM_SCATTER,
    MSM_ACCUMULATE,
    MSM_FIXUP,
    MSM_CHUNK,
    MSM_REDUCE,
    MSM_HOST_FINISH,
    G1_GENERATE,
    PLONK_ELEMENTWISE,
    PLONK_SCAN,
    PLONK_EVAL,
    MSM_PAIR,
    NUM_IDS
};
inline const char* name(int i)
{
    static const char* n[NUM_IDS] = { "ntt_pass_a", "ntt_pass_b", "ntt_small", "ntt_tables", "msm_digits", "msm_scan", "msm_scatter",
                                      "msm_accumulate", "msm_fixup", "msm_chunk", "msm_reduce", "msm_host_finish", "g1_generate",
                                      "plonk_elementwise", "plonk_scan", "plonk_eval", "msm_pair_rounds" };
    return (i >= 0 && i < NUM_IDS) ? n[i] : "?";
}
struct State
{
    bool on = false;
    double total_ms[NUM_IDS] = {};
    unsigned long long count[NUM_IDS] = {};
#ifndef BBG_EMULATE
    struct Rec
    {
        int id;
        cudaEvent_t a, b;
    };
    std::vector<Rec> pending;
#endif
};
inline State& state()
{
    static State s;
    return s;
}
// The stopwatch belongs to the thread that drives the primary device; the per-device worker threads of the multi-GPU MSM
// (bbg_msm.cu) mute it for themselves.
inline bool& thread_muted()
{
    static thread_local bool muted = false;
    return muted;
}
inline void collect()
{
#ifndef BBG_EMULATE
    State& s = state();
    for (auto& r : s.pending)
    {
        cudaEventSynchronize(r.b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        s.total_ms[r.id] += ms;
        s.count[r.id] += 1;
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    s.pending.clear();
#endif
}
inline void reset()
{
    collect();
    State& s = state();
    for (int i = 0; i < NUM_IDS; ++i)
    {
        s.total_ms[i] = 0;
        s.count[i] = 0;
    }
}
// RAII: everything launched on `st` during the scope's lifetime is attributed to `id`
struct Scope
{
#ifndef BBG_EMULATE
    int id_;
    cudaStream_t st_;
    cudaEvent_t a_ = nullptr, b_ = nullptr;
    bool live_;
    Scope(int i, cudaStream_t st) : id_(i), st_(st), live_(state().on && !thread_muted())
    {
        if (!live_) return;
        cudaEventCreate(&a_);
        cudaEventCreate(&b_);
        cudaEventRecord(a_, st_);
    }
    ~Scope()
    {
        if (!live_) return;
        cudaEventRecord(b_, st_);
        state().pending.push_back({ id_, a_, b_ });
    }
#else
    Scope(int, cudaStream_t) {}
#endif
};
inline void add_host_ms(int i, double ms)
{
    if (!state().on || thread_muted()) return;
    state().total_ms[i] += ms;
    state().count[i] += 1;
}
} // namespace bbg_prof
