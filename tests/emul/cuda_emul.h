// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// A minimal "CUDA on CPU threads" shim so the kernels under barretenberg_b200/csrc can be compiled
// with g++ (-DBBG_EMULATE) and exercised against the oracle on machines without a GPU.  It exists
// to debug index arithmetic and limb logic before spending B200 minutes; it is never built into
// libbbgpu.so and nothing in the product path can reach it.
//
// Model: one launch = for every block (sequentially) run blockDim.x OS threads; __syncthreads() is a
// pthread barrier; __shared__ statics are plain statics (blocks never overlap); dynamic shared
// memory is one heap buffer per launch.  Kernels launched with BBG_LAUNCH_NOSYNC promise not to
// call __syncthreads(): their blocks are spread over the host cores and threads run in a loop.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <atomic>
#include <functional>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(x) __attribute__((aligned(x)))

struct uint4 { uint32_t x, y, z, w; } __attribute__((aligned(16)));
struct uint2 { uint32_t x, y; } __attribute__((aligned(8)));
struct dim3
{
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
typedef void* cudaStream_t;

namespace emul
{
extern thread_local dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
extern thread_local pthread_barrier_t* t_barrier;
extern thread_local unsigned char* t_dyn_smem;

inline void syncthreads()
{
    if (t_barrier) pthread_barrier_wait(t_barrier);
}

inline void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body)
{
    const unsigned nthreads = block.x * block.y * block.z;
    unsigned char* smem = (unsigned char*)aligned_alloc(128, ((smem_bytes + 127) / 128 + 1) * 128);
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, nullptr, nthreads);
    std::vector<std::thread> pool;
    pool.reserve(nthreads);
    for (unsigned t = 0; t < nthreads; ++t)
    {
        pool.emplace_back([&, t]() {
            t_barrier = &bar;
            t_dyn_smem = smem;
            t_blockDim = block;
            t_gridDim = grid;
            t_threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            for (unsigned bz = 0; bz < grid.z; ++bz)
                for (unsigned by = 0; by < grid.y; ++by)
                    for (unsigned bx = 0; bx < grid.x; ++bx)
                    {
                        t_blockIdx = dim3(bx, by, bz);
                        body();
                        pthread_barrier_wait(&bar); // block boundary: statics may be reused
                    }
        });
    }
    for (auto& th : pool) th.join();
    pthread_barrier_destroy(&bar);
    free(smem);
}

inline void launch_nosync(dim3 grid, dim3 block, const std::function<void()>& body)
{
    const unsigned nblocks = grid.x * grid.y * grid.z;
    unsigned nworkers = std::thread::hardware_concurrency();
    if (nworkers == 0) nworkers = 4;
    if (nworkers > nblocks) nworkers = nblocks;
    std::atomic<unsigned> next(0);
    std::vector<std::thread> pool;
    for (unsigned w = 0; w < nworkers; ++w)
    {
        pool.emplace_back([&]() {
            t_barrier = nullptr;
            t_dyn_smem = nullptr;
            t_blockDim = block;
            t_gridDim = grid;
            for (;;)
            {
                unsigned b = next.fetch_add(1);
                if (b >= nblocks) break;
                t_blockIdx = dim3(b % grid.x, (b / grid.x) % grid.y, b / (grid.x * grid.y));
                for (unsigned tz = 0; tz < block.z; ++tz)
                    for (unsigned ty = 0; ty < block.y; ++ty)
                        for (unsigned tx = 0; tx < block.x; ++tx)
                        {
                            t_threadIdx = dim3(tx, ty, tz);
                            body();
                        }
            }
        });
    }
    for (auto& th : pool) th.join();
}
} // namespace emul

#define threadIdx (emul::t_threadIdx)
#define blockIdx (emul::t_blockIdx)
#define blockDim (emul::t_blockDim)
#define gridDim (emul::t_gridDim)
#define __syncthreads() emul::syncthreads()

inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicMax(unsigned* p, unsigned v)
{
    unsigned old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
inline unsigned __brev(unsigned x)
{
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}
inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
inline int __clzll(long long x) { return x == 0 ? 64 : __builtin_clzll((unsigned long long)x); }
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s)
{
    const unsigned long long both = ((unsigned long long)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) r |= (unsigned)((both >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
}
template <typename T> inline T __ldg(const T* p) { return *p; }
inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
