// Integer multiply-add throughput microbenchmarks: the measured IMAD roofline denominator that
// bench.py reports field-multiplication throughput against ("measure it").
#include "bbg_internal.h"

namespace bbg
{
namespace
{
constexpr int MB_THREADS = 256;
constexpr int MB_CHAINS = 8;

// mode 0: 8 independent mad.lo.u32 chains per thread;  1: mad.wide.u32 with 64-bit accumulators
template <int MODE> __global__ void __launch_bounds__(MB_THREADS) imad_kernel(uint32_t* out, uint32_t seed, int iters)
{
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t x = seed ^ tid, y = (seed * 2654435761u) | 1u;
    if (MODE == 0)
    {
        uint32_t acc[MB_CHAINS];
#pragma unroll
        for (int i = 0; i < MB_CHAINS; ++i) acc[i] = tid + i;
        for (int it = 0; it < iters; ++it)
        {
#pragma unroll
            for (int u = 0; u < 4; ++u)
            {
#pragma unroll
                for (int i = 0; i < MB_CHAINS; ++i) acc[i] = acc[i] * x + y;
            }
        }
        uint32_t s = 0;
#pragma unroll
        for (int i = 0; i < MB_CHAINS; ++i) s ^= acc[i];
        out[tid] = s;
    }
    else
    {
        uint64_t acc[MB_CHAINS];
#pragma unroll
        for (int i = 0; i < MB_CHAINS; ++i) acc[i] = tid + i;
        for (int it = 0; it < iters; ++it)
        {
#pragma unroll
            for (int u = 0; u < 4; ++u)
            {
#pragma unroll
                for (int i = 0; i < MB_CHAINS; ++i) acc[i] = (uint64_t)((uint32_t)acc[i] ^ x) * y + acc[i];
            }
        }
        uint64_t s = 0;
#pragma unroll
        for (int i = 0; i < MB_CHAINS; ++i) s ^= acc[i];
        out[tid] = (uint32_t)s ^ (uint32_t)(s >> 32);
    }
}

// mode 2: the carry-chained pairs the Montgomery product is made of (4 independent 256-bit rows per thread)
__global__ void __launch_bounds__(MB_THREADS) imad_chain_kernel(uint32_t* out, uint32_t seed, int iters)
{
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc[4][8], x[8], top[4] = { 0, 0, 0, 0 };
#pragma unroll
    for (int i = 0; i < 8; ++i)
    {
        x[i] = (seed + i) * 2654435761u ^ tid;
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r][i] = tid + i + r;
    }
    uint32_t y = seed | 1u;
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int r = 0; r < 4; ++r) cc::mad_row_carry(acc[r], top[r], x, y + r);
        y += top[0];
    }
    uint32_t s = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r)
    {
#pragma unroll
        for (int i = 0; i < 8; ++i) s ^= acc[r][i];
        s ^= top[r];
    }
    out[tid] = s;
}

// modes 3/4: Montgomery products, two independent dependency chains per thread
template <typename F> __global__ void __launch_bounds__(MB_THREADS) field_mul_kernel(uint32_t* out, uint32_t seed, int iters)
{
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    // every chain must depend on the thread id, or ptxas moves it to the uniform datapath (UIMAD) and the
    // vector-pipe throughput is overstated
    fe a = F::one(), b = F::one(), m = F::one();
    a.v[0] ^= tid;
    b.v[1] ^= seed + tid * 2654435761u;
    m.v[0] += (seed ^ tid) & 0xff;
    for (int it = 0; it < iters; ++it)
    {
        a = F::mul(a, m);
        b = F::mul(b, m);
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= a.v[i] ^ b.v[i];
    out[tid] = s;
}
// modes 5-7: CHAINS independent product chains per thread at the NTT kernel's occupancy (2 x 256 threads per SM)
template <int CHAINS> __global__ void __launch_bounds__(MB_THREADS, 2) field_mul_ilp_kernel(uint32_t* out, uint32_t seed, int iters)
{
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    fe a[CHAINS], m = Fr::one();
    m.v[0] += (seed ^ tid) & 0xff;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c)
    {
        a[c] = Fr::one();
        a[c].v[c & 7] ^= tid + c;
    }
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) a[c] = Fr::mul(a[c], m);
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c)
    {
#pragma unroll
        for (int i = 0; i < 8; ++i) s ^= a[c].v[i];
    }
    out[tid] = s;
}
} // namespace

// returns ops (32x32 multiply-adds, or field products) issued by the whole grid
int microbench_launch(int mode, int iters, uint32_t* d_out, int blocks, double* ops, cudaStream_t st)
{
    const double threads = (double)blocks * MB_THREADS;
    switch (mode)
    {
    case 0:
        BBG_LAUNCH_NOSYNC(imad_kernel<0>, dim3(blocks), dim3(MB_THREADS), st, d_out, 12345u, iters);
        *ops = threads * iters * 4.0 * MB_CHAINS;
        break;
    case 1:
        BBG_LAUNCH_NOSYNC(imad_kernel<1>, dim3(blocks), dim3(MB_THREADS), st, d_out, 12345u, iters);
        *ops = threads * iters * 4.0 * MB_CHAINS;
        break;
    case 2:
        BBG_LAUNCH_NOSYNC(imad_chain_kernel, dim3(blocks), dim3(MB_THREADS), st, d_out, 12345u, iters);
        *ops = threads * iters * 16.0; // 4 rows x 4 wide multiply-adds
        break;
    case 3:
        BBG_LAUNCH_NOSYNC(field_mul_kernel<Fq>, dim3(blocks), dim3(MB_THREADS), st, d_out, 12345u, iters);
        *ops = threads * iters * 2.0;
        break;
    case 4:
        BBG_LAUNCH_NOSYNC(field_mul_kernel<Fr>, dim3(blocks), dim3(MB_THREADS), st, d_out, 12345u, iters);
        *ops = threads * iters * 2.0;
        break;
    case 5:
    case 6:
    case 7:
    {
        // 2 resident CTAs per SM exactly: grid = 2 x #SM, the 16 warps/SM the NTT pass kernels run at
        const int sm2 = 2 * bbg_rt::num_sms();
        const double thr = (double)sm2 * MB_THREADS;
        if (mode == 5) { BBG_LAUNCH_NOSYNC(field_mul_ilp_kernel<1>, dim3(sm2), dim3(MB_THREADS), st, d_out, 12345u, iters * 8); *ops = thr * iters * 8.0; }
        if (mode == 6) { BBG_LAUNCH_NOSYNC(field_mul_ilp_kernel<2>, dim3(sm2), dim3(MB_THREADS), st, d_out, 12345u, iters * 4); *ops = thr * iters * 8.0; }
        if (mode == 7) { BBG_LAUNCH_NOSYNC(field_mul_ilp_kernel<4>, dim3(sm2), dim3(MB_THREADS), st, d_out, 12345u, iters * 2); *ops = thr * iters * 8.0; }
        break;
    }
    default: return 1007;
    }
    return bbg_rt::last_error();
}
} // namespace bbg
