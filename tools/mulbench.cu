// Development aid: device throughput of the Montgomery product (Fr::mul) against the precomputed-quotient product
// (Fr::mul_const), alone and inside a DIF butterfly, and of the dedicated square (Fr::sqr) against mul(x, x), plus a
// device-vs-device agreement check of each pair.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/mulbench tools/mulbench.cu && build/mulbench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../barretenberg_b200/csrc/bbg_field.cuh"
using namespace bbg;

constexpr int CHAINS = 4;
template <int MODE> __global__ void __launch_bounds__(256, 2) bench_kernel(fe* out, const fe* in, const fe* tw, int iters)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    fe x[CHAINS], y[CHAINS];
    for (int c = 0; c < CHAINS; ++c) { x[c] = load_fe(in + (t * CHAINS + c) % 4096); y[c] = load_fe(in + (t * CHAINS + c + 7) % 4096); }
    const fe wm = load_fe(tw + 0), wp = load_fe(tw + 1), wq = load_fe(tw + 2);
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c)
        {
            if (MODE == 0) x[c] = Fr::mul(x[c], wm);
            if (MODE == 1) x[c] = Fr::mul_const(x[c], wp, wq);
            if (MODE == 2) { const fe u = x[c], v = y[c]; x[c] = Fr::add(u, v); y[c] = Fr::mul(Fr::sub_lazy(u, v), wm); }
            if (MODE == 3) { const fe u = x[c], v = y[c]; x[c] = Fr::add(u, v); y[c] = Fr::mul_const(Fr::sub_lazy(u, v), wp, wq); }
            if (MODE == 4) x[c] = Fr::mul(x[c], x[c]);
            if (MODE == 5) x[c] = Fr::sqr(x[c]);
        }
    }
    fe acc = x[0];
    for (int c = 1; c < CHAINS; ++c) acc = Fr::add(acc, x[c]);
    if (MODE >= 2) for (int c = 0; c < CHAINS; ++c) acc = Fr::add(acc, y[c]);
    store_fe(out + t, Fr::reduce(acc));
}

int main()
{
    const int blocks = 148 * 8, threads = 256, n = blocks * threads, iters = 512;
    fe *d_in, *d_tw, *d_out[6];
    cudaMalloc(&d_in, 4096 * sizeof(fe));
    cudaMalloc(&d_tw, 3 * sizeof(fe));
    for (int m = 0; m < 6; ++m) cudaMalloc(&d_out[m], n * sizeof(fe));
    fe* h_in = new fe[4096];
    uint64_t s = 12345;
    for (int i = 0; i < 4096; ++i)
    {
        for (int l = 0; l < 8; ++l) { s = s * 6364136223846793005ull + 1442695040888963407ull; h_in[i].v[l] = (uint32_t)(s >> 32); }
        h_in[i].v[7] &= 0x3fffffffu;
        h_in[i] = Fr::reduce(h_in[i]);
    }
    fe tw[3];
    tw[0] = h_in[5];                       // Montgomery form of some w
    tw[1] = Fr::from_mont(tw[0]);          // plain w
    tw[2] = Fr::const_quotient(tw[0]);     // floor(w 2^256 / p)
    cudaMemcpy(d_in, h_in, 4096 * sizeof(fe), cudaMemcpyHostToDevice);
    cudaMemcpy(d_tw, tw, sizeof(tw), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms[6];
    for (int rep = 0; rep < 3; ++rep)
    {
        for (int m = 0; m < 6; ++m)
        {
            cudaEventRecord(e0);
            if (m == 0) bench_kernel<0><<<blocks, threads>>>(d_out[0], d_in, d_tw, iters);
            if (m == 1) bench_kernel<1><<<blocks, threads>>>(d_out[1], d_in, d_tw, iters);
            if (m == 2) bench_kernel<2><<<blocks, threads>>>(d_out[2], d_in, d_tw, iters);
            if (m == 3) bench_kernel<3><<<blocks, threads>>>(d_out[3], d_in, d_tw, iters);
            if (m == 4) bench_kernel<4><<<blocks, threads>>>(d_out[4], d_in, d_tw, iters);
            if (m == 5) bench_kernel<5><<<blocks, threads>>>(d_out[5], d_in, d_tw, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms[m], e0, e1);
        }
    }
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(err)); return 2; }
    fe* h[6];
    for (int m = 0; m < 6; ++m) { h[m] = new fe[n]; cudaMemcpy(h[m], d_out[m], n * sizeof(fe), cudaMemcpyDeviceToHost); }
    size_t bad01 = 0, bad23 = 0, bad45 = 0;
    for (int i = 0; i < n; ++i)
    {
        if (!Fr::eq_raw(h[0][i], h[1][i])) ++bad01;
        if (!Fr::eq_raw(h[2][i], h[3][i])) ++bad23;
        if (!Fr::eq_raw(h[4][i], h[5][i])) ++bad45;
    }
    const double muls = (double)n * CHAINS * iters;
    const char* names[6] = { "mul (Montgomery)", "mul_const", "butterfly + mul", "butterfly + mul_const", "mul(x, x)", "sqr(x)" };
    for (int m = 0; m < 6; ++m) printf("{\"mode\": \"%s\", \"ms\": %.4f, \"products_per_s\": %.4e}\n", names[m], ms[m], muls / (ms[m] * 1e-3));
    printf("{\"mismatch_mul\": %zu, \"mismatch_butterfly\": %zu, \"mismatch_sqr\": %zu, \"of\": %d}\n", bad01, bad23, bad45, n);
    return (bad01 || bad23 || bad45) ? 1 : 0;
}
