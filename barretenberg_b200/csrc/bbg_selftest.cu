// Device-side self test of the field and group primitives, element by element, through the C ABI
// (bbg_field_selftest / bbg_g1_selftest in include/bbgpu.h).  The device bodies of bbg_field.cuh are inline PTX
// (#ifdef __CUDA_ARCH__) and differ from the portable host bodies the CPU emulation tests run, so the parity tests
// feed the reference's known-answer vectors (test/test_fq.cpp:51-133, test_fr.cpp:51-88, test_g1.cpp:41-122) and
// 10^6 seeded operand pairs through these kernels and compare limb for limb with the oracle.
// Not a product path: nothing in the MSM / NTT / prover code calls into this file.
#include "bbg_internal.h"

namespace bbg
{
namespace
{
size_t g_selftest_launches = 0;

// raw = the integer the device routine leaves (coarse, [0, 2p) unless noted), exactly as the reference's lazy routines do
template <typename F> __global__ void field_selftest_kernel(int op, const fe* a, const fe* b, fe* out, size_t count)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const fe x = load_fe(a + i);
    const fe y = b != nullptr ? load_fe(b + i) : F::zero();
    fe r;
    switch (op)
    {
    case 0: r = F::mul(x, y); break;                  // field_impl_int128.tcc:187-225 __mul_with_coarse_reduction: (xy + Mp) / 2^256
    case 1: r = F::sqr(x); break;                     // :227-263 __sqr_with_coarse_reduction: same integer as mul(x, x)
    case 2: r = F::reduce(F::mul_const(x, F::from_mont(y), F::const_quotient(y))); break; // == __mul(x, y), y canonical
    case 3: r = F::add(x, y); break;                  // :72-96 __add_with_coarse_reduction
    case 4: r = F::sub(x, y); break;                  // :120-140 __sub_with_coarse_reduction
    case 5: r = F::reduce(x); break;                  // reduce_once
    case 6: r = F::neg(x); break;                     // 2p - x in the lazy range
    case 7: r = F::to_mont(x); break;                 // field.hpp:224-232
    case 8: r = F::from_mont(x); break;               // field.hpp:233-236
    case 9: r = F::invert(x); break;                  // field.hpp:345-348
    case 10: r = F::sub_lazy(x, y); break;            // x - y + 2p, no correction: the butterfly difference of the NTT
    case 11: r = F::mul_full(x, y); break;            // __mul: canonical
    case 12: r = F::mul_const(x, F::from_mont(y), F::const_quotient(y)); break; // raw: must lie in [0, 2p)
    case 13: r = F::invert_binary(x); break;
    case 14: r = F::mul2(x, y, F::add(x, y), F::sub(x, y)); break; // x y + (x + y)(x - y) under one reduction (the mixed addition's y3)          // the same inverse by the binary extended Euclid (the pair-sum rounds of the MSM)
    default: r = F::zero();
    }
    store_fe(out + i, r);
}

// points in and out in the reference's affine image (64 bytes, infinity flag = bit 63 of y limb 3)
__global__ void g1_selftest_kernel(int op, const fe* p, const fe* q, fe* out, size_t count)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const affine_pt P = load_affine(p + 2 * i);
    const affine_pt Q = load_affine(q + 2 * i);
    const xyzz_pt p0 = G1::affine_is_infinity(P) ? G1::infinity() : G1::from_affine(P);
    const xyzz_pt q0 = G1::affine_is_infinity(Q) ? G1::infinity() : G1::from_affine(Q);
    xyzz_pt r;
    switch (op)
    {
    case 0: r = G1::affine_is_infinity(Q) ? p0 : G1::madd(p0, Q); break;                   // group.hpp:211-301 mixed_add (P + P, P - P, inf + Q inside)
    case 1: r = G1::add(G1::dbl(p0), G1::dbl(q0)); break;                                  // :361-448 add on non-trivial z: 2P + 2Q
    case 2: r = G1::dbl(G1::dbl(p0)); break;                                               // :153-209 dbl twice: 4P
    case 3: r = G1::madd(G1::madd(G1::madd(G1::infinity(), P), Q), P); break;              // accumulate from infinity: 2P + Q
    case 4: r = G1::add(p0, q0); break;                                                    // add with z = 1 on both sides, P = Q and P = -Q included
    case 5: r = G1::from_affine(G1::endo_table_entry(P)); break;                           // scalar_multiplication.cpp:131-140 odd entry
    case 6: r = G1::affine_is_infinity(P) ? G1::infinity() : G1::dbl_affine(P); break;     // 2P from the affine image
    default: r = G1::infinity();
    }
    store_affine(out + 2 * i, G1::to_affine(r));
}
} // namespace

size_t selftest_launch_count() { return g_selftest_launches; }

// d_a, d_b, d_out: count field elements each (d_b may be null for unary ops)
int field_selftest_device(int field, int op, const void* d_a, const void* d_b, void* d_out, size_t count, cudaStream_t st)
{
    if (count == 0) return 0;
    const dim3 grid((unsigned)((count + 127) / 128)), block(128);
    if (field == 0) BBG_LAUNCH_NOSYNC(field_selftest_kernel<Fq>, grid, block, st, op, (const fe*)d_a, (const fe*)d_b, (fe*)d_out, count);
    else BBG_LAUNCH_NOSYNC(field_selftest_kernel<Fr>, grid, block, st, op, (const fe*)d_a, (const fe*)d_b, (fe*)d_out, count);
    ++g_selftest_launches;
    return bbg_rt::last_error();
}

int g1_selftest_device(int op, const void* d_p, const void* d_q, void* d_out, size_t count, cudaStream_t st)
{
    if (count == 0) return 0;
    BBG_LAUNCH_NOSYNC(g1_selftest_kernel, dim3((unsigned)((count + 63) / 64)), dim3(64), st, op, (const fe*)d_p, (const fe*)d_q, (fe*)d_out, count);
    ++g_selftest_launches;
    return bbg_rt::last_error();
}
} // namespace bbg
