// HBM-resident PLONK prover rounds for sm_100a (device-resident polynomials, the prover's
// element-wise loops, and its two serial recurrences as parallel scans).
//
// Replaces the data-parallel body of waffle::Prover::construct_proof for circuits built from the reference's four widgets
//   waffle/proof_system/prover/prover.cpp:65-690, widgets/{arithmetic,bool,mimc,sequential}_widget.cpp, permutation.hpp:13-88,
//   polynomials/polynomial_arithmetic.cpp:337-373 (evaluate), :478-560 (divide_by_pseudo_vanishing_polynomial),
//   :562-591 (compute_kate_opening_coefficients), fields/field.hpp:503-522 (batch_invert)
// The Fiat-Shamir transcript (keccak, challenge.hpp) and the handful of scalar formulas (linearizer.hpp,
// get_lagrange_evaluations) stay in the host shim (barretenberg_b200/shim/prover_gpu.cpp), which runs the reference's
// own header code for them; everything that touches n field elements happens here, between rounds nothing but
// commitments (12 limbs each) and evaluations cross PCIe.
//
// Value contract: the proof is a deterministic function of (witness, circuit, SRS); every quantity below is the same
// field element / group element the reference computes, so the proof is identical limb for limb
// (tests/test_gpu_prover_dropin.py, tests/test_emul_prover.py).  Intermediate polynomials live in [0, 2p) ("coarse");
// everything that leaves the device is canonical.
//
// Streams: the work stream the C ABI hands in, a second stream for the 4n coset transforms that only need earlier rounds'
// results, two streams for the pipelined wire commitments of round 1, and an upload stream fed by a helper thread.
// Circuit constants (permutation / selector polynomials in all the forms the rounds need) stay on the device between
// proofs behind a fingerprint of the host buffers (proving-key cache).
//
// Layout in HBM for a circuit of n = 2^k gates (field element = 32 B; 86 n elements, n = 2^20 -> 2.9 GB in total):
//   w_lag[3][n]  witness, Lagrange form          w_coef[3][n]   coefficient form       w4[3][4n]  coset evaluations
//   sigma_lag[3][n], sigma[3][n]  permutation polys, Lagrange / coefficient form       s4[3][4n]  their coset evaluations
//   z[n], z4[4n] grand product                   q[11][n], q2[9][2n], q4[2][4n]  selectors (all four widget kinds)   l1[2n]
//   quot_large[4n], quot_mid[2n], r[n], tmp[2][n] (scan inputs / opening polynomials)
#include "bbg_internal.h"
#include "bbg_host_g1.h"
#include "bbg_hostcopy.h"
#include "bbg_plonk.h"

#include <chrono>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#ifndef BBG_EMULATE
#include <atomic>
#include <thread>
#endif

namespace bbg
{
namespace plonkk
{
// fr.hpp:66-69 multiplicative_generator (5) and :76-79 alternate_multiplicative_generator, Montgomery form
BBG_HD fe gen_k1()
{
    return fe{ { 0x9FFFFFE6u, 0x1B0D0EF9u, 0xA32A913Fu, 0xEABA68A3u, 0xD8DD0689u, 0x47D8EB76u, 0x20F5BBC3u, 0x15D00855u } };
}
BBG_HD fe gen_k2()
{
    return fe{ { 0x4FFFFFDBu, 0x3057819Eu, 0x6832BB01u, 0x307F6D86u, 0x484E3A89u, 0x5C65EC9Fu, 0x73D3D9F8u, 0x0180A965u } };
}

// w^e for e < domain size as lo[e & mask] * hi[e >> lo_log] (hi == nullptr for domains that fit the low table)
struct PowTable
{
    const fe* lo;
    const fe* hi;
    int lo_log;
};
BBG_D fe root_pow(const PowTable& t, uint32_t e)
{
    fe a = load_fe(t.lo + (e & ((1u << t.lo_log) - 1)));
    if (t.hi != nullptr) a = Fr::mul(a, load_fe(t.hi + (e >> t.lo_log)));
    return a;
}

// out[i] = base^i, canonical
__global__ void powers_kernel(fe* out, fe base, unsigned count)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    store_fe(out + i, Fr::reduce(Fr::pow_u64(base, i)));
}

// permutation.hpp:13-88: sigma_k[i] = K_col * w_n^idx for mapping entry (col << 30 | idx).  (The reference looks w^idx up
// as +-roots[idx mod n/2]; w^(n/2) = -1 makes that the plain power.)
__global__ void sigma_from_mapping_kernel(fe* out, const uint32_t* map, PowTable small, unsigned n, unsigned total)
{
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x)
    {
        const uint32_t m = map[idx];
        const uint32_t raw = (m & ((1u << 29) - 1u)) & (n - 1);
        fe v = root_pow(small, raw);
        switch ((m >> 30) & 3u)
        {
        case 2u: v = Fr::mul(v, gen_k2()); break;
        case 1u: v = Fr::mul(v, gen_k1()); break;
        default: break;
        }
        store_fe(out + idx, v);
    }
}

// prover.cpp:146-186: the six accumulator inputs of the grand product, already multiplied three by three:
//   num[i] = (w_l + beta w^i + gamma)(w_r + beta k1 w^i + gamma)(w_o + beta k2 w^i + gamma)
//   den[i] = (w_l + beta sigma_1 + gamma)(w_r + beta sigma_2 + gamma)(w_o + beta sigma_3 + gamma)
__global__ void z_terms_kernel(fe* num, fe* den, const fe* w_lag, const fe* sigma, PowTable small, fe beta, fe gamma, unsigned n)
{
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const fe wl = Fr::add(load_fe(w_lag + i), gamma);
        const fe wr = Fr::add(load_fe(w_lag + n + i), gamma);
        const fe wo = Fr::add(load_fe(w_lag + 2 * (size_t)n + i), gamma);
        const fe rb = Fr::mul(root_pow(small, i), beta);
        fe t = Fr::add(wl, rb);
        t = Fr::mul(t, Fr::add(wr, Fr::mul(rb, gen_k1())));
        t = Fr::mul(t, Fr::add(wo, Fr::mul(rb, gen_k2())));
        store_fe(num + i, t);
        fe d = Fr::add(wl, Fr::mul(load_fe(sigma + i), beta));
        d = Fr::mul(d, Fr::add(wr, Fr::mul(load_fe(sigma + n + i), beta)));
        d = Fr::mul(d, Fr::add(wo, Fr::mul(load_fe(sigma + 2 * (size_t)n + i), beta)));
        store_fe(den + i, d);
    }
}

// ---- exclusive prefix product (prover.cpp:188-201 "6 non-parallelizable processes") ------------------------------
// level 0: every thread owns `run` consecutive elements; level 1: one CTA scans the per-run products.
constexpr int SCAN_THREADS = 1024;
__global__ void prod_reduce_kernel(const fe* in, fe* aggs, unsigned n, unsigned run, size_t in_stride, unsigned aggs_stride)
{
    const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned first = k * run;
    if (first >= n) return;
    const fe* src = in + blockIdx.y * in_stride;
    const unsigned count = n - first < run ? n - first : run;
    fe acc = load_fe(src + first);
    for (unsigned j = 1; j < count; ++j) acc = Fr::mul(acc, load_fe(src + first + j));
    store_fe(aggs + (size_t)blockIdx.y * aggs_stride + k, acc);
}
// in place: aggs[k] <- prod_{m<k} aggs[m]
__global__ void __launch_bounds__(SCAN_THREADS) prod_spine_kernel(fe* aggs_all, unsigned count, unsigned aggs_stride)
{
    __shared__ fe sh[SCAN_THREADS];
    fe* aggs = aggs_all + (size_t)blockIdx.x * aggs_stride;
    const unsigned t = threadIdx.x, T = blockDim.x;
    const unsigned per = (count + T - 1) / T;
    const unsigned lo = t * per, hi = (lo + per < count) ? lo + per : count;
    fe a = Fr::one();
    for (unsigned k = lo; k < hi; ++k) a = Fr::mul(a, load_fe(aggs + k));
    sh[t] = a;
    __syncthreads();
    for (unsigned d = 1; d < T; d <<= 1)
    {
        fe v = sh[t];
        if (t >= d) v = Fr::mul(v, sh[t - d]);
        __syncthreads();
        sh[t] = v;
        __syncthreads();
    }
    fe carry = t == 0 ? Fr::one() : sh[t - 1];
    for (unsigned k = lo; k < hi; ++k)
    {
        const fe c = load_fe(aggs + k);
        store_fe(aggs + k, carry);
        carry = Fr::mul(carry, c);
    }
}
// prover.cpp:203-222 fused with the final scan level: z[i] = P_num[i] / P_den[i], P[i] = prod_{k<i} (exclusive).  The
// reference inverts all n denominators with one serial Montgomery-trick sweep (field.hpp:503-522); here every thread
// inverts its own run (one Fermat inversion per run).
constexpr int ZRUN = 32;
__global__ void __launch_bounds__(64) z_finish_kernel(fe* z, const fe* num, const fe* den, const fe* num_carry, const fe* den_carry, unsigned n)
{
    const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned first = k * ZRUN;
    if (first >= n) return;
    const int count = (int)(n - first < (unsigned)ZRUN ? n - first : ZRUN);
    fe pd[ZRUN], prefix[ZRUN];
    fe acc = Fr::one();
    fe dcarry = load_fe(den_carry + k);
    for (int i = 0; i < count; ++i)
    {
        pd[i] = dcarry; // exclusive prefix product of the denominators at first + i
        prefix[i] = acc;
        acc = Fr::mul(acc, dcarry);
        dcarry = Fr::mul(dcarry, load_fe(den + first + i));
    }
    fe inv = Fr::invert(acc);
    // numerators: exclusive prefix products at first + i, needed from the top down
    fe pn[ZRUN];
    fe ncarry = load_fe(num_carry + k);
    for (int i = 0; i < count; ++i)
    {
        pn[i] = ncarry;
        ncarry = Fr::mul(ncarry, load_fe(num + first + i));
    }
    for (int i = count - 1; i >= 0; --i)
    {
        const fe d_inv = Fr::mul(inv, prefix[i]);
        inv = Fr::mul(inv, pd[i]);
        store_fe(z + first + i, Fr::mul(pn[i], d_inv));
    }
}

// dst[b][i] = i < n_src ? src[b][i] : 0
__global__ void pad_copy_kernel(fe* dst, const fe* src, unsigned n_src, unsigned n_dst, size_t src_stride, size_t dst_stride)
{
    const fe* s = src + blockIdx.y * src_stride;
    fe* d = dst + blockIdx.y * dst_stride;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_dst; i += gridDim.x * blockDim.x)
        store_fe(d + i, i < n_src ? load_fe(s + i) : Fr::zero());
}
struct QuotientConsts
{
    fe g;          // coset generator
    fe g_beta;     // g * beta
    fe beta, gamma, alpha, alpha_sqr, alpha_cube;
    fe one;          // Montgomery one
    fe z_scale;      // alpha for the passes that have to scale Z's coset evaluations themselves (they are kept unscaled)
    fe neg_root_inv; // -w_n^-1 = -w_n^(n-1)  (divide_by_pseudo_vanishing_polynomial "numerator_constant")
    fe vinv[4];      // 1 / ((g w_S^j)^n - 1), j < S = 2 (mid domain) or 4 (large domain)
};

// prover.cpp:279-285 (permutation term), :296-323 (identity term), polynomial_arithmetic.cpp:478-560 (the division by
// Z_H*(X) on the large domain), one pass:
//   q[i] = ( (w_l + b x + c)(w_r + b k1 x + c)(w_o + b k2 x + c) aZ(x)  -  s1 s2 s3 aZ(x w) ) (x - w^(n-1)) / (x^n - 1),
//   x = g w_4n^i,  s_k = w_k + b sigma_k + c.
// The reference transforms b sigma_k(X) + w_k(X) + c to the 4n coset every proof (:246-273).  The transform is linear and
// sigma_k's coset evaluations are a circuit constant (sigma4, part of the proving-key cache), so s_k is formed here from
// sigma4 and the wires' evaluations: three products per point instead of three 4n-point transforms per proof.
template <bool DIVIDE>
__global__ void quotient_large_kernel(fe* q, const fe* sigma4, const fe* w4, const fe* z4, PowTable large, QuotientConsts c, unsigned n4)
{
    const unsigned mask = n4 - 1;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x)
    {
        const fe root = root_pow(large, i);
        const fe bx = Fr::mul(root, c.g_beta);
        const fe wl = Fr::add(load_fe(w4 + i), c.gamma);
        const fe wr = Fr::add(load_fe(w4 + (size_t)n4 + i), c.gamma);
        const fe wo = Fr::add(load_fe(w4 + 2 * (size_t)n4 + i), c.gamma);
        fe id = Fr::add(wl, bx);
        id = Fr::mul(id, Fr::add(wr, Fr::mul(bx, gen_k1())));
        id = Fr::mul(id, Fr::add(wo, Fr::mul(bx, gen_k2())));
        id = Fr::mul(id, load_fe(z4 + i));
        fe pm = Fr::add(wl, Fr::mul(load_fe(sigma4 + i), c.beta));
        pm = Fr::mul(pm, Fr::add(wr, Fr::mul(load_fe(sigma4 + (size_t)n4 + i), c.beta)));
        pm = Fr::mul(pm, Fr::add(wo, Fr::mul(load_fe(sigma4 + 2 * (size_t)n4 + i), c.beta)));
        pm = Fr::mul(pm, load_fe(z4 + ((i + 4) & mask)));
        fe v = Fr::sub(id, pm);
        if (DIVIDE)
        {
            v = Fr::mul(v, c.vinv[i & 3]); // (carries alpha: Z's coset evaluations are unscaled)
            v = Fr::mul(v, Fr::add(Fr::mul(root, c.g), c.neg_root_inv));
        }
        else
        {
            v = Fr::mul(v, c.z_scale);
        }
        store_fe(q + i, v);
    }
}

// prover.cpp:325-391 (the two L_1 boundary terms), arithmetic_widget.cpp:60-97 (gate identity) and the division by
// Z_H*(X) on the mid domain, one pass over i < 2n (aZ = alpha Z: z4 was transformed with the constant alpha):
//   q[i] = ( (aZ(x w) - a) a L1[i+4] + (aZ(x) - a) a^2 L1[i] + qm wl wr + ql wl + qr wr + qo wo + qc ) (x - w^(n-1)) / (x^n - 1)
__global__ void quotient_mid_kernel(fe* q, const fe* z4, const fe* l1, const fe* w4, const fe* q2, PowTable mid, QuotientConsts c, fe alpha_base,
                                    unsigned n2)
{
    const unsigned n4 = 2 * n2, mask2 = n2 - 1, mask4 = n4 - 1;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x)
    {
        // z4 holds Z itself (the reference transforms alpha Z, :275): (aZ - a) a = a^2 (Z - 1), (aZ - a) a^2 = a^3 (Z - 1)
        fe t6 = Fr::sub(load_fe(z4 + ((2 * i + 4) & mask4)), c.one);
        t6 = Fr::mul(Fr::mul(t6, c.alpha_sqr), load_fe(l1 + ((i + 4) & mask2)));
        fe t4 = Fr::sub(load_fe(z4 + 2 * i), c.one);
        t4 = Fr::mul(Fr::mul(t4, c.alpha_cube), load_fe(l1 + i));
        const fe wl = load_fe(w4 + 2 * i), wr = load_fe(w4 + (size_t)n4 + 2 * i), wo = load_fe(w4 + 2 * (size_t)n4 + 2 * i);
        fe a = Fr::mul(Fr::mul(wl, load_fe(q2 + i)), wr);
        a = Fr::add(a, Fr::mul(wl, load_fe(q2 + (size_t)n2 + i)));
        fe b = Fr::mul(wr, load_fe(q2 + 2 * (size_t)n2 + i));
        b = Fr::add(b, Fr::mul(wo, load_fe(q2 + 3 * (size_t)n2 + i)));
        a = Fr::add(Fr::add(a, b), load_fe(q2 + 4 * (size_t)n2 + i));
        a = Fr::mul(a, alpha_base); // the selectors' coset evaluations are cached unscaled
        fe v = Fr::add(Fr::add(t4, t6), a);
        v = Fr::mul(v, c.vinv[i & 1]);
        const fe x = Fr::mul(root_pow(mid, i), c.g);
        v = Fr::mul(v, Fr::add(x, c.neg_root_inv));
        store_fe(q + i, v);
    }
}

// ---- widget mixes other than the standard composer's: the same sums, one widget at a time -------------------------
// the two L_1 boundary terms only (prover.cpp:325-391), undivided
__global__ void quotient_mid_base_kernel(fe* q, const fe* z4, const fe* l1, QuotientConsts c, unsigned n2)
{
    const unsigned n4 = 2 * n2, mask2 = n2 - 1, mask4 = n4 - 1;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x)
    {
        // z4 holds Z itself (the reference transforms alpha Z, :275): (aZ - a) a = a^2 (Z - 1), (aZ - a) a^2 = a^3 (Z - 1)
        fe t6 = Fr::sub(load_fe(z4 + ((2 * i + 4) & mask4)), c.one);
        t6 = Fr::mul(Fr::mul(t6, c.alpha_sqr), load_fe(l1 + ((i + 4) & mask2)));
        fe t4 = Fr::sub(load_fe(z4 + 2 * i), c.one);
        t4 = Fr::mul(Fr::mul(t4, c.alpha_cube), load_fe(l1 + i));
        store_fe(q + i, Fr::add(t4, t6));
    }
}
// arithmetic_widget.cpp:80-97
__global__ void arith_mid_add_kernel(fe* q, const fe* w4, const fe* q2, fe scale, unsigned n2)
{
    const unsigned n4 = 2 * n2;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x)
    {
        const fe wl = load_fe(w4 + 2 * i), wr = load_fe(w4 + (size_t)n4 + 2 * i), wo = load_fe(w4 + 2 * (size_t)n4 + 2 * i);
        fe a = Fr::mul(Fr::mul(wl, load_fe(q2 + i)), wr);
        a = Fr::add(a, Fr::mul(wl, load_fe(q2 + (size_t)n2 + i)));
        fe b = Fr::mul(wr, load_fe(q2 + 2 * (size_t)n2 + i));
        b = Fr::add(b, Fr::mul(wo, load_fe(q2 + 3 * (size_t)n2 + i)));
        a = Fr::add(Fr::add(a, b), load_fe(q2 + 4 * (size_t)n2 + i));
        store_fe(q + i, Fr::add(load_fe(q + i), Fr::mul(a, scale)));
    }
}
// bool_widget.cpp:76-97: (w^2 - w) q_b alpha_base alpha^k for the three wires
struct Scale3
{
    fe s[3];
};
__global__ void bool_mid_add_kernel(fe* q, const fe* w4, const fe* q2, Scale3 scale, unsigned n2)
{
    const unsigned n4 = 2 * n2;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x)
    {
        fe acc = load_fe(q + i);
#pragma unroll
        for (int k = 0; k < 3; ++k)
        {
            const fe w = load_fe(w4 + (size_t)k * n4 + 2 * i);
            acc = Fr::add(acc, Fr::mul(Fr::mul(Fr::sub(Fr::sqr(w), w), load_fe(q2 + (size_t)k * n2 + i)), scale.s[k]));
        }
        store_fe(q + i, acc);
    }
}
// sequential_widget.cpp:56-59: w_o(X w) q_o_next
__global__ void seq_mid_add_kernel(fe* q, const fe* wo4, const fe* q2, fe scale, unsigned n2)
{
    const unsigned mask4 = 2 * n2 - 1;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x)
        store_fe(q + i, Fr::add(load_fe(q + i), Fr::mul(Fr::mul(load_fe(wo4 + ((2 * i + 4) & mask4)), load_fe(q2 + i)), scale)));
}
// mimc_widget.cpp:68-86 on the large domain: T0 = w_o + w_l + q_c;  ((T0^3 - w_r) + (w_r^2 T0 - w_o(X w)) alpha) q_mimc
__global__ void mimc_large_add_kernel(fe* q, const fe* w4, const fe* q4, fe alpha, fe scale, unsigned n4)
{
    const unsigned mask = n4 - 1;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x)
    {
        const fe wl = load_fe(w4 + i), wr = load_fe(w4 + (size_t)n4 + i), wo = load_fe(w4 + 2 * (size_t)n4 + i);
        const fe t0 = Fr::add(Fr::add(wo, wl), load_fe(q4 + (size_t)n4 + i));
        fe t1 = Fr::sub(Fr::mul(Fr::sqr(t0), t0), wr);
        fe t2 = Fr::sub(Fr::mul(Fr::sqr(wr), t0), load_fe(w4 + 2 * (size_t)n4 + ((i + 4) & mask)));
        t1 = Fr::add(t1, Fr::mul(t2, alpha));
        store_fe(q + i, Fr::add(load_fe(q + i), Fr::mul(Fr::mul(t1, load_fe(q4 + i)), scale)));
    }
}
// polynomial_arithmetic.cpp:478-560 on its own: v[i] *= (g w^i - w_n^(n-1)) / ((g w^i)^n - 1)
__global__ void divide_vanishing_kernel(fe* v, PowTable table, QuotientConsts c, unsigned s_mask, unsigned count, int canonical = 0)
{
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
    {
        fe x = Fr::mul(load_fe(v + i), c.vinv[i & s_mask]);
        x = Fr::mul(x, Fr::add(Fr::mul(root_pow(table, i), c.g), c.neg_root_inv));
        store_fe(v + i, canonical ? Fr::reduce(x) : x);
    }
}
// dst[i] (+)= sum_j c_j src_j[i]
constexpr int MAX_TERMS = 12;
struct LinTerms
{
    const fe* src[MAX_TERMS];
    fe c[MAX_TERMS];
    int count;
};
template <bool ACCUMULATE> __global__ void axpy_terms_kernel(fe* dst, LinTerms t, unsigned n)
{
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        fe acc = ACCUMULATE ? load_fe(dst + i) : Fr::zero();
        for (int j = 0; j < t.count; ++j) acc = Fr::add(acc, Fr::mul(load_fe(t.src[j] + i), t.c[j]));
        store_fe(dst + i, acc);
    }
}

// a[i] += b[i]
__global__ void add_into_kernel(fe* a, const fe* b, unsigned count)
{
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
        store_fe(a + i, Fr::add(load_fe(a + i), load_fe(b + i)));
}

// ---- polynomial evaluation (polynomial_arithmetic.cpp:337-373), batched over (polynomial, point) jobs -------------
constexpr int EVAL_THREADS = 256;
constexpr int EVAL_RUN = 16;
constexpr int EVAL_SPAN = EVAL_THREADS * EVAL_RUN;
constexpr int MAX_EVAL_JOBS = 12;
struct EvalJobs
{
    const fe* poly[MAX_EVAL_JOBS];
    unsigned len[MAX_EVAL_JOBS];
    fe point[MAX_EVAL_JOBS];
};
// sum over the CTA of val_t * mult^t  (mult = per-thread step); result valid in thread 0
BBG_D fe block_horner_reduce(fe val, fe mult, fe* sh)
{
    const unsigned t = threadIdx.x, T = blockDim.x;
    for (unsigned d = 1; d < T; d <<= 1)
    {
        sh[t] = val;
        __syncthreads();
        if ((t & (2 * d - 1)) == 0 && t + d < T) val = Fr::add(val, Fr::mul(sh[t + d], mult));
        __syncthreads();
        mult = Fr::sqr(mult);
    }
    return val;
}
// partial[job][block] = sum_{j in block span} f_j z^(j - span start)
__global__ void __launch_bounds__(EVAL_THREADS) eval_partial_kernel(EvalJobs jobs, fe* partial, unsigned partial_stride)
{
    __shared__ fe sh[EVAL_THREADS];
    const int job = blockIdx.y;
    const unsigned len = jobs.len[job];
    if ((size_t)blockIdx.x * EVAL_SPAN >= len) return;
    const fe* f = jobs.poly[job];
    const fe z = jobs.point[job];
    const unsigned first = blockIdx.x * EVAL_SPAN + threadIdx.x * EVAL_RUN;
    fe acc = Fr::zero();
    for (int j = EVAL_RUN - 1; j >= 0; --j)
    {
        acc = Fr::mul(acc, z);
        if (first + j < len) acc = Fr::add(acc, load_fe(f + first + j));
    }
    const fe r = block_horner_reduce(acc, Fr::pow_u64(z, EVAL_RUN), sh);
    if (threadIdx.x == 0) store_fe(partial + (size_t)job * partial_stride + blockIdx.x, r);
}
// out[job] = sum_b partial[job][b] (z^SPAN)^b, canonical
__global__ void __launch_bounds__(SCAN_THREADS) eval_final_kernel(EvalJobs jobs, const fe* partial, unsigned partial_stride, fe* out)
{
    __shared__ fe sh[SCAN_THREADS];
    const int job = blockIdx.x;
    const unsigned len = jobs.len[job];
    const unsigned nb = (len + EVAL_SPAN - 1) / EVAL_SPAN;
    const unsigned t = threadIdx.x, T = blockDim.x;
    const unsigned per = (nb + T - 1) / T;
    const fe step = Fr::pow_u64(jobs.point[job], EVAL_SPAN);
    const fe* p = partial + (size_t)job * partial_stride;
    fe acc = Fr::zero();
    for (int j = (int)per - 1; j >= 0; --j)
    {
        acc = Fr::mul(acc, step);
        const unsigned b = t * per + (unsigned)j;
        if (b < nb) acc = Fr::add(acc, load_fe(p + b));
    }
    const fe r = block_horner_reduce(acc, Fr::pow_u64(step, per), sh);
    if (t == 0) store_fe(out + job, Fr::reduce(r));
}

// prover.cpp:552-590: the two batched opening polynomials before the division
struct OpeningConsts
{
    fe nu[7];           // nu^1 .. nu^7
    fe z_pow_n, z_pow_2n;
};
__global__ void opening_combine_kernel(fe* opening, fe* shifted, const fe* quot, const fe* r, const fe* w_coef, const fe* sigma, const fe* z,
                                       OpeningConsts c, unsigned n)
{
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        fe t8 = Fr::add(Fr::mul(load_fe(quot + (size_t)n + i), c.z_pow_n), Fr::mul(load_fe(quot + 2 * (size_t)n + i), c.z_pow_2n));
        fe t1 = Fr::add(Fr::mul(load_fe(r + i), c.nu[0]), Fr::mul(load_fe(w_coef + i), c.nu[1]));
        fe t3 = Fr::add(Fr::mul(load_fe(w_coef + (size_t)n + i), c.nu[2]), Fr::mul(load_fe(w_coef + 2 * (size_t)n + i), c.nu[3]));
        // (the reference holds beta sigma and multiplies by beta^-1 here, :586; sigma is kept unscaled)
        fe t4 = Fr::add(Fr::mul(load_fe(sigma + i), c.nu[4]), Fr::mul(load_fe(sigma + (size_t)n + i), c.nu[5]));
        fe v = Fr::add(Fr::add(t3, t1), Fr::add(t4, t8));
        store_fe(opening + i, Fr::add(v, load_fe(quot + i)));
        store_fe(shifted + i, Fr::mul(load_fe(z + i), c.nu[6]));
    }
}

// ---- (F(X) - F(z)) / (X - z) (polynomial_arithmetic.cpp:562-591) --------------------------------------------------
// The reference walks the recurrence w_i = (f_i - w_{i-1}) / (-z) upwards, serially.  The same quotient read from
// the top is w_{i-1} = f_i + z w_i (w_{n-1} = 0): a suffix scan with the affine maps x -> f_i + z x.
// blockIdx.y selects the polynomial (stride n) and its point.
struct KatePoints
{
    fe z[2];
    fe z_run[2];  // z^run
    fe z_span[2]; // (z^run)^per  (per = aggregates per spine thread)
};
__global__ void kate_reduce_kernel(const fe* in, fe* aggs, KatePoints pts, unsigned n, unsigned run, unsigned aggs_stride)
{
    const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned first = k * run;
    if (first >= n) return;
    const fe* f = in + (size_t)blockIdx.y * n;
    const fe z = pts.z[blockIdx.y];
    fe acc = Fr::zero();
    for (int j = (int)run - 1; j >= 0; --j)
    {
        acc = Fr::mul(acc, z);
        if (first + j < n) acc = Fr::add(acc, load_fe(f + first + j));
    }
    store_fe(aggs + (size_t)blockIdx.y * aggs_stride + k, acc);
}
// in place: aggs[k] <- sum_{m>k} aggs[m] (z^run)^(m-k-1)
__global__ void __launch_bounds__(SCAN_THREADS) kate_spine_kernel(fe* aggs_all, KatePoints pts, unsigned count, unsigned aggs_stride)
{
    __shared__ fe sh[SCAN_THREADS];
    fe* aggs = aggs_all + (size_t)blockIdx.x * aggs_stride;
    const fe M = pts.z_run[blockIdx.x];
    const unsigned t = threadIdx.x, T = blockDim.x;
    const unsigned per = (count + T - 1) / T;
    const unsigned lo = t * per, hi = (lo + per < count) ? lo + per : count;
    fe a = Fr::zero();
    for (unsigned k = hi; k > lo; --k) a = Fr::add(Fr::mul(a, M), load_fe(aggs + k - 1));
    // a = sum_{k in [lo, hi)} c_k M^(k - lo); ranges that end early (hi < lo + per) only occur above every non-zero aggregate
    fe mult = pts.z_span[blockIdx.x];
    sh[t] = a;
    __syncthreads();
    for (unsigned d = 1; d < T; d <<= 1)
    {
        fe v = sh[t];
        if (t + d < T) v = Fr::add(v, Fr::mul(sh[t + d], mult));
        __syncthreads();
        sh[t] = v;
        __syncthreads();
        mult = Fr::sqr(mult);
    }
    fe carry = (t + 1 < T) ? sh[t + 1] : Fr::zero();
    // a thread whose range is cut short by `count` sits at the top: its carry is zero and stays aligned
    if (lo < count && hi < lo + per) carry = Fr::zero();
    for (unsigned k = hi; k > lo; --k)
    {
        const fe c = load_fe(aggs + k - 1);
        store_fe(aggs + k - 1, carry);
        carry = Fr::add(c, Fr::mul(carry, M));
    }
}
// out of place (in place the loads cannot be hoisted above the stores and every thread walks its run at memory latency)
__global__ void kate_apply_kernel(fe* __restrict__ out, const fe* __restrict__ in, const fe* __restrict__ aggs, KatePoints pts, unsigned n, unsigned run,
                                  unsigned aggs_stride, int canonical = 0)
{
    const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned first = k * run;
    if (first >= n) return;
    const fe* f = in + (size_t)blockIdx.y * n;
    fe* w = out + (size_t)blockIdx.y * n;
    const fe z = pts.z[blockIdx.y];
    fe carry = load_fe(aggs + (size_t)blockIdx.y * aggs_stride + k);
    const unsigned last = (first + run < n) ? first + run : n;
    unsigned j = last;
    for (; j >= first + 4; j -= 4)
    {
        const fe c0 = load_fe(f + j - 1), c1 = load_fe(f + j - 2), c2 = load_fe(f + j - 3), c3 = load_fe(f + j - 4);
        store_fe(w + j - 1, canonical ? Fr::reduce(carry) : carry);
        carry = Fr::add(c0, Fr::mul(carry, z));
        store_fe(w + j - 2, canonical ? Fr::reduce(carry) : carry);
        carry = Fr::add(c1, Fr::mul(carry, z));
        store_fe(w + j - 3, canonical ? Fr::reduce(carry) : carry);
        carry = Fr::add(c2, Fr::mul(carry, z));
        store_fe(w + j - 4, canonical ? Fr::reduce(carry) : carry);
        carry = Fr::add(c3, Fr::mul(carry, z));
    }
    for (; j > first; --j)
    {
        const fe c = load_fe(f + j - 1);
        store_fe(w + j - 1, canonical ? Fr::reduce(carry) : carry);
        carry = Fr::add(c, Fr::mul(carry, z));
    }
}
} // namespace plonkk

// =================================================================================================================
// Host driver
// =================================================================================================================
namespace plonk
{
using namespace plonkk;

namespace
{
size_t g_plonk_launches = 0;
const fe ROOT_2_28 = { { 0x80D13D9Cu, 0x636E7355u, 0x2445FFD6u, 0xA22BF374u, 0x1EB203D8u, 0x56452AC0u, 0x2963F9E7u, 0x1860EF94u } };

fe host_root_of_unity(unsigned log_n) // field.hpp:487-494
{
    fe r = ROOT_2_28;
    for (unsigned i = 28; i > log_n; --i) r = Fr::reduce(Fr::sqr(r));
    return r;
}
fe host_pow(fe base, uint64_t e) { return Fr::reduce(Fr::pow_u64(base, e)); }
fe from_u64(const uint64_t* p)
{
    fe r;
    memcpy(r.v, p, 32);
    return r;
}
unsigned grid_for(size_t count, unsigned block)
{
    size_t g = (count + block - 1) / block;
    const size_t cap = (size_t)bbg_rt::num_sms() * 8;
    if (g > cap) g = cap;
    return g == 0 ? 1u : (unsigned)g;
}
} // namespace

constexpr int MAX_WIDGETS = 4;
constexpr int MAX_SELECTORS = 11; // 5 arithmetic + 3 bool + 2 MiMC + 1 sequential
inline int selectors_of(int kind)
{
    switch (kind)
    {
    case WIDGET_ARITHMETIC: return 5; // q_m, q_l, q_r, q_o, q_c          (arithmetic_widget.hpp:45-49)
    case WIDGET_BOOL: return 3;       // q_bl, q_br, q_bo                 (bool_widget.hpp:45-47)
    case WIDGET_MIMC: return 2;       // q_mimc_selector, q_mimc_coefficient (mimc_widget.hpp:43-44)
    case WIDGET_SEQUENTIAL: return 1; // q_o_next                         (sequential_widget.hpp:46)
    }
    return -1;
}

struct Prover
{
    unsigned log_n = 0;
    size_t n = 0;
    void* arena = nullptr;
    fe *w_lag = nullptr, *w_coef = nullptr, *w4 = nullptr;
    fe *sigma = nullptr, *sigma_lag = nullptr, *s4 = nullptr;
    fe *z = nullptr, *z4 = nullptr;
    fe *q = nullptr, *q2 = nullptr, *q4 = nullptr, *l1 = nullptr;
    // widget list of the circuit (prover.hpp:60), in the prover's order; selectors are stored back to back in q[]
    int num_widgets = 0, num_selectors = 0;
    int widget_kind[MAX_WIDGETS] = {};
    int widget_first_selector[MAX_WIDGETS] = {};
    fe *quot_large = nullptr, *quot_mid = nullptr, *r = nullptr, *tmp = nullptr;
    fe *aggs = nullptr, *eval_partial = nullptr, *eval_out = nullptr;
    uint32_t* map = nullptr;
    fe* pow_mem = nullptr;
    PowTable pow_small{}, pow_mid{}, pow_large{};
    unsigned aggs_stride = 0, partial_stride = 0;
    const void* d_srs = nullptr;
    bool have_witness = false, have_perm = false, have_selectors = false, tables_ready = false;
    bool sigma_ready = false; // sigma_lag[] holds this proof's Lagrange values
    // Proving-key cache: the permutation polynomials (Lagrange + coefficient form) and the selectors (coefficient form +
    // coset evaluations, all unscaled) are circuit constants.  Every proof the helper thread hashes the host buffers it
    // was handed; when the hash equals the one the device copies were built from, the copies and the 13 transforms of
    // rounds 2-3 that only depend on them are skipped.  BBG_PLONK_KEY_CACHE=0 switches this off.
    bool key_cache_enabled = true;
    bool key_valid = false;          // device constants are complete and belong to key_hash
    uint64_t key_hash[2] = { 0, 0 };
    uint64_t pending_hash[2] = { 0, 0 };
    bool constants_cached = false;   // this proof: the helper thread found the constants unchanged
    fe beta{};                       // this proof's beta (sigma is stored unscaled)
    bool l1_ready = false;    // l1[] depends on the circuit size only: computed once
    // this proof's inputs, in the order the rounds need them: w_l, w_r, w_o, the three mappings, the five selectors
    enum { ITEM_WL = 0, ITEM_WR, ITEM_WO, ITEM_MAP, ITEM_SEL, NUM_ITEMS };
    const void* host_src[6 + MAX_SELECTORS] = {};
    bool uploads_started = false;
#ifndef BBG_EMULATE
    // One helper thread copies all of them through its own pinned ring and stream while the rounds already run on
    // what has arrived: uploaded (host side) / ev_item (device side) say how far it is.
    int device = 0;
    cudaStream_t upload_stream = nullptr;
    cudaEvent_t ev_item[NUM_ITEMS] = {};
    cudaEvent_t ev_fence = nullptr;
    // Second work stream: the coset transforms of round 3 that only need earlier rounds' results (wires after round 1,
    // beta sigma + w + gamma after the beta / gamma challenge, Z after its ifft) are queued here as soon as their inputs
    // exist and run in the shadow of the commitments' latency-bound phases.  BBG_PLONK_OVERLAP=0 keeps one stream.
    cudaStream_t side = nullptr;
    cudaStream_t msm_stream[2] = {}; // round 1: the three wire commitments alternate between two streams / MSM workspaces
    cudaEvent_t ev_msm[3] = {};
    cudaEvent_t ev_wire[3] = {}, ev_sigma = nullptr, ev_z = nullptr, ev_side = nullptr;
    bool overlap = true;
    bbg_hostcopy::Ring upload_ring;
    std::thread uploader;
    std::atomic<int> uploaded{ 0 }; // items whose copies have all been queued (event recorded)
    int upload_error = 0;
#endif
};

#ifndef BBG_EMULATE
static int join_uploader(Prover* p)
{
    if (p->uploader.joinable()) p->uploader.join();
    const int e = p->upload_error;
    p->upload_error = 0;
    return e;
}
#endif

static int build_pow_table(Prover* p, unsigned log_size, fe*& cursor, PowTable* out, cudaStream_t st)
{
    const int lo_log = (int)(log_size < 11 ? log_size : 11);
    const fe w = host_root_of_unity(log_size);
    out->lo = cursor;
    out->lo_log = lo_log;
    BBG_LAUNCH_NOSYNC(powers_kernel, dim3(((1u << lo_log) + 127) / 128), dim3(128), st, cursor, w, 1u << lo_log);
    cursor += (size_t)1 << lo_log;
    ++g_plonk_launches;
    if (log_size > (unsigned)lo_log)
    {
        const unsigned hi_count = 1u << (log_size - lo_log);
        out->hi = cursor;
        BBG_LAUNCH_NOSYNC(powers_kernel, dim3((hi_count + 127) / 128), dim3(128), st, cursor, host_pow(w, (uint64_t)1 << lo_log), hi_count);
        cursor += hi_count;
        ++g_plonk_launches;
    }
    else out->hi = nullptr;
    (void)p;
    return bbg_rt::last_error();
}

int create(unsigned log_n, Prover** out)
{
    // every domain needs at least 4 points for the Z(X w) index shifts; above 2^23 gates the arena (86 n field elements)
    // and the 32-bit element indices of the 4n passes would need another look
    if (log_n < 2 || log_n > 23) return 1002;
    Prover* p = new Prover();
    p->log_n = log_n;
    p->n = (size_t)1 << log_n;
    const size_t n = p->n;
    const unsigned run_aggs = (unsigned)((n + ZRUN - 1) / ZRUN);
    p->aggs_stride = (run_aggs + 7) & ~7u;
    p->partial_stride = (unsigned)((4 * n + EVAL_SPAN - 1) / EVAL_SPAN + 8);
    size_t pow_elems = 64;
    for (unsigned lg = log_n; lg <= log_n + 2; ++lg) pow_elems += ((size_t)1 << (lg < 11 ? lg : 11)) + (lg > 11 ? (size_t)1 << (lg - 11) : 0);
    // element counts, in the order of the header comment
    const size_t counts[] = { 3 * n, 3 * n, 12 * n, 3 * n, 12 * n, n, 4 * n, MAX_SELECTORS * n, 18 * n, 2 * n, 4 * n, 2 * n, n, 2 * n,
                              2 * (size_t)p->aggs_stride, (size_t)MAX_EVAL_JOBS * p->partial_stride, MAX_EVAL_JOBS, pow_elems, 8 * n, 3 * n };
    size_t total = 0;
    for (size_t c : counts) total += (c + 7) & ~(size_t)7;
    const size_t bytes = total * 32 + 3 * n * 4 + 256; // (2.9 GB at n = 2^20)
    int e = bbg_rt::dev_alloc(&p->arena, bytes);
    if (e != 0)
    {
        delete p;
        return e;
    }
    fe* cur = (fe*)p->arena;
    auto take = [&](size_t c) { fe* r = cur; cur += (c + 7) & ~(size_t)7; return r; };
    p->w_lag = take(counts[0]);
    p->w_coef = take(counts[1]);
    p->w4 = take(counts[2]);
    p->sigma = take(counts[3]);
    p->s4 = take(counts[4]);
    p->z = take(counts[5]);
    p->z4 = take(counts[6]);
    p->q = take(counts[7]);
    p->q2 = take(counts[8]);
    p->l1 = take(counts[9]);
    p->quot_large = take(counts[10]);
    p->quot_mid = take(counts[11]);
    p->r = take(counts[12]);
    p->tmp = take(counts[13]);
    p->aggs = take(counts[14]);
    p->eval_partial = take(counts[15]);
    p->eval_out = take(counts[16]);
    p->pow_mem = take(counts[17]);
    p->q4 = take(counts[18]);
    p->sigma_lag = take(counts[19]);
    {
        const char* e = getenv("BBG_PLONK_KEY_CACHE");
        p->key_cache_enabled = !(e != nullptr && e[0] == '0');
    }
    p->map = (uint32_t*)cur;
#ifndef BBG_EMULATE
    cudaGetDevice(&p->device);
    e = (int)cudaStreamCreateWithFlags(&p->upload_stream, cudaStreamNonBlocking);
    for (int i = 0; i < Prover::NUM_ITEMS && e == 0; ++i) e = (int)cudaEventCreateWithFlags(&p->ev_item[i], cudaEventDisableTiming);
    if (e == 0) e = (int)cudaEventCreateWithFlags(&p->ev_fence, cudaEventDisableTiming);
    if (e == 0) e = (int)cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == 0; ++i) e = (int)cudaStreamCreateWithFlags(&p->msm_stream[i], cudaStreamNonBlocking);
    for (int i = 0; i < 3 && e == 0; ++i) e = (int)cudaEventCreateWithFlags(&p->ev_msm[i], cudaEventDisableTiming);
    for (int i = 0; i < 3 && e == 0; ++i) e = (int)cudaEventCreateWithFlags(&p->ev_wire[i], cudaEventDisableTiming);
    if (e == 0) e = (int)cudaEventCreateWithFlags(&p->ev_sigma, cudaEventDisableTiming);
    if (e == 0) e = (int)cudaEventCreateWithFlags(&p->ev_z, cudaEventDisableTiming);
    if (e == 0) e = (int)cudaEventCreateWithFlags(&p->ev_side, cudaEventDisableTiming);
    {
        const char* ov = getenv("BBG_PLONK_OVERLAP");
        p->overlap = !(ov != nullptr && ov[0] == '0');
    }
    if (e != 0)
    {
        destroy(p);
        return e;
    }
#endif
    *out = p;
    return 0;
}

void destroy(Prover* p)
{
    if (p == nullptr) return;
#ifndef BBG_EMULATE
    join_uploader(p);
    if (p->upload_stream)
    {
        cudaStreamSynchronize(p->upload_stream);
        cudaStreamDestroy(p->upload_stream);
    }
    for (int i = 0; i < Prover::NUM_ITEMS; ++i)
        if (p->ev_item[i]) cudaEventDestroy(p->ev_item[i]);
    if (p->ev_fence) cudaEventDestroy(p->ev_fence);
    if (p->side)
    {
        cudaStreamSynchronize(p->side);
        cudaStreamDestroy(p->side);
    }
    for (int i = 0; i < 2; ++i)
        if (p->msm_stream[i])
        {
            cudaStreamSynchronize(p->msm_stream[i]);
            cudaStreamDestroy(p->msm_stream[i]);
        }
    for (int i = 0; i < 3; ++i)
        if (p->ev_msm[i]) cudaEventDestroy(p->ev_msm[i]);
    for (int i = 0; i < 3; ++i)
        if (p->ev_wire[i]) cudaEventDestroy(p->ev_wire[i]);
    if (p->ev_sigma) cudaEventDestroy(p->ev_sigma);
    if (p->ev_z) cudaEventDestroy(p->ev_z);
    if (p->ev_side) cudaEventDestroy(p->ev_side);
    p->upload_ring.release();
#endif
    if (p->arena) bbg_rt::dev_free(p->arena);
    delete p;
}

static int ensure_tables(Prover* p, cudaStream_t st)
{
    if (p->tables_ready) return 0;
    fe* cursor = p->pow_mem;
    BBG_CHECK(build_pow_table(p, p->log_n, cursor, &p->pow_small, st));
    BBG_CHECK(build_pow_table(p, p->log_n + 1, cursor, &p->pow_mid, st));
    BBG_CHECK(build_pow_table(p, p->log_n + 2, cursor, &p->pow_large, st));
    p->tables_ready = true;
    return 0;
}

// The host buffers handed to the three set_* calls must stay valid and unmodified until bbg_plonk_round_quotient has
// returned (they are the Prover's own polynomials / mapping vectors and the widget's selector polynomials); the copies
// start with bbg_plonk_round_wires.
int set_witness(Prover* p, const uint64_t* w_l, const uint64_t* w_r, const uint64_t* w_o, cudaStream_t)
{
    if (w_l == nullptr || w_r == nullptr || w_o == nullptr) return 1007;
    if (p->uploads_started) return 1007; // a proof is in flight
    p->host_src[0] = w_l;
    p->host_src[1] = w_r;
    p->host_src[2] = w_o;
    p->have_witness = true;
    return 0;
}

int set_permutation(Prover* p, const uint32_t* m1, const uint32_t* m2, const uint32_t* m3, cudaStream_t)
{
    if (m1 == nullptr || m2 == nullptr || m3 == nullptr) return 1007;
    if (p->uploads_started) return 1007;
    p->host_src[3] = m1;
    p->host_src[4] = m2;
    p->host_src[5] = m3;
    p->have_perm = true;
    return 0;
}

int set_widgets(Prover* p, const int* kinds, int count, const uint64_t* const* selectors_lagrange, cudaStream_t)
{
    if (kinds == nullptr || selectors_lagrange == nullptr || count < 1 || count > MAX_WIDGETS) return 1007;
    if (p->uploads_started) return 1007;
    int total = 0;
    for (int w = 0; w < count; ++w)
    {
        const int k = selectors_of(kinds[w]);
        if (k < 0 || total + k > MAX_SELECTORS) return 1007;
        for (int prev = 0; prev < w; ++prev)
            if (kinds[prev] == kinds[w]) return 1007; // each widget kind at most once (what the reference's composers build)
        p->widget_kind[w] = kinds[w];
        p->widget_first_selector[w] = total;
        total += k;
    }
    for (int k = 0; k < total; ++k)
    {
        if (selectors_lagrange[k] == nullptr) return 1007;
        p->host_src[6 + k] = selectors_lagrange[k];
    }
    p->num_widgets = count;
    p->num_selectors = total;
    p->have_selectors = true;
    return 0;
}

// 128-bit fingerprint of the circuit constants (mappings, widget kinds, selectors): four multiply-xorshift lanes per slice,
// slices hashed by a few threads and folded in order.  Memory-bound (~170 MB at n = 2^20, a few ms behind round 1).
static void hash_slice(const uint64_t* w, size_t words, uint64_t out[2])
{
    uint64_t h0 = 0x9E3779B97F4A7C15ULL, h1 = 0xC2B2AE3D27D4EB4FULL, h2 = 0x165667B19E3779F9ULL, h3 = 0x27D4EB2F165667C5ULL;
    size_t i = 0;
    for (; i + 4 <= words; i += 4)
    {
        h0 = (h0 ^ w[i]) * 0xFF51AFD7ED558CCDULL;
        h0 ^= h0 >> 29;
        h1 = (h1 ^ w[i + 1]) * 0xC4CEB9FE1A85EC53ULL;
        h1 ^= h1 >> 31;
        h2 = (h2 ^ w[i + 2]) * 0x9FB21C651E98DF25ULL;
        h2 ^= h2 >> 30;
        h3 = (h3 ^ w[i + 3]) * 0xD6E8FEB86659FD93ULL;
        h3 ^= h3 >> 32;
    }
    for (; i < words; ++i)
    {
        h0 = (h0 ^ w[i]) * 0xFF51AFD7ED558CCDULL;
        h0 ^= h0 >> 29;
    }
    out[0] = (h0 ^ (h1 << 1 | h1 >> 63)) * 0x9E3779B97F4A7C15ULL ^ h2;
    out[1] = (h2 ^ (h3 << 7 | h3 >> 57)) * 0xC2B2AE3D27D4EB4FULL ^ h0 ^ h1;
}
static void fold_hash(uint64_t acc[2], const uint64_t h[2])
{
    acc[0] = (acc[0] ^ h[0]) * 0xFF51AFD7ED558CCDULL;
    acc[0] ^= acc[0] >> 32;
    acc[1] = (acc[1] ^ h[1] ^ acc[0]) * 0xC4CEB9FE1A85EC53ULL;
    acc[1] ^= acc[1] >> 29;
}
static void hash_constants(const Prover* p, uint64_t out[2])
{
    struct Part
    {
        const uint64_t* w;
        size_t words;
    };
    std::vector<Part> parts;
    const size_t slice_words = (size_t)1 << 20; // 8 MiB
    auto add = [&](const void* ptr, size_t bytes) {
        const uint64_t* w = (const uint64_t*)ptr;
        size_t words = bytes / 8;
        while (words > 0)
        {
            const size_t take = words < slice_words ? words : slice_words;
            parts.push_back({ w, take });
            w += take;
            words -= take;
        }
    };
    for (int k = 0; k < 3; ++k) add(p->host_src[3 + k], p->n * 4);
    for (int k = 0; k < p->num_selectors; ++k) add(p->host_src[6 + k], p->n * 32);
    std::vector<uint64_t> h(2 * parts.size());
#ifndef BBG_EMULATE
    const unsigned hc = std::thread::hardware_concurrency();
    const int threads = hc >= 16 ? 8 : (hc >= 4 ? 4 : 1);
    std::atomic<size_t> next{ 0 };
    auto work = [&]() {
        for (;;)
        {
            const size_t i = next.fetch_add(1);
            if (i >= parts.size()) break;
            hash_slice(parts[i].w, parts[i].words, &h[2 * i]);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
#else
    for (size_t i = 0; i < parts.size(); ++i) hash_slice(parts[i].w, parts[i].words, &h[2 * i]);
#endif
    out[0] = 0x243F6A8885A308D3ULL ^ (uint64_t)p->n;
    out[1] = 0x13198A2E03707344ULL ^ (uint64_t)p->num_selectors;
    for (int w = 0; w < p->num_widgets; ++w)
    {
        const uint64_t kind[2] = { (uint64_t)p->widget_kind[w] + 1, (uint64_t)w };
        fold_hash(out, kind);
    }
    for (size_t i = 0; i < parts.size(); ++i) fold_hash(out, &h[2 * i]);
}

// queue every input copy (helper thread; synchronous in the emulation build)
static int start_uploads(Prover* p, cudaStream_t st)
{
    const size_t n = p->n;
    p->sigma_ready = false;
    p->uploads_started = true;
#ifndef BBG_EMULATE
    BBG_CHECK(join_uploader(p));
    p->uploaded.store(0);
    // kernels of an earlier proof may still read these buffers: the copies wait for the work stream
    BBG_CHECK(cudaEventRecord(p->ev_fence, st));
    BBG_CHECK(cudaStreamWaitEvent(p->upload_stream, p->ev_fence, 0));
    p->uploader = std::thread([p, n]() {
        cudaSetDevice(p->device);
        int e = 0;
        auto copy = [&](void* dst, const void* src, size_t bytes) {
            if (e == 0) e = bbg_hostcopy::h2d_ring(p->upload_ring, dst, src, bytes, p->upload_stream);
        };
        auto done = [&](int item) {
            if (e == 0) e = (int)cudaEventRecord(p->ev_item[item], p->upload_stream);
            p->uploaded.store(item + 1);
        };
        for (int k = 0; k < 3; ++k)
        {
            copy(p->w_lag + (size_t)k * n, p->host_src[k], n * 32);
            done(Prover::ITEM_WL + k);
        }
        // circuit constants: unchanged since the device copies were built?
        bool cached = false;
        if (p->key_cache_enabled)
        {
            hash_constants(p, p->pending_hash);
            cached = p->key_valid && p->pending_hash[0] == p->key_hash[0] && p->pending_hash[1] == p->key_hash[1];
        }
        p->constants_cached = cached;
        if (!cached) p->key_valid = false; // the device constants are about to be overwritten
        if (!cached)
            for (int k = 0; k < 3; ++k) copy(p->map + (size_t)k * n, p->host_src[3 + k], n * 4);
        done(Prover::ITEM_MAP);
        if (!cached)
            for (int k = 0; k < p->num_selectors; ++k) copy(p->q + (size_t)k * n, p->host_src[6 + k], n * 32);
        done(Prover::ITEM_SEL);
        p->upload_error = e;
    });
#else
    for (int k = 0; k < 3; ++k) BBG_CHECK(bbg_hostcopy::h2d(p->w_lag + (size_t)k * n, p->host_src[k], n * 32, st));
    bool cached = false;
    if (p->key_cache_enabled)
    {
        hash_constants(p, p->pending_hash);
        cached = p->key_valid && p->pending_hash[0] == p->key_hash[0] && p->pending_hash[1] == p->key_hash[1];
    }
    p->constants_cached = cached;
    if (!cached)
    {
        p->key_valid = false;
        for (int k = 0; k < 3; ++k) BBG_CHECK(bbg_hostcopy::h2d(p->map + (size_t)k * n, p->host_src[3 + k], n * 4, st));
        for (int k = 0; k < p->num_selectors; ++k) BBG_CHECK(bbg_hostcopy::h2d(p->q + (size_t)k * n, p->host_src[6 + k], n * 32, st));
    }
#endif
    return 0;
}
// make the work stream wait for input `item`
static int wait_item(Prover* p, int item, cudaStream_t st)
{
#ifndef BBG_EMULATE
    while (p->uploaded.load() <= item) std::this_thread::yield();
    if (item == Prover::NUM_ITEMS - 1) BBG_CHECK(join_uploader(p));
    BBG_CHECK(cudaStreamWaitEvent(st, p->ev_item[item], 0));
#else
    (void)p;
    (void)item;
    (void)st;
#endif
    return 0;
}

int set_srs(Prover* p, const void* d_table)
{
    p->d_srs = d_table;
    return 0;
}

// commitments to `count` polynomials of n coefficients, `stride` elements apart, in one batched MSM pipeline
static int commit(Prover* p, const fe* d_scalars, size_t stride, int count, uint64_t* out_xyz, cudaStream_t st)
{
    hostg1::hxyzz r[4];
    const void* ptrs[4];
    if (count < 1 || count > 4) return 1007;
    for (int k = 0; k < count; ++k) ptrs[k] = d_scalars + (size_t)k * stride;
    BBG_CHECK(msm_device_batched(ptrs, (size_t)count, p->d_srs, p->n, r, st));
    for (int k = 0; k < count; ++k) hostg1::to_normalized_jacobian(r[k], out_xyz + 12 * k);
    return 0;
}

// prover.cpp:126-135 compute_wire_coefficients + :65-89 compute_wire_commitments (and, off the critical path of the
// transcript, permutation.hpp's sigma polynomials)
// Side-stream plumbing (no-ops in the emulation build and with BBG_PLONK_OVERLAP=0: everything stays on `st`).
static cudaStream_t side_stream(Prover* p, cudaStream_t st)
{
#ifndef BBG_EMULATE
    if (p->overlap) return p->side;
#endif
    (void)p;
    return st;
}
// `waiter` proceeds only after everything queued on `signaller` so far
static int order_after(Prover* p, cudaStream_t waiter, cudaStream_t signaller, int which)
{
#ifndef BBG_EMULATE
    if (waiter == signaller) return 0;
    cudaEvent_t ev = which < 3 ? p->ev_wire[which] : (which == 3 ? p->ev_sigma : (which == 4 ? p->ev_z : p->ev_side));
    BBG_CHECK(cudaEventRecord(ev, signaller));
    BBG_CHECK(cudaStreamWaitEvent(waiter, ev, 0));
#else
    (void)p;
    (void)waiter;
    (void)signaller;
    (void)which;
#endif
    return 0;
}
// a transform queued on the side stream uses the NTT driver's second scratch buffer
struct SideScratch
{
    bool on;
    SideScratch(cudaStream_t side, cudaStream_t st) : on(side != st) { if (on) ntt_use_side_scratch(true); }
    ~SideScratch() { if (on) ntt_use_side_scratch(false); }
};

// BBG_PLONK_TRACE=1: host-side timeline of the rounds on stderr (development aid)
struct Trace
{
    bool on;
    std::chrono::steady_clock::time_point t0;
    Trace() : on(getenv("BBG_PLONK_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void mark(const char* what, cudaStream_t st)
    {
        if (!on) return;
        bbg_rt::sync(st);
        fprintf(stderr, "  [plonk trace] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
};

int round_wires(Prover* p, uint64_t* out_xyz /* 3 x 12 */, cudaStream_t st)
{
    Trace tr;
    if (!p->have_witness || !p->have_perm || !p->have_selectors || p->d_srs == nullptr) return 1007;
    BBG_CHECK(ensure_tables(p, st));
    const size_t n = p->n;
    BBG_CHECK(start_uploads(p, st));
    int tickets[3] = { -1, -1, -1 };
    // wire by wire: the commitment to w_l is computed while w_r and w_o are still on the PCIe bus
    for (int k = 0; k < 3; ++k)
    {
        BBG_CHECK(wait_item(p, Prover::ITEM_WL + k, st));
        tr.mark("wire uploaded", st);
        BBG_CHECK(bbg_rt::d2d(p->w_coef + (size_t)k * n, p->w_lag + (size_t)k * n, n * 32, st));
        BBG_CHECK(ntt_device(p->w_coef + (size_t)k * n, n, 1, p->log_n, OP_IFFT, nullptr, st));
        tr.mark("wire ifft", st);
        {
            // the wire on the 4n coset (prover.cpp:400-414) only needs its coefficients: queued now, used in round 3
            const size_t n4 = 4 * n;
            cudaStream_t sd = side_stream(p, st);
            BBG_CHECK(order_after(p, sd, st, k));
            SideScratch scratch(sd, st);
            BBG_LAUNCH_NOSYNC(pad_copy_kernel, dim3(grid_for(n4, 256), 1), dim3(256), sd, p->w4 + (size_t)k * n4, (const fe*)(p->w_coef + (size_t)k * n),
                              (unsigned)n, (unsigned)n4, n, n4);
            ++g_plonk_launches;
            BBG_CHECK(ntt_device(p->w4 + (size_t)k * n4, n4, 1, p->log_n + 2, OP_COSET_FFT, nullptr, sd));
        }
        {
            // the commitment is only queued here — wires alternate between two streams and MSM workspaces, so the sort /
            // fix-up / reduction kernels of one commitment run under the accumulate pass of the next — and finished below
            cudaStream_t ms = st;
            int workspace = 0;
#ifndef BBG_EMULATE
            if (p->overlap)
            {
                ms = p->msm_stream[k & 1];
                workspace = k & 1;
                BBG_CHECK(cudaEventRecord(p->ev_msm[k], st));
                BBG_CHECK(cudaStreamWaitEvent(ms, p->ev_msm[k], 0));
            }
#endif
            const void* scalars[1] = { p->w_coef + (size_t)k * n };
            BBG_CHECK(msm_launch(workspace, scalars, 1, p->d_srs, n, ms, &tickets[k]));
        }
        tr.mark("wire commitment queued", st);
    }
    for (int k = 0; k < 3; ++k)
    {
        hostg1::hxyzz r;
        BBG_CHECK(msm_finish(tickets[k], &r));
        hostg1::to_normalized_jacobian(r, out_xyz + 12 * k);
    }
    tr.mark("wire commitments", st);
    return bbg_rt::last_error();
}

// prover.cpp:137-225 compute_z_coefficients + :91-107 compute_z_commitment
int round_grand_product(Prover* p, const uint64_t* beta_, const uint64_t* gamma_, uint64_t out_xyz[12], cudaStream_t st)
{
    const size_t n = p->n;
    const fe beta = from_u64(beta_), gamma = from_u64(gamma_);
    fe* num = p->tmp;
    fe* den = p->tmp + n;
    if (!p->uploads_started) return 1007;
    BBG_CHECK(wait_item(p, Prover::ITEM_MAP, st)); // the mappings arrived behind round 1
    {
        bbg_prof::Scope prof(bbg_prof::PLONK_ELEMENTWISE, st);
        // permutation.hpp:13-88 (prover.cpp:659-661)
        if (!p->constants_cached)
        {
            BBG_LAUNCH_NOSYNC(sigma_from_mapping_kernel, dim3(grid_for(3 * n, 256)), dim3(256), st, p->sigma_lag, (const uint32_t*)p->map, p->pow_small,
                              (unsigned)n, (unsigned)(3 * n));
            ++g_plonk_launches;
        }
        p->sigma_ready = true;
        p->beta = beta;
        BBG_LAUNCH_NOSYNC(z_terms_kernel, dim3(grid_for(n, 128)), dim3(128), st, num, den, (const fe*)p->w_lag, (const fe*)p->sigma_lag, p->pow_small, beta,
                          gamma, (unsigned)n);
    }
    const size_t n4 = 4 * n;
    cudaStream_t sd = side_stream(p, st);
    {
        // a cold proving key brings sigma into coefficient form and onto the 4n coset (circuit constants, kept unscaled:
        // the quotient pass forms beta sigma + w + gamma itself); nothing to do when the key is cached
        if (!p->constants_cached)
        {
            BBG_CHECK(order_after(p, sd, st, 3));
            SideScratch scratch(sd, st);
            BBG_CHECK(bbg_rt::d2d(p->sigma, p->sigma_lag, 3 * n * 32, sd));
            BBG_CHECK(ntt_device(p->sigma, n, 3, p->log_n, OP_IFFT, nullptr, sd));
            BBG_LAUNCH_NOSYNC(pad_copy_kernel, dim3(grid_for(n4, 256), 3), dim3(256), sd, p->s4, (const fe*)p->sigma, (unsigned)n, (unsigned)n4, n, n4);
            ++g_plonk_launches;
            BBG_CHECK(ntt_device(p->s4, n4, 3, p->log_n + 2, OP_COSET_FFT, nullptr, sd));
        }
    }
    const unsigned runs = (unsigned)((n + ZRUN - 1) / ZRUN);
    bbg_prof::Scope* prof_scan = new bbg_prof::Scope(bbg_prof::PLONK_SCAN, st);
    BBG_LAUNCH_NOSYNC(prod_reduce_kernel, dim3((runs + 127) / 128, 2), dim3(128), st, (const fe*)p->tmp, p->aggs, (unsigned)n, (unsigned)ZRUN, n, p->aggs_stride);
    BBG_LAUNCH(prod_spine_kernel, dim3(2), dim3(SCAN_THREADS), 0, st, p->aggs, runs, p->aggs_stride);
    BBG_LAUNCH_NOSYNC(z_finish_kernel, dim3((runs + 63) / 64), dim3(64), st, p->z, (const fe*)num, (const fe*)den, (const fe*)p->aggs,
                      (const fe*)(p->aggs + p->aggs_stride), (unsigned)n);
    delete prof_scan;
    g_plonk_launches += 4;
    BBG_CHECK(ntt_device(p->z, n, 1, p->log_n, OP_IFFT, nullptr, st));
    {
        // Z on the 4n coset (:275; unscaled, alpha is applied by the quotient passes): queued behind the commitment
        BBG_CHECK(order_after(p, sd, st, 4));
        SideScratch scratch(sd, st);
        BBG_LAUNCH_NOSYNC(pad_copy_kernel, dim3(grid_for(n4, 256), 1), dim3(256), sd, p->z4, (const fe*)p->z, (unsigned)n, (unsigned)n4, n, n4);
        ++g_plonk_launches;
        BBG_CHECK(ntt_device(p->z4, n4, 1, p->log_n + 2, OP_COSET_FFT, nullptr, sd));
    }
    BBG_CHECK(commit(p, p->z, n, 1, out_xyz, st));
    return bbg_rt::last_error();
}

// prover.cpp:393-463 compute_quotient_polynomial (after the z commitment) + :109-124 compute_quotient_commitment
int round_quotient(Prover* p, const uint64_t* beta_, const uint64_t* gamma_, const uint64_t* alpha_, const uint64_t* alpha_base_,
                   uint64_t* out_xyz /* 3 x 12 */, cudaStream_t st)
{
    const size_t n = p->n, n2 = 2 * n, n4 = 4 * n;
    const fe beta = from_u64(beta_), gamma = from_u64(gamma_), alpha = from_u64(alpha_), alpha_base = from_u64(alpha_base_);
    Trace tr;
    if (!p->sigma_ready || !p->uploads_started) return 1007;
    BBG_CHECK(wait_item(p, Prover::ITEM_SEL, st)); // the selectors arrived behind rounds 1 and 2
    p->uploads_started = false;                    // every input of this proof is on the device
    // the 4n coset evaluations of the wires, of beta sigma + w + gamma and of Z were queued in rounds 1 and 2
    BBG_CHECK(order_after(p, st, side_stream(p, st), 5));
    const bool cached = p->constants_cached;
    tr.mark("side stream joined (w4, z4)", st);
    // L_1 on the 2n coset (:349-351)
    if (!p->l1_ready) BBG_CHECK(lagrange_fft_device(p->l1, p->log_n, p->log_n + 1, st));
    p->l1_ready = true;
    tr.mark("z4 + l1", st);
    // selectors: Lagrange -> coefficients (arithmetic_widget.cpp:62-66 and the other widgets' first lines)
    const int S = p->num_selectors;
    if (!cached) BBG_CHECK(ntt_device(p->q, n, (size_t)S, p->log_n, OP_IFFT, nullptr, st));

    QuotientConsts c;
    c.g = gen_k1();
    c.beta = beta;
    c.gamma = gamma;
    c.alpha = alpha;
    c.alpha_sqr = Fr::reduce(Fr::sqr(alpha));
    c.alpha_cube = Fr::reduce(Fr::mul(c.alpha_sqr, alpha));
    c.one = Fr::one();
    c.z_scale = alpha;
    c.g_beta = Fr::reduce(Fr::mul(c.g, beta));
    {
        const fe w = host_root_of_unity(p->log_n);
        c.neg_root_inv = Fr::reduce(Fr::neg(Fr::invert(w)));
    }
    // (g w_S^j)^n - 1 = g^n w_S^j - 1, S = target / source size (compute_multiplicative_subgroup, :104-127)
    fe gn = c.g;
    for (unsigned i = 0; i < p->log_n; ++i) gn = Fr::reduce(Fr::sqr(gn));
    auto fill_vinv = [&](unsigned log_s) {
        const fe ws = host_root_of_unity(log_s);
        fe cur = gn;
        for (unsigned j = 0; j < 4; ++j)
        {
            c.vinv[j] = j < (1u << log_s) ? Fr::invert(Fr::reduce(Fr::sub(cur, Fr::one()))) : Fr::zero();
            cur = Fr::reduce(Fr::mul(cur, ws));
        }
    };
    // selector polynomials sel .. sel + count - 1 -> their (unscaled) coset evaluations on the 2n / 4n domain: circuit
    // constants, computed once per proving key; the widgets' alpha powers are applied in the passes that consume them
    auto to_coset = [&](fe* dst, int sel, int count, size_t size, unsigned log_size) -> int {
        if (cached) return 0;
        BBG_LAUNCH_NOSYNC(pad_copy_kernel, dim3(grid_for(size, 256), (unsigned)count), dim3(256), st, dst, (const fe*)(p->q + (size_t)sel * n), (unsigned)n,
                          (unsigned)size, n, size);
        ++g_plonk_launches;
        return ntt_device(dst, size, (size_t)count, log_size, OP_COSET_FFT, nullptr, st);
    };
    const bool standard = p->num_widgets == 1 && p->widget_kind[0] == WIDGET_ARITHMETIC;
    if (standard)
    {
        // StandardComposer circuits: the gate identity is fused into the mid-domain pass (arithmetic_widget.cpp:68-97)
        BBG_CHECK(to_coset(p->q2, 0, 5, n2, p->log_n + 1));
        tr.mark("selectors ifft + coset_fft", st);
        fill_vinv(2);
        for (int j = 0; j < 4; ++j) c.vinv[j] = Fr::reduce(Fr::mul(c.vinv[j], alpha)); // Z's coset evaluations are unscaled
        BBG_LAUNCH_NOSYNC(quotient_large_kernel<true>, dim3(grid_for(n4, 128)), dim3(128), st, p->quot_large, (const fe*)p->s4, (const fe*)p->w4,
                          (const fe*)p->z4, p->pow_large, c, (unsigned)n4);
        fill_vinv(1);
        BBG_LAUNCH_NOSYNC(quotient_mid_kernel, dim3(grid_for(n2, 128)), dim3(128), st, p->quot_mid, (const fe*)p->z4, (const fe*)p->l1, (const fe*)p->w4,
                          (const fe*)p->q2, p->pow_mid, c, alpha_base, (unsigned)n2);
        g_plonk_launches += 2;
    }
    else
    {
        // any other widget mix: permutation / boundary terms first, then every widget adds its term in the prover's order
        // with the reference's alpha_base chain (prover.cpp:436-441), then the two divisions by Z_H*
        BBG_LAUNCH_NOSYNC(quotient_large_kernel<false>, dim3(grid_for(n4, 128)), dim3(128), st, p->quot_large, (const fe*)p->s4, (const fe*)p->w4,
                          (const fe*)p->z4, p->pow_large, c, (unsigned)n4);
        BBG_LAUNCH_NOSYNC(quotient_mid_base_kernel, dim3(grid_for(n2, 128)), dim3(128), st, p->quot_mid, (const fe*)p->z4, (const fe*)p->l1, c, (unsigned)n2);
        g_plonk_launches += 2;
        fe ab = alpha_base;
        fe* q2_cursor = p->q2;
        for (int w = 0; w < p->num_widgets; ++w)
        {
            const int sel = p->widget_first_selector[w];
            switch (p->widget_kind[w])
            {
            case WIDGET_ARITHMETIC: // arithmetic_widget.cpp:60-99
                BBG_CHECK(to_coset(q2_cursor, sel, 5, n2, p->log_n + 1));
                BBG_LAUNCH_NOSYNC(arith_mid_add_kernel, dim3(grid_for(n2, 128)), dim3(128), st, p->quot_mid, (const fe*)p->w4, (const fe*)q2_cursor, ab,
                                  (unsigned)n2);
                q2_cursor += 5 * n2;
                ab = Fr::reduce(Fr::mul(ab, alpha));
                break;
            case WIDGET_BOOL: // bool_widget.cpp:62-100
            {
                Scale3 sc;
                fe k = ab;
                for (int j = 0; j < 3; ++j)
                {
                    sc.s[j] = k; // alpha_base, alpha_base alpha, alpha_base alpha^2 (bool_widget.cpp:72-74)
                    k = Fr::reduce(Fr::mul(k, alpha));
                }
                BBG_CHECK(to_coset(q2_cursor, sel, 3, n2, p->log_n + 1));
                BBG_LAUNCH_NOSYNC(bool_mid_add_kernel, dim3(grid_for(n2, 128)), dim3(128), st, p->quot_mid, (const fe*)p->w4, (const fe*)q2_cursor, sc,
                                  (unsigned)n2);
                q2_cursor += 3 * n2;
                ab = k; // alpha_base * alpha^3
                break;
            }
            case WIDGET_MIMC: // mimc_widget.cpp:57-89
                BBG_CHECK(to_coset(p->q4, sel, 2, n4, p->log_n + 2));
                BBG_LAUNCH_NOSYNC(mimc_large_add_kernel, dim3(grid_for(n4, 128)), dim3(128), st, p->quot_large, (const fe*)p->w4, (const fe*)p->q4, alpha, ab,
                                  (unsigned)n4);
                ab = Fr::reduce(Fr::mul(ab, c.alpha_sqr));
                break;
            case WIDGET_SEQUENTIAL: // sequential_widget.cpp:47-62
            {
                const fe old_alpha = Fr::reduce(Fr::mul(ab, Fr::invert(alpha)));
                BBG_CHECK(to_coset(q2_cursor, sel, 1, n2, p->log_n + 1));
                BBG_LAUNCH_NOSYNC(seq_mid_add_kernel, dim3(grid_for(n2, 128)), dim3(128), st, p->quot_mid, (const fe*)(p->w4 + 2 * n4), (const fe*)q2_cursor,
                                  old_alpha, (unsigned)n2);
                q2_cursor += n2;
                break;
            }
            }
            ++g_plonk_launches;
        }
        tr.mark("selectors ifft + coset_fft + widget terms", st);
        fill_vinv(2);
        BBG_LAUNCH_NOSYNC(divide_vanishing_kernel, dim3(grid_for(n4, 128)), dim3(128), st, p->quot_large, p->pow_large, c, 3u, (unsigned)n4);
        fill_vinv(1);
        BBG_LAUNCH_NOSYNC(divide_vanishing_kernel, dim3(grid_for(n2, 128)), dim3(128), st, p->quot_mid, p->pow_mid, c, 1u, (unsigned)n2);
        g_plonk_launches += 2;
    }
    if (!cached && p->key_cache_enabled)
    {
        // every device-side constant now belongs to the hash the helper thread computed for this proof
        p->key_hash[0] = p->pending_hash[0];
        p->key_hash[1] = p->pending_hash[1];
        p->key_valid = true;
    }
    tr.mark("quotient kernels", st);
    BBG_CHECK(ntt_device(p->quot_mid, n2, 1, p->log_n + 1, OP_COSET_IFFT, nullptr, st));
    BBG_CHECK(ntt_device(p->quot_large, n4, 1, p->log_n + 2, OP_COSET_IFFT, nullptr, st));
    BBG_LAUNCH_NOSYNC(add_into_kernel, dim3(grid_for(n2, 256)), dim3(256), st, p->quot_large, (const fe*)p->quot_mid, (unsigned)n2);
    g_plonk_launches += 3;
    tr.mark("coset_ifft x2 + add", st);
    BBG_CHECK(commit(p, p->quot_large, n, 3, out_xyz, st));
    tr.mark("3 commitments", st);
    return bbg_rt::last_error();
}

static int run_evals(Prover* p, const EvalJobs& jobs, int count, uint64_t* out, cudaStream_t st)
{
    unsigned max_blocks = 1;
    for (int j = 0; j < count; ++j)
    {
        const unsigned nb = (jobs.len[j] + EVAL_SPAN - 1) / EVAL_SPAN;
        if (nb > max_blocks) max_blocks = nb;
    }
    BBG_LAUNCH(eval_partial_kernel, dim3(max_blocks, (unsigned)count), dim3(EVAL_THREADS), 0, st, jobs, p->eval_partial, p->partial_stride);
    BBG_LAUNCH(eval_final_kernel, dim3((unsigned)count), dim3(SCAN_THREADS), 0, st, jobs, (const fe*)p->eval_partial, p->partial_stride, p->eval_out);
    g_plonk_launches += 2;
    BBG_CHECK(bbg_rt::d2h(out, p->eval_out, (size_t)count * 32, st));
    return bbg_rt::sync(st);
}

// prover.cpp:465-477 (+ :455-463 shifted wire, mimc_widget.cpp:91-94): w_l, w_r, w_o, beta sigma_1, beta sigma_2 at z; Z at z w;
// the quotient's 3n coefficients at z; then w_o at z w when a widget needs it and q_mimc_coefficient at z (zero otherwise)
int round_evaluations(Prover* p, const uint64_t* zeta_, const uint64_t* zeta_omega_, uint64_t* out /* 9 x 4 */, cudaStream_t st)
{
    const size_t n = p->n;
    EvalJobs jobs;
    const fe zeta = from_u64(zeta_), zw = from_u64(zeta_omega_);
    const fe* polys[7] = { p->w_coef, p->w_coef + n, p->w_coef + 2 * n, p->sigma, p->sigma + n, p->z, p->quot_large };
    for (int j = 0; j < MAX_EVAL_JOBS; ++j)
    {
        jobs.poly[j] = nullptr;
        jobs.len[j] = 0;
        jobs.point[j] = Fr::zero();
    }
    for (int j = 0; j < 7; ++j)
    {
        jobs.poly[j] = polys[j];
        jobs.len[j] = (unsigned)(j == 6 ? 3 * n : n);
        jobs.point[j] = j == 5 ? zw : zeta;
    }
    int count = 7, shifted_at = -1, mimc_at = -1;
    for (int w = 0; w < p->num_widgets; ++w)
    {
        const int kind = p->widget_kind[w];
        if ((kind == WIDGET_MIMC || kind == WIDGET_SEQUENTIAL) && shifted_at < 0)
        {
            shifted_at = count;
            jobs.poly[count] = p->w_coef + 2 * n;
            jobs.len[count] = (unsigned)n;
            jobs.point[count++] = zw;
        }
        if (kind == WIDGET_MIMC)
        {
            mimc_at = count;
            jobs.poly[count] = p->q + (size_t)(p->widget_first_selector[w] + 1) * n;
            jobs.len[count] = (unsigned)n;
            jobs.point[count++] = zeta;
        }
    }
    uint64_t tmp[MAX_EVAL_JOBS * 4];
    BBG_CHECK(run_evals(p, jobs, count, tmp, st));
    // sigma is stored unscaled: hand back [beta sigma_k](z) as the reference's prover holds it (:473-477 divide it out again)
    for (int j = 3; j <= 4; ++j)
    {
        const fe v = Fr::mul_full(from_u64(tmp + 4 * j), p->beta);
        memcpy(tmp + 4 * j, v.v, 32);
    }
    memset(out, 0, 9 * 32);
    memcpy(out, tmp, 7 * 32);
    if (shifted_at >= 0) memcpy(out + 28, tmp + 4 * shifted_at, 32);
    if (mimc_at >= 0) memcpy(out + 32, tmp + 4 * mimc_at, 32);
    return 0;
}

// prover.cpp:479-503 and the widgets' compute_linear_contribution: r[i] = s_0 z[i] + s_1 [beta sigma_3][i] + sum_k s_(2+k) q_k[i]
// over the circuit's selectors in coefficient form (the host has folded wire evaluations and the alpha chain into s_k)
int round_linearise(Prover* p, const uint64_t* scalars /* (2 + selectors) x 4 */, const uint64_t* zeta_, uint64_t out_eval[4], cudaStream_t st)
{
    const size_t n = p->n;
    LinTerms t;
    t.count = 0;
    auto term = [&](const fe* src, const uint64_t* c) {
        t.src[t.count] = src;
        t.c[t.count++] = from_u64(c);
    };
    term(p->z, scalars);
    term(p->sigma + 2 * n, scalars + 4);
    t.c[1] = Fr::mul_full(t.c[1], p->beta); // s_1 is meant for beta sigma_3; sigma is stored unscaled
    bool first = true;
    for (int k = 0; k < p->num_selectors; ++k)
    {
        if (t.count == MAX_TERMS)
        {
            if (first) BBG_LAUNCH_NOSYNC(axpy_terms_kernel<false>, dim3(grid_for(n, 128)), dim3(128), st, p->r, t, (unsigned)n);
            else BBG_LAUNCH_NOSYNC(axpy_terms_kernel<true>, dim3(grid_for(n, 128)), dim3(128), st, p->r, t, (unsigned)n);
            first = false;
            t.count = 0;
            ++g_plonk_launches;
        }
        term(p->q + (size_t)k * n, scalars + 8 + 4 * k);
    }
    if (first) BBG_LAUNCH_NOSYNC(axpy_terms_kernel<false>, dim3(grid_for(n, 128)), dim3(128), st, p->r, t, (unsigned)n);
    else BBG_LAUNCH_NOSYNC(axpy_terms_kernel<true>, dim3(grid_for(n, 128)), dim3(128), st, p->r, t, (unsigned)n);
    ++g_plonk_launches;
    EvalJobs jobs;
    for (int j = 0; j < MAX_EVAL_JOBS; ++j)
    {
        jobs.poly[j] = nullptr;
        jobs.len[j] = 0;
        jobs.point[j] = Fr::zero();
    }
    jobs.poly[0] = p->r;
    jobs.len[0] = (unsigned)n;
    jobs.point[0] = from_u64(zeta_);
    return run_evals(p, jobs, 1, out_eval, st);
}

// prover.cpp:505-655 compute_opening_elements after the nu challenge
// wire_shift: nu-power coefficients of w_l, w_r, w_o in the shifted opening polynomial (:597-631; zero = not needed);
// selector_terms: coefficients of the selectors in the opening polynomial (the widgets' compute_opening_poly_contribution)
int round_openings(Prover* p, const uint64_t* nu_powers /* 7 x 4 */, const uint64_t* beta_inv_, const uint64_t* zeta_, const uint64_t* zeta_omega_,
                   const uint64_t* wire_shift /* 3 x 4 */, const uint64_t* selector_terms /* selectors x 4 */, uint64_t* out_xyz /* 2 x 12 */,
                   cudaStream_t st)
{
    const size_t n = p->n;
    OpeningConsts c;
    for (int i = 0; i < 7; ++i) c.nu[i] = from_u64(nu_powers + 4 * i);
    (void)beta_inv_; // the reference divides beta back out of its beta-scaled sigma (:586); sigma is stored unscaled here
    const fe zeta = from_u64(zeta_), zw = from_u64(zeta_omega_);
    c.z_pow_n = host_pow(zeta, n);
    c.z_pow_2n = host_pow(zeta, 2 * n);
    fe* opening = p->tmp;
    fe* shifted = p->tmp + n;
    BBG_LAUNCH_NOSYNC(opening_combine_kernel, dim3(grid_for(n, 128)), dim3(128), st, opening, shifted, (const fe*)p->quot_large, (const fe*)p->r,
                      (const fe*)p->w_coef, (const fe*)p->sigma, (const fe*)p->z, c, (unsigned)n);
    {
        LinTerms t;
        t.count = 0;
        for (int k = 0; k < 3; ++k)
        {
            const fe ck = from_u64(wire_shift + 4 * k);
            if (Fr::is_zero_raw(ck)) continue;
            t.src[t.count] = p->w_coef + (size_t)k * n;
            t.c[t.count++] = ck;
        }
        if (t.count > 0)
        {
            BBG_LAUNCH_NOSYNC(axpy_terms_kernel<true>, dim3(grid_for(n, 128)), dim3(128), st, shifted, t, (unsigned)n);
            ++g_plonk_launches;
        }
        t.count = 0;
        for (int k = 0; k < p->num_selectors; ++k)
        {
            const fe ck = from_u64(selector_terms + 4 * k);
            if (Fr::is_zero_raw(ck)) continue;
            t.src[t.count] = p->q + (size_t)k * n;
            t.c[t.count++] = ck;
        }
        if (t.count > 0)
        {
            BBG_LAUNCH_NOSYNC(axpy_terms_kernel<true>, dim3(grid_for(n, 128)), dim3(128), st, opening, t, (unsigned)n);
            ++g_plonk_launches;
        }
    }
    const unsigned run = ZRUN;
    const unsigned runs = (unsigned)((n + run - 1) / run);
    const unsigned per = (runs + SCAN_THREADS - 1) / SCAN_THREADS;
    KatePoints pts;
    pts.z[0] = zeta;
    pts.z[1] = zw;
    for (int k = 0; k < 2; ++k)
    {
        pts.z_run[k] = host_pow(pts.z[k], run);
        pts.z_span[k] = host_pow(pts.z_run[k], per);
    }
    BBG_LAUNCH_NOSYNC(kate_reduce_kernel, dim3((runs + 127) / 128, 2), dim3(128), st, (const fe*)p->tmp, p->aggs, pts, (unsigned)n, run, p->aggs_stride);
    BBG_LAUNCH(kate_spine_kernel, dim3(2), dim3(SCAN_THREADS), 0, st, p->aggs, pts, runs, p->aggs_stride);
    fe* quotients = p->w4; // free since the quotient round; 2 x n
    BBG_LAUNCH_NOSYNC(kate_apply_kernel, dim3((runs + 127) / 128, 2), dim3(128), st, quotients, (const fe*)p->tmp, (const fe*)p->aggs, pts, (unsigned)n, run,
                      p->aggs_stride);
    g_plonk_launches += 4;
    BBG_CHECK(commit(p, quotients, n, 2, out_xyz, st));
    return bbg_rt::last_error();
}

// =================================================================================================================
// The same kernels behind the reference's stand-alone helpers (polynomial_arithmetic.cpp:337-373, :478-591), for callers
// that keep the reference's round structure (shim/polynomial_arithmetic_gpu.cpp): device buffers in, no Prover object.
// =================================================================================================================
namespace
{
struct HelperBuf
{
    void* p = nullptr;
    size_t bytes = 0;
    int ensure(size_t need)
    {
        if (need <= bytes) return 0;
        if (p) bbg_rt::dev_free(p);
        p = nullptr;
        bytes = 0;
        const int e = bbg_rt::dev_alloc(&p, need);
        if (e == 0) bytes = need;
        return e;
    }
} g_helper;
} // namespace

void release_helpers()
{
    if (g_helper.p) bbg_rt::dev_free(g_helper.p);
    g_helper.p = nullptr;
    g_helper.bytes = 0;
}

// F(z) = sum_i coeffs[i] z^i over len coefficients, canonical, to the host
int evaluate_device(const void* d_poly, size_t len, const uint64_t* z, uint64_t out[4], cudaStream_t st)
{
    if (len >= ((size_t)1 << 32)) return 1002;
    if (len == 0)
    {
        memset(out, 0, 32);
        return 0;
    }
    const unsigned nb = (unsigned)((len + EVAL_SPAN - 1) / EVAL_SPAN);
    BBG_CHECK(g_helper.ensure(((size_t)nb + 16) * 32));
    fe* partial = (fe*)g_helper.p;
    fe* result = partial + nb + 8;
    EvalJobs jobs;
    for (int j = 0; j < MAX_EVAL_JOBS; ++j)
    {
        jobs.poly[j] = nullptr;
        jobs.len[j] = 0;
        jobs.point[j] = Fr::zero();
    }
    jobs.poly[0] = (const fe*)d_poly;
    jobs.len[0] = (unsigned)len;
    jobs.point[0] = from_u64(z);
    BBG_LAUNCH(eval_partial_kernel, dim3(nb, 1), dim3(EVAL_THREADS), 0, st, jobs, partial, nb + 8);
    BBG_LAUNCH(eval_final_kernel, dim3(1), dim3(SCAN_THREADS), 0, st, jobs, (const fe*)partial, nb + 8, result);
    g_plonk_launches += 2;
    BBG_CHECK(bbg_rt::d2h(out, result, 32, st));
    return bbg_rt::sync(st);
}

// coeffs[i] *= (g w_T^i - w_n^(n-1)) / ((g w_T^i)^n - 1) on the 2^log_target coset, n = 2^log_src; canonical, in place
int divide_by_pseudo_vanishing_device(void* d_coeffs, unsigned log_src, unsigned log_target, cudaStream_t st)
{
    if (log_target < log_src || log_target - log_src > 2 || log_target > 28 || log_src < 1) return 1002; // subgroup of 1, 2 or 4
    const unsigned log_s = log_target - log_src;
    const size_t T = (size_t)1 << log_target;
    const int lo_log = (int)(log_target < 11 ? log_target : 11);
    const size_t lo_count = (size_t)1 << lo_log, hi_count = log_target > (unsigned)lo_log ? (size_t)1 << (log_target - lo_log) : 0;
    BBG_CHECK(g_helper.ensure((lo_count + hi_count + 8) * 32));
    fe* lo = (fe*)g_helper.p;
    fe* hi = lo + lo_count;
    const fe w = host_root_of_unity(log_target);
    BBG_LAUNCH_NOSYNC(powers_kernel, dim3((unsigned)((lo_count + 127) / 128)), dim3(128), st, lo, w, (unsigned)lo_count);
    if (hi_count) BBG_LAUNCH_NOSYNC(powers_kernel, dim3((unsigned)((hi_count + 127) / 128)), dim3(128), st, hi, host_pow(w, lo_count), (unsigned)hi_count);
    PowTable table{ lo, hi_count ? hi : nullptr, lo_log };
    QuotientConsts c;
    c.g = gen_k1();
    c.beta = c.gamma = c.alpha = c.alpha_sqr = c.alpha_cube = c.g_beta = c.z_scale = Fr::zero();
    c.one = Fr::one();
    c.neg_root_inv = Fr::reduce(Fr::neg(Fr::invert(host_root_of_unity(log_src))));
    fe gn = c.g;
    for (unsigned i = 0; i < log_src; ++i) gn = Fr::reduce(Fr::sqr(gn));
    const fe ws = host_root_of_unity(log_s);
    fe cur = gn;
    for (unsigned j = 0; j < 4; ++j)
    {
        c.vinv[j] = j < (1u << log_s) ? Fr::invert(Fr::reduce(Fr::sub(cur, Fr::one()))) : Fr::zero();
        cur = Fr::reduce(Fr::mul(cur, ws));
    }
    BBG_LAUNCH_NOSYNC(divide_vanishing_kernel, dim3(grid_for(T, 128)), dim3(128), st, (fe*)d_coeffs, table, c, (1u << log_s) - 1u, (unsigned)T, 1);
    g_plonk_launches += 3;
    return bbg_rt::last_error();
}

// d_dest = (F(X) - F(z)) / (X - z) for F = d_src (n coefficients; d_dest must not alias d_src), canonical; F(z) to the host
int kate_opening_device(const void* d_src, void* d_dest, size_t n, const uint64_t* z_, uint64_t f_out[4], cudaStream_t st)
{
    if (n == 0 || n >= ((size_t)1 << 32) || d_src == d_dest) return 1007;
    BBG_CHECK(evaluate_device(d_src, n, z_, f_out, st)); // (uses, then releases, the helper buffer before the scan does)
    const unsigned run = ZRUN;
    const unsigned runs = (unsigned)((n + run - 1) / run);
    const unsigned per = (runs + SCAN_THREADS - 1) / SCAN_THREADS;
    const unsigned stride = (runs + 7) & ~7u;
    BBG_CHECK(g_helper.ensure(((size_t)stride + 8) * 32));
    fe* aggs = (fe*)g_helper.p;
    KatePoints pts;
    pts.z[0] = pts.z[1] = from_u64(z_);
    pts.z_run[0] = pts.z_run[1] = host_pow(pts.z[0], run);
    pts.z_span[0] = pts.z_span[1] = host_pow(pts.z_run[0], per);
    BBG_LAUNCH_NOSYNC(kate_reduce_kernel, dim3((runs + 127) / 128, 1), dim3(128), st, (const fe*)d_src, aggs, pts, (unsigned)n, run, stride);
    BBG_LAUNCH(kate_spine_kernel, dim3(1), dim3(SCAN_THREADS), 0, st, aggs, pts, runs, stride);
    BBG_LAUNCH_NOSYNC(kate_apply_kernel, dim3((runs + 127) / 128, 1), dim3(128), st, (fe*)d_dest, (const fe*)d_src, (const fe*)aggs, pts, (unsigned)n, run,
                      stride, 1);
    g_plonk_launches += 3;
    return bbg_rt::last_error();
}

size_t launch_count() { return g_plonk_launches; }
} // namespace plonk
} // namespace bbg
