// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// extern "C" harness around the UNMODIFIED reference sources (compiled where they lie under
// /root/reference/src by oracle/Makefile into oracle/_ref/libbb_ref.so).  It exists to
//   (1) pin oracle/bb_oracle.c (our plain-C restatement) against the real reference,
//   (2) generate the golden vectors under tests/golden/,
//   (3) serve as the "reference" CPU baseline of bench.py (multithreaded x86-asm path).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.  Nothing here is copied from the reference: it only *calls* it.
//
// All field elements cross this ABI as 4 x uint64 little-endian limbs (Montgomery form unless
// stated), affine points as 8 x uint64 (x,y), Jacobian points as 12 x uint64 (x,y,z).
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <omp.h>

#include <barretenberg/curves/bn254/fq.hpp>
#include <barretenberg/curves/bn254/fr.hpp>
#include <barretenberg/curves/bn254/g1.hpp>
#include <barretenberg/curves/bn254/scalar_multiplication.hpp>
#include <barretenberg/groups/wnaf.hpp>
#include <barretenberg/polynomials/evaluation_domain.hpp>
#include <barretenberg/polynomials/polynomial_arithmetic.hpp>

using namespace barretenberg;

namespace
{
template <typename T> T* aligned_copy(const uint64_t* src, size_t count)
{
    T* p = (T*)aligned_alloc(64, sizeof(T) * (count ? count : 1));
    memcpy((void*)p, src, sizeof(T) * count);
    return p;
}
inline fq::field_t ldq(const uint64_t* p) { fq::field_t r; memcpy(&r, p, 32); return r; }
inline fr::field_t ldr(const uint64_t* p) { fr::field_t r; memcpy(&r, p, 32); return r; }
} // namespace

#pragma GCC visibility push(default)
extern "C" {

int ref_omp_threads() { return omp_get_max_threads(); }
void ref_set_omp_threads(int n) { omp_set_num_threads(n); }

// ---- fields (fields/field.hpp, field_impl_asm.tcc) ----------------------------------------
#define FIELD_BINOP(NAME, F, OP, LD)                                                               \
    void NAME(const uint64_t* a, const uint64_t* b, uint64_t* r)                                   \
    {                                                                                              \
        F::field_t x = LD(a), y = LD(b), z;                                                        \
        F::OP(x, y, z);                                                                            \
        memcpy(r, &z, 32);                                                                         \
    }
FIELD_BINOP(ref_fq_mul, fq, __mul, ldq)
FIELD_BINOP(ref_fq_mul_coarse, fq, __mul_with_coarse_reduction, ldq)
FIELD_BINOP(ref_fq_add, fq, __add, ldq)
FIELD_BINOP(ref_fq_sub, fq, __sub, ldq)
FIELD_BINOP(ref_fr_mul, fr, __mul, ldr)
FIELD_BINOP(ref_fr_mul_coarse, fr, __mul_with_coarse_reduction, ldr)
FIELD_BINOP(ref_fr_add, fr, __add, ldr)
FIELD_BINOP(ref_fr_sub, fr, __sub, ldr)
FIELD_BINOP(ref_fr_add_coarse, fr, __add_with_coarse_reduction, ldr)
FIELD_BINOP(ref_fr_sub_coarse, fr, __sub_with_coarse_reduction, ldr)

void ref_fq_sqr(const uint64_t* a, uint64_t* r) { fq::field_t x = ldq(a), z; fq::__sqr(x, z); memcpy(r, &z, 32); }
void ref_fr_sqr(const uint64_t* a, uint64_t* r) { fr::field_t x = ldr(a), z; fr::__sqr(x, z); memcpy(r, &z, 32); }
void ref_fq_invert(const uint64_t* a, uint64_t* r) { fq::field_t x = ldq(a), z; fq::__invert(x, z); memcpy(r, &z, 32); }
void ref_fr_invert(const uint64_t* a, uint64_t* r) { fr::field_t x = ldr(a), z; fr::__invert(x, z); memcpy(r, &z, 32); }
void ref_fq_to_mont(const uint64_t* a, uint64_t* r) { fq::field_t x = ldq(a), z; fq::__to_montgomery_form(x, z); memcpy(r, &z, 32); }
void ref_fq_from_mont(const uint64_t* a, uint64_t* r) { fq::field_t x = ldq(a), z; fq::__from_montgomery_form(x, z); memcpy(r, &z, 32); }
void ref_fr_to_mont(const uint64_t* a, uint64_t* r) { fr::field_t x = ldr(a), z; fr::__to_montgomery_form(x, z); memcpy(r, &z, 32); }
void ref_fr_from_mont(const uint64_t* a, uint64_t* r) { fr::field_t x = ldr(a), z; fr::__from_montgomery_form(x, z); memcpy(r, &z, 32); }
void ref_fq_neg(const uint64_t* a, uint64_t* r) { fq::field_t x = ldq(a), z; fq::__neg(x, z); memcpy(r, &z, 32); }
void ref_fq_mul_beta(const uint64_t* a, uint64_t* r) { fq::field_t x = ldq(a), z; fq::__mul_beta(x, z); memcpy(r, &z, 32); }
void ref_fr_reduce_once(const uint64_t* a, uint64_t* r) { fr::field_t x = ldr(a), z; fr::reduce_once(x, z); memcpy(r, &z, 32); }

// batched variants (n independent ops) so python can throw 10^5..10^6 seeded pairs at them
void ref_fq_mul_n(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t n) { for (size_t i = 0; i < n; ++i) ref_fq_mul(a + 4 * i, b + 4 * i, r + 4 * i); }
void ref_fr_mul_n(const uint64_t* a, const uint64_t* b, uint64_t* r, size_t n) { for (size_t i = 0; i < n; ++i) ref_fr_mul(a + 4 * i, b + 4 * i, r + 4 * i); }
void ref_fr_to_mont_n(const uint64_t* a, uint64_t* r, size_t n) { for (size_t i = 0; i < n; ++i) ref_fr_to_mont(a + 4 * i, r + 4 * i); }
void ref_fr_from_mont_n(const uint64_t* a, uint64_t* r, size_t n) { for (size_t i = 0; i < n; ++i) ref_fr_from_mont(a + 4 * i, r + 4 * i); }

// constants (curves/bn254/fq.hpp, fr.hpp, g1.hpp) — index selects which
void ref_constant(int which, uint64_t* r)
{
    switch (which)
    {
    case 0: memcpy(r, &fq::modulus, 32); break;
    case 1: memcpy(r, &fq::r_squared, 32); break;
    case 2: memcpy(r, &fq::one, 32); break;
    case 3: memcpy(r, &fq::beta, 32); break;
    case 4: memcpy(r, &fr::modulus, 32); break;
    case 5: memcpy(r, &fr::r_squared, 32); break;
    case 6: memcpy(r, &fr::one, 32); break;
    case 7: memcpy(r, &fr::beta, 32); break;
    case 8: memcpy(r, &fr::root_of_unity, 32); break;
    case 9: memcpy(r, &fr::multiplicative_generator, 32); break;
    case 10: memcpy(r, &fr::multiplicative_generator_inverse, 32); break;
    case 11: { g1::affine_element o = g1::affine_one(); memcpy(r, &o.x, 32); break; }
    case 12: { g1::affine_element o = g1::affine_one(); memcpy(r, &o.y, 32); break; }
    case 13: memcpy(r, &Bn254G1Params::b, 32); break;
    default: memset(r, 0, 32);
    }
}
uint64_t ref_r_inv(int fr_not_fq) { return fr_not_fq ? FrParams::r_inv : Bn254FqParams::r_inv; }

// ---- endomorphism split + wNAF (field.hpp:413-485, wnaf.hpp:38-55) ------------------------
// k: non-Montgomery, < r. Exactly the aliased call of scalar_multiplication.cpp:292:
// out[0..1] = k1 (low 128 bits), out[2..3] = k2 (low 128 bits).
void ref_split_endo(const uint64_t* k, uint64_t* out)
{
    fr::field_t s = ldr(k);
    fr::split_into_endomorphism_scalars(s, s, *(fr::field_t*)&s.data[2]);
    memcpy(out, &s, 32);
}
void ref_split_endo_n(const uint64_t* k, uint64_t* out, size_t n) { for (size_t i = 0; i < n; ++i) ref_split_endo(k + 4 * i, out + 4 * i); }

// wnaf: writes entries at wnaf[i*num_points]; returns skew
int ref_fixed_wnaf(const uint64_t* scalar128, uint32_t* wnaf, size_t num_points, size_t wnaf_bits)
{
    uint64_t s[2] = { scalar128[0], scalar128[1] };
    bool skew = false;
    wnaf::fixed_wnaf(s, wnaf, skew, num_points, wnaf_bits);
    return skew ? 1 : 0;
}

// ---- G1 (groups/group.hpp) -----------------------------------------------------------------
void ref_g1_mixed_add(const uint64_t* p1, const uint64_t* p2_affine, uint64_t* out)
{
    g1::element a, r; g1::affine_element b;
    memcpy(&a, p1, 96); memcpy(&b, p2_affine, 64);
    g1::mixed_add(a, b, r);
    memcpy(out, &r, 96);
}
void ref_g1_add(const uint64_t* p1, const uint64_t* p2, uint64_t* out)
{
    g1::element a, b, r;
    memcpy(&a, p1, 96); memcpy(&b, p2, 96);
    g1::add(a, b, r);
    memcpy(out, &r, 96);
}
void ref_g1_dbl(const uint64_t* p1, uint64_t* out)
{
    g1::element a, r;
    memcpy(&a, p1, 96);
    g1::dbl(a, r);
    memcpy(out, &r, 96);
}
void ref_g1_normalize(const uint64_t* p1, uint64_t* out)
{
    g1::element a;
    memcpy(&a, p1, 96);
    g1::element r = g1::normalize(a);
    memcpy(out, &r, 96);
}
void ref_g1_batch_normalize(uint64_t* pts, size_t n)
{
    g1::element* p = aligned_copy<g1::element>(pts, n);
    g1::batch_normalize(p, n);
    memcpy(pts, p, 96 * n);
    free(p);
}
// affine result of scalar (Montgomery) * affine point; infinity => y msb set
void ref_g1_group_exponentiation(const uint64_t* affine, const uint64_t* scalar_mont, uint64_t* out_affine)
{
    g1::affine_element a; memcpy(&a, affine, 64);
    fr::field_t s = ldr(scalar_mont);
    g1::affine_element r = g1::group_exponentiation(a, s);
    memcpy(out_affine, &r, 64);
}
int ref_g1_on_curve(const uint64_t* affine)
{
    g1::affine_element a; memcpy(&a, affine, 64);
    return g1::on_curve(a) ? 1 : 0;
}

// points[i] = (start + i*step) * G for i < n, normalised affine (Montgomery coords).  Built by
// repeated g1::mixed_add + g1::batch_normalize (the closed-form MSM check).
void ref_g1_arith_progression(const uint64_t* start_mont, const uint64_t* step_mont, uint64_t* out_affine, size_t n)
{
    if (n == 0) return;
    fr::field_t a0 = ldr(start_mont), d = ldr(step_mont);
    g1::affine_element base = g1::group_exponentiation(g1::affine_one(), a0);
    g1::affine_element step = g1::group_exponentiation(g1::affine_one(), d);
    g1::element* acc = (g1::element*)aligned_alloc(64, sizeof(g1::element) * n);
    fq::__copy(base.x, acc[0].x); fq::__copy(base.y, acc[0].y); fq::__copy(fq::one, acc[0].z);
    if (g1::is_point_at_infinity(base)) g1::set_infinity(acc[0]);
    for (size_t i = 1; i < n; ++i)
    {
        if (g1::is_point_at_infinity(step)) { g1::copy(&acc[i - 1], &acc[i]); continue; }
        g1::mixed_add(acc[i - 1], step, acc[i]);
    }
    g1::batch_normalize(acc, n);
    for (size_t i = 0; i < n; ++i)
    {
        memcpy(out_affine + 8 * i, &acc[i].x, 32);
        memcpy(out_affine + 8 * i + 4, &acc[i].y, 32);
    }
    free(acc);
}

// ---- MSM (curves/bn254/scalar_multiplication.cpp) ------------------------------------------
// points_n_affine: n affine points -> writes the interleaved 2n table (table may alias nothing here)
void ref_generate_pippenger_point_table(const uint64_t* points_n_affine, uint64_t* table_2n, size_t n)
{
    g1::affine_element* t = (g1::affine_element*)aligned_alloc(64, sizeof(g1::affine_element) * 2 * (n ? n : 1));
    memcpy((void*)t, points_n_affine, 64 * n);
    scalar_multiplication::generate_pippenger_point_table(t, t, n);
    memcpy(table_2n, t, 128 * n);
    free(t);
}
// un-normalised Jacobian result, exactly what pippenger returns
void ref_pippenger(const uint64_t* scalars_mont, const uint64_t* table_2n, size_t n, size_t forced_bucket_width, uint64_t* out_jac)
{
    fr::field_t* s = aligned_copy<fr::field_t>(scalars_mont, n);
    g1::affine_element* t = aligned_copy<g1::affine_element>(table_2n, 2 * n);
    g1::element r = scalar_multiplication::pippenger(s, t, n, forced_bucket_width);
    memcpy(out_jac, &r, 96);
    free(s); free(t);
}
// zero-copy variants for benchmarking (caller guarantees 32-byte alignment)
void ref_pippenger_inplace(uint64_t* scalars_mont, uint64_t* table_2n, size_t n, size_t forced_bucket_width, uint64_t* out_jac)
{
    g1::element r = scalar_multiplication::pippenger((fr::field_t*)scalars_mont, (g1::affine_element*)table_2n, n, forced_bucket_width);
    memcpy(out_jac, &r, 96);
}
// B same-size MSMs over ONE table; outputs normalised (z = one) like the reference
void ref_batched_scalar_multiplications(uint64_t* const* scalars_mont, uint64_t* table_2n, size_t n, size_t batches, uint64_t* out_jac /*batches*12*/)
{
    scalar_multiplication::multiplication_state* st = new scalar_multiplication::multiplication_state[batches];
    for (size_t i = 0; i < batches; ++i)
    {
        st[i].points = (g1::affine_element*)table_2n;
        st[i].scalars = (fr::field_t*)scalars_mont[i];
        st[i].num_elements = n;
    }
    scalar_multiplication::batched_scalar_multiplications(st, batches);
    for (size_t i = 0; i < batches; ++i) memcpy(out_jac + 12 * i, &st[i].output, 96);
    delete[] st;
}
size_t ref_get_optimal_bucket_width(size_t n) { return scalar_multiplication::get_optimal_bucket_width(n); }

// ---- NTT (polynomials/polynomial_arithmetic.cpp, evaluation_domain.cpp) ---------------------
struct ref_domain { evaluation_domain* d; };
void* ref_domain_new(size_t n)
{
    evaluation_domain* d = new evaluation_domain(n);
    if (n >= 2) d->compute_lookup_table(); // log2(n)-1 rounds; n=2 has zero rounds (vector stays empty)
    return (void*)d;
}
void ref_domain_free(void* h) { delete (evaluation_domain*)h; }
// which: 0 root, 1 root_inverse, 2 domain, 3 domain_inverse, 4 generator, 5 generator_inverse
void ref_domain_constant(void* h, int which, uint64_t* r)
{
    evaluation_domain* d = (evaluation_domain*)h;
    const fr::field_t* src[6] = { &d->root, &d->root_inverse, &d->domain, &d->domain_inverse, &d->generator, &d->generator_inverse };
    memcpy(r, src[which], 32);
}
size_t ref_domain_num_threads(void* h) { return ((evaluation_domain*)h)->num_threads; }
// op: 0 fft, 1 ifft, 2 coset_fft, 3 coset_ifft, 4 fft_with_constant, 5 ifft_with_constant, 6 coset_fft_with_constant
void ref_ntt(void* h, int op, uint64_t* coeffs /*in place, caller-aligned*/, const uint64_t* constant)
{
    evaluation_domain& d = *(evaluation_domain*)h;
    fr::field_t* c = (fr::field_t*)coeffs;
    fr::field_t k = constant ? ldr(constant) : fr::one;
    switch (op)
    {
    case 0: polynomial_arithmetic::fft(c, d); break;
    case 1: polynomial_arithmetic::ifft(c, d); break;
    case 2: polynomial_arithmetic::coset_fft(c, d); break;
    case 3: polynomial_arithmetic::coset_ifft(c, d); break;
    case 4: polynomial_arithmetic::fft_with_constant(c, d, k); break;
    case 5: polynomial_arithmetic::ifft_with_constant(c, d, k); break;
    case 6: polynomial_arithmetic::coset_fft_with_constant(c, d, k); break;
    }
}
// Horner evaluation at z (polynomial_arithmetic.cpp evaluate) — used by the n=16 fft test
void ref_poly_evaluate(const uint64_t* coeffs, const uint64_t* z, size_t n, uint64_t* out)
{
    fr::field_t* c = aligned_copy<fr::field_t>(coeffs, n);
    fr::field_t zz = ldr(z);
    fr::field_t r = polynomial_arithmetic::evaluate(c, zz, n);
    memcpy(out, &r, 32);
    free(c);
}

// polynomial_arithmetic.cpp:381-476; l_1 must hold target_n elements (caller-aligned)
void ref_compute_lagrange_polynomial_fft(uint64_t* l_1, size_t src_n, size_t target_n)
{
    evaluation_domain s(src_n), t(target_n);
    polynomial_arithmetic::compute_lagrange_polynomial_fft((fr::field_t*)l_1, s, t);
}

// polynomial_arithmetic.cpp:478-560; coeffs holds target_n elements (caller-aligned), in place
void ref_divide_by_pseudo_vanishing_polynomial(uint64_t* coeffs, size_t src_n, size_t target_n)
{
    evaluation_domain s(src_n), t(target_n);
    polynomial_arithmetic::divide_by_pseudo_vanishing_polynomial((fr::field_t*)coeffs, s, t);
}
// polynomial_arithmetic.cpp:562-591; returns F(z) in f_out, the quotient in dest (caller-aligned, n elements, lazily reduced)
void ref_compute_kate_opening_coefficients(const uint64_t* src, uint64_t* dest, const uint64_t* z, size_t n, uint64_t* f_out)
{
    fr::field_t zz = ldr(z);
    fr::field_t f = polynomial_arithmetic::compute_kate_opening_coefficients((const fr::field_t*)src, (fr::field_t*)dest, zz, n);
    memcpy(f_out, &f, 32);
}

void* ref_aligned_alloc(size_t bytes) { return aligned_alloc(64, (bytes + 63) & ~(size_t)63); }
void ref_aligned_free(void* p) { free(p); }

} // extern "C"
#pragma GCC visibility pop
