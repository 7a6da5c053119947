// Drop-in replacement for the radix-2 NTT entry points of the reference's
//   src/barretenberg/polynomials/polynomial_arithmetic.cpp
// Same namespace and signatures (polynomial_arithmetic.hpp:28-39): fft, ifft, fft_with_constant, ifft_with_constant,
// coset_fft, coset_fft_with_constant, coset_ifft — in place on `coeffs`, length domain.size, natural order in and
// out, canonical outputs.  barretenberg::polynomial (polynomial.cpp:246-313), the prover and the widgets call these.
//
// How to link: build the reference's polynomial_arithmetic.cpp with
//   -Dfft=cpu_reference_fft -Difft=cpu_reference_ifft -Dfft_with_constant=cpu_reference_fft_with_constant
//   -Difft_with_constant=cpu_reference_ifft_with_constant -Dcoset_fft=cpu_reference_coset_fft
//   -Dcoset_fft_with_constant=cpu_reference_coset_fft_with_constant -Dcoset_ifft=cpu_reference_coset_ifft
//   -Dcompute_lagrange_polynomial_fft=cpu_reference_compute_lagrange_polynomial_fft
//   -Devaluate=cpu_reference_evaluate -Ddivide_by_pseudo_vanishing_polynomial=cpu_reference_divide_by_pseudo_vanishing_polynomial
//   -Dcompute_kate_opening_coefficients=cpu_reference_compute_kate_opening_coefficients
// (copy_polynomial, add, mul, compress_fft, get_lagrange_evaluations, ... keep their reference bodies: memory-bound or
// scalar host code that a PCIe round trip would only slow down) and add this file for the names above.
// evaluation_domain.cpp is unchanged: only domain.log2_size crosses the boundary; root, 1/n, the coset generator
// and every twiddle are derived on the device from the curve constants.
#include <cstdio>
#include <cstdlib>

#include <barretenberg/polynomials/polynomial_arithmetic.hpp>

#include "bbgpu.h"
#include "shim_stats.h"

namespace
{
void run(barretenberg::fr::field_t* coeffs, const barretenberg::evaluation_domain& domain, int op, const barretenberg::fr::field_t* constant,
         const char* what)
{
    int e = bbg_shim::ensure_library();
    bbg_shim::Timer timer(what);
    if (e == 0) e = bbg_ntt_fr((uint64_t*)coeffs, (unsigned)domain.log2_size, op, (const uint64_t*)constant);
    if (e != 0)
    {
        fprintf(stderr, "bbgpu shim: %s failed: %s (no CPU fallback)\n", what, bbg_error_string(e));
        abort();
    }
}
} // namespace

namespace barretenberg
{
namespace polynomial_arithmetic
{
void fft(fr::field_t* coeffs, const evaluation_domain& domain) { run(coeffs, domain, BBG_FFT, nullptr, "fft"); }
void ifft(fr::field_t* coeffs, const evaluation_domain& domain) { run(coeffs, domain, BBG_IFFT, nullptr, "ifft"); }
void fft_with_constant(fr::field_t* coeffs, const evaluation_domain& domain, const fr::field_t& value)
{
    run(coeffs, domain, BBG_FFT_WITH_CONSTANT, &value, "fft_with_constant");
}
void ifft_with_constant(fr::field_t* coeffs, const evaluation_domain& domain, const fr::field_t& value)
{
    run(coeffs, domain, BBG_IFFT_WITH_CONSTANT, &value, "ifft_with_constant");
}
void coset_fft(fr::field_t* coeffs, const evaluation_domain& domain) { run(coeffs, domain, BBG_COSET_FFT, nullptr, "coset_fft"); }
void coset_fft_with_constant(fr::field_t* coeffs, const evaluation_domain& domain, const fr::field_t& constant)
{
    run(coeffs, domain, BBG_COSET_FFT_WITH_CONSTANT, &constant, "coset_fft_with_constant");
}
void coset_ifft(fr::field_t* coeffs, const evaluation_domain& domain) { run(coeffs, domain, BBG_COSET_IFFT, nullptr, "coset_ifft"); }

// polynomial_arithmetic.cpp:381-476 (first widening into SURVEY.md §8f): output-only, so no upload is needed.
// Build the reference file with -Dcompute_lagrange_polynomial_fft=cpu_reference_compute_lagrange_polynomial_fft as well.
void compute_lagrange_polynomial_fft(fr::field_t* l_1_coefficients, const evaluation_domain& src_domain, const evaluation_domain& target_domain)
{
    int e = bbg_shim::ensure_library();
    bbg_shim::Timer timer("compute_lagrange_polynomial_fft");
    if (e == 0) e = bbg_compute_lagrange_polynomial_fft((uint64_t*)l_1_coefficients, (unsigned)src_domain.log2_size, (unsigned)target_domain.log2_size);
    if (e != 0)
    {
        fprintf(stderr, "bbgpu shim: compute_lagrange_polynomial_fft failed: %s (no CPU fallback)\n", bbg_error_string(e));
        abort();
    }
}

// ---- the three heavy stand-alone helpers -------------------------------------------------------
// Callers: polynomial::evaluate (polynomial.cpp), prover.cpp:443-444 / :638-640 and the widgets in the reference's round
// structure.  Build the reference file additionally with
//   -Devaluate=cpu_reference_evaluate -Ddivide_by_pseudo_vanishing_polynomial=cpu_reference_divide_by_pseudo_vanishing_polynomial
//   -Dcompute_kate_opening_coefficients=cpu_reference_compute_kate_opening_coefficients
namespace
{
int ensure_library() { return bbg_shim::ensure_library(); }
void die(const char* what, int e)
{
    fprintf(stderr, "bbgpu shim: %s failed: %s (no CPU fallback)\n", what, bbg_error_string(e));
    abort();
}
} // namespace

// polynomial_arithmetic.cpp:337-373
fr::field_t evaluate(const fr::field_t* coeffs, const fr::field_t& z, const size_t n)
{
    int e = ensure_library();
    bbg_shim::Timer timer("evaluate");
    fr::field_t r;
    if (e == 0) e = bbg_fr_evaluate((const uint64_t*)coeffs, n, z.data, r.data);
    if (e != 0) die("evaluate", e);
    return r;
}

// polynomial_arithmetic.cpp:478-560
void divide_by_pseudo_vanishing_polynomial(fr::field_t* coeffs, const evaluation_domain& src_domain, const evaluation_domain& target_domain)
{
    int e = ensure_library();
    bbg_shim::Timer timer("divide_by_pseudo_vanishing_polynomial");
    if (e == 0) e = bbg_fr_divide_by_pseudo_vanishing_polynomial((uint64_t*)coeffs, (unsigned)src_domain.log2_size, (unsigned)target_domain.log2_size);
    if (e != 0) die("divide_by_pseudo_vanishing_polynomial", e);
}

// polynomial_arithmetic.cpp:562-591 (dest may alias src: polynomial::compute_kate_opening_coefficients passes its own buffer twice)
fr::field_t compute_kate_opening_coefficients(const fr::field_t* src, fr::field_t* dest, const fr::field_t& z, const size_t n)
{
    int e = ensure_library();
    bbg_shim::Timer timer("compute_kate_opening_coefficients");
    fr::field_t f;
    if (e == 0) e = bbg_fr_compute_kate_opening_coefficients((const uint64_t*)src, (uint64_t*)dest, z.data, n, f.data);
    if (e != 0) die("compute_kate_opening_coefficients", e);
    return f;
}
} // namespace polynomial_arithmetic
} // namespace barretenberg
