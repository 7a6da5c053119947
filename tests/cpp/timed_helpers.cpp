// DEVELOPMENT PROFILING AID (build/prover_gpu_prof only): times the element-wise polynomial helpers that still run
// the reference's CPU bodies (SURVEY.md §8f "next" rows), so a prove can be split into boundary / helpers / rest.
#include <barretenberg/polynomials/polynomial_arithmetic.hpp>

#include "shim_stats.h"

namespace barretenberg
{
namespace polynomial_arithmetic
{
fr::field_t cpu_reference_evaluate(const fr::field_t* coeffs, const fr::field_t& z, const size_t n);
void cpu_reference_divide_by_pseudo_vanishing_polynomial(fr::field_t* coeffs, const evaluation_domain& src_domain, const evaluation_domain& target_domain);
fr::field_t cpu_reference_compute_kate_opening_coefficients(const fr::field_t* src, fr::field_t* dest, const fr::field_t& z, const size_t n);
void cpu_reference_copy_polynomial(fr::field_t* src, fr::field_t* dest, size_t num_src_coefficients, size_t num_target_coefficients);

fr::field_t evaluate(const fr::field_t* coeffs, const fr::field_t& z, const size_t n)
{
    bbg_shim::Timer t("cpu:evaluate");
    return cpu_reference_evaluate(coeffs, z, n);
}
void divide_by_pseudo_vanishing_polynomial(fr::field_t* c, const evaluation_domain& s, const evaluation_domain& d)
{
    bbg_shim::Timer t("cpu:divide_by_pseudo_vanishing");
    cpu_reference_divide_by_pseudo_vanishing_polynomial(c, s, d);
}
fr::field_t compute_kate_opening_coefficients(const fr::field_t* src, fr::field_t* dest, const fr::field_t& z, const size_t n)
{
    bbg_shim::Timer t("cpu:compute_kate_opening_coeffs");
    return cpu_reference_compute_kate_opening_coefficients(src, dest, z, n);
}
void copy_polynomial(fr::field_t* src, fr::field_t* dest, size_t a, size_t b)
{
    bbg_shim::Timer t("cpu:copy_polynomial");
    cpu_reference_copy_polynomial(src, dest, a, b);
}
} // namespace polynomial_arithmetic
} // namespace barretenberg
