// Drop-in replacement for waffle::Prover::construct_proof / Prover::reset of the reference's
//   src/barretenberg/waffle/proof_system/prover/prover.cpp:657-690
// (SURVEY.md §8f rows 1-3).  Same class, same signatures (prover.hpp:41-42): the prover links against this file instead
// of the reference's two member functions, everything else in prover.cpp keeps its reference body.
//
// For circuits whose only widget is the arithmetic widget (what StandardComposer::preprocess builds,
// standard_composer.cpp:199-218) the whole proof is computed on the GPU with every polynomial resident in HBM
// (include/bbgpu.h bbg_plonk_*, barretenberg_b200/csrc/bbg_plonk.cu).  What stays here is exactly what the reference
// does with scalars: the Fiat-Shamir transcript (challenge.hpp), compute_linear_terms (linearizer.hpp),
// get_lagrange_evaluations — the reference's own header / library code, called unchanged.
// Any other widget mix runs the reference's round structure (its construct_proof, renamed), whose MSM and NTT calls
// still go to the GPU through scalar_multiplication_gpu.cpp / polynomial_arithmetic_gpu.cpp.  There is no CPU path
// for the hot loops in either case; a CUDA failure aborts.
//
// How to link: build the reference's prover.cpp with
//   -Dconstruct_proof=cpu_reference_construct_proof -Dreset=cpu_reference_reset
// build prover_gpu_glue.cpp with the same two defines (it forwards to the renamed bodies), and add this file.
#include <cstdio>
#include <cstdlib>

#include <barretenberg/polynomials/polynomial_arithmetic.hpp>
#include <barretenberg/waffle/proof_system/challenge.hpp>
#include <barretenberg/waffle/proof_system/linearizer.hpp>
#include <barretenberg/waffle/proof_system/prover/prover.hpp>
#include <barretenberg/waffle/proof_system/widgets/arithmetic_widget.hpp>

#include "bbgpu.h"
#include "shim_stats.h"

namespace bbg_shim
{
// prover_gpu_glue.cpp: the reference's own (renamed) member functions
waffle::plonk_proof reference_construct_proof(waffle::Prover& prover);
void reference_reset(waffle::Prover& prover);
} // namespace bbg_shim

using namespace barretenberg;

namespace
{
void check(int e, const char* what)
{
    if (e != 0)
    {
        fprintf(stderr, "bbgpu shim: %s failed: %s (no CPU fallback)\n", what, bbg_error_string(e));
        abort();
    }
}

template <typename Fn> int timed(const char* name, Fn fn)
{
    bbg_shim::Timer t(name);
    return fn();
}

void init_library()
{
    static bool ready = false;
    if (ready) return;
    const char* dev = getenv("BBG_DEVICE");
    check(bbg_init(dev ? atoi(dev) : 0), "bbg_init");
    bbg_set_auto_srs_cache(1);
    bbg_shim::stats().after_init();
    ready = true;
}

// one device-side prover per circuit size, kept for the life of the process (2.1 GB of HBM at n = 2^20)
bbg_plonk_prover* device_prover(size_t log2_n)
{
    static bbg_plonk_prover* cached = nullptr;
    static size_t cached_log = 0;
    if (cached != nullptr && cached_log == log2_n) return cached;
    if (cached != nullptr) bbg_plonk_destroy(cached);
    cached = nullptr;
    check(bbg_plonk_create((unsigned)log2_n, &cached), "bbg_plonk_create");
    cached_log = log2_n;
    return cached;
}

// the normalised Jacobian the MSM returns -> affine proof element, exactly as prover.cpp does it
void to_affine(const uint64_t* xyz, g1::affine_element& out)
{
    g1::element p;
    for (int i = 0; i < 4; ++i)
    {
        p.x.data[i] = xyz[i];
        p.y.data[i] = xyz[4 + i];
        p.z.data[i] = xyz[8 + i];
    }
    g1::jacobian_to_affine(p, out); // prover.cpp:119-121, :651-652
}
void copy_xy(const uint64_t* xyz, g1::affine_element& out)
{
    for (int i = 0; i < 4; ++i) // prover.cpp:76-81, :103-104 (fq::__copy of x and y)
    {
        out.x.data[i] = xyz[i];
        out.y.data[i] = xyz[4 + i];
    }
}
void load(const uint64_t* limbs, fr::field_t& out)
{
    for (int i = 0; i < 4; ++i) out.data[i] = limbs[i];
}
} // namespace

namespace waffle
{
plonk_proof Prover::construct_proof()
{
    ProverArithmeticWidget* arith = widgets.size() == 1 ? dynamic_cast<ProverArithmeticWidget*>(widgets[0].get()) : nullptr;
    const char* mode = getenv("BBG_PLONK_RESIDENT");
    const bool resident = arith != nullptr && n >= 4 && n <= ((size_t)1 << 20) && (n & (n - 1)) == 0 && !(mode != nullptr && mode[0] == '0') &&
                          w_l.get_size() >= n && w_r.get_size() >= n && w_o.get_size() >= n && sigma_1_mapping.size() >= n &&
                          sigma_2_mapping.size() >= n && sigma_3_mapping.size() >= n && arith->q_m.get_size() >= n;
    if (!resident) return bbg_shim::reference_construct_proof(*this);

    init_library();
    bbg_shim::Timer timer("construct_proof(resident)");
    const evaluation_domain& domain = circuit_state.small_domain;
    bbg_plonk_prover* dev = device_prover(domain.log2_size);

    // ---- inputs -------------------------------------------------------------------------------------------------
    check(timed("bbg_plonk_set_srs", [&]() { return bbg_plonk_set_srs(dev, (const uint64_t*)reference_string.monomials, n); }), "bbg_plonk_set_srs");
    check(timed("bbg_plonk_set_witness", [&]() { return bbg_plonk_set_witness(dev, (const uint64_t*)w_l.get_coefficients(), (const uint64_t*)w_r.get_coefficients(),
                                (const uint64_t*)w_o.get_coefficients()); }), "bbg_plonk_set_witness");
    check(timed("bbg_plonk_set_permutation", [&]() { return bbg_plonk_set_permutation(dev, sigma_1_mapping.data(), sigma_2_mapping.data(), sigma_3_mapping.data()); }), "bbg_plonk_set_permutation");
    check(timed("bbg_plonk_set_arithmetic_selectors", [&]() { return bbg_plonk_set_arithmetic_selectors(dev, (const uint64_t*)arith->q_m.get_coefficients(), (const uint64_t*)arith->q_l.get_coefficients(),
                                             (const uint64_t*)arith->q_r.get_coefficients(), (const uint64_t*)arith->q_o.get_coefficients(),
                                             (const uint64_t*)arith->q_c.get_coefficients()); }), "bbg_plonk_set_arithmetic_selectors");

    // ---- round 1: wire commitments (prover.cpp:65-89, :126-135) --------------------------------------------------
    uint64_t pts[3 * 12];
    check(timed("bbg_plonk_round_wires", [&]() { return bbg_plonk_round_wires(dev, pts); }), "bbg_plonk_round_wires");
    copy_xy(pts, proof.W_L);
    copy_xy(pts + 12, proof.W_R);
    copy_xy(pts + 24, proof.W_O);
    challenges.gamma = compute_gamma(proof);
    challenges.beta = compute_beta(proof, challenges.gamma);

    // ---- round 2: grand product (:137-225, :91-107) ---------------------------------------------------------------
    check(timed("bbg_plonk_round_grand_product", [&]() { return bbg_plonk_round_grand_product(dev, challenges.beta.data, challenges.gamma.data, pts); }), "bbg_plonk_round_grand_product");
    copy_xy(pts, proof.Z_1);
    challenges.alpha = compute_alpha(proof);

    // ---- round 3: quotient (:227-463, :109-124) -------------------------------------------------------------------
    fr::field_t alpha_base = fr::sqr(fr::sqr(challenges.alpha));
    fr::mul(challenges.alpha, alpha_base); // (prover.cpp:437 discards this product: the widgets start at alpha^4)
    check(timed("bbg_plonk_round_quotient", [&]() { return bbg_plonk_round_quotient(dev, challenges.beta.data, challenges.gamma.data, challenges.alpha.data, alpha_base.data, pts); }), "bbg_plonk_round_quotient");
    to_affine(pts, proof.T_LO);
    to_affine(pts + 12, proof.T_MID);
    to_affine(pts + 24, proof.T_HI);
    challenges.z = compute_evaluation_challenge(proof);

    // ---- round 4: evaluations and the linearisation polynomial (:465-503) ------------------------------------------
    fr::field_t beta_inv;
    fr::__invert(challenges.beta, beta_inv);
    fr::field_t shifted_z;
    fr::__mul(challenges.z, domain.root, shifted_z);
    uint64_t evals[7 * 4];
    check(timed("bbg_plonk_round_evaluations", [&]() { return bbg_plonk_round_evaluations(dev, challenges.z.data, shifted_z.data, evals); }), "bbg_plonk_round_evaluations");
    load(evals, proof.w_l_eval);
    load(evals + 4, proof.w_r_eval);
    load(evals + 8, proof.w_o_eval);
    load(evals + 12, proof.sigma_1_eval);
    load(evals + 16, proof.sigma_2_eval);
    load(evals + 20, proof.z_1_shifted_eval);
    fr::field_t t_eval;
    load(evals + 24, t_eval);
    // we scaled the sigma polynomials up by beta, so scale back down (:475-477)
    fr::__mul(proof.sigma_1_eval, beta_inv, proof.sigma_1_eval);
    fr::__mul(proof.sigma_2_eval, beta_inv, proof.sigma_2_eval);

    polynomial_arithmetic::lagrange_evaluations lagrange_evals = polynomial_arithmetic::get_lagrange_evaluations(challenges.z, domain);
    plonk_linear_terms linear_terms = compute_linear_terms(proof, challenges, lagrange_evals.l_1, n);
    fr::field_t scalars[7];
    scalars[0] = linear_terms.z_1;
    fr::__mul(linear_terms.sigma_3, beta_inv, scalars[1]); // (:488-490)
    scalars[2] = fr::mul(proof.w_l_eval, proof.w_r_eval);  // arithmetic_widget.cpp:101
    scalars[3] = proof.w_l_eval;
    scalars[4] = proof.w_r_eval;
    scalars[5] = proof.w_o_eval;
    scalars[6] = fr::sqr(fr::sqr(challenges.alpha)); // prover.cpp:495
    check(timed("bbg_plonk_round_linearise", [&]() { return bbg_plonk_round_linearise(dev, (const uint64_t*)scalars, challenges.z.data, proof.linear_eval.data); }), "bbg_plonk_round_linearise");

    // ---- round 5: opening proofs (:505-655) -----------------------------------------------------------------------
    challenges.nu = compute_linearisation_challenge(proof, t_eval);
    fr::field_t nu_powers[8];
    fr::__copy(challenges.nu, nu_powers[0]);
    for (size_t i = 1; i < 8; ++i) fr::__mul(nu_powers[i - 1], nu_powers[0], nu_powers[i]);
    check(timed("bbg_plonk_round_openings", [&]() { return bbg_plonk_round_openings(dev, (const uint64_t*)nu_powers, beta_inv.data, challenges.z.data, shifted_z.data, pts); }), "bbg_plonk_round_openings");
    to_affine(pts, proof.PI_Z);
    to_affine(pts + 12, proof.PI_Z_OMEGA);
    return proof;
}

void Prover::reset()
{
    // The resident path never transforms the host polynomials (w_l, w_r, w_o and the selectors stay in Lagrange form),
    // so there is nothing to undo.  The reference's round structure leaves a copy of the witness in
    // circuit_state.w_l_fft (prover.cpp:128, :400) until its own reset() empties it (:678): that is the marker.
    if (circuit_state.w_l_fft.get_size() != 0) bbg_shim::reference_reset(*this);
}
} // namespace waffle
