// Drop-in replacement for waffle::Prover::construct_proof / Prover::reset of the reference's
//   src/barretenberg/waffle/proof_system/prover/prover.cpp:657-690
// (SURVEY.md §8f rows 1-3).  Same class, same signatures (prover.hpp:41-42): the prover links against this file instead
// of the reference's two member functions, everything else in prover.cpp keeps its reference body.
//
// For circuits built from the reference's four widgets — arithmetic, bool, MiMC, sequential, each at most once, i.e.
// everything the Standard / Bool / MiMC / Extended composers build (standard_composer.cpp:199-218, bool_composer.cpp:123-124,
// mimc_composer.cpp:225-226, extended_composer.cpp:512-514) — the whole proof is computed on the GPU with every
// polynomial resident in HBM (include/bbgpu.h bbg_plonk_*, barretenberg_b200/csrc/bbg_plonk.cu).  What stays here is exactly what the reference
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
#include <barretenberg/waffle/proof_system/widgets/bool_widget.hpp>
#include <barretenberg/waffle/proof_system/widgets/mimc_widget.hpp>
#include <barretenberg/waffle/proof_system/widgets/sequential_widget.hpp>

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
    check(bbg_shim::ensure_library(), "bbg_init");
}

// one device-side prover per circuit size, kept for the life of the process (2.9 GB of HBM at n = 2^20)
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
    // ---- which widgets? -------------------------------------------------------------------------------------------
    int kinds[4];
    const uint64_t* selectors[11];
    int num_selectors = 0;
    int first_selector[4] = { 0, 0, 0, 0 };
    bool seen[4] = { false, false, false, false };
    bool known = widgets.size() >= 1 && widgets.size() <= 4;
    auto take = [&](const polynomial& q) {
        known = known && q.get_size() >= n;
        selectors[num_selectors++] = (const uint64_t*)q.get_coefficients();
    };
    for (size_t i = 0; known && i < widgets.size(); ++i)
    {
        ProverBaseWidget* w = widgets[i].get();
        first_selector[i] = num_selectors;
        if (auto* a = dynamic_cast<ProverArithmeticWidget*>(w))
        {
            kinds[i] = BBG_WIDGET_ARITHMETIC;
            take(a->q_m), take(a->q_l), take(a->q_r), take(a->q_o), take(a->q_c);
        }
        else if (auto* bw = dynamic_cast<ProverBoolWidget*>(w))
        {
            kinds[i] = BBG_WIDGET_BOOL;
            take(bw->q_bl), take(bw->q_br), take(bw->q_bo);
        }
        else if (auto* m = dynamic_cast<ProverMiMCWidget*>(w))
        {
            kinds[i] = BBG_WIDGET_MIMC;
            take(m->q_mimc_selector), take(m->q_mimc_coefficient);
        }
        else if (auto* sq = dynamic_cast<ProverSequentialWidget*>(w))
        {
            kinds[i] = BBG_WIDGET_SEQUENTIAL;
            take(sq->q_o_next);
        }
        else
        {
            known = false;
            break;
        }
        known = known && !seen[kinds[i]];
        seen[kinds[i]] = true;
    }
    // (prover.cpp:445-454) only the output wire is ever needed at z w by the reference's widgets
    bool needs_w_l_shifted = false, needs_w_r_shifted = false, needs_w_o_shifted = false;
    for (size_t i = 0; i < widgets.size(); ++i)
    {
        needs_w_l_shifted |= widgets[i]->version.has_dependency(WidgetVersionControl::Dependencies::REQUIRES_W_L_SHIFTED);
        needs_w_r_shifted |= widgets[i]->version.has_dependency(WidgetVersionControl::Dependencies::REQUIRES_W_R_SHIFTED);
        needs_w_o_shifted |= widgets[i]->version.has_dependency(WidgetVersionControl::Dependencies::REQUIRES_W_O_SHIFTED);
    }
    const char* mode = getenv("BBG_PLONK_RESIDENT");
    const bool resident = known && !needs_w_l_shifted && !needs_w_r_shifted && n >= 4 && n <= ((size_t)1 << 23) && (n & (n - 1)) == 0 &&
                          !(mode != nullptr && mode[0] == '0') && w_l.get_size() >= n && w_r.get_size() >= n && w_o.get_size() >= n &&
                          sigma_1_mapping.size() >= n && sigma_2_mapping.size() >= n && sigma_3_mapping.size() >= n;
    if (!resident) return bbg_shim::reference_construct_proof(*this);

    init_library();
    bbg_shim::Timer timer("construct_proof(resident)");
    const evaluation_domain& domain = circuit_state.small_domain;
    bbg_plonk_prover* dev = device_prover(domain.log2_size);

    // ---- inputs (queued; the copies run behind round 1) ------------------------------------------------------------
    check(bbg_plonk_set_srs(dev, (const uint64_t*)reference_string.monomials, n), "bbg_plonk_set_srs");
    check(bbg_plonk_set_witness(dev, (const uint64_t*)w_l.get_coefficients(), (const uint64_t*)w_r.get_coefficients(),
                                (const uint64_t*)w_o.get_coefficients()),
          "bbg_plonk_set_witness");
    check(bbg_plonk_set_permutation(dev, sigma_1_mapping.data(), sigma_2_mapping.data(), sigma_3_mapping.data()), "bbg_plonk_set_permutation");
    check(bbg_plonk_set_widgets(dev, kinds, (int)widgets.size(), selectors), "bbg_plonk_set_widgets");

    // ---- round 1: wire commitments (prover.cpp:65-89, :126-135) --------------------------------------------------
    uint64_t pts[3 * 12];
    check(timed("bbg_plonk_round_wires", [&]() { return bbg_plonk_round_wires(dev, pts); }), "bbg_plonk_round_wires");
    copy_xy(pts, proof.W_L);
    copy_xy(pts + 12, proof.W_R);
    copy_xy(pts + 24, proof.W_O);
    challenges.gamma = compute_gamma(proof);
    challenges.beta = compute_beta(proof, challenges.gamma);

    // ---- round 2: grand product (:137-225, :91-107) ---------------------------------------------------------------
    check(timed("bbg_plonk_round_grand_product", [&]() { return bbg_plonk_round_grand_product(dev, challenges.beta.data, challenges.gamma.data, pts); }),
          "bbg_plonk_round_grand_product");
    copy_xy(pts, proof.Z_1);
    challenges.alpha = compute_alpha(proof);

    // ---- round 3: quotient (:227-463, :109-124) -------------------------------------------------------------------
    fr::field_t alpha_base = fr::sqr(fr::sqr(challenges.alpha));
    fr::mul(challenges.alpha, alpha_base); // (prover.cpp:437 discards this product: the widgets start at alpha^4)
    check(timed("bbg_plonk_round_quotient",
                [&]() { return bbg_plonk_round_quotient(dev, challenges.beta.data, challenges.gamma.data, challenges.alpha.data, alpha_base.data, pts); }),
          "bbg_plonk_round_quotient");
    to_affine(pts, proof.T_LO);
    to_affine(pts + 12, proof.T_MID);
    to_affine(pts + 24, proof.T_HI);
    challenges.z = compute_evaluation_challenge(proof);

    // ---- round 4: evaluations and the linearisation polynomial (:465-503) ------------------------------------------
    fr::field_t beta_inv;
    fr::__invert(challenges.beta, beta_inv);
    fr::field_t shifted_z;
    fr::__mul(challenges.z, domain.root, shifted_z);
    uint64_t evals[9 * 4];
    check(timed("bbg_plonk_round_evaluations", [&]() { return bbg_plonk_round_evaluations(dev, challenges.z.data, shifted_z.data, evals); }),
          "bbg_plonk_round_evaluations");
    load(evals, proof.w_l_eval);
    load(evals + 4, proof.w_r_eval);
    load(evals + 8, proof.w_o_eval);
    load(evals + 12, proof.sigma_1_eval);
    load(evals + 16, proof.sigma_2_eval);
    load(evals + 20, proof.z_1_shifted_eval);
    fr::field_t t_eval;
    load(evals + 24, t_eval);
    if (needs_w_o_shifted) load(evals + 28, proof.w_o_shifted_eval);          // (:461-463)
    if (seen[BBG_WIDGET_MIMC]) load(evals + 32, proof.q_mimc_coefficient_eval); // (mimc_widget.cpp:91-94, called from :469-472)
    // we scaled the sigma polynomials up by beta, so scale back down (:475-477)
    fr::__mul(proof.sigma_1_eval, beta_inv, proof.sigma_1_eval);
    fr::__mul(proof.sigma_2_eval, beta_inv, proof.sigma_2_eval);

    polynomial_arithmetic::lagrange_evaluations lagrange_evals = polynomial_arithmetic::get_lagrange_evaluations(challenges.z, domain);
    plonk_linear_terms linear_terms = compute_linear_terms(proof, challenges, lagrange_evals.l_1, n);
    fr::field_t scalars[2 + 11];
    scalars[0] = linear_terms.z_1;
    fr::__mul(linear_terms.sigma_3, beta_inv, scalars[1]); // (:488-490)
    {
        // the widgets' compute_linear_contribution, with the sum over i left to the device: every selector gets the
        // scalar it is multiplied with; alpha_base is chained through the widgets exactly as prover.cpp:495-499 does
        fr::field_t ab = fr::sqr(fr::sqr(challenges.alpha));
        const fr::field_t& alpha = challenges.alpha;
        for (size_t i = 0; i < widgets.size(); ++i)
        {
            fr::field_t* s = &scalars[2 + first_selector[i]];
            switch (kinds[i])
            {
            case BBG_WIDGET_ARITHMETIC: // arithmetic_widget.cpp:101-121
                s[0] = fr::mul(fr::mul(proof.w_l_eval, proof.w_r_eval), ab);
                s[1] = fr::mul(proof.w_l_eval, ab);
                s[2] = fr::mul(proof.w_r_eval, ab);
                s[3] = fr::mul(proof.w_o_eval, ab);
                s[4] = ab;
                ab = fr::mul(ab, alpha);
                break;
            case BBG_WIDGET_BOOL: // bool_widget.cpp:106-121
                s[0] = fr::mul(fr::sub(fr::sqr(proof.w_l_eval), proof.w_l_eval), ab);
                s[1] = fr::mul(fr::mul(fr::sub(fr::sqr(proof.w_r_eval), proof.w_r_eval), ab), alpha);
                s[2] = fr::mul(fr::mul(fr::sub(fr::sqr(proof.w_o_eval), proof.w_o_eval), ab), fr::sqr(alpha));
                ab = fr::mul(ab, fr::mul(fr::sqr(alpha), alpha));
                break;
            case BBG_WIDGET_MIMC: // mimc_widget.cpp:96-110
            {
                fr::field_t mimc_T0 = fr::add(fr::add(proof.w_o_eval, proof.w_l_eval), proof.q_mimc_coefficient_eval);
                fr::field_t mimc_a = fr::sqr(mimc_T0);
                mimc_a = fr::mul(mimc_a, mimc_T0);
                mimc_a = fr::sub(mimc_a, proof.w_r_eval);
                fr::field_t mimc_term = fr::mul(fr::sub(fr::mul(fr::sqr(proof.w_r_eval), mimc_T0), proof.w_o_shifted_eval), alpha);
                mimc_term = fr::mul(fr::add(mimc_term, mimc_a), ab);
                s[0] = mimc_term;  // x q_mimc_selector
                s[1] = fr::zero;   // q_mimc_coefficient does not enter r(X)
                ab = fr::mul(ab, fr::sqr(alpha));
                break;
            }
            case BBG_WIDGET_SEQUENTIAL: // sequential_widget.cpp:64-75
            {
                fr::field_t old_alpha = fr::mul(ab, fr::invert(alpha));
                s[0] = fr::mul(proof.w_o_shifted_eval, old_alpha);
                break;
            }
            }
        }
    }
    check(timed("bbg_plonk_round_linearise", [&]() { return bbg_plonk_round_linearise(dev, (const uint64_t*)scalars, challenges.z.data, proof.linear_eval.data); }),
          "bbg_plonk_round_linearise");

    // ---- round 5: opening proofs (:505-655) -----------------------------------------------------------------------
    challenges.nu = compute_linearisation_challenge(proof, t_eval);
    fr::field_t nu_powers[8];
    fr::__copy(challenges.nu, nu_powers[0]);
    for (size_t i = 1; i < 8; ++i) fr::__mul(nu_powers[i - 1], nu_powers[0], nu_powers[i]);
    fr::field_t wire_shift[3] = { fr::zero, fr::zero, fr::zero };
    fr::field_t selector_terms[11];
    for (int k = 0; k < 11; ++k) selector_terms[k] = fr::zero;
    {
        fr::field_t nu_base = nu_powers[7]; // (:593)
        if (needs_w_o_shifted)              // (:597-631; w_l / w_r are never requested by the reference's widgets)
        {
            wire_shift[2] = nu_base;
            nu_base = fr::mul(nu_base, challenges.nu);
        }
        for (size_t i = 0; i < widgets.size(); ++i) // (:633-636)
        {
            if (kinds[i] == BBG_WIDGET_MIMC) // mimc_widget.cpp:112-120: poly += q_mimc_coefficient * nu_base
            {
                selector_terms[first_selector[i] + 1] = nu_base;
                nu_base = fr::mul(nu_base, nu_powers[0]);
            }
        }
    }
    check(timed("bbg_plonk_round_openings",
                [&]() {
                    return bbg_plonk_round_openings(dev, (const uint64_t*)nu_powers, beta_inv.data, challenges.z.data, shifted_z.data,
                                                    (const uint64_t*)wire_shift, (const uint64_t*)selector_terms, pts);
                }),
          "bbg_plonk_round_openings");
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
