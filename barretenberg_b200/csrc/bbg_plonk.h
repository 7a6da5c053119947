// Internal C++ interface of the HBM-resident PLONK prover rounds (bbg_plonk.cu); the C ABI wrappers live in
// bbg_capi.cu (bbg_plonk_* in include/bbgpu.h).
#pragma once
#include "bbg_rt.h"

namespace bbg
{
namespace plonk
{
enum widget_kind // == bbg_plonk_widget in include/bbgpu.h
{
    WIDGET_ARITHMETIC = 0,
    WIDGET_BOOL = 1,
    WIDGET_MIMC = 2,
    WIDGET_SEQUENTIAL = 3,
};
struct Prover;
int create(unsigned log_n, Prover** out);
void destroy(Prover* p);
int set_witness(Prover* p, const uint64_t* w_l, const uint64_t* w_r, const uint64_t* w_o, cudaStream_t st);
int set_permutation(Prover* p, const uint32_t* m1, const uint32_t* m2, const uint32_t* m3, cudaStream_t st);
int set_widgets(Prover* p, const int* kinds, int count, const uint64_t* const* selectors_lagrange, cudaStream_t st);
int set_srs(Prover* p, const void* d_table);
int round_wires(Prover* p, uint64_t* out_xyz, cudaStream_t st);
int round_grand_product(Prover* p, const uint64_t* beta, const uint64_t* gamma, uint64_t out_xyz[12], cudaStream_t st);
int round_quotient(Prover* p, const uint64_t* beta, const uint64_t* gamma, const uint64_t* alpha, const uint64_t* alpha_base, uint64_t* out_xyz,
                   cudaStream_t st);
int round_evaluations(Prover* p, const uint64_t* zeta, const uint64_t* zeta_omega, uint64_t* out, cudaStream_t st);
int round_linearise(Prover* p, const uint64_t* scalars, const uint64_t* zeta, uint64_t out_eval[4], cudaStream_t st);
int round_openings(Prover* p, const uint64_t* nu_powers, const uint64_t* beta_inv, const uint64_t* zeta, const uint64_t* zeta_omega,
                   const uint64_t* wire_shift, const uint64_t* selector_terms, uint64_t* out_xyz, cudaStream_t st);
// stand-alone helpers on device buffers (polynomial_arithmetic.cpp:337-373, :478-560, :562-591)
int evaluate_device(const void* d_poly, size_t len, const uint64_t* z, uint64_t out[4], cudaStream_t st);
int divide_by_pseudo_vanishing_device(void* d_coeffs, unsigned log_src, unsigned log_target, cudaStream_t st);
int kate_opening_device(const void* d_src, void* d_dest, size_t n, const uint64_t* z, uint64_t f_out[4], cudaStream_t st);
void release_helpers();
size_t launch_count();
} // namespace plonk
} // namespace bbg
