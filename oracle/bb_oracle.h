/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * bb_oracle: a plain-C (gcc, unsigned __int128), single-threaded CPU restatement of the two
 * reference hot paths — bn254 G1 Pippenger MSM and the radix-2 Fr NTT family — used ONLY as the
 * checker for the CUDA path (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg).
 * The product (barretenberg_b200/csrc + libbbgpu.so) never includes, links or calls this file.
 *
 * Parity status: PINNED.  tests/test_oracle_pinned.py checks every function below against
 *   (1) the reference's own known-answer vectors (tests/golden/reference_kats.json, restated from
 *       /root/reference/test/test_fq.cpp, test_fr.cpp, test_g1.cpp, test_wnaf.cpp), and
 *   (2) the unmodified reference compiled by oracle/Makefile into oracle/_ref/libbb_ref.so, on
 *       seeded inputs, limb for limb (also frozen into tests/golden/ref_vectors.npz by
 *       tests/golden/make_golden.py so the check survives on machines without the reference).
 *
 * Conventions: field element = 4 x uint64 little-endian limbs, Montgomery form with
 * R = 2^256 unless stated; affine point = 8 x uint64 (x,y); Jacobian point = 12 x uint64 (x,y,z);
 * point at infinity <=> bit 63 of y limb 3 set (reference groups/group.hpp:133-151).
 * "path:line" citations are relative to /root/reference/src/barretenberg/.
 */
#ifndef BB_ORACLE_H
#define BB_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* field selector */
enum { ORC_FQ = 0, ORC_FR = 1 };

/* ---- field arithmetic: fields/field_impl_int128.tcc:72-263 -------------------------------- */
void orc_mul(int field, const uint64_t a[4], const uint64_t b[4], uint64_t r[4]);        /* __mul: [0,p) for inputs < 2p */
void orc_mul_coarse(int field, const uint64_t a[4], const uint64_t b[4], uint64_t r[4]); /* __mul_with_coarse_reduction: [0,2p) */
void orc_sqr(int field, const uint64_t a[4], uint64_t r[4]);
void orc_add(int field, const uint64_t a[4], const uint64_t b[4], uint64_t r[4]);
void orc_sub(int field, const uint64_t a[4], const uint64_t b[4], uint64_t r[4]);
void orc_add_coarse(int field, const uint64_t a[4], const uint64_t b[4], uint64_t r[4]);
void orc_sub_coarse(int field, const uint64_t a[4], const uint64_t b[4], uint64_t r[4]);
void orc_reduce_once(int field, const uint64_t a[4], uint64_t r[4]);
void orc_neg(int field, const uint64_t a[4], uint64_t r[4]);           /* field.hpp:118-121: p - a */
void orc_to_mont(int field, const uint64_t a[4], uint64_t r[4]);       /* field.hpp:224-232 */
void orc_from_mont(int field, const uint64_t a[4], uint64_t r[4]);     /* field.hpp:233-236 */
void orc_invert(int field, const uint64_t a[4], uint64_t r[4]);        /* field.hpp:345-348 (Fermat) */
void orc_pow_small(int field, const uint64_t a[4], uint64_t e, uint64_t r[4]); /* field.hpp:290-332 */
void orc_constant(int which, uint64_t r[4]);                           /* same numbering as ref_constant() */
void orc_mul_n(int field, const uint64_t* a, const uint64_t* b, uint64_t* r, size_t n);
/* element-wise batches, op codes of bbg_field_selftest / bbg_g1_selftest (include/bbgpu.h); residue-only ops canonical */
void orc_field_op_n(int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* r, size_t n);
void orc_g1_op_n(int op, const uint64_t* p_affine, const uint64_t* q_affine, uint64_t* out_affine, size_t n);

/* ---- scalar decomposition: field.hpp:413-485, groups/wnaf.hpp:15-55 ----------------------- */
void orc_split_endo(const uint64_t k[4], uint64_t out[4]); /* out[0..1]=k1, out[2..3]=k2 (aliased form, scalar_multiplication.cpp:292) */
int  orc_fixed_wnaf(const uint64_t scalar128[2], uint32_t* wnaf, size_t num_points, size_t wnaf_bits); /* returns skew */

/* ---- G1: groups/group.hpp:153-534 ---------------------------------------------------------- */
void orc_g1_dbl(const uint64_t p[12], uint64_t out[12]);
void orc_g1_mixed_add(const uint64_t p1[12], const uint64_t p2_affine[8], uint64_t out[12]);
void orc_g1_add(const uint64_t p1[12], const uint64_t p2[12], uint64_t out[12]);
void orc_g1_normalize(const uint64_t p[12], uint64_t out[12]);
void orc_g1_batch_normalize(uint64_t* pts, size_t n);
/* value-equivalent (plain double-and-add, NOT the reference's endo-wNAF ladder): affine k*P */
void orc_g1_scalar_mul(const uint64_t affine[8], const uint64_t scalar_mont[4], uint64_t out_affine[8]);
int  orc_g1_on_curve(const uint64_t affine[8]);
/* out[i] = (start + i*step) * G, i < n, normalised affine — the synthetic point sets of SURVEY.md §8c-3
 * (repeated mixed_add + one batch_normalize, like ref_g1_arith_progression in ref_capi.cpp) */
void orc_g1_arith_progression(const uint64_t start_mont[4], const uint64_t step_mont[4], uint64_t* out_affine, size_t n);

/* ---- MSM: curves/bn254/scalar_multiplication.cpp ------------------------------------------- */
size_t orc_get_optimal_bucket_width(size_t n);                                             /* :21-81 */
void orc_generate_pippenger_point_table(const uint64_t* points_n, uint64_t* table_2n, size_t n); /* :131-140 */
void orc_pippenger(const uint64_t* scalars_mont, const uint64_t* table_2n, size_t n,
                   size_t forced_bucket_width, uint64_t out_jac[12]);                      /* :457-476, :576-648 */
/* value-equivalent to batched_scalar_multiplications (:650-772) for ONE batch: pippenger + normalise.
 * (The thread partition of the reference changes the Jacobian representative, not the point.) */
void orc_msm_normalized(const uint64_t* scalars_mont, const uint64_t* table_2n, size_t n, uint64_t out_jac[12]);

/* ---- NTT: polynomials/evaluation_domain.cpp, polynomial_arithmetic.cpp ---------------------- */
typedef struct orc_domain orc_domain;
orc_domain* orc_domain_new(size_t n);                                                      /* evaluation_domain.cpp:57-75,172-178 */
void orc_domain_free(orc_domain* d);
void orc_domain_constant(const orc_domain* d, int which, uint64_t r[4]); /* 0 root,1 root_inverse,2 domain,3 domain_inverse,4 generator,5 generator_inverse */
/* op: 0 fft, 1 ifft, 2 coset_fft, 3 coset_ifft, 4 fft_with_constant, 5 ifft_with_constant, 6 coset_fft_with_constant
 * (polynomial_arithmetic.cpp:266-315); coeffs in place, length d->size */
void orc_ntt(const orc_domain* d, int op, uint64_t* coeffs, const uint64_t* constant);
/* polynomial_arithmetic.cpp:381-476 (value-equivalent: one inversion per element instead of the batched sweep) */
void orc_compute_lagrange_polynomial_fft(uint64_t* l_1, size_t log2_src, size_t log2_target);
void orc_poly_evaluate(const uint64_t* coeffs, const uint64_t z[4], size_t n, uint64_t out[4]); /* :337-373, Horner-equivalent */

#ifdef __cplusplus
}
#endif
#endif
