/* bbgpu — C ABI of the B200-native MSM / NTT library (libbbgpu.so).
 *
 * This is the drop-in boundary for the two data-parallel hot paths of the Barretenberg "waffle" PLONK
 * prover.  The reference has no FFI layer: its boundary is the set of free C++ functions below, which
 * the shim sources under barretenberg_b200/shim/ re-define with identical signatures and forward here.
 * Citations are relative to /root/reference/src/barretenberg/.
 *
 *   curves/bn254/scalar_multiplication.hpp:60-61   g1::element pippenger(fr::field_t*, g1::affine_element*, size_t, size_t)
 *   curves/bn254/scalar_multiplication.hpp:88-96   void batched_scalar_multiplications(multiplication_state*, size_t)
 *   curves/bn254/scalar_multiplication.hpp:41      void generate_pippenger_point_table(affine*, affine*, size_t)
 *   polynomials/polynomial_arithmetic.hpp:28-39    fft / ifft / coset_fft / coset_ifft / fft_with_constant /
 *                                                  ifft_with_constant / coset_fft_with_constant
 *
 * Data conventions (identical to the reference):
 *   field element  = 4 x uint64 little-endian limbs, Montgomery form (R = 2^256), 32 bytes
 *   affine point   = x, y                      (64 bytes); infinity <=> bit 63 of y limb 3
 *   Jacobian point = x, y, z                   (96 bytes)
 *   point table    = 2n affine entries [P_0, phi(P_0), P_1, phi(P_1), ...], phi(x,y) = (beta x, -y)
 *   NTT            : natural order in/out, in place, outputs canonical in [0,p); inputs may be in [0,2p)
 *   MSM            : scalars may be in [0,2p); output normalised (z = fq::one, x,y canonical) or the
 *                    infinity flag set — the form batched_scalar_multiplications leaves in .output
 *
 * Every function returns 0 on success, a cudaError_t value (< 1000) or a BBG_E_* code otherwise.
 * There is NO CPU fallback: without a usable CUDA device every compute entry point fails.
 * The library is not re-entrant; calls are serialised by an internal mutex (the reference's callers are
 * single-threaded, scalar_multiplication.cpp:731 keeps its OpenMP region inside the callee).
 */
#ifndef BBGPU_H
#define BBGPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BBG_E_BAD_SIZE 1002      /* NTT: log2_n outside [1, 28] */
#define BBG_E_BAD_OP 1003
#define BBG_E_NULL_CONSTANT 1004
#define BBG_E_NOT_INITIALISED 1005
#define BBG_E_NO_DEVICE 1006
#define BBG_E_BAD_ARGUMENT 1007
#define BBG_E_TOO_LARGE 1008     /* MSM: more than 2^27 table entries */

/* NTT operation selector == the reference entry point being replaced */
enum bbg_ntt_op
{
    BBG_FFT = 0,                     /* polynomial_arithmetic.cpp:266-269 */
    BBG_IFFT = 1,                    /* :271-277 */
    BBG_COSET_FFT = 2,               /* :287-291 */
    BBG_COSET_IFFT = 3,              /* :311-315 */
    BBG_FFT_WITH_CONSTANT = 4,       /* :279-285 */
    BBG_IFFT_WITH_CONSTANT = 5,      /* :301-309 */
    BBG_COSET_FFT_WITH_CONSTANT = 6, /* :293-299 */
};

/* ---- lifetime ------------------------------------------------------------------------------- */
int bbg_init(int device);            /* select device, create the work stream; idempotent */
/* One process, several GPUs of one box (the reference splits an MSM into point ranges inside
 * batched_scalar_multiplications, scalar_multiplication.cpp:688-761, so its caller gets every core without asking).
 * devices[0] becomes the primary device — everything bbg_init(devices[0]) gives, every NTT, every device pointer of this
 * API lives there — and each further device gets a worker thread, a stream and its own copy of every registered point
 * table.  From then on an MSM of >= 2^15 points over a registered / cached table (bbg_msm_g1, _batched, _launch, the
 * *_dev forms when d_table points into such a table, and the commitments of the resident prover) is cut into contiguous
 * point ranges, one per device, exactly as :703-728 cuts it per thread; the per-device sums are added on the calling
 * thread (:750-765).  Results are the same group elements, hence bit-identical outputs.  May be called after bbg_init on
 * the same primary device; idempotent for the same list.  BBG_MULTI_MIN_POINTS / BBG_MULTI_PRIMARY_SHARE tune the split. */
int bbg_init_multi(const int* devices, int count);
int bbg_device_count(void);          /* devices driven by this instance (1 after bbg_init) */
int bbg_shutdown(void);              /* free tables, caches, workspace */
int bbg_set_stream(void* cuda_stream); /* run on a caller-owned cudaStream_t (e.g. torch's current stream) */
const char* bbg_error_string(int code);
uint64_t bbg_launch_count(void);     /* kernels launched by this library so far (bench.py: gpu_launches) */

/* ---- NTT, host buffers (what the shims call) ------------------------------------------------- */
/* coeffs: n = 2^log2_n field elements, transformed in place.  constant: one field element, used by the
 * *_WITH_CONSTANT ops, ignored (may be NULL) otherwise. */
int bbg_ntt_fr(uint64_t* coeffs, unsigned log2_n, int op, const uint64_t* constant);
/* batch independent polynomials of the same size (prover wire / permutation / quotient polynomials) */
int bbg_ntt_fr_batched(uint64_t* const* coeffs, size_t batch, unsigned log2_n, int op, const uint64_t* constant);

/* ---- NTT, device-resident --------------------------------------------------------------------- */
/* d_coeffs: device pointer to `batch` polynomials, `stride_elems` field elements apart, in place */
int bbg_ntt_fr_dev(void* d_coeffs, size_t stride_elems, size_t batch, unsigned log2_n, int op, const uint64_t* constant);

/* ---- element-wise helpers next to the NTT (widened one at a time) ---------------- */
/* polynomial_arithmetic::compute_lagrange_polynomial_fft (polynomials/polynomial_arithmetic.hpp:54,
 * polynomial_arithmetic.cpp:381-476; caller prover.cpp:351): writes the 2^log2_target coset evaluations of
 * L_1(X) = (X^n - 1) / (n (X - 1)), n = 2^log2_src, canonical Montgomery limbs.  Output only. */
int bbg_compute_lagrange_polynomial_fft(uint64_t* l_1_coefficients, unsigned log2_src, unsigned log2_target);
int bbg_compute_lagrange_polynomial_fft_dev(void* d_out, unsigned log2_src, unsigned log2_target);

/* ---- MSM ------------------------------------------------------------------------------------- */
/* Register a 2n-entry point table (reference_string.cpp:20-23 owns it for the prover's lifetime): uploaded
 * once; later MSM calls whose `points` pointer lies inside [table, table + 2n) reuse the device copy
 * (batched_scalar_multiplications passes &points[2 * offset], scalar_multiplication.cpp:720-723). */
int bbg_srs_register(const uint64_t* table_2n, size_t n);
int bbg_srs_unregister(const uint64_t* table_2n);
/* Auto cache (off by default, the shims switch it on): an unregistered table of >= 1024 points is kept on the device
 * the first time an MSM sees it; every hit re-checks a fingerprint of sampled host entries, so a buffer that was
 * freed and rewritten behind the same address is re-uploaded, never trusted. */
int bbg_set_auto_srs_cache(int enable);
/* Fixed-base form for registered / cached tables (the GPU side of generate_pippenger_precompute_table and
 * pippenger_precomputed, scalar_multiplication.cpp:90-129, :478-573; reference test test_scalar_multiplication.cpp:233-269):
 * with the switch on, every table that is registered (or auto-cached) also gets its pre-doubled windows
 * pre[w][j] = 2^(c w) table[j] built on the device, W x the table's memory, on every device of the instance.  MSMs that
 * name such a table, or a sub-range of it, then add the digits of ALL windows of a scalar into one bucket set — no
 * doublings between windows, 1 / W of the buckets to reduce, a wider window — and give the same group element, hence
 * the same normalised output.  The SRS of a prover is fixed, so this is built once per ReferenceString.  Off by default
 * (the shims switch it on; BBG_SRS_PRECOMPUTE=0 keeps it off); skipped silently for a table whose windows would not fit
 * BBG_SRS_PRECOMPUTE_MAX_MB (default: a quarter of the free device memory). */
int bbg_set_srs_precompute(int enable);
/* device copy of a registered table (or of the sub-range starting at the given entry), for device-resident callers of
 * bbg_msm_g1_dev / _partial_dev*; window_bits / windows (may be NULL) = the fixed-base windows behind it, 0 when none */
int bbg_srs_device_table(const uint64_t* table_2n, void** d_table, int* window_bits, int* windows);
/* sum_i scalars[i] * P_i over table entries points_table[0 .. 2n); out_xyz = 12 limbs, normalised */
int bbg_msm_g1(const uint64_t* scalars, const uint64_t* points_table, size_t n, uint64_t out_xyz[12]);
/* the same sum over n PLAIN affine points (64 bytes each, no endomorphism entries): what pippenger_low_memory and
 * pippenger_precomputed are handed (scalar_multiplication.hpp:52, :79-81; .cpp:142-263, :478-573 apply the endomorphism on
 * the fly); the 2n-entry table is built on the device.  Unlike pippenger_low_memory the scalars are not overwritten. */
int bbg_msm_g1_points(const uint64_t* scalars, const uint64_t* points_n, size_t n, uint64_t out_xyz[12]);
/* generate_pippenger_precompute_table (scalar_multiplication.hpp:83-86, .cpp:90-129): from n plain points,
 * table[i * n + j] = 2^((bits_per_bucket + 1)(i + 1)) P_j for i < WNAF_SIZE(bits_per_bucket + 1) - 1, canonical affine
 * coordinates — byte for byte what the reference writes.  (The device keeps its own fixed-base windows for the MSM itself,
 * bbg_set_srs_precompute; this call fills the caller-visible table.) */
int bbg_generate_pippenger_precompute_table(const uint64_t* points_n, uint64_t* table, size_t n, unsigned bits_per_bucket);
/* `batches` MSMs of the same size n (multiplication_state[], scalar_multiplication.hpp:88-94) */
int bbg_msm_g1_batched(const uint64_t* const* scalars, const uint64_t* const* points_tables, size_t n, size_t batches,
                       uint64_t* out_xyz /* batches x 12 */);
/* device-resident: d_scalars (n x 32 B) and d_table (2n x 64 B) already in HBM; result to host */
int bbg_msm_g1_dev(const void* d_scalars, const void* d_table, size_t n, uint64_t out_xyz[12]);
/* multi-GPU building blocks: a rank's un-normalised partial (16 limbs: X, Y, ZZ, ZZZ with x = X/ZZ,
 * y = Y/ZZZ; ZZ = 0 <=> infinity) and the final fold of gathered partials into a normalised point */
int bbg_msm_g1_partial_dev(const void* d_scalars, const void* d_table, size_t n, uint64_t out_xyzz[16]);
int bbg_g1_fold_partials(const uint64_t* partials_xyzz /* count x 16 */, size_t count, uint64_t out_xyz[12]);
/* The same MSM in two steps, for device-resident callers with independent work to queue in between (a prover's wire
 * commitment next to the transforms of the other wires): launch returns once the kernels are queued on a second stream,
 * ordered behind everything queued on the work stream so far; finish waits for them and folds the windows.  d_scalars and
 * d_table must not be modified in between.  Up to 6 tickets may be pending. */
int bbg_msm_g1_partial_dev_launch(const void* d_scalars, const void* d_table, size_t n, int* ticket);
int bbg_msm_g1_partial_finish(int ticket, uint64_t out_xyzz[16]);
/* bbg_msm_g1 in two steps with HOST buffers (what a caller of pippenger() with independent transforms to run in between
 * would use: prover.cpp:65-124 commits to a wire while the next wires are still being transformed): launch copies the
 * scalars to the device and queues the MSM on the library's second stream, then returns; the NTT / MSM calls made before
 * finish run beside it (their PCIe copies hide the MSM, whose own traffic is 32 bytes per point).  finish waits, folds the
 * windows on the host and writes the normalised Jacobian point like bbg_msm_g1.  scalars may be reused once launch
 * returns when they are pageable, after finish when they are pinned; up to 6 tickets may be pending. */
int bbg_msm_g1_launch(const uint64_t* scalars, const uint64_t* points_table, size_t n, int* ticket);
int bbg_msm_g1_finish(int ticket, uint64_t out_xyz[12]);
/* table[2i] = points[i], table[2i+1] = (beta x_i, -y_i): the layout of generate_pippenger_point_table
 * (scalar_multiplication.cpp:131-140) computed on the device; table may alias points */
int bbg_generate_pippenger_point_table(const uint64_t* points_n, uint64_t* table_2n, size_t n);

/* device-resident variant (d_table must not alias d_points) */
int bbg_generate_pippenger_point_table_dev(const void* d_points, void* d_table, size_t n);
/* Synthetic point sets for sizes beyond the SRS (BASELINE configs[3], "random multiples of the G1 generator"):
 * d_points[i] = (start + i * step) * G for i < n, affine, canonical Montgomery coordinates, written to device
 * memory (n x 64 bytes).  start / step are Fr elements in Montgomery form. */
int bbg_g1_generate_multiples_dev(const uint64_t start[4], const uint64_t step[4], void* d_points, size_t n);

/* ---- the reference's stand-alone polynomial helpers on host buffers ------------------------------
 * For callers that keep the reference's round structure (shim/polynomial_arithmetic_gpu.cpp); the resident rounds below
 * use the same kernels without leaving HBM.
 * polynomial_arithmetic::evaluate (polynomial_arithmetic.hpp:41, .cpp:337-373): out = sum_i coeffs[i] z^i, canonical */
int bbg_fr_evaluate(const uint64_t* coeffs, size_t n, const uint64_t z[4], uint64_t out[4]);
/* divide_by_pseudo_vanishing_polynomial (.hpp:57, .cpp:478-560): coeffs = 2^log2_target coset evaluations, multiplied in
 * place by (x - w_n^(n-1)) / (x^n - 1), n = 2^log2_src, x = g w_T^i; log2_target - log2_src in {0, 1, 2}; canonical */
int bbg_fr_divide_by_pseudo_vanishing_polynomial(uint64_t* coeffs, unsigned log2_src, unsigned log2_target);
/* compute_kate_opening_coefficients (.hpp:61, .cpp:562-591): dest = (F(X) - F(z)) / (X - z) for F = src (n coefficients),
 * canonical (the reference's serial recurrence leaves lazily reduced values: same field elements); f_out = F(z);
 * dest may alias src */
int bbg_fr_compute_kate_opening_coefficients(const uint64_t* src, uint64_t* dest, const uint64_t z[4], size_t n, uint64_t f_out[4]);

/* ---- prover construction ----------------------------------------------------
 * evaluation_domain::compute_lookup_table (polynomials/evaluation_domain.hpp:35, evaluation_domain.cpp:33-54, :172-178):
 * roots = 2 * 2^log2_size field elements (host); per direction round i (m = 2^(i+1), i = 0 .. log2_size - 2) holds
 * w_(2m)^j, j < m, at offset 2^(i+1) - 2: forward rounds in roots[0, size), inverse rounds in roots[size, 2 size).
 * Canonical Montgomery values (the reference's serial chain leaves them lazily reduced in [0, 2p): same field elements).
 * Shim: barretenberg_b200/shim/evaluation_domain_gpu.cpp. */
int bbg_fr_domain_lookup_table(uint64_t* roots, unsigned log2_size);
/* io::read_transcript's G1 part + generate_pippenger_point_table (io/io.hpp:76-98, :157-182, reference_string.cpp:20-23,
 * scalar_multiplication.cpp:131-140): g1_bytes = the (n - 1) x 64 raw bytes that follow the transcript manifest (x then y,
 * four 64-bit limbs each, least significant limb first, big-endian bytes inside a limb, plain values); writes the 2n-entry
 * table [G, phi(G), P_1, phi(P_1), ...] to table_2n (host) and keeps the device copy registered as the SRS behind that
 * address (as bbg_srs_register would, without a second upload).  Shim: barretenberg_b200/shim/reference_string_gpu.cpp. */
int bbg_srs_from_transcript(const uint8_t* g1_bytes, size_t n, uint64_t* table_2n);

/* ---- HBM-resident PLONK prover rounds -------------------------------------
 * Stand behind waffle::Prover::construct_proof (waffle/proof_system/prover/prover.cpp:657-666) for circuits built from
 * the arithmetic, bool, MiMC and sequential widgets (Standard / Bool / MiMC / Extended composers): the witness, the permutation mappings and the selectors are
 * uploaded once per proof, every polynomial of the proof stays in HBM, and only commitments (12 limbs, the same
 * normalised form as bbg_msm_g1) and evaluations (4 limbs, canonical) come back.  The Fiat-Shamir transcript stays
 * with the caller (challenge.hpp), which feeds the challenges back in between rounds.
 * Host shim: barretenberg_b200/shim/prover_gpu.cpp.  n = 2^log2_n gates, 2 <= log2_n <= 23. */
typedef struct bbg_plonk_prover bbg_plonk_prover;
int bbg_plonk_create(unsigned log2_n, bbg_plonk_prover** out);
int bbg_plonk_destroy(bbg_plonk_prover* p);
/* prover.hpp:44-46 w_l, w_r, w_o in Lagrange form, n elements each (host, not modified) */
int bbg_plonk_set_witness(bbg_plonk_prover* p, const uint64_t* w_l, const uint64_t* w_r, const uint64_t* w_o);
/* prover.hpp:56-58 sigma_k_mapping: n entries, wire column << 30 | row (permutation.hpp:13-88) */
int bbg_plonk_set_permutation(bbg_plonk_prover* p, const uint32_t* sigma_1_mapping, const uint32_t* sigma_2_mapping, const uint32_t* sigma_3_mapping);
/* The circuit's widgets in the prover's order (prover.hpp:60 `widgets`), each kind at most once, and their selector
 * polynomials in Lagrange form (host, not modified), back to back in this order:
 *   arithmetic  q_m, q_l, q_r, q_o, q_c                 (arithmetic_widget.hpp:45-49)
 *   bool        q_bl, q_br, q_bo                        (bool_widget.hpp:45-47)
 *   MiMC        q_mimc_selector, q_mimc_coefficient     (mimc_widget.hpp:43-44)
 *   sequential  q_o_next                                (sequential_widget.hpp:46)
 * "selectors" below = the total number of these polynomials (5 for a StandardComposer circuit). */
enum bbg_plonk_widget
{
    BBG_WIDGET_ARITHMETIC = 0,
    BBG_WIDGET_BOOL = 1,
    BBG_WIDGET_MIMC = 2,
    BBG_WIDGET_SEQUENTIAL = 3,
};
int bbg_plonk_set_widgets(bbg_plonk_prover* p, const int* kinds, int count, const uint64_t* const* selectors);
/* shorthand for the StandardComposer's single arithmetic widget */
int bbg_plonk_set_arithmetic_selectors(bbg_plonk_prover* p, const uint64_t* q_m, const uint64_t* q_l, const uint64_t* q_r, const uint64_t* q_o,
                                       const uint64_t* q_c);
/* ReferenceString::monomials, the 2n-entry point table (goes through the SRS cache of bbg_srs_register) */
int bbg_plonk_set_srs(bbg_plonk_prover* p, const uint64_t* points_table, size_t n);
/* compute_wire_coefficients + compute_wire_commitments (prover.cpp:126-135, :65-89): out = W_L, W_R, W_O (3 x 12) */
int bbg_plonk_round_wires(bbg_plonk_prover* p, uint64_t* out_xyz);
/* compute_z_coefficients + compute_z_commitment (:137-225, :91-107): out = Z_1 */
int bbg_plonk_round_grand_product(bbg_plonk_prover* p, const uint64_t beta[4], const uint64_t gamma[4], uint64_t out_xyz[12]);
/* rest of compute_quotient_polynomial + compute_quotient_commitment (:227-463, :109-124) and every widget's
 * compute_quotient_contribution (arithmetic_widget.cpp:60-99, bool_widget.cpp:62-100, mimc_widget.cpp:57-89,
 * sequential_widget.cpp:47-62); alpha_base = the first widget's scaling (alpha^4 in prover.cpp:436-441), chained through
 * the widgets as the reference does: out = T_LO, T_MID, T_HI (3 x 12) */
int bbg_plonk_round_quotient(bbg_plonk_prover* p, const uint64_t beta[4], const uint64_t gamma[4], const uint64_t alpha[4], const uint64_t alpha_base[4],
                             uint64_t* out_xyz);
/* compute_linearisation_coefficients, first half (:465-477): out (9 x 4) = w_l(z), w_r(z), w_o(z), [beta sigma_1](z),
 * [beta sigma_2](z), Z(z w), t(z) over the quotient's first 3n coefficients, w_o(z w) when a MiMC / sequential widget is
 * present (:455-463), q_mimc_coefficient(z) (mimc_widget.cpp:91-94); zero where not applicable */
int bbg_plonk_round_evaluations(bbg_plonk_prover* p, const uint64_t zeta[4], const uint64_t zeta_omega[4], uint64_t* out_evals);
/* second half (:479-503 and the widgets' compute_linear_contribution): r[i] = s_0 z[i] + s_1 [beta sigma_3][i] +
 * sum_k s_(2+k) selector_k[i] (coefficient forms); scalars = (2 + selectors) x 4, the caller folds wire evaluations and
 * the alpha chain into them; out = r(zeta) */
int bbg_plonk_round_linearise(bbg_plonk_prover* p, const uint64_t* scalars, const uint64_t zeta[4], uint64_t out_linear_eval[4]);
/* compute_opening_elements after the nu challenge (:505-655; compute_kate_opening_coefficients,
 * polynomial_arithmetic.cpp:562-591): nu_powers = nu^1..nu^7 (7 x 4); wire_shift_terms (3 x 4) = coefficients of w_l, w_r,
 * w_o in the shifted opening polynomial (:597-631, zero = wire not needed); selector_terms (selectors x 4) = coefficients
 * of the selectors in the opening polynomial (compute_opening_poly_contribution, zero for most); beta_inv is accepted
 * for symmetry with the reference's formula and not used (sigma is held unscaled on the device);
 * out = PI_Z, PI_Z_OMEGA (2 x 12) */
int bbg_plonk_round_openings(bbg_plonk_prover* p, const uint64_t* nu_powers, const uint64_t beta_inv[4], const uint64_t zeta[4],
                             const uint64_t zeta_omega[4], const uint64_t* wire_shift_terms, const uint64_t* selector_terms, uint64_t* out_xyz);

/* ---- caller-owned host buffers -------------------------------------------------------------------
 * The reference's callers pass pageable memory (aligned_alloc, types.hpp:25) and reuse the same long-lived buffers call
 * after call (barretenberg::polynomial members, ReferenceString::monomials).  With the cache on, a pageable buffer of
 * >= 1 MiB that keeps coming back behind the same address and size (6th copy; BBG_HOST_REGISTER_AFTER) is page-locked in
 * place (cudaHostRegister) and from then on copied at the pinned rate, without the staging memcpy; page-locking costs
 * about ten staged copies, so one-proof temporaries are left alone.  Contract: hand a buffer to bbg_host_buffer_forget BEFORE
 * freeing it (the driver keeps DMA mappings of the physical pages; an address range unmapped and mapped again would be
 * read through the old ones).  shim/host_buffer_free_wrap.cpp does that for a prover linked with -Wl,--wrap=free.
 * Off by default; BBG_HOST_REGISTER_MAX_MB caps the page-locked total (default 16384). */
int bbg_set_host_register_cache(int enable);
int bbg_host_buffer_forget(const void* host_ptr); /* any pointer into the buffer; unknown pointers are ignored; thread-safe */
/* what page-locking has cost so far: wall milliseconds inside cudaHostRegister, bytes and buffers registered */
int bbg_host_register_stats(double* register_ms, uint64_t* registered_bytes, uint64_t* registrations);

/* ---- device self test (tests only; bbg_selftest.cu) --------------------------------------------------
 * Element-wise field / group primitives on the device, host buffers in and out, for limb-for-limb comparison with the
 * reference's known-answer vectors (test/test_fq.cpp:51-133, test_fr.cpp:51-88, test_g1.cpp:41-122) and the oracle.
 * field: 0 = Fq, 1 = Fr.  op: 0 mul (coarse: (ab + Mp) / 2^256), 1 sqr (coarse), 2 mul by a constant (canonical),
 * 3 add (coarse), 4 sub (coarse), 5 reduce_once, 6 neg (lazy range), 7 to_montgomery_form, 8 from_montgomery_form,
 * 9 invert, 10 a - b + 2p uncorrected, 11 mul (canonical), 12 mul by a constant (raw, must lie in [0, 2p)).
 * b may be NULL for the unary ops. */
int bbg_field_selftest(int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t count);
/* p, q, out: affine points (64 bytes).  op: 0 mixed add P + Q, 1 2P + 2Q through the general addition, 2 4P through two
 * doublings, 3 ((inf + P) + Q) + P by mixed additions, 4 general addition P + Q, 5 the endomorphism table entry
 * (beta x, -y), 6 2P from the affine image. */
int bbg_g1_selftest(int op, const uint64_t* p, const uint64_t* q, uint64_t* out, size_t count);

/* ---- device memory helpers (tests, bench, device-resident callers) ---------------------------- */
int bbg_dev_alloc(void** d_ptr, size_t bytes);
int bbg_dev_free(void* d_ptr);
int bbg_copy_h2d(void* d_dst, const void* h_src, size_t bytes);
int bbg_copy_d2h(void* h_dst, const void* d_src, size_t bytes);
int bbg_sync(void);
/* CUDA-event stopwatch on the library's work stream */
int bbg_timer_start(void);
int bbg_timer_stop(float* elapsed_ms);

/* ---- measurement ------------------------------------------------------------------------------ */
/* Per-kernel stopwatch (CUDA events on the work stream), off by default: bench.py enables it for a separate
 * untimed pass to attribute step time to kernels.  ids 0 .. bbg_profile_count()-1, names via bbg_profile_name. */
int bbg_profile_enable(int on);           /* also resets the counters */
int bbg_profile_count(void);
const char* bbg_profile_name(int id);
int bbg_profile_read(int id, double* total_ms, uint64_t* launches);
/* Dependency-free integer multiply-add throughput on all SMs: the IMAD roofline denominator.
 * mode 0: mad.lo.u32   1: mad.wide.u32 (64-bit accumulate)   2: carry-chained mad.lo.cc/madc.hi.cc pairs
 *      3: Fq Montgomery products (reports field products/s)   4: Fr Montgomery products
 * ops_per_second: 32x32 multiply-adds (modes 0-2) or field products (3-4) per second.
 * Measured on B200: mode 0 = 1.85e13/s (64 IMAD/clk/SM), modes 1-2 = 8.9e12/s: a 32x32->64 multiply-add
 * (IMAD.WIDE) occupies the integer multiply pipe for two issue slots. */
int bbg_microbench(int mode, int iters, double* ops_per_second, float* elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif
