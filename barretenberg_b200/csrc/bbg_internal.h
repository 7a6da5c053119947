// Internal declarations shared by the CUDA translation units behind include/bbgpu.h.
#pragma once
#include "bbg_rt.h"
#include "bbg_g1.cuh"

#include <stdio.h>
#include <stdlib.h>
// BBG_DEBUG=1: name the failing call on stderr (development aid; the error code is returned either way)
inline int bbg_trace_error(int e, const char* expr, const char* file, int line)
{
    static const bool on = getenv("BBG_DEBUG") != nullptr;
    if (on) fprintf(stderr, "bbgpu: error %d from %s at %s:%d\n", e, expr, file, line);
    return e;
}
#define BBG_CHECK(expr)                                                                            \
    do                                                                                             \
    {                                                                                              \
        int bbg_err_ = (int)(expr);                                                                \
        if (bbg_err_ != 0) return bbg_trace_error(bbg_err_, #expr, __FILE__, __LINE__);            \
    } while (0)

namespace bbg
{
// NTT operation codes == the reference's seven entry points (polynomial_arithmetic.hpp:28-39)
enum ntt_op
{
    OP_FFT = 0,
    OP_IFFT = 1,
    OP_COSET_FFT = 2,
    OP_COSET_IFFT = 3,
    OP_FFT_WITH_CONSTANT = 4,
    OP_IFFT_WITH_CONSTANT = 5,
    OP_COSET_FFT_WITH_CONSTANT = 6,
};

// device-resident NTT: batch polynomials of 2^log_n elements, `stride` elements apart, in place.
int ntt_device(void* d_coeffs, size_t stride, size_t batch, unsigned log_n, int op, const uint64_t* constant, cudaStream_t stream);
#ifndef BBG_EMULATE
bool ntt_host_blocks_applicable(unsigned log_n);
// the transform of ONE polynomial in a (partly) page-locked host buffer with uploads / passes / downloads overlapped block by block
int ntt_host_blocks(void* h_coeffs, size_t safe_lo, size_t safe_hi, void* d_coeffs, unsigned log_n, int op, const uint64_t* constant, cudaStream_t st,
                    cudaStream_t copy_in, cudaStream_t copy_out);
#endif
void ntt_use_side_scratch(bool on); // route the following ntt_device calls to a second scratch buffer (second stream)
int ntt_release_tables();
// coset evaluations of L_1 on the 2^log_target domain for a 2^log_src circuit (compute_lagrange_polynomial_fft)
int lagrange_fft_device(void* d_out, unsigned log_src, unsigned log_target, cudaStream_t stream);
size_t ntt_launch_count();

// device-resident MSM over a device point table (2n affine entries, reference layout) and device scalars.
// Writes the un-normalised sum as XYZZ (16 x uint64: X, Y, ZZ, ZZZ) to the HOST buffer out_xyzz_host;
// synchronises the stream (the window fold is the host-side finish, bbg_host_g1.h).
int msm_device(const void* d_scalars, const void* d_table, size_t n, void* out_xyzz_host, cudaStream_t stream);
// `batch` same-size MSMs over one table in a single pipeline; out_xyzz_host: batch x 16 uint64
int msm_device_batched(const void* const* d_scalars, size_t batch, const void* d_table, size_t n, void* out_xyzz_host, cudaStream_t stream);
// the same in two steps: msm_launch queues the kernels on `stream` using workspace 0 or 1 and hands back a ticket;
// msm_finish waits for that MSM and folds its windows on the host (up to 6 tickets may be pending)
int msm_launch(int workspace, const void* const* d_scalars, size_t batch, const void* d_table, size_t n, cudaStream_t stream, int* ticket_out);
int msm_finish(int ticket, void* out_xyzz_host);
// msm_launch with the scalars either on the primary device or — host_scalars — still in the caller's host buffers: the
// primary's point range is uploaded into `staging` (batch * n * 32 bytes of primary-device memory) behind `stream`, and
// in a multi-GPU instance every other device uploads its own range over its own PCIe link
int msm_launch_any(int workspace, const void* const* scalars, bool host_scalars, void* staging, size_t batch, const void* d_table, size_t n,
                   cudaStream_t stream, int* ticket_out);
bool msm_ticket_pending(int ticket);
int msm_release_workspace();
// Multi-GPU (one process, several devices): devices[0] = the primary device the library was initialised on; the others
// get a worker thread each.  Registered point tables are replicated to every device; msm_launch then fans an MSM over
// such a table out by contiguous point range (scalar_multiplication.cpp:703-728) and msm_finish adds the per-device sums.
int msm_multi_init(const int* devices, int count);
int msm_multi_device_count();
int msm_multi_replicate(const void* d_table_primary, size_t bytes, cudaStream_t stream);
int msm_multi_drop_replica(const void* d_table_primary);
int msm_multi_quiesce();
// Fixed-base form (generate_pippenger_precompute_table / pippenger_precomputed, scalar_multiplication.cpp:90-129,
// :478-573): pre-doubled windows of a table on every device; MSMs that name the table (or a sub-range) then add the digits
// of all windows into one bucket set.  Build after msm_multi_replicate, drop before msm_multi_drop_replica.
int msm_fixed_base_build(const void* d_table_primary, size_t n_srs, cudaStream_t stream);
int msm_fixed_base_drop(const void* d_table_primary);
int msm_fixed_base_info(const void* d_table_primary, int* c, int* W);
size_t msm_launch_count();
// d_points[i] = (start + i * step) * G (affine, canonical), i < n; start / step: Fr Montgomery limbs (host)
int g1_generate_progression_device(const uint64_t* start_mont, const uint64_t* step_mont, void* d_points, size_t n, cudaStream_t stream);
// raw transcript G1 bytes ((n - 1) x 64) -> the 2n-entry point table, generator first (io.hpp:157-182 + :131-140)
int g1_table_from_transcript_device(const void* d_g1_bytes, void* d_table, size_t n, cudaStream_t stream);
// evaluation_domain::compute_lookup_table (evaluation_domain.cpp:33-54, :172-178): 2 * 2^log_size elements on the device
int domain_lookup_table_device(void* d_roots, unsigned log_size, cudaStream_t stream);
// out[i * n + j] = 2^(bits_per_window (i + 1)) P_j, i < rounds - 1 (generate_pippenger_precompute_table's layout)
int g1_precompute_plain_device(const void* d_points, void* d_out, size_t n, int bits_per_window, int rounds, cudaStream_t stream);
// table[2i] = P_i, table[2i+1] = (beta x_i, -y_i) on device (generate_pippenger_point_table layout)
int g1_build_endo_table_device(const void* d_points, void* d_table, size_t n, cudaStream_t stream);
} // namespace bbg
