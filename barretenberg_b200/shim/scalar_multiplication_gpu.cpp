// Drop-in replacement for the GPU-backed entry points of the reference's
//   src/barretenberg/curves/bn254/scalar_multiplication.cpp
// Same namespace, same signatures (scalar_multiplication.hpp:41, :60-61, :88-96), so the waffle prover, the
// widgets, preprocess.hpp, the verifier and ReferenceString link against these unchanged.  Compiled against the
// reference's own headers (-I<reference>/src); everything heavy happens behind the C ABI of include/bbgpu.h.
//
// How to link (INTEGRATION.md has the CMake lines): build the reference's scalar_multiplication.cpp with
//   -Dpippenger=cpu_reference_pippenger -Dbatched_scalar_multiplications=cpu_reference_batched_scalar_multiplications
//   -Dgenerate_pippenger_point_table=cpu_reference_generate_pippenger_point_table
//   -Dalt_pippenger=cpu_reference_alt_pippenger -Dpippenger_low_memory=cpu_reference_pippenger_low_memory
//   -Dpippenger_precomputed=cpu_reference_pippenger_precomputed
//   -Dgenerate_pippenger_precompute_table=cpu_reference_generate_pippenger_precompute_table
// (or the objcopy --redefine-sym form of INTEGRATION.md) so its other helper symbols (compute_wnaf_state, pippenger_internal,
// ...) keep their reference CPU bodies, and add this file for the seven names above.  Replace pippenger and
// batched_scalar_multiplications TOGETHER (the reference's batched version calls pippenger from inside an OpenMP
// region with sub-range pointers, scalar_multiplication.cpp:731-738).
//
// Error behaviour: the reference has no error returns (size mismatch prints and returns, :677-685; everything
// else is an ASSERT compiled out in release).  A CUDA failure here prints the bbgpu error and abort()s — there
// is deliberately no CPU fallback.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <barretenberg/curves/bn254/scalar_multiplication.hpp>

#include "bbgpu.h"
#include "shim_stats.h"

namespace
{
[[noreturn]] void die(const char* what, int code)
{
    fprintf(stderr, "bbgpu shim: %s failed: %s (no CPU fallback)\n", what, bbg_error_string(code));
    abort();
}
void ensure_init()
{
    const int e = bbg_shim::ensure_library();
    if (e != 0) die("bbg_init", e);
}
static_assert(sizeof(barretenberg::fr::field_t) == 32, "field_t layout");
static_assert(sizeof(barretenberg::g1::affine_element) == 64, "affine_element layout");
static_assert(sizeof(barretenberg::g1::element) == 96, "element layout");
} // namespace

// Optional: print the BBG_SHIM_STATS=1 call / kernel statistics (call before main returns).
extern "C" void bbg_shim_report(void) { bbg_shim::stats().report(); }

namespace barretenberg
{
namespace scalar_multiplication
{
// scalar_multiplication.cpp:131-140 — table may alias points (reference_string.cpp:23 calls it in place)
void generate_pippenger_point_table(g1::affine_element* points, g1::affine_element* table, size_t num_points)
{
    ensure_init();
    bbg_shim::Timer timer("generate_pippenger_point_table");
    int e = bbg_generate_pippenger_point_table((const uint64_t*)points, (uint64_t*)table, num_points);
    if (e != 0) die("generate_pippenger_point_table", e);
    // the table just written is the SRS the prover will commit against: keep it on the device
    if (num_points >= 1024)
    {
        e = bbg_srs_register((const uint64_t*)table, num_points);
        if (e != 0) die("bbg_srs_register", e);
    }
}

// scalar_multiplication.cpp:457-476.  forced_bucket_width does not change the value (SURVEY.md §8 note 3) and is
// ignored.  The returned point is already normalised (z = fq::one), a valid representative of the same group
// element the reference returns un-normalised; n == 0 and all-zero scalars give the infinity flag.
g1::element pippenger(fr::field_t* scalars, g1::affine_element* points, size_t num_initial_points, size_t /*forced_bucket_width*/)
{
    ensure_init();
    bbg_shim::Timer timer("pippenger");
    g1::element out;
    int e = bbg_msm_g1((const uint64_t*)scalars, (const uint64_t*)points, num_initial_points, (uint64_t*)&out);
    if (e != 0) die("pippenger", e);
    return out;
}

// scalar_multiplication.cpp:317-455 (bucket-ordered prototype over the same 2n-entry table): the same group element as
// pippenger; called by the reference's tests and benchmarks (test_scalar_multiplication.cpp:197-230,
// bench_barretenberg.cpp:487-498).
g1::element alt_pippenger(fr::field_t* scalars, g1::affine_element* points, size_t num_initial_points, size_t forced_bucket_width)
{
    return pippenger(scalars, points, num_initial_points, forced_bucket_width);
}
// scalar_multiplication.cpp:142-263: `points` are the n PLAIN points (no endomorphism entries — the reference applies
// beta on the fly, :222-225), test_scalar_multiplication.cpp:164-195.  Unlike the reference the scalars are NOT overwritten
// with their non-Montgomery endomorphism halves (:144-147, :171).
g1::element pippenger_low_memory(fr::field_t* scalars, g1::affine_element* points, size_t num_points)
{
    ensure_init();
    bbg_shim::Timer timer("pippenger_low_memory");
    g1::element out;
    int e = bbg_msm_g1_points((const uint64_t*)scalars, (const uint64_t*)points, num_points, (uint64_t*)&out);
    if (e != 0) die("pippenger_low_memory", e);
    return out;
}
// scalar_multiplication.cpp:90-129: the caller-visible table of pre-doubled points, computed on the device and written in
// the reference's layout (table[i * n + j] = 2^((bits + 1)(i + 1)) P_j, canonical affine — byte for byte the reference's
// output), and the same vector of round pointers (:118-124).
std::vector<g1::affine_element*> generate_pippenger_precompute_table(g1::affine_element* points, g1::affine_element* table, size_t num_points,
                                                                     size_t bits_per_bucket)
{
    ensure_init();
    bbg_shim::Timer timer("generate_pippenger_precompute_table");
    const size_t num_rounds = (127 + bits_per_bucket + 1) / (bits_per_bucket + 1); // WNAF_SIZE(bits_per_bucket + 1)
    int e = bbg_generate_pippenger_precompute_table((const uint64_t*)points, (uint64_t*)table, num_points, (unsigned)bits_per_bucket);
    if (e != 0) die("generate_pippenger_precompute_table", e);
    std::vector<g1::affine_element*> result(num_rounds);
    result[num_rounds - 1] = points;
    for (size_t i = 0; i + 1 < num_rounds; ++i) result[num_rounds - 2 - i] = &table[i * num_points];
    return result;
}

// scalar_multiplication.cpp:478-488: round_points[r] = 2^((bits + 1)(num_rounds - 1 - r)) P, the last entry being the n
// plain points themselves (:120-124).  The sum does not depend on the window width, so the device runs its own fixed-base
// form: the first call over a point set builds the 2n-entry table and the pre-doubled windows for the device's own width
// in HBM and keeps them under the points' address (bbg_msm_g1_points; bbg_set_srs_precompute), later calls reuse them.
// The caller's CPU-side tables (round_points[0 .. num_rounds - 2]) are not read.
g1::element pippenger_precomputed(fr::field_t* scalars, const std::vector<g1::affine_element*>& round_points, const size_t num_initial_points)
{
    ensure_init();
    bbg_shim::Timer timer("pippenger_precomputed");
    g1::element out;
    if (round_points.empty()) die("pippenger_precomputed (no round points)", BBG_E_BAD_ARGUMENT);
    int e = bbg_msm_g1_points((const uint64_t*)scalars, (const uint64_t*)round_points.back(), num_initial_points, (uint64_t*)&out);
    if (e != 0) die("pippenger_precomputed", e);
    return out;
}

// scalar_multiplication.cpp:650-772: writes only mul_state[i].output, normalised
void batched_scalar_multiplications(multiplication_state* mul_state, size_t num_batches)
{
    if (num_batches == 0) return;
    const size_t num_elements = mul_state[0].num_elements;
    for (size_t i = 1; i < num_batches; ++i)
    {
        if (mul_state[i].num_elements != num_elements)
        {
            printf("batched_scalar_multiplications err: each scalar mul must be same size.\n");
            return;
        }
    }
    ensure_init();
    bbg_shim::Timer timer("batched_scalar_multiplications", num_batches);
    const uint64_t** scalars = (const uint64_t**)malloc(sizeof(uint64_t*) * num_batches * 2);
    const uint64_t** tables = scalars + num_batches;
    uint64_t* outs = (uint64_t*)malloc(96 * num_batches);
    for (size_t i = 0; i < num_batches; ++i)
    {
        scalars[i] = (const uint64_t*)mul_state[i].scalars;
        tables[i] = (const uint64_t*)mul_state[i].points;
    }
    int e = bbg_msm_g1_batched(scalars, tables, num_elements, num_batches, outs);
    if (e != 0) die("batched_scalar_multiplications", e);
    for (size_t i = 0; i < num_batches; ++i) memcpy((void*)&mul_state[i].output, outs + 12 * i, 96);
    free(outs);
    free((void*)scalars);
}
} // namespace scalar_multiplication
} // namespace barretenberg
