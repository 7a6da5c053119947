// Drop-in replacement for waffle::ReferenceString::ReferenceString(size_t) of the reference's
//   src/barretenberg/waffle/reference_string/reference_string.cpp:15-34      
// Same class, same signature (reference_string.hpp:17).  The reference reads the transcript, byte-swaps and converts
// 2 (n - 1) coordinates to Montgomery form on one host thread (io/io.hpp:76-98) and then builds the interleaved point
// table; here the raw G1 bytes go to the device in one copy, a single kernel does byte order + Montgomery form + the
// endomorphism half of the table, and the device copy stays registered as the SRS every later MSM uses — the table is
// never uploaded again.  The two G2 points and the pairing precomputation (verifier side) use the reference's own code.
//
// How to link: make the reference's definition of this one constructor weak and add this file:
//   objcopy --weaken-symbol=_ZN6waffle15ReferenceStringC1Em --weaken-symbol=_ZN6waffle15ReferenceStringC2Em reference_string.o
// (every other member of the class keeps its reference body).
#include <cstdio>
#include <cstdlib>

#include <barretenberg/io/io.hpp>
#include <barretenberg/waffle/reference_string/reference_string.hpp>

#include "bbgpu.h"
#include "shim_stats.h"

namespace waffle
{
ReferenceString::ReferenceString(const size_t num_points)
{
    degree = num_points;
    if (num_points == 0)
    {
        monomials = nullptr;
        precomputed_g2_lines = nullptr;
        return;
    }
    int e = bbg_shim::ensure_library();
    bbg_shim::Timer timer("ReferenceString(n)");
    monomials = (barretenberg::g1::affine_element*)(aligned_alloc(64, sizeof(barretenberg::g1::affine_element) * (2 * degree + 2)));
    precomputed_g2_lines = (barretenberg::pairing::miller_lines*)(aligned_alloc(64, sizeof(barretenberg::pairing::miller_lines) * 2));

    // the file, with the reference's own readers for everything but the G1 points (io.hpp:140-181)
    std::vector<char> buffer = barretenberg::io::read_file_into_buffer(BARRETENBERG_SRS_PATH);
    barretenberg::io::Manifest manifest;
    const size_t manifest_size = sizeof(barretenberg::io::Manifest);
    if (buffer.size() < manifest_size)
    {
        fprintf(stderr, "bbgpu shim: cannot read the transcript %s\n", BARRETENBERG_SRS_PATH);
        abort();
    }
    barretenberg::io::read_manifest(buffer, manifest);
    const size_t g1_bytes = sizeof(barretenberg::fq::field_t) * 2 * (degree - 1);
    const size_t g2_offset = manifest_size + sizeof(barretenberg::fq::field_t) * 2 * manifest.num_g1_points;
    const size_t g2_bytes = sizeof(barretenberg::fq2::field_t) * 2 * 2;
    if (manifest.num_g1_points + 1 < degree || buffer.size() < g2_offset + g2_bytes)
    {
        fprintf(stderr, "bbgpu shim: transcript %s holds %u G1 points, %zu needed\n", BARRETENBERG_SRS_PATH, manifest.num_g1_points, degree - 1);
        abort();
    }
    if (e == 0) e = bbg_srs_from_transcript((const uint8_t*)&buffer[manifest_size], degree, (uint64_t*)monomials);
    if (e != 0)
    {
        fprintf(stderr, "bbgpu shim: ReferenceString(%zu) failed: %s (no CPU fallback)\n", degree, bbg_error_string(e));
        abort();
    }
    (void)g1_bytes;

    barretenberg::g2::affine_element* g2_buffer = (barretenberg::g2::affine_element*)(aligned_alloc(32, sizeof(barretenberg::g2::affine_element) * 2));
    barretenberg::io::read_g2_elements_from_buffer(g2_buffer, &buffer[g2_offset], g2_bytes);
    barretenberg::g2::copy_affine(g2_buffer[1], g2_x);
    aligned_free(g2_buffer);

    barretenberg::g2::element g2_x_jac;
    barretenberg::g2::affine_to_jacobian(g2_x, g2_x_jac);
    barretenberg::pairing::precompute_miller_lines(barretenberg::g2::one(), precomputed_g2_lines[0]);
    barretenberg::pairing::precompute_miller_lines(g2_x_jac, precomputed_g2_lines[1]);
}
} // namespace waffle
