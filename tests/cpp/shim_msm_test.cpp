// TEST INFRASTRUCTURE.  The reference's own tests of the Pippenger variants (test/test_scalar_multiplication.cpp:164-313:
// pippenger_low_memory, pippenger_internal_alt, precomputed_pippenger, batched_scalar_multiplication) restated against the
// GPU shims: each entry point is called through its reference C++ signature (shim/scalar_multiplication_gpu.cpp ->
// libbbgpu.so) and compared, normalised, with the reference's own CPU body of the same function, kept linkable under a
// cpu_reference_ name by tools/redefine_syms.py.  Prints one JSON object; exit code 0 iff everything matched.
//   shim_msm_test <num_points> [seed]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <barretenberg/curves/bn254/fr.hpp>
#include <barretenberg/curves/bn254/g1.hpp>
#include <barretenberg/curves/bn254/scalar_multiplication.hpp>

using namespace barretenberg;

namespace barretenberg
{
namespace scalar_multiplication
{
// the reference's CPU bodies (renamed symbols of scalar_multiplication.cpp)
g1::element cpu_reference_pippenger(fr::field_t* scalars, g1::affine_element* points, size_t num_initial_points, size_t forced_bucket_width);
g1::element cpu_reference_alt_pippenger(fr::field_t* scalars, g1::affine_element* points, size_t num_initial_points, size_t forced_bucket_width);
g1::element cpu_reference_pippenger_low_memory(fr::field_t* scalars, g1::affine_element* points, size_t num_points);
g1::element cpu_reference_pippenger_precomputed(fr::field_t* scalars, const std::vector<g1::affine_element*>& round_points, const size_t num_initial_points);
void cpu_reference_generate_pippenger_point_table(g1::affine_element* points, g1::affine_element* table, size_t num_points);
std::vector<g1::affine_element*> cpu_reference_generate_pippenger_precompute_table(g1::affine_element* points, g1::affine_element* table,
                                                                                   size_t num_points, size_t bits_per_bucket);
void cpu_reference_batched_scalar_multiplications(multiplication_state* mul_state, size_t num_batches);
} // namespace scalar_multiplication
} // namespace barretenberg

static uint64_t splitmix(uint64_t& s)
{
    uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

static bool same_point(const g1::element& a, const g1::element& b)
{
    const bool ia = g1::is_point_at_infinity(a), ib = g1::is_point_at_infinity(b);
    if (ia || ib) return ia == ib;
    g1::element ca = a, cb = b;
    g1::element x = g1::normalize(ca), y = g1::normalize(cb);
    return memcmp(&x.x, &y.x, 32) == 0 && memcmp(&x.y, &y.y, 32) == 0;
}

int main(int argc, char** argv)
{
    const size_t n = argc > 1 ? (size_t)atol(argv[1]) : 1000;
    uint64_t seed = argc > 2 ? (uint64_t)atol(argv[2]) : 7;
    fr::field_t* scalars = (fr::field_t*)aligned_alloc(32, sizeof(fr::field_t) * (n + 1));
    fr::field_t* scratch = (fr::field_t*)aligned_alloc(32, sizeof(fr::field_t) * (n + 1));
    g1::affine_element* plain = (g1::affine_element*)aligned_alloc(64, sizeof(g1::affine_element) * (n + 1));
    g1::affine_element* table = (g1::affine_element*)aligned_alloc(64, sizeof(g1::affine_element) * (2 * n + 2));
    g1::affine_element* table_cpu = (g1::affine_element*)aligned_alloc(64, sizeof(g1::affine_element) * (2 * n + 2));
    // points: an arithmetic progression of multiples of the generator (distinct, cheap); scalars: canonical 253-bit values
    {
        std::vector<g1::element> jac(n ? n : 1);
        fr::field_t k0 = { { splitmix(seed), splitmix(seed), splitmix(seed), splitmix(seed) >> 4 } };
        g1::affine_element base = g1::group_exponentiation(g1::affine_one(), k0);
        g1::element cur;
        g1::affine_to_jacobian(base, cur);
        for (size_t i = 0; i < n; ++i)
        {
            jac[i] = cur;
            g1::mixed_add(cur, g1::affine_one(), cur);
        }
        if (n) g1::batch_normalize(&jac[0], n);
        for (size_t i = 0; i < n; ++i)
        {
            fq::__copy(jac[i].x, plain[i].x);
            fq::__copy(jac[i].y, plain[i].y);
        }
    }
    for (size_t i = 0; i < n; ++i)
    {
        for (int k = 0; k < 4; ++k) scalars[i].data[k] = splitmix(seed);
        scalars[i].data[3] >>= 3;
    }
    if (n > 3)
    {
        scalars[1] = fr::zero;
        scalars[2] = fr::one;
    }
    memcpy(table, plain, sizeof(g1::affine_element) * n);
    memcpy(table_cpu, plain, sizeof(g1::affine_element) * n);
    scalar_multiplication::generate_pippenger_point_table(table, table, n);                    // GPU shim, in place as reference_string.cpp:23
    scalar_multiplication::cpu_reference_generate_pippenger_point_table(table_cpu, table_cpu, n);
    const bool ok_table = memcmp(table, table_cpu, sizeof(g1::affine_element) * 2 * n) == 0;

    memcpy(scratch, scalars, sizeof(fr::field_t) * n);
    const g1::element expect = n ? scalar_multiplication::cpu_reference_pippenger(scratch, table_cpu, n, 0) : g1::element{};
    const g1::element got_pip = scalar_multiplication::pippenger(scalars, table, n);
    const bool ok_pip = n == 0 ? g1::is_point_at_infinity(got_pip) : same_point(got_pip, expect);

    const g1::element got_alt = scalar_multiplication::alt_pippenger(scalars, table, n);
    memcpy(scratch, scalars, sizeof(fr::field_t) * n);
    const bool ok_alt = n == 0 ? g1::is_point_at_infinity(got_alt)
                               : (same_point(got_alt, scalar_multiplication::cpu_reference_alt_pippenger(scratch, table_cpu, n, 0)) && same_point(got_alt, expect));

    // pippenger_low_memory: n plain points; the reference overwrites its scalars, so it gets a copy
    const g1::element got_low = scalar_multiplication::pippenger_low_memory(scalars, plain, n);
    memcpy(scratch, scalars, sizeof(fr::field_t) * n);
    const bool ok_low = n == 0 ? g1::is_point_at_infinity(got_low)
                               : (same_point(got_low, scalar_multiplication::cpu_reference_pippenger_low_memory(scratch, plain, n)) && same_point(got_low, expect));

    // pippenger_precomputed over the reference's own pre-doubled tables
    bool ok_pre = true;
    if (n > 0)
    {
        const size_t bits = scalar_multiplication::get_optimal_bucket_width(n);
        const size_t rounds = (127 + bits) / (bits + 1);
        const size_t pre_bytes = sizeof(g1::affine_element) * n * (rounds - 1);
        g1::affine_element* pre = (g1::affine_element*)aligned_alloc(32, pre_bytes + 64);
        g1::affine_element* pre_cpu = (g1::affine_element*)aligned_alloc(32, pre_bytes + 64);
        // the table itself: device-built vs the reference's CPU body, byte for byte, and the same round pointers
        std::vector<g1::affine_element*> round_points = scalar_multiplication::generate_pippenger_precompute_table(plain, pre, n, bits);
        std::vector<g1::affine_element*> round_cpu = scalar_multiplication::cpu_reference_generate_pippenger_precompute_table(plain, pre_cpu, n, bits);
        ok_pre = round_points.size() == round_cpu.size() && round_points.size() == rounds && memcmp(pre, pre_cpu, pre_bytes) == 0;
        for (size_t r = 0; ok_pre && r < rounds; ++r)
            ok_pre = r + 1 == rounds ? (round_points[r] == plain && round_cpu[r] == plain) : (round_points[r] - pre == round_cpu[r] - pre_cpu);
        for (int rep = 0; rep < 2; ++rep) // second call: the device's own fixed-base tables are cached behind `plain`
        {
            const g1::element got_pre = scalar_multiplication::pippenger_precomputed(scalars, round_points, n);
            memcpy(scratch, scalars, sizeof(fr::field_t) * n);
            ok_pre = ok_pre && same_point(got_pre, scalar_multiplication::cpu_reference_pippenger_precomputed(scratch, round_cpu, n)) && same_point(got_pre, expect);
        }
        free(pre);
        free(pre_cpu);
    }

    // batched_scalar_multiplications: 5 MSMs (test_scalar_multiplication.cpp:272-313), outputs already normalised
    bool ok_batch = true;
    if (n > 0)
    {
        const size_t B = 5;
        std::vector<fr::field_t*> sc(B);
        scalar_multiplication::multiplication_state gpu_state[B], cpu_state[B];
        for (size_t b = 0; b < B; ++b)
        {
            sc[b] = (fr::field_t*)aligned_alloc(32, sizeof(fr::field_t) * n);
            for (size_t i = 0; i < n; ++i)
            {
                for (int k = 0; k < 4; ++k) sc[b][i].data[k] = splitmix(seed);
                sc[b][i].data[3] >>= 3;
            }
            gpu_state[b].points = table;
            gpu_state[b].scalars = sc[b];
            gpu_state[b].num_elements = n;
            cpu_state[b] = gpu_state[b];
            cpu_state[b].points = table_cpu;
        }
        scalar_multiplication::batched_scalar_multiplications(gpu_state, B);
        scalar_multiplication::cpu_reference_batched_scalar_multiplications(cpu_state, B);
        for (size_t b = 0; b < B; ++b)
        {
            ok_batch = ok_batch && same_point(gpu_state[b].output, cpu_state[b].output);
            // the shim leaves the normalised form the reference leaves (scalar_multiplication.cpp:766-771)
            if (!g1::is_point_at_infinity(cpu_state[b].output)) ok_batch = ok_batch && memcmp(&gpu_state[b].output, &cpu_state[b].output, 96) == 0;
            free(sc[b]);
        }
    }
    const bool ok = ok_table && ok_pip && ok_alt && ok_low && ok_pre && ok_batch;
    printf("{\"n\": %zu, \"ok\": %s, \"point_table\": %s, \"pippenger\": %s, \"alt_pippenger\": %s, \"pippenger_low_memory\": %s, "
           "\"pippenger_precomputed\": %s, \"batched_scalar_multiplications\": %s}\n",
           n, ok ? "true" : "false", ok_table ? "true" : "false", ok_pip ? "true" : "false", ok_alt ? "true" : "false", ok_low ? "true" : "false",
           ok_pre ? "true" : "false", ok_batch ? "true" : "false");
    free(scalars);
    free(scratch);
    free(plain);
    free(table);
    free(table_cpu);
    return ok ? 0 : 1;
}
