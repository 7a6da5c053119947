// TEST / BENCH INFRASTRUCTURE for BASELINE configs[4]: the reference's own waffle StandardComposer prover,
// unmodified, driven end to end.  This one source file is linked twice by tests/cpp/Makefile:
//   build/prover_cpu : every object from the reference sources               (the CPU baseline)
//   build/prover_gpu : the same objects, except that pippenger / batched_scalar_multiplications /
//                      generate_pippenger_point_table and the seven fft entry points come from
//                      barretenberg_b200/shim/*.cpp -> libbbgpu.so           (the drop-in under test)
// Circuit: the one of test/benchmarks/bench_plonk.cpp:25-37, with SEEDED witnesses instead of getentropy so both
// binaries prove the same statement; the prover draws no randomness, so the two proofs must be identical.
//   usage: prover_harness <log2_gates> [repeat]
// prints one JSON line: sizes, timings, verify result and every proof element (hex limbs).
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include <barretenberg/curves/bn254/fr.hpp>
#include <barretenberg/waffle/composer/standard_composer.hpp>
#include <barretenberg/waffle/proof_system/preprocess.hpp>
#include <barretenberg/waffle/proof_system/prover/prover.hpp>
#include <barretenberg/waffle/proof_system/verifier/verifier.hpp>
#include <barretenberg/waffle/stdlib/field/field.hpp>

using namespace barretenberg;

// present only in the shim build (barretenberg_b200/shim/scalar_multiplication_gpu.cpp)
extern "C" void bbg_shim_report(void) __attribute__((weak));

static uint64_t sm_state = 0x853c49e6748fea9bULL;
static uint64_t splitmix()
{
    uint64_t z = (sm_state += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static fr::field_t seeded_element()
{
    fr::field_t r = { { splitmix(), splitmix(), splitmix(), splitmix() & 0x0fffffffffffffffULL } };
    fr::__to_montgomery_form(r, r);
    return r;
}

static void generate_test_plonk_circuit(waffle::StandardComposer& composer, size_t num_gates)
{
    plonk::stdlib::field_t<waffle::StandardComposer> a(plonk::stdlib::witness_t(&composer, seeded_element()));
    plonk::stdlib::field_t<waffle::StandardComposer> b(plonk::stdlib::witness_t(&composer, seeded_element()));
    plonk::stdlib::field_t<waffle::StandardComposer> c(&composer);
    for (size_t i = 0; i < (num_gates / 4) - 4; ++i)
    {
        c = a + b;
        c = a * c;
        a = b * b;
        b = c * c;
    }
}

static void print_fe(const char* name, const uint64_t* d, bool last = false)
{
    printf("\"%s\": \"%016lx%016lx%016lx%016lx\"%s", name, d[3], d[2], d[1], d[0], last ? "" : ", ");
}
static void print_pt(const char* name, const g1::affine_element& p)
{
    printf("\"%s\": \"%016lx%016lx%016lx%016lx:%016lx%016lx%016lx%016lx\", ", name, p.x.data[3], p.x.data[2], p.x.data[1], p.x.data[0], p.y.data[3],
           p.y.data[2], p.y.data[1], p.y.data[0]);
}
static double ms_since(std::chrono::steady_clock::time_point t0)
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

int main(int argc, char** argv)
{
    const size_t log_gates = argc > 1 ? strtoull(argv[1], nullptr, 10) : 12;
    const int repeat = argc > 2 ? atoi(argv[2]) : 1;
    const size_t num_gates = (size_t)1 << log_gates;
    {
        // the reference's transcript reader does not check the file size (io.hpp:157-182): refuse to run into a short SRS
        FILE* f = fopen(BARRETENBERG_SRS_PATH, "rb");
        long have = 0;
        if (f != nullptr)
        {
            fseek(f, 0, SEEK_END);
            have = ftell(f);
            fclose(f);
        }
        const long need = 28 + 64 * (long)(num_gates - 1) + 256;
        if (have < need)
        {
            fprintf(stderr, "prover_harness: %s holds %ld bytes, %ld needed for 2^%zu gates (run build/make_srs %zu %s)\n", BARRETENBERG_SRS_PATH, have, need,
                    log_gates, num_gates, BARRETENBERG_SRS_PATH);
            return 2;
        }
    }

    auto t0 = std::chrono::steady_clock::now();
    waffle::StandardComposer composer = waffle::StandardComposer(num_gates);
    generate_test_plonk_circuit(composer, num_gates);
    waffle::Prover prover = composer.preprocess();
    const double setup_ms = ms_since(t0);

    t0 = std::chrono::steady_clock::now();
    waffle::Verifier verifier = waffle::preprocess(prover);
    const double vk_ms = ms_since(t0);

    waffle::plonk_proof proof;
    double best_prove_ms = 1e300, first_prove_ms = 0;
    for (int r = 0; r < repeat; ++r)
    {
        if (r > 0) prover.reset();
        t0 = std::chrono::steady_clock::now();
        proof = prover.construct_proof();
        const double ms = ms_since(t0);
        if (r == 0) first_prove_ms = ms;
        if (ms < best_prove_ms) best_prove_ms = ms;
    }
    t0 = std::chrono::steady_clock::now();
    const bool ok = verifier.verify_proof(proof);
    const double verify_ms = ms_since(t0);

    printf("{\"log2_gates\": %zu, \"n\": %zu, \"setup_ms\": %.3f, \"verifier_key_ms\": %.3f, \"prove_ms_first\": %.3f, \"prove_ms_best\": %.3f, "
           "\"repeat\": %d, \"verify_ms\": %.3f, \"verified\": %s, \"proof\": {",
           log_gates, prover.n, setup_ms, vk_ms, first_prove_ms, best_prove_ms, repeat, verify_ms, ok ? "true" : "false");
    print_pt("W_L", proof.W_L);
    print_pt("W_R", proof.W_R);
    print_pt("W_O", proof.W_O);
    print_pt("Z_1", proof.Z_1);
    print_pt("T_LO", proof.T_LO);
    print_pt("T_MID", proof.T_MID);
    print_pt("T_HI", proof.T_HI);
    print_pt("PI_Z", proof.PI_Z);
    print_pt("PI_Z_OMEGA", proof.PI_Z_OMEGA);
    print_fe("w_l_eval", proof.w_l_eval.data);
    print_fe("w_r_eval", proof.w_r_eval.data);
    print_fe("w_o_eval", proof.w_o_eval.data);
    print_fe("sigma_1_eval", proof.sigma_1_eval.data);
    print_fe("sigma_2_eval", proof.sigma_2_eval.data);
    print_fe("z_1_shifted_eval", proof.z_1_shifted_eval.data);
    print_fe("linear_eval", proof.linear_eval.data, true);
    printf("}}\n");
    if (bbg_shim_report) bbg_shim_report();
    return ok ? 0 : 1;
}
