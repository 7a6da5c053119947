// TEST / BENCH INFRASTRUCTURE for BASELINE configs[4]: the reference's own waffle StandardComposer prover,
// unmodified, driven end to end.  This one source file is linked twice by tests/cpp/Makefile:
//   build/prover_cpu : every object from the reference sources               (the CPU baseline)
//   build/prover_gpu : the same objects, except that pippenger / batched_scalar_multiplications /
//                      generate_pippenger_point_table and the seven fft entry points come from
//                      barretenberg_b200/shim/*.cpp -> libbbgpu.so           (the drop-in under test)
// Circuit: the one of test/benchmarks/bench_plonk.cpp:25-37, with SEEDED witnesses instead of getentropy so both
// binaries prove the same statement; the prover draws no randomness, so the two proofs must be identical.
//   usage: prover_harness <log2_gates> [repeat] [standard|bool|bool_degenerate|mimc|extended]
// The three other composers exercise the bool / MiMC / sequential widgets with the circuits of the reference's own
// composer tests (test/composer/test_{bool,mimc,extended}_composer.cpp), scaled to ~2^log2_gates gates and seeded.
// prints one JSON line: sizes, timings, verify result and every proof element (hex limbs).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include <barretenberg/curves/bn254/fr.hpp>
#include <barretenberg/waffle/composer/bool_composer.hpp>
#include <barretenberg/waffle/composer/extended_composer.hpp>
#include <barretenberg/waffle/composer/mimc_composer.hpp>
#include <barretenberg/waffle/composer/standard_composer.hpp>
#include <barretenberg/waffle/stdlib/uint32/uint32.hpp>
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

// test_bool_composer.cpp:110-137, with seeded bits
// full_size = false: the reference's own circuit (0 / 1 witnesses only).  Its quotient has degree < 2n, so T_HI is the point
// at infinity, whose limbs the reference leaves unspecified (group.hpp:143-146) and hashes into the next challenge
// (challenge.hpp:15-23): such proofs verify on both sides but are not comparable beyond that commitment ("bool_degenerate").
static void generate_bool_circuit(waffle::BoolComposer& composer, size_t num_gates, bool full_size = true)
{
    for (size_t i = 0; i + 5 <= num_gates - 2; i += 5)
    {
        if (!full_size)
        {
            const uint64_t bits = splitmix();
            fr::field_t a = (bits & 1) ? fr::one : fr::zero;
            fr::field_t b = (bits & 2) ? fr::one : fr::zero;
            uint32_t a_idx = composer.add_variable(a);
            uint32_t b_idx = composer.add_variable(b);
            uint32_t c_idx = composer.add_variable(fr::add(a, b));
            composer.create_bool_gate(a_idx);
            composer.create_bool_gate(b_idx);
            composer.create_add_gate({ a_idx, b_idx, c_idx, fr::one, fr::one, fr::neg_one(), fr::zero });
            i -= 2; // (three gates per turn instead of five)
            continue;
        }
        // two gates on full-size field elements keep every commitment of the proof away from the point at infinity
        // (the reference leaves the limbs of an infinity commitment unspecified, and they enter its transcript)
        fr::field_t d = seeded_element(), e = seeded_element();
        uint32_t d_idx = composer.add_variable(d);
        uint32_t e_idx = composer.add_variable(e);
        uint32_t f_idx = composer.add_variable(fr::add(d, e));
        uint32_t g_idx = composer.add_variable(fr::mul(d, e));
        composer.create_add_gate({ d_idx, e_idx, f_idx, fr::one, fr::one, fr::neg_one(), fr::zero });
        composer.create_mul_gate({ d_idx, e_idx, g_idx, fr::one, fr::neg_one(), fr::zero });
        const uint64_t bits = splitmix();
        fr::field_t a = (bits & 1) ? fr::one : fr::zero;
        fr::field_t b = (bits & 2) ? fr::one : fr::zero;
        fr::field_t c = fr::add(a, b);
        uint32_t a_idx = composer.add_variable(a);
        uint32_t b_idx = composer.add_variable(b);
        uint32_t c_idx = composer.add_variable(c);
        composer.create_bool_gate(a_idx);
        composer.create_bool_gate(b_idx);
        composer.create_add_gate({ a_idx, b_idx, c_idx, fr::one, fr::one, fr::neg_one(), fr::zero });
    }
}
// test_mimc_composer.cpp:14-45, with seeded round constants
static void generate_mimc_circuit(waffle::MiMCComposer& composer, size_t num_gates)
{
    const size_t rounds = num_gates > 8 ? num_gates - 5 : 3;
    fr::field_t x = seeded_element();
    fr::field_t k = seeded_element();
    uint32_t x_in_idx = composer.add_variable(x);
    uint32_t k_idx = composer.add_variable(k);
    for (size_t i = 0; i < rounds; ++i)
    {
        fr::field_t c = seeded_element();
        fr::field_t T0 = fr::add(fr::add(x, k), c);
        fr::field_t x_cubed = fr::sqr(T0);
        x_cubed = fr::mul(x_cubed, T0);
        uint32_t x_cubed_idx = composer.add_variable(x_cubed);
        fr::field_t x_out = fr::sqr(x_cubed);
        x_out = fr::mul(x_out, T0);
        uint32_t x_out_idx = composer.add_variable(x_out);
        composer.create_mimc_gate({ x_in_idx, x_cubed_idx, k_idx, x_out_idx, c });
        x_in_idx = x_out_idx;
        x = x_out;
    }
}
// test_extended_composer.cpp:252-271, :356-372: uint32 arithmetic and logic (adjacent gates get merged: q_o_next)
static void generate_extended_circuit(waffle::ExtendedComposer& composer, size_t num_gates)
{
    typedef plonk::stdlib::uint32<waffle::ExtendedComposer> uint32;
    typedef plonk::stdlib::witness_t<waffle::ExtendedComposer> witness_t;
    const size_t blocks = num_gates / 512 > 0 ? num_gates / 512 : 1; // one block is a few hundred gates
    for (size_t i = 0; i < blocks; ++i)
    {
        uint32 a = witness_t(&composer, (uint32_t)splitmix());
        uint32 b = witness_t(&composer, (uint32_t)splitmix());
        uint32 c = a * b;
        uint32 d = a * c;
        uint32 e = (~d) & b;
        uint32 f = e + 1;
        c.get_witness_index();
        d.get_witness_index();
        f.get_witness_index();
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

template <typename Composer> static int run(Composer& composer, size_t log_gates, int repeat, double circuit_ms, const char* kind)
{
    auto t0 = std::chrono::steady_clock::now();
    waffle::Prover prover = composer.preprocess();
    const double setup_ms = circuit_ms + ms_since(t0);

    t0 = std::chrono::steady_clock::now();
    waffle::Verifier verifier = waffle::preprocess(prover);
    const double vk_ms = ms_since(t0);

    waffle::plonk_proof proof, first_proof;
    double best_prove_ms = 1e300, first_prove_ms = 0;
    int repeat_mismatches = 0; // the prover draws no randomness: every repetition must reproduce the first proof exactly
    for (int r = 0; r < repeat; ++r)
    {
        if (r > 0) prover.reset();
        t0 = std::chrono::steady_clock::now();
        proof = prover.construct_proof();
        const double ms = ms_since(t0);
        if (r == 0) first_prove_ms = ms;
        if (ms < best_prove_ms) best_prove_ms = ms;
        if (r == 0) first_proof = proof;
        else if (memcmp(&proof.W_L, &first_proof.W_L, 9 * sizeof(g1::affine_element)) != 0 ||
                 memcmp(&proof.w_l_eval, &first_proof.w_l_eval, 7 * sizeof(fr::field_t)) != 0)
            ++repeat_mismatches;
    }
    t0 = std::chrono::steady_clock::now();
    const bool ok = verifier.verify_proof(proof);
    const double verify_ms = ms_since(t0);

    printf("{\"log2_gates\": %zu, \"composer\": \"%s\", \"widgets\": %zu, \"n\": %zu, \"setup_ms\": %.3f, \"verifier_key_ms\": %.3f, \"prove_ms_first\": %.3f, "
           "\"prove_ms_best\": %.3f, \"repeat\": %d, \"repeat_mismatches\": %d, \"verify_ms\": %.3f, \"verified\": %s, \"proof\": {",
           log_gates, kind, prover.widgets.size(), prover.n, setup_ms, vk_ms, first_prove_ms, best_prove_ms, repeat, repeat_mismatches, verify_ms,
           ok ? "true" : "false");
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
    // plonk_proof fields the prover only writes when a widget asks for them (uninitialised memory otherwise)
    bool has_shifted = false;
    for (size_t i = 0; i < prover.widgets.size(); ++i)
        has_shifted |= prover.widgets[i]->version.has_dependency(waffle::WidgetVersionControl::Dependencies::REQUIRES_W_O_SHIFTED);
    if (has_shifted) print_fe("w_o_shifted_eval", proof.w_o_shifted_eval.data);
    if (strcmp(kind, "mimc") == 0) print_fe("q_mimc_coefficient_eval", proof.q_mimc_coefficient_eval.data);
    print_fe("linear_eval", proof.linear_eval.data, true);
    printf("}}\n");
    if (bbg_shim_report) bbg_shim_report();
    return (ok && repeat_mismatches == 0) ? 0 : 1;
}

int main(int argc, char** argv)
{
    const size_t log_gates = argc > 1 ? strtoull(argv[1], nullptr, 10) : 12;
    const int repeat = argc > 2 ? atoi(argv[2]) : 1;
    const std::string kind = argc > 3 ? argv[3] : "standard";
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
    if (kind == "standard")
    {
        waffle::StandardComposer composer = waffle::StandardComposer(num_gates);
        generate_test_plonk_circuit(composer, num_gates);
        return run(composer, log_gates, repeat, ms_since(t0), kind.c_str());
    }
    if (kind == "bool")
    {
        waffle::BoolComposer composer = waffle::BoolComposer();
        generate_bool_circuit(composer, num_gates);
        return run(composer, log_gates, repeat, ms_since(t0), kind.c_str());
    }
    if (kind == "bool_degenerate")
    {
        waffle::BoolComposer composer = waffle::BoolComposer();
        generate_bool_circuit(composer, num_gates, false);
        return run(composer, log_gates, repeat, ms_since(t0), kind.c_str());
    }
    if (kind == "mimc")
    {
        waffle::MiMCComposer composer = waffle::MiMCComposer(num_gates);
        generate_mimc_circuit(composer, num_gates);
        return run(composer, log_gates, repeat, ms_since(t0), kind.c_str());
    }
    if (kind == "extended")
    {
        waffle::ExtendedComposer composer = waffle::ExtendedComposer();
        generate_extended_circuit(composer, num_gates);
        return run(composer, log_gates, repeat, ms_since(t0), kind.c_str());
    }
    fprintf(stderr, "unknown composer %s\n", kind.c_str());
    return 2;
}
