"""BASELINE configs[4] at test size: the reference's unmodified waffle prover linked against the shims
(barretenberg_b200/shim -> libbbgpu.so) must produce the SAME proof, field for field, as the all-CPU reference
build, and the reference verifier must accept it.  Binaries are prebuilt by tests/cpp/Makefile (build/):
  prover_gpu          MSM / NTT shims + the HBM-resident Prover::construct_proof (shim/prover_gpu.cpp)
  prover_gpu_classic  MSM / NTT shims only, the reference's own round structure"""
import json
import os
import subprocess

import pytest

import helpers as H

pytestmark = pytest.mark.gpu
B = os.path.join(H.ROOT, "build")


def have_binaries():
    return all(os.path.exists(os.path.join(B, f)) for f in ("make_srs", "prover_cpu", "prover_gpu"))


@pytest.fixture(scope="module")
def srs():
    H.require_built("make_srs", "prover_cpu", "prover_gpu")  # -m gpu: missing prebuilt binaries fail, they do not skip
    path = os.path.join(B, "srs", "transcript.dat")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    need = 64 * 16383 + 28 + 256 + 64
    if not os.path.exists(path) or os.path.getsize(path) < need:
        subprocess.check_call([os.path.join(B, "make_srs"), "16384", path], cwd=H.ROOT)
    return path


def pow2_threads():
    """evaluation_domain silently needs a power-of-two OpenMP team (SURVEY.md §5 hazard): with 24 host cores the all-CPU
    reference prover produces a proof that does not verify"""
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    t = 1
    while t * 2 <= cores:
        t *= 2
    return t


def run(binary, log_gates, repeat=1, env=None, composer="standard"):
    e = dict(os.environ)
    e["OMP_NUM_THREADS"] = str(pow2_threads())
    e.update(env or {})
    out = subprocess.run([os.path.join(B, binary), str(log_gates), str(repeat), composer], cwd=H.ROOT, capture_output=True, text=True, timeout=600,
                         env=e)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("log_gates", [5, 8, 12, 14])  # 2^4 gives the recipe zero gates: the reference itself throws bad_alloc
@pytest.mark.parametrize("repeat,env", [(1, {}), (2, {}), (2, {"BBG_PLONK_KEY_CACHE": "0"})])
def test_prover_gpu_matches_cpu_reference(srs, log_gates, repeat, env):
    """repeat 1: cold proving key; repeat 2: second proof after Prover::reset() with the circuit constants found unchanged
    on the device (proving-key cache), or rebuilt when the cache is switched off"""
    cpu = run("prover_cpu", log_gates)
    gpu = run("prover_gpu", log_gates, repeat=repeat, env=env)
    assert cpu["verified"] and gpu["verified"]
    assert gpu["n"] == cpu["n"]
    for k, v in cpu["proof"].items():
        assert gpu["proof"][k] == v, k


@pytest.mark.parametrize("composer", ["bool", "mimc", "extended"])
@pytest.mark.parametrize("log_gates", [7, 11, 13])
def test_resident_prover_other_widget_mixes(srs, log_gates, composer):
    """bool / MiMC / sequential widgets inside the resident rounds (test/composer/test_*_composer.cpp circuits, scaled)"""
    cpu = run("prover_cpu", log_gates, composer=composer)
    gpu = run("prover_gpu", log_gates, repeat=2, composer=composer)
    assert cpu["verified"] and gpu["verified"]
    assert gpu["n"] == cpu["n"] and gpu["widgets"] == cpu["widgets"] >= 2
    for k, v in cpu["proof"].items():
        assert gpu["proof"][k] == v, k


@pytest.mark.parametrize("binary,env", [("prover_gpu_classic", {}), ("prover_gpu", {"BBG_PLONK_RESIDENT": "0"})])
@pytest.mark.parametrize("log_gates", [5, 12])
def test_prover_classic_round_structure_matches_cpu_reference(srs, log_gates, binary, env):
    """the ten-entry-point drop-in alone (reference round structure), also reachable from the resident build"""
    H.require_built(binary)
    cpu = run("prover_cpu", log_gates)
    gpu = run(binary, log_gates, repeat=2, env=env)
    assert cpu["verified"] and gpu["verified"]
    for k, v in cpu["proof"].items():
        assert gpu["proof"][k] == v, k


@pytest.mark.parametrize("log_gates,repeat,composer", [(12, 60, "standard"), (14, 30, "mimc"), (14, 30, "extended")])
def test_resident_prover_is_reproducible_across_streams(srs, log_gates, repeat, composer):
    """The resident rounds run on four streams (work, second NTT stream, two commitment streams) plus an upload thread.
    The prover draws no randomness, so every repetition must reproduce the first proof bit for bit: a missing ordering
    between streams would show up here (the harness counts repetitions that differ)."""
    gpu = run("prover_gpu", log_gates, repeat=repeat, composer=composer)
    assert gpu["verified"] and gpu["repeat_mismatches"] == 0
    cpu = run("prover_cpu", log_gates, composer=composer)
    for k, v in cpu["proof"].items():
        assert gpu["proof"][k] == v, k


@pytest.mark.parametrize("log_gates,composer", [(14, "standard"), (16, "standard"), (15, "mimc")])
def test_resident_prover_on_every_gpu_of_the_box(srs, log_gates, composer):
    """BBG_NUM_GPUS = all devices: the shim brings the library up with bbg_init_multi, the SRS is replicated and every
    commitment of the resident rounds is cut into point ranges, one per device (scalar_multiplication.cpp:703-728 cuts per
    thread).  The proof must stay identical to the all-CPU reference's, and reproducible.  On a one-GPU box this is the
    single-device path (a one-device list), still checked."""
    import torch

    g = torch.cuda.device_count()
    need = 64 * ((1 << log_gates) - 1) + 28 + 256 + 64
    if os.path.getsize(srs) < need:
        subprocess.check_call([os.path.join(B, "make_srs"), str(1 << log_gates), srs], cwd=H.ROOT)
    env = {"BBG_NUM_GPUS": str(g), "BBG_MULTI_MIN_POINTS": "4096", "BBG_MULTI_MIN_SHARD": "512"}
    cpu = run("prover_cpu", log_gates, composer=composer)
    gpu = run("prover_gpu", log_gates, repeat=4, env=env, composer=composer)
    assert cpu["verified"] and gpu["verified"] and gpu["repeat_mismatches"] == 0
    for k, v in cpu["proof"].items():
        assert gpu["proof"][k] == v, k
    classic = run("prover_gpu_classic", log_gates, repeat=2, env=env, composer=composer)
    assert classic["verified"]
    for k, v in cpu["proof"].items():
        assert classic["proof"][k] == v, k


@pytest.mark.parametrize("binary", ["prover_gpu", "prover_gpu_classic"])
@pytest.mark.parametrize("log_gates", [6, 10])
def test_degenerate_circuit_with_infinity_commitment_verifies(srs, log_gates, binary):
    """A circuit of 0 / 1 witnesses only (the reference's own bool-composer test circuit): the quotient has degree < 2n, so
    T_HI is the point at infinity, whose coordinates the reference leaves unspecified (groups/group.hpp:143-146) and then
    hashes into the next challenge (waffle/proof_system/challenge.hpp:15-23).  Byte identity is therefore only defined up to
    that commitment; what must hold on both builds: the reference verifier accepts the proof, and everything committed
    before the infinity point is identical."""
    H.require_built(binary)
    cpu = run("prover_cpu", log_gates, composer="bool_degenerate")
    gpu = run(binary, log_gates, repeat=2, composer="bool_degenerate")
    assert cpu["verified"], "the reference rejects its own degenerate proof"
    assert gpu["verified"] and gpu["repeat_mismatches"] == 0
    assert gpu["n"] == cpu["n"]
    for k in ("W_L", "W_R", "W_O", "Z_1", "T_LO", "T_MID"):
        assert gpu["proof"][k] == cpu["proof"][k], k
