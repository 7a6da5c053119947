"""CPU check of the HBM-resident PLONK prover rounds (barretenberg_b200/csrc/bbg_plonk.cu + shim/prover_gpu.cpp):
build/prover_gpu_emul is the reference's waffle prover with Prover::construct_proof replaced by the resident shim,
linked against the CPU kernel-EMULATION build of the library (tests/emul, test infrastructure only).  Its proof must
equal, field for field, the proof of the all-CPU reference build (build/prover_cpu), and the reference verifier must
accept it.  The same comparison runs on the real GPU library in tests/test_gpu_prover_dropin.py."""
import json
import os
import subprocess

import pytest

import helpers as H

B = os.path.join(H.ROOT, "build")


def have_binaries():
    return all(os.path.exists(os.path.join(B, f)) for f in ("make_srs", "prover_cpu", "prover_gpu_emul"))


@pytest.fixture(scope="module")
def srs():
    if not have_binaries():
        pytest.skip("build/prover_{cpu,gpu_emul} not built (needs the reference tree: make -C tests/cpp)")
    path = os.path.join(B, "srs", "transcript.dat")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    need = 64 * 1023 + 28 + 256 + 64
    if not os.path.exists(path) or os.path.getsize(path) < need:
        subprocess.check_call([os.path.join(B, "make_srs"), "1024", path], cwd=H.ROOT)
    return path


def run(binary, log_gates, composer="standard", env=None):
    e = dict(os.environ)
    e["OMP_NUM_THREADS"] = "4"
    e.update(env or {})
    out = subprocess.run([os.path.join(B, binary), str(log_gates), "1", composer], cwd=H.ROOT, capture_output=True, text=True, timeout=900, env=e)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("composer,log_gates", [("standard", 5), ("standard", 9), ("bool", 6), ("mimc", 6), ("extended", 7)])
def test_resident_prover_emulated_matches_cpu_reference(srs, composer, log_gates):
    """standard = arithmetic widget; bool = bool + arithmetic; mimc = MiMC + arithmetic (shifted output wire, selector in
    the opening polynomial); extended = bool + arithmetic + sequential (test/composer/test_*_composer.cpp circuits)"""
    cpu = run("prover_cpu", log_gates, composer)
    emu = run("prover_gpu_emul", log_gates, composer)
    assert cpu["verified"] and emu["verified"]
    assert emu["n"] == cpu["n"] and emu["widgets"] == cpu["widgets"]
    for k, v in cpu["proof"].items():
        assert emu["proof"][k] == v, k


def test_degenerate_circuit_with_infinity_commitment_verifies_emulated(srs):
    """0 / 1 witnesses only: T_HI is the point at infinity (unspecified limbs on the reference side, group.hpp:143-146, which
    enter the next challenge): the proof must verify on both builds and agree up to that commitment."""
    cpu = run("prover_cpu", 6, "bool_degenerate")
    emu = run("prover_gpu_emul", 6, "bool_degenerate")
    assert cpu["verified"] and emu["verified"]
    for k in ("W_L", "W_R", "W_O", "Z_1", "T_LO", "T_MID"):
        assert emu["proof"][k] == cpu["proof"][k], k
