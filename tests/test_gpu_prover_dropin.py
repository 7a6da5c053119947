"""BASELINE configs[4] at test size: the reference's unmodified waffle prover linked against the shims
(barretenberg_b200/shim -> libbbgpu.so) must produce the SAME proof, field for field, as the all-CPU reference
build, and the reference verifier must accept it.  Binaries are prebuilt by tests/cpp/Makefile (build/)."""
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
    if not have_binaries():
        pytest.skip("build/prover_{cpu,gpu} not built (needs the reference tree: make -C tests/cpp)")
    path = os.path.join(B, "srs", "transcript.dat")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    need = 64 * 16383 + 28 + 256 + 64
    if not os.path.exists(path) or os.path.getsize(path) < need:
        subprocess.check_call([os.path.join(B, "make_srs"), "16384", path], cwd=H.ROOT)
    return path


def run(binary, log_gates):
    out = subprocess.run([os.path.join(B, binary), str(log_gates)], cwd=H.ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("log_gates", [5, 8, 12, 14])  # 2^4 gives the recipe zero gates: the reference itself throws bad_alloc
def test_prover_gpu_matches_cpu_reference(srs, log_gates):
    cpu = run("prover_cpu", log_gates)
    gpu = run("prover_gpu", log_gates)
    assert cpu["verified"] and gpu["verified"]
    assert gpu["n"] == cpu["n"]
    for k, v in cpu["proof"].items():
        assert gpu["proof"][k] == v, k
