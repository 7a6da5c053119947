"""The reference's own tests of the Pippenger entry points (test/test_scalar_multiplication.cpp:113-313: pippenger,
pippenger_low_memory, pippenger_internal_alt, precomputed_pippenger, batched_scalar_multiplication) against the C++ shims:
build/shim_msm_test calls every replaced function through its reference signature (shim/scalar_multiplication_gpu.cpp ->
libbbgpu.so) and compares it with the reference's own CPU body of the same function (kept under a cpu_reference_ symbol
by tools/redefine_syms.py).  `_emul` = the same binary linked against the CPU kernel-emulation build."""
import json
import os
import subprocess

import pytest

import helpers as H

B = os.path.join(H.ROOT, "build")
FIELDS = ("point_table", "pippenger", "alt_pippenger", "pippenger_low_memory", "pippenger_precomputed", "batched_scalar_multiplications")


def run(binary, n, seed=7, env=None):
    e = dict(os.environ)
    e.setdefault("OMP_NUM_THREADS", "8")
    e.update(env or {})
    out = subprocess.run([os.path.join(B, binary), str(n), str(seed)], cwd=H.ROOT, capture_output=True, text=True, timeout=900, env=e)
    line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else ""
    assert out.returncode == 0, (out.returncode, line, out.stderr[-1500:])
    return json.loads(line)


@pytest.mark.parametrize("n", [0, 1, 2, 37, 1000])
def test_shim_entry_points_emulation(n):
    if not os.path.exists(os.path.join(B, "shim_msm_test_emul")):
        pytest.skip("build/shim_msm_test_emul not built (needs the reference tree: make -C tests/cpp)")
    r = run("shim_msm_test_emul", n)
    assert r["ok"] and all(r[f] for f in FIELDS), r


@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 37, 1000, 10000, 1 << 16])
def test_shim_entry_points_gpu(n):
    H.require_built("shim_msm_test")
    r = run("shim_msm_test", n)
    assert r["n"] == n and r["ok"] and all(r[f] for f in FIELDS), r


@pytest.mark.gpu
def test_shim_entry_points_gpu_all_devices():
    """The same calls with every GPU of the box behind the shim (BBG_NUM_GPUS, bbg_init_multi): point ranges fanned out
    per device, identical results.  One device: the single-GPU path, still checked."""
    import torch

    H.require_built("shim_msm_test")
    g = torch.cuda.device_count()
    r = run("shim_msm_test", 1 << 17, seed=11, env={"BBG_NUM_GPUS": str(g), "BBG_MULTI_MIN_POINTS": "4096", "BBG_MULTI_MIN_SHARD": "512"})
    assert r["ok"] and all(r[f] for f in FIELDS), r
