"""Fr/Fq mul_const (product with a precomputed quotient, bbg_field.cuh) against the Montgomery product, on the host path
of the same header: random multiplicands over the whole allowed range [0, 4p), edge constants, result range [0, 2p);
and the dedicated square sqr(a) against mul(a, a) as integers (random a in [0, 2p) plus carry / shifted-bit edge cases).
The device paths (PTX carry chains) are compared with the Montgomery product on the GPU by tools/mulbench.cu and, through
the NTT and the MSM, by every GPU parity test."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_mul_const_and_sqr_match_montgomery_product(tmp_path):
    exe = str(tmp_path / "field_const_mul_test")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "field_const_mul_test.cpp")], check=True)
    out = subprocess.run([exe, "60000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "Fr: 60000 cases, 0 bad" in out.stdout and "Fq: 60000 cases, 0 bad" in out.stdout
    assert "Fr sqr: 60000 cases, 0 bad" in out.stdout and "Fq sqr: 60000 cases, 0 bad" in out.stdout
