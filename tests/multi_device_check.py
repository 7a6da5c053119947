"""Run by tests/test_gpu_parity.py::test_multi_device_instance in its own process: one library instance driving every
GPU of the box (bbg_init_multi) — MSMs over registered point tables are cut into point ranges, one per device
(scalar_multiplication.cpp:703-728 cuts them per thread), and must give the same normalised points as the closed form
sum k_i (a0 + i d) G = ((sum k_i (a0 + i d)) mod r) G and as the compiled reference."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import barretenberg_b200 as bb  # noqa: E402
import helpers as H  # noqa: E402
from barretenberg_b200 import synthetic as S  # noqa: E402


def main():
    g = torch.cuda.device_count()
    assert g >= 1
    os.environ.setdefault("BBG_MULTI_MIN_POINTS", "4096")
    os.environ.setdefault("BBG_MULTI_MIN_SHARD", "512")
    lib = bb.Library(devices=list(range(g)))
    assert lib.device_count() == g
    for log_n, precompute in ((13, False), (16, False), (18, False), (14, True), (17, True), (19, True)):
        # precompute: every device also builds the fixed-base windows of its replica when the table is registered
        # (bbg_set_srs_precompute) and the per-device point ranges take the one-bucket-set form
        lib.set_srs_precompute(precompute)
        n = 1 << log_n
        a0, d = 0x1234567 + log_n, 0x89ABC
        # table built on the primary device, then registered from the host copy (what ReferenceString + the shim do)
        d_points = lib.dev_alloc(n * 64)
        d_table = lib.dev_alloc(n * 128)
        lib.generate_multiples_dev(S.to_limbs(S.mont(a0)), S.to_limbs(S.mont(d)), d_points, n)
        lib.generate_pippenger_point_table_dev(d_points, d_table, n)
        table = np.zeros((2 * n, 8), dtype=np.uint64)
        lib.d2h(table, d_table)
        lib.dev_free(d_points)
        lib.dev_free(d_table)
        keep = lib.srs_register(table)
        if precompute and n >= 1024:
            assert lib.srs_device_table(keep)[2] >= 6, "fixed-base windows were not built"
        scs = [S.random_field(100 * log_n + i, n) for i in range(4)]
        scs[1][: n // 2] = scs[1][0]  # one giant bucket per window in the first half
        scs[2][:] = 0
        scs[2][n - 1] = S.random_field(5, 1)[0]  # only the last device's range is non-zero
        expect = [H.closed_form_msm(s, a0, d) for s in scs]
        for i, s in enumerate(scs):
            assert (lib.msm(s, table) == expect[i]).all(), ("host msm", log_n, i)
        got = lib.msm_batched(scs, [table] * 4)
        for i in range(4):
            assert (got[i] == expect[i]).all(), ("batched", log_n, i)
        # sub-range of the registered table (batched_scalar_multiplications passes &points[2 * offset])
        off = n // 4 + 3
        m = n // 2
        sub = lib.msm(np.ascontiguousarray(scs[0][:m]), table[2 * off:], m)
        assert (sub == H.closed_form_msm(scs[0][:m], a0 + off * d, d)).all(), ("sub-range", log_n)
        # launched (host scalars) with transforms in between
        tickets = [lib.msm_launch(scs[i], table, n) for i in (0, 3)]
        x = S.random_field(9, 1 << 12)
        assert (lib.ntt("ifft", lib.ntt("fft", x.copy())) == x).all()
        for t, i in zip(tickets, (0, 3)):
            assert (lib.msm_finish(t) == expect[i]).all(), ("launch", log_n, i)
        lib.srs_unregister(table)
    lib.set_srs_precompute(False)
    if H.have_ref():
        n = 1 << 16
        table, a0, d = H.generator_multiples_table(3, n)
        sc = H.random_scalars_mont(4, n)
        lib.srs_register(table)
        assert (lib.msm(sc, table) == H.oracle_msm(sc, table)).all()
        lib.srs_unregister(table)
    print("multi-device check ok: %d device(s)" % g)


if __name__ == "__main__":
    main()
