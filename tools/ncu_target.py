"""Target for `ncu --set full`: one warm-up round and one measured round of the dominant kernels — fft 2^20 and coset_fft 2^22
on a batch of 8 polynomials (two pass kernels each) and one MSM 2^20 over a registered table (fixed-base windows).
usage under ncu: -k regex:'ntt_pass|msm_accumulate' -s 5 -c 5 python tools/ncu_target.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import barretenberg_b200 as bb  # noqa: E402
from barretenberg_b200 import synthetic as S  # noqa: E402

lib = bb.Library()
n = 1 << 20
d_pts = lib.dev_alloc(n * 64)
d_tab = lib.dev_alloc(n * 128)
lib.generate_multiples_dev(S.to_limbs(S.mont(12345)), S.to_limbs(S.mont(777)), d_pts, n)
lib.generate_pippenger_point_table_dev(d_pts, d_tab, n)
h_tab = np.zeros((2 * n, 8), dtype=np.uint64)
lib.d2h(h_tab, d_tab)
lib.set_srs_precompute(True)
keep = lib.srs_register(h_tab)
d_fixed, c, w = lib.srs_device_table(keep)
d_sc = lib.dev_alloc(n * 32)
lib.h2d(d_sc, S.random_field(5, n))
x = np.random.default_rng(1).integers(0, 1 << 60, size=(8 * 4 * n, 4), dtype=np.uint64)
d_x = lib.dev_alloc(x.nbytes)
lib.h2d(d_x, x)
for _ in range(2):  # round 0: tables, workspaces; round 1: the one to look at
    lib.ntt_dev("fft", d_x, 20, batch=8)
    lib.ntt_dev("coset_fft", d_x, 22, batch=8)
    lib.msm_dev(d_sc, d_fixed, n)
lib.sync()
print("ncu target done: fixed-base c=%d windows=%d" % (c, w))
