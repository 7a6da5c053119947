"""Target for `ncu --set full` on the pair-sum rounds: two MSMs 2^20 over a registered table (fixed-base windows).
usage under ncu: -k regex:'msm_pair_round|msm_accumulate' -s 4 -c 4 python tools/ncu_pair_target.py"""
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
for _ in range(2):
    lib.msm_dev(d_sc, d_fixed, n)
lib.sync()
print("ncu target done: fixed-base c=%d windows=%d" % (c, w))
