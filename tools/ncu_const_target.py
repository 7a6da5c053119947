"""Target for ncu on the degenerate MSM: every scalar equal (one giant bucket per window and half scalar), 2^20 points,
plain windows.  usage under ncu: -k regex:'msm_accumulate|msm_fixup' -s 2 -c 3 python tools/ncu_const_target.py"""
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
sc = S.random_field(5, n)
sc = np.ascontiguousarray(np.tile(sc[0], (n, 1)))
d_sc = lib.dev_alloc(n * 32)
lib.h2d(d_sc, sc)
for _ in range(2):
    lib.msm_dev(d_sc, d_tab, n)
lib.sync()
print("done")
