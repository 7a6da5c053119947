"""Quick device-resident timings (CUDA events on the library stream) — development aid, not the bench."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import barretenberg_b200 as bb  # noqa: E402
import helpers as H  # noqa: E402

lib = bb.default_library()
res = {}
for log_n in (16, 20, 22):
    n = 1 << log_n
    x = H.random_scalars_mont(1, n)
    d = lib.dev_alloc(x.nbytes)
    lib.h2d(d, x)
    for op in ("fft", "ifft", "coset_fft"):
        lib.ntt_dev(op, d, log_n)
        lib.sync()
        best = 1e9
        for _ in range(5):
            lib.timer_start()
            lib.ntt_dev(op, d, log_n)
            best = min(best, lib.timer_stop())
        res["ntt_%s_2^%d_ms" % (op, log_n)] = best
    lib.dev_free(d)
for log_n in (16, 20):
    n = 1 << log_n
    t0 = time.time()
    table, a0, dd = H.generator_multiples_table(77, n)
    sc = H.random_scalars_mont(78, n)
    res["gen_table_2^%d_s" % log_n] = time.time() - t0
    d_s, d_t = lib.dev_alloc(sc.nbytes), lib.dev_alloc(table.nbytes)
    lib.h2d(d_s, sc)
    lib.h2d(d_t, table)
    lib.msm_dev(d_s, d_t, n)
    best = 1e9
    for _ in range(3):
        t0 = time.time()
        lib.msm_dev(d_s, d_t, n)
        best = min(best, (time.time() - t0) * 1e3)
    res["msm_dev_2^%d_wall_ms" % log_n] = best
    t0 = time.time()
    lib.msm(sc, table)
    res["msm_host_unregistered_2^%d_wall_ms" % log_n] = (time.time() - t0) * 1e3
    lib.dev_free(d_s)
    lib.dev_free(d_t)
print(json.dumps(res, indent=1))
