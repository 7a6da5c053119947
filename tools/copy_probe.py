"""Development probe: host<->device copy rates for pageable vs pinned buffers through the C ABI."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import barretenberg_b200 as bb  # noqa: E402

lib = bb.default_library()
n = 1 << 22
nbytes = n * 32
res = {}
page = np.ones((n, 4), dtype=np.uint64)
pinned = torch.ones((n, 4), dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
d = lib.dev_alloc(nbytes)


def rate(fn, reps=3):
    fn()
    best = 1e9
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t)
    return nbytes / best / 1e9


t = time.perf_counter()
tmp = page.copy()
res["numpy_copy_GBps"] = nbytes / (time.perf_counter() - t) / 1e9
res["h2d_pageable_driver_GBps"] = rate(lambda: lib.h2d(d, page))
res["d2h_pageable_driver_GBps"] = rate(lambda: lib.d2h(page, d))
res["h2d_pinned_GBps"] = rate(lambda: lib.h2d(d, pinned))
res["d2h_pinned_GBps"] = rate(lambda: lib.d2h(pinned, d))
for name, buf in (("pageable", page), ("pinned", pinned)):
    buf[:] = 1
    lib.ntt("coset_fft", buf)
    best = 1e9
    for _ in range(3):
        t = time.perf_counter()
        lib.ntt("coset_fft", buf)
        best = min(best, time.perf_counter() - t)
    res["ntt_coset_2p22_host_%s_ms" % name] = best * 1e3
lib.ntt_dev("coset_fft", d, 22)
lib.sync()
t = time.perf_counter()
lib.ntt_dev("coset_fft", d, 22)
lib.sync()
res["ntt_coset_2p22_device_ms"] = (time.perf_counter() - t) * 1e3
print(json.dumps(res, indent=1))
