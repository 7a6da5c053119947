"""Development probe: one blocking host-buffer transform (the reference's signature) on a pageable buffer that the
registration cache has page-locked, per call, for fft 2^20 and coset_fft 2^22 (the two shapes of the bench step).
BBG_NTT_HOST_BLOCKS=0 switches the block-pipelined copies off, BBG_NTT_HOST_BLOCK_KB sets the width of a block's row pieces."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import barretenberg_b200 as bb  # noqa: E402

lib = bb.Library()
lib.set_host_register_cache(True)
res = {"host_blocks": os.environ.get("BBG_NTT_HOST_BLOCKS", "1"), "block_kb": os.environ.get("BBG_NTT_HOST_BLOCK_KB", "default")}
for op, log_n in (("fft", 20), ("coset_fft", 21), ("coset_fft", 22)):
    n = 1 << log_n
    raw = np.zeros(n * 4 + 512 + 8, dtype=np.uint64)
    off = (-(raw.ctypes.data // 8) % 512) + 4  # 32 bytes past a page boundary, like an aligned_alloc(32, ..) block
    buf = raw[off:off + n * 4].reshape(n, 4)
    buf[:] = np.random.default_rng(log_n).integers(0, 1 << 60, size=(n, 4), dtype=np.uint64)
    for _ in range(5):
        lib.ntt(op, buf)
    best, tot = 1e9, 0.0
    reps = 10
    for _ in range(reps):
        t = time.perf_counter()
        lib.ntt(op, buf)
        dt = (time.perf_counter() - t) * 1e3
        best = min(best, dt)
        tot += dt
    res["%s_2p%d_ms_best" % (op, log_n)] = round(best, 3)
    res["%s_2p%d_ms_mean" % (op, log_n)] = round(tot / reps, 3)
    lib.host_buffer_forget(buf)
res["registrations"] = lib.host_register_stats()["registrations"]
print(json.dumps(res), flush=True)

# the bench step's shape: 8 buffers per size, each visited once per phase (fft, ifft at 2^20; coset_fft at 2^22)
import ctypes as C  # noqa: E402

libc = C.CDLL(None)
libc.aligned_alloc.restype = C.c_void_p
libc.aligned_alloc.argtypes = [C.c_size_t, C.c_size_t]


def pageable(shape):
    nbytes = int(np.prod(shape)) * 8
    p = libc.aligned_alloc(64, (nbytes + 63) // 64 * 64)
    return np.ctypeslib.as_array((C.c_uint64 * (nbytes // 8)).from_address(p)).reshape(shape)


P = 8
small = [pageable((1 << 20, 4)) for _ in range(P)]
big = [pageable((1 << 22, 4)) for _ in range(P)]
for b in small + big:
    b[:] = 12345
phases = {}
for it in range(8):
    t0 = time.perf_counter()
    for b in small:
        lib.ntt("fft", b)
    t1 = time.perf_counter()
    for b in small:
        lib.ntt("ifft", b)
    t2 = time.perf_counter()
    for b in big:
        lib.ntt("coset_fft", b)
    t3 = time.perf_counter()
    phases = {"fft_x8_ms": round((t1 - t0) * 1e3, 2), "ifft_x8_ms": round((t2 - t1) * 1e3, 2), "coset_x8_ms": round((t3 - t2) * 1e3, 2)}
    if it in (0, 3, 7):
        print(json.dumps({"step_iteration": it, **phases, "registrations": lib.host_register_stats()["registrations"]}), flush=True)
