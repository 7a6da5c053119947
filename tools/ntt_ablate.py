"""Development aid: time the NTT pass kernels of the product library and of the ablation builds under build/ablate
(bbg_ntt.cu compiled with -DBBG_NTT_ABLATE=N: wrong results by construction, timing only).  One process per library
(the C ABI keeps global state).  usage: python tools/ntt_ablate.py [lib.so ...]   -> one JSON line per library"""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(path):
    import numpy as np

    import barretenberg_b200 as bb

    lib = bb.Library(path)
    out = {"lib": os.path.basename(path), "env": {k: v for k, v in os.environ.items() if k.startswith("BBG_NTT_")}}
    # ifft(fft(x)) == x through this library (the ablation builds fail this by construction)
    x0 = np.random.default_rng(7).integers(0, 1 << 60, size=(2 << 20, 4), dtype=np.uint64)
    d0 = lib.dev_alloc(x0.nbytes)
    lib.h2d(d0, x0)
    lib.ntt_dev("coset_fft", d0, 20, batch=2)
    lib.ntt_dev("coset_ifft", d0, 20, batch=2)
    y0 = np.zeros_like(x0)
    lib.d2h(y0, d0)
    lib.dev_free(d0)
    out["round_trip_ok"] = bool((x0 == y0).all())
    for log_n, op in ((20, "fft"), (21, "coset_fft"), (22, "coset_fft")):
        n, batch = 1 << log_n, 8
        x = np.random.default_rng(1).integers(0, 1 << 60, size=(batch * n, 4), dtype=np.uint64)
        d = lib.dev_alloc(x.nbytes)
        lib.h2d(d, x)
        lib.profile_enable(True)
        for _ in range(3):
            lib.check(lib.lib.bbg_ntt_fr_dev(C.c_void_p(d), n, batch, log_n, lib.OPS[op], None))
        lib.sync()
        lib.profile_enable(True)  # reset
        reps = 10
        lib.timer_start()
        for _ in range(reps):
            lib.check(lib.lib.bbg_ntt_fr_dev(C.c_void_p(d), n, batch, log_n, lib.OPS[op], None))
        ms = lib.timer_stop() / reps
        prof = lib.profile_read()
        out["%s_2p%d_batch8_ms" % (op, log_n)] = round(ms, 4)
        for k in ("ntt_pass_a", "ntt_pass_b"):
            if k in prof:
                out["%s_2p%d_%s_ms" % (op, log_n, k)] = round(prof[k][0] / prof[k][1], 4)
        lib.dev_free(d)
    # single polynomial (what a prover's one-at-a-time calls and a rank with one polynomial see)
    for log_n, op in ((20, "fft"), (22, "coset_fft"), (12, "fft")):
        n = 1 << log_n
        x = np.random.default_rng(2).integers(0, 1 << 60, size=(n, 4), dtype=np.uint64)
        d = lib.dev_alloc(x.nbytes)
        lib.h2d(d, x)
        for _ in range(3):
            lib.ntt_dev(op, d, log_n, batch=1)
        lib.sync()
        reps = 20
        lib.timer_start()
        for _ in range(reps):
            lib.ntt_dev(op, d, log_n, batch=1)
        out["%s_2p%d_batch1_ms" % (op, log_n)] = round(lib.timer_stop() / reps, 4)
        lib.dev_free(d)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--one":
        one(sys.argv[2])
    else:
        found = []
        for sub in ("ablate", "variants"):
            d = os.path.join(ROOT, "build", sub)
            if os.path.isdir(d):
                found += sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(".so"))
        libs = sys.argv[1:] or [os.path.join(ROOT, "barretenberg_b200", "libbbgpu.so")] + found
        for p in libs:
            subprocess.run([sys.executable, os.path.abspath(__file__), "--one", p], timeout=300)
