"""Development aid: time the device-resident MSM of the product library (and of variant builds under build/variants/) for
several accumulate slice lengths (BBG_MSM_SLICE override; 0 = the planner's own choice) and sizes.  One process per run.
usage: python tools/msm_variants.py [--slices 0,56,64,74] [--logs 17,20] [lib.so ...]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(path, log_n=20):
    import numpy as np

    import barretenberg_b200 as bb
    from barretenberg_b200 import synthetic as S

    lib = bb.Library(path)
    n = 1 << log_n
    d_pts = lib.dev_alloc(n * 64)
    d_tab = lib.dev_alloc(n * 128)
    lib.generate_multiples_dev(S.to_limbs(S.mont(12345)), S.to_limbs(S.mont(777)), d_pts, n)
    lib.generate_pippenger_point_table_dev(d_pts, d_tab, n)
    sc = S.random_field(5, n)
    d_sc = lib.dev_alloc(n * 32)
    lib.h2d(d_sc, sc)
    ref = lib.msm_dev(d_sc, d_tab, n)
    lib.profile_enable(True)
    best = 1e9
    import time
    for _ in range(12):
        lib.sync()
        t = time.perf_counter()
        out = lib.msm_dev(d_sc, d_tab, n)
        best = min(best, (time.perf_counter() - t) * 1e3)
        assert (out == ref).all()
    prof = lib.profile_read()
    print(json.dumps({"lib": os.path.basename(path), "log_n": log_n, "slice": os.environ.get("BBG_MSM_SLICE", "auto"), "msm_ms_best": round(best, 4),
                      "accumulate_ms": round(prof["msm_accumulate"][0] / prof["msm_accumulate"][1], 4),
                      "fixup_ms": round(prof["msm_fixup"][0] / prof["msm_fixup"][1], 4), "x_limb0": int(ref[0])}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) == 4 and sys.argv[1] == "--one":
        one(sys.argv[2], int(sys.argv[3]))
    else:
        args = sys.argv[1:]
        slices, logs = ["0"], ["20"]
        if "--slices" in args:
            i = args.index("--slices")
            slices = args[i + 1].split(",")
            del args[i:i + 2]
        if "--logs" in args:
            i = args.index("--logs")
            logs = args[i + 1].split(",")
            del args[i:i + 2]
        vdir = os.path.join(ROOT, "build", "variants")
        libs = args or [os.path.join(ROOT, "barretenberg_b200", "libbbgpu.so")] + sorted(
            os.path.join(vdir, f) for f in (os.listdir(vdir) if os.path.isdir(vdir) else []) if f.endswith(".so"))
        for p in libs:
            for lg in logs:
                for sl in slices:
                    env = dict(os.environ)
                    env.pop("BBG_MSM_SLICE", None)
                    if sl != "0":
                        env["BBG_MSM_SLICE"] = sl
                    subprocess.run([sys.executable, os.path.abspath(__file__), "--one", p, lg], timeout=300, env=env)
