"""Development aid: device-resident MSM over a registered table with and without the fixed-base windows
(bbg_set_srs_precompute), for several forced window widths (BBG_MSM_FIXED_WINDOW; 0 = the planner's choice).
usage: python tools/msm_fixed_base.py [--logs 20] [--windows 0,13,15,16,17,18,19,20] [--pair-rounds 0,1,2,3]   (one JSON line per case;
--pair-rounds forces the number of pair-sum rounds ahead of the accumulate pass, BBG_MSM_PAIR_ROUNDS, default: the planner's)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np

    import barretenberg_b200 as bb
    from barretenberg_b200 import synthetic as S

    args = sys.argv[1:]
    logs, windows = [20], [0]
    if "--logs" in args:
        logs = [int(v) for v in args[args.index("--logs") + 1].split(",")]
    if "--windows" in args:
        windows = [int(v) for v in args[args.index("--windows") + 1].split(",")]
    pair_rounds = [None]
    if "--pair-rounds" in args:
        pair_rounds = [int(v) for v in args[args.index("--pair-rounds") + 1].split(",")]
    lib = bb.Library()
    for log_n in logs:
        n = 1 << log_n
        d_pts = lib.dev_alloc(n * 64)
        d_tab = lib.dev_alloc(n * 128)
        lib.generate_multiples_dev(S.to_limbs(S.mont(12345)), S.to_limbs(S.mont(777)), d_pts, n)
        lib.generate_pippenger_point_table_dev(d_pts, d_tab, n)
        h_tab = np.zeros((2 * n, 8), dtype=np.uint64)
        lib.d2h(h_tab, d_tab)
        sc = S.random_field(5, n)
        if "--constant" in args:  # every scalar the same: one giant bucket per window (the prover's repetitive polynomials)
            sc = np.ascontiguousarray(np.tile(sc[0], (n, 1)))
        d_sc = lib.dev_alloc(n * 32)
        lib.h2d(d_sc, sc)
        os.environ["BBG_MSM_PAIR_ROUNDS"] = "0"
        ref = lib.msm_dev(d_sc, d_tab, n)
        os.environ.pop("BBG_MSM_PAIR_ROUNDS", None)

        def timed(d_table, label, extra):
            for r in pair_rounds:
                if r is None:
                    os.environ.pop("BBG_MSM_PAIR_ROUNDS", None)
                else:
                    os.environ["BBG_MSM_PAIR_ROUNDS"] = str(r)
                    extra = dict(extra, pair_rounds=r)
                timed_one(d_table, label, extra)
            os.environ.pop("BBG_MSM_PAIR_ROUNDS", None)

        def timed_one(d_table, label, extra):
            out = lib.msm_dev(d_sc, d_table, n)
            ok = bool((out == ref).all())
            lib.profile_enable(True)
            best = 1e9
            for _ in range(10):
                lib.sync()
                t = time.perf_counter()
                lib.msm_dev(d_sc, d_table, n)
                best = min(best, (time.perf_counter() - t) * 1e3)
            prof = lib.profile_read()
            lib.profile_enable(False)
            line = {"log_n": log_n, "form": label, "msm_ms_best": round(best, 4), "same_point": ok}
            line.update(extra)
            line["kernels_ms"] = {k: round(v[0] / max(v[1], 1), 4) for k, v in prof.items()}
            print(json.dumps(line), flush=True)

        timed(d_tab, "plain", {})
        for c in windows:
            if c:
                os.environ["BBG_MSM_FIXED_WINDOW"] = str(c)
            else:
                os.environ.pop("BBG_MSM_FIXED_WINDOW", None)
            lib.set_srs_precompute(True)
            t = time.perf_counter()
            keep = lib.srs_register(h_tab)
            build_ms = (time.perf_counter() - t) * 1e3
            d_ptr, cc, w = lib.srs_device_table(keep)
            if w:
                timed(d_ptr, "fixed_base", {"c": cc, "windows": w, "register_ms": round(build_ms, 1), "table_mib": w * n * 128 >> 20})
            else:
                print(json.dumps({"log_n": log_n, "form": "fixed_base", "c": c, "skipped": "no windows built"}), flush=True)
            lib.srs_unregister(keep)
            lib.set_srs_precompute(False)
        for p in (d_pts, d_tab, d_sc):
            lib.dev_free(p)


if __name__ == "__main__":
    main()
