"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel next to bench.py's CUDA-event shares.
usage: python tools/summarize_launches.py profiles/r01_ncu_launches.csv profiles/r01_bench_b200_1gpu.json > profiles/r01_ncu_launches_summary.md"""
import collections
import csv
import json
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot, cnt = {}, collections.Counter()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
    tot[name] = tot.get(name, 0.0) + float(r[vi].replace(",", "")) / 1e6
    cnt[name] += 1
# one-off set-up kernels (tables, synthetic points, fixed-base windows) are listed but kept out of the shares: the step's
# CUDA-event attribution has no counterpart for them
SETUP = ("msm_precompute", "g1_progression", "gen_matrix", "gen_powers", "gen_subtw", "endo_table", "precompute_plain", "srs_from_transcript")
setup_ms = sum(v for k, v in tot.items() if any(x in k for x in SETUP))
total = sum(tot.values()) - setup_ms
print("# r01 ncu launch list summary (bench.py --steps 2 --warmup 3 --steps-only, 1x B200, final kernels of the round)\n")
print("Source: %s (`ncu --metrics gpu__time_duration.sum --clock-control none -c 400`). Per-launch times under ncu are" % sys.argv[1])
print("cold-cache and serialised: compare SHARES with bench.py's CUDA-event `kernels` table, not absolutes. The capture covers set-up")
print("(table generation, parity gate) + untimed steps + 2 timed steps: every step launches the same kernels.\n")
print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    if any(x in k for x in SETUP):
        print("| `%s` (set-up, once) | %d | %.3f | — |" % (k, cnt[k], v))
    else:
        print("| `%s` | %d | %.3f | %.1f%% |" % (k, cnt[k], v, 100 * v / total))
ntt = sum(v for k, v in tot.items() if "ntt_pass" in k)
acc = sum(v for k, v in tot.items() if "msm_accumulate" in k)
print("\nNTT passes %.1f%%, `msm_accumulate` %.1f%% of the profiled kernel time.\n" % (100 * ntt / total, 100 * acc / total))
if len(sys.argv) > 2:
    b = json.load(open(sys.argv[2]))
    print("bench.py CUDA-event shares of the same step (%s `kernels`):\n" % sys.argv[2])
    for k, v in sorted(b["kernels"].items(), key=lambda x: -x[1]["share"]):
        print("* `%s`: %.1f%% (%.3f ms per launch)" % (k, 100 * v["share"], v["ms_per_launch"]))
    pa = sum(v["share"] for k, v in b["kernels"].items() if k.startswith("ntt_pass"))
    ev_acc = b["kernels"]["msm_accumulate"]["share"]
    print("\nNTT passes %.1f%%, `msm_accumulate` %.1f%% by CUDA events; under ncu %.1f%% / %.1f%% (difference %.1f / %.1f points)."
          % (100 * pa, 100 * ev_acc, 100 * ntt / total, 100 * acc / total, abs(100 * pa - 100 * ntt / total), abs(100 * ev_acc - 100 * acc / total)))
