"""profiles/ncu_kernel_metrics.json from an `ncu --set full` report of tools/ncu_target.py: per captured launch the duration,
DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) and pipe activity, keyed the way bench.py names kernels.
bench.py copies the entry of its dominant kernel into the `roofline` object of its JSON line (traffic, fmaheavy pipe-active).
usage: python tools/ncu_metrics_json.py report.ncu-rep "<where the capture came from>" > profiles/ncu_kernel_metrics.json"""
import csv
import json
import re
import subprocess
import sys

rep, note = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def num(r, name):
    v = float(r[col[name]].replace(",", ""))
    u = units[col[name]].lower()
    scale = {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u, 1.0)
    return v * scale


out = {"_note": note, "_metrics": "ms = gpu__time_duration.sum; traffic_bytes = dram__bytes_read.sum + dram__bytes_write.sum; "
       "fmaheavy_pct / alu_pct = sm__pipe_{fmaheavy,alu}_cycles_active.avg.pct_of_peak_sustained_elapsed; issue_pct = "
       "smsp__issue_active.avg.pct_of_peak_sustained_active; all per launch, ncu --set full --clock-control none"}
seen = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    m = re.search(r"ntt_pass_kernel<(\d+), *(\d|true|false)>", name)
    if m:
        key = "ntt_pass_%s_L%s" % ("a" if m.group(2) in ("1", "true") else "b", m.group(1))
    elif "msm_accumulate" in name:
        key = "msm_accumulate"
    else:
        key = re.sub(r"\(.*", "", name).split("::")[-1]
    grid = int(float(r[col["launch__grid_size"]].replace(",", "")))
    if key in seen and seen[key] >= grid:
        continue  # keep the largest launch of each kernel (the batch-8 / full-size one)
    seen[key] = grid
    out[key] = {
        "ms": round(num(r, "gpu__time_duration.sum"), 6),
        "traffic_bytes": num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum"),
        "fmaheavy_pct": round(num(r, "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"), 2),
        "alu_pct": round(num(r, "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed"), 2),
        "issue_pct": round(num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), 2),
        "registers": int(num(r, "launch__registers_per_thread")),
        "grid": grid,
    }
print(json.dumps(out, indent=1))
