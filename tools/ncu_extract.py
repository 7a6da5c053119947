"""Extract the raw-page metrics worth reading from an `ncu --set full` report (first launch of every kernel name, or of every
(kernel, grid size) pair with --by-grid).  usage: python tools/ncu_extract.py report.ncu-rep [--by-grid] > summary.md"""
import csv
import re
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem"]

rep = sys.argv[1]
by_grid = "--by-grid" in sys.argv
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
gi = hdr.index("launch__grid_size") if "launch__grid_size" in hdr else None
idx = {w: hdr.index(w) for w in WANT if w in hdr}
seen = set()
print("# `ncu --set full --clock-control none` raw-page extract of %s\n" % rep)
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
    key = (name, r[gi]) if (by_grid and gi is not None) else name
    if key in seen:
        continue
    seen.add(key)
    print("\n## `%s`%s\n\n| metric | value |\n|---|---|" % (name, (" (grid %s)" % r[gi]) if by_grid and gi is not None else ""))
    for w, i in idx.items():
        print("| %s | %s %s |" % (w, r[i], units[i]))
