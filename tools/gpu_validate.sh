#!/bin/bash
# Full single-GPU validation of a build, as run at the end of round 2 (under gpurun: `bash tools/gpu_validate.sh TAG`).
# Everything it writes goes to gpurun_out/TAG_*; the ncu report itself stays in /tmp (it is larger than what gpurun copies back).
T=${1:-val}
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${T}_pytest.log
python bench.py > gpurun_out/${T}_bench1.json 2> gpurun_out/${T}_bench1.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/${T}_bench1.json
ncu --set full --clock-control none --import-source on -k regex:'ntt_pass|msm_accumulate' -s 5 -c 5 -o /tmp/${T}_full -f python tools/ncu_target.py > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/${T}_full.ncu-rep --page raw --csv > gpurun_out/${T}_full_raw.csv 2>/dev/null
python tools/ncu_extract.py /tmp/${T}_full.ncu-rep --by-grid > gpurun_out/${T}_ncu_full.md 2>/dev/null
python tools/ncu_metrics_json.py /tmp/${T}_full.ncu-rep "ncu --set full of tools/ncu_target.py, one B200" > gpurun_out/${T}_ncu_kernel_metrics.json 2>/dev/null
python bench.py --steps 2 --warmup 3 --steps-only > gpurun_out/${T}_bench_steps.json 2>/dev/null; echo "steps rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --steps-only > gpurun_out/${T}_ncu_launch.log 2>&1; echo "ncu list rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_ref1.json 2> gpurun_out/${T}_ref1.err; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
