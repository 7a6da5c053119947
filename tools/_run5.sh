python -m pytest tests -m gpu -x -q > gpurun_out/r2N_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2N_pytest.log
python bench.py > gpurun_out/r2N_bench1.json 2> gpurun_out/r2N_bench1.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2N_bench1.json
python bench.py --steps 2 --warmup 3 --steps-only > gpurun_out/r2N_bench_steps.json 2>/dev/null; echo "steps rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2N_launches.csv python bench.py --steps 2 --warmup 3 --steps-only > gpurun_out/r2N_ncu_launch.log 2>&1; echo "ncu list rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
