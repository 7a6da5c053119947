python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2j_pytest.log
python bench.py > gpurun_out/r2j_bench1.json 2> gpurun_out/r2j_bench1.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r2j_bench1.json
python tools/ncu_target.py > gpurun_out/r2j_target.log 2>&1; echo "target rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'ntt_pass|msm_accumulate' -s 5 -c 5 -o gpurun_out/r2j_full -f python tools/ncu_target.py > gpurun_out/r2j_ncu_full.log 2>&1; echo "ncu full rc=$?"
python bench.py --steps 2 --warmup 3 --device-only > gpurun_out/r2j_bench_dev.json 2>/dev/null; echo "dev rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2j_launches.csv python bench.py --steps 2 --warmup 3 --device-only > gpurun_out/r2j_ncu_launch.log 2>&1; echo "ncu list rc=$?"
ls -la gpurun_out/ | tail -8
