python bench.py > gpurun_out/r2C_bench1.json 2> gpurun_out/r2C_bench1.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2C_bench1.json
ncu --set full --clock-control none --import-source on -k regex:'ntt_pass|msm_accumulate' -s 5 -c 5 -o /tmp/r2C_full -f python tools/ncu_target.py > gpurun_out/r2C_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/r2C_full.ncu-rep --page raw --csv > gpurun_out/r2C_full_raw.csv 2>/dev/null
python tools/ncu_extract.py /tmp/r2C_full.ncu-rep --by-grid > gpurun_out/r2C_ncu_full.md 2>/dev/null
python tools/ncu_metrics_json.py /tmp/r2C_full.ncu-rep "ncu --set full of tools/ncu_target.py, final r02 kernels (256-bit pass loads), one B200" > gpurun_out/r2C_ncu_kernel_metrics.json 2>gpurun_out/r2C_metrics.err
ncu -i /tmp/r2C_full.ncu-rep --page source --csv -k regex:ntt_pass_kernel -c 1 > gpurun_out/r2C_source_pass.csv 2>/dev/null; gzip -f gpurun_out/r2C_source_pass.csv
python bench.py --steps 2 --warmup 3 --device-only > gpurun_out/r2C_bench_dev.json 2>/dev/null; echo "dev rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2C_launches.csv python bench.py --steps 2 --warmup 3 --device-only > gpurun_out/r2C_ncu_launch.log 2>&1; echo "ncu list rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2C_ref1.json 2> gpurun_out/r2C_ref1.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2C_ref1.json
ls -la gpurun_out/ | tail -12
