python -m pytest tests -m gpu -x -q > gpurun_out/r2H_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2H_pytest.log
python bench.py > gpurun_out/r2H_bench1.json 2> gpurun_out/r2H_bench1.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2H_bench1.json
ncu --set full --clock-control none --import-source on -k regex:'ntt_pass|msm_accumulate' -s 5 -c 5 -o /tmp/r2H_full -f python tools/ncu_target.py > gpurun_out/r2H_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/r2H_full.ncu-rep --page raw --csv > gpurun_out/r2H_full_raw.csv 2>/dev/null
python tools/ncu_extract.py /tmp/r2H_full.ncu-rep --by-grid > gpurun_out/r2H_ncu_full.md 2>/dev/null
python tools/ncu_metrics_json.py /tmp/r2H_full.ncu-rep "ncu --set full of tools/ncu_target.py, final r02 kernels (256-bit pass loads, mul2 in the mixed addition), one B200" > gpurun_out/r2H_ncu_kernel_metrics.json 2>gpurun_out/r2H_metrics.err
python bench.py --steps 2 --warmup 3 --steps-only > gpurun_out/r2H_bench_steps.json 2>/dev/null; echo "steps rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2H_launches.csv python bench.py --steps 2 --warmup 3 --steps-only > gpurun_out/r2H_ncu_launch.log 2>&1; echo "ncu list rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2H_ref1.json 2> gpurun_out/r2H_ref1.err; echo "ref rc=$?"
N=1048576; ./build/make_srs $N build/srs/transcript.dat >/dev/null 2>&1
export OMP_NUM_THREADS=16
BBG_SRS_PRECOMPUTE=1 BBG_PLONK_TRACE=1 ./build/prover_gpu 20 6 > gpurun_out/r2H_prove.json 2> gpurun_out/r2H_prove.err; echo "prove rc=$?"; cut -c1-300 gpurun_out/r2H_prove.json
ls -la gpurun_out/ | grep r2H
