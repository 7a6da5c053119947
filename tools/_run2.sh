O=gpurun_out/r2B_ntt_host.jsonl
: > $O
BBG_NTT_HOST_BLOCKS_MIN_LOG=20 python tools/ntt_host_probe.py >> $O 2>>gpurun_out/r2B_err.log
python tools/ntt_host_probe.py >> $O 2>>gpurun_out/r2B_err.log
cat $O
python bench.py --steps 5 --warmup 3 > gpurun_out/r2B_bench1.json 2> gpurun_out/r2B_bench1.err; tail -c 600 gpurun_out/r2B_bench1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2B_bench1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','components_ms')}, d['e2e']['value'], d['e2e_pinned']['value'], d['cpu_baseline']['value'], d['prove']['gpu_prove_ms'], d['roofline']['frac'])
PY
