O=gpurun_out/r2M_msm_sizes.jsonl
: > $O
python tools/msm_fixed_base.py --logs 17,18,19,20 --windows 0 2>>gpurun_out/r2M_err.log | cut -c1-600 >> $O
python - <<'PY'
import json
for l in open('gpurun_out/r2M_msm_sizes.jsonl'):
    d=json.loads(l)
    k=d['kernels_ms']; print(' ',d['log_n'],d['form'],d.get('c'),d['msm_ms_best'],d['same_point'],'acc',k['msm_accumulate'],'chunk',k['msm_chunk'],'reduce',k['msm_reduce'])
PY
python -m pytest tests/test_gpu_parity.py tests/test_shim_msm.py tests/test_gpu_prover_dropin.py -m gpu -x -q -k "msm or shim or prover" 2>&1 | tail -3
tail -2 gpurun_out/r2M_err.log
