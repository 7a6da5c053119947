O=gpurun_out/r2G_msm_mul2.jsonl
: > $O
python tools/msm_fixed_base.py --logs 20,17 --windows 0 2>>gpurun_out/r2G_err.log | cut -c1-600 >> $O
cat $O
python -m pytest tests/test_field_selftest.py tests/test_gpu_parity.py tests/test_shim_msm.py -m gpu -x -q -k "selftest or field or msm or shim" 2>&1 | tail -3
tail -2 gpurun_out/r2G_err.log
