O=gpurun_out/r2K_msm_const.jsonl
: > $O
python tools/msm_fixed_base.py --logs 20 --windows 0 2>>gpurun_out/r2K_err.log | cut -c1-600 >> $O
python tools/msm_fixed_base.py --logs 20 --windows 0 --constant 2>>gpurun_out/r2K_err.log | cut -c1-600 >> $O
cat $O
python -m pytest tests/test_gpu_parity.py tests/test_shim_msm.py -m gpu -x -q -k "msm or shim" 2>&1 | tail -3
tail -2 gpurun_out/r2K_err.log
