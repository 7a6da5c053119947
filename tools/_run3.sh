O=gpurun_out/r2f_fixed_base.jsonl
python tools/msm_fixed_base.py --logs 20 --windows 0,16 > $O 2> gpurun_out/r2f_err.log
python tools/msm_fixed_base.py --logs 17,18,19 --windows 0 >> $O 2>> gpurun_out/r2f_err.log
cut -c1-420 $O
python -m pytest tests/test_gpu_parity.py tests/test_shim_msm.py -m gpu -x -q -k "msm or shim" 2>&1 | tail -3
tail -3 gpurun_out/r2f_err.log
