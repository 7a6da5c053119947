O=gpurun_out/r2x_msm_digits.jsonl
: > $O
python tools/msm_fixed_base.py --logs 20 --windows 0 2>>gpurun_out/r2x_err.log | cut -c1-600 >> $O
python tools/msm_fixed_base.py --logs 20 --windows 0 --constant 2>>gpurun_out/r2x_err.log | cut -c1-600 >> $O
python tools/msm_fixed_base.py --logs 17 --windows 0 2>>gpurun_out/r2x_err.log | cut -c1-600 >> $O
cat $O
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "msm" 2>&1 | tail -3
tail -2 gpurun_out/r2x_err.log
