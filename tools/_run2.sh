python tools/msm_variants.py --logs 20,17 2>gpurun_out/r2t_err.log | tee gpurun_out/r2t_msm_wide.jsonl
python tools/msm_fixed_base.py --logs 20 --windows 0 2>>gpurun_out/r2t_err.log | cut -c1-400 | tee -a gpurun_out/r2t_msm_wide.jsonl
python -m pytest tests/test_gpu_parity.py tests/test_shim_msm.py -m gpu -x -q -k "msm or shim" 2>&1 | tail -3
tail -2 gpurun_out/r2t_err.log
