set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2b_pytest.log
python tools/msm_fixed_base.py --logs 20 --windows 0,13,15,16,17,18,19,20 > gpurun_out/r2b_fixed_base_2p20.jsonl 2> gpurun_out/r2b_fixed_base.err; echo "fb rc=$?"
python tools/msm_fixed_base.py --logs 17,18 --windows 0 >> gpurun_out/r2b_fixed_base_2p20.jsonl 2>> gpurun_out/r2b_fixed_base.err
cat gpurun_out/r2b_fixed_base_2p20.jsonl | cut -c1-400
N=1048576; ./build/make_srs $N build/srs/transcript.dat >/dev/null 2>&1
export OMP_NUM_THREADS=16
for pre in 1 0; do BBG_SRS_PRECOMPUTE=$pre ./build/prover_gpu 20 6 > gpurun_out/r2b_prover_gpu_20_pre$pre.json 2> gpurun_out/r2b_prover_gpu_20_pre$pre.err; echo "prove pre=$pre rc=$?"; cut -c1-330 gpurun_out/r2b_prover_gpu_20_pre$pre.json; done
BBG_SRS_PRECOMPUTE=1 ./build/prover_gpu_classic 20 3 > gpurun_out/r2b_prover_classic_20.json 2> gpurun_out/r2b_prover_classic_20.err; echo "classic rc=$?"; cut -c1-330 gpurun_out/r2b_prover_classic_20.json
