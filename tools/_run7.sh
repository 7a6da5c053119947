nvidia-smi -L | wc -l
python -m pytest tests -m gpu -x -q -k "nccl or multi_device or all_devices" > gpurun_out/r2m_pytest8.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2m_pytest8.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2m_bench_8gpu.json 2> gpurun_out/r2m_bench_8gpu.err; echo "bench8 rc=$?"; cut -c1-200 gpurun_out/r2m_bench_8gpu.json; tail -3 gpurun_out/r2m_bench_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/r2m_ref_8gpu.json 2> gpurun_out/r2m_ref_8gpu.err; echo "ref8 rc=$?"; cut -c1-200 gpurun_out/r2m_ref_8gpu.json
