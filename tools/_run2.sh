L=barretenberg_b200/libbbgpu.so
O=gpurun_out/r2c_ntt_sweep.jsonl
: > $O
python tools/ntt_ablate.py $L build/variants/lib_coarse_sync.so >> $O 2>gpurun_out/r2c_err.log
for s in 15000 30000 45000 60000 90000; do
  BBG_NTT_STAGGER_A=$s BBG_NTT_STAGGER_B=$s python tools/ntt_ablate.py $L >> $O 2>>gpurun_out/r2c_err.log
done
for s in 20000 60000; do
  BBG_NTT_STAGGER_CTA_B=$s python tools/ntt_ablate.py $L >> $O 2>>gpurun_out/r2c_err.log
  BBG_NTT_STAGGER_A=45000 BBG_NTT_STAGGER_B=45000 BBG_NTT_STAGGER_CTA_B=$s python tools/ntt_ablate.py $L >> $O 2>>gpurun_out/r2c_err.log
done
BBG_NTT_STAGGER_A=45000 BBG_NTT_STAGGER_B=45000 python tools/ntt_ablate.py build/variants/lib_coarse_sync.so >> $O 2>>gpurun_out/r2c_err.log
cat $O
tail -5 gpurun_out/r2c_err.log
