export OMP_NUM_THREADS=16
./build/make_srs 1048576 build/srs/transcript.dat >/dev/null 2>&1
BBG_PLONK_TRACE=1 ./build/prover_gpu 20 4 > /dev/null 2> gpurun_out/r2l_trace.txt; echo rc=$?
BBG_SHIM_STATS=1 ./build/prover_gpu 20 4 > gpurun_out/r2l_prove.json 2> gpurun_out/r2l_stats.txt; echo rc=$?
cut -c1-330 gpurun_out/r2l_prove.json
tail -60 gpurun_out/r2l_trace.txt
cat gpurun_out/r2l_stats.txt | tail -70
