#!/bin/bash
# configs[4]: reference waffle prover — all-CPU  vs  MSM+NTT on the GPU through the ten-entry-point shims ("classic")
#             vs  the HBM-resident Prover::construct_proof ("resident").  usage: prover_compare.sh <log2_gates> [repeat]
set -e
cd "$(dirname "$0")/.."
LG=${1:-16}
REP=${2:-3}
N=$((1 << LG))
mkdir -p build/srs gpurun_out
need=$((64 * (N - 1) + 28 + 256 + 64))
have=$(stat -c %s build/srs/transcript.dat 2>/dev/null || echo 0)
if [ "$have" -lt "$need" ]; then ./build/make_srs $N build/srs/transcript.dat 1>&2; fi
export OMP_NUM_THREADS=$(python3 -c "import os;c=os.cpu_count();p=1
while p*2<=c:p*=2
print(p)")
./build/prover_cpu $LG $REP > gpurun_out/prover_cpu_$LG.json
./build/prover_gpu_classic $LG $REP > gpurun_out/prover_gpu_classic_$LG.json
./build/prover_gpu $LG $REP > gpurun_out/prover_gpu_$LG.json
# one more resident run with the per-round / per-kernel stopwatch on (not used for the headline time)
BBG_SHIM_STATS=1 ./build/prover_gpu $LG $REP > /dev/null 2> gpurun_out/prover_gpu_${LG}_stats.txt || true
python3 - <<PY
import json
c=json.load(open("gpurun_out/prover_cpu_$LG.json")); g=json.load(open("gpurun_out/prover_gpu_$LG.json")); k=json.load(open("gpurun_out/prover_gpu_classic_$LG.json"))
same=all(c["proof"][f]==g["proof"][f] for f in c["proof"]); same_k=all(c["proof"][f]==k["proof"][f] for f in c["proof"])
print(json.dumps({"log2_gates":$LG,"omp_threads":int("$OMP_NUM_THREADS"),"cpu_prove_ms":c["prove_ms_best"],
 "gpu_resident_prove_ms_first":g["prove_ms_first"],"gpu_resident_prove_ms_best":g["prove_ms_best"],
 "gpu_classic_prove_ms_first":k["prove_ms_first"],"gpu_classic_prove_ms_best":k["prove_ms_best"],
 "cpu_verified":c["verified"],"gpu_resident_verified":g["verified"],"gpu_classic_verified":k["verified"],
 "resident_proof_identical":same,"classic_proof_identical":same_k,
 "cpu_setup_ms":c["setup_ms"],"gpu_setup_ms":g["setup_ms"],"cpu_vk_ms":c["verifier_key_ms"],"gpu_vk_ms":g["verifier_key_ms"]}))
PY
