#!/bin/bash
# Development aid: build/variants/lib_<name>.so = the product library with ONE translation unit recompiled with extra
# -D flags (timing experiments: tools/ntt_ablate.py, tools/msm_variants.py pick the libraries up from build/variants/).
# usage: tools/build_variant.sh <name> <unit: ntt|msm|plonk|capi> <-Dflags...>
set -e
cd "$(dirname "$0")/.."
NAME=$1; UNIT=$2; shift 2
C=barretenberg_b200/csrc
mkdir -p build/variants/obj
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --extended-lambda "$@" \
  -ccbin /usr/bin/g++ -Xcompiler -fPIC,-fvisibility=default -c -o build/variants/obj/${NAME}_${UNIT}.o $C/bbg_${UNIT}.cu
OBJS=""
for u in capi ntt msm plonk microbench selftest; do
  if [ "$u" = "$UNIT" ]; then OBJS="$OBJS build/variants/obj/${NAME}_${UNIT}.o"; else OBJS="$OBJS $C/build/bbg_$u.o"; fi
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o build/variants/lib_${NAME}.so $OBJS
echo "built build/variants/lib_${NAME}.so"
