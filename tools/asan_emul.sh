#!/bin/bash
# TEST INFRASTRUCTURE: the kernels' index logic under AddressSanitizer.  compute-sanitizer is not available on the GPU pool,
# so the CPU kernel-emulation build (tests/emul: every "device" buffer is a heap allocation) is compiled with
# -fsanitize=address and the resident prover is run through it for all four composers, two proofs each (cold and warm
# proving key).  Any out-of-bounds read / write of a kernel, the scans or the drivers aborts with an ASan report.
# usage: tools/asan_emul.sh [log2_gates]      (needs build/prover_gpu_emul: make -C tests/cpp)
set -e
cd "$(dirname "$0")/.."
LG=${1:-6}
OUT=$(mktemp -d)
( cd barretenberg_b200/csrc && /usr/bin/g++ -std=c++17 -O1 -g -fPIC -shared -fsanitize=address -fno-omit-frame-pointer -DBBG_EMULATE \
    -I../../tests/emul -I. -x c++ bbg_capi.cu bbg_ntt.cu bbg_msm.cu bbg_plonk.cu bbg_microbench.cu ../../tests/emul/emul_globals.cpp \
    -o "$OUT/libbbgpu_emul.so" -lpthread )
for k in standard bool mimc extended; do
  ASAN_OPTIONS=detect_leaks=0 LD_PRELOAD=$(gcc -print-file-name=libasan.so) LD_LIBRARY_PATH="$OUT" ./build/prover_gpu_emul "$LG" 2 "$k" > "$OUT/$k.json" 2> "$OUT/$k.err" \
    || { echo "$k: FAILED"; grep -m5 "ERROR\|SUMMARY" "$OUT/$k.err"; exit 1; }
  if grep -q "ERROR: AddressSanitizer" "$OUT/$k.err"; then echo "$k: ASan report"; grep -m5 "ERROR\|SUMMARY" "$OUT/$k.err"; exit 1; fi
  python3 -c "import json,sys;d=json.load(open(sys.argv[1]));assert d['verified'];print(sys.argv[2],'n =',d['n'],'widgets =',d['widgets'],'verified, no ASan report')" "$OUT/$k.json" "$k"
done
rm -rf "$OUT"
