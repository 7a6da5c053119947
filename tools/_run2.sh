O=gpurun_out/r2i_ntt_split.jsonl
: > $O
for sp in "" "20:9,21:10" "20:11,21:11" "20:9,21:11"; do
  BBG_NTT_SPLIT=$sp python tools/ntt_ablate.py barretenberg_b200/libbbgpu.so >> $O 2>>gpurun_out/r2i_err.log
done
python - <<'PY'
import json
for l in open('gpurun_out/r2i_ntt_split.jsonl'):
    d=json.loads(l); print(d['env'], d['round_trip_ok'], {k:v for k,v in d.items() if k.endswith('_ms')})
PY
tail -3 gpurun_out/r2i_err.log
