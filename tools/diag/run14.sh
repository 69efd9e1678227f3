set -x
timeout 600 python tools/emit_profile.py 15 14 gpurun_out/profiled_B200_GPU_tp.json 10 2>&1 | tail -3
python - <<'PY'
import json
d=json.load(open('gpurun_out/profiled_B200_GPU_tp.json'))
for k,v in d['latencyTableThroughput'].items(): print(k,[round(x,1) for x in v])
PY
