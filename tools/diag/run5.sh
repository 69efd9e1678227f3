set -x
mkdir -p gpurun_out
for g in 1184 800 592 444 296 148; do
  echo "== HEVM_GROUP_WARPS=$g resnet"; HEVM_GROUP_WARPS=$g timeout 300 python tools/launch_count.py 2>&1 | tail -1
  HEVM_GROUP_WARPS=$g timeout 300 python bench.py --no-cpu-baseline --no-resnet-mix --no-op-table --steps 5 > gpurun_out/gw_$g.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/gw_$g.json'));print('bench value',round(d['value']),'ms/step',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value']))"
done
