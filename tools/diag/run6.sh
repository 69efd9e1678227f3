set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_l30.py -m gpu -x -q 2>&1 | tail -3
for st in 8 12 16 24 32; do
  HEVM_STREAMS=$st timeout 300 python bench.py --no-cpu-baseline --no-resnet-mix --no-op-table --steps 5 > gpurun_out/st_$st.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/st_$st.json'));print('streams $st bench value',round(d['value']),'ms/step',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value']))"
done
for b in 16 32 48 64; do
  timeout 300 python bench.py --no-cpu-baseline --no-resnet-mix --no-op-table --steps 5 --batch $b > gpurun_out/bt_$b.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/bt_$b.json'));print('batch $b bench value',round(d['value']),'ms/step',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value']))"
done
echo "== resnet"; timeout 300 python tools/launch_count.py 2>&1 | tail -1
