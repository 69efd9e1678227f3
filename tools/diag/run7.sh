set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_l30.py -m gpu -x -q 2>&1 | tail -3
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_probe.py > gpurun_out/r02c_sanitizer_racecheck.log 2>&1; tail -4 gpurun_out/r02c_sanitizer_racecheck.log
timeout 300 python bench.py --no-cpu-baseline --no-resnet-mix --no-op-table --steps 5 > gpurun_out/pad_bench.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/pad_bench.json'));print('bench value',round(d['value']),'ms/step',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value']), d['roofline']['avg_launch_us'])"
echo "== resnet"; timeout 300 python tools/launch_count.py 2>&1 | tail -1
timeout 200 python tools/ks_time.py 13 4 1 2>&1 | tail -8
HEVM_FUSED=0 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_r02c -f python tools/ks_probe.py 13 > gpurun_out/ncu_r02c.log 2>&1; tail -2 gpurun_out/ncu_r02c.log
