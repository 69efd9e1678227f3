set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_runner_gpu.py tests/test_benchmarks.py -m gpu -x -q -k "c_abi or async or replay or bootstrap or runner or bench or encrypt" 2>&1 | tail -5
timeout 600 python bench.py --no-cpu-baseline --no-resnet-mix --no-op-table > gpurun_out/r02c_bench_quick.json 2> gpurun_out/r02c_bench_quick.err; tail -3 gpurun_out/r02c_bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02c_bench_quick.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches')})
PY
