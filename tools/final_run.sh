# Round-end evidence run on one B200: GPU tests, bench (own arm + reference arm), ncu launch list of the bench's timed region.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r02_gpu_tests.log
tail -3 gpurun_out/r02_gpu_tests.log
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; cp bench_details.json gpurun_out/r02_bench_details.json; cp profiled_B200_GPU.json gpurun_out/profiled_B200_GPU_final.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 1 --warmup 3 --ncu-range --no-cpu-baseline --no-resnet-mix --no-op-table > /dev/null 2>&1
head -c 1500 gpurun_out/r02_bench.json
