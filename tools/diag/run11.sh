set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err
echo "exit $?"
tail -5 gpurun_out/r02_bench_8gpu.err
wc -l gpurun_out/r02_bench_8gpu.json
