set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_gpu_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r02b_gpu_tests.log
tail -4 gpurun_out/r02b_gpu_tests.log
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "p2p_sharded or seal_key_directory" 2>&1 | tail -1; done
timeout 300 python tools/resnet_classes.py > gpurun_out/resnet_classes.log 2>&1; cat gpurun_out/resnet_classes.log
echo "== graph PDL"; HEVM_GRAPH_PDL=1 timeout 300 python tools/launch_count.py 2>&1 | tail -2
HEVM_GRAPH_PDL=1 timeout 600 python -m pytest tests/test_resnet_gpu.py -m gpu -q -x 2>&1 | tail -3
