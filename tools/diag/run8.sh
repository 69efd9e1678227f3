set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_gpu_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r02d_gpu_tests.log
grep -v "^$" gpurun_out/r02d_gpu_tests.log | tail -6
timeout 300 python examples/resnet20.py b200c 40 B200 GPU 2>&1 | tail -12
