set -x
timeout 400 python tools/arms_sweep.py gpurun_out/r2_arms_sweep2.json 2>&1 | tail -14
timeout 300 python -m pytest tests/test_resnet_gpu.py -m gpu -x -q -k "restated" 2>&1 | tail -2
