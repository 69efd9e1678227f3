set -x
timeout 300 python examples/resnet20.py b200c 40 B200 GPU 2>&1 | grep -E "first run|latency|rms"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c_abi or async or replay or chain or fused or scheduler" 2>&1 | tail -2
