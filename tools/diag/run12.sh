set -x
HEVM_P2P_OVERLAP=2 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "p2p_sharded" 2>&1 | tail -2
