# ResNet-20 run() latency under scheduler / launch knobs (one line per setting): which resource bounds the 0.12 s?
run() { echo "== $*"; env "$@" timeout 300 python tools/launch_count.py 2>&1 | tail -1; }
run HEVM_STREAMS=16
run HEVM_STREAMS=8
run HEVM_STREAMS=24
run HEVM_STREAMS=32
run HEVM_STREAMS=32 HEVM_RENAME_POOL=256
run HEVM_FUSED=1
run HEVM_PDL=0
run HEVM_GRAPH=0
run HEVM_STREAMS=1
run CUDA_DEVICE_MAX_CONNECTIONS=32
run CUDA_DEVICE_MAX_CONNECTIONS=8
