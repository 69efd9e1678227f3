"""Limb-sharded rotation across GPUs: bit-exactness against the single-GPU rotate, and timing.
launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/shard_probe.py [logN] [L]"""
import json, os, sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import torch, torch.distributed as dist
from dacapo_b200 import _binding, sharded

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 15
L = int(sys.argv[2]) if len(sys.argv) > 2 else 14
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ["HEVM_DEVICE"] = str(local)
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
out = sharded.measure(_binding.bind(_binding.B200_LIB), rank, world, logn, L)
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
