set -x
for m in 0 2; do
HEVM_P2P_OVERLAP=$m timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2951$m tools/shard_probe.py 16 30 > gpurun_out/shard8_ov$m.json 2> gpurun_out/shard8_ov$m.err
echo "exit $?"
python - <<PY
import json
for line in open('gpurun_out/shard8_ov$m.json'):
    if line.startswith('{'):
        d=json.loads(line)
        for lv,v in d['levels'].items():
            print('overlap $m level',lv,'rotate',v['rotate']['sharded_us'],v['rotate']['speedup'],'mulcc',v['mulcc']['sharded_us'],v['mulcc']['speedup'],v['bit_exact_vs_single_gpu'],v['rotate'].get('stages_us_rank0'))
PY
done
