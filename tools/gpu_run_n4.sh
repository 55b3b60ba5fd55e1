# 4-GPU repeatability check of the C2 step: peer-memory exchange vs NCCL, alternating, three times each
mkdir -p gpurun_out
B="--extras off --no-cpu-baseline --e2e-steps 0 --no-parity"
for rep in 1 2 3; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2971$rep bench.py --gpus 4 --steps 40 $B > gpurun_out/r02_n4_peer_$rep.json 2> gpurun_out/r02_n4_peer_$rep.err
  AFL_XGPU=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2972$rep bench.py --gpus 4 --steps 40 $B > gpurun_out/r02_n4_nccl_$rep.json 2> gpurun_out/r02_n4_nccl_$rep.err
done
for f in gpurun_out/r02_n4_*.json; do python -c "
import json
try:
    l=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(l['value'],1), round(l['ms_per_step']*1e3,1), l['breakdown_us'])
except Exception as e: print('$f ERR', e)
"; done
