# 8-GPU run: sharded == unsharded check, C2 strong scaling 1/2/4/8 with the peer-memory exchange (and NCCL at 4 and 8 for
# comparison), then the default bench at 8 GPUs with the C3/C4/C5 shards in `extra`.
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
B="--extras off --no-cpu-baseline --e2e-steps 0 --no-parity"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29601 tools/multi_gpu_check.py > gpurun_out/r02_h_multigpu_check_n8.txt 2>&1
for n in 8 4 2; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 30 $B > gpurun_out/r02_h_c2_n${n}_peer.json 2> gpurun_out/r02_h_c2_n${n}_peer.err
done
timeout 200 python bench.py --gpus 1 --steps 30 $B > gpurun_out/r02_h_c2_n1.json 2> gpurun_out/r02_h_c2_n1.err
for n in 8 4; do
  AFL_XGPU=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2962$n bench.py --gpus $n --steps 30 $B > gpurun_out/r02_h_c2_n${n}_nccl.json 2> gpurun_out/r02_h_c2_n${n}_nccl.err
done
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 8 --steps 20 --extra-steps 3 --e2e-steps 1 > gpurun_out/r02_h_bench_n8_extras.json 2> gpurun_out/r02_h_bench_n8_extras.err
grep MULTIGPU gpurun_out/r02_h_multigpu_check_n8.txt | tail -1 | cut -c1-400
for f in gpurun_out/r02_h_c2_n*.json; do echo $f; python -c "
import json,sys
try:
    l=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(l['value'],1), round(l['ms_per_step'],4), l['breakdown_us'], l['config']['parallelism'][:60])
except Exception as e: print('ERR', e)
"; done
