# 2-GPU run: peer-memory exchange vs NCCL, sharded == unsharded check
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/multi_gpu_check.py > gpurun_out/r02_f_multigpu_check_n2.txt 2>&1
B="--extras off --no-cpu-baseline --e2e-steps 0"
timeout 300 $TR bench.py --gpus 2 --steps 20 $B > gpurun_out/r02_f_c2_n2_peer.json 2> gpurun_out/r02_f_c2_n2_peer.err
AFL_XGPU=0 timeout 300 $TR bench.py --gpus 2 --steps 20 $B > gpurun_out/r02_f_c2_n2_nccl.json 2> gpurun_out/r02_f_c2_n2_nccl.err
timeout 300 python bench.py --gpus 1 --steps 20 $B > gpurun_out/r02_f_c2_n1.json 2> gpurun_out/r02_f_c2_n1.err
timeout 400 $TR bench.py --gpus 2 --rule Bulyan --n 500 --d 25000000 --f 100 --steps 3 $B > gpurun_out/r02_f_c4_n2.json 2> gpurun_out/r02_f_c4_n2.err
grep MULTIGPU gpurun_out/r02_f_multigpu_check_n2.txt | tail -2; tail -c 600 gpurun_out/r02_f_c2_n2_peer.err
