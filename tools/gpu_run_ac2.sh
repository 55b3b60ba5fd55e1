# 2-GPU run at the final code: sharded == unsharded check, then exactly the driver's commands for N=2 (both arms)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/multi_gpu_check.py > gpurun_out/r02_ac_multigpu_check_n2.txt 2>&1
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_ac_bench_ref_n2.json 2> gpurun_out/r02_ac_bench_ref_n2.err
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_ac_bench_n2.json 2> gpurun_out/r02_ac_bench_n2.err
grep MULTIGPU gpurun_out/r02_ac_multigpu_check_n2.txt | tail -2; tail -c 300 gpurun_out/r02_ac_bench_n2.err
