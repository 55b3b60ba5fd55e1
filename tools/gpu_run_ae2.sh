mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 python bench.py > gpurun_out/r02_ae_bench_n1.json 2> gpurun_out/r02_ae_bench_n1.err; echo "n1 rc=$?"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_ae_bench_n2.json 2> gpurun_out/r02_ae_bench_n2.err; echo "n2 rc=$?"
