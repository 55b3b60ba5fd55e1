mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider -k "alie or ALIE or drift" > gpurun_out/r02_ad_pytest.txt 2>&1; tail -2 gpurun_out/r02_ad_pytest.txt
B="--extras off --no-cpu-baseline --e2e-steps 0 --no-parity"
for ms in 20 200 0 20 0; do
AFL_BENCH_CLOCKS_MS=$ms timeout 200 $TR bench.py --gpus 2 --steps 20 --warmup 3 $B > gpurun_out/r02_ad_n2_clk${ms}_$RANDOM.json 2>> gpurun_out/r02_ad_n2.err
done
timeout 200 python bench.py --rule ALIE --clients 1000 --dim 25000000 --byzantine 240 --steps 5 --extras off --no-cpu-baseline --e2e-steps 0 > gpurun_out/r02_ad_alie.json 2> gpurun_out/r02_ad_alie.err
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_ad_bench_n2.json 2> gpurun_out/r02_ad_bench_n2.err; echo "driver-cmd rc=$?"
