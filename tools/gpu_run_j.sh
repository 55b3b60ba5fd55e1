# final single-GPU evidence run
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/r02_j_pytest_all.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_j_smoke.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_j_bench_n1.json 2> gpurun_out/r02_j_bench_n1.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_j_bench_reference_arm.json 2> gpurun_out/r02_j_bench_reference_arm.err
B="--extras off --no-cpu-baseline --e2e-steps 0"
timeout 300 python bench.py --rule Krum --clients 1000 --dim 3125000 --dtype bf16 --steps 5 $B > gpurun_out/r02_j_krum1000_bf16.json 2> gpurun_out/r02_j_krum1000_bf16.err
timeout 300 python bench.py --rule TrimmedMean --clients 2000 --dim 1000000 --dtype bf16 --byzantine 480 --steps 3 $B > gpurun_out/r02_j_tm_n2000.json 2> gpurun_out/r02_j_tm_n2000.err
timeout 600 python tools/attack_sweep.py 1000 262144 > gpurun_out/r02_j_attack_sweep_n1000.json 2> gpurun_out/r02_j_attack_sweep_n1000.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_j_launches_bench_n1.csv python bench.py --steps 2 --warmup 3 $B --no-parity > gpurun_out/r02_j_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gram_bf16x2_kernel -s 3 -c 1 -o gpurun_out/r02_ncu_gram_bf16x2_c2 -f python tools/run_kernel.py gram 4 > gpurun_out/r02_j_ncu_c2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gram_pair_kernel -s 1 -c 1 -o gpurun_out/r02_ncu_gram_pair_n1000 -f python tools/run_kernel.py pair1000 2 > gpurun_out/r02_j_ncu_pair.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:trimmed_mean_kernel -s 1 -c 1 -o gpurun_out/r02_ncu_tm_general_bf16 -f python tools/run_kernel.py tm_bf16 2 > gpurun_out/r02_j_ncu_tm.log 2>&1
tail -3 gpurun_out/r02_j_pytest_all.txt; tail -1 gpurun_out/r02_j_smoke.txt; tail -c 300 gpurun_out/r02_j_bench_n1.err
