timeout 200 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:gram_bf16x2 -c 1 -f -o gpurun_out/gram_bf16x2_c2 python tools/run_kernel.py gram 2 > gpurun_out/ncu_full.log 2>&1
