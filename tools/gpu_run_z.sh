# final evidence refresh (single GPU)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/r02_z_pytest_all.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_z_smoke.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_z_bench_n1.json 2> gpurun_out/r02_z_bench_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_z_launches_bench_n1.csv python bench.py --steps 2 --warmup 3 --extras off --no-cpu-baseline --e2e-steps 0 --no-parity > gpurun_out/r02_z_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gram_pair_kernel -s 1 -c 1 -o gpurun_out/r02_ncu_gram_pair_n1000 -f python tools/run_kernel.py pair1000 2 > gpurun_out/r02_z_ncu_pair.log 2>&1
tail -3 gpurun_out/r02_z_pytest_all.txt; tail -1 gpurun_out/r02_z_smoke.txt; tail -c 300 gpurun_out/r02_z_bench_n1.err
