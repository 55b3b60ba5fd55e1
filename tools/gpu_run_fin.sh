mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/r02_fin_pytest.txt 2>&1
tail -3 gpurun_out/r02_fin_pytest.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_fin_smoke.txt 2>&1
tail -1 gpurun_out/r02_fin_smoke.txt
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_fin_bench_ref.json 2> gpurun_out/r02_fin_bench_ref.err
timeout 600 python bench.py > gpurun_out/r02_fin_bench_n1.json 2> gpurun_out/r02_fin_bench_n1.err
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_fin_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --extras off > gpurun_out/r02_fin_bench_under_ncu.log 2>&1
