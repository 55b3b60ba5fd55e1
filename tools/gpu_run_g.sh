mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider -x > gpurun_out/r02_g_pytest_all.txt 2>&1
B="--extras off --no-cpu-baseline --e2e-steps 0"
timeout 300 python bench.py --steps 20 $B > gpurun_out/r02_g_c2.json 2> gpurun_out/r02_g_c2.err
timeout 300 python bench.py --rule Krum --clients 1000 --dim 524288 --steps 5 $B > gpurun_out/r02_g_krum1000_524k.json 2> gpurun_out/r02_g_krum1000_524k.err
timeout 300 python bench.py --rule Bulyan --clients 500 --dim 2500000 --byzantine 100 --steps 5 $B > gpurun_out/r02_g_bulyan500.json 2> gpurun_out/r02_g_bulyan500.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_g_smoke.txt 2>&1
tail -3 gpurun_out/r02_g_pytest_all.txt; tail -2 gpurun_out/r02_g_smoke.txt
