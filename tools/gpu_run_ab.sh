mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider -k "trimmed or bulyan or golden or smoke or harness or properties" > gpurun_out/r02_ab_pytest.txt 2>&1
B="--extras off --no-cpu-baseline --e2e-steps 0"
timeout 300 python bench.py --rule Bulyan --clients 500 --dim 2500000 --byzantine 100 --steps 5 $B > gpurun_out/r02_ab_bulyan500.json 2> gpurun_out/r02_ab_bulyan500.err
timeout 300 python bench.py --rule Bulyan --clients 500 --dim 25000000 --byzantine 100 --steps 3 $B > gpurun_out/r02_ab_c4.json 2> gpurun_out/r02_ab_c4.err
timeout 300 python bench.py --rule Bulyan --clients 1000 --dim 25000000 --byzantine 240 --steps 3 $B > gpurun_out/r02_ab_c5_bulyan.json 2> gpurun_out/r02_ab_c5_bulyan.err
tail -3 gpurun_out/r02_ab_pytest.txt
