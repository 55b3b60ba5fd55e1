mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scale_parity.py -m gpu -q --timeout 400 -p no:cacheprovider -k "gram or bulyan or alie or two_devices" > gpurun_out/r02_b_pytest_scale.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/r02_b_pytest_parity.txt 2>&1
B="--extras off --no-cpu-baseline --e2e-steps 0"
timeout 300 python bench.py --rule Bulyan --n 500 --d 2500000 --f 100 --steps 5 $B > gpurun_out/r02_b_bulyan500.json 2> gpurun_out/r02_b_bulyan500.err
timeout 300 python bench.py --rule Krum --n 1000 --d 524288 --steps 5 $B > gpurun_out/r02_b_krum1000_524k.json 2> gpurun_out/r02_b_krum1000_524k.err
timeout 300 python bench.py --rule Krum --n 1000 --d 3125000 --steps 5 $B > gpurun_out/r02_b_krum1000_3m.json 2> gpurun_out/r02_b_krum1000_3m.err
timeout 300 python bench.py --steps 20 $B > gpurun_out/r02_b_c2_center.json 2> gpurun_out/r02_b_c2_center.err
AFL_GRAM_CENTER=0 timeout 300 python bench.py --steps 20 $B > gpurun_out/r02_b_c2_nocenter.json 2> gpurun_out/r02_b_c2_nocenter.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gram_pair_kernel -s 1 -c 1 -o gpurun_out/r02_ncu_gram_pair_n1000 python tools/run_kernel.py pair1000 2 > gpurun_out/r02_b_ncu.log 2>&1
tail -3 gpurun_out/r02_b_pytest_scale.txt; tail -3 gpurun_out/r02_b_pytest_parity.txt
