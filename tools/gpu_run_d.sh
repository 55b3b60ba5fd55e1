mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scale_parity.py -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/r02_d_pytest_scale.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/r02_d_pytest_parity.txt 2>&1
B="--extras off --no-cpu-baseline --e2e-steps 0"
timeout 300 python bench.py --rule TrimmedMean --n 1000 --d 10000000 --dtype bf16 --steps 5 $B > gpurun_out/r02_d_tm_c3.json 2> gpurun_out/r02_d_tm_c3.err
timeout 300 python bench.py --rule Bulyan --n 500 --d 2500000 --f 100 --steps 5 $B > gpurun_out/r02_d_bulyan500.json 2> gpurun_out/r02_d_bulyan500.err
timeout 300 python bench.py --rule Krum --n 1000 --d 524288 --steps 5 $B > gpurun_out/r02_d_krum1000_524k.json 2> gpurun_out/r02_d_krum1000_524k.err
timeout 300 python bench.py --steps 20 $B > gpurun_out/r02_d_c2.json 2> gpurun_out/r02_d_c2.err
AFL_GRAM_CENTER=0 timeout 300 python bench.py --steps 20 $B > gpurun_out/r02_d_c2_nocenter.json 2> gpurun_out/r02_d_c2_nocenter.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:trimmed_mean_packed -s 1 -c 1 -o gpurun_out/r02_ncu_tm_packed -f python tools/run_kernel.py tm_bf16 2 > gpurun_out/r02_d_ncu_tm.log 2>&1
tail -3 gpurun_out/r02_d_pytest_scale.txt; tail -3 gpurun_out/r02_d_pytest_parity.txt
