mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scale_parity.py -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/r02_c_pytest_scale.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/r02_c_pytest_parity.txt 2>&1
B="--extras off --no-cpu-baseline --e2e-steps 0"
timeout 300 python bench.py --rule TrimmedMean --n 1000 --d 10000000 --dtype bf16 --steps 5 $B > gpurun_out/r02_c_tm_c3.json 2> gpurun_out/r02_c_tm_c3.err
timeout 300 python bench.py --rule TrimmedMean --n 300 --d 4000000 --dtype bf16 --f 200 --steps 5 $B > gpurun_out/r02_c_tm_300.json 2> gpurun_out/r02_c_tm_300.err
timeout 300 python bench.py --rule Bulyan --n 500 --d 2500000 --f 100 --steps 5 $B > gpurun_out/r02_c_bulyan500.json 2> gpurun_out/r02_c_bulyan500.err
timeout 300 python bench.py --rule Krum --n 1000 --d 524288 --steps 5 $B > gpurun_out/r02_c_krum1000_524k.json 2> gpurun_out/r02_c_krum1000_524k.err
timeout 300 python bench.py --steps 20 $B > gpurun_out/r02_c_c2.json 2> gpurun_out/r02_c_c2.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gram_pair_kernel -s 1 -c 1 -o gpurun_out/r02_ncu_gram_pair_n1000 -f python tools/run_kernel.py pair1000 2 > gpurun_out/r02_c_ncu_pair.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:trimmed_mean_packed -s 1 -c 1 -o gpurun_out/r02_ncu_tm_packed -f python tools/run_kernel.py tm_bf16 2 > gpurun_out/r02_c_ncu_tm.log 2>&1
tail -3 gpurun_out/r02_c_pytest_scale.txt; tail -3 gpurun_out/r02_c_pytest_parity.txt
