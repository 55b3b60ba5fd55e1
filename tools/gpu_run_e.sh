mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scale_parity.py -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/r02_e_pytest_scale.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/r02_e_pytest_parity.txt 2>&1
B="--extras off --no-cpu-baseline --e2e-steps 0"
timeout 300 python bench.py --rule TrimmedMean --n 1000 --d 10000000 --dtype bf16 --steps 5 $B > gpurun_out/r02_e_tm_c3.json 2> gpurun_out/r02_e_tm_c3.err
AFL_TM_KERNEL=general timeout 300 python bench.py --rule TrimmedMean --n 1000 --d 10000000 --dtype bf16 --steps 5 $B > gpurun_out/r02_e_tm_c3_general.json 2> gpurun_out/r02_e_tm_c3_general.err
timeout 300 python bench.py --rule Bulyan --n 500 --d 2500000 --f 100 --steps 5 $B > gpurun_out/r02_e_bulyan500.json 2> gpurun_out/r02_e_bulyan500.err
timeout 300 python bench.py --rule Krum --n 1000 --d 524288 --steps 5 $B > gpurun_out/r02_e_krum1000_524k.json 2> gpurun_out/r02_e_krum1000_524k.err
timeout 300 python bench.py --rule Krum --n 1000 --d 3125000 --steps 5 $B > gpurun_out/r02_e_krum1000_3m.json 2> gpurun_out/r02_e_krum1000_3m.err
timeout 300 python bench.py --steps 20 $B > gpurun_out/r02_e_c2.json 2> gpurun_out/r02_e_c2.err
AFL_GRAM_CENTER=0 timeout 300 python bench.py --steps 20 $B > gpurun_out/r02_e_c2_nocenter.json 2> gpurun_out/r02_e_c2_nocenter.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gram_pair_kernel -s 1 -c 1 -o gpurun_out/r02_ncu_gram_pair_n1000 -f python tools/run_kernel.py pair1000 2 > gpurun_out/r02_e_ncu_pair.log 2>&1
tail -3 gpurun_out/r02_e_pytest_scale.txt; tail -3 gpurun_out/r02_e_pytest_parity.txt
