mkdir -p gpurun_out
B="--extras off --no-cpu-baseline --e2e-steps 0"
timeout 300 python bench.py --rule Krum --clients 1000 --dim 524288 --steps 5 $B > gpurun_out/r02_l_krum1000_center.json 2> gpurun_out/r02_l_krum1000_center.err
AFL_GRAM_CENTER=0 timeout 300 python bench.py --rule Krum --clients 1000 --dim 524288 --steps 5 $B > gpurun_out/r02_l_krum1000_nocenter.json 2> gpurun_out/r02_l_krum1000_nocenter.err
AFL_GRAM_FLUSH=8 timeout 300 python bench.py --rule Krum --clients 1000 --dim 524288 --steps 5 $B > gpurun_out/r02_l_krum1000_flush8.json 2> gpurun_out/r02_l_krum1000_flush8.err
AFL_GRAM_SPLITS=8 timeout 300 python bench.py --rule Krum --clients 1000 --dim 524288 --steps 5 $B > gpurun_out/r02_l_krum1000_splits8.json 2> gpurun_out/r02_l_krum1000_splits8.err
