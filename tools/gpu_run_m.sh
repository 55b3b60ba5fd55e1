mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scale_parity.py tests/test_gpu_parity.py -m gpu -q --timeout 300 -p no:cacheprovider -k "gram or krum or bulyan or alie or bf16_clients or identical" > gpurun_out/r02_m_pytest.txt 2>&1
B="--extras off --no-cpu-baseline --e2e-steps 0"
timeout 300 python bench.py --rule Bulyan --clients 500 --dim 2500000 --byzantine 100 --steps 5 $B > gpurun_out/r02_m_bulyan500.json 2> gpurun_out/r02_m_bulyan500.err
timeout 300 python bench.py --rule Krum --clients 1000 --dim 524288 --steps 5 $B > gpurun_out/r02_m_krum1000_524k.json 2> gpurun_out/r02_m_krum1000_524k.err
timeout 300 python bench.py --rule Krum --clients 1000 --dim 25000000 --steps 3 $B > gpurun_out/r02_m_krum1000_25m.json 2> gpurun_out/r02_m_krum1000_25m.err
tail -3 gpurun_out/r02_m_pytest.txt
