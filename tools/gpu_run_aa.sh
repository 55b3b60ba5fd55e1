mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider -k "trimmed or bulyan or golden or smoke or harness or properties" > gpurun_out/r02_aa_pytest.txt 2>&1
B="--extras off --no-cpu-baseline --e2e-steps 0"
timeout 300 python bench.py --rule TrimmedMean --clients 1000 --dim 10000000 --dtype bf16 --steps 5 $B > gpurun_out/r02_aa_tm_c3.json 2> gpurun_out/r02_aa_tm_c3.err
timeout 300 python bench.py --rule TrimmedMean --clients 1000 --dim 4000000 --dtype f32 --steps 5 $B > gpurun_out/r02_aa_tm_f32.json 2> gpurun_out/r02_aa_tm_f32.err
timeout 300 python bench.py --rule Bulyan --clients 500 --dim 2500000 --byzantine 100 --steps 5 $B > gpurun_out/r02_aa_bulyan500.json 2> gpurun_out/r02_aa_bulyan500.err
tail -3 gpurun_out/r02_aa_pytest.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:trimmed_mean_kernel -s 1 -c 1 -o gpurun_out/r02_ncu_tm_general_bf16_v5 -f python tools/run_kernel.py tm_bf16 2 > gpurun_out/r02_aa_ncu_tm.log 2>&1
