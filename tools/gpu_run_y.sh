mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scale_parity.py -m gpu -q --timeout 300 -p no:cacheprovider -k "tile_pairs or bf16_clients" > gpurun_out/r02_y_pytest.txt 2>&1
tail -15 gpurun_out/r02_y_pytest.txt
