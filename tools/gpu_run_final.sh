mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/r02_final_pytest.txt 2>&1
tail -3 gpurun_out/r02_final_pytest.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_final_smoke.txt 2>&1
tail -1 gpurun_out/r02_final_smoke.txt
