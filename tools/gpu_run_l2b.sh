mkdir -p gpurun_out
O=gpurun_out/r02_l2_tma_bench_b.txt
: > $O
for cfg in "64 3 8192 64 2 2" "64 3 8192 64 2 1" "128 1 8192 64 2 2" "64 2 8192 64 2 3"; do
  timeout 30 tools/l2_tma_bench $cfg >> $O 2>&1
done
cat $O
