mkdir -p gpurun_out
O=gpurun_out/r02_l2_tma_bench.txt
: > $O
for cfg in "64 7" "64 4" "64 3" "128 3" "32 14" "64 7 8192 64 1" "64 7 524288 2" "128 3 524288 2" "64 7 65536 8"; do
  timeout 60 tools/l2_tma_bench $cfg >> $O 2>&1
done
cat $O
