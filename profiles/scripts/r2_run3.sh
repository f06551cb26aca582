# round 2, GPU call 3: tensor-core block kernel v2 (W in TMEM, 4 stages, prefetch)
cd $GRAFT_REPO_ROOT/profiles/microbench
for args in "26 8 4 - 8 0" "26 8 4 - 8 1" "28 8 20 - 8 0" "28 8 20 - 6 0" "28 3 20 - 6 0" "28 0 20 3,4,9,10,17,20 6 0" "30 10 35 - 8 0" "30 10 35 - 6 0"; do
  echo "== tc_block_bench $args"; timeout 120 ./tc_block_bench $args; echo "exit $?"
done > ../../gpurun_out/r2_tc_block_bench_v2.txt 2>&1
cat ../../gpurun_out/r2_tc_block_bench_v2.txt
