# round 2, GPU call 18: proxy fence in the MMA warp instead of the fill threads
cd $GRAFT_REPO_ROOT/profiles/microbench
{
for args in "22 0 8 - 8" "20 0 0 - 6" "20 0 0 1,3,4,9,17,19 6"; do
  echo "== tc_rev_bench $args"; timeout 120 ./tc_rev_bench $args; echo "exit $?"
done
for args in "0 28 8 - 6" "0 30 10 - 8" "0 30 10 - 6" "0 30 0 - 6"; do
  echo "== tc_rev_bench $args"; timeout 120 ./tc_rev_bench $args; echo "exit $?"
done
for args in "0 28 8 - 6"; do
  echo "== tc_rev_trace_bench $args"; timeout 120 ./tc_rev_trace_bench $args; echo "exit $?"
done
for args in "28 8 2 - 6 0" "26 0 2 - 8 0"; do
  echo "== tc_block_bench $args"; timeout 120 ./tc_block_bench $args; echo "exit $?"
done
} > ../../gpurun_out/r2_tc_rev_bench_v5.txt 2>&1
cat ../../gpurun_out/r2_tc_rev_bench_v5.txt
