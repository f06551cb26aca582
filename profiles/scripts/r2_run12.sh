# round 2, GPU call 12: elected-lane MMA issue + register-pipelined fill, all three tensor-core kernels
cd $GRAFT_REPO_ROOT/profiles/microbench
{
for args in "22 0 8 - 8" "22 0 0 - 6" "22 0 0 1,3,4,9,17,20 8"; do
  echo "== tc_rev_bench $args"; timeout 120 ./tc_rev_bench $args; echo "exit $?"
done
for args in "0 28 8 - 8" "0 30 10 - 8" "0 30 10 - 6" "0 30 0 - 8" "0 30 24 - 8"; do
  echo "== tc_rev_bench $args"; timeout 120 ./tc_rev_bench $args; echo "exit $?"
done
for args in "0 28 8 - 8" "0 28 8 - 6"; do
  echo "== tc_rev_trace_bench $args"; timeout 120 ./tc_rev_trace_bench $args; echo "exit $?"
done
for args in "28 8 4 - 8 0" "28 8 4 - 6 0" "30 10 4 - 8 0" "30 0 4 - 8 0"; do
  echo "== tc_block_bench $args"; timeout 120 ./tc_block_bench $args; echo "exit $?"
done
for args in "24 30 8 -"; do
  echo "== tc_grad_bench $args"; timeout 180 ./tc_grad_bench $args; echo "exit $?"
done
} > ../../gpurun_out/r2_tc_rev_bench_v2.txt 2>&1
cat ../../gpurun_out/r2_tc_rev_bench_v2.txt
