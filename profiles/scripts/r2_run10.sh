# round 2, GPU call 10: fused tensor-core reverse step (tc_rev.cuh) -- correctness against host double, throughput
cd $GRAFT_REPO_ROOT/profiles/microbench
{
for args in "22 0 8 - 8" "22 0 8 - 6" "22 0 0 - 8" "22 0 3 - 8" "22 0 0 1,3,4,9,17,20 8" "20 0 14 - 8"; do
  echo "== tc_rev_bench $args"; timeout 120 ./tc_rev_bench $args; echo "exit $?"
done
for args in "0 28 8 - 8" "0 30 10 - 8" "0 30 10 - 6" "0 30 0 - 8" "0 30 24 - 8"; do
  echo "== tc_rev_bench $args"; timeout 120 ./tc_rev_bench $args; echo "exit $?"
done
} > ../../gpurun_out/r2_tc_rev_bench_v1.txt 2>&1
cat ../../gpurun_out/r2_tc_rev_bench_v1.txt
