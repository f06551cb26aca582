# round 2, GPU call 22: bulk L2 prefetch (0 / 4 / 8 tiles ahead), packed f32x2 slicing
cd $GRAFT_REPO_ROOT/profiles/microbench
{
for args in "22 0 8 - 8 6" "20 0 0 - 6 6" "20 0 0 1,3,4,9,17,19 6 6"; do
  echo "== tc_rev_bench $args"; timeout 120 ./tc_rev_bench $args; echo "exit $?"
done
for b in tc_rev_bench tc_rev_pf0_bench tc_rev_pf8_bench tc_rev_nofadd2_bench; do
  for args in "0 30 10 - 6 6"; do
    echo "== $b $args"; timeout 120 ./$b $args; echo "exit $?"
  done
done
for args in "0 30 0 - 6 6" "0 28 8 - 6 6" "0 30 10 - 8 6"; do
  echo "== tc_rev_bench $args"; timeout 120 ./tc_rev_bench $args; echo "exit $?"
done
echo "== tc_rev_trace_bench"; timeout 120 ./tc_rev_trace_bench 0 28 8 - 6 6
for b in tc_block_bench tc_block_pf0_bench; do
  echo "== $b 28 8 1 - 6 0"; timeout 120 ./$b 28 8 1 - 6 0; echo "exit $?"
done
} > ../../gpurun_out/r2_tc_rev_bench_v7.txt 2>&1
cat ../../gpurun_out/r2_tc_rev_bench_v7.txt
