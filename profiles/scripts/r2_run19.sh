# round 2, GPU call 19: H single accumulator + all W slices in TMEM + dedicated output staging (stage cycle = fill + MMA)
cd $GRAFT_REPO_ROOT/profiles/microbench
{
for args in "24 0 8 - 8 6" "22 0 0 - 6 6" "22 0 0 1,3,4,9,17,19 6 3" "20 0 3 - 6 6"; do
  echo "== tc_rev_bench $args"; timeout 120 ./tc_rev_bench $args; echo "exit $?"
done
for args in "0 28 8 - 6 6" "0 30 10 - 8 6" "0 30 10 - 6 6" "0 30 10 - 6 3" "0 30 0 - 6 6"; do
  echo "== tc_rev_bench $args"; timeout 120 ./tc_rev_bench $args; echo "exit $?"
done
for args in "0 28 8 - 6 6"; do
  echo "== tc_rev_trace_bench $args"; timeout 120 ./tc_rev_trace_bench $args; echo "exit $?"
done
} > ../../gpurun_out/r2_tc_rev_bench_v6.txt 2>&1
cat ../../gpurun_out/r2_tc_rev_bench_v6.txt
