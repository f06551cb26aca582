# round 2, GPU call 11: cycle timeline of the fused reverse kernel (CTA 0, tiles 8..15)
cd $GRAFT_REPO_ROOT/profiles/microbench
{
for args in "0 28 8 - 8" "0 28 8 - 6"; do
  echo "== tc_rev_trace_bench $args"; timeout 120 ./tc_rev_trace_bench $args; echo "exit $?"
done
} > ../../gpurun_out/r2_tc_rev_trace_v1.txt 2>&1
cat ../../gpurun_out/r2_tc_rev_trace_v1.txt
