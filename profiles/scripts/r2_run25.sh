# round 2, GPU call 25: deterministic per-warp slicing grid -- correctness, run-to-run identity, throughput
cd $GRAFT_REPO_ROOT/profiles/microbench
{
for rep in 1 2; do
for args in "22 0 8 - 6 6" "20 0 0 1,3,4,9,17,19 6 6"; do
  echo "== tc_rev_bench $args (run $rep)"; timeout 120 ./tc_rev_bench $args; echo "exit $?"
done
done
for args in "0 30 10 - 6 6" "0 28 8 - 6 6"; do
  echo "== tc_rev_bench $args"; timeout 120 ./tc_rev_bench $args; echo "exit $?"
done
echo "== tc_block_bench 26 8 2 - 6 0"; timeout 120 ./tc_block_bench 26 8 2 - 6 0; echo "exit $?"
} > ../../gpurun_out/r2_tc_rev_bench_v8.txt 2>&1
cat ../../gpurun_out/r2_tc_rev_bench_v8.txt
cd ../..
timeout 300 python -m pytest tests/test_tc_gpu.py -q -x -k "not 30q" 2>&1 | tail -3
