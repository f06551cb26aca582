# round 2, GPU call 8: tensor-core kernels v5 (no loads in flight across the proxy fence) + ncu
cd $GRAFT_REPO_ROOT/profiles/microbench
{
for args in "28 8 4 - 8 0" "30 10 4 - 8 0" "30 0 4 - 8 0"; do
  echo "== tc_block_bench $args"; timeout 120 ./tc_block_bench $args; echo "exit $?"
done
for args in "24 30 8 -"; do
  echo "== tc_grad_bench $args"; timeout 180 ./tc_grad_bench $args; echo "exit $?"
done
} > ../../gpurun_out/r2_tc_block_bench_v5.txt 2>&1
grep -E "^==|one block|gradient|after" ../../gpurun_out/r2_tc_block_bench_v5.txt
timeout 120 ./tc_block_bench 28 8 2 - 8 0 > ../../gpurun_out/r2_ncu_tc_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tc_block_fwd -s 3 -c 1 -o ../../gpurun_out/r2_tc_fwd_28q_v5 ./tc_block_bench 28 8 2 - 8 0 > ../../gpurun_out/r2_ncu_tc.log 2>&1; echo "ncu fwd exit $?"
