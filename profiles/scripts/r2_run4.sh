# round 2, GPU call 4: tensor-core block kernels v3 (9-bit slices, L2 prefetch, blocks on any positions) + gradient kernel
cd $GRAFT_REPO_ROOT/profiles/microbench
{
for args in "28 8 20 - 8 0" "28 8 20 - 6 0" "30 10 35 - 8 0" "30 10 35 - 6 0" "28 0 10 - 8 0" "28 1 10 - 8 0" "28 0 10 0,2,5,9,17,20 8 0"; do
  echo "== tc_block_bench $args"; timeout 120 ./tc_block_bench $args; echo "exit $?"
done
for args in "24 30 8 -" "24 0 3 -" "24 28 0 -" "24 0 0 1,3,4,9,17,20"; do
  echo "== tc_grad_bench $args"; timeout 180 ./tc_grad_bench $args; echo "exit $?"
done
} > ../../gpurun_out/r2_tc_block_bench_v3.txt 2>&1
cat ../../gpurun_out/r2_tc_block_bench_v3.txt
