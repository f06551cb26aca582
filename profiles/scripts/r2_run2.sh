# round 2, GPU call 2: rest of the GPU suite, tensor-core block microbenchmark
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/parity_large.jsonl
cd profiles/microbench
for args in "26 8 4" "28 8 10" "28 3 10" "28 20 10" "28 0 10 3,4,9,10,17,20" "30 10 20"; do
  echo "== tc_block_bench $args"; timeout 120 ./tc_block_bench $args; echo "exit $?"
done > ../../gpurun_out/r2_tc_block_bench.txt 2>&1
cd ../..
cat gpurun_out/r2_tc_block_bench.txt
timeout 1800 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/r2_pytest_gpu.log
