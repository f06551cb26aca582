# round 2, GPU call 7: tensor-core kernels v4 (coalesced item mapping), executor tests, A/B
cd $GRAFT_REPO_ROOT/profiles/microbench
{
for args in "28 8 10 - 8 0" "30 10 10 - 8 0" "30 3 4 - 8 0" "28 0 10 - 8 0" "28 0 10 0,2,5,9,17,20 8 0" "30 24 4 - 8 0"; do
  echo "== tc_block_bench $args"; timeout 120 ./tc_block_bench $args; echo "exit $?"
done
for args in "24 30 8 -" "24 28 0 -" "24 0 0 1,3,4,9,17,20"; do
  echo "== tc_grad_bench $args"; timeout 180 ./tc_grad_bench $args; echo "exit $?"
done
} > ../../gpurun_out/r2_tc_block_bench_v4.txt 2>&1
grep -E "^==|one block|gradient|after" ../../gpurun_out/r2_tc_block_bench_v4.txt
cd ../..
timeout 900 python -m pytest tests/test_tc_gpu.py -q -x --durations=5 > gpurun_out/r2_pytest_tc.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/r2_pytest_tc.log
for tc in 1; do
  timeout 600 python bench.py --qubits 30 --depth 40 --steps 2 --warmup 1 --no-cpu-baseline --secondary 0 --tc $tc > gpurun_out/r2_bench_30q_tc$tc.json 2> gpurun_out/r2_bench_30q_tc$tc.err; echo "bench tc=$tc exit $?"; python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_30q_tc$tc.json"))
print({k:d[k] for k in ("value","ms_per_step","profile_ms","check")})
PY
  tail -3 gpurun_out/r2_bench_30q_tc$tc.err
done
