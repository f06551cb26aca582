# round 2, GPU call 5: tensor-core blocks through the executor (option tc): parity tests, then A/B bench at 30 q
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_tc_gpu.py -q -x --durations=5 > gpurun_out/r2_pytest_tc.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/r2_pytest_tc.log
for tc in 0 1; do
  timeout 600 python bench.py --qubits 30 --depth 40 --steps 2 --warmup 1 --no-cpu-baseline --secondary 0 --tc $tc > gpurun_out/r2_bench_30q_tc$tc.json 2> gpurun_out/r2_bench_30q_tc$tc.err; echo "bench tc=$tc exit $?"; python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_30q_tc$tc.json"))
print({k:d[k] for k in ("value","ms_per_step","profile_ms","check")})
PY
  tail -3 gpurun_out/r2_bench_30q_tc$tc.err
done
