# round 2, GPU call 13: fused reverse step in the executor -- tc tests, 30 q A/B, 32 q headline
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_tc_gpu.py -q -x --durations=5 > gpurun_out/r2_pytest_tc_v2.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/r2_pytest_tc_v2.log
for rev in 1 0; do
  timeout 600 python bench.py --qubits 30 --depth 40 --steps 2 --warmup 1 --no-cpu-baseline --secondary 0 --tc 1 --tc-rev $rev > gpurun_out/r2_bench_30q_tc1_rev$rev.json 2> gpurun_out/r2_bench_30q_tc1_rev$rev.err; echo "bench 30q rev=$rev exit $?"
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_30q_tc1_rev$rev.json"))
print({k:d[k] for k in ("value","ms_per_step","profile_ms","check")})
PY
done
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --secondary 0 --tc 1 > gpurun_out/r2_bench_32q_tc1_rev1.json 2> gpurun_out/r2_bench_32q_tc1_rev1.err; echo "bench 32q exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_32q_tc1_rev1.json"))
print({k:d[k] for k in ("value","ms_per_step","profile_ms","check","roofline")})
PY
