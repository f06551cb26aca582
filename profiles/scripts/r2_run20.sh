# round 2, GPU call 20: v6 kernels in the executor (products 6 default): tc + large parity tests, 32 q bench
cd $GRAFT_REPO_ROOT
rm -f gpurun_out/parity_large.jsonl
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_large_gpu.py -q -x --durations=5 > gpurun_out/r2_pytest_tc_v3.log 2>&1; echo "pytest exit $?"; tail -10 gpurun_out/r2_pytest_tc_v3.log
cat gpurun_out/parity_large.jsonl
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --secondary 0 > gpurun_out/r2_bench_32q_v6.json 2> gpurun_out/r2_bench_32q_v6.err; echo "bench 32q exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_32q_v6.json"))
print({k:d[k] for k in ("value","ms_per_step","profile_ms","check","clocks")})
PY
