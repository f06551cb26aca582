# round 2, GPU call 9: headline bench at 32 q with the tensor-core blocks (tc=1) vs tc=0 (2 steps each)
cd $GRAFT_REPO_ROOT
for tc in 1; do
  timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --secondary 0 --tc $tc > gpurun_out/r2_bench_32q_tc$tc.json 2> gpurun_out/r2_bench_32q_tc$tc.err; echo "bench tc=$tc exit $?"
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_32q_tc$tc.json"))
print({k:d[k] for k in ("value","ms_per_step","profile_ms","check")})
PY
  tail -3 gpurun_out/r2_bench_32q_tc$tc.err
done
