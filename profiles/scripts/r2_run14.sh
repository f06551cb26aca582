# round 2, GPU call 14: tensor-core blocks as the f32 default -- full GPU suite, default bench, launch list, ncu of the fused reverse kernel
cd $GRAFT_REPO_ROOT
rm -f gpurun_out/parity_large.jsonl
timeout 1500 python -m pytest tests -q -x -m gpu --durations=8 > gpurun_out/r2_pytest_gpu_v2.log 2>&1; echo "pytest exit $?"; tail -14 gpurun_out/r2_pytest_gpu_v2.log
cat gpurun_out/parity_large.jsonl
timeout 1200 python bench.py > gpurun_out/r2_bench_default_v2.json 2> gpurun_out/r2_bench_default_v2.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_default_v2.json"))
print({k:d[k] for k in ("value","ms_per_step","profile_ms","check","e2e","gpu_launches","clocks")})
print({k:(v.get("value"), v.get("check")) for k,v in d.get("secondary",{}).items()})
PY
tail -3 gpurun_out/r2_bench_default_v2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches_32q_tc.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --secondary 0 --no-check > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu launches exit $?"
cd profiles/microbench
timeout 120 ./tc_rev_bench 0 28 8 - 8 > ../../gpurun_out/r2_ncu_rev_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tc_block_rev -s 3 -c 1 -o ../../gpurun_out/r2_tc_rev_28q_v2 ./tc_rev_bench 0 28 8 - 8 > ../../gpurun_out/r2_ncu_rev.log 2>&1; echo "ncu rev exit $?"
