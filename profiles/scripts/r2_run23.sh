# round 2, GPU call 23: final kernels (v7: packed slicing, no L2 prefetch) -- full GPU suite, default bench, launch list, ncu
cd $GRAFT_REPO_ROOT
rm -f gpurun_out/parity_large.jsonl
timeout 1200 python -m pytest tests -q -x -m gpu --durations=6 > gpurun_out/r2_pytest_gpu_v3.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/r2_pytest_gpu_v3.log
cat gpurun_out/parity_large.jsonl
timeout 900 python bench.py > gpurun_out/r2_bench_default_v3.json 2> gpurun_out/r2_bench_default_v3.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_default_v3.json"))
print({k:d[k] for k in ("value","ms_per_step","profile_ms","e2e","gpu_launches","clocks","roofline")})
print({k:(v.get("value"), v.get("check",{}).get("ok")) for k,v in d.get("secondary",{}).items()})
PY
tail -2 gpurun_out/r2_bench_default_v3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches_32q_tc_v3.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --secondary 0 --no-check > gpurun_out/r2_ncu_launches_v3.log 2>&1; echo "ncu launches exit $?"
cd profiles/microbench
timeout 120 ./tc_rev_bench 0 28 8 - 6 6 > ../../gpurun_out/r2_ncu_rev_plain_v3.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tc_block_rev -s 3 -c 1 -o ../../gpurun_out/r2_tc_rev_28q_v7 ./tc_rev_bench 0 28 8 - 6 6 > ../../gpurun_out/r2_ncu_rev_v3.log 2>&1; echo "ncu rev exit $?"
timeout 120 ./tc_block_bench 26 8 1 - 6 0 > ../../gpurun_out/r2_ncu_fwd_plain_v3.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tc_block_fwd -s 1 -c 1 -o ../../gpurun_out/r2_tc_fwd_26q_v7 ./tc_block_bench 26 8 1 - 6 0 > ../../gpurun_out/r2_ncu_fwd_v3.log 2>&1; echo "ncu fwd exit $?"
