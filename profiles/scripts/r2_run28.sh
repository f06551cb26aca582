# round 2, GPU call 28: final state -- smoke(), full GPU suite, default bench
cd $GRAFT_REPO_ROOT
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_final.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/r2_smoke_final.log
rm -f gpurun_out/parity_large.jsonl
timeout 600 python -m pytest tests -q -x -m gpu > gpurun_out/r2_pytest_gpu_final.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2_pytest_gpu_final.log
timeout 600 python bench.py > gpurun_out/r2_bench_default_final.json 2> gpurun_out/r2_bench_default_final.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_default_final.json"))
print({k:d[k] for k in ("value","ms_per_step","profile_ms","e2e","gpu_launches","clocks")})
print({k:d["roofline"][k] for k in ("kernel","frac","dram_frac","tensor_frac","share_of_step")})
print({k:(v.get("value"), v.get("check",{}).get("ok")) for k,v in d.get("secondary",{}).items()})
print(d.get("cpu_baseline"))
PY
