set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_soa2.log 2>&1; tail -3 gpurun_out/pytest_gpu_soa2.log
for cfg in "1 1 0" "2 1 0" "1 1 11"; do set -- $cfg
  timeout 300 python bench.py --qubits 28 --depth 40 --steps 2 --warmup 1 --fuse $1 --soa $2 --tile-bits $3 --no-cpu-baseline > gpurun_out/b28_f$1_s$2_t$3.log 2>&1
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/b28_f$1_s$2_t$3.log").read().strip().splitlines()[-1])
    print("fuse $1 soa $2 T $3: value %.1f ms/step %.1f" % (d["value"], d["ms_per_step"]), d["profile_ms"], d["gpu_launches"])
except Exception as e: print("fuse $1 soa $2 FAILED", e)
P
done
timeout 600 python bench.py --steps 1 --warmup 1 --fuse 1 --soa 1 --no-cpu-baseline > gpurun_out/b32_f1_s1_v2.log 2>&1; tail -1 gpurun_out/b32_f1_s1_v2.log
