set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_soa.log 2>&1; tail -3 gpurun_out/pytest_gpu_soa.log
for cfg in "1 1" "2 1" "2 0" "1 0"; do set -- $cfg
  timeout 300 python bench.py --qubits 28 --depth 40 --steps 2 --warmup 1 --fuse $1 --soa $2 --no-cpu-baseline > gpurun_out/b28_f$1_s$2.log 2>&1
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/b28_f$1_s$2.log").read().strip().splitlines()[-1])
    print("fuse $1 soa $2: value %.1f ms/step %.1f" % (d["value"], d["ms_per_step"]), d["profile_ms"])
except Exception as e: print("fuse $1 soa $2 FAILED", e)
P
done
timeout 600 python bench.py --steps 1 --warmup 1 --fuse 1 --soa 1 --no-cpu-baseline > gpurun_out/b32_f1_s1.log 2>&1; tail -1 gpurun_out/b32_f1_s1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tile_bwd_soa -s 4 -c 1 -o gpurun_out/prof_bwd_soa -f python bench.py --qubits 28 --depth 20 --steps 1 --warmup 0 --fuse 1 --no-cpu-baseline > gpurun_out/ncu_bwd_soa.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tile_fwd_soa -s 4 -c 1 -o gpurun_out/prof_fwd_soa -f python bench.py --qubits 28 --depth 20 --steps 1 --warmup 0 --fuse 1 --no-cpu-baseline > gpurun_out/ncu_fwd_soa.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
