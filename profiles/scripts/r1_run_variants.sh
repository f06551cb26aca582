# usage: bash profiles/scripts/r1_run_variants.sh "<variant> ..." [qubits] [depth] [extra bench args]
cd $GRAFT_REPO_ROOT
Q=${2:-28}; D=${3:-40}
for v in $1; do
  QDC_LIB_VARIANT=$v timeout 300 python bench.py --qubits $Q --depth $D --steps 2 --warmup 1 --fuse 1 --no-cpu-baseline $4 > gpurun_out/var_$v.log 2>&1
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/var_$v.log").read().strip().splitlines()[-1])
    print("variant $v: value %.1f ms/step %.1f" % (d["value"], d["ms_per_step"]), d["profile_ms"], "grad_norm", d["check"]["grad_norm"])
except Exception as e: print("variant $v FAILED", e)
P
done
