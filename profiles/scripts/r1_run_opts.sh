# usage: bash profiles/scripts/r1_run_opts.sh "<bench args set 1>|<bench args set 2>|..." [qubits] [depth]
cd $GRAFT_REPO_ROOT
Q=${2:-28}; D=${3:-40}
IFS='|' read -ra SETS <<< "$1"
k=0
for a in "${SETS[@]}"; do
  k=$((k+1))
  timeout 300 python bench.py --qubits $Q --depth $D --steps 2 --warmup 1 --no-cpu-baseline $a > gpurun_out/opt_$k.log 2>&1
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/opt_$k.log").read().strip().splitlines()[-1])
    print("[$a]: value %.1f ms/step %.1f" % (d["value"], d["ms_per_step"]), d["profile_ms"], "grad_norm", d["check"]["grad_norm"])
except Exception as e: print("[$a] FAILED", e)
P
done
