# round 2, GPU call 31 (2 GPUs): remap victims chosen by the cost model -- one sharded test and the 33 q point
cd $GRAFT_REPO_ROOT
timeout 100 python -m pytest tests/test_sharded_gpu.py -q -x -k "2-2-1-f32-autodiff or (tensor_core and 2-vqse)" > gpurun_out/r2_pytest_sharded_2gpu_v2.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2_pytest_sharded_2gpu_v2.log
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu-baseline --secondary 0 > gpurun_out/r2_bench_2gpu_33q_v2.json 2> gpurun_out/r2_bench_2gpu_33q_v2.err; echo "bench 2gpu exit $?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_2gpu_33q_v2.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","full_state_gate_applies_per_s","ms_per_step","profile_ms","gpu_launches")}, d["check"]["ok"], d["check"]["max_rel_err_gradient"])
PY
