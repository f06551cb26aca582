# round 2, GPU call 26 (8 GPUs): sharded tests at world 8 (merged exchanges, k = 3) and the 35 q weak-scaling point
cd $GRAFT_REPO_ROOT
timeout 200 python -m pytest tests/test_sharded_gpu.py -q -x -k "(tensor_core and 8-brickwork) or (tensor_core and 8-vqse) or 8-2-1-f32-brickwork" --durations=5 > gpurun_out/r2_pytest_sharded_8gpu.log 2>&1; echo "pytest exit $?"; tail -9 gpurun_out/r2_pytest_sharded_8gpu.log
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 1 --warmup 1 --no-cpu-baseline --secondary 0 > gpurun_out/r2_bench_8gpu_35q.json 2> gpurun_out/r2_bench_8gpu_35q.err; echo "bench 8gpu exit $?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_8gpu_35q.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","full_state_gate_applies_per_s","ms_per_step","profile_ms","check")})
PY
tail -2 gpurun_out/r2_bench_8gpu_35q.err
