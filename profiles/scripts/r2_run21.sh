# round 2, GPU call 21 (2 GPUs): sharded execution with the tensor-core blocks -- tests and the 33 q weak-scaling point
cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 900 python -m pytest tests/test_sharded_gpu.py -q -x -k "tensor_core or (brickwork and f32)" --durations=5 > gpurun_out/r2_pytest_sharded_2gpu.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/r2_pytest_sharded_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu-baseline --secondary 0 > gpurun_out/r2_bench_2gpu_33q.json 2> gpurun_out/r2_bench_2gpu_33q.err; echo "bench 2gpu exit $?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_2gpu_33q.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","full_state_gate_applies_per_s","ms_per_step","profile_ms","check")})
PY
tail -3 gpurun_out/r2_bench_2gpu_33q.err
