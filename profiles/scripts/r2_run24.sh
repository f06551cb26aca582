# round 2, GPU call 24 (4 GPUs): merged multi-qubit exchanges (k = 2) -- sharded tests at world 4
cd $GRAFT_REPO_ROOT
timeout 420 python -m pytest tests/test_sharded_gpu.py -q -x -k "(tensor_core and 4-brickwork) or (tensor_core and 4-vqse) or 4-2-1-f32-brickwork or 4-2-1-f64-autodiff" --durations=5 > gpurun_out/r2_pytest_sharded_4gpu.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/r2_pytest_sharded_4gpu.log
