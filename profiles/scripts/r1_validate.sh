cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r1b.log 2>&1; tail -3 gpurun_out/pytest_gpu_r1b.log
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b32_f32_soa.log 2>&1; tail -1 gpurun_out/b32_f32_soa.log
timeout 900 python bench.py --steps 1 --warmup 1 --precision f64 --no-cpu-baseline > gpurun_out/b32_f64_merged.log 2>&1; tail -1 gpurun_out/b32_f64_merged.log
