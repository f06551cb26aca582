cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r1e.log 2>&1; tail -5 gpurun_out/pytest_gpu_r1e.log
bash profiles/scripts/r1_run_opts.sh "--fuse 2|--fuse 1" 28 40
bash profiles/scripts/r1_run_opts.sh "--workload vqse --depth 26 --fuse 2" 28 26
