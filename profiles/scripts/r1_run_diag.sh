cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r1d.log 2>&1; tail -5 gpurun_out/pytest_gpu_r1d.log
python profiles/scripts/gate_kind_bench.py 2>&1 | grep -E "^(diag|q1 )" | tee gpurun_out/gate_kind_bench2.txt
bash profiles/scripts/r1_run_opts.sh "--workload vqse --depth 26|--workload vqse --depth 26 --fuse 1|--workload vqse --depth 26 --precision f64" 28 26
