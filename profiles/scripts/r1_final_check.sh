cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q -k "not autodiff or not sharded" > gpurun_out/pytest_gpu_final.log 2>&1; tail -3 gpurun_out/pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -2 gpurun_out/smoke_final.log
B="python bench.py --qubits 28 --depth 20 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 300 $B > gpurun_out/ncu_plain2.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1_launches_brickwork28q_final.csv $B > gpurun_out/ncu_launches2.log 2>&1
wc -l gpurun_out/r1_launches_brickwork28q_final.csv
