cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; tail -3 gpurun_out/pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -2 gpurun_out/smoke_final.log
python profiles/scripts/small_circuit_latency.py 2>&1 | grep ours | tee gpurun_out/small_circuit_latency2.txt
/usr/bin/time -v timeout 900 python bench.py > gpurun_out/bench_default_final.log 2> gpurun_out/bench_default_final.err; tail -1 gpurun_out/bench_default_final.log; grep -E "Elapsed|Maximum resident" gpurun_out/bench_default_final.err
