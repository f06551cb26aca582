# round 2, GPU call 1 (one B200): GPU test suite incl. the 30 / 32 q parity tests, the default bench line with its
# parity check and the secondary configs, the same-config reference arm, and a full-set ncu capture of one reverse
# pass in the shipped configuration (19-gate windows; 30 q so that ncu's save / restore of the buffers stays short).
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/parity_large.jsonl
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/r2_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench exit $?"; tail -c 2500 gpurun_out/r2_bench_default.json; tail -5 gpurun_out/r2_bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref exit $?"; cat gpurun_out/r2_bench_ref.json; tail -5 gpurun_out/r2_bench_ref.err
CMD="python bench.py --qubits 30 --depth 24 --steps 1 --warmup 1 --no-cpu-baseline --no-check --secondary 0"
timeout 600 $CMD > gpurun_out/r2_ncu_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_bwd_soa -s 12 -c 1 -o gpurun_out/r2_bwd_soa_30q $CMD > gpurun_out/r2_ncu.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/r2_ncu.log
