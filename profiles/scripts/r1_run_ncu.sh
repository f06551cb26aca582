cd $GRAFT_REPO_ROOT
B="python bench.py --qubits 28 --depth 20 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 300 $B > gpurun_out/ncu_plain.log 2>&1 || exit 1
tail -c 600 gpurun_out/ncu_plain.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1_launches_brickwork28q.csv $B > gpurun_out/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_tile_bwd_soa -s 8 -c 1 -o gpurun_out/prof_bwd_soa_final -f $B > gpurun_out/ncu_b.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_tile_fwd_rbs -s 8 -c 1 -o gpurun_out/prof_fwd_rbs_final -f $B > gpurun_out/ncu_f.log 2>&1
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_tile_bwd_soa -s 2 -c 1 --csv --log-file gpurun_out/r1_traffic_32q_bwd.csv python bench.py --qubits 32 --depth 6 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_t.log 2>&1
tail -3 gpurun_out/r1_traffic_32q_bwd.csv
ls -la gpurun_out/*final*.ncu-rep
