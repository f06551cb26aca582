# round 2, GPU call 27 (4 GPUs): merged exchanges with the balanced partner schedule -- tests + exchange A/B
cd $GRAFT_REPO_ROOT
timeout 200 python -m pytest tests/test_sharded_gpu.py -q -x -k "(tensor_core and 4-brickwork) or 4-2-1-f32-brickwork" > gpurun_out/r2_pytest_sharded_4gpu_v2.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2_pytest_sharded_4gpu_v2.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 profiles/scripts/multiswap_bench.py 30 24 > gpurun_out/r2_multiswap_bench_4gpu.txt 2>&1; echo "bench exit $?"; grep "{" gpurun_out/r2_multiswap_bench_4gpu.txt
