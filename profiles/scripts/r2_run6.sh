# round 2, GPU call 6: ncu of the tensor-core block kernels (microbenchmarks)
cd $GRAFT_REPO_ROOT/profiles/microbench
timeout 120 ./tc_block_bench 28 8 2 - 8 0 > ../../gpurun_out/r2_ncu_tc_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tc_block_fwd -s 3 -c 1 -o ../../gpurun_out/r2_tc_fwd_28q ./tc_block_bench 28 8 2 - 8 0 > ../../gpurun_out/r2_ncu_tc.log 2>&1; echo "ncu fwd exit $?"; tail -2 ../../gpurun_out/r2_ncu_tc.log
timeout 120 ./tc_grad_bench 0 28 8 - > ../../gpurun_out/r2_ncu_tcg_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tc_block_grad -s 1 -c 1 -o ../../gpurun_out/r2_tc_grad_28q ./tc_grad_bench 0 28 8 - > ../../gpurun_out/r2_ncu_tcg.log 2>&1; echo "ncu grad exit $?"; tail -2 ../../gpurun_out/r2_ncu_tcg.log
