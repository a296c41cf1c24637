timeout 120 tools/i8_gemm_test 1 1024 8192 16 -2 > gpurun_out/plain_i8.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_i8_sliced -s 1 -c 2 -f -o gpurun_out/i8_gemm_r01b tools/i8_gemm_test 1 1024 8192 16 -2 > gpurun_out/ncu_i8.log 2>&1
tail -2 gpurun_out/ncu_i8.log
