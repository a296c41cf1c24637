timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_r01i.json 2> gpurun_out/bench_r01i.err; tail -2 gpurun_out/bench_r01i.err
timeout 120 tools/tf32_gemm_test 0 1 1024 1024 8192 16 -2 > gpurun_out/plain_t32.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 1 -c 2 -f -o gpurun_out/tf32_gemm_r01 tools/tf32_gemm_test 0 1 1024 1024 8192 16 -2 > gpurun_out/ncu_t32.log 2>&1
tail -3 gpurun_out/ncu_t32.log
