for cfg in "0 256 256 2" "1 512 512 2" "2 512 512 2" "1 384 320 3"; do
  I8_NS=6 timeout 100 tools/i8_gemm_test $cfg | grep -v "^  bad" | tr '\n' ' '; echo "rc=$?"
done
I8_NS=6 timeout 120 tools/i8_gemm_test 1 1024 8192 16 -10 | tail -3
I8_NS=6 timeout 120 tools/i8_gemm_test 0 1024 8192 16 -10 | tail -3
timeout 300 python -m pytest tests/test_gpu_int8.py -x -q 2>&1 | tail -3
for p in int8x6 int8; do timeout 300 python tools/predict_bench.py 2000000 1024 $p 2>&1 | grep "predict_y +"; done
