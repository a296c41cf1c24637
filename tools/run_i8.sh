for cfg in "0 128 64 1" "0 256 256 2" "1 512 512 2" "2 512 512 2" "1 384 320 3"; do
  timeout 100 tools/i8_gemm_test $cfg | grep -v "^  bad" | tr '\n' ' '; echo "rc=$?"
done
timeout 120 tools/i8_gemm_test 1 512 8192 16 -20 | tail -3
timeout 120 tools/i8_gemm_test 1 1024 8192 16 -10 | tail -3
timeout 120 tools/i8_gemm_test 0 1024 8192 16 -10 | tail -3
