for cfg in "0 0 256 64 512 2" "0 0" "0 1 512 512 512 2" "0 2 512 512 512 2" "0 1 384 384 768 3" "0 2 384 384 768 3" "0 0 128 256 256 1" "0 1 128 128 256 2"; do
  timeout 60 tools/tf32_gemm_test $cfg | grep -v "  bad" | tr '\n' ' '; echo "rc=$?"
done
timeout 120 tools/tf32_gemm_test 0 1 512 512 8192 16 -20 | tail -2
timeout 120 tools/tf32_gemm_test 0 2 512 512 8192 8 -20 | tail -2
timeout 120 tools/tf32_gemm_test 0 1 1024 1024 8192 16 -10 | tail -2
timeout 120 tools/tf32_gemm_test 0 0 1024 1024 8192 16 -10 | tail -2
