for cfg in "0 0 256 64 512 2" "0 0" "0 1 512 512 512 2" "0 2 512 512 512 2" "0 1 384 384 768 3"; do
  timeout 60 tools/tf32_gemm_test $cfg | grep -v "  bad"; echo "rc=$?"
done
timeout 120 tools/tf32_gemm_test 0 1 512 512 8192 16 -20; echo "rc=$?"
timeout 120 tools/tf32_gemm_test 0 2 512 512 8192 8 -20; echo "rc=$?"
timeout 120 tools/tf32_gemm_test 0 1 1024 1024 8192 16 -10; echo "rc=$?"
timeout 120 tools/tf32_gemm_test 0 0 1024 1024 8192 16 -10; echo "rc=$?"
