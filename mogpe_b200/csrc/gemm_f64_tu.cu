// One operand orientation of the FP64 DMMA GEMM (gemm_f64.cuh) per translation unit: compiled four times with
// -DMOGPE_GEMM_F64_TT=0 (NN), 1 (TN: A transposed), 2 (NT: B transposed), 3 (TT).
#ifndef MOGPE_GEMM_F64_TT
#error "compile with -DMOGPE_GEMM_F64_TT=0..3"
#endif
#include "gemm_f64.cuh"

namespace mogpe {
#if MOGPE_GEMM_F64_TT == 0
cudaError_t launch_gemm_nn(const GemmParams& p, int batch, cudaStream_t st, int cfg) { return launch_gemm_t<false, false>(p, batch, st, cfg); }
#elif MOGPE_GEMM_F64_TT == 1
cudaError_t launch_gemm_tn(const GemmParams& p, int batch, cudaStream_t st, int cfg) { return launch_gemm_t<true, false>(p, batch, st, cfg); }
#elif MOGPE_GEMM_F64_TT == 2
cudaError_t launch_gemm_nt(const GemmParams& p, int batch, cudaStream_t st, int cfg) { return launch_gemm_t<false, true>(p, batch, st, cfg); }
#else
cudaError_t launch_gemm_tt(const GemmParams& p, int batch, cudaStream_t st, int cfg) { return launch_gemm_t<true, true>(p, batch, st, cfg); }
#endif
}  // namespace mogpe
