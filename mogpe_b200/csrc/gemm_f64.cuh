// Batched FP64 GEMM on the sm_100a DMMA pipe (mma.sync m8n8k4 f64 -> SASS DMMA.8x8x4).
//
// tcgen05.mma has no f64 kind, so the FP64 tensor path on B200 is the warp-level DMMA; measured issue
// peak on this pool: 37.0 TFLOP/s (tools/fp64_probe.cu), cuBLAS DGEMM 35.4.
//
// C[M,N] = alpha * opA(A)[M,K] * opB(B)[K,N] + beta * C, all row-major in memory:
//   TA == false: A(m,k) = A[m*lda + k]      TA == true: A(m,k) = A[k*lda + m]
//   TB == false: B(k,n) = B[k*ldb + n]      TB == true: B(k,n) = B[n*ldb + k]
// Tiles keep the orientation they have in memory, so every global read is a 16-byte cp.async along the
// contiguous axis, and the row pitches (+4 doubles) make every DMMA fragment read bank-conflict free.
// Triangular operands restrict the k range per tile (kmode) instead of multiplying zeros.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mogpe {

enum : int { KM_LEFT_LOWER = 1, KM_LEFT_UPPER = 2, KM_RIGHT_LOWER = 4, KM_RIGHT_UPPER = 8 };

struct GemmParams {
  const double* A;
  const double* B;
  double* C;
  long long sA, sB, sC;  // batch strides in elements
  const int* mapA;       // optional batch -> index maps (device), nullptr = identity
  const int* mapB;
  const int* mapC;
  int lda, ldb, ldc;
  int M, N, K;
  int kmode;
  int tril_out;  // 1: only tiles with col0 <= row0 are produced; elements above the diagonal are zeroed
  double alpha, beta;
  // optional fused column scaling of the B operand: B(k,n) *= colscale[batch_map*scs + n]  (TB == false only)
  const double* colscale;
  long long scs;
  const int* mapS;
  // optional fused per-k scaling of the A operand: A(m,k) *= kscale[batch_map*sks + k]  (fuses diag(dvar) into
  // Sbar = 2 tril(A diag(dvar) LTA^T) without a separate pass over LTA)
  const double* kscale;
  long long sks;
  const int* mapK;
  // optional fused column statistics of the RESULT tile (deterministic: one partial per row block, summed later):
  //   stat_ss [batch][M/BM][ldstat]      += sum over the tile's rows of C(m,n)^2
  //   stat_dot[v][M/BM][ldstat] (v = statvec_begin[batch] .. statvec_begin[batch+1], index via statvec[])
  //                                        = sum_m C(m,n) * statV[vec][m]
  double* stat_ss;
  double* stat_dot;
  const double* statV;      // [NV][M]
  const int* statvec_begin; // [batch + 1] (indexed by mapped C batch)
  const int* statvec;       // vector ids
  int ldstat;
  int stat_rowblocks;       // row blocks of the WHOLE matrix (M_total / BM), set by the launcher
  // row slice: this launch produces rows [row_base, row_base + M) of a result with M_total rows (triangular k ranges
  // and statistics row-block indices are absolute).  row_base must be a multiple of the row tile.
  int row_base;
  int M_total;              // 0 -> M
  // split-K: the k range of every output tile is cut into `ksplit` pieces handled by different CTAs, which ADD
  // alpha * partial into C with FP64 atomics (C must hold the beta-scaled prior value: zeros for beta = 0, set by the
  // launcher).  For [M x M] results reduced over the batch dimension when M is small: 50 CTAs cannot fill 148 SMs.
  int ksplit;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// The kernel templates below are instantiated in four separate translation units (gemm_f64_tu.cu compiled with
// -DMOGPE_GEMM_F64_TT=0..3: one per operand orientation), so that a rebuild after a change elsewhere does not pay for
// ~90 kernel instantiations; api.cu only sees the declarations.
int& gemm_cfg_override();  // defined in api.cu; < 0: automatic tile configuration
cudaError_t launch_gemm_nn(const GemmParams& p, int batch, cudaStream_t st, int cfg);
cudaError_t launch_gemm_tn(const GemmParams& p, int batch, cudaStream_t st, int cfg);
cudaError_t launch_gemm_nt(const GemmParams& p, int batch, cudaStream_t st, int cfg);
cudaError_t launch_gemm_tt(const GemmParams& p, int batch, cudaStream_t st, int cfg);

#ifdef MOGPE_GEMM_F64_TT
template <int BM, int BN, int BK, int WM, int WN, bool TA, bool TB, int STAGES>
struct GemmCfg {
  static constexpr int WARPS_M = BM / WM;
  static constexpr int WARPS_N = BN / WN;
  static constexpr int THREADS = WARPS_M * WARPS_N * 32;
  static constexpr int MI = WM / 8;
  static constexpr int NI = WN / 8;
  // smem tile shapes follow the memory orientation
  static constexpr int A_ROWS = TA ? BK : BM;
  static constexpr int A_COLS = TA ? BM : BK;
  static constexpr int A_LD = A_COLS + 4;
  static constexpr int B_ROWS = TB ? BN : BK;
  static constexpr int B_COLS = TB ? BK : BN;
  static constexpr int B_LD = B_COLS + 4;
  static constexpr int A_ELEMS = A_ROWS * A_LD;
  static constexpr int B_ELEMS = B_ROWS * B_LD;
  static constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS + BK;  // + BK per-k scales
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_ELEMS * sizeof(double);
  // resident CTAs per SM the register allocation must allow (64x64 tiles: 4 CTAs = 16 warps keep the DMMA pipe fed)
  static constexpr int MIN_CTAS = (BM * BN <= 4096) ? (THREADS <= 128 ? 4 : 3) : ((BM * BN <= 8192) ? 2 : 1);
};

// One k-tile (BK deep) of MMAs for row-fragment range [I0, I1) of this warp.  TRIL: compile-time pattern
// "quarter(j) <= quarter(i)" used on the diagonal tiles of a lower-triangular RESULT.  Everything is a
// compile-time constant, so the DMMA stream has no predicates and ptxas can software-pipeline the LDS.
template <class Cfg, int BK, bool TA, bool TB, int I0, int I1, bool TRIL, bool KS, bool CS>
__device__ __forceinline__ void mma_ktile(const double* __restrict__ As, const double* __restrict__ Bs,
                                          double (&acc)[Cfg::MI][Cfg::NI][2], const double (&cs)[Cfg::NI], int wr,
                                          int wc, int lr, int lc) {
  const double* __restrict__ Ks = Bs + Cfg::B_ELEMS;
#pragma unroll
  for (int kk = 0; kk < BK / 4; kk++) {
    double af[Cfg::MI], bf[Cfg::NI];
    const int k = kk * 4 + lc;
    const double sk = KS ? Ks[k] : 1.0;
#pragma unroll
    for (int i = I0; i < I1; i++) {
      const int m = (i * Cfg::WARPS_M + wr) * 8 + lr;
      af[i] = TA ? As[k * Cfg::A_LD + m] : As[m * Cfg::A_LD + k];
      if (KS) af[i] *= sk;
    }
#pragma unroll
    for (int j = 0; j < Cfg::NI; j++) {
      const int n = (j * Cfg::WARPS_N + wc) * 8 + lr;
      bf[j] = TB ? Bs[n * Cfg::B_LD + k] : Bs[k * Cfg::B_LD + n];
      if (CS) bf[j] *= cs[j];  // only instantiated for scaled launches: a DMUL shares the FP64 pipe with DMMA
    }
#pragma unroll
    for (int i = I0; i < I1; i++)
#pragma unroll
      for (int j = 0; j < Cfg::NI; j++)
        if (!TRIL || (j * 4 / Cfg::NI) <= (i / (Cfg::MI / 4))) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
  }
}

template <int BM, int BN, int BK, int WM, int WN, bool TA, bool TB, int STAGES, bool STATS, bool SCALE>
__global__ void __launch_bounds__(GemmCfg<BM, BN, BK, WM, WN, TA, TB, STAGES>::THREADS,
                                  GemmCfg<BM, BN, BK, WM, WN, TA, TB, STAGES>::MIN_CTAS)
    gemm_dmma_kernel(const GemmParams p) {
  using Cfg = GemmCfg<BM, BN, BK, WM, WN, TA, TB, STAGES>;
  static_assert(Cfg::MI % 4 == 0, "fragment quarters");
  extern __shared__ __align__(16) double smem[];

  const int ks = p.ksplit > 1 ? p.ksplit : 1;
  const int batch = blockIdx.z / ks, split = blockIdx.z % ks;
  // heavier (lower) row blocks first for lower-triangular left operands
  const int brow = (p.kmode & KM_LEFT_LOWER) ? (gridDim.y - 1 - blockIdx.y) : blockIdx.y;
  const int bcol = blockIdx.x;
  const int row0 = p.row_base + brow * BM, col0 = bcol * BN;
  if (p.tril_out && col0 > row0 + BM - 1) {
    // strictly-upper tile of a lower-triangular result: exact zeros (consumers read whole tiles)
    if (p.beta == 0.0 && ks == 1) {
      double* Cz = p.C + (long long)(p.mapC ? p.mapC[batch] : batch) * p.sC;
      for (int t = threadIdx.x; t < BM * BN; t += blockDim.x)
        Cz[(long long)(row0 + t / BN) * p.ldc + col0 + (t % BN)] = 0.0;
    }
    return;
  }

  int k_begin = 0, k_end = p.K;
  if (p.kmode & KM_LEFT_LOWER) k_end = min(k_end, row0 + BM);
  if (p.kmode & KM_LEFT_UPPER) k_begin = max(k_begin, row0);
  if (p.kmode & KM_RIGHT_LOWER) k_begin = max(k_begin, col0);
  if (p.kmode & KM_RIGHT_UPPER) k_end = min(k_end, col0 + BN);
  k_begin = (k_begin / BK) * BK;
  int nk = (k_end > k_begin) ? (k_end - k_begin + BK - 1) / BK : 0;
  if (ks > 1) {  // this CTA's share of the k tiles
    const int per = (nk + ks - 1) / ks, kt0 = split * per, kt1 = min(nk, kt0 + per);
    k_begin += kt0 * BK;
    nk = max(0, kt1 - kt0);
    if (nk == 0) return;
  }

  const double* __restrict__ Ab = p.A + (long long)(p.mapA ? p.mapA[batch] : batch) * p.sA;
  const double* __restrict__ Bb = p.B + (long long)(p.mapB ? p.mapB[batch] : batch) * p.sB;
  double* __restrict__ Cb = p.C + (long long)(p.mapC ? p.mapC[batch] : batch) * p.sC;
  const double* __restrict__ Sb =
      p.colscale ? p.colscale + (long long)(p.mapS ? p.mapS[batch] : batch) * p.scs : nullptr;
  const double* __restrict__ Kb =
      p.kscale ? p.kscale + (long long)(p.mapK ? p.mapK[batch] : batch) * p.sks : nullptr;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  // 8-row / 8-column fragments are INTERLEAVED across the warp grid: fragment i of warp-row wr covers tile rows
  // (i*WARPS_M + wr)*8 .. +7, so fragment quarter q = [q*MI/4, (q+1)*MI/4) covers tile rows [q*BM/4, (q+1)*BM/4)
  // for EVERY warp.  Skipping the quarters that lie in the zero half of a triangular operand (or above the
  // diagonal of a triangular result) then shortens all warps equally instead of idling half of them.
  const int wr = warp / Cfg::WARPS_N, wc = warp % Cfg::WARPS_N;
  const int lr = lane >> 2, lc = lane & 3;

  double acc[Cfg::MI][Cfg::NI][2];
#pragma unroll
  for (int i = 0; i < Cfg::MI; i++)
#pragma unroll
    for (int j = 0; j < Cfg::NI; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  auto load_stage = [&](int stage, int kt) {
    double* As = smem + stage * Cfg::STAGE_ELEMS;
    double* Bs = As + Cfg::A_ELEMS;
    const int k0 = k_begin + kt * BK;
    // A tile: A_ROWS x A_COLS doubles, 2 doubles per cp.async
    constexpr int A_CH = Cfg::A_COLS / 2;
    for (int c = tid; c < Cfg::A_ROWS * A_CH; c += Cfg::THREADS) {
      const int r = c / A_CH, cc = (c % A_CH) * 2;
      const double* g = TA ? (Ab + (long long)(k0 + r) * p.lda + row0 + cc)
                           : (Ab + (long long)(row0 + r) * p.lda + k0 + cc);
      cp_async16(As + r * Cfg::A_LD + cc, g);
    }
    constexpr int B_CH = Cfg::B_COLS / 2;
    for (int c = tid; c < Cfg::B_ROWS * B_CH; c += Cfg::THREADS) {
      const int r = c / B_CH, cc = (c % B_CH) * 2;
      const double* g = TB ? (Bb + (long long)(col0 + r) * p.ldb + k0 + cc)
                           : (Bb + (long long)(k0 + r) * p.ldb + col0 + cc);
      cp_async16(Bs + r * Cfg::B_LD + cc, g);
    }
    if (SCALE && Kb && tid < BK / 2) cp_async16(Bs + Cfg::B_ELEMS + tid * 2, Kb + k0 + tid * 2);
  };

  // per-thread column scales for its NI fragment columns (fused dLTA = 2 * LTA * dvar scaling)
  double cs[Cfg::NI];
#pragma unroll
  for (int j = 0; j < Cfg::NI; j++) cs[j] = 1.0;
  if (SCALE && Sb) {
#pragma unroll
    for (int j = 0; j < Cfg::NI; j++) cs[j] = TB ? 1.0 : Sb[col0 + (j * Cfg::WARPS_N + wc) * 8 + lr];
  }

#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) {
    if (s < nk) load_stage(s, s);
    cp_async_commit();
  }

  const bool left_lower = p.kmode & KM_LEFT_LOWER, left_upper = p.kmode & KM_LEFT_UPPER;
  const bool tril_diag = p.tril_out && row0 == col0 && BM == BN && (Cfg::NI % 4 == 0);
  constexpr int QR = BM / 4;  // rows per fragment quarter
  constexpr int Q1 = Cfg::MI / 4, Q2 = Cfg::MI / 2, Q3 = 3 * Cfg::MI / 4, Q4 = Cfg::MI;
  for (int kt = 0; kt < nk; kt++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nxt = kt + STAGES - 1;
      if (nxt < nk) load_stage(nxt % STAGES, nxt);
      cp_async_commit();
    }
    const double* As = smem + (kt % STAGES) * Cfg::STAGE_ELEMS;
    const double* Bs = As + Cfg::A_ELEMS;
    // quarters [q0, q1) of the row fragments that can hold non-zeros of the left operand in this k-tile
    const int d = k_begin + kt * BK - row0;  // k-tile start relative to the tile's first row
    int q0 = 0, q1 = 4;
    if (left_lower && d > 0) q0 = min(3, d / QR);                  // A(m,k) = 0 for k > m
    if (left_upper) q1 = max(1, min(4, (d + BK - 1) / QR + 1));    // A(m,k) = 0 for k < m
    if (SCALE && Kb) {  // per-k scaled products are triangular-result products (no left-operand structure)
      if (tril_diag) mma_ktile<Cfg, BK, TA, TB, 0, Q4, true, true, false>(As, Bs, acc, cs, wr, wc, lr, lc);
      else mma_ktile<Cfg, BK, TA, TB, 0, Q4, false, true, false>(As, Bs, acc, cs, wr, wc, lr, lc);
    } else if (tril_diag) {
      mma_ktile<Cfg, BK, TA, TB, 0, Q4, true, false, SCALE>(As, Bs, acc, cs, wr, wc, lr, lc);
    } else if (q0 == 0 && q1 == 4) {
      mma_ktile<Cfg, BK, TA, TB, 0, Q4, false, false, SCALE>(As, Bs, acc, cs, wr, wc, lr, lc);
    } else if (q0 == 1) {
      mma_ktile<Cfg, BK, TA, TB, Q1, Q4, false, false, SCALE>(As, Bs, acc, cs, wr, wc, lr, lc);
    } else if (q0 == 2) {
      mma_ktile<Cfg, BK, TA, TB, Q2, Q4, false, false, SCALE>(As, Bs, acc, cs, wr, wc, lr, lc);
    } else if (q0 == 3) {
      mma_ktile<Cfg, BK, TA, TB, Q3, Q4, false, false, SCALE>(As, Bs, acc, cs, wr, wc, lr, lc);
    } else if (q1 == 1) {
      mma_ktile<Cfg, BK, TA, TB, 0, Q1, false, false, SCALE>(As, Bs, acc, cs, wr, wc, lr, lc);
    } else if (q1 == 2) {
      mma_ktile<Cfg, BK, TA, TB, 0, Q2, false, false, SCALE>(As, Bs, acc, cs, wr, wc, lr, lc);
    } else {
      mma_ktile<Cfg, BK, TA, TB, 0, Q3, false, false, SCALE>(As, Bs, acc, cs, wr, wc, lr, lc);
    }
  }
  cp_async_wait<0>();

  if (STATS && p.stat_ss) {
    // column statistics of this tile: reduce over the thread's MI row fragments, then over the 8 lanes that share a
    // column pair (lane bits 2..4), then over the warp rows through shared memory; one plain store per column.
    __syncthreads();  // all warps are done with the operand tiles: reuse them as scratch
    double* red = smem;  // [WARPS_M][BN] per quantity
    const int cb = p.mapC ? p.mapC[batch] : batch;
    const int vb = p.statvec_begin ? p.statvec_begin[cb] : 0, ve = p.statvec_begin ? p.statvec_begin[cb + 1] : 0;
    for (int qi = -1; qi < ve - vb; qi++) {
      const double* vv = (qi >= 0) ? p.statV + (long long)p.statvec[vb + qi] * (p.M_total ? p.M_total : p.M) + row0 : nullptr;
      double w[Cfg::MI];
#pragma unroll
      for (int i = 0; i < Cfg::MI; i++) w[i] = vv ? vv[(i * Cfg::WARPS_M + wr) * 8 + lr] : 0.0;
#pragma unroll
      for (int j = 0; j < Cfg::NI; j++) {
        double s0 = 0, s1 = 0;
#pragma unroll
        for (int i = 0; i < Cfg::MI; i++) {
          const double c0 = p.alpha * acc[i][j][0], c1 = p.alpha * acc[i][j][1];
          s0 += (qi < 0) ? c0 * c0 : c0 * w[i];
          s1 += (qi < 0) ? c1 * c1 : c1 * w[i];
        }
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          s0 += __shfl_xor_sync(0xffffffffu, s0, o);
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        }
        if (lr == 0) {
          const int c = (j * Cfg::WARPS_N + wc) * 8 + lc * 2;
          red[wr * BN + c] = s0;
          red[wr * BN + c + 1] = s1;
        }
      }
      __syncthreads();
      for (int c = tid; c < BN; c += Cfg::THREADS) {
        double t = 0;
#pragma unroll
        for (int q = 0; q < Cfg::WARPS_M; q++) t += red[q * BN + c];
        const int arb = row0 / BM;  // absolute row block
        double* dst = (qi < 0) ? p.stat_ss + ((long long)cb * p.stat_rowblocks + arb) * p.ldstat
                               : p.stat_dot + ((long long)p.statvec[vb + qi] * p.stat_rowblocks + arb) * p.ldstat;
        dst[col0 + c] = t;
      }
      __syncthreads();
    }
  }

  // epilogue: thread owns (row = lr, cols = 2*lc, 2*lc+1) of each 8x8 fragment
#pragma unroll
  for (int i = 0; i < Cfg::MI; i++) {
    const int r = row0 + (i * Cfg::WARPS_M + wr) * 8 + lr;
#pragma unroll
    for (int j = 0; j < Cfg::NI; j++) {
      const int c = col0 + (j * Cfg::WARPS_N + wc) * 8 + lc * 2;
      double2* dst = reinterpret_cast<double2*>(Cb + (long long)r * p.ldc + c);
      double2 v;
      v.x = p.alpha * acc[i][j][0];
      v.y = p.alpha * acc[i][j][1];
      if (ks > 1) {
        if (!p.tril_out || c <= r) atomicAdd(&dst->x, v.x);
        if (!p.tril_out || c + 1 <= r) atomicAdd(&dst->y, v.y);
        continue;
      }
      if (p.beta != 0.0) {
        const double2 o = *dst;
        v.x += p.beta * o.x;
        v.y += p.beta * o.y;
      }
      if (p.tril_out) {
        if (c > r) v.x = 0.0;
        if (c + 1 > r) v.y = 0.0;
      }
      *dst = v;
    }
  }
}

// ---- host-side dispatch ------------------------------------------------------------------------
template <int BM, int BN, int BK, int WM, int WN, bool TA, bool TB, int STAGES, bool STATS = false, bool SCALE = false>
inline cudaError_t launch_gemm_cfg(const GemmParams& p, int batch, cudaStream_t st) {
  using Cfg = GemmCfg<BM, BN, BK, WM, WN, TA, TB, STAGES>;
  if (!SCALE && (p.colscale || p.kscale)) return cudaErrorInvalidValue;  // scaled launches use the dedicated paths
  auto kern = gemm_dmma_kernel<BM, BN, BK, WM, WN, TA, TB, STAGES, STATS, SCALE>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int ks = p.ksplit > 1 ? p.ksplit : 1;
  if (ks > 1 && (STATS || p.stat_ss || (p.beta != 0.0 && p.beta != 1.0))) return cudaErrorInvalidValue;
  dim3 grid(p.N / BN, p.M / BM, batch * ks);
  GemmParams q = p;
  q.stat_rowblocks = (p.M_total ? p.M_total : p.M) / BM;
  if (p.row_base % BM) return cudaErrorInvalidValue;
  kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(q);
  return cudaGetLastError();
}

// Tile configurations (BM, BN, BK, WM, WN, stages).  cfg < 0: automatic choice.

template <bool TA, bool TB>
inline cudaError_t launch_gemm_t(const GemmParams& p, int batch, cudaStream_t st, int cfg) {
  if (cfg < 0) cfg = gemm_cfg_override();
  if (cfg < 0) {
    // measured on B200 (tools/gemm_bench.py, profiles/gemm_sweep_r01.txt): 64x64 tiles win everywhere because
    // 4 CTAs/SM (16 warps with independent barriers) keep the DMMA pipe fed; BK = 32 with 3 stages is best for the
    // long-K products with a triangular result and for the M^3 products, BK = 16 with 2 stages for [M x M] x [M x B]
    cfg = (p.K % 32 == 0 && (p.tril_out || p.N < 1024)) ? 11 : 7;
  }
  const bool ok128 = (p.M % 128 == 0) && (p.N % 128 == 0);
  if (p.stat_ss) return launch_gemm_cfg<64, 64, 16, 32, 32, TA, TB, 2, true>(p, batch, st);  // BM = GEMM_STAT_BM
  if (p.kscale) return launch_gemm_cfg<64, 64, 32, 32, 32, TA, TB, 3, false, true>(p, batch, st);
  if (p.colscale) return launch_gemm_cfg<64, 64, 16, 32, 32, TA, TB, 2, false, true>(p, batch, st);
  switch (cfg) {
    case 0: if (ok128) return launch_gemm_cfg<128, 128, 16, 64, 32, TA, TB, 3>(p, batch, st); break;
    case 2: if (p.N % 128 == 0) return launch_gemm_cfg<64, 128, 16, 32, 32, TA, TB, 4>(p, batch, st); break;
    case 3: if (p.M % 128 == 0) return launch_gemm_cfg<128, 64, 16, 32, 32, TA, TB, 4>(p, batch, st); break;
    case 4: if (ok128) return launch_gemm_cfg<128, 128, 16, 32, 32, TA, TB, 3>(p, batch, st); break;
    case 5: if (ok128) return launch_gemm_cfg<128, 128, 16, 64, 32, TA, TB, 4>(p, batch, st); break;
    case 6: return launch_gemm_cfg<64, 64, 16, 32, 32, TA, TB, 3>(p, batch, st);
    case 7: return launch_gemm_cfg<64, 64, 16, 32, 32, TA, TB, 2>(p, batch, st);
    case 8: if (p.N % 128 == 0 && p.K % 32 == 0) return launch_gemm_cfg<64, 128, 32, 32, 32, TA, TB, 2>(p, batch, st); break;
    case 9: if (p.N % 128 == 0) return launch_gemm_cfg<64, 128, 16, 32, 32, TA, TB, 3>(p, batch, st); break;
    case 10: if (p.K % 32 == 0) return launch_gemm_cfg<64, 64, 32, 32, 32, TA, TB, 2>(p, batch, st); break;
    case 11: if (p.K % 32 == 0) return launch_gemm_cfg<64, 64, 32, 32, 32, TA, TB, 3>(p, batch, st); break;
    case 12: return launch_gemm_cfg<64, 64, 16, 32, 16, TA, TB, 2>(p, batch, st);      // 8 warps of 32x16
    case 13: return launch_gemm_cfg<64, 64, 16, 32, 16, TA, TB, 3>(p, batch, st);
    case 14: if (p.M % 128 == 0) return launch_gemm_cfg<128, 64, 16, 32, 32, TA, TB, 2>(p, batch, st); break;  // 8 warps
    case 15: if (p.N % 128 == 0) return launch_gemm_cfg<64, 128, 16, 32, 32, TA, TB, 2>(p, batch, st); break;
    case 16: if (p.K % 32 == 0) return launch_gemm_cfg<64, 64, 32, 32, 16, TA, TB, 2>(p, batch, st); break;
    case 17: if (p.K % 32 == 0) return launch_gemm_cfg<64, 64, 32, 32, 32, TA, TB, 4>(p, batch, st); break;
    default: break;
  }
  return launch_gemm_cfg<64, 64, 16, 32, 32, TA, TB, 4>(p, batch, st);
}

#endif  // MOGPE_GEMM_F64_TT

// Row blocks per matrix that the automatic configuration will use for an [M x M] x [M x B] product with fused
// column statistics (every configuration chosen for those products has BM = 64).
constexpr int GEMM_STAT_BM = 64;

inline cudaError_t launch_gemm(const GemmParams& p, int batch, bool ta, bool tb, cudaStream_t st, int cfg = -1) {
  if (p.M % 64 || p.N % 64 || p.K % 16) return cudaErrorInvalidValue;
  if (!ta && !tb) return launch_gemm_nn(p, batch, st, cfg);
  if (ta && !tb) return launch_gemm_tn(p, batch, st, cfg);
  if (!ta && tb) return launch_gemm_nt(p, batch, st, cfg);
  return launch_gemm_tt(p, batch, st, cfg);
}

}  // namespace mogpe
