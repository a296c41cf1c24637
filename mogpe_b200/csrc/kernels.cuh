// Kernel matrices, factorisation, streaming reductions and gradient kernels (FP64).
// Every kernel is batched over the Q kernel groups / P latent GPs of the model.
#pragma once
#include "model_dev.cuh"

namespace mogpe {

// ------------------------------------------------------------------------------------------------
// SquaredExponential Kuu (+ jitter), padded with the identity.
// Expanded squared-distance form |z/l|^2 + |z'/l|^2 - 2 (z/l).(z'/l) as gpflow's square_distance.
// grid (Mp/16, Mp/16, Q), block (16,16)
// ------------------------------------------------------------------------------------------------
__global__ void kuu_kernel(const ModelDev* __restrict__ md, const double* __restrict__ params, Layout lo,
                           double jitter, double* __restrict__ Kuu) {
  const int g = blockIdx.z, Mp = md->Mp, D = md->D, Mg = md->group_M[g];
  const int i = blockIdx.y * 16 + threadIdx.y, j = blockIdx.x * 16 + threadIdx.x;
  const double* ls = params + lo.ls + (long long)g * D;
  const double* Z = params + lo.Z + (long long)g * Mp * D;
  double v;
  if (i < Mg && j < Mg) {
    double si = 0, sj = 0, dot = 0;
    for (int d = 0; d < D; d++) {
      const double a = Z[i * D + d] / ls[d], b = Z[j * D + d] / ls[d];
      si += a * a;
      sj += b * b;
      dot += a * b;
    }
    const double r2 = -2.0 * dot + (si + sj);
    v = params[lo.kvar + g] * exp(-0.5 * r2);
    if (i == j) v += jitter;
  } else {
    v = (i == j) ? 1.0 : 0.0;
  }
  Kuu[((long long)g * Mp + i) * Mp + j] = v;
}

// ------------------------------------------------------------------------------------------------
// Kuf[g][m][n] = k(z_m, x_n); zero for padded rows / columns.  X streamed coalesced, Z staged in smem.
// grid (Bp/128, Mp/32, Q), block 128.  Dynamic smem: 32*(DMAX+1) doubles.
// ------------------------------------------------------------------------------------------------
template <int DMAX>
__global__ void __launch_bounds__(128) kuf_kernel(const ModelDev* __restrict__ md, const double* __restrict__ params,
                                                  Layout lo, const double* __restrict__ X, int N, int Bp,
                                                  double* __restrict__ Kuf) {
  extern __shared__ double zs[];  // [32][DMAX] scaled Z rows (zero beyond D / Mg), then [32] squared norms
  const int g = blockIdx.z, Mp = md->Mp, D = md->D, Mg = md->group_M[g];
  const int m0 = blockIdx.y * 32;
  const double* ls = params + lo.ls + (long long)g * D;
  const double* Z = params + lo.Z + (long long)g * Mp * D;
  double* zn = zs + 32 * DMAX;
  for (int t = threadIdx.x; t < 32 * DMAX; t += blockDim.x) {
    const int r = t / DMAX, d = t % DMAX;
    zs[t] = (d < D && m0 + r < Mg) ? Z[(m0 + r) * D + d] / ls[d] : 0.0;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    double s = 0;
#pragma unroll
    for (int d = 0; d < DMAX; d++) s += zs[threadIdx.x * DMAX + d] * zs[threadIdx.x * DMAX + d];
    zn[threadIdx.x] = s;
  }
  __syncthreads();
  const int n = blockIdx.x * 128 + threadIdx.x;
  const double kv = params[lo.kvar + g];
  const bool valid = n < N;
  double xs[DMAX];  // registers: every loop over d is fully unrolled
  double xn = 0;
#pragma unroll
  for (int d = 0; d < DMAX; d++) {
    xs[d] = (valid && d < D) ? X[(long long)n * D + d] / ls[d] : 0.0;
    xn += xs[d] * xs[d];
  }
  double* out = Kuf + ((long long)g * Mp + m0) * Bp + n;
  // four rows per iteration: four independent exp chains in flight per thread (the chain is latency-bound otherwise)
#pragma unroll 2
  for (int r = 0; r < 32; r += 4) {
    double v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      double dot = 0;
#pragma unroll
      for (int d = 0; d < DMAX; d++) dot += zs[(r + u) * DMAX + d] * xs[d];
      const double r2 = -2.0 * dot + (zn[r + u] + xn);
      v[u] = kv * exp(-0.5 * r2);
    }
#pragma unroll
    for (int u = 0; u < 4; u++) out[(long long)(r + u) * Bp] = (valid && m0 + r + u < Mg) ? v[u] : 0.0;
  }
}

// ------------------------------------------------------------------------------------------------
// Blocked left-looking Cholesky FUSED with the triangular inverse, one launch per 32-wide block column j.
// Per group, Mp/32 CTAs:
//   b <  j : Linv[j, b] = -Ljj^-1 sum_{k=b}^{j-1} L[j,k] Linv[k,b]      (row block j of L^-1 is complete now)
//   b == j : L[j,j] = chol(D),  Linv[j,j] = Ljj^-1
//   b >  j : L[b, j] = (Kuu[b,j] - sum_{k<j} L[b,k] L[j,k]^T) Ljj^-T
// with D = Kuu[j,j] - sum_{k<j} L[j,k] L[j,k]^T recomputed by every CTA (cheap, avoids a second launch).
// The explicit inverse turns every later triangular solve into a DMMA GEMM.
// Reads Kuu, writes L / Linv (separate buffers: no read/write race on the diagonal block).
// grid (Mp/32, Q), block 256.
// ------------------------------------------------------------------------------------------------
constexpr int CHOL_CH = 128;          // k-chunk of the block products (one cp.async round trip each)
constexpr int CHOL_PA = CHOL_CH + 4;  // row pitch = 4 (mod 16) doubles: conflict-free DMMA fragment reads
constexpr int CHOL_PC = 36;
constexpr size_t CHOL_SMEM_BYTES = (32 * CHOL_PA + CHOL_CH * CHOL_PC + 2 * 32 * 33) * sizeof(double);
static_assert(CHOL_CH * CHOL_PC >= 32 * CHOL_PA, "Pb must fit in the Pc region");

__device__ __forceinline__ void cp_async16_k(void* smem, const void* gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all_k() { asm volatile("cp.async.wait_all;\n" ::); }

__device__ __forceinline__ void dmma884_k(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Cholesky of a 32x32 block held in shared memory, by ONE warp: lane r keeps row r in registers, pivots and
// multipliers travel by shuffles (no block barriers on the sequential critical path).  rdiag[c] = 1 / L[c][c].
__device__ __forceinline__ void chol32_warp(double (*Dg)[33], double* rdiag, int lane, int* fail_flag, int fail_code) {
  double a[32];
  bool bad = false;
#pragma unroll
  for (int c = 0; c < 32; c++) a[c] = Dg[lane][c];
#pragma unroll
  for (int c = 0; c < 32; c++) {
    const double piv = __shfl_sync(0xffffffffu, a[c], c);
    if (!(piv > 0.0)) bad = true;   // non-positive or NaN pivot: matrix is not positive definite
    const double rs = rsqrt(piv);  // NaN for a non-positive pivot: also propagates to the ELBO
    a[c] = (lane == c) ? piv * rs : a[c] * rs;  // rows above the pivot hold junk in column c; never read
    if (lane == c) rdiag[c] = rs;
#pragma unroll
    for (int cc = c + 1; cc < 32; cc++) {
      const double lcc = __shfl_sync(0xffffffffu, a[c], cc);  // L[cc][c]
      a[cc] -= a[c] * lcc;
    }
  }
#pragma unroll
  for (int c = 0; c < 32; c++) Dg[lane][c] = (c <= lane) ? a[c] : 0.0;
  if (bad && lane == 0) atomicMax(fail_flag, fail_code);
}

// Solve T y = b for the 32 columns of Rb in place (T lower-triangular 32x32 in smem), ONE warp, lane = column.
// Right-looking: once y[r] is known it is eliminated from all later rows (independent FMAs).
__device__ __forceinline__ void trsm32_cols_warp(const double (*T)[33], const double* rdiag, double (*Rb)[33], int lane) {
  double y[32];
#pragma unroll
  for (int r = 0; r < 32; r++) y[r] = Rb[r][lane];
#pragma unroll
  for (int r = 0; r < 32; r++) {
    y[r] *= rdiag[r];
#pragma unroll
    for (int rr = r + 1; rr < 32; rr++) y[rr] -= T[rr][r] * y[r];
  }
#pragma unroll
  for (int r = 0; r < 32; r++) Rb[r][lane] = y[r];
}

// Solve x T^T = c for the 32 rows of Cb in place, ONE warp, lane = row.
__device__ __forceinline__ void trsm32_rows_warp(const double (*T)[33], const double* rdiag, double (*Cb)[33], int lane) {
  double x[32];
#pragma unroll
  for (int c = 0; c < 32; c++) x[c] = Cb[lane][c];
#pragma unroll
  for (int c = 0; c < 32; c++) {
    x[c] *= rdiag[c];
#pragma unroll
    for (int cc = c + 1; cc < 32; cc++) x[cc] -= T[cc][c] * x[c];
  }
#pragma unroll
  for (int c = 0; c < 32; c++) Cb[lane][c] = x[c];
}

// Step j of the fused Cholesky + inverse (launch j = -1 .. Mp/32 - 1), grid (Mp/32, Q), block 256.
//   B-part, CTA b != j (j >= 0): one 32x32 block product on the DMMA pipe, then a single-warp 32x32 triangular
//     solve against Ljj (all CTAs in parallel):
//        b > j : L[b,j]    = (Kuu[b,j] - sum_{k<j} L[b,k] L[j,k]^T) Ljj^-T
//        b < j : Linv[j,b] = -Ljj^-1 sum_{k=b}^{j-1} L[j,k] Linv[k,b]
//   A-part (look-ahead), CTA b == j+1: it already streams row block j+1 of L for its B-part, so it accumulates
//     sum_{k<j} L[j+1,k] L[j+1,k]^T in the same sweep, adds the outer product of the block it just produced and
//     factors the NEXT diagonal block (one warp, registers + shuffles).  Its inverse Linv[j+1,j+1] is computed
//     by CTA b == j+1 of the next launch, off the critical path.
// The 32x32 outputs are split over the 8 warps (fragment row i = warp/2, fragment columns 2*(warp%2) + {0,1}):
// no cross-warp reduction, deterministic, 8 accumulator registers per product.
// The only sequential work per launch is one block product + one 32x32 factor/invert by ONE CTA per group.
// One step of the chain for CTA (bx, g); shared by the per-step kernel and the single-launch chain kernel (where blocks of
// L / L^-1 written by OTHER CTAs earlier in the same launch are read through the L2: cp.async.cg and ld.global.cg).
__device__ __forceinline__ void chol_inv_step(const double* __restrict__ Kuu, double* __restrict__ L, double* __restrict__ Linv,
                                              int Mp, int j, int* __restrict__ fail_flag, int bx, int g, double* chol_smem,
                                              double* rdiag) {
  double(*Pa)[CHOL_PA] = reinterpret_cast<double(*)[CHOL_PA]>(chol_smem);
  double(*Pb)[CHOL_PA] = reinterpret_cast<double(*)[CHOL_PA]>(chol_smem + 32 * CHOL_PA);
  double(*Pc)[CHOL_PC] = reinterpret_cast<double(*)[CHOL_PC]>(chol_smem + 32 * CHOL_PA);  // Linv[k, b] chunk (k rows)
  double(*Dg)[33] = reinterpret_cast<double(*)[33]>(chol_smem + 32 * CHOL_PA + CHOL_CH * CHOL_PC);
  double(*Cg)[33] = Dg + 32;
  const int nb = Mp / 32;
  const int b = (bx + j + 1) % nb;  // the look-ahead CTA (critical path) is scheduled first
  const double* Kg = Kuu + (long long)g * Mp * Mp;
  double* Lg = L + (long long)g * Mp * Mp;
  double* Xg = Linv + (long long)g * Mp * Mp;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
  const int d = j + 1;
  // launch j also lets CTA b == j invert Ljj (factored by the previous launch): off the critical path
  const bool do_b = (j >= 0), do_a = (b == d) && (d < nb);
  if (!do_b && !do_a) return;
  const int fi = warp >> 1, fq0 = (warp & 1) * 2;  // this warp's output fragments
  double accC[2][2] = {{0, 0}, {0, 0}}, accD[2][2] = {{0, 0}, {0, 0}};

  if (do_b) {
    const int kend = (b == j) ? 0 : j * 32, kbeg = (b > j) ? 0 : b * 32;  // b == j: no product, Cg = I
    for (int k0 = kbeg; k0 < kend; k0 += CHOL_CH) {
      const int kw = min(CHOL_CH, kend - k0);  // multiple of 32
      for (int t = tid; t < 32 * (kw / 2); t += 256) {
        const int r = t / (kw / 2), c = (t % (kw / 2)) * 2;
        cp_async16_k(&Pa[r][c], Lg + (long long)(j * 32 + r) * Mp + k0 + c);
        if (b > j) cp_async16_k(&Pb[r][c], Lg + (long long)(b * 32 + r) * Mp + k0 + c);
      }
      if (b < j)
        for (int t = tid; t < kw * 16; t += 256) {
          const int r = t >> 4, c = (t & 15) * 2;
          cp_async16_k(&Pc[r][c], Xg + (long long)(k0 + r) * Mp + b * 32 + c);
        }
      cp_async_wait_all_k();
      __syncthreads();
      if (b > j) {
#pragma unroll 4
        for (int k = lc; k < kw; k += 4) {  // C += L[b,k] L[j,k]^T ; look-ahead: D += L[b,k] L[b,k]^T
          const double a = Pb[fi * 8 + lr][k];
          dmma884_k(accC[0][0], accC[0][1], a, Pa[fq0 * 8 + lr][k]);
          dmma884_k(accC[1][0], accC[1][1], a, Pa[(fq0 + 1) * 8 + lr][k]);
          if (do_a) {
            dmma884_k(accD[0][0], accD[0][1], a, Pb[fq0 * 8 + lr][k]);
            dmma884_k(accD[1][0], accD[1][1], a, Pb[(fq0 + 1) * 8 + lr][k]);
          }
        }
      } else {
#pragma unroll 4
        for (int k = lc; k < kw; k += 4) {  // R += L[j,k] Linv[k,b]
          const double a = Pa[fi * 8 + lr][k];
          dmma884_k(accC[0][0], accC[0][1], a, Pc[k][fq0 * 8 + lr]);
          dmma884_k(accC[1][0], accC[1][1], a, Pc[k][(fq0 + 1) * 8 + lr]);
        }
      }
      __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int r = fi * 8 + lr, c = (fq0 + u) * 8 + lc * 2;
      const double2 base = (b > j) ? *reinterpret_cast<const double2*>(Kg + (long long)(b * 32 + r) * Mp + j * 32 + c)
                                   : make_double2(0.0, 0.0);
      Cg[r][c] = ((b == j && r == c) ? 1.0 : base.x) - accC[u][0];
      Cg[r][c + 1] = ((b == j && r == c + 1) ? 1.0 : base.y) - accC[u][1];
    }
    for (int t = tid; t < 1024; t += 256)
      Dg[t >> 5][t & 31] = __ldcg(Lg + (long long)(j * 32 + (t >> 5)) * Mp + j * 32 + (t & 31));  // Ljj (previous step)
    if (tid < 32) rdiag[tid] = 1.0 / __ldcg(Lg + (long long)(j * 32 + tid) * Mp + j * 32 + tid);
    __syncthreads();
    // one warp solves against Ljj (every CTA in parallel; nothing waits for an explicit Ljj^-1):
    //   b > j : X Ljj^T = C      b < j : Ljj Y = -R
    if (tid < 32) {
      if (b > j) trsm32_rows_warp(Dg, rdiag, Cg, tid);
      else trsm32_cols_warp(Dg, rdiag, Cg, tid);
    }
    __syncthreads();
    double o[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int t = tid + u * 256;
      o[u] = Cg[t >> 5][t & 31];
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int t = tid + u * 256, r = t >> 5, c = t & 31;
      if (b > j) {
        Lg[(long long)(b * 32 + r) * Mp + j * 32 + c] = o[u];
        Lg[(long long)(j * 32 + r) * Mp + b * 32 + c] = 0.0;   // mirrored upper block of L
        Xg[(long long)(j * 32 + r) * Mp + b * 32 + c] = 0.0;   // and of Linv
        Cg[r][c] = o[u];                                        // keep L[b,j] for the look-ahead
      } else {
        Xg[(long long)(j * 32 + r) * Mp + b * 32 + c] = o[u];
      }
    }
    if (!do_a) return;
    __syncthreads();
  }

  // ---- A-part: D = Kuu[d,d] - sum_{k<j*32} L[d,k] L[d,k]^T (accD) - L[d,j] L[d,j]^T (block just produced, in Cg)
#pragma unroll
  for (int u = 0; u < 2; u++) {
    const int r = fi * 8 + lr, c = (fq0 + u) * 8 + lc * 2;
    const double2 base = *reinterpret_cast<const double2*>(Kg + (long long)(d * 32 + r) * Mp + d * 32 + c);
    double s0 = 0, s1 = 0;
    if (j >= 0) {
#pragma unroll 8
      for (int q = 0; q < 32; q++) {
        s0 += Cg[r][q] * Cg[c][q];
        s1 += Cg[r][q] * Cg[c + 1][q];
      }
    }
    Dg[r][c] = base.x - accD[u][0] - s0;
    Dg[r][c + 1] = base.y - accD[u][1] - s1;
  }
  __syncthreads();
  if (tid < 32) chol32_warp(Dg, rdiag, tid, fail_flag, g + 1);  // flag = 1 + index of the failing group
  __syncthreads();
  for (int t = tid; t < 1024; t += 256) {
    const int r = t >> 5, c = t & 31;
    Lg[(long long)(d * 32 + r) * Mp + d * 32 + c] = Dg[r][c];
  }
}


__global__ void __launch_bounds__(256, 2) chol_inv_step_kernel(const double* __restrict__ Kuu, double* __restrict__ L,
                                                                double* __restrict__ Linv, int Mp, int j,
                                                                int* __restrict__ fail_flag) {
  extern __shared__ __align__(16) double chol_smem[];  // Pa | Pb/Pc (never live together) | Dg | Cg
  __shared__ double rdiag[32];
  chol_inv_step(Kuu, L, Linv, Mp, j, fail_flag, blockIdx.x, blockIdx.y, chol_smem, rdiag);
}

// The whole chain in ONE launch: every CTA loops over the steps j = -1 .. Mp/32 - 1 and the Mp/32 CTAs of a group meet at a
// counter in global memory between steps (release: __threadfence + atomicAdd; acquire: ld.acquire.gpu spin by one thread).
// All CTAs of the grid must be co-resident (the host checks the occupancy and falls back to per-step launches otherwise); a
// bounded spin turns an impossible wait into a reported failure instead of a hang.
// grid (Mp/32, Q), block 256; sync [Q] zero-initialised.
__global__ void __launch_bounds__(256, 2) chol_inv_chain_kernel(const double* __restrict__ Kuu, double* __restrict__ L,
                                                                 double* __restrict__ Linv, int Mp,
                                                                 int* __restrict__ fail_flag, unsigned* __restrict__ sync) {
  extern __shared__ __align__(16) double chol_smem[];
  __shared__ double rdiag[32];
  __shared__ int s_abort;
  const int g = blockIdx.y, nb = Mp / 32;
  if (threadIdx.x == 0) s_abort = 0;
  for (int j = -1; j < nb; j++) {
    chol_inv_step(Kuu, L, Linv, Mp, j, fail_flag, blockIdx.x, g, chol_smem, rdiag);
    if (j == nb - 1) break;  // nothing reads after the last step inside this launch
    __syncthreads();          // every thread's stores of this step are issued
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(sync + g, 1u);
      const unsigned target = (unsigned)nb * (unsigned)(j + 2);
      unsigned seen;
      long long spins = 0;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(sync + g) : "memory");
        if (++spins > 4000000LL) {  // seconds of L2 round trips: the grid is not co-resident
          s_abort = 1;
          atomicMax(fail_flag, 1000 + g);
          break;
        }
      } while (seen < target);
    }
    __syncthreads();
    if (s_abort) return;
  }
}

// ------------------------------------------------------------------------------------------------
// Clean, padded lower-triangular copy of q_sqrt (GEMM operands must hold exact zeros above the diagonal)
// grid (Mp/16, Mp/16, P), block (16,16)
// ------------------------------------------------------------------------------------------------
__global__ void prep_qsqrt_kernel(const ModelDev* __restrict__ md, const double* __restrict__ params, Layout lo,
                                  double* __restrict__ Sc) {
  const int l = blockIdx.z, Mp = md->Mp, Mg = md->group_M[md->latent_group[l]];
  const int i = blockIdx.y * 16 + threadIdx.y, j = blockIdx.x * 16 + threadIdx.x;
  const long long idx = ((long long)l * Mp + i) * Mp + j;
  Sc[idx] = (i < Mg && j <= i) ? params[lo.q_sqrt + idx] : 0.0;
}

// ------------------------------------------------------------------------------------------------
// Mean vectors: V[v][m] = q_mu[l][m] (+ sum_{k<=m} S_l[m][k] eps_u[l][r][k] for inducing-sampled latents)
// (TFP MultivariateNormalTriL.sample, mogpe/keras/gps.py:20-29).  One warp per (vector, row m).
// grid (Mp/8, NV), block 256
// ------------------------------------------------------------------------------------------------
__global__ void make_V_kernel(const ModelDev* __restrict__ md, const double* __restrict__ params, Layout lo,
                              const double* __restrict__ Sc, const double* __restrict__ eps_u, int S,
                              double* __restrict__ V) {
  const int v = blockIdx.y, Mp = md->Mp;
  const int l = md->vec_latent[v], r = v - md->voff[l];
  const int Mg = md->group_M[md->latent_group[l]];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * 8 + warp;
  double s = 0.0;
  const bool sampled = (md->R[l] == S) && !md->marg[l] && eps_u != nullptr;
  if (m < Mg && sampled) {
    const double* row = Sc + ((long long)l * Mp + m) * Mp;
    const double* e = eps_u + ((long long)l * S + r) * Mp;
    for (int k = lane; k <= m; k += 32) s += row[k] * e[k];
  }
  s = warp_sum(s);
  if (lane == 0) V[(long long)v * Mp + m] = (m < Mg) ? params[lo.q_mu + (long long)l * Mp + m] + s : 0.0;
}

// ------------------------------------------------------------------------------------------------
// One streaming pass over A[g] ([Mp, Bp], n contiguous): column sums of squares and the means
// mean[v][n] = sum_m A[m][n] V[v][m] + c_l for up to 8 vectors of the group per pass (warp-free: one
// column per thread, coalesced across the warp).
// grid (Bp/128, Q), block 128, dynamic smem 8*Mp doubles
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) a_stats_kernel(const ModelDev* __restrict__ md, const double* __restrict__ params,
                                                      Layout lo, const double* __restrict__ A,
                                                      const double* __restrict__ V, int Bp,
                                                      double* __restrict__ var0ss, double* __restrict__ mean) {
  extern __shared__ double vs[];  // [8][Mp]
  const int g = blockIdx.y, Mp = md->Mp;
  const int n = blockIdx.x * 128 + threadIdx.x;
  const double* Ag = A + (long long)g * Mp * Bp + n;
  const int vb = md->gvec_begin[g], ve = md->gvec_begin[g + 1];
  bool first = true;
  for (int v0 = vb; v0 < ve || first; v0 += 8) {
    const int nv = min(8, ve - v0);
    __syncthreads();
    for (int t = threadIdx.x; t < nv * Mp; t += 128) vs[t] = V[(long long)md->gvec[v0 + t / Mp] * Mp + (t % Mp)];
    __syncthreads();
    double ss = 0, acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 4
    for (int m = 0; m < Mp; m++) {
      const double a = Ag[(long long)m * Bp];
      ss += a * a;
#pragma unroll
      for (int q = 0; q < 8; q++)
        if (q < nv) acc[q] += a * vs[q * Mp + m];
    }
    if (first) var0ss[(long long)g * Bp + n] = ss;
    for (int q = 0; q < nv; q++) {
      const int v = md->gvec[v0 + q];
      mean[(long long)v * Bp + n] = acc[q] + params[lo.mean_c + md->vec_latent[v]];
    }
    first = false;
  }
}

// Sum the per-row-block partial column statistics written by the GEMM epilogues (deterministic order):
//   var0ss[g][n] = sum_rb ss_part[g][rb][n];  mean[v][n] = sum_rb dot_part[v][rb][n] + c_latent(v);
//   varq[l][n] = sum_rb lta_part[slot(l)][rb][n].     grid (Bc/128, Q + NV + nlta), block 128
__global__ void __launch_bounds__(128) stats_finish_kernel(const ModelDev* __restrict__ md, const double* __restrict__ params,
                                                           Layout lo, const double* __restrict__ ss_part,
                                                           const double* __restrict__ dot_part,
                                                           const double* __restrict__ lta_part, int rbn, int Bp,
                                                           double* __restrict__ var0ss, double* __restrict__ mean,
                                                           double* __restrict__ varq) {
  const int n = blockIdx.x * 128 + threadIdx.x, Q = md->Q, NV = md->NV;
  int y = blockIdx.y;
  const double* src;
  double* dst;
  double add = 0.0;
  if (y < Q) {
    src = ss_part + (long long)y * rbn * Bp;
    dst = var0ss + (long long)y * Bp;
  } else if (y < Q + NV) {
    y -= Q;
    src = dot_part + (long long)y * rbn * Bp;
    dst = mean + (long long)y * Bp;
    add = params[lo.mean_c + md->vec_latent[y]];
  } else {
    y -= Q + NV;
    src = lta_part + (long long)y * rbn * Bp;
    dst = varq + (long long)md->lta_latent[y] * Bp;
  }
  double t = 0;
  for (int rb = 0; rb < rbn; rb++) t += src[(long long)rb * Bp + n];
  dst[n] = t + add;
}

// varq[l][n] = sum_m LTA[slot][m][n]^2.   grid (Bp/128, nlta), block 128
__global__ void __launch_bounds__(128) col_sumsq_kernel(const ModelDev* __restrict__ md, const double* __restrict__ LTA,
                                                        int Bp, double* __restrict__ varq) {
  const int slot = blockIdx.y, Mp = md->Mp, l = md->lta_latent[slot];
  const int n = blockIdx.x * 128 + threadIdx.x;
  const double* p = LTA + (long long)slot * Mp * Bp + n;
  double ss = 0;
  for (int m = 0; m < Mp; m += 8) {
    double a[8];
#pragma unroll
    for (int u = 0; u < 8; u++) a[u] = p[(long long)(m + u) * Bp];
#pragma unroll
    for (int u = 0; u < 8; u++) ss += a[u] * a[u];
  }
  varq[(long long)l * Bp + n] = ss;
}

// ------------------------------------------------------------------------------------------------
// KL terms (Appendix A.5, whitened gauss_kl) and their gradients w.r.t. q_mu, q_sqrt.
// grid (Mp/8, P), block 256: one warp per row of q_sqrt (coalesced).  out[2] += KL of gating latents,
// out[3] += KL of expert latents.  grad convention: gradient of the ELBO (= data term - KL).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) kl_kernel(const ModelDev* __restrict__ md, const double* __restrict__ params,
                                                 Layout lo, double klw, double* __restrict__ out, double* __restrict__ grads) {
  __shared__ double red[8];
  const int l = blockIdx.y, Mp = md->Mp, Mg = md->group_M[md->latent_group[l]];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + warp;
  double acc = 0;
  if (i < Mg) {
    const double* row = params + lo.q_sqrt + ((long long)l * Mp + i) * Mp;
    double* grow = grads ? grads + lo.q_sqrt + ((long long)l * Mp + i) * Mp : nullptr;
    for (int j = lane; j <= i; j += 32) {
      const double s = row[j];
      acc += s * s;
      if (j == i) acc -= 1.0 + log(s * s);
      if (grow) grow[j] -= klw * ((i == j) ? (s - 1.0 / s) : s);
    }
    if (lane == 0) {
      const double q = params[lo.q_mu + (long long)l * Mp + i];
      acc += q * q;
      if (grads) grads[lo.q_mu + (long long)l * Mp + i] -= klw * q;
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int w = 0; w < 8; w++) s += red[w];
    if (s != 0.0) atomicAdd(out + (l >= md->Pe ? 2 : 3), 0.5 * klw * s);
  }
}

// per-latent KL (same arithmetic as kl_kernel, no gradients): kl[l] += ...   grid (Mp/8, P), block 256
__global__ void __launch_bounds__(256) kl_per_latent_kernel(const ModelDev* __restrict__ md, const double* __restrict__ params,
                                                            Layout lo, double* __restrict__ kl) {
  __shared__ double red[8];
  const int l = blockIdx.y, Mp = md->Mp, Mg = md->group_M[md->latent_group[l]];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + warp;
  double acc = 0;
  if (i < Mg) {
    const double* row = params + lo.q_sqrt + ((long long)l * Mp + i) * Mp;
    for (int j = lane; j <= i; j += 32) {
      const double s = row[j];
      acc += s * s;
      if (j == i) acc -= 1.0 + log(s * s);
    }
    if (lane == 0) {
      const double q = params[lo.q_mu + (long long)l * Mp + i];
      acc += q * q;
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int w = 0; w < 8; w++) s += red[w];
    if (s != 0.0) atomicAdd(kl + l, 0.5 * s);
  }
}

// f_mean[s][n][l] = mean[voff[l] + s][n] ; f_var[n][l] = kvar[g] - var0ss[g][n]   (all latents inducing-sampled)
// grid (ceil(N/128)), block 128
__global__ void inducing_out_kernel(const ModelDev* __restrict__ md, const double* __restrict__ params, Layout lo,
                                    const double* __restrict__ mean, const double* __restrict__ var0ss, int N, int Bp,
                                    double* __restrict__ f_mean, double* __restrict__ f_var) {
  const int n = blockIdx.x * 128 + threadIdx.x;
  if (n >= N) return;
  const int P = md->P, S = md->S;
  for (int l = 0; l < P; l++) {
    const int g = md->latent_group[l];
    if (f_var) f_var[(long long)n * P + l] = params[lo.kvar + g] - var0ss[(long long)g * Bp + n];
    if (f_mean)
      for (int s = 0; s < S; s++) f_mean[((long long)s * N + n) * P + l] = mean[(long long)(md->voff[l] + s) * Bp + n];
  }
}

__global__ void finish_elbo_kernel(double* out) { out[0] = out[1] - out[2] - out[3]; }

// ------------------------------------------------------------------------------------------------
// Backward: Abar[g][m][n] = -2 A[m][n] sum_{l in g} dfvar[l][n] + sum_{v in g} V[v][m] dmean[v][n]
// and, in the same pass over A, dV[v][m] += sum_n A[g][m][n] dmean[v][n] (dV zero-initialised): the products
// are transposed through shared memory, reduced per row by one warp and added with one atomic per (row, vector)
// per CTA.   grid (Bc/128, Mp/32, Q), block 128
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) build_abar_kernel(const ModelDev* __restrict__ md, const double* __restrict__ A,
                                                         const double* __restrict__ V,
                                                         const double* __restrict__ dmean,
                                                         const double* __restrict__ dfvar, int Bp,
                                                         double* __restrict__ Abar, double* __restrict__ dV) {
  __shared__ double vs[32 * 33];
  __shared__ double pr[32][129];
  const int g = blockIdx.z, Mp = md->Mp, m0 = blockIdx.y * 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = blockIdx.x * 128 + tid;
  double dv = 0;
  for (int q = md->glat_begin[g]; q < md->glat_begin[g + 1]; q++) dv += dfvar[(long long)md->glat[q] * Bp + n];
  double a[32], acc[32];
  const double* Ag = A + ((long long)g * Mp + m0) * Bp + n;
#pragma unroll
  for (int r = 0; r < 32; r++) {
    a[r] = Ag[(long long)r * Bp];
    acc[r] = -2.0 * dv * a[r];
  }
  const int vb = md->gvec_begin[g], ve = md->gvec_begin[g + 1];
  for (int v0 = vb; v0 < ve; v0 += 32) {
    const int nv = min(32, ve - v0);
    __syncthreads();
    for (int t = tid; t < nv * 32; t += 128) {
      const int q = t / 32, r = t % 32;
      vs[q * 33 + r] = V[(long long)md->gvec[v0 + q] * Mp + m0 + r];
    }
    __syncthreads();
    for (int q = 0; q < nv; q++) {
      const int vec = md->gvec[v0 + q];
      const double dm = dmean[(long long)vec * Bp + n];
#pragma unroll
      for (int r = 0; r < 32; r++) {
        acc[r] += vs[q * 33 + r] * dm;
        pr[r][tid] = a[r] * dm;
      }
      __syncthreads();
#pragma unroll
      for (int rr = 0; rr < 8; rr++) {
        const int row = warp * 8 + rr;
        double sum = (pr[row][lane] + pr[row][lane + 32]) + (pr[row][lane + 64] + pr[row][lane + 96]);
        sum = warp_sum(sum);
        if (lane == 0) atomicAdd(dV + (long long)vec * Mp + m0 + row, sum);
      }
      __syncthreads();
    }
  }
  double* out = Abar + ((long long)g * Mp + m0) * Bp + n;
#pragma unroll
  for (int r = 0; r < 32; r++) out[(long long)r * Bp] = acc[r];
}

// grads.q_mu[l][m] += sum_r dV[v(l,r)][m];  grads.q_sqrt[l][m][k] += sum_s dV[v(l,s)][m] eps_u[l][s][k] (k <= m)
// for inducing-sampled latents.   grid (Mp/16, Mp/16, P), block (16,16)
__global__ void q_grads_kernel(const ModelDev* __restrict__ md, Layout lo, const double* __restrict__ dV,
                               const double* __restrict__ eps_u, int S, double* __restrict__ grads) {
  const int l = blockIdx.z, Mp = md->Mp, Mg = md->group_M[md->latent_group[l]];
  const int m = blockIdx.y * 16 + threadIdx.y, k = blockIdx.x * 16 + threadIdx.x;
  if (m >= Mg || k > m) return;
  const int R = md->R[l], v0 = md->voff[l];
  if (k == 0) {
    double s = 0;
    for (int r = 0; r < R; r++) s += dV[(long long)(v0 + r) * Mp + m];
    grads[lo.q_mu + (long long)l * Mp + m] += s;
  }
  if (!md->marg[l] && eps_u != nullptr) {
    double s = 0;
    for (int r = 0; r < R; r++) s += dV[(long long)(v0 + r) * Mp + m] * eps_u[((long long)l * S + r) * Mp + k];
    grads[lo.q_sqrt + ((long long)l * Mp + m) * Mp + k] += s;
  }
}

// ------------------------------------------------------------------------------------------------
// Kernel-hyperparameter and Z gradients from Kuf-bar (= W): E = W .* Kuf,
//   dZ[m,d]  = -sum_n E (z_md - x_nd) / l_d^2 ;  dl_d = sum_mn E (z_md - x_nd)^2 / l_d^3 ; dvar = sum E / var
// One warp per ROWS rows m (coalesced over n; x_n is loaded once for the ROWS rows).
// grid (Mp/(8*ROWS), Q), block 256.  D <= DMAX.
// ------------------------------------------------------------------------------------------------
// REGEN: Kuf is not read but re-evaluated from X and Z (the INT8-sliced step never materialises it in FP64): one FP64 exp
// per element in exchange for half of the kernel's HBM reads.
template <int DMAX, int ROWS, bool REGEN = false>
__global__ void __launch_bounds__(256) kuf_grad_kernel(const ModelDev* __restrict__ md, const double* __restrict__ params,
                                                       Layout lo, const double* __restrict__ X, int N, int Bp,
                                                       const double* __restrict__ W, const double* __restrict__ Kuf,
                                                       double* __restrict__ grads, int n_chunk) {
  const int g = blockIdx.y, Mp = md->Mp, D = md->D, Mg = md->group_M[g];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = (blockIdx.x * 8 + warp) * ROWS;
  if (m0 >= Mg) return;
  // blockIdx.z: slice of the data columns (small models have too few rows to fill the machine; every output is an
  // atomicAdd already, so slices simply add up)
  const int n_lo = blockIdx.z * n_chunk, n_hi = min(N, n_lo + n_chunk);
  const double* ls = params + lo.ls + (long long)g * D;
  double zz[ROWS][DMAX], a1[ROWS][DMAX], a2[ROWS][DMAX], se[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; r++) {
    se[r] = 0.0;
#pragma unroll
    for (int d = 0; d < DMAX; d++) {
      zz[r][d] = (d < D && m0 + r < Mg) ? params[lo.Z + ((long long)g * Mp + m0 + r) * D + d] : 0.0;
      a1[r][d] = a2[r][d] = 0.0;
    }
  }
  const double* w = W + ((long long)g * Mp + m0) * Bp;
  const double* kf = REGEN ? nullptr : Kuf + ((long long)g * Mp + m0) * Bp;
  double il[DMAX], zn[ROWS];  // REGEN: 1 / lengthscale and |z / l|^2 (the forward's expanded squared-distance form)
  const double kvar = params[lo.kvar + g];
  if (REGEN) {
#pragma unroll
    for (int d = 0; d < DMAX; d++) il[d] = d < D ? 1.0 / ls[d] : 0.0;
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
      zn[r] = 0.0;
#pragma unroll
      for (int d = 0; d < DMAX; d++) zn[r] += (zz[r][d] * il[d]) * (zz[r][d] * il[d]);
    }
  }
#pragma unroll(8 / ROWS)
  for (int n = n_lo + lane; n < n_hi; n += 32) {
    double x[DMAX], e[ROWS];
#pragma unroll
    for (int d = 0; d < DMAX; d++) x[d] = (d < D) ? X[(long long)n * D + d] : 0.0;
    if (REGEN) {
      double xn = 0.0;
#pragma unroll
      for (int d = 0; d < DMAX; d++) xn += (x[d] * il[d]) * (x[d] * il[d]);
#pragma unroll
      for (int r = 0; r < ROWS; r++) {
        double dot = 0.0;
#pragma unroll
        for (int d = 0; d < DMAX; d++) dot += (zz[r][d] * il[d]) * (x[d] * il[d]);
        const double r2 = -2.0 * dot + (zn[r] + xn);
        e[r] = (m0 + r < Mg) ? w[(long long)r * Bp + n] * (kvar * exp(-0.5 * r2)) : 0.0;
      }
    } else {
#pragma unroll
      for (int r = 0; r < ROWS; r++) e[r] = w[(long long)r * Bp + n] * kf[(long long)r * Bp + n];  // rows >= Mg: Kuf = 0
    }
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
      se[r] += e[r];
#pragma unroll
      for (int d = 0; d < DMAX; d++) {
        const double df = zz[r][d] - x[d];
        a1[r][d] += e[r] * df;
        a2[r][d] += e[r] * df * df;
      }
    }
  }
  double dkv = 0.0, dls[DMAX];
#pragma unroll
  for (int d = 0; d < DMAX; d++) dls[d] = 0.0;
#pragma unroll
  for (int r = 0; r < ROWS; r++) {
    dkv += warp_sum(se[r]);
#pragma unroll
    for (int d = 0; d < DMAX; d++)
      if (d < D) {
        const double s1 = warp_sum(a1[r][d]);
        dls[d] += warp_sum(a2[r][d]);
        if (lane == 0 && m0 + r < Mg) {
          const double l = ls[d];
          atomicAdd(grads + lo.Z + ((long long)g * Mp + m0 + r) * D + d, -s1 / (l * l));
        }
      }
  }
  if (lane == 0) {
    atomicAdd(grads + lo.kvar + g, dkv / params[lo.kvar + g]);
    for (int d = 0; d < D && d < DMAX; d++) {
      const double l = ls[d];
      atomicAdd(grads + lo.ls + (long long)g * D + d, dls[d] / (l * l * l));
    }
  }
}

// Same for Kuu with the symmetrised Kuu-bar = (Kb + Kb^T)/2 (Kb = L^-T Phi L^-1 from the Cholesky
// backward).  Both triangles of Kuu depend on Z, so row m collects the (m, m') and (m', m) terms.
template <int DMAX>
__global__ void __launch_bounds__(256) kuu_grad_kernel(const ModelDev* __restrict__ md, const double* __restrict__ params,
                                                       Layout lo, double jitter, const double* __restrict__ Kb,
                                                       const double* __restrict__ Kuu, double* __restrict__ grads) {
  const int g = blockIdx.y, Mp = md->Mp, D = md->D, Mg = md->group_M[g];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * 8 + warp;
  if (m >= Mg) return;
  const double* ls = params + lo.ls + (long long)g * D;
  const double* Zg = params + lo.Z + (long long)g * Mp * D;
  const double* kb = Kb + (long long)g * Mp * Mp;
  const double* ku = Kuu + (long long)g * Mp * Mp;
  double zz[DMAX], a1[DMAX], a2[DMAX];
#pragma unroll
  for (int d = 0; d < DMAX; d++) {
    zz[d] = d < D ? Zg[m * D + d] : 0.0;
    a1[d] = a2[d] = 0.0;
  }
  double se = 0;
  for (int j = lane; j < Mg; j += 32) {
    const double kbs = 0.5 * (kb[(long long)m * Mp + j] + kb[(long long)j * Mp + m]);
    const double kv = ku[(long long)m * Mp + j] - (j == m ? jitter : 0.0);
    const double e = kbs * kv;
    se += e;
#pragma unroll
    for (int d = 0; d < DMAX; d++)
      if (d < D) {
        const double df = zz[d] - Zg[j * D + d];
        a1[d] += e * df;
        a2[d] += e * df * df;
      }
  }
  se = warp_sum(se);
  if (lane == 0) atomicAdd(grads + lo.kvar + g, se / params[lo.kvar + g]);
#pragma unroll
  for (int d = 0; d < DMAX; d++)
    if (d < D) {
      const double s1 = warp_sum(a1[d]), s2 = warp_sum(a2[d]);
      if (lane == 0) {
        const double l = ls[d];
        atomicAdd(grads + lo.Z + ((long long)g * Mp + m) * D + d, -2.0 * s1 / (l * l));
        atomicAdd(grads + lo.ls + (long long)g * D + d, s2 / (l * l * l));
      }
    }
}

// Phi: halve the diagonal of a (already lower-triangular) matrix batch.  grid (Q), block 256
__global__ void halve_diag_kernel(double* __restrict__ P, int Mp) {
  double* p = P + (long long)blockIdx.x * Mp * Mp;
  for (int i = threadIdx.x; i < Mp; i += blockDim.x) p[(long long)i * Mp + i] *= 0.5;
}

// Minibatch assembly on the device (reference: tf.data shuffle + batch, mogpe/training/utils.py:105-116):
// dst[i, :] = src[idx[i], :].   grid (ceil(n*C/256)), block 256
__global__ void gather_rows_kernel(const double* __restrict__ src, const long long* __restrict__ idx, long long n, int C,
                                   double* __restrict__ dst) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * C) return;
  const long long i = t / C;
  const int c = (int)(t % C);
  dst[t] = src[idx[i] * C + c];
}

// ------------------------------------------------------------------------------------------------
// Device-side optimiser step: the reference's train_step applies tf.optimizers.Adam to the
// UNCONSTRAINED variables (mosvgpe.py:57-69).  `u` holds them in the packed layout; positive parameters are
// softplus-transformed (gpflow.utilities.positive; noise: + lower bound), q_sqrt's FillTriangular is a pure
// permutation, so Adam (elementwise) acts on the lower-triangular entries in place.
// var_rep[i] = index of the representative entry of the variable that packed entry i belongs to (-1: not a
// variable).  Tied entries (a scalar lengthscale broadcast over D, an expert object used twice, ...) sum their
// gradients into the representative, which is updated and broadcast back.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int param_kind(long long i, Layout lo) {  // 0 identity, 1 softplus, 2 softplus + noise lower bound
  if (i < lo.Z) return 1;          // lengthscales, kernel variances
  if (i >= lo.noise) return 2;     // Gaussian likelihood variances
  return 0;                        // Z, q_mu, q_sqrt (lower triangle), mean_c
}

// var_rep encoding: r >= 0 representative index; -1 not a variable; r < -1: representative -2 - r AND the entry is a positive
// (softplus) parameter whatever its position -- the diagonal of a q_diag=True q_sqrt, which gpflow constrains with
// positive() where the full q_sqrt is an unconstrained FillTriangular.
__host__ __device__ __forceinline__ int var_rep_index(int v) { return v >= -1 ? v : -2 - v; }
__host__ __device__ __forceinline__ bool var_rep_softplus(int v) { return v < -1; }

// gu[i] = dELBO/du_i (chain through the bijector); tied entries are accumulated into their representative
__global__ void opt_chain_kernel(long long n, Layout lo, const double* __restrict__ u, const double* __restrict__ g,
                                 const int* __restrict__ var_rep, double* __restrict__ gu) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int rep = var_rep_index(var_rep[i]);
  if (rep < 0) return;
  double v = g[i];
  if (param_kind(i, lo) != 0 || var_rep_softplus(var_rep[i])) v *= 1.0 / (1.0 + exp(-u[i]));  // d softplus / du = sigmoid(u)
  if (rep == i) atomicAdd(gu + i, v);
  else atomicAdd(gu + rep, v);
}

// Adam (TF semantics: lr_t = lr sqrt(1 - b2^t) / (1 - b1^t); epsilon outside the sqrt) on loss = -ELBO
// lr_t_dev (optional): the bias-corrected step size computed on the device by train_step_prep_kernel (graph replay).
__global__ void opt_adam_kernel(long long n, const int* __restrict__ var_rep, const double* __restrict__ gu,
                                double* __restrict__ u, double* __restrict__ m, double* __restrict__ v, double lr_t,
                                double b1, double b2, double eps, const double* __restrict__ lr_t_dev) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || var_rep_index(var_rep[i]) != i) return;
  if (lr_t_dev) lr_t = *lr_t_dev;
  const double gl = -gu[i];
  const double mi = m[i] + (1.0 - b1) * (gl - m[i]);
  const double vi = v[i] + (1.0 - b2) * (gl * gl - v[i]);
  m[i] = mi;
  v[i] = vi;
  u[i] -= lr_t * mi / (sqrt(vi) + eps);
}

// Graph-replayed training step: advance the device-side random-stream position and Adam step count, and derive the
// bias-corrected step size.  state = {rng position, adam t}; one thread.
__global__ void train_step_prep_kernel(unsigned long long* __restrict__ state, unsigned long long rng_consumed, double lr,
                                       double b1, double b2, double* __restrict__ lr_t_dev) {
  state[0] += rng_consumed;
  const unsigned long long t = ++state[1];
  *lr_t_dev = lr * sqrt(1.0 - pow(b2, (double)t)) / (1.0 - pow(b1, (double)t));
}

// params[i] = T(u[rep(i)]) (and u of tied entries follows its representative); non-variables are left untouched
__global__ void opt_transform_kernel(long long n, Layout lo, const int* __restrict__ var_rep, double* __restrict__ u,
                                     double noise_lower, double* __restrict__ params) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int rep = var_rep_index(var_rep[i]);
  if (rep < 0) return;
  const double ui = u[rep];
  if (rep != i) u[i] = ui;
  const int kind = var_rep_softplus(var_rep[i]) ? 1 : param_kind(i, lo);
  double p = ui;
  if (kind != 0) p = (ui > 0 ? ui : 0.0) + log1p(exp(-fabs(ui))) + (kind == 2 ? noise_lower : 0.0);  // stable softplus
  params[i] = p;
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 + Box-Muller standard normals (counter-based: element i depends only on seed, offset+i)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

// `base` (optional, device memory) is added to `offset`: launches captured in a CUDA graph keep the stream position on
// the device so that replaying the graph continues the random stream (mogpe_train_step).
__global__ void fill_normal_kernel(double* __restrict__ dst, long long n, unsigned long long seed,
                                   unsigned long long offset, const unsigned long long* __restrict__ base) {
  const long long pair = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long i = pair * 2;
  if (i >= n) return;
  if (base) offset += *base;
  const unsigned long long ctr = offset / 2 + pair;
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0x6d6f6770u, 0x65623230u};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; r++) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  // two 53-bit uniforms in (0, 1]
  const double u1 = ((double)(((unsigned long long)c[0] << 21) ^ (c[1] >> 11)) + 1.0) * (1.0 / 9007199254740992.0);
  const double u2 = ((double)(((unsigned long long)c[2] << 21) ^ (c[3] >> 11)) + 1.0) * (1.0 / 9007199254740992.0);
  const double rad = sqrt(-2.0 * log(u1));
  double s, co;
  sincospi(2.0 * u2, &s, &co);
  dst[i] = rad * co;
  if (i + 1 < n) dst[i + 1] = rad * s;
}

}  // namespace mogpe

namespace mogpe {
// Library-generated Softmax Monte-Carlo draws for one chunk of test points (gpflow Softmax.predict_mean_and_var draws
// from the global RNG): dst[r][i][k], i in [0, nc), is element offset + ((r * Ntot) + n0 + i) * K + k of the Philox stream,
// so the result does not depend on how the points are chunked.   grid (ceil(nc*K/256), R), block 256
__global__ void fill_normal_mc_kernel(double* __restrict__ dst, int nc, int K, long long Ntot, long long n0,
                                      unsigned long long seed, unsigned long long offset) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (j >= (long long)nc * K) return;
  const unsigned long long e = offset + ((unsigned long long)r * Ntot + n0) * K + j;
  const unsigned long long ctr = e >> 1;
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0x6d6f6770u, 0x65623230u};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int q = 0; q < 10; q++) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  const double u1 = ((double)(((unsigned long long)c[0] << 21) ^ (c[1] >> 11)) + 1.0) * (1.0 / 9007199254740992.0);
  const double u2 = ((double)(((unsigned long long)c[2] << 21) ^ (c[3] >> 11)) + 1.0) * (1.0 / 9007199254740992.0);
  const double rad = sqrt(-2.0 * log(u1));
  double sn, co;
  sincospi(2.0 * u2, &sn, &co);
  dst[(long long)r * nc * K + j] = rad * ((e & 1) ? sn : co);
}
}  // namespace mogpe
