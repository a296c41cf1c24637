// Streaming kernels of the INT8-sliced TRAINING step (gemm_i8.cuh products inside mogpe_elbo_fwd_bwd): operand bounds,
// the fused build-Abar + transpose + slice pass, and row-maximum -> scale conversion.
//
// Every FP64 operand of a sliced product is written as 2^e * sum_p q_p 2^(-7 (p + 1)) with |x| <= 2^(e-1).  Where the scale
// 2^e of an operand cannot be measured without an extra pass over it, a rigorous a-priori BOUND is used instead: eight byte
// slices carry 56 bits, the tolerance of the path needs ~35, so a bound that is loose by a factor 2^5 .. 2^10 costs nothing.
#pragma once
#include "gemm_i8.cuh"
#include "model_dev.cuh"

namespace mogpe {
namespace oz {

__device__ __forceinline__ void ld_global_256(const double* p, double (&v)[4]) {
  asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}

// Bounds of the sliced operands that are known before the products run (S = tril(q_sqrt), V = mean vectors):
//   cmax2[slot] = max_m |S[:, m]|_2^2,  rmax2[slot] = max_m |S[m, :]|_2^2   (per marginal-latent slot),
//   vmax[v]     = max_m |V[v][m]|                                           (per mean vector),
// accumulated with atomicMax on the bit patterns of non-negative doubles (the buffers are zeroed by the caller).
// Block (x, slot) sweeps the 32-column block x of S (column norms, rows >= 32 x) and the 32-row slab x (row norms).
// grid (Mp/32, nlta + NV), block 256
__global__ void __launch_bounds__(256) bounds_kernel(const ModelDev* __restrict__ md, const double* __restrict__ Sc,
                                                     const double* __restrict__ V, unsigned long long* __restrict__ cmax2,
                                                     unsigned long long* __restrict__ rmax2,
                                                     unsigned long long* __restrict__ vmax_bits) {
  __shared__ double colp[8][33];
  const int Mp = md->Mp, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, cb = blockIdx.x * 32;
  if ((int)blockIdx.y < md->nlta) {
    const int slot = blockIdx.y, l = md->lta_latent[slot];
    const double* S = Sc + (long long)l * Mp * Mp;
    double cs = 0.0;
    for (int r = cb + warp; r < Mp; r += 8) {
      const double x = S[(long long)r * Mp + cb + lane];  // exact zeros above the diagonal (clean copy)
      cs += x * x;
    }
    colp[warp][lane] = cs;
    double rmax = 0.0;
    for (int r = cb + warp; r < cb + 32; r += 8) {
      double s2 = 0.0;
      for (int c = lane; c <= r; c += 32) {
        const double x = S[(long long)r * Mp + c];
        s2 += x * x;
      }
      rmax = fmax(rmax, warp_sum(s2));
    }
    __syncthreads();
    if (warp == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; w++) t += colp[w][lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t = fmax(t, __shfl_xor_sync(0xffffffffu, t, o));
      if (lane == 0) atomicMax(cmax2 + slot, (unsigned long long)__double_as_longlong(t));
    }
    if (lane == 0) atomicMax(rmax2 + slot, (unsigned long long)__double_as_longlong(rmax));
  } else {
    const int v = blockIdx.y - md->nlta;
    if (tid < 32) {
      double m = fabs(V[(long long)v * Mp + cb + tid]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (tid == 0) atomicMax(vmax_bits + v, (unsigned long long)__double_as_longlong(m));
    }
  }
}

// lta_scale[slot] = pow2_scale(max_m |S[:, m]|_2 sqrt(kvar_g))  (|LTA[m][n]| = |S[:, m] . a_n| <= |S[:, m]|_2 |a_n|_2 and
// |a_n|_2^2 = sum_m A[m][n]^2 <= kvar);  srowmax[latent] = max_m |S[m, :]|_2  (bound of S LTA);  vmax[v] as doubles.
// one block, 256 threads
__global__ void bounds_finish_kernel(const ModelDev* __restrict__ md, const double* __restrict__ params, Layout lo,
                                     const unsigned long long* __restrict__ cmax2, const unsigned long long* __restrict__ rmax2,
                                     const unsigned long long* __restrict__ vmax_bits, double* __restrict__ lta_scale,
                                     double* __restrict__ srowmax, double* __restrict__ vmax) {
  for (int i = threadIdx.x; i < md->nlta; i += blockDim.x) {
    const int l = md->lta_latent[i], g = md->latent_group[l];
    lta_scale[i] = pow2_scale(1.0001 * sqrt(__longlong_as_double((long long)cmax2[i])) * sqrt(params[lo.kvar + g]));
    srowmax[l] = 1.0001 * sqrt(__longlong_as_double((long long)rmax2[i]));
  }
  for (int i = threadIdx.x; i < md->NV; i += blockDim.x) vmax[i] = __longlong_as_double((long long)vmax_bits[i]);
}

// out_bits[b] = max_n |v[b][n]| (n < N) as bit patterns (zero-initialised by the caller).  grid (ceil(N/1024), batch), block 256
__global__ void __launch_bounds__(256) absmax_rows_kernel(const double* __restrict__ v, int N, long long stride,
                                                          unsigned long long* __restrict__ out_bits) {
  const int b = blockIdx.y;
  double m = 0.0;
  for (int n = blockIdx.x * 1024 + threadIdx.x; n < min(N, (int)(blockIdx.x + 1) * 1024); n += 256)
    m = fmax(m, fabs(v[(long long)b * stride + n]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits + b, (unsigned long long)__double_as_longlong(m));
}

// Row scales of LTA_l diag(dfvar_l) from a bound instead of a pass over the matrix:
//   scale[slot][m] = pow2_scale(max_n |LTA[slot][m][n]| * max_n |dfvar[latent(slot)][n]|)
// (row maxima of LTA collected by the epilogue of the LTA product, the adjoint maxima by absmax_rows_kernel).
__global__ void lta_d_scale_kernel(const ModelDev* __restrict__ md, const unsigned long long* __restrict__ lta_rowmax,
                                   const unsigned long long* __restrict__ dmax, int Mp, double* __restrict__ scale) {
  const int slot = blockIdx.y, m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= Mp) return;
  const double d = __longlong_as_double((long long)dmax[md->lta_latent[slot]]);
  scale[(long long)slot * Mp + m] = pow2_scale(1.0000001 * __longlong_as_double((long long)lta_rowmax[(long long)slot * Mp + m]) * d);
}

// scale[i] = pow2_scale(max[i]) from the bit patterns accumulated by atomicMax (non-negative doubles order like integers)
__global__ void max_to_scale_kernel(const unsigned long long* __restrict__ mx, double* __restrict__ scale, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) scale[i] = pow2_scale(__longlong_as_double((long long)mx[i]));
}

// Backward, fused: Abar (never materialised in FP64) -> transposed byte planes, the B operand of W = L^-T Abar.
//   Abar[g][m][n] = -2 A[m][n] sum_{l in g} dfvar[l][n] + sum_{v in g} V[v][m] dmean[v][n]
//                   + sum_{marginal l in g} 2 dfvar[l][n] (S_l LTA_l)[m][n]                 (P2[slot(l)] = S_l LTA_l)
//   dV[v][m]     += sum_n A[g][m][n] dmean[v][n]                                            (dV zero-initialised)
// A comes from its [m][n] byte planes (written by the epilogue of the A product; 2^e = a_scale[g]); the result is sliced
// against a per-COLUMN scale from a rigorous bound (|A[m][n]| <= sqrt(var0ss[n]), |P2[m][n]| <= max_m |S[m, :]|_2 sqrt(varq[n]),
// |V[v][m]| <= vmax[v]) -- a column scale commutes with the left multiplication by L^-T (gemm_i8 b_row_scale) -- and goes out
// as AbarT8[g][plane][m / 64][n][64] (the blocked plane layout of gemm_i8.cuh, like the A planes it reads) through a
// shared-memory transpose.        grid (Bc/64, Mp/64, Q), block 256
// sign-extended byte J of w
template <int J>
__device__ __forceinline__ int sbyte(uint32_t w) {
  return (int)(w << (24 - 8 * J)) >> 24;
}

// value of element J (0..3) of four consecutive columns from the NS plane words w[s] (byte J of every word), in units of
// the operand scale:  sum_s d_s 2^(-7 (s + 1)),  exact (two integers of 28 bits, combined in FP64)
template <int NS, int J>
__device__ __forceinline__ double unslice(const uint32_t (&w)[NS]) {
  constexpr int H = NS / 2;
  int hi = sbyte<J>(w[0]), lo = sbyte<J>(w[H]);
#pragma unroll
  for (int k = 1; k < H; k++) {
    hi = hi * 128 + sbyte<J>(w[k]);
    lo = lo * 128 + sbyte<J>(w[H + k]);
  }
  constexpr double P1 = 1.0 / (double)(1 << (7 * H));
  return fma(int_to_double(lo), P1, int_to_double(hi)) * P1;
}

// The same for all four columns at once, with byte-wise SIMD: flipping the sign bits turns the signed digits into unsigned bytes
// d + 128, a 4 x 4 byte transpose gathers the digits of one element into one word, two DP4A with the weights (128, 1) pair
// them up and one IMAD joins the pairs; the bias 128 * sum 128^k rides in the accumulator input of the second DP4A.
// 12 integer instructions per element instead of 22.
template <int NS>
__device__ __forceinline__ void unslice4(const uint32_t (&w)[NS], double (&a)[4]) {
  constexpr int H = NS / 2;
  static_assert(H == 3 || H == 4, "six or eight slices");
  constexpr unsigned BIAS = H == 4 ? 128u * 2113665u : 128u * 16513u;
  int half[2][4];
#pragma unroll
  for (int h2 = 0; h2 < 2; h2++) {
    uint32_t t[4];
    transpose4x4_bytes(w[h2 * H] ^ 0x80808080u, w[h2 * H + 1] ^ 0x80808080u, w[h2 * H + 2] ^ 0x80808080u,
                       H == 4 ? (w[h2 * H + H - 1] ^ 0x80808080u) : 0u, t);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const unsigned top = __dp4a(t[j], 0x00000180u, 0u);  // u0 * 128 + u1
      const unsigned low = __dp4a(t[j], H == 4 ? 0x01800000u : 0x00010000u, 0u - BIAS);  // u2 * 128 + u3 (or u2) - bias
      half[h2][j] = (int)(top * (H == 4 ? 16384u : 128u) + low);
    }
  }
  constexpr double P1 = 1.0 / (double)(1 << (7 * H));
#pragma unroll
  for (int j = 0; j < 4; j++) a[j] = fma(int_to_double(half[1][j]), P1, int_to_double(half[0][j])) * P1;
}

// (ISL: integer slicing, slice_words_int -- measured same-box: 462 us against 417 with the FP64 slicing here, so off)
template <int NS, bool ISL = false>
__global__ void __launch_bounds__(256) abar8_kernel(const ModelDev* __restrict__ md, const int8_t* __restrict__ A8,
                                                    const double* __restrict__ a_scale, const double* __restrict__ V,
                                                    const double* __restrict__ dmean, const double* __restrict__ dfvar,
                                                    const double* __restrict__ var0ss, const double* __restrict__ varq,
                                                    const double* __restrict__ P2, const double* __restrict__ vmax,
                                                    const double* __restrict__ srowmax, int N, int Bp,
                                                    int8_t* __restrict__ AbarT8, double* __restrict__ nscale,
                                                    double* __restrict__ dV) {
  constexpr int PITCH = 65;  // words per [n] row of the staging tile (64 m + 1: column groups of a warp spread over the banks)
  __shared__ uint32_t stage[2][64][PITCH];  // [slice half][n][m]: the four digit bytes of planes 4 half .. 4 half + 3
  __shared__ double s_inv[64], s_dv[64];
  __shared__ int s_ie[64];  // exponent of 1 / scale (integer slicing)
  const int g = blockIdx.z, Mp = md->Mp, m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tid = threadIdx.x;
  const int lb = md->glat_begin[g], le = md->glat_begin[g + 1], vb = md->gvec_begin[g], ve = md->gvec_begin[g + 1];
  if (tid < 64) {
    const int n = n0 + tid;
    double dv = 0.0, b = 0.0;
    if (n < N) {
      for (int q = lb; q < le; q++) {
        const int l = md->glat[q];
        const double d = dfvar[(long long)l * Bp + n];
        dv += d;
        if (md->marg[l]) b += 2.0 * fabs(d) * srowmax[l] * sqrt(fmax(varq[(long long)l * Bp + n], 0.0));
      }
      b += 2.0 * fabs(dv) * sqrt(fmax(var0ss[(long long)g * Bp + n], 0.0));
      for (int q = vb; q < ve; q++) {
        const int v = md->gvec[q];
        b += vmax[v] * fabs(dmean[(long long)v * Bp + n]);
      }
    }
    const double sc = pow2_scale(1.001 * b);
    // column n = 4 i + j lives at [16 j + i]: the 16 lanes of a tile row read consecutive words (no bank conflicts)
    s_inv[(tid & 3) * 16 + (tid >> 2)] = 1.0 / sc;
    s_ie[(tid & 3) * 16 + (tid >> 2)] = 1023 - ((__double2hiint(sc) >> 20) & 0x7ff);
    s_dv[(tid & 3) * 16 + (tid >> 2)] = dv;
    if (blockIdx.y == 0) nscale[(long long)g * Bp + n] = sc;
  }
  __syncthreads();
  const int c4 = (tid & 15) * 4, r = tid >> 4;
  const long long plane = (long long)Mp * Bp;
  const int8_t* Ag = A8 + (long long)g * NS * plane;
  const double asc = a_scale[g];
  // Columns in [N, Bc) need no special case: their A / LTA planes are zero (zero Kuf^T rows) and the likelihood kernel wrote
  // zero adjoints there, so every term below vanishes.
  const int nn = n0 + c4;
#pragma unroll 2
  for (int pass = 0; pass < 4; pass++) {
    const int ml = pass * 16 + r, m = m0 + ml;
    uint32_t w[NS];
#pragma unroll
    // (A planes in the blocked layout [n / 64][m][64]: the 64 x 64 tile of a plane is one contiguous 4 KB run)
    for (int s2 = 0; s2 < NS; s2++)
      w[s2] = *reinterpret_cast<const uint32_t*>(Ag + s2 * plane + ((long long)blockIdx.x * Mp + m) * 64 + c4);
    double a[4];
    unslice4<NS>(w, a);
#pragma unroll
    for (int j = 0; j < 4; j++) a[j] *= asc;
    double ab[4];
#pragma unroll
    for (int j = 0; j < 4; j++) ab[j] = -2.0 * s_dv[j * 16 + (tid & 15)] * a[j];
    for (int q = vb; q < ve; q++) {
      const int v = md->gvec[q];
      const double vm = V[(long long)v * Mp + m];
      double dm[4];
      ld_global_256(dmean + (long long)v * Bp + nn, dm);
      double part = 0.0;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        ab[j] += vm * dm[j];
        part += a[j] * dm[j];
      }
      // the 16 threads of one tile row are 16 consecutive lanes
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      part += __shfl_xor_sync(0xffffffffu, part, 4);
      part += __shfl_xor_sync(0xffffffffu, part, 8);
      if ((tid & 15) == 0) atomicAdd(dV + (long long)v * Mp + m, part);
    }
    for (int q = lb; q < le; q++) {
      const int l = md->glat[q];
      if (!md->marg[l]) continue;
      double p2[4], dl[4];
      ld_global_256(P2 + ((long long)md->lta_slot[l] * Mp + m) * Bp + nn, p2);
      ld_global_256(dfvar + (long long)l * Bp + nn, dl);
#pragma unroll
      for (int j = 0; j < 4; j++) ab[j] += 2.0 * dl[j] * p2[j];
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      uint32_t wh, wl;
      if constexpr (ISL) slice_words_int<NS>(ab[j], s_ie[j * 16 + (tid & 15)], wh, wl);
      else slice_words<NS>(ab[j] * s_inv[j * 16 + (tid & 15)], wh, wl);
      stage[0][c4 + j][ml] = wh;
      stage[1][c4 + j][ml] = wl;
    }
  }
  __syncthreads();
  // read back 16 consecutive m of one (half, n), transpose the digit bytes into plane words and store 16 bytes per plane
  int8_t* Og = AbarT8 + (long long)g * NS * Bp * Mp;
  // (a warp covers 16 columns x 2 groups of 16 m: with the odd pitch the 32 row starts fall into 32 different banks)
  for (int u = tid; u < 2 * 64 * 4; u += 256) {
    const int half = u >> 8, wq = (u >> 5) & 7, l = u & 31;
    const int n = (l >> 1) + 16 * (wq >> 1), mg = (l & 1) + 2 * (wq & 1);
    const uint32_t* sp = &stage[half][n][mg * 16];
    uint32_t o4[4][4];
#pragma unroll
    for (int k4 = 0; k4 < 4; k4++) {
      uint32_t tr[4];
      transpose4x4_bytes(sp[4 * k4], sp[4 * k4 + 1], sp[4 * k4 + 2], sp[4 * k4 + 3], tr);
#pragma unroll
      for (int b4 = 0; b4 < 4; b4++) o4[b4][k4] = tr[b4];
    }
    // (blocked over m: [m / 64][n][64] -- the tile is contiguous again)
    int8_t* ob = Og + ((long long)blockIdx.y * Bp + n0 + n) * 64 + mg * 16;
#pragma unroll
    for (int b4 = 4 - NS / 2; b4 < 4; b4++) {  // byte b4 of a slice word = digit 3 - b4 of its half
      const int pl = half * (NS / 2) + 3 - b4;
      *reinterpret_cast<uint4*>(ob + (long long)pl * Bp * Mp) = make_uint4(o4[b4][0], o4[b4][1], o4[b4][2], o4[b4][3]);
    }
  }
}

}  // namespace oz
}  // namespace mogpe

namespace mogpe {
namespace oz {

// exp(x) for x <= 0 to ~1e-14 relative (17 FP64 instructions instead of the ~40 of the correctly rounded library routine):
// k = round(x log2 e) by the 2^52 trick, f = x - k ln 2 in two pieces (|f| <= 0.347), a degree-11 Taylor polynomial, and
// 2^k assembled in the exponent field (flushed to zero below the normal range).  Used ONLY where Kuf is re-evaluated for the
// gradient kernels (tolerance 1e-6); the forward Kuf keeps the library exp.
__device__ __forceinline__ double exp_neg_fast(double x) {
  const double C = 6755399441055744.0;  // 1.5 * 2^52
  const double t = fma(x, 1.4426950408889634074, C);
  const int k = __double2loint(t);
  const double kd = t - C;
  double f = fma(kd, -6.93147180369123816490e-01, x);  // ln2 high part (gpflow-independent constants of fdlibm)
  f = fma(kd, -1.90821492927058770002e-10, f);         // ln2 low part
  double p = 1.0 / 39916800.0;
  p = fma(p, f, 1.0 / 3628800.0);
  p = fma(p, f, 1.0 / 362880.0);
  p = fma(p, f, 1.0 / 40320.0);
  p = fma(p, f, 1.0 / 5040.0);
  p = fma(p, f, 1.0 / 720.0);
  p = fma(p, f, 1.0 / 120.0);
  p = fma(p, f, 1.0 / 24.0);
  p = fma(p, f, 1.0 / 6.0);
  p = fma(p, f, 0.5);
  p = fma(p, f, 1.0);
  p = fma(p, f, 1.0);
  const double s = k > -1022 ? __hiloint2double((k + 1023) << 20, 0) : 0.0;
  return p * s;
}

// Inputs of a kernel group in ITS length scales, once per step:  xs[g][n][d] = x_nd / l_gd  (zero padded to DMAX; columns
// n >= N get 1e150 in d = 0, which drives their kernel value to exactly zero in kuf_grad_slice_kernel).   grid (Bp/256, Q)
template <int DMAX>
__global__ void xs_prep_kernel(const ModelDev* __restrict__ md, const double* __restrict__ params, Layout lo,
                               const double* __restrict__ X, int N, int Bp, double* __restrict__ xs) {
  const int g = blockIdx.y, D = md->D;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= Bp) return;
  const double* ls = params + lo.ls + (long long)g * D;
  double* o = xs + ((long long)g * Bp + n) * DMAX;
#pragma unroll
  for (int d = 0; d < DMAX; d++) o[d] = (n < N) ? (d < D ? X[(long long)n * D + d] / ls[d] : 0.0) : (d == 0 ? 1e150 : 0.0);
}

// Kuf-bar (= W) consumer of the INT8 step, ONE pass over W [g][m][n]:
//   (a) kernel-hyperparameter and Z gradients  E = W .* Kuf,  dZ[m,d] = -sum_n E (z_md - x_nd) / l_d^2,
//       dl_d = sum_mn E (z_md - x_nd)^2 / l_d^3,  dvar = sum E / var   (as kuf_grad_kernel) with Kuf RE-EVALUATED from the scaled
//       inputs of xs_prep_kernel and Z (the step holds it only as byte planes; one FP64 exp per element instead of a second
//       8-byte read): with u = (z - x) / l,  r2 = |u|^2 directly (no |x|^2 + |z|^2 - 2 x.z cancellation), and the same u give
//       the gradient sums  sum E u  (-> dZ = -. / l)  and  sum E u^2  (-> dl = . / l);
//   (b) the [m][n] byte planes of W for Lbar = -tril(W A^T), sliced against the exact row maxima that the epilogue of the
//       W product collected (wmax, bit patterns) -- no separate slicing pass over W.
// One warp per row m, every lane four consecutive columns per step (128 columns per warp step: 1 KB of W, 128 B per byte
// plane).  n_chunk must be a multiple of 128.       grid (Mp/8, Q, nsplit), block 256
// ISL: W is re-sliced with integer operations (slice_words_int): 455 -> 436 us same-box -- this kernel is bound by its FP64
// instructions, and the slicing is the part of them that integer units can take over.
template <int DMAX, int NS, bool ISL = true>
__global__ void __launch_bounds__(256) kuf_grad_slice_kernel(const ModelDev* __restrict__ md, const double* __restrict__ params,
                                                             Layout lo, const double* __restrict__ XS, int Bc, int Bp,
                                                             const double* __restrict__ W,
                                                             const unsigned long long* __restrict__ wmax,
                                                             double* __restrict__ grads, int8_t* __restrict__ W8,
                                                             double* __restrict__ w_scale, int n_chunk) {
  const int g = blockIdx.y, Mp = md->Mp, D = md->D, Mg = md->group_M[g];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * 8 + warp;
  const int n_lo = blockIdx.z * n_chunk, n_hi = min(Bc, n_lo + n_chunk);
  const double* ls = params + lo.ls + (long long)g * D;
  const double kvar = params[lo.kvar + g];
  const bool live = m < Mg;  // padded rows: no gradient (sums discarded below), W planes are still written (zeros in, zeros out)
  double zs[DMAX], a1[DMAX], a2[DMAX], se = 0.0;
#pragma unroll
  for (int d = 0; d < DMAX; d++) {
    zs[d] = (d < D && live) ? params[lo.Z + ((long long)g * Mp + m) * D + d] / ls[d] : 0.0;
    a1[d] = a2[d] = 0.0;
  }
  const double sc = pow2_scale(__longlong_as_double((long long)wmax[(long long)g * Mp + m]));
  const double inv = 1.0 / sc;
  const int inv_e = 1023 - ((__double2hiint(sc) >> 20) & 0x7ff);  // sc = 2^e: only the exponent is needed by the integer slicing
  if (lane == 0 && blockIdx.z == 0) w_scale[(long long)g * Mp + m] = sc;
  const double* w = W + ((long long)g * Mp + m) * Bp;
  const double* xg = XS + (long long)g * Bp * DMAX;
  int8_t* o8 = W8 + ((long long)g * NS * Mp + m) * Bp;
  const long long plane = (long long)Mp * Bp;
  for (int n4 = n_lo + lane * 4; n4 < n_hi; n4 += 128) {
    double wv[4];
    ld_global_256(w + n4, wv);
    uint32_t wh[4], wl[4];
    // branch-free over the four columns (independent exp chains in flight); padding columns have r2 ~ 1e300 -> E = 0
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if constexpr (ISL) slice_words_int<NS>(wv[j], inv_e, wh[j], wl[j]);
      else slice_words<NS>(wv[j] * inv, wh[j], wl[j]);
      double u[DMAX];
#pragma unroll
      for (int d4 = 0; d4 < DMAX; d4 += 4) ld_global_256(xg + (long long)(n4 + j) * DMAX + d4, *reinterpret_cast<double(*)[4]>(&u[d4]));
      double r2 = 0.0;
#pragma unroll
      for (int d = 0; d < DMAX; d++) {
        u[d] = zs[d] - u[d];
        r2 = fma(u[d], u[d], r2);
      }
      const double ev = wv[j] * (kvar * exp_neg_fast(fmax(-0.5 * r2, -1000.0)));
      se += ev;
#pragma unroll
      for (int d = 0; d < DMAX; d++) {
        const double t = ev * u[d];
        a1[d] += t;
        a2[d] = fma(t, u[d], a2[d]);
      }
    }
    uint32_t th[4], tl[4];
    transpose4x4_bytes(wh[0], wh[1], wh[2], wh[3], th);
    transpose4x4_bytes(wl[0], wl[1], wl[2], wl[3], tl);
#pragma unroll
    for (int b4 = 4 - NS / 2; b4 < 4; b4++) {  // byte b4 of a slice word = digit 3 - b4 of its half
      *reinterpret_cast<uint32_t*>(o8 + (long long)(3 - b4) * plane + n4) = th[b4];
      *reinterpret_cast<uint32_t*>(o8 + (long long)(NS / 2 + 3 - b4) * plane + n4) = tl[b4];
    }
  }
  if (!live) return;
  se = warp_sum(se);
  double dls[DMAX];
#pragma unroll
  for (int d = 0; d < DMAX; d++) {
    dls[d] = 0.0;
    if (d < D) {
      const double s1 = warp_sum(a1[d]);
      dls[d] = warp_sum(a2[d]);
      if (lane == 0) atomicAdd(grads + lo.Z + ((long long)g * Mp + m) * D + d, -s1 / ls[d]);
    }
  }
  if (lane == 0) {
    atomicAdd(grads + lo.kvar + g, se / kvar);
    for (int d = 0; d < D && d < DMAX; d++) atomicAdd(grads + lo.ls + (long long)g * D + d, dls[d] / ls[d]);
  }
}

}  // namespace oz
}  // namespace mogpe
