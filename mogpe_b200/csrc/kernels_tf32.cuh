// Streaming kernels of the TF32 mode (see gemm_tf32.cuh): operand preparation for the tcgen05 products.
// The kernel matrices are still EVALUATED in FP64 (gpflow's expanded squared-distance form) and the
// predictive means are FP64 dot products  mean[n] = sum_k Kuf[k][n] alpha[k],  alpha = L^-T v  (no M x M x B product
// needed for the mean); only the two variance products run on the tensor cores in split FP32 planes.
#pragma once
#include "gemm_tf32.cuh"
#include "model_dev.cuh"

namespace mogpe {
namespace t32 {

// dst [batch][2][n_p][n_p] <- src[batch][n][n]^T (zero padded): dst(r, c) = src(c, r).  grid (n_p/32, n_p/32, batch),
// block (32, 8); 32x32 tiles through shared memory so that both sides are coalesced.  src batch via `map` (slot -> index).
__global__ void split_planes_T_kernel(const double* __restrict__ src, int n, long long src_batch_stride, int src_ld,
                                      const int* __restrict__ map, float* __restrict__ dst, int n_p) {
  __shared__ double tile[32][33];
  const int b = blockIdx.z, sb = map ? map[b] : b;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;  // dst tile rows r0.., cols c0..  <- src rows c0.., cols r0..
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int sr = c0 + j, sc = r0 + threadIdx.x;
    tile[j][threadIdx.x] = (sr < n && sc < n) ? src[(long long)sb * src_batch_stride + (long long)sr * src_ld + sc] : 0.0;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    const double x = tile[threadIdx.x][j];
    const float hi = tf32_hi((float)x);
    float* d = dst + ((long long)b * 2 * n_p + r) * n_p + c;
    d[0] = hi;
    d[(long long)n_p * n_p] = (float)(x - (double)hi);
  }
}

// alpha[v][k] = sum_{m >= k} Linv[g(v)][m][k] V[v][m]   (alpha = L^-T v; Linv lower-triangular), zero for k >= Mp.
// grid (Mp32/128, NV), block 128: thread = k, coalesced over k for every m.
__global__ void __launch_bounds__(128) alpha_kernel(const ModelDev* __restrict__ md, const double* __restrict__ Linv,
                                                    const double* __restrict__ V, int Mp32, double* __restrict__ alpha) {
  const int v = blockIdx.y, Mp = md->Mp, g = md->latent_group[md->vec_latent[v]];
  const int k = blockIdx.x * 128 + threadIdx.x;
  double acc = 0.0;
  if (k < Mp) {
    const double* Lg = Linv + (long long)g * Mp * Mp;
    const double* Vv = V + (long long)v * Mp;
    for (int m = k; m < Mp; m++) acc += Lg[(long long)m * Mp + k] * Vv[m];
  }
  alpha[(long long)v * Mp32 + k] = acc;
}

// Kuf^T in split planes + the FP64 mean partials.
//   KufT[g][plane][n][k] = k(z_k, x_n)   (zero for padded k / n),   n in [0, Bc), k in [0, Mp32)
//   dot_part[v][blockIdx.y][n] = sum_{k in this 128-block} Kuf[k][n] alpha[v][k]     for the vectors v of group g
// grid (Bc/64, Mp32/128, Q), block 128: warp w handles points n_base + w, w + 4, ...; lane covers k = k0 + lane + 32 j,
// so every store is a coalesced 128-byte row segment and the mean needs one warp reduction per point.
template <int DMAX>
__global__ void __launch_bounds__(128) kuf32_kernel(const ModelDev* __restrict__ md, const double* __restrict__ params,
                                                    Layout lo, const double* __restrict__ X, int N, int Bp, int Mp32,
                                                    float* __restrict__ KufT, const double* __restrict__ alpha,
                                                    double* __restrict__ dot_part, int rbn) {
  const int g = blockIdx.z, Mp = md->Mp, D = md->D, Mg = md->group_M[g];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.y * 128;
  const double* ls = params + lo.ls + (long long)g * D;
  const double* Z = params + lo.Z + (long long)g * Mp * D;
  const double kv = params[lo.kvar + g];
  double il[DMAX];  // 1 / lengthscale: one FP64 division per thread instead of D per point
#pragma unroll
  for (int d = 0; d < DMAX; d++) il[d] = d < D ? 1.0 / ls[d] : 0.0;
  double z[4][DMAX], zn[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int k = k0 + lane + 32 * j;
    zn[j] = 0.0;
#pragma unroll
    for (int d = 0; d < DMAX; d++) {
      z[j][d] = (d < D && k < Mg) ? Z[(long long)k * D + d] * il[d] : 0.0;
      zn[j] += z[j][d] * z[j][d];
    }
  }
  const int v0 = md->gvec_begin[g], nv = md->gvec_begin[g + 1] - v0;
  float* hi_base = KufT + (long long)g * 2 * Bp * Mp32;
  float* lo_base = hi_base + (long long)Bp * Mp32;
  for (int i = warp; i < 64; i += 4) {
    const int n = blockIdx.x * 64 + i;
    const bool valid = n < N;
    double x[DMAX], xn = 0.0;
#pragma unroll
    for (int d = 0; d < DMAX; d++) {
      x[d] = (valid && d < D) ? X[(long long)n * D + d] * il[d] : 0.0;
      xn += x[d] * x[d];
    }
    double kval[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int k = k0 + lane + 32 * j;
      double dot = 0.0;
#pragma unroll
      for (int d = 0; d < DMAX; d++) dot += z[j][d] * x[d];
      const double r2 = -2.0 * dot + (zn[j] + xn);
      kval[j] = (valid && k < Mg) ? kv * exp(-0.5 * r2) : 0.0;
      const float hi = tf32_hi((float)kval[j]);
      hi_base[(long long)n * Mp32 + k] = hi;
      lo_base[(long long)n * Mp32 + k] = (float)(kval[j] - (double)hi);
    }
    for (int q = 0; q < nv; q++) {
      const int v = md->gvec[v0 + q];
      const double* av = alpha + (long long)v * Mp32 + k0 + lane;
      double s = kval[0] * av[0] + kval[1] * av[32] + kval[2] * av[64] + kval[3] * av[96];
      s = warp_sum(s);
      if (lane == 0) dot_part[((long long)v * rbn + blockIdx.y) * Bp + n] = s;
    }
  }
}

}  // namespace t32
}  // namespace mogpe

// ---- INT8 (integer-sliced, FP64-accurate) mode: gemm_i8.cuh -------------------------------------------------------------
#include "gemm_i8.cuh"

namespace mogpe {
namespace oz {

// Per-group scales of the sliced operands: kuf_scale[g] = 2^e >= 2 kvar (|Kuf| <= kvar), a_scale[g] = 2^e >= 2 sqrt(kvar)
// (sum_m A[m][n]^2 <= kvar, so |A| <= sqrt(kvar)).   one block, Q threads
__global__ void group_scales_kernel(const ModelDev* __restrict__ md, const double* __restrict__ params, Layout lo,
                                    double* __restrict__ kuf_scale, double* __restrict__ a_scale) {
  const int g = threadIdx.x;
  if (g >= md->Q) return;
  const double kv = params[lo.kvar + g];
  kuf_scale[g] = pow2_scale(kv);
  a_scale[g] = pow2_scale(sqrt(kv));
}

// Kuf^T in NS byte planes + the FP64 mean partials (same roles as t32::kuf32_kernel).
//   KufT8[g][plane][n][k]: slices of k(z_k, x_n) / kuf_scale[g];   dot_part[v][blockIdx.y][n] = sum_k Kuf[k][n] alpha[v][k]
// grid (Bc/64, Mp32/128, Q), block 128: warp w handles points n_base + w, w + 4, ...; lane covers the FOUR CONSECUTIVE
// inducing indices k0 + 4 lane .. + 3, so the four bytes of a plane go out as one 32-bit store (128 B per warp).
// (ISL: integer slicing -- measured same-box 364 us against 330 with the FP64 slicing: this kernel is issue-bound, off)
template <int DMAX, int NS, bool ISL = false>
__global__ void __launch_bounds__(128) kuf8_kernel(const ModelDev* __restrict__ md, const double* __restrict__ params,
                                                   Layout lo, const double* __restrict__ X, int N, int Bp, int Mp32,
                                                   const double* __restrict__ kuf_scale, int8_t* __restrict__ KufT,
                                                   const double* __restrict__ alpha, double* __restrict__ dot_part, int rbn) {
  const int g = blockIdx.z, Mp = md->Mp, D = md->D, Mg = md->group_M[g];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.y * 128 + 4 * lane;
  const double* ls = params + lo.ls + (long long)g * D;
  const double* Z = params + lo.Z + (long long)g * Mp * D;
  const double kv = params[lo.kvar + g];
  const double inv_scale = 1.0 / kuf_scale[g];
  const int inv_e = 1023 - ((__double2hiint(kuf_scale[g]) >> 20) & 0x7ff);  // the scale is a power of two
  double il[DMAX];  // 1 / lengthscale: one FP64 division per thread instead of D per point (1 ulp from x / l)
#pragma unroll
  for (int d = 0; d < DMAX; d++) il[d] = d < D ? 1.0 / ls[d] : 0.0;
  double z[4][DMAX], zn[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int k = k0 + j;
    zn[j] = 0.0;
#pragma unroll
    for (int d = 0; d < DMAX; d++) {
      z[j][d] = (d < D && k < Mg) ? Z[(long long)k * D + d] * il[d] : 0.0;
      zn[j] += z[j][d] * z[j][d];
    }
  }
  const int v0 = md->gvec_begin[g], nv = md->gvec_begin[g + 1] - v0;
  int8_t* base = KufT + (long long)g * NS * Bp * Mp32;
  const long long plane = (long long)Bp * Mp32;
  for (int i = warp; i < 64; i += 4) {
    const int n = blockIdx.x * 64 + i;
    const bool valid = n < N;
    double x[DMAX], xn = 0.0;
#pragma unroll
    for (int d = 0; d < DMAX; d++) {
      x[d] = (valid && d < D) ? X[(long long)n * D + d] * il[d] : 0.0;
      xn += x[d] * x[d];
    }
    double kval[4];
    uint32_t wh[4], wl[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      double dot = 0.0;
#pragma unroll
      for (int d = 0; d < DMAX; d++) dot += z[j][d] * x[d];
      const double r2 = -2.0 * dot + (zn[j] + xn);
      kval[j] = (valid && k0 + j < Mg) ? kv * exp(-0.5 * r2) : 0.0;
      if constexpr (ISL) slice_words_int<NS>(kval[j], inv_e, wh[j], wl[j]);
      else slice_words<NS>(kval[j] * inv_scale, wh[j], wl[j]);
    }
    {  // four consecutive inducing indices -> one word per byte plane (4 x 4 byte transposes)
      uint32_t th[4], tl[4];
      transpose4x4_bytes(wh[0], wh[1], wh[2], wh[3], th);
      transpose4x4_bytes(wl[0], wl[1], wl[2], wl[3], tl);
#pragma unroll
      for (int b4 = 4 - NS / 2; b4 < 4; b4++) {  // byte b4 of a slice word = digit 3 - b4 of its half
        *reinterpret_cast<uint32_t*>(base + (3 - b4) * plane + (long long)n * Mp32 + k0) = th[b4];
        *reinterpret_cast<uint32_t*>(base + (NS / 2 + 3 - b4) * plane + (long long)n * Mp32 + k0) = tl[b4];
      }
    }
    for (int q2 = 0; q2 < nv && alpha != nullptr; q2++) {  // (alpha == nullptr: the means come from the A product's epilogue)
      const int v = md->gvec[v0 + q2];
      const double* av = alpha + (long long)v * Mp32 + k0;
      double s = kval[0] * av[0] + kval[1] * av[1] + kval[2] * av[2] + kval[3] * av[3];
      s = warp_sum(s);
      if (lane == 0) dot_part[((long long)v * rbn + blockIdx.y) * Bp + n] = s;
    }
  }
}

}  // namespace oz
}  // namespace mogpe
