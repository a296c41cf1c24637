// Stand-alone check of the tcgen05 3xTF32 GEMM (mogpe_b200/csrc/gemm_tf32.cuh) against a CPU FP64 product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tf32_gemm_test tools/tf32_gemm_test.cu
//   tools/tf32_gemm_test <unused> <kmode 0|1|2> [rows K N batch reps]   D[m][n] = sum_k A[m][k] B[n][k]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../mogpe_b200/csrc/gemm_tf32.cuh"

using namespace mogpe;

static float host_tf32_hi(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xffffe000u;
  memcpy(&x, &u, 4);
  return x;
}

#define CK(x)                                                                                 \
  do {                                                                                        \
    cudaError_t e_ = (x);                                                                     \
    if (e_ != cudaSuccess) {                                                                  \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);        \
      return 2;                                                                               \
    }                                                                                         \
  } while (0)

int main(int argc, char** argv) {
  const int a_mn = 0;
  const int kmode = argc > 2 ? atoi(argv[2]) : 0;
  const int rows = argc > 3 ? atoi(argv[3]) : 256;  // M (= K for the triangular modes)
  const int K = argc > 4 ? atoi(argv[4]) : 256;
  const int N = argc > 5 ? atoi(argv[5]) : 512;
  const int batch = argc > 6 ? atoi(argv[6]) : 2;
  int reps = argc > 7 ? atoi(argv[7]) : 0;
  const bool check = reps >= 0;
  if (reps < 0) reps = -reps;
  printf("tf32x3 gemm test: a_mn_major=%d kmode=%d rows=%d K=%d N=%d batch=%d\n", a_mn, kmode, rows, K, N, batch);
  std::mt19937_64 rng(1234);
  std::normal_distribution<double> nd(0.0, 1.0);
  // logical A(m,k), stored [m][k] (K-major) or [k][m] (MN-major); B stored [k][n]
  std::vector<double> A((size_t)batch * rows * K), B((size_t)batch * N * K);  // B stored [n][k]
  for (int b = 0; b < batch; b++)
    for (int m = 0; m < rows; m++)
      for (int k = 0; k < K; k++) {
        double v = nd(rng);
        if (kmode == 1 && k > m) v = 0;
        if (kmode == 2 && k < m) v = 0;
        const size_t idx = a_mn ? ((size_t)b * K + k) * rows + m : ((size_t)b * rows + m) * K + k;
        A[idx] = v;
      }
  for (auto& v : B) v = nd(rng);
  double *dA, *dB;
  CK(cudaMalloc(&dA, A.size() * 8));
  CK(cudaMalloc(&dB, B.size() * 8));
  CK(cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice));
  float *A32, *B32, *O32;
  CK(cudaMalloc(&A32, A.size() * 8));
  CK(cudaMalloc(&B32, B.size() * 8));
  CK(cudaMalloc(&O32, (size_t)batch * 2 * rows * N * 4));
  CK(cudaMemset(O32, 0xff, (size_t)batch * 2 * rows * N * 4));
  const int a_rows = a_mn ? K : rows, a_cols = a_mn ? rows : K;
  t32::split_planes_kernel<<<dim3((a_cols + 255) / 256, a_rows, batch), 256>>>(dA, a_rows, a_cols, (long long)a_rows * a_cols,
                                                                               a_cols, A32, a_rows, a_cols, 0);
  t32::split_planes_kernel<<<dim3((K + 255) / 256, N, batch), 256>>>(dB, N, K, (long long)K * N, K, B32, N, K, 0);
  CK(cudaGetLastError());
  CUtensorMap mapA, mapB;
  bool ok = t32::make_map_kmajor(&mapA, A32, batch, rows, K, t32::BM);
  if (!ok) { printf("tensor map A failed\n"); return 3; }
  if (!t32::make_map_kmajor(&mapB, B32, batch, N, K, t32::BN)) { printf("tensor map B failed\n"); return 3; }
  double* ss;
  int* err;
  CK(cudaMalloc(&ss, (size_t)batch * (rows / 128) * N * 8));
  CK(cudaMalloc(&err, 4));
  CK(cudaMemset(err, 0, 4));
  t32::Params p;
  memset(&p, 0, sizeof(p));
  p.m_tiles = rows / 128; p.K = K; p.kmode = kmode;
  p.out = O32; p.out_batch_stride = (long long)2 * rows * N; p.out_plane_stride = (long long)rows * N; p.ldo = rows;  // [n][m]
  p.ss_part = ss; p.ld_ss = N; p.err_flag = err;
  for (int b = 0; b < batch; b++) { p.mapA[b] = batch - 1 - b; p.mapB[b] = b; }  // output b = A[batch-1-b] * B[b]
  CK(t32::launch(mapA, mapB, p, N / 256, batch, 0));
  cudaError_t se = cudaDeviceSynchronize();
  if (se != cudaSuccess) {
    printf("kernel failed: %s\n", cudaGetErrorString(se));
    return 4;
  }
  std::vector<float> O((size_t)batch * 2 * rows * N);
  std::vector<double> SS((size_t)batch * (rows / 128) * N);
  CK(cudaMemcpy(O.data(), O32, O.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(SS.data(), ss, SS.size() * 8, cudaMemcpyDeviceToHost));
  double max_rel = 0, max_ss_rel = 0;
  int bad_printed = 0;
  for (int b = 0; b < batch && check; b++) {
    const int ba = batch - 1 - b;
    std::vector<double> colss((size_t)(rows / 128) * N, 0.0);
    for (int m = 0; m < rows; m++)
      for (int n = 0; n < N; n++) {
        double acc = 0, mag = 0;
        for (int k = 0; k < K; k++) {
          const double a = a_mn ? A[((size_t)ba * K + k) * rows + m] : A[((size_t)ba * rows + m) * K + k];
          const double bb = B[((size_t)b * N + n) * K + k];
          acc += a * bb;
          mag += fabs(a * bb);
        }
        const double got = (double)O[((size_t)b * 2 * N + n) * rows + m] + (double)O[((size_t)b * 2 * N + N + n) * rows + m];
        const double rel = fabs(got - acc) / (mag + 1e-30);
        if (!(rel <= max_rel)) max_rel = rel;
        if (!(rel < 1e-4) && bad_printed < 8) {
          printf("  bad b=%d m=%d n=%d got=%.8g want=%.8g\n", b, m, n, got, acc);
          bad_printed++;
        }
        colss[(size_t)(m / 128) * N + n] += acc * acc;
      }
    for (size_t i = 0; i < colss.size(); i++) {
      const double rel = fabs(SS[(size_t)b * colss.size() + i] - colss[i]) / (colss[i] + 1e-30);
      if (!(rel <= max_ss_rel)) max_ss_rel = rel;
    }
  }
  printf("max |err| / sum|a b| = %.3e   max rel err of column sums of squares = %.3e\n", max_rel, max_ss_rel);
  const bool pass = max_rel < 2e-5 && max_ss_rel < 1e-4;
  printf(pass ? "PASS\n" : "FAIL\n");
  if (reps > 0) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    p.out = nullptr;
    for (int variant = 0; variant < 2; variant++) {
      if (variant == 1) { p.out = O32; }
      t32::launch(mapA, mapB, p, N / 256, batch, 0);
      cudaEventRecord(e0);
      for (int r = 0; r < reps; r++) t32::launch(mapA, mapB, p, N / 256, batch, 0);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double frac = kmode == 0 ? 1.0 : 0.5;
      const double flops = 2.0 * rows * (double)K * N * batch * frac;
      printf("%s: %.3f ms/launch, %.1f TFLOP/s algorithmic (x3 TF32 MMAs: %.1f TF/s tensor)\n",
             variant ? "with output planes" : "stats only", ms / reps, flops / (ms / reps * 1e-3) * 1e-12,
             3 * flops / (ms / reps * 1e-3) * 1e-12);
    }
  }
  return pass ? 0 : 1;
}
