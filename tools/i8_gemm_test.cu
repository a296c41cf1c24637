// Stand-alone check of the integer-sliced FP64-accurate GEMM (mogpe_b200/csrc/gemm_i8.cuh) against a CPU FP64 product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/i8_gemm_test tools/i8_gemm_test.cu
//   tools/i8_gemm_test <kmode 0|1|2> [M N batch reps]     D[m][n] = sum_k A[m][k] B[n][k],  A is M x M
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../mogpe_b200/csrc/gemm_i8.cuh"

using namespace mogpe;

#define CK(x)                                                                                 \
  do {                                                                                        \
    cudaError_t e_ = (x);                                                                     \
    if (e_ != cudaSuccess) {                                                                  \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);        \
      return 2;                                                                               \
    }                                                                                         \
  } while (0)

// dst [batch][NS][rows][cols] <- src [batch][rows][cols] with one fixed scale
template <int NSL>
__global__ void slice_fixed_kernel(const double* __restrict__ src, long long n_per_batch, double inv_scale,
                                   int8_t* __restrict__ dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (i >= n_per_batch) return;
  int8_t q[NSL];
  oz::slice8<NSL>(src[(long long)b * n_per_batch + i], inv_scale, q);
  for (int s = 0; s < NSL; s++) dst[((long long)b * NSL + s) * n_per_batch + i] = q[s];
}

template <int NSL>
int run(int argc, char** argv) {
  const int kmode = argc > 1 ? atoi(argv[1]) : 0;
  const int M = argc > 2 ? atoi(argv[2]) : 256;
  const int N = argc > 3 ? atoi(argv[3]) : 256;
  const int batch = argc > 4 ? atoi(argv[4]) : 2;
  int reps = argc > 5 ? atoi(argv[5]) : 0;
  const bool check = reps >= 0;
  if (reps < 0) reps = -reps;
  const int K = M;
  printf("int8-sliced gemm test: kmode=%d M=K=%d N=%d batch=%d NS=%d\n", kmode, M, N, batch, NSL);
  std::mt19937_64 rng(99);
  std::normal_distribution<double> nd(0.0, 1.0);
  std::vector<double> A((size_t)batch * M * K), B((size_t)batch * N * K);
  for (int b = 0; b < batch; b++)
    for (int m = 0; m < M; m++) {
      const double rowmag = exp(3.0 * nd(rng));  // rows of very different magnitude (as L^-1 has)
      for (int k = 0; k < K; k++) {
        double v = nd(rng) * rowmag * exp(2.0 * nd(rng));
        if (kmode == 1 && k > m) v = 0;
        if (kmode == 2 && k < m) v = 0;
        A[((size_t)b * M + m) * K + k] = v;
      }
    }
  double bmax = 0;
  for (auto& v : B) { v = nd(rng); bmax = fmax(bmax, fabs(v)); }
  double *dA, *dB, *dScale, *ss;
  int8_t *A8, *B8, *O8;
  int* err;
  CK(cudaMalloc(&dA, A.size() * 8));
  CK(cudaMalloc(&dB, B.size() * 8));
  CK(cudaMalloc(&dScale, (size_t)batch * M * 8));
  CK(cudaMalloc(&A8, A.size() * NSL));
  CK(cudaMalloc(&B8, B.size() * NSL));
  CK(cudaMalloc(&O8, (size_t)batch * NSL * N * M));
  CK(cudaMalloc(&ss, (size_t)batch * (M / 128) * N * 8));
  CK(cudaMalloc(&err, 4));
  CK(cudaMemset(err, 0, 4));
  CK(cudaMemset(O8, 0x7f, (size_t)batch * NSL * N * M));
  CK(cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice));
  oz::row_scale_kernel<<<dim3(M, batch), 128>>>(dA, M, (long long)M * K, K, nullptr, 0, 0, dScale, M);
  oz::slice_square_kernel<NSL><<<dim3(M / 32, M / 32, batch), dim3(32, 8)>>>(dA, M, (long long)M * K, K, nullptr, 0, 0, dScale, A8, M);
  const double bscale = oz::pow2_scale(bmax);
  slice_fixed_kernel<NSL><<<dim3((unsigned)(((long long)N * K + 255) / 256), batch), 256>>>(dB, (long long)N * K, 1.0 / bscale, B8);
  CK(cudaGetLastError());
  // output scale: |D| <= rowmax(A) * bmax * K is far too pessimistic for slicing; the test uses the true max of D
  std::vector<double> Dref((size_t)batch * M * N);
  double dmax = 0;
  if (check) {
    for (int b = 0; b < batch; b++)
      for (int m = 0; m < M; m++)
        for (int n = 0; n < N; n++) {
          double acc = 0;
          for (int k = 0; k < K; k++) acc += A[((size_t)(batch - 1 - b) * M + m) * K + k] * B[((size_t)b * N + n) * K + k];
          Dref[((size_t)b * M + m) * N + n] = acc;
          dmax = fmax(dmax, fabs(acc));
        }
  } else {
    dmax = 1e6;
  }
  const double oscale = oz::pow2_scale(dmax);
  CUtensorMap mapA, mapB;
  if (!oz::make_map_i8(&mapA, A8, batch, M, K, oz::BM, NSL)) { printf("tensor map A failed\n"); return 3; }
  if (!oz::make_map_i8(&mapB, B8, batch, N, K, oz::BN, NSL)) { printf("tensor map B failed\n"); return 3; }
  oz::Params p;
  memset(&p, 0, sizeof(p));
  p.m_tiles = M / 128; p.K = K; p.kmode = kmode;
  p.a_scale = dScale; p.a_scale_stride = M;
  p.out = O8; p.out_batch_stride = (long long)NSL * N * M; p.out_plane_stride = (long long)N * M; p.ldo = M;
  p.ss_part = ss; p.ld_ss = N; p.err_flag = err;
  std::vector<double> hs(2 * batch);
  for (int b = 0; b < batch; b++) {
    p.mapA[b] = batch - 1 - b; p.mapB[b] = b; hs[b] = bscale; hs[batch + b] = oscale;
  }
  double* dsc;
  CK(cudaMalloc(&dsc, hs.size() * 8));
  CK(cudaMemcpy(dsc, hs.data(), hs.size() * 8, cudaMemcpyHostToDevice));
  p.b_scale = dsc; p.out_scale = dsc + batch;
  CK(oz::launch<NSL>(mapA, mapB, p, N / oz::BN, batch, 0));
  cudaError_t se = cudaDeviceSynchronize();
  if (se != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(se)); return 4; }
  bool pass = true;
  if (check) {
    std::vector<int8_t> O((size_t)batch * NSL * N * M);
    std::vector<double> SS((size_t)batch * (M / 128) * N);
    CK(cudaMemcpy(O.data(), O8, O.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(SS.data(), ss, SS.size() * 8, cudaMemcpyDeviceToHost));
    double max_rel = 0, max_ss_rel = 0;
    int bad = 0;
    for (int b = 0; b < batch; b++) {
      std::vector<double> colss((size_t)(M / 128) * N, 0.0);
      for (int m = 0; m < M; m++) {
        double arow = 0;
        for (int k = 0; k < K; k++) arow = fmax(arow, fabs(A[((size_t)(batch - 1 - b) * M + m) * K + k]));
        for (int n = 0; n < N; n++) {
          double got = 0, w = 1.0 / 128.0;
          for (int s = 0; s < NSL; s++) { got += w * (double)O[(((size_t)b * NSL + s) * N + n) * M + m]; w /= 128.0; }
          got *= oscale;
          const double want = Dref[((size_t)b * M + m) * N + n];
          // error unit: the slicing resolution of the OUTPUT planes (2^-57 oscale) plus the product error bound
          const double rel = fabs(got - want) / (arow * bmax * K + 1e-300);
          if (rel > max_rel) max_rel = rel;
          const double tol = NSL == 8 ? 1e-12 : 3e-8;  // product truncation ~ (NS+1) K 2^(-7 NS), relative to the row scale
          if (!(fabs(got - want) <= tol * arow * bmax * K + oscale * (NSL == 8 ? 1e-15 : 1e-11))) {
            if (bad < 6)
              printf("  bad b=%d m=%d n=%d got=%.17g want=%.17g\n", b, m, n, got, want);
            bad++;
          }
          colss[(size_t)(m / 128) * N + n] += want * want;
        }
      }
      for (size_t i = 0; i < colss.size(); i++) {
        const double rel = fabs(SS[(size_t)b * colss.size() + i] - colss[i]) / (colss[i] + 1e-300);
        if (rel > max_ss_rel) max_ss_rel = rel;
      }
    }
    printf("max |err| / (rowmax(A) max(B) K) = %.3e (re-sliced output)   max rel err of column sums of squares = %.3e\n", max_rel,
           max_ss_rel);
    pass = bad == 0 && max_ss_rel < (NSL == 8 ? 1e-10 : 1e-6);
    printf(pass ? "PASS\n" : "FAIL\n");
  }
  if (reps > 0) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int variant = 0; variant < 3; variant++) {
      p.out = variant == 1 ? O8 : nullptr;
      p.ss_part = variant == 2 ? nullptr : ss;
      oz::launch<NSL>(mapA, mapB, p, N / oz::BN, batch, 0);
      cudaEventRecord(e0);
      for (int r = 0; r < reps; r++) oz::launch<NSL>(mapA, mapB, p, N / oz::BN, batch, 0);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double frac = kmode == 0 ? 1.0 : 0.5;
      const double flops = 2.0 * M * (double)K * N * batch * frac;
      printf("%s: %.3f ms/launch, %.1f TFLOP/s FP64-equivalent algorithmic (x%d INT8 MMAs: %.0f TOP/s)\n",
             variant == 1 ? "with re-sliced output" : variant == 2 ? "no epilogue outputs (main loop + TMEM read + combine only)" : "stats only", ms / reps, flops / (ms / reps * 1e-3) * 1e-12, NSL * (NSL + 1) / 2,
             NSL * (NSL + 1) / 2 * flops / (ms / reps * 1e-3) * 1e-12);
    }
  }
  return pass ? 0 : 1;
}

int main(int argc, char** argv) {
  const int ns = getenv("I8_NS") ? atoi(getenv("I8_NS")) : 8;
  return ns == 6 ? run<6>(argc, argv) : run<8>(argc, argv);
}
