// FP64 throughput probe for B200 (sm_100a): DFMA vs DMMA.8x8x4 issue rate.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_probe fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

template<int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<int ILP>
__global__ void dmma_kernel(double* out, int iters, double a, double b) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { c0[i] = threadIdx.x; c1[i] = i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int clk=0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("device %s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, clk);
  double* out; CK(cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sms = p.multiProcessorCount;
  for (int threads : {128, 256, 512, 1024}) {
    for (int bps : {1, 2}) {
      if (threads * bps > 2048) continue;
      int iters = 20000;
      dfma_kernel<8><<<sms * bps, threads>>>(out, 100, 1.0000001, 1e-9);
      CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      dfma_kernel<8><<<sms * bps, threads>>>(out, iters, 1.0000001, 1e-9);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double flops = 2.0 * 8 * iters * (double)threads * sms * bps;
      printf("DFMA threads=%4d bps=%d: %.2f TFLOP/s (%.3f ms)\n", threads, bps, flops / ms * 1e-9, ms);
      dmma_kernel<8><<<sms * bps, threads>>>(out, 100, 1.0000001, 1e-9);
      CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      dmma_kernel<8><<<sms * bps, threads>>>(out, iters, 1.0000001, 1e-9);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      cudaEventElapsedTime(&ms, e0, e1);
      flops = 2.0 * 256 * 8 * iters * (double)(threads / 32) * sms * bps;
      printf("DMMA threads=%4d bps=%d: %.2f TFLOP/s (%.3f ms)\n", threads, bps, flops / ms * 1e-9, ms);
    }
  }
  return 0;
}
