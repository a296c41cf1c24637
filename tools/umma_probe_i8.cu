// Micro-probe of tcgen05.mma kind::i8 operand layouts (no TMA): threads write the operand tiles into shared memory in
// a chosen canonical layout, one thread issues ONE MMA (M = 128, N = 64..256, K = 8), the result is read back from
// TMEM and compared with the exact product.  Used to pin down the shared-memory descriptor semantics on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_probe_i8 tools/umma_probe_i8.cu
// Signed 8-bit operands, exact int32 accumulation in TMEM, K = 32 per instruction: the building block of an
// integer-sliced (Ozaki-scheme) FP64-accurate product.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>
#include <stdint.h>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__);         \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

enum Layout : int { K_NOSW = 0, K_SW64 = 1, K_SW128 = 2, MN_SW128 = 3, MN_NOSW = 4 };

struct Variant {
  int a_layout, b_layout, N;
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
  const char* name;
};

// byte offset of element (r, k) of an operand tile (r = M or N index, k in [0, 8)), for rows R
__host__ __device__ inline uint32_t elem_offset(int layout, int r, int k, uint32_t lbo, uint32_t sbo) {
  switch (layout) {
    case K_NOSW:  // core matrix = 8 rows x 16 bytes, contiguous 128 B; K cores LBO apart, row groups SBO apart
      return (uint32_t)(r / 8) * sbo + (uint32_t)(k / 16) * lbo + (uint32_t)(r % 8) * 16 + (uint32_t)(k % 16);
    case K_SW64: {  // rows of 64 B (64 int8, only the first 32 used), 16-byte chunk ^= (addr >> 7) & 3
      const uint32_t lin = (uint32_t)r * 64 + (uint32_t)k;
      return lin ^ (((lin >> 7) & 3) << 4);
    }
    case K_SW128: {  // rows of 128 B, chunk ^= (addr >> 7) & 7
      const uint32_t lin = (uint32_t)r * 128 + (uint32_t)k;
      return lin ^ (((lin >> 7) & 7) << 4);
    }
    case MN_SW128: {  // [r / 32 blocks, LBO apart][8 k rows of 128 B][32 r], chunk ^= k
      const uint32_t lin = (uint32_t)k * 128 + (uint32_t)(r % 32) * 4;
      return (uint32_t)(r / 32) * lbo + (lin ^ (((lin >> 7) & 7) << 4));
    }
    case MN_NOSW:  // core matrix = 8 k rows x 16 bytes (4 r); r cores: stride sbo?/lbo? -> (r/4)*lbo ; k groups sbo
      return (uint32_t)(r / 4) * lbo + (uint32_t)k * 16 + (uint32_t)(r % 4) * 4;
  }
  return 0;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1) probe_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int a_layout,
                                                       int b_layout, int N, uint32_t a_lbo, uint32_t a_sbo, uint32_t b_lbo,
                                                       uint32_t b_sbo, int* __restrict__ D, int issue_mode) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* g = raw + (base - smem_u32(raw));
  uint8_t* sa = g;            // 32 KB region
  uint8_t* sb = g + 32768;    // 64 KB region
  uint64_t* bar = reinterpret_cast<uint64_t*>(g + 98304);
  uint32_t* slot = reinterpret_cast<uint32_t*>(g + 98304 + 16);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 98304 / 4; i += 128) reinterpret_cast<uint32_t*>(g)[i] = 0;
  __syncthreads();
  for (int i = tid; i < 128 * 32; i += 128) {
    const int r = i / 32, k = i % 32;
    *reinterpret_cast<int8_t*>(sa + elem_offset(a_layout, r, k, a_lbo, a_sbo)) = A[r * 32 + k];
  }
  for (int i = tid; i < N * 32; i += 128) {
    const int r = i / 32, k = i % 32;
    *reinterpret_cast<int8_t*>(sb + elem_offset(b_layout, r, k, b_lbo, b_sbo)) = B[r * 32 + k];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async-proxy (tensor core) reads
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  const bool a_mn = a_layout >= MN_SW128, b_mn = b_layout >= MN_SW128;
  auto lt = [](int l) -> uint32_t { return l == K_SW64 ? 4u : (l == K_SW128 || l == MN_SW128) ? 2u : 0u; };
  const uint64_t ad = make_desc(smem_u32(sa), a_lbo, a_sbo, lt(a_layout));
  const uint64_t bd = make_desc(smem_u32(sb), b_lbo, b_sbo, lt(b_layout));
  uint32_t idesc = (2u << 4) /* D = S32 */ | (1u << 7) /* A = signed int8 */ | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
                   ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  if (issue_mode == 0) {
    if (tid == 0) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(0u)
          : "memory");
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
  } else {
    if (warp == 0) {  // converged warp, elect.sync as CUTLASS does
      uint32_t pred = 0;
      asm volatile(
          "{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred)::"memory");
      if (pred) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(0u)
            : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
      }
      __syncwarp();
    }
  }
  // wait for the MMA
  {
    uint32_t done = 0;
    for (long long spin = 0; spin < (1LL << 26) && !done; spin++) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(smem_u32(bar)), "r"(0u)
          : "memory");
    }
    if (!done) __trap();
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c = 0; c < N; c += 8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c)
                 : "memory");
    for (int i = 0; i < 8; i++) D[(size_t)tid * N + c + i] = (int)r[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

int main() {
  const Variant vs[] = {
      {K_NOSW, K_NOSW, 64, 128, 256, 128, 256, "A K-major no-swizzle, B K-major no-swizzle (LBO = K-core stride 128, SBO = 8-row stride 256)"},
      {K_NOSW, K_NOSW, 64, 256, 128, 256, 128, "same data, LBO/SBO swapped in the descriptor (expect FAIL unless semantics are swapped)"},
      {K_SW128, K_SW128, 64, 16, 1024, 16, 1024, "A, B K-major 128B swizzle (SBO 1024)"},
      {K_SW64, K_SW64, 64, 16, 512, 16, 512, "A, B K-major 64B swizzle (SBO 512)"},
      {K_SW64, K_SW64, 256, 16, 512, 16, 512, "A, B K-major 64B swizzle, N = 256"},
      {K_SW128, K_SW128, 256, 16, 1024, 16, 1024, "A, B K-major 128B swizzle, N = 256"},
  };
  std::vector<int8_t> A(128 * 32), B(256 * 32);
  srand(7);
  for (auto& v : A) v = (int8_t)(rand() % 256 - 128);
  for (auto& v : B) v = (int8_t)(rand() % 256 - 128);
  int8_t *dA, *dB;
  int* dD;
  CK(cudaMalloc(&dA, A.size()));
  CK(cudaMalloc(&dB, B.size()));
  CK(cudaMalloc(&dD, 128 * 256 * 4));
  CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
  const int smem = 98304 + 1024 + 64;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int mode = 0; mode < 2; mode++)
    for (const Variant& v : vs) {
      // the tile must fit its region for the layout's address range
      CK(cudaMemset(dD, 0xff, 128 * 256 * 4));
      probe_kernel<<<1, 128, smem>>>(dA, dB, v.a_layout, v.b_layout, v.N, v.a_lbo, v.a_sbo, v.b_lbo, v.b_sbo, dD, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("[mode %d] %s: kernel error %s\n", mode, v.name, cudaGetErrorString(e));
        return 1;
      }
      std::vector<int> D(128 * v.N);
      CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0, maxabs = 0;
      long long nz = 0;
      for (int m = 0; m < 128; m++)
        for (int n = 0; n < v.N; n++) {
          double want = 0;
          for (int k = 0; k < 32; k++) want += (double)A[m * 32 + k] * (double)B[n * 32 + k];
          const double got = D[(size_t)m * v.N + n];
          maxerr = fmax(maxerr, fabs(got - want));
          maxabs = fmax(maxabs, fabs(want));
          if (got != 0) nz++;
        }
      printf("[issue %s] %-100s max|err| %.3e (max|want| %.2f, nonzero %lld) %s\n", mode ? "elect" : "tid0", v.name, maxerr, maxabs, nz,
             maxerr == 0 ? "EXACT" : "WRONG");
    }
  return 0;
}
