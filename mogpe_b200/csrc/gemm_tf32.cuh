// Batched 3xTF32 GEMM on the sm_100a 5th-generation tensor cores: tcgen05.mma (kind::tf32) issued by one thread,
// operands staged in shared memory by TMA (cp.async.bulk.tensor, 64B swizzle), FP32 accumulators in TMEM, read back
// with tcgen05.ld for a fused epilogue.  This is the "TF32 / FP32 mode" of the hot path (BASELINE.json north_star;
// SURVEY 8d C5 "FP64 vs TF32"): the M x M x B products of the sparse-GP conditional
//      A   = L^-1 Kuf           (reference: tf.linalg.triangular_solve in gpflow base_conditional)
//      LTA = tril(q_sqrt)^T A   (reference: tf.linalg.matmul(q_sqrt, A, transpose_a=True))
// run here; Kuu, its Cholesky factor and L^-1 stay FP64 (M^3, small) so the factorisation never fails in low precision.
//
// Precision: every FP64 operand x is split into two FP32 planes  hi = fp32(x) truncated to 11 significant bits (exactly
// a TF32 number), lo = fp32(x - hi); the product is hi*hi + lo*hi + hi*lo (three MMAs per k-step, FP32 accumulate).  The
// dropped lo*lo term and the TF32 truncation of lo leave ~2^-20 relative error per product: FP32-class results at one
// third of the TF32 tensor peak (~10x the FP64 DMMA peak).
//
// Operand layout: BOTH operands K-major (tools/umma_probe.cu: on sm_100a kind::tf32 with an MN-major operand silently
// produces zeros, profiles/umma_probe_r01.log), i.e. D[m][n] = sum_k Aop[m][k] * Bop[n][k] with k contiguous in memory
// for both.  So Kuf and A are kept TRANSPOSED ([point n][inducing index k]) in this mode and q_sqrt^T is materialised.
// Tile: two D[128 x 256] row tiles per CTA (UMMA M = 128, N = 256, K = 8 per instruction) sharing one B stage, BK = 16 per
// pipeline stage, 3 stages of 64 KB:  A0, A1 [plane][128 rows][16 fp32], B [plane][256 rows][16 fp32], rows of 64 B,
// 64B-swizzled by TMA.
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-9 = epilogue
// (two per 32-lane TMEM quarter, four 32-column chunks each).  Triangular operands shorten the k loop per row tile.
// Epilogue: column sums of squares of the tile (butterfly transpose-reduce over the 32 lanes, then over the 4 warps)
// written as one FP64 partial per row tile (deterministic), and optionally the tile itself, TRANSPOSED ([n][m], warp-wide
// coalesced 128-byte stores), as hi / lo planes so that it can be the K-major B operand of the next product.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mogpe {
namespace t32 {

constexpr int BM = 128, BN = 256, BK = 16, STAGES = 3, UK = 8;
constexpr int A_PLANE_BYTES = BM * BK * 4;          // 8 KB
constexpr int B_PLANE_BYTES = BN * BK * 4;          // 16 KB
constexpr int A_STAGE_BYTES = 2 * A_PLANE_BYTES;    // hi + lo
constexpr int B_STAGE_BYTES = 2 * B_PLANE_BYTES;
constexpr int STAGE_BYTES = 2 * A_STAGE_BYTES + B_STAGE_BYTES;  // two row tiles + one column tile: 64 KB
constexpr int RED_BYTES = 2 * 4 * BN * 4;                    // cross-warp column-sum scratch, one per row tile
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + RED_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;
constexpr int THREADS = 320;  // warp 0: TMA, warp 1: MMA, warps 2-9: epilogue (two per TMEM lane quarter)
constexpr int TMEM_COLS = 512;  // two 128 x 256 FP32 accumulators

enum : int { K_FULL = 0, K_LOWER = 1, K_UPPER = 2 };

struct Params {
  int m_tiles;   // row tiles of 128
  int K;         // reduction length (multiple of 16)
  int kmode;     // K_FULL; K_LOWER: A(m,k) = 0 for k > m; K_UPPER: A(m,k) = 0 for k < m
  float* out;              // optional TRANSPOSED result planes [batch][2][n][ldo] (ldo >= m_tiles*128)
  long long out_batch_stride;  // floats per batch (= 2 * n_rows * ldo)
  long long out_plane_stride;  // floats per plane
  int ldo;
  double* ss_part;  // optional [batch][m_tiles][ld_ss]: column sums of squares of each row tile
  int ld_ss;
  int* err_flag;    // set to 1 if a barrier wait times out (the kernel traps instead of hanging)
  int mapA[64], mapB[64];  // batch -> outermost tensor-map coordinate of the A / B operand
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded spin: a descriptor / coordinate bug must end as a trapped launch, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err_flag) {
  const uint32_t a = smem_u32(bar);
  uint32_t done = 0;
  for (long long spin = 0; spin < (1LL << 28); spin++) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (done) return;
  }
  if (err_flag) atomicExch(err_flag, 1);
  __trap();
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// Shared-memory matrix descriptor (sm_100 UMMA): start address, leading / stride byte offsets (16-byte units),
// descriptor version 1, swizzle mode in bits 61-63 (2 = 128B, 4 = 64B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

// Instruction descriptor: FP32 accumulator, TF32 x TF32, M = 128, N = 256, both operands K-major.
__device__ __forceinline__ uint32_t make_idesc() {
  uint32_t d = 0;
  d |= 1u << 4;                      // c_format = F32
  d |= 2u << 7;                      // a_format = TF32
  d |= 2u << 10;                     // b_format = TF32
  d |= (uint32_t)(BN >> 3) << 17;
  d |= (uint32_t)(BM >> 4) << 24;
  return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// hi = x with the 13 low mantissa bits cleared (a TF32 number whether the tensor core truncates or rounds)
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// One CTA = TWO adjacent row tiles (2p, 2p+1) x one 256-column tile: both accumulators (2 x 256 TMEM columns = all 512)
// are fed from the same B stage, so the L2 -> shared-memory traffic per MMA is 2/3 of the one-tile version (the kernel
// is bound by that traffic).  With a triangular operand the two tiles have k ranges that differ by 128: the loop runs
// in the direction in which the SHORTER tile finishes first (ascending k for K_LOWER, descending for K_UPPER); its
// accumulator is committed early and its epilogue overlaps the last 8 stages of the other tile.
// grid (m_pairs, n_tiles, batch): the pairs of one column tile are neighbours in launch order (longest k loop first), so
// they run at the same time and the B tile they share comes from DRAM once and from L2 afterwards (with the pair index
// outermost ncu showed 2.3x the algorithmic DRAM reads, profiles/tf32_gemm_ncu_r01.txt).
__global__ void __launch_bounds__(THREADS, 1) gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                  const __grid_constant__ CUtensorMap mapB,
                                                                  const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // 1024-byte aligned tile area (swizzle atoms)
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  float* red = reinterpret_cast<float*>(gbase + STAGES * STAGE_BYTES);  // [2 tiles][4 warps][BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(gbase + STAGES * STAGE_BYTES + RED_BYTES);
  uint64_t* full = bars;              // [STAGES] TMA -> MMA
  uint64_t* empty = bars + STAGES;    // [STAGES] MMA -> TMA
  uint64_t* acc_full = bars + 2 * STAGES;  // [2] accumulator of tile 0 / 1 complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.y, batch = blockIdx.z;
  const int m_pairs = (p.m_tiles + 1) / 2;
  const int pair = (p.kmode == K_UPPER) ? (int)blockIdx.x : m_pairs - 1 - (int)blockIdx.x;
  const int n0 = n_tile * BN;
  const bool has1 = 2 * pair + 1 < p.m_tiles;
  // k range [kb_t, ke_t) of each tile, and of the pair
  int kb_t[2], ke_t[2];
#pragma unroll
  for (int t = 0; t < 2; t++) {
    const int m0 = (2 * pair + t) * BM;
    kb_t[t] = (p.kmode == K_UPPER) ? m0 : 0;
    ke_t[t] = (p.kmode == K_LOWER) ? min(p.K, m0 + BM) : p.K;
  }
  if (!has1) { kb_t[1] = 0; ke_t[1] = 0; }
  const int k_lo = has1 ? min(kb_t[0], kb_t[1]) : kb_t[0], k_hi = has1 ? max(ke_t[0], ke_t[1]) : ke_t[0];
  const int num_kb = (k_hi - k_lo) / BK;
  const bool descending = p.kmode == K_UPPER;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&acc_full[0], 1);
    mbar_init(&acc_full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapB)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      const int ca = p.mapA[batch], cb = p.mapB[batch];
      for (int kb = 0; kb < num_kb; kb++) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        const int k0 = descending ? k_hi - (kb + 1) * BK : k_lo + kb * BK;
        const bool use0 = k0 >= kb_t[0] && k0 < ke_t[0], use1 = k0 >= kb_t[1] && k0 < ke_t[1];
        mbar_wait(&empty[s], ph ^ 1, p.err_flag);
        mbar_expect_tx(&full[s], B_STAGE_BYTES + (use0 ? A_STAGE_BYTES : 0) + (use1 ? A_STAGE_BYTES : 0));
        const uint32_t sa = base + s * STAGE_BYTES, sb = sa + 2 * A_STAGE_BYTES;
        if (use0) tma_load_4d(sa, &mapA, smem_u32(&full[s]), k0, 2 * pair * BM, 0, ca);                        // {k, m, plane, batch}
        if (use1) tma_load_4d(sa + A_STAGE_BYTES, &mapA, smem_u32(&full[s]), k0, (2 * pair + 1) * BM, 0, ca);
        tma_load_4d(sb, &mapB, smem_u32(&full[s]), k0, n0, 0, cb);                                            // {k, n, plane, batch}
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc();
      uint32_t acc[2] = {0, 0};
      for (int kb = 0; kb < num_kb; kb++) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        const int k0 = descending ? k_hi - (kb + 1) * BK : k_lo + kb * BK;
        mbar_wait(&full[s], ph, p.err_flag);
        tc_fence_after();
        const uint32_t sa = base + s * STAGE_BYTES, sb = sa + 2 * A_STAGE_BYTES;
#pragma unroll
        for (int t = 0; t < 2; t++) {
          if (!(k0 >= kb_t[t] && k0 < ke_t[t])) continue;
          const uint32_t d_tmem = tmem_base + (uint32_t)(t * BN);
#pragma unroll
          for (int j = 0; j < BK / UK; j++) {
            uint64_t ad[2], bd[2];
#pragma unroll
            for (int pl = 0; pl < 2; pl++) {
              // K-major, 64B swizzle: 8-row groups 512 B apart (SBO); the k-step advances the start address by 32 B
              ad[pl] = make_smem_desc(sa + t * A_STAGE_BYTES + pl * A_PLANE_BYTES + j * UK * 4, 16, 8 * BK * 4, 4);
              bd[pl] = make_smem_desc(sb + pl * B_PLANE_BYTES + j * UK * 4, 16, 8 * BK * 4, 4);
            }
            umma_tf32(d_tmem, ad[1], bd[0], idesc, acc[t]);  // lo * hi
            acc[t] = 1;
            umma_tf32(d_tmem, ad[0], bd[1], idesc, 1);       // hi * lo
            umma_tf32(d_tmem, ad[0], bd[0], idesc, 1);       // hi * hi
          }
        }
        umma_commit(&empty[s]);  // smem slot is free once these MMAs have read it
#pragma unroll
        for (int t = 0; t < 2; t++) {
          const bool last = descending ? (k0 == kb_t[t]) : (k0 + BK == ke_t[t]);
          if (last && ke_t[t] > kb_t[t]) umma_commit(&acc_full[t]);  // this tile's accumulator is complete
        }
      }
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
    const int q = warp & 3;
    const int t_first = descending ? 1 : 0;  // the tile with the shorter k range finishes first
#pragma unroll 1
    for (int it = 0; it < 2; it++) {
      const int t = it == 0 ? t_first : 1 - t_first;
      if (t == 1 && !has1) continue;
      const int m_tile = 2 * pair + t, m0 = m_tile * BM;
      mbar_wait(&acc_full[t], 0, p.err_flag);
      tc_fence_after();
      float* redt = red + t * 4 * BN;
      float* ocol_hi = nullptr;  // element (n0, this thread's row m) of the transposed result
      float* ocol_lo = nullptr;
      if (p.out) {
        ocol_hi = p.out + (long long)batch * p.out_batch_stride + (long long)n0 * p.ldo + (m0 + q * 32 + lane);
        ocol_lo = ocol_hi + p.out_plane_stride;
      }
      const int c_lo = ((warp - 2) >> 2) * (BN / 64);  // the two warps of a lane quarter split the 8 column chunks
#pragma unroll 1
      for (int c = c_lo; c < c_lo + BN / 64; c++) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * BN + c * 32), v);
        if (p.out) {
#pragma unroll
          for (int i = 0; i < 32; i++) {  // one coalesced 128-byte store per warp and column
            const float hi = tf32_hi(v[i]);
            const long long o = (long long)(c * 32 + i) * p.ldo;
            ocol_hi[o] = hi;
            ocol_lo[o] = v[i] - hi;
          }
        }
        if (p.ss_part) {
#pragma unroll
          for (int i = 0; i < 32; i++) v[i] *= v[i];
          // butterfly transpose-reduce: after the last step lane c holds the sum over the warp's 32 rows of column c
#pragma unroll
          for (int o = 16; o >= 1; o >>= 1) {
            const bool up = (lane & o) != 0;
#pragma unroll
            for (int i = 0; i < o; i++) {
              const float send = up ? v[i] : v[i + o];
              const float keep = up ? v[i + o] : v[i];
              v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
          }
          redt[q * BN + c * 32 + lane] = v[0];
        }
      }
      if (p.ss_part) {
        asm volatile("bar.sync 1, 256;" ::: "memory");  // the eight epilogue warps only
        const int tt = threadIdx.x - 64;
        double* dst = p.ss_part + ((long long)batch * p.m_tiles + m_tile) * p.ld_ss + n0;
        if (tt < BN)  // 256 epilogue threads, one column each
          dst[tt] = ((double)redt[tt] + (double)redt[BN + tt]) + ((double)redt[2 * BN + tt] + (double)redt[3 * BN + tt]);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// ---- host side: tensor maps ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  return fn;
}

// K-major operand planes [batch][2][rows][cols] (cols = k contiguous): box {16 k, box_rows, 2 planes, 1}, 64B swizzle
inline bool make_map_kmajor(CUtensorMap* map, const float* ptr, int batch, int rows, int cols, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[4] = {(cuuint64_t)cols, (cuuint64_t)rows, 2, (cuuint64_t)batch};
  cuuint64_t strides[3] = {(cuuint64_t)cols * 4, (cuuint64_t)rows * cols * 4, (cuuint64_t)2 * rows * cols * 4};
  cuuint32_t box[4] = {BK, (cuuint32_t)box_rows, 2, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline cudaError_t launch(const CUtensorMap& mapA, const CUtensorMap& mapB, const Params& p, int n_tiles, int batch,
                          cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  gemm_tf32x3_kernel<<<dim3((p.m_tiles + 1) / 2, n_tiles, batch), THREADS, SMEM_BYTES, st>>>(mapA, mapB, p);
  return cudaGetLastError();
}

// ---- FP64 -> hi / lo FP32 planes ---------------------------------------------------------------------------------
// dst [batch][2][rows_p][cols_p] <- src [batch][rows][cols] (zero padded).  grid (ceil(cols_p/256), rows_p, batch)
// lower_only: entries above the diagonal are written as zero whatever the source holds there.
__global__ void split_planes_kernel(const double* __restrict__ src, int rows, int cols, long long src_batch_stride,
                                    int src_ld, float* __restrict__ dst, int rows_p, int cols_p, int lower_only) {
  const int c = blockIdx.x * 256 + threadIdx.x, r = blockIdx.y, b = blockIdx.z;
  if (c >= cols_p) return;
  double x = 0.0;
  if (r < rows && c < cols && !(lower_only && c > r)) x = src[(long long)b * src_batch_stride + (long long)r * src_ld + c];
  const float hi = tf32_hi((float)x);
  const float lo = (float)(x - (double)hi);
  float* d = dst + ((long long)b * 2 * rows_p + r) * cols_p + c;
  d[0] = hi;
  d[(long long)rows_p * cols_p] = lo;
}

}  // namespace t32
}  // namespace mogpe
