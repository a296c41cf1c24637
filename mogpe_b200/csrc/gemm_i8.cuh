// FP64-accurate batched GEMM on the sm_100a INT8 tensor cores (integer slicing, "Ozaki scheme"), for the same products as
// gemm_tf32.cuh:   D[m][n] = sum_k Aop[m][k] * Bop[n][k],  both operands K-major.
//
// tcgen05.mma has no f64 kind and the FP64 DMMA pipe peaks at 37 TFLOP/s on this part, but kind::i8 multiplies signed bytes
// with EXACT int32 accumulation in TMEM at 4.5 POP/s.  Every FP64 operand row is written as
//      x = 2^e * sum_{p=0}^{NS-1} q_p * 2^(-7 (p + 1)),    q_p in [-64, 64]  (signed bytes),   |x| <= 2^(e-1)
// (e: one power-of-two scale per operand row; successive round-to-nearest residuals), and the product is
//      D[m][n] = 2^(ea[m] + eb[n]) * sum_{d=0}^{NS-1} 2^(-7 (d + 2)) * T_d[m][n],    T_d = sum_{i+j=d} sum_k qa_i[m][k] qb_j[n][k]
// with every T_d an exact integer (|T_d| <= (d+1) K 2^12 << 2^31).  Only the slice pairs with i + j >= NS are dropped:
// an error of ~ (NS+1) K 2^(-7 NS) relative to rowmax(A) * rowmax(B), i.e. 1e-13 for NS = 8, K = 1024 -- FP64-class, and
// unlike a floating-point accumulation it does not grow with cancellation inside the sum.
// Cost: NS (NS + 1) / 2 = 36 slice pairs per 32-wide k-step (INT8 K = 32 per instruction), issued as 12 wide MMAs (see the
// MMA loop), against 4 DMMA k-steps of 8.
//
// Tile: D[128 x 64] per CTA, the NS accumulators T_0 .. T_7 occupy all 512 TMEM columns (8 x 64).  Pipeline stage = 64
// k-values (64-byte rows, 64B swizzle): A [NS planes][128 rows][64 B] + B [NS][64 rows][64 B] = 96 KB, 2 stages; each stage
// feeds 2 x 36 MMAs (>= 2300 tensor cycles), so two stages cover the TMA latency.  Warp roles as in gemm_tf32x3_kernel, but
// eight epilogue warps (two per TMEM lane quarter, one 32-column half each): the epilogue cannot overlap the main loop
// (the accumulators fill TMEM), so it is made short instead.
// Epilogue: T_d -> FP64 (tcgen05.ld, 8 int->double conversions per element), row / column scales, column sums of squares in
// FP64 (butterfly transpose-reduce), and optionally the tile re-sliced into NS byte planes, transposed ([n][m]), as the
// B operand of the next product.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "gemm_tf32.cuh"  // mbarrier / TMA / descriptor helpers, tensor-map encoder

namespace mogpe {
namespace oz {

// NS = byte slices per operand (template parameter): 8 -> 56 significant bits, FP64-class (the 1e-8 tolerance); 6 -> 42 bits,
// ~1e-7-class (the 1e-4 "fast mode" tolerance with a wide margin even for cond(L) ~ 1e3) at 21 instead of 36 slice pairs.
constexpr int MAX_NS = 8;
constexpr int BM = 128, BN = 64, BKB = 64, UKB = 32, STAGES = 2;
constexpr int A_PLANE_BYTES = BM * BKB;            // 8 KB
constexpr int B_PLANE_BYTES = BN * BKB;            // 4 KB
constexpr int RED_BYTES = 4 * BN * 8;              // cross-warp column sums (FP64)
template <int NS>
struct Cfg {
  static constexpr int A_STAGE_BYTES = NS * A_PLANE_BYTES;  // 64 KB for NS = 8
  static constexpr int B_STAGE_BYTES = NS * B_PLANE_BYTES;  // 32 KB
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + RED_BYTES + 1024 + 256;
  static constexpr int H = NS / 2;  // accumulators T_0 .. T_(H-1) and T_H .. T_(NS-1) are combined in two int64 Horner sums
};
constexpr int THREADS = 320;  // warp 0: TMA, warp 1: MMA, warps 2-9: epilogue (two per TMEM lane quarter, 32 columns each)
constexpr int TMEM_COLS = 512;  // NS accumulators of 64 columns (allocations are powers of two)

struct Params {
  int m_tiles;  // row tiles of 128
  int n_tiles;  // column tiles of 64 (set by launch)
  int batch;    // (set by launch)
  int K;        // reduction length (multiple of 64)
  int kmode;    // t32::K_FULL / K_LOWER / K_UPPER
  const double* a_scale;     // [A batches][m_tiles * 128]: 2^ea of each operand row
  long long a_scale_stride;  // doubles per A batch
  const double* b_scale;     // [B batches] (device): 2^eb, one scale for all rows of a B operand batch
  int8_t* out;               // optional re-sliced TRANSPOSED result planes [batch][NS][n][ldo]
  long long out_batch_stride, out_plane_stride;  // bytes
  int ldo;
  const double* out_scale;   // [batch] (device): 2^e of the re-sliced result (|D| / 2^e <= 1/2)
  double* ss_part;           // optional [batch][m_tiles][ld_ss] column sums of squares per row tile
  int ld_ss;
  // optional FP64 result, ADDED (atomicAdd) into out64[mapC[batch]][m][n] (not transposed): the caller pre-sets it (zeros, or
  // the value to accumulate onto).  tril64: only elements with n <= m; tiles wholly above the diagonal are skipped.
  // ksplit > 1: the k range of every tile is cut into pieces handled as separate tiles (few large-K tiles otherwise).
  double* out64;
  long long out64_batch_stride;
  int ld64;
  double alpha64;
  int tril64;
  int ksplit;
  const double* b_row_scale;       // optional [B batches][b_row_scale_stride]: 2^e per ROW of the B operand (instead of b_scale)
  long long b_row_scale_stride;
  int* err_flag;
  int mapA[64], mapB[64], mapC[64];
};

__device__ __forceinline__ constexpr uint32_t make_idesc_i8(int n) {
  return (2u << 4)      // D = S32
         | (1u << 7)    // A = signed int8
         | (1u << 10)   // B = signed int8
         | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// Called by ALL lanes of the (converged) MMA warp: elect.sync picks the issuing lane, so the SASS is one predicated UTCIMMA
// instead of the per-thread ELECT loop that a divergent `if (lane == 0)` region compiles to (these MMAs are only 32-48 cycles
// long: the issue path must be short).
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(t32::smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ void tmem_ld32_i(uint32_t taddr, int* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// Tile decode shared by the three warp roles.  Tiles are numbered with the row tile fastest, so the CTAs that run at the
// same time work on neighbouring tiles and share their B tiles in L2.
struct TileInfo {
  int m_tile, n_tile, batch, m0, n0, k_begin, num_kb;
  bool skip;  // tile wholly above the diagonal of a lower-triangular result
};
__device__ __forceinline__ TileInfo decode_tile(const Params& p, int t) {
  using namespace t32;
  TileInfo ti;
  const int ks = p.ksplit > 1 ? p.ksplit : 1;
  const int split = t % ks;
  t /= ks;
  if (p.tril64) {
    // lower-triangular result: only the (row tile m, column tile n) pairs with 64 n <= 128 m + 127, i.e. n <= 2 m + 1, are
    // enumerated: pair index j in [0, m_tiles (m_tiles + 1)), row tile m holds 2 (m + 1) column tiles
    const int per_batch = p.m_tiles * (p.m_tiles + 1);
    const int j = t % per_batch;
    int m = (int)((sqrtf(4.0f * (float)j + 1.0f) - 1.0f) * 0.5f);
    while (m * (m + 1) > j) m--;
    while ((m + 1) * (m + 2) <= j) m++;
    ti.m_tile = m;
    ti.n_tile = j - m * (m + 1);
    ti.batch = t / per_batch;
  } else {
    const int mr = t % p.m_tiles, rest = t / p.m_tiles;
    ti.m_tile = (p.kmode == K_UPPER) ? mr : p.m_tiles - 1 - mr;
    ti.n_tile = rest % p.n_tiles;
    ti.batch = rest / p.n_tiles;
  }
  ti.m0 = ti.m_tile * BM;
  ti.n0 = ti.n_tile * BN;
  int k_begin = 0, k_end = p.K;
  if (p.kmode == K_LOWER) k_end = min(p.K, ti.m0 + BM);
  if (p.kmode == K_UPPER) k_begin = ti.m0 / BKB * BKB;
  int nkb = (k_end - k_begin) / BKB;
  if (ks > 1) {  // this piece of the k range
    const int per = (nkb + ks - 1) / ks, b0 = split * per, b1 = min(nkb, b0 + per);
    k_begin += b0 * BKB;
    nkb = max(0, b1 - b0);
  }
  ti.k_begin = k_begin;
  ti.num_kb = nkb;
  ti.skip = nkb == 0;
  return ti;
}

// PERSISTENT kernel: gridDim.x CTAs (one per SM) loop over tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...  TMEM, the
// barriers and the tensor-map prefetch are set up once, and the TMA producer runs ahead into the next tile while the
// epilogue of the current one drains TMEM, so the per-tile prologue (first TMA round trip: ~4k of ~25k cycles) is hidden.
// The host picks gridDim.x coprime to m_tiles: every CTA then cycles through all row tiles (triangular operands give them
// different k lengths) and the static schedule is balanced.
template <int NS>
__global__ void __launch_bounds__(THREADS, 1) gemm_i8_sliced_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                     const __grid_constant__ CUtensorMap mapB,
                                                                     const __grid_constant__ Params p) {
  using namespace t32;
  constexpr int A_STAGE_BYTES = Cfg<NS>::A_STAGE_BYTES, STAGE_BYTES = Cfg<NS>::STAGE_BYTES, H = Cfg<NS>::H;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  double* red = reinterpret_cast<double*>(gbase + STAGES * STAGE_BYTES);  // [4][BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(gbase + STAGES * STAGE_BYTES + RED_BYTES);
  uint64_t* full = bars;                     // [STAGES] TMA -> MMA
  uint64_t* empty = bars + STAGES;           // [STAGES] MMA -> TMA
  uint64_t* acc_full = bars + 2 * STAGES;    // MMA -> epilogue: the accumulators of this tile are complete
  uint64_t* acc_empty = bars + 2 * STAGES + 1;  // epilogue (8 warps) -> MMA: TMEM has been drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = (p.tril64 ? p.m_tiles * (p.m_tiles + 1) : p.m_tiles * p.n_tiles) * p.batch * (p.ksplit > 1 ? p.ksplit : 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 8);
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
    // ===== TMA producer: one running stage counter across tiles =====
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const TileInfo ti = decode_tile(p, t);
        if (ti.skip) continue;
        const int ca = p.mapA[ti.batch], cb = p.mapB[ti.batch];
        for (int kb = 0; kb < ti.num_kb; kb++, it++) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1, p.err_flag);
          mbar_expect_tx(&full[s], STAGE_BYTES);
          const int k0 = ti.k_begin + kb * BKB;
          const uint32_t sa = base + s * STAGE_BYTES, sb = sa + A_STAGE_BYTES;
          tma_load_4d(sa, &mapA, smem_u32(&full[s]), k0, ti.m0, 0, ca);  // {k, m, slice, batch}: all NS planes in one copy
          tma_load_4d(sb, &mapB, smem_u32(&full[s]), k0, ti.n0, 0, cb);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the loop (converged); one elected lane issues =====
    int it = 0, tile_i = -1;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const TileInfo ti = decode_tile(p, t);
      if (ti.skip) continue;
      tile_i++;
      if (tile_i > 0) {  // the epilogue must have drained the previous tile's accumulators
        mbar_wait(acc_empty, (uint32_t)((tile_i - 1) & 1), p.err_flag);
        tc_fence_after();
      }
      for (int kb = 0; kb < ti.num_kb; kb++, it++) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&full[s], ph, p.err_flag);
        tc_fence_after();
        const uint32_t sa = base + s * STAGE_BYTES, sb = sa + A_STAGE_BYTES;
        // K-major, 64B swizzle: 8-row groups 512 B apart; planes and the 32-byte k-step only move the start address field
        const uint64_t ad0 = make_smem_desc(sa, 16, 8 * BKB, 4), bd0 = make_smem_desc(sb, 16, 8 * BKB, 4);
        const uint32_t first = kb == 0 ? 0u : 1u;
        // Slice i of A meets slices j = 0 .. NS-1-i of B, and pair (i, j) belongs to accumulator T_(i+j) = TMEM columns
        // (i + j) * 64.  The B planes are contiguous in shared memory (plane j = rows 64 j .. 64 j + 63 of one tall K-major
        // tile), so ONE wide MMA  A_i x [B_0; ...; B_(NS-1-i)]^T  with N = 64 (NS - i) columns starting at TMEM column 64 i
        // covers all pairs of slice i (split in two when N > 256): 12 MMAs per k-step instead of 36, and A_i is read from
        // shared memory once instead of NS - i times (the 36-MMA version was bound by that re-read, 63 instead of 32 cycles
        // per 64 columns).  Slice 0 goes first and overwrites every accumulator on the first k-step of a tile.
#pragma unroll
        for (int j = 0; j < BKB / UKB; j++) {
#pragma unroll
          for (int i = 0; i < NS; i++) {
            const uint64_t ad = ad0 + (uint64_t)((i * A_PLANE_BYTES + j * UKB) >> 4);
#pragma unroll
            for (int c0 = 0; c0 < BN * (NS - i); c0 += 256) {
              const int n1 = BN * (NS - i) - c0 < 256 ? BN * (NS - i) - c0 : 256;
              const uint64_t bd = bd0 + (uint64_t)((c0 * BKB + j * UKB) >> 4);
              umma_i8(tmem_base + (uint32_t)(i * BN + c0), ad, bd, make_idesc_i8(n1), (j == 0 && i == 0) ? first : 1u);
            }
          }
        }
        __syncwarp();
        umma_commit_elect(&empty[s]);
      }
      umma_commit_elect(acc_full);
    }
  } else {
    // ===== epilogue: warps 2..9, TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 =====
    const int q = warp & 3;
    const int c = (warp - 2) >> 2;
    int tile_i = -1;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const TileInfo ti = decode_tile(p, t);
      if (ti.skip) continue;
      tile_i++;
      const int batch = ti.batch, m_tile = ti.m_tile, n0 = ti.n0;
      mbar_wait(acc_full, (uint32_t)(tile_i & 1), p.err_flag);
      tc_fence_after();
      const int row = ti.m0 + q * 32 + lane;
      // 2^(ea + eb) * 2^(-14): T_d carries 2^(-7 (d + 2))
      const double rs = p.a_scale[(long long)p.mapA[batch] * p.a_scale_stride + row] *
                        (p.b_row_scale ? 1.0 : p.b_scale[p.mapB[batch]]) * (1.0 / 16384.0);
      int8_t* ocol = nullptr;
      if (p.out) ocol = p.out + (long long)batch * p.out_batch_stride + (long long)n0 * p.ldo + row;
      const double oinv = p.out ? 1.0 / p.out_scale[batch] : 0.0;
      // sum_d T_d 128^-d, exactly: two int64 Horner sums (T_0..T_3 and T_4..T_7; |sum| < 2^50), two int -> double
      // conversions per element instead of eight
      double v[32];
      {
        long long acc[32];
        int tt[32];
        tmem_ld32_i(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), tt);
#pragma unroll
        for (int i = 0; i < 32; i++) acc[i] = tt[i];
#pragma unroll 1
        for (int d = 1; d < H; d++) {
          tmem_ld32_i(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(d * BN + c * 32), tt);
#pragma unroll
          for (int i = 0; i < 32; i++) acc[i] = acc[i] * 128 + tt[i];
        }
        constexpr double W_HI = 1.0 / (double)(1ull << (7 * (H - 1)));   // 128^-(H-1)
        constexpr double W_LO = 1.0 / (double)(1ull << (7 * (NS - 1)));  // 128^-(NS-1)
#pragma unroll
        for (int i = 0; i < 32; i++) v[i] = (double)acc[i] * W_HI;
        tmem_ld32_i(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(H * BN + c * 32), tt);
#pragma unroll
        for (int i = 0; i < 32; i++) acc[i] = tt[i];
#pragma unroll 1
        for (int d = H + 1; d < NS; d++) {
          tmem_ld32_i(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(d * BN + c * 32), tt);
#pragma unroll
          for (int i = 0; i < 32; i++) acc[i] = acc[i] * 128 + tt[i];
        }
#pragma unroll
        for (int i = 0; i < 32; i++) v[i] = fma((double)acc[i], W_LO, v[i]) * rs;
      }
      // every accumulator value of this warp is in registers: TMEM may be overwritten by the next tile's MMAs
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(acc_empty)) : "memory");
      if (p.b_row_scale) {
        const double* bs = p.b_row_scale + (long long)p.mapB[batch] * p.b_row_scale_stride + n0 + c * 32;
#pragma unroll
        for (int i = 0; i < 32; i++) v[i] *= bs[i];
      }
      if (p.out64) {
        double* o = p.out64 + (long long)p.mapC[batch] * p.out64_batch_stride + (long long)row * p.ld64 + n0 + c * 32;
#pragma unroll
        for (int i = 0; i < 32; i++)
          if (!p.tril64 || n0 + c * 32 + i <= row) atomicAdd(o + i, p.alpha64 * v[i]);
      }
      if (p.out) {
#pragma unroll
        for (int i = 0; i < 32; i++) {
          double r = v[i] * oinv;  // |r| <= 1/2
          int8_t* o = ocol + (long long)(c * 32 + i) * p.ldo;
#pragma unroll
          for (int s = 0; s < NS; s++) {
            // round to nearest with the 1.5 * 2^52 trick: the low word of y holds the integer, no conversion instructions
            const double x = r * 128.0;
            const double y = x + 6755399441055744.0;
            r = x - (y - 6755399441055744.0);
            o[(long long)s * p.out_plane_stride] = (int8_t)__double2loint(y);
          }
        }
      }
      if (p.ss_part) {
#pragma unroll
        for (int i = 0; i < 32; i++) v[i] *= v[i];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
          const bool up = (lane & o) != 0;
#pragma unroll
          for (int i = 0; i < o; i++) {
            const double send = up ? v[i] : v[i + o];
            const double keep = up ? v[i + o] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
          }
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");  // the previous tile's sums have been read by everyone
        red[q * BN + c * 32 + lane] = v[0];
        asm volatile("bar.sync 1, 256;" ::: "memory");  // the eight epilogue warps only
        const int e = threadIdx.x - 64;
        if (e < BN) {
          double* dst = p.ss_part + ((long long)batch * p.m_tiles + m_tile) * p.ld_ss + n0;
          dst[e] = (red[e] + red[BN + e]) + (red[2 * BN + e] + red[3 * BN + e]);
        }
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// byte planes [batch][NS][rows][cols] (cols = k contiguous): box {64 k, box_rows, NS planes, 1}, 64B swizzle
inline bool make_map_i8(CUtensorMap* map, const int8_t* ptr, int batch, int rows, int cols, int box_rows, int NS = MAX_NS) {
  t32::EncodeTiledFn fn = t32::encode_fn();
  if (!fn) return false;
  cuuint64_t dims[4] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)NS, (cuuint64_t)batch};
  cuuint64_t strides[3] = {(cuuint64_t)cols, (cuuint64_t)rows * cols, (cuuint64_t)NS * rows * cols};
  cuuint32_t box[4] = {BKB, (cuuint32_t)box_rows, (cuuint32_t)NS, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<int8_t*>(ptr), dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int NS = MAX_NS>
inline cudaError_t launch(const CUtensorMap& mapA, const CUtensorMap& mapB, const Params& p, int n_tiles, int batch,
                          cudaStream_t st) {
  constexpr int SMEM_BYTES = Cfg<NS>::SMEM_BYTES;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_i8_sliced_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms < 1) sms = 148;
  }
  // persistent grid: one CTA per SM, count coprime to m_tiles so that every CTA cycles through all row tiles
  auto gcd = [](int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; };
  const long long total = (long long)(p.tril64 ? p.m_tiles * (p.m_tiles + 1) : p.m_tiles * n_tiles) * batch * (p.ksplit > 1 ? p.ksplit : 1);
  int grid = (int)(total < sms ? total : sms);
  while (grid > 1 && total > grid && gcd(grid, p.m_tiles) != 1) grid--;
  Params q = p;
  q.n_tiles = n_tiles;
  q.batch = batch;
  gemm_i8_sliced_kernel<NS><<<grid, THREADS, SMEM_BYTES, st>>>(mapA, mapB, q);
  return cudaGetLastError();
}

// ---- slicing of FP64 operands -------------------------------------------------------------------------------------
// 2^e with 2^(e-1) >= x (so that |x| / 2^e <= 1/2); 1 for x == 0
__host__ __device__ inline double pow2_scale(double maxabs) {
  if (!(maxabs > 0.0)) return 1.0;
  int e;
  frexp(maxabs, &e);          // maxabs = f * 2^e, f in [0.5, 1)
  return ldexp(1.0, e + 1);   // maxabs < 2^e  ->  maxabs / 2^(e+1) < 1/2
}

template <int NS = MAX_NS>
__device__ __forceinline__ void slice8(double x, double inv_scale, int8_t* q) {
  double r = x * inv_scale;
#pragma unroll
  for (int s = 0; s < NS; s++) {
    // round to nearest with the 1.5 * 2^52 trick: the low word of y holds the integer, no conversion instructions
    const double t = r * 128.0;
    const double y = t + 6755399441055744.0;
    r = t - (y - 6755399441055744.0);
    q[s] = (int8_t)__double2loint(y);
  }
}

// Row scales of src(r, :) (or of src(:, r) when transposed): scale[b][r] = pow2_scale(max |.|), 1 for padded rows.
// lower_only: entries above the diagonal of src are ignored (treated as zero).   grid (rows_p, batch), block 128
__global__ void __launch_bounds__(128) row_scale_kernel(const double* __restrict__ src, int n, long long src_batch_stride,
                                                        int src_ld, const int* __restrict__ map, int transposed,
                                                        int lower_only, double* __restrict__ scale, int rows_p) {
  __shared__ double red[4];
  const int r = blockIdx.x, b = blockIdx.y, sb = map ? map[b] : b;
  double m = 0.0;
  if (r < n) {
    const double* s = src + (long long)sb * src_batch_stride;
    for (int c = threadIdx.x; c < n; c += 128) {
      const int rr = transposed ? c : r, cc = transposed ? r : c;  // element (r, c) of the operand = src(rr, cc)
      if (lower_only && cc > rr) continue;
      m = fmax(m, fabs(s[(long long)rr * src_ld + cc]));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) scale[(long long)b * rows_p + r] = pow2_scale(fmax(fmax(red[0], red[1]), fmax(red[2], red[3])));
}

// dst [batch][NS][n_p][n_p] byte planes of the operand (r, c) -> src(r, c) or src(c, r) (transposed), zero padded, each
// row sliced against its own scale.   grid (n_p/32, n_p/32, batch), block (32, 8)
template <int NS>
__global__ void slice_square_kernel(const double* __restrict__ src, int n, long long src_batch_stride, int src_ld,
                                    const int* __restrict__ map, int transposed, int lower_only,
                                    const double* __restrict__ scale, int8_t* __restrict__ dst, int n_p) {
  __shared__ double tile[32][33];
  const int b = blockIdx.z, sb = map ? map[b] : b;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const double* s = src + (long long)sb * src_batch_stride;
  for (int j = threadIdx.y; j < 32; j += 8) {
    // coalesced read of src: operand rows r0.. / cols c0..  = src rows (transposed ? c0.. : r0..)
    const int sr = (transposed ? c0 : r0) + j, sc = (transposed ? r0 : c0) + threadIdx.x;
    double x = 0.0;
    if (sr < n && sc < n && !(lower_only && sc > sr)) x = s[(long long)sr * src_ld + sc];
    tile[j][threadIdx.x] = x;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    const double x = transposed ? tile[threadIdx.x][j] : tile[j][threadIdx.x];
    int8_t qv[NS];
    slice8<NS>(x, 1.0 / scale[(long long)b * n_p + r], qv);
    int8_t* d = dst + ((long long)b * NS * n_p + r) * n_p + c;
#pragma unroll
    for (int p2 = 0; p2 < NS; p2++) d[(long long)p2 * n_p * n_p] = qv[p2];
  }
}

}  // namespace oz
}  // namespace mogpe

// ---- slicing of [rows][n] matrices along n (the reduction index of the M x M results of the backward) -------------------
namespace mogpe {
namespace oz {

// scale[b][r] = pow2_scale(max_n |src[b][r][n] * cs[n]|) for n < N (cs optional column scale, batch via cs_map); padded rows 1.
// grid (rows_p, batch), block 256
__global__ void __launch_bounds__(256) row_scale_n_kernel(const double* __restrict__ src, int rows, int N,
                                                          long long src_batch_stride, int ld, const int* __restrict__ map,
                                                          const double* __restrict__ cs, long long cs_stride,
                                                          const int* __restrict__ cs_map, double* __restrict__ scale,
                                                          int rows_p) {
  __shared__ double red[8];
  const int r = blockIdx.x, b = blockIdx.y, sb = map ? map[b] : b;
  double m = 0.0;
  if (r < rows) {
    const double* s = src + (long long)sb * src_batch_stride + (long long)r * ld;
    const double* c = cs ? cs + (long long)(cs_map ? cs_map[b] : b) * cs_stride : nullptr;
    for (int n = threadIdx.x; n < N; n += 256) m = fmax(m, fabs(c ? s[n] * c[n] : s[n]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; w++) m = fmax(m, red[w]);
    scale[(long long)b * rows_p + r] = pow2_scale(fmax(m, red[0]));
  }
}

// dst[b][plane][r][n] byte planes of src[b][r][n] * cs[n] / scale[b][r]  (zero for r >= rows or n >= N); every thread slices
// four consecutive n and stores one 32-bit word per plane (coalesced 128 B per warp).  grid (ld/1024, rows_p, batch), block 256
template <int NS>
__global__ void __launch_bounds__(256) slice_rows_n_kernel(const double* __restrict__ src, int rows, int N,
                                                           long long src_batch_stride, int ld, const int* __restrict__ map,
                                                           const double* __restrict__ cs, long long cs_stride,
                                                           const int* __restrict__ cs_map, const double* __restrict__ scale,
                                                           int8_t* __restrict__ dst, int rows_p) {
  const int n4 = (blockIdx.x * 256 + threadIdx.x) * 4, r = blockIdx.y, b = blockIdx.z, sb = map ? map[b] : b;
  if (n4 >= ld) return;
  uint32_t packed[NS];
#pragma unroll
  for (int s = 0; s < NS; s++) packed[s] = 0;
  if (r < rows && n4 < N) {
    const double* sp = src + (long long)sb * src_batch_stride + (long long)r * ld + n4;
    const double* c = cs ? cs + (long long)(cs_map ? cs_map[b] : b) * cs_stride + n4 : nullptr;
    const double inv = 1.0 / scale[(long long)b * rows_p + r];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      double x = (n4 + j < N) ? sp[j] : 0.0;
      if (c) x *= c[j];
      int8_t q[NS];
      slice8<NS>(x, inv, q);
#pragma unroll
      for (int s = 0; s < NS; s++) packed[s] |= (uint32_t)(uint8_t)q[s] << (8 * j);
    }
  }
  int8_t* d = dst + (((long long)b * NS) * rows_p + r) * ld + n4;
#pragma unroll
  for (int s = 0; s < NS; s++) *reinterpret_cast<uint32_t*>(d + (long long)s * rows_p * ld) = packed[s];
}

// scale[b][r] = group_scale[map[b]] for every row (operands whose bound is known per group, e.g. |A| <= sqrt(kvar))
__global__ void fill_row_scale_kernel(const double* __restrict__ group_scale, const int* __restrict__ map,
                                      double* __restrict__ scale, int rows_p) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (r < rows_p) scale[(long long)b * rows_p + r] = group_scale[map ? map[b] : b];
}

}  // namespace oz
}  // namespace mogpe
