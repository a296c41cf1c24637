// Device-side model description shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mogpe_b200.h"

namespace mogpe {

constexpr int MAXQ = MOGPE_MAX_GROUPS;
constexpr int MAXP = MOGPE_MAX_LATENTS;
constexpr int MAXV = MOGPE_MAX_LATENTS * MOGPE_MAX_SAMPLES;  // mean vectors (latent, sample)

// One copy lives in device memory (Handle::d_md); it is re-uploaded when the call mode changes.
struct ModelDev {
  int D, F, K, G, Q, P, Pe, Mp, flavour;
  int group_M[MAXQ];
  int latent_group[MAXP];
  // ---- per call mode (bound, S) ----
  int S;
  int bound;           // MOGPE_BOUND_* or -1 for predict
  int NV;              // total number of mean vectors
  int R[MAXP];         // mean vectors of latent l: S if inducing-sampled, else 1
  int voff[MAXP];      // first vector index of latent l
  int marg[MAXP];      // latent uses the marginal variance (needs q_sqrt^T A)
  int lta_slot[MAXP];  // slot in the LTA buffer, or -1
  int nlta;
  int lta_latent[MAXP];    // slot -> latent
  int lta_group[MAXP];     // slot -> group (GEMM batch maps)
  int vec_latent[MAXV];    // vector -> latent
  int gvec_begin[MAXQ + 1];  // vectors of group g: gvec[gvec_begin[g] .. gvec_begin[g+1])
  int gvec[MAXV];
  int glat_begin[MAXQ + 1];  // latents of group g
  int glat[MAXP];
};

// Offsets into the packed parameter buffer (doubles); mirrors mogpe_layout.
struct Layout {
  long long ls, kvar, Z, q_mu, q_sqrt, mean_c, noise, total;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace mogpe
