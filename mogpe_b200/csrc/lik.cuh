// Fused gating / expert-likelihood / mixture log-sum pass: one thread per data point computes the
// per-point variational expectation AND its adjoints w.r.t. every latent mean / variance in one sweep.
//
// Reference semantics:
//   further_gating: experts via inducing samples (variance without q_sqrt), gating via marginal h samples
//   tight:          gating also via inducing samples + analytic Bernoulli expectation
//   further:        experts and gating via marginal samples; density scale = noise variance
// The reference evaluates log(sum_k p_k pi_k) in probability space; this kernel evaluates the same
// quantity in log space (log-sum-exp), equal to rounding wherever the reference is finite (quirk 6).
#pragma once
#include "model_dev.cuh"

namespace mogpe {

#define MOGPE_LOG_2PI 1.8378770664093454835606594728112
#define MOGPE_INV_SQRT2 0.70710678118654752440084436210485
#define MOGPE_INV_SQRT_2PI 0.39894228040143267793994605993438

// gpflow.likelihoods.utils.inv_probit with its 1e-3 jitter (quirk 3)
__device__ __forceinline__ double inv_probit(double x) {
  return 0.5 * (1.0 + erf(x * MOGPE_INV_SQRT2)) * (1.0 - 2e-3) + 1e-3;
}
__device__ __forceinline__ double inv_probit_grad(double x) {
  return (1.0 - 2e-3) * MOGPE_INV_SQRT_2PI * exp(-0.5 * x * x);
}

struct LikArgs {
  const ModelDev* md;
  const double* params;
  Layout lo;
  const double* mean;    // [NV][Bp]
  const double* var0ss;  // [Q][Bp]
  const double* varq;    // [P][Bp]
  const double* Y;       // [N][F]
  const double* eps_h;   // [S][N][G]
  const double* eps_f;   // [S][N][Pe]
  const double* eps_mc;  // [S][R][N][G] Softmax Monte-Carlo draws (tight bound, G > 1)
  int num_mc;
  int N, Bp;
  double scale;
  double* dmean;  // [NV][Bp]
  double* dfvar;  // [P][Bp]
  double* out;    // out[1] += scaled data term
  double* grads;  // nullable
};

// per-thread scratch in shared memory, laid out [slot][thread] (conflict-free)
struct Scratch {
  double* base;
  int stride;
  __device__ __forceinline__ double& operator()(int slot) const { return base[slot * stride]; }
};

// Gating log-probabilities for sample s -> lpi[k] (K entries), marginal-sample mode.
// Also used in backward to recompute pi.
__device__ __forceinline__ void gating_forward_sample(const ModelDev* md, const LikArgs& a, int n, int s, bool valid,
                                                      const Scratch& lpi, int lpi0) {
  const int G = md->G, K = md->K, Pe = md->Pe;
  const bool keras = md->flavour == MOGPE_FLAVOUR_KERAS;
  if (G == 1) {
    const int l = Pe, g = md->latent_group[l];
    double hm = 0, hv = 1, e = 0;
    if (valid) {
      hm = a.mean[(long long)md->voff[l] * a.Bp + n];
      hv = a.params[a.lo.kvar + g] - a.var0ss[(long long)g * a.Bp + n] + a.varq[(long long)l * a.Bp + n];
      e = a.eps_h[((long long)s * a.N + n)];
    }
    const double sd = keras ? sqrt(hv) : hv;
    const double h = hm + sd * e;
    const double p0 = inv_probit(h);
    const double p1 = keras ? inv_probit(-h) : 1.0 - p0;
    lpi(lpi0 + 0) = log(p0);
    lpi(lpi0 + 1) = log(p1);
  } else {
    double mx = -1e300;
    for (int k = 0; k < K; k++) {
      const int l = Pe + k, g = md->latent_group[l];
      double h = 0;
      if (valid) {
        const double hm = a.mean[(long long)md->voff[l] * a.Bp + n];
        const double hv = a.params[a.lo.kvar + g] - a.var0ss[(long long)g * a.Bp + n] + a.varq[(long long)l * a.Bp + n];
        const double sd = keras ? sqrt(hv) : hv;
        h = hm + sd * a.eps_h[((long long)s * a.N + n) * G + k];
      }
      lpi(lpi0 + k) = h;
      mx = fmax(mx, h);
    }
    double se = 0;
    for (int k = 0; k < K; k++) se += exp(lpi(lpi0 + k) - mx);
    const double lse = mx + log(se);
    for (int k = 0; k < K; k++) lpi(lpi0 + k) -= lse;
  }
}

// Backward of gating_forward_sample given dlpi[k] (adjoints of log pi_k for sample s).
__device__ __forceinline__ void gating_backward_sample(const ModelDev* md, const LikArgs& a, int n, int s,
                                                       const Scratch& lpi, int lpi0, const Scratch& dlpi, int d0) {
  const int G = md->G, K = md->K, Pe = md->Pe;
  const bool keras = md->flavour == MOGPE_FLAVOUR_KERAS;
  if (G == 1) {
    const int l = Pe, g = md->latent_group[l];
    const double hm = a.mean[(long long)md->voff[l] * a.Bp + n];
    const double hv = a.params[a.lo.kvar + g] - a.var0ss[(long long)g * a.Bp + n] + a.varq[(long long)l * a.Bp + n];
    const double e = a.eps_h[((long long)s * a.N + n)];
    const double sd = keras ? sqrt(hv) : hv;
    const double h = hm + sd * e;
    const double gp = inv_probit_grad(h);
    const double p0 = exp(lpi(lpi0 + 0)), p1 = exp(lpi(lpi0 + 1));
    const double dh = dlpi(d0 + 0) * gp / p0 - dlpi(d0 + 1) * gp / p1;
    a.dmean[(long long)md->voff[l] * a.Bp + n] += dh;
    a.dfvar[(long long)l * a.Bp + n] += dh * e * (keras ? 0.5 / sd : 1.0);
  } else {
    double sd_all = 0;
    for (int k = 0; k < K; k++) sd_all += dlpi(d0 + k);
    for (int k = 0; k < K; k++) {
      const int l = Pe + k, g = md->latent_group[l];
      const double hv = a.params[a.lo.kvar + g] - a.var0ss[(long long)g * a.Bp + n] + a.varq[(long long)l * a.Bp + n];
      const double sd = keras ? sqrt(hv) : hv;
      const double e = a.eps_h[((long long)s * a.N + n) * G + k];
      const double dh = dlpi(d0 + k) - exp(lpi(lpi0 + k)) * sd_all;
      a.dmean[(long long)md->voff[l] * a.Bp + n] += dh;
      a.dfvar[(long long)l * a.Bp + n] += dh * e * (keras ? 0.5 / sd : 1.0);
    }
  }
}

// grid (Bp/blockDim), dynamic smem = (4*S*K + K*F + 6*K)*blockDim doubles.  Threads n >= N write zero adjoints
// (the padded columns feed the B-long GEMMs of the backward pass).
__global__ void lik_kernel(const LikArgs a) {
  extern __shared__ double sm[];
  const ModelDev* md = a.md;
  const int S = md->S, K = md->K, F = md->F, G = md->G, Pe = md->Pe, P = md->P;
  const int bound = md->bound;
  const bool keras = md->flavour == MOGPE_FLAVOUR_KERAS;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = n < a.N;
  const int SK = S * K;
  Scratch lp{sm + threadIdx.x, (int)blockDim.x};  // slots: [0,SK) lp, [SK,2SK) lpi, [2SK,3SK) dlp, [3SK,4SK) dlpi
  const int LP = 0, LPI = SK, DLP = 2 * SK, DLPI = 3 * SK, DN = 4 * SK;  // DN: Pe noise-gradient slots
  const bool want = a.grads != nullptr;

  // zero this column's adjoints
  for (int v = 0; v < md->NV; v++) a.dmean[(long long)v * a.Bp + n] = 0.0;
  for (int l = 0; l < P; l++) a.dfvar[(long long)l * a.Bp + n] = 0.0;

  double val = 0.0;
  const double w_sample = a.scale;

  if (bound == MOGPE_BOUND_FURTHER && !keras) {
    // ---- legacy "further" (mogpe/mixture_of_experts/svgp.py:232-330): per-output mixture log_prob, one
    // sample index shared by gating and experts; f, h sampled with scale = variance (quirk 2).
    const double w = w_sample / S;
    for (int s = 0; s < S; s++) {
      gating_forward_sample(md, a, n, s, valid, lp, LPI);
      for (int k = 0; k < K; k++) lp(DLPI + k) = 0.0;
      for (int f = 0; f < F; f++) {
        double mx = -1e300;
        for (int k = 0; k < K; k++) {
          const int l = k * F + f, g = md->latent_group[l];
          double ln = 0;
          if (valid) {
            const double fm = a.mean[(long long)md->voff[l] * a.Bp + n];
            const double fv = a.params[a.lo.kvar + g] - a.var0ss[(long long)g * a.Bp + n] + a.varq[(long long)l * a.Bp + n];
            const double fs = fm + fv * a.eps_f[((long long)s * a.N + n) * Pe + l];
            const double sc = a.params[a.lo.noise + l];
            const double z = (a.Y[(long long)n * F + f] - fs) / sc;
            ln = -0.5 * z * z - log(sc) - 0.5 * MOGPE_LOG_2PI;
          }
          lp(LP + k) = ln + lp(LPI + k);
          mx = fmax(mx, lp(LP + k));
        }
        double se = 0;
        for (int k = 0; k < K; k++) se += exp(lp(LP + k) - mx);
        const double lse = mx + log(se);
        if (valid) val += w * lse;
        if (want && valid) {
          for (int k = 0; k < K; k++) {
            const int l = k * F + f, g = md->latent_group[l];
            const double r = w * exp(lp(LP + k) - lse);
            const double fm = a.mean[(long long)md->voff[l] * a.Bp + n];
            const double fv = a.params[a.lo.kvar + g] - a.var0ss[(long long)g * a.Bp + n] + a.varq[(long long)l * a.Bp + n];
            const double e = a.eps_f[((long long)s * a.N + n) * Pe + l];
            const double fs = fm + fv * e;
            const double sc = a.params[a.lo.noise + l];
            const double z = (a.Y[(long long)n * F + f] - fs) / sc;
            const double dmu = r * z / sc;
            a.dmean[(long long)md->voff[l] * a.Bp + n] += dmu;
            a.dfvar[(long long)l * a.Bp + n] += dmu * e;
            lp(DN + l) = (s == 0 ? 0.0 : lp(DN + l)) + r * (z * z - 1.0) / sc;  // dnoise
            lp(DLPI + k) += r;
          }
        }
      }
      if (want && valid) gating_backward_sample(md, a, n, s, lp, LPI, lp, DLPI);
    }
    if (want) {
      for (int l = 0; l < Pe; l++) {
        const double s = warp_sum(valid ? lp(DN + l) : 0.0);
        if ((threadIdx.x & 31) == 0) atomicAdd(a.grads + a.lo.noise + l, s);
      }
    }
  } else {
    // ---- S x S structure: log sum_k p_k[s1] pi_k[s2], mean over (s1, s2)
    // 1. expert log densities lp[s][k]
    for (int s = 0; s < S; s++)
      for (int k = 0; k < K; k++) {
        double acc = 0;
        if (valid)
          for (int f = 0; f < F; f++) {
            const int l = k * F + f, g = md->latent_group[l];
            const double fvar0 = a.params[a.lo.kvar + g] - a.var0ss[(long long)g * a.Bp + n];
            double mu, sc;
            if (bound != MOGPE_BOUND_FURTHER) {
              mu = a.mean[(long long)(md->voff[l] + s) * a.Bp + n];
              const double yv = fvar0 + a.params[a.lo.noise + l];
              sc = keras ? yv : sqrt(yv);  // quirk 1
            } else {  // keras further: f_s = mean + sqrt(var) eps, scale = noise variance
              const double fv = fvar0 + a.varq[(long long)l * a.Bp + n];
              mu = a.mean[(long long)md->voff[l] * a.Bp + n] + sqrt(fv) * a.eps_f[((long long)s * a.N + n) * Pe + l];
              sc = a.params[a.lo.noise + l];
            }
            const double z = (a.Y[(long long)n * F + f] - mu) / sc;
            acc += -0.5 * z * z - log(sc) - 0.5 * MOGPE_LOG_2PI;
          }
        lp(LP + s * K + k) = acc;
        lp(DLP + s * K + k) = 0.0;
        lp(DLPI + s * K + k) = 0.0;
      }
    // 2. gating log probabilities lpi[s][k]
    if (bound == MOGPE_BOUND_TIGHT && G > 1) {
      // inducing-sampled gating + Monte-Carlo Softmax expectation: pi_k = mean_r softmax_k(hm + sqrt(hv) eps_r)
      const int HM = DN + Pe, SD = HM + K, PR = SD + K;  // scratch slots: hm[K], sd[K], p_r[K]
      for (int s = 0; s < S; s++) {
        for (int k = 0; k < K; k++) {
          const int l = Pe + k, g = md->latent_group[l];
          lp(HM + k) = valid ? a.mean[(long long)(md->voff[l] + s) * a.Bp + n] : 0.0;
          lp(SD + k) = valid ? sqrt(a.params[a.lo.kvar + g] - a.var0ss[(long long)g * a.Bp + n]) : 1.0;
          lp(LPI + s * K + k) = 0.0;
        }
        for (int r = 0; r < a.num_mc; r++) {
          double mx = -1e300;
          for (int k = 0; k < K; k++) {
            const double e = valid ? a.eps_mc[(((long long)s * a.num_mc + r) * a.N + n) * G + k] : 0.0;
            const double h = lp(HM + k) + lp(SD + k) * e;
            lp(PR + k) = h;
            mx = fmax(mx, h);
          }
          double se = 0;
          for (int k = 0; k < K; k++) {
            lp(PR + k) = exp(lp(PR + k) - mx);
            se += lp(PR + k);
          }
          for (int k = 0; k < K; k++) lp(LPI + s * K + k) += lp(PR + k) / se;
        }
        for (int k = 0; k < K; k++) lp(LPI + s * K + k) = log(lp(LPI + s * K + k) / (double)a.num_mc);
      }
    } else if (bound == MOGPE_BOUND_TIGHT) {
      // inducing-sampled gating + analytic Bernoulli expectation (keras/gating_networks.py:159-171,243-244)
      const int l = Pe, g = md->latent_group[l];
      for (int s = 0; s < S; s++) {
        double t = 0;
        if (valid) {
          const double hm = a.mean[(long long)(md->voff[l] + s) * a.Bp + n];
          const double hv = a.params[a.lo.kvar + g] - a.var0ss[(long long)g * a.Bp + n];
          t = hm / sqrt(1.0 + hv);
        }
        const double p0 = inv_probit(t);
        const double p1 = keras ? inv_probit(-t) : 1.0 - p0;
        lp(LPI + s * K + 0) = log(p0);
        lp(LPI + s * K + 1) = log(p1);
      }
    } else {
      for (int s = 0; s < S; s++) gating_forward_sample(md, a, n, s, valid, lp, LPI + s * K);
    }
    // 3. marginalise the indicator, average samples
    const double w = w_sample / ((double)S * S);
    for (int s1 = 0; s1 < S; s1++)
      for (int s2 = 0; s2 < S; s2++) {
        double mx = -1e300;
        for (int k = 0; k < K; k++) mx = fmax(mx, lp(LP + s1 * K + k) + lp(LPI + s2 * K + k));
        double se = 0;
        for (int k = 0; k < K; k++) se += exp(lp(LP + s1 * K + k) + lp(LPI + s2 * K + k) - mx);
        const double lse = mx + log(se);
        if (valid) val += w * lse;
        if (want)
          for (int k = 0; k < K; k++) {
            const double r = w * exp(lp(LP + s1 * K + k) + lp(LPI + s2 * K + k) - lse);
            lp(DLP + s1 * K + k) += r;
            lp(DLPI + s2 * K + k) += r;
          }
      }
    // 4. backward through the expert densities
    if (want) {
      for (int k = 0; k < K; k++)
        for (int f = 0; f < F; f++) {
          const int l = k * F + f, g = md->latent_group[l];
          double dnoise = 0;
          if (valid) {
            const double fvar0 = a.params[a.lo.kvar + g] - a.var0ss[(long long)g * a.Bp + n];
            const double y = a.Y[(long long)n * F + f];
            if (bound != MOGPE_BOUND_FURTHER) {
              const double yv = fvar0 + a.params[a.lo.noise + l];
              const double sc = keras ? yv : sqrt(yv);
              double dyv = 0;
              for (int s = 0; s < S; s++) {
                const double mu = a.mean[(long long)(md->voff[l] + s) * a.Bp + n];
                const double z = (y - mu) / sc;
                const double d = lp(DLP + s * K + k);
                a.dmean[(long long)(md->voff[l] + s) * a.Bp + n] += d * z / sc;
                dyv += d * (z * z - 1.0) / sc * (keras ? 1.0 : 0.5 / sc);
              }
              a.dfvar[(long long)l * a.Bp + n] += dyv;
              dnoise = dyv;
            } else {
              const double fv = fvar0 + a.varq[(long long)l * a.Bp + n];
              const double sf = sqrt(fv), sc = a.params[a.lo.noise + l];
              const double fm = a.mean[(long long)md->voff[l] * a.Bp + n];
              double dfm = 0, dfv = 0;
              for (int s = 0; s < S; s++) {
                const double e = a.eps_f[((long long)s * a.N + n) * Pe + l];
                const double z = (y - (fm + sf * e)) / sc;
                const double d = lp(DLP + s * K + k);
                const double dmu = d * z / sc;
                dfm += dmu;
                dfv += dmu * e * 0.5 / sf;
                dnoise += d * (z * z - 1.0) / sc;
              }
              a.dmean[(long long)md->voff[l] * a.Bp + n] += dfm;
              a.dfvar[(long long)l * a.Bp + n] += dfv;
            }
          }
          dnoise = warp_sum(dnoise);
          if ((threadIdx.x & 31) == 0) atomicAdd(a.grads + a.lo.noise + l, dnoise);
        }
      // 5. backward through the gating
      if (valid) {
        if (bound == MOGPE_BOUND_TIGHT && G > 1) {
          // d log pi_k = w_k d pi_k, w_k = dlpi_k / pi_k;  d pi_k / d h_j = mean_r p_rk (delta_kj - p_rj)
          const int HM = DN + Pe, SD = HM + K, PR = SD + K, WK = PR + K, DH = WK + K, DS = DH + K;
          for (int s = 0; s < S; s++) {
            for (int k = 0; k < K; k++) {
              const int l = Pe + k, g = md->latent_group[l];
              lp(HM + k) = a.mean[(long long)(md->voff[l] + s) * a.Bp + n];
              lp(SD + k) = sqrt(a.params[a.lo.kvar + g] - a.var0ss[(long long)g * a.Bp + n]);
              lp(WK + k) = lp(DLPI + s * K + k) / exp(lp(LPI + s * K + k));
              lp(DH + k) = 0.0;
              lp(DS + k) = 0.0;
            }
            for (int r = 0; r < a.num_mc; r++) {
              const double* e = a.eps_mc + (((long long)s * a.num_mc + r) * a.N + n) * G;
              double mx = -1e300;
              for (int k = 0; k < K; k++) {
                lp(PR + k) = lp(HM + k) + lp(SD + k) * e[k];
                mx = fmax(mx, lp(PR + k));
              }
              double se = 0;
              for (int k = 0; k < K; k++) {
                lp(PR + k) = exp(lp(PR + k) - mx);
                se += lp(PR + k);
              }
              double t = 0;
              for (int k = 0; k < K; k++) {
                lp(PR + k) /= se;
                t += lp(WK + k) * lp(PR + k);
              }
              for (int k = 0; k < K; k++) {
                const double dh = lp(PR + k) * (lp(WK + k) - t);
                lp(DH + k) += dh;
                lp(DS + k) += dh * e[k];
              }
            }
            for (int k = 0; k < K; k++) {
              const int l = Pe + k;
              a.dmean[(long long)(md->voff[l] + s) * a.Bp + n] += lp(DH + k) / (double)a.num_mc;
              a.dfvar[(long long)l * a.Bp + n] += lp(DS + k) / (double)a.num_mc * 0.5 / lp(SD + k);
            }
          }
        } else if (bound == MOGPE_BOUND_TIGHT) {
          const int l = Pe, g = md->latent_group[l];
          const double hv = a.params[a.lo.kvar + g] - a.var0ss[(long long)g * a.Bp + n];
          const double rs = rsqrt(1.0 + hv);
          double dhv = 0;
          for (int s = 0; s < S; s++) {
            const double hm = a.mean[(long long)(md->voff[l] + s) * a.Bp + n];
            const double t = hm * rs;
            const double gp = inv_probit_grad(t);
            const double p0 = exp(lp(LPI + s * K + 0)), p1 = exp(lp(LPI + s * K + 1));
            const double dt = lp(DLPI + s * K + 0) * gp / p0 - lp(DLPI + s * K + 1) * gp / p1;
            a.dmean[(long long)(md->voff[l] + s) * a.Bp + n] += dt * rs;
            dhv += dt * (-0.5) * hm * rs * rs * rs;
          }
          a.dfvar[(long long)l * a.Bp + n] += dhv;
        } else {
          for (int s = 0; s < S; s++) gating_backward_sample(md, a, n, s, lp, LPI + s * K, lp, DLPI + s * K);
        }
      }
    }
  }

  // data term
  val = warp_sum(val);
  if ((threadIdx.x & 31) == 0) atomicAdd(a.out + 1, val);

  // mean-function constants and kernel variance (through k(x,x) = variance) gradients
  if (want) {
    for (int l = 0; l < P; l++) {
      double dc = 0;
      for (int r = 0; r < md->R[l]; r++) dc += a.dmean[(long long)(md->voff[l] + r) * a.Bp + n];
      dc = warp_sum(dc);
      const double dk = warp_sum(a.dfvar[(long long)l * a.Bp + n]);
      if ((threadIdx.x & 31) == 0) {
        atomicAdd(a.grads + a.lo.mean_c + l, dc);
        atomicAdd(a.grads + a.lo.kvar + md->latent_group[l], dk);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Fast path of lik_kernel for bound = further_gating -- the reference's DEFAULT bound (mosvgpe.py:109-168, legacy
// svgp.py:129-230) and the one bench.py times: experts via inducing samples, gating via marginal h samples -- with up to
// LIK_FG_SMAX samples.  Same arithmetic as lik_kernel, different decomposition: KT = 2^ceil(log2 K) lanes work on ONE data
// point, lane k owns expert k (its F latents) and, with Softmax gating, gating latent k (Bernoulli: the single gating latent
// is evaluated by lanes 0 and 1); the log-sum-exp over experts, the softmax and the adjoint sums are shuffle reductions
// inside the lane group, every adjoint is written exactly once (no zero + read-modify-write sweeps over dmean / dfvar), and
// the per-latent reductions over the data points (mean-function, kernel-variance and noise gradients) go through shared
// memory: one atomic per latent and block.  lik_kernel's one-thread-per-point body took 54 us for 8192 points (latency-bound:
// ~2 warps per SM, every intermediate through a shared-memory scratch and dozens of dependent global read-modify-writes).
// grid (ceil(Bc / (256 / KT))), block 256.
// ------------------------------------------------------------------------------------------------
constexpr int LIK_FG_SMAX = 4;

template <int KT>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
  for (int o = KT / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int KT>
__device__ __forceinline__ double group_max(double v) {
#pragma unroll
  for (int o = KT / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// sum over the data points of a warp for the lanes that own the same expert index (lanes k, k + KT, ...)
template <int KT>
__device__ __forceinline__ double points_sum(double v) {
#pragma unroll
  for (int o = 16; o >= KT; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int KT>
__global__ void __launch_bounds__(256) lik_fg_kernel(const LikArgs a) {
  constexpr int PTS = 256 / KT;  // data points per block
  constexpr int WARPS = 8;
  // per-latent block reductions: [mean_c | kvar | noise] x up to MOGPE_MAX_LATENTS, one slot per warp
  extern __shared__ double red_sm[];  // [WARPS][3][P]
  const ModelDev* md = a.md;
  const int S = md->S, K = md->K, F = md->F, G = md->G, Pe = md->Pe, P = md->P;
  const bool keras = md->flavour == MOGPE_FLAVOUR_KERAS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kk = lane & (KT - 1);
  const int n = blockIdx.x * PTS + tid / KT;
  const bool valid = n < a.N, act = kk < K;
  const bool want = a.grads != nullptr;
  const long long Bp = a.Bp;
  for (int i = tid; i < WARPS * 3 * P; i += 256) red_sm[i] = 0.0;
  __syncthreads();
  double* red = red_sm + (long long)warp * 3 * P;

  // ---- gating: log pi_k[s] from marginal h samples --------------------------------------------------------------
  double lpi[LIK_FG_SMAX], eh[LIK_FG_SMAX];
  double g_sd = 1.0, g_hm = 0.0;
  const int gl = (G == 1) ? Pe : Pe + kk;  // gating latent evaluated by this lane
  const bool g_act = (G == 1) ? (kk < 2) : act;
  if (valid && g_act) {
    const int gg = md->latent_group[gl];
    g_hm = a.mean[(long long)md->voff[gl] * Bp + n];
    const double hv = a.params[a.lo.kvar + gg] - a.var0ss[(long long)gg * Bp + n] + a.varq[(long long)gl * Bp + n];
    g_sd = keras ? sqrt(hv) : hv;  // quirk 2: the legacy flavour samples with scale = variance
  }
#pragma unroll
  for (int s = 0; s < LIK_FG_SMAX; s++) {
    lpi[s] = 0.0;
    eh[s] = 0.0;
    if (s < S) {
      if (valid && g_act) eh[s] = a.eps_h[((long long)s * a.N + n) * G + (G == 1 ? 0 : kk)];
      const double h = g_hm + g_sd * eh[s];
      if (G == 1) {
        const double p0 = inv_probit(h);
        const double p1 = keras ? inv_probit(-h) : 1.0 - p0;
        lpi[s] = log(kk == 0 ? p0 : p1);
      } else {
        const double hh = act ? h : -1e300;
        const double mx = group_max<KT>(hh);
        const double se = group_sum<KT>(act ? exp(hh - mx) : 0.0);
        lpi[s] = hh - (mx + log(se));
      }
    }
  }
  // ---- expert log densities lp_k[s] = sum_f log N(y_f; mean, scale) ------------------------------------------------
  double lp[LIK_FG_SMAX];
#pragma unroll
  for (int s = 0; s < LIK_FG_SMAX; s++) lp[s] = 0.0;
  if (valid && act)
    for (int f = 0; f < F; f++) {
      const int l = kk * F + f, g = md->latent_group[l];
      const double yv = a.params[a.lo.kvar + g] - a.var0ss[(long long)g * Bp + n] + a.params[a.lo.noise + l];
      const double sc = keras ? yv : sqrt(yv);  // quirk 1: the Keras flavour uses the predictive variance as the scale
      const double y = a.Y[(long long)n * F + f], lsc = log(sc) + 0.5 * MOGPE_LOG_2PI;
#pragma unroll
      for (int s = 0; s < LIK_FG_SMAX; s++)
        if (s < S) {
          const double z = (y - a.mean[(long long)(md->voff[l] + s) * Bp + n]) / sc;
          lp[s] += -0.5 * z * z - lsc;
        }
    }
  // ---- marginalise the indicator, average the S x S sample pairs ----------------------------------------------------
  double dlp[LIK_FG_SMAX], dlpi[LIK_FG_SMAX], val = 0.0;
#pragma unroll
  for (int s = 0; s < LIK_FG_SMAX; s++) dlp[s] = dlpi[s] = 0.0;
  const double w = a.scale / ((double)S * S);
#pragma unroll
  for (int s1 = 0; s1 < LIK_FG_SMAX; s1++)
#pragma unroll
    for (int s2 = 0; s2 < LIK_FG_SMAX; s2++)
      if (s1 < S && s2 < S) {
        const double t = act ? lp[s1] + lpi[s2] : -1e300;
        const double mx = group_max<KT>(t);
        const double se = group_sum<KT>(act ? exp(t - mx) : 0.0);
        const double lse = mx + log(se);
        if (valid && kk == 0) val += w * lse;
        if (want && valid && act) {
          const double r = w * exp(t - lse);
          dlp[s1] += r;
          dlpi[s2] += r;
        }
      }
  // ---- adjoints of the expert latents (written once; every lane runs the loop so that the shuffles stay convergent) ---------
  for (int f = 0; f < F; f++) {
    const int l = act ? kk * F + f : 0;
    double dyv = 0.0, dc = 0.0;
    if (valid && act) {
      const int g = md->latent_group[l];
      const double yv = a.params[a.lo.kvar + g] - a.var0ss[(long long)g * Bp + n] + a.params[a.lo.noise + l];
      const double sc = keras ? yv : sqrt(yv);
      const double y = a.Y[(long long)n * F + f];
#pragma unroll
      for (int s = 0; s < LIK_FG_SMAX; s++)
        if (s < S) {
          const double z = (y - a.mean[(long long)(md->voff[l] + s) * Bp + n]) / sc;
          const double dm = dlp[s] * z / sc;
          a.dmean[(long long)(md->voff[l] + s) * Bp + n] = dm;
          dc += dm;
          dyv += dlp[s] * (z * z - 1.0) / sc * (keras ? 1.0 : 0.5 / sc);
        }
    } else if (act) {
      for (int s = 0; s < S; s++) a.dmean[(long long)(md->voff[l] + s) * Bp + n] = 0.0;
    }
    if (act) a.dfvar[(long long)l * Bp + n] = dyv;
    if (want) {
      const double sc_ = points_sum<KT>(dc), sk = points_sum<KT>(dyv);
      if (lane < KT && act) {
        red[0 * P + l] += sc_;
        red[1 * P + l] += sk;
        red[2 * P + l] += sk;  // d / d noise = d / d y_var
      }
    }
  }
  // ---- adjoints of the gating latents --------------------------------------------------------------------------------------
  {
    double dhm = 0.0, dhv = 0.0;
    if (G == 1) {
#pragma unroll
      for (int s = 0; s < LIK_FG_SMAX; s++)
        if (s < S) {
          const double other = __shfl_xor_sync(0xffffffffu, dlpi[s], 1);  // lane 0 <-> lane 1 of the group
          const double d0 = kk == 0 ? dlpi[s] : other, d1 = kk == 0 ? other : dlpi[s];
          const double h = g_hm + g_sd * eh[s];
          const double gp = inv_probit_grad(h);
          const double p0 = inv_probit(h), p1 = keras ? inv_probit(-h) : 1.0 - p0;
          const double dh = d0 * gp / p0 - d1 * gp / p1;
          dhm += dh;
          dhv += dh * eh[s] * (keras ? 0.5 / g_sd : 1.0);
        }
    } else {
#pragma unroll
      for (int s = 0; s < LIK_FG_SMAX; s++)
        if (s < S) {
          const double sd_all = group_sum<KT>(dlpi[s]);
          const double dh = act ? dlpi[s] - exp(lpi[s]) * sd_all : 0.0;
          dhm += dh;
          dhv += dh * eh[s] * (keras ? 0.5 / g_sd : 1.0);
        }
    }
    const bool owner = (G == 1) ? (kk == 0) : act;
    if (owner) {
      if (!valid) dhm = dhv = 0.0;
      a.dmean[(long long)md->voff[gl] * Bp + n] = dhm;
      a.dfvar[(long long)gl * Bp + n] = dhv;
    }
    if (want) {
      const double sm = points_sum<KT>(owner ? dhm : 0.0), sv = points_sum<KT>(owner ? dhv : 0.0);
      if (lane < KT && owner) {
        red[0 * P + gl] += sm;
        red[1 * P + gl] += sv;
      }
    }
  }
  // ---- block reductions -> one atomic per latent -------------------------------------------------------------------------
  val = warp_sum(val);
  __syncthreads();
  if (want)
    for (int i = tid; i < 3 * P; i += 256) {
      double t = 0.0;
#pragma unroll
      for (int wq = 0; wq < WARPS; wq++) t += red_sm[(long long)wq * 3 * P + i];
      const int which = i / P, l = i % P;
      if (t != 0.0) {
        if (which == 0) atomicAdd(a.grads + a.lo.mean_c + l, t);
        else if (which == 1) atomicAdd(a.grads + a.lo.kvar + md->latent_group[l], t);
        else if (l < Pe) atomicAdd(a.grads + a.lo.noise + l, t);
      }
    }
  __shared__ double vred[WARPS];
  if (lane == 0) vred[warp] = val;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int wq = 0; wq < WARPS; wq++) t += vred[wq];
    atomicAdd(a.out + 1, t);
  }
}

// ------------------------------------------------------------------------------------------------
// Prediction outputs (Appendix A.6-A.7).  One thread per test point.
// ------------------------------------------------------------------------------------------------
struct PredArgs {
  const ModelDev* md;
  const double* params;
  Layout lo;
  const double* mean;    // [P][Bp]  (R = 1 for every latent)
  const double* var0ss;  // [Q][Bp]
  const double* varq;    // [P][Bp]
  int n0, nchunk, Bp;
  long long Ntot;
  const double* eps_mc;  // [R][mc_ld][K] (rows n - mc_n0) or null
  long long mc_ld, mc_n0;
  int num_mc;
  double* f_mean;  // [Ntot][P] or null
  double* f_var;
  double* mix;     // [Ntot][K] or null
  double* y_mean;  // [Ntot][F] or null
  double* y_var;
};

__global__ void predict_out_kernel(const PredArgs a) {
  extern __shared__ double sm[];
  const ModelDev* md = a.md;
  const int K = md->K, F = md->F, G = md->G, Pe = md->Pe, P = md->P;
  const bool keras = md->flavour == MOGPE_FLAVOUR_KERAS;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.nchunk) return;
  const long long n = (long long)a.n0 + i;
  Scratch pi{sm + threadIdx.x, (int)blockDim.x};  // K slots: mixing probs; K more: softmax scratch
  auto fmean = [&](int l) { return a.mean[(long long)md->voff[l] * a.Bp + i]; };
  auto fvar = [&](int l) {
    const int g = md->latent_group[l];
    return a.params[a.lo.kvar + g] - a.var0ss[(long long)g * a.Bp + i] + a.varq[(long long)l * a.Bp + i];
  };
  if (a.f_mean)
    for (int l = 0; l < P; l++) a.f_mean[n * P + l] = fmean(l);
  if (a.f_var)
    for (int l = 0; l < P; l++) a.f_var[n * P + l] = fvar(l);
  if (!a.mix && !a.y_mean && !a.y_var) return;
  if (G == 1) {
    const double t = fmean(Pe) / sqrt(1.0 + fvar(Pe));
    const double p0 = inv_probit(t);
    pi(0) = p0;
    pi(1) = keras ? inv_probit(-t) : 1.0 - p0;
  } else {
    // gpflow Softmax.predict_mean_and_var: Monte-Carlo mean of softmax(mu + sqrt(v) eps) over num_mc draws
    for (int k = 0; k < K; k++) pi(k) = 0.0;
    for (int r = 0; r < a.num_mc; r++) {
      double mx = -1e300;
      for (int k = 0; k < K; k++) {
        const double h = fmean(Pe + k) + sqrt(fvar(Pe + k)) * a.eps_mc[((long long)r * a.mc_ld + (n - a.mc_n0)) * K + k];
        pi(K + k) = h;
        mx = fmax(mx, h);
      }
      double se = 0;
      for (int k = 0; k < K; k++) {
        pi(K + k) = exp(pi(K + k) - mx);
        se += pi(K + k);
      }
      for (int k = 0; k < K; k++) pi(k) += pi(K + k) / se;
    }
    for (int k = 0; k < K; k++) pi(k) /= (double)a.num_mc;
  }
  if (a.mix)
    for (int k = 0; k < K; k++) a.mix[n * K + k] = pi(k);
  if (a.y_mean || a.y_var) {
    for (int f = 0; f < F; f++) {
      double mean = 0;
      for (int k = 0; k < K; k++) mean += pi(k) * fmean(k * F + f);
      double var;
      if (keras) {  // tfd.Mixture of MVNDiag(scale_diag = y_var): component stddev is y_var (quirk 1)
        double e2 = 0, m2 = 0;
        for (int k = 0; k < K; k++) {
          const int l = k * F + f;
          const double yv = fvar(l) + a.params[a.lo.noise + l], mu = fmean(l);
          e2 += pi(k) * yv * yv;
          m2 += pi(k) * mu * mu;
        }
        var = e2 + m2 - mean * mean;
      } else {  // MixtureSameFamily: law of total variance, centred
        double e2 = 0, c2 = 0;
        for (int k = 0; k < K; k++) {
          const int l = k * F + f;
          const double d = fmean(l) - mean;
          e2 += pi(k) * (fvar(l) + a.params[a.lo.noise + l]);
          c2 += pi(k) * d * d;
        }
        var = e2 + c2;
      }
      if (a.y_mean) a.y_mean[n * F + f] = mean;
      if (a.y_var) a.y_var[n * F + f] = var;
    }
  }
}

}  // namespace mogpe
