/*
 * mogpe_b200 -- C-ABI of the B200-native hot path of aidanscannell/mogpe.
 *
 * The reference has no FFI/plugin interface: its hot path is Python methods calling GPflow / TFP / TF
 * (SURVEY.md 8b).  This header is the boundary a replacement sits behind: the Python classes in
 * mogpe_b200/ (same names and signatures as the reference's) call these entry points through ctypes
 * with plain device/host pointers and sizes -- no framework types cross it.  Each entry point names
 * the reference code it replaces (paths relative to the reference tree).
 *
 * Conventions
 *   - float64 everywhere (reference: tf.keras.backend.set_floatx("float64"),
 *     mogpe/keras/mixture_of_experts/mosvgpe.py:13).
 *   - X [N, D], Y [N, F] row-major (reference convention).
 *   - "group" = one (inducing inputs Z, kernel) pair: Kuu, its Cholesky factor and A = L^-1 Kuf are per
 *     group.  "latent" = one latent GP (one q_mu column + one q_sqrt matrix) and belongs to a group.
 *     Latent order: expert k output f -> k*F + f  (P_e = K*F of them), then the G gating latents.
 *   - Parameters cross the boundary in CONSTRAINED space in one packed buffer (mogpe_param_layout);
 *     gradients come back in the same layout.  The host chains softplus / FillTriangular Jacobians
 *     (reference: gpflow Parameter transforms).
 *   - All inducing dimensions are padded to Mp (mogpe_layout.Mp); padded rows/cols are ignored on input
 *     and zero in gradients.
 *   - Standard-normal draws are explicit inputs (RNG parity with TF is impossible);
 *     mogpe_fill_normal generates them on the device for production use.
 *   - Ownership: the caller owns every buffer; the handle owns only scratch.  Errors: integer status,
 *     text via mogpe_last_error.  A handle is bound to one device; calls on a handle are serialised by
 *     the caller; work is enqueued on the caller's stream.
 */
#ifndef MOGPE_B200_H
#define MOGPE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOGPE_MAX_GROUPS 64
#define MOGPE_MAX_LATENTS 64
#define MOGPE_MAX_SAMPLES 16

typedef enum {
  MOGPE_OK = 0,
  MOGPE_ERR_INVALID = 1,
  MOGPE_ERR_CUDA = 2,
  MOGPE_ERR_UNSUPPORTED = 3,
  MOGPE_ERR_NCCL = 4,
  MOGPE_ERR_NUMERIC = 5 /* Cholesky of Kuu + jitter*I failed (not positive definite / NaN parameters) */
} mogpe_status;

/* bound ids: mogpe/keras/mixture_of_experts/mosvgpe.py:83-107 ; mogpe/mixture_of_experts/svgp.py:102-124 */
enum { MOGPE_BOUND_FURTHER_GATING = 0, MOGPE_BOUND_TIGHT = 1, MOGPE_BOUND_FURTHER = 2 };
/* Precision of the dense M x M x B products.  F64: FP64 DMMA everywhere (gpflow's default_float, results within 1e-8 of
 * the reference).  TF32: mogpe_predict runs A = L^-1 Kuf and q_sqrt^T A as 3xTF32 products on the tcgen05 tensor cores
 * (FP32-class accuracy, tolerance class 1e-4); kernel matrices, Cholesky, L^-1 and the predictive means stay FP64.
 * INT8: the same two products with every FP64 operand cut into eight signed-byte slices and multiplied on the INT8
 * tensor cores with exact int32 accumulation (integer slicing / Ozaki scheme): FP64-class accuracy (the 1e-8 tolerance
 * holds), about twice the throughput of the FP64 DMMA path.
 * INT8X6: the same with six slices (42 significant bits, 21 instead of 36 slice pairs): errors ~1e-7 of the largest variance
 * even for ill-conditioned Kuu -- the fast mode that meets the 1e-4 tolerance class robustly.
 * mogpe_elbo_fwd_bwd uses the mode only when it is called WITHOUT gradients (evaluation: test_step / elbo()); a call with
 * gradients always runs in FP64. */
enum { MOGPE_PRECISION_F64 = 0, MOGPE_PRECISION_TF32 = 1, MOGPE_PRECISION_INT8 = 2, MOGPE_PRECISION_INT8X6 = 3 };
/* API flavour selects the documented numerical quirks */
enum { MOGPE_FLAVOUR_KERAS = 0, MOGPE_FLAVOUR_LEGACY = 1 };

typedef struct mogpe_desc {
  int32_t input_dim;      /* D */
  int32_t output_dim;     /* F */
  int32_t num_experts;    /* K */
  int32_t num_gating;     /* G: 1 = Bernoulli gating (K must be 2), else Softmax with G == K */
  int32_t num_groups;     /* Q */
  int32_t num_latents;    /* P = K*F + G */
  int32_t max_batch;      /* largest N passed to the ELBO call; predict streams in chunks of this */
  int32_t flavour;        /* MOGPE_FLAVOUR_* */
  int32_t device;         /* CUDA device ordinal */
  int32_t precision;      /* MOGPE_PRECISION_*: arithmetic of the M x M x B products of mogpe_predict */
  double jitter;          /* gpflow default_jitter() = 1e-6 */
  const int32_t* group_num_inducing; /* [Q] true M per group */
  const int32_t* latent_group;       /* [P] */
} mogpe_desc;

/* Offsets (in doubles) into the packed parameter / gradient buffer. */
typedef struct mogpe_layout {
  int64_t lengthscales; /* [Q, D]        kernel lengthscales (ARD; host broadcasts scalars) */
  int64_t variance;     /* [Q]           kernel variance */
  int64_t Z;            /* [Q, Mp, D]    inducing inputs */
  int64_t q_mu;         /* [P, Mp] */
  int64_t q_sqrt;       /* [P, Mp, Mp]   lower-triangular */
  int64_t mean_c;       /* [P]           Constant mean function (0 for Zero) */
  int64_t noise;        /* [K*F]         Gaussian likelihood variance per expert output */
  int64_t total;
  int32_t Mp;
  int32_t reserved;
} mogpe_layout;

typedef struct mogpe_handle mogpe_handle;

/* Replaces the construction of the gpflow SVGP objects inside SVGPExpert / SVGPGatingNetwork
 * (mogpe/keras/experts.py:54-65, mogpe/keras/gating_networks.py:108-119). */
mogpe_status mogpe_create(const mogpe_desc* desc, mogpe_handle** out);
void mogpe_destroy(mogpe_handle* h);
const char* mogpe_last_error(const mogpe_handle* h); /* h may be NULL: last create error */
mogpe_status mogpe_param_layout(const mogpe_handle* h, mogpe_layout* out);

/* Minibatch ELBO, forward and (if grads != NULL) backward.
 * Replaces MixtureOfSVGPExperts.maximum_log_likelihood_objective / elbo / lower_bound_* and the
 * GradientTape in train_step (mogpe/keras/mixture_of_experts/mosvgpe.py:57-274;
 * mogpe/mixture_of_experts/svgp.py:67-469), including everything they call in experts.py,
 * gating_networks.py, gps.py and GPflow (conditional, gauss_kl, likelihoods).
 *
 *   X [N, D], Y [N, F], params, grads: device pointers.
 *   eps_u [P, S, Mp]  draws for u ~ q(u) of inducing-sampled latents (others ignored; may be NULL when
 *                     no latent is inducing-sampled under `bound`)
 *   eps_h [S, N, G]   marginal gating draws (further_gating, further)
 *   eps_f [S, N, K*F] marginal expert draws (further)
 *   scale             num_data / N, or 1.0 when num_data is None (mosvgpe.py:309-316)
 *   out   [4] device: {elbo, variational expectation (scaled), KL gating, KL experts}
 */
mogpe_status mogpe_elbo_fwd_bwd(mogpe_handle* h, const double* X, const double* Y, int32_t N,
                                const double* params, const double* eps_u, const double* eps_h,
                                const double* eps_f, int32_t bound, int32_t num_samples, double scale,
                                double* out, double* grads, void* cuda_stream);

/* Same call with HOST X / Y / eps (pinned or pageable) and a host result: copies in, runs, copies the
 * 4 result doubles out and synchronises the stream.  params / grads stay device-resident (model state).
 * This is the end-to-end entry the Python train_step uses for host-side data. */
mogpe_status mogpe_elbo_fwd_bwd_host(mogpe_handle* h, const double* X_host, const double* Y_host, int32_t N,
                                     const double* params, const double* eps_u_host,
                                     const double* eps_h_host, const double* eps_f_host, int32_t bound,
                                     int32_t num_samples, double scale, double* out_host, double* grads,
                                     void* cuda_stream);

/* Asynchronous variant for training loops: same as mogpe_elbo_fwd_bwd_host (library random stream for the draws), but it
 * returns as soon as the work is enqueued.  X / Y are copied into one of a small ring of pinned staging slots (the call
 * blocks only if the slot it needs is still in flight); the 4 result doubles are written to out_dev (device; NULL ->
 * internal) and copied asynchronously into
 * out_host_pinned (caller-owned pinned memory, mogpe_malloc_host) and are valid after the stream has been synchronised
 * (mogpe_stream_sync / mogpe_check_numerics). */
mogpe_status mogpe_elbo_fwd_bwd_host_async(mogpe_handle* h, const double* X_host, const double* Y_host, int32_t N,
                                           const double* params, int32_t bound, int32_t num_samples, double scale,
                                           double* out_dev, double* out_host_pinned, double* grads, void* cuda_stream);

/* Prediction.  Replaces predict_y / predict_experts_f / predict_experts_y / predict_mixing_probs
 * (mogpe/keras/mixture_of_experts/base.py:65-77, mosvgpe.py:276-290, keras/gating_networks.py:173-249;
 * mogpe/mixture_of_experts/base.py:58-115, mogpe/experts/svgp.py:171-227,
 * mogpe/gating_networks/svgp.py:49-105).  Any output pointer may be NULL.
 *   f_mean, f_var [N, P]  marginal q(f) per latent (with the q_sqrt term, no sampling)
 *   mix_probs [N, K]      Bernoulli: analytic probit; Softmax: mean over eps_mc [R, N, K] draws
 *   y_mean, y_var [N, F]  mixture mean / variance (flavour-specific formula)
 * X and outputs are device pointers; N is streamed in chunks of max_batch.
 *   Softmax gating (G > 1): the mixing probabilities are gpflow's Monte-Carlo mean over num_mc draws; eps_mc [num_mc, N, K]
 *   supplies them (parity tests), eps_mc == NULL takes them from the library random stream (generated per chunk on the
 *   device, indexed by the global point index: independent of max_batch; this is what makes 1e8-point sweeps possible).
 */
mogpe_status mogpe_predict(mogpe_handle* h, const double* X, int64_t N, const double* params, double* f_mean,
                           double* f_var, double* mix_probs, double* y_mean, double* y_var,
                           const double* eps_mc, int32_t num_mc, void* cuda_stream);

/* Conditional given inducing samples u_s ~ q(u) for EVERY latent (mogpe/keras/gps.py:14-50,
 * mogpe/gps/svgp.py:130-152: SVGPExpert.predict_f(num_inducing_samples=S), SVGPGatingNetwork.predict_h(...)).
 *   eps_u [P, S, Mp] (NULL -> library stream); f_mean [S, N, P] includes the mean function;
 *   f_var [N, P] is the same for every sample (no q_sqrt term).  N <= max_batch. */
mogpe_status mogpe_predict_f_inducing_samples(mogpe_handle* h, const double* X, int32_t N, const double* params,
                                              const double* eps_u, int32_t num_samples, double* f_mean,
                                              double* f_var, void* cuda_stream);

/* prior_kl of every latent GP (gpflow gauss_kl, whitened; mogpe/keras/experts.py:205-207,
 * mogpe/keras/gating_networks.py:251-253, mogpe/experts/svgp.py:161-169).  kl [P] device. */
mogpe_status mogpe_prior_kl(mogpe_handle* h, const double* params, double* kl, void* cuda_stream);

/* Monte-Carlo draws for the Softmax likelihood expectation inside the `tight` bound with G > 1 gating functions
 * (gpflow Softmax.predict_mean_and_var: mean over num_mc draws of softmax(mu + sqrt(var) eps); the legacy flavour maps
 * it over the S inducing samples, mogpe/mixture_of_experts/svgp.py:355-364).  eps_mc [S, num_mc, N, G] (device) is used
 * by the next mogpe_elbo_fwd_bwd call(s); NULL -> drawn from the library stream each call.  Default num_mc = 100. */
mogpe_status mogpe_set_mc(mogpe_handle* h, const double* eps_mc, int32_t num_mc);

/* Numerical status of the calls issued since the last check (synchronises the stream).  The reference's
 * tf.linalg.cholesky raises InvalidArgumentError when Kuu + jitter*I is not positive definite; here the factorisation
 * records the first failing group and this call (made by the *_host entry and by the Python layer whenever it reads
 * results back) returns MOGPE_ERR_NUMERIC with a message naming it. */
mogpe_status mogpe_check_numerics(mogpe_handle* h, void* cuda_stream);

/* Library-owned random stream.  When an eps_* pointer that `bound` needs is NULL, the library draws it on
 * the device from Philox(seed, counter) -- as the reference draws inside tf.random / tfd -- and advances the
 * counter.  mogpe_debug_copy("eps_u" | "eps_h" | "eps_f") returns the draws of the last call so the oracle
 * can replay them. */
mogpe_status mogpe_set_seed(mogpe_handle* h, uint64_t seed);

/* Data-parallel ranks each evaluate the KL terms redundantly; with weight 1/world_size the all-reduced
 * ELBO and gradients equal the single-process ones. Default 1. */
mogpe_status mogpe_set_kl_weight(mogpe_handle* h, double w);

/* Counter-based N(0,1) generator (Philox4x32-10 + Box-Muller) for the eps_* inputs.  Stand-in for the
 * tf.random / tfd sampling inside the reference's bounds (mogpe/keras/gps.py:29, gpflow sample_mvn). */
mogpe_status mogpe_fill_normal(mogpe_handle* h, double* dst, int64_t n, uint64_t seed, uint64_t offset,
                               void* cuda_stream);

/* Data-parallel gradient sum (north_star: one allreduce per step over NVLink).  The unique id is
 * created on rank 0 and distributed by the caller (torch.distributed / MPI / file). */
mogpe_status mogpe_nccl_unique_id(char out[128]);
mogpe_status mogpe_comm_init(mogpe_handle* h, const char id[128], int32_t rank, int32_t world_size);
mogpe_status mogpe_allreduce_sum(mogpe_handle* h, double* buf, int64_t n, void* cuda_stream);
/* The two [M x M] results of the backward that are reduced over the minibatch (Lbar = -tril(W A^T), Sbar = 2 tril(A diag(d)
 * LTA^T)) run on the INT8 tensor cores with integer-sliced operands and exact accumulation (FP64-class accuracy, section 4c of
 * DESIGN.md) when the padded M is >= 512 (M > 448; above 512 the library pads to multiples of 128) and the batch has >= 2048 rows; on (default) / off selects it. */
mogpe_status mogpe_set_int8_backward(mogpe_handle* h, int32_t on);
/* The WHOLE gradient step on the INT8 tensor cores under the same shape conditions (default on; needs the INT8 backward on):
 * A = L^-1 Kuf, S^T A, S (S^T A) and W = L^-T Abar also run as integer-sliced products (tcgen05 kind::i8, eight exact byte
 * slices, FP64-class results: same tolerances as the FP64 path); operands are exchanged between the kernels as byte planes
 * and Kuf / A / Abar are never written in FP64.  Replaces the conditional + backward of gpflow's SVGP.predict_f under
 * tf.GradientTape (mogpe/keras/gps.py:31-48, mogpe/keras/mixture_of_experts/mosvgpe.py:57-69). */
mogpe_status mogpe_set_int8_step(mogpe_handle* h, int32_t on);
/* Byte slices loaded by the GRADIENT-ONLY products of the INT8 step (S LTA, W = L^-T Abar, Lbar, Sbar): 7 (default: the
 * seven most significant planes, 49 bits, 28 slice pairs), 8 (56 bits, 36 pairs) or 6; never more than the forward products
 * load (mogpe_set_int8_forward_slices). */
mogpe_status mogpe_set_int8_grad_slices(mogpe_handle* h, int32_t slices);
/* Byte slices loaded by the two FORWARD products of the INT8 step (A = L^-1 Kuf, S^T A), whose results enter the ELBO:
 * 8 (default: 56 bits, 36 slice pairs), 7 (49 bits, 28 pairs: still float64-class -- ELBO within 1.1e-10, gradients within
 * 1.6e-8 of the FP64 step on the ill-conditioned stress model of tests/test_gpu_engine.py -- but only 0.07 ms faster at the
 * bench shape, so opt-in), or 6 = FAST TRAINING MODE, the "1e-4 class" BASELINE.json's north_star allows beside float64:
 * 42 bits, 21 slice pairs, and the gradient-only products are capped at six as well (4.55 instead of 5.1 ms/step at C4).
 * Measured in fast mode: ELBO within 4e-9, gradients within 1.1e-6 of the float64 step
 * (profiles/i8_grad_slices_accuracy_r02.log); tests/test_gpu_bench_shapes.py holds it to 1e-5 / 1e-4. */
mogpe_status mogpe_set_int8_forward_slices(mogpe_handle* h, int32_t slices);

/* on != 0: mogpe_elbo_fwd_bwd performs the data-parallel reduction itself and returns globally summed gradients and
 * results -- the part of the gradient buffer that is final before the Cholesky backward (q_mu, q_sqrt, mean, noise and, when
 * out == grads + layout.total, the 4 results: 99.9 % of the bytes) is all-reduced on a communication stream underneath
 * the rest of the backward, the kernel-hyperparameter / Z head at the end.  The caller must not all-reduce again. */
mogpe_status mogpe_set_dp_overlap(mogpe_handle* h, int32_t on);

/* Optimiser step on the device (reference: optimizer.apply_gradients in train_step,
 * mogpe/keras/mixture_of_experts/mosvgpe.py:61-63, tf.optimizers.Adam with TF defaults).
 *   var_rep [total] (host): for every entry of the packed buffer, the index of the representative entry of the
 *     variable it belongs to (itself if untied), or -1 for entries that are not variables (padding, the strict upper
 *     triangle of q_sqrt, non-trainable parameters).  A value r < -1 means representative -2 - r AND a positive (softplus)
 *     parameter regardless of its position: the diagonal of a q_diag=True q_sqrt (gpflow: positive(), not FillTriangular).
 *   u [total] (device): the unconstrained variables in the packed layout (lengthscales / variances / noise hold the
 *     pre-softplus values; everything else is stored as is).
 * mogpe_opt_transform writes params = T(u); mogpe_opt_adam_step chains `grads` (d ELBO / d constrained, as returned by
 * mogpe_elbo_fwd_bwd) to the unconstrained variables, sums tied entries, applies one Adam step to loss = -ELBO and
 * refreshes params.  t is the 1-based step count. */
mogpe_status mogpe_opt_setup(mogpe_handle* h, const int32_t* var_rep, double noise_lower_bound);
mogpe_status mogpe_opt_transform(mogpe_handle* h, double* u, double* params, void* cuda_stream);
/* Adam's moment estimates (tf.optimizers.Adam slots "m" / "v"), [total] doubles each in the packed layout, HOST pointers:
 * set == 0 copies them out, set != 0 copies them in.  Lets the host carry the optimiser state over a workspace rebuild or a
 * change of the trainable set (mogpe_opt_setup zeroes them). */
mogpe_status mogpe_opt_state(mogpe_handle* h, double* m_host, double* v_host, int32_t set);
mogpe_status mogpe_opt_adam_step(mogpe_handle* h, double* u, double* params, const double* grads, double lr,
                                 double beta1, double beta2, double epsilon, int64_t t, void* cuda_stream);

/* One whole training step on the library's own stream: [copy the minibatch in from host memory through a pinned ring],
 * ELBO forward + backward with draws from the library random stream, bijector chain, Adam (TF semantics, step count t),
 * params = T(u), and an asynchronous copy of the 4 result doubles to out_host_pinned (nullable).  Replaces one iteration
 * of the Keras fit loop / monitored_training_loop (mogpe/keras/mixture_of_experts/mosvgpe.py:57-69,
 * mogpe/training/training_loops.py:28-30).  With use_graph != 0 the launches (~45 kernels) are captured into a CUDA graph the
 * second time a given argument set is seen and replayed afterwards: the random-stream position and the Adam step count
 * live on the device, so replays are exact continuations.  Small models are launch-bound; this removes the launch cost.
 * use_graph: 0 = never, 1 = only when the step is launch-bound (Q M^2 B < 2e8), 2 = always.
 * Nothing is synchronised: results are valid after mogpe_check_numerics / mogpe_stream_sync(device, NULL). */
mogpe_status mogpe_train_step(mogpe_handle* h, const double* X, const double* Y, int32_t data_on_host, int32_t N,
                              double* params, int32_t bound, int32_t num_samples, double scale, double* out,
                              double* out_host_pinned, double* grads, double* u, double lr, double beta1, double beta2,
                              double epsilon, int64_t t, int32_t use_graph);

/* The same step with the minibatch assembled by the library: X_all [n, D] / Y_all [n, F] are the whole HOST data set and
 * rows [N] (host, int64) the indices of this minibatch (one shuffled slice of tf.data's shuffle().batch(),
 * mogpe/training/utils.py:105-116); rows are gathered straight into the pinned staging slot. */
mogpe_status mogpe_train_step_rows(mogpe_handle* h, const double* X_all, const double* Y_all, const int64_t* rows, int32_t N,
                                   double* params, int32_t bound, int32_t num_samples, double scale, double* out,
                                   double* out_host_pinned, double* grads, double* u, double lr, double beta1, double beta2,
                                   double epsilon, int64_t t, int32_t use_graph);

/* Number of captured training-step graphs and whether capture failed and was switched off (diagnostics / tests). */
mogpe_status mogpe_graph_stats(mogpe_handle* h, int32_t* captured, int32_t* disabled);

/* Minibatch assembly from a device-resident data set (replaces the tf.data shuffle/batch pipeline of
 * mogpe/training/utils.py:105-116): dst[i, :] = src[idx[i], :], src [N, cols], idx [n] int64, all device pointers. */
mogpe_status mogpe_gather_rows(mogpe_handle* h, const double* src, const int64_t* idx, int64_t n, int32_t cols,
                               double* dst, void* cuda_stream);

/* Device memory helpers so the Python host needs no tensor framework. */
mogpe_status mogpe_malloc(int32_t device, int64_t bytes, void** out);
mogpe_status mogpe_free(int32_t device, void* p);
mogpe_status mogpe_malloc_host(int64_t bytes, void** out); /* pinned */
mogpe_status mogpe_free_host(void* p);
mogpe_status mogpe_memcpy_h2d(int32_t device, void* dst, const void* src, int64_t bytes, void* cuda_stream);
mogpe_status mogpe_memcpy_d2h(int32_t device, void* dst, const void* src, int64_t bytes, void* cuda_stream);
mogpe_status mogpe_memcpy_d2h_async(int32_t device, void* dst_pinned, const void* src, int64_t bytes, void* cuda_stream);
mogpe_status mogpe_memset_zero(int32_t device, void* dst, int64_t bytes, void* cuda_stream);
mogpe_status mogpe_stream_sync(int32_t device, void* cuda_stream);

/* Measurement hooks (bench.py roofline): with profiling on, every DMMA GEMM launch is bracketed by CUDA
 * events on the caller's stream.  mogpe_profile_read synchronises, returns the accumulated GEMM device
 * time (ms), the algorithmic FLOPs of those launches, their count, and the count of ALL kernels this
 * handle launched since the last read, then resets the counters. */
mogpe_status mogpe_profile_enable(mogpe_handle* h, int32_t on);
mogpe_status mogpe_profile_read(mogpe_handle* h, double* gemm_ms, double* gemm_flops, int64_t* gemm_launches,
                                int64_t* all_launches);
/* INT8-sliced product launches (gemm_i8_sliced_kernel) of the step since the last read: device time (ms), ALGORITHMIC
 * FP64-equivalent FLOPs, launch count, and the INT8 tensor operations behind them (36 slice-pair operations per FLOP with
 * eight slices, 28 with seven). */
mogpe_status mogpe_profile_read_int8(mogpe_handle* h, double* ms, double* flops, int64_t* launches, double* tensor_ops);
/* Profiling mode also times the streaming kernels alone on the caller's stream: ms[4], bytes[4] for
 * {kuf, lik, build_abar, kuf_grad} (accumulated since the last read; bytes = algorithmic bytes). */
mogpe_status mogpe_profile_read_streaming(mogpe_handle* h, double* ms, double* bytes);

/* Roofline denominator measured on the spot: issue rate of the FP64 tensor pipe (independent mma.sync.m8n8k4.f64 chains,
 * SASS DMMA.8x8x4, 2 x 256 threads per SM, ~5 ms), in TFLOP/s.  MEASURED_PEAKS.json holds no FP64 figure. */
mogpe_status mogpe_probe_fp64_tensor_peak(int32_t device, double* tflops);
/* The same for the INT8 tensor pipe (tcgen05.mma kind::i8, 128 x 256 x 32 from shared-memory operands, one issuing warp per
 * SM), in TOP/s (2 operations per multiply-accumulate): the denominator for gemm_i8_sliced_kernel, whose 36 slice-pair
 * products per FP64-equivalent multiply-add run on this pipe. */
mogpe_status mogpe_probe_int8_tensor_peak(int32_t device, double* tops);

/* Test hook: copy a named intermediate of the last call to the host (Kuu, L, Linv, Kuf, A, ...).
 * Returns the element count in *n_out; copies min(capacity, count) doubles. */
mogpe_status mogpe_debug_copy(mogpe_handle* h, const char* name, double* dst_host, int64_t capacity,
                              int64_t* n_out);

/* Test hook: batched FP64 GEMM on the library's DMMA kernel, C = alpha*op(A)*op(B) + beta*C (row-major,
 * device pointers, dims multiples of 64).  kmode: bit0 left-lower, bit1 left-upper, bit2 right-lower,
 * bit3 right-upper (restricts the k range per tile); tril_out: only lower-triangular tiles are produced;
 * cfg: tile configuration id (-1 = automatic). */
mogpe_status mogpe_debug_gemm(int32_t device, const double* A, const double* B, double* C, int32_t batch,
                              int32_t M, int32_t N, int32_t K, int32_t transA, int32_t transB, int32_t kmode,
                              int32_t tril_out, double alpha, double beta, int32_t cfg, void* cuda_stream);
/* Test/tuning hook: force a GEMM tile configuration for all later launches (-1 = automatic). */
mogpe_status mogpe_debug_set_gemm_cfg(int32_t cfg);

const char* mogpe_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MOGPE_B200_H */
