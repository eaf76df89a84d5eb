/* mtrl_b200 C-ABI.
 *
 * The reference (reginald-mclean/mtrl) has no native layer at all: its hot path is NumPy on the
 * host (mtrl/rl/buffers.py:494-549) followed by one jitted XLA executable
 * (mtrl/rl/algorithms/mtsac.py:1173-1251).  This header is the boundary a binding for that path
 * would attach to; each entry names the reference symbol it replaces.  All pointers named
 * "device" are CUDA device pointers owned by the caller (the Python side allocates them through
 * PyTorch); `stream` is a cudaStream_t passed as void*.  Every function returns 0 on success or a
 * negative MTRL_ERR_* code and leaves a message readable through mtrl_last_error().  No entry
 * synchronises the host unless its comment says so.  Handles are not re-entrant.
 */
#ifndef MTRL_B200_H_
#define MTRL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Last error text of the calling thread ("" if none). */
const char* mtrl_last_error(void);
/* Library ABI version; bumped whenever a struct below changes layout. */
int mtrl_abi_version(void);
/* n host -> device copies on `stream` in one call (the five arrays of a batch handed to `update`, mtrl/types.py:30-35,
 * mtrl/rl/algorithms/base.py:221): cudaMemcpyAsync semantics per entry (pinned sources must stay untouched until the stream
 * has passed the copies; pageable sources are staged before the call returns). */
int mtrl_memcpy_h2d_batch(int n, void* const* dst, const void* const* src, const long long* bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Grouped TF32 GEMM (tcgen05 / TMEM / TMA).  Replaces the dot_generals XLA emits for
 * nn.Dense in MultiHeadNetwork (mtrl/nn/multi_head.py:34-44) and their VJPs
 * (jax.value_and_grad at mtrl/rl/algorithms/mtsac.py:587-596, 689-691).
 * D[M][N] = sum_k A(m,k) * B(n,k); fp32 storage, tf32 operands (or hi + lo tf32 pairs, see A_lo), fp32 accumulate.
 * ------------------------------------------------------------------------------------------ */
enum {
  MTRL_EPI_STORE = 0,       /* D = acc                                                      */
  MTRL_EPI_BIAS_RELU = 1,   /* D = tf32(relu(acc + bias[n]))     forward Dense + ReLU        */
  MTRL_EPI_RELU_MASK = 2,   /* D = tf32(acc) where mask[m][n] > 0 else 0   ReLU VJP          */
  MTRL_EPI_ATOMIC_ADD = 3,  /* D += acc (float4 atomics; required when k_splits > 1)         */
  MTRL_EPI_STORE_TF32 = 4   /* D = tf32(acc)                                                 */
};

typedef struct mtrl_gemm_problem {
  const float* A;   /* device; a_major 0: [M][K] row-major (K contiguous); 1: [K][M] (M contiguous) */
  long long lda;    /* row pitch of A in floats (multiple of 4)                                     */
  int a_major;
  const float* B;   /* device; b_major 0: [N][K] row-major (K contiguous); 1: [K][N] (N contiguous) */
  long long ldb;
  int b_major;
  float* D;         /* device; [M][N] row-major                                                    */
  long long ldd;
  int M, N, K;      /* N must be a multiple of 4                                                    */
  int block_n;      /* tile width, multiple of 16 in [16, 256]                                      */
  int k_splits;     /* >= 1; > 1 needs MTRL_EPI_ATOMIC_ADD and a zeroed D                           */
  int epilogue;     /* MTRL_EPI_*                                                                   */
  const float* bias;    /* device [N] for MTRL_EPI_BIAS_RELU                                        */
  const float* mask;    /* device [M][N] for MTRL_EPI_RELU_MASK                                     */
  long long ldmask;
  const unsigned* mask_bits; /* MTRL_EPI_RELU_MASK alternative to `mask`: bit (n % 32) of word [m][n / 32] set <=>
                                the forward activation was > 0 (rows of ldbits words)                            */
  unsigned* relu_bits_out;   /* optional, MTRL_EPI_BIAS_RELU: emit those bits for the backward pass               */
  long long ldbits;
  float* colsum_partial; /* optional, MTRL_EPI_RELU_MASK only: device [ceil(M/32)][N] receiving the column sums of
                            every 32-row group of D (bias gradients are their sum over groups); NULL to skip  */
  int schedule_first;    /* != 0: this problem's tiles are dealt to the workers before all others (outputs that travel
                            over NVLink: their stores then overlap the remaining tiles instead of the launch's tail)  */
  int phase;             /* 0 .. 7.  Problems of phase p start only after EVERY problem of the launch with a smaller phase
                            has completed and its outputs are visible (a grid-wide barrier inside the persistent kernel):
                            the consecutive Dense layers of one network pass, or of its backward, run as ONE launch
                            instead of one launch per layer.  All zero = a plain grouped launch.                      */
  /* fp32x3 mode ("3xTF32", the precision the reference's fp32 CPU dots have).  A_lo / B_lo: the tf32 remainders of the
   * operands (x = hi + lo, both tf32; same shape, pitch and layout as A / B).  Both set: the contraction accumulates
   * A B + A B_lo + A_lo B in the same fp32 TMEM accumulator (three tcgen05.mma passes per k-block).  D_lo (needs an
   * epilogue that rounds: BIAS_RELU, RELU_MASK, STORE_TF32; same pitch as D): also emit the tf32 remainder of the
   * unrounded result, so that D / D_lo feed the next contraction as a hi / lo pair; column-sum partials are then taken
   * from the unrounded values. */
  const float* A_lo;
  const float* B_lo;
  float* D_lo;
  /* Fused output head (MTRL_EPI_BIAS_RELU only; all four set, or head_w NULL): the own-task head of a multi-head network
   * (nn.vmap(Dense) heads, mtrl/nn/multi_head.py:50-66) evaluated on the tile while it is still in registers, so the
   * last trunk activation is not read back for it:
   *   head_out[m][j] += sum_n D[m][n] * head_w[task(m)][n][j],   task(m) = head_tile_task[m / 128]
   * (float atomics, one per row, output and column slice of the tile; the caller zeroes head_out and adds the head bias).
   * D is the value the epilogue stores (tf32-rounded; the unrounded hi + lo value with D_lo).  head_w is the Flax head
   * kernel (tasks, N, head_dim) row-major; head_dim in {1, 2, 4, 8}. */
  const float* head_w;
  float* head_out;
  const int* head_tile_task;
  int head_dim;
} mtrl_gemm_problem_t;

typedef struct mtrl_gemm_plan mtrl_gemm_plan_t;

#define MTRL_GEMM_MAX_PROBLEMS 24   /* problems per launch */
#define MTRL_GEMM_MAX_PHASES 8      /* dependent phases per launch */

/* Encodes the TMA descriptors for up to 24 problems that will run as one persistent launch. */
int mtrl_gemm_plan_create(mtrl_gemm_plan_t** out, const mtrl_gemm_problem_t* problems, int n);
/* Same with the tile shape chosen by the caller: ctas = 2 -> 256-row tiles over CTA pairs (cta_group::2; fewest operand
 * bytes per flop, the default), 1 -> 128-row tiles on single CTAs (twice as many, half-size units: better when a launch
 * has too few 256-row units to fill the 74 pairs evenly, e.g. one rank's rows of a task-sharded batch), 0 -> default. */
int mtrl_gemm_plan_create_ex(mtrl_gemm_plan_t** out, const mtrl_gemm_problem_t* problems, int n, int ctas);
/* OR-ed into `ctas`: stream-K schedule.  The launch's k-blocks, not its tiles, are dealt evenly to the SMs (every SM gets a
 * contiguous run of k-blocks that may start and end inside a tile); a tile cut into several units is finished by the unit
 * that completes last, which adds the others' parked accumulators to its own before the fused epilogue.  For launches of
 * only a few tiles per SM (one rank's rows of a task-sharded batch, MT10), where whole-tile rounds leave SMs idle. */
#define MTRL_GEMM_STREAMK 16
/* OR-ed into `ctas`: the phases of the launch are ordered by per-row-tile dependencies instead of grid-wide barriers.  Every
 * problem of phase p > 0 must take as its (K-major) A operand the D of one problem of phase p - 1 with the same rows (the next
 * Dense layer of the same network pass); a tile of it starts as soon as the producing problem has stored every column tile of
 * those rows, so SMs that run out of tiles of one layer continue with the next instead of idling at a barrier -- for chains of
 * short layers (few rows: MT10, one rank's share of a sharded batch) whose tile count does not fill whole rounds of SMs. */
#define MTRL_GEMM_ROWDEPS 32
int mtrl_gemm_plan_run(mtrl_gemm_plan_t* plan, void* stream);
int mtrl_gemm_plan_units(const mtrl_gemm_plan_t* plan);
/* 2 when the plan runs as CTA pairs (tcgen05 cta_group::2, 256-row tiles), 1 for single-CTA tiles. */
int mtrl_gemm_plan_ctas(const mtrl_gemm_plan_t* plan);
/* Debug aid: device long long[8] receiving per-role cycle sums (see csrc/gemm_tcgen05.cu); NULL disables. */
int mtrl_gemm_plan_set_debug(mtrl_gemm_plan_t* plan, long long* dbg);
void mtrl_gemm_plan_destroy(mtrl_gemm_plan_t* plan);

/* ------------------------------------------------------------------------------------------
 * Replay sampler.  Replaces MultiTaskReplayBuffer.add / .sample
 * (mtrl/rl/buffers.py:426-474, 494-549) and the np.random.Generator(PCG64) it draws from
 * (buffers.py:260, 523-527).  Storage arrays are device fp32 (capacity, T, dim), reference layout
 * (buffers.py:293-306), owned by the caller.  Array order everywhere: observations, actions,
 * next_observations, dones, rewards (= ReplayBufferSamples, mtrl/types.py:30-35).
 * ------------------------------------------------------------------------------------------ */
typedef struct mtrl_sampler mtrl_sampler_t;

int mtrl_sampler_create(mtrl_sampler_t** out, int capacity, int num_tasks, int obs_dim, int act_dim,
                        float* obs, float* actions, float* next_obs, float* dones, float* rewards);
void mtrl_sampler_destroy(mtrl_sampler_t* s);
/* state4 = {state_hi, state_lo, inc_hi, inc_lo}: numpy's PCG64 `bit_generator.state` dict split in
 * 64-bit words; has_uint32/uinteger are the buffered 32-bit half.  Both calls synchronise `stream`. */
int mtrl_sampler_set_state(mtrl_sampler_t* s, const uint64_t* state4, uint32_t has_uint32, uint32_t uinteger,
                           void* stream);
int mtrl_sampler_get_state(mtrl_sampler_t* s, uint64_t* state4, uint32_t* has_uint32, uint32_t* uinteger,
                           void* stream);
/* buffers.py:453-457.  Sources are (T, dim) fp32, host or device. */
int mtrl_sampler_add(mtrl_sampler_t* s, int pos, const float* obs, const float* actions, const float* next_obs,
                     const float* dones, const float* rewards, void* stream);
/* buffers.py:520-549: one shared index vector of n_per_task draws in [0, max(fill, n_per_task)); outputs
 * are device (n_per_task*T, dim), row i <-> (sample i / T, task i % T).  norm_mode 1 applies
 * rewards' = (double(r) - shift[t]) / den[t] (buffers.py:531-538); shift/den are device double[T]. */
int mtrl_sampler_sample(mtrl_sampler_t* s, int fill, int n_per_task, long long* idx_out, float* obs_out,
                        float* actions_out, float* next_obs_out, float* dones_out, float* rewards_out,
                        int norm_mode, const double* shift, const double* den, void* stream);
/* The index draw alone, `rng.integers(0, high, size=n)` (buffers.py:523-527): n device int64 values, generator advanced
 * exactly as numpy's PCG64 + 32-bit Lemire rejection does.  1 <= high < 2^32, n <= the sampler's index capacity. */
int mtrl_sampler_draw(mtrl_sampler_t* s, unsigned long long high, int n, long long* idx_out, void* stream);
/* buffers.py:496-519: counts is a host int[T]; independent draws per task, rows concatenated by task. */
int mtrl_sampler_sample_per_task(mtrl_sampler_t* s, int fill, const int* counts, float* obs_out,
                                 float* actions_out, float* next_obs_out, float* dones_out, float* rewards_out,
                                 void* stream);

/* ------------------------------------------------------------------------------------------
 * MT-SAC update.  Replaces MTSAC.update / _update_inner (mtrl/rl/algorithms/mtsac.py:1173-1251):
 * update_critic (MSE branch, :513-621), update_actor (:623-711), update_alpha (:713-731) on
 * MultiHeadNetwork actor / critic ensemble (mtrl/nn/multi_head.py:21-68, mtrl/rl/networks.py:21-67,
 * 208-222), optax clip_by_global_norm + adam (mtrl/config/optim.py:26-43) and the Polyak target
 * update (mtsac.py:607-613).
 * ------------------------------------------------------------------------------------------ */
#define MTRL_MAX_DEPTH 4
#define MTRL_VARIANT_MTSAC 0
#define MTRL_VARIANT_SAC 1
/* Arithmetic of the trunk contractions (storage and accumulation are fp32 in both):
 *   MTRL_PRECISION_TF32   operands rounded to tf32 (what XLA's default f32 dot precision does on NVIDIA GPUs), fastest;
 *   MTRL_PRECISION_FP32X3 every operand kept as a (hi, lo) pair of tf32 values and three tensor-core passes per k-block
 *                         ("3xTF32"): reproduces fp32 products to ~2^-22, i.e. the numerics of the reference's fp32 CPU
 *                         path, at 3x the tensor time.  The parity mode for the 1e-3 tolerance on every parameter leaf. */
#define MTRL_PRECISION_TF32 0
#define MTRL_PRECISION_FP32X3 1

typedef struct mtrl_sac_config {
  int num_tasks;        /* T: width of the one-hot block that ends every observation             */
  int task_begin;       /* first task owned by this handle (0 on one GPU)                         */
  int num_local_tasks;  /* tasks owned by this handle (T on one GPU)                              */
  int obs_dim;          /* observation length including the one-hot (39 + T for Meta-World)       */
  int action_dim;       /* 1..8                                                                   */
  int width;            /* hidden width, multiple of 4                                            */
  int depth;            /* trunk layers, 1..MTRL_MAX_DEPTH (reference default 3)                  */
  int num_critics;      /* ensemble size, 1..4 (reference default 2)                              */
  int max_rows;         /* packed-row capacity: >= sum_t roundup(rows of task t, 128); mult of 128 */
  int max_batch;        /* largest batch (rows) one update call may pass                          */
  float gamma, tau;
  float actor_lr, critic_lr, alpha_lr;
  float adam_b1, adam_b2, adam_eps;
  float actor_max_grad_norm, critic_max_grad_norm, alpha_max_grad_norm; /* <= 0: no clipping      */
  float log_std_min, log_std_max;
  float target_entropy;
  int clip_q;           /* AlgorithmConfig.clip: clamp target and prediction to +-5000            */
  int use_task_weights; /* MTSACConfig.use_task_weights                                           */
  unsigned long long noise_seed; /* Philox seed used when eps_c / eps_a are NULL                  */
  int variant;          /* MTRL_VARIANT_MTSAC: MTSAC._update_inner (mtsac.py:1173-1247);
                           MTRL_VARIANT_SAC: single-task SAC._update_inner (sac.py:262-383): num_tasks = 1 (the network
                           is a plain MLP, mtrl/nn/base.py:11-63, its last Dense is the one "head"), alpha updated first,
                           critic loss 0.5 * sum_e mean_b, parameter-norm logs of the pre-update parameters        */
  int precision;        /* MTRL_PRECISION_*                                                                        */
  int use_layer_norm;        /* MTRL_VARIANT_SAC only: VanillaNetworkConfig.use_layer_norm (mtrl/config/nn.py:33-39): a
                                flax LayerNorm (eps 1e-6, scale + bias) before every Dense of the MLP except the first
                                (mtrl/nn/base.py:35-37, 52-53)                                                        */
  int use_skip_connections;  /* MTRL_VARIANT_SAC only: VanillaNetworkConfig.use_skip_connections: hidden layers whose
                                input has the hidden width add it to their activation (mtrl/nn/base.py:46-49)         */
} mtrl_sac_config_t;

/* Flat fp32 layout of one network (all ensemble members).  [member trunks | 32 reduction slots |
 * member heads]; the trunk prefix (plus the slots) is what ranks all-reduce.  Flax names:
 * layer_i/kernel (in, W) at trunk(e) + kernel_off[i], layer_i/bias (W) at trunk(e) + bias_off[i], (LayerNorm_i below,)
 * VmapDense_0/kernel (T_local, W, head) at heads(e) + head_kernel_off, bias (T_local, head). */
typedef struct mtrl_net_layout {
  long long total;
  long long trunk_total;        /* members * member_trunk_stride                                  */
  long long slots_off;          /* == trunk_total; 32 floats                                      */
  long long heads_base;
  long long member_trunk_stride;
  long long member_head_stride;
  long long kernel_off[MTRL_MAX_DEPTH];
  long long bias_off[MTRL_MAX_DEPTH];
  long long head_kernel_off;
  long long head_bias_off;
  int in_dim, head_dim, members, num_local_tasks, width, depth;
  /* MLP with use_layer_norm: LayerNorm_k/scale (W) and /bias (W), k = 0 .. depth-1, inside the member's trunk at
   * trunk(e) + ln_scale_off[k] / ln_bias_off[k]; LayerNorm_k normalises the input of layer_{k+1} (the output Dense for
   * k = depth-1).  use_layer_norm = 0: the arrays are unused. */
  long long ln_scale_off[MTRL_MAX_DEPTH];
  long long ln_bias_off[MTRL_MAX_DEPTH];
  int use_layer_norm, reserved;
} mtrl_net_layout_t;

typedef struct mtrl_sac_layout {
  mtrl_net_layout_t actor;
  mtrl_net_layout_t critic;
  long long workspace_bytes;
  int k_actor;   /* padded input pitch of the actor  (floats) */
  int k_critic;  /* padded input pitch of the critic (floats) */
} mtrl_sac_layout_t;

int mtrl_sac_query_layout(const mtrl_sac_config_t* cfg, mtrl_sac_layout_t* out);

/* All device memory, allocated by the caller with the sizes mtrl_sac_query_layout reports
 * (net arrays: layout.total floats, 128-byte aligned; zero-initialise grads/m/v/steps). */
typedef struct mtrl_sac_buffers {
  float *actor_params, *actor_grads, *actor_m, *actor_v, *actor_shadow;
  float *critic_params, *critic_grads, *critic_m, *critic_v, *critic_shadow;
  float *critic_target, *critic_target_shadow;
  float *log_alpha, *alpha_m, *alpha_v; /* [num_local_tasks]                                       */
  int* steps;                           /* int[4]: actor, critic, alpha Adam counts; noise counter */
  float* logs;                          /* float[16], see MTRL_LOG_*                               */
  void* workspace;
} mtrl_sac_buffers_t;

/* Indices into logs[] in the order the reference merges its log dicts
 * (mtsac.py:616-621, 704-709, 728-731). */
enum {
  MTRL_LOG_QF_VALUES = 0,
  MTRL_LOG_QF_LOSS = 1,
  MTRL_LOG_CRITIC_GRAD_MAGNITUDE = 2,
  MTRL_LOG_CRITIC_PARAMS_NORM = 3,
  MTRL_LOG_ACTOR_LOSS = 4,
  MTRL_LOG_ACTOR_GRAD_MAGNITUDE = 5,
  MTRL_LOG_ACTOR_PARAMS_NORM = 6,
  MTRL_LOG_EXPLORE_LOSS = 7,
  MTRL_LOG_ALPHA_LOSS = 8,
  MTRL_LOG_ALPHA = 9,
  MTRL_LOG_COUNT = 16
};

typedef struct mtrl_sac mtrl_sac_t;

int mtrl_sac_create(mtrl_sac_t** out, const mtrl_sac_config_t* cfg, const mtrl_sac_buffers_t* buffers);
void mtrl_sac_destroy(mtrl_sac_t* h);
/* Recompute the tf32 operand copies after the caller wrote params / target directly. */
int mtrl_sac_refresh_shadows(mtrl_sac_t* h, void* stream);

/* One MTSAC.update on `batch` rows (device fp32, ReplayBufferSamples field order, any row order;
 * the task of a row is argmax of its trailing one-hot).  global_batch is the B every loss mean
 * divides by (== batch on one GPU).  eps_c / eps_a: device (batch, action_dim) standard normal
 * draws for the critic-target and actor samples, or NULL for in-kernel Philox. */
int mtrl_sac_update(mtrl_sac_t* h, const float* obs, const float* actions, const float* next_obs,
                    const float* dones, const float* rewards, int batch, int global_batch, const float* eps_c,
                    const float* eps_a, void* stream);
/* The same update cut at the two points where ranks exchange trunk gradients (all-reduce sum over
 * grads[0 .. trunk_total + 32) of the critic after phase 1 and of the actor after phase 2). */
int mtrl_sac_phase1_critic_grads(mtrl_sac_t* h, const float* obs, const float* actions, const float* next_obs,
                                 const float* dones, const float* rewards, int batch, int global_batch,
                                 const float* eps_c, const float* eps_a, void* stream);
int mtrl_sac_phase2_critic_step_actor_grads(mtrl_sac_t* h, void* stream);
int mtrl_sac_phase3_actor_step_alpha(mtrl_sac_t* h, void* stream);
/* MTSAC.sample_action / eval_action (mtsac.py:70-84, 299-311): actions for n observation rows (device fp32
 * (n, obs_dim), any owned tasks, n <= max_rows).  deterministic = 0: a = tanh(mu + sigma eps) with eps = device
 * (n, action_dim) standard normal draws or NULL for in-kernel Philox; deterministic = 1: the mode tanh(mu)
 * (nn/distributions.py:15-16).  actions_out: device fp32 (n, action_dim).  Rows of a task this handle does not own
 * set the status word (mtrl_sac_read_status_async) and come back as zeros.  Not re-entrant with mtrl_sac_update. */
int mtrl_sac_act(mtrl_sac_t* h, const float* obs, int n, const float* eps, int deterministic, float* actions_out,
                 void* stream);
/* MultiHeadNetwork.__call__ (mtrl/nn/multi_head.py:21-68) of one of the handle's networks on n arbitrary rows (device fp32
 * obs (n, obs_dim), actions (n, action_dim) for the critics; any owned tasks, any order; n <= max_rows):
 *   net 0  the actor          -> out (n, 2 * action_dim): what ContinuousActionPolicy splits into mean / log_std (networks.py:36-37)
 *   net 1  the critic ensemble -> out (num_critics, n, 1): QValueFunction on concatenate((actions, obs)) (networks.py:55-67, 208-222)
 *   net 2  the target critics, same shape
 * Trunk layers run as the update's tcgen05 GEMMs in the handle's precision, then each row's own head.  Uses the update's
 * activation buffers as scratch: not re-entrant with mtrl_sac_update.  Rows of tasks this handle does not own set status
 * word 0 (mtrl_sac_read_status_async) and come back as zeros. */
int mtrl_mlp_forward(mtrl_sac_t* h, int net, const float* obs, const float* actions, int n, float* out, void* stream);
/* optax.chain(clip_by_global_norm(max_grad_norm), adam(lr, b1, b2, eps)) + apply_updates (mtrl/config/optim.py:26-43,
 * mtrl/rl/algorithms/utils.py:11-46), optionally followed by optax.incremental_update(params, target, tau)
 * (mtsac.py:607-613), on caller-owned flat device fp32 buffers of n elements (n % 4 == 0, 16-byte aligned; target may be
 * NULL; max_grad_norm <= 0: no clipping).  step: device int, the Adam count, incremented.  scratch: device double[5]
 * receiving [0] the squared gradient norm (before clipping), [1] |params'|^2, [3] |params|^2 before the step. */
int mtrl_adam_polyak_step(float* params, const float* grads, float* m, float* v, float* target, long long n, int* step, float lr,
                          float b1, float b2, float eps, float max_grad_norm, float tau, double* scratch, void* stream);
/* The fused SAC loss pass alone: min over the target ensemble, entropy term, Bellman target, clip, weighted MSE and dL/dQ
 * (mode 0: mtsac.py:547-566) or the actor loss alpha logp - min_e Q_e with the arg-min routing of its gradient (mode 1:
 * mtsac.py:659-666), from the LAST trunk activations and the heads' parameters.  Rows are packed in 128-row tiles, all
 * rows of a tile belonging to one task (tile_task), padding rows marked by row_valid < 0. */
typedef struct mtrl_sac_losses_args {
  int mode;                      /* 0 critic loss, 1 actor loss                                                    */
  int rows, width, num_critics;  /* rows: multiple of 128                                                          */
  int global_batch;              /* the B every mean divides by                                                    */
  int clip_q;                    /* mode 0: clamp target and prediction to +-5000 (AlgorithmConfig.clip)           */
  float gamma;
  const float* H_target[4];      /* mode 0: [rows][width] last trunk activation of every target critic             */
  const float* H_online[4];      /* [rows][width] of every online critic                                           */
  const float* w_target[4];      /* (T, width, 1) head kernels                                                     */
  const float* b_target[4];      /* (T, 1)                                                                         */
  const float* w_online[4];
  const float* b_online[4];
  const int* tile_task;          /* [rows / 128]                                                                   */
  const int* row_valid;          /* [rows]: < 0 marks a padding row                                                */
  const float* rewards;          /* mode 0: [rows]                                                                 */
  const float* dones;            /* mode 0: [rows]                                                                 */
  const float* logp_next;        /* mode 0: [rows] log pi(a'|s')                                                   */
  const float* logp;             /* mode 1: [rows] log pi(a|s)                                                     */
  const float* alpha;            /* [T] exp(log_alpha)                                                             */
  const float* task_weights;     /* [T] (ones when use_task_weights is off)                                        */
  float* dq;                     /* out [num_critics][rows]: dL/dQ_e                                               */
  double* acc;                   /* out double[4]: mode 0: [0] sum w (Q - y)^2, [1] sum Q; mode 1: [2] sum w (alpha logp - min Q) */
} mtrl_sac_losses_args_t;
int mtrl_sac_losses_fwd_bwd(const mtrl_sac_losses_args_t* args, void* stream);
/* Per-task gradients, the first half of MTSAC.compute_weights (mtsac.py:870-1170: batch split by task,
 * jax.vmap(jax.value_and_grad(critic_loss / actor_loss)) over tasks).  critic_tg: device fp32 (T, critic layout.total),
 * actor_tg: (T, actor layout.total), row t = gradient of task t in the flat network layout (zero outside task t's own
 * head).  Values are the full-batch gradients restricted to task t's rows = (n_t / B) x the reference's per-task-mean
 * gradients.  No parameter is updated.  Needs every task on this handle and equally many rows per task (status 3
 * otherwise, mtrl_sac_read_status_async). */
int mtrl_sac_task_grads(mtrl_sac_t* h, const float* obs, const float* actions, const float* next_obs, const float* dones,
                        const float* rewards, int batch, const float* eps_c, const float* eps_a, float* critic_tg,
                        float* actor_tg, void* stream);
/* Gram matrix of a per-task gradient matrix: gram (T, T) = rows rows^T, fp32 (flat_grads @ flat_grads.T of
 * compute_gram_metrics, mtsac.py:747; the input of vmap_cos_sim and compute_conflict_metrics, utils.py:49-174).
 * rows: device fp32 (T, ld) with P valid columns, T <= 64. */
int mtrl_task_gram(const float* rows, long long ld, int T, long long P, float* gram, void* stream);
/* The element-wise reductions of compute_conflict_metrics (mtrl/rl/algorithms/utils.py:75-101) over a per-task gradient
 * matrix rows (T, ld), P valid columns, every element multiplied by `scale` first:
 *   mismatch (T, T) fp32 : #{p : |x_a| < eps and |x_b| > tau}   (compute_sparsity_mismatch before its normalisation)
 *   row_stats (T, 2) f64 : {sum_p |x_t|, #{p : |x_t| < eps}}     (participation ratio l1, near-zero counts)
 * Columns that are layout padding are zero and therefore counted as near-zero: the caller subtracts them. */
int mtrl_task_elementwise(const float* rows, long long ld, int T, long long P, float scale, float eps, float tau,
                          float* mismatch, double* row_stats, void* stream);
/* compute_support_metrics (mtrl/rl/algorithms/mtsac.py:774-860) building blocks on a per-task gradient matrix rows (T, ld):
 * mtrl_task_abs_order_stats: out2 (T, 2) device fp32 = the ranks[t]-th and (ranks[t] + 1)-th smallest |x| of row t (radix
 *   select; the two neighbours jnp.quantile(|g|, 0.8) interpolates between, :804-806).  ranks: HOST long long[T];
 *   scratch: device, T * (32 + 2048) bytes.  Synchronises the stream once.
 * mtrl_task_support_pairs: out3 (3, T, T) device fp32 = for supports {|x_t| >= thr[t]}: support intersections, sign
 *   conflicts (x_a x_b < 0), genuine conflicts (both in support and sign conflict) (:809-835). */
int mtrl_task_abs_order_stats(const float* rows, long long ld, int T, long long P, const long long* ranks, float* out2,
                              void* scratch, void* stream);
int mtrl_task_support_pairs(const float* rows, long long ld, int T, long long P, const float* thr, float* out3, void* stream);
/* PCGradConfig (mtrl/config/optim.py:62-76): optax.chain(pcgrad(num_tasks), clip_by_global_norm, adam).  After this
 * call mtrl_sac_update splits the critic's and / or the actor's loss by task (mtsac.py:568-585, 677-687), runs the
 * per-task gradients through pcgrad (mtrl/optim/pcgrad.py:22-136, in coefficient space over the Gram matrix) and feeds
 * its averaged output to clip + adam.  critic_tg / actor_tg: device fp32 (T, layout.total) work matrices; scratch: device
 * fp32 of 2 T^2 + 2 T + 8 floats = [Gram critic | Gram actor | weights critic | weights actor | stats critic[4] |
 * stats actor[4]], stats = n_grad_conflicts, avg_grad_magnitude, avg_grad_magnitude_before_surgery, norm of the plain
 * mean gradient; perm_*: device int[T] row permutations (pcgrad.py:79) the caller rewrites before every update, or
 * NULL for the identity.  T <= 64, every task on this handle, equally many rows per task. */
int mtrl_sac_enable_pcgrad(mtrl_sac_t* h, int critic, int actor, float* critic_tg, float* actor_tg, float* scratch,
                           const int* perm_critic, const int* perm_actor);
/* CAGradConfig (mtrl/config/optim.py:104-124): optax.chain(cagrad(num_tasks), clip_by_global_norm, adam), same wiring
 * as mtrl_sac_enable_pcgrad with cagrad's defaults (mtrl/optim/cagrad.py:20-41: c = 0.5, 21 SGD iterations on the task
 * weights, lr 25 / 50, momentum 0.5).  scratch: 2 T^2 + 4 T + 8 floats = the pcgrad layout (stats = norm of the
 * combined gradient, mean clipped per-task norm, best objective) followed by the critic's and the actor's softmax task
 * weights (CAGradState.task_weights). */
int mtrl_sac_enable_cagrad(mtrl_sac_t* h, int critic, int actor, float* critic_tg, float* actor_tg, float* scratch);
/* GradNormConfig (mtrl/config/optim.py:79-102): optax.chain(gradnorm(...), clip_by_global_norm, adam).  The reference's
 * gradnorm loss is independent of the task weights (mtrl/optim/gradnorm.py:134-142), so the weights never move from 1
 * and the transformation returns the SUM of the per-task gradients, each clipped to unit norm first when
 * clip_per_task (its max_grad_norm) is set (:40-57, 106-107, 155-157).  scratch as for pcgrad; stats = norm of the sum,
 * mean per-task norm after clipping. */
int mtrl_sac_enable_gradnorm(mtrl_sac_t* h, int critic, int actor, int clip_per_task, float* critic_tg, float* actor_tg,
                             float* scratch);
/* DummyMultiTaskConfig (mtrl/config/optim.py:46-59): optax.chain(dummy_multitask_optimizer(), clip_by_global_norm, adam);
 * the transformation is the mean over tasks of the per-task gradients (mtrl/optim/dummy.py:5-20) of the split losses. */
int mtrl_sac_enable_dummy(mtrl_sac_t* h, int critic, int actor, float* critic_tg, float* actor_tg, float* scratch);
/* The gradient transformations of mtrl/optim as one stand-alone operator (what the optax-protocol objects of
 * mtrl_b200.optim call): rows (T, ld) device fp32 = per-task gradients (updates with a leading task axis, raveled), P valid
 * columns (P and ld multiples of 4, T <= 64); out (P) = the transformed gradient:
 *   kind 0  pcgrad  (mtrl/optim/pcgrad.py:20-136): mean of the sequentially projected rows; perm = device int[T] row
 *           permutation (:79) or NULL for the identity
 *   kind 1  cagrad  (mtrl/optim/cagrad.py:20-237) with its defaults
 *   kind 2  gradnorm (mtrl/optim/gradnorm.py:61-163): sum of the rows, each clipped to unit norm first if clip_per_task
 *   kind 3  dummy   (mtrl/optim/dummy.py:5-20): mean of the rows
 * scratch: device fp32, T*T + 2*T + 4 floats = [Gram | weights on the rows | stats[4] | cagrad task weights]; stats as in
 * mtrl_sac_enable_pcgrad / _cagrad / _gradnorm. */
int mtrl_task_combine(int kind, const float* rows, long long ld, int T, long long P, const int* perm, int clip_per_task,
                      float* out, float* scratch, void* stream);
/* Number of kernels one mtrl_sac_update launches (for bench.py's gpu_launches). */
int mtrl_sac_launches_per_update(const mtrl_sac_t* h);
/* Bracket every GEMM launch of the following updates with CUDA events on the launch stream
 * (enable = 1) and read back their summed duration and count (synchronises on those events). */
int mtrl_sac_profile_gemms(mtrl_sac_t* h, int enable);
int mtrl_sac_profile_read(mtrl_sac_t* h, double* total_ms, int* launches);
/* The exchange kernels (csrc/comm.cuh) bracketed in the same pass: their summed duration and count as of the last
 * mtrl_sac_profile_read. */
int mtrl_sac_profile_exchange(mtrl_sac_t* h, double* total_ms, int* launches);
/* The other kernel classes bracketed in the same pass, as of the last mtrl_sac_profile_read: ms12[c] / n12[c] = summed
 * duration / number of launches of class c: 0 GEMM, 1 exchange, 2 Adam + Polyak, 3 head VJP, 4 critic loss, 5 actor head
 * (policy sample + log-prob), 6 actor loss, 7 batch packing (3 kernels per bracket), 8 gradient norms, 9 bias-gradient
 * column sums, 10 LayerNorm junctions; 11 unused. */
int mtrl_sac_profile_classes(mtrl_sac_t* h, double* ms12, int* n12);
/* Asynchronous copy of the status words of the last update into 4 pinned host ints:
 * [0] == 0 ok, 1: a row's task is outside this handle's task range, 2: rows do not fit max_rows, 3: unbalanced split;
 * [1] != 0: the peer exchange is dead -- an in-kernel wait for another rank timed out (code = 1 + barrier, 10 + barrier
 *     for the in-rank grid barrier); no rank has applied the step that was in flight, every later update is a no-op for
 *     the trunk, and the host must stop (mtrl_comm_error reads the same word synchronously). */
int mtrl_sac_read_status_async(const mtrl_sac_t* h, int* host_pinned4, void* stream);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU exchange (nothing in the single-device reference corresponds to it).  Tasks are sharded over
 * the GPUs of one box; the replicated trunk needs the sum of every rank's trunk gradients before
 * optax.clip_by_global_norm + adam (mtrl/config/optim.py:26-43).  Instead of a library all-reduce, each rank owns a
 * cudaMalloc "arena" that every other rank maps through CUDA IPC; ONE kernel per network then does
 * reduce-scatter (peer loads over NVLink) -> clip -> Adam on the owned 1/N of the trunk -> all-gather (peer stores)
 * -> Polyak / tf32 copies (csrc/comm.cuh).  Arena = [4096-byte header | caller-defined regions]; the four regions
 * a handle exchanges (critic/actor gradients and parameters) must sit at the same offsets on every rank.
 * ------------------------------------------------------------------------------------------ */
#define MTRL_IPC_HANDLE_BYTES 64
typedef struct mtrl_comm mtrl_comm_t;

/* Allocates and zeroes the arena on the current device and writes its 64-byte IPC handle to handle_out. */
int mtrl_comm_create(mtrl_comm_t** out, int rank, int world, long long arena_bytes, unsigned char* handle_out);
/* Device pointer of the local arena. */
void* mtrl_comm_arena(mtrl_comm_t* c);
/* handles: world x 64 bytes, rank-major (the caller all-gathers what mtrl_comm_create returned). */
int mtrl_comm_open_peers(mtrl_comm_t* c, const unsigned char* handles);
/* NVSwitch multicast ("NVLS") region for the parameter all-gather (optional; needs CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED
 * on every GPU).  Protocol, each step on every rank unless noted, with a host barrier after add_device and after bind:
 *   rank 0: mtrl_comm_mc_create(bytes) -> POSIX fd, passed to the other ranks by the host (SCM_RIGHTS);
 *   others: mtrl_comm_mc_import(bytes, fd);   all: mtrl_comm_mc_add_device;   all: mtrl_comm_mc_bind.
 * mtrl_comm_mc_local = this rank's copy of the region (ordinary device memory: the parameter buffers of the handle
 * live here), mtrl_comm_mc_ptr = the multicast alias (multimem.st only).  mtrl_sac_attach_comm on a comm with a bound
 * region takes the two parameter offsets relative to that region. */
int mtrl_comm_mc_supported(int* out);
int mtrl_comm_mc_create(mtrl_comm_t* c, long long bytes, int* fd_out);
int mtrl_comm_mc_import(mtrl_comm_t* c, long long bytes, int fd);
int mtrl_comm_mc_add_device(mtrl_comm_t* c);
int mtrl_comm_mc_bind(mtrl_comm_t* c);
void* mtrl_comm_mc_local(mtrl_comm_t* c);
void* mtrl_comm_mc_ptr(mtrl_comm_t* c);
long long mtrl_comm_mc_bytes(mtrl_comm_t* c);
/* Synchronous read of the arena's error word: 0 ok, otherwise the in-kernel wait that timed out (a peer never came). */
int mtrl_comm_error(mtrl_comm_t* c, int* code);
/* Synchronous read of the phase durations (microseconds, CTA 0 of this rank) of the last two fused trunk steps:
 * us14[0..6] critic, us14[7..13] actor; per set: wait for all ranks' gradients, owned-segment norms, norm exchange,
 * Adam + all-gather stores, wait for all ranks' stores, derived copies (Polyak / tf32), total. */
int mtrl_comm_phase_times(mtrl_comm_t* c, double* us14);
void mtrl_comm_destroy(mtrl_comm_t* c);
/* Switches a multi-task handle to the fused peer-memory exchange.  The handle's critic_grads, actor_grads,
 * critic_params and actor_params buffers must be arena + the given byte offsets (same offsets on every rank).
 * Afterwards mtrl_sac_update runs the whole sharded update with no host-visible exchange points. */
int mtrl_sac_attach_comm(mtrl_sac_t* h, mtrl_comm_t* c, long long off_critic_grads, long long off_actor_grads,
                         long long off_critic_params, long long off_actor_params);

/* Host-only: the ownership table the sharded exchange uses for one network layout (from mtrl_sac_query_layout) and
 * `world` ranks, as rows {begin, end, owner rank, pre_reduced} in floats into the flat buffer.  The segments tile
 * [0, trunk_total) exactly; pre_reduced rows are hidden-layer kernel row blocks whose gradient the dW GEMM epilogues
 * reduce into the owner's buffer. */
int mtrl_trunk_segments(const mtrl_net_layout_t* layout, int world, long long* out4, int max_segments, int* n_out);
/* Checkpoint support for the sharded exchange: writes 1.0 / 0.0 over [0, trunk_total) of the actor (critic = 0) or
 * critic (critic = 1) layout -- 1.0 where THIS handle holds the live Adam moments (everything without an attached
 * arena).  sum over ranks of mask * moments is the full optimiser state of the reference's TrainState.  Synchronises. */
int mtrl_sac_trunk_owner_mask(mtrl_sac_t* h, int critic, float* mask_dev);

/* ------------------------------------------------------------------------------------------
 * MT-PPO update.  Replaces MTPPO.update / _update_inner (mtrl/rl/algorithms/mtppo.py:292-317):
 * update_policy (:196-254) + update_value_function (:256-290), ContinuousActionPolicy with
 * squash_tanh = False (mtrl/rl/networks.py:29-45) and ValueFunction (:188-205) on MultiHeadNetwork
 * (num_tasks > 1) or a plain MLP (num_tasks = 1).  One full-batch step per network per call.
 * ------------------------------------------------------------------------------------------ */
typedef struct mtrl_ppo_config {
  int num_tasks;        /* heads; 1 = plain MLP (the observation then carries no one-hot)                    */
  int obs_dim;          /* >= 16                                                                             */
  int action_dim;       /* 1..8                                                                              */
  int width, depth;
  int steps_per_task;   /* rollout length per task; the batch is num_tasks * steps_per_task rows, task-major */
  float clip_eps;       /* MTPPOConfig.clip_eps (0.2)                                                        */
  int clip_vf_loss;
  float entropy_coefficient, vf_coefficient;
  int normalize_advantages;
  float policy_lr, vf_lr;
  float adam_b1, adam_b2, adam_eps;
  float policy_max_grad_norm, vf_max_grad_norm;   /* <= 0: no clipping                                       */
  float log_std_min, log_std_max;
  unsigned long long noise_seed;
  int use_layer_norm;        /* num_tasks == 1 (plain MLP) only: VanillaNetworkConfig.use_layer_norm, as in mtrl_sac_config_t */
  int use_skip_connections;  /* num_tasks == 1 only: VanillaNetworkConfig.use_skip_connections                               */
} mtrl_ppo_config_t;

typedef struct mtrl_ppo_layout {
  mtrl_net_layout_t policy;
  mtrl_net_layout_t vf;
  long long workspace_bytes;
  int k_in;      /* padded input pitch (floats)                     */
  int max_rows;  /* packed rows = num_tasks * roundup(steps, 128)   */
} mtrl_ppo_layout_t;

typedef struct mtrl_ppo_buffers {
  float *policy_params, *policy_grads, *policy_m, *policy_v, *policy_shadow;
  float *vf_params, *vf_grads, *vf_m, *vf_v, *vf_shadow;
  int* steps;     /* int[4]: policy, value-function Adam counts; unused; noise counter                       */
  float* logs;    /* float[16]: entropy_loss, policy_loss, approx_kl, clip_fracs, value_function, values     */
  void* workspace;
} mtrl_ppo_buffers_t;

typedef struct mtrl_ppo mtrl_ppo_t;

int mtrl_ppo_query_layout(const mtrl_ppo_config_t* cfg, mtrl_ppo_layout_t* out);
int mtrl_ppo_create(mtrl_ppo_t** out, const mtrl_ppo_config_t* cfg, const mtrl_ppo_buffers_t* buffers);
void mtrl_ppo_destroy(mtrl_ppo_t* h);
int mtrl_ppo_refresh_shadows(mtrl_ppo_t* h, void* stream);
/* Rollout arrays: device fp32, flattened task-major (row = task * steps_per_task + step), the fields of
 * mtrl/types.py:48-63: observations (B, obs_dim), log_probs, advantages, returns, values (B).  eps: (B, action_dim)
 * standard normal draws for the fresh policy sample of mtppo.py:205-207, or NULL for in-kernel Philox. */
int mtrl_ppo_update(mtrl_ppo_t* h, const float* obs, const float* log_probs, const float* advantages,
                    const float* returns, const float* values, const float* eps, void* stream);
int mtrl_ppo_launches_per_update(const mtrl_ppo_t* h);

/* ------------------------------------------------------------------------------------------
 * Rollout advantages.  Replaces the GAE loop of MultiTaskRolloutBuffer.get (mtrl/rl/buffers.py:650-707).
 * rewards / values / dones / advantages / returns: device fp32 (num_steps, num_tasks[, 1]) in the reference's
 * storage order; last_values / last_dones: device fp32 (num_tasks).  The last step bootstraps from the `dones`
 * argument (the reference reads the whole self.dones array there and cannot run for num_steps > 1).
 * Bit-identical to NumPy's float32 evaluation order.
 * ------------------------------------------------------------------------------------------ */
int mtrl_gae(const float* rewards, const float* values, const float* dones, const float* last_values,
             const float* last_dones, int num_steps, int num_tasks, float gamma, float gae_lambda, float* advantages,
             float* returns, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MTRL_B200_H_ */
