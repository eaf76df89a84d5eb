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

/* ------------------------------------------------------------------------------------------
 * Grouped TF32 GEMM (tcgen05 / TMEM / TMA).  Replaces the dot_generals XLA emits for
 * nn.Dense in MultiHeadNetwork (mtrl/nn/multi_head.py:34-44) and their VJPs
 * (jax.value_and_grad at mtrl/rl/algorithms/mtsac.py:587-596, 689-691).
 * D[M][N] = sum_k A(m,k) * B(n,k); fp32 storage, tf32 operands, fp32 accumulate.
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
} mtrl_gemm_problem_t;

typedef struct mtrl_gemm_plan mtrl_gemm_plan_t;

/* Encodes the TMA descriptors for up to 8 problems that will run as one persistent launch. */
int mtrl_gemm_plan_create(mtrl_gemm_plan_t** out, const mtrl_gemm_problem_t* problems, int n);
int mtrl_gemm_plan_run(mtrl_gemm_plan_t* plan, void* stream);
int mtrl_gemm_plan_units(const mtrl_gemm_plan_t* plan);
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
/* buffers.py:496-519: counts is a host int[T]; independent draws per task, rows concatenated by task. */
int mtrl_sampler_sample_per_task(mtrl_sampler_t* s, int fill, const int* counts, float* obs_out,
                                 float* actions_out, float* next_obs_out, float* dones_out, float* rewards_out,
                                 void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MTRL_B200_H_ */
