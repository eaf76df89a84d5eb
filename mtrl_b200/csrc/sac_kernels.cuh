// Non-GEMM kernels of the MT-SAC update (HBM-bound elementwise / reduction work):
// batch packing, per-task heads, the fused SAC losses, bias gradients, the multi-tensor
// Adam + global-norm clip + Polyak step and the temperature step.
// Reference semantics: mtrl/rl/algorithms/mtsac.py:513-731, 1173-1247 (see sac.cu for the sequence).
#pragma once

#include <curand_kernel.h>

#include "common.cuh"
#include "mtrl_b200.h"

namespace sac {

constexpr int kTileRows = 128;   // rows of one task are padded to a multiple of the GEMM M tile
constexpr int kMaxE = 4;
constexpr int kMaxA = 8;
constexpr int kColsumSplits = 32;

// accumulators (double, zeroed at the start of every update)
enum {
  ACC_QLOSS = 0, ACC_QSUM, ACC_CRITIC_G2, ACC_CRITIC_P2_TRUNK, ACC_CRITIC_P2_HEAD, ACC_CRITIC_HEAD_G2,
  ACC_ACTOR_LOSS, ACC_ACTOR_G2, ACC_ACTOR_P2_TRUNK, ACC_ACTOR_P2_HEAD, ACC_ACTOR_HEAD_G2,
  ACC_CRITIC_P2_OLD, ACC_ACTOR_P2_OLD,   // squared norm of the parameters BEFORE the step (sac.py:360-364 logs those)
  ACC_EXPLORE,                           // sum over rows and action dims of (a_data - a_pi)^2 (split actor loss, mtsac.py:668-673)
  ACC_COUNT = 16
};

// extra (non reference) log slots used to combine ranks: squared norms that are per-rank partial
enum { LOG_X_CRITIC_P2_TRUNK = 10, LOG_X_CRITIC_P2_HEAD = 11, LOG_X_ACTOR_P2_TRUNK = 12, LOG_X_ACTOR_P2_HEAD = 13 };

__device__ __forceinline__ float warp_sum_f(float v) { return warp_sum(v); }

__device__ __forceinline__ float block_sum(float v, float* smem) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem[w] = v;
  __syncthreads();
  if (w == 0) {
    v = lane < (blockDim.x + 31) / 32 ? smem[lane] : 0.f;
    v = warp_sum(v);
  }
  return v;  // valid in warp 0
}
__device__ __forceinline__ double block_sum(double v, double* smem) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem[w] = v;
  __syncthreads();
  if (w == 0) {
    v = lane < (blockDim.x + 31) / 32 ? smem[lane] : 0.0;
    v = warp_sum(v);
  }
  return v;
}

// ---------------------------------------------------------------------------------------------
// Packing: rows arrive in any order (the reference batch is (sample, task) interleaved,
// buffers.py:547-548); internally rows are task-major with each task padded to 128-row tiles so a
// GEMM M tile belongs to one task (own-task head only).
// ---------------------------------------------------------------------------------------------
// Task of every row = first arg-max of its trailing one-hot (jnp.argmax, multi_head.py:65).  One warp per row.
static __global__ void row_task_kernel(const float* __restrict__ obs, int B, int obs_dim, int T, int task_begin, int T_local,
                                int* __restrict__ row_slot, int* __restrict__ status) {
  MTRL_PDL_PROLOGUE();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* oh = obs + static_cast<long long>(row) * obs_dim + (obs_dim - T);
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int t = lane; t < T; t += 32) {
    const float v = oh[t];
    if (v > best) { best = v; bi = t; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (lane == 0) {
    int task = bi - task_begin;
    if (task < 0 || task >= T_local) { atomicExch(status, 1); task = -1; }
    row_slot[row] = task;
  }
}

static __global__ void pack_plan_kernel(int B, int T_local, int max_rows, int* __restrict__ row_slot,
                                 int* __restrict__ slot_src, int* __restrict__ tile_task, int* __restrict__ seg_start,
                                 int* __restrict__ status) {
  MTRL_PDL_PROLOGUE();
  extern __shared__ int sm[];
  const int nchunks = (B + 31) / 32;
  int* counts = sm;                       // [nchunks][T_local]
  int* seg = sm + nchunks * T_local;      // [T_local + 1]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  for (int i = tid; i < nchunks * T_local; i += blockDim.x) counts[i] = 0;
  for (int i = tid; i < max_rows; i += blockDim.x) slot_src[i] = -1;
  __syncthreads();
  // pass 1: per-chunk histograms of the row tasks (row_task_kernel stashed them in row_slot)
  for (int c = warp; c < nchunks; c += nwarps) {
    const int row = c * 32 + lane;
    const int task = row < B ? row_slot[row] : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, task);
    if (task >= 0 && lane == __ffs(peers) - 1) counts[c * T_local + task] = __popc(peers);
  }
  __syncthreads();
  // pass 2: exclusive scan over chunks per task
  for (int t = tid; t < T_local; t += blockDim.x) {
    int run = 0;
    for (int c = 0; c < nchunks; ++c) {
      const int v = counts[c * T_local + t];
      counts[c * T_local + t] = run;
      run += v;
    }
    seg[t] = run;  // rows of task t
  }
  __syncthreads();
  if (tid == 0) {
    int off = 0;
    for (int t = 0; t < T_local; ++t) {
      const int n = seg[t];
      const int padded = (n + kTileRows - 1) / kTileRows * kTileRows;
      seg[t] = off;
      seg_start[t] = off;
      if (off + padded > max_rows) { atomicExch(status, 2); break; }
      for (int tile = off / kTileRows; tile < (off + padded) / kTileRows; ++tile) tile_task[tile] = t;
      off += padded;
    }
    seg[T_local] = off;
    seg_start[T_local] = off;
    for (int tile = off / kTileRows; tile < max_rows / kTileRows; ++tile) tile_task[tile] = 0;
  }
  __syncthreads();
  if (*status) return;
  // pass 3: slot of every row (stable within a task)
  for (int c = warp; c < nchunks; c += nwarps) {
    const int row = c * 32 + lane;
    const int task = row < B ? row_slot[row] : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, task);
    if (task >= 0) {
      const int rank = __popc(peers & ((1u << lane) - 1u));
      const int slot = seg[task] + counts[c * T_local + task] + rank;
      row_slot[row] = slot;
      slot_src[slot] = row;
    }
  }
}

struct PackArgs {
  const float *obs, *actions, *next_obs, *dones, *rewards, *eps_c, *eps_a;
  float *Xa_next, *Xa, *Xc_next, *Xc, *rew, *done, *peps_c, *peps_a;
  float* act_data;                   // [M][A] the batch actions in packed order, unrounded (explore term of the split actor loss)
  const int* slot_src;
  const int* noise_counter;
  unsigned long long seed;
  unsigned long long noise_stream;   // mixed into the Philox subsequence: the first task this handle owns (shards draw
                                     // independent noise although they share the seed, which initialises the trunk)
  long long lo_delta;                // fp32x3: the tf32 remainder of X[i] goes to X[i + lo_delta] (0 = tf32 mode)
  int obs_dim, act_dim, Ka, Kc;
};

// Store an operand value as its tf32 hi part and, in fp32x3 mode, the remainder at p[lo_delta].
__device__ __forceinline__ void store_operand(float* p, long long lo_delta, float v) {
  const float hi = tf32_rna(v);
  *p = hi;
  if (lo_delta) p[lo_delta] = tf32_lo(v, hi);
}

// One block per packed row.  Inputs are rounded to tf32 here (they are GEMM A operands).
static __global__ void pack_rows_kernel(const PackArgs a) {
  MTRL_PDL_PROLOGUE();
  const int slot = blockIdx.x;
  const int src = a.slot_src[slot];
  const int A = a.act_dim, od = a.obs_dim;
  float* xa = a.Xa + static_cast<long long>(slot) * a.Ka;
  float* xan = a.Xa_next + static_cast<long long>(slot) * a.Ka;
  float* xc = a.Xc + static_cast<long long>(slot) * a.Kc;
  float* xcn = a.Xc_next + static_cast<long long>(slot) * a.Kc;
  const long long ld = a.lo_delta;
  if (src < 0) {
    for (int j = threadIdx.x; j < a.Ka; j += blockDim.x) { store_operand(xa + j, ld, 0.f); store_operand(xan + j, ld, 0.f); }
    for (int j = threadIdx.x; j < a.Kc; j += blockDim.x) { store_operand(xc + j, ld, 0.f); store_operand(xcn + j, ld, 0.f); }
    if (threadIdx.x < A) { a.peps_c[slot * A + threadIdx.x] = 0.f; a.peps_a[slot * A + threadIdx.x] = 0.f; a.act_data[slot * A + threadIdx.x] = 0.f; }
    if (threadIdx.x == 0) { a.rew[slot] = 0.f; a.done[slot] = 0.f; }
    return;
  }
  if (threadIdx.x < A) a.act_data[slot * A + threadIdx.x] = a.actions[static_cast<long long>(src) * A + threadIdx.x];
  const float* o = a.obs + static_cast<long long>(src) * od;
  const float* on = a.next_obs + static_cast<long long>(src) * od;
  for (int j = threadIdx.x; j < a.Ka; j += blockDim.x) {
    store_operand(xa + j, ld, j < od ? o[j] : 0.f);
    store_operand(xan + j, ld, j < od ? on[j] : 0.f);
  }
  for (int j = threadIdx.x; j < a.Kc; j += blockDim.x) {
    float v = 0.f, vn = 0.f;
    if (j < A) v = a.actions[static_cast<long long>(src) * A + j];   // (action, state) order, networks.py:61
    else if (j < A + od) { v = o[j - A]; vn = on[j - A]; }
    store_operand(xc + j, ld, v);
    store_operand(xcn + j, ld, vn);  // action columns of the next-state input are filled by the actor sample
  }
  if (threadIdx.x == 0) {
    a.rew[slot] = a.rewards[src];
    a.done[slot] = a.dones[src];
    if (a.eps_c && a.eps_a) {
      for (int d = 0; d < A; ++d) {
        a.peps_c[slot * A + d] = a.eps_c[static_cast<long long>(src) * A + d];
        a.peps_a[slot * A + d] = a.eps_a[static_cast<long long>(src) * A + d];
      }
    } else {
      curandStatePhilox4_32_10_t st;
      curand_init(a.seed, (a.noise_stream << 32) | static_cast<unsigned long long>(src),
                  static_cast<unsigned long long>(*a.noise_counter) * 32ull, &st);
      for (int d = 0; d < A; d += 4) {
        const float4 n = curand_normal4(&st);
        const float nn[4] = {n.x, n.y, n.z, n.w};
        for (int q = 0; q < 4 && d + q < A; ++q) a.peps_c[slot * A + d + q] = nn[q];
      }
      for (int d = 0; d < A; d += 4) {
        const float4 n = curand_normal4(&st);
        const float nn[4] = {n.x, n.y, n.z, n.w};
        for (int q = 0; q < 4 && d + q < A; ++q) a.peps_a[slot * A + d + q] = nn[q];
      }
    }
  }
}

// Action sampling (mtsac.py:70-84, 299-311): observations of any tasks, one block per row.  Writes the tf32 actor
// input row, the row's task (first arg-max of the trailing one-hot, relative to this handle's first task; -1 and
// status 1 if it is not owned here) and the row's noise: caller's eps, zeros (mode) or Philox.
static __global__ void act_pack_kernel(const float* __restrict__ obs, int n, int obs_dim, int Ka, int T, int task_begin,
                                       int T_local, const float* __restrict__ eps_in, int deterministic, int A,
                                       unsigned long long seed, unsigned long long call, float* __restrict__ Xa,
                                       int* __restrict__ row_task, float* __restrict__ eps_out, int* __restrict__ status,
                                       long long lo_delta) {
  const int row = blockIdx.x;
  const float* o = obs + static_cast<long long>(row) * obs_dim;
  float* xa = Xa + static_cast<long long>(row) * Ka;
  for (int j = threadIdx.x; j < Ka; j += blockDim.x) store_operand(xa + j, lo_delta, j < obs_dim ? o[j] : 0.f);
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const float* oh = o + (obs_dim - T);
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int t = lane; t < T; t += 32) {
      const float v = oh[t];
      if (v > best) { best = v; bi = t; }
    }
    for (int off = 16; off > 0; off >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) {
      int task = bi - task_begin;
      if (task < 0 || task >= T_local) { atomicExch(status, 1); task = -1; }
      row_task[row] = task;
      if (eps_in) {
        for (int d = 0; d < A; ++d) eps_out[row * A + d] = eps_in[static_cast<long long>(row) * A + d];
      } else if (deterministic) {
        for (int d = 0; d < A; ++d) eps_out[row * A + d] = 0.f;
      } else {
        // subsequences above 2^60 are never used by the update's noise ((first owned task << 32) | batch row)
        curandStatePhilox4_32_10_t st;
        curand_init(seed, (1ull << 60) + (static_cast<unsigned long long>(task_begin) << 32) + static_cast<unsigned long long>(row),
                    call * 8ull, &st);
        for (int d = 0; d < A; d += 4) {
          const float4 z = curand_normal4(&st);
          const float zz[4] = {z.x, z.y, z.z, z.w};
          for (int q = 0; q < 4 && d + q < A; ++q) eps_out[row * A + d + q] = zz[q];
        }
      }
    }
  }
}

// Stand-alone network forward (mtrl_mlp_forward): rows of any owned tasks, in the caller's order.  Writes the padded GEMM
// input row [actions | obs] (actions == nullptr: [obs], the actor's input) and the row's task.  One block per row.
static __global__ void mlp_pack_kernel(const float* __restrict__ obs, const float* __restrict__ actions, int obs_dim, int act_dim,
                                       int K, int T, int task_begin, int T_local, float* __restrict__ X, int* __restrict__ row_task,
                                       int* __restrict__ status, long long lo_delta) {
  const int row = blockIdx.x;
  const float* o = obs + static_cast<long long>(row) * obs_dim;
  const int A = actions ? act_dim : 0;
  float* x = X + static_cast<long long>(row) * K;
  for (int j = threadIdx.x; j < K; j += blockDim.x) {
    float v = 0.f;
    if (j < A) v = actions[static_cast<long long>(row) * act_dim + j];
    else if (j < A + obs_dim) v = o[j - A];
    store_operand(x + j, lo_delta, v);
  }
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const float* oh = o + (obs_dim - T);
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int t = lane; t < T; t += 32) {
      const float v = oh[t];
      if (v > best) { best = v; bi = t; }
    }
    for (int off = 16; off > 0; off >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) {
      int task = bi - task_begin;
      if (task < 0 || task >= T_local) { atomicExch(status, 1); task = -1; }
      row_task[row] = task;
    }
  }
}

// out[e][row][j] = H_e[row] . Wh_e[task(row)][:, j] + bh_e[task(row)][j]: the own-task head of MultiHeadNetwork
// (multi_head.py:50-66) for rows in arbitrary order.  One warp per (row, member); rows of foreign tasks give zeros.
struct HeadFwdArgs {
  const float* H[kMaxE];    // [n][W]
  const float* Wh[kMaxE];   // (T_local, W, HD)
  const float* bh[kMaxE];   // (T_local, HD)
  const int* row_task;
  float* out;               // [E][n][HD]
  long long h_lo_delta;
  int n, W, HD, E;
};
static __global__ void head_fwd_kernel(const HeadFwdArgs p) {
  const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (gw >= p.n * p.E) return;
  const int e = gw / p.n, row = gw % p.n;
  const int t = p.row_task[row];
  float* o = p.out + (static_cast<long long>(e) * p.n + row) * p.HD;
  const float* h = p.H[e] + static_cast<long long>(row) * p.W;
  for (int j = 0; j < p.HD; ++j) {
    float s = 0.f;
    if (t >= 0) {
      const float* w = p.Wh[e] + static_cast<long long>(t) * p.W * p.HD + j;
      for (int k = lane; k < p.W; k += 32) {
        const float hv = p.h_lo_delta ? h[k] + h[k + p.h_lo_delta] : h[k];
        s = fmaf(hv, __ldg(w + static_cast<long long>(k) * p.HD), s);
      }
    }
    s = warp_sum(s);
    if (lane == 0) o[j] = t >= 0 ? s + p.bh[e][t * p.HD + j] : 0.f;
  }
}

// alpha_t = exp(log_alpha_t) (MultiTaskTemperature, mtsac.py:60-63); w_t = T * softmax(-log_alpha)_t
// (extract_task_weights, mtsac.py:103-113) or 1.
static __global__ void alpha_prep_kernel(const float* __restrict__ log_alpha, int T_local, int use_w, float* __restrict__ alpha_val,
                                  float* __restrict__ task_w) {
  MTRL_PDL_PROLOGUE();
  if (blockIdx.x != 0) return;
  __shared__ float red[32];
  float mx = -INFINITY;
  for (int t = threadIdx.x; t < T_local; t += blockDim.x) mx = fmaxf(mx, -log_alpha[t]);
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < (blockDim.x + 31) / 32; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float s = 0.f;
  for (int t = threadIdx.x; t < T_local; t += blockDim.x) s += expf(-log_alpha[t] - mx);
  s = block_sum(s, red);
  __shared__ float denom;
  if (threadIdx.x == 0) denom = s;
  __syncthreads();
  for (int t = threadIdx.x; t < T_local; t += blockDim.x) {
    alpha_val[t] = expf(log_alpha[t]);
    task_w[t] = use_w ? expf(-log_alpha[t] - mx) / denom * static_cast<float>(T_local) : 1.f;
  }
}

// ---------------------------------------------------------------------------------------------
// Actor head + tanh-Gaussian sample and log-prob (networks.py:29-45, nn/distributions.py:6-16).
// One warp per packed row.
// ---------------------------------------------------------------------------------------------
struct ActorHeadArgs {
  long long h_lo_delta;  // fp32x3: H[i + h_lo_delta] is the tf32 remainder of H[i]; the heads then see hi + lo (0 = tf32 mode)
  const float* H;        // [M][W] last trunk activation
  const float* pre;      // optional [M][2A]: H . Wh already computed by the GEMM epilogue (fused heads); H is then not read
  const float* Wh;       // (T_local, W, 2A)
  const float* bh;       // (T_local, 2A)
  const int* tile_task;
  const int* slot_src;
  const int* row_task;   // optional: task of every row (< 0 = skip) instead of tile_task / slot_src (action sampling, where
                         // rows are not packed into per-task tiles)
  const float* eps;      // [M][A] packed
  float* Xdst;           // critic input buffer whose columns [0, A) receive tf32(a)
  long long xdst_lo_delta;  // fp32x3: the remainder of a goes to Xdst[... + xdst_lo_delta] (0 = tf32 mode)
  int ldx;
  float* act;            // [M][A] (may be null)
  float* logp;           // [M]
  float* logstd;         // [M][A] clipped (may be null)
  unsigned* inrange;     // [M] (may be null)
  int M, W;
  float ls_min, ls_max;
};

template <int A>
static __global__ void actor_head_kernel(const ActorHeadArgs p) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= p.M) return;
  const bool valid = p.row_task ? p.row_task[row] >= 0 : p.slot_src[row] >= 0;
  const int t = p.row_task ? max(p.row_task[row], 0) : p.tile_task[row / kTileRows];
  float acc[2 * A];
#pragma unroll
  for (int j = 0; j < 2 * A; ++j) acc[j] = 0.f;
  if (valid && p.pre) {
#pragma unroll
    for (int j = 0; j < 2 * A; ++j) acc[j] = p.pre[static_cast<long long>(row) * (2 * A) + j];
  } else if (valid) {
    const float* h = p.H + static_cast<long long>(row) * p.W;
    const float* w = p.Wh + static_cast<long long>(t) * p.W * (2 * A);
    for (int k = lane; k < p.W; k += 32) {
      const float hv = p.h_lo_delta ? h[k] + h[k + p.h_lo_delta] : h[k];
      const float* wk = w + static_cast<long long>(k) * (2 * A);
#pragma unroll
      for (int j = 0; j < 2 * A; ++j) acc[j] = fmaf(hv, __ldg(wk + j), acc[j]);
    }
  }
  if (!p.pre) {
#pragma unroll
    for (int j = 0; j < 2 * A; ++j) acc[j] = warp_sum(acc[j]);
  }
  float lp = 0.f;
  unsigned mask = 0;
  if (lane < A) {
    const int d = lane;
    float mean = 0.f, ls_raw = 0.f;
#pragma unroll
    for (int j = 0; j < A; ++j) {
      if (j == d) { mean = acc[j]; ls_raw = acc[A + j]; }
    }
    float a = 0.f, ls = 0.f;
    if (valid) {
      mean += p.bh[t * 2 * A + d];
      ls_raw += p.bh[t * 2 * A + A + d];
      ls = fminf(fmaxf(ls_raw, p.ls_min), p.ls_max);
      const float sd = expf(ls);
      const float e = p.eps[row * A + d];
      const float x = fmaf(sd, e, mean);
      a = tanhf(x);
      const float z = -2.f * x;
      const float softplus = z > 0.f ? z + log1pf(expf(-z)) : log1pf(expf(z));
      const float fldj = 2.f * (0.69314718055994531f - x - softplus);
      lp = -0.5f * e * e - 0.91893853320467274f - ls - fldj;
      if (ls_raw > p.ls_min && ls_raw < p.ls_max) mask = 1u << d;
    }
    if (p.Xdst) store_operand(p.Xdst + static_cast<long long>(row) * p.ldx + d, p.xdst_lo_delta, a);
    if (p.act) p.act[row * A + d] = a;
    if (p.logstd) p.logstd[row * A + d] = ls;
  }
  lp = warp_sum(lp);
  for (int o = 16; o > 0; o >>= 1) mask |= __shfl_xor_sync(0xffffffffu, mask, o);
  if (lane == 0) {
    p.logp[row] = lp;
    if (p.inrange) p.inrange[row] = mask;
  }
}

// Same computation, one block per `rows_per_block` packed rows (all of one task; rows_per_block divides 128 and is a
// multiple of 16): the task's (W, 2A) head matrix is staged once per block in shared memory, transposed to
// [2A][W + 4] so that a lane reads 4 consecutive hidden units of one output as a conflict-free float4.  A warp works
// on two rows at a time (the weight reads are shared) with all of a row's float4 activation loads in flight.
constexpr int kAhBatch = 4;

template <int A>
static __global__ void __launch_bounds__(256, 3) actor_head_tile_kernel(const ActorHeadArgs p, int rows_per_block) {
  MTRL_PDL_PROLOGUE();
  extern __shared__ __align__(16) float sw[];  // [2A][W + 4]
  const int ldw = p.W + 4;
  const int row0 = blockIdx.x * rows_per_block;
  const int t = p.tile_task[row0 / kTileRows];
  const float* wsrc = p.Wh + static_cast<long long>(t) * p.W * (2 * A);
  if constexpr ((2 * A) % 4 == 0) {
    // batches of 8 independent 16-byte loads per thread (a plain load-store loop would serialise on the L2 latency)
    const int n4 = p.W * 2 * A / 4;
    for (int i0 = threadIdx.x; i0 < n4; i0 += 8 * blockDim.x) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * blockDim.x;
        v[u] = i < n4 ? __ldg(reinterpret_cast<const float4*>(wsrc) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * blockDim.x;
        if (i < n4) {
          const int k = (i * 4) / (2 * A), j = (i * 4) % (2 * A);
          sw[j * ldw + k] = v[u].x;
          sw[(j + 1) * ldw + k] = v[u].y;
          sw[(j + 2) * ldw + k] = v[u].z;
          sw[(j + 3) * ldw + k] = v[u].w;
        }
      }
    }
  } else {
    const int n = p.W * 2 * A;
    for (int i0 = threadIdx.x; i0 < n; i0 += 8 * blockDim.x) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * blockDim.x;
        v[u] = i < n ? __ldg(wsrc + i) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * blockDim.x;
        if (i < n) sw[(i % (2 * A)) * ldw + i / (2 * A)] = v[u];
      }
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int W4 = p.W / 4;
  for (int rp = row0 + 2 * warp; rp < row0 + rows_per_block && rp < p.M; rp += 2 * nw) {
    bool valid[2];
    float acc[2][2 * A];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      valid[u] = rp + u < p.M && p.slot_src[rp + u] >= 0;
#pragma unroll
      for (int j = 0; j < 2 * A; ++j) acc[u][j] = 0.f;
    }
    if (valid[0] || valid[1]) {
      const float4* h0 = reinterpret_cast<const float4*>(p.H + static_cast<long long>(rp) * p.W);
      const float4* h1 = reinterpret_cast<const float4*>(p.H + static_cast<long long>(rp + (rp + 1 < p.M ? 1 : 0)) * p.W);
      // kAhBatch float4 loads per row in flight per lane: small enough for three resident blocks per SM (85 registers),
      // which with 16-row blocks puts a 6400-row batch in ONE wave (400 blocks <= 3 x 148 slots)
      for (int kb = lane; kb < W4; kb += 32 * kAhBatch) {
        float4 a[kAhBatch], b[kAhBatch];
#pragma unroll
        for (int u = 0; u < kAhBatch; ++u) {
          const int k4 = kb + 32 * u;
          a[u] = k4 < W4 ? h0[k4] : make_float4(0.f, 0.f, 0.f, 0.f);
          b[u] = k4 < W4 ? h1[k4] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (p.h_lo_delta) {
          const float4* l0 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(h0) + p.h_lo_delta);
          const float4* l1 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(h1) + p.h_lo_delta);
#pragma unroll
          for (int u = 0; u < kAhBatch; ++u) {
            const int k4 = kb + 32 * u;
            if (k4 < W4) {
              const float4 x = l0[k4], y = l1[k4];
              a[u].x += x.x; a[u].y += x.y; a[u].z += x.z; a[u].w += x.w;
              b[u].x += y.x; b[u].y += y.y; b[u].z += y.z; b[u].w += y.w;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < kAhBatch; ++u) {
          const int k4 = kb + 32 * u;
          if (k4 < W4) {
#pragma unroll
            for (int j = 0; j < 2 * A; ++j) {
              const float4 wv = *reinterpret_cast<const float4*>(sw + j * ldw + k4 * 4);
              acc[0][j] = fmaf(a[u].x, wv.x, fmaf(a[u].y, wv.y, fmaf(a[u].z, wv.z, fmaf(a[u].w, wv.w, acc[0][j]))));
              acc[1][j] = fmaf(b[u].x, wv.x, fmaf(b[u].y, wv.y, fmaf(b[u].z, wv.z, fmaf(b[u].w, wv.w, acc[1][j]))));
            }
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int row = rp + u;
      if (row >= p.M) break;
#pragma unroll
      for (int j = 0; j < 2 * A; ++j) acc[u][j] = warp_sum(acc[u][j]);
      float lp = 0.f;
      unsigned mask = 0;
      if (lane < A) {
        const int d = lane;
        float mean = 0.f, ls_raw = 0.f;
#pragma unroll
        for (int j = 0; j < A; ++j) {
          if (j == d) { mean = acc[u][j]; ls_raw = acc[u][A + j]; }
        }
        float a = 0.f, ls = 0.f;
        if (valid[u]) {
          mean += p.bh[t * 2 * A + d];
          ls_raw += p.bh[t * 2 * A + A + d];
          ls = fminf(fmaxf(ls_raw, p.ls_min), p.ls_max);
          const float sd = expf(ls);
          const float e = p.eps[row * A + d];
          const float x = fmaf(sd, e, mean);
          a = tanhf(x);
          const float z = -2.f * x;
          const float softplus = z > 0.f ? z + log1pf(expf(-z)) : log1pf(expf(z));
          const float fldj = 2.f * (0.69314718055994531f - x - softplus);
          lp = -0.5f * e * e - 0.91893853320467274f - ls - fldj;
          if (ls_raw > p.ls_min && ls_raw < p.ls_max) mask = 1u << d;
        }
        if (p.Xdst) store_operand(p.Xdst + static_cast<long long>(row) * p.ldx + d, p.xdst_lo_delta, a);
        if (p.act) p.act[row * A + d] = a;
        if (p.logstd) p.logstd[row * A + d] = ls;
      }
      lp = warp_sum(lp);
      for (int o = 16; o > 0; o >>= 1) mask |= __shfl_xor_sync(0xffffffffu, mask, o);
      if (lane == 0) {
        p.logp[row] = lp;
        if (p.inrange) p.inrange[row] = mask;
      }
    }
  }
}

// Copy saved policy actions into the action columns of the critic input (single-task SAC samples them before the
// critic step, whose backward still needs the buffer actions there).
static __global__ void write_actions_kernel(const float* __restrict__ act, float* __restrict__ X, int ldx, int M, int A,
                                            long long lo_delta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M * A) store_operand(X + static_cast<long long>(i / A) * ldx + (i % A), lo_delta, act[i]);
}

// ---------------------------------------------------------------------------------------------
// Critic heads + losses.  One warp per packed row; E members.
// ---------------------------------------------------------------------------------------------
struct QHeads {
  const float* H[kMaxE];   // [M][W] last trunk activation of member e
  const float* w[kMaxE];   // (T_local, W, 1)
  const float* b[kMaxE];   // (T_local, 1)
  const float* pre[kMaxE]; // optional [M]: H_e . w_e already computed by the GEMM epilogue (fused heads); H is then not read
};

// N row-by-vector dot products at once (one warp, lanes stride the float4 columns): every pass issues the 2N loads of
// 4 column groups before any arithmetic, so a warp keeps 8N 16-byte loads in flight instead of one dependent pair.
__device__ __forceinline__ float4 ld_hi_lo(const float* h, int k4, long long lo_delta) {
  float4 a = reinterpret_cast<const float4*>(h)[k4];
  if (lo_delta) {
    const float4 l = reinterpret_cast<const float4*>(h + lo_delta)[k4];
    a.x += l.x; a.y += l.y; a.z += l.z; a.w += l.w;
  }
  return a;
}

template <int N>
__device__ __forceinline__ void row_dots(const float* const (&h)[N], const float* const (&w)[N], int W, int lane, float (&out)[N],
                                         long long h_lo_delta = 0) {
  float s[N];
#pragma unroll
  for (int n = 0; n < N; ++n) s[n] = 0.f;
  const int W4 = W / 4;
  int kb = lane;
  for (; kb + 96 < W4; kb += 128) {
    float4 a[N][4], c[N][4];
#pragma unroll
    for (int n = 0; n < N; ++n)
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[n][u] = ld_hi_lo(h[n], kb + 32 * u, h_lo_delta);
        c[n][u] = __ldg(reinterpret_cast<const float4*>(w[n]) + kb + 32 * u);
      }
#pragma unroll
    for (int n = 0; n < N; ++n)
#pragma unroll
      for (int u = 0; u < 4; ++u)
        s[n] = fmaf(a[n][u].x, c[n][u].x, fmaf(a[n][u].y, c[n][u].y, fmaf(a[n][u].z, c[n][u].z, fmaf(a[n][u].w, c[n][u].w, s[n]))));
  }
  for (; kb < W4; kb += 32) {
#pragma unroll
    for (int n = 0; n < N; ++n) {
      const float4 a = ld_hi_lo(h[n], kb, h_lo_delta);
      const float4 c = __ldg(reinterpret_cast<const float4*>(w[n]) + kb);
      s[n] = fmaf(a.x, c.x, fmaf(a.y, c.y, fmaf(a.z, c.z, fmaf(a.w, c.w, s[n]))));
    }
  }
#pragma unroll
  for (int n = 0; n < N; ++n) out[n] = warp_sum(s[n]);
}

__device__ __forceinline__ float row_dot(const float* __restrict__ h, const float* __restrict__ w, int W, int lane,
                                         long long h_lo_delta = 0) {
  const float* const hh[1] = {h};
  const float* const ww[1] = {w};
  float o[1];
  row_dots<1>(hh, ww, W, lane, o, h_lo_delta);
  return o[0];
}

struct CriticLossArgs {
  QHeads target, online;
  const int* tile_task;
  const int* slot_src;
  const float *rew, *done, *logp_next, *alpha_val, *task_w;
  float* dq;       // [E][M]
  double* acc;
  int M, W, E;
  float gamma, dq_scale;  // dL/dQ_e = dq_scale * w * (Q_e - y)
  int clip;
  long long h_lo_delta;   // fp32x3: remainders of the activations (see ActorHeadArgs)
};

// y = r + (1-d) gamma (min_e Qbar_e - alpha logp')   (mtsac.py:547-553)
// L = mean_{e,b} w (Q_e - y)^2                        (mtsac.py:562-565);  dq_e = dL/dQ_e
static __global__ void critic_loss_kernel(const CriticLossArgs p) {
  MTRL_PDL_PROLOGUE();
  __shared__ double red[32];
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  double loss = 0.0, qsum = 0.0;
  if (row < p.M) {
    const bool valid = p.slot_src[row] >= 0;
    const int t = p.tile_task[row / kTileRows];
    float qt_min = INFINITY;
    float q[kMaxE];
#pragma unroll
    for (int e = 0; e < kMaxE; ++e) {
      q[e] = 0.f;
      if (e < p.E && valid) {
        const long long ro = static_cast<long long>(row) * p.W, wo = static_cast<long long>(t) * p.W;
        const float* const hh[2] = {p.target.H[e] + ro, p.online.H[e] + ro};
        const float* const ww[2] = {p.target.w[e] + wo, p.online.w[e] + wo};
        float o[2];
        if (p.online.pre[e]) {
          o[0] = p.target.pre[e][row];
          o[1] = p.online.pre[e][row];
        } else {
          row_dots<2>(hh, ww, p.W, lane, o, p.h_lo_delta);
        }
        qt_min = fminf(qt_min, o[0] + p.target.b[e][t]);
        q[e] = o[1] + p.online.b[e][t];
      }
    }
    if (lane == 0) {
      if (valid) {
        const float w = p.task_w[t];
        float y = p.rew[row] + (1.f - p.done[row]) * p.gamma * (qt_min - p.alpha_val[t] * p.logp_next[row]);
        if (p.clip) y = fminf(fmaxf(y, -5000.f), 5000.f);
#pragma unroll
        for (int e = 0; e < kMaxE; ++e) {
          if (e < p.E) {
            float qc = q[e];
            float pass = 1.f;
            if (p.clip) {
              qc = fminf(fmaxf(qc, -5000.f), 5000.f);
              pass = (q[e] > -5000.f && q[e] < 5000.f) ? 1.f : 0.f;
            }
            const float diff = qc - y;
            loss += static_cast<double>(w * diff * diff);
            qsum += static_cast<double>(qc);
            p.dq[static_cast<long long>(e) * p.M + row] = p.dq_scale * w * diff * pass;
          }
        }
      } else {
        for (int e = 0; e < p.E; ++e) p.dq[static_cast<long long>(e) * p.M + row] = 0.f;
      }
    }
  }
  loss = block_sum(loss, red);
  qsum = block_sum(qsum, red);
  if (threadIdx.x == 0) {
    atomicAdd(p.acc + ACC_QLOSS, loss);
    atomicAdd(p.acc + ACC_QSUM, qsum);
  }
}

struct ActorLossArgs {
  QHeads online;
  const int* tile_task;
  const int* slot_src;
  const float *logp, *alpha_val, *task_w;
  float* dq;  // [E][M] seeds dL/dQ_e
  double* acc;
  int M, W, E;
  float inv_b;  // 1 / B_global
  long long h_lo_delta;
};

// L = mean_b w (alpha logp - min_e Q_e(s, a))   (mtsac.py:659-666); min routes the gradient to the arg-min.
static __global__ void actor_loss_kernel(const ActorLossArgs p) {
  MTRL_PDL_PROLOGUE();
  __shared__ double red[32];
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  double loss = 0.0;
  if (row < p.M) {
    const bool valid = p.slot_src[row] >= 0;
    const int t = p.tile_task[row / kTileRows];
    float q[kMaxE];
    float qmin = INFINITY;
#pragma unroll
    for (int e = 0; e < kMaxE; ++e) q[e] = INFINITY;
    const long long ro = static_cast<long long>(row) * p.W, wo = static_cast<long long>(t) * p.W;
    if (valid && p.online.pre[0]) {
#pragma unroll
      for (int e = 0; e < kMaxE; ++e) {
        if (e < p.E) {
          q[e] = p.online.pre[e][row] + p.online.b[e][t];
          qmin = fminf(qmin, q[e]);
        }
      }
    } else if (valid && p.E == 2) {   // the reference's ensemble size: both members' loads in flight together
      const float* const hh[2] = {p.online.H[0] + ro, p.online.H[1] + ro};
      const float* const ww[2] = {p.online.w[0] + wo, p.online.w[1] + wo};
      float o[2];
      row_dots<2>(hh, ww, p.W, lane, o, p.h_lo_delta);
      q[0] = o[0] + p.online.b[0][t];
      q[1] = o[1] + p.online.b[1][t];
      qmin = fminf(q[0], q[1]);
    } else if (valid) {
#pragma unroll
      for (int e = 0; e < kMaxE; ++e) {
        if (e < p.E) {
          q[e] = row_dot(p.online.H[e] + ro, p.online.w[e] + wo, p.W, lane, p.h_lo_delta) + p.online.b[e][t];
          qmin = fminf(qmin, q[e]);
        }
      }
    }
    if (lane == 0) {
      if (valid) {
        const float w = p.task_w[t];
        loss = static_cast<double>(w * (p.alpha_val[t] * p.logp[row] - qmin));
        int ties = 0;
#pragma unroll
        for (int e = 0; e < kMaxE; ++e) ties += (e < p.E && q[e] == qmin) ? 1 : 0;
        const float seed = -w * p.inv_b / static_cast<float>(ties);
#pragma unroll
        for (int e = 0; e < kMaxE; ++e)
          if (e < p.E) p.dq[static_cast<long long>(e) * p.M + row] = (q[e] == qmin) ? seed : 0.f;
      } else {
        for (int e = 0; e < p.E; ++e) p.dq[static_cast<long long>(e) * p.M + row] = 0.f;
      }
    }
  }
  loss = block_sum(loss, red);
  if (threadIdx.x == 0) atomicAdd(p.acc + ACC_ACTOR_LOSS, loss);
}

// dL/d(head output) of the actor from dL/da (through the critics) and dL/dlogp = alpha w / B
// (SURVEY Appendix A): gx = da (1 - a^2) + gl 2a; dmu = gx; dl = (gx sigma eps - gl) [l strictly inside the clip].
struct ActorDoutArgs {
  const float* dXin;   // [E][M][16], columns [0, A) = dL/da through member e
  const float *act, *logstd, *eps, *alpha_val, *task_w;
  const unsigned* inrange;
  const int* tile_task;
  const int* slot_src;
  float* dout;         // [M][2A]
  // split actor loss (gradient-surgery optimisers): the reference's vmapped call leaves `_explore` at its default True
  // (mtsac.py:631-637, 676-682), so every task's loss also carries  - mean_{rows, dims} (a_data - a_pi)^2
  const float* act_data;   // [M][A] batch actions, packed
  int explore;
  double* acc;
  int M, E, A;
  float inv_b;
};

static __global__ void actor_dout_kernel(const ActorDoutArgs p) {
  MTRL_PDL_PROLOGUE();
  __shared__ double red[32];
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  const int A = p.A;
  double ex = 0.0;
  if (row < p.M) {
    if (p.slot_src[row] < 0) {
      for (int j = 0; j < 2 * A; ++j) p.dout[row * 2 * A + j] = 0.f;
    } else {
      const int t = p.tile_task[row / kTileRows];
      const float gl = p.alpha_val[t] * p.task_w[t] * p.inv_b;
      const unsigned m = p.inrange[row];
      const float gex = 2.f * p.inv_b / static_cast<float>(A);   // d/da of -(1/(B A)) sum (a_data - a)^2 = gex (a_data - a)
      for (int d = 0; d < A; ++d) {
        float da = 0.f;
        for (int e = 0; e < p.E; ++e) da += p.dXin[(static_cast<long long>(e) * p.M + row) * 16 + d];
        const float a = p.act[row * A + d];
        if (p.explore) {
          const float diff = p.act_data[row * A + d] - a;
          da += gex * diff;
          ex += static_cast<double>(diff * diff);
        }
        const float gx = da * (1.f - a * a) + gl * 2.f * a;
        const float sd = expf(p.logstd[row * A + d]);
        p.dout[row * 2 * A + d] = gx;
        p.dout[row * 2 * A + A + d] = ((m >> d) & 1u) ? (gx * sd * p.eps[row * A + d] - gl) : 0.f;
      }
    }
  }
  if (p.explore) {
    ex = block_sum(ex, red);
    if (threadIdx.x == 0) atomicAdd(p.acc + ACC_EXPLORE, ex);
  }
}

// ---------------------------------------------------------------------------------------------
// Head VJP (nn.vmap(Dense) heads, multi_head.py:50-66, own-task rows only):
//   dZ[row,k]   = (sum_j dout[row,j] Wh[t,k,j]) * (H[row,k] > 0)      (masked by the trunk's last ReLU)
//   dWh[t,k,j]  = sum_{rows of t} H[row,k] dout[row,j];   dbh[t,j] = sum_{rows of t} dout[row,j]
// grid (W/128, T_local, E); block = RG warps.  A lane owns 4 consecutive hidden units (float4 traffic), warp w owns
// rows [w * 128/RG, (w+1) * 128/RG) of every 128-row tile of the task and keeps 8 row loads in flight; the per-warp
// partial sums (dWh, bias-gradient column sums) are combined through shared memory in a fixed order.
// HBM-bound: reads H and writes dZ once (2 * rows * W * 4 bytes per member).
// ---------------------------------------------------------------------------------------------
struct HeadBwdArgs {
  const float* H[kMaxE];
  const float* dout[kMaxE];   // [M][HD]
  const float* Wh[kMaxE];     // (T_local, W, HD)
  float* dZ[kMaxE];           // [M][W]
  float* dWh[kMaxE];          // may be null (no weight gradients)
  float* dbh[kMaxE];
  float* colsum[kMaxE];       // may be null: [M/128][W] column sums of dZ per 128-row tile (trunk bias gradient partials)
  const int* seg_start;
  long long dz_lo_delta;      // fp32x3: remainder of dZ at dZ[... + dz_lo_delta]; column sums then use the unrounded values
  int no_mask;                // the head's input is not a ReLU output (MLP with LayerNorm / skip, ln_kernels.cuh): dZ is the plain
                              // input gradient dout Wh^T in fp32 (no ReLU gate, no rounding, no remainder), consumed by the junction
  int M, W;
};

template <int HD, int RG>
static __global__ void __launch_bounds__(RG * 32, ((RG == 8 && HD <= 8) || HD <= 2) ? 2 : 1) head_bwd_kernel(const HeadBwdArgs p) {
  MTRL_PDL_PROLOGUE();
  constexpr int RPW = kTileRows / RG;   // rows per warp per tile
  static_assert(RPW % 8 == 0, "rows per warp must be a multiple of the 8-row load batch");
  __shared__ float sd[kTileRows * HD];
  __shared__ __align__(16) float red[RG][128];
  const int e = blockIdx.z, t = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = blockIdx.x * 128 + lane * 4;
  const bool kok = k < p.W;   // W is a multiple of 4
  const int r0 = p.seg_start[t], r1 = p.seg_start[t + 1];
  const float* H = p.H[e];
  const float* dout = p.dout[e];
  float* dZ = p.dZ[e];
  float w[HD][4], acc[HD][4];
#pragma unroll
  for (int j = 0; j < HD; ++j)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      acc[j][c] = 0.f;
      w[j][c] = kok ? p.Wh[e][(static_cast<long long>(t) * p.W + k + c) * HD + j] : 0.f;
    }
  float bsum = 0.f;  // threads j < HD of block x == 0 accumulate the bias gradient
  for (int base = r0; base < r1; base += kTileRows) {
    __syncthreads();
    for (int i = threadIdx.x; i < kTileRows * HD; i += blockDim.x) sd[i] = dout[static_cast<long long>(base) * HD + i];
    __syncthreads();
    float csum[4] = {0.f, 0.f, 0.f, 0.f};
    if (kok) {
#pragma unroll
      for (int c0 = 0; c0 < RPW; c0 += 8) {
        float4 h4[8];
        const int rr = warp * RPW + c0;
#pragma unroll
        for (int r = 0; r < 8; ++r) h4[r] = *reinterpret_cast<const float4*>(H + static_cast<long long>(base + rr + r) * p.W + k);
        if (p.dz_lo_delta) {   // fp32x3: the activation is hi + lo (both buffers mirror each other at dz_lo_delta)
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const float4 l = *reinterpret_cast<const float4*>(H + p.dz_lo_delta + static_cast<long long>(base + rr + r) * p.W + k);
            h4[r].x += l.x; h4[r].y += l.y; h4[r].z += l.z; h4[r].w += l.w;
          }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float hv[4] = {h4[r].x, h4[r].y, h4[r].z, h4[r].w};
          float dz[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < HD; ++j) {
            const float d = sd[(rr + r) * HD + j];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              acc[j][c] = fmaf(hv[c], d, acc[j][c]);
              dz[c] = fmaf(d, w[j][c], dz[c]);
            }
          }
          float dlo[4];
          float* dzp = dZ + static_cast<long long>(base + rr + r) * p.W + k;
          if (p.no_mask) {
            *reinterpret_cast<float4*>(dzp) = make_float4(dz[0], dz[1], dz[2], dz[3]);
            continue;
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float full = hv[c] > 0.f ? dz[c] : 0.f;
            dz[c] = tf32_rna(full);
            dlo[c] = tf32_lo(full, dz[c]);
            csum[c] += p.dz_lo_delta ? full : dz[c];
          }
          *reinterpret_cast<float4*>(dzp) = make_float4(dz[0], dz[1], dz[2], dz[3]);
          if (p.dz_lo_delta) *reinterpret_cast<float4*>(dzp + p.dz_lo_delta) = make_float4(dlo[0], dlo[1], dlo[2], dlo[3]);
        }
      }
    }
    if (p.colsum[e]) {
      *reinterpret_cast<float4*>(&red[warp][lane * 4]) = make_float4(csum[0], csum[1], csum[2], csum[3]);
      __syncthreads();
      if (threadIdx.x < 128 && blockIdx.x * 128 + threadIdx.x < p.W) {
        float s = 0.f;
#pragma unroll
        for (int g = 0; g < RG; ++g) s += red[g][threadIdx.x];
        p.colsum[e][static_cast<long long>(base / kTileRows) * p.W + blockIdx.x * 128 + threadIdx.x] = s;
      }
    }
    if (blockIdx.x == 0 && threadIdx.x < HD)
      for (int r = 0; r < kTileRows; ++r) bsum += sd[r * HD + threadIdx.x];
  }
  // Tiles past the last task's rows (max_rows leaves room for uneven batches) belong to nobody: their dZ rows and
  // column-sum partials must read as zero for the dW / dX GEMMs and the bias-gradient sum, which run over all max_rows
  // rows -- otherwise they would see whatever an earlier kernel left there (the partials buffer is shared with the
  // 32-row partials of the dX epilogue).  The last task's blocks clear them.
  if (t == gridDim.y - 1 && kok) {
    for (int base = r1; base + kTileRows <= p.M; base += kTileRows) {
      for (int r = warp; r < kTileRows; r += RG) {
        float* dzp = dZ + static_cast<long long>(base + r) * p.W + k;
        *reinterpret_cast<float4*>(dzp) = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.dz_lo_delta && !p.no_mask) *reinterpret_cast<float4*>(dzp + p.dz_lo_delta) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (p.colsum[e] && warp == 0)
        *reinterpret_cast<float4*>(p.colsum[e] + static_cast<long long>(base / kTileRows) * p.W + k) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  if (p.dWh[e]) {
#pragma unroll
    for (int j = 0; j < HD; ++j) {
      __syncthreads();
      *reinterpret_cast<float4*>(&red[warp][lane * 4]) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
      __syncthreads();
      const int kk = blockIdx.x * 128 + threadIdx.x;
      if (threadIdx.x < 128 && kk < p.W) {
        float s = 0.f;
#pragma unroll
        for (int g = 0; g < RG; ++g) s += red[g][threadIdx.x];
        p.dWh[e][(static_cast<long long>(t) * p.W + kk) * HD + j] = s;
      }
    }
    if (blockIdx.x == 0 && threadIdx.x < HD) p.dbh[e][t * HD + threadIdx.x] = bsum;
  }
}

// ---------------------------------------------------------------------------------------------
// Bias gradients: db[k] = sum over row groups of the partial column sums that the producers of dZ
// emit (head_bwd_kernel: one row of partials per 128-row tile; the dX GEMM epilogue: one per 32 rows).
// Fixed summation order, no float atomics.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxColsumJobs = 4 * kMaxE;   // bias gradients of every member and layer of a fused backward launch, or bias /
                                            // LayerNorm scale / LayerNorm bias gradients of every member of one layer
struct ColsumJobs {
  const float* part[kMaxColsumJobs];   // [groups][W]
  float* dst[kMaxColsumJobs];
  int groups[kMaxColsumJobs];          // 0: the launch's common group count
  int njobs;
};

// block = 32 columns x 8 group slices; slices are combined in a fixed order through shared memory.
static __global__ void colsum_final_kernel(const ColsumJobs jobs, int groups, int W) {
  MTRL_PDL_PROLOGUE();
  __shared__ float red[8][33];
  const int job = blockIdx.y;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + tx;
  const float* p = jobs.part[job];
  if (jobs.groups[job] > 0) groups = jobs.groups[job];
  float a = 0.f;
  if (k < W)
    for (int g = ty; g < groups; g += 8) a += p[static_cast<long long>(g) * W + k];
  red[ty][tx] = a;
  __syncthreads();
  if (ty == 0 && k < W) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += red[j][tx];
    jobs.dst[job][k] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// Per-task gradients (MTSAC.compute_weights, mtsac.py:870-1170: jax.vmap(jax.value_and_grad(loss)) over the batch split
// by task).  A row's dZ does not depend on other rows, so task t's gradient is the ordinary backward restricted to
// task t's rows: its dW comes from per-task GEMMs (sac.cu), and these two kernels fill in the bias and head slices of
// row t of the (T, P) gradient matrix (flat network layout).
// ---------------------------------------------------------------------------------------------
// tg[t][bias_off + k] = sum of the task's partial column sums; grid (ceil(W/256), T, E)
static __global__ void task_bias_kernel(const ColsumJobs parts, float* __restrict__ tg, long long row_stride, long long member_stride,
                                        long long bias_off, int groups_per_task, int W) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = blockIdx.y, e = blockIdx.z;
  if (k >= W) return;
  const float* p = parts.part[e] + static_cast<long long>(t) * groups_per_task * W;
  float s = 0.f;
  for (int g = 0; g < groups_per_task; ++g) s += p[static_cast<long long>(g) * W + k];
  tg[static_cast<long long>(t) * row_stride + e * member_stride + bias_off + k] = s;
}

// Head parameters of task t only receive gradient from task t: copy that slice of the (already per-task) head
// gradients into row t.  grid (T, E); n_k = W * head_dim, n_b = head_dim.
static __global__ void task_heads_kernel(const float* __restrict__ grads, float* __restrict__ tg, long long row_stride,
                                         long long heads_base, long long member_head_stride, long long head_kernel_off,
                                         long long head_bias_off, int n_k, int n_b) {
  const int t = blockIdx.x, e = blockIdx.y;
  const long long kbase = heads_base + e * member_head_stride + head_kernel_off + static_cast<long long>(t) * n_k;
  const long long bbase = heads_base + e * member_head_stride + head_bias_off + static_cast<long long>(t) * n_b;
  float* dst = tg + static_cast<long long>(t) * row_stride;
  for (int i = threadIdx.x; i < n_k; i += blockDim.x) dst[kbase + i] = grads[kbase + i];
  for (int i = threadIdx.x; i < n_b; i += blockDim.x) dst[bbase + i] = grads[bbase + i];
}

// status[0] = 3 unless every task owns exactly rows [t * R, (t + 1) * R) of the packed batch and n of them are real
// (balanced batch: the reference reshapes the task-sorted batch to (num_tasks, -1, dim), mtsac.py:325).
static __global__ void check_balanced_kernel(const int* __restrict__ seg_start, const int* __restrict__ slot_src, int T, int R, int n,
                                             int* __restrict__ status) {
  for (int t = threadIdx.x; t <= T; t += blockDim.x) {
    if (seg_start[t] != t * R) atomicExch(status, 3);
    if (t < T) {
      int cnt = 0;
      for (int r = t * R; r < (t + 1) * R; ++r) cnt += slot_src[r] >= 0 ? 1 : 0;
      if (cnt != n) atomicExch(status, 3);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// PCGrad (mtrl/optim/pcgrad.py:22-136) in coefficient space.  Every projected gradient stays in the span of the task
// gradients, g_i_pc = sum_k c_ik g_k, so the whole sequential projection loop (:58-69) only needs the Gram matrix
// G = g g^T: <g_i_pc, g_j> = sum_k c_ik G_kj.  One block, thread i owns row i of C (in shared memory).
//   gram   : (T, ldg) of the UNSCALED rows; gscale = T^2 turns it into the Gram of the reference's per-task-mean
//            gradients (the 1e-8 in the projection denominator is not scale free)
//   perm   : the row permutation of :79 (position -> task); identity when null
//   w_out  : (T) weights with  mean_i g_i_pc(reference scale) = sum_k w_out[k] * (unscaled row k)
//   stats  : [0] n_grad_conflicts (:71), [1] avg_grad_magnitude after surgery (:85), [2] before (:86-88),
//            [3] norm of the plain mean gradient (what the algorithm logs as grad magnitude, mtsac.py:583-585, 619)
// ---------------------------------------------------------------------------------------------
static __global__ void pcgrad_coeff_kernel(const float* __restrict__ gram, int ldg, int T, float gscale, const int* __restrict__ perm,
                                           float* __restrict__ w_out, float* __restrict__ stats) {
  extern __shared__ float pcg_sm[];   // G[T][T] (permuted, scaled) then C[T][T]
  float* G = pcg_sm;
  float* Cm = pcg_sm + T * T;
  __shared__ float red_conf[64], red_after[64], red_before[64];
  const int i = threadIdx.x;
  for (int idx = threadIdx.x; idx < T * T; idx += blockDim.x) {
    const int a = idx / T, b = idx % T;
    const int ta = perm ? perm[a] : a, tb = perm ? perm[b] : b;
    G[idx] = gram[static_cast<long long>(ta) * ldg + tb] * gscale;
    Cm[idx] = a == b ? 1.f : 0.f;
  }
  __syncthreads();
  float conf = 0.f, after = 0.f, before = 0.f;
  if (i < T) {
    float* c = Cm + i * T;
    for (int j = 0; j < T; ++j) {
      float dot = 0.f;
      for (int k = 0; k < T; ++k) dot = fmaf(c[k], G[k * T + j], dot);
      const float proj = dot / (G[j * T + j] + 1e-8f);
      if (proj < 0.f) {
        c[j] -= proj;
        conf += 1.f;
      }
    }
    float q = 0.f;
    for (int a = 0; a < T; ++a) {
      float r = 0.f;
      for (int b = 0; b < T; ++b) r = fmaf(G[a * T + b], c[b], r);
      q = fmaf(c[a], r, q);
    }
    after = sqrtf(fmaxf(q, 0.f));
    before = sqrtf(fmaxf(G[i * T + i], 0.f));
  }
  if (i < 64) { red_conf[i] = conf; red_after[i] = after; red_before[i] = before; }
  __syncthreads();
  if (i < T) {
    // column sums of C: position k of the permuted order is task perm[k]
    float wsum = 0.f;
    for (int r = 0; r < T; ++r) wsum += Cm[r * T + i];
    // mean over tasks (1 / T) of reference-scale rows (sqrt(gscale) x the given rows); inside the update gscale = T^2,
    // i.e. the plain sum of the given rows
    w_out[perm ? perm[i] : i] = wsum * (sqrtf(gscale) / static_cast<float>(T));
  }
  if (i == 0) {
    float sc = 0.f, sa = 0.f, sb = 0.f, all = 0.f;
    for (int r = 0; r < T; ++r) { sc += red_conf[r]; sa += red_after[r]; sb += red_before[r]; }
    for (int idx = 0; idx < T * T; ++idx) all += G[idx];
    stats[0] = sc * 0.5f;
    stats[1] = sa / T;
    stats[2] = sb / T;
    stats[3] = sqrtf(fmaxf(all, 0.f)) / T;
  }
}

// CAGrad (mtrl/optim/cagrad.py:20-237), also a function of the Gram matrix only: per-task clipping to unit norm
// (:181-187) scales rows, the 21-step momentum SGD on the T task weights (:56-123) works on GG = G / scale^2, and the
// result (:125-163) is a weighted sum of the clipped rows.  One thread, double precision (T <= 64: ~1e5 flops).
//   w_out : weights on the UNSCALED rows;  stats: [0] |g| of the result, [1] mean clipped-row norm, [2] best objective,
//   [3] unused;  tw_out (T): the softmax task weights (CAGradState.task_weights)
static __global__ void cagrad_coeff_kernel(const float* __restrict__ gram, int ldg, int T, float gscale, float c, int iterations,
                                           float lr, float momentum, float* __restrict__ w_out, float* __restrict__ stats,
                                           float* __restrict__ tw_out) {
  extern __shared__ float pcg_sm[];
  double* GG = reinterpret_cast<double*>(pcg_sm);          // [T][T]
  double* clipc = GG + T * T;                                // [T]
  double* Gg = clipc + T;                                    // [T]
  double* w = Gg + T;
  double* vel = w + T;
  double* wbest = vel + T;
  double* ww = wbest + T;
  double* d = ww + T;
  if (threadIdx.x != 0) return;
  // clipped, normalised Gram
  for (int i = 0; i < T; ++i) {
    const double n = sqrt(fmax(static_cast<double>(gram[i * ldg + i]) * gscale, 0.0));
    clipc[i] = fmin(1.0, 1.0 / (n + 1e-8));
  }
  double scale = 0.0, before = 0.0;
  for (int i = 0; i < T; ++i) {
    for (int j = 0; j < T; ++j) GG[i * T + j] = static_cast<double>(gram[i * ldg + j]) * gscale * clipc[i] * clipc[j];
    scale += sqrt(GG[i * T + i] + 1e-4);
    before += sqrt(fmax(GG[i * T + i], 0.0));
  }
  scale /= T;
  double gg = 0.0;
  for (int i = 0; i < T; ++i) {
    double r = 0.0;
    for (int j = 0; j < T; ++j) {
      GG[i * T + j] /= scale * scale;
      r += GG[i * T + j];
    }
    Gg[i] = r / T;
    gg += Gg[i];
  }
  gg /= T;
  const double cn = sqrt(gg + 1e-4) * c;
  auto objective = [&](const double* wv, bool want_grad, double* grad) {
    double s = 1e-8;
    for (int i = 0; i < T; ++i) s += wv[i];
    double t1 = 0.0;
    for (int i = 0; i < T; ++i) { ww[i] = wv[i] / s; t1 += ww[i] * Gg[i]; }
    double q = 0.0;
    for (int i = 0; i < T; ++i) {
      double r = 0.0;
      for (int j = 0; j < T; ++j) r += GG[i * T + j] * ww[j];
      d[i] = r;
      q += ww[i] * r;
    }
    const double root = sqrt(q + 1e-4);
    if (want_grad) {
      double dot = 0.0;
      for (int i = 0; i < T; ++i) { d[i] = Gg[i] + cn * d[i] / root; dot += d[i] * ww[i]; }
      for (int j = 0; j < T; ++j) grad[j] = (d[j] - dot) / s;
    }
    return t1 + cn * root;
  };
  for (int i = 0; i < T; ++i) { w[i] = 0.0; vel[i] = 0.0; wbest[i] = 0.0; }
  double obj_best = INFINITY;
  double* grad = d;   // objective() leaves the gradient in d
  for (int it = 0; it < iterations - 1; ++it) {
    const double o = objective(w, true, grad);
    if (o < obj_best) { obj_best = o; for (int i = 0; i < T; ++i) wbest[i] = w[i]; }
    for (int i = 0; i < T; ++i) { vel[i] = momentum * vel[i] + grad[i]; w[i] -= lr * vel[i]; }
  }
  {
    const double o = objective(w, false, nullptr);
    if (o < obj_best) { obj_best = o; for (int i = 0; i < T; ++i) wbest[i] = w[i]; }
  }
  // softmax(w_best) and the combination weights (:139-161)
  double mx = -INFINITY, den = 0.0;
  for (int i = 0; i < T; ++i) mx = fmax(mx, wbest[i]);
  for (int i = 0; i < T; ++i) { w[i] = exp(wbest[i] - mx); den += w[i]; }
  double q = 0.0;
  for (int i = 0; i < T; ++i) w[i] /= den;
  for (int i = 0; i < T; ++i) {
    double r = 0.0;
    for (int j = 0; j < T; ++j) r += GG[i * T + j] * w[j];
    q += w[i] * r;
  }
  const double lmbda = cn / (sqrt(q + 1e-4) + 1e-4);
  for (int i = 0; i < T; ++i) {
    const double comb = (1.0 / T + w[i] * lmbda) / (1.0 + static_cast<double>(c) * c);
    // reference-scale clipped row = clip_i * sqrt(gscale) * unscaled row
    vel[i] = comb * clipc[i] * sqrt(static_cast<double>(gscale));
    w_out[i] = static_cast<float>(vel[i]);
    tw_out[i] = static_cast<float>(w[i]);
  }
  double n2 = 0.0;
  for (int i = 0; i < T; ++i)
    for (int j = 0; j < T; ++j) n2 += vel[i] * vel[j] * static_cast<double>(gram[i * ldg + j]);
  stats[0] = static_cast<float>(sqrt(fmax(n2, 0.0)));
  stats[1] = static_cast<float>(before / T);
  stats[2] = static_cast<float>(obj_best);
  stats[3] = 0.f;
}

// GradNorm as the reference implements it (mtrl/optim/gradnorm.py:60-163): its gradnorm_loss does not depend on the task
// weights (:134-142), so their gradient is zero, Adam leaves them at their normalised initial value 1 (:72-80, 145-153)
// and the transformation returns sum_i 1 * g_i (:155-157) of the (optionally, max_grad_norm set) per-task clipped
// gradients (:40-57, 106-107).  stats: [0] norm of that sum, [1] mean per-task norm after clipping.
static __global__ void gradnorm_coeff_kernel(const float* __restrict__ gram, int ldg, int T, float gscale, int clip_per_task, int mean,
                                             float* __restrict__ w_out, float* __restrict__ stats) {
  __shared__ float wv[64];
  const int i = threadIdx.x;
  if (i < T) {
    const float n = sqrtf(fmaxf(gram[i * ldg + i] * gscale, 0.f));
    const float clipc = clip_per_task ? fminf(1.f, 1.f / (n + 1e-8f)) : 1.f;
    // weight 1 on the reference-scale row = sqrt(gscale) x the unscaled row; `mean`: the dummy multi-task optimiser's
    // plain average over tasks instead (mtrl/optim/dummy.py:18)
    wv[i] = clipc * sqrtf(gscale) / (mean ? static_cast<float>(T) : 1.f);
    w_out[i] = wv[i];
  }
  __syncthreads();
  if (i == 0) {
    double n2 = 0.0, before = 0.0;
    for (int a = 0; a < T; ++a) {
      before += static_cast<double>(wv[a]) * sqrt(fmax(static_cast<double>(gram[a * ldg + a]), 0.0));
      for (int b = 0; b < T; ++b) n2 += static_cast<double>(wv[a]) * wv[b] * gram[a * ldg + b];
    }
    stats[0] = static_cast<float>(sqrt(fmax(n2, 0.0)));
    stats[1] = static_cast<float>(before / T);
    stats[2] = 0.f;
    stats[3] = 0.f;
  }
}

// out[p] = sum_k w[k] * rows[k][p]   (HBM-bound: reads the (T, P) matrix once)
static __global__ void weighted_rows_kernel(const float* __restrict__ rows, long long ld, int T, const float* __restrict__ w,
                                            float* __restrict__ out, long long P) {
  __shared__ float sw[64];
  if (threadIdx.x < T) sw[threadIdx.x] = w[threadIdx.x];
  __syncthreads();
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < P / 4; i += stride) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int k = 0; k < T; ++k) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(rows + k * ld) + i);
      acc.x = fmaf(sw[k], v.x, acc.x); acc.y = fmaf(sw[k], v.y, acc.y);
      acc.z = fmaf(sw[k], v.z, acc.z); acc.w = fmaf(sw[k], v.w, acc.w);
    }
    reinterpret_cast<float4*>(out)[i] = acc;
  }
}

// Pairwise reductions over the columns of a (T, P) matrix, T <= 64:  out[a][b] += sum_p f(rows[a][p], rows[b][p]).
//   MODE 0: f = x y                                   Gram matrix (fp32 CUDA cores)
//   MODE 1: f = [|s x| < eps] [|s y| > tau]           compute_sparsity_mismatch (utils.py:75-91): a's near-zero elements
//                                                     that b updates strongly (s = row scale, counts exact per block)
// grid.x blocks of 256 threads; thread (a-group, b-group) register-tiles 4 x 4 entries over a shared (64 x 64) tile.
template <int MODE>
static __global__ void pairwise_kernel(const float* __restrict__ rows, long long ld, int T, long long P, float* __restrict__ out, int ldo,
                                       float s, float eps, float tau) {
  __shared__ float tile[64][65];
  // 4 x 4 register tiles.  MODE 1: 16 x 16 threads cover the 64 x 64 entries.  MODE 0 (symmetric): only the tiles on or
  // above the diagonal are computed -- enumerated densely over the first threads so that whole warps go idle -- and
  // mirrored at the end.
  int ta = (threadIdx.x / 16) * 4, tb = (threadIdx.x % 16) * 4;
  bool active = true;
  if (MODE == 0) {
    const int nt = (T + 3) / 4;
    int i = threadIdx.x, r = 0;
    while (r < nt && i >= nt - r) { i -= nt - r; ++r; }
    active = r < nt;
    ta = 4 * r;
    tb = 4 * (r + i);
  }
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  const long long per = ((P + gridDim.x - 1) / gridDim.x + 63) / 64 * 64;
  const long long p0 = per * blockIdx.x, p1 = min(P, p0 + per);
  for (long long c0 = p0; c0 < p1; c0 += 64) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < 64 * 64; idx += blockDim.x) {
      const int r = idx / 64, cc = idx % 64;
      // MODE 1: out-of-range entries must be neither "near zero" nor "large": NaN fails both comparisons
      tile[r][cc] = (r < T && c0 + cc < p1) ? rows[r * ld + c0 + cc] : (MODE == 0 ? 0.f : __int_as_float(0x7fc00000));
    }
    __syncthreads();
    if (!active) continue;
#pragma unroll 8
    for (int cc = 0; cc < 64; ++cc) {
      float va[4], vb[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) va[a] = tile[ta + a][cc];
#pragma unroll
      for (int b = 0; b < 4; ++b) vb[b] = tile[tb + b][cc];
      if (MODE == 0) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(va[a], vb[b], acc[a][b]);
      } else {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const float nz = fabsf(va[a] * s) < eps ? 1.f : 0.f;
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] += (fabsf(vb[b] * s) > tau) ? nz : 0.f;
        }
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if (active && ta + a < T && tb + b < T) {
        atomicAdd(&out[(ta + a) * ldo + tb + b], acc[a][b]);
        if (MODE == 0 && ta < tb) atomicAdd(&out[(tb + b) * ldo + ta + a], acc[a][b]);
      }
}

// compute_support_metrics (mtsac.py:774-860): with per-row support thresholds thr[t] (the 0.8-quantile of |x|),
//   out3[0][a][b] += #{p : |x_a| >= thr_a and |x_b| >= thr_b}             support intersection
//   out3[1][a][b] += #{p : x_a x_b < 0}                                   sign conflicts
//   out3[2][a][b] += #{p : both in support and x_a x_b < 0}              genuine conflicts
static __global__ void support_pairs_kernel(const float* __restrict__ rows, long long ld, int T, long long P, const float* __restrict__ thr,
                                            float* __restrict__ out3) {
  __shared__ float tile[64][65];
  __shared__ float sthr[64];
  if (threadIdx.x < 64) sthr[threadIdx.x] = threadIdx.x < T ? thr[threadIdx.x] : INFINITY;
  const int ta = (threadIdx.x / 16) * 4, tb = (threadIdx.x % 16) * 4;
  float inter[4][4], conf[4][4], genu[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) { inter[a][b] = 0.f; conf[a][b] = 0.f; genu[a][b] = 0.f; }
  const long long per = ((P + gridDim.x - 1) / gridDim.x + 63) / 64 * 64;
  const long long p0 = per * blockIdx.x, p1 = min(P, p0 + per);
  for (long long c0 = p0; c0 < p1; c0 += 64) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < 64 * 64; idx += blockDim.x) {
      const int r = idx / 64, cc = idx % 64;
      tile[r][cc] = (r < T && c0 + cc < p1) ? rows[r * ld + c0 + cc] : __int_as_float(0x7fc00000);   // NaN: in no set
    }
    __syncthreads();
#pragma unroll 4
    for (int cc = 0; cc < 64; ++cc) {
      float va[4], vb[4];
      bool sa[4], sb[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) { va[a] = tile[ta + a][cc]; sa[a] = fabsf(va[a]) >= sthr[ta + a]; }
#pragma unroll
      for (int b = 0; b < 4; ++b) { vb[b] = tile[tb + b][cc]; sb[b] = fabsf(vb[b]) >= sthr[tb + b]; }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const bool both = sa[a] && sb[b];
          const bool neg = va[a] * vb[b] < 0.f;
          inter[a][b] += both ? 1.f : 0.f;
          conf[a][b] += neg ? 1.f : 0.f;
          genu[a][b] += (both && neg) ? 1.f : 0.f;
        }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if (ta + a < T && tb + b < T) {
        const int o = (ta + a) * T + tb + b;
        atomicAdd(&out3[o], inter[a][b]);
        atomicAdd(&out3[T * T + o], conf[a][b]);
        atomicAdd(&out3[2 * T * T + o], genu[a][b]);
      }
}

// Order statistics of |x| per row by radix select on the float bits (non-negative floats order like their uint32 bits):
// 4 passes of one byte, most significant first.  state[t][s] = {prefix, remaining rank} for two targets s = 0, 1 per row
// (the two neighbours the linear-interpolation quantile needs).  select_hist_kernel: grid (blocks, T); histograms of the
// current byte among the elements matching each target's prefix.  select_scan_kernel: grid T, picks the bin.
static __global__ void select_hist_kernel(const float* __restrict__ rows, long long ld, long long P, int pass,
                                          const unsigned long long* __restrict__ state, unsigned int* __restrict__ hist) {
  __shared__ unsigned int sh[2][256];
  const int t = blockIdx.y;
  for (int i = threadIdx.x; i < 512; i += blockDim.x) (&sh[0][0])[i] = 0u;
  __syncthreads();
  const int shift = 24 - 8 * pass;
  const unsigned int mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
  const unsigned int pre0 = static_cast<unsigned int>(state[(t * 2 + 0) * 2]), pre1 = static_cast<unsigned int>(state[(t * 2 + 1) * 2]);
  const float* r = rows + t * ld;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < P; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const unsigned int key = __float_as_uint(fabsf(r[i]));
    const unsigned int b = (key >> shift) & 0xffu;
    if ((key & mask) == (pre0 & mask)) atomicAdd(&sh[0][b], 1u);
    if ((key & mask) == (pre1 & mask)) atomicAdd(&sh[1][b], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 512; i += blockDim.x) {
    const unsigned int v = (&sh[0][0])[i];
    if (v) atomicAdd(&hist[t * 512 + i], v);
  }
}

static __global__ void select_scan_kernel(int pass, unsigned long long* __restrict__ state, unsigned int* __restrict__ hist) {
  const int t = blockIdx.x, s = threadIdx.x;   // 2 threads: one per target
  if (s < 2) {
    unsigned int* h = hist + t * 512 + s * 256;
    unsigned long long rank = state[(t * 2 + s) * 2 + 1];
    const int shift = 24 - 8 * pass;
    unsigned int prefix = static_cast<unsigned int>(state[(t * 2 + s) * 2]);
    for (int b = 0; b < 256; ++b) {
      const unsigned int c = h[b];
      if (rank < c) { prefix |= static_cast<unsigned int>(b) << shift; break; }
      rank -= c;
    }
    state[(t * 2 + s) * 2] = prefix;
    state[(t * 2 + s) * 2 + 1] = rank;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 512; i += blockDim.x) hist[t * 512 + i] = 0u;   // ready for the next pass
}

static __global__ void order_stats_out_kernel(const unsigned long long* __restrict__ state, int T, float* __restrict__ out2) {
  const int t = threadIdx.x;
  if (t < T) {
    out2[2 * t] = __uint_as_float(static_cast<unsigned int>(state[(t * 2 + 0) * 2]));
    out2[2 * t + 1] = __uint_as_float(static_cast<unsigned int>(state[(t * 2 + 1) * 2]));
  }
}

// Per-row sums over the columns: out[t] = {sum |s x|, count(|s x| < eps)}   (participation ratio, near-zero counts:
// utils.py:86-87, 94-101).  grid (blocks, T)
static __global__ void row_stats_kernel(const float* __restrict__ rows, long long ld, long long P, float s, float eps,
                                        double* __restrict__ out2) {
  __shared__ double red[32];
  const int t = blockIdx.y;
  const float* r = rows + t * ld;
  double l1 = 0.0, nz = 0.0;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < P; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = fabsf(r[i] * s);
    l1 += v;
    nz += v < eps ? 1.0 : 0.0;
  }
  l1 = block_sum(l1, red);
  nz = block_sum(nz, red);
  if (threadIdx.x == 0) {
    atomicAdd(&out2[2 * t], l1);
    atomicAdd(&out2[2 * t + 1], nz);
  }
}

// ---------------------------------------------------------------------------------------------
// Optimiser: optax.chain(clip_by_global_norm, adam) + apply_updates (config/optim.py:26-43,
// algorithms/utils.py:11-46) over the flat parameter buffer of a network, fused with the Polyak
// target update (mtsac.py:607-613) and the tf32 operand copies the GEMMs read.
// ---------------------------------------------------------------------------------------------
static __global__ void sumsq_kernel(const float* __restrict__ x, long long n, double* __restrict__ acc) {
  MTRL_PDL_PROLOGUE();
  __shared__ double red[32];
  double s = 0.0;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long n4 = n / 4;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  // four independent 16-byte loads in flight per thread (one per iteration leaves the kernel latency-bound)
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = x4[i + u * stride];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      s += static_cast<double>(v[u].x * v[u].x + v[u].y * v[u].y) + static_cast<double>(v[u].z * v[u].z + v[u].w * v[u].w);
  }
  for (; i < n4; i += stride) {
    const float4 v = x4[i];
    s += static_cast<double>(v.x * v.x + v.y * v.y) + static_cast<double>(v.z * v.z + v.w * v.w);
  }
  for (long long j = n4 * 4 + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; j < n; j += stride)
    s += static_cast<double>(x[j] * x[j]);
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(acc, s);
}

// grads[slot] = local head-gradient squared norm, so one all-reduce of [trunk | slots] carries it.
static __global__ void write_slot_kernel(float* __restrict__ slot, const double* __restrict__ acc) {
  MTRL_PDL_PROLOGUE();
  *slot = static_cast<float>(*acc);
}

struct AdamArgs {
  float *p, *m, *v, *shadow;
  const float* g;
  float *target, *target_shadow;     // null for networks without a target
  float *shadow_lo, *target_shadow_lo;   // fp32x3: tf32 remainders of the operand copies (null in tf32 mode)
  long long n;
  long long trunk_n;                 // elements [0, trunk_n) count into the trunk param norm, the rest into the head norm;
                                     // [trunk_n, trunk_n + 32) are the reduction slots, not parameters
  const double* g2_trunk;            // squared norm of the (all-reduced) trunk gradients
  const float* g2_heads;             // all-reduced slot: squared norm of every rank's head gradients
  const int* step;                   // Adam count before this step
  double* p2_trunk;
  double* p2_head;
  double* p2_old;                    // squared norm of the parameters before the step
  float lr, b1, b2, eps, max_norm, tau;
};

static __global__ void __launch_bounds__(256, 4) adam_kernel(const AdamArgs a) {
  MTRL_PDL_PROLOGUE();
  __shared__ double red[32];
  const double g2 = *a.g2_trunk + static_cast<double>(*a.g2_heads);
  const float gn = static_cast<float>(sqrt(g2));
  // optax.clip_by_global_norm: g if norm < max else g / norm * max
  const float scale = (a.max_norm > 0.f && !(gn < a.max_norm)) ? a.max_norm / gn : 1.f;
  const int t = *a.step + 1;
  const float bc1 = static_cast<float>(1.0 - pow(static_cast<double>(a.b1), static_cast<double>(t)));
  const float bc2 = static_cast<float>(1.0 - pow(static_cast<double>(a.b2), static_cast<double>(t)));
  double s_trunk = 0.0, s_head = 0.0, s_old = 0.0;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  // the flat buffers are 128-byte aligned and every region (trunk, 32 slots, heads) is a multiple of 32 floats
  const long long n4 = a.n / 4, slot4 = a.trunk_n / 4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    if (i >= slot4 && i < slot4 + 8) continue;
    const float4 g4 = reinterpret_cast<const float4*>(a.g)[i];
    const float4 m4 = reinterpret_cast<const float4*>(a.m)[i];
    const float4 v4 = reinterpret_cast<const float4*>(a.v)[i];
    const float4 p4 = reinterpret_cast<const float4*>(a.p)[i];
    float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.target) t4 = reinterpret_cast<const float4*>(a.target)[i];
    const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
    float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w}, pp[4] = {p4.x, p4.y, p4.z, p4.w};
    float tt[4] = {t4.x, t4.y, t4.z, t4.w}, sh[4], tsh[4], shl[4], tshl[4];
    float sq = 0.f;
    s_old += static_cast<double>(p4.x * p4.x + p4.y * p4.y + p4.z * p4.z + p4.w * p4.w);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float g = gg[q] * scale;
      mm[q] = a.b1 * mm[q] + (1.f - a.b1) * g;
      vv[q] = a.b2 * vv[q] + (1.f - a.b2) * g * g;
      pp[q] = pp[q] - a.lr * (mm[q] / bc1) / (sqrtf(vv[q] / bc2) + a.eps);
      sh[q] = tf32_rna(pp[q]);
      shl[q] = tf32_lo(pp[q], sh[q]);
      tt[q] = a.tau * pp[q] + (1.f - a.tau) * tt[q];
      tsh[q] = tf32_rna(tt[q]);
      tshl[q] = tf32_lo(tt[q], tsh[q]);
      sq += pp[q] * pp[q];
    }
    reinterpret_cast<float4*>(a.m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    reinterpret_cast<float4*>(a.v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    reinterpret_cast<float4*>(a.p)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
    if (a.shadow) reinterpret_cast<float4*>(a.shadow)[i] = make_float4(sh[0], sh[1], sh[2], sh[3]);
    if (a.shadow_lo) reinterpret_cast<float4*>(a.shadow_lo)[i] = make_float4(shl[0], shl[1], shl[2], shl[3]);
    if (a.target) {
      reinterpret_cast<float4*>(a.target)[i] = make_float4(tt[0], tt[1], tt[2], tt[3]);
      if (a.target_shadow) reinterpret_cast<float4*>(a.target_shadow)[i] = make_float4(tsh[0], tsh[1], tsh[2], tsh[3]);
      if (a.target_shadow_lo) reinterpret_cast<float4*>(a.target_shadow_lo)[i] = make_float4(tshl[0], tshl[1], tshl[2], tshl[3]);
    }
    if (i < slot4) s_trunk += static_cast<double>(sq);
    else s_head += static_cast<double>(sq);
  }
  s_trunk = block_sum(s_trunk, red);
  s_head = block_sum(s_head, red);
  s_old = block_sum(s_old, red);
  if (threadIdx.x == 0) {
    atomicAdd(a.p2_trunk, s_trunk);
    atomicAdd(a.p2_head, s_head);
    atomicAdd(a.p2_old, s_old);
  }
}

// Polyak target update alone (mtsac.py:607-613): target = tau p + (1 - tau) target and its tf32 operand copies, for the whole
// flat network except the reduction slots.  Launched on a side stream right after the critic's Adam step, so that it runs
// under the actor-phase GEMMs instead of on the critical path (nothing reads the target before the next update): few
// registers and no shared memory, so a block fits on an SM beside a resident GEMM CTA; streaming loads / stores keep the 4 x
// P bytes it moves from evicting the GEMMs' operand tiles from L2.
static __global__ void __launch_bounds__(256) polyak_kernel(const float* __restrict__ p, float* __restrict__ target,
                                                            float* __restrict__ tsh, float* __restrict__ tsh_lo, long long n,
                                                            long long trunk_n, float tau) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long n4 = n / 4, slot4 = trunk_n / 4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    if (i >= slot4 && i < slot4 + 8) continue;
    const float4 p4 = __ldcs(reinterpret_cast<const float4*>(p) + i);
    float4 t4 = __ldcs(reinterpret_cast<const float4*>(target) + i);
    t4.x = tau * p4.x + (1.f - tau) * t4.x;
    t4.y = tau * p4.y + (1.f - tau) * t4.y;
    t4.z = tau * p4.z + (1.f - tau) * t4.z;
    t4.w = tau * p4.w + (1.f - tau) * t4.w;
    __stcs(reinterpret_cast<float4*>(target) + i, t4);
    const float4 h4 = make_float4(tf32_rna(t4.x), tf32_rna(t4.y), tf32_rna(t4.z), tf32_rna(t4.w));
    reinterpret_cast<float4*>(tsh)[i] = h4;   // read by the next update's first target GEMM: may stay in L2
    if (tsh_lo)
      __stcs(reinterpret_cast<float4*>(tsh_lo) + i,
             make_float4(tf32_lo(t4.x, h4.x), tf32_lo(t4.y, h4.y), tf32_lo(t4.z, h4.z), tf32_lo(t4.w, h4.w)));
  }
}

static __global__ void shadow_kernel(const float* __restrict__ p, float* __restrict__ s, float* __restrict__ s_lo, long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float hi = tf32_rna(p[i]);
    s[i] = hi;
    if (s_lo) s_lo[i] = tf32_lo(p[i], hi);
  }
}

static __global__ void step_inc_kernel(int* step) { *step += 1; }

// Single thread: turn accumulators into the reference's log scalars and advance the Adam count.
static __global__ void finalize_critic_kernel(const double* acc, const float* g2_heads, int* steps, float* logs, float inv_eb,
                                       float loss_scale, int log_old_norm) {
  MTRL_PDL_PROLOGUE();
  logs[MTRL_LOG_QF_VALUES] = static_cast<float>(acc[ACC_QSUM] * inv_eb);
  logs[MTRL_LOG_QF_LOSS] = static_cast<float>(acc[ACC_QLOSS] * loss_scale);
  logs[MTRL_LOG_CRITIC_GRAD_MAGNITUDE] = static_cast<float>(sqrt(acc[ACC_CRITIC_G2] + static_cast<double>(*g2_heads)));
  // mtsac.py:614,620 logs the norm of the UPDATED critic; sac.py:363-364 ravels self.critic.params, the old ones
  logs[MTRL_LOG_CRITIC_PARAMS_NORM] = static_cast<float>(
      sqrt(log_old_norm ? acc[ACC_CRITIC_P2_OLD] : acc[ACC_CRITIC_P2_TRUNK] + acc[ACC_CRITIC_P2_HEAD]));
  logs[LOG_X_CRITIC_P2_TRUNK] = static_cast<float>(acc[ACC_CRITIC_P2_TRUNK]);
  logs[LOG_X_CRITIC_P2_HEAD] = static_cast<float>(acc[ACC_CRITIC_P2_HEAD]);
  steps[1] += 1;
}
static __global__ void finalize_actor_kernel(const double* acc, const float* g2_heads, int* steps, float* logs, float inv_b,
                                      int log_old_norm, float inv_a) {
  MTRL_PDL_PROLOGUE();
  // split actor loss: every task's loss is reduced by its explore term (ACC_EXPLORE stays 0 otherwise); the log is the
  // mean over tasks (mtsac.py:704 actor_loss_value.mean(); the reference's explore_loss log is the per-task vector, whose
  // mean this is)
  const double explore = acc[ACC_EXPLORE] * inv_b * inv_a;
  logs[MTRL_LOG_ACTOR_LOSS] = static_cast<float>(acc[ACC_ACTOR_LOSS] * inv_b - explore);
  logs[MTRL_LOG_ACTOR_GRAD_MAGNITUDE] = static_cast<float>(sqrt(acc[ACC_ACTOR_G2] + static_cast<double>(*g2_heads)));
  logs[MTRL_LOG_ACTOR_PARAMS_NORM] = static_cast<float>(
      sqrt(log_old_norm ? acc[ACC_ACTOR_P2_OLD] : acc[ACC_ACTOR_P2_TRUNK] + acc[ACC_ACTOR_P2_HEAD]));
  logs[LOG_X_ACTOR_P2_TRUNK] = static_cast<float>(acc[ACC_ACTOR_P2_TRUNK]);
  logs[LOG_X_ACTOR_P2_HEAD] = static_cast<float>(acc[ACC_ACTOR_P2_HEAD]);
  logs[MTRL_LOG_EXPLORE_LOSS] = static_cast<float>(explore);  // 0 on the non-split path: explore=False (mtsac.py:277, 671)
  steps[0] += 1;
  steps[3] += 1;  // noise counter
}

// Temperature step (mtsac.py:713-731): L = mean_b -log_alpha[task_b] (logp_b + target_entropy); Adam on log_alpha.
// One block; warp per task over that task's packed rows.
struct AlphaArgs {
  float *log_alpha, *m, *v;
  const float* logp;
  const int* seg_start;
  const int* slot_src;
  int* steps;
  float* logs;
  int T_local;
  float target_entropy, inv_b, lr, b1, b2, eps, max_norm;
};

static __global__ void alpha_step_kernel(const AlphaArgs a) {
  MTRL_PDL_PROLOGUE();
  extern __shared__ float sg[];  // [T_local] gradients
  __shared__ float red[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float loss = 0.f;
  for (int t = warp; t < a.T_local; t += nw) {
    float s = 0.f;
    for (int r = a.seg_start[t] + lane; r < a.seg_start[t + 1]; r += 32)
      if (a.slot_src[r] >= 0) s += a.logp[r] + a.target_entropy;
    s = warp_sum(s);
    if (lane == 0) {
      sg[t] = -s * a.inv_b;
      loss += -a.log_alpha[t] * s * a.inv_b;
    }
  }
  loss = block_sum(loss, red);
  __shared__ float sh_loss, sh_scale;
  if (threadIdx.x == 0) sh_loss = loss;
  __syncthreads();
  float g2 = 0.f;
  for (int t = threadIdx.x; t < a.T_local; t += blockDim.x) g2 += sg[t] * sg[t];
  g2 = block_sum(g2, red);
  if (threadIdx.x == 0) {
    const float gn = sqrtf(g2);
    sh_scale = (a.max_norm > 0.f && !(gn < a.max_norm)) ? a.max_norm / gn : 1.f;
  }
  __syncthreads();
  const int tstep = a.steps[2] + 1;
  const float bc1 = static_cast<float>(1.0 - pow(static_cast<double>(a.b1), static_cast<double>(tstep)));
  const float bc2 = static_cast<float>(1.0 - pow(static_cast<double>(a.b2), static_cast<double>(tstep)));
  float asum = 0.f;
  for (int t = threadIdx.x; t < a.T_local; t += blockDim.x) {
    const float g = sg[t] * sh_scale;
    const float m = a.b1 * a.m[t] + (1.f - a.b1) * g;
    const float v = a.b2 * a.v[t] + (1.f - a.b2) * g * g;
    const float la = a.log_alpha[t] - a.lr * (m / bc1) / (sqrtf(v / bc2) + a.eps);
    a.m[t] = m;
    a.v[t] = v;
    a.log_alpha[t] = la;
    asum += expf(la);
  }
  asum = block_sum(asum, red);
  __syncthreads();
  if (threadIdx.x == 0) {
    a.logs[MTRL_LOG_ALPHA_LOSS] = sh_loss;
    a.logs[MTRL_LOG_ALPHA] = asum;
    a.steps[2] = tstep;
  }
}

}  // namespace sac
