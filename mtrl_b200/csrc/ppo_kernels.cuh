// Loss / head kernels of the multi-task PPO update (/root/reference/mtrl/rl/algorithms/mtppo.py:196-290).
// Everything else (trunk GEMMs, head VJPs, bias gradients, Adam) is shared with the SAC path.
#pragma once

#include <curand_kernel.h>

#include "common.cuh"
#include "mtrl_b200.h"
#include "sac_kernels.cuh"

namespace ppo {

using sac::block_sum;
using sac::kTileRows;

enum { PACC_ADV_SUM = 0, PACC_ADV_SQ, PACC_PG, PACC_ENT, PACC_KL, PACC_CLIPFRAC, PACC_VF, PACC_VSUM,
       PACC_P_G2, PACC_P_HEAD_G2, PACC_V_G2, PACC_V_HEAD_G2, PACC_SCRATCH0, PACC_SCRATCH1, PACC_SCRATCH2, PACC_SCRATCH3,
       PACC_COUNT = 16 };

struct PackArgs {
  const float *obs, *logp, *adv, *ret, *val, *eps;   // rollout flattened task-major: row = task * steps + step
  float *X, *plogp, *padv, *pret, *pval, *peps;
  int* slot_src;
  const int* noise_counter;
  unsigned long long seed;
  int obs_dim, act_dim, K, steps, steps_pad;
};

// Sum and sum of squares of the advantages (global normalisation, mtppo.py:220-225).
static __global__ void adv_stats_kernel(const float* __restrict__ adv, long long n, double* __restrict__ acc) {
  __shared__ double red[32];
  double s = 0.0, q = 0.0;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double v = adv[i];
    s += v;
    q += v * v;
  }
  s = block_sum(s, red);
  q = block_sum(q, red);
  if (threadIdx.x == 0) {
    atomicAdd(acc + PACC_ADV_SUM, s);
    atomicAdd(acc + PACC_ADV_SQ, q);
  }
}

// One block per packed row (slot = task * steps_pad + step).
static __global__ void pack_kernel(const PackArgs a) {
  const int slot = blockIdx.x;
  const int t = slot / a.steps_pad, s = slot % a.steps_pad;
  const bool valid = s < a.steps;
  const long long src = static_cast<long long>(t) * a.steps + s;
  float* x = a.X + static_cast<long long>(slot) * a.K;
  if (!valid) {
    for (int j = threadIdx.x; j < a.K; j += blockDim.x) x[j] = 0.f;
    if (threadIdx.x < a.act_dim) a.peps[slot * a.act_dim + threadIdx.x] = 0.f;
    if (threadIdx.x == 0) {
      a.slot_src[slot] = -1;
      a.plogp[slot] = 0.f; a.padv[slot] = 0.f; a.pret[slot] = 0.f; a.pval[slot] = 0.f;
    }
    return;
  }
  const float* o = a.obs + src * a.obs_dim;
  for (int j = threadIdx.x; j < a.K; j += blockDim.x) x[j] = j < a.obs_dim ? tf32_rna(o[j]) : 0.f;
  if (threadIdx.x == 0) {
    a.slot_src[slot] = static_cast<int>(src);
    a.plogp[slot] = a.logp[src];
    const float ad = a.adv[src];
    a.padv[slot] = ad;
    a.pret[slot] = a.ret[src];
    a.pval[slot] = a.val[src];
    if (a.eps) {
      for (int d = 0; d < a.act_dim; ++d) a.peps[slot * a.act_dim + d] = a.eps[src * a.act_dim + d];
    } else {
      curandStatePhilox4_32_10_t st;
      curand_init(a.seed, static_cast<unsigned long long>(src), static_cast<unsigned long long>(*a.noise_counter) * 16ull, &st);
      for (int d = 0; d < a.act_dim; d += 4) {
        const float4 n = curand_normal4(&st);
        const float nn[4] = {n.x, n.y, n.z, n.w};
        for (int q = 0; q < 4 && d + q < a.act_dim; ++q) a.peps[slot * a.act_dim + d + q] = nn[q];
      }
    }
  }
}

struct PolicyLossArgs {
  const float* H;        // [M][W] last trunk activation of the policy
  const float* Wh;       // (T, W, 2A)
  const float* bh;       // (T, 2A)
  const int* tile_task;
  const int* slot_src;
  const float *eps, *old_logp, *adv;
  float* dout;           // [M][2A]  dL/d(head output)
  double* acc;
  int M, W;
  float ls_min, ls_max, clip_eps, ent_coef, inv_b, n_rows;
  int normalize;
};

// Clipped surrogate on the log-prob of a fresh sample + entropy bonus (mtppo.py:196-239), and its gradient with
// respect to the head output.  One block per 32 packed rows of one task, head weights staged in shared memory.
template <int A>
static __global__ void policy_loss_kernel(const PolicyLossArgs p) {
  extern __shared__ float sw[];  // [W][2A]
  __shared__ double red[32];
  const int row0 = blockIdx.x * 32;
  const int t = p.tile_task[row0 / kTileRows];
  const float* wsrc = p.Wh + static_cast<long long>(t) * p.W * (2 * A);
  for (int i = threadIdx.x; i < p.W * 2 * A / 4; i += blockDim.x)
    reinterpret_cast<float4*>(sw)[i] = __ldg(reinterpret_cast<const float4*>(wsrc) + i);
  __syncthreads();
  // advantage normalisation constants: population std, like jnp.std
  const double mean_d = p.acc[PACC_ADV_SUM] / p.n_rows;
  const double var_d = fmax(p.acc[PACC_ADV_SQ] / p.n_rows - mean_d * mean_d, 0.0);
  const float a_mean = static_cast<float>(mean_d), a_std = static_cast<float>(sqrt(var_d));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  double s_pg = 0.0, s_ent = 0.0, s_kl = 0.0, s_cf = 0.0;
  for (int row = row0 + warp; row < row0 + 32 && row < p.M; row += nw) {
    const bool valid = p.slot_src[row] >= 0;
    float acc[2 * A];
#pragma unroll
    for (int j = 0; j < 2 * A; ++j) acc[j] = 0.f;
    if (valid) {
      const float* h = p.H + static_cast<long long>(row) * p.W;
      for (int k = lane; k < p.W; k += 32) {
        const float hv = h[k];
        const float* wk = sw + k * (2 * A);
#pragma unroll
        for (int j = 0; j < 2 * A; ++j) acc[j] = fmaf(hv, wk[j], acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 2 * A; ++j) acc[j] = sac::warp_sum_f(acc[j]);
    float nl = 0.f, ent = 0.f;
    bool inr = false;
    if (lane < A && valid) {
      float ls_raw = 0.f;
#pragma unroll
      for (int j = 0; j < A; ++j)
        if (j == lane) ls_raw = acc[A + j];
      ls_raw += p.bh[t * 2 * A + A + lane];
      const float ls = fminf(fmaxf(ls_raw, p.ls_min), p.ls_max);
      const float e = p.eps[row * A + lane];
      nl = -0.5f * e * e - 0.91893853320467274f - ls;
      ent = ls + 1.4189385332046727f;  // log(2 pi e) / 2
      inr = ls_raw > p.ls_min && ls_raw < p.ls_max;
    }
    nl = sac::warp_sum_f(nl);
    ent = sac::warp_sum_f(ent);
    float g_nl = 0.f;
    if (valid) {
      const float lr = nl - p.old_logp[row];
      const float r = expf(lr);
      const float ah = p.normalize ? (p.adv[row] - a_mean) / (a_std + 1e-8f) : p.adv[row];
      const float rc = fminf(fmaxf(r, 1.f - p.clip_eps), 1.f + p.clip_eps);
      const float l1 = -ah * r, l2 = -ah * rc;
      const bool inside = r > 1.f - p.clip_eps && r < 1.f + p.clip_eps;
      g_nl = (l1 >= l2 || inside) ? -ah * r * p.inv_b : 0.f;
      if (lane == 0) {
        s_pg += static_cast<double>(fmaxf(l1, l2));
        s_ent += static_cast<double>(ent);
        s_kl += static_cast<double>((r - 1.f) - lr);
        s_cf += fabsf(r - 1.f) > p.clip_eps ? 1.0 : 0.0;
      }
    }
    if (lane < A) {
      p.dout[row * 2 * A + lane] = 0.f;                                                   // the sample's log-prob ignores the mean
      p.dout[row * 2 * A + A + lane] = (valid && inr) ? (-g_nl - p.ent_coef * p.inv_b) : 0.f;  // d nl / d log_std = -1
    }
  }
  s_pg = block_sum(s_pg, red);
  s_ent = block_sum(s_ent, red);
  s_kl = block_sum(s_kl, red);
  s_cf = block_sum(s_cf, red);
  if (threadIdx.x == 0) {
    atomicAdd(p.acc + PACC_PG, s_pg);
    atomicAdd(p.acc + PACC_ENT, s_ent);
    atomicAdd(p.acc + PACC_KL, s_kl);
    atomicAdd(p.acc + PACC_CLIPFRAC, s_cf);
  }
}

struct ValueLossArgs {
  const float* H;   // [M][W]
  const float* w;   // (T, W, 1)
  const float* b;   // (T, 1)
  const int* tile_task;
  const int* slot_src;
  const float *ret, *old_val;
  float* dq;        // [M]
  double* acc;
  int M, W;
  float clip_eps, vf_coef, inv_b;
  int clip;
};

// Clipped value loss (mtppo.py:256-277): one warp per packed row.
static __global__ void value_loss_kernel(const ValueLossArgs p) {
  __shared__ double red[32];
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  double s_l = 0.0, s_v = 0.0;
  if (row < p.M) {
    const bool valid = p.slot_src[row] >= 0;
    float v = 0.f;
    if (valid) {
      const int t = p.tile_task[row / kTileRows];
      v = sac::row_dot(p.H + static_cast<long long>(row) * p.W, p.w + static_cast<long long>(t) * p.W, p.W, lane) + p.b[t];
    }
    if (lane == 0) {
      float dv = 0.f;
      if (valid) {
        const float R = p.ret[row], vo = p.old_val[row];
        const float un = (v - R) * (v - R);
        float loss = un, g = v - R;
        if (p.clip) {
          const float d = v - vo;
          const float vc = vo + fminf(fmaxf(d, -p.clip_eps), p.clip_eps);
          const float cl = (vc - R) * (vc - R);
          if (cl > un) {
            loss = cl;
            g = (d > -p.clip_eps && d < p.clip_eps) ? (vc - R) : 0.f;
          }
        }
        s_l = 0.5 * static_cast<double>(loss);
        s_v = static_cast<double>(v);
        dv = p.vf_coef * p.inv_b * g;
      }
      p.dq[row] = dv;
    }
  }
  s_l = block_sum(s_l, red);
  s_v = block_sum(s_v, red);
  if (threadIdx.x == 0) {
    atomicAdd(p.acc + PACC_VF, s_l);
    atomicAdd(p.acc + PACC_VSUM, s_v);
  }
}

static __global__ void finalize_kernel(const double* acc, int* steps, float* logs, float inv_b) {
  logs[0] = static_cast<float>(acc[PACC_ENT] * inv_b);
  logs[1] = static_cast<float>(acc[PACC_PG] * inv_b);
  logs[2] = static_cast<float>(acc[PACC_KL] * inv_b);
  logs[3] = static_cast<float>(acc[PACC_CLIPFRAC] * inv_b);
  logs[4] = static_cast<float>(acc[PACC_VF] * inv_b);
  logs[5] = static_cast<float>(acc[PACC_VSUM] * inv_b);
  steps[0] += 1;
  steps[1] += 1;
  steps[3] += 1;
}

}  // namespace ppo
