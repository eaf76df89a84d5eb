// LayerNorm and residual junctions of the reference's MLP (mtrl/nn/base.py:32-63):
//
//   for i in range(depth):
//       n_i = LayerNorm_{i-1}(x_i) if use_layer_norm and i != 0 else x_i            (:35-37)
//       d_i = relu(Dense_i(n_i))                                                     (:38-45)
//       x_{i+1} = n_i + d_i if use_skip_connections and n_i has the hidden width else d_i   (:46-49)
//   n_depth = LayerNorm_{depth-1}(x_depth) if use_layer_norm else x_depth            (:52-53)
//   out = Dense_depth(n_depth)
//
// flax.linen.LayerNorm defaults: last axis, epsilon 1e-6, learnable scale + bias, use_fast_variance (var = E[x^2] - E[x]^2).
// The Dense layers stay tcgen05 GEMMs whose epilogue (bias + ReLU) writes d_i; a "junction" j = 1..depth turns
// (d_{j-1}, n_{j-1}) into n_j -- the next GEMM's A operand -- in one HBM pass, and on the way back turns the gradient of
// n_j into dZ_{j-1} = g(x_j) * 1[d_{j-1} > 0] (the dW / dX GEMMs' operand), the bias / scale gradients' column partials
// and the gradient that travels on through the skip connection.  HBM-bound elementwise + row-reduction work:
// forward reads and writes M x W once (the second read of a row hits L1 / L2); backward is two passes (row coefficients,
// then 128 x 128 tiles with column partials).
#pragma once

#include "common.cuh"
#include "sac_kernels.cuh"

namespace sac {

constexpr int kMaxLnPasses = 6;   // actor(next), actor(obs), 4 critic members at most in one launch

struct LnFwdPass {
  const float* D;       // [M][W] relu output d_{j-1} of the previous Dense (tf32 hi; remainder at + lo_delta in fp32x3 mode)
  const float* Nprev;   // [M][W] n_{j-1}, added when layer j-1 has a skip connection; else null
  const float* scale;   // [W] LayerNorm_{j-1}/scale, or null: no LayerNorm (n_j = x_j)
  const float* bias;    // [W]
  float* N;             // [M][W] out: n_j as a GEMM operand (hi; remainder at + lo_delta)
  float* stats;         // [M][2] out: mean, rstd of x_j (read by the backward)
};
struct LnFwdArgs {
  LnFwdPass p[kMaxLnPasses];
  int npass, M, W;
  long long lo_delta;
  float eps;
};

__device__ __forceinline__ float4 ld_pair(const float* p, long long lo_delta) {
  float4 a = *reinterpret_cast<const float4*>(p);
  if (lo_delta) {
    const float4 l = *reinterpret_cast<const float4*>(p + lo_delta);
    a.x += l.x; a.y += l.y; a.z += l.z; a.w += l.w;
  }
  return a;
}
__device__ __forceinline__ void st_pair(float* p, long long lo_delta, const float4& v) {
  const float4 h = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
  *reinterpret_cast<float4*>(p) = h;
  if (lo_delta)
    *reinterpret_cast<float4*>(p + lo_delta) = make_float4(tf32_lo(v.x, h.x), tf32_lo(v.y, h.y), tf32_lo(v.z, h.z), tf32_lo(v.w, h.w));
}

// One warp per row; grid (ceil(M / 8), npass), 256 threads.  W % 4 == 0.
static __global__ void ln_fwd_kernel(const LnFwdArgs a) {
  MTRL_PDL_PROLOGUE();
  const LnFwdPass& p = a.p[blockIdx.y];
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= a.M) return;
  const long long ro = static_cast<long long>(row) * a.W;
  float mean = 0.f, rstd = 1.f;
  if (p.scale) {
    float s = 0.f, s2 = 0.f;
    for (int k = lane * 4; k < a.W; k += 128) {
      float4 u = ld_pair(p.D + ro + k, a.lo_delta);
      if (p.Nprev) {
        const float4 n = ld_pair(p.Nprev + ro + k, a.lo_delta);
        u.x += n.x; u.y += n.y; u.z += n.z; u.w += n.w;
      }
      s += (u.x + u.y) + (u.z + u.w);
      s2 += (u.x * u.x + u.y * u.y) + (u.z * u.z + u.w * u.w);
    }
    s = warp_sum(s);
    s2 = warp_sum(s2);
    mean = s / static_cast<float>(a.W);
    const float var = fmaxf(s2 / static_cast<float>(a.W) - mean * mean, 0.f);
    rstd = rsqrtf(var + a.eps);
  }
  for (int k = lane * 4; k < a.W; k += 128) {
    float4 u = ld_pair(p.D + ro + k, a.lo_delta);
    if (p.Nprev) {
      const float4 n = ld_pair(p.Nprev + ro + k, a.lo_delta);
      u.x += n.x; u.y += n.y; u.z += n.z; u.w += n.w;
    }
    if (p.scale) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(p.scale + k));
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + k));
      u.x = (u.x - mean) * rstd * g.x + b.x;
      u.y = (u.y - mean) * rstd * g.y + b.y;
      u.z = (u.z - mean) * rstd * g.z + b.z;
      u.w = (u.w - mean) * rstd * g.w + b.w;
    }
    st_pair(p.N + ro + k, a.lo_delta, u);
  }
  if (lane == 0) {
    p.stats[2 * row] = mean;
    p.stats[2 * row + 1] = rstd;
  }
}

// Backward of junction j for one pass.  g(n_j) = dN (+ dSkip); with x^ = (x_j - mean) rstd and g = g(n_j) gamma:
//   g(x_j) = rstd (g - mean_k(g) - x^ mean_k(g x^));   dgamma = sum_rows g(n_j) x^;   dbeta = sum_rows g(n_j)
// (no LayerNorm: g(x_j) = g(n_j)).  Then dZ_{j-1} = g(x_j) 1[d_{j-1} > 0].
struct LnBwdPass {
  const float* dN;        // [M][W] fp32: dZ_j W_j^T from the dX GEMM, or the head VJP for j = depth
  const float* dSkip;     // [M][W] fp32 or null: g(x_{j+1}), which reaches n_j directly when layer j has a skip connection
  const float* D;         // [M][W] d_{j-1} (hi; + lo_delta)
  const float* Nprev;     // [M][W] n_{j-1} or null (as in the forward)
  const float* scale;     // gamma or null
  const float* stats;     // [M][2] mean, rstd
  float* rowc;            // [M][2] scratch: mean_k(g), mean_k(g x^)
  float* dZ;              // [M][W] out: GEMM operand (hi; + lo_delta)
  float* dX;              // [M][W] out or null: g(x_j) for the junction below (when layer j-1 has a skip connection)
  float* part_db;         // [M / 128][W] or null: column sums of dZ (bias gradient of Dense_{j-1})
  float* part_dg;         // [M / 128][W] or null: column sums of g(n_j) x^ (LayerNorm_{j-1}/scale gradient)
  float* part_dbeta;      // [M / 128][W] or null: column sums of g(n_j)
};
struct LnBwdArgs {
  LnBwdPass p[kMaxE];
  int npass, M, W;
  long long lo_delta;
};

// Row coefficients; one warp per row, grid (ceil(M / 8), npass).
static __global__ void ln_bwd_rows_kernel(const LnBwdArgs a) {
  MTRL_PDL_PROLOGUE();
  const LnBwdPass& p = a.p[blockIdx.y];
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= a.M || !p.scale) return;
  const long long ro = static_cast<long long>(row) * a.W;
  const float mean = p.stats[2 * row], rstd = p.stats[2 * row + 1];
  float c1 = 0.f, c2 = 0.f;
  for (int k = lane * 4; k < a.W; k += 128) {
    float4 dn = *reinterpret_cast<const float4*>(p.dN + ro + k);
    if (p.dSkip) {
      const float4 s = *reinterpret_cast<const float4*>(p.dSkip + ro + k);
      dn.x += s.x; dn.y += s.y; dn.z += s.z; dn.w += s.w;
    }
    float4 u = ld_pair(p.D + ro + k, a.lo_delta);
    if (p.Nprev) {
      const float4 n = ld_pair(p.Nprev + ro + k, a.lo_delta);
      u.x += n.x; u.y += n.y; u.z += n.z; u.w += n.w;
    }
    const float4 g = __ldg(reinterpret_cast<const float4*>(p.scale + k));
    const float gx = dn.x * g.x, gy = dn.y * g.y, gz = dn.z * g.z, gw = dn.w * g.w;
    c1 += (gx + gy) + (gz + gw);
    c2 += (gx * (u.x - mean) + gy * (u.y - mean)) + (gz * (u.z - mean) + gw * (u.w - mean));
  }
  c1 = warp_sum(c1);
  c2 = warp_sum(c2) * rstd;
  if (lane == 0) {
    p.rowc[2 * row] = c1 / static_cast<float>(a.W);
    p.rowc[2 * row + 1] = c2 / static_cast<float>(a.W);
  }
}

// 128 rows x 128 columns per block; grid (ceil(W / 128), M / 128, npass), 256 threads: a lane owns 4 consecutive
// columns, warp w rows [16 w, 16 w + 16) of the tile.  Column partials leave through shared memory in a fixed order.
static __global__ void __launch_bounds__(256) ln_bwd_tile_kernel(const LnBwdArgs a) {
  MTRL_PDL_PROLOGUE();
  __shared__ __align__(16) float red[3][8][128];
  const LnBwdPass& p = a.p[blockIdx.z];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = blockIdx.x * 128 + lane * 4;
  const bool kok = k < a.W;
  const int row0 = blockIdx.y * kTileRows + warp * 16;
  float4 g4 = make_float4(1.f, 1.f, 1.f, 1.f);
  if (p.scale && kok) g4 = __ldg(reinterpret_cast<const float4*>(p.scale + k));
  float sdb[4] = {0.f, 0.f, 0.f, 0.f}, sdg[4] = {0.f, 0.f, 0.f, 0.f}, sdbeta[4] = {0.f, 0.f, 0.f, 0.f};
  if (kok) {
#pragma unroll 4
    for (int r = 0; r < 16; ++r) {
      const int row = row0 + r;
      if (row >= a.M) break;
      const long long ro = static_cast<long long>(row) * a.W + k;
      float4 dn = *reinterpret_cast<const float4*>(p.dN + ro);
      if (p.dSkip) {
        const float4 s = *reinterpret_cast<const float4*>(p.dSkip + ro);
        dn.x += s.x; dn.y += s.y; dn.z += s.z; dn.w += s.w;
      }
      const float4 d = ld_pair(p.D + ro, a.lo_delta);
      float4 du = dn;
      if (p.scale) {
        float4 u = d;
        if (p.Nprev) {
          const float4 n = ld_pair(p.Nprev + ro, a.lo_delta);
          u.x += n.x; u.y += n.y; u.z += n.z; u.w += n.w;
        }
        const float mean = p.stats[2 * row], rstd = p.stats[2 * row + 1];
        const float c1 = p.rowc[2 * row], c2 = p.rowc[2 * row + 1];
        const float xh[4] = {(u.x - mean) * rstd, (u.y - mean) * rstd, (u.z - mean) * rstd, (u.w - mean) * rstd};
        const float dnv[4] = {dn.x, dn.y, dn.z, dn.w};
        const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
        float o[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          o[c] = rstd * (dnv[c] * gv[c] - c1 - xh[c] * c2);
          sdg[c] += dnv[c] * xh[c];
          sdbeta[c] += dnv[c];
        }
        du = make_float4(o[0], o[1], o[2], o[3]);
      }
      if (p.dX) *reinterpret_cast<float4*>(p.dX + ro) = du;
      const float4 dz = make_float4(d.x > 0.f ? du.x : 0.f, d.y > 0.f ? du.y : 0.f, d.z > 0.f ? du.z : 0.f, d.w > 0.f ? du.w : 0.f);
      st_pair(p.dZ + ro, a.lo_delta, dz);
      // bias-gradient partials: of the unrounded values in fp32x3 mode, of what the GEMMs will read in tf32 mode
      sdb[0] += a.lo_delta ? dz.x : tf32_rna(dz.x);
      sdb[1] += a.lo_delta ? dz.y : tf32_rna(dz.y);
      sdb[2] += a.lo_delta ? dz.z : tf32_rna(dz.z);
      sdb[3] += a.lo_delta ? dz.w : tf32_rna(dz.w);
    }
  }
  *reinterpret_cast<float4*>(&red[0][warp][lane * 4]) = make_float4(sdb[0], sdb[1], sdb[2], sdb[3]);
  *reinterpret_cast<float4*>(&red[1][warp][lane * 4]) = make_float4(sdg[0], sdg[1], sdg[2], sdg[3]);
  *reinterpret_cast<float4*>(&red[2][warp][lane * 4]) = make_float4(sdbeta[0], sdbeta[1], sdbeta[2], sdbeta[3]);
  __syncthreads();
  const int kk = blockIdx.x * 128 + (threadIdx.x & 127);
  const int which = threadIdx.x >> 7;   // threads 0..127: db (+ dbeta), 128..255: dgamma
  if (kk < a.W) {
    const long long o = static_cast<long long>(blockIdx.y) * a.W + kk;
    const int c = threadIdx.x & 127;
    if (which == 0) {
      float s = 0.f, t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) { s += red[0][w][c]; t += red[2][w][c]; }
      if (p.part_db) p.part_db[o] = s;
      if (p.part_dbeta) p.part_dbeta[o] = t;
    } else if (p.part_dg) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[1][w][c];
      p.part_dg[o] = s;
    }
  }
}

}  // namespace sac
