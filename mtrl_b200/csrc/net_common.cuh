// Host-side helpers shared by the update orchestrators (sac.cu, ppo.cu): flat parameter layout, GEMM problem
// builders for the three trunk contractions, head-VJP launcher.
#pragma once

#include <vector>

#include "common.cuh"
#include "mtrl_b200.h"
#include "sac_kernels.cuh"

namespace netc {

using namespace sac;

inline long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }


inline void fill_net_layout(mtrl_net_layout_t* L, int in_dim, int head_dim, int members, int t_local, int width, int depth,
                            bool use_layer_norm = false) {
  memset(L, 0, sizeof(*L));
  L->in_dim = in_dim;
  L->head_dim = head_dim;
  L->members = members;
  L->num_local_tasks = t_local;
  L->width = width;
  L->depth = depth;
  long long off = 0;
  int d = in_dim;
  for (int i = 0; i < depth; ++i) {
    L->kernel_off[i] = off;
    off = round_up(off + static_cast<long long>(d) * width, 32);
    L->bias_off[i] = off;
    off = round_up(off + width, 32);
    if (use_layer_norm) {
      L->ln_scale_off[i] = off;
      off = round_up(off + width, 32);
      L->ln_bias_off[i] = off;
      off = round_up(off + width, 32);
    }
    d = width;
  }
  L->use_layer_norm = use_layer_norm ? 1 : 0;
  L->member_trunk_stride = off;
  L->trunk_total = off * members;
  L->slots_off = L->trunk_total;
  L->heads_base = L->trunk_total + 32;
  long long h = 0;
  L->head_kernel_off = h;
  h = round_up(h + static_cast<long long>(t_local) * width * head_dim, 32);
  L->head_bias_off = h;
  h = round_up(h + static_cast<long long>(t_local) * head_dim, 32);
  L->member_head_stride = h;
  L->total = L->heads_base + h * members;
}

inline int block_n_for(int n) {
  if (n <= 16) return 16;
  const int tiles = (n + 255) / 256;
  const int bn = static_cast<int>(round_up((n + tiles - 1) / tiles, 32));  // 32-column epilogue boxes / ReLU bit words
  return bn > 256 ? 256 : bn;
}


inline float* tk(float* base, const mtrl_net_layout_t& L, int e, int l) { return base + e * L.member_trunk_stride + L.kernel_off[l]; }
inline float* tb(float* base, const mtrl_net_layout_t& L, int e, int l) { return base + e * L.member_trunk_stride + L.bias_off[l]; }
inline float* lns(float* base, const mtrl_net_layout_t& L, int e, int k) { return base + e * L.member_trunk_stride + L.ln_scale_off[k]; }
inline float* lnb(float* base, const mtrl_net_layout_t& L, int e, int k) { return base + e * L.member_trunk_stride + L.ln_bias_off[k]; }
inline float* hk(float* base, const mtrl_net_layout_t& L, int e) { return base + L.heads_base + e * L.member_head_stride + L.head_kernel_off; }
inline float* hb(float* base, const mtrl_net_layout_t& L, int e) { return base + L.heads_base + e * L.member_head_stride + L.head_bias_off; }

// The trailing *_lo arguments are the tf32 remainders of the operands / the output (fp32x3 mode, mtrl_b200.h); all
// null = plain tf32.
inline mtrl_gemm_problem_t fwd_problem(const float* X, int ldx, int K, const float* Wsh, const float* bias, float* out, int M, int W,
                                unsigned* bits_out = nullptr, const float* X_lo = nullptr, const float* Wsh_lo = nullptr,
                                float* out_lo = nullptr) {
  mtrl_gemm_problem_t p;
  memset(&p, 0, sizeof(p));
  p.A = X; p.lda = ldx; p.a_major = 0;
  p.B = Wsh; p.ldb = W; p.b_major = 1;     // Flax kernel (in, out): N contiguous
  p.D = out; p.ldd = W;
  p.M = M; p.N = W; p.K = K;
  p.block_n = block_n_for(W); p.k_splits = 1; p.epilogue = MTRL_EPI_BIAS_RELU; p.bias = bias;
  p.relu_bits_out = bits_out; p.ldbits = (W + 31) / 32;
  p.A_lo = X_lo; p.B_lo = Wsh_lo; p.D_lo = out_lo;
  return p;
}
// dZ_prev = (dZ W^T) * (H_prev > 0): A = dZ [M][W] K-major, B = W [in=N][W=K] K-major
inline mtrl_gemm_problem_t dx_problem(const float* dZ, const float* Wsh, int n_in, const unsigned* mask_bits, float* out, int M, int W,
                               float* colsum_partial, const float* dZ_lo = nullptr, const float* Wsh_lo = nullptr,
                               float* out_lo = nullptr) {
  mtrl_gemm_problem_t p;
  memset(&p, 0, sizeof(p));
  p.A = dZ; p.lda = W; p.a_major = 0;
  p.B = Wsh; p.ldb = W; p.b_major = 0;
  p.D = out; p.ldd = n_in;
  p.M = M; p.N = n_in; p.K = W;
  p.block_n = block_n_for(n_in); p.k_splits = 1; p.epilogue = MTRL_EPI_RELU_MASK;
  p.mask_bits = mask_bits; p.ldbits = (n_in + 31) / 32;
  p.colsum_partial = colsum_partial;
  p.A_lo = dZ_lo; p.B_lo = Wsh_lo; p.D_lo = out_lo;
  return p;
}
// dW = X^T dZ: A = X [rows][in] MN-major, B = dZ [rows][W] MN-major, K = rows
inline mtrl_gemm_problem_t dw_problem(const float* X, int ldx, int n_in, const float* dZ, float* dW, int M, int W, int sms,
                               int units_hint, const float* X_lo = nullptr, const float* dZ_lo = nullptr) {
  mtrl_gemm_problem_t p;
  memset(&p, 0, sizeof(p));
  p.A = X; p.lda = ldx; p.a_major = 1;
  p.B = dZ; p.ldb = W; p.b_major = 1;
  p.D = dW; p.ldd = W;
  p.M = n_in; p.N = W; p.K = M;
  p.block_n = block_n_for(W);
  const int kb = (M + 31) / 32;
  const int tiles = ((n_in + 127) / 128) * ((W + p.block_n - 1) / p.block_n);
  int splits;
  if (n_in >= 128) {
    // match the K depth of the dX units that share the launch (W/32 k-blocks) so units are uniform
    splits = (kb + (W / 32) - 1) / (W / 32 > 0 ? W / 32 : 1);
  } else {
    splits = sms / (tiles > 0 ? tiles : 1);
  }
  const int max_splits = kb / 4 > 0 ? kb / 4 : 1;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  // enough output tiles to occupy every SM without cutting K (W >= 2048): one unit per tile, plain stores -- no reduce-adds, and
  // the gradient is bit-reproducible run to run; the longest-first schedule absorbs the longer units (same-box A/B at
  // MT50/W2048: GEMM launches 1.5 % shorter).  MTRL_DW_SPLITS=n forces n.
  if (n_in >= 128 && tiles * 2 >= sms) splits = 1;
  if (n_in >= 128 && getenv("MTRL_DW_SPLITS")) splits = atoi(getenv("MTRL_DW_SPLITS")) > 0 ? atoi(getenv("MTRL_DW_SPLITS")) : splits;
  (void)units_hint;
  p.k_splits = splits;
  p.epilogue = splits > 1 ? MTRL_EPI_ATOMIC_ADD : MTRL_EPI_STORE;
  p.A_lo = X_lo; p.B_lo = dZ_lo;
  return p;
}

inline int make_plan_plain(std::vector<mtrl_gemm_plan_t*>& dst, const std::vector<mtrl_gemm_problem_t>& probs, int flags = 0) {
  mtrl_gemm_plan_t* plan = nullptr;
  MTRL_PROPAGATE(mtrl_gemm_plan_create_ex(&plan, probs.data(), static_cast<int>(probs.size()), flags));
  dst.push_back(plan);
  return MTRL_OK;
}

// Average duration (ms) of running the plans one after the other on the default stream: min over `reps` timed rounds after
// one warm-up round.  Used at handle creation to choose between equivalent launch structures.
inline int time_plans(const std::vector<mtrl_gemm_plan_t*>& plans, float* ms_out, int reps = 4) {
  cudaEvent_t e0, e1;
  MTRL_CUDA_CHECK(cudaEventCreate(&e0));
  MTRL_CUDA_CHECK(cudaEventCreate(&e1));
  float best = 1e30f;
  int rc = MTRL_OK;
  for (int rep = 0; rep <= reps && rc == MTRL_OK; ++rep) {
    cudaEventRecord(e0, nullptr);
    for (mtrl_gemm_plan_t* p : plans)
      if (rc == MTRL_OK) rc = mtrl_gemm_plan_run(p, nullptr);
    cudaEventRecord(e1, nullptr);
    if (cudaEventSynchronize(e1) != cudaSuccess) { mtrl_set_error("GEMM timing launch failed"); rc = MTRL_ERR_CUDA; }
    float t = 0.f;
    cudaEventElapsedTime(&t, e0, e1);
    if (rep > 0 && t < best) best = t;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_out = best;
  return rc;
}

// Tile-shape selection by measurement.  A launch is a static schedule of units over the SMs, so its duration is (rounds of
// units) x (unit time): 80 units of 256 x 256 on 74 CTA pairs take two rounds where 120 units of width 192 take two
// SHORTER rounds, and the 112 units of 128 x 256 that one rank's 896 rows of a task-sharded batch give fit ONE round of
// 148 single CTAs where the 64 padded 256-row units need a full round of pairs at twice the unit time.  Which shape wins
// depends on the shapes grouped into the launch (batch rows, width, members) and on how far smaller tiles are L2-bound, so
// every candidate (CTA pairs or single CTAs) x (tile width) is built and timed once at plan creation (a few launches on the
// real buffers: forward outputs are overwritten and gradient accumulators re-zeroed by every update) and the fastest plan
// is kept.  MTRL_GEMM_AUTOTUNE=0 keeps the default shape; MTRL_GEMM_CTAS=1|2 restricts the candidates to one kind.
// `flags`: MTRL_GEMM_ROWDEPS for a chain of dependent layers in one launch (every candidate is built with it).
inline int make_plan(std::vector<mtrl_gemm_plan_t*>& dst, const std::vector<mtrl_gemm_problem_t>& probs, int flags = 0) {
  const char* env = getenv("MTRL_GEMM_AUTOTUNE");
  bool wide = false;
  for (const auto& p : probs) wide = wide || p.N >= 256;
  if ((env && env[0] == '0') || !wide) return make_plan_plain(dst, probs, flags);
  {
    // many rounds of units: the last partial round costs little and every candidate launch would take milliseconds
    mtrl_gemm_plan_t* probe = nullptr;
    MTRL_PROPAGATE(mtrl_gemm_plan_create_ex(&probe, probs.data(), static_cast<int>(probs.size()), flags));
    if (mtrl_gemm_plan_units(probe) >= 6 * 74) {
      dst.push_back(probe);
      return MTRL_OK;
    }
    mtrl_gemm_plan_destroy(probe);
  }
  cudaEvent_t e0, e1;
  MTRL_CUDA_CHECK(cudaEventCreate(&e0));
  MTRL_CUDA_CHECK(cudaEventCreate(&e1));
  mtrl_gemm_plan_t* best = nullptr;
  float best_ms = 0.f;
  int rc = MTRL_OK;
  const char* force = getenv("MTRL_GEMM_CTAS");
  // MTRL_GEMM_STREAMK=1 adds the stream-K schedule (default tile width) of either kind to the candidates.  Off by default:
  // at every shape probed (profiles/r02_streamk_probe_*.txt) its fix-up -- the last unit of a cut tile reads the parked
  // partial accumulators before the fused epilogue, on the launch's critical path -- costs more than the idle SMs it fills.
  const bool try_streamk = getenv("MTRL_GEMM_STREAMK") && getenv("MTRL_GEMM_STREAMK")[0] == '1';
  const int kinds[4] = {2, 1, 2 | MTRL_GEMM_STREAMK, 1 | MTRL_GEMM_STREAMK};
  const int cands[4] = {0, 192, 128, 64};   // 0 = the caller's default
  for (int kind_flags : kinds) {
    const int kind = kind_flags & ~MTRL_GEMM_STREAMK;
    if ((kind_flags & MTRL_GEMM_STREAMK) && (!try_streamk || flags)) continue;
    if (force && (force[0] == '1' || force[0] == '2') && force[0] - '0' != kind) continue;
    for (int cand : cands) {
      if ((kind_flags & MTRL_GEMM_STREAMK) && cand != 0) continue;
      std::vector<mtrl_gemm_problem_t> q = probs;
      bool changed = cand == 0;
      for (auto& p : q)
        if (cand && p.N >= 256 && p.block_n > cand) { p.block_n = cand; changed = true; }
      if (!changed) continue;
      mtrl_gemm_plan_t* plan = nullptr;
      rc = mtrl_gemm_plan_create_ex(&plan, q.data(), static_cast<int>(q.size()), kind_flags | flags);
      if (rc != MTRL_OK) break;
      if (mtrl_gemm_plan_ctas(plan) != kind) {   // an odd SM count turns pairs into single CTAs: already covered
        mtrl_gemm_plan_destroy(plan);
        continue;
      }
      float ms = 1e30f;
      for (int rep = 0; rep < 4 && rc == MTRL_OK; ++rep) {
        cudaEventRecord(e0, nullptr);
        rc = mtrl_gemm_plan_run(plan, nullptr);
        cudaEventRecord(e1, nullptr);
        if (cudaEventSynchronize(e1) != cudaSuccess) { mtrl_set_error("GEMM autotune launch failed"); rc = MTRL_ERR_CUDA; }
        float t = 0.f;
        cudaEventElapsedTime(&t, e0, e1);
        if (rep > 0 && t < ms) ms = t;   // first run warms the instruction cache / L2
      }
      if (rc != MTRL_OK) { mtrl_gemm_plan_destroy(plan); break; }
      if (!best || ms < best_ms * 0.97f) {   // prefer the earlier candidate unless a later one is clearly faster
        if (best) mtrl_gemm_plan_destroy(best);
        best = plan;
        best_ms = ms;
      } else {
        mtrl_gemm_plan_destroy(plan);
      }
    }
    if (rc != MTRL_OK) break;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (rc != MTRL_OK) {
    if (best) mtrl_gemm_plan_destroy(best);
    return rc;
  }
  if (getenv("MTRL_GEMM_AUTOTUNE_LOG"))
    fprintf(stderr, "[mtrl gemm autotune] %zu problems (M %d N %d K %d ...): ctas %d, %d units, %.1f us\n", probs.size(), probs[0].M,
            probs[0].N, probs[0].K, mtrl_gemm_plan_ctas(best), mtrl_gemm_plan_units(best), best_ms * 1e3f);
  dst.push_back(best);
  return MTRL_OK;
}

template <int HD>
inline void launch_head_bwd_t(const HeadBwdArgs& a, int T_local, int E, cudaStream_t st) {
  dim3 grid((a.W + 127) / 128, T_local, E);
  constexpr int RG = HD <= 4 ? 16 : 8;   // wide heads keep (HD x 4) weights + sums per lane: fewer, fatter warps
  mtrl_launch(head_bwd_kernel<HD, RG>, grid, dim3(RG * 32), 0, st, a);
}


// Head VJP launch for head_dim hd (critic 1, actor 2A); returns false for an unsupported head_dim.
inline bool launch_head_bwd_any(const HeadBwdArgs& a, int hd, int T_local, int E, cudaStream_t st) {
  switch (hd) {
    case 1: launch_head_bwd_t<1>(a, T_local, E, st); break;
    case 2: launch_head_bwd_t<2>(a, T_local, E, st); break;
    case 4: launch_head_bwd_t<4>(a, T_local, E, st); break;
    case 6: launch_head_bwd_t<6>(a, T_local, E, st); break;
    case 8: launch_head_bwd_t<8>(a, T_local, E, st); break;
    case 10: launch_head_bwd_t<10>(a, T_local, E, st); break;
    case 12: launch_head_bwd_t<12>(a, T_local, E, st); break;
    case 14: launch_head_bwd_t<14>(a, T_local, E, st); break;
    case 16: launch_head_bwd_t<16>(a, T_local, E, st); break;
    default: return false;
  }
  return true;
}

}  // namespace netc
