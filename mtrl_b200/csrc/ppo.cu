// Multi-task PPO update on one device: MTPPO._update_inner (/root/reference/mtrl/rl/algorithms/mtppo.py:292-317) =
// update_policy (:196-254) then update_value_function (:256-290), one optimiser step each on the whole rollout.
// The two networks are independent given the rollout, so their trunk GEMMs share grouped launches
// ([policy, value] forward per layer; [dW_p, dW_v, dX_p, dX_v] backward per layer).
#include <vector>

#include "common.cuh"
#include "ln_kernels.cuh"
#include "mtrl_b200.h"
#include "net_common.cuh"
#include "ppo_kernels.cuh"

using namespace netc;

namespace {

struct PpoWs {
  float* X;
  float *P[MTRL_MAX_DEPTH], *V[MTRL_MAX_DEPTH];
  unsigned *bits_P[MTRL_MAX_DEPTH], *bits_V[MTRL_MAX_DEPTH];
  float* G[2][2];   // [net][ping-pong]
  float *logp, *adv, *ret, *val, *eps, *dout, *dq, *colsum_part;
  int *slot_src, *tile_task, *seg_start;
  double* acc;
  // MLP with LayerNorm / skip connections (ln_kernels.cuh; [0] policy, [1] value function): inputs n_j of Dense_j / of the
  // head (index j - 1), row statistics, and the backward's scratch (see the Workspace of sac.cu)
  float *N[2][MTRL_MAX_DEPTH], *St[2][MTRL_MAX_DEPTH];
  float *GN[2], *GS[2][2], *rowc[2], *part_dg[2], *part_dbeta[2];
};

long long ppo_carve(const mtrl_ppo_config_t& c, int K, int M, uint8_t* base, PpoWs* out) {
  long long off = 0;
  auto take = [&](long long bytes) -> uint8_t* {
    uint8_t* p = base ? base + off : nullptr;
    off = round_up(off + bytes, 256);
    return p;
  };
  auto f = [&](long long n) { return reinterpret_cast<float*>(take(n * 4 + 256)); };
  const long long W = c.width, A = c.action_dim;
  PpoWs w;
  memset(&w, 0, sizeof(w));
  w.X = f(static_cast<long long>(M) * K);
  const long long bit_words = static_cast<long long>(M) * ((W + 31) / 32);
  for (int l = 0; l < c.depth; ++l) {
    w.P[l] = f(M * W);
    w.V[l] = f(M * W);
    if (l + 1 < c.depth) {
      w.bits_P[l] = reinterpret_cast<unsigned*>(take(bit_words * 4));
      w.bits_V[l] = reinterpret_cast<unsigned*>(take(bit_words * 4));
    }
  }
  for (int n = 0; n < 2; ++n)
    for (int b = 0; b < 2; ++b) w.G[n][b] = f(M * W);
  w.logp = f(M); w.adv = f(M); w.ret = f(M); w.val = f(M);
  w.eps = f(M * A);
  w.dout = f(M * 2 * A);
  w.dq = f(M);
  w.colsum_part = f(2ll * (M / 32) * W);
  if (c.use_layer_norm || c.use_skip_connections) {
    for (int n = 0; n < 2; ++n) {
      for (int l = 0; l < c.depth; ++l) {
        w.N[n][l] = f(M * W);
        w.St[n][l] = f(static_cast<long long>(M) * 2);
      }
      w.GN[n] = f(M * W);
      w.GS[n][0] = f(M * W);
      w.GS[n][1] = f(M * W);
      w.rowc[n] = f(static_cast<long long>(M) * 2);
      w.part_dg[n] = f(static_cast<long long>(M / sac::kTileRows) * W);
      w.part_dbeta[n] = f(static_cast<long long>(M / sac::kTileRows) * W);
    }
  }
  w.slot_src = reinterpret_cast<int*>(take(static_cast<long long>(M) * 4));
  w.tile_task = reinterpret_cast<int*>(take(static_cast<long long>(M / sac::kTileRows) * 4));
  w.seg_start = reinterpret_cast<int*>(take((c.num_tasks + 1) * 4));
  w.acc = reinterpret_cast<double*>(take(ppo::PACC_COUNT * 8));
  if (out) *out = w;
  return off;
}

int ppo_validate(const mtrl_ppo_config_t& c) {
  MTRL_REQUIRE(c.num_tasks >= 1 && c.steps_per_task >= 1, "ppo config: num_tasks and steps_per_task must be positive");
  MTRL_REQUIRE(c.obs_dim >= 16 && c.obs_dim > (c.num_tasks > 1 ? c.num_tasks : 0), "ppo config: obs_dim %d too small", c.obs_dim);
  MTRL_REQUIRE(c.action_dim >= 1 && c.action_dim <= sac::kMaxA, "ppo config: action_dim %d outside [1, %d]", c.action_dim, sac::kMaxA);
  MTRL_REQUIRE(c.width >= 16 && c.width % 4 == 0, "ppo config: width %d must be a multiple of 4 and >= 16", c.width);
  MTRL_REQUIRE(c.depth >= 1 && c.depth <= MTRL_MAX_DEPTH, "ppo config: depth %d outside [1, %d]", c.depth, MTRL_MAX_DEPTH);
  MTRL_REQUIRE(static_cast<size_t>(c.width) * 2 * c.action_dim * 4 <= 200 * 1024, "ppo config: policy head does not fit shared memory");
  MTRL_REQUIRE(!(c.use_layer_norm || c.use_skip_connections) || c.num_tasks == 1,
               "ppo config: LayerNorm / skip connections belong to the plain MLP (num_tasks == 1)");
  MTRL_REQUIRE(!c.use_skip_connections || c.obs_dim != c.width, "ppo config: skip connections with obs_dim == width are not supported");
  return MTRL_OK;
}

}  // namespace

struct mtrl_ppo {
  mtrl_ppo_config_t cfg;
  mtrl_ppo_buffers_t buf;
  mtrl_ppo_layout_t lay;
  PpoWs ws;
  int sms = 148, steps_pad = 0, M = 0;
  std::vector<mtrl_gemm_plan_t*> fwd, bwd;
  int launches = 0;
  bool ln_mode = false;
};

extern "C" int mtrl_ppo_query_layout(const mtrl_ppo_config_t* cfg, mtrl_ppo_layout_t* out) {
  MTRL_REQUIRE(cfg && out, "mtrl_ppo_query_layout: null argument");
  MTRL_PROPAGATE(ppo_validate(*cfg));
  memset(out, 0, sizeof(*out));
  fill_net_layout(&out->policy, cfg->obs_dim, 2 * cfg->action_dim, 1, cfg->num_tasks, cfg->width, cfg->depth, cfg->use_layer_norm != 0);
  fill_net_layout(&out->vf, cfg->obs_dim, 1, 1, cfg->num_tasks, cfg->width, cfg->depth, cfg->use_layer_norm != 0);
  out->k_in = static_cast<int>(round_up(cfg->obs_dim, 32));
  const int steps_pad = static_cast<int>(round_up(cfg->steps_per_task, sac::kTileRows));
  out->max_rows = steps_pad * cfg->num_tasks;
  out->workspace_bytes = ppo_carve(*cfg, out->k_in, out->max_rows, nullptr, nullptr);
  return MTRL_OK;
}

extern "C" void mtrl_ppo_destroy(mtrl_ppo_t* h) {
  if (!h) return;
  for (auto* p : h->fwd) mtrl_gemm_plan_destroy(p);
  for (auto* p : h->bwd) mtrl_gemm_plan_destroy(p);
  delete h;
}

extern "C" int mtrl_ppo_refresh_shadows(mtrl_ppo_t* h, void* stream) {
  MTRL_REQUIRE(h, "mtrl_ppo_refresh_shadows: null handle");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  sac::shadow_kernel<<<h->sms * 4, 256, 0, st>>>(h->buf.policy_params, h->buf.policy_shadow, nullptr, h->lay.policy.total);
  sac::shadow_kernel<<<h->sms * 4, 256, 0, st>>>(h->buf.vf_params, h->buf.vf_shadow, nullptr, h->lay.vf.total);
  MTRL_CUDA_CHECK(cudaGetLastError());
  return MTRL_OK;
}

extern "C" int mtrl_ppo_create(mtrl_ppo_t** out, const mtrl_ppo_config_t* cfg, const mtrl_ppo_buffers_t* b) {
  MTRL_REQUIRE(out && cfg && b, "mtrl_ppo_create: null argument");
  mtrl_ppo* h = new mtrl_ppo();
  h->cfg = *cfg;
  h->buf = *b;
  h->ln_mode = cfg->use_layer_norm || cfg->use_skip_connections;
  int rc = mtrl_ppo_query_layout(cfg, &h->lay);
  if (rc != MTRL_OK) { delete h; return rc; }
  const void* need[] = {b->policy_params, b->policy_grads, b->policy_m, b->policy_v, b->policy_shadow, b->vf_params,
                        b->vf_grads, b->vf_m, b->vf_v, b->vf_shadow, b->steps, b->logs, b->workspace};
  for (const void* p : need) {
    if (!p || (reinterpret_cast<uintptr_t>(p) & 15u)) {
      delete h;
      mtrl_set_error("mtrl_ppo_create: every buffer must be non-null and 16-byte aligned");
      return MTRL_ERR_INVALID;
    }
  }
  h->M = h->lay.max_rows;
  h->steps_pad = h->M / cfg->num_tasks;
  ppo_carve(*cfg, h->lay.k_in, h->M, static_cast<uint8_t*>(b->workspace), &h->ws);
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, dev);
  cudaMemset(b->workspace, 0, h->lay.workspace_bytes);
  // static packing geometry: task t owns packed rows [t * steps_pad, (t + 1) * steps_pad)
  {
    std::vector<int> tt(h->M / sac::kTileRows), ss(cfg->num_tasks + 1);
    for (size_t i = 0; i < tt.size(); ++i) tt[i] = static_cast<int>(i * sac::kTileRows / h->steps_pad);
    for (int t = 0; t <= cfg->num_tasks; ++t) ss[t] = t * h->steps_pad;
    cudaMemcpy(h->ws.tile_task, tt.data(), tt.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(h->ws.seg_start, ss.data(), ss.size() * 4, cudaMemcpyHostToDevice);
  }
#define MTRL_PL_ATTR(A_) cudaFuncSetAttribute(ppo::policy_loss_kernel<A_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  MTRL_PL_ATTR(1) MTRL_PL_ATTR(2) MTRL_PL_ATTR(3) MTRL_PL_ATTR(4) MTRL_PL_ATTR(5) MTRL_PL_ATTR(6) MTRL_PL_ATTR(7) MTRL_PL_ATTR(8)
#undef MTRL_PL_ATTR
  const mtrl_net_layout_t& LP = h->lay.policy;
  const mtrl_net_layout_t& LV = h->lay.vf;
  const PpoWs& w = h->ws;
  const int M = h->M, W = cfg->width, D = cfg->depth, K = h->lay.k_in;
  auto cpart = [&](int n) { return w.colsum_part + static_cast<long long>(n) * (M / 32) * W; };
  // input of Dense_l (l >= 1): the previous activation, or the junction's output n_l with LayerNorm / skip connections
  const bool ln = h->ln_mode;
  auto in_p = [&](int l) { return ln ? w.N[0][l - 1] : w.P[l - 1]; };
  auto in_v = [&](int l) { return ln ? w.N[1][l - 1] : w.V[l - 1]; };
  for (int l = 0; l < D && rc == MTRL_OK; ++l) {
    std::vector<mtrl_gemm_problem_t> p;
    p.push_back(fwd_problem(l == 0 ? w.X : in_p(l), l == 0 ? K : W, l == 0 ? LP.in_dim : W, tk(b->policy_shadow, LP, 0, l),
                            tb(b->policy_params, LP, 0, l), w.P[l], M, W, (l + 1 < D && !ln) ? w.bits_P[l] : nullptr));
    p.push_back(fwd_problem(l == 0 ? w.X : in_v(l), l == 0 ? K : W, l == 0 ? LV.in_dim : W, tk(b->vf_shadow, LV, 0, l),
                            tb(b->vf_params, LV, 0, l), w.V[l], M, W, (l + 1 < D && !ln) ? w.bits_V[l] : nullptr));
    rc = make_plan(h->fwd, p);
  }
  for (int l = D - 1; l >= 0 && rc == MTRL_OK; --l) {
    const int src = ln ? 0 : (D - 1 - l) & 1, dst = src ^ 1;
    std::vector<mtrl_gemm_problem_t> p;
    p.push_back(dw_problem(l == 0 ? w.X : in_p(l), l == 0 ? K : W, l == 0 ? LP.in_dim : W, w.G[0][src],
                           tk(b->policy_grads, LP, 0, l), M, W, h->sms, 0));
    p.push_back(dw_problem(l == 0 ? w.X : in_v(l), l == 0 ? K : W, l == 0 ? LV.in_dim : W, w.G[1][src],
                           tk(b->vf_grads, LV, 0, l), M, W, h->sms, 0));
    if (l > 0 && !ln) {
      p.push_back(dx_problem(w.G[0][src], tk(b->policy_shadow, LP, 0, l), W, w.bits_P[l - 1], w.G[0][dst], M, W, cpart(0)));
      p.push_back(dx_problem(w.G[1][src], tk(b->vf_shadow, LV, 0, l), W, w.bits_V[l - 1], w.G[1][dst], M, W, cpart(1)));
    } else if (l > 0) {
      // the plain gradient of the layer's input n_l (fp32): the junction below applies the ReLU gate
      mtrl_gemm_problem_t q0 = dx_problem(w.G[0][0], tk(b->policy_shadow, LP, 0, l), W, nullptr, w.GN[0], M, W, nullptr);
      mtrl_gemm_problem_t q1 = dx_problem(w.G[1][0], tk(b->vf_shadow, LV, 0, l), W, nullptr, w.GN[1], M, W, nullptr);
      q0.epilogue = q1.epilogue = MTRL_EPI_STORE;
      p.push_back(q0);
      p.push_back(q1);
    }
    rc = make_plan(h->bwd, p);
  }
  if (rc == MTRL_OK) rc = mtrl_ppo_refresh_shadows(h, nullptr);
  if (rc != MTRL_OK) { mtrl_ppo_destroy(h); return rc; }
  cudaDeviceSynchronize();
  *out = h;
  return MTRL_OK;
}

// rollout arrays are device fp32, flattened task-major (row = task * steps_per_task + step): observations
// (B, obs_dim), log_probs / advantages / returns / values (B) (mtrl/types.py:48-63); eps (B, action_dim) or NULL.
extern "C" int mtrl_ppo_update(mtrl_ppo_t* h, const float* obs, const float* log_probs, const float* advantages,
                               const float* returns, const float* values, const float* eps, void* stream) {
  MTRL_REQUIRE(h && obs && log_probs && advantages && returns && values, "mtrl_ppo_update: null rollout pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const mtrl_ppo_config_t& c = h->cfg;
  PpoWs& w = h->ws;
  const mtrl_net_layout_t& LP = h->lay.policy;
  const mtrl_net_layout_t& LV = h->lay.vf;
  const int M = h->M, W = c.width, D = c.depth, T = c.num_tasks, A = c.action_dim;
  const long long B = static_cast<long long>(T) * c.steps_per_task;
  const float inv_b = 1.f / static_cast<float>(B);
  h->launches = 0;
  auto cpart = [&](int n) { return w.colsum_part + static_cast<long long>(n) * (M / 32) * W; };
  MTRL_CUDA_CHECK(cudaMemsetAsync(w.acc, 0, ppo::PACC_COUNT * sizeof(double), st));
  MTRL_CUDA_CHECK(cudaMemsetAsync(h->buf.policy_grads, 0, LP.total * sizeof(float), st));
  MTRL_CUDA_CHECK(cudaMemsetAsync(h->buf.vf_grads, 0, LV.total * sizeof(float), st));
  h->launches += 3;
  ppo::adv_stats_kernel<<<h->sms * 2, 256, 0, st>>>(advantages, B, w.acc);
  ppo::PackArgs pa;
  pa.obs = obs; pa.logp = log_probs; pa.adv = advantages; pa.ret = returns; pa.val = values; pa.eps = eps;
  pa.X = w.X; pa.plogp = w.logp; pa.padv = w.adv; pa.pret = w.ret; pa.pval = w.val; pa.peps = w.eps;
  pa.slot_src = w.slot_src; pa.noise_counter = h->buf.steps + 3; pa.seed = c.noise_seed;
  pa.obs_dim = c.obs_dim; pa.act_dim = A; pa.K = h->lay.k_in; pa.steps = c.steps_per_task; pa.steps_pad = h->steps_pad;
  ppo::pack_kernel<<<M, 128, 0, st>>>(pa);
  h->launches += 2;
  MTRL_CUDA_CHECK(cudaGetLastError());
  float* net_params[2] = {h->buf.policy_params, h->buf.vf_params};
  float* net_grads[2] = {h->buf.policy_grads, h->buf.vf_grads};
  const mtrl_net_layout_t* net_lay[2] = {&LP, &LV};
  float* const* act[2] = {w.P, w.V};
  const bool ln = c.use_layer_norm != 0, skip = c.use_skip_connections != 0;
  for (int l = 0; l < D; ++l) {
    MTRL_PROPAGATE(mtrl_gemm_plan_run(h->fwd[l], st));
    h->launches++;
    if (h->ln_mode) {   // junction l + 1 of both networks (ln_kernels.cuh)
      sac::LnFwdArgs ja;
      memset(&ja, 0, sizeof(ja));
      for (int n = 0; n < 2; ++n) {
        sac::LnFwdPass& p = ja.p[n];
        p.D = act[n][l];
        p.Nprev = (skip && l >= 1) ? w.N[n][l - 1] : nullptr;
        p.scale = ln ? lns(net_params[n], *net_lay[n], 0, l) : nullptr;
        p.bias = ln ? lnb(net_params[n], *net_lay[n], 0, l) : nullptr;
        p.N = w.N[n][l];
        p.stats = w.St[n][l];
      }
      ja.npass = 2; ja.M = M; ja.W = W; ja.lo_delta = 0; ja.eps = 1e-6f;
      sac::ln_fwd_kernel<<<dim3((M + 7) / 8, 2), 256, 0, st>>>(ja);
      h->launches++;
    }
  }
  const float* head_p = h->ln_mode ? w.N[0][D - 1] : w.P[D - 1];
  const float* head_v = h->ln_mode ? w.N[1][D - 1] : w.V[D - 1];
  {
    ppo::PolicyLossArgs a;
    a.H = head_p; a.Wh = hk(h->buf.policy_params, LP, 0); a.bh = hb(h->buf.policy_params, LP, 0);
    a.tile_task = w.tile_task; a.slot_src = w.slot_src; a.eps = w.eps; a.old_logp = w.logp; a.adv = w.adv;
    a.dout = w.dout; a.acc = w.acc; a.M = M; a.W = W;
    a.ls_min = c.log_std_min; a.ls_max = c.log_std_max; a.clip_eps = c.clip_eps; a.ent_coef = c.entropy_coefficient;
    a.inv_b = inv_b; a.n_rows = static_cast<float>(B); a.normalize = c.normalize_advantages;
    const size_t wbytes = static_cast<size_t>(W) * 2 * A * sizeof(float);
    dim3 grid(M / 32), block(256);
    switch (A) {
      case 1: ppo::policy_loss_kernel<1><<<grid, block, wbytes, st>>>(a); break;
      case 2: ppo::policy_loss_kernel<2><<<grid, block, wbytes, st>>>(a); break;
      case 3: ppo::policy_loss_kernel<3><<<grid, block, wbytes, st>>>(a); break;
      case 4: ppo::policy_loss_kernel<4><<<grid, block, wbytes, st>>>(a); break;
      case 5: ppo::policy_loss_kernel<5><<<grid, block, wbytes, st>>>(a); break;
      case 6: ppo::policy_loss_kernel<6><<<grid, block, wbytes, st>>>(a); break;
      case 7: ppo::policy_loss_kernel<7><<<grid, block, wbytes, st>>>(a); break;
      default: ppo::policy_loss_kernel<8><<<grid, block, wbytes, st>>>(a); break;
    }
    ppo::ValueLossArgs v;
    v.H = head_v; v.w = hk(h->buf.vf_params, LV, 0); v.b = hb(h->buf.vf_params, LV, 0);
    v.tile_task = w.tile_task; v.slot_src = w.slot_src; v.ret = w.ret; v.old_val = w.val; v.dq = w.dq; v.acc = w.acc;
    v.M = M; v.W = W; v.clip_eps = c.clip_eps; v.vf_coef = c.vf_coefficient; v.inv_b = inv_b; v.clip = c.clip_vf_loss;
    ppo::value_loss_kernel<<<(M + 7) / 8, 256, 0, st>>>(v);
    h->launches += 2;
    MTRL_CUDA_CHECK(cudaGetLastError());
  }
  {
    sac::HeadBwdArgs a;
    memset(&a, 0, sizeof(a));
    a.H[0] = head_p; a.dout[0] = w.dout; a.Wh[0] = hk(h->buf.policy_params, LP, 0); a.dZ[0] = h->ln_mode ? w.GN[0] : w.G[0][0];
    a.dWh[0] = hk(h->buf.policy_grads, LP, 0); a.dbh[0] = hb(h->buf.policy_grads, LP, 0); a.colsum[0] = h->ln_mode ? nullptr : cpart(0);
    a.seg_start = w.seg_start; a.M = M; a.W = W; a.no_mask = h->ln_mode;
    MTRL_REQUIRE(launch_head_bwd_any(a, 2 * A, T, 1, st), "ppo: unsupported head_dim %d", 2 * A);
    sac::HeadBwdArgs v;
    memset(&v, 0, sizeof(v));
    v.H[0] = head_v; v.dout[0] = w.dq; v.Wh[0] = hk(h->buf.vf_params, LV, 0); v.dZ[0] = h->ln_mode ? w.GN[1] : w.G[1][0];
    v.dWh[0] = hk(h->buf.vf_grads, LV, 0); v.dbh[0] = hb(h->buf.vf_grads, LV, 0); v.colsum[0] = h->ln_mode ? nullptr : cpart(1);
    v.seg_start = w.seg_start; v.M = M; v.W = W; v.no_mask = h->ln_mode;
    launch_head_bwd_any(v, 1, T, 1, st);
    h->launches += 2;
    MTRL_CUDA_CHECK(cudaGetLastError());
  }
  for (int l = D - 1, i = 0; l >= 0 && h->ln_mode; --l, ++i) {
    // junction l + 1 on the way back (see run_trunk_backward_ln in sac.cu), then the layer's dW / dX plan
    const int j = l + 1;
    sac::LnBwdArgs ja;
    memset(&ja, 0, sizeof(ja));
    sac::ColsumJobs jobs;
    memset(&jobs, 0, sizeof(jobs));
    for (int n = 0; n < 2; ++n) {
      sac::LnBwdPass& p = ja.p[n];
      p.dN = w.GN[n];
      p.dSkip = (skip && j >= 1 && j <= D - 1) ? w.GS[n][(j + 1) & 1] : nullptr;
      p.D = act[n][l];
      p.Nprev = (skip && l >= 1) ? w.N[n][l - 1] : nullptr;
      p.scale = ln ? lns(net_params[n], *net_lay[n], 0, l) : nullptr;
      p.stats = w.St[n][l];
      p.rowc = w.rowc[n];
      p.dZ = w.G[n][0];
      p.dX = (skip && l >= 1) ? w.GS[n][j & 1] : nullptr;
      p.part_db = cpart(n);
      p.part_dg = ln ? w.part_dg[n] : nullptr;
      p.part_dbeta = ln ? w.part_dbeta[n] : nullptr;
      jobs.part[jobs.njobs] = cpart(n);
      jobs.dst[jobs.njobs++] = tb(net_grads[n], *net_lay[n], 0, l);
      if (ln) {
        jobs.part[jobs.njobs] = w.part_dg[n];
        jobs.dst[jobs.njobs++] = lns(net_grads[n], *net_lay[n], 0, l);
        jobs.part[jobs.njobs] = w.part_dbeta[n];
        jobs.dst[jobs.njobs++] = lnb(net_grads[n], *net_lay[n], 0, l);
      }
    }
    ja.npass = 2; ja.M = M; ja.W = W; ja.lo_delta = 0;
    if (ln) sac::ln_bwd_rows_kernel<<<dim3((M + 7) / 8, 2), 256, 0, st>>>(ja);
    sac::ln_bwd_tile_kernel<<<dim3((W + 127) / 128, M / sac::kTileRows, 2), 256, 0, st>>>(ja);
    sac::colsum_final_kernel<<<dim3((W + 31) / 32, jobs.njobs), 256, 0, st>>>(jobs, M / sac::kTileRows, W);
    MTRL_PROPAGATE(mtrl_gemm_plan_run(h->bwd[i], st));
    h->launches += 4;
  }
  for (int l = D - 1, i = 0; l >= 0 && !h->ln_mode; --l, ++i) {
    sac::ColsumJobs jobs;
    memset(&jobs, 0, sizeof(jobs));
    jobs.njobs = 2;
    jobs.part[0] = cpart(0); jobs.dst[0] = tb(h->buf.policy_grads, LP, 0, l);
    jobs.part[1] = cpart(1); jobs.dst[1] = tb(h->buf.vf_grads, LV, 0, l);
    const int groups = l == D - 1 ? M / sac::kTileRows : M / 32;
    sac::colsum_final_kernel<<<dim3((W + 31) / 32, 2), 256, 0, st>>>(jobs, groups, W);
    MTRL_PROPAGATE(mtrl_gemm_plan_run(h->bwd[i], st));
    h->launches += 2;
  }
  // clip_by_global_norm + Adam per network (mtppo.py:249-252, 285-288); no target networks
  struct Net { float *p, *g, *m, *v, *sh; const mtrl_net_layout_t* L; int g2, hg2, step; float lr, mx; };
  Net nets[2] = {{h->buf.policy_params, h->buf.policy_grads, h->buf.policy_m, h->buf.policy_v, h->buf.policy_shadow, &LP,
                  ppo::PACC_P_G2, ppo::PACC_P_HEAD_G2, 0, c.policy_lr, c.policy_max_grad_norm},
                 {h->buf.vf_params, h->buf.vf_grads, h->buf.vf_m, h->buf.vf_v, h->buf.vf_shadow, &LV, ppo::PACC_V_G2,
                  ppo::PACC_V_HEAD_G2, 1, c.vf_lr, c.vf_max_grad_norm}};
  for (const Net& n : nets) {
    sac::sumsq_kernel<<<64, 256, 0, st>>>(n.g + n.L->heads_base, n.L->total - n.L->heads_base, w.acc + n.hg2);
    sac::write_slot_kernel<<<1, 1, 0, st>>>(n.g + n.L->slots_off, w.acc + n.hg2);
    sac::sumsq_kernel<<<h->sms * 2, 256, 0, st>>>(n.g, n.L->trunk_total, w.acc + n.g2);
    sac::AdamArgs a;
    memset(&a, 0, sizeof(a));
    a.p = n.p; a.m = n.m; a.v = n.v; a.shadow = n.sh; a.g = n.g; a.target = nullptr; a.target_shadow = nullptr;
    a.n = n.L->total; a.trunk_n = n.L->trunk_total;
    a.g2_trunk = w.acc + n.g2; a.g2_heads = n.g + n.L->slots_off; a.step = h->buf.steps + n.step;
    a.p2_trunk = w.acc + ppo::PACC_SCRATCH0; a.p2_head = w.acc + ppo::PACC_SCRATCH1; a.p2_old = w.acc + ppo::PACC_SCRATCH2;
    a.lr = n.lr; a.b1 = c.adam_b1; a.b2 = c.adam_b2; a.eps = c.adam_eps; a.max_norm = n.mx; a.tau = 0.f;
    sac::adam_kernel<<<h->sms * 4, 256, 0, st>>>(a);
    h->launches += 4;
  }
  ppo::finalize_kernel<<<1, 1, 0, st>>>(w.acc, h->buf.steps, h->buf.logs, inv_b);
  h->launches += 1;
  MTRL_CUDA_CHECK(cudaGetLastError());
  return MTRL_OK;
}

extern "C" int mtrl_ppo_launches_per_update(const mtrl_ppo_t* h) { return h ? h->launches : 0; }
