// MT-SAC update on one device: handle, buffer carve-up, GEMM plans and the launch sequence.
//
// Sequence = MTSAC._update_inner (/root/reference/mtrl/rl/algorithms/mtsac.py:1173-1247), non-split branch:
//   critic step (:513-621)  uses OLD actor and OLD alpha; target update uses the NEW critic (:607-613)
//   actor  step (:623-711)  uses the NEW critic and OLD alpha
//   alpha  step (:713-731)  uses log-probs of the OLD actor (the ones the actor loss sampled)
// Networks: MultiHeadNetwork (mtrl/nn/multi_head.py:21-68) trunk Dense+ReLU layers as tcgen05 GEMMs,
// own-task heads as CUDA-core row dots; critic input is (action, state) (mtrl/rl/networks.py:61).
#include <algorithm>
#include <map>
#include <vector>

#include "comm.cuh"
#include "common.cuh"
#include "ln_kernels.cuh"
#include "mtrl_b200.h"
#include "net_common.cuh"
#include "sac_kernels.cuh"

using namespace sac;
using namespace netc;

namespace {

struct Workspace {
  float *Xa_next, *Xa, *Xc_next, *Xc;
  float *An[MTRL_MAX_DEPTH], *Ao[MTRL_MAX_DEPTH];
  float *C[kMaxE][MTRL_MAX_DEPTH], *Tg[kMaxE][MTRL_MAX_DEPTH];
  float* G[kMaxE][2];
  float* dXin;
  float *rew, *done, *eps_c, *eps_a, *logp_next, *logp, *act, *act_data, *logstd, *dq, *dout, *colsum_part, *alpha_val, *task_w;
  // head outputs accumulated by the last trunk layer's GEMM epilogue (fused heads, without the head bias): actor on
  // next_obs / obs [M][2A], critic members on (a, s) / target members on (a', s') / critic members on (pi(s), s) [M]
  float *pre_an, *pre_ao, *pre_c[kMaxE], *pre_tg[kMaxE], *pre_pi[kMaxE];
  long long pre_floats;   // the region [pre_an, pre_an + pre_floats) is zeroed at the start of every update
  unsigned* inrange;
  unsigned *bits_Ao[MTRL_MAX_DEPTH], *bits_C[kMaxE][MTRL_MAX_DEPTH];  // ReLU sign bits of Ao[l] / C[e][l], [M][W/32]
  int *row_slot, *slot_src, *tile_task, *seg_start, *status;
  double* acc;
  // fp32x3 (MTRL_PRECISION_FP32X3): every GEMM operand buffer above (inputs, activations, dZ) has its tf32 remainder at
  // the same offset in a mirror region, `lo_delta` floats further on (0 in tf32 mode); the remainders of the three
  // parameter operand copies (shadows) follow.
  long long lo_delta;
  float *ash_lo, *csh_lo, *tsh_lo;
  // MLP with LayerNorm / skip connections (ln_kernels.cuh): n_j = the input of Dense_j, j = 1 .. depth, stored at index
  // j - 1 (index depth - 1 = the head's input), per chain like the activations above, and the rows' (mean, rstd);
  // backward: GN = gradient of n_j (dX GEMM / head VJP output, fp32), GS = g(x_j) travelling through a skip connection
  // (ping-pong), rowc = per-row coefficients, part_dg / part_dbeta = column partials of the LayerNorm gradients.
  float *An_n[MTRL_MAX_DEPTH], *Ao_n[MTRL_MAX_DEPTH], *C_n[kMaxE][MTRL_MAX_DEPTH], *Tg_n[kMaxE][MTRL_MAX_DEPTH];
  float *An_st[MTRL_MAX_DEPTH], *Ao_st[MTRL_MAX_DEPTH], *C_st[kMaxE][MTRL_MAX_DEPTH], *Tg_st[kMaxE][MTRL_MAX_DEPTH];
  float *GN[kMaxE], *GS[kMaxE][2], *rowc[kMaxE], *part_dg[kMaxE], *part_dbeta[kMaxE];
};

// Bump allocation; with base == nullptr only the size is computed.
long long carve(const mtrl_sac_config_t& c, int Ka, int Kc, long long actor_total, long long critic_total, uint8_t* base,
                Workspace* ws) {
  long long off = 0;
  auto take = [&](long long bytes) -> uint8_t* {
    uint8_t* p = base ? base + off : nullptr;
    off = round_up(off + bytes, 256);
    return p;
  };
  const long long M = c.max_rows, W = c.width, A = c.action_dim, E = c.num_critics;
  auto f = [&](long long n) { return reinterpret_cast<float*>(take(n * 4 + 256)); };  // +256: slack for vector tails
  Workspace w;
  memset(&w, 0, sizeof(w));
  w.Xa_next = f(M * Ka);
  w.Xa = f(M * Ka);
  w.Xc_next = f(M * Kc);
  w.Xc = f(M * Kc);
  for (int l = 0; l < c.depth; ++l) {
    w.An[l] = f(M * W);
    w.Ao[l] = f(M * W);
    for (int e = 0; e < E; ++e) {
      w.C[e][l] = f(M * W);
      w.Tg[e][l] = f(M * W);
    }
  }
  for (int e = 0; e < E; ++e) {
    w.G[e][0] = f(M * W);
    w.G[e][1] = f(M * W);
  }
  const bool ln_mode = c.use_layer_norm || c.use_skip_connections;
  if (ln_mode) {
    for (int l = 0; l < c.depth; ++l) {
      w.An_n[l] = f(M * W);
      w.Ao_n[l] = f(M * W);
      for (int e = 0; e < E; ++e) {
        w.C_n[e][l] = f(M * W);
        w.Tg_n[e][l] = f(M * W);
      }
    }
  }
  const long long mirror_bytes = off;   // everything above is a GEMM operand
  if (ln_mode) {
    for (int l = 0; l < c.depth; ++l) {
      w.An_st[l] = f(M * 2);
      w.Ao_st[l] = f(M * 2);
      for (int e = 0; e < E; ++e) {
        w.C_st[e][l] = f(M * 2);
        w.Tg_st[e][l] = f(M * 2);
      }
    }
    for (int e = 0; e < E; ++e) {
      w.GN[e] = f(M * W);
      w.GS[e][0] = f(M * W);
      w.GS[e][1] = f(M * W);
      w.rowc[e] = f(M * 2);
      w.part_dg[e] = f((M / kTileRows) * W);
      w.part_dbeta[e] = f((M / kTileRows) * W);
    }
  }
  w.dXin = f(E * M * 16);
  w.rew = f(M);
  w.done = f(M);
  w.eps_c = f(M * A);
  w.eps_a = f(M * A);
  w.logp_next = f(M);
  w.logp = f(M);
  w.act = f(M * A);
  w.act_data = f(M * A);
  w.logstd = f(M * A);
  w.dq = f(E * M);
  w.dout = f(M * 2 * A);
  // bias-gradient column partials, one slot per (layer, member): [M/32 row groups][W] (a fused backward launch fills the
  // slots of all layers before one finishing kernel sums them)
  w.colsum_part = f(static_cast<long long>(c.depth) * kMaxE * (M / 32) * W);
  w.alpha_val = f(c.num_local_tasks);
  w.task_w = f(c.num_local_tasks);
  {
    const long long o0 = off;
    w.pre_an = f(M * 2 * A);
    w.pre_ao = f(M * 2 * A);
    for (int e = 0; e < E; ++e) {
      w.pre_c[e] = f(M);
      w.pre_tg[e] = f(M);
      w.pre_pi[e] = f(M);
    }
    w.pre_floats = (off - o0) / 4;
  }
  w.inrange = reinterpret_cast<unsigned*>(take(M * 4));
  const long long bit_words = M * ((W + 31) / 32);
  for (int l = 0; l + 1 < c.depth; ++l) {
    w.bits_Ao[l] = reinterpret_cast<unsigned*>(take(bit_words * 4));
    for (int e = 0; e < E; ++e) w.bits_C[e][l] = reinterpret_cast<unsigned*>(take(bit_words * 4));
  }
  w.row_slot = reinterpret_cast<int*>(take(static_cast<long long>(c.max_batch) * 4));
  w.slot_src = reinterpret_cast<int*>(take(M * 4));
  w.tile_task = reinterpret_cast<int*>(take(M / kTileRows * 4));
  w.seg_start = reinterpret_cast<int*>(take((c.num_local_tasks + 1) * 4));
  w.status = reinterpret_cast<int*>(take(16));
  w.acc = reinterpret_cast<double*>(take(ACC_COUNT * 8));
  if (c.precision == MTRL_PRECISION_FP32X3) {
    w.lo_delta = off / 4;
    take(mirror_bytes);
    w.ash_lo = f(actor_total);
    w.csh_lo = f(critic_total);
    w.tsh_lo = f(critic_total);
  }
  if (ws) *ws = w;
  return off;
}

int validate(const mtrl_sac_config_t& c) {
  MTRL_REQUIRE(c.num_tasks >= 1 && c.num_local_tasks >= 1 && c.task_begin >= 0 &&
                   c.task_begin + c.num_local_tasks <= c.num_tasks,
               "sac config: bad task range [%d, %d) of %d", c.task_begin, c.task_begin + c.num_local_tasks, c.num_tasks);
  MTRL_REQUIRE(c.obs_dim > c.num_tasks, "sac config: obs_dim %d must exceed num_tasks %d (one-hot suffix)", c.obs_dim,
               c.num_tasks);
  MTRL_REQUIRE(c.action_dim >= 1 && c.action_dim <= kMaxA, "sac config: action_dim %d outside [1, %d]", c.action_dim, kMaxA);
  MTRL_REQUIRE(c.obs_dim + c.action_dim >= 16, "sac config: critic input narrower than 16 is not supported");
  MTRL_REQUIRE(c.width >= 16 && c.width % 4 == 0, "sac config: width %d must be a multiple of 4 and >= 16", c.width);
  MTRL_REQUIRE(c.depth >= 1 && c.depth <= MTRL_MAX_DEPTH, "sac config: depth %d outside [1, %d]", c.depth, MTRL_MAX_DEPTH);
  MTRL_REQUIRE(c.num_critics >= 1 && c.num_critics <= kMaxE, "sac config: num_critics %d outside [1, %d]", c.num_critics,
               kMaxE);
  MTRL_REQUIRE(c.max_rows >= kTileRows && c.max_rows % kTileRows == 0, "sac config: max_rows %d must be a multiple of %d",
               c.max_rows, kTileRows);
  MTRL_REQUIRE(c.max_batch >= 1 && c.max_batch <= c.max_rows, "sac config: max_batch %d outside [1, max_rows]", c.max_batch);
  MTRL_REQUIRE(c.variant == MTRL_VARIANT_MTSAC || c.variant == MTRL_VARIANT_SAC, "sac config: unknown variant %d", c.variant);
  MTRL_REQUIRE(c.variant != MTRL_VARIANT_SAC || (c.num_tasks == 1 && c.num_local_tasks == 1 && !c.use_task_weights && !c.clip_q),
               "sac config: the single-task variant needs num_tasks == 1, no task weights, no Q clipping");
  MTRL_REQUIRE(!(c.use_task_weights && c.num_local_tasks != c.num_tasks),
               "sac config: use_task_weights needs all tasks on one handle (softmax over every log_alpha)");
  MTRL_REQUIRE(c.precision == MTRL_PRECISION_TF32 || c.precision == MTRL_PRECISION_FP32X3, "sac config: unknown precision %d",
               c.precision);
  MTRL_REQUIRE(!(c.use_layer_norm || c.use_skip_connections) || c.variant == MTRL_VARIANT_SAC,
               "sac config: LayerNorm / skip connections belong to the MLP of the single-task variant (MultiHeadNetwork has neither)");
  MTRL_REQUIRE(!c.use_skip_connections || (c.obs_dim != c.width && c.obs_dim + c.action_dim != c.width),
               "sac config: skip connections with an input as wide as the hidden layers are not supported");
  return MTRL_OK;
}

}  // namespace

struct mtrl_sac {
  mtrl_sac_config_t cfg;
  mtrl_sac_buffers_t buf;
  mtrl_sac_layout_t lay;
  Workspace ws;
  int sms = 148;
  // GEMM plans
  std::vector<mtrl_gemm_plan_t*> fwd;          // [depth] actor(next), actor(obs), critic members
  std::vector<mtrl_gemm_plan_t*> fwd_target;   // [depth]
  std::vector<mtrl_gemm_plan_t*> bwd_critic;   // [depth] (index l: dW_l of every member + dZ_{l-1})
  std::vector<mtrl_gemm_plan_t*> fwd_pi;       // [depth] critic (new params) on (pi(s), s)
  std::vector<mtrl_gemm_plan_t*> bwd_pi;       // [depth] dX only
  std::vector<mtrl_gemm_plan_t*> bwd_actor;    // [depth]
  // The layers of each list above as ONE launch (make_fused_plan: row-tile dependencies for the chains, grid barriers for the
  // backward passes with weight gradients); null = layer by layer.
  enum Fused { F_FWD = 0, F_FWD_TARGET, F_FWD_PI, F_BWD_CRITIC, F_BWD_PI, F_BWD_ACTOR, F_COUNT };
  mtrl_gemm_plan_t* fused[F_COUNT] = {};
  // the output heads ride in the epilogue of the last trunk layer's GEMM (mtrl_gemm_problem_t::head_w): the loss / sampling
  // kernels then read M x head_dim numbers instead of the M x W activation
  bool fused_heads = false;
  // Polyak target update on a side stream, forked after the critic's Adam step and joined at the end of the update
  bool defer_polyak = false, polyak_pending = false;
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int launches = 0;
  int batch = 0, global_batch = 0;
  // optional CUDA-event bracketing of every GEMM launch (bench.py's live roofline measurement)
  bool prof = false;
  std::vector<cudaEvent_t> ev;
  std::vector<int> ev_tag;   // per event pair: 0 = GEMM launch, 1 = exchange kernel (trunk step / rank barrier)
  size_t ev_used = 0;
  double exchange_ms = 0.0;
  int exchange_launches = 0;
  double class_ms[12] = {};
  int class_n[12] = {};
  // fused peer-memory exchange (comm.cuh); null = single GPU, or the caller all-reduces between the phases
  mtrl_comm* comm = nullptr;
  long long off_critic_grads = 0, off_actor_grads = 0, off_critic_params = 0, off_actor_params = 0;
  std::map<int, std::vector<mtrl_gemm_plan_t*>> act_plans;   // actor forward plans of mtrl_sac_act, by row count
  std::map<std::pair<int, int>, std::vector<mtrl_gemm_plan_t*>> mlp_plans;   // mtrl_mlp_forward plans by (network, rows)
  unsigned long long act_calls = 0;
  // per-task gradient path (mtrl_sac_task_grads): grouped per-task dW plans [critic/actor][layer] -> launches
  struct TaskGradCache {
    float *critic_tg = nullptr, *actor_tg = nullptr;
    int rows_per_task = 0;
    bool fill_critic = true, fill_actor = true;                  // which networks' backward fills its matrix
    std::vector<std::vector<mtrl_gemm_plan_t*>> critic, actor;   // [layer index i = D-1-l][launch]
  } tgc;
  // PCGrad (mtrl_sac_enable_pcgrad): which optimiser chains start with pcgrad, scratch and the row permutations
  bool pcgrad_critic = false, pcgrad_actor = false;
  int surgery_mode = 0;   // 0: pcgrad, 1: cagrad, 2: gradnorm (3: gradnorm with per-task clipping), 4: dummy (mean)
  float* pcgrad_scratch = nullptr;
  const int *pcgrad_perm_critic = nullptr, *pcgrad_perm_actor = nullptr;
  TaskGradCache* tg_active = nullptr;   // set while mtrl_sac_task_grads runs: the backward also fills the (T, P) rows
  // The reference's split-loss branches (taken when an optimiser requires per-task losses, and by compute_weights) are not
  // the un-split loss evaluated per task; both differences are reproduced:
  //   split_critic: a' and log pi(a'|.) are sampled from the actor on data.OBSERVATIONS, not next_observations
  //                 (mtsac.py:515-523, 995-1003), while the target critics still see next_observations;
  //   split_actor:  the vmapped actor loss runs with `_explore` at its default True (mtsac.py:631-637, 676-682).
  bool split_critic = false, split_actor = false;
  bool ln_mode = false;   // MLP with LayerNorm and / or skip connections: junction kernels between the Dense GEMMs
  comm::Segment *d_segs_critic = nullptr, *d_segs_actor = nullptr;   // ownership tables (device)
  int nsegs_critic = 0, nsegs_actor = 0;
};

namespace {

// column partials of dZ_l of member e (bias gradient of layer l)
float* colsum_part(const mtrl_sac* h, int e, int l) {
  return h->ws.colsum_part + (static_cast<long long>(l) * kMaxE + e) * (h->cfg.max_rows / 32) * h->cfg.width;
}

// fp32x3: the tf32 remainder of a workspace operand buffer (nullptr in tf32 mode); works for interior pointers.
template <typename Ptr>
Ptr lo(const mtrl_sac* h, Ptr p) { return h->ws.lo_delta ? p + h->ws.lo_delta : nullptr; }
// ... and of a trunk kernel inside one of the three parameter operand copies
float* tk_lo(float* base_lo, const mtrl_net_layout_t& L, int e, int l) { return base_lo ? tk(base_lo, L, e, l) : nullptr; }

// The four activation chains of an update: actor on next_obs / obs, critic members (online params) and target members.
enum Chain { CH_AN = 0, CH_AO = 1, CH_C = 2, CH_TG = 3 };
float* chain_d(const mtrl_sac* h, int ch, int e, int l) {   // relu output of Dense_l
  const Workspace& w = h->ws;
  return ch == CH_AN ? w.An[l] : (ch == CH_AO ? w.Ao[l] : (ch == CH_C ? w.C[e][l] : w.Tg[e][l]));
}
float* chain_n(const mtrl_sac* h, int ch, int e, int l) {   // n_{l+1} = LN_l(x_{l+1}): the input of Dense_{l+1} / of the head
  const Workspace& w = h->ws;
  return ch == CH_AN ? w.An_n[l] : (ch == CH_AO ? w.Ao_n[l] : (ch == CH_C ? w.C_n[e][l] : w.Tg_n[e][l]));
}
float* chain_st(const mtrl_sac* h, int ch, int e, int l) {
  const Workspace& w = h->ws;
  return ch == CH_AN ? w.An_st[l] : (ch == CH_AO ? w.Ao_st[l] : (ch == CH_C ? w.C_st[e][l] : w.Tg_st[e][l]));
}
// input of Dense_l for l >= 1, and of the head
float* dense_in(const mtrl_sac* h, int ch, int e, int l) { return h->ln_mode ? chain_n(h, ch, e, l - 1) : chain_d(h, ch, e, l - 1); }
float* head_in(const mtrl_sac* h, int ch, int e) { return dense_in(h, ch, e, h->cfg.depth); }

// Rows [r0, r1) of a hidden-layer kernel (in = W rows) that rank r owns when the trunk is sharded over G ranks.
void row_block(int W, int G, int r, int* r0, int* r1) {
  const int rb = static_cast<int>(round_up((W + G - 1) / G, 32));
  *r0 = r * rb < W ? r * rb : W;
  *r1 = *r0 + rb < W ? *r0 + rb : W;
}

// Ownership table of one network's trunk (comm.cuh Segment): hidden-layer kernels by row blocks (their gradients are
// reduced by the dW GEMM epilogues), first-layer kernel + bias and the other biases as whole pieces dealt round-robin.
std::vector<comm::Segment> trunk_segments(const mtrl_net_layout_t& L, int G) {
  std::vector<comm::Segment> out;
  int piece = 0;
  for (int e = 0; e < L.members; ++e) {
    const long long base = e * L.member_trunk_stride;
    const long long first_end = L.depth > 1 ? L.kernel_off[1] : L.member_trunk_stride;
    out.push_back({(base + 0) / 4, (base + first_end) / 4, piece++ % G, 0});
    for (int l = 1; l < L.depth; ++l) {
      for (int r = 0; r < G; ++r) {
        int r0, r1;
        row_block(L.width, G, r, &r0, &r1);
        if (r1 <= r0) continue;
        const long long b = base + L.kernel_off[l] + static_cast<long long>(r0) * L.width;
        // the last block also takes the alignment padding between the kernel and its bias (always zero)
        const long long e_ = r1 == L.width ? base + L.bias_off[l] : base + L.kernel_off[l] + static_cast<long long>(r1) * L.width;
        out.push_back({b / 4, e_ / 4, r, 1});
      }
      const long long bias_end = l + 1 < L.depth ? L.kernel_off[l + 1] : L.member_trunk_stride;
      out.push_back({(base + L.bias_off[l]) / 4, (base + bias_end) / 4, piece++ % G, 0});
    }
  }
  return out;
}

// dW problems of layer l >= 1 of one member.  Single device: one problem into the local gradient buffer.  Sharded
// with the peer-memory exchange: one problem per owner rank, whose epilogue reduce-adds (TMA cp.reduce over NVLink)
// the owner's row block straight into the OWNER's gradient buffer -- the reduce-scatter happens tile by tile inside
// the GEMM, overlapped with its main loop.
void push_dw(mtrl_sac* h, std::vector<mtrl_gemm_problem_t>& dst, const float* X, int ldx, int n_in, const float* dZ, float* grads,
             long long off_grads, const mtrl_net_layout_t& L, int e, int l) {
  const int M = h->cfg.max_rows, W = h->cfg.width;
  if (!h->comm || l == 0) {
    dst.push_back(dw_problem(X, ldx, n_in, dZ, tk(grads, L, e, l), M, W, h->sms, 0, lo(h, X), lo(h, dZ)));
    return;
  }
  const mtrl_comm* c = h->comm;
  for (int r = 0; r < c->world; ++r) {
    int r0, r1;
    row_block(W, c->world, r, &r0, &r1);
    if (r1 <= r0) continue;
    float* owner_grads = reinterpret_cast<float*>(c->peer[r] + off_grads);
    mtrl_gemm_problem_t p = dw_problem(X + r0, ldx, r1 - r0, dZ, tk(owner_grads, L, e, l) + static_cast<long long>(r0) * W, M, W,
                                       h->sms, 0, lo(h, X + r0), lo(h, dZ));
    p.epilogue = MTRL_EPI_ATOMIC_ADD;   // every rank adds its rows' contribution (the buffer is zeroed per update)
    // MTRL_REMOTE_FIRST=1 (experiment): deal the tiles whose reduce-adds cross NVLink first so the transfers overlap the
    // launch's other tiles.  Measured SLOWER at 2 GPUs (2.04 vs 1.98 ms/step: the long dX units then go last and unbalance
    // the schedule), so the default keeps the plain longest-first order.
    p.schedule_first = r != c->rank && getenv("MTRL_REMOTE_FIRST") && getenv("MTRL_REMOTE_FIRST")[0] == '1';
    dst.push_back(p);
  }
}

// One launch for all layers of a trunk pass, out of the per-layer problem lists (launch order = phase order).
//   chain = true  (forward passes, the input-gradient-only backward): every layer's A operand is the previous layer's output,
//                 so the launch is ordered by per-row-tile dependencies (MTRL_GEMM_ROWDEPS): SMs that run out of tiles of one
//                 layer continue with the next.  Default where that pays under graph replay (same-box A/B): few row tiles and
//                 wide layers -- one rank's share of a sharded batch (1/8 of MT50/W2048: 0.924 -> 0.899 ms per update), MT10 at
//                 W1024 (0.526 -> 0.514) -- and not at MT10/W400 (+4 %: launches that short are cheaper than the counters) nor
//                 at MT50 on one GPU (11 rounds of tiles per layer: nothing to fill).
//   chain = false (backward passes with weight gradients): grid barriers between the layers, only on request.
// MTRL_FUSE_LAYERS=0: never; =1: always, both kinds.  Too many problems for one launch (the per-owner dW problems of a sharded
// trunk) leaves the slot empty and the caller runs layer by layer.
int make_fused_plan(mtrl_sac* h, int slot, const std::vector<std::vector<mtrl_gemm_problem_t>>& layers, bool chain) {
  if (h->fused[slot]) {
    mtrl_gemm_plan_destroy(h->fused[slot]);
    h->fused[slot] = nullptr;
  }
  const char* env = getenv("MTRL_FUSE_LAYERS");
  const bool never = h->ln_mode || layers.size() < 2 || (env && env[0] == '0');
  const bool always = env && env[0] == '1';
  // (the two-network passes keep paying up to one rank's share of a 2-GPU split: -2.2 % per update at 3200 rows; at the full 6400
  // rows the event-bracketed GEMM time still drops 1.7 % but the replayed step does not, 3.211 vs 3.195 ms, so the bound stays
  // at 3200; the four-network forward pass stops at 2048 rows.  MTRL_CHAIN_ROWS2 moves the bound.)
  const int rows2 = getenv("MTRL_CHAIN_ROWS2") ? atoi(getenv("MTRL_CHAIN_ROWS2")) : 3200;
  const bool by_shape = chain && h->cfg.width >= 1024 && h->cfg.max_rows <= (slot == mtrl_sac::F_FWD ? 2048 : rows2);
  if (never || !(always || by_shape)) return MTRL_OK;
  std::vector<mtrl_gemm_problem_t> all;
  for (size_t i = 0; i < layers.size(); ++i)
    for (mtrl_gemm_problem_t p : layers[i]) {
      p.phase = static_cast<int>(i);
      all.push_back(p);
    }
  if (static_cast<int>(all.size()) > MTRL_GEMM_MAX_PROBLEMS || static_cast<int>(layers.size()) > MTRL_GEMM_MAX_PHASES) return MTRL_OK;
  std::vector<mtrl_gemm_plan_t*> out;
  MTRL_PROPAGATE(make_plan(out, all, chain ? MTRL_GEMM_ROWDEPS : 0));
  h->fused[slot] = out[0];
  return MTRL_OK;
}

int build_backward_plans(mtrl_sac* h) {
  const mtrl_sac_config_t& c = h->cfg;
  const mtrl_net_layout_t& LA = h->lay.actor;
  const mtrl_net_layout_t& LC = h->lay.critic;
  const Workspace& w = h->ws;
  const int M = c.max_rows, W = c.width, D = c.depth, E = c.num_critics;
  const int Ka = h->lay.k_actor, Kc = h->lay.k_critic;
  float* ash = h->buf.actor_shadow;
  float* csh = h->buf.critic_shadow;
  float* ash_lo = w.ash_lo;
  float* csh_lo = w.csh_lo;
  for (auto* v : {&h->bwd_critic, &h->bwd_pi, &h->bwd_actor}) {
    for (auto* p : *v) mtrl_gemm_plan_destroy(p);
    v->clear();
  }
  // Backward: dZ_{D-1} (masked head VJP) sits in G[e][0]; layer l reads G[e][(D-1-l)&1], writes G[e][(D-l)&1].
  // MLP with LayerNorm / skip (ln_mode): the junction kernels produce dZ_l in G[e][0] for every layer; the dX problem
  // stores the plain gradient of the layer's input n_l in GN[e] (fp32, no ReLU gate: the junction below applies it).
  const bool ln = h->ln_mode;
  auto dx_any = [&](int e, int src, int dst, float* sh, float* sh_lo, const mtrl_net_layout_t& L, int l, unsigned* bits, float* csum) {
    if (!ln)
      return dx_problem(w.G[e][src], tk(sh, L, e, l), W, bits, w.G[e][dst], M, W, csum, lo(h, w.G[e][src]), tk_lo(sh_lo, L, e, l),
                        lo(h, w.G[e][dst]));
    mtrl_gemm_problem_t p = dx_problem(w.G[e][0], tk(sh, L, e, l), W, nullptr, w.GN[e], M, W, nullptr, lo(h, w.G[e][0]),
                                       tk_lo(sh_lo, L, e, l), nullptr);
    p.epilogue = MTRL_EPI_STORE;
    return p;
  };
  std::vector<std::vector<mtrl_gemm_problem_t>> all_c, all_pi, all_a;
  for (int l = D - 1; l >= 0; --l) {
    const int src = ln ? 0 : (D - 1 - l) & 1, dst = src ^ 1;
    std::vector<mtrl_gemm_problem_t> pc, ppi, pa;
    for (int e = 0; e < E; ++e) {
      const float* X = l == 0 ? w.Xc : dense_in(h, CH_C, e, l);
      push_dw(h, pc, X, l == 0 ? Kc : W, l == 0 ? LC.in_dim : W, w.G[e][src], h->buf.critic_grads, h->off_critic_grads, LC, e, l);
    }
    for (int e = 0; e < E; ++e) {
      if (l > 0) {
        pc.push_back(dx_any(e, src, dst, csh, csh_lo, LC, l, ln ? nullptr : w.bits_C[e][l - 1], colsum_part(h, e, l - 1)));
        ppi.push_back(dx_any(e, src, dst, csh, csh_lo, LC, l, ln ? nullptr : w.bits_C[e][l - 1], nullptr));
      } else {
        // actor step: only dL/da = first action_dim input columns of dZ_0 W_0^T (N = 16 rows of W_0)
        mtrl_gemm_problem_t p;
        memset(&p, 0, sizeof(p));
        p.A = w.G[e][src]; p.lda = W; p.a_major = 0;
        p.B = tk(csh, LC, e, 0); p.ldb = W; p.b_major = 0;
        p.D = w.dXin + static_cast<long long>(e) * M * 16; p.ldd = 16;
        // a 16-column tile is latency-bound per k-block: split K so the (few) tiles spread over the SMs
        p.M = M; p.N = 16; p.K = W; p.block_n = 16;
        p.k_splits = W >= 1024 ? 8 : (W >= 256 ? 2 : 1);
        p.epilogue = p.k_splits > 1 ? MTRL_EPI_ATOMIC_ADD : MTRL_EPI_STORE;
        p.A_lo = lo(h, w.G[e][src]);
        p.B_lo = tk_lo(csh_lo, LC, e, 0);
        ppi.push_back(p);
      }
    }
    push_dw(h, pa, l == 0 ? w.Xa : dense_in(h, CH_AO, 0, l), l == 0 ? Ka : W, l == 0 ? LA.in_dim : W, w.G[0][src], h->buf.actor_grads,
            h->off_actor_grads, LA, 0, l);
    if (l > 0) pa.push_back(dx_any(0, src, dst, ash, ash_lo, LA, l, ln ? nullptr : w.bits_Ao[l - 1], colsum_part(h, 0, l - 1)));
    MTRL_PROPAGATE(make_plan(h->bwd_critic, pc));
    MTRL_PROPAGATE(make_plan(h->bwd_pi, ppi));
    MTRL_PROPAGATE(make_plan(h->bwd_actor, pa));
    all_c.push_back(pc);
    all_pi.push_back(ppi);
    all_a.push_back(pa);
  }
  MTRL_PROPAGATE(make_fused_plan(h, mtrl_sac::F_BWD_CRITIC, all_c, false));
  MTRL_PROPAGATE(make_fused_plan(h, mtrl_sac::F_BWD_PI, all_pi, true));
  MTRL_PROPAGATE(make_fused_plan(h, mtrl_sac::F_BWD_ACTOR, all_a, false));
  return MTRL_OK;
}

int build_plans(mtrl_sac* h) {
  const mtrl_sac_config_t& c = h->cfg;
  const mtrl_net_layout_t& LA = h->lay.actor;
  const mtrl_net_layout_t& LC = h->lay.critic;
  const Workspace& w = h->ws;
  const int M = c.max_rows, W = c.width, D = c.depth, E = c.num_critics;
  const int Ka = h->lay.k_actor, Kc = h->lay.k_critic;
  float* ash = h->buf.actor_shadow;
  float* csh = h->buf.critic_shadow;
  float* tsh = h->buf.critic_target_shadow;
  // forward problem of one trunk layer of one chain: input / kernel / output with their fp32x3 remainders (null in tf32
  // mode).  With LayerNorm / skip the layer's input is the junction's output n_l and the ReLU gate bits are not needed
  // (the junction reads the activation itself).
  auto fwd = [&](int ch, float* X0, int K0, int in_dim, float* sh, float* sh_lo, float* params, const mtrl_net_layout_t& L, int e, int l,
                 unsigned* bits, float* head_pre) {
    float* X = l == 0 ? X0 : dense_in(h, ch, e, l);
    float* out = chain_d(h, ch, e, l);
    mtrl_gemm_problem_t p = fwd_problem(X, l == 0 ? K0 : W, l == 0 ? in_dim : W, tk(sh, L, e, l), tb(params, L, e, l), out, M, W,
                                        h->ln_mode ? nullptr : bits, lo(h, X), tk_lo(sh_lo, L, e, l), lo(h, out));
    if (h->fused_heads && l == D - 1) {
      p.head_w = hk(params, L, e);
      p.head_out = head_pre;
      p.head_tile_task = w.tile_task;
      p.head_dim = L.head_dim;
    }
    return p;
  };
  std::vector<std::vector<mtrl_gemm_problem_t>> all_f, all_t, all_pi;
  for (int l = 0; l < D; ++l) {
    std::vector<mtrl_gemm_problem_t> p;
    p.push_back(fwd(CH_AN, w.Xa_next, Ka, LA.in_dim, ash, w.ash_lo, h->buf.actor_params, LA, 0, l, nullptr, w.pre_an));
    p.push_back(fwd(CH_AO, w.Xa, Ka, LA.in_dim, ash, w.ash_lo, h->buf.actor_params, LA, 0, l, l + 1 < D ? w.bits_Ao[l] : nullptr, w.pre_ao));
    for (int e = 0; e < E; ++e)
      p.push_back(fwd(CH_C, w.Xc, Kc, LC.in_dim, csh, w.csh_lo, h->buf.critic_params, LC, e, l, l + 1 < D ? w.bits_C[e][l] : nullptr,
                      w.pre_c[e]));
    MTRL_PROPAGATE(make_plan(h->fwd, p));
    all_f.push_back(p);
  }
  for (int l = 0; l < D; ++l) {
    std::vector<mtrl_gemm_problem_t> p, q;
    for (int e = 0; e < E; ++e) {
      p.push_back(fwd(CH_TG, w.Xc_next, Kc, LC.in_dim, tsh, w.tsh_lo, h->buf.critic_target, LC, e, l, nullptr, w.pre_tg[e]));
      q.push_back(fwd(CH_C, w.Xc, Kc, LC.in_dim, csh, w.csh_lo, h->buf.critic_params, LC, e, l, l + 1 < D ? w.bits_C[e][l] : nullptr,
                      w.pre_pi[e]));
    }
    MTRL_PROPAGATE(make_plan(h->fwd_target, p));
    MTRL_PROPAGATE(make_plan(h->fwd_pi, q));
    all_t.push_back(p);
    all_pi.push_back(q);
  }
  MTRL_PROPAGATE(make_fused_plan(h, mtrl_sac::F_FWD, all_f, true));
  MTRL_PROPAGATE(make_fused_plan(h, mtrl_sac::F_FWD_TARGET, all_t, true));
  MTRL_PROPAGATE(make_fused_plan(h, mtrl_sac::F_FWD_PI, all_pi, true));
  return build_backward_plans(h);
}

#define LAUNCHED(h) ((h)->launches++)

// The deferred Polyak update (polyak_kernel): fork from `st` behind everything enqueued so far, run on the side stream.
int fork_polyak(mtrl_sac* h, cudaStream_t st) {
  const mtrl_net_layout_t& LC = h->lay.critic;
  MTRL_CUDA_CHECK(cudaEventRecord(h->ev_fork, st));
  MTRL_CUDA_CHECK(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
  polyak_kernel<<<h->sms * 2, 256, 0, h->side>>>(h->buf.critic_params, h->buf.critic_target, h->buf.critic_target_shadow, h->ws.tsh_lo,
                                                 LC.total, LC.trunk_total, h->cfg.tau);
  MTRL_CUDA_CHECK(cudaGetLastError());
  MTRL_CUDA_CHECK(cudaEventRecord(h->ev_join, h->side));
  h->polyak_pending = true;
  LAUNCHED(h);
  return MTRL_OK;
}
// ... and `st` continues only after it (end of the update; also before anything else that touches the target).
int join_polyak(mtrl_sac* h, cudaStream_t st) {
  if (!h->polyak_pending) return MTRL_OK;
  MTRL_CUDA_CHECK(cudaStreamWaitEvent(st, h->ev_join, 0));
  h->polyak_pending = false;
  return MTRL_OK;
}

// Kernel classes of the event-bracketed profiling pass (mtrl_sac_profile_gemms / _read / _classes).
enum ProfTag { PT_GEMM = 0, PT_EXCHANGE = 1, PT_ADAM = 2, PT_HEAD_BWD = 3, PT_CRITIC_LOSS = 4, PT_ACTOR_HEAD = 5, PT_ACTOR_LOSS = 6,
               PT_PACK = 7, PT_SUMSQ = 8, PT_COLSUM = 9, PT_JUNCTION = 10, PT_COUNT = 12 };

int prof_begin(mtrl_sac* h, int tag, cudaStream_t st) {
  if (!h->prof) return MTRL_OK;
  while (h->ev.size() < h->ev_used + 2) {
    cudaEvent_t e;
    MTRL_CUDA_CHECK(cudaEventCreate(&e));
    h->ev.push_back(e);
  }
  if (h->ev_tag.size() < h->ev.size() / 2) h->ev_tag.resize(h->ev.size() / 2, 0);
  h->ev_tag[h->ev_used / 2] = tag;
  MTRL_CUDA_CHECK(cudaEventRecord(h->ev[h->ev_used], st));
  return MTRL_OK;
}

int prof_end(mtrl_sac* h, cudaStream_t st) {
  if (!h->prof) return MTRL_OK;
  MTRL_CUDA_CHECK(cudaEventRecord(h->ev[h->ev_used + 1], st));
  h->ev_used += 2;
  return MTRL_OK;
}

int run_plan(mtrl_sac* h, mtrl_gemm_plan_t* plan, cudaStream_t st) {
  MTRL_PROPAGATE(prof_begin(h, 0, st));
  MTRL_PROPAGATE(mtrl_gemm_plan_run(plan, st));
  MTRL_PROPAGATE(prof_end(h, st));
  LAUNCHED(h);
  return MTRL_OK;
}

// One warp per row, weights streamed from L2 (any width; also the action-sampling path).
void launch_actor_head_rows(const ActorHeadArgs& a, int action_dim, dim3 grid, dim3 block, cudaStream_t st) {
  switch (action_dim) {
    case 1: actor_head_kernel<1><<<grid, block, 0, st>>>(a); break;
    case 2: actor_head_kernel<2><<<grid, block, 0, st>>>(a); break;
    case 3: actor_head_kernel<3><<<grid, block, 0, st>>>(a); break;
    case 4: actor_head_kernel<4><<<grid, block, 0, st>>>(a); break;
    case 5: actor_head_kernel<5><<<grid, block, 0, st>>>(a); break;
    case 6: actor_head_kernel<6><<<grid, block, 0, st>>>(a); break;
    case 7: actor_head_kernel<7><<<grid, block, 0, st>>>(a); break;
    default: actor_head_kernel<8><<<grid, block, 0, st>>>(a); break;
  }
}

int launch_actor_head(mtrl_sac* h, const float* H, const float* pre, const float* eps, float* Xdst, float* logp, bool save,
                      cudaStream_t st) {
  const mtrl_sac_config_t& c = h->cfg;
  ActorHeadArgs a;
  a.h_lo_delta = h->ws.lo_delta;
  a.H = H;
  a.pre = pre;
  a.Wh = hk(h->buf.actor_params, h->lay.actor, 0);
  a.bh = hb(h->buf.actor_params, h->lay.actor, 0);
  a.tile_task = h->ws.tile_task;
  a.slot_src = h->ws.slot_src;
  a.row_task = nullptr;
  a.eps = eps;
  a.Xdst = Xdst;
  a.xdst_lo_delta = h->ws.lo_delta;
  a.ldx = h->lay.k_critic;
  a.act = save ? h->ws.act : nullptr;
  a.logp = logp;
  a.logstd = save ? h->ws.logstd : nullptr;
  a.inrange = save ? h->ws.inrange : nullptr;
  a.M = c.max_rows;
  a.W = c.width;
  a.ls_min = c.log_std_min;
  a.ls_max = c.log_std_max;
  const size_t wbytes = static_cast<size_t>(c.width + 4) * 2 * c.action_dim * sizeof(float);
  MTRL_PROPAGATE(prof_begin(h, PT_ACTOR_HEAD, st));
  if (pre) {
    // the head's outputs came out of the GEMM epilogue: one thread per (row, action dimension) worth of work
    const int wpb = 8;
    launch_actor_head_rows(a, c.action_dim, dim3((c.max_rows + wpb - 1) / wpb), dim3(wpb * 32), st);
    MTRL_CUDA_CHECK(cudaGetLastError());
    MTRL_PROPAGATE(prof_end(h, st));
    LAUNCHED(h);
    return MTRL_OK;
  }
  if (wbytes <= 200 * 1024) {
    // 16-row blocks, three resident per SM: 6400 rows = 400 blocks = one wave
    const int rpb = c.max_rows <= 16 * 3 * h->sms ? 16 : 32;
    dim3 grid(c.max_rows / rpb), block(256);
#define MTRL_AH_TILE(A_) \
  case A_: mtrl_launch(actor_head_tile_kernel<A_>, grid, block, wbytes, st, a, rpb); break;
    switch (c.action_dim) {
      MTRL_AH_TILE(1) MTRL_AH_TILE(2) MTRL_AH_TILE(3) MTRL_AH_TILE(4) MTRL_AH_TILE(5) MTRL_AH_TILE(6) MTRL_AH_TILE(7)
      default: mtrl_launch(actor_head_tile_kernel<8>, grid, block, wbytes, st, a, rpb);
    }
#undef MTRL_AH_TILE
    MTRL_CUDA_CHECK(cudaGetLastError());
    MTRL_PROPAGATE(prof_end(h, st));
    LAUNCHED(h);
    return MTRL_OK;
  }
  const int wpb = 8;
  dim3 grid((c.max_rows + wpb - 1) / wpb), block(wpb * 32);
  launch_actor_head_rows(a, c.action_dim, grid, block, st);
  MTRL_CUDA_CHECK(cudaGetLastError());
  MTRL_PROPAGATE(prof_end(h, st));
  LAUNCHED(h);
  return MTRL_OK;
}

int launch_head_bwd(mtrl_sac* h, const HeadBwdArgs& a, int hd, int E, cudaStream_t st) {
  MTRL_PROPAGATE(prof_begin(h, PT_HEAD_BWD, st));
  if (!launch_head_bwd_any(a, hd, h->cfg.num_local_tasks, E, st)) {
    mtrl_set_error("head_bwd: unsupported head_dim %d", hd);
    return MTRL_ERR_UNSUPPORTED;
  }
  MTRL_CUDA_CHECK(cudaGetLastError());
  MTRL_PROPAGATE(prof_end(h, st));
  LAUNCHED(h);
  return MTRL_OK;
}

int launch_colsum(mtrl_sac* h, const ColsumJobs& jobs, int groups, cudaStream_t st) {
  const int W = h->cfg.width;
  dim3 g((W + 31) / 32, jobs.njobs);
  MTRL_PROPAGATE(prof_begin(h, PT_COLSUM, st));
  mtrl_launch(colsum_final_kernel, g, dim3(256), 0, st, jobs, groups, W);
  MTRL_CUDA_CHECK(cudaGetLastError());
  MTRL_PROPAGATE(prof_end(h, st));
  LAUNCHED(h);
  return MTRL_OK;
}

// One chain of an MLP pass through a junction: which activation buffers, and which parameter buffer holds its LayerNorms.
struct ChainRef {
  int ch, e;
  float* params;
  const mtrl_net_layout_t* L;
};

// Junction l + 1 of the given chains, right after their Dense_l: n_{l+1} = LayerNorm_l(d_l [+ n_l]) (ln_kernels.cuh).
int run_junctions_fwd(mtrl_sac* h, const ChainRef* chains, int n, int l, int rows, cudaStream_t st) {
  const mtrl_sac_config_t& c = h->cfg;
  LnFwdArgs a;
  memset(&a, 0, sizeof(a));
  MTRL_REQUIRE(n <= kMaxLnPasses, "run_junctions_fwd: %d chains in one launch", n);
  for (int i = 0; i < n; ++i) {
    const ChainRef& r = chains[i];
    LnFwdPass& p = a.p[i];
    p.D = chain_d(h, r.ch, r.e, l);
    p.Nprev = (c.use_skip_connections && l >= 1) ? chain_n(h, r.ch, r.e, l - 1) : nullptr;
    p.scale = c.use_layer_norm ? lns(r.params, *r.L, r.e, l) : nullptr;
    p.bias = c.use_layer_norm ? lnb(r.params, *r.L, r.e, l) : nullptr;
    p.N = chain_n(h, r.ch, r.e, l);
    p.stats = chain_st(h, r.ch, r.e, l);
  }
  a.npass = n; a.M = rows; a.W = c.width; a.lo_delta = h->ws.lo_delta; a.eps = 1e-6f;
  mtrl_launch(ln_fwd_kernel, dim3((rows + 7) / 8, n), dim3(256), 0, st, a);
  MTRL_CUDA_CHECK(cudaGetLastError());
  LAUNCHED(h);
  return MTRL_OK;
}

int launch_colsum(mtrl_sac* h, const ColsumJobs& jobs, int groups, cudaStream_t st);

// Trunk backward of an MLP with LayerNorm / skip connections: per layer l (from the top), junction l + 1 turns the
// gradient of n_{l+1} (GN: head VJP for l = depth - 1, the dX GEMM of layer l + 1 below that, plus what arrives through
// the skip connection) into dZ_l, the bias / LayerNorm gradient partials and the gradient that skips on; then the
// layer's dW / dX plan.
int run_trunk_backward_ln(mtrl_sac* h, std::vector<mtrl_gemm_plan_t*>& plans, float* grads, float* params, const mtrl_net_layout_t& L,
                          int ch, int E, bool want_wgrad, cudaStream_t st) {
  const mtrl_sac_config_t& c = h->cfg;
  Workspace& w = h->ws;
  const int D = c.depth, M = c.max_rows, W = c.width;
  const bool ln = c.use_layer_norm != 0, skip = c.use_skip_connections != 0;
  for (int l = D - 1, i = 0; l >= 0; --l, ++i) {
    const int j = l + 1;
    LnBwdArgs a;
    memset(&a, 0, sizeof(a));
    for (int e = 0; e < E; ++e) {
      LnBwdPass& p = a.p[e];
      p.dN = w.GN[e];
      p.dSkip = (skip && j >= 1 && j <= D - 1) ? w.GS[e][(j + 1) & 1] : nullptr;
      p.D = chain_d(h, ch, e, l);
      p.Nprev = (skip && l >= 1) ? chain_n(h, ch, e, l - 1) : nullptr;
      p.scale = ln ? lns(params, L, e, l) : nullptr;
      p.stats = chain_st(h, ch, e, l);
      p.rowc = w.rowc[e];
      p.dZ = w.G[e][0];
      p.dX = (skip && l >= 1) ? w.GS[e][j & 1] : nullptr;
      p.part_db = want_wgrad ? colsum_part(h, e, l) : nullptr;
      p.part_dg = (want_wgrad && ln) ? w.part_dg[e] : nullptr;
      p.part_dbeta = (want_wgrad && ln) ? w.part_dbeta[e] : nullptr;
    }
    a.npass = E; a.M = M; a.W = W; a.lo_delta = w.lo_delta;
    if (ln) {
      mtrl_launch(ln_bwd_rows_kernel, dim3((M + 7) / 8, E), dim3(256), 0, st, a);
      LAUNCHED(h);
    }
    mtrl_launch(ln_bwd_tile_kernel, dim3((W + 127) / 128, M / kTileRows, E), dim3(256), 0, st, a);
    MTRL_CUDA_CHECK(cudaGetLastError());
    LAUNCHED(h);
    if (want_wgrad) {
      ColsumJobs jobs;
      memset(&jobs, 0, sizeof(jobs));
      for (int e = 0; e < E; ++e) {
        jobs.part[jobs.njobs] = colsum_part(h, e, l);
        jobs.dst[jobs.njobs++] = tb(grads, L, e, l);
        if (ln) {
          jobs.part[jobs.njobs] = w.part_dg[e];
          jobs.dst[jobs.njobs++] = lns(grads, L, e, l);
          jobs.part[jobs.njobs] = w.part_dbeta[e];
          jobs.dst[jobs.njobs++] = lnb(grads, L, e, l);
        }
      }
      MTRL_PROPAGATE(launch_colsum(h, jobs, M / kTileRows, st));
    }
    MTRL_PROPAGATE(run_plan(h, plans[i], st));
  }
  return MTRL_OK;
}

// Trunk backward of one network: per layer, finish the bias gradient from the partial column sums its dZ producer
// left behind (head VJP: per 128-row tile; previous layer's dX GEMM epilogue: per 32 rows), then the dW / dX plan.
int run_trunk_backward(mtrl_sac* h, std::vector<mtrl_gemm_plan_t*>& plans, mtrl_gemm_plan_t* fused, float* grads,
                       const mtrl_net_layout_t& L, int E, bool want_wgrad, cudaStream_t st) {
  const int D = h->cfg.depth;
  if (h->ln_mode) {
    const bool actor = grads == h->buf.actor_grads;
    return run_trunk_backward_ln(h, plans, grads, actor ? h->buf.actor_params : h->buf.critic_params, L, actor ? CH_AO : CH_C, E,
                                 want_wgrad, st);
  }
  const bool tg = h->tg_active && (grads == h->buf.critic_grads ? h->tg_active->fill_critic : h->tg_active->fill_actor);
  // bias gradients of all layers in one finishing kernel once every layer's column partials are in their slots
  auto finish_biases = [&]() {
    ColsumJobs jobs;
    memset(&jobs, 0, sizeof(jobs));
    for (int l = D - 1; l >= 0; --l)
      for (int e = 0; e < E; ++e) {
        jobs.part[jobs.njobs] = colsum_part(h, e, l);
        jobs.dst[jobs.njobs] = tb(grads, L, e, l);
        jobs.groups[jobs.njobs++] = l == D - 1 ? h->cfg.max_rows / kTileRows : h->cfg.max_rows / 32;
      }
    return launch_colsum(h, jobs, 1, st);
  };
  const bool one_finisher = D * E <= kMaxColsumJobs;
  if (fused && !tg && (!want_wgrad || one_finisher)) {
    // every layer's dW / dX in ONE phased launch (per-task gradients read each layer's dZ before the next layer
    // overwrites it: layer by layer below)
    MTRL_PROPAGATE(run_plan(h, fused, st));
    if (want_wgrad) MTRL_PROPAGATE(finish_biases());
    return MTRL_OK;
  }
  for (int l = D - 1, i = 0; l >= 0; --l, ++i) {
    if (want_wgrad && (tg || !one_finisher)) {
      ColsumJobs jobs;
      memset(&jobs, 0, sizeof(jobs));
      jobs.njobs = E;
      for (int e = 0; e < E; ++e) {
        jobs.part[e] = colsum_part(h, e, l);
        jobs.dst[e] = tb(grads, L, e, l);
      }
      const int groups = l == D - 1 ? h->cfg.max_rows / kTileRows : h->cfg.max_rows / 32;
      MTRL_PROPAGATE(launch_colsum(h, jobs, groups, st));
      if (tg) {
        // row t of the (T, P) matrix: bias slice from the task's partial column sums, kernel slice from the per-task
        // dW GEMMs (dZ of this layer is still in place; the plan below consumes it)
        const bool critic = grads == h->buf.critic_grads;
        float* tgm = critic ? h->tg_active->critic_tg : h->tg_active->actor_tg;
        const int R = h->tg_active->rows_per_task;
        dim3 g((h->cfg.width + 255) / 256, h->cfg.num_local_tasks, E);
        task_bias_kernel<<<g, 256, 0, st>>>(jobs, tgm, L.total, L.member_trunk_stride, L.bias_off[l],
                                            l == D - 1 ? R / kTileRows : R / 32, h->cfg.width);
        MTRL_CUDA_CHECK(cudaGetLastError());
        LAUNCHED(h);
        for (auto* p : (critic ? h->tg_active->critic : h->tg_active->actor)[i]) MTRL_PROPAGATE(run_plan(h, p, st));
      }
    }
    MTRL_PROPAGATE(run_plan(h, plans[i], st));
  }
  if (want_wgrad && !tg && one_finisher) MTRL_PROPAGATE(finish_biases());
  return MTRL_OK;
}

int head_sumsq_to_slot(mtrl_sac* h, float* grads, const mtrl_net_layout_t& L, int acc_idx, cudaStream_t st) {
  const long long n = L.total - L.heads_base;
  MTRL_PROPAGATE(prof_begin(h, PT_SUMSQ, st));
  mtrl_launch(sumsq_kernel, dim3(64), dim3(256), 0, st, grads + L.heads_base, n, h->ws.acc + acc_idx);
  mtrl_launch(write_slot_kernel, dim3(1), dim3(1), 0, st, grads + L.slots_off, h->ws.acc + acc_idx);
  MTRL_CUDA_CHECK(cudaGetLastError());
  MTRL_PROPAGATE(prof_end(h, st));
  h->launches += 2;
  return MTRL_OK;
}


// Sharded clip + Adam + Polyak fused with the trunk-gradient exchange over peer memory (comm.cuh).
int launch_trunk_step(mtrl_sac* h, comm::TrunkStepArgs& a, long long off_grads, long long off_params,
                      const comm::Segment* segs, int nsegs, cudaStream_t st) {
  mtrl_comm* c = h->comm;
  a.segs = segs;
  a.nsegs = nsegs;
  a.status = h->ws.status;
  a.rank = c->rank;
  a.world = c->world;
  for (int q = 0; q < c->world; ++q) {
    a.peer_g[q] = reinterpret_cast<float*>(c->peer[q] + off_grads);
    // with a multicast region the peers' parameter copies are not mapped here: one multimem.st reaches them all
    a.peer_p[q] = c->mc_ptr ? (q == c->rank ? a.p : nullptr) : reinterpret_cast<float*>(c->peer[q] + off_params);
    a.peer_hdr[q] = reinterpret_cast<comm::Header*>(c->peer[q]);
  }
  a.mc_p = c->mc_ptr ? reinterpret_cast<float*>(c->mc_ptr + off_params) : nullptr;
  const dim3 grid(h->sms), block(512);
  MTRL_PROPAGATE(prof_begin(h, 1, st));
  switch (c->world) {
    case 2: comm::trunk_step_kernel<2><<<grid, block, 0, st>>>(a); break;
    case 4: comm::trunk_step_kernel<4><<<grid, block, 0, st>>>(a); break;
    case 8: comm::trunk_step_kernel<8><<<grid, block, 0, st>>>(a); break;
    default: comm::trunk_step_kernel<0><<<grid, block, 0, st>>>(a); break;
  }
  MTRL_CUDA_CHECK(cudaGetLastError());
  MTRL_PROPAGATE(prof_end(h, st));
  LAUNCHED(h);
  return MTRL_OK;
}

}  // namespace

extern "C" int mtrl_sac_query_layout(const mtrl_sac_config_t* cfg, mtrl_sac_layout_t* out) {
  MTRL_REQUIRE(cfg && out, "mtrl_sac_query_layout: null argument");
  MTRL_PROPAGATE(validate(*cfg));
  memset(out, 0, sizeof(*out));
  fill_net_layout(&out->actor, cfg->obs_dim, 2 * cfg->action_dim, 1, cfg->num_local_tasks, cfg->width, cfg->depth,
                  cfg->use_layer_norm != 0);
  fill_net_layout(&out->critic, cfg->action_dim + cfg->obs_dim, 1, cfg->num_critics, cfg->num_local_tasks, cfg->width,
                  cfg->depth, cfg->use_layer_norm != 0);
  out->k_actor = static_cast<int>(round_up(cfg->obs_dim, 32));
  out->k_critic = static_cast<int>(round_up(cfg->action_dim + cfg->obs_dim, 32));
  out->workspace_bytes = carve(*cfg, out->k_actor, out->k_critic, out->actor.total, out->critic.total, nullptr, nullptr);
  return MTRL_OK;
}

extern "C" int mtrl_sac_create(mtrl_sac_t** out, const mtrl_sac_config_t* cfg, const mtrl_sac_buffers_t* b) {
  MTRL_REQUIRE(out && cfg && b, "mtrl_sac_create: null argument");
  mtrl_sac* h = new mtrl_sac();
  h->cfg = *cfg;
  h->buf = *b;
  h->ln_mode = cfg->use_layer_norm || cfg->use_skip_connections;
  {
    // MTRL_FUSED_HEADS=0 keeps the stand-alone head kernels (which LayerNorm / skip networks and head widths the GEMM
    // epilogue does not take always use)
    const char* env = getenv("MTRL_FUSED_HEADS");
    const int hd = 2 * cfg->action_dim;
    h->fused_heads = !h->ln_mode && (hd == 2 || hd == 4 || hd == 8) && cfg->width % 32 == 0 && !(env && env[0] == '0');
  }
  int rc = mtrl_sac_query_layout(cfg, &h->lay);
  if (rc != MTRL_OK) { delete h; return rc; }
  const void* need[] = {b->actor_params, b->actor_grads, b->actor_m, b->actor_v, b->actor_shadow, b->critic_params,
                        b->critic_grads, b->critic_m, b->critic_v, b->critic_shadow, b->critic_target,
                        b->critic_target_shadow, b->log_alpha, b->alpha_m, b->alpha_v, b->steps, b->logs, b->workspace};
  for (const void* p : need) {
    if (!p || (reinterpret_cast<uintptr_t>(p) & 15u)) {
      delete h;
      mtrl_set_error("mtrl_sac_create: every buffer must be non-null and 16-byte aligned");
      return MTRL_ERR_INVALID;
    }
  }
  carve(*cfg, h->lay.k_actor, h->lay.k_critic, h->lay.actor.total, h->lay.critic.total, static_cast<uint8_t*>(b->workspace), &h->ws);
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, dev);
  cudaMemset(b->workspace, 0, h->lay.workspace_bytes);
  cudaFuncSetAttribute(pack_plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(actor_head_tile_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(actor_head_tile_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(actor_head_tile_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(actor_head_tile_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(actor_head_tile_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(actor_head_tile_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(actor_head_tile_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(actor_head_tile_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  {
    // Default: on an unsharded handle (same-box A/B: -1 % step time at MT50/W2048, -1...4 % at MT10/W1024).  On a task shard the
    // kernel still walks the WHOLE replicated critic while the rank's GEMM launches are 1/N as long: it shortens the exchange
    // kernels (8 GPUs: 0.337 -> 0.305 ms) but slows the GEMMs it runs under by more (862.7 -> 849.1 updates/s), so there the
    // target update stays inside the trunk step.  MTRL_DEFER_POLYAK=0|1 forces either.
    const char* env = getenv("MTRL_DEFER_POLYAK");
    h->defer_polyak = env ? env[0] != '0' : cfg->num_local_tasks == cfg->num_tasks;
    if (h->defer_polyak) {
      if (cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        h->defer_polyak = false;
      }
    }
  }
  mtrl_pdl_auto(cfg->width <= 1024);   // the plan timings below see the launch mode the updates will use
  rc = build_plans(h);
  if (rc != MTRL_OK) { mtrl_sac_destroy(h); return rc; }
  rc = mtrl_sac_refresh_shadows(h, nullptr);
  if (rc != MTRL_OK) { mtrl_sac_destroy(h); return rc; }
  cudaDeviceSynchronize();
  *out = h;
  return MTRL_OK;
}

extern "C" void mtrl_sac_destroy(mtrl_sac_t* h) {
  if (!h) return;
  for (auto* v : {&h->fwd, &h->fwd_target, &h->bwd_critic, &h->fwd_pi, &h->bwd_pi, &h->bwd_actor})
    for (auto* p : *v) mtrl_gemm_plan_destroy(p);
  for (auto* p : h->fused)
    if (p) mtrl_gemm_plan_destroy(p);
  for (auto& kv : h->act_plans)
    for (auto* p : kv.second) mtrl_gemm_plan_destroy(p);
  for (auto& kv : h->mlp_plans)
    for (auto* p : kv.second) mtrl_gemm_plan_destroy(p);
  for (auto* v : {&h->tgc.critic, &h->tgc.actor})
    for (auto& launches : *v)
      for (auto* p : launches) mtrl_gemm_plan_destroy(p);
  for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
  if (h->side) cudaStreamSynchronize(h->side);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->side) cudaStreamDestroy(h->side);
  if (h->d_segs_critic) cudaFree(h->d_segs_critic);
  if (h->d_segs_actor) cudaFree(h->d_segs_actor);
  delete h;
}

extern "C" int mtrl_sac_refresh_shadows(mtrl_sac_t* h, void* stream) {
  MTRL_REQUIRE(h, "mtrl_sac_refresh_shadows: null handle");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  shadow_kernel<<<h->sms * 4, 256, 0, st>>>(h->buf.actor_params, h->buf.actor_shadow, h->ws.ash_lo, h->lay.actor.total);
  shadow_kernel<<<h->sms * 4, 256, 0, st>>>(h->buf.critic_params, h->buf.critic_shadow, h->ws.csh_lo, h->lay.critic.total);
  shadow_kernel<<<h->sms * 4, 256, 0, st>>>(h->buf.critic_target, h->buf.critic_target_shadow, h->ws.tsh_lo, h->lay.critic.total);
  MTRL_CUDA_CHECK(cudaGetLastError());
  return MTRL_OK;
}

// ---------------------------------------------------------------------------------------------
// Steps of the update.  MT-SAC (mtsac.py:1173-1247) runs them as
//   begin -> critic_grads -> [all-reduce] -> critic_apply -> actor_sample -> actor_grads -> [all-reduce] -> actor_apply -> alpha
// single-task SAC (sac.py:262-383) as
//   begin -> actor_sample -> alpha -> critic_grads (NEW alpha) -> critic_apply -> actor_grads (NEW alpha, NEW critic) -> actor_apply
// ---------------------------------------------------------------------------------------------
namespace {

int step_alpha_prep(mtrl_sac* h, cudaStream_t st) {
  mtrl_launch(alpha_prep_kernel, dim3(1), dim3(256), 0, st, h->buf.log_alpha, h->cfg.num_local_tasks, h->cfg.use_task_weights, h->ws.alpha_val,
                                       h->ws.task_w);
  MTRL_CUDA_CHECK(cudaGetLastError());
  LAUNCHED(h);
  return MTRL_OK;
}

// memsets, alpha values, packing, and the forward passes that only need the OLD actor / critic:
// actor trunk on s' and s, critic trunk on (a_buffer, s).
int step_begin(mtrl_sac* h, const float* obs, const float* actions, const float* next_obs, const float* dones,
               const float* rewards, int batch, int global_batch, const float* eps_c, const float* eps_a, cudaStream_t st) {
  MTRL_REQUIRE(h && obs && actions && next_obs && dones && rewards, "mtrl_sac_update: null batch pointer");
  const mtrl_sac_config_t& c = h->cfg;
  MTRL_REQUIRE(batch >= 1 && batch <= c.max_batch, "mtrl_sac_update: batch %d outside [1, %d]", batch, c.max_batch);
  MTRL_REQUIRE(global_batch >= batch, "mtrl_sac_update: global_batch %d < batch %d", global_batch, batch);
  MTRL_REQUIRE((eps_c == nullptr) == (eps_a == nullptr), "mtrl_sac_update: pass both eps_c and eps_a or neither");
  Workspace& w = h->ws;
  const int M = c.max_rows, D = c.depth, T = c.num_local_tasks;
  mtrl_pdl_auto(c.width <= 1024);   // programmatic dependent launch only where the update is launch-latency bound
  MTRL_PROPAGATE(join_polyak(h, st));   // (an update abandoned between its phases)
  h->launches = 0;
  h->batch = batch;
  h->global_batch = global_batch;
  MTRL_CUDA_CHECK(cudaMemsetAsync(w.acc, 0, ACC_COUNT * sizeof(double), st));
  MTRL_CUDA_CHECK(cudaMemsetAsync(w.status, 0, 16, st));
  MTRL_CUDA_CHECK(cudaMemsetAsync(h->buf.critic_grads, 0, h->lay.critic.total * sizeof(float), st));
  MTRL_CUDA_CHECK(cudaMemsetAsync(h->buf.actor_grads, 0, h->lay.actor.total * sizeof(float), st));
  h->launches += 4;
  if (h->fused_heads) {
    MTRL_CUDA_CHECK(cudaMemsetAsync(w.pre_an, 0, w.pre_floats * sizeof(float), st));
    h->launches += 1;
  }
  MTRL_PROPAGATE(step_alpha_prep(h, st));
  const int nchunks = (batch + 31) / 32;
  const size_t smem = (static_cast<size_t>(nchunks) * T + T + 1) * sizeof(int);
  MTRL_REQUIRE(smem <= 200 * 1024, "mtrl_sac_update: batch %d x %d tasks exceeds the packing kernel's shared memory", batch, T);
  MTRL_PROPAGATE(prof_begin(h, PT_PACK, st));
  mtrl_launch(row_task_kernel, dim3((batch + 7) / 8), dim3(256), 0, st, obs, batch, c.obs_dim, c.num_tasks, c.task_begin, T, w.row_slot, w.status);
  mtrl_launch(pack_plan_kernel, dim3(1), dim3(1024), smem, st, batch, T, M, w.row_slot, w.slot_src, w.tile_task, w.seg_start, w.status);
  h->launches += 2;
  PackArgs a;
  a.obs = obs; a.actions = actions; a.next_obs = next_obs; a.dones = dones; a.rewards = rewards;
  a.eps_c = eps_c; a.eps_a = eps_a;
  a.Xa_next = w.Xa_next; a.Xa = w.Xa; a.Xc_next = w.Xc_next; a.Xc = w.Xc;
  a.rew = w.rew; a.done = w.done; a.peps_c = w.eps_c; a.peps_a = w.eps_a; a.act_data = w.act_data;
  a.slot_src = w.slot_src;
  a.noise_counter = h->buf.steps + 3;
  a.seed = c.noise_seed;
  a.noise_stream = static_cast<unsigned long long>(c.task_begin);
  a.lo_delta = w.lo_delta;
  a.obs_dim = c.obs_dim; a.act_dim = c.action_dim; a.Ka = h->lay.k_actor; a.Kc = h->lay.k_critic;
  mtrl_launch(pack_rows_kernel, M, dim3(128), 0, st, a);
  LAUNCHED(h);
  MTRL_CUDA_CHECK(cudaGetLastError());
  MTRL_PROPAGATE(prof_end(h, st));
  if (h->fused[mtrl_sac::F_FWD]) MTRL_PROPAGATE(run_plan(h, h->fused[mtrl_sac::F_FWD], st));
  for (int l = 0; l < D && !h->fused[mtrl_sac::F_FWD]; ++l) {
    MTRL_PROPAGATE(run_plan(h, h->fwd[l], st));
    if (h->ln_mode) {
      ChainRef ch[2 + kMaxE] = {{CH_AN, 0, h->buf.actor_params, &h->lay.actor}, {CH_AO, 0, h->buf.actor_params, &h->lay.actor}};
      for (int e = 0; e < c.num_critics; ++e) ch[2 + e] = {CH_C, e, h->buf.critic_params, &h->lay.critic};
      MTRL_PROPAGATE(run_junctions_fwd(h, ch, 2 + c.num_critics, l, M, st));
    }
  }
  return MTRL_OK;
}

// Critic loss and gradients (mtsac.py:526-597 / sac.py:267-300): a' ~ pi_old(s'), target critics on (a', s'), Bellman
// target with the CURRENT alpha_val, dL/dQ, head and trunk VJPs; leaves the head-gradient squared norm in the slot.
int step_critic_grads(mtrl_sac* h, cudaStream_t st) {
  const mtrl_sac_config_t& c = h->cfg;
  Workspace& w = h->ws;
  const mtrl_net_layout_t& LC = h->lay.critic;
  const int M = c.max_rows, W = c.width, D = c.depth, E = c.num_critics;
  MTRL_PROPAGATE(launch_actor_head(h, head_in(h, h->split_critic ? CH_AO : CH_AN, 0),
                                   h->fused_heads ? (h->split_critic ? w.pre_ao : w.pre_an) : nullptr, w.eps_c, w.Xc_next, w.logp_next,
                                   false, st));
  if (h->fused[mtrl_sac::F_FWD_TARGET]) MTRL_PROPAGATE(run_plan(h, h->fused[mtrl_sac::F_FWD_TARGET], st));
  for (int l = 0; l < D && !h->fused[mtrl_sac::F_FWD_TARGET]; ++l) {
    MTRL_PROPAGATE(run_plan(h, h->fwd_target[l], st));
    if (h->ln_mode) {
      ChainRef ch[kMaxE];
      for (int e = 0; e < E; ++e) ch[e] = {CH_TG, e, h->buf.critic_target, &h->lay.critic};
      MTRL_PROPAGATE(run_junctions_fwd(h, ch, E, l, M, st));
    }
  }
  {
    CriticLossArgs a;
    memset(&a, 0, sizeof(a));
    for (int e = 0; e < E; ++e) {
      a.target.H[e] = head_in(h, CH_TG, e);
      a.target.w[e] = hk(h->buf.critic_target, LC, e);
      a.target.b[e] = hb(h->buf.critic_target, LC, e);
      a.online.H[e] = head_in(h, CH_C, e);
      a.online.w[e] = hk(h->buf.critic_params, LC, e);
      a.online.b[e] = hb(h->buf.critic_params, LC, e);
      a.target.pre[e] = h->fused_heads ? w.pre_tg[e] : nullptr;
      a.online.pre[e] = h->fused_heads ? w.pre_c[e] : nullptr;
    }
    a.tile_task = w.tile_task; a.slot_src = w.slot_src;
    a.rew = w.rew; a.done = w.done; a.logp_next = w.logp_next; a.alpha_val = w.alpha_val; a.task_w = w.task_w;
    a.dq = w.dq; a.acc = w.acc;
    a.M = M; a.W = W; a.E = E;
    a.h_lo_delta = w.lo_delta;
    a.gamma = c.gamma;
    const float B = static_cast<float>(h->global_batch);
    // MT-SAC: L = mean over (E, B) of (q-y)^2 (mtsac.py:565); SAC: L = 0.5 * sum_e mean_b (q-y)^2 (sac.py:292)
    a.dq_scale = c.variant == MTRL_VARIANT_SAC ? 1.f / B : 2.f / (static_cast<float>(E) * B);
    a.clip = c.clip_q;
    MTRL_PROPAGATE(prof_begin(h, PT_CRITIC_LOSS, st));
    mtrl_launch(critic_loss_kernel, dim3((M + 7) / 8), dim3(256), 0, st, a);
    MTRL_CUDA_CHECK(cudaGetLastError());
    MTRL_PROPAGATE(prof_end(h, st));
    LAUNCHED(h);
  }
  {
    HeadBwdArgs a;
    memset(&a, 0, sizeof(a));
    for (int e = 0; e < E; ++e) {
      a.H[e] = head_in(h, CH_C, e);
      a.dout[e] = w.dq + static_cast<long long>(e) * M;
      a.Wh[e] = hk(h->buf.critic_params, LC, e);
      a.dZ[e] = h->ln_mode ? w.GN[e] : w.G[e][0];
      a.dWh[e] = hk(h->buf.critic_grads, LC, e);
      a.dbh[e] = hb(h->buf.critic_grads, LC, e);
      a.colsum[e] = h->ln_mode ? nullptr : colsum_part(h, e, D - 1);
    }
    a.seg_start = w.seg_start; a.M = M; a.W = W; a.dz_lo_delta = w.lo_delta; a.no_mask = h->ln_mode;
    MTRL_PROPAGATE(launch_head_bwd(h, a, 1, E, st));
  }
  if (h->comm) {
    // the dW epilogues are about to reduce-add into peer gradient buffers: every rank must have zeroed its own
    // (step_begin) first.  One warp; ranks left the previous update together, so this rarely waits.
    MTRL_PROPAGATE(prof_begin(h, 1, st));
    comm::rank_barrier_kernel<<<1, 32, 0, st>>>(h->comm->d_peer_hdr, h->comm->rank, h->comm->world, h->ws.status);
    MTRL_CUDA_CHECK(cudaGetLastError());
    MTRL_PROPAGATE(prof_end(h, st));
    LAUNCHED(h);
  }
  MTRL_PROPAGATE(run_trunk_backward(h, h->bwd_critic, h->fused[mtrl_sac::F_BWD_CRITIC], h->buf.critic_grads, LC, E, true, st));
  MTRL_PROPAGATE(head_sumsq_to_slot(h, h->buf.critic_grads, LC, ACC_CRITIC_HEAD_G2, st));
  return MTRL_OK;
}

// clip_by_global_norm + Adam on the critic, Polyak target update, log scalars (mtsac.py:599-621).
int step_critic_apply(mtrl_sac* h, cudaStream_t st) {
  const mtrl_sac_config_t& c = h->cfg;
  Workspace& w = h->ws;
  const mtrl_net_layout_t& LC = h->lay.critic;
  const float EB = static_cast<float>(c.num_critics) * static_cast<float>(h->global_batch);
  const float loss_scale = c.variant == MTRL_VARIANT_SAC ? 0.5f / static_cast<float>(h->global_batch) : 1.f / EB;
  if (h->comm) {
    comm::TrunkStepArgs t;
    memset(&t, 0, sizeof(t));
    t.p = h->buf.critic_params; t.m = h->buf.critic_m; t.v = h->buf.critic_v; t.shadow = h->buf.critic_shadow;
    t.g = h->buf.critic_grads; t.target = h->buf.critic_target; t.target_shadow = h->buf.critic_target_shadow;
    t.shadow_lo = w.csh_lo; t.target_shadow_lo = w.tsh_lo;
    t.n = LC.total; t.trunk_n = LC.trunk_total;
    t.step = h->buf.steps + 1;
    t.g2_trunk_out = w.acc + ACC_CRITIC_G2; t.p2_trunk = w.acc + ACC_CRITIC_P2_TRUNK; t.p2_head = w.acc + ACC_CRITIC_P2_HEAD;
    t.lr = c.critic_lr; t.b1 = c.adam_b1; t.b2 = c.adam_b2; t.eps = c.adam_eps; t.max_norm = c.critic_max_grad_norm;
    t.tau = c.tau;
    if (h->defer_polyak) t.target = t.target_shadow = t.target_shadow_lo = nullptr;
    MTRL_PROPAGATE(launch_trunk_step(h, t, h->off_critic_grads, h->off_critic_params, h->d_segs_critic, h->nsegs_critic, st));
    if (h->defer_polyak) MTRL_PROPAGATE(fork_polyak(h, st));
    mtrl_launch(finalize_critic_kernel, dim3(1), dim3(1), 0, st, w.acc, h->buf.critic_grads + LC.slots_off + 1, h->buf.steps, h->buf.logs, 1.f / EB,
                                            loss_scale, 0);
    LAUNCHED(h);
    MTRL_CUDA_CHECK(cudaGetLastError());
    return MTRL_OK;
  }
  MTRL_PROPAGATE(prof_begin(h, PT_SUMSQ, st));
  mtrl_launch(sumsq_kernel, dim3(h->sms * 4), dim3(256), 0, st, h->buf.critic_grads, LC.trunk_total, w.acc + ACC_CRITIC_G2);
  MTRL_PROPAGATE(prof_end(h, st));
  LAUNCHED(h);
  AdamArgs a;
  memset(&a, 0, sizeof(a));
  a.p = h->buf.critic_params; a.m = h->buf.critic_m; a.v = h->buf.critic_v; a.shadow = h->buf.critic_shadow;
  a.g = h->buf.critic_grads; a.target = h->buf.critic_target; a.target_shadow = h->buf.critic_target_shadow;
  a.shadow_lo = w.csh_lo; a.target_shadow_lo = w.tsh_lo;
  a.n = LC.total; a.trunk_n = LC.trunk_total;
  a.g2_trunk = w.acc + ACC_CRITIC_G2; a.g2_heads = h->buf.critic_grads + LC.slots_off;
  a.step = h->buf.steps + 1;
  a.p2_trunk = w.acc + ACC_CRITIC_P2_TRUNK; a.p2_head = w.acc + ACC_CRITIC_P2_HEAD; a.p2_old = w.acc + ACC_CRITIC_P2_OLD;
  a.lr = c.critic_lr; a.b1 = c.adam_b1; a.b2 = c.adam_b2; a.eps = c.adam_eps; a.max_norm = c.critic_max_grad_norm;
  a.tau = c.tau;
  if (h->defer_polyak) a.target = a.target_shadow = a.target_shadow_lo = nullptr;
  MTRL_PROPAGATE(prof_begin(h, PT_ADAM, st));
  mtrl_launch(adam_kernel, dim3(h->sms * 4), dim3(256), 0, st, a);
  MTRL_PROPAGATE(prof_end(h, st));
  LAUNCHED(h);
  if (h->defer_polyak) MTRL_PROPAGATE(fork_polyak(h, st));
  mtrl_launch(finalize_critic_kernel, dim3(1), dim3(1), 0, st, w.acc, h->buf.critic_grads + LC.slots_off, h->buf.steps, h->buf.logs, 1.f / EB,
                                          loss_scale, c.variant == MTRL_VARIANT_SAC);
  LAUNCHED(h);
  MTRL_CUDA_CHECK(cudaGetLastError());
  return MTRL_OK;
}

// a ~ pi_old(s), log pi(a|s) and what the tanh-Gaussian VJP needs (mtsac.py:640-642).  With write_x the sampled
// action also becomes the action columns of the critic input (only valid once the critic backward has consumed them).
int step_actor_sample(mtrl_sac* h, bool write_x, cudaStream_t st) {
  Workspace& w = h->ws;
  return launch_actor_head(h, head_in(h, CH_AO, 0), h->fused_heads ? w.pre_ao : nullptr, w.eps_a, write_x ? w.Xc : nullptr, w.logp, true,
                           st);
}

int step_write_actions(mtrl_sac* h, cudaStream_t st) {
  const mtrl_sac_config_t& c = h->cfg;
  write_actions_kernel<<<(c.max_rows * c.action_dim + 255) / 256, 256, 0, st>>>(h->ws.act, h->ws.Xc, h->lay.k_critic,
                                                                                c.max_rows, c.action_dim, h->ws.lo_delta);
  MTRL_CUDA_CHECK(cudaGetLastError());
  LAUNCHED(h);
  return MTRL_OK;
}

// Actor loss and gradients (mtsac.py:631-692): Q of the CURRENT critic params on (a, s), dL/da back through the
// critics (input gradients only), tanh-Gaussian VJP, actor head and trunk VJPs.
int step_actor_grads(mtrl_sac* h, cudaStream_t st) {
  const mtrl_sac_config_t& c = h->cfg;
  Workspace& w = h->ws;
  const mtrl_net_layout_t& LA = h->lay.actor;
  const mtrl_net_layout_t& LC = h->lay.critic;
  const int M = c.max_rows, W = c.width, D = c.depth, E = c.num_critics;
  const float inv_b = 1.f / static_cast<float>(h->global_batch);
  // dL/da accumulates over K splits (see build_plans)
  MTRL_CUDA_CHECK(cudaMemsetAsync(w.dXin, 0, static_cast<size_t>(E) * M * 16 * sizeof(float), st));
  h->launches += 1;
  if (h->fused[mtrl_sac::F_FWD_PI]) MTRL_PROPAGATE(run_plan(h, h->fused[mtrl_sac::F_FWD_PI], st));
  for (int l = 0; l < D && !h->fused[mtrl_sac::F_FWD_PI]; ++l) {
    MTRL_PROPAGATE(run_plan(h, h->fwd_pi[l], st));
    if (h->ln_mode) {
      ChainRef ch[kMaxE];
      for (int e = 0; e < E; ++e) ch[e] = {CH_C, e, h->buf.critic_params, &h->lay.critic};
      MTRL_PROPAGATE(run_junctions_fwd(h, ch, E, l, M, st));
    }
  }
  {
    ActorLossArgs a;
    memset(&a, 0, sizeof(a));
    for (int e = 0; e < E; ++e) {
      a.online.H[e] = head_in(h, CH_C, e);
      a.online.w[e] = hk(h->buf.critic_params, LC, e);
      a.online.b[e] = hb(h->buf.critic_params, LC, e);
      a.online.pre[e] = h->fused_heads ? w.pre_pi[e] : nullptr;
    }
    a.tile_task = w.tile_task; a.slot_src = w.slot_src; a.logp = w.logp; a.alpha_val = w.alpha_val; a.task_w = w.task_w;
    a.dq = w.dq; a.acc = w.acc; a.M = M; a.W = W; a.E = E; a.inv_b = inv_b; a.h_lo_delta = w.lo_delta;
    MTRL_PROPAGATE(prof_begin(h, PT_ACTOR_LOSS, st));
    mtrl_launch(actor_loss_kernel, dim3((M + 7) / 8), dim3(256), 0, st, a);
    MTRL_CUDA_CHECK(cudaGetLastError());
    MTRL_PROPAGATE(prof_end(h, st));
    LAUNCHED(h);
  }
  {
    HeadBwdArgs a;  // critic heads, input gradients only
    memset(&a, 0, sizeof(a));
    for (int e = 0; e < E; ++e) {
      a.H[e] = head_in(h, CH_C, e);
      a.dout[e] = w.dq + static_cast<long long>(e) * M;
      a.Wh[e] = hk(h->buf.critic_params, LC, e);
      a.dZ[e] = h->ln_mode ? w.GN[e] : w.G[e][0];
    }
    a.seg_start = w.seg_start; a.M = M; a.W = W; a.dz_lo_delta = w.lo_delta; a.no_mask = h->ln_mode;
    MTRL_PROPAGATE(launch_head_bwd(h, a, 1, E, st));
  }
  MTRL_PROPAGATE(run_trunk_backward(h, h->bwd_pi, h->fused[mtrl_sac::F_BWD_PI], nullptr, LC, E, false, st));
  {
    ActorDoutArgs a;
    a.dXin = w.dXin; a.act = w.act; a.logstd = w.logstd; a.eps = w.eps_a; a.alpha_val = w.alpha_val; a.task_w = w.task_w;
    a.inrange = w.inrange; a.tile_task = w.tile_task; a.slot_src = w.slot_src; a.dout = w.dout;
    a.act_data = w.act_data; a.explore = h->split_actor ? 1 : 0; a.acc = w.acc;
    a.M = M; a.E = E; a.A = c.action_dim; a.inv_b = inv_b;
    mtrl_launch(actor_dout_kernel, dim3((M + 127) / 128), dim3(128), 0, st, a);
    MTRL_CUDA_CHECK(cudaGetLastError());
    LAUNCHED(h);
  }
  {
    HeadBwdArgs a;
    memset(&a, 0, sizeof(a));
    a.H[0] = head_in(h, CH_AO, 0);
    a.dout[0] = w.dout;
    a.Wh[0] = hk(h->buf.actor_params, LA, 0);
    a.dZ[0] = h->ln_mode ? w.GN[0] : w.G[0][0];
    a.dWh[0] = hk(h->buf.actor_grads, LA, 0);
    a.dbh[0] = hb(h->buf.actor_grads, LA, 0);
    a.colsum[0] = h->ln_mode ? nullptr : colsum_part(h, 0, D - 1);
    a.seg_start = w.seg_start; a.M = M; a.W = W; a.dz_lo_delta = w.lo_delta; a.no_mask = h->ln_mode;
    MTRL_PROPAGATE(launch_head_bwd(h, a, 2 * c.action_dim, 1, st));
  }
  MTRL_PROPAGATE(run_trunk_backward(h, h->bwd_actor, h->fused[mtrl_sac::F_BWD_ACTOR], h->buf.actor_grads, LA, 1, true, st));
  MTRL_PROPAGATE(head_sumsq_to_slot(h, h->buf.actor_grads, LA, ACC_ACTOR_HEAD_G2, st));
  return MTRL_OK;
}

int step_actor_apply_inner(mtrl_sac* h, cudaStream_t st);
// Last optimiser step of every update variant: the deferred Polyak update rejoins the caller's stream here.
int step_actor_apply(mtrl_sac* h, cudaStream_t st) {
  MTRL_PROPAGATE(step_actor_apply_inner(h, st));
  return join_polyak(h, st);
}

int step_actor_apply_inner(mtrl_sac* h, cudaStream_t st) {
  const mtrl_sac_config_t& c = h->cfg;
  Workspace& w = h->ws;
  const mtrl_net_layout_t& LA = h->lay.actor;
  if (h->comm) {
    comm::TrunkStepArgs t;
    memset(&t, 0, sizeof(t));
    t.p = h->buf.actor_params; t.m = h->buf.actor_m; t.v = h->buf.actor_v; t.shadow = h->buf.actor_shadow;
    t.g = h->buf.actor_grads;
    t.shadow_lo = w.ash_lo;
    t.n = LA.total; t.trunk_n = LA.trunk_total;
    t.step = h->buf.steps + 0;
    t.g2_trunk_out = w.acc + ACC_ACTOR_G2; t.p2_trunk = w.acc + ACC_ACTOR_P2_TRUNK; t.p2_head = w.acc + ACC_ACTOR_P2_HEAD;
    t.lr = c.actor_lr; t.b1 = c.adam_b1; t.b2 = c.adam_b2; t.eps = c.adam_eps; t.max_norm = c.actor_max_grad_norm;
    MTRL_PROPAGATE(launch_trunk_step(h, t, h->off_actor_grads, h->off_actor_params, h->d_segs_actor, h->nsegs_actor, st));
    mtrl_launch(finalize_actor_kernel, dim3(1), dim3(1), 0, st, w.acc, h->buf.actor_grads + LA.slots_off + 1, h->buf.steps, h->buf.logs,
                                           1.f / static_cast<float>(h->global_batch), 0, 1.f / static_cast<float>(c.action_dim));
    LAUNCHED(h);
    MTRL_CUDA_CHECK(cudaGetLastError());
    return MTRL_OK;
  }
  MTRL_PROPAGATE(prof_begin(h, PT_SUMSQ, st));
  mtrl_launch(sumsq_kernel, dim3(h->sms * 4), dim3(256), 0, st, h->buf.actor_grads, LA.trunk_total, w.acc + ACC_ACTOR_G2);
  MTRL_PROPAGATE(prof_end(h, st));
  LAUNCHED(h);
  AdamArgs a;
  memset(&a, 0, sizeof(a));
  a.p = h->buf.actor_params; a.m = h->buf.actor_m; a.v = h->buf.actor_v; a.shadow = h->buf.actor_shadow;
  a.g = h->buf.actor_grads; a.target = nullptr; a.target_shadow = nullptr;
  a.shadow_lo = w.ash_lo;
  a.n = LA.total; a.trunk_n = LA.trunk_total;
  a.g2_trunk = w.acc + ACC_ACTOR_G2; a.g2_heads = h->buf.actor_grads + LA.slots_off;
  a.step = h->buf.steps + 0;
  a.p2_trunk = w.acc + ACC_ACTOR_P2_TRUNK; a.p2_head = w.acc + ACC_ACTOR_P2_HEAD; a.p2_old = w.acc + ACC_ACTOR_P2_OLD;
  a.lr = c.actor_lr; a.b1 = c.adam_b1; a.b2 = c.adam_b2; a.eps = c.adam_eps; a.max_norm = c.actor_max_grad_norm; a.tau = 0.f;
  MTRL_PROPAGATE(prof_begin(h, PT_ADAM, st));
  mtrl_launch(adam_kernel, dim3(h->sms * 4), dim3(256), 0, st, a);
  MTRL_PROPAGATE(prof_end(h, st));
  LAUNCHED(h);
  mtrl_launch(finalize_actor_kernel, dim3(1), dim3(1), 0, st, w.acc, h->buf.actor_grads + LA.slots_off, h->buf.steps, h->buf.logs,
                                         1.f / static_cast<float>(h->global_batch), c.variant == MTRL_VARIANT_SAC,
                                         1.f / static_cast<float>(c.action_dim));
  LAUNCHED(h);
  MTRL_CUDA_CHECK(cudaGetLastError());
  return MTRL_OK;
}

// Temperature step (mtsac.py:713-731 / sac.py:308-331) on the log-probs of step_actor_sample.
int step_alpha(mtrl_sac* h, cudaStream_t st) {
  const mtrl_sac_config_t& c = h->cfg;
  Workspace& w = h->ws;
  AlphaArgs al;
  al.log_alpha = h->buf.log_alpha; al.m = h->buf.alpha_m; al.v = h->buf.alpha_v;
  al.logp = w.logp; al.seg_start = w.seg_start; al.slot_src = w.slot_src; al.steps = h->buf.steps; al.logs = h->buf.logs;
  al.T_local = c.num_local_tasks; al.target_entropy = c.target_entropy; al.inv_b = 1.f / static_cast<float>(h->global_batch);
  al.lr = c.alpha_lr; al.b1 = c.adam_b1; al.b2 = c.adam_b2; al.eps = c.adam_eps; al.max_norm = c.alpha_max_grad_norm;
  mtrl_launch(alpha_step_kernel, dim3(1), dim3(1024), c.num_local_tasks * sizeof(float), st, al);
  LAUNCHED(h);
  MTRL_CUDA_CHECK(cudaGetLastError());
  return MTRL_OK;
}

}  // namespace

extern "C" int mtrl_sac_phase1_critic_grads(mtrl_sac_t* h, const float* obs, const float* actions, const float* next_obs,
                                            const float* dones, const float* rewards, int batch, int global_batch,
                                            const float* eps_c, const float* eps_a, void* stream) {
  MTRL_REQUIRE(h, "mtrl_sac_phase1: null handle");
  MTRL_REQUIRE(h->cfg.variant == MTRL_VARIANT_MTSAC, "the phase API is for the multi-task variant");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MTRL_PROPAGATE(step_begin(h, obs, actions, next_obs, dones, rewards, batch, global_batch, eps_c, eps_a, st));
  return step_critic_grads(h, st);
}

extern "C" int mtrl_sac_phase2_critic_step_actor_grads(mtrl_sac_t* h, void* stream) {
  MTRL_REQUIRE(h && h->batch > 0, "mtrl_sac_phase2: phase 1 has not run");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MTRL_PROPAGATE(step_critic_apply(h, st));
  MTRL_PROPAGATE(step_actor_sample(h, true, st));
  return step_actor_grads(h, st);
}

extern "C" int mtrl_sac_phase3_actor_step_alpha(mtrl_sac_t* h, void* stream) {
  MTRL_REQUIRE(h && h->batch > 0, "mtrl_sac_phase3: phase 1 has not run");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MTRL_PROPAGATE(step_actor_apply(h, st));
  return step_alpha(h, st);
}

int update_with_pcgrad(mtrl_sac* h, const float* obs, const float* actions, const float* next_obs, const float* dones,
                       const float* rewards, int batch, const float* eps_c, const float* eps_a, cudaStream_t st);

extern "C" int mtrl_sac_update(mtrl_sac_t* h, const float* obs, const float* actions, const float* next_obs,
                               const float* dones, const float* rewards, int batch, int global_batch, const float* eps_c,
                               const float* eps_a, void* stream) {
  MTRL_REQUIRE(h, "mtrl_sac_update: null handle");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (h->cfg.variant == MTRL_VARIANT_SAC) {
    // sac.py:334-351: sample (a, logp) once; alpha first; critic with the new alpha; actor with new alpha and new critic
    MTRL_PROPAGATE(step_begin(h, obs, actions, next_obs, dones, rewards, batch, global_batch, eps_c, eps_a, st));
    MTRL_PROPAGATE(step_actor_sample(h, false, st));
    MTRL_PROPAGATE(step_alpha(h, st));
    MTRL_PROPAGATE(step_alpha_prep(h, st));
    MTRL_PROPAGATE(step_critic_grads(h, st));
    MTRL_PROPAGATE(step_critic_apply(h, st));
    MTRL_PROPAGATE(step_write_actions(h, st));
    MTRL_PROPAGATE(step_actor_grads(h, st));
    return step_actor_apply(h, st);
  }
  if (h->pcgrad_critic || h->pcgrad_actor) return update_with_pcgrad(h, obs, actions, next_obs, dones, rewards, batch, eps_c, eps_a, st);
  MTRL_PROPAGATE(mtrl_sac_phase1_critic_grads(h, obs, actions, next_obs, dones, rewards, batch, global_batch, eps_c, eps_a, stream));
  MTRL_PROPAGATE(mtrl_sac_phase2_critic_step_actor_grads(h, stream));
  MTRL_PROPAGATE(mtrl_sac_phase3_actor_step_alpha(h, stream));
  return MTRL_OK;
}

extern "C" int mtrl_sac_attach_comm(mtrl_sac_t* h, mtrl_comm_t* c, long long off_critic_grads, long long off_actor_grads,
                                    long long off_critic_params, long long off_actor_params) {
  MTRL_REQUIRE(h && c, "mtrl_sac_attach_comm: null argument");
  MTRL_REQUIRE(h->cfg.variant == MTRL_VARIANT_MTSAC, "mtrl_sac_attach_comm: only the multi-task variant shards");
  MTRL_REQUIRE(c->opened, "mtrl_sac_attach_comm: call mtrl_comm_open_peers first");
  const long long need_c = h->lay.critic.total * 4, need_a = h->lay.actor.total * 4;
  const struct { const char* name; long long off; long long bytes; const float* ptr; } r[4] = {
      {"critic_grads", off_critic_grads, need_c, h->buf.critic_grads}, {"actor_grads", off_actor_grads, need_a, h->buf.actor_grads},
      {"critic_params", off_critic_params, need_c, h->buf.critic_params}, {"actor_params", off_actor_params, need_a, h->buf.actor_params}};
  for (int i = 0; i < 4; ++i) {
    const auto& x = r[i];
    if (i >= 2 && c->mc_local) {   // parameters in the multicast region
      MTRL_REQUIRE(x.off >= 0 && x.off % 128 == 0 && x.off + x.bytes <= c->mc_bytes,
                   "mtrl_sac_attach_comm: %s region [%lld, +%lld) outside the multicast region or misaligned", x.name, x.off, x.bytes);
      MTRL_REQUIRE(reinterpret_cast<const uint8_t*>(x.ptr) == c->mc_local + x.off,
                   "mtrl_sac_attach_comm: the handle's %s buffer is not mc_local + %lld", x.name, x.off);
      continue;
    }
    MTRL_REQUIRE(x.off >= MTRL_COMM_HEADER_BYTES && x.off % 128 == 0 && x.off + x.bytes <= c->arena_bytes,
                 "mtrl_sac_attach_comm: %s region [%lld, +%lld) outside the arena or misaligned", x.name, x.off, x.bytes);
    MTRL_REQUIRE(reinterpret_cast<const uint8_t*>(x.ptr) == c->arena + x.off,
                 "mtrl_sac_attach_comm: the handle's %s buffer is not arena + %lld", x.name, x.off);
  }
  MTRL_REQUIRE(h->cfg.num_critics * (c->world + 1) <= 24,
               "mtrl_sac_attach_comm: %d critics x %d ranks exceed the problems one grouped GEMM launch holds; use the "
               "all-reduce exchange", h->cfg.num_critics, c->world);
  h->comm = c;
  h->off_critic_grads = off_critic_grads;
  h->off_actor_grads = off_actor_grads;
  h->off_critic_params = off_critic_params;
  h->off_actor_params = off_actor_params;
  const std::vector<comm::Segment> sc = trunk_segments(h->lay.critic, c->world), sa = trunk_segments(h->lay.actor, c->world);
  MTRL_CUDA_CHECK(cudaMalloc(&h->d_segs_critic, sc.size() * sizeof(comm::Segment)));
  MTRL_CUDA_CHECK(cudaMalloc(&h->d_segs_actor, sa.size() * sizeof(comm::Segment)));
  MTRL_CUDA_CHECK(cudaMemcpy(h->d_segs_critic, sc.data(), sc.size() * sizeof(comm::Segment), cudaMemcpyHostToDevice));
  MTRL_CUDA_CHECK(cudaMemcpy(h->d_segs_actor, sa.data(), sa.size() * sizeof(comm::Segment), cudaMemcpyHostToDevice));
  h->nsegs_critic = static_cast<int>(sc.size());
  h->nsegs_actor = static_cast<int>(sa.size());
  // the dW problems now target the owners' buffers
  return build_backward_plans(h);
}

// Policy actions for arbitrary observations (ContinuousActionPolicy + TanhMultivariateNormalDiag, mtsac.py:70-84,
// 299-311): a = tanh(mu + sigma eps) (sample) or tanh(mu) (mode).  Uses the update's actor buffers as scratch.
extern "C" int mtrl_sac_act(mtrl_sac_t* h, const float* obs, int n, const float* eps, int deterministic, float* actions_out,
                            void* stream) {
  MTRL_REQUIRE(h && obs && actions_out, "mtrl_sac_act: null argument");
  const mtrl_sac_config_t& c = h->cfg;
  MTRL_REQUIRE(n >= 1 && n <= c.max_rows, "mtrl_sac_act: %d rows outside [1, max_rows = %d]", n, c.max_rows);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Workspace& w = h->ws;
  const mtrl_net_layout_t& LA = h->lay.actor;
  const int W = c.width, D = c.depth, Ka = h->lay.k_actor;
  auto it = h->act_plans.find(n);
  if (it == h->act_plans.end()) {
    std::vector<mtrl_gemm_plan_t*> plans;
    for (int l = 0; l < D; ++l) {
      std::vector<mtrl_gemm_problem_t> p;
      float* X = l == 0 ? w.Xa : dense_in(h, CH_AO, 0, l);
      p.push_back(fwd_problem(X, l == 0 ? Ka : W, l == 0 ? LA.in_dim : W, tk(h->buf.actor_shadow, LA, 0, l),
                              tb(h->buf.actor_params, LA, 0, l), w.Ao[l], n, W, nullptr, lo(h, X), tk_lo(w.ash_lo, LA, 0, l),
                              lo(h, w.Ao[l])));
      MTRL_PROPAGATE(make_plan_plain(plans, p));   // created lazily on the caller's stream: no timing launches here
    }
    it = h->act_plans.emplace(n, plans).first;
  }
  MTRL_CUDA_CHECK(cudaMemsetAsync(w.status, 0, 16, st));
  act_pack_kernel<<<n, 128, 0, st>>>(obs, n, c.obs_dim, Ka, c.num_tasks, c.task_begin, c.num_local_tasks, eps, deterministic,
                                     c.action_dim, c.noise_seed, h->act_calls++, w.Xa, w.slot_src, w.eps_a, w.status, w.lo_delta);
  MTRL_CUDA_CHECK(cudaGetLastError());
  for (int l = 0; l < D; ++l) {
    MTRL_PROPAGATE(mtrl_gemm_plan_run(it->second[l], st));
    if (h->ln_mode) {
      ChainRef ch = {CH_AO, 0, h->buf.actor_params, &h->lay.actor};
      MTRL_PROPAGATE(run_junctions_fwd(h, &ch, 1, l, n, st));
    }
  }
  ActorHeadArgs a;
  memset(&a, 0, sizeof(a));
  a.H = head_in(h, CH_AO, 0);
  a.h_lo_delta = w.lo_delta;
  a.Wh = hk(h->buf.actor_params, LA, 0);
  a.bh = hb(h->buf.actor_params, LA, 0);
  a.row_task = w.slot_src;
  a.eps = w.eps_a;
  a.act = actions_out;
  a.logp = w.logp;
  a.M = n;
  a.W = W;
  a.ls_min = c.log_std_min;
  a.ls_max = c.log_std_max;
  const int wpb = 8;
  dim3 grid((n + wpb - 1) / wpb), block(wpb * 32);
  launch_actor_head_rows(a, c.action_dim, grid, block, st);
  MTRL_CUDA_CHECK(cudaGetLastError());
  return MTRL_OK;
}

// MultiHeadNetwork.__call__ (mtrl/nn/multi_head.py:21-68) of one of the handle's networks on arbitrary rows, alone:
// trunk layers as the tcgen05 GEMMs of the update (same precision mode), then the row's own head.
//   net 0: the actor   -> out (n, 2 A)   = the head outputs ContinuousActionPolicy splits into mean / log_std (networks.py:36-37)
//   net 1: the critics -> out (E, n, 1)  = QValueFunction on concatenate((actions, obs)) (networks.py:55-67), ensemble axis first
//   net 2: the target critics, same shape
// Uses the update's activation buffers as scratch (not re-entrant with mtrl_sac_update); rows of tasks this handle does
// not own set status word 0 and give zeros.
extern "C" int mtrl_mlp_forward(mtrl_sac_t* h, int net, const float* obs, const float* actions, int n, float* out, void* stream) {
  MTRL_REQUIRE(h && obs && out, "mtrl_mlp_forward: null argument");
  MTRL_REQUIRE(net >= 0 && net <= 2, "mtrl_mlp_forward: net %d outside {0 actor, 1 critic, 2 target critic}", net);
  MTRL_REQUIRE(net == 0 || actions, "mtrl_mlp_forward: the critic needs actions");
  const mtrl_sac_config_t& c = h->cfg;
  MTRL_REQUIRE(n >= 1 && n <= c.max_rows, "mtrl_mlp_forward: %d rows outside [1, max_rows = %d]", n, c.max_rows);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Workspace& w = h->ws;
  const bool actor = net == 0;
  const mtrl_net_layout_t& L = actor ? h->lay.actor : h->lay.critic;
  const int W = c.width, D = c.depth, E = actor ? 1 : c.num_critics, K = actor ? h->lay.k_actor : h->lay.k_critic;
  float* X = actor ? w.Xa : (net == 1 ? w.Xc : w.Xc_next);
  float* params = actor ? h->buf.actor_params : (net == 1 ? h->buf.critic_params : h->buf.critic_target);
  float* sh = actor ? h->buf.actor_shadow : (net == 1 ? h->buf.critic_shadow : h->buf.critic_target_shadow);
  float* sh_lo = actor ? w.ash_lo : (net == 1 ? w.csh_lo : w.tsh_lo);
  const int chn = actor ? CH_AO : (net == 1 ? CH_C : CH_TG);
  auto act_buf = [&](int e, int l) { return chain_d(h, chn, e, l); };
  auto key = std::make_pair(net, n);
  auto it = h->mlp_plans.find(key);
  if (it == h->mlp_plans.end()) {
    std::vector<mtrl_gemm_plan_t*> plans;
    for (int l = 0; l < D; ++l) {
      std::vector<mtrl_gemm_problem_t> p;
      for (int e = 0; e < E; ++e) {
        float* in = l == 0 ? X : dense_in(h, chn, e, l);
        p.push_back(fwd_problem(in, l == 0 ? K : W, l == 0 ? L.in_dim : W, tk(sh, L, e, l), tb(params, L, e, l), act_buf(e, l), n, W,
                                nullptr, lo(h, in), tk_lo(sh_lo, L, e, l), lo(h, act_buf(e, l))));
      }
      MTRL_PROPAGATE(make_plan_plain(plans, p));
    }
    it = h->mlp_plans.emplace(key, plans).first;
  }
  MTRL_CUDA_CHECK(cudaMemsetAsync(w.status, 0, 16, st));
  mlp_pack_kernel<<<n, 128, 0, st>>>(obs, actor ? nullptr : actions, c.obs_dim, c.action_dim, K, c.num_tasks, c.task_begin,
                                     c.num_local_tasks, X, w.slot_src, w.status, w.lo_delta);
  MTRL_CUDA_CHECK(cudaGetLastError());
  for (int l = 0; l < D; ++l) {
    MTRL_PROPAGATE(mtrl_gemm_plan_run(it->second[l], st));
    if (h->ln_mode) {
      ChainRef ch[kMaxE];
      for (int e = 0; e < E; ++e) ch[e] = {chn, e, params, &L};
      MTRL_PROPAGATE(run_junctions_fwd(h, ch, E, l, n, st));
    }
  }
  HeadFwdArgs a;
  memset(&a, 0, sizeof(a));
  for (int e = 0; e < E; ++e) {
    a.H[e] = head_in(h, chn, e);
    a.Wh[e] = hk(params, L, e);
    a.bh[e] = hb(params, L, e);
  }
  a.row_task = w.slot_src; a.out = out; a.h_lo_delta = w.lo_delta; a.n = n; a.W = W; a.HD = L.head_dim; a.E = E;
  head_fwd_kernel<<<(n * E + 7) / 8, 256, 0, st>>>(a);
  MTRL_CUDA_CHECK(cudaGetLastError());
  return MTRL_OK;
}

// optax.chain(clip_by_global_norm(max_grad_norm), adam(lr, b1, b2, eps)) + apply_updates (mtrl/config/optim.py:26-43,
// mtrl/rl/algorithms/utils.py:11-46) and optax.incremental_update (mtsac.py:607-613) on caller-owned flat buffers, as one
// stand-alone operator: the two kernels the fused update runs (sumsq_kernel, adam_kernel).
extern "C" int mtrl_adam_polyak_step(float* params, const float* grads, float* m, float* v, float* target, long long n, int* step,
                                     float lr, float b1, float b2, float eps, float max_grad_norm, float tau, double* scratch,
                                     void* stream) {
  MTRL_REQUIRE(params && grads && m && v && step && scratch, "mtrl_adam_polyak_step: null argument");
  MTRL_REQUIRE(n >= 4 && n % 4 == 0, "mtrl_adam_polyak_step: n %lld must be a positive multiple of 4", n);
  for (const void* p : {static_cast<const void*>(params), static_cast<const void*>(grads), static_cast<const void*>(m),
                        static_cast<const void*>(v), static_cast<const void*>(target)})
    MTRL_REQUIRE((reinterpret_cast<uintptr_t>(p) & 15u) == 0, "mtrl_adam_polyak_step: buffers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // scratch: [0] squared gradient norm, [1] |params_new|^2, [2] unused (head part), [3] |params_old|^2, [4] = 0.0f slot
  MTRL_CUDA_CHECK(cudaMemsetAsync(scratch, 0, 5 * sizeof(double), st));
  sumsq_kernel<<<sms * 2, 256, 0, st>>>(grads, n, scratch);
  AdamArgs a;
  memset(&a, 0, sizeof(a));
  a.p = params; a.m = m; a.v = v; a.g = grads; a.target = target;
  a.n = n; a.trunk_n = n;   // no reduction slots in a caller-owned buffer
  a.g2_trunk = scratch; a.g2_heads = reinterpret_cast<const float*>(scratch + 4);
  a.step = step;
  a.p2_trunk = scratch + 1; a.p2_head = scratch + 2; a.p2_old = scratch + 3;
  a.lr = lr; a.b1 = b1; a.b2 = b2; a.eps = eps; a.max_norm = max_grad_norm; a.tau = tau;
  adam_kernel<<<sms * 4, 256, 0, st>>>(a);
  step_inc_kernel<<<1, 1, 0, st>>>(step);
  MTRL_CUDA_CHECK(cudaGetLastError());
  return MTRL_OK;
}

// The fused SAC loss pass alone (mtsac.py:547-566, 659-666): from the last trunk activations of the target / online
// critics (and the heads' parameters) to the loss sums and dL/dQ, for rows already packed in 128-row tiles of one task.
extern "C" int mtrl_sac_losses_fwd_bwd(const mtrl_sac_losses_args_t* a, void* stream) {
  MTRL_REQUIRE(a, "mtrl_sac_losses_fwd_bwd: null argument");
  MTRL_REQUIRE(a->num_critics >= 1 && a->num_critics <= kMaxE, "mtrl_sac_losses_fwd_bwd: num_critics %d outside [1, %d]",
               a->num_critics, kMaxE);
  MTRL_REQUIRE(a->rows >= kTileRows && a->rows % kTileRows == 0 && a->width >= 4 && a->width % 4 == 0,
               "mtrl_sac_losses_fwd_bwd: rows %d must be a multiple of %d and width %d of 4", a->rows, kTileRows, a->width);
  MTRL_REQUIRE(a->tile_task && a->row_valid && a->alpha && a->task_weights && a->dq && a->acc && a->global_batch >= 1,
               "mtrl_sac_losses_fwd_bwd: null pointer or empty batch");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int M = a->rows, E = a->num_critics;
  MTRL_CUDA_CHECK(cudaMemsetAsync(a->acc, 0, 4 * sizeof(double), st));
  // the kernels index their accumulators by ACC_*: give them a base such that ACC_QLOSS / ACC_QSUM / ACC_ACTOR_LOSS land
  // in acc[0], acc[1], acc[2]
  static_assert(ACC_QLOSS == 0 && ACC_QSUM == 1, "accumulator layout");
  if (a->mode == 0) {
    MTRL_REQUIRE(a->rewards && a->dones && a->logp_next, "mtrl_sac_losses_fwd_bwd: the critic loss needs rewards, dones, logp_next");
    CriticLossArgs p;
    memset(&p, 0, sizeof(p));
    for (int e = 0; e < E; ++e) {
      MTRL_REQUIRE(a->H_target[e] && a->H_online[e] && a->w_target[e] && a->w_online[e] && a->b_target[e] && a->b_online[e],
                   "mtrl_sac_losses_fwd_bwd: null member pointer");
      p.target.H[e] = a->H_target[e]; p.target.w[e] = a->w_target[e]; p.target.b[e] = a->b_target[e];
      p.online.H[e] = a->H_online[e]; p.online.w[e] = a->w_online[e]; p.online.b[e] = a->b_online[e];
    }
    p.tile_task = a->tile_task; p.slot_src = a->row_valid; p.rew = a->rewards; p.done = a->dones; p.logp_next = a->logp_next;
    p.alpha_val = a->alpha; p.task_w = a->task_weights; p.dq = a->dq; p.acc = a->acc; p.M = M; p.W = a->width; p.E = E;
    p.gamma = a->gamma; p.dq_scale = 2.f / (static_cast<float>(E) * static_cast<float>(a->global_batch)); p.clip = a->clip_q;
    critic_loss_kernel<<<(M + 7) / 8, 256, 0, st>>>(p);
  } else {
    MTRL_REQUIRE(a->logp, "mtrl_sac_losses_fwd_bwd: the actor loss needs logp");
    ActorLossArgs p;
    memset(&p, 0, sizeof(p));
    for (int e = 0; e < E; ++e) {
      MTRL_REQUIRE(a->H_online[e] && a->w_online[e] && a->b_online[e], "mtrl_sac_losses_fwd_bwd: null member pointer");
      p.online.H[e] = a->H_online[e]; p.online.w[e] = a->w_online[e]; p.online.b[e] = a->b_online[e];
    }
    p.tile_task = a->tile_task; p.slot_src = a->row_valid; p.logp = a->logp; p.alpha_val = a->alpha; p.task_w = a->task_weights;
    p.dq = a->dq; p.acc = a->acc - ACC_ACTOR_LOSS + 2; p.M = M; p.W = a->width; p.E = E;
    p.inv_b = 1.f / static_cast<float>(a->global_batch);
    actor_loss_kernel<<<(M + 7) / 8, 256, 0, st>>>(p);
  }
  MTRL_CUDA_CHECK(cudaGetLastError());
  return MTRL_OK;
}

// The ownership table of one network's trunk for `world` ranks (host only; no device, no handle): rows of
// {begin, end, owner, pre_reduced} in floats.  Used by tests and by hosts that want to gather sharded optimiser state.
extern "C" int mtrl_trunk_segments(const mtrl_net_layout_t* layout, int world, long long* out4, int max_segments, int* n_out) {
  MTRL_REQUIRE(layout && out4 && n_out, "mtrl_trunk_segments: null argument");
  MTRL_REQUIRE(world >= 1 && world <= MTRL_COMM_MAX_RANKS, "mtrl_trunk_segments: world %d outside [1, %d]", world, MTRL_COMM_MAX_RANKS);
  const std::vector<comm::Segment> segs = trunk_segments(*layout, world);
  MTRL_REQUIRE(static_cast<int>(segs.size()) <= max_segments, "mtrl_trunk_segments: %d segments, room for %d",
               static_cast<int>(segs.size()), max_segments);
  for (size_t i = 0; i < segs.size(); ++i) {
    out4[4 * i + 0] = segs[i].begin4 * 4;
    out4[4 * i + 1] = segs[i].end4 * 4;
    out4[4 * i + 2] = segs[i].owner;
    out4[4 * i + 3] = segs[i].pre_reduced;
  }
  *n_out = static_cast<int>(segs.size());
  return MTRL_OK;
}

// 1.0 over the trunk elements whose Adam moments this handle holds live (all of them unless the sharded exchange is
// attached, then the segments it owns), 0.0 elsewhere: lets the host assemble the full optimiser state from the ranks
// (sum over ranks of mask * moments) for a checkpoint.  Synchronises.
extern "C" int mtrl_sac_trunk_owner_mask(mtrl_sac_t* h, int critic, float* mask_dev) {
  MTRL_REQUIRE(h && mask_dev, "mtrl_sac_trunk_owner_mask: null argument");
  const mtrl_net_layout_t& L = critic ? h->lay.critic : h->lay.actor;
  std::vector<float> mask(static_cast<size_t>(L.trunk_total), h->comm ? 0.f : 1.f);
  if (h->comm) {
    for (const comm::Segment& sg : trunk_segments(L, h->comm->world))
      if (sg.owner == h->comm->rank)
        for (long long i = sg.begin4 * 4; i < sg.end4 * 4; ++i) mask[static_cast<size_t>(i)] = 1.f;
  }
  MTRL_CUDA_CHECK(cudaMemcpy(mask_dev, mask.data(), mask.size() * sizeof(float), cudaMemcpyHostToDevice));
  return MTRL_OK;
}

// Grouped per-task dW plans for the (T, P) gradient matrices at critic_tg / actor_tg with R packed rows per task.
int ensure_task_plans(mtrl_sac* h, float* critic_tg, float* actor_tg, int R) {
  const mtrl_sac_config_t& c = h->cfg;
  const int T = c.num_tasks, D = c.depth, E = c.num_critics, W = c.width;
  Workspace& w = h->ws;
  const mtrl_net_layout_t& LA = h->lay.actor;
  const mtrl_net_layout_t& LC = h->lay.critic;
  mtrl_sac::TaskGradCache& tc = h->tgc;
  if (tc.critic_tg != critic_tg || tc.actor_tg != actor_tg || tc.rows_per_task != R) {
    for (auto* v : {&tc.critic, &tc.actor}) {
      for (auto& launches : *v)
        for (auto* p : launches) mtrl_gemm_plan_destroy(p);
      v->clear();
    }
    const int Ka = h->lay.k_actor, Kc = h->lay.k_critic;
    for (int l = D - 1; l >= 0; --l) {
      const int src = (D - 1 - l) & 1;
      std::vector<mtrl_gemm_problem_t> pc, pa;
      for (int e = 0; e < E; ++e)
        for (int t = 0; t < T; ++t) {
          const float* X = (l == 0 ? w.Xc : w.C[e][l - 1]) + static_cast<long long>(t) * R * (l == 0 ? Kc : W);
          float* out = tk(critic_tg + static_cast<long long>(t) * LC.total, LC, e, l);
          const float* dZt = w.G[e][src] + static_cast<long long>(t) * R * W;
          pc.push_back(dw_problem(X, l == 0 ? Kc : W, l == 0 ? LC.in_dim : W, dZt, out, R, W, h->sms, 0, lo(h, X), lo(h, dZt)));
        }
      for (int t = 0; t < T; ++t) {
        const float* X = (l == 0 ? w.Xa : w.Ao[l - 1]) + static_cast<long long>(t) * R * (l == 0 ? Ka : W);
        float* out = tk(actor_tg + static_cast<long long>(t) * LA.total, LA, 0, l);
        const float* dZt = w.G[0][src] + static_cast<long long>(t) * R * W;
        pa.push_back(dw_problem(X, l == 0 ? Ka : W, l == 0 ? LA.in_dim : W, dZt, out, R, W, h->sms, 0, lo(h, X), lo(h, dZt)));
      }
      for (auto* pr : {&pc, &pa}) {
        std::vector<mtrl_gemm_plan_t*> launches;
        for (size_t i0 = 0; i0 < pr->size(); i0 += 24) {
          std::vector<mtrl_gemm_problem_t> chunk(pr->begin() + i0, pr->begin() + std::min(pr->size(), i0 + 24));
          for (auto& q : chunk) { q.k_splits = 1; q.epilogue = MTRL_EPI_STORE; }
          MTRL_PROPAGATE(make_plan_plain(launches, chunk));
        }
        (pr == &pc ? tc.critic : tc.actor).push_back(launches);
      }
    }
    tc.critic_tg = critic_tg;
    tc.actor_tg = actor_tg;
    tc.rows_per_task = R;
  }
  return MTRL_OK;
}

// One of the reference's multi-task gradient transformations (mtrl/optim/{pcgrad,cagrad,gradnorm,dummy}.py) on a (T, P)
// matrix of per-task gradients, in coefficient space: Gram matrix -> T weights -> out = sum_k w_k rows[k].
//   mode 0 pcgrad, 1 cagrad, 2 gradnorm, 3 gradnorm with per-task clipping, 4 dummy (plain mean over tasks)
//   gscale: rows x sqrt(gscale) are the reference-scale per-task gradients (T^2 inside the update, whose rows are the
//           full-batch gradient restricted to a task; 1 for rows that already are per-task-mean gradients)
//   gram (T, T), wts (T), stats (4), tw (T; cagrad's softmax task weights, may be null otherwise): device scratch
int combine_rows(int sms, int mode, const float* tg, long long ld, int T, long long P, float gscale, const int* perm, float* gram,
                 float* wts, float* stats, float* tw, float* out, cudaStream_t st) {
  MTRL_CUDA_CHECK(cudaMemsetAsync(gram, 0, static_cast<size_t>(T) * T * sizeof(float), st));
  pairwise_kernel<0><<<sms * 2, 256, 0, st>>>(tg, ld, T, P, gram, T, 1.f, 0.f, 0.f);
  if (mode == 1) {
    // cagrad(num_tasks) defaults (cagrad.py:20-41): c = 0.5, 21 iterations, lr 25 (< 50 tasks) or 50, momentum 0.5
    cagrad_coeff_kernel<<<1, 32, (static_cast<size_t>(T) * T + 7 * T) * sizeof(double), st>>>(gram, T, T, gscale, 0.5f, 21,
                                                                                          T < 50 ? 25.f : 50.f, 0.5f, wts, stats, tw);
  } else if (mode >= 2) {
    gradnorm_coeff_kernel<<<1, 64, 0, st>>>(gram, T, T, gscale, mode == 3, mode == 4, wts, stats);
  } else {
    pcgrad_coeff_kernel<<<1, 64, 2 * T * T * sizeof(float), st>>>(gram, T, T, gscale, perm, wts, stats);
  }
  weighted_rows_kernel<<<sms * 4, 256, 0, st>>>(tg, ld, T, wts, out, P);
  MTRL_CUDA_CHECK(cudaGetLastError());
  return MTRL_OK;
}

// The transformation in front of one network's optimiser chain: the combined gradient is written over the network's
// gradient buffer (what the rest of the chain, clip + adam, consumes).
int pcgrad_combine(mtrl_sac* h, bool critic, cudaStream_t st) {
  const int T = h->cfg.num_tasks;
  const mtrl_net_layout_t& L = critic ? h->lay.critic : h->lay.actor;
  float* tg = critic ? h->tgc.critic_tg : h->tgc.actor_tg;
  float* grads = critic ? h->buf.critic_grads : h->buf.actor_grads;
  float* gram = h->pcgrad_scratch + (critic ? 0 : T * T);
  float* wts = h->pcgrad_scratch + 2 * T * T + (critic ? 0 : T);
  float* stats = h->pcgrad_scratch + 2 * T * T + 2 * T + (critic ? 0 : 4);
  // cagrad's softmax task weights go where the other network's Gram matrix is not (scratch tail)
  float* tw = h->pcgrad_scratch + 2 * T * T + 2 * T + 8 + (critic ? 0 : T);
  MTRL_PROPAGATE(combine_rows(h->sms, h->surgery_mode, tg, L.total, T, L.total, static_cast<float>(T) * static_cast<float>(T),
                              critic ? h->pcgrad_perm_critic : h->pcgrad_perm_actor, gram, wts, stats, tw, grads, st));
  h->launches += 4;
  // the head-gradient norm that rides in the slot was accumulated from the pre-surgery gradients: start it over
  const int acc_idx = critic ? ACC_CRITIC_HEAD_G2 : ACC_ACTOR_HEAD_G2;
  MTRL_CUDA_CHECK(cudaMemsetAsync(h->ws.acc + acc_idx, 0, sizeof(double), st));
  return head_sumsq_to_slot(h, grads, L, acc_idx, st);
}

// MTSAC.update when an optimiser chain starts with pcgrad (PCGradConfig, mtrl/config/optim.py:62-76): the losses are
// split by task (mtsac.py:568-585, 677-687), the per-task gradients go through pcgrad and its output through clip + adam.
int update_with_pcgrad(mtrl_sac* h, const float* obs, const float* actions, const float* next_obs, const float* dones,
                       const float* rewards, int batch, const float* eps_c, const float* eps_a, cudaStream_t st) {
  const mtrl_sac_config_t& c = h->cfg;
  const int T = c.num_tasks, E = c.num_critics, W = c.width;
  MTRL_REQUIRE(batch % T == 0, "pcgrad: batch %d is not a multiple of the %d tasks (the split losses reshape it)", batch, T);
  const int R = static_cast<int>(round_up(batch / T, kTileRows));
  MTRL_REQUIRE(static_cast<long long>(R) * T <= c.max_rows, "pcgrad: %d rows per task do not fit max_rows", R);
  const mtrl_net_layout_t& LA = h->lay.actor;
  const mtrl_net_layout_t& LC = h->lay.critic;
  MTRL_PROPAGATE(ensure_task_plans(h, h->tgc.critic_tg, h->tgc.actor_tg, R));
  mtrl_sac::TaskGradCache& tc = h->tgc;
  tc.fill_critic = h->pcgrad_critic;
  tc.fill_actor = h->pcgrad_actor;
  MTRL_PROPAGATE(step_begin(h, obs, actions, next_obs, dones, rewards, batch, batch, eps_c, eps_a, st));
  check_balanced_kernel<<<1, 64, 0, st>>>(h->ws.seg_start, h->ws.slot_src, T, R, batch / T, h->ws.status);
  MTRL_CUDA_CHECK(cudaGetLastError());
  if (h->pcgrad_critic) MTRL_CUDA_CHECK(cudaMemsetAsync(tc.critic_tg, 0, static_cast<size_t>(T) * LC.total * sizeof(float), st));
  h->tg_active = &tc;
  h->split_critic = h->pcgrad_critic;   // split_critic_losses = the critic optimiser's requires_split_task_losses (mtsac.py:273)
  int rc = step_critic_grads(h, st);
  h->split_critic = false;
  h->tg_active = nullptr;
  MTRL_PROPAGATE(rc);
  if (h->pcgrad_critic) {
    task_heads_kernel<<<dim3(T, E), 256, 0, st>>>(h->buf.critic_grads, tc.critic_tg, LC.total, LC.heads_base, LC.member_head_stride,
                                                  LC.head_kernel_off, LC.head_bias_off, W, 1);
    MTRL_PROPAGATE(pcgrad_combine(h, true, st));
  }
  MTRL_PROPAGATE(step_critic_apply(h, st));
  MTRL_PROPAGATE(step_actor_sample(h, true, st));
  if (h->pcgrad_actor) MTRL_CUDA_CHECK(cudaMemsetAsync(tc.actor_tg, 0, static_cast<size_t>(T) * LA.total * sizeof(float), st));
  h->tg_active = &tc;
  h->split_actor = h->pcgrad_actor;     // split_actor_losses (mtsac.py:272)
  rc = step_actor_grads(h, st);
  h->split_actor = false;
  h->tg_active = nullptr;
  MTRL_PROPAGATE(rc);
  if (h->pcgrad_actor) {
    task_heads_kernel<<<dim3(T, 1), 256, 0, st>>>(h->buf.actor_grads, tc.actor_tg, LA.total, LA.heads_base, LA.member_head_stride,
                                                  LA.head_kernel_off, LA.head_bias_off, W * 2 * c.action_dim, 2 * c.action_dim);
    MTRL_PROPAGATE(pcgrad_combine(h, false, st));
  }
  MTRL_PROPAGATE(step_actor_apply(h, st));
  return step_alpha(h, st);
}

// Gram matrix of a (T, P) per-task gradient matrix (rows of ld floats): gram[T][T] = rows rows^T in fp32
// (compute_gram_metrics, mtsac.py:733-771; the input of vmap_cos_sim / compute_conflict_metrics, utils.py:49-174).
extern "C" int mtrl_task_gram(const float* rows, long long ld, int T, long long P, float* gram, void* stream) {
  MTRL_REQUIRE(rows && gram, "mtrl_task_gram: null argument");
  MTRL_REQUIRE(T >= 1 && T <= 64 && P >= 1 && ld >= P, "mtrl_task_gram: T %d outside [1, 64] or bad row length", T);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  MTRL_CUDA_CHECK(cudaMemsetAsync(gram, 0, static_cast<size_t>(T) * T * sizeof(float), st));
  pairwise_kernel<0><<<sms * 2, 256, 0, st>>>(rows, ld, T, P, gram, T, 1.f, 0.f, 0.f);
  MTRL_CUDA_CHECK(cudaGetLastError());
  return MTRL_OK;
}

// The element-wise part of compute_conflict_metrics (utils.py:75-101, 146-156) on a (T, P) per-task gradient matrix:
// mismatch[a][b] = #{p : |s rows[a][p]| < eps and |s rows[b][p]| > tau} (compute_sparsity_mismatch before the division by
// the near-zero counts) and row_stats[t] = {sum_p |s rows[t][p]|, #{p : |s rows[t][p]| < eps}} (participation ratio,
// near-zero counts).  `pad` = columns of the flat layout that are alignment padding (always zero): they are removed from
// the near-zero counts here so the caller sees the reference's d = P - pad parameters.
extern "C" int mtrl_task_elementwise(const float* rows, long long ld, int T, long long P, float scale, float eps, float tau,
                                     float* mismatch, double* row_stats, void* stream) {
  MTRL_REQUIRE(rows && mismatch && row_stats, "mtrl_task_elementwise: null argument");
  MTRL_REQUIRE(T >= 1 && T <= 64 && P >= 1 && ld >= P, "mtrl_task_elementwise: T %d outside [1, 64] or bad row length", T);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  MTRL_CUDA_CHECK(cudaMemsetAsync(mismatch, 0, static_cast<size_t>(T) * T * sizeof(float), st));
  MTRL_CUDA_CHECK(cudaMemsetAsync(row_stats, 0, static_cast<size_t>(T) * 2 * sizeof(double), st));
  pairwise_kernel<1><<<sms * 2, 256, 0, st>>>(rows, ld, T, P, mismatch, T, scale, eps, tau);
  row_stats_kernel<<<dim3(sms, T), 256, 0, st>>>(rows, ld, P, scale, eps, row_stats);
  MTRL_CUDA_CHECK(cudaGetLastError());
  return MTRL_OK;
}

// Per-row order statistics of |rows| (radix select on the float bits, 4 one-byte passes): out2[t] = {the rank0[t]-th and
// the (rank0[t] + 1)-th smallest |x| of row t} -- the two neighbours jnp.quantile interpolates between
// (compute_support_metrics, mtsac.py:804-806).  ranks: host long long[T]; scratch: device, T * (32 + 2048) bytes.
extern "C" int mtrl_task_abs_order_stats(const float* rows, long long ld, int T, long long P, const long long* ranks, float* out2,
                                         void* scratch, void* stream) {
  MTRL_REQUIRE(rows && ranks && out2 && scratch, "mtrl_task_abs_order_stats: null argument");
  MTRL_REQUIRE(T >= 1 && T <= 64 && P >= 2 && ld >= P, "mtrl_task_abs_order_stats: T %d outside [1, 64] or bad row length", T);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  unsigned long long* state = static_cast<unsigned long long*>(scratch);                 // [T][2]{prefix, rank}
  unsigned int* hist = reinterpret_cast<unsigned int*>(state + static_cast<size_t>(T) * 4);   // [T][2][256]
  std::vector<unsigned long long> init(static_cast<size_t>(T) * 4, 0ull);
  for (int t = 0; t < T; ++t) {
    MTRL_REQUIRE(ranks[t] >= 0 && ranks[t] + 1 < P, "mtrl_task_abs_order_stats: rank %lld outside [0, P - 2]", ranks[t]);
    init[(t * 2 + 0) * 2 + 1] = static_cast<unsigned long long>(ranks[t]);
    init[(t * 2 + 1) * 2 + 1] = static_cast<unsigned long long>(ranks[t] + 1);
  }
  MTRL_CUDA_CHECK(cudaMemcpyAsync(state, init.data(), init.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
  MTRL_CUDA_CHECK(cudaStreamSynchronize(st));   // `init` is a pageable host vector
  MTRL_CUDA_CHECK(cudaMemsetAsync(hist, 0, static_cast<size_t>(T) * 512 * sizeof(unsigned int), st));
  for (int pass = 0; pass < 4; ++pass) {
    select_hist_kernel<<<dim3(sms, T), 256, 0, st>>>(rows, ld, P, pass, state, hist);
    select_scan_kernel<<<T, 256, 0, st>>>(pass, state, hist);
  }
  order_stats_out_kernel<<<1, 64, 0, st>>>(state, T, out2);
  MTRL_CUDA_CHECK(cudaGetLastError());
  return MTRL_OK;
}

// compute_support_metrics' pairwise counts (mtsac.py:809-835) for supports {|x_t| >= thr[t]}: out3 (3, T, T) fp32 =
// support intersections, sign conflicts, genuine conflicts (sign conflict with both elements in their supports).
extern "C" int mtrl_task_support_pairs(const float* rows, long long ld, int T, long long P, const float* thr, float* out3,
                                       void* stream) {
  MTRL_REQUIRE(rows && thr && out3, "mtrl_task_support_pairs: null argument");
  MTRL_REQUIRE(T >= 1 && T <= 64 && P >= 1 && ld >= P, "mtrl_task_support_pairs: T %d outside [1, 64] or bad row length", T);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  MTRL_CUDA_CHECK(cudaMemsetAsync(out3, 0, static_cast<size_t>(3) * T * T * sizeof(float), st));
  support_pairs_kernel<<<sms * 2, 256, 0, st>>>(rows, ld, T, P, thr, out3);
  MTRL_CUDA_CHECK(cudaGetLastError());
  return MTRL_OK;
}

// Puts pcgrad in front of the critic's and / or the actor's optimiser chain.  critic_tg / actor_tg: device fp32
// (T, layout.total) matrices; scratch: device fp32, 2 T^2 + 2 T + 8 floats (Gram matrices, weights, statistics: per
// network n_grad_conflicts, avg_grad_magnitude, avg_grad_magnitude_before_surgery, norm of the plain mean gradient);
// perm_*: device int[T] row permutations (pcgrad.py:79), rewritten by the caller before every update, or NULL.
extern "C" int mtrl_sac_enable_pcgrad(mtrl_sac_t* h, int critic, int actor, float* critic_tg, float* actor_tg, float* scratch,
                                      const int* perm_critic, const int* perm_actor) {
  MTRL_REQUIRE(h && critic_tg && actor_tg && scratch, "mtrl_sac_enable_pcgrad: null argument");
  MTRL_REQUIRE(h->cfg.variant == MTRL_VARIANT_MTSAC && h->cfg.num_local_tasks == h->cfg.num_tasks && !h->comm,
               "mtrl_sac_enable_pcgrad: needs the multi-task variant with every task on one handle");
  MTRL_REQUIRE(h->cfg.num_tasks <= 64, "mtrl_sac_enable_pcgrad: at most 64 tasks");
  h->pcgrad_critic = critic != 0;
  h->pcgrad_actor = actor != 0;
  if (h->tgc.critic_tg != critic_tg || h->tgc.actor_tg != actor_tg) {
    h->tgc.critic_tg = critic_tg;
    h->tgc.actor_tg = actor_tg;
    h->tgc.rows_per_task = 0;   // plans are rebuilt for the new matrices at the next update
  }
  h->pcgrad_scratch = scratch;
  h->pcgrad_perm_critic = perm_critic;
  h->pcgrad_perm_actor = perm_actor;
  h->surgery_mode = 0;
  return MTRL_OK;
}

// Same wiring with the reference's gradnorm (mtrl/optim/gradnorm.py, GradNormConfig mtrl/config/optim.py:79-102).
extern "C" int mtrl_sac_enable_gradnorm(mtrl_sac_t* h, int critic, int actor, int clip_per_task, float* critic_tg, float* actor_tg,
                                        float* scratch) {
  MTRL_PROPAGATE(mtrl_sac_enable_pcgrad(h, critic, actor, critic_tg, actor_tg, scratch, nullptr, nullptr));
  h->surgery_mode = clip_per_task ? 3 : 2;
  return MTRL_OK;
}

// DummyMultiTaskConfig (mtrl/config/optim.py:46-59): optax.chain(dummy_multitask_optimizer(), clip, adam) -- the plain
// mean of the per-task gradients of the SPLIT losses (which are not the un-split loss, see mtrl_sac::split_critic).
extern "C" int mtrl_sac_enable_dummy(mtrl_sac_t* h, int critic, int actor, float* critic_tg, float* actor_tg, float* scratch) {
  MTRL_PROPAGATE(mtrl_sac_enable_pcgrad(h, critic, actor, critic_tg, actor_tg, scratch, nullptr, nullptr));
  h->surgery_mode = 4;
  return MTRL_OK;
}

// The reference's gradient transformations as a stand-alone operator (the optax protocol objects of mtrl_b200.optim):
// rows (T, ld) = per-task gradients at reference scale; out (P) = the transformed gradient.
extern "C" int mtrl_task_combine(int kind, const float* rows, long long ld, int T, long long P, const int* perm, int clip_per_task,
                                 float* out, float* scratch, void* stream) {
  MTRL_REQUIRE(rows && out && scratch, "mtrl_task_combine: null argument");
  MTRL_REQUIRE(kind >= 0 && kind <= 3, "mtrl_task_combine: kind %d outside 0 (pcgrad) .. 3 (dummy)", kind);
  MTRL_REQUIRE(T >= 1 && T <= 64 && P >= 4 && P % 4 == 0 && ld >= P && ld % 4 == 0,
               "mtrl_task_combine: T %d outside [1, 64] or row length %lld / pitch %lld not a multiple of 4", T, P, ld);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (kind == 1) cudaFuncSetAttribute(cagrad_coeff_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int mode = kind == 0 ? 0 : (kind == 1 ? 1 : (kind == 2 ? (clip_per_task ? 3 : 2) : 4));
  float* gram = scratch;
  float* wts = scratch + T * T;
  float* stats = wts + T;
  float* tw = stats + 4;
  return combine_rows(sms, mode, rows, ld, T, P, 1.f, perm, gram, wts, stats, tw, out, static_cast<cudaStream_t>(stream));
}

// Same wiring with cagrad (mtrl/optim/cagrad.py, CAGradConfig mtrl/config/optim.py:104-124) in front of the chain.
// scratch needs 2 T^2 + 4 T + 8 floats: the pcgrad layout followed by the two networks' softmax task weights.
extern "C" int mtrl_sac_enable_cagrad(mtrl_sac_t* h, int critic, int actor, float* critic_tg, float* actor_tg, float* scratch) {
  MTRL_PROPAGATE(mtrl_sac_enable_pcgrad(h, critic, actor, critic_tg, actor_tg, scratch, nullptr, nullptr));
  h->surgery_mode = 1;
  cudaFuncSetAttribute(cagrad_coeff_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  return MTRL_OK;
}

// Per-task gradients of the critic and actor losses (MTSAC.compute_weights, mtsac.py:870-1170): the batch split by
// task, jax.vmap(jax.value_and_grad(loss)) over the task axis.  Writes row t of critic_tg (T, critic layout.total) and
// actor_tg (T, actor layout.total) in the flat network layout; parameters are NOT updated.  Gradients are those of the
// full-batch losses restricted to task t's rows, i.e. (n_t / B) x the reference's per-task-mean gradients (the caller
// rescales).  Needs all tasks on this handle and the same number of rows for every task.
extern "C" int mtrl_sac_task_grads(mtrl_sac_t* h, const float* obs, const float* actions, const float* next_obs,
                                   const float* dones, const float* rewards, int batch, const float* eps_c,
                                   const float* eps_a, float* critic_tg, float* actor_tg, void* stream) {
  MTRL_REQUIRE(h && critic_tg && actor_tg, "mtrl_sac_task_grads: null argument");
  const mtrl_sac_config_t& c = h->cfg;
  MTRL_REQUIRE(c.variant == MTRL_VARIANT_MTSAC && c.num_local_tasks == c.num_tasks && !h->comm,
               "mtrl_sac_task_grads: needs the multi-task variant with every task on one handle");
  const int T = c.num_tasks, D = c.depth, E = c.num_critics, W = c.width;
  MTRL_REQUIRE(batch % T == 0, "mtrl_sac_task_grads: batch %d is not a multiple of the %d tasks", batch, T);
  const int R = static_cast<int>(round_up(batch / T, kTileRows));
  MTRL_REQUIRE(static_cast<long long>(R) * T <= c.max_rows, "mtrl_sac_task_grads: %d rows per task do not fit max_rows", R);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Workspace& w = h->ws;
  const mtrl_net_layout_t& LA = h->lay.actor;
  const mtrl_net_layout_t& LC = h->lay.critic;
  MTRL_PROPAGATE(ensure_task_plans(h, critic_tg, actor_tg, R));
  mtrl_sac::TaskGradCache& tc = h->tgc;
  tc.fill_critic = tc.fill_actor = true;
  MTRL_CUDA_CHECK(cudaMemsetAsync(critic_tg, 0, static_cast<size_t>(T) * LC.total * sizeof(float), st));
  MTRL_CUDA_CHECK(cudaMemsetAsync(actor_tg, 0, static_cast<size_t>(T) * LA.total * sizeof(float), st));
  MTRL_PROPAGATE(step_begin(h, obs, actions, next_obs, dones, rewards, batch, batch, eps_c, eps_a, st));
  check_balanced_kernel<<<1, 64, 0, st>>>(w.seg_start, w.slot_src, T, R, batch / T, w.status);
  MTRL_CUDA_CHECK(cudaGetLastError());
  h->tg_active = &tc;
  h->split_critic = true;   // compute_weights samples a' on split_data.observations too (mtsac.py:995-1003); no explore term there
  int rc = step_critic_grads(h, st);
  h->split_critic = false;
  if (rc == MTRL_OK) {
    task_heads_kernel<<<dim3(T, E), 256, 0, st>>>(h->buf.critic_grads, critic_tg, LC.total, LC.heads_base, LC.member_head_stride,
                                                  LC.head_kernel_off, LC.head_bias_off, W * 1, 1);
    rc = step_actor_sample(h, true, st);
  }
  if (rc == MTRL_OK) rc = step_actor_grads(h, st);
  if (rc == MTRL_OK)
    task_heads_kernel<<<dim3(T, 1), 256, 0, st>>>(h->buf.actor_grads, actor_tg, LA.total, LA.heads_base, LA.member_head_stride,
                                                  LA.head_kernel_off, LA.head_bias_off, W * 2 * c.action_dim, 2 * c.action_dim);
  h->tg_active = nullptr;
  MTRL_PROPAGATE(rc);
  MTRL_CUDA_CHECK(cudaGetLastError());
  return MTRL_OK;
}

extern "C" int mtrl_sac_launches_per_update(const mtrl_sac_t* h) { return h ? h->launches : 0; }

extern "C" int mtrl_sac_profile_gemms(mtrl_sac_t* h, int enable) {
  MTRL_REQUIRE(h, "mtrl_sac_profile_gemms: null handle");
  h->prof = enable != 0;
  h->ev_used = 0;
  return MTRL_OK;
}

// Sum of the event-bracketed GEMM launch durations since profiling was enabled (synchronises on the events).
extern "C" int mtrl_sac_profile_read(mtrl_sac_t* h, double* total_ms, int* launches) {
  MTRL_REQUIRE(h && total_ms && launches, "mtrl_sac_profile_read: null argument");
  double sum = 0.0, xsum = 0.0, class_ms[PT_COUNT] = {};
  int n = 0, xn = 0, class_n[PT_COUNT] = {};
  for (size_t i = 0; i + 1 < h->ev_used; i += 2) {
    MTRL_CUDA_CHECK(cudaEventSynchronize(h->ev[i + 1]));
    float ms = 0.f;
    MTRL_CUDA_CHECK(cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]));
    const int tag = h->ev_tag[i / 2];
    if (tag >= 0 && tag < PT_COUNT) { class_ms[tag] += ms; ++class_n[tag]; }
    if (tag == PT_GEMM) { sum += ms; ++n; } else if (tag == PT_EXCHANGE) { xsum += ms; ++xn; }
  }
  *total_ms = sum;
  *launches = n;
  h->exchange_ms = xsum;
  h->exchange_launches = xn;
  for (int i = 0; i < PT_COUNT; ++i) { h->class_ms[i] = class_ms[i]; h->class_n[i] = class_n[i]; }
  h->ev_used = 0;
  return MTRL_OK;
}

// Summed duration / count of the exchange kernels (sharded-Adam trunk steps, rank barrier) seen by the last
// mtrl_sac_profile_read.
extern "C" int mtrl_sac_profile_exchange(mtrl_sac_t* h, double* total_ms, int* launches) {
  MTRL_REQUIRE(h && total_ms && launches, "mtrl_sac_profile_exchange: null argument");
  *total_ms = h->exchange_ms;
  *launches = h->exchange_launches;
  return MTRL_OK;
}
// Per kernel class (see include/mtrl_b200.h MTRL_PROF_*): summed event-bracketed duration and launch count as of the last
// mtrl_sac_profile_read.
extern "C" int mtrl_sac_profile_classes(mtrl_sac_t* h, double* ms12, int* n12) {
  MTRL_REQUIRE(h && ms12 && n12, "mtrl_sac_profile_classes: null argument");
  for (int i = 0; i < PT_COUNT; ++i) { ms12[i] = h->class_ms[i]; n12[i] = h->class_n[i]; }
  return MTRL_OK;
}
// Device int[4] written by the packing kernel: [0] != 0 means the last batch was rejected
// (1: a row's task is outside this handle's range, 2: rows do not fit max_rows).
extern "C" int mtrl_sac_read_status_async(const mtrl_sac_t* h, int* host_pinned4, void* stream) {
  MTRL_REQUIRE(h && host_pinned4, "mtrl_sac_read_status_async: null argument");
  MTRL_CUDA_CHECK(cudaMemcpyAsync(host_pinned4, h->ws.status, 16, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
  return MTRL_OK;
}
