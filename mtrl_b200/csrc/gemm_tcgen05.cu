// Grouped TF32 GEMM for sm_100a: tcgen05.mma with TMEM accumulators, TMA-fed 4-stage smem ring,
// persistent warp-specialised CTAs (1 TMA warp, 1 MMA warp, 4 epilogue warps).
//
// One launch ("plan") runs a list of independent problems D = A * B^T-like contractions that
// share tile shape 128 x block_n x 32 (tf32).  These are the dense contractions XLA emits for the
// reference's MultiHeadNetwork trunk (mtrl/nn/multi_head.py:34-44: nn.Dense(width) + activation)
// and for its VJPs under jax.value_and_grad (mtrl/rl/algorithms/mtsac.py:587-596, 689-691):
//   forward   H' = relu(H W + b)          A = H  (K-major),  B = W  [K][N] (MN-major)
//   dX        dH = (dH' W^T) * (H > 0)    A = dH'(K-major),  B = W  [N][K] (K-major)
//   dW        dW = H^T dH'                A = H  [K][M] (MN-major), B = dH' [K][N] (MN-major)
// Grouping lets the two critic-ensemble members, their target copies and the actor share a launch
// (mtrl/rl/networks.py:208-222 stacks the ensemble on a leading axis; same shapes per member).
//
// Operands are fp32 containers already rounded to tf32 by their producers (common.cuh tf32_rna),
// accumulation is fp32 in TMEM, epilogues are fp32.
//
// fp32x3 mode (problem.A_lo / B_lo set): every operand arrives as a (hi, lo) pair of tf32 tensors with x = hi + lo to
// 2^-23, and the contraction is A B + A B_lo + A_lo B (three tensor-core passes; the dropped A_lo B_lo is 2^-22).  The
// tensor core adds into its fp32 accumulator with truncation, a bias that grows linearly with the number of k-steps
// (measured: 7e-9 K relative, profiles/r02_precision_table.md), so the mode does NOT let one accumulator run over K:
// the k-range is cut into chunks of `chunk_kb` k-blocks, every chunk is two sub-units with their own TMEM accumulator --
// MAIN = the (A, B) pass, SMALL = the (A, B_lo) and (A_lo, B) passes, whose sum is 2^-11 of MAIN so its truncation
// error is too -- and the epilogue warps add the finished accumulators into an fp32 RUNNING SUM (round-to-nearest, in
// registers, parked in a third TMEM region at columns [128, 256), hence block_n <= 128).  The last sub-unit's
// epilogue adds the running sum and applies the fused op.  Producer and MMA issuer see sub-units as ordinary pipeline
// stages (only the tensor maps switch).  Epilogues that round their output can also emit the remainder tensor (D_lo)
// so the next contraction gets its pair.
#include <cuda.h>  // CUtensorMap types only; the encode entry point is fetched at run time

#include <stdlib.h>

#include <algorithm>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "mtrl_b200.h"

namespace {

constexpr int kBlockM = 128;         // accumulator rows per CTA (TMEM lanes)
constexpr int kBlockK = 32;          // tf32 elements = 128 bytes = one SWIZZLE_128B row
constexpr int kMaxBlockN = 256;
constexpr int kUmmaK = 8;            // tf32: 32 bytes of K per tcgen05.mma
constexpr int kABytes = kBlockM * kBlockK * 4;        // 16 KB
constexpr int kChunkBytes = 32 * kBlockK * 4;         // one 32(MN) x 32(K) MN-major TMA box
constexpr int kTmemCols = 512;                        // two 256-column fp32 accumulators
constexpr int kEpilogueWarps = 8;                      // two warps per TMEM lane quarter, each takes half the columns
constexpr int kThreads = 64 + 32 * kEpilogueWarps;

// kCtas == 1: one CTA owns a 128 x block_n tile (A 16 KB + B 32 KB per stage, 4 stages).
// kCtas == 2: a CTA pair (cta_group::2) owns a 256 x block_n tile; each CTA stages its own 128 rows of A and
//             HALF of B's N, so a k-block costs 32 KB of L2->smem traffic per SM instead of 48 KB (the 1-CTA
//             kernel is bound by that traffic, profiles/r01_gemm_1cta_ncu_full.txt).
// Epilogue staging per epilogue warp: two 32 x 32 fp32 boxes (SWIZZLE_128B, 4 KB each) for TMA stores, 512 B of bias.
constexpr int kEpiBoxBytes = 32 * 32 * 4;
constexpr int kEpiStageBytes = kEpilogueWarps * 2 * kEpiBoxBytes;   // 64 KB
constexpr int kEpiBiasBytes = kEpilogueWarps * 512;                 // 4 KB
constexpr int kFixSlabFloats = kEpilogueWarps * 4 * 8 * 32 * 4;   // one CTA's parked accumulator (stream-K): 128 x 256 floats
constexpr int kMaxHeadDim = 8;
constexpr int kEpiHeadBytes = kEpilogueWarps * 32 * kMaxHeadDim * 4;   // 8 KB: one box's (32 columns x head_dim) head weights per warp

template <int kCtas>
struct Cfg {
  static constexpr int kBBytes = (kMaxBlockN / kCtas) * kBlockK * 4;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = kCtas == 1 ? 3 : 4;
  static constexpr int kSmemBytes =
      kStages * kStageBytes + kEpiStageBytes + kEpiBiasBytes + kEpiHeadBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct __align__(16) DevProblem {
  float* D;
  const float* bias;
  const float* mask;
  float* colsum;   // optional [ceil(M/32)][N] partial column sums (ReLU-mask epilogue)
  const unsigned* mask_bits;  // ReLU-mask epilogue: bit (n % 32) of word [m][n / 32] set <=> forward activation > 0
  unsigned* bits_out;         // bias+ReLU epilogue: optional, emits those bits
  long long ldbits;
  long long ldd;
  long long ldmask;
  int M, N, K;
  int block_n;
  int m_tiles, n_tiles, k_splits;
  int kb_total, kb_per_split;
  int epilogue;
  int a_major, b_major;
  int unit_begin, unit_count;
  uint32_t idesc;
  int b_chunks;  // 32-wide MN chunks of B each CTA loads per stage (MN-major B only)
  int mn3d;      // bit 0 / 1: A / B is MN-major and described by a 3-D map (one TMA per stage instead of one per chunk)
  int tma_out;   // output goes through smem staging + TMA store / reduce-add (needs block_n % 32 == 0)
  int x3;        // fp32x3: > 0 = k-blocks per chunk (maps 3 / 4 = A_lo / B_lo); every chunk is a MAIN and a SMALL sub-unit
  int d_lo;      // the epilogue also stores the tf32 remainder of the unrounded output through map 5
  int dep_prob;  // row-dependency plans: the problem whose D is this problem's A operand (-1: none)
  int sig_base;  // ... and, when other problems read this problem's D: first of its per-row-tile completion counters (-1: none)
  int head_dim;  // fused output head (bias+ReLU epilogue): 0 = none
  const float* head_w;      // (tasks, N, head_dim)
  float* head_out;          // [M][head_dim], accumulated with atomics
  const int* head_tile;     // task of every 128-row tile
};

// smem matrix descriptor (cute::UMMA::SmemDescriptor layout): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type [61,64): SWIZZLE_128B = 2 (K-major operands),
// SWIZZLE_128B_BASE32B = 1 (the only swizzled layout tcgen05 accepts for MN-major 32-bit operands).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= static_cast<uint64_t>(layout_type & 7u) << 61;
  return d;
}

constexpr int kMaxProblems = MTRL_GEMM_MAX_PROBLEMS;
constexpr int kMapsPer = 6;
constexpr int kMaxPhases = MTRL_GEMM_MAX_PHASES;

// Passed by value as a __grid_constant__ kernel parameter (the usual home of TMA descriptors).
struct GemmParams {
  CUtensorMap maps[kMapsPer * kMaxProblems];   // A, B, D, A_lo, B_lo, D_lo of every problem
  DevProblem probs[kMaxProblems];
  int nprob;
  int total_units;
  long long* dbg;  // optional: 8 cycle counters summed over CTAs (see mtrl_gemm_plan_set_debug)
  // Static schedule built by the host (longest-processing-time-first over the workers, phase by phase): worker w runs,
  // in phase p, units sched[nworkers * nphases + 1 + i] for i in [sched[w * nphases + p], sched[w * nphases + p + 1]).
  const int* sched;
  // Phases: between two phases every CTA meets at a grid-wide barrier (phase_cnt[p], one ticket counter per boundary,
  // monotonic over launches: a launch adds exactly gridDim.x tickets to each) after its TMA stores have completed.
  int nphases;
  unsigned long long* phase_cnt;
  // Stream-K plans (MTRL_GEMM_STREAMK): the schedule's entries index unit_tab, explicit {problem, tile, kb0 | kb1 << 16, fix} units
  // whose k-ranges cut the launch's k-blocks evenly over the workers instead of whole tiles.  A tile cut into `pieces` units whose
  // epilogue is not a plain accumulation (fix >= 0) is finished by whichever unit completes LAST: the others park their raw
  // accumulators in fix_ws and count themselves in fix_cnt; the last one adds them to its own before the fused epilogue.
  // Row-dependency plans (MTRL_GEMM_ROWDEPS): no grid barrier between phases.  A unit of a later phase starts as soon as the
  // row tile of its A operand is complete: every epilogue warp of a producing unit adds one to row_cnt[sig_base + m_tile] once its
  // bulk stores have landed; the TMA producer of a consuming unit waits for (launches so far + 1) x arrivals per launch.  Counters
  // are monotonic over launches (launch_cnt gives the generation), so nothing is reset between launches.
  int rowdeps;
  unsigned long long* row_cnt;
  unsigned long long* launch_cnt;
  const int4* unit_tab;
  const int4* fix_tab;      // [fix] = {pieces, slab offset in floats (lo, hi), 0}
  float* fix_ws;
  unsigned* fix_cnt;        // [fix][cta of the pair][4] = {arrived, parked (x epilogue warps), consumed, -}
};

struct UnitCoord {
  int p, m_tile, n_tile, kb0, kb1, fix;
};

__device__ __forceinline__ UnitCoord decode_unit(const DevProblem* __restrict__ probs, int nprob,
                                                 int unit, const int4* __restrict__ unit_tab = nullptr) {
  if (unit_tab) {
    const int4 e = __ldg(unit_tab + unit);
    UnitCoord c;
    c.p = e.x;
    c.n_tile = e.y % probs[e.x].n_tiles;
    c.m_tile = e.y / probs[e.x].n_tiles;
    c.kb0 = e.z & 0xFFFF;
    c.kb1 = static_cast<int>(static_cast<unsigned>(e.z) >> 16);
    c.fix = e.w;
    return c;
  }
  int p = 0;
  while (p + 1 < nprob && unit >= probs[p + 1].unit_begin) ++p;
  const DevProblem& P = probs[p];
  int u = unit - P.unit_begin;
  UnitCoord c;
  c.p = p;
  c.n_tile = u % P.n_tiles;
  u /= P.n_tiles;
  c.m_tile = u % P.m_tiles;
  int split = u / P.m_tiles;
  c.kb0 = split * P.kb_per_split;
  c.kb1 = min(P.kb_total, c.kb0 + P.kb_per_split);
  c.fix = -1;
  return c;
}

// A unit's k-range [kb0, kb1) runs as `nsub` sub-units, each with its own accumulator: one in tf32 mode; in fp32x3 mode
// two per chunk of P.x3 k-blocks (see the header comment).  Sub-unit `sub` covers k-blocks [s0, s1) in `nst` pipeline
// stages; stage `st` loads k-block kk with operand pass 0 (A, B), 1 (A, B_lo) or 2 (A_lo, B).
struct SubUnit {
  int s0, nst, small;
  __device__ __forceinline__ void stage(int st, int* kk, int* pass) const {
    *kk = s0 + (small ? st >> 1 : st);
    *pass = small ? 1 + (st & 1) : 0;
  }
};
__device__ __forceinline__ int num_subunits(const DevProblem& P, const UnitCoord& c) {
  return P.x3 ? 2 * ((c.kb1 - c.kb0 + P.x3 - 1) / P.x3) : 1;
}
__device__ __forceinline__ SubUnit get_subunit(const DevProblem& P, const UnitCoord& c, int sub) {
  SubUnit u;
  if (!P.x3) {
    u.s0 = c.kb0; u.nst = c.kb1 - c.kb0; u.small = 0;
    return u;
  }
  u.small = sub & 1;
  u.s0 = c.kb0 + (sub >> 1) * P.x3;
  const int s1 = min(c.kb1, u.s0 + P.x3);
  u.nst = (s1 - u.s0) * (u.small ? 2 : 1);
  return u;
}

template <int kCtas>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tf32_grouped_kernel(const __grid_constant__ GemmParams params) {
  using C = Cfg<kCtas>;
  constexpr int kStages = C::kStages;
  const DevProblem* probs = params.probs;
  const CUtensorMap* maps = params.maps;
  const int nprob = params.nprob;
  const int total_units = params.total_units;
  pdl_launch_dependents();   // the next kernel may be launched (it waits for this grid's completion itself)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = smem_base + kStages * C::kStageBytes;   // 1024-aligned staging boxes, then bias
  const uint32_t bar_base = epi_base + kEpiStageBytes + kEpiBiasBytes + kEpiHeadBytes;
  // barrier layout (8 bytes each): full[kStages], empty[kStages], tmem_full[2], tmem_empty[2]
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  const uint32_t phase_bar = bar_base + 8u * (2 * kStages + 5);   // epilogue leader -> producer: the grid passed the phase barrier
  const uint32_t fix_slot = bar_base + 8u * (2 * kStages + 6);    // two 4-byte ticket slots (stream-K fix-up), alternating per unit
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = kCtas == 2 ? cluster_ctarank() : 0u;   // 0 = leader of the pair
  const int worker = blockIdx.x / kCtas;                       // CTA (or pair) index
  const int nworkers = gridDim.x / kCtas;
  const int nphases = params.nphases;
  const int* __restrict__ sched_units = params.sched + nworkers * nphases + 1;
  const int* __restrict__ sched_off = params.sched + worker * nphases;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);   // the (leader's) producer arrive + TMA transaction bytes
      mbar_init(empty_bar(s), 1);  // one tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kEpilogueWarps * kCtas);  // one arrive per epilogue warp of every CTA of the pair
    }
    mbar_init(phase_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    if (kCtas == 2) {
      tmem_alloc_pair(tmem_slot, kTmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (kCtas == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation) touches no memory another kernel
  // writes, so it may run while the previous kernel of the stream drains; from here on its results are needed.
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer (every CTA) =====================
    // The whole warp walks the schedule (warp-uniform values stay in uniform registers); one elected lane issues.
    {
      int stage = 0;
      uint32_t phase = 0;
      long long t_wait = 0, t_issue = 0;
      unsigned long long gen = 0;   // row-dependency plans: launches of this plan before this one
      if (params.rowdeps) {
        if (lane == 0) gen = atomicAdd(params.launch_cnt, 1ull) / gridDim.x;
        gen = __shfl_sync(0xffffffffu, gen, 0);
      }
      for (int ph = 0; ph < nphases; ++ph) {
      if (ph > 0 && !params.rowdeps) {
        // the operands of this phase are outputs of the previous one, written by other CTAs' TMA stores: wait until the
        // whole grid has passed the phase barrier, then order this (async-proxy) reader behind it
        mbar_wait(phase_bar, static_cast<uint32_t>(ph - 1) & 1u);
        fence_proxy_async_all();
      }
      for (int si = sched_off[ph]; si < sched_off[ph + 1]; ++si) {
        const int unit = sched_units[si];
        const UnitCoord c = decode_unit(probs, nprob, unit, params.unit_tab);
        const DevProblem& P = probs[c.p];
        const CUtensorMap* mapA0 = maps + kMapsPer * c.p;
        const int n_cta = P.block_n / kCtas;                       // B columns staged by this CTA
        const int m0 = (c.m_tile * kCtas + static_cast<int>(rank)) * kBlockM;
        const int n0 = c.n_tile * P.block_n + static_cast<int>(rank) * n_cta;
        const uint32_t b_bytes = P.b_major ? static_cast<uint32_t>(P.b_chunks) * kChunkBytes
                                           : static_cast<uint32_t>(n_cta) * kBlockK * 4u;
        const int nsub = num_subunits(P, c);
        if (params.rowdeps && P.dep_prob >= 0) {
          // the A rows of this tile are the D rows of tile row m_tile of the producing problem: all of its column tiles (every
          // epilogue warp of every CTA that stored a piece of them) must have counted themselves in for THIS launch
          const DevProblem& Q = probs[P.dep_prob];
          const unsigned long long want = (gen + 1ull) * static_cast<unsigned long long>(Q.n_tiles * kEpilogueWarps * kCtas);
          const unsigned long long* cnt = params.row_cnt + Q.sig_base + c.m_tile;
          if (lane == 0) {
            unsigned long long seen;
            const long long t_spin = clock64();
            do {
              asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(cnt) : "memory");
            } while (seen < want && clock64() - t_spin < (1ll << 32));   // bounded: a logic error must not hang the GPU
          }
          __syncwarp();
          fence_proxy_async_all();   // order the TMA (async-proxy) reads below behind the acquire
        }
        for (int sub = 0; sub < nsub; ++sub) {
        const SubUnit su = get_subunit(P, c, sub);
        for (int sst = 0; sst < su.nst; ++sst) {
          int kk, pass;
          su.stage(sst, &kk, &pass);
          const long long t0 = params.dbg ? clock64() : 0;
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const long long t1 = params.dbg ? clock64() : 0;
          t_wait += t1 - t0;
          const uint32_t sa = smem_base + stage * C::kStageBytes;
          const uint32_t sb = sa + kABytes;
          const uint32_t fb = full_bar(stage);
          const CUtensorMap* mapA = mapA0 + (pass == 2 ? 3 : 0);
          const CUtensorMap* mapB = mapA0 + (pass == 1 ? 4 : 1);
          const int k0 = kk * kBlockK;
          if (elect_one()) {
          if (rank == 0) mbar_expect_tx(fb, kCtas * (kABytes + b_bytes));
          auto load = [&](uint32_t dst, const CUtensorMap* map, int c0, int c1) {
            if (kCtas == 2) tma_load_2d_pair(dst, map, fb, c0, c1);
            else tma_load_2d(dst, map, fb, c0, c1);
          };
          auto load3 = [&](uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2) {
            if (kCtas == 2) tma_load_3d_pair(dst, map, fb, c0, c1, c2);
            else tma_load_3d(dst, map, fb, c0, c1, c2);
          };
          if (!P.a_major) {
            load(sa, mapA, k0, m0);
          } else if (P.mn3d & 1) {
            load3(sa, mapA, 0, k0, m0 >> 5);
          } else {
#pragma unroll
            for (int j = 0; j < kBlockM / 32; ++j) load(sa + j * kChunkBytes, mapA, m0 + 32 * j, k0);
          }
          if (!P.b_major) {
            load(sb, mapB, k0, n0);
          } else if (P.mn3d & 2) {
            load3(sb, mapB, 0, k0, n0 >> 5);
          } else {
            for (int j = 0; j < P.b_chunks; ++j) load(sb + j * kChunkBytes, mapB, n0 + 32 * j, k0);
          }
          }  // elect_one
          __syncwarp();
          if (params.dbg) t_issue += clock64() - t1;
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        }  // sub-units
      }
      }  // phases
      if (params.dbg && lane == 0) {
        atomicAdd(reinterpret_cast<unsigned long long*>(params.dbg + 0), static_cast<unsigned long long>(t_wait));
        atomicAdd(reinterpret_cast<unsigned long long*>(params.dbg + 1), static_cast<unsigned long long>(t_issue));
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      long long t_wfull = 0, t_wtempty = 0, t_issue = 0, t_total0 = params.dbg ? clock64() : 0;
      for (int si = sched_off[0]; si < sched_off[nphases]; ++si) {   // the issuer is paced by the producer: no phase logic here
        const int unit = sched_units[si];
        const UnitCoord c = decode_unit(probs, nprob, unit, params.unit_tab);
        const DevProblem& P = probs[c.p];
        const uint32_t idesc = P.idesc;
        // K-major (SWIZZLE_128B): rows are 128 B, 8-row swizzle atoms 1024 B apart (SBO);
        //   one MMA consumes 32 B of K, so advance the start address by 32 B.
        // MN-major (SWIZZLE_128B_BASE32B): each TMA box holds 32 MN elements (128 B) x 32 k-rows;
        //   boxes are kChunkBytes apart (LBO), 4-k-row swizzle atoms 512 B apart (SBO);
        //   one MMA consumes 8 k-rows = 1024 B.
        const uint32_t a_lbo = P.a_major ? kChunkBytes : 16u;
        const uint32_t b_lbo = P.b_major ? kChunkBytes : 16u;
        const uint32_t a_sbo = P.a_major ? 512u : 1024u;
        const uint32_t b_sbo = P.b_major ? 512u : 1024u;
        const uint32_t a_lt = P.a_major ? 1u : 2u;
        const uint32_t b_lt = P.b_major ? 1u : 2u;
        const uint32_t a_kstep = P.a_major ? 1024u : 32u;
        const uint32_t b_kstep = P.b_major ? 1024u : 32u;
        const int nsub = num_subunits(P, c);
        for (int sub = 0; sub < nsub; ++sub) {
        const SubUnit su = get_subunit(P, c, sub);
        const long long ta = params.dbg ? clock64() : 0;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        if (params.dbg) t_wtempty += clock64() - ta;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc) * kMaxBlockN;
        for (int sst = 0; sst < su.nst; ++sst) {
          const long long t0 = params.dbg ? clock64() : 0;
          mbar_wait(full_bar(stage), phase);
          const long long t1 = params.dbg ? clock64() : 0;
          t_wfull += t1 - t0;
          tc_fence_after();
          const uint32_t sa = smem_base + stage * C::kStageBytes;
          const uint32_t sb = sa + kABytes;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              const uint64_t adesc = make_smem_desc(sa + k * a_kstep, a_lbo, a_sbo, a_lt);
              const uint64_t bdesc = make_smem_desc(sb + k * b_kstep, b_lbo, b_sbo, b_lt);
              const uint32_t accum = (sst > 0 || k > 0) ? 1u : 0u;
              if (kCtas == 2) umma_tf32_pair(d_tmem, adesc, bdesc, idesc, accum);
              else umma_tf32(d_tmem, adesc, bdesc, idesc, accum);
            }
            // frees the smem slot (in both CTAs of a pair) when these MMAs retire
            if (kCtas == 2) umma_commit_pair(empty_bar(stage), 3); else umma_commit(empty_bar(stage));
          }
          __syncwarp();
          if (params.dbg) t_issue += clock64() - t1;
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        // accumulator complete -> epilogue warps (of both CTAs)
        if (elect_one()) {
          if (kCtas == 2) umma_commit_pair(tfull_bar(acc), 3); else umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
        }  // sub-units
      }
      if (params.dbg && lane == 0) {
        atomicAdd(reinterpret_cast<unsigned long long*>(params.dbg + 2), static_cast<unsigned long long>(t_wfull));
        atomicAdd(reinterpret_cast<unsigned long long*>(params.dbg + 3), static_cast<unsigned long long>(t_wtempty));
        atomicAdd(reinterpret_cast<unsigned long long*>(params.dbg + 4), static_cast<unsigned long long>(t_issue));
        atomicAdd(reinterpret_cast<unsigned long long*>(params.dbg + 5), static_cast<unsigned long long>(clock64() - t_total0));
      }
    }
  } else {
    // ===================== epilogue (warps 2..9 of every CTA) =====================
    // Warp (quarter, col_half) drains TMEM lanes [32*quarter, +32) x its half of the tile's columns in 32-column
    // boxes: TMEM -> registers -> fused op -> swizzled smem box -> one TMA store (or reduce-add) per box, so global
    // memory sees full 128-byte lines instead of 32 scattered 16-byte pieces per instruction.
    const int quarter = warp & 3;
    const int ew = warp - 2;
    const int col_half = ew >> 2;
    const uint32_t stg0 = epi_base + static_cast<uint32_t>(ew) * 2u * kEpiBoxBytes;
    const uint32_t sbias = epi_base + kEpiStageBytes + static_cast<uint32_t>(ew) * 512u;
    const uint32_t shead = epi_base + kEpiStageBytes + kEpiBiasBytes + static_cast<uint32_t>(ew) * (32u * kMaxHeadDim * 4u);
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t nbox = 0;  // running count of TMA boxes this warp has issued (selects the staging buffer)
    uint32_t nfix = 0;  // stream-K fix-up units seen so far (selects the ticket slot)
    long long t_wait = 0, t_work = 0;
    for (int ph = 0; ph < nphases; ++ph) {
    for (int si = sched_off[ph]; si < sched_off[ph + 1]; ++si) {
        const int unit = sched_units[si];
      const UnitCoord c = decode_unit(probs, nprob, unit, params.unit_tab);
      const DevProblem& P = probs[c.p];
      const CUtensorMap* mapD = maps + kMapsPer * c.p + 2;
      const CUtensorMap* mapDlo = maps + kMapsPer * c.p + 5;
      const bool d_lo = P.d_lo != 0;
      const int row0 = (c.m_tile * kCtas + static_cast<int>(rank)) * kBlockM + quarter * 32;
      const int row = row0 + lane;
      const int n0 = c.n_tile * P.block_n;
      const bool row_ok = row < P.M;
      float* drow = P.D + static_cast<long long>(row) * P.ldd;
      const float* mrow = P.mask ? P.mask + static_cast<long long>(row) * P.ldmask : nullptr;
      const int chunks = P.block_n >> 4;
      const int c0 = min(((chunks + 3) >> 2) << 1, chunks);   // 16-column chunks of the first half (even unless tiny)
      const int cbeg = col_half ? c0 * 16 : 0;
      const int cend = col_half ? P.block_n : c0 * 16;
      const int epi = P.epilogue;
      // fused output head: the (32 columns x hd) weights of a box are fetched one box ahead (the first box's before the wait
      // for the accumulator), so their latency never sits between two TMEM loads
      const int hd = row0 < P.M ? P.head_dim : 0;   // (a pair's second half may lie past the last row tile)
      const float* head_wt = nullptr;
      float4 hw_next[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
      auto head_fetch = [&](int cc) {
        const int col0 = n0 + cc;
        const int nval = (cc < cend ? min(32, P.N - col0) : 0) * hd;   // valid floats (columns past N contribute nothing)
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int q = lane + 32 * u;
          hw_next[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (q < 8 * hd && 4 * q < nval) hw_next[u] = __ldg(reinterpret_cast<const float4*>(head_wt + static_cast<long long>(col0) * hd) + q);
        }
      };
      if (hd) {
        head_wt = P.head_w + static_cast<long long>(__ldg(P.head_tile + (row0 >> 7))) * P.N * hd;
        head_fetch(cbeg);
      }
      // fp32x3: all sub-units but the last are added (fp32, round to nearest) into the running sum this warp keeps for
      // its lanes x columns in the third TMEM region, columns [128, 256) (block_n <= 128 in that mode)
      const int nsub = num_subunits(P, c);
      const uint32_t t_sum = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + 128u;
      for (int sub = 0; sub + 1 < nsub; ++sub) {
        const long long ts0 = params.dbg ? clock64() : 0;
        mbar_wait(tfull_bar(acc), acc_phase);
        if (params.dbg) t_wait += clock64() - ts0;
        tc_fence_after();
        const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc) * kMaxBlockN;
        for (int cc = cbeg; cc < cend; cc += 16) {
          uint32_t v[16], sv[16];
          tmem_ld16(t_acc + cc, v);
          if (sub > 0) tmem_ld16(t_sum + cc, sv);
          tmem_ld_wait();
          if (sub > 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(sv[i]));
          }
          tmem_st16(t_sum + cc, v);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (kCtas == 2) mbar_arrive_cluster(tempty_bar(acc), 0); else mbar_arrive(tempty_bar(acc));
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
      const long long t0 = params.dbg ? clock64() : 0;
      mbar_wait(tfull_bar(acc), acc_phase);
      const long long t1 = params.dbg ? clock64() : 0;
      t_wait += t1 - t0;
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                             static_cast<uint32_t>(acc) * kMaxBlockN;
      // Stream-K fix-up (this unit is one of `fix_pieces` k-ranges of its tile and the epilogue is not a plain accumulation):
      // one ticket per CTA once the accumulator is complete; the LAST unit to arrive finishes the tile (fix_role 2), the others
      // park their raw accumulators (fix_role 1).  Nobody waits for a unit that has not arrived yet, so the static schedule
      // cannot deadlock; the finisher only waits for units already in their epilogue.
      int fix_role = 0, fix_pieces = 0;
      float* fix_slab = nullptr;
      unsigned* fix_cnt = nullptr;
      if (c.fix >= 0) {
        const int4 fi = __ldg(params.fix_tab + c.fix);
        fix_pieces = fi.x;
        fix_cnt = params.fix_cnt + (static_cast<long long>(c.fix) * kCtas + rank) * 4;
        const uint32_t slot = fix_slot + 4u * (nfix & 1u);
        if (warp == 2 && lane == 0) {
          const unsigned t = atomicAdd(fix_cnt, 1u);
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(slot), "r"(t) : "memory");
        }
        asm volatile("bar.sync 2, %0;" ::"n"(32 * kEpilogueWarps) : "memory");
        unsigned ticket;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(ticket) : "r"(slot) : "memory");
        ++nfix;
        fix_role = static_cast<int>(ticket) == fix_pieces - 1 ? 2 : 1;
        const long long slab_off = (static_cast<long long>(static_cast<unsigned>(fi.z)) << 32) | static_cast<unsigned>(fi.y);
        // slabs: [piece][cta of the pair][kFixSlabFloats]
        fix_slab = params.fix_ws + slab_off + static_cast<long long>(rank) * kFixSlabFloats;
        if (fix_role == 1) {
          fix_slab += static_cast<long long>(ticket) * kCtas * kFixSlabFloats;
        } else {
          // every earlier arrival has parked all of its eight warps' boxes
          if (lane == 0) {
            const unsigned want = static_cast<unsigned>(fix_pieces - 1) * kEpilogueWarps;
            const long long t_spin = clock64();
            unsigned seen;
            do {
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(fix_cnt + 1) : "memory");
            } while (seen < want && clock64() - t_spin < (1ll << 31));   // bounded: a logic error must not hang the GPU
          }
          __syncwarp();
        }
      }
      if (epi == MTRL_EPI_BIAS_RELU) {
        // this warp's (<= 128) bias values, staged once per tile
        const int col = n0 + cbeg + 4 * lane;
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cbeg + 4 * lane < cend && col + 4 <= P.N) b = __ldg(reinterpret_cast<const float4*>(P.bias + col));
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sbias + 16u * lane), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w)
                     : "memory");
        __syncwarp();
      }
      // ReLU bits of the forward activation for this warp's boxes (one 32-bit word per row and box)
      unsigned mbits[4] = {0u, 0u, 0u, 0u};
      if (epi == MTRL_EPI_RELU_MASK && P.mask_bits && row_ok) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int col = n0 + cbeg + 32 * b;
          if (cbeg + 32 * b < cend && col < P.N) mbits[b] = __ldg(P.mask_bits + static_cast<long long>(row) * P.ldbits + (col >> 5));
        }
      }
      // fused output head: this thread's row x this warp's columns, summed over the boxes of the tile
      float hs[kMaxHeadDim];
#pragma unroll
      for (int j = 0; j < kMaxHeadDim; ++j) hs[j] = 0.f;
      int bi = 0;
      for (int cc = cbeg; cc < cend; cc += 32, ++bi) {
        const int ncols = min(32, cend - cc);
        uint32_t v0[16], v1[16];
        tmem_ld16(t_row + cc, v0);
        if (ncols == 32) tmem_ld16(t_row + cc + 16, v1);
        tmem_ld_wait();
        float o[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          o[i] = __uint_as_float(v0[i]);
          o[16 + i] = ncols == 32 ? __uint_as_float(v1[i]) : 0.f;
        }
        if (nsub > 1) {   // the earlier sub-units of this tile
          tmem_ld16(t_sum + cc, v0);
          if (ncols == 32) tmem_ld16(t_sum + cc + 16, v1);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            o[i] += __uint_as_float(v0[i]);
            if (ncols == 32) o[16 + i] += __uint_as_float(v1[i]);
          }
        }
        if (fix_role) {
          // slab layout [epilogue warp][box][16-byte piece j][lane]: every load / store instruction moves 512 contiguous bytes
          float4* fp = reinterpret_cast<float4*>(fix_slab) + ((ew * 4 + bi) * 8) * 32 + lane;
          if (fix_role == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (4 * j < ncols) __stcg(fp + j * 32, make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]));
            continue;   // nothing else for a parked unit: no fused op, no output, no staging box
          }
          for (int piece = 0; piece + 1 < fix_pieces; ++piece) {
            const float4* sp = fp + static_cast<long long>(piece) * kCtas * (kFixSlabFloats / 4);
            float4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 4 * j < ncols ? __ldcg(sp + j * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              o[4 * j] += v[j].x; o[4 * j + 1] += v[j].y; o[4 * j + 2] += v[j].z; o[4 * j + 3] += v[j].w;
            }
          }
        }
        const int col0 = n0 + cc;
        // d_lo: o[] keeps the UNROUNDED fp32 result through the fused op; the (hi, lo) split happens at the store
        if (epi == MTRL_EPI_BIAS_RELU) {
          unsigned bits = 0u;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 b;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                         : "r"(sbias + static_cast<uint32_t>(cc - cbeg + 4 * j) * 4u));
            o[4 * j + 0] = fmaxf(o[4 * j + 0] + b.x, 0.f);
            o[4 * j + 1] = fmaxf(o[4 * j + 1] + b.y, 0.f);
            o[4 * j + 2] = fmaxf(o[4 * j + 2] + b.z, 0.f);
            o[4 * j + 3] = fmaxf(o[4 * j + 3] + b.w, 0.f);
          }
          if (!d_lo) {
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = tf32_rna(o[i]);
          }
          if (P.bits_out) {
#pragma unroll
            for (int i = 0; i < 32; ++i) bits |= (o[i] > 0.f ? 1u : 0u) << i;
            if (row_ok && col0 < P.N) P.bits_out[static_cast<long long>(row) * P.ldbits + (col0 >> 5)] = bits;
          }
          if (hd) {
            // the box's 32 x hd head weights: staged by the warp (coalesced), read back as broadcast float4s
            __syncwarp();   // the previous box's reads are done
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int q = lane + 32 * u;
              if (q < 8 * hd)
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(shead + 16u * q), "f"(hw_next[u].x), "f"(hw_next[u].y),
                             "f"(hw_next[u].z), "f"(hw_next[u].w)
                             : "memory");
            }
            __syncwarp();
            head_fetch(cc + 32);
            auto head_box = [&](auto hdc) {
              constexpr int HD = decltype(hdc)::value;
#pragma unroll
              for (int q = 0; q < 8 * HD; ++q) {
                float4 w4;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w4.x), "=f"(w4.y), "=f"(w4.z), "=f"(w4.w)
                             : "r"(shead + 16u * q));
                const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                  const int idx = 4 * q + r;   // flat (column, output) index of the box
                  hs[idx % HD] = fmaf(o[idx / HD], wv[r], hs[idx % HD]);
                }
              }
            };
            if (hd == 1) head_box(std::integral_constant<int, 1>());
            else if (hd == 2) head_box(std::integral_constant<int, 2>());
            else if (hd == 4) head_box(std::integral_constant<int, 4>());
            else head_box(std::integral_constant<int, 8>());
          }
        } else if (epi == MTRL_EPI_RELU_MASK) {
          if (P.mask_bits) {
            const unsigned mb = bi == 0 ? mbits[0] : (bi == 1 ? mbits[1] : (bi == 2 ? mbits[2] : mbits[3]));
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = ((mb >> i) & 1u) ? (d_lo ? o[i] : tf32_rna(o[i])) : 0.f;
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int col = col0 + 4 * j;
              float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
              if (row_ok && 4 * j < ncols && col + 4 <= P.N) h = __ldg(reinterpret_cast<const float4*>(mrow + col));
              o[4 * j + 0] = h.x > 0.f ? (d_lo ? o[4 * j + 0] : tf32_rna(o[4 * j + 0])) : 0.f;
              o[4 * j + 1] = h.y > 0.f ? (d_lo ? o[4 * j + 1] : tf32_rna(o[4 * j + 1])) : 0.f;
              o[4 * j + 2] = h.z > 0.f ? (d_lo ? o[4 * j + 2] : tf32_rna(o[4 * j + 2])) : 0.f;
              o[4 * j + 3] = h.w > 0.f ? (d_lo ? o[4 * j + 3] : tf32_rna(o[4 * j + 3])) : 0.f;
            }
          }
          if (P.colsum) {
            // Column sums over this warp's 32 rows (bias-gradient partials): butterfly that halves the values a lane
            // holds each step, 31 shuffles for 32 columns; lane L ends with column L.
            float r[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = (row_ok && col0 + i < P.N) ? o[i] : 0.f;
#pragma unroll
            for (int half = 16, off = 16; half >= 1; half >>= 1, off >>= 1) {
              const bool upper = (lane & off) != 0;
#pragma unroll
              for (int i = 0; i < half; ++i) {
                const float send = upper ? r[i] : r[i + half];
                const float keep = upper ? r[i + half] : r[i];
                r[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
              }
            }
            const int group = row0 >> 5;
            if (lane < ncols && col0 + lane < P.N && row0 < P.M)
              P.colsum[static_cast<long long>(group) * P.N + col0 + lane] = r[0];
          }
        } else if (epi == MTRL_EPI_STORE_TF32) {
          if (!d_lo) {
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = tf32_rna(o[i]);
          }
        }
        if (P.tma_out && d_lo) {
          // (hi, lo) pair: both staging boxes of this warp per iteration, so the previous pair must have been read
          if (nbox >= 2) {
            if (lane == 0) tma_store_wait_read<0>();
            __syncwarp();
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float hi[4], lo[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              hi[q] = tf32_rna(o[4 * j + q]);
              lo[q] = tf32_lo(o[4 * j + q], hi[q]);
            }
            const uint32_t dst = stg0 + static_cast<uint32_t>(lane) * 128u + (static_cast<uint32_t>(j ^ (lane & 7)) << 4);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(hi[0]), "f"(hi[1]), "f"(hi[2]), "f"(hi[3]) : "memory");
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + kEpiBoxBytes), "f"(lo[0]), "f"(lo[1]), "f"(lo[2]), "f"(lo[3])
                         : "memory");
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            // one bulk group per box, in box order: the single-box path's wait_read<1> bookkeeping stays valid when a
            // launch mixes problems with and without D_lo
            tma_store_2d(mapD, stg0, col0, row0);
            tma_store_commit();
            tma_store_2d(mapDlo, stg0 + kEpiBoxBytes, col0, row0);
            tma_store_commit();
          }
          nbox += 2;
        } else if (P.tma_out) {
          const uint32_t stg = stg0 + (nbox & 1u) * kEpiBoxBytes;
          if (nbox >= 2) {
            if (lane == 0) tma_store_wait_read<1>();   // the store issued two boxes ago has released this buffer
            __syncwarp();
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t dst = stg + static_cast<uint32_t>(lane) * 128u + (static_cast<uint32_t>(j ^ (lane & 7)) << 4);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(o[4 * j]), "f"(o[4 * j + 1]), "f"(o[4 * j + 2]),
                         "f"(o[4 * j + 3])
                         : "memory");
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (epi == MTRL_EPI_ATOMIC_ADD) tma_reduce_add_2d(mapD, stg, col0, row0);
            else tma_store_2d(mapD, stg, col0, row0);
            tma_store_commit();
          }
          ++nbox;
        } else if (row_ok) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int col = col0 + 4 * j;
            if (4 * j < ncols && col + 4 <= P.N) {
              const float4 q = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
              if (epi == MTRL_EPI_ATOMIC_ADD) atomicAdd(reinterpret_cast<float4*>(drow + col), q);
              else *reinterpret_cast<float4*>(drow + col) = q;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kCtas == 2) mbar_arrive_cluster(tempty_bar(acc), 0); else mbar_arrive(tempty_bar(acc));
      }
      if (hd && row_ok && cbeg < cend && fix_role != 1) {
#pragma unroll
        for (int j = 0; j < kMaxHeadDim; ++j)
          if (j < hd) atomicAdd(P.head_out + static_cast<long long>(row) * hd + j, hs[j]);
      }
      if (fix_role == 1) {
        __threadfence();   // this warp's parked boxes are visible before it counts itself in
        __syncwarp();
        if (lane == 0) atomicAdd(fix_cnt + 1, 1u);
      } else if (fix_role == 2) {
        __syncwarp();
        if (lane == 0 && atomicAdd(fix_cnt + 2, 1u) == kEpilogueWarps - 1) {
          // the last warp of the finisher re-arms the counters for the next launch of this plan
          fix_cnt[0] = 0u;
          fix_cnt[1] = 0u;
          fix_cnt[2] = 0u;
        }
      }
      if (params.rowdeps && P.sig_base >= 0) {
        // this warp's part of the tile (rows of this CTA x its columns) is in global memory: count it in for the consumers
        if (lane == 0) {
          tma_store_wait<0>();
          __threadfence();
          fence_proxy_async_all();
          atomicAdd(params.row_cnt + P.sig_base + c.m_tile, 1ull);
        }
        __syncwarp();
      }
      if (params.dbg) t_work += clock64() - t1;
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
    if (ph + 1 < nphases && !params.rowdeps) {
      // Phase barrier.  Writer side: this warp's bulk stores have completed, its plain stores (ReLU bits, column-sum
      // partials, atomics) are fenced, all eight epilogue warps of the CTA have done so; then one thread takes a ticket
      // and waits until every CTA of the grid has (tickets are monotonic over launches: generation = ticket / gridDim.x).
      if (lane == 0) tma_store_wait<0>();
      __threadfence();
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpilogueWarps) : "memory");
      if (warp == 2 && lane == 0) {
        fence_proxy_async_all();
        unsigned long long* cnt = params.phase_cnt + ph;
        const unsigned long long ticket = atomicAdd(cnt, 1ull);
        const unsigned long long target = (ticket / gridDim.x + 1ull) * gridDim.x;
        unsigned long long seen;
        const long long t_spin = clock64();
        do {
          asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(cnt) : "memory");
          // every CTA of the grid is resident (grid <= SM count, one CTA per SM), so this cannot wait long; a bound keeps a
          // logic error from hanging the GPU (the results are then wrong and the parity tests say so)
        } while (seen < target && clock64() - t_spin < (1ll << 32));
        fence_proxy_async_all();
        mbar_arrive(phase_bar);   // release the producer warp of this CTA into the next phase
      }
    }
    }  // phases
    if (lane == 0) tma_store_wait<0>();   // all bulk stores of this warp have landed before the CTA retires
    if (params.dbg && lane == 0 && warp == 2) {
      atomicAdd(reinterpret_cast<unsigned long long*>(params.dbg + 6), static_cast<unsigned long long>(t_wait));
      atomicAdd(reinterpret_cast<unsigned long long*>(params.dbg + 7), static_cast<unsigned long long>(t_work));
    }
  }

  tc_fence_before();
  if (kCtas == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (kCtas == 2) tmem_dealloc_pair(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode_fn(EncodeTiledFn* out) {
  static EncodeTiledFn cached = nullptr;
  if (!cached) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    MTRL_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    MTRL_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess,
                 "cuTensorMapEncodeTiled not available from the driver");
    cached = reinterpret_cast<EncodeTiledFn>(fn);
  }
  *out = cached;
  return MTRL_OK;
}

// 2-D fp32 tensor map, zero fill out of bounds.
// inner = contiguous extent (elements), outer = rows, pitch in elements.
int encode_map(CUtensorMap* map, const float* base, long long inner, long long outer, long long pitch,
               int box_inner, int box_outer, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn enc;
  MTRL_PROPAGATE(get_encode_fn(&enc));
  MTRL_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15u) == 0, "GEMM operand base %p not 16-byte aligned",
               (const void*)base);
  MTRL_REQUIRE((pitch * 4) % 16 == 0, "GEMM operand pitch %lld floats is not a multiple of 16 bytes", pitch);
  MTRL_REQUIRE(box_inner == 32 && box_outer >= 1 && box_outer <= 256, "bad TMA box %dx%d", box_inner,
               box_outer);
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(pitch) * 4u};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_outer)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MTRL_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return MTRL_OK;
}

// MN-major operand [K rows][MN] viewed as (32, K, MN/32): one box {32, 32, chunks} fills `chunks` consecutive
// 4 KB [32 k][32 mn] smem blocks, the layout the MN-major descriptors expect.  Needs MN % 32 == 0 (a partial
// last chunk would alias the next row instead of being zero-filled).  Returns false if the driver refuses it.
bool encode_map_mn3d(CUtensorMap* map, const float* base, long long mn, long long k, long long pitch, int chunks) {
  EncodeTiledFn enc;
  if (get_encode_fn(&enc) != MTRL_OK) return false;
  if (mn % 32 != 0 || chunks < 1 || chunks > 8 || (reinterpret_cast<uintptr_t>(base) & 15u) || (pitch * 4) % 16) return false;
  cuuint64_t dims[3] = {32, static_cast<cuuint64_t>(k), static_cast<cuuint64_t>(mn / 32)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(pitch) * 4u, 128u};
  cuuint32_t box[3] = {32, static_cast<cuuint32_t>(kBlockK), static_cast<cuuint32_t>(chunks)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace

struct mtrl_gemm_plan {
  GemmParams params;
  int grid = 0;
  int ctas = 1;  // 1: one CTA per tile; 2: CTA pairs (cta_group::2)
  int* d_sched = nullptr;
  unsigned long long* d_phase_cnt = nullptr;
  int4 *d_unit_tab = nullptr, *d_fix_tab = nullptr;   // stream-K plans only
  float* d_fix_ws = nullptr;
  unsigned* d_fix_cnt = nullptr;
  unsigned long long* d_row_cnt = nullptr;   // row-dependency plans: [0] launches x grid, [1 ..] per row tile arrivals
  bool streamk = false;
  ~mtrl_gemm_plan() {
    if (d_sched) cudaFree(d_sched);
    if (d_phase_cnt) cudaFree(d_phase_cnt);
    if (d_unit_tab) cudaFree(d_unit_tab);
    if (d_fix_tab) cudaFree(d_fix_tab);
    if (d_fix_ws) cudaFree(d_fix_ws);
    if (d_fix_cnt) cudaFree(d_fix_cnt);
    if (d_row_cnt) cudaFree(d_row_cnt);
  }
};

namespace {
// MTRL_GEMM_CTAS=1 forces the single-CTA kernel (debugging / A-B comparison); default is CTA pairs.
int preferred_ctas(int sms) {
  const char* e = getenv("MTRL_GEMM_CTAS");
  if (e && e[0] == '1') return 1;
  return (sms % 2 == 0) ? 2 : 1;
}
}  // namespace

extern "C" int mtrl_gemm_plan_create(mtrl_gemm_plan_t** out, const mtrl_gemm_problem_t* problems, int n) {
  return mtrl_gemm_plan_create_ex(out, problems, n, 0);
}

extern "C" int mtrl_gemm_plan_create_ex(mtrl_gemm_plan_t** out, const mtrl_gemm_problem_t* problems, int n, int ctas_req) {
  MTRL_REQUIRE(out && problems && n >= 1 && n <= kMaxProblems,
               "mtrl_gemm_plan_create: need 1..%d problems per launch, got %d", kMaxProblems, n);
  const bool streamk = (ctas_req & MTRL_GEMM_STREAMK) != 0;
  const bool rowdeps = (ctas_req & MTRL_GEMM_ROWDEPS) != 0;
  ctas_req &= ~(MTRL_GEMM_STREAMK | MTRL_GEMM_ROWDEPS);
  MTRL_REQUIRE(!(streamk && rowdeps), "mtrl_gemm_plan_create_ex: stream-K and row dependencies cannot be combined");
  MTRL_REQUIRE(ctas_req >= 0 && ctas_req <= 2, "mtrl_gemm_plan_create_ex: ctas %d outside {0, 1, 2}", ctas_req);
  int dev_id = 0, sms = 148;
  cudaGetDevice(&dev_id);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_id);
  const int ctas = (ctas_req == 2 && sms % 2) ? 1 : (ctas_req ? ctas_req : preferred_ctas(sms));
  mtrl_gemm_plan* plan = new mtrl_gemm_plan();
  struct Guard {
    mtrl_gemm_plan* p;
    ~Guard() { delete p; }
  } guard{plan};
  plan->ctas = ctas;
  // fp32x3: k-blocks (of 32) one accumulator runs over before it is folded into the running sum (MTRL_X3_CHUNK, 1..64)
  int x3_chunk = 4;
  if (const char* e = getenv("MTRL_X3_CHUNK")) x3_chunk = atoi(e) < 1 ? 1 : (atoi(e) > 64 ? 64 : atoi(e));
  GemmParams& P = plan->params;
  memset(&P, 0, sizeof(P));
  int units = 0;
  bool any_x3 = false;
  for (int i = 0; i < n; ++i) any_x3 = any_x3 || problems[i].A_lo != nullptr;
  for (int i = 0; i < n; ++i) {
    const mtrl_gemm_problem_t& p = problems[i];
    MTRL_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, "problem %d: empty shape", i);
    MTRL_REQUIRE(p.N % 4 == 0, "problem %d: N=%d must be a multiple of 4", i, p.N);
    MTRL_REQUIRE(p.block_n >= 16 && p.block_n <= kMaxBlockN && p.block_n % 16 == 0,
                 "problem %d: block_n=%d must be a multiple of 16 in [16,256]", i, p.block_n);
    MTRL_REQUIRE(p.k_splits >= 1, "problem %d: k_splits must be >= 1", i);
    MTRL_REQUIRE(p.k_splits == 1 || p.epilogue == MTRL_EPI_ATOMIC_ADD,
                 "problem %d: split-K needs the atomic-add epilogue", i);
    MTRL_REQUIRE(p.epilogue != MTRL_EPI_BIAS_RELU || p.bias, "problem %d: bias epilogue without bias", i);
    MTRL_REQUIRE(p.epilogue != MTRL_EPI_RELU_MASK || p.mask || p.mask_bits, "problem %d: mask epilogue without mask", i);
    MTRL_REQUIRE((reinterpret_cast<uintptr_t>(p.D) & 15u) == 0 && p.ldd % 4 == 0,
                 "problem %d: D must be 16-byte aligned with ldd %% 4 == 0", i);
    // A CTA of a pair stages block_n / 2 columns of B; MN-major B arrives in 32-column TMA boxes, so the
    // half must be a multiple of 32 there.
    int block_n = p.block_n;
    MTRL_REQUIRE((p.A_lo == nullptr) == (p.B_lo == nullptr), "problem %d: fp32x3 needs both A_lo and B_lo", i);
    // columns [128, 256) of TMEM hold the running sum of fp32x3 tiles; the MMA issuer runs up to two accumulators ahead of
    // the epilogue, so a wider tile of ANY problem of the launch could overwrite a sum that is still needed
    if (any_x3 && block_n > 128) block_n = 128;
    if (ctas == 2 && p.b_major && block_n % 64 != 0) block_n = (block_n + 63) / 64 * 64;
    const int tile_m = kBlockM * ctas;
    const int n_cta = block_n / ctas;
    DevProblem& d = P.probs[i];
    d.D = p.D;
    d.bias = p.bias;
    d.mask = p.mask;
    d.colsum = p.epilogue == MTRL_EPI_RELU_MASK ? p.colsum_partial : nullptr;
    d.mask_bits = p.epilogue == MTRL_EPI_RELU_MASK ? p.mask_bits : nullptr;
    d.bits_out = p.epilogue == MTRL_EPI_BIAS_RELU ? p.relu_bits_out : nullptr;
    d.ldbits = p.ldbits;
    d.ldd = p.ldd;
    d.ldmask = p.ldmask;
    d.M = p.M;
    d.N = p.N;
    d.K = p.K;
    d.block_n = block_n;
    d.m_tiles = (p.M + tile_m - 1) / tile_m;
    d.n_tiles = (p.N + block_n - 1) / block_n;
    d.x3 = p.A_lo ? x3_chunk : 0;
    d.kb_total = (p.K + kBlockK - 1) / kBlockK;
    int splits = p.k_splits < d.kb_total ? p.k_splits : d.kb_total;
    d.kb_per_split = (d.kb_total + splits - 1) / splits;
    d.k_splits = (d.kb_total + d.kb_per_split - 1) / d.kb_per_split;
    d.epilogue = p.epilogue;
    d.a_major = p.a_major ? 1 : 0;
    d.b_major = p.b_major ? 1 : 0;
    d.b_chunks = (n_cta + 31) / 32;
    d.unit_begin = units;
    d.unit_count = d.m_tiles * d.n_tiles * d.k_splits;
    units += d.unit_count;
    // cute::UMMA::InstrDescriptor: c_format F32 (1) [4,6), a/b_format TF32 (2) [7,10)/[10,13),
    // a_major [15], b_major [16], N>>3 [17,23), M>>4 [24,29)  (M = 256 for a CTA pair).
    d.idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(d.a_major) << 15) |
              (static_cast<uint32_t>(d.b_major) << 16) | (static_cast<uint32_t>(block_n >> 3) << 17) |
              (static_cast<uint32_t>(tile_m >> 4) << 24);
    // A: K-major -> [M][K] rows of K; MN-major -> [K][M] rows of M.  Each CTA loads its own 128 rows.
    MTRL_REQUIRE(!(d.mask_bits || d.bits_out) || (block_n % 32 == 0 && p.ldbits * 32 >= p.N),
                 "problem %d: ReLU bit masks need block_n %% 32 == 0 and ldbits >= N / 32", i);
    // output through TMA: 32 x 32 fp32 boxes of D [M][N] (rows of ldd floats), SWIZZLE_128B staging
    const bool allow_tma_out = !(getenv("MTRL_GEMM_NO_TMA_OUT") && getenv("MTRL_GEMM_NO_TMA_OUT")[0] == '1');
    d.tma_out = 0;
    if (allow_tma_out && block_n % 32 == 0 &&
        encode_map(&P.maps[kMapsPer * i + 2], p.D, p.N, p.M, p.ldd, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B) == MTRL_OK)
      d.tma_out = 1;
    d.dep_prob = -1;
    d.sig_base = -1;
    d.head_dim = 0;
    d.head_w = nullptr;
    d.head_out = nullptr;
    d.head_tile = nullptr;
    if (p.head_w) {
      MTRL_REQUIRE(p.epilogue == MTRL_EPI_BIAS_RELU, "problem %d: the fused head needs the bias + ReLU epilogue", i);
      MTRL_REQUIRE(p.head_out && p.head_tile_task, "problem %d: head_w without head_out / head_tile_task", i);
      MTRL_REQUIRE(p.head_dim == 1 || p.head_dim == 2 || p.head_dim == 4 || p.head_dim == 8,
                   "problem %d: fused head_dim %d not in {1, 2, 4, 8}", i, p.head_dim);
      MTRL_REQUIRE(block_n % 32 == 0 && (reinterpret_cast<uintptr_t>(p.head_w) & 15u) == 0,
                   "problem %d: the fused head needs block_n %% 32 == 0 and a 16-byte aligned head_w", i);
      d.head_dim = p.head_dim;
      d.head_w = p.head_w;
      d.head_out = p.head_out;
      d.head_tile = p.head_tile_task;
    }
    d.d_lo = 0;
    if (p.D_lo) {
      MTRL_REQUIRE(p.epilogue == MTRL_EPI_BIAS_RELU || p.epilogue == MTRL_EPI_RELU_MASK || p.epilogue == MTRL_EPI_STORE_TF32,
                   "problem %d: D_lo needs an epilogue that rounds its output", i);
      MTRL_REQUIRE(d.tma_out, "problem %d: D_lo needs the TMA output path (block_n %% 32 == 0)", i);
      MTRL_PROPAGATE(encode_map(&P.maps[kMapsPer * i + 5], p.D_lo, p.N, p.M, p.ldd, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B));
      d.d_lo = 1;
    }
    const bool allow3d = !(getenv("MTRL_GEMM_NO_3D") && getenv("MTRL_GEMM_NO_3D")[0] == '1');
    d.mn3d = 0;
    // the remainder operands (fp32x3) share shape, pitch and layout with their hi tensors: same map kinds
    for (int part = 0; part < (d.x3 ? 2 : 1); ++part) {
      const float* pa = part ? p.A_lo : p.A;
      const float* pb = part ? p.B_lo : p.B;
      CUtensorMap* ma = &P.maps[kMapsPer * i + (part ? 3 : 0)];
      CUtensorMap* mb = &P.maps[kMapsPer * i + (part ? 4 : 1)];
      if (!d.a_major) {
        MTRL_PROPAGATE(encode_map(ma, pa, p.K, p.M, p.lda, kBlockK, kBlockM, CU_TENSOR_MAP_SWIZZLE_128B));
      } else if ((part ? (d.mn3d & 1) != 0 : allow3d) && encode_map_mn3d(ma, pa, p.M, p.K, p.lda, kBlockM / 32)) {
        d.mn3d |= 1;
      } else {
        MTRL_REQUIRE(!(part && (d.mn3d & 1)), "problem %d: A_lo cannot use the 3-D map its hi tensor uses", i);
        MTRL_PROPAGATE(encode_map(ma, pa, p.M, p.K, p.lda, 32, kBlockK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
      }
      if (!d.b_major) {
        MTRL_PROPAGATE(encode_map(mb, pb, p.K, p.N, p.ldb, kBlockK, n_cta, CU_TENSOR_MAP_SWIZZLE_128B));
      } else if ((part ? (d.mn3d & 2) != 0 : (allow3d && n_cta % 32 == 0)) && encode_map_mn3d(mb, pb, p.N, p.K, p.ldb, n_cta / 32)) {
        d.mn3d |= 2;
      } else {
        MTRL_REQUIRE(!(part && (d.mn3d & 2)), "problem %d: B_lo cannot use the 3-D map its hi tensor uses", i);
        MTRL_PROPAGATE(encode_map(mb, pb, p.N, p.K, p.ldb, 32, kBlockK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
      }
    }
  }
  P.nprob = n;
  P.total_units = units;
  int nphases = 1;
  for (int i = 0; i < n; ++i) {
    MTRL_REQUIRE(problems[i].phase >= 0 && problems[i].phase < kMaxPhases, "problem %d: phase %d outside [0, %d)", i,
                 problems[i].phase, kMaxPhases);
    nphases = std::max(nphases, problems[i].phase + 1);
  }
  P.nphases = nphases;
  int row_counters = 0;
  if (rowdeps) {
    // who reads whose output: a problem of a later phase whose A operand is the D of an earlier problem waits per row tile
    for (int i = 0; i < n; ++i) {
      const mtrl_gemm_problem_t& p = problems[i];
      for (int j = 0; j < n; ++j) {
        const mtrl_gemm_problem_t& q = problems[j];
        if (j == i || q.phase >= p.phase) continue;
        MTRL_REQUIRE(p.B != q.D && (!p.B_lo || p.B_lo != q.D_lo),
                     "row-dependency plan: problem %d reads the output of problem %d as its B operand (needs a grid barrier)", i, j);
        if (p.A != q.D) continue;
        MTRL_REQUIRE(q.phase == p.phase - 1 && !p.a_major && p.lda == q.ldd && p.M == q.M && p.K == q.N && P.probs[j].tma_out &&
                         P.probs[j].k_splits == 1 && q.epilogue != MTRL_EPI_ATOMIC_ADD && (p.A_lo == nullptr || p.A_lo == q.D_lo),
                     "row-dependency plan: problem %d cannot chain on problem %d (phase, layout, shape or epilogue)", i, j);
        MTRL_REQUIRE(P.probs[i].dep_prob < 0, "row-dependency plan: problem %d has two producers", i);
        P.probs[i].dep_prob = j;
        if (P.probs[j].sig_base < 0) {
          P.probs[j].sig_base = row_counters;
          row_counters += P.probs[j].m_tiles;
        }
      }
      MTRL_REQUIRE(p.phase == 0 || P.probs[i].dep_prob >= 0,
                   "row-dependency plan: problem %d of phase %d reads no output of the previous phase", i, p.phase);
    }
  }
  const int workers = sms / ctas;
  // with phases every CTA must be resident at once (they meet at in-kernel barriers): never more workers than CTA slots
  const int nworkers = (units < workers && !streamk) ? units : workers;
  plan->grid = nworkers * ctas;
  plan->streamk = streamk;
  if (streamk) {
    // Stream-K: per phase, the k-blocks of all tiles (weighted by tile width) are laid end to end and cut into `nworkers` equal
    // shares; a worker's share is a run of whole tiles with at most one partial tile at either end.  Pieces shorter than
    // kMinPiece k-blocks are not cut off (the neighbour keeps them), so shares differ by a few k-blocks at most.
    constexpr int kMinPiece = 4;
    struct Piece { int p, tile, kb0, kb1; };
    std::vector<int4> unit_tab, fix_tab;
    std::vector<std::vector<int>> lists(static_cast<size_t>(nworkers) * nphases);
    long long ws_floats = 0;
    for (int ph = 0; ph < nphases; ++ph) {
      struct Tile { int p, tile, kb0, kb1; long long w; };
      std::vector<Tile> tiles;
      long long total = 0;
      for (int i = 0; i < n; ++i) {
        if (problems[i].phase != ph) continue;
        const DevProblem& d = P.probs[i];
        MTRL_REQUIRE(d.kb_total < 65536, "problem %d: K too long for a stream-K plan", i);
        const long long w = (d.block_n > 160 ? d.block_n : 160) * (d.x3 ? 3 : 1);
        for (int t = 0; t < d.m_tiles * d.n_tiles; ++t)
          for (int sp = 0; sp < d.k_splits; ++sp) {   // existing split-K units (accumulating epilogue) stay separate runs
            const int kb0 = sp * d.kb_per_split, kb1 = std::min(d.kb_total, kb0 + d.kb_per_split);
            tiles.push_back({i, t, kb0, kb1, w});
            total += (kb1 - kb0) * w;
          }
      }
      if (tiles.empty()) continue;
      // (no per-unit constant here, unlike the LPT cost model: every worker ends up with about the same number of units, and
      // a constant that the running sum counts per UNIT but the total per TILE would starve the last worker's predecessors)
      const long long overhead = 0;
      size_t ti = 0;
      int kb = tiles[0].kb0;
      long long done = 0;
      for (int wk = 0; wk < nworkers && ti < tiles.size(); ++wk) {
        const long long target = total * (wk + 1) / nworkers;   // cumulative boundary: rounding never accumulates
        while (ti < tiles.size() && (done < target || wk == nworkers - 1)) {
          const Tile& t = tiles[ti];
          const long long room = target - done;
          int take = t.kb1 - kb;
          if (wk != nworkers - 1 && (take * t.w + overhead) > room) {
            take = static_cast<int>((room - overhead) / t.w);
            if (take < kMinPiece) break;                              // not worth a unit: the next worker starts here
            if (t.kb1 - (kb + take) < kMinPiece) take = t.kb1 - kb;   // do not leave a sliver behind
          }
          const bool whole = kb == t.kb0 && kb + take == t.kb1;
          unit_tab.push_back(make_int4(t.p, t.tile, kb | ((kb + take) << 16), whole ? -1 : -2 - static_cast<int>(ti)));
          lists[static_cast<size_t>(wk) * nphases + ph].push_back(static_cast<int>(unit_tab.size()) - 1);
          done += take * t.w + overhead;
          kb += take;
          if (kb == t.kb1) {
            ++ti;
            if (ti < tiles.size()) kb = tiles[ti].kb0;
          }
        }
      }
      MTRL_REQUIRE(ti == tiles.size(), "stream-K schedule left %zu tiles unassigned", tiles.size() - ti);
      // pieces of one tile: accumulate-only epilogues need no fix-up (fix = -1); the others share a fix entry
      std::vector<int> npieces(tiles.size(), 0), fix_of(tiles.size(), -1);
      for (const int4& u : unit_tab)
        if (u.w <= -2) npieces[-2 - u.w]++;
      for (int4& u : unit_tab) {
        if (u.w > -2) continue;
        const int t = -2 - u.w;
        const DevProblem& d = P.probs[tiles[t].p];
        if (d.epilogue == MTRL_EPI_ATOMIC_ADD) { u.w = -1; continue; }
        if (fix_of[t] < 0) {
          fix_of[t] = static_cast<int>(fix_tab.size());
          fix_tab.push_back(make_int4(npieces[t], static_cast<int>(ws_floats & 0xFFFFFFFFll), static_cast<int>(ws_floats >> 32), 0));
          ws_floats += static_cast<long long>(npieces[t] - 1) * ctas * kFixSlabFloats;
        }
        u.w = fix_of[t];
      }
      for (int4& u : unit_tab)
        if (u.w <= -2) u.w = -1;   // (units of earlier phases are already resolved; nothing left at <= -2 here)
    }
    P.total_units = static_cast<int>(unit_tab.size());
    if (getenv("MTRL_GEMM_STREAMK_DUMP")) {
      for (int wk = 0; wk < nworkers; ++wk)
        for (int ph = 0; ph < nphases; ++ph) {
          fprintf(stderr, "[stream-K] worker %d phase %d:", wk, ph);
          for (int u : lists[static_cast<size_t>(wk) * nphases + ph]) {
            const int4 e = unit_tab[u];
            fprintf(stderr, " (p%d t%d kb %d-%d fix %d)", e.x, e.y, e.z & 0xFFFF, static_cast<unsigned>(e.z) >> 16, e.w);
          }
          fprintf(stderr, "\n");
        }
      for (size_t f = 0; f < fix_tab.size(); ++f) fprintf(stderr, "[stream-K] fix %zu: pieces %d\n", f, fix_tab[f].x);
    }
    std::vector<int> table(static_cast<size_t>(nworkers) * nphases + 1 + unit_tab.size());
    int off = 0;
    for (size_t k = 0; k < lists.size(); ++k) {
      table[k] = off;
      for (int u : lists[k]) table[static_cast<size_t>(nworkers) * nphases + 1 + off++] = u;
    }
    table[static_cast<size_t>(nworkers) * nphases] = off;
    MTRL_CUDA_CHECK(cudaMalloc(&plan->d_sched, table.size() * sizeof(int)));
    MTRL_CUDA_CHECK(cudaMemcpy(plan->d_sched, table.data(), table.size() * sizeof(int), cudaMemcpyHostToDevice));
    P.sched = plan->d_sched;
    MTRL_CUDA_CHECK(cudaMalloc(&plan->d_phase_cnt, kMaxPhases * sizeof(unsigned long long)));
    MTRL_CUDA_CHECK(cudaMemset(plan->d_phase_cnt, 0, kMaxPhases * sizeof(unsigned long long)));
    P.phase_cnt = plan->d_phase_cnt;
    MTRL_CUDA_CHECK(cudaMalloc(&plan->d_unit_tab, std::max<size_t>(unit_tab.size(), 1) * sizeof(int4)));
    MTRL_CUDA_CHECK(cudaMemcpy(plan->d_unit_tab, unit_tab.data(), unit_tab.size() * sizeof(int4), cudaMemcpyHostToDevice));
    P.unit_tab = plan->d_unit_tab;
    MTRL_CUDA_CHECK(cudaMalloc(&plan->d_fix_tab, std::max<size_t>(fix_tab.size(), 1) * sizeof(int4)));
    MTRL_CUDA_CHECK(cudaMemcpy(plan->d_fix_tab, fix_tab.data(), fix_tab.size() * sizeof(int4), cudaMemcpyHostToDevice));
    P.fix_tab = plan->d_fix_tab;
    MTRL_CUDA_CHECK(cudaMalloc(&plan->d_fix_ws, std::max<long long>(ws_floats, 4) * sizeof(float)));
    P.fix_ws = plan->d_fix_ws;
    const size_t cnt_bytes = std::max<size_t>(fix_tab.size(), 1) * ctas * 4 * sizeof(unsigned);
    MTRL_CUDA_CHECK(cudaMalloc(&plan->d_fix_cnt, cnt_bytes));
    MTRL_CUDA_CHECK(cudaMemset(plan->d_fix_cnt, 0, cnt_bytes));
    P.fix_cnt = plan->d_fix_cnt;
  } else {
    // Longest-processing-time-first assignment, phase by phase.  Unit cost ~ k-blocks x tile width (narrow tiles are
    // bounded by the per-k-block TMA / issue latency, not by the MMA) + a constant for prologue and epilogue drain.
    // Units of equal cost keep their index order, so a launch of uniform units degenerates to the round-robin it
    // replaces (neighbouring workers share operand tiles in L2).
    struct U { int unit; long long cost; int first; };
    const bool lpt = !(getenv("MTRL_GEMM_NO_LPT") && getenv("MTRL_GEMM_NO_LPT")[0] == '1');
    std::vector<std::vector<int>> lists(static_cast<size_t>(nworkers) * nphases);
    for (int ph = 0; ph < nphases; ++ph) {
      std::vector<U> us;
      for (int i = 0; i < n; ++i) {
        if (problems[i].phase != ph) continue;
        const DevProblem& d = P.probs[i];
        for (int u = 0; u < d.unit_count; ++u) {
          const int split = u / (d.n_tiles * d.m_tiles);
          const int kb0 = split * d.kb_per_split;
          const int kb1 = kb0 + d.kb_per_split < d.kb_total ? kb0 + d.kb_per_split : d.kb_total;
          const long long per_kb = d.block_n > 160 ? d.block_n : 160;
          us.push_back({d.unit_begin + u, static_cast<long long>(kb1 - kb0) * per_kb * (d.x3 ? 3 : 1) + 4 * 256,
                        problems[i].schedule_first ? 1 : 0});
        }
      }
      // schedule_first problems (outputs reduce-added into a peer GPU) lead every worker's list; cost order within a class
      if (lpt)
        std::stable_sort(us.begin(), us.end(), [](const U& a, const U& b) { return a.first != b.first ? a.first > b.first : a.cost > b.cost; });
      std::vector<long long> load(nworkers, 0);
      for (size_t i = 0; i < us.size(); ++i) {
        int best = static_cast<int>(i % nworkers);
        if (lpt) {
          best = 0;
          for (int w = 1; w < nworkers; ++w)
            if (load[w] < load[best]) best = w;
        }
        lists[static_cast<size_t>(best) * nphases + ph].push_back(us[i].unit);
        load[best] += us[i].cost;
      }
    }
    std::vector<int> table(static_cast<size_t>(nworkers) * nphases + 1 + units);
    int off = 0;
    for (size_t k = 0; k < lists.size(); ++k) {
      table[k] = off;
      for (int u : lists[k]) table[static_cast<size_t>(nworkers) * nphases + 1 + off++] = u;
    }
    table[static_cast<size_t>(nworkers) * nphases] = off;
    MTRL_CUDA_CHECK(cudaMalloc(&plan->d_sched, table.size() * sizeof(int)));
    MTRL_CUDA_CHECK(cudaMemcpy(plan->d_sched, table.data(), table.size() * sizeof(int), cudaMemcpyHostToDevice));
    P.sched = plan->d_sched;
    MTRL_CUDA_CHECK(cudaMalloc(&plan->d_phase_cnt, kMaxPhases * sizeof(unsigned long long)));
    MTRL_CUDA_CHECK(cudaMemset(plan->d_phase_cnt, 0, kMaxPhases * sizeof(unsigned long long)));
    P.phase_cnt = plan->d_phase_cnt;
  }
  if (rowdeps) {
    const size_t bytes = (static_cast<size_t>(row_counters) + 1) * sizeof(unsigned long long);
    MTRL_CUDA_CHECK(cudaMalloc(&plan->d_row_cnt, bytes));
    MTRL_CUDA_CHECK(cudaMemset(plan->d_row_cnt, 0, bytes));
    P.rowdeps = 1;
    P.row_cnt = plan->d_row_cnt + 1;
    P.launch_cnt = plan->d_row_cnt;
  }
  static bool attr_set = false;
  if (!attr_set) {
    MTRL_CUDA_CHECK(cudaFuncSetAttribute(gemm_tf32_grouped_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg<1>::kSmemBytes));
    MTRL_CUDA_CHECK(cudaFuncSetAttribute(gemm_tf32_grouped_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg<2>::kSmemBytes));
    attr_set = true;
  }
  guard.p = nullptr;
  *out = plan;
  return MTRL_OK;
}

extern "C" int mtrl_gemm_plan_run(mtrl_gemm_plan_t* plan, void* stream) {
  MTRL_REQUIRE(plan, "mtrl_gemm_plan_run: null plan");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (plan->ctas == 1) {
    MTRL_CUDA_CHECK(mtrl_launch(gemm_tf32_grouped_kernel<1>, dim3(plan->grid), dim3(kThreads), Cfg<1>::kSmemBytes, st, plan->params));
    return MTRL_OK;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(plan->grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = Cfg<2>::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = mtrl_pdl_enabled() ? 2 : 1;
  MTRL_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_tf32_grouped_kernel<2>, plan->params));
  return MTRL_OK;
}

// Debug aid: dbg = device long long[8], zeroed by the caller.  Cycle sums over CTAs of
// [0] producer wait-empty, [1] producer TMA issue, [2] MMA wait-full, [3] MMA wait-tmem-empty, [4] MMA issue,
// [5] MMA thread total, [6] epilogue wait-tmem-full (warp 2), [7] epilogue work (warp 2).
extern "C" int mtrl_gemm_plan_set_debug(mtrl_gemm_plan_t* plan, long long* dbg) {
  MTRL_REQUIRE(plan, "mtrl_gemm_plan_set_debug: null plan");
  plan->params.dbg = dbg;
  return MTRL_OK;
}

extern "C" int mtrl_gemm_plan_units(const mtrl_gemm_plan_t* plan) { return plan ? plan->params.total_units : 0; }
extern "C" int mtrl_gemm_plan_ctas(const mtrl_gemm_plan_t* plan) { return plan ? plan->ctas : 0; }

extern "C" void mtrl_gemm_plan_destroy(mtrl_gemm_plan_t* plan) {
  delete plan;
}
