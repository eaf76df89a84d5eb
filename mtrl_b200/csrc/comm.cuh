// Trunk-gradient exchange between the task shards of one box WITHOUT a library collective:
// every rank maps every peer's exchange arena (CUDA IPC over NVLink / NVSwitch) and ONE kernel per network does
//
//   reduce-scatter (peer loads)  ->  global-norm clip  ->  Adam on the owned 1/N of the trunk  ->
//   all-gather (peer stores of the new parameters)      ->  local Polyak / tf32 operand copies
//
// i.e. optax.chain(clip_by_global_norm, adam) + apply_updates (mtrl/config/optim.py:26-43,
// mtrl/rl/algorithms/utils.py:11-46) and incremental_update (mtrl/rl/algorithms/mtsac.py:607-613) with the Adam state
// of the replicated trunk sharded over the ranks (each element is updated by exactly one rank, then broadcast, so the
// replicas stay bit-identical).  Head parameters belong to one rank and are stepped locally in the same launch.
#pragma once

#include "common.cuh"
#include "mtrl_b200.h"
#include "sac_kernels.cuh"

#define MTRL_COMM_MAX_RANKS 8
#define MTRL_COMM_HEADER_BYTES 4096

namespace comm {

// First MTRL_COMM_HEADER_BYTES of every arena.  `flag`, `inbox_*` are written by peers; the rest is local.
struct Header {
  unsigned int flag[4][MTRL_COMM_MAX_RANKS];      // [barrier][source rank] = epoch stamp
  double inbox_g2[MTRL_COMM_MAX_RANKS];           // squared norm of the reduced trunk shard rank q owns
  float inbox_head_g2[MTRL_COMM_MAX_RANKS];       // squared norm of rank q's (local) head gradients
  unsigned long long grid_count;                  // in-rank arrivals, monotonic
  unsigned int epoch;                             // stamp of the next exchange (starts at 1)
  int error;                                      // != 0: a wait timed out (peer missing)
  double shard_g2[2];                             // this rank's shard accumulator, double-buffered by epoch parity
  unsigned int sync_epoch;                        // stamp of the next rank_barrier_kernel (flag[3], starts at 1)
  unsigned long long phase_ns[2][8];              // [epoch parity][phase]: globaltimer stamps of CTA 0 in the last two
                                                  // trunk steps: start, barrier 0, norms, barrier 1, Adam + all-gather,
                                                  // barrier 2, derived copies (mtrl_comm_phase_times)
  unsigned long long timeout_ns;                  // how long an in-kernel wait for a peer may last (MTRL_COMM_TIMEOUT_S)
};

// Ownership of the replicated trunk: the flat buffer is cut into segments, each stepped (Adam) by exactly one rank.
// pre_reduced segments are the row blocks of the hidden-layer kernels whose gradient every rank's dW GEMM epilogue
// reduce-adds straight into the owner's buffer over NVLink (sac.cu build_plans); the rest (first-layer kernel,
// biases: small) is summed by the owner with peer loads.
struct Segment {
  long long begin4, end4;   // float4 units into the flat network buffer
  int owner;
  int pre_reduced;
};
static_assert(sizeof(Header) <= MTRL_COMM_HEADER_BYTES, "comm header overflows its page");

#ifdef __CUDACC__

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// Bulk peer (or peer-written) data: ld.global.cg -- served by the L2 of the GPU that owns the line (peer memory is
// not cached in the local L2), never by a possibly stale local L1 line.  A plain intrinsic rather than volatile asm so
// that the compiler can keep several loads of an unrolled loop in flight.
__device__ __forceinline__ float4 ld_sys_f4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ double ld_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_sys_f32(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// A wait for a peer that exceeds Header::timeout_ns (default 30 s, MTRL_COMM_TIMEOUT_S; a missing peer must not hang the
// box) sets Header::error.  The error is STICKY and fatal for the exchange: the kernel that saw it, and every later
// exchange kernel of this rank, returns without reducing, stepping Adam or broadcasting anything (a late peer then times
// out at its next barrier too, so no rank applies a step the others did not), and the code is copied into the update's
// status word, where the host's normal status check raises (mtrl_sac_read_status_async, word 1).

struct TrunkStepArgs {
  float *p, *m, *v, *shadow, *g, *target, *target_shadow;   // local buffers of this network (target may be null)
  float *shadow_lo, *target_shadow_lo;                      // fp32x3: tf32 remainders of the operand copies (or null)
  long long n, trunk_n;                                    // as AdamArgs
  float* peer_g[MTRL_COMM_MAX_RANKS];                      // every rank's gradient buffer ([rank] == g)
  float* peer_p[MTRL_COMM_MAX_RANKS];                      // every rank's parameter buffer ([rank] == p)
  float* mc_p;                                             // multicast alias of the parameter buffers (one multimem.st reaches
                                                           // every rank's copy, this rank's included), or null
  Header* peer_hdr[MTRL_COMM_MAX_RANKS];
  const Segment* segs;                                     // device table covering [0, trunk_n)
  int nsegs;
  int rank, world;
  const int* step;
  int* status;                                             // the update's status words: [1] receives Header::error
  double* g2_trunk_out;                                    // global trunk gradient squared norm (log scalar input)
  double *p2_trunk, *p2_head;
  float lr, b1, b2, eps, max_norm, tau;
};

// All CTAs: wait until every rank has stamped barrier `b` of this rank's header with `epoch`.
__device__ __forceinline__ void wait_ranks(Header* H, int b, int world, unsigned epoch) {
  if (threadIdx.x < world) {
    const unsigned long long t0 = globaltimer_ns();
    const unsigned long long limit = H->timeout_ns;
    while (static_cast<int>(ld_acquire_sys(&H->flag[b][threadIdx.x]) - epoch) < 0) {
      if (globaltimer_ns() - t0 > limit || *reinterpret_cast<volatile int*>(&H->error) != 0) {
        atomicExch(&H->error, 1 + b);
        break;
      }
      __nanosleep(100);
    }
  }
  __syncthreads();
}

// Every CTA arrives; CTA 0 waits for the whole grid, then stamps barrier `b` in every rank's header.
__device__ __forceinline__ void grid_arrive_then_signal(const TrunkStepArgs& a, Header* H, int b, unsigned epoch,
                                                        unsigned long long target, bool with_norms) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    atomicAdd(&H->grid_count, 1ull);
  }
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) {
      const unsigned long long t0 = globaltimer_ns();
      const unsigned long long limit = H->timeout_ns;
      while (ld_acquire_gpu(&H->grid_count) < target) {
        if (globaltimer_ns() - t0 > limit) {
          atomicExch(&H->error, 10 + b);
          break;
        }
      }
      __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x < a.world) {
      Header* R = a.peer_hdr[threadIdx.x];
      if (with_norms) {
        R->inbox_g2[a.rank] = ld_sys_f64(&H->shard_g2[epoch & 1]);
        R->inbox_head_g2[a.rank] = a.g[a.trunk_n];   // slot 0: local head-gradient squared norm (write_slot_kernel)
        __threadfence_system();
      }
      st_release_sys(&R->flag[b][a.rank], epoch);
    }
  }
}

template <int WORLD>
__device__ __forceinline__ float4 reduce_ranks(const TrunkStepArgs& a, long long i4) {
  float4 x[WORLD];
#pragma unroll
  for (int q = 0; q < WORLD; ++q) x[q] = ld_sys_f4(a.peer_g[q] + i4 * 4);
  float4 s = x[0];
#pragma unroll
  for (int q = 1; q < WORLD; ++q) {
    s.x += x[q].x; s.y += x[q].y; s.z += x[q].z; s.w += x[q].w;
  }
  return s;
}
template <>
__device__ __forceinline__ float4 reduce_ranks<0>(const TrunkStepArgs& a, long long i4) {
  float4 s = ld_sys_f4(a.peer_g[0] + i4 * 4);
  for (int q = 1; q < a.world; ++q) {
    const float4 x = ld_sys_f4(a.peer_g[q] + i4 * 4);
    s.x += x.x; s.y += x.y; s.z += x.z; s.w += x.w;
  }
  return s;
}

__device__ __forceinline__ void adam4(const TrunkStepArgs& a, float scale, float bc1, float bc2, const float4& g4, float4& m4,
                                      float4& v4, float4& p4) {
  const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
  float* mm = reinterpret_cast<float*>(&m4);
  float* vv = reinterpret_cast<float*>(&v4);
  float* pp = reinterpret_cast<float*>(&p4);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float g = gg[q] * scale;
    mm[q] = a.b1 * mm[q] + (1.f - a.b1) * g;
    vv[q] = a.b2 * vv[q] + (1.f - a.b2) * g * g;
    pp[q] = pp[q] - a.lr * (mm[q] / bc1) / (sqrtf(vv[q] / bc2) + a.eps);
  }
}

// shadow = tf32(p); target = tau p + (1 - tau) target; target_shadow = tf32(target); returns |p|^2.
// derived4t takes the already loaded target element.
// hi (and, when lo is non-null, remainder) operand copies of four values
__device__ __forceinline__ void store_operand4(float* hi, float* lo, long long i4, const float4& x) {
  const float4 h = make_float4(tf32_rna(x.x), tf32_rna(x.y), tf32_rna(x.z), tf32_rna(x.w));
  reinterpret_cast<float4*>(hi)[i4] = h;
  if (lo) reinterpret_cast<float4*>(lo)[i4] = make_float4(tf32_lo(x.x, h.x), tf32_lo(x.y, h.y), tf32_lo(x.z, h.z), tf32_lo(x.w, h.w));
}
__device__ __forceinline__ float derived4t(const TrunkStepArgs& a, long long i4, const float4& p4, float4 t4) {
  store_operand4(a.shadow, a.shadow_lo, i4, p4);
  if (a.target) {
    t4.x = a.tau * p4.x + (1.f - a.tau) * t4.x;
    t4.y = a.tau * p4.y + (1.f - a.tau) * t4.y;
    t4.z = a.tau * p4.z + (1.f - a.tau) * t4.z;
    t4.w = a.tau * p4.w + (1.f - a.tau) * t4.w;
    reinterpret_cast<float4*>(a.target)[i4] = t4;
    store_operand4(a.target_shadow, a.target_shadow_lo, i4, t4);
  }
  return p4.x * p4.x + p4.y * p4.y + p4.z * p4.z + p4.w * p4.w;
}
__device__ __forceinline__ float derived4(const TrunkStepArgs& a, long long i4, const float4& p4) {
  float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (a.target) t4 = reinterpret_cast<const float4*>(a.target)[i4];
  return derived4t(a, i4, p4, t4);
}

// Every rank has finished what precedes this kernel in its stream (one warp; used before the first gradient
// GEMM of an update so that no rank reduce-adds into a peer buffer that is not zeroed yet).
static __global__ void rank_barrier_kernel(Header* const* peer_hdr_dev, int rank, int world, int* status) {
  __shared__ Header* hdr[MTRL_COMM_MAX_RANKS];
  if (threadIdx.x < world) hdr[threadIdx.x] = peer_hdr_dev[threadIdx.x];
  __syncthreads();
  Header* H = hdr[rank];
  if (*reinterpret_cast<volatile int*>(&H->error) != 0) {   // sticky: the exchange is dead
    if (threadIdx.x == 0) status[1] = H->error;
    return;
  }
  const unsigned epoch = *reinterpret_cast<volatile unsigned*>(&H->sync_epoch);
  __syncthreads();
  if (threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(&hdr[threadIdx.x]->flag[3][rank], epoch);
  }
  wait_ranks(H, 3, world, epoch);
  if (threadIdx.x == 0) {
    *reinterpret_cast<volatile unsigned*>(&H->sync_epoch) = epoch + 1;
    if (*reinterpret_cast<volatile int*>(&H->error) != 0) status[1] = H->error;
  }
}

// Launch with one CTA per SM (all CTAs must be co-resident: they meet at in-kernel barriers).
template <int WORLD>
static __global__ void __launch_bounds__(512, 1) trunk_step_kernel(const TrunkStepArgs a) {
  __shared__ double red[32];
  __shared__ float s_scale;
  // the ownership table is walked three times: keep it in shared memory (a dependent global load per entry costs more
  // than the entry's work once the segments are small)
  constexpr int kMaxSmemSegs = 128;
  __shared__ Segment s_segs[kMaxSmemSegs];
  const bool segs_in_smem = a.nsegs <= kMaxSmemSegs;
  if (segs_in_smem)
    for (int i = threadIdx.x; i < a.nsegs; i += blockDim.x) s_segs[i] = a.segs[i];
  __syncthreads();
  const Segment* segs = segs_in_smem ? s_segs : a.segs;
  Header* H = a.peer_hdr[a.rank];
  // every CTA reads the (sticky) error word at the same points, right after a barrier wait, so they agree on leaving
  auto dead = [&]() {
    const int e = *reinterpret_cast<volatile int*>(&H->error);
    if (e != 0 && blockIdx.x == 0 && threadIdx.x == 0) a.status[1] = e;
    return e != 0;
  };
  if (dead()) return;
  const unsigned epoch = *reinterpret_cast<volatile unsigned*>(&H->epoch);
  const unsigned long long grid_base = static_cast<unsigned long long>(epoch - 1) * 2ull * gridDim.x;
  const int world = WORLD ? WORLD : a.world;
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;

  if (blockIdx.x == 0 && threadIdx.x == 0) H->phase_ns[epoch & 1][0] = globaltimer_ns();
  // ---- barrier 0: every rank's gradient kernels are complete (by stream order, once its kernel runs), so the
  //      reduce-adds they sent into this rank's pre-reduced segments have landed too ----
  if (blockIdx.x == 0 && threadIdx.x < world) st_release_sys(&a.peer_hdr[threadIdx.x]->flag[0][a.rank], epoch);
  wait_ranks(H, 0, world, epoch);
  if (dead()) return;   // a peer never finished its gradient kernels: nothing has been touched yet
  if (blockIdx.x == 0 && threadIdx.x == 0) H->phase_ns[epoch & 1][1] = globaltimer_ns();

  // ---- owned segments: finish the reduction where the GEMMs have not done it, and take the squared norm ----
  {
    double s = 0.0;
    for (int si = 0; si < a.nsegs; ++si) {
      const Segment sg = segs[si];
      if (sg.owner != a.rank) continue;
      if (sg.pre_reduced) {
#pragma unroll 4
        for (long long i = sg.begin4 + tid; i < sg.end4; i += stride) {
          const float4 g4 = ld_sys_f4(a.g + i * 4);   // written by peers' reduce-adds: read at L2
          s += static_cast<double>(g4.x * g4.x + g4.y * g4.y) + static_cast<double>(g4.z * g4.z + g4.w * g4.w);
        }
      } else {
        for (long long i = sg.begin4 + tid; i < sg.end4; i += stride) {
          const float4 g4 = reduce_ranks<WORLD>(a, i);
          reinterpret_cast<float4*>(a.g)[i] = g4;
          s += static_cast<double>(g4.x * g4.x + g4.y * g4.y) + static_cast<double>(g4.z * g4.z + g4.w * g4.w);
        }
      }
    }
    s = sac::block_sum(s, red);
    if (threadIdx.x == 0) atomicAdd(&H->shard_g2[epoch & 1], s);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) H->phase_ns[epoch & 1][2] = globaltimer_ns();
  // ---- barrier 1: exchange the shard norms and the head norms ----
  grid_arrive_then_signal(a, H, 1, epoch, grid_base + gridDim.x, true);
  wait_ranks(H, 1, world, epoch);
  if (dead()) return;   // incomplete norms: no clip scale, no Adam, no broadcast
  if (blockIdx.x == 0 && threadIdx.x == 0) H->phase_ns[epoch & 1][3] = globaltimer_ns();
  if (threadIdx.x == 0) {
    double g2 = 0.0, h2 = 0.0;
    for (int q = 0; q < world; ++q) {
      g2 += ld_sys_f64(&H->inbox_g2[q]);
      h2 += static_cast<double>(ld_sys_f32(&H->inbox_head_g2[q]));
    }
    const float gn = static_cast<float>(sqrt(g2 + static_cast<double>(static_cast<float>(h2))));
    // optax.clip_by_global_norm: g if norm < max else g / norm * max
    s_scale = (a.max_norm > 0.f && !(gn < a.max_norm)) ? a.max_norm / gn : 1.f;
    if (blockIdx.x == 0) {
      *a.g2_trunk_out = g2;
      a.g[a.trunk_n + 1] = static_cast<float>(h2);   // slot 1: head norm summed over ranks (read by the finalize kernel)
    }
  }
  __syncthreads();
  const float scale = s_scale;
  const int t = *a.step + 1;
  const float bc1 = static_cast<float>(1.0 - pow(static_cast<double>(a.b1), static_cast<double>(t)));
  const float bc2 = static_cast<float>(1.0 - pow(static_cast<double>(a.b2), static_cast<double>(t)));

  // ---- Adam on the owned segments; all-gather by storing the new parameters into every rank ----
  double p2_trunk = 0.0, p2_head = 0.0;
  for (int si = 0; si < a.nsegs; ++si) {
    const Segment sg = segs[si];
    if (sg.owner != a.rank) continue;
    for (long long i = sg.begin4 + tid; i < sg.end4; i += stride) {
      const float4 g4 = sg.pre_reduced ? ld_sys_f4(a.g + i * 4) : reinterpret_cast<const float4*>(a.g)[i];
      float4 m4 = reinterpret_cast<const float4*>(a.m)[i];
      float4 v4 = reinterpret_cast<const float4*>(a.v)[i];
      float4 p4 = reinterpret_cast<const float4*>(a.p)[i];
      adam4(a, scale, bc1, bc2, g4, m4, v4, p4);
      reinterpret_cast<float4*>(a.m)[i] = m4;
      reinterpret_cast<float4*>(a.v)[i] = v4;
      if (a.mc_p) {
        // NVLS: the switch replicates the store into every rank's copy
        asm volatile("multimem.st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a.mc_p + i * 4), "f"(p4.x), "f"(p4.y), "f"(p4.z),
                     "f"(p4.w)
                     : "memory");
      } else {
#pragma unroll
        for (int q = 0; q < (WORLD ? WORLD : MTRL_COMM_MAX_RANKS); ++q)
          if (q < world) reinterpret_cast<float4*>(a.peer_p[q])[i] = p4;
      }
      p2_trunk += static_cast<double>(derived4(a, i, p4));
    }
  }
  // ---- heads (local to this rank) ----
  for (long long i = (a.trunk_n + 32) / 4 + tid; i < a.n / 4; i += stride) {
    const float4 g4 = reinterpret_cast<const float4*>(a.g)[i];
    float4 m4 = reinterpret_cast<const float4*>(a.m)[i];
    float4 v4 = reinterpret_cast<const float4*>(a.v)[i];
    float4 p4 = reinterpret_cast<const float4*>(a.p)[i];
    adam4(a, scale, bc1, bc2, g4, m4, v4, p4);
    reinterpret_cast<float4*>(a.m)[i] = m4;
    reinterpret_cast<float4*>(a.v)[i] = v4;
    reinterpret_cast<float4*>(a.p)[i] = p4;
    p2_head += static_cast<double>(derived4(a, i, p4));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) H->phase_ns[epoch & 1][4] = globaltimer_ns();
  // ---- barrier 2: every rank has stored its segments everywhere ----
  grid_arrive_then_signal(a, H, 2, epoch, grid_base + 2ull * gridDim.x, false);
  wait_ranks(H, 2, world, epoch);
  if (dead()) return;   // a peer's segments never arrived: leave the replicas' derived copies alone, the host raises
  if (blockIdx.x == 0 && threadIdx.x == 0) H->phase_ns[epoch & 1][5] = globaltimer_ns();

  // ---- derived copies of the segments the peers own ----
  for (int si = 0; si < a.nsegs; ++si) {
    const Segment sg = segs[si];
    if (sg.owner == a.rank) continue;
    // two elements per pass: both parameter and both target loads are issued before the dependent stores
    long long i = sg.begin4 + tid;
    for (; i + stride < sg.end4; i += 2 * stride) {
      const float4 pa = ld_sys_f4(a.p + i * 4), pb = ld_sys_f4(a.p + (i + stride) * 4);
      float4 ta = make_float4(0.f, 0.f, 0.f, 0.f), tb = ta;
      if (a.target) {
        ta = reinterpret_cast<const float4*>(a.target)[i];
        tb = reinterpret_cast<const float4*>(a.target)[i + stride];
      }
      p2_trunk += static_cast<double>(derived4t(a, i, pa, ta));
      p2_trunk += static_cast<double>(derived4t(a, i + stride, pb, tb));
    }
    for (; i < sg.end4; i += stride) {
      const float4 p4 = ld_sys_f4(a.p + i * 4);
      p2_trunk += static_cast<double>(derived4(a, i, p4));
    }
  }
  p2_trunk = sac::block_sum(p2_trunk, red);
  p2_head = sac::block_sum(p2_head, red);
  if (threadIdx.x == 0) {
    atomicAdd(a.p2_trunk, p2_trunk);
    atomicAdd(a.p2_head, p2_head);
    if (blockIdx.x == 0) {
      H->phase_ns[epoch & 1][6] = globaltimer_ns();
      H->shard_g2[(epoch + 1) & 1] = 0.0;   // next exchange's accumulator (nobody touches it during this one)
      *reinterpret_cast<volatile unsigned*>(&H->epoch) = epoch + 1;
    }
  }
}

#endif  // __CUDACC__

}  // namespace comm

struct mtrl_comm {
  int rank = 0, world = 1;
  long long arena_bytes = 0;
  uint8_t* arena = nullptr;
  uint8_t* peer[MTRL_COMM_MAX_RANKS] = {};
  bool opened = false;
  comm::Header** d_peer_hdr = nullptr;   // device copy of the header pointers (rank_barrier_kernel)
  // NVSwitch multicast region (comm.cu): the parameters live in mc_local (this rank's own physical copy); a multimem.st
  // to mc_ptr + offset lands at mc_local + offset of EVERY rank.  All zero: peers' parameter copies are written one by one.
  unsigned long long mc_handle = 0, mc_mem = 0;
  long long mc_bytes = 0;
  uint8_t *mc_local = nullptr, *mc_ptr = nullptr;
};
