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
#include <cuda.h>  // CUtensorMap types only; the encode entry point is fetched at run time

#include "common.cuh"
#include "mtrl_b200.h"

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 32;          // tf32 elements = 128 bytes = one SWIZZLE_128B row
constexpr int kMaxBlockN = 256;
constexpr int kUmmaK = 8;            // tf32: 32 bytes of K per tcgen05.mma
constexpr int kStages = 4;
constexpr int kABytes = kBlockM * kBlockK * 4;        // 16 KB
constexpr int kBBytes = kMaxBlockN * kBlockK * 4;     // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;        // 48 KB
constexpr int kChunkBytes = 32 * kBlockK * 4;         // one 32(MN) x 32(K) MN-major TMA box
constexpr int kTmemCols = 512;                        // two 256-column fp32 accumulators
constexpr int kThreads = 192;
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;

struct __align__(16) DevProblem {
  float* D;
  const float* bias;
  const float* mask;
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
  int b_chunks;  // number of 32-wide MN chunks of B per stage (MN-major B only)
  int pad;
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

constexpr int kMaxProblems = 8;

// Passed by value as a __grid_constant__ kernel parameter (the usual home of TMA descriptors).
struct GemmParams {
  CUtensorMap maps[2 * kMaxProblems];
  DevProblem probs[kMaxProblems];
  int nprob;
  int total_units;
};

struct UnitCoord {
  int p, m_tile, n_tile, kb0, kb1;
};

__device__ __forceinline__ UnitCoord decode_unit(const DevProblem* __restrict__ probs, int nprob,
                                                 int unit) {
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
  return c;
}

__global__ void __launch_bounds__(kThreads, 1)
gemm_tf32_grouped_kernel(const __grid_constant__ GemmParams params) {
  const DevProblem* probs = params.probs;
  const CUtensorMap* maps = params.maps;
  const int nprob = params.nprob;
  const int total_units = params.total_units;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kStages * kStageBytes;
  // barrier layout (8 bytes each): full[kStages], empty[kStages], tmem_full[2], tmem_empty[2]
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);  // one arrive per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        const UnitCoord c = decode_unit(probs, nprob, unit);
        const DevProblem& P = probs[c.p];
        const CUtensorMap* mapA = maps + 2 * c.p;
        const CUtensorMap* mapB = maps + 2 * c.p + 1;
        const int m0 = c.m_tile * kBlockM;
        const int n0 = c.n_tile * P.block_n;
        const uint32_t b_bytes =
            P.b_major ? static_cast<uint32_t>(P.b_chunks) * kChunkBytes
                      : static_cast<uint32_t>(P.block_n) * kBlockK * 4u;
        for (int kb = c.kb0; kb < c.kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * kStageBytes;
          const uint32_t sb = sa + kABytes;
          mbar_expect_tx(full_bar(stage), kABytes + b_bytes);
          const int k0 = kb * kBlockK;
          if (!P.a_major) {
            tma_load_2d(sa, mapA, full_bar(stage), k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < kBlockM / 32; ++j)
              tma_load_2d(sa + j * kChunkBytes, mapA, full_bar(stage), m0 + 32 * j, k0);
          }
          if (!P.b_major) {
            tma_load_2d(sb, mapB, full_bar(stage), k0, n0);
          } else {
            for (int j = 0; j < P.b_chunks; ++j)
              tma_load_2d(sb + j * kChunkBytes, mapB, full_bar(stage), n0 + 32 * j, k0);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        const UnitCoord c = decode_unit(probs, nprob, unit);
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
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc) * kMaxBlockN;
        for (int kb = c.kb0; kb < c.kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * kStageBytes;
          const uint32_t sb = sa + kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint64_t adesc = make_smem_desc(sa + k * a_kstep, a_lbo, a_sbo, a_lt);
            const uint64_t bdesc = make_smem_desc(sb + k * b_kstep, b_lbo, b_sbo, b_lt);
            umma_tf32(d_tmem, adesc, bdesc, idesc, (kb > c.kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));  // frees the smem slot when these MMAs retire
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32) are this warp's
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
      const UnitCoord c = decode_unit(probs, nprob, unit);
      const DevProblem& P = probs[c.p];
      const int row = c.m_tile * kBlockM + quarter * 32 + lane;
      const int n0 = c.n_tile * P.block_n;
      const bool row_ok = row < P.M;
      float* drow = P.D + static_cast<long long>(row) * P.ldd;
      const float* mrow = P.mask ? P.mask + static_cast<long long>(row) * P.ldmask : nullptr;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                             static_cast<uint32_t>(acc) * kMaxBlockN;
      for (int cc = 0; cc < P.block_n; cc += 16) {
        uint32_t v[16];
        tmem_ld16(t_row + cc, v);
        tmem_ld_wait();
        const int col0 = n0 + cc;
        if (row_ok) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int col = col0 + 4 * q;
            if (col + 4 <= P.N) {
              float4 o;
              o.x = __uint_as_float(v[4 * q + 0]);
              o.y = __uint_as_float(v[4 * q + 1]);
              o.z = __uint_as_float(v[4 * q + 2]);
              o.w = __uint_as_float(v[4 * q + 3]);
              if (P.epilogue == MTRL_EPI_BIAS_RELU) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(P.bias + col));
                o.x = tf32_rna(fmaxf(o.x + b.x, 0.f));
                o.y = tf32_rna(fmaxf(o.y + b.y, 0.f));
                o.z = tf32_rna(fmaxf(o.z + b.z, 0.f));
                o.w = tf32_rna(fmaxf(o.w + b.w, 0.f));
                *reinterpret_cast<float4*>(drow + col) = o;
              } else if (P.epilogue == MTRL_EPI_RELU_MASK) {
                const float4 h = __ldg(reinterpret_cast<const float4*>(mrow + col));
                o.x = h.x > 0.f ? tf32_rna(o.x) : 0.f;
                o.y = h.y > 0.f ? tf32_rna(o.y) : 0.f;
                o.z = h.z > 0.f ? tf32_rna(o.z) : 0.f;
                o.w = h.w > 0.f ? tf32_rna(o.w) : 0.f;
                *reinterpret_cast<float4*>(drow + col) = o;
              } else if (P.epilogue == MTRL_EPI_ATOMIC_ADD) {
                atomicAdd(reinterpret_cast<float4*>(drow + col), o);
              } else if (P.epilogue == MTRL_EPI_STORE_TF32) {
                o.x = tf32_rna(o.x);
                o.y = tf32_rna(o.y);
                o.z = tf32_rna(o.z);
                o.w = tf32_rna(o.w);
                *reinterpret_cast<float4*>(drow + col) = o;
              } else {
                *reinterpret_cast<float4*>(drow + col) = o;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
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

}  // namespace

struct mtrl_gemm_plan {
  GemmParams params;
  int grid = 0;
};

extern "C" int mtrl_gemm_plan_create(mtrl_gemm_plan_t** out, const mtrl_gemm_problem_t* problems, int n) {
  MTRL_REQUIRE(out && problems && n >= 1 && n <= kMaxProblems,
               "mtrl_gemm_plan_create: need 1..%d problems per launch, got %d", kMaxProblems, n);
  mtrl_gemm_plan* plan = new mtrl_gemm_plan();
  struct Guard {
    mtrl_gemm_plan* p;
    ~Guard() { delete p; }
  } guard{plan};
  GemmParams& P = plan->params;
  memset(&P, 0, sizeof(P));
  int units = 0;
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
    MTRL_REQUIRE(p.epilogue != MTRL_EPI_RELU_MASK || p.mask, "problem %d: mask epilogue without mask", i);
    MTRL_REQUIRE((reinterpret_cast<uintptr_t>(p.D) & 15u) == 0 && p.ldd % 4 == 0,
                 "problem %d: D must be 16-byte aligned with ldd %% 4 == 0", i);
    DevProblem& d = P.probs[i];
    d.D = p.D;
    d.bias = p.bias;
    d.mask = p.mask;
    d.ldd = p.ldd;
    d.ldmask = p.ldmask;
    d.M = p.M;
    d.N = p.N;
    d.K = p.K;
    d.block_n = p.block_n;
    d.m_tiles = (p.M + kBlockM - 1) / kBlockM;
    d.n_tiles = (p.N + p.block_n - 1) / p.block_n;
    d.kb_total = (p.K + kBlockK - 1) / kBlockK;
    int splits = p.k_splits < d.kb_total ? p.k_splits : d.kb_total;
    d.kb_per_split = (d.kb_total + splits - 1) / splits;
    d.k_splits = (d.kb_total + d.kb_per_split - 1) / d.kb_per_split;
    d.epilogue = p.epilogue;
    d.a_major = p.a_major ? 1 : 0;
    d.b_major = p.b_major ? 1 : 0;
    d.b_chunks = (p.block_n + 31) / 32;
    d.unit_begin = units;
    d.unit_count = d.m_tiles * d.n_tiles * d.k_splits;
    units += d.unit_count;
    // cute::UMMA::InstrDescriptor: c_format F32 (1) [4,6), a/b_format TF32 (2) [7,10)/[10,13),
    // a_major [15], b_major [16], N>>3 [17,23), M>>4 [24,29).
    d.idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(d.a_major) << 15) |
              (static_cast<uint32_t>(d.b_major) << 16) | (static_cast<uint32_t>(p.block_n >> 3) << 17) |
              (static_cast<uint32_t>(kBlockM >> 4) << 24);
    // A: K-major -> [M][K] rows of K; MN-major -> [K][M] rows of M.
    if (!d.a_major)
      MTRL_PROPAGATE(encode_map(&P.maps[2 * i], p.A, p.K, p.M, p.lda, kBlockK, kBlockM, CU_TENSOR_MAP_SWIZZLE_128B));
    else
      MTRL_PROPAGATE(encode_map(&P.maps[2 * i], p.A, p.M, p.K, p.lda, 32, kBlockK,
                                CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
    if (!d.b_major)
      MTRL_PROPAGATE(encode_map(&P.maps[2 * i + 1], p.B, p.K, p.N, p.ldb, kBlockK, p.block_n,
                                CU_TENSOR_MAP_SWIZZLE_128B));
    else
      MTRL_PROPAGATE(encode_map(&P.maps[2 * i + 1], p.B, p.N, p.K, p.ldb, 32, kBlockK,
                                CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
  }
  P.nprob = n;
  P.total_units = units;
  int dev_id = 0, sms = 148;
  cudaGetDevice(&dev_id);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_id);
  plan->grid = units < sms ? units : sms;
  static bool attr_set = false;
  if (!attr_set) {
    MTRL_CUDA_CHECK(cudaFuncSetAttribute(gemm_tf32_grouped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kSmemBytes));
    attr_set = true;
  }
  guard.p = nullptr;
  *out = plan;
  return MTRL_OK;
}

extern "C" int mtrl_gemm_plan_run(mtrl_gemm_plan_t* plan, void* stream) {
  MTRL_REQUIRE(plan, "mtrl_gemm_plan_run: null plan");
  gemm_tf32_grouped_kernel<<<plan->grid, kThreads, kSmemBytes, static_cast<cudaStream_t>(stream)>>>(plan->params);
  MTRL_CUDA_CHECK(cudaGetLastError());
  return MTRL_OK;
}

extern "C" int mtrl_gemm_plan_units(const mtrl_gemm_plan_t* plan) { return plan ? plan->params.total_units : 0; }

extern "C" void mtrl_gemm_plan_destroy(mtrl_gemm_plan_t* plan) {
  delete plan;
}
