// Multi-task replay sampler on the device: the index stream of numpy's
// `Generator(PCG64).integers(0, high, size=n)` reproduced bit for bit, and the fancy-index gathers of
// MultiTaskReplayBuffer.sample (/root/reference/mtrl/rl/buffers.py:494-549) as vectorised slab copies.
//
// Storage keeps the reference layout (capacity, T, dim) fp32 (buffers.py:293-306), so one sampled
// index is one contiguous T*dim slab per array and the (sample, task)-interleaved output order of
// buffers.py:547-548 equals the source order: the gather is a set of contiguous copies.
#include "common.cuh"
#include "mtrl_b200.h"

namespace {

struct PcgState {
  unsigned long long state_hi, state_lo, inc_hi, inc_lo;
  unsigned int has_uint32, uinteger;
  unsigned int pad0, pad1;
};

// PCG_DEFAULT_MULTIPLIER_128 = 0x2360ED051FC65DA4'4385DF649FCCF645
constexpr unsigned long long kMulHi = 0x2360ED051FC65DA4ull;
constexpr unsigned long long kMulLo = 0x4385DF649FCCF645ull;

__device__ __forceinline__ unsigned long long pcg_next64(PcgState& s) {
  // state = state * MULT + inc (mod 2^128)
  const unsigned long long lo = s.state_lo * kMulLo;
  unsigned long long hi = __umul64hi(s.state_lo, kMulLo) + s.state_lo * kMulHi + s.state_hi * kMulLo;
  const unsigned long long nlo = lo + s.inc_lo;
  hi += s.inc_hi + (nlo < lo ? 1ull : 0ull);
  s.state_lo = nlo;
  s.state_hi = hi;
  // XSL-RR: rotr64(hi ^ lo, hi >> 58)
  const unsigned long long x = hi ^ nlo;
  const unsigned int rot = static_cast<unsigned int>(hi >> 58);
  return (x >> rot) | (x << ((64u - rot) & 63u));
}

__device__ __forceinline__ unsigned int pcg_next32(PcgState& s) {
  // pcg64_next32: low half first; the high half is buffered and survives across calls
  if (s.has_uint32) {
    s.has_uint32 = 0;
    return s.uinteger;
  }
  const unsigned long long n = pcg_next64(s);
  s.has_uint32 = 1;
  s.uinteger = static_cast<unsigned int>(n >> 32);
  return static_cast<unsigned int>(n & 0xffffffffull);
}

// ---- 128-bit helpers for jumping the LCG ahead (PCG's advance(): state_k = M_k * state_0 + P_k) ----
struct U128 {
  unsigned long long hi, lo;
};
__device__ __forceinline__ U128 mul128(U128 a, U128 b) {
  U128 r;
  r.lo = a.lo * b.lo;
  r.hi = __umul64hi(a.lo, b.lo) + a.lo * b.hi + a.hi * b.lo;
  return r;
}
__device__ __forceinline__ U128 add128(U128 a, U128 b) {
  U128 r;
  r.lo = a.lo + b.lo;
  r.hi = a.hi + b.hi + (r.lo < a.lo ? 1ull : 0ull);
  return r;
}
// (M, P) with state_{+delta} = M * state + P, O(log delta) (Brown, "Random number generation with arbitrary strides")
__device__ __forceinline__ void lcg_jump(U128 inc, unsigned int delta, U128* M, U128* P) {
  U128 acc_m = {0ull, 1ull}, acc_p = {0ull, 0ull};
  U128 cur_m = {kMulHi, kMulLo}, cur_p = inc;
  while (delta) {
    if (delta & 1u) {
      acc_m = mul128(acc_m, cur_m);
      acc_p = add128(mul128(acc_p, cur_m), cur_p);
    }
    cur_p = mul128(add128(cur_m, U128{0ull, 1ull}), cur_p);
    cur_m = mul128(cur_m, cur_m);
    delta >>= 1;
  }
  *M = acc_m;
  *P = acc_p;
}
__device__ __forceinline__ unsigned long long xsl_rr(U128 s) {
  const unsigned long long x = s.hi ^ s.lo;
  const unsigned int rot = static_cast<unsigned int>(s.hi >> 58);
  return (x >> rot) | (x << ((64u - rot) & 63u));
}

// The sequential definition (one thread): Lemire rejection consumes a data-dependent number of draws.
__device__ void draw_indices_sequential(PcgState* __restrict__ st, long long* __restrict__ out, int n, unsigned int high) {
  PcgState s = *st;
  const unsigned int thr = static_cast<unsigned int>((0x100000000ull - high) % high);
  for (int i = 0; i < n; ++i) {
    unsigned long long m = static_cast<unsigned long long>(pcg_next32(s)) * high;
    unsigned int left = static_cast<unsigned int>(m);
    if (left < high) {
      while (left < thr) {
        m = static_cast<unsigned long long>(pcg_next32(s)) * high;
        left = static_cast<unsigned int>(m);
      }
    }
    out[i] = static_cast<long long>(m >> 32);
  }
  *st = s;
}

// One warp.  The 32-bit stream numpy consumes is [buffered half, if any] ++ (low, high) halves of successive 64-bit
// outputs, and draw i takes stream element i unless a Lemire rejection occurred before it (probability
// (2^32 mod high) / 2^32 < 2.4e-5 per draw for ring sizes <= 1e5).  Every lane therefore jumps the 128-bit LCG straight
// to the outputs its draws need (lcg_jump, O(log n)), and only if some lane sees a rejection does lane 0 redo the call
// sequentially from the untouched state.  Same stream and same final state bit for bit, ~2 us instead of ~12.
__global__ void draw_indices_kernel(PcgState* __restrict__ st, long long* __restrict__ out, int n,
                                    unsigned int high) {
  MTRL_PDL_PROLOGUE();
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  if (high <= 1u) {  // rng == 0: numpy fills with `low` and consumes nothing
    for (int i = lane; i < n; i += 32) out[i] = 0;
    return;
  }
  const PcgState s0 = *st;
  const U128 inc = {s0.inc_hi, s0.inc_lo};
  const U128 base = {s0.state_hi, s0.state_lo};
  const unsigned int thr = static_cast<unsigned int>((0x100000000ull - high) % high);
  const int has0 = s0.has_uint32 ? 1 : 0;
  U128 M16, P16;
  lcg_jump(inc, 16u, &M16, &P16);
  bool rejected = false;
  U128 cur = base;
  bool have_cur = false;
  for (int i = lane; i < n; i += 32) {
    unsigned int v;
    if (i == 0 && has0) {
      v = s0.uinteger;
    } else {
      const int q = i - has0;            // index into the fresh 32-bit stream
      if (!have_cur) {
        U128 M, P;
        lcg_jump(inc, static_cast<unsigned int>(q >> 1) + 1u, &M, &P);   // 64-bit output #j is produced by state_{j+1}
        cur = add128(mul128(M, base), P);
        have_cur = true;
      } else {
        cur = add128(mul128(M16, cur), P16);                             // i += 32  =>  j += 16, same half
      }
      const unsigned long long o = xsl_rr(cur);
      v = (q & 1) ? static_cast<unsigned int>(o >> 32) : static_cast<unsigned int>(o & 0xffffffffull);
    }
    const unsigned long long m = static_cast<unsigned long long>(v) * high;
    const unsigned int left = static_cast<unsigned int>(m);
    if (left < high && left < thr) rejected = true;
    out[i] = static_cast<long long>(m >> 32);
  }
  if (__any_sync(0xffffffffu, rejected)) {
    __syncwarp();
    if (lane == 0) draw_indices_sequential(st, out, n, high);
    return;
  }
  if (lane == 0) {
    const int fresh = n - has0;                  // fresh 32-bit values consumed
    PcgState s = s0;
    if (fresh > 0) {
      const unsigned int total64 = static_cast<unsigned int>((fresh + 1) >> 1);
      U128 M, P;
      lcg_jump(inc, total64, &M, &P);
      const U128 fin = add128(mul128(M, base), P);
      s.state_hi = fin.hi;
      s.state_lo = fin.lo;
      s.uinteger = static_cast<unsigned int>(xsl_rr(fin) >> 32);   // the high half buffered by the last 64-bit output
      s.has_uint32 = (fresh & 1) ? 1u : 0u;
    } else {
      s.has_uint32 = has0 && n >= 1 ? 0u : s0.has_uint32;          // only the buffered half was consumed
    }
    *st = s;
  }
}

struct GatherArgs {
  const float* src[5];   // obs, actions, next_obs, dones, rewards storage (capacity, T, dim)
  float* dst[5];
  int dim[5];            // per-task row length
  int vec[5];            // vector width usable for the slab copy (4, 2 or 1 floats)
  int num_tasks;
  int norm_mode;         // 0 none; 1 rewards' = (double(r) - shift[t]) / den[t]  (array 4 only)
  const double* shift;
  const double* den;
};

template <int V>
__device__ __forceinline__ void copy_vec(float* __restrict__ d, const float* __restrict__ s, int n, int tid,
                                         int nthreads) {
  if (V == 4) {
    const float4* s4 = reinterpret_cast<const float4*>(s);
    float4* d4 = reinterpret_cast<float4*>(d);
    for (int i = tid; i < n / 4; i += nthreads) d4[i] = __ldg(s4 + i);
  } else if (V == 2) {
    const float2* s2 = reinterpret_cast<const float2*>(s);
    float2* d2 = reinterpret_cast<float2*>(d);
    for (int i = tid; i < n / 2; i += nthreads) d2[i] = __ldg(s2 + i);
  } else {
    for (int i = tid; i < n; i += nthreads) d[i] = __ldg(s + i);
  }
}

// grid = (chunks, samples, 5 arrays).  Output row (i*T + t) = storage[idx[i], t, :]  (buffers.py:540-548).
__global__ void gather_slabs_kernel(const GatherArgs a, const long long* __restrict__ idx, int chunk_elems) {
  MTRL_PDL_PROLOGUE();
  const int arr = blockIdx.z;
  const int i = blockIdx.y;
  const int slab = a.num_tasks * a.dim[arr];
  const int e0 = blockIdx.x * chunk_elems;
  if (e0 >= slab) return;
  const int n = min(chunk_elems, slab - e0);
  const long long row = idx[i];
  const float* s = a.src[arr] + row * slab + e0;
  float* d = a.dst[arr] + static_cast<long long>(i) * slab + e0;
  if (arr == 4 && a.norm_mode) {
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      const int t = (e0 + j) / a.dim[arr];
      d[j] = static_cast<float>((static_cast<double>(__ldg(s + j)) - a.shift[t]) / a.den[t]);
    }
    return;
  }
  const int v = a.vec[arr];
  if (v == 4) copy_vec<4>(d, s, n, threadIdx.x, blockDim.x);
  else if (v == 2) copy_vec<2>(d, s, n, threadIdx.x, blockDim.x);
  else copy_vec<1>(d, s, n, threadIdx.x, blockDim.x);
}

// Per-task-count path (buffers.py:496-519): rows of task t are storage[idx_t[i], t, :], tasks concatenated.
// grid = (count_t, 5); one block copies one row.
__global__ void gather_task_rows_kernel(const GatherArgs a, const long long* __restrict__ idx, int task,
                                        int out_row0) {
  const int arr = blockIdx.y;
  const int i = blockIdx.x;
  const int dim = a.dim[arr];
  const float* s = a.src[arr] + (idx[i] * a.num_tasks + task) * static_cast<long long>(dim);
  float* d = a.dst[arr] + static_cast<long long>(out_row0 + i) * dim;
  for (int j = threadIdx.x; j < dim; j += blockDim.x) d[j] = __ldg(s + j);
}

}  // namespace

struct mtrl_sampler {
  int capacity, num_tasks, obs_dim, act_dim;
  float* store[5];  // obs, actions, next_obs, dones, rewards
  PcgState* state;  // device
  long long* idx;   // device scratch for drawn indices
  int idx_cap;
};

static int dims_of(const mtrl_sampler* s, int arr) {
  switch (arr) {
    case 0: case 2: return s->obs_dim;
    case 1: return s->act_dim;
    default: return 1;
  }
}

extern "C" int mtrl_sampler_create(mtrl_sampler_t** out, int capacity, int num_tasks, int obs_dim, int act_dim,
                                   float* obs, float* actions, float* next_obs, float* dones, float* rewards) {
  MTRL_REQUIRE(out && capacity > 0 && num_tasks > 0 && obs_dim > 0 && act_dim > 0, "mtrl_sampler_create: bad shape");
  MTRL_REQUIRE(obs && actions && next_obs && dones && rewards, "mtrl_sampler_create: null storage pointer");
  mtrl_sampler* s = new mtrl_sampler();
  s->capacity = capacity;
  s->num_tasks = num_tasks;
  s->obs_dim = obs_dim;
  s->act_dim = act_dim;
  s->store[0] = obs;
  s->store[1] = actions;
  s->store[2] = next_obs;
  s->store[3] = dones;
  s->store[4] = rewards;
  // an index draw never exceeds the ring capacity (sampling more rows than that is refused like numpy's IndexError), so
  // the scratch is sized from it: sample(ndarray) may legitimately ask one task for up to 128 * T rows (buffers.py:498)
  s->idx_cap = capacity > 4096 ? capacity : 4096;
  if (cudaMalloc(&s->state, sizeof(PcgState)) != cudaSuccess ||
      cudaMalloc(&s->idx, sizeof(long long) * s->idx_cap) != cudaSuccess) {
    mtrl_set_error("mtrl_sampler_create: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
    delete s;
    return MTRL_ERR_CUDA;
  }
  cudaMemset(s->state, 0, sizeof(PcgState));
  *out = s;
  return MTRL_OK;
}

extern "C" void mtrl_sampler_destroy(mtrl_sampler_t* s) {
  if (!s) return;
  cudaFree(s->state);
  cudaFree(s->idx);
  delete s;
}

// state4 = {state_hi, state_lo, inc_hi, inc_lo} of numpy's PCG64 state dict (Appendix B of SURVEY.md).
extern "C" int mtrl_sampler_set_state(mtrl_sampler_t* s, const uint64_t* state4, uint32_t has_uint32,
                                      uint32_t uinteger, void* stream) {
  MTRL_REQUIRE(s && state4, "mtrl_sampler_set_state: null argument");
  PcgState h;
  memset(&h, 0, sizeof(h));
  h.state_hi = state4[0];
  h.state_lo = state4[1];
  h.inc_hi = state4[2];
  h.inc_lo = state4[3];
  h.has_uint32 = has_uint32;
  h.uinteger = uinteger;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MTRL_CUDA_CHECK(cudaMemcpyAsync(s->state, &h, sizeof(h), cudaMemcpyHostToDevice, st));
  MTRL_CUDA_CHECK(cudaStreamSynchronize(st));  // h is a stack object
  return MTRL_OK;
}

// Synchronises `stream` (checkpointing only).
extern "C" int mtrl_sampler_get_state(mtrl_sampler_t* s, uint64_t* state4, uint32_t* has_uint32, uint32_t* uinteger,
                                      void* stream) {
  MTRL_REQUIRE(s && state4 && has_uint32 && uinteger, "mtrl_sampler_get_state: null argument");
  PcgState h;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MTRL_CUDA_CHECK(cudaMemcpyAsync(&h, s->state, sizeof(h), cudaMemcpyDeviceToHost, st));
  MTRL_CUDA_CHECK(cudaStreamSynchronize(st));
  state4[0] = h.state_hi;
  state4[1] = h.state_lo;
  state4[2] = h.inc_hi;
  state4[3] = h.inc_lo;
  *has_uint32 = h.has_uint32;
  *uinteger = h.uinteger;
  return MTRL_OK;
}

// MultiTaskReplayBuffer.add (buffers.py:453-457): write one (T, dim) row per array at ring position `pos`.
// Sources may be host or device pointers (cudaMemcpyDefault), each (T, dim) contiguous fp32.
extern "C" int mtrl_sampler_add(mtrl_sampler_t* s, int pos, const float* obs, const float* actions,
                                const float* next_obs, const float* dones, const float* rewards, void* stream) {
  MTRL_REQUIRE(s && pos >= 0 && pos < s->capacity, "mtrl_sampler_add: pos %d outside [0, %d)", pos,
               s ? s->capacity : 0);
  const float* src[5] = {obs, actions, next_obs, dones, rewards};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int a = 0; a < 5; ++a) {
    MTRL_REQUIRE(src[a], "mtrl_sampler_add: null source %d", a);
    const size_t row = static_cast<size_t>(s->num_tasks) * dims_of(s, a);
    MTRL_CUDA_CHECK(cudaMemcpyAsync(s->store[a] + static_cast<size_t>(pos) * row, src[a], row * sizeof(float),
                                    cudaMemcpyDefault, st));
  }
  return MTRL_OK;
}

static int fill_gather_args(const mtrl_sampler* s, GatherArgs* g, float* const* outs, int norm_mode,
                            const double* shift, const double* den) {
  for (int a = 0; a < 5; ++a) {
    MTRL_REQUIRE(outs[a], "sampler: null output %d", a);
    g->src[a] = s->store[a];
    g->dst[a] = outs[a];
    g->dim[a] = dims_of(s, a);
    const long long slab = static_cast<long long>(s->num_tasks) * g->dim[a];
    const bool al16 = ((reinterpret_cast<uintptr_t>(g->src[a]) | reinterpret_cast<uintptr_t>(g->dst[a])) & 15u) == 0;
    const bool al8 = ((reinterpret_cast<uintptr_t>(g->src[a]) | reinterpret_cast<uintptr_t>(g->dst[a])) & 7u) == 0;
    g->vec[a] = (slab % 4 == 0 && al16) ? 4 : ((slab % 2 == 0 && al8) ? 2 : 1);
  }
  g->num_tasks = s->num_tasks;
  g->norm_mode = norm_mode;
  g->shift = shift;
  g->den = den;
  MTRL_REQUIRE(!norm_mode || (shift && den), "sampler: reward normalisation needs shift and den arrays");
  return MTRL_OK;
}

// MultiTaskReplayBuffer.sample(int) (buffers.py:520-549).  fill = pos if not full else capacity.
// Outputs are (n_per_task*T, dim) device arrays in the field order of ReplayBufferSamples.
// idx_out (device int64[n_per_task]) is optional.
extern "C" int mtrl_sampler_sample(mtrl_sampler_t* s, int fill, int n_per_task, long long* idx_out, float* obs_out,
                                   float* actions_out, float* next_obs_out, float* dones_out, float* rewards_out,
                                   int norm_mode, const double* shift, const double* den, void* stream) {
  MTRL_REQUIRE(s && n_per_task > 0 && n_per_task <= s->idx_cap, "mtrl_sampler_sample: n_per_task %d outside (0, %d]",
               n_per_task, s ? s->idx_cap : 0);
  MTRL_REQUIRE(fill >= 0 && fill <= s->capacity, "mtrl_sampler_sample: fill %d outside [0, %d]", fill, s->capacity);
  const int high = fill > n_per_task ? fill : n_per_task;  // buffers.py:525
  // numpy would raise IndexError on self.rewards[sample_idx]; an out-of-range gather is refused up front.
  MTRL_REQUIRE(high <= s->capacity, "index out of bounds: sampling %d per task from a buffer of capacity %d", n_per_task,
               s->capacity);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* outs[5] = {obs_out, actions_out, next_obs_out, dones_out, rewards_out};
  GatherArgs g;
  MTRL_PROPAGATE(fill_gather_args(s, &g, outs, norm_mode, shift, den));
  MTRL_CUDA_CHECK(mtrl_launch(draw_indices_kernel, dim3(1), dim3(32), 0, st, s->state, s->idx, n_per_task, static_cast<unsigned int>(high)));
  if (idx_out)
    MTRL_CUDA_CHECK(cudaMemcpyAsync(idx_out, s->idx, sizeof(long long) * n_per_task, cudaMemcpyDeviceToDevice, st));
  const int threads = 256;
  const int chunk = threads * 4 * 4;  // 4 float4 per thread per block
  const int max_slab = s->num_tasks * s->obs_dim;
  dim3 grid((max_slab + chunk - 1) / chunk, n_per_task, 5);
  MTRL_CUDA_CHECK(mtrl_launch(gather_slabs_kernel, grid, dim3(threads), 0, st, g, s->idx, chunk));
  return MTRL_OK;
}

// `self._rng.integers(low=0, high=high, size=n)` alone (buffers.py:523-527): n int64 draws into idx_out (device),
// advancing the generator exactly as numpy does.  1 <= high <= 2^32.
extern "C" int mtrl_sampler_draw(mtrl_sampler_t* s, unsigned long long high, int n, long long* idx_out, void* stream) {
  MTRL_REQUIRE(s && idx_out, "mtrl_sampler_draw: null argument");
  MTRL_REQUIRE(n > 0 && n <= s->idx_cap, "mtrl_sampler_draw: n %d outside (0, %d]", n, s->idx_cap);
  MTRL_REQUIRE(high >= 1ull && high <= 0xffffffffull, "mtrl_sampler_draw: high outside [1, 2^32 - 1] (numpy's 32-bit path)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  draw_indices_kernel<<<1, 32, 0, st>>>(s->state, s->idx, n, static_cast<unsigned int>(high));
  MTRL_CUDA_CHECK(cudaGetLastError());
  MTRL_CUDA_CHECK(cudaMemcpyAsync(idx_out, s->idx, sizeof(long long) * n, cudaMemcpyDeviceToDevice, st));
  return MTRL_OK;
}

// MultiTaskReplayBuffer.sample(ndarray) (buffers.py:496-519): independent draws per task, rows concatenated
// by task.  counts is a host int[T].  Reward normalisation is not applied on this path (as in the reference).
extern "C" int mtrl_sampler_sample_per_task(mtrl_sampler_t* s, int fill, const int* counts, float* obs_out,
                                            float* actions_out, float* next_obs_out, float* dones_out,
                                            float* rewards_out, void* stream) {
  MTRL_REQUIRE(s && counts, "mtrl_sampler_sample_per_task: null argument");
  MTRL_REQUIRE(fill >= 0 && fill <= s->capacity, "mtrl_sampler_sample_per_task: fill %d outside [0, %d]", fill,
               s->capacity);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* outs[5] = {obs_out, actions_out, next_obs_out, dones_out, rewards_out};
  GatherArgs g;
  MTRL_PROPAGATE(fill_gather_args(s, &g, outs, 0, nullptr, nullptr));
  int row0 = 0;
  for (int t = 0; t < s->num_tasks; ++t) {
    const int n = counts[t];
    MTRL_REQUIRE(n >= 0 && n <= s->idx_cap, "mtrl_sampler_sample_per_task: count %d for task %d", n, t);
    if (n == 0) continue;
    const int high = fill > n ? fill : n;  // buffers.py:505
    MTRL_REQUIRE(high <= s->capacity, "index out of bounds: sampling %d rows of task %d from capacity %d", n, t,
                 s->capacity);
    draw_indices_kernel<<<1, 32, 0, st>>>(s->state, s->idx, n, static_cast<unsigned int>(high));
    gather_task_rows_kernel<<<dim3(n, 5), 128, 0, st>>>(g, s->idx, t, row0);
    MTRL_CUDA_CHECK(cudaGetLastError());
    row0 += n;
  }
  return MTRL_OK;
}
