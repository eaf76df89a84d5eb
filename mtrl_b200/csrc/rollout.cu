// Generalised advantage estimation over a rollout: MultiTaskRolloutBuffer.get(compute_advantages=True)
// (/root/reference/mtrl/rl/buffers.py:650-707; the recurrence it adapts is openai/baselines ppo2/runner.py).
// Arrays keep the reference layout (timestep, task, 1).  The recurrence is sequential in time and independent across
// tasks: one thread per task walks the timesteps backwards in batches of 16 so that the four loads of a batch are in
// flight together (tasks are adjacent in memory: a warp's loads coalesce), and evaluates exactly the float32
// operation order NumPy uses -- results are bit-identical to the CPU restatement (oracle/rollout_oracle.py).
#include "common.cuh"
#include "mtrl_b200.h"

namespace {

constexpr int kBatch = 16;

__global__ void gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values, const float* __restrict__ dones,
                           const float* __restrict__ last_values, const float* __restrict__ last_dones, int S, int T, float gamma,
                           float lam, float* __restrict__ adv, float* __restrict__ ret) {
  const int task = blockIdx.x * blockDim.x + threadIdx.x;
  if (task >= T) return;
  float last = 0.f;   // last_gae_lamda
  for (int hi = S - 1; hi >= 0; hi -= kBatch) {
    float r[kBatch], v[kBatch], nv[kBatch], nd[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int t = hi - u;
      if (t < 0) break;
      const long long i = static_cast<long long>(t) * T + task;
      r[u] = rewards[i];
      v[u] = values[i];
      // buffers.py:676-681 with the last step's `self.dones` read as the `dones` argument
      nv[u] = t == S - 1 ? last_values[task] : values[i + T];
      nd[u] = t == S - 1 ? last_dones[task] : dones[i + T];
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int t = hi - u;
      if (t < 0) break;
      const long long i = static_cast<long long>(t) * T + task;
      // NumPy evaluates left to right in float32, no contraction:
      //   delta = rewards + next_nonterminal * gamma * next_values - values                       (:682-686)
      //   adv   = delta + next_nonterminal * gamma * gae_lambda * last_gae_lamda                  (:687-689)
      const float nn = __fsub_rn(1.0f, nd[u]);
      const float ng = __fmul_rn(nn, gamma);
      const float delta = __fsub_rn(__fadd_rn(r[u], __fmul_rn(ng, nv[u])), v[u]);
      last = __fadd_rn(delta, __fmul_rn(__fmul_rn(ng, lam), last));
      adv[i] = last;
      ret[i] = __fadd_rn(last, v[u]);   // returns = advantages + values (:690)
    }
  }
}

}  // namespace

extern "C" int mtrl_gae(const float* rewards, const float* values, const float* dones, const float* last_values,
                        const float* last_dones, int num_steps, int num_tasks, float gamma, float gae_lambda, float* advantages,
                        float* returns, void* stream) {
  MTRL_REQUIRE(rewards && values && dones && last_values && last_dones && advantages && returns, "mtrl_gae: null argument");
  MTRL_REQUIRE(num_steps >= 1 && num_tasks >= 1, "mtrl_gae: empty rollout (%d steps, %d tasks)", num_steps, num_tasks);
  const int threads = 32;   // few tasks: spread the warps over SMs
  gae_kernel<<<(num_tasks + threads - 1) / threads, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      rewards, values, dones, last_values, last_dones, num_steps, num_tasks, gamma, gae_lambda, advantages, returns);
  MTRL_CUDA_CHECK(cudaGetLastError());
  return MTRL_OK;
}
