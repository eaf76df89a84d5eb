// Exchange arena: one cudaMalloc block per rank, exported to the other ranks of the box through CUDA IPC so that
// kernels can load / store peer memory over NVLink (comm.cuh).  Replaces nothing in the reference (it is single
// device); it is the multi-GPU plumbing of SURVEY 8(e).
#include <stdlib.h>

#include "comm.cuh"

static_assert(sizeof(cudaIpcMemHandle_t) == MTRL_IPC_HANDLE_BYTES, "IPC handle size");

extern "C" int mtrl_comm_create(mtrl_comm_t** out, int rank, int world, long long arena_bytes, unsigned char* handle_out) {
  MTRL_REQUIRE(out && handle_out, "mtrl_comm_create: null argument");
  MTRL_REQUIRE(world >= 1 && world <= MTRL_COMM_MAX_RANKS && rank >= 0 && rank < world,
               "mtrl_comm_create: rank %d / world %d outside [1, %d]", rank, world, MTRL_COMM_MAX_RANKS);
  MTRL_REQUIRE(arena_bytes >= MTRL_COMM_HEADER_BYTES, "mtrl_comm_create: arena smaller than its %d-byte header",
               MTRL_COMM_HEADER_BYTES);
  mtrl_comm* c = new mtrl_comm();
  c->rank = rank;
  c->world = world;
  c->arena_bytes = arena_bytes;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, static_cast<size_t>(arena_bytes));
  if (e != cudaSuccess) {
    delete c;
    mtrl_set_error("mtrl_comm_create: cudaMalloc(%lld) failed: %s", arena_bytes, cudaGetErrorString(e));
    return MTRL_ERR_CUDA;
  }
  c->arena = static_cast<uint8_t*>(p);
  c->peer[rank] = c->arena;
  cudaMemset(p, 0, static_cast<size_t>(arena_bytes));
  comm::Header h;
  memset(&h, 0, sizeof(h));
  h.epoch = 1;
  h.sync_epoch = 1;
  double timeout_s = 30.0;
  if (const char* e = getenv("MTRL_COMM_TIMEOUT_S")) timeout_s = atof(e) > 0.0 ? atof(e) : timeout_s;
  h.timeout_ns = static_cast<unsigned long long>(timeout_s * 1e9);
  cudaMemcpy(p, &h, sizeof(h), cudaMemcpyHostToDevice);
  cudaIpcMemHandle_t ih;
  e = cudaIpcGetMemHandle(&ih, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    delete c;
    mtrl_set_error("mtrl_comm_create: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return MTRL_ERR_CUDA;
  }
  memcpy(handle_out, &ih, sizeof(ih));
  cudaDeviceSynchronize();
  *out = c;
  return MTRL_OK;
}

extern "C" void* mtrl_comm_arena(mtrl_comm_t* c) { return c ? c->arena : nullptr; }

extern "C" int mtrl_comm_open_peers(mtrl_comm_t* c, const unsigned char* handles) {
  MTRL_REQUIRE(c && handles, "mtrl_comm_open_peers: null argument");
  MTRL_REQUIRE(!c->opened, "mtrl_comm_open_peers: already opened");
  for (int q = 0; q < c->world; ++q) {
    if (q == c->rank) continue;
    cudaIpcMemHandle_t ih;
    memcpy(&ih, handles + static_cast<size_t>(q) * MTRL_IPC_HANDLE_BYTES, sizeof(ih));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      mtrl_set_error("mtrl_comm_open_peers: cudaIpcOpenMemHandle(rank %d) failed: %s", q, cudaGetErrorString(e));
      return MTRL_ERR_CUDA;
    }
    c->peer[q] = static_cast<uint8_t*>(p);
  }
  comm::Header* hp[MTRL_COMM_MAX_RANKS] = {};
  for (int q = 0; q < c->world; ++q) hp[q] = reinterpret_cast<comm::Header*>(c->peer[q]);
  MTRL_CUDA_CHECK(cudaMalloc(&c->d_peer_hdr, sizeof(hp)));
  MTRL_CUDA_CHECK(cudaMemcpy(c->d_peer_hdr, hp, sizeof(hp), cudaMemcpyHostToDevice));
  c->opened = true;
  return MTRL_OK;
}

extern "C" int mtrl_comm_error(mtrl_comm_t* c, int* code) {
  MTRL_REQUIRE(c && code, "mtrl_comm_error: null argument");
  comm::Header h;
  MTRL_CUDA_CHECK(cudaMemcpy(&h, c->arena, sizeof(h), cudaMemcpyDeviceToHost));
  *code = h.error;
  return MTRL_OK;
}

extern "C" int mtrl_comm_phase_times(mtrl_comm_t* c, double* us14) {
  MTRL_REQUIRE(c && us14, "mtrl_comm_phase_times: null argument");
  comm::Header h;
  MTRL_CUDA_CHECK(cudaMemcpy(&h, c->arena, sizeof(h), cudaMemcpyDeviceToHost));
  // the update runs critic then actor: the older stamp set is the critic's
  const int first = h.phase_ns[0][0] <= h.phase_ns[1][0] ? 0 : 1;
  for (int k = 0; k < 2; ++k) {
    const unsigned long long* t = h.phase_ns[k == 0 ? first : 1 - first];
    for (int i = 0; i < 6; ++i) us14[k * 7 + i] = t[i + 1] >= t[i] ? (t[i + 1] - t[i]) * 1e-3 : -1.0;
    us14[k * 7 + 6] = t[6] >= t[0] ? (t[6] - t[0]) * 1e-3 : -1.0;
  }
  return MTRL_OK;
}

extern "C" void mtrl_comm_destroy(mtrl_comm_t* c) {
  if (!c) return;
  for (int q = 0; q < c->world; ++q)
    if (q != c->rank && c->peer[q]) cudaIpcCloseMemHandle(c->peer[q]);
  if (c->d_peer_hdr) cudaFree(c->d_peer_hdr);
  if (c->arena) cudaFree(c->arena);
  delete c;
}
