// Exchange arena: one cudaMalloc block per rank, exported to the other ranks of the box through CUDA IPC so that
// kernels can load / store peer memory over NVLink (comm.cuh).  Replaces nothing in the reference (it is single
// device); it is the multi-GPU plumbing of SURVEY 8(e).
#include <cuda.h>
#include <stdlib.h>
#include <unistd.h>

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

// ---------------------------------------------------------------------------------------------
// NVSwitch multicast ("NVLS") region for the parameter all-gather.  One multicast object spans the box; every rank binds
// a physical allocation of its own to it and maps (a) that allocation for ordinary local access and (b) the multicast
// address, where ONE multimem.st lands in all ranks' copies -- the switch replicates the store, so the owner of a trunk
// segment sends its new parameters over NVLink once instead of once per peer.  Driver entry points are looked up at run
// time (the library links no libcuda, so it still loads on a box without a driver).
// ---------------------------------------------------------------------------------------------
namespace {

template <typename Fn>
int driver_fn(const char* name, Fn* out) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || !fn || qres != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    mtrl_set_error("driver entry point %s is not available", name);
    return MTRL_ERR_UNSUPPORTED;
  }
  *out = reinterpret_cast<Fn>(fn);
  return MTRL_OK;
}

#define MTRL_DRV(name, ...)                                             \
  typedef CUresult (*name##_t)(__VA_ARGS__);                            \
  name##_t name##_p = nullptr;                                          \
  MTRL_PROPAGATE(driver_fn(#name, &name##_p))

#define MTRL_CU_CHECK(call)                                                       \
  do {                                                                            \
    const CUresult r_ = (call);                                                   \
    if (r_ != CUDA_SUCCESS) {                                                     \
      mtrl_set_error("%s failed with CUresult %d", #call, static_cast<int>(r_));  \
      return MTRL_ERR_CUDA;                                                       \
    }                                                                             \
  } while (0)

int current_device(CUdevice* dev, int* ordinal) {
  MTRL_DRV(cuDeviceGet, CUdevice*, int);
  MTRL_CUDA_CHECK(cudaGetDevice(ordinal));
  MTRL_CUDA_CHECK(cudaFree(nullptr));   // make sure the primary context exists
  MTRL_CU_CHECK(cuDeviceGet_p(dev, *ordinal));
  return MTRL_OK;
}

long long round_up_ll(long long x, long long m) { return (x + m - 1) / m * m; }

int mc_prop(const mtrl_comm* c, long long bytes, CUmulticastObjectProp* prop, size_t* gran) {
  MTRL_DRV(cuMulticastGetGranularity, size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags);
  memset(prop, 0, sizeof(*prop));
  prop->numDevices = static_cast<unsigned>(c->world);
  prop->handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  prop->size = static_cast<size_t>(bytes);
  MTRL_CU_CHECK(cuMulticastGetGranularity_p(gran, prop, CU_MULTICAST_GRANULARITY_RECOMMENDED));
  prop->size = static_cast<size_t>(round_up_ll(bytes, static_cast<long long>(*gran)));
  return MTRL_OK;
}

}  // namespace

extern "C" int mtrl_comm_mc_supported(int* out) {
  MTRL_REQUIRE(out, "mtrl_comm_mc_supported: null argument");
  *out = 0;
  MTRL_DRV(cuDeviceGetAttribute, int*, CUdevice_attribute, CUdevice);
  CUdevice dev;
  int ordinal = 0;
  MTRL_PROPAGATE(current_device(&dev, &ordinal));
  int v = 0;
  MTRL_CU_CHECK(cuDeviceGetAttribute_p(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev));
  *out = v;
  return MTRL_OK;
}

extern "C" int mtrl_comm_mc_create(mtrl_comm_t* c, long long bytes, int* fd_out) {
  MTRL_REQUIRE(c && fd_out && bytes > 0, "mtrl_comm_mc_create: bad argument");
  MTRL_REQUIRE(!c->mc_handle, "mtrl_comm_mc_create: the multicast object exists already");
  MTRL_DRV(cuMulticastCreate, CUmemGenericAllocationHandle*, const CUmulticastObjectProp*);
  MTRL_DRV(cuMemExportToShareableHandle, void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
  CUmulticastObjectProp prop;
  size_t gran = 0;
  MTRL_PROPAGATE(mc_prop(c, bytes, &prop, &gran));
  CUmemGenericAllocationHandle mc = 0;
  MTRL_CU_CHECK(cuMulticastCreate_p(&mc, &prop));
  int fd = -1;
  MTRL_CU_CHECK(cuMemExportToShareableHandle_p(&fd, mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  c->mc_handle = mc;
  c->mc_bytes = static_cast<long long>(prop.size);
  *fd_out = fd;
  return MTRL_OK;
}

extern "C" int mtrl_comm_mc_import(mtrl_comm_t* c, long long bytes, int fd) {
  MTRL_REQUIRE(c && bytes > 0 && fd >= 0, "mtrl_comm_mc_import: bad argument");
  MTRL_REQUIRE(!c->mc_handle, "mtrl_comm_mc_import: the multicast object exists already");
  MTRL_DRV(cuMemImportFromShareableHandle, CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType);
  CUmulticastObjectProp prop;
  size_t gran = 0;
  MTRL_PROPAGATE(mc_prop(c, bytes, &prop, &gran));
  CUmemGenericAllocationHandle mc = 0;
  MTRL_CU_CHECK(cuMemImportFromShareableHandle_p(&mc, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)),
                                                 CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
  c->mc_handle = mc;
  c->mc_bytes = static_cast<long long>(prop.size);
  return MTRL_OK;
}

extern "C" int mtrl_comm_mc_add_device(mtrl_comm_t* c) {
  MTRL_REQUIRE(c && c->mc_handle, "mtrl_comm_mc_add_device: no multicast object");
  MTRL_DRV(cuMulticastAddDevice, CUmemGenericAllocationHandle, CUdevice);
  CUdevice dev;
  int ordinal = 0;
  MTRL_PROPAGATE(current_device(&dev, &ordinal));
  MTRL_CU_CHECK(cuMulticastAddDevice_p(static_cast<CUmemGenericAllocationHandle>(c->mc_handle), dev));
  return MTRL_OK;
}

extern "C" int mtrl_comm_mc_bind(mtrl_comm_t* c) {
  MTRL_REQUIRE(c && c->mc_handle && !c->mc_local, "mtrl_comm_mc_bind: no multicast object, or bound already");
  MTRL_DRV(cuMemCreate, CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
  MTRL_DRV(cuMemGetAllocationGranularity, size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
  MTRL_DRV(cuMemAddressReserve, CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
  MTRL_DRV(cuMemMap, CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
  MTRL_DRV(cuMemSetAccess, CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
  MTRL_DRV(cuMulticastBindMem, CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long);
  CUdevice dev;
  int ordinal = 0;
  MTRL_PROPAGATE(current_device(&dev, &ordinal));
  const size_t size = static_cast<size_t>(c->mc_bytes);
  CUmemAllocationProp ap;
  memset(&ap, 0, sizeof(ap));
  ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ap.location.id = ordinal;
  ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t agran = 0;
  MTRL_CU_CHECK(cuMemGetAllocationGranularity_p(&agran, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  MTRL_REQUIRE(size % agran == 0, "mtrl_comm_mc_bind: multicast size %zu is not a multiple of the allocation granularity %zu", size,
               agran);
  CUmemGenericAllocationHandle mem = 0;
  MTRL_CU_CHECK(cuMemCreate_p(&mem, size, &ap, 0));
  const CUmemGenericAllocationHandle mc = static_cast<CUmemGenericAllocationHandle>(c->mc_handle);
  MTRL_CU_CHECK(cuMulticastBindMem_p(mc, 0, mem, 0, size, 0));
  CUmemAccessDesc ad;
  memset(&ad, 0, sizeof(ad));
  ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ad.location.id = ordinal;
  ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  CUdeviceptr local = 0, mcva = 0;
  MTRL_CU_CHECK(cuMemAddressReserve_p(&local, size, agran, 0, 0));
  MTRL_CU_CHECK(cuMemMap_p(local, size, 0, mem, 0));
  MTRL_CU_CHECK(cuMemSetAccess_p(local, size, &ad, 1));
  MTRL_CU_CHECK(cuMemAddressReserve_p(&mcva, size, agran, 0, 0));
  MTRL_CU_CHECK(cuMemMap_p(mcva, size, 0, mc, 0));
  MTRL_CU_CHECK(cuMemSetAccess_p(mcva, size, &ad, 1));
  c->mc_mem = mem;
  c->mc_local = reinterpret_cast<uint8_t*>(local);
  c->mc_ptr = reinterpret_cast<uint8_t*>(mcva);
  MTRL_CUDA_CHECK(cudaMemset(c->mc_local, 0, size));
  MTRL_CUDA_CHECK(cudaDeviceSynchronize());
  return MTRL_OK;
}

extern "C" void* mtrl_comm_mc_local(mtrl_comm_t* c) { return c ? c->mc_local : nullptr; }
extern "C" void* mtrl_comm_mc_ptr(mtrl_comm_t* c) { return c ? c->mc_ptr : nullptr; }
extern "C" long long mtrl_comm_mc_bytes(mtrl_comm_t* c) { return c ? c->mc_bytes : 0; }

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
