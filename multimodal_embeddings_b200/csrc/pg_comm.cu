// pg_comm.cu — K6: the one exchange step of the path.  Per-rank INTEGER corpus histograms (plain_text widths
// in 1-px bins, column centres in per-mille bins; accumulated by K4/K5) are summed over the ranks with one
// ncclAllReduce(ncclSum) over NVLink 5 / NVSwitch (no reference analogue — the reference is a
// single process).  Integer sums are order-independent, so the totals are bit-identical for 1/2/4/8 GPUs.
//
// NCCL is bound at run time (dlopen), not at link time: the library must load on a machine without NCCL (the
// CPU test-suite checks its exports), and a host process that already carries an NCCL (PyTorch's) shares it.
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "pg_common.cuh"

namespace {
// the part of NCCL's ABI used here (stable since NCCL 2.0: nccl.h)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;                    // ncclSuccess = 0
constexpr int kNcclUint32 = 3, kNcclSum = 0;  // ncclDataType_t / ncclRedOp_t values

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};

NcclApi* nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the host process's own copy, if it has one
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) return;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(h, "ncclAllReduce"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(dlsym(h, "ncclGetVersion"));
    if (api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce) api.handle = h;
  });
  return api.handle ? &api : nullptr;
}

int nccl_fail(const char* what, ncclResult_t rc) {
  NcclApi* n = nccl();
  pg_set_error("NCCL error in %s: %s", what, (n && n->GetErrorString) ? n->GetErrorString(rc) : "?");
  return PG_ERR_CUDA;
}
}  // namespace

struct PgComm {
  ncclComm_t comm = nullptr;
  int32_t world = 1, rank = 0;
};

extern "C" int pg_comm_nccl_version(void) {
  NcclApi* n = nccl();
  int v = 0;
  if (!n || !n->GetVersion || n->GetVersion(&v) != 0) return 0;
  return v;
}

extern "C" int pg_comm_unique_id(uint8_t id[128]) {
  PG_REQUIRE(id != nullptr, "id");
  NcclApi* n = nccl();
  if (!n) {
    pg_set_error("unsupported: libnccl.so.2 not found (needed only for the corpus-histogram exchange)");
    return PG_ERR_UNSUPPORTED;
  }
  ncclUniqueId u;
  const ncclResult_t rc = n->GetUniqueId(&u);
  if (rc != 0) return nccl_fail("ncclGetUniqueId", rc);
  std::memcpy(id, u.internal, 128);
  return PG_OK;
}

extern "C" int pg_comm_create(const uint8_t id[128], int32_t world, int32_t rank, PgComm** out) {
  PG_REQUIRE(id != nullptr && out != nullptr && world >= 1 && rank >= 0 && rank < world, "communicator arguments");
  NcclApi* n = nccl();
  if (!n) {
    pg_set_error("unsupported: libnccl.so.2 not found (needed only for the corpus-histogram exchange)");
    return PG_ERR_UNSUPPORTED;
  }
  ncclUniqueId u;
  std::memcpy(u.internal, id, 128);
  auto* c = new PgComm();
  c->world = world;
  c->rank = rank;
  const ncclResult_t rc = n->CommInitRank(&c->comm, world, u, rank);  // collective: every rank calls it
  if (rc != 0) {
    delete c;
    return nccl_fail("ncclCommInitRank", rc);
  }
  *out = c;
  return PG_OK;
}

extern "C" void pg_comm_destroy(PgComm* c) {
  if (!c) return;
  NcclApi* n = nccl();
  if (n && c->comm) n->CommDestroy(c->comm);
  delete c;
}

extern "C" void* pg_comm_nccl(PgComm* c) { return c ? (void*)c->comm : nullptr; }

extern "C" int pg_hist_allreduce(uint32_t* hist, size_t n_bins, void* nccl_comm, void* stream) {
  PG_REQUIRE(hist != nullptr && nccl_comm != nullptr, "hist / communicator");
  if (n_bins == 0) return PG_OK;
  NcclApi* n = nccl();
  if (!n) {
    pg_set_error("unsupported: libnccl.so.2 not found");
    return PG_ERR_UNSUPPORTED;
  }
  const ncclResult_t rc = n->AllReduce(hist, hist, n_bins, kNcclUint32, kNcclSum, (ncclComm_t)nccl_comm, (cudaStream_t)stream);
  if (rc != 0) return nccl_fail("ncclAllReduce", rc);
  return PG_OK;
}


// ---- pinned host staging ---------------------------------------------------------------------------------------
// Page-locked host memory for the buffers that cross PCIe every step (the scans' file bytes).  write_combined != 0
// asks for cudaHostAllocWriteCombined: the CPU fills it with streaming stores and never reads it back, and device
// reads of it do not snoop the CPU caches — worth having when several GPUs pull from host memory at once.
extern "C" int pg_pinned_alloc(size_t bytes, int32_t write_combined, void** out) {
  PG_REQUIRE(out && bytes > 0, "pinned alloc arguments");
  *out = nullptr;
  PG_CUDA_TRY(cudaHostAlloc(out, bytes, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
  return PG_OK;
}
extern "C" int pg_pinned_free(void* p) {
  if (p) PG_CUDA_TRY(cudaFreeHost(p));
  return PG_OK;
}
