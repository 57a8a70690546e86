// K11: data-parallel gradient exchange.  The reference has no collective at all (single process, single
// device — SURVEY.md §2.1); sharding the minibatch over the GPUs of a box adds exactly one exchange per
// optimizer step: a sum all-reduce of the flat gradient arena (and of a few loss partial sums), issued on
// the update stream so that it is captured into the same CUDA graph as the kernels around it.
//
// NCCL is bound at run time with dlopen (the library ships inside the PyTorch wheel; libd3b.so has no
// link-time dependency on it): only ncclGetUniqueId / ncclCommInitRank / ncclAllReduce / ncclCommDestroy.
#include <dlfcn.h>

#include "common.cuh"

namespace {

typedef struct { char internal[128]; } UniqueId;  // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef void* Comm;
typedef int (*GetUniqueIdFn)(UniqueId*);
typedef int (*CommInitRankFn)(Comm*, int, UniqueId, int);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, Comm, cudaStream_t);
typedef int (*CommDestroyFn)(Comm);
typedef const char* (*GetErrorStringFn)(int);

struct Nccl {
  void* handle = nullptr;
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  AllReduceFn all_reduce = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  GetErrorStringFn error_string = nullptr;
} g_nccl;

constexpr int kNcclFloat32 = 7;  // ncclFloat32
constexpr int kNcclSum = 0;      // ncclSum

int nccl_fail(const char* what, int rc) {
  return d3b::set_err(D3B_ERR_CUDA, "%s: NCCL error %d (%s)", what, rc,
                      g_nccl.error_string ? g_nccl.error_string(rc) : "?");
}

}  // namespace

extern "C" int d3b_comm_load(const char* libnccl_path) {
  if (g_nccl.handle) return D3B_OK;
  const char* path = (libnccl_path && libnccl_path[0]) ? libnccl_path : "libnccl.so.2";
  void* h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) return d3b::set_err(D3B_ERR_ARG, "comm_load: dlopen(%s) failed: %s", path, dlerror());
  g_nccl.get_unique_id = (GetUniqueIdFn)dlsym(h, "ncclGetUniqueId");
  g_nccl.comm_init_rank = (CommInitRankFn)dlsym(h, "ncclCommInitRank");
  g_nccl.all_reduce = (AllReduceFn)dlsym(h, "ncclAllReduce");
  g_nccl.comm_destroy = (CommDestroyFn)dlsym(h, "ncclCommDestroy");
  g_nccl.error_string = (GetErrorStringFn)dlsym(h, "ncclGetErrorString");
  if (!g_nccl.get_unique_id || !g_nccl.comm_init_rank || !g_nccl.all_reduce || !g_nccl.comm_destroy)
    return d3b::set_err(D3B_ERR_ARG, "comm_load: %s does not export the NCCL entry points", path);
  g_nccl.handle = h;
  return D3B_OK;
}

extern "C" int d3b_comm_unique_id(void* id_out_128) {
  D3B_REQUIRE(g_nccl.handle, "comm_unique_id: call d3b_comm_load first");
  D3B_REQUIRE(id_out_128, "comm_unique_id: null pointer");
  UniqueId id;
  int rc = g_nccl.get_unique_id(&id);
  if (rc) return nccl_fail("ncclGetUniqueId", rc);
  memcpy(id_out_128, &id, sizeof(id));
  return D3B_OK;
}

extern "C" int d3b_comm_init(const void* id_128, int world_size, int rank, void** comm_out) {
  D3B_REQUIRE(g_nccl.handle, "comm_init: call d3b_comm_load first");
  D3B_REQUIRE(id_128 && comm_out && world_size >= 1 && rank >= 0 && rank < world_size, "comm_init: bad arguments");
  UniqueId id;
  memcpy(&id, id_128, sizeof(id));
  Comm c = nullptr;
  int rc = g_nccl.comm_init_rank(&c, world_size, id, rank);
  if (rc) return nccl_fail("ncclCommInitRank", rc);
  *comm_out = c;
  return D3B_OK;
}

extern "C" int d3b_allreduce_sum(void* comm, float* buf, int64_t n, void* stream) {
  D3B_REQUIRE(g_nccl.handle && comm, "allreduce_sum: communicator not initialised");
  D3B_REQUIRE(n >= 0 && (buf || n == 0), "allreduce_sum: bad arguments");
  if (n == 0) return D3B_OK;
  int rc = g_nccl.all_reduce(buf, buf, (size_t)n, kNcclFloat32, kNcclSum, (Comm)comm, (cudaStream_t)stream);
  if (rc) return nccl_fail("ncclAllReduce", rc);
  d3b::count_launch();
  return D3B_OK;
}

extern "C" int d3b_comm_destroy(void* comm) {
  if (comm && g_nccl.handle) {
    int rc = g_nccl.comm_destroy((Comm)comm);
    if (rc) return nccl_fail("ncclCommDestroy", rc);
  }
  return D3B_OK;
}
