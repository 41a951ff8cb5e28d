// Data-parallel plumbing: one NCCL communicator per process (one process per GPU),
// used for the sum all-reduce of dW / bias grads over NVLink 5 / NVSwitch.
// The reference has no communication layer; the
// contract is "all-reduced dW == single-process reference dW on the full batch".
//
// NCCL is bound at first use with dlopen so that libmgb200.so itself loads on a
// machine without NCCL or without a GPU.
#include <dlfcn.h>
#include <string.h>

#include "common.cuh"

namespace mgb {
namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat32 = 7, ncclFloat64 = 8 };
enum { ncclSum = 0 };

struct Nccl {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

Nccl g_nccl;
ncclComm_t g_comm = nullptr;
int g_world = 1, g_rank = 0;

int load_nccl() {
  if (g_nccl.handle) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return fail("NCCL not found (dlopen libnccl.so.2): %s", dlerror());
#define BIND(field, sym)                                                  \
  *(void**)(&g_nccl.field) = dlsym(h, sym);                               \
  if (!g_nccl.field) return fail("NCCL symbol %s missing", sym);
  BIND(GetUniqueId, "ncclGetUniqueId");
  BIND(CommInitRank, "ncclCommInitRank");
  BIND(CommDestroy, "ncclCommDestroy");
  BIND(AllReduce, "ncclAllReduce");
  BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
  g_nccl.handle = h;
  return 0;
}

#define MGB_NCCL(call)                                                                  \
  do {                                                                                  \
    ncclResult_t r__ = (call);                                                          \
    if (r__ != 0) return fail("%s -> NCCL error %d: %s", #call, r__, g_nccl.GetErrorString(r__)); \
  } while (0)

}  // namespace
}  // namespace mgb

using namespace mgb;

extern "C" int mgb_comm_unique_id(void* id128) {
  if (!id128) return fail("mgb_comm_unique_id: null");
  if (load_nccl()) return 1;
  ncclUniqueId id;
  MGB_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return 0;
}

extern "C" int mgb_comm_init_rank(const void* id128, int world_size, int rank) {
  if (!id128 || world_size < 1 || rank < 0 || rank >= world_size)
    return fail("mgb_comm_init_rank: bad arguments (world=%d rank=%d)", world_size, rank);
  if (g_comm) return fail("mgb_comm_init_rank: communicator already initialised");
  if (load_nccl()) return 1;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  MGB_NCCL(g_nccl.CommInitRank(&g_comm, world_size, id, rank));
  g_world = world_size;
  g_rank = rank;
  return 0;
}

extern "C" int mgb_comm_destroy(void) {
  if (!g_comm) return 0;
  MGB_NCCL(g_nccl.CommDestroy(g_comm));
  g_comm = nullptr;
  g_world = 1;
  g_rank = 0;
  return 0;
}

extern "C" int mgb_allreduce_sum(void* buf, size_t n, int dtype, void* stream) {
  if (n == 0) return 0;
  if (!g_comm) return fail("mgb_allreduce_sum: communicator not initialised");
  if (!buf) return fail("mgb_allreduce_sum: null buffer");
  int t = dtype == MGB_F32 ? ncclFloat32 : (dtype == MGB_F64 ? ncclFloat64 : -1);
  if (t < 0) return fail("mgb_allreduce_sum: bad dtype %d", dtype);
  MGB_NCCL(g_nccl.AllReduce(buf, buf, n, t, ncclSum, g_comm, as_stream(stream)));
  return 0;
}
