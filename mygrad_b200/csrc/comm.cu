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

// ---- peer-memory all-reduce: OUR collective over NVLink / NVSwitch -----------------------------
//
// dW (and every broadcast-parameter grad) is a few hundred KB: an all-reduce of that size is pure
// latency, and it always follows a small reduction of our own (the split-K partials of the wgrad
// GEMM).  peer_reduce_allreduce does both in ONE kernel: every rank owns a cudaMalloc'ed region
// (flags + two data slots) that all other ranks map with CUDA IPC; CTA c of rank r
//   A. sums its chunk of the local split-K partials (float64, fixed order) into rank r's slot,
//   B. publishes "chunk c of rank r, epoch e" into every peer's flag array (release, system scope)
//      and waits until its own flag array shows epoch e for chunk c from every rank,
//   C. reads chunk c of every rank's slot over NVLink (peer loads) and adds them in RANK ORDER in
//      float64 — every rank computes bit-identical sums — and writes the result.
// Chunks only depend on the same chunk index on the other GPUs, so there is no grid-wide barrier;
// slots alternate with the epoch parity: a rank can only start epoch e+2 after every peer has
// published e+1, i.e. finished reading epoch e (kernels of one rank run in stream order).
//
// Forward progress and failure behaviour:
//   * the kernel is launched COOPERATIVELY: all of its CTAs are co-resident or the launch does not start, so a
//     CTA never waits for a peer CTA that cannot be scheduled because other work holds the SMs;
//   * every call must be issued on ONE stream, in the same order on all ranks (epoch / slot parity are
//     host-side counters): a call on another stream is refused;
//   * a peer that never arrives (it failed before the collective, or was killed) no longer traps the
//     context: after kPeerTimeoutNs the waiting CTAs give up and raise a flag in mapped host memory; the
//     result of that call is undefined and every later call — and mgb_peer_status — reports the time-out.

constexpr int kPeerMaxWorld = 16;
constexpr int kPeerChunks = 128;            // CTAs per call == flag slots per source rank
constexpr size_t kPeerFlagBytes = (size_t)kPeerMaxWorld * kPeerChunks * sizeof(uint32_t);

struct PeerTable {
  char* base[kPeerMaxWorld];                // region of each rank as mapped into THIS process
};

constexpr unsigned long long kPeerTimeoutNs = 30ull * 1000 * 1000 * 1000;

struct PeerState {
  bool ready = false;
  int rank = 0, world = 1;
  size_t slot_bytes = 0;
  char* local = nullptr;
  PeerTable table{};
  uint32_t epoch = 0;
  unsigned int* err_host = nullptr;       // mapped, pinned: [0] = epoch of the first timed-out call (0 = none), [1] = rank waited for
  unsigned int* err_dev = nullptr;
  bool stream_known = false;
  cudaStream_t stream = nullptr;
};
PeerState g_peer;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// 16-byte load from (possibly peer) memory at system scope: never served from a stale L1 line
__device__ __forceinline__ float4 ld_peer16(const void* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];\n"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}

template <typename T>
__global__ void __launch_bounds__(256) peer_reduce_allreduce_kernel(const T* partial, int splits,
                                                                    long long n, T* out,
                                                                    PeerTable tab, int rank, int world,
                                                                    size_t slot_off, uint32_t epoch,
                                                                    unsigned int* err) {
  constexpr int V = 16 / (int)sizeof(T);                  // elements per 16-byte pack
  const int c = blockIdx.x;
  // chunks are whole packs, so that every remote access is one aligned 16-byte transaction
  const long long packs = (n + V - 1) / V;
  const long long per = (packs + kPeerChunks - 1) / kPeerChunks;
  const long long lo = (long long)c * per * V, hi = min(n, lo + per * V);
  T* mine = reinterpret_cast<T*>(tab.base[rank] + kPeerFlagBytes + slot_off);
  // A: local split-K reduction into this rank's slot
  for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    double acc = 0.0;
    for (int s = 0; s < splits; ++s) acc += (double)partial[(long long)s * n + i];
    mine[i] = (T)acc;
  }
  __syncthreads();
  // B: publish chunk c to every rank (own flag array included), then wait for every rank's chunk c.
  // The release is cumulative over the CTA's writes ordered before it by the barrier.
  if (threadIdx.x < world) {
    uint32_t* flag = reinterpret_cast<uint32_t*>(tab.base[threadIdx.x]) + rank * kPeerChunks + c;
    st_release_sys(flag, epoch);
    const uint32_t* seen = reinterpret_cast<const uint32_t*>(tab.base[rank]) + threadIdx.x * kPeerChunks + c;
    unsigned long long t0 = 0, now = 0;
    unsigned int spin = 0;
    while ((int32_t)(ld_acquire_sys(seen) - epoch) < 0) {     // epochs only grow
      if ((++spin & 1023u) == 0) {                  // a peer that never arrives must not hang the GPU
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        if (now - t0 > kPeerTimeoutNs) {
          if (atomicCAS(err, 0u, epoch) == 0u) err[1] = (unsigned int)threadIdx.x;
          __threadfence_system();
          break;
        }
      }
    }
  }
  __syncthreads();
  // C: sum the chunk over ranks in rank order: all ranks' packs are requested first (independent
  // 16-byte NVLink reads in flight), then added in float64
  for (long long i = lo + (long long)threadIdx.x * V; i < hi; i += (long long)blockDim.x * V) {
    float4 pack[kPeerMaxWorld];
#pragma unroll
    for (int r = 0; r < kPeerMaxWorld; ++r)
      if (r < world)
        pack[r] = ld_peer16(reinterpret_cast<const T*>(tab.base[r] + kPeerFlagBytes + slot_off) + i);
    double acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = 0.0;
#pragma unroll
    for (int r = 0; r < kPeerMaxWorld; ++r)
      if (r < world) {
        const T* e = reinterpret_cast<const T*>(&pack[r]);
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] += (double)e[k];
      }
#pragma unroll
    for (int k = 0; k < V; ++k)
      if (i + k < hi) out[i + k] = (T)acc[k];
  }
}

}  // namespace

// used by the wgrad epilogue (conv.cu) and by mgb_peer_allreduce_sum
bool peer_ready() { return g_peer.ready && g_peer.world > 1; }

int peer_reduce_allreduce(int dtype, const void* partial, int splits, long long n, void* out, cudaStream_t st) {
  if (!peer_ready()) return fail("peer all-reduce: mgb_peer_connect has not been called");
  const size_t esz = dtype == MGB_F32 ? 4 : 8;
  if ((size_t)n * esz > g_peer.slot_bytes)
    return fail("peer all-reduce: %lld elements exceed the %zu-byte slot", n, g_peer.slot_bytes);
  if (dtype != MGB_F32 && dtype != MGB_F64) return fail("peer all-reduce: bad dtype %d", dtype);
  if (g_peer.err_host && g_peer.err_host[0] != 0)
    return fail("peer all-reduce: call %u timed out after %llu s waiting for rank %u — the ranks are out of step "
                "(a rank failed before its collective, or the calls were not issued in the same order everywhere)",
                g_peer.err_host[0], kPeerTimeoutNs / 1000000000ull, g_peer.err_host[1]);
  if (g_peer.stream_known && st != g_peer.stream)
    return fail("peer all-reduce: every call must be issued on one stream (epoch and slot parity assume stream "
                "order); first call used %p, this one %p", (void*)g_peer.stream, (void*)st);
  g_peer.stream_known = true;
  g_peer.stream = st;
  if (n == 0) return 0;
  const uint32_t epoch = ++g_peer.epoch;
  const size_t slot_off = (epoch & 1u) ? g_peer.slot_bytes : 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kPeerChunks); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;      // all CTAs co-resident: cross-rank waits cannot starve
  at[0].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (dtype == MGB_F32)
    MGB_CUDA(cudaLaunchKernelEx(&cfg, peer_reduce_allreduce_kernel<float>, (const float*)partial, splits, n,
                                (float*)out, g_peer.table, g_peer.rank, g_peer.world, slot_off, epoch,
                                g_peer.err_dev));
  else
    MGB_CUDA(cudaLaunchKernelEx(&cfg, peer_reduce_allreduce_kernel<double>, (const double*)partial, splits, n,
                                (double*)out, g_peer.table, g_peer.rank, g_peer.world, slot_off, epoch,
                                g_peer.err_dev));
  count_launch();
  return 0;
}

}  // namespace mgb

using namespace mgb;

extern "C" int mgb_peer_create(int rank, int world_size, size_t slot_bytes, void* ipc_handle64) {
  if (world_size < 1 || world_size > kPeerMaxWorld || rank < 0 || rank >= world_size || !ipc_handle64)
    return fail("mgb_peer_create: bad arguments (world=%d rank=%d, at most %d ranks)", world_size, rank,
                kPeerMaxWorld);
  if (g_peer.local) return fail("mgb_peer_create: already created");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  slot_bytes = align_up(slot_bytes, 256);
  const size_t total = kPeerFlagBytes + 2 * slot_bytes;
  void* ptr = nullptr;
  MGB_CUDA(cudaMalloc(&ptr, total));
  MGB_CUDA(cudaMemset(ptr, 0, total));
  MGB_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
  if (e != cudaSuccess) {
    cudaFree(ptr);
    return fail("cudaIpcGetMemHandle -> %s", cudaGetErrorString(e));
  }
  memcpy(ipc_handle64, &h, sizeof(h));
  void* eh = nullptr;
  if (cudaHostAlloc(&eh, 64, cudaHostAllocMapped) != cudaSuccess) {
    cudaFree(ptr);
    return fail("mgb_peer_create: cudaHostAlloc for the time-out flag failed");
  }
  memset(eh, 0, 64);
  void* ed = nullptr;
  if (cudaHostGetDevicePointer(&ed, eh, 0) != cudaSuccess) {
    cudaFreeHost(eh); cudaFree(ptr);
    return fail("mgb_peer_create: cudaHostGetDevicePointer failed");
  }
  g_peer.err_host = (unsigned int*)eh; g_peer.err_dev = (unsigned int*)ed;
  g_peer.local = (char*)ptr;
  g_peer.rank = rank; g_peer.world = world_size; g_peer.slot_bytes = slot_bytes; g_peer.epoch = 0;
  return 0;
}

extern "C" int mgb_peer_connect(const void* handles) {
  if (!g_peer.local || !handles) return fail("mgb_peer_connect: mgb_peer_create first");
  for (int r = 0; r < g_peer.world; ++r) {
    if (r == g_peer.rank) { g_peer.table.base[r] = g_peer.local; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)r * sizeof(h), sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail("cudaIpcOpenMemHandle(rank %d) -> %s", r, cudaGetErrorString(e));
    g_peer.table.base[r] = (char*)p;
  }
  g_peer.ready = true;
  return 0;
}

extern "C" int mgb_peer_destroy(void) {
  if (!g_peer.local) return 0;
  cudaDeviceSynchronize();
  if (g_peer.ready)
    for (int r = 0; r < g_peer.world; ++r)
      if (r != g_peer.rank && g_peer.table.base[r]) cudaIpcCloseMemHandle(g_peer.table.base[r]);
  cudaFree(g_peer.local);
  if (g_peer.err_host) cudaFreeHost(g_peer.err_host);
  g_peer = PeerState{};
  return 0;
}

extern "C" int mgb_peer_status(int* timed_out_call, int* waited_for_rank) {
  if (timed_out_call) *timed_out_call = g_peer.err_host ? (int)g_peer.err_host[0] : 0;
  if (waited_for_rank) *waited_for_rank = g_peer.err_host ? (int)g_peer.err_host[1] : 0;
  return 0;
}

extern "C" int mgb_peer_slot_bytes(size_t* bytes) {
  if (!bytes) return fail("mgb_peer_slot_bytes: null");
  *bytes = peer_ready() ? g_peer.slot_bytes : 0;
  return 0;
}

extern "C" int mgb_peer_allreduce_sum(void* buf, size_t n, int dtype, void* stream) {
  return peer_reduce_allreduce(dtype, buf, 1, (long long)n, buf, as_stream(stream));
}

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
