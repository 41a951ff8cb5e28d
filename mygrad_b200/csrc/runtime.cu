// Runtime plumbing behind the C ABI: errors, memory, streams, events.
#include <vector>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace mgb {

static thread_local char t_err[1024] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}

int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
  return 1;
}

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) {
      cached = n;
      cached_dev = dev;
    }
  }
  return cached;
}

}  // namespace mgb

using namespace mgb;

extern "C" {

const char* mgb_last_error(void) { return t_err; }
const char* mgb_version(void) { return "mgb200 0.1 (sm_100a)"; }

int mgb_launch_count(uint64_t* count) {
  if (!count) return fail("mgb_launch_count: null");
  *count = g_launches.load();
  return 0;
}

int mgb_device_count(int* count) {
  if (!count) return fail("mgb_device_count: null");
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) {
    *count = 0;
    return fail("cudaGetDeviceCount -> %s", cudaGetErrorString(e));
  }
  return 0;
}

int mgb_set_device(int ordinal) {
  MGB_CUDA(cudaSetDevice(ordinal));
  return 0;
}

int mgb_get_device(int* ordinal) {
  MGB_CUDA(cudaGetDevice(ordinal));
  return 0;
}

int mgb_device_info(int* sms, int* major, int* minor, size_t* total_mem) {
  int dev = 0;
  MGB_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  MGB_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sms) *sms = p.multiProcessorCount;
  if (major) *major = p.major;
  if (minor) *minor = p.minor;
  if (total_mem) *total_mem = p.totalGlobalMem;
  return 0;
}

int mgb_malloc_async(void** ptr, size_t bytes, void* stream) {
  if (!ptr) return fail("mgb_malloc_async: null");
  if (bytes == 0) bytes = 16;  // zero-size tensors still get a distinct pointer
  static thread_local int pool_dev = -1;
  int dev = 0;
  MGB_CUDA(cudaGetDevice(&dev));
  if (dev != pool_dev) {
    // keep freed blocks in the stream-ordered pool instead of returning them to the OS
    cudaMemPool_t pool;
    MGB_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    uint64_t thresh = UINT64_MAX;
    MGB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
    pool_dev = dev;
  }
  MGB_CUDA(cudaMallocAsync(ptr, bytes, as_stream(stream)));
  return 0;
}

int mgb_free_async(void* ptr, void* stream) {
  if (!ptr) return 0;
  MGB_CUDA(cudaFreeAsync(ptr, as_stream(stream)));
  return 0;
}

int mgb_host_alloc(void** host_ptr, size_t bytes) {
  if (!host_ptr) return fail("mgb_host_alloc: null");
  MGB_CUDA(cudaMallocHost(host_ptr, bytes ? bytes : 16));
  return 0;
}

int mgb_host_free(void* host_ptr) {
  if (!host_ptr) return 0;
  MGB_CUDA(cudaFreeHost(host_ptr));
  return 0;
}

int mgb_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream) {
  if (bytes == 0) return 0;
  MGB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, as_stream(stream)));
  return 0;
}

int mgb_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream) {
  if (bytes == 0) return 0;
  MGB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, as_stream(stream)));
  return 0;
}

int mgb_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream) {
  if (bytes == 0) return 0;
  MGB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
  return 0;
}

int mgb_memset(void* dst, int value, size_t bytes, void* stream) {
  if (bytes == 0) return 0;
  MGB_CUDA(cudaMemsetAsync(dst, value, bytes, as_stream(stream)));
  return 0;
}

int mgb_stream_create(void** stream) {
  if (!stream) return fail("mgb_stream_create: null");
  cudaStream_t s;
  MGB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  *stream = s;
  return 0;
}

int mgb_stream_destroy(void* stream) {
  if (!stream) return 0;
  MGB_CUDA(cudaStreamDestroy(as_stream(stream)));
  return 0;
}

int mgb_stream_synchronize(void* stream) {
  MGB_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return 0;
}

int mgb_device_synchronize(void) {
  MGB_CUDA(cudaDeviceSynchronize());
  return 0;
}

int mgb_event_create(void** event) {
  if (!event) return fail("mgb_event_create: null");
  cudaEvent_t e;
  MGB_CUDA(cudaEventCreate(&e));
  *event = e;
  return 0;
}

int mgb_event_destroy(void* event) {
  if (!event) return 0;
  MGB_CUDA(cudaEventDestroy(reinterpret_cast<cudaEvent_t>(event)));
  return 0;
}

int mgb_event_record(void* event, void* stream) {
  MGB_CUDA(cudaEventRecord(reinterpret_cast<cudaEvent_t>(event), as_stream(stream)));
  return 0;
}

int mgb_event_synchronize(void* event) {
  MGB_CUDA(cudaEventSynchronize(reinterpret_cast<cudaEvent_t>(event)));
  return 0;
}

int mgb_event_elapsed_ms(void* start, void* stop, float* ms) {
  MGB_CUDA(cudaEventElapsedTime(ms, reinterpret_cast<cudaEvent_t>(start),
                                reinterpret_cast<cudaEvent_t>(stop)));
  return 0;
}

int mgb_stream_wait_event(void* stream, void* event) {
  MGB_CUDA(cudaStreamWaitEvent(as_stream(stream), reinterpret_cast<cudaEvent_t>(event), 0));
  return 0;
}

// ---- CUDA graphs: record everything this thread enqueues on `stream`, replay it with one launch ----
// (launch-bound steps: config 1 is 8 kernels of a few microseconds each behind ~150 us of host work)

int mgb_graph_begin(void* stream) {
  if (!stream) return fail("mgb_graph_begin: the legacy default stream cannot be captured; create a stream");
  MGB_CUDA(cudaStreamBeginCapture(as_stream(stream), cudaStreamCaptureModeThreadLocal));
  return 0;
}

int mgb_graph_end(void* stream, void** graph_exec, int* kernel_nodes) {
  if (!graph_exec) return fail("mgb_graph_end: null");
  cudaGraph_t graph = nullptr;
  MGB_CUDA(cudaStreamEndCapture(as_stream(stream), &graph));
  if (!graph) return fail("mgb_graph_end: the capture was invalidated");
  if (kernel_nodes) {
    size_t n = 0;
    *kernel_nodes = 0;
    if (cudaGraphGetNodes(graph, nullptr, &n) == cudaSuccess && n) {
      std::vector<cudaGraphNode_t> nodes(n);
      if (cudaGraphGetNodes(graph, nodes.data(), &n) == cudaSuccess)
        for (size_t i = 0; i < n; ++i) {
          cudaGraphNodeType t;
          if (cudaGraphNodeGetType(nodes[i], &t) == cudaSuccess && t == cudaGraphNodeTypeKernel) ++*kernel_nodes;
        }
    }
  }
  cudaGraphExec_t exec = nullptr;
  cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return fail("mgb_graph_end: cudaGraphInstantiate -> %s", cudaGetErrorString(e));
  *graph_exec = exec;
  return 0;
}

int mgb_graph_launch(void* graph_exec, void* stream) {
  if (!graph_exec) return fail("mgb_graph_launch: null");
  MGB_CUDA(cudaGraphLaunch(reinterpret_cast<cudaGraphExec_t>(graph_exec), as_stream(stream)));
  return 0;
}

int mgb_graph_destroy(void* graph_exec) {
  if (!graph_exec) return 0;
  MGB_CUDA(cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(graph_exec)));
  return 0;
}

}  // extern "C"
