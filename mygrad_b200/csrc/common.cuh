// Shared helpers for libmgb200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>

#include "../../include/mgb.h"

namespace mgb {

// thread-local error text returned by mgb_last_error()
void set_error(const char* fmt, ...);
int fail(const char* fmt, ...);  // sets error, returns 1

extern std::atomic<uint64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

#define MGB_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess)                                                         \
      return ::mgb::fail("%s:%d %s -> %s", __FILE__, __LINE__, #call,              \
                         cudaGetErrorString(e__));                                  \
  } while (0)

#define MGB_LAUNCH_CHECK()                                                          \
  do {                                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess)                                                         \
      return ::mgb::fail("%s:%d kernel launch -> %s", __FILE__, __LINE__,          \
                         cudaGetErrorString(e__));                                  \
    ::mgb::count_launch();                                                          \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int sm_count();  // cached multiprocessor count of the current device

// comm.cu: fused "sum `splits` partials of n elements, then sum over ranks through peer memory"
bool peer_ready();
int peer_reduce_allreduce(int dtype, const void* partial, int splits, long long n, void* out, cudaStream_t st);

}  // namespace mgb
