// Measurement helpers for bench.py: register-resident MMA loops that give the
// tensor-pipe ceilings the roofline fractions are quoted against (MEASURED_PEAKS.json
// has HBM and bf16 only; FP64 DMMA and legacy TF32 mma.sync are measured here).
#include "common.cuh"

namespace mgb {
namespace {

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* sink, int iters) {
  double c[8][2];
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile(
          "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
          : "+d"(c[i][0]), "+d"(c[i][1])
          : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) sink[0] = s;
}

__global__ void __launch_bounds__(256) tf32_peak_kernel(float* sink, int iters) {
  float c[8][4];
  uint32_t a[4], b[2];
  a[0] = a[1] = a[2] = a[3] = __float_as_uint(1.0f);
  b[0] = b[1] = __float_as_uint(0.5f);
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile(
          "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
          "{%8,%9}, {%0,%1,%2,%3};\n"
          : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
          : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456f) sink[0] = s;
}

}  // namespace
}  // namespace mgb

using namespace mgb;

extern "C" int mgb_probe_mma_peak(int kind, int iters, double* tflops, void* stream) {
  if (!tflops || iters < 1) return fail("mgb_probe_mma_peak: bad arguments");
  cudaStream_t st = as_stream(stream);
  void* sink = nullptr;
  MGB_CUDA(cudaMalloc(&sink, 64));
  cudaEvent_t e0, e1;
  MGB_CUDA(cudaEventCreate(&e0));
  MGB_CUDA(cudaEventCreate(&e1));
  int ctas = sm_count() * 4, threads = 256;
  double flops_per_warp_iter = kind == 0 ? 8.0 * 2 * 8 * 8 * 4 : 8.0 * 2 * 16 * 8 * 8;
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    MGB_CUDA(cudaEventRecord(e0, st));
    if (kind == 0) dmma_peak_kernel<<<ctas, threads, 0, st>>>((double*)sink, iters);
    else if (kind == 1) tf32_peak_kernel<<<ctas, threads, 0, st>>>((float*)sink, iters);
    else return fail("mgb_probe_mma_peak: kind=%d", kind);
    MGB_LAUNCH_CHECK();
    MGB_CUDA(cudaEventRecord(e1, st));
    MGB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    MGB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  double total = flops_per_warp_iter * (double)iters * (threads / 32) * ctas;
  *tflops = total / (best * 1e-3) / 1e12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  return 0;
}
