// Which cp.async.bulk.tensor boxes of 8-byte elements does the hardware accept?  One variant per process
// (a fault kills the context).  build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_f64_probe tma_f64_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap map, double* out, int bytes, int c0, int c1, int c2, int c3, int c4, int rank) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(bar), dst = (uint32_t)__cvta_generic_to_shared(smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar_a), "r"(bytes) : "memory");
    if (rank == 5)
      asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n"
                   ::"r"(dst), "l"(&map), "r"(bar_a), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
    else if (rank == 4)
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n"
                   ::"r"(dst), "l"(&map), "r"(bar_a), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
                   ::"r"(dst), "l"(&map), "r"(bar_a), "r"(c0), "r"(c1), "r"(c2) : "memory");
  }
  __syncthreads();
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar_a) : "memory");
  for (int i = threadIdx.x; i < bytes / 8; i += blockDim.x) out[i] = reinterpret_cast<double*>(smem)[i];
}

int main(int argc, char** argv) {
  int v = argc > 1 ? atoi(argv[1]) : 0;
  const int X2 = 16, X1 = 16, X0 = 1, C = 64, N = 2;
  size_t n = (size_t)N * C * X0 * X1 * X2;
  std::vector<double> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = (double)i;
  double *d, *out;
  cudaMalloc(&d, n * 8); cudaMalloc(&out, 65536);
  cudaMemcpy(d, h.data(), n * 8, cudaMemcpyHostToDevice);
  cuuint64_t xp = (cuuint64_t)X0 * X1 * X2;
  CUtensorMap m;
  CUresult r;
  int rank = 5, bytes = 0;
  CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
  cuuint64_t dims[5], str[4]; cuuint32_t box[5], es[5] = {1, 1, 1, 1, 1};
  auto set = [&](int i, cuuint64_t dim, cuuint64_t stride_bytes, cuuint32_t b) { dims[i] = dim; if (i) str[i - 1] = stride_bytes; box[i] = b; };
  switch (v) {
    case 0: set(0, X2, 0, 8); set(1, C, xp * 8, 16); set(2, X1, X2 * 8, 8); set(3, X0, X1 * X2 * 8, 1); set(4, N, C * xp * 8, 2); break;
    case 1: set(0, X2, 0, 8); set(1, X1, X2 * 8, 8); set(2, X0, X1 * X2 * 8, 1); set(3, C, xp * 8, 16); set(4, N, C * xp * 8, 2); break;   // natural order
    case 2: rank = 4; set(0, X2, 0, 8); set(1, C, xp * 8, 16); set(2, X1, X2 * 8, 8); set(3, N, C * xp * 8, 2); break;
    case 3: set(0, X2, 0, 8); set(1, C, xp * 8, 16); set(2, X1, X2 * 8, 16); set(3, X0, X1 * X2 * 8, 1); set(4, N, C * xp * 8, 1); break;
    case 4: set(0, X2, 0, 8); set(1, C, xp * 8, 8); set(2, X1, X2 * 8, 8); set(3, X0, X1 * X2 * 8, 1); set(4, N, C * xp * 8, 2); break;
    case 5: rank = 3; set(0, X2, 0, 8); set(1, C, xp * 8, 16); set(2, X1, X2 * 8, 8); break;
    case 6: rank = 3; set(0, X2, 0, 8); set(1, X1, X2 * 8, 8); set(2, C, xp * 8, 16); break;
    case 7: rank = 3; set(0, X2, 0, 16); set(1, C, xp * 8, 16); set(2, X1, X2 * 8, 8); break;   // inner 128 B
    case 8: dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; rank = 3; set(0, 2 * X2, 0, 16); set(1, C, xp * 8, 16); set(2, X1, X2 * 8, 8); break;  // same bytes as 5, as float32
    case 9: rank = 3; set(0, X2, 0, 8); set(1, X1, X2 * 8, 2); set(2, C, xp * 8, 16); break;   // [c][2][8] under SWIZZLE_128B
    default: return 2;
  }
  bytes = 8; for (int i = 0; i < rank; ++i) bytes *= box[i];
  if (dt == CU_TENSOR_MAP_DATA_TYPE_FLOAT32) bytes /= 2;
  r = cuTensorMapEncodeTiled(&m, dt, rank, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, v == 9 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("variant %d rank %d encode rc %d bytes %d\n", v, rank, (int)r, bytes);
  if (r) return 1;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 64);
  int cc[5] = {0, 0, 0, 0, 0};
  for (int i = 0; i < 5 && i + 2 < argc; ++i) cc[i] = atoi(argv[i + 2]);
  probe<<<1, 128, 65536 + 64>>>(m, out, bytes, cc[0], cc[1], cc[2], cc[3], cc[4], rank);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<double> o(8192);
  cudaMemcpy(o.data(), out, 65536, cudaMemcpyDeviceToHost);
  printf("variant %d: %s; smem[0..3]=%g %g %g %g smem[8..9]=%g %g\n", v, cudaGetErrorString(e), o[0], o[1], o[2], o[3], o[8], o[9]);
  if (v == 9)
    for (int row = 0; row < 10; ++row) {
      printf("row %2d:", row);
      for (int i = 0; i < 16; ++i) printf(" %4g", o[row * 16 + i]);
      printf("\n");
    }
  return e != cudaSuccess;
}
