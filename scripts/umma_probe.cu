// Micro-benchmark: how fast does one SM (or one SM pair) retire kind::tf32 UMMAs?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I. -o scripts/umma_probe scripts/umma_probe.cu
//   scripts/umma_probe            (prints one line per configuration)
//
// Each CTA (or CTA pair, cta_group::2) issues `iters` k32-stages of the 3xTF32 pattern used by the
// conv kernels — per k8-step one UMMA of N1 columns (a_hi x [b_hi | b_lo]) and, if N2 > 0, one of N2
// columns (a_lo x b_hi) — on zeroed operands that stay in shared memory, optionally while another
// thread streams bulk copies into shared memory at a set rate (the TMA writes of the real kernels
// compete with the operand reads for the same 128 B/clk).  Reports SM clocks per k8-step.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "mygrad_b200/csrc/conv_halo.cuh"

using namespace mgb::conv;

namespace mgb {   // the headers refer to these; the probe links nothing else
void set_error(const char*, ...) {}
int fail(const char*, ...) { return 1; }
std::atomic<uint64_t> g_launches{0};
int sm_count() { return 148; }
bool peer_ready() { return false; }
int peer_reduce_allreduce(int, const void*, int, long long, void*, cudaStream_t) { return 1; }
}  // namespace mgb

struct ProbeParams {
  int n1, n2;          // UMMA widths per k8-step (n2 = 0: one UMMA per step)
  int iters;           // k32 stages
  int mode;            // 0: N1 (+N2) per k8-step, one accumulator; 1: three UMMAs of N1, one accumulator;
                       // 2: three UMMAs of N1 into three accumulators; 3: N1 alternating between two accumulators
  int inflight;        // writer chunks in flight (<= 4)
  int wr_bytes;        // bulk-copy chunk (0 = no writer)
  int wr_period;       // clocks between chunks
  const float* src;    // global source of the bulk copies
  unsigned long long* out;   // per CTA: total clocks, issue clocks, writer bytes
};

__device__ __forceinline__ void umma2_tf32_acc(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                               uint32_t b_hi, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.eq.b32 p, 0, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %5, p;\n\t}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc)
      : "memory");
}

__device__ __forceinline__ void bulk_copy(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(tc::smem_u32(bar))
      : "memory");
}

template <int CG>
__global__ void __launch_bounds__(128, 1) probe_kernel(const ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  // [A_hi 16K][A_lo 16K][B 2*bn rows x 128 B <= 32K][writer scratch 2 x 32K][barriers]
  uint8_t* a_hi = smem;
  uint8_t* b_op = smem + 32768;
  uint8_t* scratch = smem + 65536;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 65536 + 65536);
  uint64_t* done_bar = bars;           // MMAs retired
  uint64_t* wr_bar = bars + 1;         // [4] writer chunks landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  volatile uint32_t* stop = tmem_slot + 2;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank = 0;
  if (CG == 2) rank = tma::cluster_ctarank();
  for (int i = threadIdx.x; i < 65536 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    tc::mbar_init(done_bar, 1);
    for (int i = 0; i < 4; ++i) tc::mbar_init(wr_bar + i, 1);
    *stop = 0;
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  tc::fence_proxy_async();
  if (warp == 1) { if (CG == 2) tma::tmem2_alloc(tmem_slot, 512); else tc::tmem_alloc(tmem_slot, 512); }
  tc::tc_fence_before();
  __syncthreads();
  if (CG == 2) tma::cluster_sync_all();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 1) {
    const uint32_t tmem_u = __reduce_or_sync(0xffffffffu, tmem_base);
    if ((CG == 1 || rank == 0) && tc::elect_one()) {
      const uint32_t mfield = (uint32_t)((CG == 2 ? 256 : 128) >> 4) << 24;
      const uint32_t idesc1 = (tc::idesc_tf32(p.n1) & ~(0x1Fu << 24)) | mfield;
      const uint32_t idesc2 = (tc::idesc_tf32(p.n2 > 0 ? p.n2 : 8) & ~(0x1Fu << 24)) | mfield;
      const uint32_t dhi = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t ah = (tc::smem_u32(a_hi) >> 4) | 0x10000u, al = ah + (16384u >> 4);
      const uint32_t bh = (tc::smem_u32(b_op) >> 4) | 0x10000u;
      const long long t0 = clock64();
      for (int it = 0; it < p.iters; ++it) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (p.mode == 1 || p.mode == 2) {
            const uint32_t step = p.mode == 2 ? (uint32_t)p.n1 : 0u;
            halo::umma_tf32_acc(tmem_u, al + 2 * j, dhi, bh + 2 * j, dhi, idesc1);
            halo::umma_tf32_acc(tmem_u + step, ah + 2 * j, dhi, bh + 2 * j + 1024, dhi, idesc1);
            halo::umma_tf32_acc(tmem_u + 2 * step, ah + 2 * j, dhi, bh + 2 * j, dhi, idesc1);
          } else if (p.mode == 3) {
            halo::umma_tf32_acc(tmem_u + (uint32_t)((j & 1) * p.n1), ah + 2 * j, dhi, bh + 2 * j, dhi, idesc1);
          } else if (CG == 2) {
            umma2_tf32_acc(tmem_u, ah + 2 * j, dhi, bh + 2 * j, dhi, idesc1);
            if (p.n2 > 0) umma2_tf32_acc(tmem_u, al + 2 * j, dhi, bh + 2 * j, dhi, idesc2);
          } else {
            halo::umma_tf32_acc(tmem_u, ah + 2 * j, dhi, bh + 2 * j, dhi, idesc1);
            if (p.n2 > 0) halo::umma_tf32_acc(tmem_u, al + 2 * j, dhi, bh + 2 * j, dhi, idesc2);
          }
        }
      }
      const long long t1 = clock64();
      if (CG == 2) tma::umma2_commit(done_bar); else tc::umma_commit(done_bar);
      tc::mbar_wait(done_bar, 0);
      const long long t2 = clock64();
      p.out[blockIdx.x * 4 + 0] = (unsigned long long)(t2 - t0);
      p.out[blockIdx.x * 4 + 1] = (unsigned long long)(t1 - t0);
      *stop = 1;
    } else if (CG == 2 && rank == 1 && lane == 0) {
      tc::mbar_wait(done_bar, 0);       // multicast commit
      *stop = 1;
    }
  } else if (warp == 2 && lane == 0 && p.wr_bytes > 0) {
    // writer: bulk copies global -> shared at wr_bytes per wr_period clocks until the MMAs are done
    unsigned long long bytes = 0;
    long long next = clock64();
    int k = 0;
    while (!*stop) {
      const int F = p.inflight;
      const int s = k % F;
      if (k >= F) tc::mbar_wait(wr_bar + s, (uint32_t)((k / F - 1) & 1));
      while (clock64() < next && !*stop) {}
      next += p.wr_period;
      tma::mbar_expect_tx(wr_bar + s, (uint32_t)p.wr_bytes);
      bulk_copy(tc::smem_u32(scratch + s * 16384), p.src + (size_t)(k & 63) * 8192, (uint32_t)p.wr_bytes, wr_bar + s);
      bytes += p.wr_bytes;
      ++k;
    }
    for (int s = 0; s < p.inflight; ++s) {       // drain
      const int kk = (k - 1 - s);
      if (kk >= 0) tc::mbar_wait(wr_bar + (kk % p.inflight), (uint32_t)((kk / p.inflight) & 1));
    }
    p.out[blockIdx.x * 4 + 2] = bytes;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (CG == 2) tma::cluster_sync_all();
  if (warp == 1) { if (CG == 2) tma::tmem2_dealloc(tmem_base, 512); else tc::tmem_dealloc(tmem_base, 512); }
}

template <int CG>
void run(int n1, int n2, int wr_bytes, int wr_period, const float* src, unsigned long long* d_out, int grid, int mode = 0) {
  ProbeParams p{n1, n2, 2000, mode, 4, wr_bytes, wr_period, src, d_out};
  size_t smem = 65536 + 65536 + 256 + 1024;
  cudaFuncSetAttribute(probe_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaMemset(d_out, 0, grid * 4 * 8);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, probe_kernel<CG>, p);
    if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); exit(1); }
  }
  std::vector<unsigned long long> h(grid * 4);
  cudaMemcpy(h.data(), d_out, grid * 4 * 8, cudaMemcpyDeviceToHost);
  double tot = 0, iss = 0, wr = 0; int n = 0;
  for (int i = 0; i < grid; ++i) {
    if (h[i * 4] == 0) { wr += (double)h[i * 4 + 2]; continue; }
    tot += (double)h[i * 4]; iss += (double)h[i * 4 + 1]; wr += (double)h[i * 4 + 2]; ++n;
  }
  const double steps = 2000.0 * 4;
  const double clk = tot / n / steps;
  // useful MACs per SM per k8-step: 128 rows x (n1 + n2) columns x 8 (per CTA; the pair does twice that on two SMs)
  const double ideal = mode == 1 || mode == 2 ? 1.5 * n1 : (n1 + n2) / 2.0;     // tensor clocks per k8-step at 128x N x8 per N/2 clk
  printf("mode %d cta_group %d  N1 %3d N2 %3d  writer %5d B / %4d clk : %7.1f clk per k8-step (issue %6.1f), ideal %5.1f -> %.2f of tensor peak; "
         "writer achieved %.1f B/clk/SM\n", mode, CG, n1, n2, wr_bytes, wr_period, clk, iss / n / steps, ideal, ideal / clk,
         wr / grid / (tot / n));
}

int main() {
  float* src; unsigned long long* d_out;
  cudaMalloc(&src, 64 * 8192 * 4 + 65536);
  cudaMemset(src, 0, 64 * 8192 * 4 + 65536);
  cudaMalloc(&d_out, 148 * 4 * 8);
  const int grid = 148;
  const int widths[][2] = {{256, 0}, {128, 0}, {64, 0}, {256, 128}, {128, 64}, {64, 32}, {192, 0}};
  for (auto& w : widths) run<1>(w[0], w[1], 0, 0, src, d_out, grid);
  for (auto& w : widths) run<2>(w[0], w[1], 0, 0, src, d_out, grid);
  // with shared-memory write traffic: 16 KB chunks every T clocks
  for (int n : {192, 160, 128, 96}) { run<1>(n, 0, 0, 0, src, d_out, grid, 1); if (3 * n <= 512) run<1>(n, 0, 0, 0, src, d_out, grid, 2); }
  for (int n : {256, 128, 64}) run<1>(n, 0, 0, 0, src, d_out, grid, 3);
  const int periods[] = {512, 384, 256, 192, 128};
  for (int T : periods) {
    run<1>(256, 128, 16384, T, src, d_out, grid);
    run<1>(128, 64, 16384, T, src, d_out, grid);
    run<2>(256, 128, 16384, T, src, d_out, grid);
    run<2>(128, 64, 16384, T, src, d_out, grid);
  }
  printf("PROBE DONE\n");
  return 0;
}
