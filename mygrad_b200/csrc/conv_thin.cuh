// "Thin" convolutions: few channels on one side of the GEMM (the first layer of a CNN: C = 3).
//
// With C * taps = 27 the implicit GEMM has K = 27 (forward) or N = 27 (dW): the 128-wide MMA tiles of
// conv_gemm.cuh pad that to 32..128, every operand element is fetched by its own 4-byte cp.async with a
// table look-up, and the layer is 9-17 x off its HBM bound (config 5, layer 1: forward 0.40 ms and dW 0.75 ms
// for 293 MB of traffic = 0.045 ms).  These shapes do not need tensor cores (3.6 GFLOP): they need the bytes
// moved once and an FMA loop that the load/store pipe can feed.
//
//   thin_fwdlike_kernel  D[pixel, n] = sum_k gather(pixel, k) * B[n][k]   (FwdParams of conv_gemm.cuh: forward
//                        and every dgrad class).  A thread owns PX pixels (a block stride apart: coalesced
//                        stores) and NB output channels in registers; B sits in shared memory as [n][K padded
//                        to 4] and is read with one broadcast LDS.128 per 4 k; per 4-k chunk that is
//                        4 PX gathers + NB LDS for 4 NB PX FMAs.
//   thin_wgrad_kernel    D[f, k] = sum_pixels dy[pixel, f] * gather_x(pixel, k)   (WgradParams).  A lane owns a
//                        pixel, a warp owns an 8 x KB tile of D in registers (8 coalesced dy loads + KB gathers
//                        for 8 KB FMAs), the warps of a block tile F x K and walk the same pixels (the gathers
//                        of the other warps hit L1); lanes are summed by shuffles at the end and every block
//                        writes one slab of partials for the fixed-order split-K reduction.
// Sums run in the input type (float32 FMA for float32, runs of a few hundred terms) and are combined in float64
// by the split-K reduction, as in the other kernels.
#pragma once
#include "conv_gemm.cuh"

namespace mgb {
namespace conv {
namespace thin {

constexpr int THREADS = 256;

// grid = (pixel groups of THREADS * PX, ceil(N / NB)); dynamic smem = (NB * KP) * sizeof(T) + K * sizeof(TapEntry)
template <typename T, int NB, int PX>
__global__ void __launch_bounds__(THREADS) thin_fwdlike_kernel(const __grid_constant__ FwdParams<T> p, int KP) {
  extern __shared__ __align__(16) unsigned char thin_smem[];
  T* ws = reinterpret_cast<T*>(thin_smem);                                       // [NB][KP]
  TapEntry* tabs = reinterpret_cast<TapEntry*>(thin_smem + (size_t)NB * KP * sizeof(T));
  const int n0 = blockIdx.y * NB;
  for (int i = threadIdx.x; i < NB * KP; i += THREADS) {
    const int n = i / KP, k = i - n * KP;
    ws[i] = (n0 + n < p.N && k < p.K) ? p.B[(long long)(n0 + n) * p.ldb + k] : (T)0;
  }
  for (int k = threadIdx.x; k < p.K; k += THREADS) tabs[k] = load_tap(p.tab, k);
  __syncthreads();

  PixCoord pc[PX];
  const long long pix0 = (long long)blockIdx.x * (THREADS * PX) + threadIdx.x;
#pragma unroll
  for (int j = 0; j < PX; ++j) {
    const long long pix = pix0 + (long long)j * THREADS;
    const bool v = pix < p.M;
    int n, q0, q1, q2;
    split_pixel(p.ga, v ? (int)pix : 0, n, q0, q1, q2);
    pc[j] = pix_coord(p.ga, n, q0, q1, q2, v);
  }
  T acc[PX][NB];
#pragma unroll
  for (int j = 0; j < PX; ++j)
#pragma unroll
    for (int n = 0; n < NB; ++n) acc[j][n] = (T)0;

  for (int k0 = 0; k0 < KP; k0 += 4) {
    T xv[PX][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int k = k0 + kk;
      if (k < p.K) {
        const TapEntry te = tabs[k];
#pragma unroll
        for (int j = 0; j < PX; ++j)
          xv[j][kk] = tap_in_bounds(p.ga, pc[j], te) ? __ldg(p.src + pc[j].off + te.off) : (T)0;
      } else {
#pragma unroll
        for (int j = 0; j < PX; ++j) xv[j][kk] = (T)0;
      }
    }
#pragma unroll
    for (int n = 0; n < NB; ++n) {
      T w4[4];
      if (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(w4) = *reinterpret_cast<const float4*>(ws + n * KP + k0);
      } else {
        *reinterpret_cast<double2*>(w4) = *reinterpret_cast<const double2*>(ws + n * KP + k0);
        *reinterpret_cast<double2*>(w4 + 2) = *reinterpret_cast<const double2*>(ws + n * KP + k0 + 2);
      }
#pragma unroll
      for (int j = 0; j < PX; ++j)
        acc[j][n] += xv[j][0] * w4[0] + xv[j][1] * w4[1] + xv[j][2] * w4[2] + xv[j][3] * w4[3];
    }
  }

#pragma unroll
  for (int j = 0; j < PX; ++j) {
    const long long pix = pix0 + (long long)j * THREADS;
    if (pix >= p.M) continue;
    int n, q0, q1, q2;
    split_pixel(p.ga, (int)pix, n, q0, q1, q2);
    const long long rowoff = (long long)n * p.om.sample_stride +
                             (long long)(q0 * p.om.omul[0] + p.om.obase[0]) * p.om.os[0] +
                             (long long)(q1 * p.om.omul[1] + p.om.obase[1]) * p.om.os[1] +
                             (long long)(q2 * p.om.omul[2] + p.om.obase[2]) * p.om.os[2];
#pragma unroll
    for (int nn = 0; nn < NB; ++nn)
      if (n0 + nn < p.N) p.out[rowoff + (long long)(n0 + nn) * p.om.chan_stride] = acc[j][nn];
  }
}

// grid = (pixel slabs, F groups of 8 * FG); block = FG * KG <= 8 warps (an 8 x KB tile is 128 registers): warp (fg, kg) owns filters [8 fg, 8 fg + 8) of the
// block's group and taps [KB kg, KB kg + KB).  dynamic smem = N * sizeof(TapEntry) (the x-gather table)
template <typename T, int KB>
__global__ void __launch_bounds__(256) thin_wgrad_kernel(const __grid_constant__ WgradParams<T> p, int FG, int KG,
                                                         long long pix_per_block) {
  extern __shared__ __align__(16) unsigned char thin_smem[];
  TapEntry* tabs = reinterpret_cast<TapEntry*>(thin_smem);
  for (int k = threadIdx.x; k < p.N; k += blockDim.x) tabs[k] = load_tap(p.tabB, k);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int fg = warp % FG, kg = warp / FG;
  const int f0 = (blockIdx.y * FG + fg) * 8, k0 = kg * KB;
  long long foff[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) foff[i] = f0 + i < p.M ? (long long)load_tap(p.tabA, f0 + i).off : -1;

  T acc[8][KB];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int k = 0; k < KB; ++k) acc[i][k] = (T)0;

  const long long pb = (long long)blockIdx.x * pix_per_block, pe = min((long long)p.K, pb + pix_per_block);
  for (long long base = pb; base < pe; base += 32) {
    const long long pix = base + lane;
    const bool pv = pix < pe;
    int n, q0, q1, q2;
    split_pixel(p.ga, pv ? (int)pix : 0, n, q0, q1, q2);
    const PixCoord pa = pix_coord(p.ga, n, q0, q1, q2, pv);
    const PixCoord pxc = pix_coord(p.gb, n, q0, q1, q2, pv);
    T dyv[8], xv[KB];
#pragma unroll
    for (int i = 0; i < 8; ++i) dyv[i] = (pv && foff[i] >= 0) ? __ldg(p.srcA + pa.off + foff[i]) : (T)0;
#pragma unroll
    for (int k = 0; k < KB; ++k) {
      xv[k] = (T)0;
      if (k0 + k < p.N) {
        const TapEntry te = tabs[k0 + k];
        if (tap_in_bounds(p.gb, pxc, te)) xv[k] = __ldg(p.srcB + pxc.off + te.off);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int k = 0; k < KB; ++k) acc[i][k] += dyv[i] * xv[k];
  }

  T* out = p.partial + (long long)blockIdx.x * p.M * p.N;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int k = 0; k < KB; ++k) {
      T v = acc[i][k];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && f0 + i < p.M && k0 + k < p.N) out[(long long)(f0 + i) * p.N + k0 + k] = v;
    }
}

// ---- 2-D thin layers with the input staged in shared memory ---------------------------------------
//
// The kernels above pay ~12 instructions of table look-up and bounds checking per gathered element, which is most
// of their time at K = 27.  For 1-D / 2-D convolutions a BAND of output rows of one sample needs a small input
// patch ([C][band rows + tap extent][row + tap extent], zero-filled where it leaves the image): it is copied to
// shared memory once (coalesced, bounds-checked there), after which a tap is a constant shared-memory offset and the
// operand loads are conflict-free LDS with no checks.  dy / y rows of a band are contiguous, so those accesses are
// plain coalesced loads / stores of base + filter * Gp + pixel.

template <typename T>
struct Thin2dParams {
  const T* x;                  // (N, C, X1, X2)
  const T* w;                  // (F, C, W1, W2) forward only
  const T* dy;                 // (N, F, G1, G2)  wgrad only
  T* out;                      // forward: y (N, F, G1, G2); wgrad: partial [blocks][F][C*Wp]
  int N, C, F, X1, X2, G1, G2, W1, W2;
  int s1, s2, p1, p2, d1, d2;
  int R, nbands;               // output rows per band, bands per sample
  int RB, CB, CBp;             // band rows, columns, padded columns (shared-memory row stride)
};

template <typename T>
__device__ __forceinline__ void thin2d_load_band(const Thin2dParams<T>& p, T* band, int n, int r0, int tid, int nthreads) {
  // a warp takes whole (channel, band row) rows, its lanes walk the columns: one division per row, none per element
  const int nrows = p.C * p.RB;
  const T* xs = p.x + (long long)n * p.C * p.X1 * p.X2;
  const int x1_0 = r0 * p.s1 - p.p1, x2_0 = -p.p2;
  const int lane = tid & 31, nwarps = nthreads >> 5;
  for (int row = tid >> 5; row < nrows; row += nwarps) {
    const int c = row / p.RB, rb = row - c * p.RB;
    const int x1 = x1_0 + rb;
    const bool rv = (unsigned)x1 < (unsigned)p.X1;
    const T* src = xs + ((long long)c * p.X1 + (rv ? x1 : 0)) * p.X2;
    T* dst = band + row * p.CBp;
    for (int cb = lane; cb < p.CBp; cb += 32) {
      const int x2 = x2_0 + cb;
      const bool v = rv && cb < p.CB && (unsigned)x2 < (unsigned)p.X2;
      cp_async_zfill<sizeof(T)>(dst + cb, v ? src + x2 : xs, v);
    }
  }
}

// forward: grid = (blocks over bands, ceil(F / NB)); smem = [NB][KP] filters + KP tap offsets + one band.
// A thread keeps the KP <= KMAX gathered inputs of its PX pixels in registers and walks the block's NB filters:
// per filter KP / 4 broadcast LDS.128 of weights and KP * PX FMAs into PX accumulators that are stored at once —
// no accumulator array, ~90 registers, three blocks per SM.
template <typename T, int KMAX, int PX>
__global__ void __launch_bounds__(THREADS) thin2d_fwd_kernel(const __grid_constant__ Thin2dParams<T> p, int KP, int NB) {
  extern __shared__ __align__(16) unsigned char thin_smem[];
  const int K = p.C * p.W1 * p.W2;
  T* ws = reinterpret_cast<T*>(thin_smem);
  int* koff = reinterpret_cast<int*>(ws + NB * KP);
  T* band = reinterpret_cast<T*>(koff + KP);
  const int f0 = blockIdx.y * NB;
  for (int i = threadIdx.x; i < NB * KP; i += THREADS) {
    const int n = i / KP, k = i - n * KP;
    ws[i] = (f0 + n < p.F && k < K) ? p.w[(long long)(f0 + n) * K + k] : (T)0;
  }
  for (int k = threadIdx.x; k < KP; k += THREADS) {
    int o = 0;
    if (k < K) {
      const int c = k / (p.W1 * p.W2), t = k - c * (p.W1 * p.W2);
      const int t1 = t / p.W2, t2 = t - t1 * p.W2;
      o = (c * p.RB + t1 * p.d1) * p.CBp + t2 * p.d2;
    }
    koff[k] = o;
  }
  const int nf = min(NB, p.F - f0);
  const long long Gp = (long long)p.G1 * p.G2;
  const int total_bands = p.N * p.nbands;
  for (int bi = blockIdx.x; bi < total_bands; bi += gridDim.x) {
    const int n = bi / p.nbands, r0 = (bi - n * p.nbands) * p.R;
    const int npix = min(p.R, p.G1 - r0) * p.G2;
    __syncthreads();                                   // the previous band (and, first time, ws / koff) is done with
    thin2d_load_band(p, band, n, r0, threadIdx.x, THREADS);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    T* yb = p.out + ((long long)n * p.F + f0) * Gp + (long long)r0 * p.G2;
    for (int q0 = 0; q0 < npix; q0 += THREADS * PX) {
      T xv[PX][KMAX];
      int qq[PX];
#pragma unroll
      for (int j = 0; j < PX; ++j) {
        const int q = q0 + j * THREADS + threadIdx.x;
        const int r = q / p.G2, g2 = q - r * p.G2;
        const int xo = q < npix ? r * p.s1 * p.CBp + g2 * p.s2 : 0;
        qq[j] = q < npix ? q : -1;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) xv[j][k] = k < KP ? band[koff[k] + xo] : (T)0;
      }
      for (int nn = 0; nn < nf; ++nn) {
        T acc[PX];
#pragma unroll
        for (int j = 0; j < PX; ++j) acc[j] = (T)0;
        const T* wr = ws + nn * KP;
#pragma unroll
        for (int k0 = 0; k0 < KMAX; k0 += 4) {
          if (k0 < KP) {
            T w4[4];
            if (sizeof(T) == 4) {
              *reinterpret_cast<float4*>(w4) = *reinterpret_cast<const float4*>(wr + k0);
            } else {
              *reinterpret_cast<double2*>(w4) = *reinterpret_cast<const double2*>(wr + k0);
              *reinterpret_cast<double2*>(w4 + 2) = *reinterpret_cast<const double2*>(wr + k0 + 2);
            }
#pragma unroll
            for (int j = 0; j < PX; ++j) {
              acc[j] = fma(xv[j][k0], w4[0], acc[j]);
              acc[j] = fma(xv[j][k0 + 1], w4[1], acc[j]);
              acc[j] = fma(xv[j][k0 + 2], w4[2], acc[j]);
              acc[j] = fma(xv[j][k0 + 3], w4[3], acc[j]);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < PX; ++j)
          if (qq[j] >= 0) yb[(long long)nn * Gp + qq[j]] = acc[j];
      }
    }
  }
}

// wgrad: grid = (blocks over bands, F groups of 8 * FG); block = FG * KG <= 8 warps; smem = two bands (double buffer)
template <typename T, int KB>
__global__ void __launch_bounds__(256) thin2d_wgrad_kernel(const __grid_constant__ Thin2dParams<T> p, int FG, int KG) {
  extern __shared__ __align__(16) unsigned char thin_smem[];
  const int K = p.C * p.W1 * p.W2;
  const int band_elems = p.C * p.RB * p.CBp;
  T* bands = reinterpret_cast<T*>(thin_smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int fg = warp % FG, kg = warp / FG;
  const int f0 = (blockIdx.y * FG + fg) * 8, k0 = kg * KB;
  int toff[KB];
#pragma unroll
  for (int k = 0; k < KB; ++k) {
    int o = -1;
    if (k0 + k < K) {
      const int kk = k0 + k;
      const int c = kk / (p.W1 * p.W2), t = kk - c * (p.W1 * p.W2);
      const int t1 = t / p.W2, t2 = t - t1 * p.W2;
      o = (c * p.RB + t1 * p.d1) * p.CBp + t2 * p.d2;
    }
    toff[k] = o;
  }
  T acc[8][KB];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int k = 0; k < KB; ++k) acc[i][k] = (T)0;

  const long long Gp = (long long)p.G1 * p.G2;
  const int total_bands = p.N * p.nbands;
  int buf = 0;
  if ((int)blockIdx.x < total_bands) {
    const int n = blockIdx.x / p.nbands, r0 = (blockIdx.x - n * p.nbands) * p.R;
    thin2d_load_band(p, bands, n, r0, threadIdx.x, blockDim.x);
  }
  cp_async_commit();
  for (int bi = blockIdx.x; bi < total_bands; bi += gridDim.x, buf ^= 1) {
    const int nxt = bi + gridDim.x;
    if (nxt < total_bands) {
      const int n2 = nxt / p.nbands, r2 = (nxt - n2 * p.nbands) * p.R;
      thin2d_load_band(p, bands + (buf ^ 1) * band_elems, n2, r2, threadIdx.x, blockDim.x);
    }
    cp_async_commit();
    cp_async_wait<1>();                                // this band has landed; the next one is in flight
    __syncthreads();
    const T* band = bands + buf * band_elems;
    const int n = bi / p.nbands, r0 = (bi - n * p.nbands) * p.R;
    const int npix = min(p.R, p.G1 - r0) * p.G2;
    const T* dyb = p.dy + ((long long)n * p.F + f0) * Gp + (long long)r0 * p.G2;
    for (int q0 = 0; q0 < npix; q0 += 32) {
      const int q = q0 + lane;
      const bool pv = q < npix;
      const int r = q / p.G2, g2 = q - r * p.G2;
      const int xo = pv ? r * p.s1 * p.CBp + g2 * p.s2 : 0;
      T dyv[8], xv[KB];
#pragma unroll
      for (int i = 0; i < 8; ++i) dyv[i] = (pv && f0 + i < p.F) ? __ldg(dyb + (long long)i * Gp + q) : (T)0;
#pragma unroll
      for (int k = 0; k < KB; ++k) xv[k] = toff[k] >= 0 ? band[toff[k] + xo] : (T)0;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int k = 0; k < KB; ++k) acc[i][k] = fma(dyv[i], xv[k], acc[i][k]);
    }
    __syncthreads();                                   // everyone is done with `buf` before it is refilled
  }
  cp_async_wait<0>();

  T* out = p.out + (long long)blockIdx.x * p.F * K;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int k = 0; k < KB; ++k) {
      T v = acc[i][k];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && f0 + i < p.F && k0 + k < K) out[(long long)(f0 + i) * K + k0 + k] = v;
    }
}

// ---- the same two float32 kernels on mma.sync.m16n8k8 (3xTF32) ------------------------------------
//
// The CUDA-core kernels above are instruction-bound (ncu: 2370 instructions per pixel for 864 FMAs forward,
// 3300 for dW).  One m16n8k8 MMA replaces 32 warp-FMA instructions, so even with three MMAs per product
// (a_lo*b_hi + a_hi*b_lo + a_hi*b_hi) and the hi / lo split of the gathered operand done in registers, a warp
// spends ~10-15 instructions per pixel.  Fragments are gathered straight from the staged band: a tap is a
// constant offset, a pixel is a row offset.

// forward: M = 16 pixels per m-tile (MT per warp pass), N = 8 filters per n-tile (NT), K = 8 taps per step.
// smem = whi / wlo [KP8][NP] (filters pre-split, NP = 8 NT (+8): conflict-free B fragments) + koff[KP8] + band
template <int NT, int MT>
__global__ void __launch_bounds__(256) thin2d_fwd_mma_kernel(const __grid_constant__ Thin2dParams<float> p, int KP8,
                                                             int NP) {
  extern __shared__ __align__(16) unsigned char thin_smem[];
  const int K = p.C * p.W1 * p.W2;
  uint32_t* whi = reinterpret_cast<uint32_t*>(thin_smem);
  uint32_t* wlo = whi + KP8 * NP;
  int* koff = reinterpret_cast<int*>(wlo + KP8 * NP);
  int* pixoff = koff + KP8;                              // band offset of pixel q of a band: [R * G2], built once
  float* band = reinterpret_cast<float*>(pixoff + ((p.R * p.G2 + 3) & ~3));
  const int f0 = blockIdx.y * (NT * 8);
  for (int q = threadIdx.x; q < p.R * p.G2; q += 256) {
    const int r = q / p.G2, c = q - r * p.G2;
    pixoff[q] = r * p.s1 * p.CBp + c * p.s2;
  }
  for (int i = threadIdx.x; i < KP8 * NP; i += 256) {
    const int k = i / NP, n = i - k * NP;
    const float v = (k < K && n < NT * 8 && f0 + n < p.F) ? p.w[(long long)(f0 + n) * K + k] : 0.f;
    split_tf32(v, whi[i], wlo[i]);
  }
  for (int k = threadIdx.x; k < KP8; k += 256) {
    int o = 0;
    if (k < K) {
      const int c = k / (p.W1 * p.W2), t = k - c * (p.W1 * p.W2);
      const int t1 = t / p.W2, t2 = t - t1 * p.W2;
      o = (c * p.RB + t1 * p.d1) * p.CBp + t2 * p.d2;
    }
    koff[k] = o;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const long long Gp = (long long)p.G1 * p.G2;
  const int total_bands = p.N * p.nbands;
  for (int bi = blockIdx.x; bi < total_bands; bi += gridDim.x) {
    const int n = bi / p.nbands, r0 = (bi - n * p.nbands) * p.R;
    const int npix = min(p.R, p.G1 - r0) * p.G2;
    __syncthreads();
    thin2d_load_band(p, band, n, r0, threadIdx.x, 256);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    float* yb = p.out + ((long long)n * p.F + f0) * Gp + (long long)r0 * p.G2;
    for (int q0 = warp * (MT * 16); q0 < npix; q0 += 8 * MT * 16) {
      int xo[MT][2];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int q = q0 + mt * 16 + g + h * 8;
          xo[mt][h] = q < npix ? pixoff[q] : 0;
        }
      float acc[MT][NT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
      for (int ks = 0; ks < KP8; ks += 8) {
        const int o0 = koff[ks + t], o1 = koff[ks + t + 4];
        uint32_t ahi[MT][4], alo[MT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          split_tf32(band[o0 + xo[mt][0]], ahi[mt][0], alo[mt][0]);     // (pixel g,     tap t)
          split_tf32(band[o0 + xo[mt][1]], ahi[mt][1], alo[mt][1]);     // (pixel g + 8, tap t)
          split_tf32(band[o1 + xo[mt][0]], ahi[mt][2], alo[mt][2]);     // (pixel g,     tap t + 4)
          split_tf32(band[o1 + xo[mt][1]], ahi[mt][3], alo[mt][3]);     // (pixel g + 8, tap t + 4)
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int b0 = (ks + t) * NP + nt * 8 + g, b1 = b0 + 4 * NP;
          const uint32_t bh[2] = {whi[b0], whi[b1]}, bl[2] = {wlo[b0], wlo[b1]};
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            mma_tf32_16x8x8(acc[mt][nt], alo[mt], bh);
            mma_tf32_16x8x8(acc[mt][nt], ahi[mt], bl);
            mma_tf32_16x8x8(acc[mt][nt], ahi[mt], bh);
          }
        }
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int q = q0 + mt * 16 + g + h * 8;
          if (q >= npix) continue;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const int f = nt * 8 + 2 * t;
            if (f0 + f < p.F) yb[(long long)f * Gp + q] = acc[mt][nt][2 * h];
            if (f0 + f + 1 < p.F) yb[(long long)(f + 1) * Gp + q] = acc[mt][nt][2 * h + 1];
          }
        }
    }
  }
}

// dW: M = 16 filters per m-tile (MT), N = 8 taps per n-tile (NT), K = 8 pixels per step (8 columns of one band
// row).  Every warp walks its own (row, column group) items and keeps the whole MT x NT tile; the warps of a
// block are summed through shared memory in warp order (deterministic) and the block writes one slab of partials.
// grid = (blocks over bands, filter groups of 16 MT); smem = two bands (double buffer; the first doubles as the
// reduction scratch)
template <int MT, int NT>
__global__ void __launch_bounds__(256) thin2d_wgrad_mma_kernel(const __grid_constant__ Thin2dParams<float> p) {
  extern __shared__ __align__(16) unsigned char thin_smem[];
  const int K = p.C * p.W1 * p.W2;
  const int band_elems = p.C * p.RB * p.CBp;
  float* bands = reinterpret_cast<float*>(thin_smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int f0 = blockIdx.y * (MT * 16);
  int toff[NT];                                        // B fragment: tap = 8 nt + g
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    int o = -1;
    const int kk = nt * 8 + g;
    if (kk < K) {
      const int c = kk / (p.W1 * p.W2), tt = kk - c * (p.W1 * p.W2);
      const int t1 = tt / p.W2, t2 = tt - t1 * p.W2;
      o = (c * p.RB + t1 * p.d1) * p.CBp + t2 * p.d2;
    }
    toff[nt] = o;
  }
  float acc[MT][NT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;

  const long long Gp = (long long)p.G1 * p.G2;
  const int total_bands = p.N * p.nbands;
  const int ncg = (p.G2 + 7) / 8;
  int buf = 0;
  if ((int)blockIdx.x < total_bands) {
    const int n = blockIdx.x / p.nbands, r0 = (blockIdx.x - n * p.nbands) * p.R;
    thin2d_load_band(p, bands, n, r0, threadIdx.x, 256);
  }
  cp_async_commit();
  for (int bi = blockIdx.x; bi < total_bands; bi += gridDim.x, buf ^= 1) {
    const int nxt = bi + gridDim.x;
    if (nxt < total_bands) {
      const int n2 = nxt / p.nbands, r2 = (nxt - n2 * p.nbands) * p.R;
      thin2d_load_band(p, bands + (buf ^ 1) * band_elems, n2, r2, threadIdx.x, 256);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const float* band = bands + buf * band_elems;
    const int n = bi / p.nbands, r0 = (bi - n * p.nbands) * p.R;
    const int rows = min(p.R, p.G1 - r0);
    const float* dyb = p.dy + ((long long)n * p.F + f0) * Gp + (long long)r0 * p.G2;
    for (int item = warp; item < rows * ncg; item += 8) {
      const int r = item / ncg, c0 = (item - r * ncg) * 8;
      const int ca = c0 + t, cb = c0 + t + 4;
      const bool va = ca < p.G2, vb = cb < p.G2;
      const float* dyr = dyb + (long long)r * p.G2;
      uint32_t ahi[MT][4], alo[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int fa = mt * 16 + g, fb = fa + 8;
        const bool ma = f0 + fa < p.F, mb = f0 + fb < p.F;
        split_tf32((va && ma) ? __ldg(dyr + (long long)fa * Gp + ca) : 0.f, ahi[mt][0], alo[mt][0]);   // (filter g,     pixel t)
        split_tf32((va && mb) ? __ldg(dyr + (long long)fb * Gp + ca) : 0.f, ahi[mt][1], alo[mt][1]);   // (filter g + 8, pixel t)
        split_tf32((vb && ma) ? __ldg(dyr + (long long)fa * Gp + cb) : 0.f, ahi[mt][2], alo[mt][2]);   // (filter g,     pixel t + 4)
        split_tf32((vb && mb) ? __ldg(dyr + (long long)fb * Gp + cb) : 0.f, ahi[mt][3], alo[mt][3]);
      }
      const int xa = va ? r * p.s1 * p.CBp + ca * p.s2 : 0, xb = vb ? r * p.s1 * p.CBp + cb * p.s2 : 0;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        uint32_t bh[2], bl[2];
        split_tf32(toff[nt] >= 0 ? band[toff[nt] + xa] : 0.f, bh[0], bl[0]);      // (pixel t,     tap g)
        split_tf32(toff[nt] >= 0 ? band[toff[nt] + xb] : 0.f, bh[1], bl[1]);      // (pixel t + 4, tap g)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          mma_tf32_16x8x8(acc[mt][nt], alo[mt], bh);
          mma_tf32_16x8x8(acc[mt][nt], ahi[mt], bl);
          mma_tf32_16x8x8(acc[mt][nt], ahi[mt], bh);
        }
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();
  __syncthreads();

  // sum the 8 warps in warp order into shared memory, then write the slab: D[(filter), (tap)]
  float* red = bands;                                  // MT * 16 x NT * 8 floats
  for (int w = 0; w < 8; ++w) {
    if (warp == w) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int row = mt * 16 + g + (e >> 1) * 8, col = nt * 8 + 2 * t + (e & 1);
            float* dst = red + row * (NT * 8) + col;
            *dst = (w == 0 ? 0.f : *dst) + acc[mt][nt][e];
          }
    }
    __syncthreads();
  }
  float* out = p.out + (long long)blockIdx.x * p.F * K;
  for (int i = threadIdx.x; i < MT * 16 * NT * 8; i += 256) {
    const int row = i / (NT * 8), col = i - row * (NT * 8);
    if (f0 + row < p.F && col < K) out[(long long)(f0 + row) * K + col] = red[i];
  }
}

}  // namespace thin
}  // namespace conv
}  // namespace mgb
