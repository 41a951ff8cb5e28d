// FP64 convolution GEMMs on DMMA (mma.sync.m8n8k4.f64) with TMA-staged operands and a dedicated
// producer warp.
//
// conv_gemm.cuh gathers its operands with 8-byte cp.async issued by the MMA warps themselves: per DMMA
// those warps execute ~4.7 other instructions (tap-table look-ups, bounds checks, address arithmetic),
// every k-slab ends in a __syncthreads, and the DMMA pipe idles a quarter of the time (ncu: 74.7 % active
// at C2; with the main-loop copies removed the same kernel runs 4.31 -> 3.67 ms).  Here
//   * one elected thread of a producer warpgroup feeds the ring: per 16-deep k-slab ONE cp.async.bulk.tensor.5d
//     for the pixel operand (a box of BM pixels x 16 channels of one filter tap; tap offsets are
//     coordinate offsets, the conv stride is the map's traversal stride, padding is TMA zero fill) and
//     one 2-D TMA for the packed filter matrix, completing on an mbarrier;
//   * eight MMA warps only wait on that mbarrier, read fragments and issue DMMAs; they hand a slot back
//     with one mbarrier arrive per warp, so warps drift apart instead of meeting at a block barrier.
// Shared-memory layouts are chosen so that the fragment reads are conflict-free without padding (TMA
// cannot pad).  The pixel operand is described as (x2, C, x1, x0, n), so a slab is
// [pixel group][16 k][row of pixels along x2].  TMA wants the innermost start coordinate 16-byte aligned
// (measured: scripts/probes/tma_f64_probe.cu — an odd coordinate of 8-byte elements is an illegal
// instruction), and filter taps shift x2 by single pixels: the box is therefore 12 pixels wide, starts at the
// even coordinate at or below the one wanted, and the MMA warps read 8 pixels at an offset of 0 or 1.
// 96-byte rows also spread the 4 k x 8 pixel fragment over all banks exactly twice (2 wavefronts = the
// minimum for 256 bytes).  K-major operands (the packed filters) are 128-byte rows under SWIZZLE_128B.
#pragma once
#include "conv_tma.cuh"

namespace mgb {
namespace conv {
namespace f64t {

constexpr int BK = 16;
// pixels per row of the pixel operand in shared memory for a gather stride PM along the innermost axis: TMA
// cannot stride dimension 0 (elementStrides[0] is ignored), so the row is contiguous — 8 pixels PM apart,
// + 1 for an odd start, rounded up so that 4 k-rows of 8 lanes spread over the banks (12 at PM = 1)
constexpr int row_pixels(int pm) { return 8 * pm + 4; }
constexpr int kMaxTaps = 128;
constexpr int CONSUMERS = 8;
// two MMA warpgroups + one producer warpgroup (only one thread of it works): register budgets move between
// warpgroups with setmaxnreg — 128 accumulator registers per MMA thread do not fit an even split of the file
constexpr int THREADS = (CONSUMERS + 4) * 32;

template <int R> __device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(R)); }
template <int R> __device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(R)); }

struct Taps {
  int n;
  short d[kMaxTaps][3];        // coordinate delta of each tap (K order: tap-major, channel fastest)
};

struct FwdLikeParams {
  double* out;
  OutMap om;
  int P[3];                    // pixel grid per sample
  int pmul[3], base[3];        // source coord = pixel * pmul + base + tap delta (pmul[2] == the kernel's PM)
  int box[3], boxn;            // pixel box of one M tile: box[2] == 8, product == BM
  int tiles[3], tiles_n;
  int Nbatch, Nout;
  int nchunk;                  // 16-channel chunks per tap
  Taps taps;
};

template <int BM, int BN, int STAGES, int PM>
constexpr size_t fwdlike_smem_bytes() {
  return (size_t)STAGES * (BM / 8 * BK * row_pixels(PM) * 8 + BN * BK * 8) + 1024 + 256;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(tc::smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// D[pixel, channel] = sum_{tap, c} T[pixel * pmul + base + delta(tap), c] * B[channel][tap][c]
// grid = (M tiles, ceil(Nout / BN)); 8 MMA warps of 64 x 32 + a producer warpgroup
template <int BM, int BN, int STAGES, int PM>
__global__ void __launch_bounds__(THREADS, 1)
f64_fwdlike_tma_kernel(const __grid_constant__ FwdLikeParams p, const __grid_constant__ CUtensorMap map_a,
                       const __grid_constant__ CUtensorMap map_b) {
  static_assert((BM / 64) * (BN / 32) == CONSUMERS, "8 MMA warps of 64 x 32");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int AW = row_pixels(PM);
  constexpr int A_STAGE = BM / 8 * BK * AW * 8, B_STAGE = BN * BK * 8;
  static_assert((STAGES * A_STAGE) % 1024 == 0, "the swizzled B ring starts 1024-byte aligned");
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + STAGES * A_STAGE;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(b_ring + STAGES * B_STAGE);
  uint64_t* empty_bar = full_bar + STAGES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma::prefetch_tmap(&map_a); tma::prefetch_tmap(&map_b);
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(full_bar + s, 1);
      tc::mbar_init(empty_bar + s, CONSUMERS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  int t = blockIdx.x;
  const int t2 = t % p.tiles[2]; t /= p.tiles[2];
  const int t1 = t % p.tiles[1]; t /= p.tiles[1];
  const int t0 = t % p.tiles[0]; t /= p.tiles[0];
  const int n0 = t * p.boxn;
  const int g0 = t0 * p.box[0], g1 = t1 * p.box[1], g2 = t2 * p.box[2];
  const int col0 = blockIdx.y * BN;
  const int S = p.taps.n * p.nchunk;

  if (warp >= CONSUMERS) {
    reg_dealloc<40>();
    if (warp == CONSUMERS && tc::elect_one()) {
      // ===== producer =====
      const int x0 = g0 * p.pmul[0] + p.base[0], x1 = g1 * p.pmul[1] + p.base[1], x2 = g2 * PM + p.base[2];
      int tap = 0, cc = 0;
      for (int s = 0; s < S; ++s) {
        const int stage = s % STAGES;
        if (s >= STAGES) tc::mbar_wait(empty_bar + stage, (uint32_t)((s / STAGES - 1) & 1));
        tma::mbar_expect_tx(full_bar + stage, (uint32_t)(A_STAGE + B_STAGE));
        tma::tma_load_5d(tc::smem_u32(a_ring + stage * A_STAGE), &map_a, full_bar + stage,
                         (x2 + p.taps.d[tap][2]) & ~1, cc * BK, x1 + p.taps.d[tap][1], x0 + p.taps.d[tap][0], n0);
        tma::tma_load_2d(tc::smem_u32(b_ring + stage * B_STAGE), &map_b, full_bar + stage, s * BK, col0);
        if (++cc == p.nchunk) { cc = 0; ++tap; }
      }
    }
    return;
  }

  // ===== MMA warps =====
  reg_alloc<224>();
  const int warp_m = (warp % (BM / 64)) * 64, warp_n = (warp / (BM / 64)) * 32;
  const int lr = lane >> 2, lk = lane & 3;
  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // fragment offsets that do not depend on the stage
  const int a_off = ((warp_m >> 3) * BK + lk) * AW + lr * PM;           // doubles: [group][k][AW]
  int b_off[4];                                                        // bytes, chunk swizzle applied per kk
#pragma unroll
  for (int j = 0; j < 4; ++j) b_off[j] = (warp_n + j * 8 + lr) * 128 + (lk & 1) * 8;

  int ctap = 0, ccc = 0;
  for (int s = 0; s < S; ++s) {
    const int stage = s % STAGES;
    const int shift = (g2 + p.base[2] + p.taps.d[ctap][2]) & 1;          // the box started one pixel early
    if (++ccc == p.nchunk) { ccc = 0; ++ctap; }
    tc::mbar_wait(full_bar + stage, (uint32_t)((s / STAGES) & 1));
    const double* As = reinterpret_cast<const double*>(a_ring + stage * A_STAGE) + a_off + shift;
    const uint8_t* Bs = b_ring + stage * B_STAGE;
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
      double a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = As[(i * BK + kk * 4) * AW];
      const int chunk = ((2 * kk + (lk >> 1)) ^ lr) << 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const double*>(Bs + b_off[j] + chunk);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[i][j], a[i], b[j]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_bar + stage);
  }

  // epilogue: rows of a thread are pixels (group = 8 consecutive pixels along the last axis)
  const int lc = (lane & 3) * 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = warp_m + i * 8 + lr;
    int g = r >> 3;
    const int i2 = r & 7;
    const int i1 = g % p.box[1]; g /= p.box[1];
    const int i0 = g % p.box[0]; g /= p.box[0];
    const int n = n0 + g;
    const int q0 = g0 + i0, q1 = g1 + i1, q2 = g2 + i2;
    if (n >= p.Nbatch || q0 >= p.P[0] || q1 >= p.P[1] || q2 >= p.P[2]) continue;
    const long long rowoff = (long long)n * p.om.sample_stride +
                             (long long)(q0 * p.om.omul[0] + p.om.obase[0]) * p.om.os[0] +
                             (long long)(q1 * p.om.omul[1] + p.om.obase[1]) * p.om.os[1] +
                             (long long)(q2 * p.om.omul[2] + p.om.obase[2]) * p.om.os[2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = col0 + warp_n + j * 8 + lc;
      if (c < p.Nout) p.out[rowoff + (long long)c * p.om.chan_stride] = acc[i][j][0];
      if (c + 1 < p.Nout) p.out[rowoff + (long long)(c + 1) * p.om.chan_stride] = acc[i][j][1];
    }
  }
}

// ---- wgrad: D[f, (tap, c)] = sum over pixels of dy[pixel, f] * x[pixel * s + tap * d - p, c] ----
//
// K walks the output pixels 16 at a time (2 rows x 8 pixels of one image), split over CTAs.  Per k-slab the
// producer brings dy as [128 filters][2 rows][8 pixels] under SWIZZLE_128B — TMA gives every 64-byte inner row
// its own 128-byte line (measured: scripts/probes/tma_f64_probe.cu variant 9), so the slab occupies 32 KB
// for 16 KB of data and line (2 f + row) holds chunk c at 16-byte slot c ^ (line & 7): 8 filters x 4 k of an
// A fragment land on all banks exactly twice — and, for each of the
// N tile's (tap, 32-channel) blocks, x as [32 channels][2 rows][WX pixels]: contiguous rows of 8 S2 + 2 pixels
// (the tap shifts the start by single pixels and TMA wants it even, see above; TMA cannot stride dimension 0);
// at S2 = 1 the 160-byte rows put the 8 columns x 4 k of a B fragment on all banks exactly twice.  Conv
// strides along the outer axes are the map's traversal strides.
constexpr int wg_row_pixels(int s2) { return 8 * s2 + 2; }
constexpr int WG_A_BYTES = 128 * BK * 8;             // dy slab: bytes TMA delivers
constexpr int WG_A_STAGE = 2 * WG_A_BYTES;           // ... and its footprint (64-byte rows in 128-byte lines)
constexpr int wg_blk_bytes(int s2) { return 32 * 2 * wg_row_pixels(s2) * 8; }   // one (tap, 32-channel) block of x

struct WgradParams {
  double* partial;             // [splits][F][C][Wp]
  int G[3];                    // output extents
  int S[3], Pd[3], D[3], W[3];
  int F, C, Wp, cblocks, nblocks_total;
  int tiles1, tiles2;          // ceil(G1 / 2), ceil(G2 / 8)
  int boxes_total, boxes_per_split;
};

template <int NB, int STAGES, int S2>
constexpr size_t wgrad_smem_bytes() {
  return (size_t)STAGES * (WG_A_STAGE + NB * wg_blk_bytes(S2)) + 1024 + 256;
}

// grid = (ceil(F / 128), ceil(nblocks / NB), splits); NB = 3 or 4 blocks of 32 columns per N tile
template <int NB, int STAGES, int S2>
__global__ void __launch_bounds__(THREADS, 1)
f64_wgrad_tma_kernel(const __grid_constant__ WgradParams p, const __grid_constant__ CUtensorMap map_dy,
                     const __grid_constant__ CUtensorMap map_x) {
  constexpr int BN = NB * 32, WN = BN / 4, NI = WN / 8;          // 8 MMA warps: 2 (M) x 4 (N), 64 x WN each
  constexpr int WX = wg_row_pixels(S2), WG_BLK = wg_blk_bytes(S2);
  constexpr int STAGE = WG_A_STAGE + NB * WG_BLK;
  static_assert(STAGE % 1024 == 0, "swizzled dy slabs start 1024-byte aligned");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
  uint64_t* empty_bar = full_bar + STAGES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma::prefetch_tmap(&map_dy); tma::prefetch_tmap(&map_x);
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(full_bar + s, 1);
      tc::mbar_init(empty_bar + s, CONSUMERS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  const int m0 = blockIdx.x * 128;
  const int blk0 = blockIdx.y * NB;
  const int nvalid = min(NB, p.nblocks_total - blk0);
  const int kb0 = blockIdx.z * p.boxes_per_split;
  const int kb1 = min(p.boxes_total, kb0 + p.boxes_per_split);
  const int S = max(0, kb1 - kb0);

  // tap of block b -> coordinate deltas
  auto tap_delta = [&](int blk, int& d0, int& d1, int& d2, int& cb) {
    int tap = blk / p.cblocks;
    cb = blk - tap * p.cblocks;
    const int t2 = tap % p.W[2]; tap /= p.W[2];
    const int t1 = tap % p.W[1], t0 = tap / p.W[1];
    d0 = t0 * p.D[0] - p.Pd[0]; d1 = t1 * p.D[1] - p.Pd[1]; d2 = t2 * p.D[2] - p.Pd[2];
  };

  if (warp >= CONSUMERS) {
    reg_dealloc<40>();
    if (warp == CONSUMERS && tc::elect_one()) {
      // ===== producer =====
      int bd0[NB], bd1[NB], bd2[NB], bcb[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) tap_delta(min(blk0 + b, p.nblocks_total - 1), bd0[b], bd1[b], bd2[b], bcb[b]);
      const uint32_t tx = (uint32_t)(WG_A_BYTES + nvalid * WG_BLK);
      for (int s = 0; s < S; ++s) {
        const int stage = s % STAGES;
        if (s >= STAGES) tc::mbar_wait(empty_bar + stage, (uint32_t)((s / STAGES - 1) & 1));
        int kb = kb0 + s;
        const int t2 = kb % p.tiles2; kb /= p.tiles2;
        const int t1 = kb % p.tiles1; kb /= p.tiles1;
        const int g0 = kb % p.G[0];
        const int n = kb / p.G[0];
        const int g1 = t1 * 2, g2 = t2 * 8;
        const uint32_t a_dst = tc::smem_u32(smem + stage * STAGE);
        tma::mbar_expect_tx(full_bar + stage, tx);
        tma::tma_load_5d(a_dst, &map_dy, full_bar + stage, g2, g1, g0, m0, n);
#pragma unroll
        for (int b = 0; b < NB; ++b)
          if (b < nvalid)
            tma::tma_load_5d(a_dst + WG_A_STAGE + b * WG_BLK, &map_x, full_bar + stage, (g2 * S2 + bd2[b]) & ~1,
                             g1 * p.S[1] + bd1[b], g0 * p.S[0] + bd0[b], bcb[b] * 32, n);
      }
    }
    return;
  }

  // ===== MMA warps =====
  reg_alloc<224>();
  const int warp_m = (warp & 1) * 64, warp_n = (warp >> 1) * WN;
  const int lr = lane >> 2, lk = lane & 3;
  double acc[8][NI][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  int a_off[8];                 // bytes; the row's line and the swizzled 16-byte slot are added per kk
#pragma unroll
  for (int i = 0; i < 8; ++i) a_off[i] = (warp_m + i * 8 + lr) * 256 + (lk & 1) * 8;
  int b_off[NI];                // bytes: column row + the block's odd-start shift + lk
#pragma unroll
  for (int j = 0; j < NI; ++j) {
    const int col = warp_n + j * 8;
    int d0, d1, d2, cb;
    tap_delta(min(blk0 + col / 32, p.nblocks_total - 1), d0, d1, d2, cb);
    b_off[j] = WG_A_STAGE + (col + lr) * (2 * WX * 8) + ((d2 & 1) + lk * S2) * 8;      // g2 * S2 is even (g2 % 8 == 0)
  }

  for (int s = 0; s < S; ++s) {
    const int stage = s % STAGES;
    tc::mbar_wait(full_bar + stage, (uint32_t)((s / STAGES) & 1));
    const uint8_t* st = smem + stage * STAGE;
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
      double a[8], b[NI];
      const int row = kk >> 1;                                   // k = kk * 4 + lk -> (row, column) of the 2 x 8 box
      const int chunk = row * 128 + (((2 * (kk & 1) + (lk >> 1)) ^ ((lr & 3) * 2 + row)) << 4);
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const double*>(st + a_off[i] + chunk);
      const int koff = (row * WX + (kk & 1) * 4 * S2) * 8;
#pragma unroll
      for (int j = 0; j < NI; ++j) b[j] = *reinterpret_cast<const double*>(st + b_off[j] + koff);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) dmma(acc[i][j], a[i], b[j]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_bar + stage);
  }

  // epilogue: partial[split][f][c][tap]
  double* out = p.partial + (long long)blockIdx.z * p.F * p.C * p.Wp;
  const int lc = (lane & 3) * 2;
#pragma unroll
  for (int j = 0; j < NI; ++j) {
    const int col = warp_n + j * 8 + lc;          // even: col and col + 1 share a block
    const int blk = blk0 + col / 32;
    if (blk >= p.nblocks_total) continue;
    const int tap = blk / p.cblocks, cb = blk - tap * p.cblocks;
    const int c = cb * 32 + (col & 31);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int f = m0 + warp_m + i * 8 + lr;
      if (f >= p.F) continue;
      if (c < p.C) out[((long long)f * p.C + c) * p.Wp + tap] = acc[i][j][0];
      if (c + 1 < p.C) out[((long long)f * p.C + c + 1) * p.Wp + tap] = acc[i][j][1];
    }
  }
}

// packed filters for the kernel above: out[row][tap][ch] (ch padded to a multiple of 16 with zeros)
//   rows_are_f = 1 (forward): row = f, ch = c;   0 (dgrad): row = c, ch = f
struct PackTaps {
  int n;
  short widx[kMaxTaps];        // flat filter-tap index of each tap
};

__global__ void pack_w_f64_kernel(const double* __restrict__ w, double* __restrict__ out, int F, int C, int Wp,
                                  int rows_are_f, int chpad, const __grid_constant__ PackTaps taps) {
  const int rows = rows_are_f ? F : C, chs = rows_are_f ? C : F;
  const long long total = (long long)rows * taps.n * chpad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % chpad);
    const int t = (int)((i / chpad) % taps.n);
    const int row = (int)(i / ((long long)chpad * taps.n));
    double v = 0.0;
    if (ch < chs) {
      const int f = rows_are_f ? row : ch, c = rows_are_f ? ch : row;
      v = w[((long long)f * C + c) * Wp + taps.widx[t]];
    }
    out[i] = v;
  }
}

}  // namespace f64t
}  // namespace conv
}  // namespace mgb
