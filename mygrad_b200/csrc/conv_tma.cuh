// FP32 conv fwd / dgrad on tcgen05 with TMA-staged operands (persistent, warp-specialised).
//
// Same math and operand format as conv_tc.cuh (3xTF32 over channels-last hi/lo planes), but
// the shared-memory ring is filled by the Tensor Memory Accelerator instead of cp.async:
//
//   * the planes are described to TMA as a rank-5 tensor (Cpad, X2, X1, X0, N); one
//     `cp.async.bulk.tensor.5d` brings the 32-channel slice of a BOX of 128 pixels
//     (b2 x b1 x b0 x bn) of one filter tap straight into a K-major SWIZZLE_128B tile.
//     Tap offsets are coordinate offsets, the conv stride is the map's traversal stride,
//     and out-of-bounds coordinates are zero-filled by the hardware == zero padding.
//   * the packed filter bank is a rank-2 tensor (K, rows): one 2-D TMA per plane and stage.
//   * one thread issues TMA, one thread issues the UMMAs, four warps run the epilogue;
//     the kernel is persistent (one CTA per SM looping over tiles) and the accumulator is
//     double-buffered in TMEM, so the epilogue of tile i overlaps the main loop of tile i+1.
#pragma once
#include <cuda.h>

#include "conv_tc.cuh"

namespace mgb {
namespace conv {
namespace tma {

using tc::BK;
using tc::BM;

constexpr int THREADS = 192;            // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue
constexpr int A_TILE_BYTES = BM * 128;

struct TmaParams {
  const tc::TcTap* taps;                // d0..d2 = coordinate delta of each tap (off unused)
  float* out;
  OutMap om;
  int P[3];                             // pixel grid per sample
  int pmul[3], base[3];                 // source coord = pixel*pmul + base + tap delta
  int box[3], boxn;                     // pixels per tile along each spatial axis / batch (powers of two)
  int tiles[3], tiles_n;                // tile counts
  int Nbatch, Nout;                     // batch size, real output channels
  int nchunk, ntaps;
  int bn, ny;                           // channel tile, number of channel tiles
  int tmem_cols;
  int total_tiles;                      // tiles_n*tiles[0]*tiles[1]*tiles[2]*ny
  // ncat = 1: the B_hi and B_lo tiles (adjacent in the stage, same K-major layout) are read as ONE
  // operand of 2*bn rows, so a_hi*[b_hi | b_lo] is a single UMMA of N = 2*bn into accumulator columns
  // [0, 2bn) and a_lo*b_hi a second one into [0, bn): two UMMAs and (A + 2B) + (A + B) bytes of
  // shared-memory reads per k-step instead of three UMMAs and 3(A + B).  The epilogue adds the halves.
  int ncat;
  int acc_cols;                         // accumulator columns per buffer: bn or 2*bn
};

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(tc::smem_u32(bar)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                            int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n" ::"r"(dst),
      "l"(map), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];\n" ::"r"(dst),
      "l"(map), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];\n" ::"l"(map) : "memory");
}

inline size_t smem_bytes(int bn, int stages) {
  return (size_t)stages * (2 * A_TILE_BYTES + 2 * (size_t)bn * 128) + 1024 + 256;
}

struct TileCoord {
  int n0, t0, t1, t2, y;   // batch origin, spatial tile indices, channel tile
};

__device__ __forceinline__ TileCoord decode_tile(const TmaParams& p, int tile) {
  TileCoord c;
  c.y = tile % p.ny; tile /= p.ny;
  c.t2 = tile % p.tiles[2]; tile /= p.tiles[2];
  c.t1 = tile % p.tiles[1]; tile /= p.tiles[1];
  c.t0 = tile % p.tiles[0]; tile /= p.tiles[0];
  c.n0 = tile * p.boxn;
  return c;
}

template <int STAGES>
__global__ void __launch_bounds__(THREADS, 1)
conv_tma_kernel(const __grid_constant__ TmaParams p, const __grid_constant__ CUtensorMap map_a_hi,
                const __grid_constant__ CUtensorMap map_a_lo, const __grid_constant__ CUtensorMap map_b_hi,
                const __grid_constant__ CUtensorMap map_b_lo) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int b_tile_bytes = p.bn * 128;
  constexpr int a_tile_bytes = A_TILE_BYTES;
  const int stage_bytes = 2 * a_tile_bytes + 2 * b_tile_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;    // accumulator buffer ready for the epilogue
  uint64_t* tempty_bar = tfull_bar + 2;        // accumulator buffer drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.ntaps * p.nchunk;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a_hi); prefetch_tmap(&map_a_lo); prefetch_tmap(&map_b_hi); prefetch_tmap(&map_b_lo);
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(full_bar + s, 1);
      tc::mbar_init(empty_bar + s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(tfull_bar + b, 1);
      tc::mbar_init(tempty_bar + b, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem_base = tc::smem_u32(smem);

  if (warp == 0) {
    if (tc::elect_one()) {
      // ===== TMA producer =====
      int it = 0;   // ring position, runs across tiles
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const TileCoord tcd = decode_tile(p, tile);
        const int x0 = tcd.t0 * p.box[0] * p.pmul[0] + p.base[0];
        const int x1 = tcd.t1 * p.box[1] * p.pmul[1] + p.base[1];
        const int x2 = tcd.t2 * p.box[2] * p.pmul[2] + p.base[2];
        for (int s = 0; s < S; ++s, ++it) {
          const int stage = it % STAGES;
          if (it >= STAGES) tc::mbar_wait(empty_bar + stage, (uint32_t)((it / STAGES - 1) & 1));
          const int tap = s / p.nchunk, cc = s - tap * p.nchunk;
          const int4 tv = __ldg(reinterpret_cast<const int4*>(p.taps) + tap);
          const uint32_t a_hi = smem_base + stage * stage_bytes;
          mbar_expect_tx(full_bar + stage, (uint32_t)stage_bytes);
          const int xs = x2 + tv.w;
          tma_load_5d(a_hi, &map_a_hi, full_bar + stage, cc * BK, xs, x1 + tv.z, x0 + tv.y, tcd.n0);
          tma_load_5d(a_hi + a_tile_bytes, &map_a_lo, full_bar + stage, cc * BK, xs, x1 + tv.z,
                      x0 + tv.y, tcd.n0);
          tma_load_2d(a_hi + 2 * a_tile_bytes, &map_b_hi, full_bar + stage, s * BK, tcd.y * p.bn);
          tma_load_2d(a_hi + 2 * a_tile_bytes + b_tile_bytes, &map_b_lo, full_bar + stage, s * BK,
                      tcd.y * p.bn);
        }
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {
      // ===== MMA issuer =====
      const uint32_t idesc = tc::idesc_tf32(p.bn);
      const uint32_t idesc2 = tc::idesc_tf32(2 * p.bn);   // a_hi x [b_hi | b_lo]
      int it = 0, local = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
        const int buf = local & 1;
        if (local >= 2) tc::mbar_wait(tempty_bar + buf, (uint32_t)((local / 2 - 1) & 1));
        tc::tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * p.acc_cols);
        for (int s = 0; s < S; ++s, ++it) {
          const int stage = it % STAGES;
          tc::mbar_wait(full_bar + stage, (uint32_t)((it / STAGES) & 1));
          tc::tc_fence_after();
          const uint32_t a_hi = smem_base + stage * stage_bytes;
          const uint32_t a_lo = a_hi + a_tile_bytes;
          const uint32_t b_hi = a_lo + a_tile_bytes;
          const uint32_t b_lo = b_hi + b_tile_bytes;
          if (p.ncat) {
#pragma unroll
            for (int j = 0; j < BK / 8; ++j) {
              const uint64_t dah = tc::smem_desc_k_sw128(a_hi + j * 32), dal = tc::smem_desc_k_sw128(a_lo + j * 32);
              const uint64_t dbh = tc::smem_desc_k_sw128(b_hi + j * 32);
              tc::umma_tf32(tmem_d, dah, dbh, idesc2, (s | j) ? 1u : 0u);   // cols [0,bn) hi*hi, [bn,2bn) hi*lo
              tc::umma_tf32(tmem_d, dal, dbh, idesc, 1u);                   // cols [0,bn) += lo*hi
            }
          } else {
#pragma unroll
            for (int j = 0; j < BK / 8; ++j) {
              const uint64_t dah = tc::smem_desc_k_sw128(a_hi + j * 32), dal = tc::smem_desc_k_sw128(a_lo + j * 32);
              const uint64_t dbh = tc::smem_desc_k_sw128(b_hi + j * 32), dbl = tc::smem_desc_k_sw128(b_lo + j * 32);
              tc::umma_tf32(tmem_d, dal, dbh, idesc, (s | j) ? 1u : 0u);
              tc::umma_tf32(tmem_d, dah, dbl, idesc, 1u);
              tc::umma_tf32(tmem_d, dah, dbh, idesc, 1u);
            }
          }
          tc::umma_commit(empty_bar + stage);
        }
        tc::umma_commit(tfull_bar + buf);
      }
    }
  } else {
    // ===== epilogue warps 2..5: TMEM lane quarter = warp % 4 =====
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;               // tile row == TMEM lane
    // row -> (in, i0, i1, i2) inside the box (all extents are powers of two)
    const int i2 = r % p.box[2];
    const int r1 = r / p.box[2];
    const int i1 = r1 % p.box[1];
    const int r0 = r1 / p.box[1];
    const int i0 = r0 % p.box[0];
    const int in = r0 / p.box[0];
    int local = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
      const int buf = local & 1;
      const TileCoord tcd = decode_tile(p, tile);
      const int n = tcd.n0 + in;
      const int g0 = tcd.t0 * p.box[0] + i0, g1 = tcd.t1 * p.box[1] + i1, g2 = tcd.t2 * p.box[2] + i2;
      const bool pv = n < p.Nbatch && g0 < p.P[0] && g1 < p.P[1] && g2 < p.P[2];
      long long rowoff = -1;
      if (pv)
        rowoff = (long long)n * p.om.sample_stride +
                 (long long)(g0 * p.om.omul[0] + p.om.obase[0]) * p.om.os[0] +
                 (long long)(g1 * p.om.omul[1] + p.om.obase[1]) * p.om.os[1] +
                 (long long)(g2 * p.om.omul[2] + p.om.obase[2]) * p.om.os[2];
      tc::mbar_wait(tfull_bar + buf, (uint32_t)((local / 2) & 1));
      tc::tc_fence_after();
      const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * p.acc_cols);
      const int ch0 = tcd.y * p.bn;
      for (int col = 0; col < p.bn; col += 16) {
        uint32_t v[16];
        tc::tmem_ld16(lane_base + (uint32_t)col, v);
        if (p.ncat) {
          uint32_t u[16];
          tc::tmem_ld16(lane_base + (uint32_t)(p.bn + col), u);   // the a_hi*b_lo half
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
        }
        if (rowoff >= 0) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            int ch = ch0 + col + j;
            if (ch < p.Nout) p.out[rowoff + (long long)ch * p.om.chan_stride] = __uint_as_float(v[j]);
          }
        }
      }
      tc::tc_fence_before();
      tc::mbar_arrive(tempty_bar + buf);             // 128 arrivals free the accumulator buffer
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ---- the same GEMM on a CTA pair: tcgen05 cta_group::2 ------------------------------------------
//
// Two CTAs of one cluster (same TPC) run ONE UMMA of M = 256: CTA r owns pixel box 2q+r (its
// own A tile and 128 TMEM lanes) and HALF of the filter tile (rows r*bn/2 ...), so every SM
// reads only half of B from its shared memory per MMA and TMA writes half of B per stage —
// the kernel above is bound by exactly that shared-memory traffic.  The leader CTA (rank 0)
// issues the MMAs; both CTAs' TMA transactions complete on the leader's full barrier;
// tcgen05.commit multicasts "stage free" / "accumulator ready" to both CTAs.

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// address of `local` (a shared::cta address of this CTA) inside CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma2_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0,
                                             int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n" ::"r"(dst),
      "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0,
                                             int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];\n" ::"r"(dst),
      "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
// the same box written to the same shared-memory offset of every CTA in `mask`; each destination CTA's
// barrier (same offset) receives the complete_tx
__device__ __forceinline__ void tma_load_5d_mc(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                               int c1, int c2, int c3, int c4, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%3, %4, %5, %6, %7}], [%2], %8;\n" ::"r"(dst),
      "l"(map), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "h"(mask)
      : "memory");
}
// single-CTA MMAs, but "done" is signalled to the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
          tc::smem_u32(bar)),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma2_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at the same offset in BOTH CTAs once the MMAs issued so far retire
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
          tc::smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void tmem2_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(tc::smem_u32(slot)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem2_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(addr), "r"(cols) : "memory");
}

// params: as TmaParams; total_tiles counts PAIR tiles (two pixel boxes x one channel tile),
// pix_tiles = number of real pixel boxes
template <int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
conv_tma2_kernel(const __grid_constant__ TmaParams p, const __grid_constant__ CUtensorMap map_a_hi,
                 const __grid_constant__ CUtensorMap map_a_lo, const __grid_constant__ CUtensorMap map_b_hi,
                 const __grid_constant__ CUtensorMap map_b_lo, int pix_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int half_n = p.bn / 2;
  const int b_tile_bytes = half_n * 128;                 // this CTA's half of the filter tile
  const int stage_bytes = 2 * A_TILE_BYTES + 2 * b_tile_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int S = p.ntaps * p.nchunk;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a_hi); prefetch_tmap(&map_a_lo); prefetch_tmap(&map_b_hi); prefetch_tmap(&map_b_lo);
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(full_bar + s, 2);      // leader's expect_tx arrive + the peer's arrive
      tc::mbar_init(empty_bar + s, 1);     // multicast commit
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(tfull_bar + b, 1);     // multicast commit
      tc::mbar_init(tempty_bar + b, 256);  // epilogue threads of both CTAs (leader's copy is the one waited on)
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) tmem2_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc::tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // barriers of both CTAs are initialised before any remote signal
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem_base = tc::smem_u32(smem);

  // pair tile -> (channel tile y, pixel-box pair q); this CTA's pixel box = 2q + rank
  auto my_box = [&](int pair, TileCoord* tcd, bool* real) {
    const int y = pair % p.ny, q = pair / p.ny;
    const int box = 2 * q + (int)rank;
    *real = box < pix_tiles;
    int t = *real ? box : 0;
    tcd->y = y;
    tcd->t2 = t % p.tiles[2]; t /= p.tiles[2];
    tcd->t1 = t % p.tiles[1]; t /= p.tiles[1];
    tcd->t0 = t % p.tiles[0]; t /= p.tiles[0];
    tcd->n0 = *real ? t * p.boxn : p.Nbatch;      // a padding box reads past the batch: all zeros
  };

  if (warp == 0) {
    if (tc::elect_one()) {
      // ===== TMA producer (both CTAs): own A box, own half of B; completion on the LEADER's barrier =====
      int it = 0;
      for (int pair = cluster_id; pair < p.total_tiles; pair += nclusters) {
        TileCoord tcd;
        bool real;
        my_box(pair, &tcd, &real);
        const int x0 = tcd.t0 * p.box[0] * p.pmul[0] + p.base[0];
        const int x1 = tcd.t1 * p.box[1] * p.pmul[1] + p.base[1];
        const int x2 = tcd.t2 * p.box[2] * p.pmul[2] + p.base[2];
        for (int s = 0; s < S; ++s, ++it) {
          const int stage = it % STAGES;
          if (it >= STAGES) tc::mbar_wait(empty_bar + stage, (uint32_t)((it / STAGES - 1) & 1));
          const int tap = s / p.nchunk, cc = s - tap * p.nchunk;
          const int4 tv = __ldg(reinterpret_cast<const int4*>(p.taps) + tap);
          const uint32_t a_hi = smem_base + stage * stage_bytes;
          const uint32_t bar = map_to_cta(tc::smem_u32(full_bar + stage), 0);
          if (rank == 0) mbar_expect_tx(full_bar + stage, (uint32_t)(2 * stage_bytes));
          else mbar_arrive_cluster(bar);
          tma2_load_5d(a_hi, &map_a_hi, bar, cc * BK, x2 + tv.w, x1 + tv.z, x0 + tv.y, tcd.n0);
          tma2_load_5d(a_hi + A_TILE_BYTES, &map_a_lo, bar, cc * BK, x2 + tv.w, x1 + tv.z, x0 + tv.y, tcd.n0);
          tma2_load_2d(a_hi + 2 * A_TILE_BYTES, &map_b_hi, bar, s * BK, tcd.y * p.bn + (int)rank * half_n);
          tma2_load_2d(a_hi + 2 * A_TILE_BYTES + b_tile_bytes, &map_b_lo, bar, s * BK,
                       tcd.y * p.bn + (int)rank * half_n);
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && tc::elect_one()) {
      // ===== MMA issuer: leader CTA only, M = 256 over the pair =====
      const uint32_t idesc = (tc::idesc_tf32(p.bn) & ~(0x1Fu << 24)) | ((256u >> 4) << 24);
      int it = 0, local = 0;
      for (int pair = cluster_id; pair < p.total_tiles; pair += nclusters, ++local) {
        const int buf = local & 1;
        if (local >= 2) tc::mbar_wait(tempty_bar + buf, (uint32_t)((local / 2 - 1) & 1));
        tc::tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * p.bn);
        for (int s = 0; s < S; ++s, ++it) {
          const int stage = it % STAGES;
          tc::mbar_wait(full_bar + stage, (uint32_t)((it / STAGES) & 1));
          tc::tc_fence_after();
          const uint32_t a_hi = smem_base + stage * stage_bytes;
          const uint32_t a_lo = a_hi + A_TILE_BYTES;
          const uint32_t b_hi = a_lo + A_TILE_BYTES;
          const uint32_t b_lo = b_hi + b_tile_bytes;
#pragma unroll
          for (int j = 0; j < BK / 8; ++j) {
            const uint64_t dah = tc::smem_desc_k_sw128(a_hi + j * 32), dal = tc::smem_desc_k_sw128(a_lo + j * 32);
            const uint64_t dbh = tc::smem_desc_k_sw128(b_hi + j * 32), dbl = tc::smem_desc_k_sw128(b_lo + j * 32);
            umma2_tf32(tmem_d, dal, dbh, idesc, (s | j) ? 1u : 0u);
            umma2_tf32(tmem_d, dah, dbl, idesc, 1u);
            umma2_tf32(tmem_d, dah, dbh, idesc, 1u);
          }
          umma2_commit(empty_bar + stage);
        }
        umma2_commit(tfull_bar + buf);
      }
    }
  } else {
    // ===== epilogue (both CTAs): own 128 TMEM lanes =====
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int i2 = r % p.box[2];
    const int r1 = r / p.box[2];
    const int i1 = r1 % p.box[1];
    const int r0 = r1 / p.box[1];
    const int i0 = r0 % p.box[0];
    const int in = r0 / p.box[0];
    int local = 0;
    for (int pair = cluster_id; pair < p.total_tiles; pair += nclusters, ++local) {
      const int buf = local & 1;
      TileCoord tcd;
      bool real;
      my_box(pair, &tcd, &real);
      const int n = tcd.n0 + in;
      const int g0 = tcd.t0 * p.box[0] + i0, g1 = tcd.t1 * p.box[1] + i1, g2 = tcd.t2 * p.box[2] + i2;
      const bool pv = real && n < p.Nbatch && g0 < p.P[0] && g1 < p.P[1] && g2 < p.P[2];
      long long rowoff = -1;
      if (pv)
        rowoff = (long long)n * p.om.sample_stride +
                 (long long)(g0 * p.om.omul[0] + p.om.obase[0]) * p.om.os[0] +
                 (long long)(g1 * p.om.omul[1] + p.om.obase[1]) * p.om.os[1] +
                 (long long)(g2 * p.om.omul[2] + p.om.obase[2]) * p.om.os[2];
      tc::mbar_wait(tfull_bar + buf, (uint32_t)((local / 2) & 1));
      tc::tc_fence_after();
      const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * p.bn);
      const int ch0 = tcd.y * p.bn;
      for (int col = 0; col < p.bn; col += 16) {
        uint32_t v[16];
        tc::tmem_ld16(lane_base + (uint32_t)col, v);
        if (rowoff >= 0) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            int ch = ch0 + col + j;
            if (ch < p.Nout) p.out[rowoff + (long long)ch * p.om.chan_stride] = __uint_as_float(v[j]);
          }
        }
      }
      tc::tc_fence_before();
      mbar_arrive_cluster(map_to_cta(tc::smem_u32(tempty_bar + buf), 0));   // the leader waits for all 256
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // nobody frees TMEM / exits while the pair still uses it
  if (warp == 1) tmem2_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ---- wgrad with TMA-staged MN-major operands -------------------------------------------------
//
// K = output pixels, walked box by box (b2 x b1 x b0 x bn = 16 pixels per stage); for each
// box one TMA per 32-channel block brings [16 pixels][32 channels] of dy (A) and of the
// tap-shifted x (B) into SWIZZLE_128B_ATOM_32B tiles = UMMA's SWIZZLE_128B_BASE32B.
// Out-of-range pixels (ragged boxes, padding) arrive as zeros and contribute nothing.

constexpr int WG_PIX = 16;                 // pixels per stage of the un-blocked flavour

struct WgradTmaParams {
  const tc::TcTap* taps;                   // d0..d2 = k*dil per tap
  float* partial;                          // [splits][F][C][Wp]
  int box[3], boxn;                        // 16-pixel box (powers of two)
  int tiles[3], tiles_n;                   // boxes per axis
  int S[3], Pd[3];                         // conv stride / padding
  int F, C, Wp, cblocks, nblocks_total, nbpc;
  int boxes_total, boxes_per_split;
  int tmem_cols;
  int pix;                                 // pixels per stage: 16 or 32
  // "blocked" tensor maps: the batch and the outermost spatial axis are merged into one
  // dimension (u = n*X0 + x0: legal when that axis is unpadded and the boxes tile it exactly, and
  // trivially when X0 = 1), the channel axis is split into (32, block) and
  // the block index is the slowest box dimension, so ONE TMA brings several 32-channel
  // blocks ([block][pixel][128 B], exactly the UMMA operand order): 4 f-blocks of dy, and
  // all c-blocks of one tap of x.  Small boxes are what limits TMA here, not bytes.
  int blocked;
  int X0, G0;                              // outermost spatial extents of x and dy (for the merged axis)
  unsigned long long* prof;                // diagnostics (MGB_WGRAD_PROF): per-CTA cycle counters of the MMA thread
  // mc > 1: the `mc` CTAs along blockIdx.y (the N tiles of one F tile and K split) form a thread-block
  // cluster and share the dy operand: CTA 0 fetches the hi plane, CTA 1 the lo plane, each ONCE, multicast
  // into the shared memory of all of them (the kernel is bound by L2 -> SM bytes; dy was fetched mc times).
  // "Stage free" commits are multicast too: a stage is refilled only when every CTA has consumed it.
  int mc;
};

inline size_t wgrad_smem_bytes(int nbpc, int stages, int pix) {
  return (size_t)stages * (2 * 4 * (size_t)pix * 128 + 2 * (size_t)nbpc * pix * 128) + 1024 + 256;
}

template <int STAGES>
__global__ void __launch_bounds__(THREADS, 1)
wgrad_tma_kernel(const __grid_constant__ WgradTmaParams p, const __grid_constant__ CUtensorMap map_dy_hi,
                 const __grid_constant__ CUtensorMap map_dy_lo, const __grid_constant__ CUtensorMap map_x_hi,
                 const __grid_constant__ CUtensorMap map_x_lo) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int WG_BLK_BYTES = p.pix * 128;                 // one 32-channel block of one plane
  const int a_plane = 4 * WG_BLK_BYTES;                 // 128 f-channels
  const int b_plane = p.nbpc * WG_BLK_BYTES;
  const int stage_bytes = 2 * a_plane + 2 * b_plane;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM;
  const int blk0 = blockIdx.y * p.nbpc;
  const int nvalid = min(p.nbpc, p.nblocks_total - blk0);     // (tap, c-block) blocks of this CTA
  const int kb0 = blockIdx.z * p.boxes_per_split;
  const int kb1 = min(p.boxes_total, kb0 + p.boxes_per_split);
  const int S = max(0, kb1 - kb0);

  const int mc = p.mc;
  const uint32_t crank = mc > 1 ? cluster_ctarank() : 0u;
  const uint16_t mc_mask = (uint16_t)((1u << mc) - 1u);
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_dy_hi); prefetch_tmap(&map_dy_lo); prefetch_tmap(&map_x_hi); prefetch_tmap(&map_x_lo);
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(full_bar + s, 1);
      tc::mbar_init(empty_bar + s, (uint32_t)mc);      // one (multicast) commit per CTA of the cluster
    }
    tc::mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc::tc_fence_before();
  __syncthreads();
  if (mc > 1) cluster_sync_all();             // every CTA's barriers exist before a peer's TMA or commit signals them
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem_base = tc::smem_u32(smem);

  if (warp == 0) {
    if (tc::elect_one()) {
      // ===== TMA producer =====
      const uint32_t tx = (uint32_t)(2 * a_plane + 2 * nvalid * WG_BLK_BYTES);
      for (int s = 0; s < S; ++s) {
        const int stage = s % STAGES;
        if (s >= STAGES) tc::mbar_wait(empty_bar + stage, (uint32_t)((s / STAGES - 1) & 1));
        int kb = kb0 + s;
        const int t2 = kb % p.tiles[2]; kb /= p.tiles[2];
        const int t1 = kb % p.tiles[1]; kb /= p.tiles[1];
        const int t0 = kb % p.tiles[0]; kb /= p.tiles[0];
        const int n0 = kb * p.boxn;
        const int g0 = t0 * p.box[0], g1 = t1 * p.box[1], g2 = t2 * p.box[2];
        const uint32_t a_hi = smem_base + stage * stage_bytes;
        mbar_expect_tx(full_bar + stage, tx);
        const uint32_t b_hi = a_hi + 2 * a_plane;
        const int x0 = g0 * p.S[0] - p.Pd[0], x1 = g1 * p.S[1] - p.Pd[1], x2 = g2 * p.S[2] - p.Pd[2];
        if (p.blocked) {
          // coordinates (ch-in-block, x2, x1, u = n*X0 + x0, block)
          const int uy = n0 * p.G0 + g0, ux = n0 * p.X0 + x0;
          if (mc > 1) {
            if (crank == 0u) tma_load_5d_mc(a_hi, &map_dy_hi, full_bar + stage, 0, g2, g1, uy, m0 / 32, mc_mask);
            if (crank == 1u) tma_load_5d_mc(a_hi + a_plane, &map_dy_lo, full_bar + stage, 0, g2, g1, uy, m0 / 32, mc_mask);
          } else {
            tma_load_5d(a_hi, &map_dy_hi, full_bar + stage, 0, g2, g1, uy, m0 / 32);
            tma_load_5d(a_hi + a_plane, &map_dy_lo, full_bar + stage, 0, g2, g1, uy, m0 / 32);
          }
          for (int b = 0; b < nvalid; b += p.cblocks) {
            const int tap = (blk0 + b) / p.cblocks;
            const int4 tv = __ldg(reinterpret_cast<const int4*>(p.taps) + tap);
            tma_load_5d(b_hi + b * WG_BLK_BYTES, &map_x_hi, full_bar + stage, 0, x2 + tv.w, x1 + tv.z,
                        ux + tv.y, 0);
            tma_load_5d(b_hi + b_plane + b * WG_BLK_BYTES, &map_x_lo, full_bar + stage, 0, x2 + tv.w,
                        x1 + tv.z, ux + tv.y, 0);
          }
        } else {
          for (int mb = 0; mb < 4; ++mb) {
            tma_load_5d(a_hi + mb * WG_BLK_BYTES, &map_dy_hi, full_bar + stage, m0 + mb * 32, g2, g1, g0, n0);
            tma_load_5d(a_hi + a_plane + mb * WG_BLK_BYTES, &map_dy_lo, full_bar + stage, m0 + mb * 32, g2,
                        g1, g0, n0);
          }
          for (int b = 0; b < nvalid; ++b) {
            const int gb = blk0 + b;
            const int tap = gb / p.cblocks, cb = gb - tap * p.cblocks;
            const int4 tv = __ldg(reinterpret_cast<const int4*>(p.taps) + tap);
            tma_load_5d(b_hi + b * WG_BLK_BYTES, &map_x_hi, full_bar + stage, cb * 32, x2 + tv.w, x1 + tv.z,
                        x0 + tv.y, n0);
            tma_load_5d(b_hi + b_plane + b * WG_BLK_BYTES, &map_x_lo, full_bar + stage, cb * 32, x2 + tv.w,
                        x1 + tv.z, x0 + tv.y, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {
      // ===== MMA issuer: A and B MN-major, N = nvalid*32 =====
      const uint32_t idesc = tc::idesc_tf32(nvalid * 32) | (1u << 15) | (1u << 16);
      long long wf = 0, t0 = 0, c0 = 0;
      if (p.prof) t0 = clock64();
      for (int s = 0; s < S; ++s) {
        const int stage = s % STAGES;
        if (p.prof) c0 = clock64();
        tc::mbar_wait(full_bar + stage, (uint32_t)((s / STAGES) & 1));
        if (p.prof) wf += clock64() - c0;
        tc::tc_fence_after();
        const uint32_t a_hi = smem_base + stage * stage_bytes;
        const uint32_t a_lo = a_hi + a_plane;
        const uint32_t b_hi = a_lo + a_plane;
        const uint32_t b_lo = b_hi + b_plane;
        for (int j = 0; j < p.pix / 8; ++j) {
          // block (32 channels) stride = pix*128 B (LBO), 4-pixel k-atom stride = 512 B (SBO)
          const uint64_t dah = tc::smem_desc_mn_sw128(a_hi + j * 1024, (uint32_t)WG_BLK_BYTES, 512u);
          const uint64_t dal = tc::smem_desc_mn_sw128(a_lo + j * 1024, (uint32_t)WG_BLK_BYTES, 512u);
          const uint64_t dbh = tc::smem_desc_mn_sw128(b_hi + j * 1024, (uint32_t)WG_BLK_BYTES, 512u);
          const uint64_t dbl = tc::smem_desc_mn_sw128(b_lo + j * 1024, (uint32_t)WG_BLK_BYTES, 512u);
          tc::umma_tf32(tmem_base, dal, dbh, idesc, (s | j) ? 1u : 0u);
          tc::umma_tf32(tmem_base, dah, dbl, idesc, 1u);
          tc::umma_tf32(tmem_base, dah, dbh, idesc, 1u);
        }
        if (mc > 1) umma_commit_mc(empty_bar + stage, mc_mask);
        else tc::umma_commit(empty_bar + stage);
      }
      if (S > 0) tc::umma_commit(accum_bar);
      if (p.prof) {
        const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        p.prof[cta * 4 + 0] = (unsigned long long)wf;
        p.prof[cta * 4 + 1] = (unsigned long long)(clock64() - t0);
        p.prof[cta * 4 + 2] = (unsigned long long)S;
      }
    }
  } else {
    // ===== epilogue: lane = f, columns = (block, channel) -> partial[f][c][tap] =====
    const int quarter = warp & 3;
    const int f = m0 + quarter * 32 + lane;
    float* out = p.partial + (long long)blockIdx.z * p.F * p.C * p.Wp;
    if (S > 0) {
      tc::mbar_wait(accum_bar, 0);
      tc::tc_fence_after();
    }
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    for (int col = 0; col < nvalid * 32; col += 16) {
      uint32_t v[16];
      if (S > 0) {
        tc::tmem_ld16(lane_base + (uint32_t)col, v);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0u;
      }
      const int gb = blk0 + (col >> 5);
      const int tap = gb / p.cblocks, cb = gb - tap * p.cblocks;
      if (f < p.F) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          int c = cb * 32 + (col & 31) + j;
          if (c < p.C) out[((long long)f * p.C + c) * p.Wp + tap] = __uint_as_float(v[j]);
        }
      }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (mc > 1) cluster_sync_all();             // no CTA leaves while a peer may still write into it or signal it
  if (warp == 1) tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

}  // namespace tma
}  // namespace conv
}  // namespace mgb
