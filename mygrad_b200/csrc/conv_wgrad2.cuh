// wgrad with a ROW HALO of x: one TMA box serves every filter tap along one spatial axis.
//
// What bounds wgrad_tma_kernel (conv_tma.cuh), measured with MGB_WGRAD_PROF at C2: every CTA pulls
// 80 KB per 32-pixel stage from L2 (32 KB of dy + six tap-shifted 8 KB blocks of x) and a stage takes
// 1890 clk = 42 B/clk per SM — the chip-wide L2 -> SM limit (~6300 B/clk over 148 SMs), not shared memory and
// not the tensor pipe (1152 clk).  A CTA-pair variant with swapped operand roles (M = x blocks, dy shared
// across the pair) moved the same 240 KB per 32 pixels over the three CTAs and ran at the same 42 B/clk.
// The bytes are what has to go: x is fetched once per TAP, nine times at 3x3.
//
// Here a CTA owns ONE tap of the innermost filter axis (and of the outermost, for 3-D) and ALL taps of the
// middle axis: its x box is the pixel box grown by the tap extent along that axis, stored as
//     [halo row][c-block][column][32 channels]            (tensor-map dimension order (c, x2, block, x1, n*x0))
// so that the (row tap dr, c-block cb) operand block is the same buffer entered (dr * cbc + cb) * b2 * 128 B later:
// a uniform block stride (the descriptor's LBO), whole 512-byte swizzle atoms, no data movement.  The N tile of
// the UMMA is W1 * cbc blocks; K walks the 32 pixels of the box, 8 per UMMA (two 4-pixel k-atoms: the next
// row, SBO = one halo row, when the box is 4 pixels wide, else the next 4 pixels of the same row).
// At C2 a stage moves 32 KB of dy + 20 KB of x instead of 80 KB.
//
// Needs: conv stride 1 along the middle axis, dilation 1 along it when a CTA holds several c-blocks, boxes that
// do not span samples, W1 * cbc <= 8.  Everything else goes to wgrad_tma_kernel.
//
// FOLD (F <= 64): a 128-lane UMMA with 64 filters wastes half of every instruction.  The upper 64 lanes then get
// the SAME filters' dy shifted by delta = D2 pixels along the innermost axis:
//     D[64 + f, (tap, c)] = sum_pixels dy[pixel + delta, f] * x[pixel + tap, c] = dw[f, c, tap - 1 column tap]
// so one CTA covers the column taps dc and dc - 1 and two CTAs cover three.  The pixel boxes start delta columns
// early (those columns of the unshifted dy are TMA zero fill), conv stride 1 along that axis.
#pragma once
#include "conv_tma.cuh"

namespace mgb {
namespace conv {
namespace tma {

struct WgradRowsParams {
  float* partial;                          // [splits][F][C][Wp]
  int b1, b2;                              // pixel box: rows x columns, b1 * b2 = 32, b2 >= 4
  int tiles1, tiles2;                      // boxes per image along the middle / innermost axis
  int G0, X0;                              // outermost extents of dy / x (merged with the batch: u = n * G0 + g0)
  int S0, S2, Pd1, Pd2, D0, D1, D2;        // stride / padding / dilation that the coordinates need
  int W0, W1, W2;
  int F, C, Wp, cblocks, cbc;              // c-blocks in all / per CTA
  int rows;                                // halo rows: b1 + (W1 - 1) * D1
  int boxes_total, boxes_per_split;
  int tmem_cols;
  int fold;                                // 1: lanes 64..127 hold dy shifted by D2 columns (F <= 64), see above
  unsigned long long* prof;                // MGB_WGRAD_PROF: [cta][4] = wait-full clocks, loop clocks, stages
};

inline size_t wgrad_rows_smem_bytes(int rows, int cbc, int b2, int stages) {
  return (size_t)stages * (2 * 4 * 32 * 128 + 2 * (size_t)rows * cbc * b2 * 128) + 1024 + 256;
}

// grid = (F tiles of 128, W0 * W2 * (cblocks / cbc), K splits)
template <int STAGES>
__global__ void __launch_bounds__(THREADS, 1)
wgrad_rows_kernel(const __grid_constant__ WgradRowsParams p, const __grid_constant__ CUtensorMap map_dy_hi,
                  const __grid_constant__ CUtensorMap map_dy_lo, const __grid_constant__ CUtensorMap map_x_hi,
                  const __grid_constant__ CUtensorMap map_x_lo) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int a_plane = 4 * 32 * 128;                  // 128 filters x 32 pixels
  const int row_bytes = p.cbc * p.b2 * 128;              // one halo row: [c-block][column][128 B]
  const int b_plane = p.rows * row_bytes;
  const int stage_bytes = 2 * a_plane + 2 * b_plane;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM;
  int y = blockIdx.y;
  const int ncbg = p.cblocks / p.cbc;
  const int cb0 = (y % ncbg) * p.cbc; y /= ncbg;
  const int ndc = p.fold ? (p.W2 + 1) / 2 : p.W2;        // column-tap CTAs
  const int jdc = y % ndc;
  const int dc = p.fold ? min(2 * jdc + 1, p.W2 - 1) : jdc;
  const bool upper = p.fold && dc >= 1 && (2 * jdc + 1 <= p.W2 - 1);   // lanes 64..127 own column tap dc - 1
  const int shift = p.fold ? p.D2 : 0;                    // boxes start this many columns early
  const int t0 = y / ndc;
  const int nblk = p.W1 * p.cbc;                         // operand blocks of the N tile: (dr, cb)
  const int kb0 = blockIdx.z * p.boxes_per_split;
  const int kb1 = min(p.boxes_total, kb0 + p.boxes_per_split);
  const int S = max(0, kb1 - kb0);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_dy_hi); prefetch_tmap(&map_dy_lo); prefetch_tmap(&map_x_hi); prefetch_tmap(&map_x_lo);
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(full_bar + s, 1);
      tc::mbar_init(empty_bar + s, 1);
    }
    tc::mbar_init(accum_bar, 1);
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
      // ===== TMA producer: per stage one box of dy and one row halo of x per plane =====
      const uint32_t tx = (uint32_t)stage_bytes;
      for (int s = 0; s < S; ++s) {
        const int stage = s % STAGES;
        if (s >= STAGES) tc::mbar_wait(empty_bar + stage, (uint32_t)((s / STAGES - 1) & 1));
        int kb = kb0 + s;
        const int t2 = kb % p.tiles2; kb /= p.tiles2;
        const int t1 = kb % p.tiles1; kb /= p.tiles1;
        const int u = kb;                                 // n * G0 + g0
        const int n = u / p.G0, g0 = u - n * p.G0;
        const int g1 = t1 * p.b1, g2 = t2 * p.b2 - shift;
        const int ux = n * p.X0 + g0 * p.S0 + t0 * p.D0;  // the outermost axis is unpadded
        const int x1 = g1 - p.Pd1;
        const int x2 = g2 * p.S2 - p.Pd2 + dc * p.D2;
        const uint32_t a_hi = smem_base + stage * stage_bytes;
        const uint32_t b_hi = a_hi + 2 * a_plane;
        mbar_expect_tx(full_bar + stage, tx);
        if (p.fold) {       // two filter blocks unshifted, the same two shifted by D2 columns
          tma_load_5d(a_hi, &map_dy_hi, full_bar + stage, 0, g2, g1, u, 0);
          tma_load_5d(a_hi + a_plane / 2, &map_dy_hi, full_bar + stage, 0, g2 + shift, g1, u, 0);
          tma_load_5d(a_hi + a_plane, &map_dy_lo, full_bar + stage, 0, g2, g1, u, 0);
          tma_load_5d(a_hi + a_plane + a_plane / 2, &map_dy_lo, full_bar + stage, 0, g2 + shift, g1, u, 0);
        } else {
          tma_load_5d(a_hi, &map_dy_hi, full_bar + stage, 0, g2, g1, u, m0 / 32);
          tma_load_5d(a_hi + a_plane, &map_dy_lo, full_bar + stage, 0, g2, g1, u, m0 / 32);
        }
        tma_load_5d(b_hi, &map_x_hi, full_bar + stage, 0, x2, cb0, x1, ux);
        tma_load_5d(b_hi + b_plane, &map_x_lo, full_bar + stage, 0, x2, cb0, x1, ux);
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {
      // ===== MMA issuer: A (dy) and B (x) MN-major, N = nblk * 32 =====
      const uint32_t idesc = tc::idesc_tf32(nblk * 32) | (1u << 15) | (1u << 16);
      const uint32_t b_lbo = (uint32_t)(p.b2 * 128 * (p.cbc == 1 ? p.D1 : 1));
      const uint32_t b_sbo = p.b2 == 4 ? (uint32_t)row_bytes : 512u;
      const int apr = p.b2 / 4;                           // k-atoms (4 pixels) per box row
      long long wf = 0, t_start = 0, c0 = 0;
      if (p.prof) t_start = clock64();
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
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int atom = 2 * j;
          const uint32_t boff = (uint32_t)((atom / apr) * row_bytes + (atom % apr) * 512);
          const uint64_t dah = tc::smem_desc_mn_sw128(a_hi + j * 1024, 4096u, 512u);
          const uint64_t dal = tc::smem_desc_mn_sw128(a_lo + j * 1024, 4096u, 512u);
          const uint64_t dbh = tc::smem_desc_mn_sw128(b_hi + boff, b_lbo, b_sbo);
          const uint64_t dbl = tc::smem_desc_mn_sw128(b_lo + boff, b_lbo, b_sbo);
          tc::umma_tf32(tmem_base, dal, dbh, idesc, (s | j) ? 1u : 0u);
          tc::umma_tf32(tmem_base, dah, dbl, idesc, 1u);
          tc::umma_tf32(tmem_base, dah, dbh, idesc, 1u);
        }
        tc::umma_commit(empty_bar + stage);
      }
      if (S > 0) tc::umma_commit(accum_bar);
      if (p.prof) {
        const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        p.prof[cta * 4 + 0] = (unsigned long long)wf;
        p.prof[cta * 4 + 1] = (unsigned long long)(clock64() - t_start);
        p.prof[cta * 4 + 2] = (unsigned long long)S;
      }
    }
  } else {
    // ===== epilogue: lane = f, columns = (dr, cb, channel) -> partial[f][c][tap] =====
    const int quarter = warp & 3;
    int f = m0 + quarter * 32 + lane;
    int dcw = dc;                                          // column tap this lane writes
    bool wr = true;
    if (p.fold && f >= 64) { f -= 64; dcw = dc - 1; wr = upper; }
    float* out = p.partial + (long long)blockIdx.z * p.F * p.C * p.Wp;
    if (S > 0) {
      tc::mbar_wait(accum_bar, 0);
      tc::tc_fence_after();
    }
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    for (int col = 0; col < nblk * 32; col += 16) {
      uint32_t v[16];
      if (S > 0) {
        tc::tmem_ld16(lane_base + (uint32_t)col, v);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0u;
      }
      const int b = col >> 5;
      const int dr = b / p.cbc, cb = cb0 + (b - dr * p.cbc);
      const int tap = (t0 * p.W1 + dr) * p.W2 + dcw;
      if (wr && f < p.F) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int c = cb * 32 + (col & 31) + j;
          if (c < p.C) out[((long long)f * p.C + c) * p.Wp + tap] = __uint_as_float(v[j]);
        }
      }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

}  // namespace tma
}  // namespace conv
}  // namespace mgb
