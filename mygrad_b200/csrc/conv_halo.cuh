// FP32 conv fwd / dgrad on tcgen05 with INPUT-HALO reuse across filter taps.
//
// conv_tma_kernel (conv_tma.cuh) re-loads the 128-pixel A tile once per filter tap: at a 3x3
// filter every input pixel crosses TMA -> shared memory nine times, and those writes compete with
// the UMMA operand reads for the same 128 B/clk of shared-memory bandwidth (the bound of that
// kernel).  Here ONE TMA brings the whole input halo of the tile
//
//     [rows + tap extent] x [bn samples] x [8 + tap extent columns] x 32 channels
//
// and every (row tap, column tap) of the filter is a UMMA whose A descriptor merely STARTS at a
// different 128-byte row of that halo:
//
//   * an M tile is (bh rows) x (bn samples) x (8 columns), bh*bn = 16.  The tensor map lists the
//     dimensions as (C, X_col, N, X_row, X_outer), so the halo lands in shared memory as
//     [row][sample][column][128 B]: the 8 pixels of one UMMA core-matrix group are 8 consecutive
//     128-byte rows and group g = row*bn + sample sits at g * (halo_cols*128) bytes — a constant
//     stride, which is all the K-major SWIZZLE_128B descriptor needs (SBO = halo_cols*128; the
//     hardware applies the swizzle XOR on absolute shared-memory address bits, so a descriptor may
//     start on any 128-byte row and SBO need not be a multiple of 1024).
//   * out-of-range rows / columns / samples are zero-filled by TMA == zero padding and ragged tiles.
//   * taps along the third spatial axis of a 3-D convolution are an outer loop (one halo each).
//
// Two rings: A (halo per 32-channel chunk and outer tap, 2 stages, own producer thread) and
// B (packed filter slice per tap, hi|lo adjacent).  Math as in conv_tma.cuh: 3xTF32 with the
// N-concatenated B operand when 2*bn <= 256.  Accumulators double-buffered in TMEM.
//
// Applies to stride-1 gathers (forward with stride 1, every dgrad residue class).
#pragma once
#include "conv_tma.cuh"

namespace mgb {
namespace conv {
namespace halo {

using tc::BK;

constexpr int THREADS = 224;     // warp 0: B producer, warp 1: MMA + TMEM, warps 2-5: epilogue, warp 6: A producer
constexpr int AST = 2;           // halo stages
constexpr int kMaxTap = 32;
constexpr int kMaxInner = 256;   // (row taps) x (column taps) the tap-offset table holds

struct HaloParams {
  float* out;
  OutMap om;
  int P[3];                      // pixel grid per sample (normalised to 3 axes)
  int base[3];                   // source coord = pixel + base + tap delta
  int ra, oa;                    // row axis (0 or 1) and the other ("outer") axis
  int bh, bnb;                   // rows and samples per M tile (bh*bnb == 16); 8 columns
  int tiles_r, tiles_c, tiles_n; // tile counts along rows, columns, batch
  int Nbatch, Nout;
  int nchunk;                    // cpad / 32
  int cnt[3];                    // taps per axis
  short delta[3][kMaxTap];       // coordinate delta of each tap, per axis
  int minr, minc;                // smallest row / column delta (halo origin)
  int hr, hc;                    // halo rows / columns
  int a_plane;                   // bytes of one halo plane in smem (rounded to 1024)
  int a_tx;                      // bytes TMA delivers per plane (hr*bnb*hc*128)
  int bn, ny;                    // channel tile, channel tiles
  int bst;                       // B stages
  int ncat, acc_cols, tmem_cols;
  int total_tiles;
  int dbg;                       // timing experiments only (MGB_HALO2_DBG): 1 = MMA thread skips the B-stage wait
  int mtiles;                    // pair kernel (conv_halo2.cuh): real M tiles; total_tiles then counts PAIR tiles
  // (halo byte offset >> 4) of each inner tap, row tap major: kernel parameters are read through the
  // uniform datapath (ULDC), so the descriptors the MMA thread builds from them stay in uniform registers
  uint32_t tapoff[kMaxInner];
  unsigned long long* prof;      // PROF instantiation only: per-CTA cycle counters of the MMA thread
  // fused 2x2 / stride-2 max-pool over (row axis, column axis) in the epilogue: the
  // accumulator tile never goes to HBM as y; each window's maximum and the position of its FIRST maximum
  // in row-major window order (pooling.py:111, NaN counts as maximal) are written instead
  int pool;                      // 0: store y;  1: store pooled + argmax (needs bnb <= 2, even P[ra], P[2])
  float* pout;                   // [N][Nout][P[ra]/2][P[2]/2]
  unsigned char* pidx;           // same shape: 2*(row & 1) + (col & 1) of the first maximum
  long long p_sample_stride, p_chan_stride;
  int pw;                        // pooled columns per row = P[2] / 2
};

__device__ __forceinline__ uint64_t smem_desc_k_sw128_sbo(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// D[tmem] (+)= A * B with the descriptors given as (low, high) 32-bit words
__device__ __forceinline__ void umma_tf32_w(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                            uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_acc(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                              uint32_t b_hi, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.eq.b32 p, 0, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_addr(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t addr, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}

// maximum that propagates NaN (IEEE 754-2019 `maximum`): NaN if either input is NaN
__device__ __forceinline__ float fmax_nan(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;\n" : "=f"(r) : "f"(a), "f"(b));
  return r;
}

// `if (pred) { *a = v; *b = (uint8_t)k; }` as two predicated stores
__device__ __forceinline__ void st_global_if(float* a, float v, unsigned char* b, int k, bool pred) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.s32 q, %4, 0;\n\t"
      "@q st.global.f32 [%0], %1;\n\t"
      "@q st.global.u8 [%2], %3;\n\t}\n" ::"l"(a),
      "f"(v), "l"(b), "r"(k), "r"((int)pred)
      : "memory");
}

struct HaloTile {
  int n0, xo, r0, c0, y;
};

__device__ __forceinline__ HaloTile decode_tile(const HaloParams& p, int tile) {
  HaloTile t;
  t.y = tile % p.ny; tile /= p.ny;
  t.c0 = (tile % p.tiles_c) * 8; tile /= p.tiles_c;
  t.r0 = (tile % p.tiles_r) * p.bh; tile /= p.tiles_r;
  t.xo = tile % p.P[p.oa]; tile /= p.P[p.oa];
  t.n0 = tile * p.bnb;
  return t;
}

inline size_t smem_bytes(int a_plane, int bn, int bst) {
  return (size_t)AST * 2 * a_plane + (size_t)bst * 2 * bn * 128 + 1024 + 256;
}

template <bool PROF>
__global__ void __launch_bounds__(THREADS, 1)
conv_halo_kernel(const __grid_constant__ HaloParams p, const __grid_constant__ CUtensorMap map_a_hi,
                 const __grid_constant__ CUtensorMap map_a_lo, const __grid_constant__ CUtensorMap map_b_hi,
                 const __grid_constant__ CUtensorMap map_b_lo) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int a_stage = 2 * p.a_plane;
  const int b_tile = p.bn * 128;
  const int b_stage = 2 * b_tile;
  uint8_t* b_ring = smem + AST * a_stage;
  uint64_t* afull = reinterpret_cast<uint64_t*>(b_ring + p.bst * b_stage);
  uint64_t* aempty = afull + AST;
  uint64_t* bfull = aempty + AST;
  uint64_t* bempty = bfull + p.bst;
  uint64_t* tfull = bempty + p.bst;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  volatile uint32_t* ts_issue = tmem_slot + 2;     // PROF: issue time of the TMA filling each B stage

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ra = p.ra, oa = p.oa;
  const int cnt_o = p.cnt[oa], cnt_r = p.cnt[ra], cnt_c = p.cnt[2];

  if (warp == 0 && lane == 0) {
    tma::prefetch_tmap(&map_a_hi); tma::prefetch_tmap(&map_a_lo);
    tma::prefetch_tmap(&map_b_hi); tma::prefetch_tmap(&map_b_lo);
    for (int s = 0; s < AST; ++s) { tc::mbar_init(afull + s, 1); tc::mbar_init(aempty + s, 1); }
    for (int s = 0; s < p.bst; ++s) { tc::mbar_init(bfull + s, 1); tc::mbar_init(bempty + s, 1); }
    for (int b = 0; b < 2; ++b) { tc::mbar_init(tfull + b, 1); tc::mbar_init(tempty + b, 128); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t a_base = tc::smem_u32(smem);
  const uint32_t b_base = tc::smem_u32(b_ring);

  if (warp == 6) {
    if (tc::elect_one()) {
      // ===== A producer: one halo (hi, lo) per (tile, 32-channel chunk, outer tap) =====
      int ita = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const HaloTile t = decode_tile(p, tile);
        const int xr = t.r0 + p.base[ra] + p.minr;
        const int xc = t.c0 + p.base[2] + p.minc;
        for (int cc = 0; cc < p.nchunk; ++cc)
          for (int jo = 0; jo < cnt_o; ++jo, ++ita) {
            const int sa = ita % AST;
            if (ita >= AST) tc::mbar_wait(aempty + sa, (uint32_t)((ita / AST - 1) & 1));
            const uint32_t dst = a_base + sa * a_stage;
            tma::mbar_expect_tx(afull + sa, (uint32_t)(2 * p.a_tx));
            const int xo = t.xo + p.base[oa] + p.delta[oa][jo];
            tma::tma_load_5d(dst, &map_a_hi, afull + sa, cc * BK, xc, t.n0, xr, xo);
            tma::tma_load_5d(dst + p.a_plane, &map_a_lo, afull + sa, cc * BK, xc, t.n0, xr, xo);
          }
      }
    }
  } else if (warp == 0) {
    if (tc::elect_one()) {
      // ===== B producer: packed filter slice (hi | lo) of one tap and chunk =====
      int itb = 0;
      long long pw = 0, pt0 = 0, pc = 0;
      if (PROF) pt0 = clock64();
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int y = tile % p.ny;
        for (int cc = 0; cc < p.nchunk; ++cc)
          for (int jo = 0; jo < cnt_o; ++jo)
            for (int jr = 0; jr < cnt_r; ++jr) {
              const int j0 = ra == 0 ? jr : jo, j1 = ra == 0 ? jo : jr;
              const int tap0 = (j0 * p.cnt[1] + j1) * cnt_c;
              for (int jc = 0; jc < cnt_c; ++jc, ++itb) {
                const int sb = itb % p.bst;
                if (PROF) pc = clock64();
                if (itb >= p.bst) tc::mbar_wait(bempty + sb, (uint32_t)((itb / p.bst - 1) & 1));
                if (PROF) pw += clock64() - pc;
                const uint32_t dst = b_base + sb * b_stage;
                tma::mbar_expect_tx(bfull + sb, (uint32_t)b_stage);
                const int k0 = ((tap0 + jc) * p.nchunk + cc) * BK;
                if (PROF) ts_issue[sb] = (uint32_t)clock64();
                tma::tma_load_2d(dst, &map_b_hi, bfull + sb, k0, y * p.bn);
                tma::tma_load_2d(dst + b_tile, &map_b_lo, bfull + sb, k0, y * p.bn);
              }
            }
      }
      if (PROF && p.prof) {
        p.prof[blockIdx.x * 8 + 6] = (unsigned long long)pw;
        p.prof[blockIdx.x * 8 + 7] = (unsigned long long)(clock64() - pt0);
      }
    }
  } else if (warp == 1) {
    // the TMEM base address was read from shared memory (a per-thread value to the compiler): a warp
    // reduction hands it back in a uniform register, so tcgen05.mma needs no per-lane broadcast loop
    const uint32_t tmem_base_u = __reduce_or_sync(0xffffffffu, tmem_base);
    if (tc::elect_one()) {
      // ===== MMA issuer =====
      // One thread feeds the tensor pipe with UMMAs of only 128 x N x 8 (kind::tf32 has K = 8), so
      // the instruction count per UMMA is what bounds this kernel at small N.  Descriptors are kept
      // as 32-bit words: the high word is a kernel constant, the low word is (address >> 4) | LBO and
      // advances by plain adds; ring positions and phases are counters, not divisions; the tap
      // offsets come from a table in the kernel parameters.
      const uint32_t idesc = tc::idesc_tf32(p.bn);
      const uint32_t idesc2 = tc::idesc_tf32(2 * p.bn);
      const uint32_t a_dhi = (((uint32_t)p.hc * 128u) >> 4) | (1u << 14) | (2u << 29);   // SBO | version | SWIZZLE_128B
      const uint32_t b_dhi = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t a_lo0 = (a_base >> 4) | 0x10000u;          // LBO field = 1 (unused for swizzled K-major)
      const uint32_t b_lo0 = (b_base >> 4) | 0x10000u;
      const uint32_t a_stage16 = (uint32_t)a_stage >> 4, a_plane16 = (uint32_t)p.a_plane >> 4;
      const uint32_t b_stage16 = (uint32_t)b_stage >> 4, b_tile16 = (uint32_t)b_tile >> 4;
      const uint32_t bfull0 = tc::smem_u32(bfull), bempty0 = tc::smem_u32(bempty);
      const int ninner = cnt_r * cnt_c;
      const bool ncat = p.ncat != 0;
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
      int local = 0;
      long long wa = 0, wb = 0, wt = 0, wi = 0, wc = 0, t_start = 0, c0 = 0, c1 = 0, lat_sum = 0, lat_n = 0;
      if (PROF) t_start = clock64();
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
        const int buf = local & 1;
        if (PROF) c0 = clock64();
        if (local >= 2) tc::mbar_wait(tempty + buf, (uint32_t)((local / 2 - 1) & 1));
        if (PROF) wt += clock64() - c0;
        tc::tc_fence_after();
        const uint32_t tmem_d = tmem_base_u + (uint32_t)(buf * p.acc_cols);
        bool first = true;
        for (int st = 0; st < p.nchunk * cnt_o; ++st) {
          if (PROF) c0 = clock64();
          tc::mbar_wait(afull + sa, pa);
          if (PROF) wa += clock64() - c0;
          tc::tc_fence_after();
          const uint32_t halo16 = a_lo0 + sa * a_stage16;
          uint32_t off16 = p.tapoff[0];
          for (int t = 0; t < ninner; ++t) {
            const uint32_t ah = halo16 + off16, al = ah + a_plane16;
            const uint32_t bh = b_lo0 + sb * b_stage16, bl = bh + b_tile16;
            if (t + 1 < ninner) off16 = p.tapoff[t + 1];          // fetched before the barrier wait
            if (PROF) c0 = clock64();
            mbar_wait_addr(bfull0 + sb * 8, pb);
            if (PROF) {
              c1 = clock64(); wb += c1 - c0;
              if (c1 - c0 > 60) { lat_sum += (uint32_t)c1 - ts_issue[sb]; ++lat_n; }
            }
            tc::tc_fence_after();
            if (ncat) {
              if (first) umma_tf32_w(tmem_d, ah, a_dhi, bh, b_dhi, idesc2, 0u);
              else       umma_tf32_acc(tmem_d, ah, a_dhi, bh, b_dhi, idesc2);
              umma_tf32_acc(tmem_d, al, a_dhi, bh, b_dhi, idesc);
#pragma unroll
              for (int j = 1; j < BK / 8; ++j) {
                umma_tf32_acc(tmem_d, ah + 2 * j, a_dhi, bh + 2 * j, b_dhi, idesc2);
                umma_tf32_acc(tmem_d, al + 2 * j, a_dhi, bh + 2 * j, b_dhi, idesc);
              }
            } else {
              if (first) umma_tf32_w(tmem_d, al, a_dhi, bh, b_dhi, idesc, 0u);
              else       umma_tf32_acc(tmem_d, al, a_dhi, bh, b_dhi, idesc);
              umma_tf32_acc(tmem_d, ah, a_dhi, bl, b_dhi, idesc);
              umma_tf32_acc(tmem_d, ah, a_dhi, bh, b_dhi, idesc);
#pragma unroll
              for (int j = 1; j < BK / 8; ++j) {
                umma_tf32_acc(tmem_d, al + 2 * j, a_dhi, bh + 2 * j, b_dhi, idesc);
                umma_tf32_acc(tmem_d, ah + 2 * j, a_dhi, bl + 2 * j, b_dhi, idesc);
                umma_tf32_acc(tmem_d, ah + 2 * j, a_dhi, bh + 2 * j, b_dhi, idesc);
              }
            }
            first = false;
            if (PROF) { c0 = clock64(); wi += c0 - c1; }
            umma_commit_addr(bempty0 + sb * 8);
            if (PROF) wc += clock64() - c0;
            if (++sb == (uint32_t)p.bst) { sb = 0; pb ^= 1u; }
          }
          tc::umma_commit(aempty + sa);
          if (++sa == AST) { sa = 0; pa ^= 1u; }
        }
        tc::umma_commit(tfull + buf);
      }
      if (PROF && p.prof) {
        unsigned long long* o = p.prof + blockIdx.x * 8;
        o[0] = (unsigned long long)wa; o[1] = (unsigned long long)wb; o[2] = (unsigned long long)wt;
        o[3] = (unsigned long long)(clock64() - t_start); o[4] = (unsigned long long)wi; o[5] = (unsigned long long)wc;
        p.prof[8 * 1024 + blockIdx.x * 2] = (unsigned long long)lat_sum;
        p.prof[8 * 1024 + blockIdx.x * 2 + 1] = (unsigned long long)lat_n;
      }
    }
  } else {
    // ===== epilogue warps 2..5: TMEM lane quarter = warp % 4; lane = (row, sample, column) =====
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int ic = r & 7;
    const int g = r >> 3;
    const int in = g % p.bnb, ir = g / p.bnb;
    int local = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
      const int buf = local & 1;
      const HaloTile t = decode_tile(p, tile);
      const int n = t.n0 + in;
      int gq[3];
      gq[ra] = t.r0 + ir; gq[oa] = t.xo; gq[2] = t.c0 + ic;
      const bool pv = n < p.Nbatch && gq[ra] < p.P[ra] && gq[2] < p.P[2];
      long long rowoff = -1;
      if (pv && !p.pool)
        rowoff = (long long)n * p.om.sample_stride +
                 (long long)(gq[0] * p.om.omul[0] + p.om.obase[0]) * p.om.os[0] +
                 (long long)(gq[1] * p.om.omul[1] + p.om.obase[1]) * p.om.os[1] +
                 (long long)(gq[2] * p.om.omul[2] + p.om.obase[2]) * p.om.os[2];
      // pooled epilogue: the 2x2 window of this lane's pixel lives in four lanes of the same warp — the column
      // partner is lane ^ 1, the row partner lane ^ rowbit (lanes are (row, sample, column) with 8 columns
      // innermost; tiles start on even rows and columns).
      const int rowbit = p.bnb == 1 ? 8 : 16;
      const int wbase = lane & ~(1 | rowbit);
      const int wpos = ((lane & rowbit) ? 2 : 0) | (lane & 1);          // this lane's position in its window
      long long poff = -1;
      if (p.pool && pv)
        poff = (long long)n * p.p_sample_stride + (long long)(gq[ra] >> 1) * p.pw + (gq[2] >> 1);
      tc::mbar_wait(tfull + buf, (uint32_t)((local / 2) & 1));
      tc::tc_fence_after();
      const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * p.acc_cols);
      const int ch0 = t.y * p.bn;
      for (int col = 0; col < p.bn; col += 16) {
        uint32_t v[16];
        tc::tmem_ld16(lane_base + (uint32_t)col, v);
        if (p.ncat) {
          uint32_t u[16];
          tc::tmem_ld16(lane_base + (uint32_t)(p.bn + col), u);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));

        }
        if (p.pool) {
          // Per channel: the window maximum by two butterfly steps (max.NaN propagates NaNs), then every lane
          // votes "my value is the maximum" (equal, or NaN when the maximum is NaN) and the FIRST voter in
          // row-major window order is the reference's argmax (pooling.py:111: first strict maximum, the first
          // NaN wins).  That lane stores ITS OWN value — bit-exact even for -0.0 / NaN payloads — and its
          // position.  8 lanes of a warp store per channel: two 16-byte runs (4 pooled columns x 2 samples/rows).
          long long o = poff + (long long)(ch0 + col) * p.p_chan_stride;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float x = __uint_as_float(v[j]);
            float m = fmax_nan(x, __shfl_xor_sync(0xffffffffu, x, 1));
            m = fmax_nan(m, __shfl_xor_sync(0xffffffffu, m, rowbit));
            const unsigned votes = __ballot_sync(0xffffffffu, x == m || x != x) >> wbase;
            const unsigned nib = (votes & 3u) | ((votes >> (rowbit - 2)) & 12u);   // window order: bits 0..3
            const int k = __ffs((int)nib) - 1;
            // predicated, not branched: a divergent branch per channel (8 of 32 lanes store) costs far more
            // than the two stores themselves
            st_global_if(p.pout + o, x, p.pidx + o, k, poff >= 0 && k == wpos && ch0 + col + j < p.Nout);
            o += p.p_chan_stride;
          }
        } else if (rowoff >= 0) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            int ch = ch0 + col + j;
            if (ch < p.Nout) p.out[rowoff + (long long)ch * p.om.chan_stride] = __uint_as_float(v[j]);
          }
        }
      }
      tc::tc_fence_before();
      tc::mbar_arrive(tempty + buf);
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

}  // namespace halo
}  // namespace conv
}  // namespace mgb
