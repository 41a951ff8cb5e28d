// The halo-reuse convolution kernel (conv_halo.cuh) on a CTA PAIR: tcgen05 cta_group::2.
//
// What bounds the single-CTA kernel (measured, profiles/r02_umma_probe_*.log and MGB_HALO_PROF): a
// kind::tf32 UMMA covers K = 8 only, so per k-step the tensor pipe reads the whole 128-row A slab (4 KB)
// plus N x 32 B of B from shared memory; TMA writes the next stages into the same shared memory; and
// every CTA streams the WHOLE packed filter bank from L2 once per M tile (C2 forward: 3.7 GB of the
// 4.3 GB that cross L2 -> SM per launch; the filter stages arrive ~2100 clk after they were requested).
// At N = 64 (dX of a 64-channel input) the two UMMAs of a k-step need 14 KB in 96 tensor clocks: more
// than the 128 B/clk shared memory delivers.
//
// A pair of CTAs on one TPC runs ONE UMMA of M = 256: CTA r owns M tile 2q + r (its own input halo
// and its own 128 TMEM lanes) and HALF of the filter tile, so per SM the B bytes read per UMMA, the B
// bytes TMA writes per stage and the B bytes pulled from L2 all halve, and the same shared memory holds
// twice as many filter stages.
//
// 3xTF32 with the N-concatenated B operand across the pair: each CTA stages [b_hi rows | b_lo rows] of
// ITS half of the channel tile (bn/2 + bn/2 rows).  cta_group::2 takes columns [0, N/2) of B from CTA 0
// and [N/2, N) from CTA 1, so
//   UMMA 1:  a_hi x B (N = 2 bn)  -> D[0, bn/2)   = hi*hi (lower channels)   D[bn/2, bn)    = hi*lo (lower)
//                                    D[bn, 3bn/2) = hi*hi (upper channels)   D[3bn/2, 2bn)  = hi*lo (upper)
//   UMMA 2:  a_lo x B_hi (N = bn) into D + bn/2:  D[bn/2, bn) += lo*hi (lower), D[bn, 3bn/2) += lo*hi (upper)
// and the epilogue adds the two column blocks of each channel half.  Needs bn % 32 == 0.
//
// Barriers: the leader's "full" barriers collect both CTAs' TMA bytes (announced by the leader alone),
// tcgen05.commit multicasts "stage free" / "accumulator ready" to both CTAs, the epilogue threads of
// both CTAs arrive on the leader's "accumulator drained" barrier.
#pragma once
#include "conv_halo.cuh"

namespace mgb {
namespace conv {
namespace halo {

__device__ __forceinline__ void umma2_tf32_w(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                             uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %5, p;\n\t}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
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
// arrive on the barrier at this shared-memory offset in BOTH CTAs once the MMAs issued so far retire
__device__ __forceinline__ void umma2_commit_addr(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(bar),
      "h"((uint16_t)3)
      : "memory");
}

// smem: [A ring: AST x (hi, lo) halo planes][B ring: bst x bn rows x 128 B][barriers]
inline size_t smem_bytes2(int a_plane, int bn, int bst) {
  return (size_t)AST * 2 * a_plane + (size_t)bst * bn * 128 + 1024 + 512;
}

// HaloParams as for conv_halo_kernel with: total_tiles = PAIR tiles (pair q -> channel tile q % ny and
// M tiles 2*(q / ny) + rank), mtiles = number of real M tiles, ncat = 1, acc_cols = 2 bn.
template <bool PROF>
__global__ void __launch_bounds__(THREADS, 1)
conv_halo2_kernel(const __grid_constant__ HaloParams p, const __grid_constant__ CUtensorMap map_a_hi,
                  const __grid_constant__ CUtensorMap map_a_lo, const __grid_constant__ CUtensorMap map_b_hi,
                  const __grid_constant__ CUtensorMap map_b_lo) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int a_stage = 2 * p.a_plane;
  const int half_n = p.bn / 2;
  const int b_half = half_n * 128;               // this CTA's b_hi (or b_lo) rows of one stage
  const int b_stage = 2 * b_half;
  uint8_t* b_ring = smem + AST * a_stage;
  uint64_t* afull = reinterpret_cast<uint64_t*>(b_ring + p.bst * b_stage);
  uint64_t* aempty = afull + AST;
  uint64_t* bfull = aempty + AST;
  uint64_t* bempty = bfull + p.bst;
  uint64_t* tfull = bempty + p.bst;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = tma::cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int ra = p.ra, oa = p.oa;
  const int cnt_o = p.cnt[oa], cnt_r = p.cnt[ra], cnt_c = p.cnt[2];

  if (warp == 0 && lane == 0) {
    tma::prefetch_tmap(&map_a_hi); tma::prefetch_tmap(&map_a_lo);
    tma::prefetch_tmap(&map_b_hi); tma::prefetch_tmap(&map_b_lo);
    // full barriers: ONE arrival — the leader's expect_tx, which announces the bytes of BOTH CTAs; the peer's
    // TMA only completes transactions on it.  (A per-stage remote arrive of the peer costs > 1000 clk: that
    // is what made the first cta_group::2 kernel slower than the single-CTA one.)  The peer cannot run a
    // whole phase ahead: its next use of a slot waits for that slot's multicast "empty" commit.
    // empty / tfull: one multicast commit; tempty: one arrival per epilogue warp of both CTAs.
    for (int s = 0; s < AST; ++s) { tc::mbar_init(afull + s, 1); tc::mbar_init(aempty + s, 1); }
    for (int s = 0; s < p.bst; ++s) { tc::mbar_init(bfull + s, 1); tc::mbar_init(bempty + s, 1); }
    for (int b = 0; b < 2; ++b) { tc::mbar_init(tfull + b, 1); tc::mbar_init(tempty + b, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) tma::tmem2_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc::tc_fence_before();
  __syncthreads();
  tma::cluster_sync_all();                  // both CTAs' barriers exist before any remote signal
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t a_base = tc::smem_u32(smem);
  const uint32_t b_base = tc::smem_u32(b_ring);

  // pair tile -> this CTA's M tile (a phantom tile past the batch when the M-tile count is odd: TMA
  // zero-fills it and the epilogue stores nothing)
  auto my_tile = [&](int q) {
    const int y = q % p.ny, mt = 2 * (q / p.ny) + (int)rank;
    HaloTile t;
    if (mt < p.mtiles) {
      t = decode_tile(p, mt * p.ny + y);
    } else {
      t.y = y; t.c0 = 0; t.r0 = 0; t.xo = 0; t.n0 = p.Nbatch;
    }
    return t;
  };

  if (warp == 6) {
    if (tc::elect_one()) {
      // ===== A producer (both CTAs): own halo; bytes complete on the LEADER's barrier =====
      int ita = 0;
      for (int q = cluster_id; q < p.total_tiles; q += nclusters) {
        const HaloTile t = my_tile(q);
        const int xr = t.r0 + p.base[ra] + p.minr;
        const int xc = t.c0 + p.base[2] + p.minc;
        for (int cc = 0; cc < p.nchunk; ++cc)
          for (int jo = 0; jo < cnt_o; ++jo, ++ita) {
            const int sa = ita % AST;
            if (ita >= AST) tc::mbar_wait(aempty + sa, (uint32_t)((ita / AST - 1) & 1));
            const uint32_t dst = a_base + sa * a_stage;
            const uint32_t bar = tma::map_to_cta(tc::smem_u32(afull + sa), 0);
            if (rank == 0) tma::mbar_expect_tx(afull + sa, (uint32_t)(4 * p.a_tx));
            const int xo = t.xo + p.base[oa] + p.delta[oa][jo];
            tma::tma2_load_5d(dst, &map_a_hi, bar, cc * BK, xc, t.n0, xr, xo);
            tma::tma2_load_5d(dst + p.a_plane, &map_a_lo, bar, cc * BK, xc, t.n0, xr, xo);
          }
      }
    }
  } else if (warp == 0) {
    if (tc::elect_one()) {
      // ===== B producer (both CTAs): own half of the channel tile, b_hi rows then b_lo rows =====
      int itb = 0;
      long long pw = 0, pt0 = 0, pc = 0;
      if (PROF) pt0 = clock64();
      for (int q = cluster_id; q < p.total_tiles; q += nclusters) {
        const int row0 = (q % p.ny) * p.bn + (int)rank * half_n;
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
                const uint32_t bar = tma::map_to_cta(tc::smem_u32(bfull + sb), 0);
                if (rank == 0) tma::mbar_expect_tx(bfull + sb, (uint32_t)(2 * b_stage));
                const int k0 = ((tap0 + jc) * p.nchunk + cc) * BK;
                tma::tma2_load_2d(dst, &map_b_hi, bar, k0, row0);
                tma::tma2_load_2d(dst + b_half, &map_b_lo, bar, k0, row0);
              }
            }
      }
      if (PROF && p.prof) {
        p.prof[blockIdx.x * 8 + 6] = (unsigned long long)pw;
        p.prof[blockIdx.x * 8 + 7] = (unsigned long long)(clock64() - pt0);
      }
    }
  } else if (warp == 1) {
    const uint32_t tmem_base_u = __reduce_or_sync(0xffffffffu, tmem_base);
    if (rank == 0 && tc::elect_one()) {
      // ===== MMA issuer: leader CTA only, M = 256 over the pair =====
      const uint32_t m256 = (256u >> 4) << 24;
      const uint32_t idesc1 = (tc::idesc_tf32(2 * p.bn) & ~(0x1Fu << 24)) | m256;     // a_hi x [hi | lo] halves
      const uint32_t idesc2 = (tc::idesc_tf32(p.bn) & ~(0x1Fu << 24)) | m256;         // a_lo x hi halves
      const uint32_t a_dhi = (((uint32_t)p.hc * 128u) >> 4) | (1u << 14) | (2u << 29);
      const uint32_t b_dhi = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t a_lo0 = (a_base >> 4) | 0x10000u;
      const uint32_t b_lo0 = (b_base >> 4) | 0x10000u;
      const uint32_t a_stage16 = (uint32_t)a_stage >> 4, a_plane16 = (uint32_t)p.a_plane >> 4;
      const uint32_t b_stage16 = (uint32_t)b_stage >> 4;
      const uint32_t bfull0 = tc::smem_u32(bfull), bempty0 = tc::smem_u32(bempty);
      const int ninner = cnt_r * cnt_c;
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
      int local = 0;
      long long wa = 0, wb = 0, wt = 0, wi = 0, wc = 0, t_start = 0, c0 = 0, c1 = 0;
      unsigned long long ns0 = 0;
      if (PROF) { t_start = clock64(); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0)); }
      for (int q = cluster_id; q < p.total_tiles; q += nclusters, ++local) {
        const int buf = local & 1;
        if (PROF) c0 = clock64();
        if (local >= 2) tc::mbar_wait(tempty + buf, (uint32_t)((local / 2 - 1) & 1));
        if (PROF) wt += clock64() - c0;
        tc::tc_fence_after();
        const uint32_t tmem_d = tmem_base_u + (uint32_t)(buf * p.acc_cols);
        const uint32_t tmem_d2 = tmem_d + (uint32_t)half_n;
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
            const uint32_t bh = b_lo0 + sb * b_stage16;
            if (t + 1 < ninner) off16 = p.tapoff[t + 1];
            if (PROF) c0 = clock64();
            if (!(PROF && (p.dbg & 1))) mbar_wait_addr(bfull0 + sb * 8, pb);
            if (PROF) { c1 = clock64(); wb += c1 - c0; }
            tc::tc_fence_after();
            if (first) umma2_tf32_w(tmem_d, ah, a_dhi, bh, b_dhi, idesc1, 0u);
            else       umma2_tf32_acc(tmem_d, ah, a_dhi, bh, b_dhi, idesc1);
            umma2_tf32_acc(tmem_d2, al, a_dhi, bh, b_dhi, idesc2);
#pragma unroll
            for (int j = 1; j < BK / 8; ++j) {
              umma2_tf32_acc(tmem_d, ah + 2 * j, a_dhi, bh + 2 * j, b_dhi, idesc1);
              umma2_tf32_acc(tmem_d2, al + 2 * j, a_dhi, bh + 2 * j, b_dhi, idesc2);
            }
            first = false;
            if (PROF) { c0 = clock64(); wi += c0 - c1; }
            umma2_commit_addr(bempty0 + sb * 8);
            if (PROF) wc += clock64() - c0;
            if (++sb == (uint32_t)p.bst) { sb = 0; pb ^= 1u; }
          }
          tma::umma2_commit(aempty + sa);
          if (++sa == AST) { sa = 0; pa ^= 1u; }
        }
        tma::umma2_commit(tfull + buf);
      }
      if (PROF && p.prof) {
        unsigned long long* o = p.prof + blockIdx.x * 8;
        o[0] = (unsigned long long)wa; o[1] = (unsigned long long)wb; o[2] = (unsigned long long)wt;
        o[3] = (unsigned long long)(clock64() - t_start); o[4] = (unsigned long long)wi; o[5] = (unsigned long long)wc;
        unsigned long long ns1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
        p.prof[4096 + blockIdx.x] = ns1 - ns0;       // wall time of the same span: clocks / ns = SM clock in GHz
      }
    }
  } else {
    // ===== epilogue warps 2..5 (both CTAs): own 128 TMEM lanes; lane = (row, sample, column) =====
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int ic = r & 7;
    const int g = r >> 3;
    const int in = g % p.bnb, ir = g / p.bnb;
    const uint32_t tempty_leader0 = tma::map_to_cta(tc::smem_u32(tempty), 0);
    int local = 0;
    for (int q = cluster_id; q < p.total_tiles; q += nclusters, ++local) {
      const int buf = local & 1;
      const HaloTile t = my_tile(q);
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
      const int rowbit = p.bnb == 1 ? 8 : 16;
      const int wbase = lane & ~(1 | rowbit);
      const int wpos = ((lane & rowbit) ? 2 : 0) | (lane & 1);
      long long poff = -1;
      if (p.pool && pv)
        poff = (long long)n * p.p_sample_stride + (long long)(gq[ra] >> 1) * p.pw + (gq[2] >> 1);
      tc::mbar_wait(tfull + buf, (uint32_t)((local / 2) & 1));
      tc::tc_fence_after();
      const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * p.acc_cols);
      const int ch0 = t.y * p.bn;
      for (int col = 0; col < p.bn; col += 16) {
        // channel col sits in D[col] + D[col + bn/2] (lower half) or D[col + bn/2] + D[col + bn] (upper half)
        const uint32_t c1 = (uint32_t)(col < half_n ? col : col + half_n);
        uint32_t v[16], u[16];
        tc::tmem_ld16(lane_base + c1, v);
        tc::tmem_ld16(lane_base + c1 + (uint32_t)half_n, u);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
        if (p.pool) {
          long long o = poff + (long long)(ch0 + col) * p.p_chan_stride;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float x = __uint_as_float(v[j]);
            float m = fmax_nan(x, __shfl_xor_sync(0xffffffffu, x, 1));
            m = fmax_nan(m, __shfl_xor_sync(0xffffffffu, m, rowbit));
            const unsigned votes = __ballot_sync(0xffffffffu, x == m || x != x) >> wbase;
            const unsigned nib = (votes & 3u) | ((votes >> (rowbit - 2)) & 12u);
            const int k = __ffs((int)nib) - 1;
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
      __syncwarp();
      if (lane == 0) tma::mbar_arrive_cluster(tempty_leader0 + (uint32_t)(buf * 8));   // 8 warps free the buffer
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  tma::cluster_sync_all();                  // nobody frees TMEM or exits while the pair still uses it
  if (warp == 1) tma::tmem2_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

}  // namespace halo
}  // namespace conv
}  // namespace mgb
