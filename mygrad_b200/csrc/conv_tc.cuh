// FP32 convolution on the 5th-generation tensor cores: tcgen05.mma.kind::tf32 with the
// accumulator in TMEM and a 3xTF32 split (a = a_hi + a_lo, both exactly TF32):
//
//     acc += a_lo*b_hi + a_hi*b_lo + a_hi*b_hi          (fp32 accumulate in TMEM)
//
// Data flow of one conv GEMM  D[pixels, channels_out] = A[pixels, (tap, c)] * B[(tap, c), channels_out]:
//
//   1. a bandwidth-bound prep pass rewrites the gathered tensor (x for fwd, dy for dgrad)
//      from NCX to channels-last [N][X...][Cpad] and splits it into hi / lo TF32 planes;
//      the filter bank is packed to [rows][tap][Cpad] hi / lo the same way (tiny).
//   2. the GEMM kernel is warp-specialised: 4 producer warps gather 128-byte channel
//      vectors (one filter tap, 32 channels) per pixel row with 16-byte cp.async into the
//      canonical K-major SWIZZLE_128B layout (zero-fill = padding), a single thread of
//      warp 4 issues the tcgen05.mma stream, mbarriers pipeline the shared-memory ring,
//      and the 4 producer warps turn into the epilogue: tcgen05.ld the TMEM accumulator
//      and store it to the NCX output (lanes = consecutive pixels => coalesced).
//
// Descriptor encodings follow the PTX ISA "tcgen05 matrix/instruction descriptor" tables.
#pragma once
#include <climits>

#include "conv_gemm.cuh"

namespace mgb {
namespace conv {
namespace tc {

constexpr int BM = 128;              // UMMA M: pixels per CTA tile (TMEM lanes)
constexpr int BK = 32;               // tf32 per smem row = 128 bytes = one SWIZZLE_128B row
constexpr int PRODUCERS = 128;       // 4 warps: producers, then epilogue
constexpr int THREADS = PRODUCERS + 32;
constexpr int A_TILE_BYTES = BM * 128;

struct __align__(16) TcTap {
  int off;             // ((d0*X1 + d1)*X2 + d2) * cpad   (elements of the channels-last source)
  int d0, d1, d2;
};

struct TcGeom {
  long long sample_stride;   // X0*X1*X2*cpad
  int P[3];                  // pixel grid of the GEMM M dimension, per sample
  int pmul[3], base[3];      // source coord = pixel*pmul + base (+ tap delta)
  int X[3];                  // source spatial extents
  int pp;                    // P0*P1*P2
  int cpad;                  // channels per source pixel (multiple of 32)
  int nchunk;                // cpad / 32
  int ntaps;
};

struct TcParams {
  const float* a_hi;
  const float* a_lo;
  const TcTap* taps;
  const float* b_hi;         // [rows >= gridDim.y*bn][K], K = ntaps*cpad, K contiguous
  const float* b_lo;
  float* out;
  OutMap om;
  TcGeom g;
  int M, N;                  // GEMM rows (pixels over the batch) and real output channels
  int bn;                    // N tile: multiple of 16, 16..256
  int tmem_cols;             // power of two >= max(32, bn)
};

// ---- PTX wrappers -----------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// bounded spin: a protocol bug must trap, not hang the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar), done = 0;
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

// true in exactly one lane of a converged warp (elect.sync): unlike `lane == 0`, ptxas then knows
// that a single thread issues the tcgen05 instructions that follow and emits them straight-line
// instead of wrapping each one in an ELECT / broadcast loop over the active lanes
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "@P mov.s32 %0, 1;\n\t}\n"
      : "+r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
  int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz));
}

// arrive on `bar` (without bumping its pending count) once all cp.async issued so far by
// this thread have landed: the copy engine publishes the stage, the thread never blocks
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(slot)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(addr), "r"(cols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], TF32 inputs, FP32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every MMA issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
                   smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
        "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// shared-memory matrix descriptor: K-major operand, SWIZZLE_128B, rows of 128 bytes,
// 8-row groups 1024 bytes apart (SBO); LBO is unused for swizzled K-major (set to 1)
__device__ __forceinline__ uint64_t smem_desc_k_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);          // [0,14)  start address >> 4
  d |= (uint64_t)1 << 16;                          // [16,30) leading byte offset >> 4
  d |= (uint64_t)(1024 >> 4) << 32;                // [32,46) stride byte offset >> 4
  d |= (uint64_t)1 << 46;                          // [46,48) descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                          // [61,64) SWIZZLE_128B
  return d;
}

// instruction descriptor: D=F32, A=B=TF32, both K-major, M=128, N=bn
__host__ __device__ inline uint32_t idesc_tf32(int bn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

inline size_t smem_bytes(int bn, int stages) {
  return (size_t)stages * (2 * A_TILE_BYTES + 2 * (size_t)bn * 128) + 1024 /*align*/ + 256 /*barriers*/;
}

// ---- the GEMM kernel -----------------------------------------------------------------------

template <int STAGES>
__global__ void __launch_bounds__(THREADS, 1) conv_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int b_tile_bytes = p.bn * 128;
  const int stage_bytes = 2 * A_TILE_BYTES + 2 * b_tile_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * p.bn;
  const int S = p.g.ntaps * p.g.nchunk;          // K stages of 32 channels of one tap
  const long long K = (long long)S * BK;

  if (warp == 4) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(full_bar + s, PRODUCERS);
        mbar_init(empty_bar + s, 1);
      }
      mbar_init(accum_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncwarp();
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ===== producers =====
    // Thread t owns 16-byte chunk (t & 7) of rows (t >> 3) + 16*i: one warp instruction then
    // covers 4 whole 128-byte rows, so every L2 request is a full line (a lone 16-byte
    // cp.async still moves a 32-byte sector).
    const int chunk = tid & 7, rg = tid >> 3;
    constexpr int ROWS = BM / 16;   // 8 A rows per thread
    long long pix_off[ROWS];
    int cc0[ROWS], cc1[ROWS], cc2[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
      int pix = m0 + rg + 16 * i;
      bool pv = pix < p.M;
      int pp = pv ? pix : 0;
      int n = pp / p.g.pp;
      int rem = pp - n * p.g.pp;
      int q = rem / p.g.P[2];
      int q2 = rem - q * p.g.P[2];
      int q0 = q / p.g.P[1];
      int q1 = q - q0 * p.g.P[1];
      cc0[i] = pv ? q0 * p.g.pmul[0] + p.g.base[0] : INT_MIN / 2;   // invalid rows fail every bounds check
      cc1[i] = q1 * p.g.pmul[1] + p.g.base[1];
      cc2[i] = q2 * p.g.pmul[2] + p.g.base[2];
      pix_off[i] = (long long)n * p.g.sample_stride +
                   (((long long)(pv ? cc0[i] : 0) * p.g.X[1] + cc1[i]) * p.g.X[2] + cc2[i]) *
                       (long long)p.g.cpad + chunk * 4;
    }
    const uint32_t so = (uint32_t)(chunk ^ (rg & 7)) << 4;   // (rg + 16 i) & 7 == rg & 7
    const uint32_t smem_base = smem_u32(smem);
    const int b_iters = p.bn / 16;

    for (int s = 0; s < S; ++s) {
      const int stage = s % STAGES;
      if (s >= STAGES) mbar_wait(empty_bar + stage, (uint32_t)((s / STAGES - 1) & 1));
      const int tap = s / p.g.nchunk, cc = s - tap * p.g.nchunk;
      const int4 tv = __ldg(reinterpret_cast<const int4*>(p.taps) + tap);
      const long long toff = (long long)tv.x + cc * BK;
      const uint32_t a_hi_s = smem_base + stage * stage_bytes + rg * 128 + so;
      const uint32_t a_lo_s = a_hi_s + A_TILE_BYTES;
#pragma unroll
      for (int i = 0; i < ROWS; ++i) {
        const bool v = (unsigned)(cc0[i] + tv.y) < (unsigned)p.g.X[0] &&
                       (unsigned)(cc1[i] + tv.z) < (unsigned)p.g.X[1] &&
                       (unsigned)(cc2[i] + tv.w) < (unsigned)p.g.X[2];
        const long long aoff = v ? (pix_off[i] + toff) : 0;
        cp_async16(a_hi_s + i * 2048, p.a_hi + aoff, v);
        cp_async16(a_lo_s + i * 2048, p.a_lo + aoff, v);
      }
      const uint32_t b_hi_s = smem_base + stage * stage_bytes + 2 * A_TILE_BYTES + rg * 128 + so;
      const uint32_t b_lo_s = b_hi_s + b_tile_bytes;
      const long long boff0 = (long long)(n0 + rg) * K + (long long)s * BK + chunk * 4;
      for (int i = 0; i < b_iters; ++i) {
        const long long boff = boff0 + (long long)(16 * i) * K;
        cp_async16(b_hi_s + i * 2048, p.b_hi + boff, true);
        cp_async16(b_lo_s + i * 2048, p.b_lo + boff, true);
      }
      // publish asynchronously: the barrier completes when all 128 producers' copies landed.
      // (The consumer issues the generic->async proxy fence after its wait.)
      cp_async_arrive_noinc(full_bar + stage);
    }

    // epilogue row of this thread: TMEM lane == tile row == tid
    const int r = tid;
    const int pix = m0 + r;
    const bool pv = pix < p.M;
    int n, q0, q1, q2;
    {
      int pp = pv ? pix : 0;
      n = pp / p.g.pp;
      int rem = pp - n * p.g.pp;
      int q = rem / p.g.P[2];
      q2 = rem - q * p.g.P[2];
      q0 = q / p.g.P[1];
      q1 = q - q0 * p.g.P[1];
    }

    // ===== epilogue: TMEM -> registers -> NCX output =====
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    long long rowoff = -1;
    if (pv)
      rowoff = (long long)n * p.om.sample_stride +
               (long long)(q0 * p.om.omul[0] + p.om.obase[0]) * p.om.os[0] +
               (long long)(q1 * p.om.omul[1] + p.om.obase[1]) * p.om.os[1] +
               (long long)(q2 * p.om.omul[2] + p.om.obase[2]) * p.om.os[2];
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int col = 0; col < p.bn; col += 16) {
      uint32_t v[16];
      tmem_ld16(lane_base + (uint32_t)col, v);
      if (rowoff >= 0) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          int ch = n0 + col + j;
          if (ch < p.N) p.out[rowoff + (long long)ch * p.om.chan_stride] = __uint_as_float(v[j]);
        }
      }
    }
  } else if (elect_one()) {
    // ===== MMA issuer: one thread =====
    const uint32_t idesc = idesc_tf32(p.bn);
    const uint32_t smem_base = smem_u32(smem);
    for (int s = 0; s < S; ++s) {
      const int stage = s % STAGES;
      mbar_wait(full_bar + stage, (uint32_t)((s / STAGES) & 1));
      fence_proxy_async();     // cp.async (generic proxy) writes -> tensor-core (async proxy) reads
      tc_fence_after();
      const uint32_t a_hi = smem_base + stage * stage_bytes;
      const uint32_t a_lo = a_hi + A_TILE_BYTES;
      const uint32_t b_hi = a_lo + A_TILE_BYTES;
      const uint32_t b_lo = b_hi + b_tile_bytes;
#pragma unroll
      for (int j = 0; j < BK / 8; ++j) {   // 4 k-steps of 8 tf32 (32 bytes) inside the 128-byte row
        const uint64_t dah = smem_desc_k_sw128(a_hi + j * 32), dal = smem_desc_k_sw128(a_lo + j * 32);
        const uint64_t dbh = smem_desc_k_sw128(b_hi + j * 32), dbl = smem_desc_k_sw128(b_lo + j * 32);
        umma_tf32(tmem_base, dal, dbh, idesc, (s | j) ? 1u : 0u);   // small cross terms first
        umma_tf32(tmem_base, dah, dbl, idesc, 1u);
        umma_tf32(tmem_base, dah, dbh, idesc, 1u);
      }
      umma_commit(empty_bar + stage);      // frees the smem stage when these MMAs retire
    }
    umma_commit(accum_bar);                // accumulator complete
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ---- wgrad on the tensor cores --------------------------------------------------------------
//
//   dw[f, c, tap] = sum_pix dy[pix, f] * x[pix (+) tap, c]
//
// GEMM with M = f (128 per CTA), N = (tap, c) blocks of 32 channels, K = output pixels,
// split-K over pixels.  Both operands come from the channels-last planes, so for a fixed
// pixel (= k) the 32 channels of a block are one contiguous 128-byte row: the canonical
// MN-major layout.  For 32-bit (TF32) MN-major operands the only legal shared-memory layout
// is SWIZZLE_128B_BASE32B: atoms of 4 k-rows x 128 B, 32-byte chunks XOR-swizzled with the
// row index (Swizzle<2,5,2> on the byte address).  A stage is [k-atom][block][4 k-rows][128 B].

constexpr int WG_KS = 16;               // pixels per pipeline stage (two UMMA K=8 groups)
constexpr int WG_A_PLANE = WG_KS * 512; // 128 f-channels * 4 B per pixel

struct TcWgradParams {
  const float* a_hi;     // dy planes [N*Gp][fpad]
  const float* a_lo;
  const float* b_hi;     // x planes  [N*Xp][cpad]
  const float* b_lo;
  const TcTap* taps;     // every filter tap: off = ((d0*X1+d1)*X2+d2)*cpad, d = k*dil
  float* partial;        // [splits][F][C][Wp]
  int fpad, cpad, cblocks;
  int F, C, Wp;
  int G[3], S[3], Pd[3], X[3];
  int gp;                // G0*G1*G2
  int Ktot;              // N * gp
  int k_per_split;       // multiple of WG_KS
  int nblocks_total;     // Wp * cblocks
  int nbpc;              // (tap, c-block) blocks per CTA, <= 8
  int tmem_cols;
};

// MN-major SWIZZLE_128B_BASE32B descriptor: LBO = byte distance between 32-element MN
// blocks, SBO = byte distance between atoms of 4 k-rows
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;          // SWIZZLE_128B_BASE32B
  return d;
}

// byte offset of 16-byte chunk c (0..7) inside k-row `row` of a BASE32B atom
__device__ __forceinline__ uint32_t base32b_chunk(uint32_t c, uint32_t row) {
  return ((((c >> 1) ^ (row & 3u)) << 5) | ((c & 1u) << 4));
}

inline size_t wgrad_smem_bytes(int nbpc, int stages) {
  return (size_t)stages * (2 * WG_A_PLANE + 2 * (size_t)WG_KS * nbpc * 128) + 1024 + 256;
}

template <int STAGES>
__global__ void __launch_bounds__(THREADS, 1) wgrad_tc_kernel(const __grid_constant__ TcWgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int bn = p.nbpc * 32;
  const int b_plane = WG_KS * p.nbpc * 128;          // bytes of one B plane per stage
  const int b_atom = p.nbpc * 512;                   // bytes of one k-atom (4 pixels) of B
  const int b_group = 2 * b_atom;                    // bytes of one UMMA K=8 slice of B
  const int stage_bytes = 2 * WG_A_PLANE + 2 * b_plane;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * BM;
  const int blk0 = blockIdx.y * p.nbpc;              // first (tap, c-block) block of this CTA
  const int kbeg = blockIdx.z * p.k_per_split;
  const int kend = min(p.Ktot, kbeg + p.k_per_split);
  const int S = (kend - kbeg + WG_KS - 1) / WG_KS;

  if (warp == 4) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(full_bar + s, PRODUCERS);
        mbar_init(empty_bar + s, 1);
      }
      mbar_init(accum_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncwarp();
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    const uint32_t smem_base = smem_u32(smem);
    // A (dy): chunk32 = 16-byte chunk of the 512-byte f-vector, 4 pixels per pass
    const int a_chunk = tid & 31, a_p0 = tid >> 5;
    const bool a_fvalid = (m0 + a_chunk * 4) < p.fpad;
    const uint32_t a_blk = (uint32_t)(a_chunk >> 3) * 512, a_c = (uint32_t)(a_chunk & 7);
    // B (x): one pixel per thread and stage, chunk c of every block
    const int b_c = tid & 7, b_p = tid >> 3;                 // b_p in [0,16)
    const uint32_t b_dst0 = (uint32_t)(b_p >> 2) * b_atom + (uint32_t)(b_p & 3) * 128 +
                            base32b_chunk((uint32_t)b_c, (uint32_t)b_p);

    for (int s = 0; s < S; ++s) {
      const int stage = s % STAGES;
      if (s >= STAGES) mbar_wait(empty_bar + stage, (uint32_t)((s / STAGES - 1) & 1));
      const int k0 = kbeg + s * WG_KS;
      const uint32_t st_base = smem_base + stage * stage_bytes;
      // ---- A planes
#pragma unroll
      for (int i = 0; i < WG_KS / 4; ++i) {
        const int pl = a_p0 + 4 * i;                          // pixel within the stage
        const int kp = k0 + pl;
        const bool v = a_fvalid && kp < kend;
        const long long off = v ? ((long long)kp * p.fpad + m0 + a_chunk * 4) : 0;
        const uint32_t dst = st_base + (uint32_t)(pl >> 2) * 2048 + a_blk + (uint32_t)(pl & 3) * 128 +
                             base32b_chunk(a_c, (uint32_t)pl);
        cp_async16(dst, p.a_hi + off, v);
        cp_async16(dst + WG_A_PLANE, p.a_lo + off, v);
      }
      // ---- B planes: decode this thread's pixel, then one row per (tap, c-block) block
      {
        const int kp = k0 + b_p;
        const bool pv = kp < kend;
        const int kk = pv ? kp : 0;
        const int n = kk / p.gp;
        int rem = kk - n * p.gp;
        const int q = rem / p.G[2];
        const int g2 = rem - q * p.G[2];
        const int g0 = q / p.G[1];
        const int g1 = q - g0 * p.G[1];
        const int c0 = g0 * p.S[0] - p.Pd[0], c1 = g1 * p.S[1] - p.Pd[1], c2 = g2 * p.S[2] - p.Pd[2];
        const long long xoff =
            ((((long long)n * p.X[0] + c0) * p.X[1] + c1) * p.X[2] + c2) * (long long)p.cpad + b_c * 4;
        const uint32_t dst0 = st_base + 2 * WG_A_PLANE + b_dst0;
        for (int b = 0; b < p.nbpc; ++b) {
          const int gb = blk0 + b;
          const int tap = gb / p.cblocks, cb = gb - tap * p.cblocks;
          bool v = pv && gb < p.nblocks_total;
          int4 tv = make_int4(0, 0, 0, 0);
          if (v) tv = __ldg(reinterpret_cast<const int4*>(p.taps) + tap);
          v = v && (unsigned)(c0 + tv.y) < (unsigned)p.X[0] && (unsigned)(c1 + tv.z) < (unsigned)p.X[1] &&
              (unsigned)(c2 + tv.w) < (unsigned)p.X[2];
          const long long off = v ? (xoff + tv.x + cb * 32) : 0;
          const uint32_t dst = dst0 + (uint32_t)b * 512;
          cp_async16(dst, p.b_hi + off, v);
          cp_async16(dst + b_plane, p.b_lo + off, v);
        }
      }
      cp_async_arrive_noinc(full_bar + stage);
    }

    // ===== epilogue: lane = f, columns = (block, channel-in-block) -> partial[f][c][tap] =====
    float* out = p.partial + (long long)blockIdx.z * p.F * p.C * p.Wp;
    const int f = m0 + tid;
    if (S > 0) {
      mbar_wait(accum_bar, 0);
      tc_fence_after();
    }
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int col = 0; col < bn; col += 16) {
      uint32_t v[16];
      if (S > 0) {
        tmem_ld16(lane_base + (uint32_t)col, v);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0u;
      }
      const int gb = blk0 + (col >> 5);
      const int tap = gb / p.cblocks, cb = gb - tap * p.cblocks;
      if (f < p.F && gb < p.nblocks_total) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          int c = cb * 32 + (col & 31) + j;
          if (c < p.C) out[((long long)f * p.C + c) * p.Wp + tap] = __uint_as_float(v[j]);
        }
      }
    }
  } else if (elect_one()) {
    // ===== MMA issuer =====
    const uint32_t idesc = idesc_tf32(bn) | (1u << 15) | (1u << 16);   // A and B MN-major
    const uint32_t smem_base = smem_u32(smem);
    for (int s = 0; s < S; ++s) {
      const int stage = s % STAGES;
      mbar_wait(full_bar + stage, (uint32_t)((s / STAGES) & 1));
      fence_proxy_async();     // cp.async (generic proxy) writes -> tensor-core (async proxy) reads
      tc_fence_after();
      const uint32_t a_hi = smem_base + stage * stage_bytes;
      const uint32_t a_lo = a_hi + WG_A_PLANE;
      const uint32_t b_hi = a_lo + WG_A_PLANE;
      const uint32_t b_lo = b_hi + b_plane;
#pragma unroll
      for (int j = 0; j < WG_KS / 8; ++j) {
        // LBO = 512 B between 32-channel blocks, SBO = distance between 4-pixel k-atoms
        const uint64_t dah = smem_desc_mn_sw128(a_hi + j * 4096, 512u, 2048u);
        const uint64_t dal = smem_desc_mn_sw128(a_lo + j * 4096, 512u, 2048u);
        const uint64_t dbh = smem_desc_mn_sw128(b_hi + j * b_group, 512u, (uint32_t)b_atom);
        const uint64_t dbl = smem_desc_mn_sw128(b_lo + j * b_group, 512u, (uint32_t)b_atom);
        umma_tf32(tmem_base, dal, dbh, idesc, (s | j) ? 1u : 0u);
        umma_tf32(tmem_base, dah, dbl, idesc, 1u);
        umma_tf32(tmem_base, dah, dbh, idesc, 1u);
      }
      umma_commit(empty_bar + stage);
    }
    if (S > 0) umma_commit(accum_bar);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ---- prep kernels ----------------------------------------------------------------------------

// x [N][C][P] (NCX) -> hi, lo [N][P][cpad] (channels last), TF32 split, zero channel padding
__global__ void __launch_bounds__(256) nchw_to_nhwc_split(const float* __restrict__ x,
                                                          float* __restrict__ hi,
                                                          float* __restrict__ lo, int C, long long P,
                                                          int cpad) {
  __shared__ float tile[32][33];
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32, n = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int c = c0 + ty + 8 * i;
    long long pp = p0 + tx;
    tile[ty + 8 * i][tx] = (c < C && pp < P) ? __ldg(x + ((long long)n * C + c) * P + pp) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long pp = p0 + ty + 8 * i;
    int c = c0 + tx;
    if (pp < P && c < cpad) {
      float v = tile[tx][ty + 8 * i];
      uint32_t h, l;
      split_tf32(v, h, l);
      long long o = ((long long)n * P + pp) * cpad + c;
      hi[o] = __uint_as_float(h);
      lo[o] = __uint_as_float(l);
    }
  }
}

// the same transposition with 16-byte accesses on both sides (needs P % 4 == 0): a CTA moves
// 32 channels x 128 pixels.  Loads are float4 along the pixels, stores float4 along the channels (8
// lanes cover one pixel's 128-byte row of a plane); the shared tile is XOR-swizzled at float4
// granularity — element (ch, px) lives at column ((px >> 2) ^ ((ch >> 2) & 7)) * 4 + (px & 3) — which
// makes both the float4 writes (one channel per warp) and the scalar reads (8 channel quads x 4
// consecutive pixels per warp) bank-conflict free without padding.
__global__ void __launch_bounds__(256) nchw_to_nhwc_split_v4(const float* __restrict__ x,
                                                             float* __restrict__ hi,
                                                             float* __restrict__ lo, int C, long long P,
                                                             int cpad) {
  __shared__ __align__(16) float tile[32][128];
  const long long p0 = (long long)blockIdx.x * 128;
  const int c0 = blockIdx.y * 32, n = blockIdx.z;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = threadIdx.x + 256 * i;
    const int ch = idx >> 5, q = idx & 31;                  // one channel per warp, 32 float4 along the pixels
    const int c = c0 + ch;
    const long long pp = p0 + 4 * q;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < C && pp < P) v = __ldg(reinterpret_cast<const float4*>(x + ((long long)n * C + c) * P + pp));
    *reinterpret_cast<float4*>(&tile[ch][(q ^ ((ch >> 2) & 7)) * 4]) = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = threadIdx.x + 256 * i;
    const int cq = idx & 7, px = idx >> 3;                  // 8 lanes = the 32 channels of one pixel
    const long long pp = p0 + px;
    if (pp < P) {
      const int col = (((px >> 2) ^ cq) << 2) | (px & 3);
      float4 h4, l4;
      float* hp = reinterpret_cast<float*>(&h4);
      float* lp = reinterpret_cast<float*>(&l4);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint32_t h, l;
        split_tf32(tile[4 * cq + k][col], h, l);
        hp[k] = __uint_as_float(h);
        lp[k] = __uint_as_float(l);
      }
      const long long o = ((long long)n * P + pp) * cpad + c0 + 4 * cq;
      *reinterpret_cast<float4*>(hi + o) = h4;
      *reinterpret_cast<float4*>(lo + o) = l4;
    }
  }
}

// generic filter pack: out[row][tap_j][ch] = w[f][c][tap] with (row, ch) = (f, c) for fwd and
// (c, f) for dgrad; rows/chs beyond the real extents are zero.  tap_list maps tap_j -> tap.
struct PackArgs {
  int rows, chs;          // padded extents of the packed matrix (rows, channels per tap)
  int F, C, Wp;           // real filter bank extents
  int ntaps;
  int swap;               // 0: row = f, ch = c (fwd)   1: row = c, ch = f (dgrad)
};

__global__ void __launch_bounds__(256) pack_filters_split(const float* __restrict__ w,
                                                          const int* __restrict__ tap_list,
                                                          float* __restrict__ hi,
                                                          float* __restrict__ lo, PackArgs a) {
  long long total = (long long)a.rows * a.ntaps * a.chs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int ch = (int)(i % a.chs);
    long long t = i / a.chs;
    int tj = (int)(t % a.ntaps);
    int row = (int)(t / a.ntaps);
    int f = a.swap ? ch : row, c = a.swap ? row : ch;
    float v = 0.f;
    if (f < a.F && c < a.C) v = __ldg(w + ((long long)f * a.C + c) * a.Wp + tap_list[tj]);
    uint32_t h, l;
    split_tf32(v, h, l);
    hi[i] = __uint_as_float(h);
    lo[i] = __uint_as_float(l);
  }
}

}  // namespace tc
}  // namespace conv
}  // namespace mgb
