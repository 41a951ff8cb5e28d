// Implicit-GEMM convolution core for sm_100a: FP64 on DMMA (mma.sync m8n8k4.f64)
// and FP32 as 3xTF32 on mma.sync m16n8k8.tf32, operands gathered straight from
// the NCX tensors with cp.async (zero-fill = padding) into a multi-stage shared
// memory ring.  No im2col matrix is ever materialised.
//
// All three convolutions of ConvND (src/mygrad/nnet/layers/conv.py) are one GEMM
// D[M,N] = A[M,K] * B[K,N] whose operands are "gathers":
//
//   element(pixel, tap) = T[n, ch(tap), coord(pixel) + delta(tap)]   (0 outside T)
//
//   fwd   (:81-103)  A = gather of x  (pixel = output position, tap = (c,k))   B = w[F][C*W]
//   dgrad (:105-138) A = gather of dy (pixel = input position of one stride-residue
//                    class, tap = (f, k in class))                             B = packed w
//   wgrad (:140-155) A = gather of dy (pixel along K, tap = f)
//                    B = gather of x  (pixel along K, tap = (c,k)),  split-K over pixels.
#pragma once
#include "common.cuh"

namespace mgb {
namespace conv {

struct __align__(16) TapEntry {
  int off;         // ch*chan_stride + sum_i delta_i * xs_i  (elements, per sample)
  int d0, d1, d2;  // coordinate deltas, for the bounds check
};

struct GatherGeom {
  long long sample_stride;  // elements between samples of the gathered tensor
  int P[3];                 // pixel grid
  int pmul[3], base[3];     // source coord = pixel*pmul + base (+ tap delta)
  int X[3];                 // source extents (bounds)
  int xs[3];                // source element strides
  int pp;                   // P0*P1*P2
};

struct OutMap {             // (pixel, channel) -> output element
  long long sample_stride, chan_stride;
  int omul[3], obase[3], os[3];
};

template <typename T>
struct FwdParams {
  const T* src;
  const TapEntry* tab;      // K entries
  const T* B;               // [N][ldb], K contiguous
  T* out;
  long long ldb;
  GatherGeom ga;
  OutMap om;
  int M, N, K;
};

template <typename T>
struct WgradParams {
  const T* srcA;
  const TapEntry* tabA;     // M entries
  const T* srcB;
  const TapEntry* tabB;     // N entries
  T* partial;               // [splits][M][N]
  GatherGeom ga, gb;        // same pixel grid, different coordinate maps
  int M, N, K;
  int k_per_split;          // multiple of BK
};

struct PixCoord {
  long long off;
  int c0, c1, c2;
  bool valid;
};

__device__ __forceinline__ void split_pixel(const GatherGeom& g, int pix, int& n, int& p0, int& p1,
                                            int& p2) {
  n = pix / g.pp;
  int r = pix - n * g.pp;
  int q = r / g.P[2];
  p2 = r - q * g.P[2];
  p0 = q / g.P[1];
  p1 = q - p0 * g.P[1];
}

__device__ __forceinline__ PixCoord pix_coord(const GatherGeom& g, int n, int p0, int p1, int p2,
                                              bool valid) {
  PixCoord pc;
  pc.valid = valid;
  pc.c0 = p0 * g.pmul[0] + g.base[0];
  pc.c1 = p1 * g.pmul[1] + g.base[1];
  pc.c2 = p2 * g.pmul[2] + g.base[2];
  pc.off = (long long)n * g.sample_stride + (long long)pc.c0 * g.xs[0] +
           (long long)pc.c1 * g.xs[1] + (long long)pc.c2 * g.xs[2];
  return pc;
}

__device__ __forceinline__ bool tap_in_bounds(const GatherGeom& g, const PixCoord& pc,
                                              const TapEntry& te) {
  return pc.valid && (unsigned)(pc.c0 + te.d0) < (unsigned)g.X[0] &&
         (unsigned)(pc.c1 + te.d1) < (unsigned)g.X[1] &&
         (unsigned)(pc.c2 + te.d2) < (unsigned)g.X[2];
}

__device__ __forceinline__ TapEntry load_tap(const TapEntry* tab, int idx) {
  int4 v = __ldg(reinterpret_cast<const int4*>(tab) + idx);
  TapEntry te;
  te.off = v.x; te.d0 = v.y; te.d1 = v.z; te.d2 = v.w;
  return te;
}

template <int BYTES>
__device__ __forceinline__ void cp_async_zfill(void* smem, const void* gmem, bool pred) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int sz = pred ? BYTES : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;\n" ::"r"(s), "l"(gmem), "n"(BYTES),
               "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ---- MMA atoms -------------------------------------------------------------

__device__ __forceinline__ void dmma_8x8x4(double (&c)[2], double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c[0]), "+d"(c[1])
      : "d"(a), "d"(b));
}

__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4],
                                                const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// x = hi + lo with hi, lo both exactly representable in TF32 (round-to-nearest)
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(hi) : "f"(x));
  float rem = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(lo) : "f"(rem));
}

// ---- tile configuration ------------------------------------------------------

template <typename T, int BM_, int BN_, int BK_, int WM_, int WN_, int STAGES_>
struct Cfg {
  static constexpr int BM = BM_, BN = BN_, BK = BK_, WM = WM_, WN = WN_, STAGES = STAGES_;
  static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
  static constexpr int THREADS = WARPS_M * WARPS_N * 32;
  static constexpr bool F64 = sizeof(T) == 8;
  // fragment geometry
  static constexpr int FM = F64 ? 8 : 16;   // rows per MMA
  static constexpr int FK = F64 ? 4 : 8;    // k per MMA
  static constexpr int MI = WM / FM, NI = WN / 8;
  static constexpr int ACC = F64 ? 2 : 4;   // accumulator registers per fragment
  // shared-memory leading dimensions, chosen conflict-free for the fragment reads
  //   M-major A  [k][m]: f64 ld % 16 == 4 ; f32 ld % 32 == 8
  //   K-major    [r][k]: f64 ld % 16 == 4 ; f32 ld % 32 == 4
  static constexpr int LDA_M = BM + (F64 ? 4 : 8);
  static constexpr int LDK = BK + 4;
  static constexpr int A_STAGE_M = BK * LDA_M;   // fwd-like A stage (elements)
  static constexpr int A_STAGE_K = BM * LDK;     // wgrad A stage
  static constexpr int B_STAGE = BN * LDK;
  static constexpr size_t SMEM_FWD = (size_t)STAGES * (A_STAGE_M + B_STAGE) * sizeof(T);
  static constexpr size_t SMEM_WGRAD = (size_t)STAGES * (A_STAGE_K + B_STAGE) * sizeof(T);
  static_assert(THREADS % BM == 0, "A loader mapping");
  static_assert(THREADS % BK == 0, "B loader mapping");
  static_assert(BK % FK == 0 && WM % FM == 0 && WN % 8 == 0, "fragment tiling");
  static_assert((F64 ? (LDA_M % 16 == 4) : (LDA_M % 32 == 8)), "LDA_M bank layout");
  static_assert((F64 ? (LDK % 16 == 4) : (LDK % 32 == 4)), "LDK bank layout");
};

// One BK-deep slab of MMAs for a warp.  A_MMAJOR: A stage is [k][m] (fwd-like),
// otherwise [m][k] (wgrad).  B stage is always [n][k].
// `hook(kk)` runs before k-step kk: the caller uses it to issue one slice of the next tile's
// cp.async traffic, so that address arithmetic fills the issue slots between tensor ops.
template <typename T, typename C, bool A_MMAJOR, typename Hook>
__device__ __forceinline__ void mma_slab(const T* __restrict__ As, const T* __restrict__ Bs,
                                         float (&accf)[C::MI][C::NI][4],
                                         double (&accd)[C::MI][C::NI][2], int warp_m, int warp_n,
                                         int lane, Hook hook) {
  const int lr = lane >> 2, lk = lane & 3;
  if constexpr (C::F64) {
#pragma unroll
    for (int kk = 0; kk < C::BK / 4; ++kk) {
      hook(kk);
      double a[C::MI], b[C::NI];
      const int k = kk * 4 + lk;
#pragma unroll
      for (int i = 0; i < C::MI; ++i) {
        int m = warp_m + i * 8 + lr;
        a[i] = A_MMAJOR ? As[k * C::LDA_M + m] : As[m * C::LDK + k];
      }
#pragma unroll
      for (int j = 0; j < C::NI; ++j) b[j] = Bs[(warp_n + j * 8 + lr) * C::LDK + k];
#pragma unroll
      for (int i = 0; i < C::MI; ++i)
#pragma unroll
        for (int j = 0; j < C::NI; ++j) dmma_8x8x4(accd[i][j], a[i], b[j]);
    }
  } else {
#pragma unroll
    for (int kk = 0; kk < C::BK / 8; ++kk) {
      hook(kk);
      uint32_t ahi[C::MI][4], alo[C::MI][4], bhi[C::NI][2], blo[C::NI][2];
      const int k0 = kk * 8 + lk, k1 = k0 + 4;
#pragma unroll
      for (int i = 0; i < C::MI; ++i) {
        int r0 = warp_m + i * 16 + lr, r1 = r0 + 8;
        float v0 = A_MMAJOR ? As[k0 * C::LDA_M + r0] : As[r0 * C::LDK + k0];
        float v1 = A_MMAJOR ? As[k0 * C::LDA_M + r1] : As[r1 * C::LDK + k0];
        float v2 = A_MMAJOR ? As[k1 * C::LDA_M + r0] : As[r0 * C::LDK + k1];
        float v3 = A_MMAJOR ? As[k1 * C::LDA_M + r1] : As[r1 * C::LDK + k1];
        split_tf32(v0, ahi[i][0], alo[i][0]);
        split_tf32(v1, ahi[i][1], alo[i][1]);
        split_tf32(v2, ahi[i][2], alo[i][2]);
        split_tf32(v3, ahi[i][3], alo[i][3]);
      }
#pragma unroll
      for (int j = 0; j < C::NI; ++j) {
        int n = warp_n + j * 8 + lr;
        split_tf32(Bs[n * C::LDK + k0], bhi[j][0], blo[j][0]);
        split_tf32(Bs[n * C::LDK + k1], bhi[j][1], blo[j][1]);
      }
      // small cross terms first, the hi*hi product last.  Three separate sweeps keep the
      // dependent MMAs on one accumulator MI*NI issues apart (mma.sync latency ~30 cycles).
#pragma unroll
      for (int i = 0; i < C::MI; ++i)
#pragma unroll
        for (int j = 0; j < C::NI; ++j) mma_tf32_16x8x8(accf[i][j], alo[i], bhi[j]);
#pragma unroll
      for (int i = 0; i < C::MI; ++i)
#pragma unroll
        for (int j = 0; j < C::NI; ++j) mma_tf32_16x8x8(accf[i][j], ahi[i], blo[j]);
#pragma unroll
      for (int i = 0; i < C::MI; ++i)
#pragma unroll
        for (int j = 0; j < C::NI; ++j) mma_tf32_16x8x8(accf[i][j], ahi[i], bhi[j]);
    }
  }
}

// visit every accumulator element of the warp: f(row_slot, row_in_tile, col_in_tile, value)
// row_slot enumerates the distinct rows a thread owns (compile-time constant per call).
template <typename T, typename C, typename F>
__device__ __forceinline__ void for_each_acc(const float (&accf)[C::MI][C::NI][4],
                                             const double (&accd)[C::MI][C::NI][2], int warp_m,
                                             int warp_n, int lane, F f) {
  const int lr = lane >> 2, lc = (lane & 3) * 2;
#pragma unroll
  for (int i = 0; i < C::MI; ++i)
#pragma unroll
    for (int j = 0; j < C::NI; ++j) {
      if constexpr (C::F64) {
        int r = warp_m + i * 8 + lr, c = warp_n + j * 8 + lc;
        f(i, r, c, (T)accd[i][j][0]);
        f(i, r, c + 1, (T)accd[i][j][1]);
      } else {
        int r = warp_m + i * 16 + lr, c = warp_n + j * 8 + lc;
        f(2 * i, r, c, (T)accf[i][j][0]);
        f(2 * i, r, c + 1, (T)accf[i][j][1]);
        f(2 * i + 1, r + 8, c, (T)accf[i][j][2]);
        f(2 * i + 1, r + 8, c + 1, (T)accf[i][j][3]);
      }
    }
}

// ---- fwd-like kernel: A gathered with pixels along M, B plain [N][K] ---------------

template <typename T, typename C>
__global__ void __launch_bounds__(C::THREADS, 1)
conv_fwdlike_kernel(const __grid_constant__ FwdParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* As = reinterpret_cast<T*>(smem_raw);
  T* Bs = As + C::STAGES * C::A_STAGE_M;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_m = (warp % C::WARPS_M) * C::WM, warp_n = (warp / C::WARPS_M) * C::WN;
  const int m0 = blockIdx.x * C::BM, n0 = blockIdx.y * C::BN;

  // A loader: fixed pixel row, strided tap columns
  constexpr int AKSTEP = C::THREADS / C::BM;
  const int arow = tid % C::BM, akc0 = tid / C::BM;
  PixCoord pc;
  {
    int pix = m0 + arow, n, q0, q1, q2;
    bool v = pix < p.M;
    split_pixel(p.ga, v ? pix : 0, n, q0, q1, q2);
    pc = pix_coord(p.ga, n, q0, q1, q2, v);
  }
  // B loader: fixed k column, strided rows
  constexpr int BNSTEP = C::THREADS / C::BK;
  const int bkc = tid % C::BK, bn0 = tid / C::BK;

  const int KT = (p.K + C::BK - 1) / C::BK;

  constexpr int PARTS = C::BK / C::FK;            // one slice per MMA k-step
  constexpr int A_ELEMS = C::BK / AKSTEP, B_ELEMS = C::BN / BNSTEP;
  // slice `part` of PARTS of the tile's copies (part < 0: everything)
  auto load_tile = [&](int stage, int kt, int part) {
    const int k0 = kt * C::BK;
    T* as = As + stage * C::A_STAGE_M;
    T* bs = Bs + stage * C::B_STAGE;
#pragma unroll
    for (int j = 0; j < A_ELEMS; ++j) {
      if (part >= 0 && (j * PARTS) / A_ELEMS != part) continue;
      int kc = akc0 + j * AKSTEP, k = k0 + kc;
      bool v = k < p.K;
      TapEntry te = load_tap(p.tab, v ? k : 0);
      v = v && tap_in_bounds(p.ga, pc, te);
      const T* g = v ? (p.src + pc.off + te.off) : p.src;
      cp_async_zfill<sizeof(T)>(as + kc * C::LDA_M + arow, g, v);
    }
    const bool kv = (k0 + bkc) < p.K;
#pragma unroll
    for (int j = 0; j < B_ELEMS; ++j) {
      if (part >= 0 && (j * PARTS) / B_ELEMS != part) continue;
      int n = bn0 + j * BNSTEP;
      bool v = kv && (n0 + n) < p.N;
      const T* g = v ? (p.B + (long long)(n0 + n) * p.ldb + k0 + bkc) : p.B;
      cp_async_zfill<sizeof(T)>(bs + n * C::LDK + bkc, g, v);
    }
  };

  float accf[C::MI][C::NI][4];
  double accd[C::MI][C::NI][2];
#pragma unroll
  for (int i = 0; i < C::MI; ++i)
#pragma unroll
    for (int j = 0; j < C::NI; ++j) {
      accf[i][j][0] = accf[i][j][1] = accf[i][j][2] = accf[i][j][3] = 0.f;
      accd[i][j][0] = accd[i][j][1] = 0.0;
    }

#pragma unroll
  for (int s = 0; s < C::STAGES - 1; ++s) {
    if (s < KT) load_tile(s, s, -1);
    cp_async_commit();
  }
  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<C::STAGES - 2>();
    __syncthreads();
    const int nk = kt + C::STAGES - 1;
    const bool more = nk < KT;
    const int stage = kt % C::STAGES;
    mma_slab<T, C, true>(As + stage * C::A_STAGE_M, Bs + stage * C::B_STAGE, accf, accd, warp_m,
                         warp_n, lane, [&](int part) {
                           if (more) load_tile(nk % C::STAGES, nk, part);
                         });
    cp_async_commit();
  }
  cp_async_wait<0>();

  // epilogue: scatter (pixel, channel) to the NCX output.  Rows owned by a thread
  // are fixed, so decode each pixel once.
  constexpr int ROWS = C::F64 ? C::MI : 2 * C::MI;
  long long rowoff[ROWS];
  const int lr = lane >> 2;
#pragma unroll
  for (int i = 0; i < ROWS; ++i) {
    int r = C::F64 ? (warp_m + i * 8 + lr) : (warp_m + (i >> 1) * 16 + (i & 1) * 8 + lr);
    int pix = m0 + r;
    if (pix < p.M) {
      int n, q0, q1, q2;
      split_pixel(p.ga, pix, n, q0, q1, q2);
      rowoff[i] = (long long)n * p.om.sample_stride +
                  (long long)(q0 * p.om.omul[0] + p.om.obase[0]) * p.om.os[0] +
                  (long long)(q1 * p.om.omul[1] + p.om.obase[1]) * p.om.os[1] +
                  (long long)(q2 * p.om.omul[2] + p.om.obase[2]) * p.om.os[2];
    } else {
      rowoff[i] = -1;
    }
  }
  for_each_acc<T, C>(accf, accd, warp_m, warp_n, lane, [&](int i, int, int c, T v) {
    int col = n0 + c;
    if (rowoff[i] >= 0 && col < p.N) p.out[rowoff[i] + (long long)col * p.om.chan_stride] = v;
  });
}

// ---- wgrad kernel: both operands gathered with pixels along K; split-K ---------------

template <typename T, typename C>
__global__ void __launch_bounds__(C::THREADS, 1)
conv_wgrad_kernel(const __grid_constant__ WgradParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* As = reinterpret_cast<T*>(smem_raw);
  T* Bs = As + C::STAGES * C::A_STAGE_K;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_m = (warp % C::WARPS_M) * C::WM, warp_n = (warp / C::WARPS_M) * C::WN;
  const int m0 = blockIdx.x * C::BM, n0 = blockIdx.y * C::BN;
  const int kbeg = blockIdx.z * p.k_per_split;
  const int kend = min(p.K, kbeg + p.k_per_split);
  const int KT = (kend - kbeg + C::BK - 1) / C::BK;

  constexpr int RSTEP = C::THREADS / C::BK;
  const int kc = tid % C::BK, r0 = tid / C::BK;

  constexpr int PARTS = C::BK / C::FK;
  constexpr int A_ROWS = C::BM / RSTEP, B_ROWS = C::BN / RSTEP;
  // slice `part` of PARTS of the tile's copies (part < 0: everything).  The pixel is decoded
  // again in every slice: cheaper than keeping two coordinate sets live across the MMAs.
  auto load_tile = [&](int stage, int kt, int part) {
    int pix = kbeg + kt * C::BK + kc;
    bool pv = pix < kend;
    int n, q0, q1, q2;
    split_pixel(p.ga, pv ? pix : 0, n, q0, q1, q2);
    PixCoord pa = pix_coord(p.ga, n, q0, q1, q2, pv);
    PixCoord pb = pix_coord(p.gb, n, q0, q1, q2, pv);
    T* as = As + stage * C::A_STAGE_K;
    T* bs = Bs + stage * C::B_STAGE;
#pragma unroll
    for (int j = 0; j < A_ROWS; ++j) {
      if (part >= 0 && (j * PARTS) / A_ROWS != part) continue;
      int m = r0 + j * RSTEP;
      bool v = (m0 + m) < p.M;
      TapEntry te = load_tap(p.tabA, v ? (m0 + m) : 0);
      v = v && tap_in_bounds(p.ga, pa, te);
      const T* g = v ? (p.srcA + pa.off + te.off) : p.srcA;
      cp_async_zfill<sizeof(T)>(as + m * C::LDK + kc, g, v);
    }
#pragma unroll
    for (int j = 0; j < B_ROWS; ++j) {
      if (part >= 0 && (j * PARTS) / B_ROWS != part) continue;
      int nn = r0 + j * RSTEP;
      bool v = (n0 + nn) < p.N;
      TapEntry te = load_tap(p.tabB, v ? (n0 + nn) : 0);
      v = v && tap_in_bounds(p.gb, pb, te);
      const T* g = v ? (p.srcB + pb.off + te.off) : p.srcB;
      cp_async_zfill<sizeof(T)>(bs + nn * C::LDK + kc, g, v);
    }
  };

  float accf[C::MI][C::NI][4];
  double accd[C::MI][C::NI][2];
#pragma unroll
  for (int i = 0; i < C::MI; ++i)
#pragma unroll
    for (int j = 0; j < C::NI; ++j) {
      accf[i][j][0] = accf[i][j][1] = accf[i][j][2] = accf[i][j][3] = 0.f;
      accd[i][j][0] = accd[i][j][1] = 0.0;
    }

#pragma unroll
  for (int s = 0; s < C::STAGES - 1; ++s) {
    if (s < KT) load_tile(s, s, -1);
    cp_async_commit();
  }
  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<C::STAGES - 2>();
    __syncthreads();
    const int nk = kt + C::STAGES - 1;
    const bool more = nk < KT;
    const int stage = kt % C::STAGES;
    mma_slab<T, C, false>(As + stage * C::A_STAGE_K, Bs + stage * C::B_STAGE, accf, accd, warp_m,
                          warp_n, lane, [&](int part) {
                            // whole tile up front: re-decoding the pixel per slice costs more
                            // than it hides (measured: f64 wgrad 4.88 ms vs 5.20 ms sliced)
                            if (more && part == 0) load_tile(nk % C::STAGES, nk, -1);
                          });
    cp_async_commit();
  }
  cp_async_wait<0>();

  T* out = p.partial + (long long)blockIdx.z * p.M * p.N;
  for_each_acc<T, C>(accf, accd, warp_m, warp_n, lane, [&](int, int r, int c, T v) {
    int row = m0 + r, col = n0 + c;
    if (row < p.M && col < p.N) out[(long long)row * p.N + col] = v;
  });
}

}  // namespace conv
}  // namespace mgb
