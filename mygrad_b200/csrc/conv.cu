// Host-side planning and C-ABI entry points for the implicit-GEMM convolutions.
// Kernels are in conv_gemm.cuh; this file turns an mgb_conv_desc into gather
// tables (built on the device, in the caller's workspace) and launches.
#include <vector>

#include <cstdlib>
#include <cstring>

#include "conv_gemm.cuh"
#include "conv_tc.cuh"
#include "conv_tma.cuh"
#include "conv_halo.cuh"
#include "conv_halo2.cuh"
#include "conv_wgrad2.cuh"
#include "conv_f64_tma.cuh"
#include "conv_thin.cuh"

namespace mgb {
namespace conv {
namespace {

constexpr int kMaxW = 32;  // max filter extent per axis (dgrad residue tables are passed by value)

struct Plan {
  int dtype;
  long long N, C, F;
  int X[3], W[3], S[3], Pd[3], D[3], G[3];
  long long Xp, Wp, Gp;  // products
};

int make_plan(const mgb_conv_desc* d, Plan* out) {
  if (!d) return fail("conv: null descriptor");
  if (d->dtype != MGB_F32 && d->dtype != MGB_F64) return fail("conv: bad dtype %d", d->dtype);
  if (d->nd < 1 || d->nd > MGB_MAX_SPATIAL)
    return fail("conv: %d spatial dims unsupported (1..%d)", d->nd, MGB_MAX_SPATIAL);
  if (d->N < 0 || d->C < 0 || d->F < 0) return fail("conv: negative extent");
  Plan p;
  p.dtype = d->dtype;
  p.N = d->N; p.C = d->C; p.F = d->F;
  int shift = 3 - d->nd;
  for (int i = 0; i < 3; ++i) { p.X[i] = 1; p.W[i] = 1; p.S[i] = 1; p.Pd[i] = 0; p.D[i] = 1; p.G[i] = 1; }
  for (int i = 0; i < d->nd; ++i) {
    long long X = d->X[i], W = d->W[i], S = d->stride[i], P = d->pad[i], D = d->dil[i];
    if (X < 0 || W < 1 || S < 1 || P < 0 || D < 1)
      return fail("conv: bad axis %d (X=%lld W=%lld stride=%lld pad=%lld dil=%lld)", i, X, W, S, P, D);
    if (W > kMaxW) return fail("conv: filter extent %lld > %d unsupported", W, kMaxW);
    long long num = X + 2 * P - ((W - 1) * D + 1);
    if (num < 0 || num % S != 0)
      return fail("conv: stride and kernel dimensions are incompatible on axis %d", i);
    long long G = num / S + 1;
    if (X + 2 * P > INT32_MAX / 4 || G > INT32_MAX / 4) return fail("conv: axis too long");
    p.X[i + shift] = (int)X; p.W[i + shift] = (int)W; p.S[i + shift] = (int)S;
    p.Pd[i + shift] = (int)P; p.D[i + shift] = (int)D; p.G[i + shift] = (int)G;
  }
  p.Xp = (long long)p.X[0] * p.X[1] * p.X[2];
  p.Wp = (long long)p.W[0] * p.W[1] * p.W[2];
  p.Gp = (long long)p.G[0] * p.G[1] * p.G[2];
  const long long lim = INT32_MAX;
  if (p.C * p.Xp > lim || p.F * p.Gp > lim) return fail("conv: per-sample extent exceeds int32");
  if (p.N * p.Gp > lim || p.N * p.Xp > lim) return fail("conv: pixel count exceeds int32");
  if (p.C * p.Wp > lim || p.F * p.Wp > lim) return fail("conv: filter bank too large");
  *out = p;
  return 0;
}

// ---- device-side table builders ------------------------------------------------

struct XTableArgs {  // taps of the x-gather: index (c, t0, t1, t2)
  int C, W[3], D[3], xs[3];
  long long chan_stride;
};

__global__ void build_x_table(TapEntry* tab, XTableArgs a) {
  int wp = a.W[0] * a.W[1] * a.W[2];
  int total = a.C * wp;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
    int c = k / wp, t = k - c * wp;
    int t2 = t % a.W[2]; t /= a.W[2];
    int t1 = t % a.W[1];
    int t0 = t / a.W[1];
    TapEntry e;
    e.d0 = t0 * a.D[0]; e.d1 = t1 * a.D[1]; e.d2 = t2 * a.D[2];
    e.off = (int)(c * a.chan_stride + (long long)e.d0 * a.xs[0] + (long long)e.d1 * a.xs[1] +
                  (long long)e.d2 * a.xs[2]);
    tab[k] = e;
  }
}

__global__ void build_chan_table(TapEntry* tab, int nchan, long long chan_stride) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nchan; k += gridDim.x * blockDim.x) {
    TapEntry e;
    e.off = (int)(k * chan_stride);
    e.d0 = e.d1 = e.d2 = 0;
    tab[k] = e;
  }
}

// one stride-residue class of dgrad: the taps k with (r - k*dil) % stride == 0 per axis
struct DgradClass {
  int cnt[3];
  short kidx[3][kMaxW];
  short e[3][kMaxW];   // (r - k*dil)/stride, exact
  int F, C, W[3], gs[3];
  long long Gp;
};

// table over (f, tap_j) and packed weights wp[c][(f, tap_j)] = w[f][c][tap]
template <typename T>
__global__ void build_dgrad_class(TapEntry* tab, T* wpack, const T* __restrict__ w, DgradClass a) {
  int ntap = a.cnt[0] * a.cnt[1] * a.cnt[2];
  int K = a.F * ntap;
  int wp = a.W[0] * a.W[1] * a.W[2];
  long long total = (long long)K * (a.C + 1);  // slot C of each k builds the table entry
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i / K), k = (int)(i - (long long)c * K);
    int f = k / ntap, j = k - f * ntap;
    int j2 = j % a.cnt[2]; j /= a.cnt[2];
    int j1 = j % a.cnt[1];
    int j0 = j / a.cnt[1];
    if (c == a.C) {
      TapEntry e;
      e.d0 = a.e[0][j0]; e.d1 = a.e[1][j1]; e.d2 = a.e[2][j2];
      e.off = (int)(f * a.Gp + (long long)e.d0 * a.gs[0] + (long long)e.d1 * a.gs[1] +
                    (long long)e.d2 * a.gs[2]);
      tab[k] = e;
    } else {
      int tap = (a.kidx[0][j0] * a.W[1] + a.kidx[1][j1]) * a.W[2] + a.kidx[2][j2];
      wpack[(long long)c * K + k] = w[((long long)f * a.C + c) * wp + tap];
    }
  }
}

template <typename T>
__global__ void splitk_reduce(const T* __restrict__ partial, T* __restrict__ out, long long n,
                              int splits) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int s = 0; s < splits; ++s) acc += (double)partial[(long long)s * n + i];  // fixed order
    out[i] = (T)acc;
  }
}

// set by mgb_conv_wgrad_dp for the duration of the call: the split-K reduction then also sums over
// the data-parallel ranks (comm.cu: one kernel over NVLink peer memory instead of reduce + ncclAllReduce)
thread_local bool g_wgrad_dp = false;

inline int small_grid(long long n);

template <typename T>
int finish_dw(const T* partial, T* dw, long long n, int splits, cudaStream_t st) {
  if (g_wgrad_dp && peer_ready())
    return peer_reduce_allreduce(sizeof(T) == 4 ? MGB_F32 : MGB_F64, partial, splits, n, dw, st);
  splitk_reduce<T><<<small_grid(n), 256, 0, st>>>(partial, dw, n, splits);
  MGB_LAUNCH_CHECK();
  return 0;
}

inline int small_grid(long long n) {
  long long b = (n + 255) / 256;
  long long cap = (long long)sm_count() * 8;
  return (int)std::max(1LL, std::min(b, cap));
}

// ---- tile configurations ---------------------------------------------------------

template <typename T> struct Tiles;
template <> struct Tiles<double> {
  using Big = Cfg<double, 128, 128, 16, 64, 32, 4>;    // 8 warps
  using Mid = Cfg<double, 128, 64, 16, 32, 32, 4>;     // 8 warps
  using Small = Cfg<double, 128, 32, 16, 32, 16, 4>;   // 8 warps
  using WBig = Cfg<double, 128, 128, 16, 64, 32, 4>;
  using WBig96 = Cfg<double, 128, 96, 16, 64, 24, 4>;   // N = 576 = 6 x 96: no padded MMAs at C2
  using WMid = Cfg<double, 64, 128, 16, 32, 32, 4>;
  using WSmall = Cfg<double, 32, 128, 16, 16, 32, 4>;
};
template <> struct Tiles<float> {
  using Big = Cfg<float, 128, 128, 32, 64, 32, 4>;
  using Mid = Cfg<float, 128, 64, 32, 32, 32, 4>;
  using Small = Cfg<float, 128, 32, 32, 32, 16, 4>;
  using WBig = Cfg<float, 128, 128, 32, 64, 32, 4>;
  using WBig96 = Cfg<float, 128, 96, 32, 64, 24, 4>;
  using WMid = Cfg<float, 64, 128, 32, 32, 32, 4>;
  using WSmall = Cfg<float, 32, 128, 32, 16, 32, 4>;
};

template <typename T, typename C>
int launch_fwdlike(const FwdParams<T>& p, cudaStream_t st) {
  static thread_local bool configured = false;
  if (!configured) {
    MGB_CUDA(cudaFuncSetAttribute(conv_fwdlike_kernel<T, C>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_FWD));
    configured = true;
  }
  dim3 grid((p.M + C::BM - 1) / C::BM, (p.N + C::BN - 1) / C::BN);
  conv_fwdlike_kernel<T, C><<<grid, C::THREADS, C::SMEM_FWD, st>>>(p);
  MGB_LAUNCH_CHECK();
  return 0;
}

// thin GEMMs (conv_thin.cuh): few taps x channels (K <= 64: a first layer) or very few output channels (dX of one).
// MGB_CONV_THIN=0 keeps them on the MMA tiles.
inline bool thin_enabled() {
  const char* e = getenv("MGB_CONV_THIN");
  return !(e && e[0] == '0');
}

template <typename T>
bool thin_fwdlike_applies(const FwdParams<T>& p, int* nb, int* kp, size_t* smem) {
  if (!thin_enabled() || !(p.K <= 64 || (p.N <= 8 && p.K <= 1024))) return false;
  *nb = p.N <= 4 ? 4 : p.N <= 16 ? 16 : 32;
  *kp = (p.K + 3) / 4 * 4;
  *smem = (size_t)(*nb) * (*kp) * sizeof(T) + (size_t)p.K * sizeof(TapEntry);
  return *smem <= 48 * 1024;
}

template <typename T, int NB, int PX>
int launch_thin_fwdlike(const FwdParams<T>& p, int kp, size_t smem, cudaStream_t st) {
  const long long groups = ((long long)p.M + thin::THREADS * PX - 1) / (thin::THREADS * PX);
  dim3 grid((unsigned)groups, (unsigned)((p.N + NB - 1) / NB));
  thin::thin_fwdlike_kernel<T, NB, PX><<<grid, thin::THREADS, smem, st>>>(p, kp);
  MGB_LAUNCH_CHECK();
  return 0;
}

// ---- 1-D / 2-D thin layers with the input band staged in shared memory (conv_thin.cuh) ----
struct Thin2dPlan { int R, nbands, RB, CB, CBp; size_t band_bytes; };

bool thin2d_plan(const Plan& p, size_t esz, size_t max_band_bytes, Thin2dPlan* out) {
  if (!thin_enabled() || p.X[0] != 1 || p.W[0] != 1 || p.G[0] != 1 || p.C * p.Wp > 64) return false;
  if (p.N * p.Gp == 0 || p.N > INT32_MAX / 4) return false;
  Thin2dPlan t;
  t.CB = (p.G[2] - 1) * p.S[2] + (p.W[2] - 1) * p.D[2] + 1;
  t.CBp = t.CB | 1;                                    // odd row stride: strided pixel columns spread over the banks
  int R = std::min(p.G[1], std::max(1, 1024 / p.G[2]));
  for (;; R /= 2) {
    t.RB = (R - 1) * p.S[1] + (p.W[1] - 1) * p.D[1] + 1;
    t.band_bytes = (size_t)p.C * t.RB * t.CBp * esz;
    if (t.band_bytes <= max_band_bytes) break;
    if (R == 1) return false;
  }
  t.R = R;
  t.nbands = (p.G[1] + R - 1) / R;
  if ((long long)p.N * t.nbands > INT32_MAX) return false;
  *out = t;
  return true;
}

template <typename T>
thin::Thin2dParams<T> thin2d_params(const Plan& p, const Thin2dPlan& t) {
  thin::Thin2dParams<T> q;
  q.x = nullptr; q.w = nullptr; q.dy = nullptr; q.out = nullptr;
  q.N = (int)p.N; q.C = (int)p.C; q.F = (int)p.F;
  q.X1 = p.X[1]; q.X2 = p.X[2]; q.G1 = p.G[1]; q.G2 = p.G[2]; q.W1 = p.W[1]; q.W2 = p.W[2];
  q.s1 = p.S[1]; q.s2 = p.S[2]; q.p1 = p.Pd[1]; q.p2 = p.Pd[2]; q.d1 = p.D[1]; q.d2 = p.D[2];
  q.R = t.R; q.nbands = t.nbands; q.RB = t.RB; q.CB = t.CB; q.CBp = t.CBp;
  return q;
}

// forward of a thin 1-D / 2-D layer: returns 0 ok, 1 error, 2 not applicable
template <typename T>
int thin2d_fwd(const Plan& p, const T* x, const T* w, T* y, cudaStream_t st) {
  const int K = (int)(p.C * p.Wp), KP = (K + 3) / 4 * 4;
  const int nb = (int)std::min<long long>(p.F, 64);   // filters per block (a run-time loop)
  const size_t fixed = (size_t)nb * KP * sizeof(T) + (size_t)KP * sizeof(int);
  Thin2dPlan t;
  if (fixed > 16 * 1024 || !thin2d_plan(p, sizeof(T), 48 * 1024 - 16 * 1024, &t)) return 2;
  thin::Thin2dParams<T> q = thin2d_params<T>(p, t);
  q.x = x; q.w = w; q.out = y;
  if constexpr (sizeof(T) == 4) {          // float32: mma.sync 3xTF32 on fragments gathered from the band
    const char* em = getenv("MGB_THIN_MMA");
    if (!(em && em[0] == '0')) {
      const int fb = (int)std::min<long long>(p.F, 64);
      const int nt = fb <= 8 ? 1 : fb <= 16 ? 2 : fb <= 32 ? 4 : 8;
      const int kp8 = (K + 7) / 8 * 8, np = nt * 8 + ((nt * 8) % 32 == 0 ? 8 : 0);
      const size_t sm = (size_t)2 * kp8 * np * 4 + (size_t)kp8 * 4 + (((size_t)t.R * p.G[2] + 3) & ~(size_t)3) * 4 + t.band_bytes;
      if (sm <= 48 * 1024) {
        const int gy2 = (int)((p.F + nt * 8 - 1) / (nt * 8));
        const int gx2 = (int)std::min<long long>(p.N * t.nbands, std::max(1, 4 * sm_count() / gy2));
        const dim3 grid(gx2, gy2);
        if (nt == 1) thin::thin2d_fwd_mma_kernel<1, 2><<<grid, 256, sm, st>>>(q, kp8, np);
        else if (nt == 2) thin::thin2d_fwd_mma_kernel<2, 2><<<grid, 256, sm, st>>>(q, kp8, np);
        else if (nt == 4) thin::thin2d_fwd_mma_kernel<4, 2><<<grid, 256, sm, st>>>(q, kp8, np);
        else thin::thin2d_fwd_mma_kernel<8, 2><<<grid, 256, sm, st>>>(q, kp8, np);
        MGB_LAUNCH_CHECK();
        return 0;
      }
    }
  }
  const size_t smem = fixed + t.band_bytes;
  const int gy = (int)((p.F + nb - 1) / nb);
  const int gx = (int)std::min<long long>(p.N * t.nbands, std::max(1, 6 * sm_count() / gy));
  if (KP <= 32) thin::thin2d_fwd_kernel<T, 32, 2><<<dim3(gx, gy), thin::THREADS, smem, st>>>(q, KP, nb);
  else thin::thin2d_fwd_kernel<T, 64, 1><<<dim3(gx, gy), thin::THREADS, smem, st>>>(q, KP, nb);
  MGB_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int run_fwdlike(const FwdParams<T>& p, cudaStream_t st) {
  if (p.M == 0 || p.N == 0) return 0;
  {
    int nb, kp;
    size_t smem;
    if (thin_fwdlike_applies(p, &nb, &kp, &smem)) {
      if (nb == 4 && (long long)p.M < 2LL * sm_count() * thin::THREADS * 4) return launch_thin_fwdlike<T, 4, 1>(p, kp, smem, st);
      if (nb == 4) return launch_thin_fwdlike<T, 4, 4>(p, kp, smem, st);
      if (nb == 16) return launch_thin_fwdlike<T, 16, 2>(p, kp, smem, st);
      return launch_thin_fwdlike<T, 32, 2>(p, kp, smem, st);
    }
  }
  if (p.N > 64) return launch_fwdlike<T, typename Tiles<T>::Big>(p, st);
  if (p.N > 32) return launch_fwdlike<T, typename Tiles<T>::Mid>(p, st);
  return launch_fwdlike<T, typename Tiles<T>::Small>(p, st);
}

template <typename C>
void wgrad_shape(int M, int N, int K, int* tiles, int* splits, int* kps) {
  int t = ((M + C::BM - 1) / C::BM) * ((N + C::BN - 1) / C::BN);
  int kt = (K + C::BK - 1) / C::BK;
  int sms = sm_count();
  // fill the machine once or twice over; keep at least 8 k-slabs per split
  int want = std::max(1, (2 * sms) / t);
  int s = std::max(1, std::min(want, std::max(1, kt / 8)));
  int per = (kt + s - 1) / s;   // slabs per split
  s = (kt + per - 1) / per;
  *tiles = t; *splits = std::max(1, s); *kps = per * C::BK;
}

// which wgrad tile family: by M (= F); family 3 = 128x96 when that pads N less than 128x128
inline int wgrad_family(int M, int N = 0) {
  if (M > 64) {
    long long pad128 = (long long)((N + 127) / 128) * 128, pad96 = (long long)((N + 95) / 96) * 96;
    return (N > 0 && pad96 < pad128) ? 3 : 0;
  }
  return M > 32 ? 1 : 2;
}

template <typename T>
void wgrad_plan(int M, int N, int K, int* splits, int* kps) {
  int tiles;
  switch (wgrad_family(M, N)) {
    case 0: wgrad_shape<typename Tiles<T>::WBig>(M, N, K, &tiles, splits, kps); break;
    case 3: wgrad_shape<typename Tiles<T>::WBig96>(M, N, K, &tiles, splits, kps); break;
    case 1: wgrad_shape<typename Tiles<T>::WMid>(M, N, K, &tiles, splits, kps); break;
    default: wgrad_shape<typename Tiles<T>::WSmall>(M, N, K, &tiles, splits, kps); break;
  }
}

template <typename T, typename C>
int launch_wgrad(const WgradParams<T>& p, int splits, cudaStream_t st) {
  static thread_local bool configured = false;
  if (!configured) {
    MGB_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel<T, C>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_WGRAD));
    configured = true;
  }
  dim3 grid((p.M + C::BM - 1) / C::BM, (p.N + C::BN - 1) / C::BN, splits);
  conv_wgrad_kernel<T, C><<<grid, C::THREADS, C::SMEM_WGRAD, st>>>(p);
  MGB_LAUNCH_CHECK();
  return 0;
}

// ---- geometry helpers ---------------------------------------------------------------

GatherGeom geom_x_by_output(const Plan& p) {  // gather from x, pixels = output grid
  GatherGeom g;
  g.sample_stride = p.C * p.Xp;
  for (int i = 0; i < 3; ++i) { g.P[i] = p.G[i]; g.pmul[i] = p.S[i]; g.base[i] = -p.Pd[i]; g.X[i] = p.X[i]; }
  g.xs[2] = 1; g.xs[1] = p.X[2]; g.xs[0] = p.X[1] * p.X[2];
  g.pp = (int)p.Gp;
  return g;
}

GatherGeom geom_dy_by_output(const Plan& p) {  // gather from dy, pixels = output grid (1x1 taps)
  GatherGeom g;
  g.sample_stride = p.F * p.Gp;
  for (int i = 0; i < 3; ++i) { g.P[i] = p.G[i]; g.pmul[i] = 1; g.base[i] = 0; g.X[i] = p.G[i]; }
  g.xs[2] = 1; g.xs[1] = p.G[2]; g.xs[0] = p.G[1] * p.G[2];
  g.pp = (int)p.Gp;
  return g;
}

XTableArgs x_table_args(const Plan& p) {
  XTableArgs a;
  a.C = (int)p.C;
  for (int i = 0; i < 3; ++i) { a.W[i] = p.W[i]; a.D[i] = p.D[i]; }
  a.xs[2] = 1; a.xs[1] = p.X[2]; a.xs[0] = p.X[1] * p.X[2];
  a.chan_stride = p.Xp;
  return a;
}

inline int round_up(long long v, int m) { return (int)((v + m - 1) / m * m); }

inline long long ceil_div_ll(long long a, long long b) {  // b > 0
  return a >= 0 ? (a + b - 1) / b : -((-a) / b);
}

struct DgradClassPlan {
  DgradClass dev;
  int Q[3], qlo[3], r[3];
  int ntap;
  long long K;        // F * ntap
  size_t tab_off, wp_off;
};

// enumerate the stride-residue classes that own at least one dx element
void dgrad_classes(const Plan& p, size_t esz, std::vector<DgradClassPlan>* out, size_t* ws_bytes) {
  size_t off = 0;
  for (int r0 = 0; r0 < p.S[0]; ++r0)
    for (int r1 = 0; r1 < p.S[1]; ++r1)
      for (int r2 = 0; r2 < p.S[2]; ++r2) {
        int r[3] = {r0, r1, r2};
        DgradClassPlan c;
        bool empty = false;
        c.ntap = 1;
        for (int i = 0; i < 3; ++i) {
          // padded coordinate s*q + r must land in [pad, X + pad)
          long long qlo = ceil_div_ll(p.Pd[i] - r[i], p.S[i]);
          long long qhi = ceil_div_ll((long long)p.X[i] + p.Pd[i] - r[i], p.S[i]);
          c.qlo[i] = (int)qlo; c.Q[i] = (int)(qhi - qlo); c.r[i] = r[i];
          if (c.Q[i] <= 0) empty = true;
          int cnt = 0;
          for (int k = 0; k < p.W[i]; ++k) {
            int num = r[i] - k * p.D[i];
            if (((num % p.S[i]) + p.S[i]) % p.S[i] == 0) {
              c.dev.kidx[i][cnt] = (short)k;
              c.dev.e[i][cnt] = (short)(num / p.S[i]);  // exact division
              ++cnt;
            }
          }
          c.dev.cnt[i] = cnt;
          c.ntap *= cnt;
        }
        if (empty || c.ntap == 0) continue;  // nothing to write / stays zero from the memset
        c.K = p.F * c.ntap;
        c.dev.F = (int)p.F; c.dev.C = (int)p.C;
        for (int i = 0; i < 3; ++i) c.dev.W[i] = p.W[i];
        c.dev.gs[2] = 1; c.dev.gs[1] = p.G[2]; c.dev.gs[0] = p.G[1] * p.G[2];
        c.dev.Gp = p.Gp;
        c.tab_off = off; off += align_up((size_t)c.K * sizeof(TapEntry), 1024);
        // packed filters: [c][(f, tap)] for the cp.async kernels, [c][tap][f padded to 16] for the float64 TMA kernel
        c.wp_off = off;  off += align_up((size_t)round_up(p.F, 16) * c.ntap * p.C * esz, 1024) + 1024;
        out->push_back(c);
      }
  *ws_bytes = off;
}

bool dgrad_needs_memset(const Plan& p, const std::vector<DgradClassPlan>& cls) {
  long long covered = 0;
  for (auto& c : cls) covered += (long long)c.Q[0] * c.Q[1] * c.Q[2];
  return covered != p.Xp;
}

// ---- FP32 on tcgen05 (conv_tc.cuh): planning ---------------------------------------------

// MGB_CONV_F32_PATH=mma forces the mma.sync 3xTF32 kernels (A/B comparisons, fallback tests)
bool tc_enabled() {
  const char* e = getenv("MGB_CONV_F32_PATH");   // read per call: tests flip it at run time
  return !(e && strcmp(e, "mma") == 0);
}


// the tensor-core path wants a GEMM worth its fixed costs: >= 16 input and output channels
bool tc_eligible(const Plan& p, int which) {
  if (p.dtype != MGB_F32 || !tc_enabled()) return false;
  long long cin = which == 0 ? p.C : p.F, cout = which == 0 ? p.F : p.C;
  if (cin < 16 || cout < 16) return false;
  if (which == 2) {
    const char* e = getenv("MGB_WGRAD_F32_PATH");   // "mma" keeps wgrad on the mma.sync kernel
    if (e && strcmp(e, "mma") == 0) return false;
    if ((long long)round_up(p.F, 32) * p.Gp > INT32_MAX || p.N * p.Gp > INT32_MAX - 64) return false;
  }
  if ((long long)round_up(cin, 32) * std::max(p.Xp, p.Gp) > INT32_MAX) return false;
  return true;
}

struct TapBuild {          // taps of one GEMM: per-axis lists of (filter index, coordinate delta)
  int cnt[3];
  short kidx[3][kMaxW];
  short delta[3][kMaxW];
  int W[3];                // filter extents (to flatten kidx into the filter-bank tap index)
  int X[3];                // source spatial extents
  int cpad;
};

__global__ void build_tc_taps(tc::TcTap* taps, int* tap_list, TapBuild a) {
  int ntap = a.cnt[0] * a.cnt[1] * a.cnt[2];
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < ntap; j += gridDim.x * blockDim.x) {
    int t = j;
    int j2 = t % a.cnt[2]; t /= a.cnt[2];
    int j1 = t % a.cnt[1];
    int j0 = t / a.cnt[1];
    tc::TcTap e;
    e.d0 = a.delta[0][j0]; e.d1 = a.delta[1][j1]; e.d2 = a.delta[2][j2];
    e.off = (int)((((long long)e.d0 * a.X[1] + e.d1) * a.X[2] + e.d2) * a.cpad);
    taps[j] = e;
    tap_list[j] = (a.kidx[0][j0] * a.W[1] + a.kidx[1][j1]) * a.W[2] + a.kidx[2][j2];
  }
}

struct TcLayout {          // workspace carve-up of one tensor-core GEMM
  int cpad, bn, ny, rows, ntaps;
  long long K;
  size_t taps_off, list_off, bhi_off, blo_off, end;
};

TcLayout tc_layout(size_t start, long long cin, long long cout, int ntaps) {
  TcLayout L;
  L.cpad = round_up(cin, 32);
  L.bn = std::min(128, round_up(cout, 16));
  L.ny = (int)((cout + L.bn - 1) / L.bn);
  L.rows = L.ny * L.bn;
  L.ntaps = ntaps;
  L.K = (long long)ntaps * L.cpad;
  size_t off = align_up(start, 256);
  L.taps_off = off; off += align_up((size_t)ntaps * sizeof(tc::TcTap), 256);
  L.list_off = off; off += align_up((size_t)ntaps * sizeof(int), 256);
  L.bhi_off = off;  off += align_up((size_t)L.rows * L.K * sizeof(float), 256);
  L.blo_off = off;  off += align_up((size_t)L.rows * L.K * sizeof(float), 256);
  L.end = off;
  return L;
}

// channels-last hi/lo planes of the gathered tensor
struct TcPlanes { size_t hi_off, lo_off, end; int cpad; };
TcPlanes tc_planes(size_t start, long long n, long long chans, long long spatial) {
  TcPlanes P;
  P.cpad = round_up(chans, 32);
  size_t bytes = align_up((size_t)n * spatial * P.cpad * sizeof(float), 1024);
  P.hi_off = align_up(start, 1024);
  P.lo_off = P.hi_off + bytes;
  P.end = P.lo_off + bytes;
  return P;
}

int tc_prep_planes(const float* src, float* hi, float* lo, long long n, long long chans,
                   long long spatial, int cpad, cudaStream_t st) {
  if (n * spatial == 0) return 0;
  if (n > 65535) return fail("conv (tcgen05 path): batch %lld > 65535 unsupported", n);
  const bool wide = spatial % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
                    !(getenv("MGB_PREP_V4") && getenv("MGB_PREP_V4")[0] == '0');
  if (wide) {    // 16-byte loads and stores, 32 channels x 128 pixels per CTA
    dim3 grid((unsigned)((spatial + 127) / 128), (unsigned)(cpad / 32), (unsigned)n);
    tc::nchw_to_nhwc_split_v4<<<grid, 256, 0, st>>>(src, hi, lo, (int)chans, spatial, cpad);
  } else {
    dim3 grid((unsigned)((spatial + 31) / 32), (unsigned)(cpad / 32), (unsigned)n);
    tc::nchw_to_nhwc_split<<<grid, 256, 0, st>>>(src, hi, lo, (int)chans, spatial, cpad);
  }
  MGB_LAUNCH_CHECK();
  return 0;
}

int tc_launch(const tc::TcParams& tp, int ny, cudaStream_t st) {
  constexpr int STAGES = 3;
  size_t smem = tc::smem_bytes(tp.bn, STAGES);
  MGB_CUDA(cudaFuncSetAttribute(tc::conv_tc_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  dim3 grid((tp.M + tc::BM - 1) / tc::BM, ny);
  tc::conv_tc_kernel<STAGES><<<grid, tc::THREADS, smem, st>>>(tp);
  MGB_LAUNCH_CHECK();
  return 0;
}

int pow2_cols(int bn) { int c = 32; while (c < bn) c <<= 1; return c; }

// ---- TMA variant of the fwd-like tensor-core GEMM (conv_tma.cuh) -------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn tma_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

// MGB_CONV_F32_PATH: "tma" (default) | "cpasync" (cp.async-fed tcgen05 kernel) | "mma"
bool tma_enabled() {
  const char* e = getenv("MGB_CONV_F32_PATH");
  return !(e && (strcmp(e, "cpasync") == 0 || strcmp(e, "mma") == 0));
}

inline int next_pow2(long long v) { int r = 1; while (r < v) r <<= 1; return r; }

// pixel box (powers of two, 128 pixels) with the fewest tiles; false when no box fits TMA's limits
bool choose_box(const int P[3], const int pmul[3], long long nbatch, int box[3], int* boxn,
                int pixels = 128, bool merged_outer = false) {
  // merged_outer: the batch and axis 0 are one TMA dimension: boxes must tile axis 0 exactly and
  // may not span samples
  long long best = -1;
  for (int b0 = 1; b0 <= pixels; b0 <<= 1)
    for (int b1 = 1; b0 * b1 <= pixels; b1 <<= 1)
      for (int b2 = 1; b0 * b1 * b2 <= pixels; b2 <<= 1) {
        int bn = pixels / (b0 * b1 * b2);
        if (b0 > next_pow2(P[0]) || b1 > next_pow2(P[1]) || b2 > next_pow2(P[2]) || bn > next_pow2(nbatch)) continue;
        if (b0 * pmul[0] > 256 || b1 * pmul[1] > 256 || b2 * pmul[2] > 256 || bn > 256) continue;
        if (merged_outer && (bn != 1 || P[0] % b0 != 0)) continue;
        long long tiles = ((P[0] + b0 - 1) / b0) * (long long)((P[1] + b1 - 1) / b1) * ((P[2] + b2 - 1) / b2) *
                          ((nbatch + bn - 1) / bn);
        // fewer tiles first; then wider contiguous runs for the epilogue stores
        long long cost = tiles * 1024 - b2;
        if (best < 0 || cost < best) { best = cost; box[0] = b0; box[1] = b1; box[2] = b2; *boxn = bn; }
      }
  return best >= 0;
}

struct TmaJob {                   // one fwd-like GEMM on channels-last planes
  const float *a_hi, *a_lo;       // planes [N][X0][X1][X2][cpad]
  const float *b_hi, *b_lo;       // packed filters [rows][K]
  const tc::TcTap* taps;
  float* out;
  OutMap om;
  int P[3], pmul[3], base[3], X[3];
  long long nbatch;
  int cpad, ntaps, nout, bn, ny, rows;
  const TapBuild* tb = nullptr;   // per-axis tap lists (host copy) — the halo kernel addresses taps by axis
  // fused 2x2/2 max-pool epilogue (halo kernel only): pooled values + argmax positions instead of `out`
  int pool = 0;
  float* pout = nullptr;
  unsigned char* pidx = nullptr;
};

// ---- halo-reuse variant (conv_halo.cuh): returns 0 ok, 1 error, 2 not applicable -----------
// MGB_CONV_HALO: unset = by cost model, "0" = never, "1" = whenever the geometry allows it
int halo_launch(const TmaJob& j, cudaStream_t st, bool dry_run = false) {
  static_assert(halo::kMaxTap >= kMaxW, "tap table too small");
  EncodeTiledFn enc = tma_encoder();
  const char* eh = getenv("MGB_CONV_HALO");
  if ((!enc && !dry_run) || !tma_enabled() || !j.tb || (eh && eh[0] == '0' && !j.pool)) return 2;
  const bool force = (eh && eh[0] == '1') || j.pool;
  // the pooled epilogue pairs rows inside one warp: 2-D grids with even extents, at most 2 samples per tile
  if (j.pool && (j.P[0] != 1 || j.P[1] % 2 || j.P[2] % 2)) return 2;
  const TapBuild& tb = *j.tb;
  for (int i = 0; i < 3; ++i) if (j.pmul[i] != 1 || j.X[i] < 1 || j.P[i] < 1 || tb.cnt[i] < 1) return 2;
  if (tb.cnt[0] * tb.cnt[2] > halo::kMaxInner || tb.cnt[1] * tb.cnt[2] > halo::kMaxInner) return 2;
  if (j.nbatch < 1) return 2;
  const bool ncat = j.bn % 8 == 0 && 2 * j.bn <= 256;
  // CTA pairs (tcgen05 cta_group::2, conv_halo2.cuh): each SM stages, reads and fetches from L2 only half of
  // the filter tile.  MGB_CONV_2CTA=0 keeps the single-CTA kernel, =1 is the default where it applies.
  const char* e2 = getenv("MGB_CONV_2CTA");
  const bool pair = !(e2 && e2[0] == '0') && ncat && j.bn % 32 == 0;
  const int b_stage = pair ? j.bn * 128 : 2 * j.bn * 128;
  const long long smem_cap = 227 * 1024;
  int mind[3], maxd[3];
  for (int i = 0; i < 3; ++i) {
    mind[i] = maxd[i] = tb.delta[i][0];
    for (int k = 1; k < tb.cnt[i]; ++k) { mind[i] = std::min(mind[i], (int)tb.delta[i][k]); maxd[i] = std::max(maxd[i], (int)tb.delta[i][k]); }
  }
  const int hc = 8 + maxd[2] - mind[2];
  if (hc > 256) return 2;
  // pick the row axis and the (rows, samples) split of the 16 groups with the fewest tiles
  long long best = -1;
  int b_ra = 1, b_bh = 16, b_bn = 1, b_plane = 0, b_bst = 0;
  const int tiles_c = (j.P[2] + 7) / 8;
  for (int ra = 1; ra >= 0; --ra) {
    const int oa = 1 - ra;
    if (ra == 0 && j.P[0] == 1) continue;                     // <= 2 spatial dims: rows are axis 1
    for (int bh = 16; bh >= 1; bh >>= 1) {
      const int bnb = 16 / bh;
      if (bh > next_pow2(j.P[ra]) || bnb > next_pow2(j.nbatch)) continue;
      if (j.pool && bnb > 2) continue;
      const int hr = bh + maxd[ra] - mind[ra];
      if (hr > 256) continue;
      const long long tx = (long long)hr * bnb * hc * 128;
      const long long plane = (tx + 1023) / 1024 * 1024;
      const long long left = smem_cap - (1024 + 512) - halo::AST * 2 * plane;
      // filter stages arrive ~2000 clk after they are requested (L2 under load): the ring must cover that
      const int bst = (int)std::min<long long>(pair ? 16 : 6, left / b_stage);
      if (bst < 2) continue;
      const long long tiles = (long long)((j.P[ra] + bh - 1) / bh) * ((j.nbatch + bnb - 1) / bnb) * j.P[oa] * tiles_c;
      // shared-memory bytes moved per tile: operand reads + filter writes per tap, halo writes per outer tap
      const long long ntap = (long long)tb.cnt[0] * tb.cnt[1] * tb.cnt[2];
      const long long reads = ncat ? 32768 + 384LL * j.bn : 3 * (16384 + 128LL * j.bn);
      const long long cost = tiles * (ntap * (reads + b_stage) + (long long)tb.cnt[oa] * 2 * tx);
      if (best < 0 || cost < best) { best = cost; b_ra = ra; b_bh = bh; b_bn = bnb; b_plane = (int)plane; b_bst = bst; }
    }
  }
  if (best < 0) return 2;
  if (dry_run) return 0;
  if (!force) {
    // the plain TMA kernel re-loads a 2 x 16 KB A tile per tap but tiles with 128-pixel boxes of any shape
    int box[3], boxn;
    if (choose_box(j.P, j.pmul, j.nbatch, box, &boxn)) {
      long long tiles = (j.nbatch + boxn - 1) / boxn;
      for (int i = 0; i < 3; ++i) tiles *= (j.P[i] + box[i] - 1) / box[i];
      const long long ntap = (long long)tb.cnt[0] * tb.cnt[1] * tb.cnt[2];
      const long long reads = ncat ? 32768 + 384LL * j.bn : 3 * (16384 + 128LL * j.bn);
      const long long plain = tiles * ntap * (reads + b_stage + 32768);
      if (best * 100 > plain * 95) return 2;
    }
  }
  halo::HaloParams hp;
  const int ra = b_ra, oa = 1 - b_ra;
  hp.out = j.out; hp.om = j.om;
  for (int i = 0; i < 3; ++i) {
    hp.P[i] = j.P[i]; hp.base[i] = j.base[i]; hp.cnt[i] = tb.cnt[i];
    for (int k = 0; k < tb.cnt[i]; ++k) hp.delta[i][k] = tb.delta[i][k];
  }
  hp.ra = ra; hp.oa = oa; hp.bh = b_bh; hp.bnb = b_bn;
  hp.tiles_r = (j.P[ra] + b_bh - 1) / b_bh; hp.tiles_c = tiles_c;
  hp.tiles_n = (int)((j.nbatch + b_bn - 1) / b_bn);
  hp.Nbatch = (int)j.nbatch; hp.Nout = j.nout; hp.nchunk = j.cpad / 32;
  hp.minr = mind[ra]; hp.minc = mind[2];
  hp.hr = b_bh + maxd[ra] - mind[ra]; hp.hc = hc;
  hp.a_plane = b_plane; hp.a_tx = hp.hr * b_bn * hc * 128;
  {
    const int row_pitch = b_bn * hc * 128;      // bytes between halo rows
    for (int jr = 0; jr < tb.cnt[ra]; ++jr)
      for (int jc = 0; jc < tb.cnt[2]; ++jc)
        hp.tapoff[jr * tb.cnt[2] + jc] =
            (uint32_t)((tb.delta[ra][jr] - mind[ra]) * row_pitch + (tb.delta[2][jc] - mind[2]) * 128) >> 4;
  }
  hp.bn = j.bn; hp.ny = j.ny; hp.bst = b_bst;
  hp.ncat = ncat ? 1 : 0; hp.acc_cols = ncat ? 2 * j.bn : j.bn; hp.tmem_cols = pow2_cols(2 * hp.acc_cols);
  const long long mtiles = (long long)hp.tiles_r * hp.tiles_c * hp.tiles_n * j.P[oa];
  const long long total = pair ? (mtiles + 1) / 2 * j.ny : mtiles * j.ny;
  if (mtiles * j.ny > INT32_MAX) return 2;
  hp.total_tiles = (int)total;
  hp.mtiles = (int)mtiles;
  hp.pool = j.pool; hp.pout = j.pout; hp.pidx = j.pidx;
  hp.pw = j.P[2] / 2;
  hp.p_chan_stride = (long long)(j.P[1] / 2) * (j.P[2] / 2);
  hp.p_sample_stride = hp.p_chan_stride * j.nout;
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  {
    // (Cpad, X_col, N, X_row, X_outer): the halo lands as [row][sample][column][128 B]
    const cuuint64_t sx[3] = {(cuuint64_t)j.X[1] * j.X[2] * j.cpad * 4, (cuuint64_t)j.X[2] * j.cpad * 4,
                              (cuuint64_t)j.cpad * 4};
    cuuint64_t dims[5] = {(cuuint64_t)j.cpad, (cuuint64_t)j.X[2], (cuuint64_t)j.nbatch, (cuuint64_t)j.X[ra],
                          (cuuint64_t)j.X[oa]};
    cuuint64_t strides[4] = {sx[2], sx[0] * j.X[0], sx[ra], sx[oa]};
    cuuint32_t box[5] = {32u, (cuuint32_t)hc, (cuuint32_t)b_bn, (cuuint32_t)hp.hr, 1u};
    cuuint32_t estr[5] = {1u, 1u, 1u, 1u, 1u};
    CUresult r1 = enc(&ma_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)j.a_hi, dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&ma_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)j.a_lo, dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) return 2;
  }
  {
    long long K = (long long)j.ntaps * j.cpad;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)j.rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 4};
    cuuint32_t box[2] = {32u, (cuuint32_t)(pair ? j.bn / 2 : j.bn)};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r1 = enc(&mb_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)j.b_hi, dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&mb_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)j.b_lo, dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) return 2;
  }
  hp.prof = nullptr;
  hp.dbg = getenv("MGB_HALO2_DBG") ? atoi(getenv("MGB_HALO2_DBG")) : 0;
  if (pair) {
    size_t smem2 = halo::smem_bytes2(hp.a_plane, j.bn, hp.bst);
    const bool prof = getenv("MGB_HALO_PROF") != nullptr;
    static unsigned long long* prof_buf2 = nullptr;
    if (prof) {
      if (!prof_buf2) MGB_CUDA(cudaMalloc(&prof_buf2, 8 * 1024 * sizeof(unsigned long long)));
      MGB_CUDA(cudaMemsetAsync(prof_buf2, 0, 8 * 1024 * sizeof(unsigned long long), st));
      hp.prof = prof_buf2;
    }
    auto kern = prof ? halo::conv_halo2_kernel<true> : halo::conv_halo2_kernel<false>;
    MGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)std::min<long long>(2 * total, sm_count() / 2 * 2));
    cfg.blockDim = dim3(halo::THREADS);
    cfg.dynamicSmemBytes = smem2;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    MGB_CUDA(cudaLaunchKernelEx(&cfg, kern, hp, ma_hi, ma_lo, mb_hi, mb_lo));
    count_launch();
    if (prof) {
      const int grid = (int)cfg.gridDim.x;
      std::vector<unsigned long long> h(8 * 1024);
      MGB_CUDA(cudaStreamSynchronize(st));
      MGB_CUDA(cudaMemcpy(h.data(), prof_buf2, h.size() * 8, cudaMemcpyDeviceToHost));
      const double stages = (double)total / (grid / 2) * hp.nchunk * tb.cnt[0] * tb.cnt[1] * tb.cnt[2];
      double s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, peer[2] = {0, 0}, ns = 0;
      for (int i = 0; i < grid; i += 2) ns += (double)h[4096 + i] / (grid / 2);
      for (int i = 0; i < grid; i += 2) {
        for (int k = 0; k < 8; ++k) s[k] += (double)h[8 * i + k] / (grid / 2);
        peer[0] += (double)h[8 * (i + 1) + 6] / (grid / 2); peer[1] += (double)h[8 * (i + 1) + 7] / (grid / 2);
      }
      fprintf(stderr, "[halo2 prof] bn=%d pair tiles=%d bst=%d hr=%d hc=%d bnb=%d, per tap-stage: wait A %.0f, wait B %.0f, "
              "wait tmem %.0f, umma issue %.0f, commit %.0f, total %.0f clk; B producer wait empty: leader %.0f, peer %.0f of %.0f\n",
              j.bn, hp.total_tiles, hp.bst, hp.hr, hp.hc, hp.bnb, s[0] / stages, s[1] / stages, s[2] / stages, s[4] / stages,
              s[5] / stages, s[3] / stages, s[6] / stages, peer[0] / stages, s[7] / stages);
      fprintf(stderr, "[halo2 prof] MMA loop %.0f clk in %.1f us: SM clock %.0f MHz\n", s[3], ns / 1e3, s[3] / (ns / 1e3));
    }
    return 0;
  }
  size_t smem = halo::smem_bytes(hp.a_plane, j.bn, hp.bst);
  int grid = (int)std::min<long long>(total, sm_count());
  if (getenv("MGB_HALO_PROF")) {
    // diagnostic build of the same kernel: where does the MMA thread spend its cycles?  (synchronous)
    static unsigned long long* prof_buf = nullptr;
    if (!prof_buf) MGB_CUDA(cudaMalloc(&prof_buf, 10 * 1024 * sizeof(unsigned long long)));
    hp.prof = prof_buf;
    MGB_CUDA(cudaFuncSetAttribute(halo::conv_halo_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    halo::conv_halo_kernel<true><<<grid, halo::THREADS, smem, st>>>(hp, ma_hi, ma_lo, mb_hi, mb_lo);
    MGB_LAUNCH_CHECK();
    const double stages = (double)total / grid * hp.nchunk * tb.cnt[0] * tb.cnt[1] * tb.cnt[2];
    const double stages_dummy = stages;
    std::vector<unsigned long long> h(10 * 1024);
    MGB_CUDA(cudaStreamSynchronize(st));
    MGB_CUDA(cudaMemcpy(h.data(), prof_buf, h.size() * 8, cudaMemcpyDeviceToHost));
    double ls = 0, ln = 0;
    for (int i = 0; i < grid; ++i) { ls += (double)h[8 * 1024 + 2 * i]; ln += (double)h[8 * 1024 + 2 * i + 1]; }
    fprintf(stderr, "[halo prof] B stages the MMA thread had to wait for: %.0f of %.0f per CTA, issue->ready %.0f clk\n", ln / grid, stages_dummy, ln > 0 ? ls / ln : 0.0);
    double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < grid; ++i) for (int k = 0; k < 8; ++k) s[k] += (double)h[8 * i + k] / grid;
    fprintf(stderr, "[halo prof] bn=%d tiles=%d bst=%d, per tap-stage: wait A %.0f, wait B %.0f, wait tmem %.0f, "
            "umma issue %.0f, commit %.0f, total %.0f clk; B producer: wait empty %.0f of %.0f\n", j.bn, hp.total_tiles,
            hp.bst, s[0] / stages, s[1] / stages, s[2] / stages, s[4] / stages, s[5] / stages, s[3] / stages,
            s[6] / stages, s[7] / stages);
    return 0;
  }
  MGB_CUDA(cudaFuncSetAttribute(halo::conv_halo_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  halo::conv_halo_kernel<false><<<grid, halo::THREADS, smem, st>>>(hp, ma_hi, ma_lo, mb_hi, mb_lo);
  MGB_LAUNCH_CHECK();
  return 0;
}

// returns 0 ok, 1 error, 2 "not applicable, use the cp.async kernel"
int tma_launch(const TmaJob& j, cudaStream_t st) {
  EncodeTiledFn enc = tma_encoder();
  if (!enc || !tma_enabled()) return 2;
  for (int i = 0; i < 3; ++i) if (j.X[i] < 1) return 2;
  tma::TmaParams tp;
  if (!choose_box(j.P, j.pmul, j.nbatch, tp.box, &tp.boxn)) return 2;
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  {
    cuuint64_t dims[5] = {(cuuint64_t)j.cpad, (cuuint64_t)j.X[2], (cuuint64_t)j.X[1], (cuuint64_t)j.X[0],
                          (cuuint64_t)j.nbatch};
    cuuint64_t strides[4];
    strides[0] = (cuuint64_t)j.cpad * 4;
    strides[1] = strides[0] * j.X[2];
    strides[2] = strides[1] * j.X[1];
    strides[3] = strides[2] * j.X[0];
    cuuint32_t box[5] = {32u, (cuuint32_t)(tp.box[2] * j.pmul[2]), (cuuint32_t)(tp.box[1] * j.pmul[1]),
                         (cuuint32_t)(tp.box[0] * j.pmul[0]), (cuuint32_t)tp.boxn};
    cuuint32_t estr[5] = {1u, (cuuint32_t)j.pmul[2], (cuuint32_t)j.pmul[1], (cuuint32_t)j.pmul[0], 1u};
    CUresult r1 = enc(&ma_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)j.a_hi, dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&ma_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)j.a_lo, dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) return 2;
  }
  // CTA pairs (tcgen05 cta_group::2, conv_tma2_kernel) are opt-in: MGB_CONV_2CTA=1.  Measured at C2
  // (N tile 128): 0.82 ms vs 0.52 ms for the single-CTA kernel — the UMMA stream itself runs slower
  // (the issuing thread never waits for data), so halving the B bytes per SM does not pay at this N.
  const char* e2 = getenv("MGB_CONV_2CTA");
  const bool pair = (e2 && e2[0] == '1') && j.bn >= 32 && j.bn % 32 == 0;
  {
    long long K = (long long)j.ntaps * j.cpad;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)j.rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 4};
    cuuint32_t box[2] = {32u, (cuuint32_t)(pair ? j.bn / 2 : j.bn)};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r1 = enc(&mb_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)j.b_hi, dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&mb_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)j.b_lo, dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) return 2;
  }
  tp.taps = j.taps; tp.out = j.out; tp.om = j.om;
  long long total = j.ny;
  for (int i = 0; i < 3; ++i) {
    tp.P[i] = j.P[i]; tp.pmul[i] = j.pmul[i]; tp.base[i] = j.base[i];
    tp.tiles[i] = (j.P[i] + tp.box[i] - 1) / tp.box[i];
    total *= tp.tiles[i];
  }
  tp.tiles_n = (int)((j.nbatch + tp.boxn - 1) / tp.boxn);
  total *= tp.tiles_n;
  if (total > INT32_MAX) return 2;
  tp.Nbatch = (int)j.nbatch; tp.Nout = j.nout; tp.nchunk = j.cpad / 32; tp.ntaps = j.ntaps;
  tp.bn = j.bn; tp.ny = j.ny; tp.total_tiles = (int)total;
  // N-concatenated B operand (see TmaParams::ncat): default on, MGB_CONV_NCAT=0 restores three UMMAs
  const char* e3 = getenv("MGB_CONV_NCAT");
  tp.ncat = (!pair && !(e3 && e3[0] == '0') && j.bn % 8 == 0 && 2 * j.bn <= 256) ? 1 : 0;
  tp.acc_cols = tp.ncat ? 2 * j.bn : j.bn;
  tp.tmem_cols = pow2_cols(2 * tp.acc_cols);
  if (pair) {
    constexpr int STAGES = 4;
    const int pix_tiles = (int)(total / j.ny);
    const long long pairs = (long long)((pix_tiles + 1) / 2) * j.ny;
    tp.total_tiles = (int)pairs;
    size_t smem = tma::smem_bytes(j.bn / 2, STAGES);
    MGB_CUDA(cudaFuncSetAttribute(tma::conv_tma2_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    int grid = (int)std::min<long long>(2 * pairs, sm_count() / 2 * 2);
    tma::conv_tma2_kernel<STAGES><<<grid, tma::THREADS, smem, st>>>(tp, ma_hi, ma_lo, mb_hi, mb_lo, pix_tiles);
    MGB_LAUNCH_CHECK();
    return 0;
  }
  constexpr int STAGES = 3;
  size_t smem = tma::smem_bytes(j.bn, STAGES);
  MGB_CUDA(cudaFuncSetAttribute(tma::conv_tma_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  int grid = (int)std::min<long long>(total, sm_count());
  tma::conv_tma_kernel<STAGES><<<grid, tma::THREADS, smem, st>>>(tp, ma_hi, ma_lo, mb_hi, mb_lo);
  MGB_LAUNCH_CHECK();
  return 0;
}

size_t tc_fwd_ws(const Plan& p) {
  TcPlanes P = tc_planes(0, p.N, p.C, p.Xp);
  return tc_layout(P.end, p.C, p.F, (int)p.Wp).end;
}

// `planes` (optional): channels-last hi|lo planes made by mgb_conv_make_planes; when null
// they are built in the workspace.
int tc_fwd(const Plan& p, const float* x, const float* w, float* y, const void* planes, void* ws,
           size_t ws_bytes, cudaStream_t st, float* pout = nullptr, unsigned char* pidx = nullptr) {
  long long M = p.N * p.Gp;
  if (M * p.F == 0) return 0;
  TcPlanes P = tc_planes(0, p.N, p.C, p.Xp);
  TcLayout L = tc_layout(P.end, p.C, p.F, (int)p.Wp);
  if (ws_bytes < L.end || !ws) return fail("conv fwd: workspace %zu < %zu bytes", ws_bytes, L.end);
  char* base = (char*)ws;
  const char* pl = planes ? (const char*)planes : base;
  float *ahi = (float*)(pl + P.hi_off), *alo = (float*)(pl + P.lo_off);
  float *bhi = (float*)(base + L.bhi_off), *blo = (float*)(base + L.blo_off);
  tc::TcTap* taps = (tc::TcTap*)(base + L.taps_off);
  int* list = (int*)(base + L.list_off);
  if (!planes && tc_prep_planes(x, ahi, alo, p.N, p.C, p.Xp, P.cpad, st)) return 1;
  TapBuild tb;
  for (int i = 0; i < 3; ++i) {
    tb.cnt[i] = p.W[i]; tb.W[i] = p.W[i]; tb.X[i] = p.X[i];
    for (int k = 0; k < p.W[i]; ++k) { tb.kidx[i][k] = (short)k; tb.delta[i][k] = (short)(k * p.D[i]); }
  }
  tb.cpad = P.cpad;
  build_tc_taps<<<1, 256, 0, st>>>(taps, list, tb);
  MGB_LAUNCH_CHECK();
  tc::PackArgs pa{L.rows, L.cpad, (int)p.F, (int)p.C, (int)p.Wp, L.ntaps, 0};
  tc::pack_filters_split<<<small_grid((long long)L.rows * L.K), 256, 0, st>>>(w, list, bhi, blo, pa);
  MGB_LAUNCH_CHECK();
  {
    TmaJob j;
    j.a_hi = ahi; j.a_lo = alo; j.b_hi = bhi; j.b_lo = blo; j.taps = taps; j.out = y;
    j.om.sample_stride = p.F * p.Gp; j.om.chan_stride = p.Gp;
    for (int i = 0; i < 3; ++i) {
      j.om.omul[i] = 1; j.om.obase[i] = 0;
      j.P[i] = p.G[i]; j.pmul[i] = p.S[i]; j.base[i] = -p.Pd[i]; j.X[i] = p.X[i];
    }
    j.om.os[2] = 1; j.om.os[1] = p.G[2]; j.om.os[0] = p.G[1] * p.G[2];
    j.nbatch = p.N; j.cpad = P.cpad; j.ntaps = L.ntaps; j.nout = (int)p.F; j.bn = L.bn; j.ny = L.ny;
    j.rows = L.rows; j.tb = &tb;
    if (pout) { j.pool = 1; j.pout = pout; j.pidx = pidx; }
    int rc = halo_launch(j, st);
    if (pout) return rc == 2 ? fail("conv_pool: the fused kernel does not apply to this geometry") : rc;
    if (rc == 2) rc = tma_launch(j, st);
    if (rc != 2) return rc;
  }
  tc::TcParams tp;
  tp.a_hi = ahi; tp.a_lo = alo; tp.taps = taps; tp.b_hi = bhi; tp.b_lo = blo; tp.out = y;
  tp.om.sample_stride = p.F * p.Gp; tp.om.chan_stride = p.Gp;
  for (int i = 0; i < 3; ++i) {
    tp.om.omul[i] = 1; tp.om.obase[i] = 0;
    tp.g.P[i] = p.G[i]; tp.g.pmul[i] = p.S[i]; tp.g.base[i] = -p.Pd[i]; tp.g.X[i] = p.X[i];
  }
  tp.om.os[2] = 1; tp.om.os[1] = p.G[2]; tp.om.os[0] = p.G[1] * p.G[2];
  tp.g.sample_stride = p.Xp * P.cpad; tp.g.pp = (int)p.Gp; tp.g.cpad = P.cpad;
  tp.g.nchunk = P.cpad / 32; tp.g.ntaps = L.ntaps;
  tp.M = (int)M; tp.N = (int)p.F; tp.bn = L.bn; tp.tmem_cols = pow2_cols(L.bn);
  return tc_launch(tp, L.ny, st);
}

size_t tc_dgrad_ws(const Plan& p, const std::vector<DgradClassPlan>& cls) {
  TcPlanes P = tc_planes(0, p.N, p.F, p.Gp);
  size_t off = P.end;
  for (auto& c : cls) off = tc_layout(off, p.F, p.C, c.ntap).end;
  return off;
}

int tc_dgrad(const Plan& p, const float* dy, const float* w, float* dx, const void* planes, void* ws,
             size_t ws_bytes, cudaStream_t st) {
  long long n_dx = p.N * p.C * p.Xp;
  if (n_dx == 0) return 0;
  std::vector<DgradClassPlan> cls;
  size_t unused = 0;
  dgrad_classes(p, sizeof(float), &cls, &unused);
  size_t need = tc_dgrad_ws(p, cls);
  if (ws_bytes < need || !ws) return fail("conv dgrad: workspace %zu < %zu bytes", ws_bytes, need);
  if (dgrad_needs_memset(p, cls)) MGB_CUDA(cudaMemsetAsync(dx, 0, (size_t)n_dx * sizeof(float), st));
  char* base = (char*)ws;
  TcPlanes P = tc_planes(0, p.N, p.F, p.Gp);
  const char* pl = planes ? (const char*)planes : base;
  float *ahi = (float*)(pl + P.hi_off), *alo = (float*)(pl + P.lo_off);
  if (!planes && tc_prep_planes(dy, ahi, alo, p.N, p.F, p.Gp, P.cpad, st)) return 1;
  size_t off = P.end;
  for (auto& c : cls) {
    TcLayout L = tc_layout(off, p.F, p.C, c.ntap);
    off = L.end;
    float *bhi = (float*)(base + L.bhi_off), *blo = (float*)(base + L.blo_off);
    tc::TcTap* taps = (tc::TcTap*)(base + L.taps_off);
    int* list = (int*)(base + L.list_off);
    TapBuild tb;
    for (int i = 0; i < 3; ++i) {
      tb.cnt[i] = c.dev.cnt[i]; tb.W[i] = p.W[i]; tb.X[i] = p.G[i];
      for (int k = 0; k < c.dev.cnt[i]; ++k) { tb.kidx[i][k] = c.dev.kidx[i][k]; tb.delta[i][k] = c.dev.e[i][k]; }
    }
    tb.cpad = P.cpad;
    build_tc_taps<<<1, 256, 0, st>>>(taps, list, tb);
    MGB_LAUNCH_CHECK();
    tc::PackArgs pa{L.rows, L.cpad, (int)p.F, (int)p.C, (int)p.Wp, L.ntaps, 1};
    tc::pack_filters_split<<<small_grid((long long)L.rows * L.K), 256, 0, st>>>(w, list, bhi, blo, pa);
    MGB_LAUNCH_CHECK();
    {
      TmaJob j;
      j.a_hi = ahi; j.a_lo = alo; j.b_hi = bhi; j.b_lo = blo; j.taps = taps; j.out = dx;
      j.om.sample_stride = p.C * p.Xp; j.om.chan_stride = p.Xp;
      for (int i = 0; i < 3; ++i) {
        j.om.omul[i] = p.S[i]; j.om.obase[i] = p.S[i] * c.qlo[i] + c.r[i] - p.Pd[i];
        j.P[i] = c.Q[i]; j.pmul[i] = 1; j.base[i] = c.qlo[i]; j.X[i] = p.G[i];
      }
      j.om.os[2] = 1; j.om.os[1] = p.X[2]; j.om.os[0] = p.X[1] * p.X[2];
      j.nbatch = p.N; j.cpad = P.cpad; j.ntaps = L.ntaps; j.nout = (int)p.C; j.bn = L.bn; j.ny = L.ny;
      j.rows = L.rows; j.tb = &tb;
      int rc = halo_launch(j, st);
      if (rc == 2) rc = tma_launch(j, st);
      if (rc == 1) return 1;
      if (rc == 0) continue;
    }
    tc::TcParams tp;
    tp.a_hi = ahi; tp.a_lo = alo; tp.taps = taps; tp.b_hi = bhi; tp.b_lo = blo; tp.out = dx;
    tp.om.sample_stride = p.C * p.Xp; tp.om.chan_stride = p.Xp;
    for (int i = 0; i < 3; ++i) {
      tp.om.omul[i] = p.S[i]; tp.om.obase[i] = p.S[i] * c.qlo[i] + c.r[i] - p.Pd[i];
      tp.g.P[i] = c.Q[i]; tp.g.pmul[i] = 1; tp.g.base[i] = c.qlo[i]; tp.g.X[i] = p.G[i];
    }
    tp.om.os[2] = 1; tp.om.os[1] = p.X[2]; tp.om.os[0] = p.X[1] * p.X[2];
    tp.g.sample_stride = p.Gp * P.cpad; tp.g.pp = c.Q[0] * c.Q[1] * c.Q[2]; tp.g.cpad = P.cpad;
    tp.g.nchunk = P.cpad / 32; tp.g.ntaps = L.ntaps;
    tp.M = (int)(p.N * tp.g.pp); tp.N = (int)p.C; tp.bn = L.bn; tp.tmem_cols = pow2_cols(L.bn);
    if (tc_launch(tp, L.ny, st)) return 1;
  }
  return 0;
}

struct TcWgradLayout {
  TcPlanes X, Y;            // channels-last planes of x and dy
  int cblocks, nblocks, nt, nbpc, mt, splits, kps;
  int tma_splits;           // split count of the TMA variant (boxes of 16 pixels)
  size_t taps_off, list_off, partial_off, end;
};

TcWgradLayout tc_wgrad_layout(const Plan& p) {
  TcWgradLayout L;
  L.X = tc_planes(0, p.N, p.C, p.Xp);
  L.Y = tc_planes(L.X.end, p.N, p.F, p.Gp);
  L.cblocks = L.X.cpad / 32;
  L.nblocks = (int)p.Wp * L.cblocks;
  L.nt = (L.nblocks + 7) / 8;
  L.nbpc = (L.nblocks + L.nt - 1) / L.nt;
  L.mt = (int)((p.F + tc::BM - 1) / tc::BM);
  long long K = p.N * p.Gp;
  int stages_total = (int)((K + tc::WG_KS - 1) / tc::WG_KS);
  int want = std::max(1, sm_count() / (L.mt * L.nt));
  int splits = std::max(1, std::min(want, std::max(1, stages_total / 32)));
  int per = (stages_total + splits - 1) / splits;
  L.splits = (stages_total + per - 1) / per;
  L.kps = per * tc::WG_KS;
  size_t off = align_up(L.Y.end, 256);
  L.taps_off = off; off += align_up((size_t)p.Wp * sizeof(tc::TcTap), 256);
  L.list_off = off; off += align_up((size_t)p.Wp * sizeof(int), 256);
  L.tma_splits = std::max(1, sm_count() / (L.mt * L.nt));
  L.partial_off = off;
  off += align_up((size_t)std::max(L.splits, L.tma_splits) * p.F * p.C * p.Wp * sizeof(float), 256);
  L.end = off;
  return L;
}

// returns 0 ok, 1 error, 2 not applicable
int wgrad_tma_launch(const Plan& p, TcWgradLayout& L, const float* xhi, const float* xlo, const float* yhi,
                     const float* ylo, const tc::TcTap* taps, float* partial, cudaStream_t st) {
  EncodeTiledFn enc = tma_encoder();
  if (!enc || !tma_enabled()) return 2;
  for (int i = 0; i < 3; ++i) if (p.X[i] < 1 || p.G[i] < 1) return 2;
  tma::WgradTmaParams wp;
  // blocked maps need the channel-block axis as a fifth dimension, so batch and axis 0 share one
  // (trivially with <= 2 spatial dims; otherwise axis 0 must be unpadded and tiled exactly by the
  // boxes), and the CTA's (tap, c-block) ranges must be whole taps
  // <= 2 spatial dims.  (X0 = G0 = 1 alone is not enough: a 3-tap outer axis with padding 1 also has them, and its
  // taps would walk into the neighbouring SAMPLE of the merged (batch, axis 0) dimension — found by scripts/fuzz_conv.py)
  const bool flat = p.X[0] == 1 && p.G[0] == 1 && p.W[0] == 1 && p.Pd[0] == 0;
  const bool mergeable = !flat && p.Pd[0] == 0;               // 3-D with an unpadded outer axis
  wp.blocked = ((flat || mergeable) && L.nbpc % L.cblocks == 0) ? 1 : 0;
  // pixels per pipeline stage: 32 (2 stages) by default; MGB_WGRAD_PIX=16 -> 16-pixel stages, 5 of them
  const char* epx = getenv("MGB_WGRAD_PIX");
  const int blocked_pix = (epx && atoi(epx) == 16) ? 16 : 32;
  wp.pix = wp.blocked ? blocked_pix : tma::WG_PIX;
  if (wp.blocked && !flat) {
    if (!choose_box(p.G, p.S, p.N, wp.box, &wp.boxn, wp.pix, true)) { wp.blocked = 0; wp.pix = tma::WG_PIX; }
  }
  if ((!wp.blocked || flat) && !choose_box(p.G, p.S, p.N, wp.box, &wp.boxn, wp.pix)) return 2;
  wp.X0 = p.X[0]; wp.G0 = p.G[0];
  CUtensorMap mdy_hi, mdy_lo, mx_hi, mx_lo;
  auto encode = [&](CUtensorMap* m, const float* base, int cpad, const int X[3], const int mul[3], int nblk) {
    cuuint64_t dims[5], strides[4];
    cuuint32_t box[5], estr[5];
    if (wp.blocked) {   // (32, X2, X1, N*X0, cpad/32)
      dims[0] = 32; dims[1] = X[2]; dims[2] = X[1]; dims[3] = (cuuint64_t)p.N * X[0]; dims[4] = cpad / 32;
      strides[0] = (cuuint64_t)cpad * 4; strides[1] = strides[0] * X[2]; strides[2] = strides[1] * X[1];
      strides[3] = 128;
      box[0] = 32; box[1] = wp.box[2] * mul[2]; box[2] = wp.box[1] * mul[1];
      box[3] = flat ? wp.boxn : wp.box[0] * mul[0]; box[4] = nblk;
      estr[0] = 1; estr[1] = mul[2]; estr[2] = mul[1]; estr[3] = flat ? 1 : mul[0]; estr[4] = 1;
    } else {            // (cpad, X2, X1, X0, N)
      dims[0] = cpad; dims[1] = X[2]; dims[2] = X[1]; dims[3] = X[0]; dims[4] = p.N;
      strides[0] = (cuuint64_t)cpad * 4; strides[1] = strides[0] * X[2]; strides[2] = strides[1] * X[1];
      strides[3] = strides[2] * X[0];
      box[0] = 32; box[1] = wp.box[2] * mul[2]; box[2] = wp.box[1] * mul[1]; box[3] = wp.box[0] * mul[0];
      box[4] = wp.boxn;
      estr[0] = 1; estr[1] = mul[2]; estr[2] = mul[1]; estr[3] = mul[0]; estr[4] = 1;
    }
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)base, dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  const int one[3] = {1, 1, 1};
  if (!encode(&mdy_hi, yhi, L.Y.cpad, p.G, one, 4) || !encode(&mdy_lo, ylo, L.Y.cpad, p.G, one, 4) ||
      !encode(&mx_hi, xhi, L.X.cpad, p.X, p.S, L.cblocks) || !encode(&mx_lo, xlo, L.X.cpad, p.X, p.S, L.cblocks))
    return 2;
  long long boxes = 1;
  for (int i = 0; i < 3; ++i) {
    wp.tiles[i] = (p.G[i] + wp.box[i] - 1) / wp.box[i];
    wp.S[i] = p.S[i]; wp.Pd[i] = p.Pd[i];
    boxes *= wp.tiles[i];
  }
  wp.tiles_n = (int)((p.N + wp.boxn - 1) / wp.boxn);
  boxes *= wp.tiles_n;
  if (boxes > INT32_MAX) return 2;
  int splits = (int)std::max<long long>(1, std::min<long long>(L.tma_splits, boxes / 8));
  long long per = (boxes + splits - 1) / splits;
  splits = (int)((boxes + per - 1) / per);
  L.tma_splits = splits;
  wp.taps = taps; wp.partial = partial;
  wp.F = (int)p.F; wp.C = (int)p.C; wp.Wp = (int)p.Wp; wp.cblocks = L.cblocks;
  wp.nblocks_total = L.nblocks; wp.nbpc = L.nbpc;
  wp.boxes_total = (int)boxes; wp.boxes_per_split = (int)per;
  wp.tmem_cols = pow2_cols(L.nbpc * 32);
  dim3 grid(L.mt, L.nt, splits);
  wp.prof = nullptr;
  static unsigned long long* wprof = nullptr;
  const bool prof = getenv("MGB_WGRAD_PROF") != nullptr;
  if (prof) {
    if (!wprof) MGB_CUDA(cudaMalloc(&wprof, 4096 * 4 * sizeof(unsigned long long)));
    MGB_CUDA(cudaMemsetAsync(wprof, 0, 4096 * 4 * sizeof(unsigned long long), st));
    if ((long long)L.mt * L.nt * splits <= 4096) wp.prof = wprof;
  }
  struct ProfPrint {
    bool on; unsigned long long* buf; int n, pix, nbpc; cudaStream_t st;
    ~ProfPrint() {
      if (!on) return;
      std::vector<unsigned long long> h((size_t)n * 4);
      cudaStreamSynchronize(st);
      cudaMemcpy(h.data(), buf, h.size() * 8, cudaMemcpyDeviceToHost);
      double wf = 0, tot = 0, S = 0;
      for (int i = 0; i < n; ++i) { wf += (double)h[4 * i]; tot += (double)h[4 * i + 1]; S += (double)h[4 * i + 2]; }
      fprintf(stderr, "[wgrad prof] %d CTAs, pix %d, nbpc %d: per stage wait full %.0f, total %.0f clk (tensor floor %.0f)\n",
              n, pix, nbpc, wf / S, tot / S, (pix / 8) * 3.0 * nbpc * 16.0);
    }
  } prof_print{wp.prof != nullptr, wprof, (int)(L.mt * L.nt * splits), wp.pix, L.nbpc, st};
  // dy multicast across the N tiles of a cluster (WgradTmaParams::mc): blocked maps only (dy is one TMA per
  // plane).  Opt-in (MGB_WGRAD_MC=1): measured at C2 it cuts the L2 -> SM bytes by 27 % and changes nothing —
  // 1940 vs 1876 clk per 32-pixel stage, and only 45 clusters of 3 fit on 148 SMs (0.636 vs 0.581 ms).  The
  // kernel is bound by shared-memory bytes (120 KB of operand reads + 81 KB of TMA writes per stage at
  // 128 B/clk = 1570 clk), which multicast does not reduce.
  const char* emc = getenv("MGB_WGRAD_MC");
  wp.mc = (wp.blocked && L.nt >= 2 && L.nt <= 8 && emc && emc[0] == '1') ? L.nt : 1;
  auto launch = [&](auto kern, int stages) -> int {
    size_t smem = tma::wgrad_smem_bytes(L.nbpc, stages, wp.pix);
    MGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(tma::THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = (unsigned)wp.mc; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (wp.mc > 1) {
      // all clusters must be resident at once (a second wave would double the time): fit the split count
      cfg.gridDim = dim3(L.mt, L.nt, 1);
      int max_clusters = 0;
      if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) != cudaSuccess || max_clusters < L.mt) {
        cudaGetLastError();
        wp.mc = 1; at[0].val.clusterDim.y = 1;
      } else if ((long long)splits * L.mt > max_clusters) {
        splits = std::max(1, max_clusters / L.mt);
        per = (boxes + splits - 1) / splits;
        splits = (int)((boxes + per - 1) / per);
        L.tma_splits = splits; wp.boxes_per_split = (int)per;
        prof_print.n = L.mt * L.nt * splits;
      }
    }
    cfg.gridDim = dim3(L.mt, L.nt, splits);
    MGB_CUDA(cudaLaunchKernelEx(&cfg, kern, wp, mdy_hi, mdy_lo, mx_hi, mx_lo));
    return 0;
  };
  int rc;
  if (wp.pix == 16 && wp.blocked) rc = launch(tma::wgrad_tma_kernel<5>, 5);
  else if (wp.pix == 32) rc = launch(tma::wgrad_tma_kernel<2>, 2);
  else rc = launch(tma::wgrad_tma_kernel<4>, 4);
  if (rc) return rc;
  MGB_LAUNCH_CHECK();
  return 0;
}


// ---- wgrad with a row halo of x (conv_wgrad2.cuh): returns 0 ok, 1 error, 2 not applicable ----
// MGB_WGRAD_ROWS: unset = when the N tile is at least 128 columns wide, "0" = never, "1" = whenever legal
int wgrad_rows_launch(const Plan& p, TcWgradLayout& L, const float* xhi, const float* xlo, const float* yhi,
                      const float* ylo, float* partial, cudaStream_t st) {
  EncodeTiledFn enc = tma_encoder();
  const char* er = getenv("MGB_WGRAD_ROWS");
  if (!enc || !tma_enabled() || (er && er[0] == '0')) return 2;
  for (int i = 0; i < 3; ++i) if (p.X[i] < 1 || p.G[i] < 1) return 2;
  // <= 2 spatial dims.  (X0 = G0 = 1 alone is not enough: a 3-tap outer axis with padding 1 also has them, and its
  // taps would walk into the neighbouring SAMPLE of the merged (batch, axis 0) dimension — found by scripts/fuzz_conv.py)
  const bool flat = p.X[0] == 1 && p.G[0] == 1 && p.W[0] == 1 && p.Pd[0] == 0;
  const bool mergeable = !flat && p.Pd[0] == 0;               // 3-D with an unpadded outer axis
  if (!(flat || mergeable) || p.S[1] != 1) return 2;
  // c-blocks per CTA: the N tile is W1 * cbc blocks of 32 columns, at most 8
  int cbc = 0;
  for (int c = L.cblocks; c >= 1; --c)
    if (L.cblocks % c == 0 && p.W[1] * c <= 8) { cbc = c; break; }
  if (cbc == 0 || (cbc > 1 && p.D[1] != 1)) return 2;
  const int ncols = p.W[1] * cbc * 32;
  // F <= 64: fold two column taps into the two halves of the 128 UMMA lanes (conv_wgrad2.cuh); MGB_WGRAD_FOLD=0 disables
  const char* efold = getenv("MGB_WGRAD_FOLD");
  const bool fold = p.F <= 64 && p.S[2] == 1 && p.W[2] >= 2 && !(efold && efold[0] == '0');
  if (ncols < (fold ? 96 : 128) && !(er && er[0] == '1')) return 2;
  const int ext = (p.W[1] - 1) * p.D[1];
  // pixel box b1 x b2 = 32 with b2 >= 4: fewest x bytes per stage over all boxes
  int b1 = 0, b2 = 0;
  long long best = -1;
  for (int c2 = 4; c2 <= 32; c2 <<= 1) {
    const int c1 = 32 / c2;
    if (c2 * p.S[2] > 256 || c1 + ext > 256) continue;
    const long long tiles = (long long)((p.G[1] + c1 - 1) / c1) * ((p.G[2] + c2 - 1) / c2);
    const long long cost = tiles * ((long long)(c1 + ext) * c2 * cbc + 4 * 32);     // x rows + dy, in 128-byte rows per plane
    if (best < 0 || cost < best) { best = cost; b1 = c1; b2 = c2; }
  }
  if (best < 0) return 2;
  tma::WgradRowsParams wp;
  wp.b1 = b1; wp.b2 = b2;
  wp.fold = fold ? 1 : 0;
  wp.tiles1 = (p.G[1] + b1 - 1) / b1; wp.tiles2 = (p.G[2] + (fold ? p.D[2] : 0) + b2 - 1) / b2;
  wp.G0 = p.G[0]; wp.X0 = p.X[0];
  wp.S0 = p.S[0]; wp.S2 = p.S[2]; wp.Pd1 = p.Pd[1]; wp.Pd2 = p.Pd[2];
  wp.D0 = p.D[0]; wp.D1 = p.D[1]; wp.D2 = p.D[2];
  wp.W0 = p.W[0]; wp.W1 = p.W[1]; wp.W2 = p.W[2];
  wp.F = (int)p.F; wp.C = (int)p.C; wp.Wp = (int)p.Wp; wp.cblocks = L.cblocks; wp.cbc = cbc;
  wp.rows = b1 + ext;
  wp.tmem_cols = pow2_cols(ncols);
  const long long boxes = (long long)p.N * p.G[0] * wp.tiles1 * wp.tiles2;
  if (boxes > INT32_MAX || (long long)p.N * std::max(p.G[0], p.X[0]) > INT32_MAX) return 2;
  const int ny = p.W[0] * (fold ? (p.W[2] + 1) / 2 : p.W[2]) * (L.cblocks / cbc);
  int splits = (int)std::max<long long>(1, std::min<long long>(std::max(1, sm_count() / (L.mt * ny)), boxes / 8));
  const long long per = (boxes + splits - 1) / splits;
  splits = (int)((boxes + per - 1) / per);
  if ((size_t)splits * p.F * p.C * p.Wp * sizeof(float) > L.end - L.partial_off) return 2;
  wp.boxes_total = (int)boxes; wp.boxes_per_split = (int)per;
  wp.partial = partial;
  // dy: (32, G2, G1, N*G0, Fpad/32), box (32, b2, b1, 1, 4) -> [f-block][row][column][128 B]
  // x : (32, X2, Cpad/32, X1, N*X0), box (32, b2 (stride S2), cbc, rows, 1) -> [row][c-block][column][128 B]
  CUtensorMap mdy_hi, mdy_lo, mx_hi, mx_lo;
  auto encode_dy = [&](CUtensorMap* m, const float* base) {
    const cuuint64_t row = (cuuint64_t)L.Y.cpad * 4;
    cuuint64_t dims[5] = {32, (cuuint64_t)p.G[2], (cuuint64_t)p.G[1], (cuuint64_t)p.N * p.G[0], (cuuint64_t)L.Y.cpad / 32};
    cuuint64_t strides[4] = {row, row * p.G[2], row * p.G[2] * p.G[1], 128};
    cuuint32_t box[5] = {32, (cuuint32_t)b2, (cuuint32_t)b1, 1, (cuuint32_t)(fold ? 2 : 4)};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)base, dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  auto encode_x = [&](CUtensorMap* m, const float* base) {
    const cuuint64_t row = (cuuint64_t)L.X.cpad * 4;
    cuuint64_t dims[5] = {32, (cuuint64_t)p.X[2], (cuuint64_t)L.X.cpad / 32, (cuuint64_t)p.X[1], (cuuint64_t)p.N * p.X[0]};
    cuuint64_t strides[4] = {row, 128, row * p.X[2], row * p.X[2] * p.X[1]};
    cuuint32_t box[5] = {32, (cuuint32_t)(b2 * p.S[2]), (cuuint32_t)cbc, (cuuint32_t)wp.rows, 1};
    cuuint32_t estr[5] = {1, (cuuint32_t)p.S[2], 1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)base, dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  if (!encode_dy(&mdy_hi, yhi) || !encode_dy(&mdy_lo, ylo) || !encode_x(&mx_hi, xhi) || !encode_x(&mx_lo, xlo)) return 2;
  wp.prof = nullptr;
  static unsigned long long* wprof = nullptr;
  const long long nctas = (long long)L.mt * ny * splits;
  if (getenv("MGB_WGRAD_PROF") && nctas <= 4096) {
    if (!wprof) MGB_CUDA(cudaMalloc(&wprof, 4096 * 4 * sizeof(unsigned long long)));
    MGB_CUDA(cudaMemsetAsync(wprof, 0, 4096 * 4 * sizeof(unsigned long long), st));
    wp.prof = wprof;
  }
  const size_t stage = 2 * 4 * 32 * 128 + 2 * (size_t)wp.rows * cbc * b2 * 128;
  const int stages = (int)std::min<size_t>(4, (227 * 1024 - 1280) / stage);
  if (stages < 2) return 2;
  auto launch = [&](auto kern) -> int {
    const size_t smem = tma::wgrad_rows_smem_bytes(wp.rows, cbc, b2, stages);
    MGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3(L.mt, ny, splits), tma::THREADS, smem, st>>>(wp, mdy_hi, mdy_lo, mx_hi, mx_lo);
    return 0;
  };
  int rc = stages == 4 ? launch(tma::wgrad_rows_kernel<4>) : stages == 3 ? launch(tma::wgrad_rows_kernel<3>)
                                                                        : launch(tma::wgrad_rows_kernel<2>);
  if (rc) return rc;
  MGB_LAUNCH_CHECK();
  L.tma_splits = splits;
  if (wp.prof) {
    std::vector<unsigned long long> h((size_t)nctas * 4);
    cudaStreamSynchronize(st);
    cudaMemcpy(h.data(), wprof, h.size() * 8, cudaMemcpyDeviceToHost);
    double wf = 0, tot = 0, S = 0;
    for (long long i = 0; i < nctas; ++i) { wf += (double)h[4 * i]; tot += (double)h[4 * i + 1]; S += (double)h[4 * i + 2]; }
    fprintf(stderr, "[wgrad prof rows] %lld CTAs, box %dx%d, %d halo rows, N %d, %d stages of %zu B: per stage wait full %.0f, total %.0f clk (tensor floor %.0f)\n",
            nctas, b1, b2, wp.rows, ncols, stages, stage, wf / S, tot / S, 4 * 3.0 * ncols / 2);
  }
  return 0;
}

int tc_wgrad(const Plan& p, const float* dy, const float* x, float* dw, const void* dy_planes,
             const void* x_planes, void* ws, size_t ws_bytes, cudaStream_t st) {
  long long n_dw = p.F * p.C * p.Wp;
  if (n_dw == 0) return 0;
  if (p.N * p.Gp == 0) {   // an empty shard still takes part in the cross-rank sum
    MGB_CUDA(cudaMemsetAsync(dw, 0, (size_t)n_dw * sizeof(float), st));
    return (g_wgrad_dp && peer_ready()) ? finish_dw<float>(dw, dw, n_dw, 1, st) : 0;
  }
  TcWgradLayout L = tc_wgrad_layout(p);
  if (ws_bytes < L.end || !ws) return fail("conv wgrad: workspace %zu < %zu bytes", ws_bytes, L.end);
  char* base = (char*)ws;
  // external planes start at offset 0 of their own buffer (tc_planes(0, ...) layout)
  TcPlanes X0 = tc_planes(0, p.N, p.C, p.Xp), Y0 = tc_planes(0, p.N, p.F, p.Gp);
  float *xhi = x_planes ? (float*)((char*)x_planes + X0.hi_off) : (float*)(base + L.X.hi_off);
  float *xlo = x_planes ? (float*)((char*)x_planes + X0.lo_off) : (float*)(base + L.X.lo_off);
  float *yhi = dy_planes ? (float*)((char*)dy_planes + Y0.hi_off) : (float*)(base + L.Y.hi_off);
  float *ylo = dy_planes ? (float*)((char*)dy_planes + Y0.lo_off) : (float*)(base + L.Y.lo_off);
  tc::TcTap* taps = (tc::TcTap*)(base + L.taps_off);
  int* list = (int*)(base + L.list_off);
  float* partial = (float*)(base + L.partial_off);
  if (!x_planes && tc_prep_planes(x, xhi, xlo, p.N, p.C, p.Xp, L.X.cpad, st)) return 1;
  if (!dy_planes && tc_prep_planes(dy, yhi, ylo, p.N, p.F, p.Gp, L.Y.cpad, st)) return 1;
  TapBuild tb;
  for (int i = 0; i < 3; ++i) {
    tb.cnt[i] = p.W[i]; tb.W[i] = p.W[i]; tb.X[i] = p.X[i];
    for (int k = 0; k < p.W[i]; ++k) { tb.kidx[i][k] = (short)k; tb.delta[i][k] = (short)(k * p.D[i]); }
  }
  tb.cpad = L.X.cpad;
  // the row-halo kernel addresses taps by box coordinates: no tap table
  {
    const int rc = wgrad_rows_launch(p, L, xhi, xlo, yhi, ylo, partial, st);
    if (rc == 1) return 1;
    if (rc == 0) return finish_dw<float>(partial, dw, n_dw, L.tma_splits, st);
  }
  build_tc_taps<<<1, 256, 0, st>>>(taps, list, tb);
  MGB_LAUNCH_CHECK();
  tc::TcWgradParams wp;
  wp.a_hi = yhi; wp.a_lo = ylo; wp.b_hi = xhi; wp.b_lo = xlo; wp.taps = taps; wp.partial = partial;
  wp.fpad = L.Y.cpad; wp.cpad = L.X.cpad; wp.cblocks = L.cblocks;
  wp.F = (int)p.F; wp.C = (int)p.C; wp.Wp = (int)p.Wp;
  for (int i = 0; i < 3; ++i) { wp.G[i] = p.G[i]; wp.S[i] = p.S[i]; wp.Pd[i] = p.Pd[i]; wp.X[i] = p.X[i]; }
  wp.gp = (int)p.Gp; wp.Ktot = (int)(p.N * p.Gp); wp.k_per_split = L.kps;
  wp.nblocks_total = L.nblocks; wp.nbpc = L.nbpc; wp.tmem_cols = pow2_cols(L.nbpc * 32);
  // the per-tap TMA variant; the cp.async kernel is the fallback
  {
    int rc = wgrad_tma_launch(p, L, xhi, xlo, yhi, ylo, taps, partial, st);
    if (rc == 1) return 1;
    if (rc == 0) {
      return finish_dw<float>(partial, dw, n_dw, L.tma_splits, st);
    }
  }
  constexpr int STAGES = 4;
  size_t smem = tc::wgrad_smem_bytes(L.nbpc, STAGES);
  MGB_CUDA(cudaFuncSetAttribute(tc::wgrad_tc_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  dim3 grid(L.mt, L.nt, L.splits);
  tc::wgrad_tc_kernel<STAGES><<<grid, tc::THREADS, smem, st>>>(wp);
  MGB_LAUNCH_CHECK();
  return finish_dw<float>(partial, dw, n_dw, L.splits, st);
}

// ---- conv -> max-pool fusion -----------------------------------------------

// the fused forward applies to: float32 tensor-core eligible, 2 spatial dims, conv stride 1, pool (2,2)
// stride (2,2) over exactly the conv output (even extents), and a halo tiling with <= 2 samples per tile
bool conv_pool_applies(const Plan& p, const mgb_pool_desc* pd) {
  if (!pd || pd->nd != 2 || pd->mode != 0 || pd->dtype != MGB_F32) return false;
  if (!tc_eligible(p, 0) || !tc_eligible(p, 1) || !tc_eligible(p, 2)) return false;
  if (p.X[0] != 1 || p.W[0] != 1 || p.G[0] != 1) return false;                  // 2 spatial dims (normalised to 3)
  for (int i = 0; i < 3; ++i) if (p.S[i] != 1) return false;
  for (int i = 0; i < 2; ++i)
    if (pd->P[i] != 2 || pd->stride[i] != 2 || pd->X[i] != p.G[i + 1] || p.G[i + 1] % 2) return false;
  if (pd->lead != p.N * p.F || p.N * p.F * p.Gp == 0) return false;
  TapBuild tb;
  for (int i = 0; i < 3; ++i) {
    tb.cnt[i] = p.W[i]; tb.W[i] = p.W[i]; tb.X[i] = p.X[i];
    for (int k = 0; k < p.W[i]; ++k) { tb.kidx[i][k] = (short)k; tb.delta[i][k] = (short)(k * p.D[i]); }
  }
  TcLayout L = tc_layout(0, p.C, p.F, (int)p.Wp);
  TmaJob j;
  for (int i = 0; i < 3; ++i) { j.P[i] = p.G[i]; j.pmul[i] = 1; j.base[i] = -p.Pd[i]; j.X[i] = p.X[i]; }
  j.nbatch = p.N; j.cpad = L.cpad; j.ntaps = L.ntaps; j.nout = (int)p.F; j.bn = L.bn; j.ny = L.ny; j.rows = L.rows;
  j.tb = &tb; j.pool = 1;
  return halo_launch(j, 0, /*dry_run=*/true) == 0;
}

// dy of the fused layer, written straight as the channels-last TF32 hi / lo planes the dgrad and wgrad
// kernels read (never as an NCHW tensor): dy[n][h][w][f] = dp[n][f][h/2][w/2] where the window's first
// maximum sits, else 0 (pooling.py:109-149 for a tiling pool).  A CTA turns 32 channels x 64 pooled pixels
// of dp / argmax (read along the pixels) into 256 output pixels x 128 B per plane (written along the channels).
template <bool FROM_Y>
__global__ void __launch_bounds__(256) pool_bwd_planes_kernel(const float* __restrict__ dp,
                                                              const unsigned char* __restrict__ am,
                                                              const float* __restrict__ y, float* __restrict__ dy,
                                                              float* __restrict__ hi, float* __restrict__ lo,
                                                              int F, int PH, int PW, int cpad) {
  __shared__ float val[32][65];
  __shared__ unsigned char pos[32][68];
  const int PP = PH * PW;
  const int W = 2 * PW;
  const int q0 = blockIdx.x * 64, c0 = blockIdx.y * 32, n = blockIdx.z;
  {
    // a warp walks 4 channels; its 32 lanes cover 32 consecutive windows of one channel, so every access is a
    // contiguous run (128 B of dp, 256 B per row of y / dy) instead of eight 32-byte pieces
    const int lane = threadIdx.x & 31, w4 = (threadIdx.x >> 5) * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int ch = w4 + k, c = c0 + ch;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int ql = lane + 32 * i, q = q0 + ql;
        float v = 0.f;
        unsigned char a = 255;
        if (c < F && q < PP) {
          const long long o = ((long long)n * F + c) * PP + q;
          v = __ldg(dp + o);
          if (FROM_Y) {
            // the unfused layer: recompute the window's first maximum from y (pooling.py:111) and write dy
            // (the public y.grad) in the same pass; 8-byte accesses, two rows per window
            const int ph = q / PW, pw = q - ph * PW;
            const long long yo = ((long long)n * F + c) * (4LL * PP) + (long long)(2 * ph) * W + 2 * pw;
            const float2 r0 = __ldg(reinterpret_cast<const float2*>(y + yo));
            const float2 r1 = __ldg(reinterpret_cast<const float2*>(y + yo + W));
            float b = r0.x;
            int kk = 0;
            if (r0.y > b || (r0.y != r0.y && b == b)) { b = r0.y; kk = 1; }
            if (r1.x > b || (r1.x != r1.x && b == b)) { b = r1.x; kk = 2; }
            if (r1.y > b || (r1.y != r1.y && b == b)) { b = r1.y; kk = 3; }
            a = (unsigned char)kk;
            *reinterpret_cast<float2*>(dy + yo) = make_float2(kk == 0 ? v : 0.f, kk == 1 ? v : 0.f);
            *reinterpret_cast<float2*>(dy + yo + W) = make_float2(kk == 2 ? v : 0.f, kk == 3 ? v : 0.f);
          } else {
            a = __ldg(am + o);
          }
        }
        val[ch][ql] = v;
        pos[ch][ql] = a;
      }
    }
  }
  __syncthreads();
  const int cq = threadIdx.x & 7;                               // 8 lanes = the 32 channels of one output pixel
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int item = (threadIdx.x >> 3) + 32 * it;              // 256 output pixels: (pooled pixel, position)
    const int ql = item >> 2, ps = item & 3;
    const int q = q0 + ql;
    if (q < PP) {
      const int ph = q / PW, pw = q - ph * PW;
      const long long px = (long long)(2 * ph + (ps >> 1)) * W + 2 * pw + (ps & 1);
      float4 h4, l4;
      float* hp = reinterpret_cast<float*>(&h4);
      float* lp = reinterpret_cast<float*>(&l4);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float v = pos[4 * cq + k][ql] == ps ? val[4 * cq + k][ql] : 0.f;
        uint32_t h, l;
        split_tf32(v, h, l);
        hp[k] = __uint_as_float(h);
        lp[k] = __uint_as_float(l);
      }
      const long long o = ((long long)n * (4LL * PP) + px) * cpad + c0 + 4 * cq;
      *reinterpret_cast<float4*>(hi + o) = h4;
      *reinterpret_cast<float4*>(lo + o) = l4;
    }
  }
}

// ---- the three convolutions ------------------------------------------------------------

// ---- FP64 fwd / dgrad with TMA-staged operands (conv_f64_tma.cuh): 0 ok, 1 error, 2 not applicable ----
// MGB_F64_PATH=cpasync keeps float64 on the cp.async kernels of conv_gemm.cuh (A/B comparisons, tests)
struct F64Job {
  const double* src;            // gathered tensor [N][CH][X0][X1][X2]
  int X[3];                     // its spatial extents
  long long CH, nbatch;         // its channels (contracted), samples
  int P[3], pmul[3], base[3];   // pixel grid; source coord = pixel * pmul + base + tap delta
  int ntaps;
  short d[f64t::kMaxTaps][3];   // tap coordinate deltas
  short widx[f64t::kMaxTaps];   // flat filter tap of each tap
  const double* w;              // filters (F, C, Wp)
  long long F, C, Wp;
  int rows_are_f;               // output channels are filters (fwd) or input channels (dgrad)
  double* out;
  OutMap om;
  double* wpack;                // rows x ntaps x chpad doubles of workspace
  int pitch2 = 0;               // row length of src in elements (0: X[2]); even — see f64_pitched
};

// TMA wants every global stride a multiple of 16 bytes: a float64 tensor with an odd innermost extent is first
// copied to rows one element longer (the extra element is never read: the map's extent stays X2)
inline int f64_pitch(int X2) { return X2 + (X2 & 1); }
inline size_t f64_pitched_bytes(long long rows, int X2) {
  return (X2 & 1) ? align_up((size_t)rows * f64_pitch(X2) * sizeof(double), 1024) + 1024 : 0;
}
__global__ void pitch_rows_f64_kernel(const double* __restrict__ src, double* __restrict__ dst, long long rows, int X2,
                                      int X2p) {
  const long long total = rows * X2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / X2;
    dst[r * X2p + (i - r * X2)] = src[i];
  }
}
int f64_pitch_copy(const double* src, double* dst, long long rows, int X2, cudaStream_t st) {
  pitch_rows_f64_kernel<<<small_grid(rows * X2), 256, 0, st>>>(src, dst, rows, X2, f64_pitch(X2));
  MGB_LAUNCH_CHECK();
  return 0;
}

inline size_t f64_pack_bytes(long long rows, long long ntaps, long long ch) {
  return align_up((size_t)rows * ntaps * round_up(ch, 16) * sizeof(double), 1024);
}

// forward workspace of the mma.sync / DMMA path: tap table, then (float64) the packed filters of the TMA kernel
inline size_t fwd_tab_bytes(const Plan& p) { return align_up((size_t)(p.C * p.Wp) * sizeof(TapEntry), 1024); }
inline size_t fwd_ws(const Plan& p) {   // + 1024: the packed filters start at the next 1024-byte boundary of whatever buffer arrives
  return fwd_tab_bytes(p) + (p.dtype == MGB_F64 ? f64_pack_bytes(p.F, p.Wp, p.C) + 1024 +
                                                      f64_pitched_bytes(p.N * p.C * p.X[0] * p.X[1], p.X[2]) : 0);
}
inline double* align1024(void* ptr) { return (double*)(((uintptr_t)ptr + 1023) & ~(uintptr_t)1023); }

bool f64_tma_enabled() {
  const char* e = getenv("MGB_F64_PATH");
  return !(e && strcmp(e, "cpasync") == 0) && tma_enabled();
}

// dry: decide only (0 = would run)
int f64_fwdlike_tma(const F64Job& j, cudaStream_t st, bool dry = false) {
  EncodeTiledFn enc = tma_encoder();
  if (!enc || !f64_tma_enabled()) return 2;
  const long long rows = j.rows_are_f ? j.F : j.C;
  if (j.ntaps < 1 || j.ntaps > f64t::kMaxTaps || j.CH < 32 || rows < 48) return 2;
  const int chpad = round_up(j.CH, 16);
  if ((long long)chpad * 100 > j.CH * 115) return 2;                           // padded channels are wasted DMMAs
  const int pitch2 = j.pitch2 ? j.pitch2 : j.X[2];
  if (!dry && (pitch2 % 2 || ((uintptr_t)j.src & 15) || ((uintptr_t)j.wpack & 1023))) return 2;   // TMA strides are multiples of 16 B
  for (int i = 0; i < 3; ++i) if (j.pmul[i] > 8 || j.P[i] < 1 || j.X[i] < 1) return 2;   // TMA traversal strides are <= 8
  if (j.pmul[2] > 2) return 2;           // rows are contiguous along the innermost axis: instantiated for strides 1 and 2
  const bool wide = rows > 64;                                                   // 128 x 128 tiles, else 256 x 64
  const int BMt = wide ? 128 : 256;
  // pixel box: 8 along the last axis, BM / 8 over the other axes and the batch (powers of two), fewest tiles
  int box[3] = {1, 1, 8}, boxn = 1;
  long long best = -1;
  const int rest = BMt / 8;
  for (int b0 = 1; b0 <= rest; b0 <<= 1)
    for (int b1 = 1; b0 * b1 <= rest; b1 <<= 1) {
      const int bn = rest / (b0 * b1);
      if (b0 * j.pmul[0] > 256 || b1 * j.pmul[1] > 256) continue;
      const long long tiles = (long long)((j.P[0] + b0 - 1) / b0) * ((j.P[1] + b1 - 1) / b1) * ((j.P[2] + 7) / 8) *
                              ((j.nbatch + bn - 1) / bn);
      if (best < 0 || tiles < best) { best = tiles; box[0] = b0; box[1] = b1; boxn = bn; }
    }
  const long long pixels = j.nbatch * j.P[0] * j.P[1] * j.P[2];
  if (best < 0 || best > INT32_MAX || best * BMt * 100 > pixels * 130) return 2;   // ragged tiles would cost more than they save
  if (dry) return 0;
  f64t::FwdLikeParams fp;
  fp.out = j.out; fp.om = j.om;
  for (int i = 0; i < 3; ++i) {
    fp.P[i] = j.P[i]; fp.pmul[i] = j.pmul[i]; fp.base[i] = j.base[i]; fp.box[i] = box[i];
    fp.tiles[i] = (j.P[i] + box[i] - 1) / box[i];
  }
  fp.boxn = boxn; fp.tiles_n = (int)((j.nbatch + boxn - 1) / boxn);
  fp.Nbatch = (int)j.nbatch; fp.Nout = (int)rows; fp.nchunk = chpad / 16;
  fp.taps.n = j.ntaps;
  f64t::PackTaps pt;
  pt.n = j.ntaps;
  for (int t = 0; t < j.ntaps; ++t) {
    for (int i = 0; i < 3; ++i) fp.taps.d[t][i] = j.d[t][i];
    pt.widx[t] = j.widx[t];
  }
  const long long Ktot = (long long)j.ntaps * chpad;
  f64t::pack_w_f64_kernel<<<small_grid(rows * Ktot), 256, 0, st>>>(j.w, j.wpack, (int)j.F, (int)j.C, (int)j.Wp,
                                                                    j.rows_are_f, chpad, pt);
  MGB_LAUNCH_CHECK();
  CUtensorMap ma, mb;
  {
    const cuuint64_t xp = (cuuint64_t)j.X[0] * j.X[1] * pitch2;
    cuuint64_t dims[5] = {(cuuint64_t)j.X[2], (cuuint64_t)j.CH, (cuuint64_t)j.X[1], (cuuint64_t)j.X[0], (cuuint64_t)j.nbatch};
    cuuint64_t strides[4] = {xp * 8, (cuuint64_t)pitch2 * 8, (cuuint64_t)j.X[1] * pitch2 * 8, (cuuint64_t)j.CH * xp * 8};
    cuuint32_t bx[5] = {(cuuint32_t)f64t::row_pixels(j.pmul[2]), 16, (cuuint32_t)(box[1] * j.pmul[1]),
                        (cuuint32_t)(box[0] * j.pmul[0]), (cuuint32_t)boxn};
    cuuint32_t estr[5] = {1, 1, (cuuint32_t)j.pmul[1], (cuuint32_t)j.pmul[0], 1};
    if (enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, (void*)j.src, dims, strides, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return 2;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)Ktot * 8};
    cuuint32_t bx[2] = {16, (cuuint32_t)(wide ? 128 : 64)};
    cuuint32_t estr[2] = {1, 1};
    if (enc(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)j.wpack, dims, strides, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return 2;
  }
  auto launch = [&](auto kern, size_t smem, int bn) -> int {
    MGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3((unsigned)best, (unsigned)((rows + bn - 1) / bn)), f64t::THREADS, smem, st>>>(fp, ma, mb);
    return 0;
  };
  int lrc;
  if (wide && j.pmul[2] == 1) lrc = launch(f64t::f64_fwdlike_tma_kernel<128, 128, 5, 1>, f64t::fwdlike_smem_bytes<128, 128, 5, 1>(), 128);
  else if (wide)              lrc = launch(f64t::f64_fwdlike_tma_kernel<128, 128, 4, 2>, f64t::fwdlike_smem_bytes<128, 128, 4, 2>(), 128);
  else if (j.pmul[2] == 1)    lrc = launch(f64t::f64_fwdlike_tma_kernel<256, 64, 4, 1>, f64t::fwdlike_smem_bytes<256, 64, 4, 1>(), 64);
  else                        lrc = launch(f64t::f64_fwdlike_tma_kernel<256, 64, 2, 2>, f64t::fwdlike_smem_bytes<256, 64, 2, 2>(), 64);
  if (lrc) return lrc;
  MGB_LAUNCH_CHECK();
  return 0;
}


template <typename T>
int fwd_t(const Plan& p, const T* x, const T* w, T* y, void* ws, size_t ws_bytes, cudaStream_t st) {
  long long K = p.C * p.Wp, M = p.N * p.Gp;
  if (M * p.F == 0) return 0;
  if (K == 0) { MGB_CUDA(cudaMemsetAsync(y, 0, (size_t)(M * p.F) * sizeof(T), st)); return 0; }
  size_t need = fwd_ws(p);
  if (ws_bytes < need || !ws) return fail("conv fwd: workspace %zu < %zu bytes", ws_bytes, need);
  FwdParams<T> fp;
  fp.om.sample_stride = p.F * p.Gp; fp.om.chan_stride = p.Gp;
  for (int i = 0; i < 3; ++i) { fp.om.omul[i] = 1; fp.om.obase[i] = 0; }
  fp.om.os[2] = 1; fp.om.os[1] = p.G[2]; fp.om.os[0] = p.G[1] * p.G[2];
  if constexpr (sizeof(T) == 8) {
    if (p.Wp <= f64t::kMaxTaps) {           // TMA-staged DMMA kernel first
      F64Job j;
      j.src = x; j.CH = p.C; j.nbatch = p.N;
      for (int i = 0; i < 3; ++i) { j.X[i] = p.X[i]; j.P[i] = p.G[i]; j.pmul[i] = p.S[i]; j.base[i] = -p.Pd[i]; }
      j.ntaps = (int)p.Wp;
      for (int t = 0; t < j.ntaps; ++t) {
        int r = t;
        const int t2 = r % p.W[2]; r /= p.W[2];
        const int t1 = r % p.W[1], t0 = r / p.W[1];
        j.d[t][0] = (short)(t0 * p.D[0]); j.d[t][1] = (short)(t1 * p.D[1]); j.d[t][2] = (short)(t2 * p.D[2]);
        j.widx[t] = (short)t;
      }
      j.w = w; j.F = p.F; j.C = p.C; j.Wp = p.Wp; j.rows_are_f = 1;
      j.out = y; j.om = fp.om; j.wpack = align1024((char*)ws + fwd_tab_bytes(p));
      bool go = true;
      if (p.X[2] & 1) {                      // odd rows: pitched copy first (only when the kernel will run)
        go = f64_fwdlike_tma(j, st, true) == 0;
        if (go) {
          double* xp2 = align1024((char*)j.wpack + f64_pack_bytes(p.F, p.Wp, p.C));
          if (f64_pitch_copy((const double*)x, xp2, p.N * p.C * p.X[0] * p.X[1], p.X[2], st)) return 1;
          j.src = xp2; j.pitch2 = f64_pitch(p.X[2]);
        }
      }
      const int rc = go ? f64_fwdlike_tma(j, st) : 2;
      if (rc != 2) return rc;
    }
  }
  {
    const int rc = thin2d_fwd<T>(p, x, w, y, st);      // thin 1-D / 2-D layers: input band in shared memory
    if (rc != 2) return rc;
  }
  TapEntry* tab = (TapEntry*)ws;
  build_x_table<<<small_grid(K), 256, 0, st>>>(tab, x_table_args(p));
  MGB_LAUNCH_CHECK();
  fp.src = x; fp.tab = tab; fp.B = w; fp.ldb = K; fp.out = y;
  fp.ga = geom_x_by_output(p);
  fp.M = (int)M; fp.N = (int)p.F; fp.K = (int)K;
  return run_fwdlike<T>(fp, st);
}

template <typename T>
int dgrad_t(const Plan& p, const T* dy, const T* w, T* dx, void* ws, size_t ws_bytes,
            cudaStream_t st) {
  long long n_dx = p.N * p.C * p.Xp;
  if (n_dx == 0) return 0;
  if (p.F * p.Gp * p.Wp == 0) { MGB_CUDA(cudaMemsetAsync(dx, 0, (size_t)n_dx * sizeof(T), st)); return 0; }
  std::vector<DgradClassPlan> cls;
  size_t need = 0;
  dgrad_classes(p, sizeof(T), &cls, &need);
  const size_t need_classes = need;
  if (sizeof(T) == 8) need += f64_pitched_bytes(p.N * p.F * p.G[0] * p.G[1], p.G[2]);
  if (ws_bytes < need || (need && !ws)) return fail("conv dgrad: workspace %zu < %zu bytes", ws_bytes, need);
  if (dgrad_needs_memset(p, cls)) MGB_CUDA(cudaMemsetAsync(dx, 0, (size_t)n_dx * sizeof(T), st));
  const double* dy_pitched = nullptr;      // float64 with odd rows: made once, when the first class wants it
  for (auto& c : cls) {
    TapEntry* tab = (TapEntry*)((char*)ws + c.tab_off);
    T* wpack = (T*)((char*)ws + c.wp_off);
    FwdParams<T> fp;
    fp.ga.sample_stride = p.F * p.Gp;
    for (int i = 0; i < 3; ++i) {
      fp.ga.P[i] = c.Q[i]; fp.ga.pmul[i] = 1; fp.ga.base[i] = c.qlo[i]; fp.ga.X[i] = p.G[i];
      fp.om.omul[i] = p.S[i];
      fp.om.obase[i] = p.S[i] * c.qlo[i] + c.r[i] - p.Pd[i];
    }
    fp.ga.xs[2] = 1; fp.ga.xs[1] = p.G[2]; fp.ga.xs[0] = p.G[1] * p.G[2];
    fp.ga.pp = c.Q[0] * c.Q[1] * c.Q[2];
    fp.om.sample_stride = p.C * p.Xp; fp.om.chan_stride = p.Xp;
    fp.om.os[2] = 1; fp.om.os[1] = p.X[2]; fp.om.os[0] = p.X[1] * p.X[2];
    if constexpr (sizeof(T) == 8) {
      if (c.ntap <= f64t::kMaxTaps) {         // TMA-staged DMMA kernel first
        F64Job j;
        j.src = dy; j.CH = p.F; j.nbatch = p.N;
        for (int i = 0; i < 3; ++i) { j.X[i] = p.G[i]; j.P[i] = c.Q[i]; j.pmul[i] = 1; j.base[i] = c.qlo[i]; }
        j.ntaps = c.ntap;
        for (int t = 0; t < c.ntap; ++t) {
          int r = t;
          const int j2 = r % c.dev.cnt[2]; r /= c.dev.cnt[2];
          const int j1 = r % c.dev.cnt[1], j0 = r / c.dev.cnt[1];
          j.d[t][0] = c.dev.e[0][j0]; j.d[t][1] = c.dev.e[1][j1]; j.d[t][2] = c.dev.e[2][j2];
          j.widx[t] = (short)((c.dev.kidx[0][j0] * p.W[1] + c.dev.kidx[1][j1]) * p.W[2] + c.dev.kidx[2][j2]);
        }
        j.w = w; j.F = p.F; j.C = p.C; j.Wp = p.Wp; j.rows_are_f = 0;
        j.out = dx; j.om = fp.om; j.wpack = align1024(wpack);
        bool go = true;
        if (p.G[2] & 1) {
          go = f64_fwdlike_tma(j, st, true) == 0;
          if (go && !dy_pitched) {
            double* d2 = align1024((char*)ws + need_classes);
            if (f64_pitch_copy((const double*)dy, d2, p.N * p.F * p.G[0] * p.G[1], p.G[2], st)) return 1;
            dy_pitched = d2;
          }
          if (go) { j.src = dy_pitched; j.pitch2 = f64_pitch(p.G[2]); }
        }
        const int rc = go ? f64_fwdlike_tma(j, st) : 2;
        if (rc == 1) return 1;
        if (rc == 0) continue;
      }
    }
    build_dgrad_class<T><<<small_grid(c.K * (p.C + 1)), 256, 0, st>>>(tab, wpack, w, c.dev);
    MGB_LAUNCH_CHECK();
    fp.src = dy; fp.tab = tab; fp.B = wpack; fp.ldb = c.K; fp.out = dx;
    fp.M = (int)(p.N * fp.ga.pp); fp.N = (int)p.C; fp.K = (int)c.K;
    if (run_fwdlike<T>(fp, st)) return 1;
  }
  return 0;
}

// ---- FP64 wgrad with TMA-staged operands (conv_f64_tma.cuh) ----
struct F64WgradPlan {
  int NB, mt, nt, cblocks, nblocks, tiles1, tiles2, splits;
  long long boxes, per;
};

// geometry-only part of the decision (also sizes the workspace): false = not applicable
bool f64_wgrad_tma_plan(const Plan& p, F64WgradPlan* out) {
  if (p.dtype != MGB_F64 || p.F < 96 || p.C < 32) return false;
  if (p.S[0] > 8 || p.S[1] > 8 || p.S[2] > 2 || p.N * p.Gp == 0) return false;   // TMA traversal strides <= 8; innermost: 1 or 2
  F64WgradPlan w;
  w.cblocks = (int)((p.C + 31) / 32);
  if ((long long)w.cblocks * 32 * 100 > p.C * 115) return false;               // padded channels are wasted DMMAs
  w.nblocks = (int)p.Wp * w.cblocks;
  const int pad3 = (w.nblocks + 2) / 3 * 3, pad4 = (w.nblocks + 3) / 4 * 4;
  w.NB = pad3 < pad4 ? 3 : 4;
  if ((long long)std::min(pad3, pad4) * 100 > (long long)w.nblocks * 134) return false;
  w.tiles1 = (p.G[1] + 1) / 2; w.tiles2 = (p.G[2] + 7) / 8;
  w.boxes = (long long)p.N * p.G[0] * w.tiles1 * w.tiles2;
  if (w.boxes > INT32_MAX || w.boxes * 16 * 100 > p.N * p.Gp * 130) return false;   // ragged 2 x 8 boxes
  w.mt = (int)((p.F + 127) / 128); w.nt = (w.nblocks + w.NB - 1) / w.NB;
  // K splits: the split count whose CTA count fills whole waves of the machine best (one CTA per SM), with at
  // least 32 k-slabs per split; ties go to fewer splits (less partial-sum traffic)
  const int tiles = w.mt * w.nt, sms = sm_count();
  int splits = 1;
  double best_eff = 0.0;
  for (int s2 = 1; s2 <= 64 && (s2 == 1 || w.boxes / s2 >= 32); ++s2) {
    const long long ctas = (long long)tiles * s2;
    const double eff = (double)ctas / (double)(((ctas + sms - 1) / sms) * sms);
    if (eff > best_eff + 0.02) { best_eff = eff; splits = s2; }
  }
  w.per = (w.boxes + splits - 1) / splits;
  w.splits = (int)((w.boxes + w.per - 1) / w.per);
  *out = w;
  return true;
}

// returns 0 ok (sets *splits), 1 error, 2 not applicable
// `pitched`: workspace for the even-pitch copies of dy and x when their rows are odd (f64_pitched_bytes each)
int f64_wgrad_tma(const Plan& p, const double* dy, const double* x, double* partial, void* pitched, int* splits,
                  cudaStream_t st) {
  EncodeTiledFn enc = tma_encoder();
  F64WgradPlan w;
  if (!enc || !f64_tma_enabled() || !f64_wgrad_tma_plan(p, &w)) return 2;
  if (((uintptr_t)dy & 15) || ((uintptr_t)x & 15)) return 2;
  const int gp2 = f64_pitch(p.G[2]), xp2 = f64_pitch(p.X[2]);
  if (p.G[2] & 1) {
    double* d2 = align1024(pitched);
    if (f64_pitch_copy(dy, d2, p.N * p.F * p.G[0] * p.G[1], p.G[2], st)) return 1;
    dy = d2;
    pitched = (char*)d2 + f64_pitched_bytes(p.N * p.F * p.G[0] * p.G[1], p.G[2]) - 1024;
  }
  if (p.X[2] & 1) {
    double* x2 = align1024(pitched);
    if (f64_pitch_copy(x, x2, p.N * p.C * p.X[0] * p.X[1], p.X[2], st)) return 1;
    x = x2;
  }
  CUtensorMap mdy, mx;
  {
    cuuint64_t dims[5] = {(cuuint64_t)p.G[2], (cuuint64_t)p.G[1], (cuuint64_t)p.G[0], (cuuint64_t)p.F, (cuuint64_t)p.N};
    const cuuint64_t gpp = (cuuint64_t)p.G[0] * p.G[1] * gp2;
    cuuint64_t strides[4] = {(cuuint64_t)gp2 * 8, (cuuint64_t)p.G[1] * gp2 * 8, gpp * 8, (cuuint64_t)p.F * gpp * 8};
    cuuint32_t bx[5] = {8, 2, 1, 128, 1}, estr[5] = {1, 1, 1, 1, 1};
    if (enc(&mdy, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, (void*)dy, dims, strides, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return 2;
  }
  {
    cuuint64_t dims[5] = {(cuuint64_t)p.X[2], (cuuint64_t)p.X[1], (cuuint64_t)p.X[0], (cuuint64_t)p.C, (cuuint64_t)p.N};
    const cuuint64_t xpp = (cuuint64_t)p.X[0] * p.X[1] * xp2;
    cuuint64_t strides[4] = {(cuuint64_t)xp2 * 8, (cuuint64_t)p.X[1] * xp2 * 8, xpp * 8, (cuuint64_t)p.C * xpp * 8};
    cuuint32_t bx[5] = {(cuuint32_t)f64t::wg_row_pixels(p.S[2]), (cuuint32_t)(2 * p.S[1]), 1, 32, 1};
    cuuint32_t estr[5] = {1, (cuuint32_t)p.S[1], 1, 1, 1};
    if (enc(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, (void*)x, dims, strides, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return 2;
  }
  f64t::WgradParams wp;
  wp.partial = partial;
  for (int i = 0; i < 3; ++i) { wp.G[i] = p.G[i]; wp.S[i] = p.S[i]; wp.Pd[i] = p.Pd[i]; wp.D[i] = p.D[i]; wp.W[i] = p.W[i]; }
  wp.F = (int)p.F; wp.C = (int)p.C; wp.Wp = (int)p.Wp; wp.cblocks = w.cblocks; wp.nblocks_total = w.nblocks;
  wp.tiles1 = w.tiles1; wp.tiles2 = w.tiles2;
  wp.boxes_total = (int)w.boxes; wp.boxes_per_split = (int)w.per;
  const dim3 grid(w.mt, w.nt, w.splits);
  auto launch = [&](auto kern, size_t smem) -> int {
    MGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, f64t::THREADS, smem, st>>>(wp, mdy, mx);
    return 0;
  };
  int lrc;
  if (w.NB == 3 && p.S[2] == 1)      lrc = launch(f64t::f64_wgrad_tma_kernel<3, 4, 1>, f64t::wgrad_smem_bytes<3, 4, 1>());
  else if (w.NB == 3)                lrc = launch(f64t::f64_wgrad_tma_kernel<3, 3, 2>, f64t::wgrad_smem_bytes<3, 3, 2>());
  else if (p.S[2] == 1)              lrc = launch(f64t::f64_wgrad_tma_kernel<4, 4, 1>, f64t::wgrad_smem_bytes<4, 4, 1>());
  else                               lrc = launch(f64t::f64_wgrad_tma_kernel<4, 3, 2>, f64t::wgrad_smem_bytes<4, 3, 2>());
  if (lrc) return lrc;
  MGB_LAUNCH_CHECK();
  *splits = w.splits;
  return 0;
}

// thin wgrad (conv_thin.cuh): C * taps <= 64.  Plan: KB taps per warp, KG x FG warps per block, slabs of pixels.
struct ThinWgradPlan { int KB, KG, FG, fgroups, splits; long long per; };
template <typename T>
bool thin_wgrad_plan(int M, int N, int K, ThinWgradPlan* out) {
  if (!thin_enabled() || N > 64 || M < 1 || K < 1) return false;
  ThinWgradPlan w;
  w.KB = sizeof(T) == 4 ? 16 : 8;
  w.KG = (N + w.KB - 1) / w.KB;                       // <= 4 (float) / 8 (double)
  w.FG = std::max(1, std::min((M + 7) / 8, 8 / w.KG));
  w.fgroups = (M + 8 * w.FG - 1) / (8 * w.FG);
  long long want = std::max<long long>(1, (3LL * sm_count()) / w.fgroups);
  w.splits = (int)std::max<long long>(1, std::min<long long>(want, ((long long)K + 1023) / 1024));
  w.per = (((long long)K + w.splits - 1) / w.splits + 31) / 32 * 32;
  w.splits = (int)(((long long)K + w.per - 1) / w.per);
  *out = w;
  return true;
}

template <typename T>
size_t wgrad_ws(const Plan& p, int* splits, int* kps) {
  int M = (int)p.F, N = (int)(p.C * p.Wp), K = (int)(p.N * p.Gp);
  wgrad_plan<T>(M, N, K, splits, kps);
  {
    ThinWgradPlan tw;
    if (thin_wgrad_plan<T>(M, N, K, &tw)) *splits = std::max(*splits, tw.splits);   // sizes the partial slabs below
  }
  size_t off = align_up((size_t)M * sizeof(TapEntry), 256);
  off += align_up((size_t)N * sizeof(TapEntry), 256);
  int slabs = *splits;
  F64WgradPlan w;
  if (sizeof(T) == 8 && f64_wgrad_tma_plan(p, &w)) slabs = std::max(slabs, w.splits);
  off += align_up((size_t)slabs * M * N * sizeof(T), 1024);
  if (sizeof(T) == 8)      // even-pitch copies of dy / x for the TMA kernel
    off += f64_pitched_bytes(p.N * p.F * p.G[0] * p.G[1], p.G[2]) + f64_pitched_bytes(p.N * p.C * p.X[0] * p.X[1], p.X[2]);
  return off;
}

template <typename T>
int wgrad_t(const Plan& p, const T* dy, const T* x, T* dw, void* ws, size_t ws_bytes,
            cudaStream_t st) {
  int M = (int)p.F, N = (int)(p.C * p.Wp), K = (int)(p.N * p.Gp);
  if ((long long)M * N == 0) return 0;
  if (K == 0) {   // an empty shard still takes part in the cross-rank sum
    MGB_CUDA(cudaMemsetAsync(dw, 0, (size_t)M * N * sizeof(T), st));
    return (g_wgrad_dp && peer_ready()) ? finish_dw<T>(dw, dw, (long long)M * N, 1, st) : 0;
  }
  int splits, kps;
  size_t need = wgrad_ws<T>(p, &splits, &kps);
  if (ws_bytes < need || !ws) return fail("conv wgrad: workspace %zu < %zu bytes", ws_bytes, need);
  TapEntry* tabA = (TapEntry*)ws;
  TapEntry* tabB = (TapEntry*)((char*)ws + align_up((size_t)M * sizeof(TapEntry), 256));
  T* partial = (T*)((char*)tabB + align_up((size_t)N * sizeof(TapEntry), 256));
  if constexpr (sizeof(T) == 8) {        // TMA-staged DMMA kernel first
    int s2 = 0;
    const size_t pitched_bytes = f64_pitched_bytes(p.N * p.F * p.G[0] * p.G[1], p.G[2]) +
                                 f64_pitched_bytes(p.N * p.C * p.X[0] * p.X[1], p.X[2]);
    const int rc = f64_wgrad_tma(p, dy, x, partial, (char*)ws + need - pitched_bytes, &s2, st);
    if (rc == 1) return 1;
    if (rc == 0) return finish_dw<T>(partial, dw, (long long)M * N, s2, st);
  }
  {
    ThinWgradPlan tw;
    Thin2dPlan t2;
    if (thin_wgrad_plan<T>(M, N, K, &tw) && thin2d_plan(p, sizeof(T), 24 * 1024, &t2)) {
      constexpr int KB = sizeof(T) == 4 ? 16 : 8;
      thin::Thin2dParams<T> q = thin2d_params<T>(p, t2);
      q.x = x; q.dy = dy; q.out = partial;
      if constexpr (sizeof(T) == 4) {      // float32: mma.sync 3xTF32, one slab of partials per block
        const char* em = getenv("MGB_THIN_MMA");
        if (!(em && em[0] == '0')) {
          const int fb = std::min(M, 64);
          const int mt = fb <= 16 ? 1 : fb <= 32 ? 2 : 4, nt = N <= 32 ? 4 : 8;
          const int gy = (M + mt * 16 - 1) / (mt * 16);
          const int bx = (int)std::min<long long>(p.N * t2.nbands, std::min(splits, std::max(1, 3 * sm_count() / gy)));
          const size_t sm = std::max<size_t>(2 * t2.band_bytes, (size_t)mt * 16 * nt * 8 * 4);
          const dim3 grid(bx, gy);
          if (mt == 1 && nt == 4) thin::thin2d_wgrad_mma_kernel<1, 4><<<grid, 256, sm, st>>>(q);
          else if (mt == 2 && nt == 4) thin::thin2d_wgrad_mma_kernel<2, 4><<<grid, 256, sm, st>>>(q);
          else if (mt == 4 && nt == 4) thin::thin2d_wgrad_mma_kernel<4, 4><<<grid, 256, sm, st>>>(q);
          else if (mt == 1) thin::thin2d_wgrad_mma_kernel<1, 8><<<grid, 256, sm, st>>>(q);
          else if (mt == 2) thin::thin2d_wgrad_mma_kernel<2, 8><<<grid, 256, sm, st>>>(q);
          else thin::thin2d_wgrad_mma_kernel<4, 8><<<grid, 256, sm, st>>>(q);
          MGB_LAUNCH_CHECK();
          return finish_dw<T>(partial, dw, (long long)M * N, bx, st);
        }
      }
      const int blocks = (int)std::min<long long>(p.N * t2.nbands, std::min(splits, std::max(1, sm_count() / tw.fgroups)));
      thin::thin2d_wgrad_kernel<T, KB><<<dim3(blocks, tw.fgroups), tw.FG * tw.KG * 32, 2 * t2.band_bytes, st>>>(q, tw.FG, tw.KG);
      MGB_LAUNCH_CHECK();
      return finish_dw<T>(partial, dw, (long long)M * N, blocks, st);
    }
  }
  build_chan_table<<<small_grid(M), 256, 0, st>>>(tabA, M, p.Gp);
  MGB_LAUNCH_CHECK();
  build_x_table<<<small_grid(N), 256, 0, st>>>(tabB, x_table_args(p));
  MGB_LAUNCH_CHECK();
  WgradParams<T> wp;
  wp.srcA = dy; wp.tabA = tabA; wp.srcB = x; wp.tabB = tabB; wp.partial = partial;
  wp.ga = geom_dy_by_output(p);
  wp.gb = geom_x_by_output(p);
  wp.M = M; wp.N = N; wp.K = K; wp.k_per_split = kps;
  {
    ThinWgradPlan tw;
    if (thin_wgrad_plan<T>(M, N, K, &tw)) {
      constexpr int KB = sizeof(T) == 4 ? 16 : 8;
      const size_t smem = (size_t)N * sizeof(TapEntry);
      thin::thin_wgrad_kernel<T, KB><<<dim3(tw.splits, tw.fgroups), tw.FG * tw.KG * 32, smem, st>>>(wp, tw.FG, tw.KG, tw.per);
      MGB_LAUNCH_CHECK();
      return finish_dw<T>(partial, dw, (long long)M * N, tw.splits, st);
    }
  }
  int rc;
  switch (wgrad_family(M, N)) {
    case 0: rc = launch_wgrad<T, typename Tiles<T>::WBig>(wp, splits, st); break;
    case 3: rc = launch_wgrad<T, typename Tiles<T>::WBig96>(wp, splits, st); break;
    case 1: rc = launch_wgrad<T, typename Tiles<T>::WMid>(wp, splits, st); break;
    default: rc = launch_wgrad<T, typename Tiles<T>::WSmall>(wp, splits, st); break;
  }
  if (rc) return rc;
  return finish_dw<T>(partial, dw, (long long)M * N, splits, st);
}

}  // namespace
}  // namespace conv
}  // namespace mgb

using namespace mgb;
using namespace mgb::conv;

extern "C" int mgb_conv_workspace_bytes(const mgb_conv_desc* d, int which, size_t* bytes) {
  Plan p;
  if (make_plan(d, &p)) return 1;
  if (!bytes) return fail("mgb_conv_workspace_bytes: null");
  size_t esz = p.dtype == MGB_F32 ? 4 : 8;
  size_t need = 0;
  if (which == 0) {
    need = tc_eligible(p, 0) ? tc_fwd_ws(p) : fwd_ws(p);
  } else if (which == 1) {
    std::vector<DgradClassPlan> cls;
    dgrad_classes(p, esz, &cls, &need);
    if (p.dtype == MGB_F64) need += f64_pitched_bytes(p.N * p.F * p.G[0] * p.G[1], p.G[2]);
    if (tc_eligible(p, 1)) need = tc_dgrad_ws(p, cls);
  } else if (which == 2) {
    int s, k;
    if (p.F * p.C * p.Wp == 0 || p.N * p.Gp == 0) need = 0;
    else if (tc_eligible(p, 2)) need = tc_wgrad_layout(p).end;
    else need = p.dtype == MGB_F32 ? wgrad_ws<float>(p, &s, &k) : wgrad_ws<double>(p, &s, &k);
  } else {
    return fail("mgb_conv_workspace_bytes: which=%d", which);
  }
  *bytes = need;
  return 0;
}

extern "C" int mgb_conv_planes_bytes(const mgb_conv_desc* d, int tensor, size_t* bytes) {
  Plan p;
  if (make_plan(d, &p)) return 1;
  if (!bytes || (tensor != 0 && tensor != 1)) return fail("mgb_conv_planes_bytes: bad arguments");
  // x planes feed fwd and wgrad, dy planes feed dgrad and wgrad
  bool used = tensor == 0 ? (tc_eligible(p, 0) || tc_eligible(p, 2)) : (tc_eligible(p, 1) || tc_eligible(p, 2));
  *bytes = !used ? 0 : (tensor == 0 ? tc_planes(0, p.N, p.C, p.Xp).end : tc_planes(0, p.N, p.F, p.Gp).end);
  return 0;
}

extern "C" int mgb_conv_make_planes(const mgb_conv_desc* d, int tensor, const void* src, void* planes,
                                    void* stream) {
  Plan p;
  if (make_plan(d, &p)) return 1;
  if (p.dtype != MGB_F32 || (tensor != 0 && tensor != 1)) return fail("mgb_conv_make_planes: float32 x (0) or dy (1) only");
  long long ch = tensor == 0 ? p.C : p.F, sp = tensor == 0 ? p.Xp : p.Gp;
  TcPlanes P = tc_planes(0, p.N, ch, sp);
  if (p.N * sp == 0) return 0;
  if (!src || !planes) return fail("mgb_conv_make_planes: null buffer");
  return tc_prep_planes((const float*)src, (float*)((char*)planes + P.hi_off),
                        (float*)((char*)planes + P.lo_off), p.N, ch, sp, P.cpad, as_stream(stream));
}

extern "C" int mgb_conv_fwd_p(const mgb_conv_desc* d, const void* x, const void* w, void* y,
                              const void* x_planes, void* ws, size_t ws_bytes, void* stream) {
  Plan p;
  if (make_plan(d, &p)) return 1;
  if (tc_eligible(p, 0))
    return tc_fwd(p, (const float*)x, (const float*)w, (float*)y, x_planes, ws, ws_bytes, as_stream(stream));
  return mgb_conv_fwd(d, x, w, y, ws, ws_bytes, stream);
}

extern "C" int mgb_conv_dgrad_p(const mgb_conv_desc* d, const void* dy, const void* w, void* dx,
                                const void* dy_planes, void* ws, size_t ws_bytes, void* stream) {
  Plan p;
  if (make_plan(d, &p)) return 1;
  if (tc_eligible(p, 1))
    return tc_dgrad(p, (const float*)dy, (const float*)w, (float*)dx, dy_planes, ws, ws_bytes, as_stream(stream));
  return mgb_conv_dgrad(d, dy, w, dx, ws, ws_bytes, stream);
}

extern "C" int mgb_conv_wgrad_p(const mgb_conv_desc* d, const void* dy, const void* x, void* dw,
                                const void* dy_planes, const void* x_planes, void* ws, size_t ws_bytes,
                                void* stream) {
  Plan p;
  if (make_plan(d, &p)) return 1;
  if (tc_eligible(p, 2))
    return tc_wgrad(p, (const float*)dy, (const float*)x, (float*)dw, dy_planes, x_planes, ws, ws_bytes,
                    as_stream(stream));
  return mgb_conv_wgrad(d, dy, x, dw, ws, ws_bytes, stream);
}

extern "C" int mgb_conv_wgrad_dp(const mgb_conv_desc* d, const void* dy, const void* x, void* dw,
                                 const void* dy_planes, const void* x_planes, void* ws, size_t ws_bytes,
                                 void* stream) {
  g_wgrad_dp = true;
  int rc = mgb_conv_wgrad_p(d, dy, x, dw, dy_planes, x_planes, ws, ws_bytes, stream);
  g_wgrad_dp = false;
  return rc;
}

extern "C" int mgb_conv_fwd(const mgb_conv_desc* d, const void* x, const void* w, void* y,
                            void* ws, size_t ws_bytes, void* stream) {
  Plan p;
  if (make_plan(d, &p)) return 1;
  if (tc_eligible(p, 0))
    return tc_fwd(p, (const float*)x, (const float*)w, (float*)y, nullptr, ws, ws_bytes, as_stream(stream));
  if (p.dtype == MGB_F32)
    return fwd_t<float>(p, (const float*)x, (const float*)w, (float*)y, ws, ws_bytes, as_stream(stream));
  return fwd_t<double>(p, (const double*)x, (const double*)w, (double*)y, ws, ws_bytes, as_stream(stream));
}

extern "C" int mgb_conv_dgrad(const mgb_conv_desc* d, const void* dy, const void* w, void* dx,
                              void* ws, size_t ws_bytes, void* stream) {
  Plan p;
  if (make_plan(d, &p)) return 1;
  if (tc_eligible(p, 1))
    return tc_dgrad(p, (const float*)dy, (const float*)w, (float*)dx, nullptr, ws, ws_bytes, as_stream(stream));
  if (p.dtype == MGB_F32)
    return dgrad_t<float>(p, (const float*)dy, (const float*)w, (float*)dx, ws, ws_bytes, as_stream(stream));
  return dgrad_t<double>(p, (const double*)dy, (const double*)w, (double*)dx, ws, ws_bytes, as_stream(stream));
}

extern "C" int mgb_conv_wgrad(const mgb_conv_desc* d, const void* dy, const void* x, void* dw,
                              void* ws, size_t ws_bytes, void* stream) {
  Plan p;
  if (make_plan(d, &p)) return 1;
  if (tc_eligible(p, 2))
    return tc_wgrad(p, (const float*)dy, (const float*)x, (float*)dw, nullptr, nullptr, ws, ws_bytes, as_stream(stream));
  if (p.dtype == MGB_F32)
    return wgrad_t<float>(p, (const float*)dy, (const float*)x, (float*)dw, ws, ws_bytes, as_stream(stream));
  return wgrad_t<double>(p, (const double*)dy, (const double*)x, (double*)dw, ws, ws_bytes, as_stream(stream));
}

// ---- conv -> max-pool fusion: C ABI ---------------------------------------------------------------

extern "C" int mgb_conv_pool_supported(const mgb_conv_desc* d, const mgb_pool_desc* pd, int* yes) {
  Plan p;
  if (make_plan(d, &p)) return 1;
  if (!yes) return fail("mgb_conv_pool_supported: null");
  *yes = conv_pool_applies(p, pd) ? 1 : 0;
  return 0;
}

extern "C" int mgb_conv_pool_fwd(const mgb_conv_desc* d, const mgb_pool_desc* pd, const void* x, const void* w,
                                 void* pooled, void* argmax, const void* x_planes, void* ws, size_t ws_bytes,
                                 void* stream) {
  Plan p;
  if (make_plan(d, &p)) return 1;
  if (!conv_pool_applies(p, pd))
    return fail("mgb_conv_pool_fwd: geometry not supported by the fused kernel (ask mgb_conv_pool_supported)");
  if (!pooled || !argmax) return fail("mgb_conv_pool_fwd: null output");
  return tc_fwd(p, (const float*)x, (const float*)w, nullptr, x_planes, ws, ws_bytes, as_stream(stream),
                (float*)pooled, (unsigned char*)argmax);
}

extern "C" int mgb_pool_bwd_planes(const mgb_conv_desc* d, const mgb_pool_desc* pd, const void* dp,
                                   const void* argmax, void* dy_planes, void* stream) {
  Plan p;
  if (make_plan(d, &p)) return 1;
  if (!conv_pool_applies(p, pd)) return fail("mgb_pool_bwd_planes: geometry not supported by the fused kernels");
  if (!dp || !argmax || !dy_planes) return fail("mgb_pool_bwd_planes: null buffer");
  if (p.N > 65535) return fail("mgb_pool_bwd_planes: batch %lld > 65535 unsupported", p.N);
  TcPlanes P = tc_planes(0, p.N, p.F, p.Gp);
  const int PH = p.G[1] / 2, PW = p.G[2] / 2;
  dim3 grid((unsigned)((PH * PW + 63) / 64), (unsigned)(P.cpad / 32), (unsigned)p.N);
  pool_bwd_planes_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(
      (const float*)dp, (const unsigned char*)argmax, nullptr, nullptr, (float*)((char*)dy_planes + P.hi_off),
      (float*)((char*)dy_planes + P.lo_off), (int)p.F, PH, PW, P.cpad);
  MGB_LAUNCH_CHECK();
  return 0;
}

// max_pool backward of a conv_nd output, for the UNFUSED pair: dy (the public y.grad) and the channels-last
// planes of dy that the convolution's dX / dW read, in one pass over y and dp (instead of mgb_pool_bwd followed by
// mgb_conv_make_planes, which re-reads dy).  Same geometry as the fused pair; 8-byte alignment of y / dy rows
// (even conv output width) is part of that geometry.
extern "C" int mgb_pool_bwd_dy_planes(const mgb_conv_desc* d, const mgb_pool_desc* pd, const void* dp, const void* y,
                                      void* dy, void* dy_planes, void* stream) {
  Plan p;
  if (make_plan(d, &p)) return 1;
  if (!conv_pool_applies(p, pd)) return fail("mgb_pool_bwd_dy_planes: geometry not supported by the fused kernels");
  if (!dp || !y || !dy || !dy_planes) return fail("mgb_pool_bwd_dy_planes: null buffer");
  if (((uintptr_t)y | (uintptr_t)dy) & 7) return fail("mgb_pool_bwd_dy_planes: y and dy must be 8-byte aligned");
  if (p.N > 65535) return fail("mgb_pool_bwd_dy_planes: batch %lld > 65535 unsupported", p.N);
  TcPlanes P = tc_planes(0, p.N, p.F, p.Gp);
  const int PH = p.G[1] / 2, PW = p.G[2] / 2;
  dim3 grid((unsigned)((PH * PW + 63) / 64), (unsigned)(P.cpad / 32), (unsigned)p.N);
  pool_bwd_planes_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(
      (const float*)dp, nullptr, (const float*)y, (float*)dy, (float*)((char*)dy_planes + P.hi_off),
      (float*)((char*)dy_planes + P.lo_off), (int)p.F, PH, PW, P.cpad);
  MGB_LAUNCH_CHECK();
  return 0;
}
