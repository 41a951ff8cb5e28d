// Batch normalisation over every axis but axis 1 — HBM-bound.
//
// Reference: BatchNorm.__call__ / backward_var, src/mygrad/nnet/layers/batchnorm.py:26-111:
//   mean/var over the non-channel axes (population variance), y = (x - mean)/sqrt(var + eps),
//   optional per-channel gamma, beta;
//   dx = gamma/std * (g - mean(g) - x_norm * mean(g * x_norm)), dgamma = sum(g * x_norm),
//   dbeta = sum(g).
//
// x is (N, C, S) with S = product of the trailing axes: channel c owns N runs of S contiguous
// elements.  All per-channel moments are accumulated in float64 and merged in a FIXED order
// (deterministic): a single pass over x with shifted-data sums per thread, converted to
// (count, mean, M2) triples that the warp / block / split levels merge with Chan's update.
// x_norm is never stored: the backward recomputes it from x and the saved moments, so the
// forward moves 2 reads + 1 write of x and the backward 4 reads + 1 write.
#include <algorithm>

#include "common.cuh"

namespace mgb {
namespace {

constexpr int kThreads = 256;

struct Moments { double n, mean, m2; };

__device__ __forceinline__ Moments merge(const Moments& a, const Moments& b) {
  if (b.n == 0.0) return a;
  if (a.n == 0.0) return b;
  Moments r;
  r.n = a.n + b.n;
  const double d = b.mean - a.mean;
  r.mean = a.mean + d * (b.n / r.n);
  r.m2 = a.m2 + b.m2 + d * d * (a.n * b.n / r.n);
  return r;
}

__device__ __forceinline__ Moments warp_merge(Moments m) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    Moments o;
    o.n = __shfl_down_sync(0xffffffffu, m.n, off);
    o.mean = __shfl_down_sync(0xffffffffu, m.mean, off);
    o.m2 = __shfl_down_sync(0xffffffffu, m.m2, off);
    m = merge(m, o);     // lane i merges lane i+off: a fixed tree
  }
  return m;
}

template <typename T, int V> struct Vec;
template <> struct Vec<float, 4> { using type = float4; };
template <> struct Vec<double, 2> { using type = double2; };
template <> struct Vec<float, 1> { using type = float; };
template <> struct Vec<double, 1> { using type = double; };

// sums of one 16-byte pack: formed in T (a handful of terms: the rounding is far below the
// tolerance) and added to the float64 accumulators once per pack — float32 data would otherwise be
// bound by FP64 conversions and adds, not by HBM
template <typename T, int V>
__device__ __forceinline__ void pack_moments(const T* e, T K, double& a1, double& a2) {
  T p1 = (T)0, p2 = (T)0;
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const T d = e[k] - K;
    p1 += d;
    p2 = fma(d, d, p2);
  }
  a1 += (double)p1;
  a2 += (double)p2;
}

template <typename T, int V>
__device__ __forceinline__ void pack_gsums(const T* ge, const T* xe, T mean_t, T inv_t, double& ag, double& agx) {
  T p1 = (T)0, p2 = (T)0;
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const T xn = (xe[k] - mean_t) * inv_t;
    p1 += ge[k];
    p2 = fma(ge[k], xn, p2);
  }
  ag += (double)p1;
  agx += (double)p2;
}

// grid (C, splits): CTA (c, s) covers samples n = s, s + splits, ... of channel c.  Threads sum
// (v - K) and (v - K)^2 in float64 with K = the channel's first element (shifted-data variance:
// three flops per element, no cancellation as long as K lies within the data), convert to a
// Welford triple once, and the triples are merged with Chan's update.
template <typename T, int V>
__global__ void __launch_bounds__(kThreads) bn_stats_kernel(const T* __restrict__ x, long long N, long long C,
                                                            long long S, int splits, double* __restrict__ part) {
  using VT = typename Vec<T, V>::type;
  const long long c = blockIdx.x;
  const T Kt = x[c * S];
  const double K = (double)Kt;
  double cnt = 0.0, s1 = 0.0, s2 = 0.0;
  const long long SV = S / V;
  double a1[4] = {0.0, 0.0, 0.0, 0.0}, a2[4] = {0.0, 0.0, 0.0, 0.0};
  // one WARP per run (sample), eight 16-byte loads in flight per lane: a run of a few thousand
  // elements keeps all 32 lanes busy, and 8 x 512 B contiguous per warp and iteration is what it takes
  // to cover the HBM latency with ~40 resident warps per SM
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int U = 8, WARPS = kThreads / 32;
  for (long long n = blockIdx.y + (long long)warp * splits; n < N; n += (long long)WARPS * splits) {
    const VT* run = reinterpret_cast<const VT*>(x + (n * C + c) * S);
    long long i = lane;
    for (; i + (U - 1) * 32 < SV; i += U * 32) {
      VT pack[U];
#pragma unroll
      for (int u = 0; u < U; ++u) pack[u] = run[i + u * 32];
#pragma unroll
      for (int u = 0; u < U; ++u) pack_moments<T, V>(reinterpret_cast<const T*>(&pack[u]), Kt, a1[u & 3], a2[u & 3]);
      cnt += (double)(U * V);
    }
    for (; i < SV; i += 32) {
      const VT pack = run[i];
      pack_moments<T, V>(reinterpret_cast<const T*>(&pack), Kt, a1[0], a2[0]);
      cnt += (double)V;
    }
  }
  s1 = (a1[0] + a1[1]) + (a1[2] + a1[3]);
  s2 = (a2[0] + a2[1]) + (a2[2] + a2[3]);
  Moments m{cnt, 0.0, 0.0};
  if (cnt > 0.0) {
    m.mean = K + s1 / cnt;
    m.m2 = fmax(0.0, s2 - s1 * s1 / cnt);
  }
  __shared__ Moments sh[kThreads / 32];
  m = warp_merge(m);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    Moments t = sh[0];
    for (int w = 1; w < kThreads / 32; ++w) t = merge(t, sh[w]);
    double* o = part + ((long long)blockIdx.y * C + c) * 3;
    o[0] = t.n; o[1] = t.mean; o[2] = t.m2;
  }
}

// moments[0..C) = mean, moments[C..2C) = population variance, moments[2C..3C) = count
__global__ void bn_stats_finish(const double* __restrict__ part, long long C, int splits,
                                double* __restrict__ moments) {
  const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (c >= C) return;
  Moments t{0.0, 0.0, 0.0};
  for (int s = 0; s < splits; ++s) {
    const double* p = part + ((long long)s * C + c) * 3;
    t = merge(t, Moments{p[0], p[1], p[2]});
  }
  moments[c] = t.mean;
  moments[C + c] = t.n > 0.0 ? t.m2 / t.n : 0.0;
  moments[2 * C + c] = t.n;
}

// y = (x - mean) / sqrt(var + eps) [* gamma] [+ beta] in x's dtype and the reference's operation
// order (batchnorm.py:59-73); grid (runs = N*C, chunks of S)
template <typename T, int V>
__global__ void __launch_bounds__(kThreads) bn_apply_kernel(const T* __restrict__ x, long long C, long long S,
                                                            const double* __restrict__ moments,
                                                            const T* __restrict__ gamma, const T* __restrict__ beta,
                                                            double eps, T* __restrict__ y) {
  using VT = typename Vec<T, V>::type;
  const long long run = blockIdx.x;
  const long long c = run % C;
  // (x - mean) * (1/std): one rounding away from the reference's division, far inside the tolerance
  const T mean_t = (T)moments[c], inv_t = (T)(1.0 / sqrt(moments[C + c] + eps));
  const bool has_g = gamma != nullptr, has_b = beta != nullptr;
  const T g = has_g ? gamma[c] : (T)1, b = has_b ? beta[c] : (T)0;
  const VT* xr = reinterpret_cast<const VT*>(x + run * S);
  VT* yr = reinterpret_cast<VT*>(y + run * S);
  const long long SV = S / V;
  const long long step = (long long)gridDim.y * kThreads;
  auto one = [&](VT pack) {
    T* e = reinterpret_cast<T*>(&pack);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      T v = (e[k] - mean_t) * inv_t;
      if (has_g) v = v * g;
      if (has_b) v = v + b;
      e[k] = v;
    }
    return pack;
  };
  long long i = (long long)blockIdx.y * kThreads + threadIdx.x;
  for (; i + 3 * step < SV; i += 4 * step) {
    VT pack[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) pack[u] = xr[i + u * step];
#pragma unroll
    for (int u = 0; u < 4; ++u) yr[i + u * step] = one(pack[u]);
  }
  for (; i < SV; i += step) yr[i] = one(xr[i]);
}

// per-channel sums of g and g * x_norm; grid (C, splits); part[(s*C + c)*2 + {0,1}]
template <typename T, int V>
__global__ void __launch_bounds__(kThreads) bn_bwd_reduce_kernel(const T* __restrict__ g, const T* __restrict__ x,
                                                                 long long N, long long C, long long S, int splits,
                                                                 const double* __restrict__ moments, double eps,
                                                                 double* __restrict__ part) {
  using VT = typename Vec<T, V>::type;
  const long long c = blockIdx.x;
  const T mean_t = (T)moments[c], inv_t = (T)(1.0 / sqrt(moments[C + c] + eps));
  double ag[2] = {0.0, 0.0}, agx[2] = {0.0, 0.0};
  const long long SV = S / V;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int U = 4, WARPS = kThreads / 32;               // one warp per run, 2 x 4 loads in flight per lane
  for (long long n = blockIdx.y + (long long)warp * splits; n < N; n += (long long)WARPS * splits) {
    const long long off = (n * C + c) * S;
    const VT* gr = reinterpret_cast<const VT*>(g + off);
    const VT* xr = reinterpret_cast<const VT*>(x + off);
    long long i = lane;
    for (; i + (U - 1) * 32 < SV; i += U * 32) {
      VT gp[U], xp[U];
#pragma unroll
      for (int u = 0; u < U; ++u) { gp[u] = gr[i + u * 32]; xp[u] = xr[i + u * 32]; }
#pragma unroll
      for (int u = 0; u < U; ++u)
        pack_gsums<T, V>(reinterpret_cast<const T*>(&gp[u]), reinterpret_cast<const T*>(&xp[u]), mean_t, inv_t,
                         ag[u & 1], agx[u & 1]);
    }
    for (; i < SV; i += 32) {
      const VT gp = gr[i], xp = xr[i];
      pack_gsums<T, V>(reinterpret_cast<const T*>(&gp), reinterpret_cast<const T*>(&xp), mean_t, inv_t, ag[0],
                       agx[0]);
    }
  }
  double sg = ag[0] + ag[1], sgx = agx[0] + agx[1];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sg += __shfl_down_sync(0xffffffffu, sg, o);
    sgx += __shfl_down_sync(0xffffffffu, sgx, o);
  }
  __shared__ double sh[2][kThreads / 32];
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = sg; sh[1][threadIdx.x >> 5] = sgx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) { a += sh[0][w]; b += sh[1][w]; }
    double* o = part + ((long long)blockIdx.y * C + c) * 2;
    o[0] = a; o[1] = b;
  }
}

// sums[0..C) = sum g, sums[C..2C) = sum g*x_norm
__global__ void bn_bwd_finish(const double* __restrict__ part, long long C, int splits, double* __restrict__ sums) {
  const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (c >= C) return;
  double a = 0.0, b = 0.0;
  for (int s = 0; s < splits; ++s) {
    a += part[((long long)s * C + c) * 2];
    b += part[((long long)s * C + c) * 2 + 1];
  }
  sums[c] = a;
  sums[C + c] = b;
}

// dx = (g - mean_g - x_norm * sum_gx / n) / std [* gamma]   (batchnorm.py:82-99, same operation order)
template <typename T, int V>
__global__ void __launch_bounds__(kThreads) bn_bwd_dx_kernel(const T* __restrict__ g, const T* __restrict__ x,
                                                             long long C, long long S,
                                                             const double* __restrict__ moments,
                                                             const double* __restrict__ sums,
                                                             const T* __restrict__ gamma, double eps,
                                                             T* __restrict__ dx) {
  using VT = typename Vec<T, V>::type;
  const long long run = blockIdx.x;
  const long long c = run % C;
  const double cnt = moments[2 * C + c];
  const double inv = 1.0 / sqrt(moments[C + c] + eps);
  const T mean_t = (T)moments[c], inv_t = (T)inv;
  const T mean_g = (T)(sums[c] / cnt);
  const T k_gx = (T)(sums[C + c] / cnt);                   // mean(g * x_norm)
  const bool has_g = gamma != nullptr;
  const T scale = (T)(has_g ? inv * (double)gamma[c] : inv);   // gamma / std
  const VT* gr = reinterpret_cast<const VT*>(g + run * S);
  const VT* xr = reinterpret_cast<const VT*>(x + run * S);
  VT* dr = reinterpret_cast<VT*>(dx + run * S);
  const long long SV = S / V;
  const long long step = (long long)gridDim.y * kThreads;
  auto one = [&](const VT& gp, const VT& xp) {
    const T* ge = reinterpret_cast<const T*>(&gp);
    const T* xe = reinterpret_cast<const T*>(&xp);
    VT out;
    T* oe = reinterpret_cast<T*>(&out);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const T xn = (xe[k] - mean_t) * inv_t;
      oe[k] = ((ge[k] - mean_g) - xn * k_gx) * scale;
    }
    return out;
  };
  long long i = (long long)blockIdx.y * kThreads + threadIdx.x;
  for (; i + step < SV; i += 2 * step) {                    // four loads in flight
    const VT g0 = gr[i], g1 = gr[i + step], x0 = xr[i], x1 = xr[i + step];
    dr[i] = one(g0, x0);
    dr[i + step] = one(g1, x1);
  }
  for (; i < SV; i += step) dr[i] = one(gr[i], xr[i]);
}

template <typename T> constexpr int kVec = 16 / (int)sizeof(T);

inline bool vec_ok(long long S, int V, const void* a, const void* b = nullptr, const void* c = nullptr) {
  auto al = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return S % V == 0 && al(a) && al(b) && al(c);
}

// cross-rank merge of the per-rank moments (data-parallel batch norm), three tiny phases around
// two sum all-reduces:  0: buf = (n, n*mean)      -> all-reduce 2C
//                       1: moments.mean/count = global; buf = M2 + n*(mean - gmean)^2  -> all-reduce C
//                       2: moments.var = buf / count
__global__ void bn_merge_kernel(int phase, long long C, double* __restrict__ moments, double* __restrict__ buf) {
  const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (phase == 0) {
    const double n = moments[2 * C + c];
    buf[c] = n;
    buf[C + c] = n * moments[c];
  } else if (phase == 1) {
    const double n = moments[2 * C + c], mean = moments[c], var = moments[C + c];
    const double gn = buf[c];
    const double gmean = gn > 0.0 ? buf[C + c] / gn : 0.0;
    const double d = mean - gmean;
    buf[2 * C + c] = var * n + n * d * d;
    moments[c] = gmean;
    moments[2 * C + c] = gn;
  } else {
    const double gn = moments[2 * C + c];
    moments[C + c] = gn > 0.0 ? buf[2 * C + c] / gn : 0.0;
  }
}

int pick_splits(long long N, long long C) {
  // about four CTAs per SM in total, never more splits than samples; among the nearby counts take the
  // one whose runs-per-CTA fills the eight warps of a CTA most evenly (one warp owns whole runs)
  const long long want = std::max(1LL, std::min(std::min((4LL * sm_count() + C - 1) / C, N), 1024LL));
  long long best = want;
  double best_waste = 1e30;
  for (long long s = std::max(1LL, want / 2); s <= std::min(std::min(2 * want, N), 1024LL); ++s) {
    const long long runs = (N + s - 1) / s;                       // runs of the fullest CTA
    const long long rounds = (runs + 7) / 8;
    const double waste = (double)(rounds * 8 * s) / (double)N     // idle warp slots
                         * (s * C < 2LL * sm_count() ? 4.0 : 1.0)  // too few CTAs to fill the machine
                         + 0.02 * (double)std::llabs(s - want) / (double)want;
    if (waste < best_waste) { best_waste = waste; best = s; }
  }
  return (int)best;
}

dim3 apply_grid(long long runs, long long SV) {
  // one CTA per run unless the run is long: every thread then keeps four 16-byte loads in flight
  long long chunks = std::max(1LL, std::min(SV / (8LL * kThreads), 64LL));
  return dim3((unsigned)runs, (unsigned)chunks);
}

template <typename T>
int stats_t(long long N, long long C, long long S, const T* x, double* moments, double* ws, cudaStream_t st) {
  const int splits = pick_splits(N, C);
  dim3 grid((unsigned)C, (unsigned)splits);
  if (vec_ok(S, kVec<T>, x))
    bn_stats_kernel<T, kVec<T>><<<grid, kThreads, 0, st>>>(x, N, C, S, splits, ws);
  else
    bn_stats_kernel<T, 1><<<grid, kThreads, 0, st>>>(x, N, C, S, splits, ws);
  MGB_LAUNCH_CHECK();
  bn_stats_finish<<<(unsigned)((C + 127) / 128), 128, 0, st>>>(ws, C, splits, moments);
  MGB_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int apply_t(long long N, long long C, long long S, const T* x, const double* moments, const T* gamma, const T* beta,
            double eps, T* y, cudaStream_t st) {
  if (vec_ok(S, kVec<T>, x, y))
    bn_apply_kernel<T, kVec<T>><<<apply_grid(N * C, S / kVec<T>), kThreads, 0, st>>>(x, C, S, moments, gamma, beta,
                                                                                    eps, y);
  else
    bn_apply_kernel<T, 1><<<apply_grid(N * C, S), kThreads, 0, st>>>(x, C, S, moments, gamma, beta, eps, y);
  MGB_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int bwd_reduce_t(long long N, long long C, long long S, const T* g, const T* x, const double* moments, double eps,
                 double* sums, double* ws, cudaStream_t st) {
  const int splits = pick_splits(N, C);
  dim3 grid((unsigned)C, (unsigned)splits);
  if (vec_ok(S, kVec<T>, g, x))
    bn_bwd_reduce_kernel<T, kVec<T>><<<grid, kThreads, 0, st>>>(g, x, N, C, S, splits, moments, eps, ws);
  else
    bn_bwd_reduce_kernel<T, 1><<<grid, kThreads, 0, st>>>(g, x, N, C, S, splits, moments, eps, ws);
  MGB_LAUNCH_CHECK();
  bn_bwd_finish<<<(unsigned)((C + 127) / 128), 128, 0, st>>>(ws, C, splits, sums);
  MGB_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int bwd_dx_t(long long N, long long C, long long S, const T* g, const T* x, const double* moments,
             const double* sums, const T* gamma, double eps, T* dx, cudaStream_t st) {
  if (vec_ok(S, kVec<T>, g, x, dx))
    bn_bwd_dx_kernel<T, kVec<T>><<<apply_grid(N * C, S / kVec<T>), kThreads, 0, st>>>(g, x, C, S, moments, sums,
                                                                                     gamma, eps, dx);
  else
    bn_bwd_dx_kernel<T, 1><<<apply_grid(N * C, S), kThreads, 0, st>>>(g, x, C, S, moments, sums, gamma, eps, dx);
  MGB_LAUNCH_CHECK();
  return 0;
}

int check_shape(long long N, long long C, long long S, const char* what) {
  if (N < 0 || C < 0 || S < 0) return fail("%s: negative extent", what);
  if (C > 2147483647LL || N * C > 2147483647LL) return fail("%s: N*C = %lld exceeds the grid limit", what, N * C);
  return 0;
}

}  // namespace
}  // namespace mgb

using namespace mgb;

extern "C" {

int mgb_batchnorm_workspace_bytes(int64_t n, int64_t c, size_t* bytes) {
  if (!bytes) return fail("mgb_batchnorm_workspace_bytes: null output");
  *bytes = (size_t)std::max<int64_t>(1, c) * 1024 * 3 * sizeof(double);
  (void)n;
  return 0;
}

int mgb_batchnorm_stats(int dtype, int64_t n, int64_t c, int64_t s, const void* x, void* moments, void* ws,
                        size_t ws_bytes, void* stream) {
  if (check_shape(n, c, s, "mgb_batchnorm_stats")) return 1;
  if (c == 0) return 0;
  size_t need = 0;
  mgb_batchnorm_workspace_bytes(n, c, &need);
  if (!ws || ws_bytes < need) return fail("mgb_batchnorm_stats: workspace %zu < %zu bytes", ws_bytes, need);
  cudaStream_t st = as_stream(stream);
  if (n * s == 0) {   // empty reduction: mean/var of nothing; the host layer rejects this before calling
    MGB_CUDA(cudaMemsetAsync(moments, 0, (size_t)c * 3 * sizeof(double), st));
    return 0;
  }
  if (dtype == MGB_F32) return stats_t<float>(n, c, s, (const float*)x, (double*)moments, (double*)ws, st);
  if (dtype == MGB_F64) return stats_t<double>(n, c, s, (const double*)x, (double*)moments, (double*)ws, st);
  return fail("mgb_batchnorm_stats: unsupported dtype %d", dtype);
}

int mgb_batchnorm_fwd(int dtype, int64_t n, int64_t c, int64_t s, const void* x, const void* moments,
                      const void* gamma, const void* beta, double eps, void* y, void* stream) {
  if (check_shape(n, c, s, "mgb_batchnorm_fwd")) return 1;
  if (n * c * s == 0) return 0;
  cudaStream_t st = as_stream(stream);
  if (dtype == MGB_F32)
    return apply_t<float>(n, c, s, (const float*)x, (const double*)moments, (const float*)gamma,
                          (const float*)beta, eps, (float*)y, st);
  if (dtype == MGB_F64)
    return apply_t<double>(n, c, s, (const double*)x, (const double*)moments, (const double*)gamma,
                           (const double*)beta, eps, (double*)y, st);
  return fail("mgb_batchnorm_fwd: unsupported dtype %d", dtype);
}

int mgb_batchnorm_bwd_reduce(int dtype, int64_t n, int64_t c, int64_t s, const void* grad, const void* x,
                             const void* moments, double eps, void* sums, void* ws, size_t ws_bytes,
                             void* stream) {
  if (check_shape(n, c, s, "mgb_batchnorm_bwd_reduce")) return 1;
  if (c == 0) return 0;
  size_t need = 0;
  mgb_batchnorm_workspace_bytes(n, c, &need);
  if (!ws || ws_bytes < need) return fail("mgb_batchnorm_bwd_reduce: workspace %zu < %zu bytes", ws_bytes, need);
  cudaStream_t st = as_stream(stream);
  if (n * s == 0) {
    MGB_CUDA(cudaMemsetAsync(sums, 0, (size_t)c * 2 * sizeof(double), st));
    return 0;
  }
  if (dtype == MGB_F32)
    return bwd_reduce_t<float>(n, c, s, (const float*)grad, (const float*)x, (const double*)moments, eps,
                               (double*)sums, (double*)ws, st);
  if (dtype == MGB_F64)
    return bwd_reduce_t<double>(n, c, s, (const double*)grad, (const double*)x, (const double*)moments, eps,
                                (double*)sums, (double*)ws, st);
  return fail("mgb_batchnorm_bwd_reduce: unsupported dtype %d", dtype);
}

int mgb_batchnorm_bwd_dx(int dtype, int64_t n, int64_t c, int64_t s, const void* grad, const void* x,
                         const void* moments, const void* sums, const void* gamma, double eps, void* dx,
                         void* stream) {
  if (check_shape(n, c, s, "mgb_batchnorm_bwd_dx")) return 1;
  if (n * c * s == 0) return 0;
  cudaStream_t st = as_stream(stream);
  if (dtype == MGB_F32)
    return bwd_dx_t<float>(n, c, s, (const float*)grad, (const float*)x, (const double*)moments,
                           (const double*)sums, (const float*)gamma, eps, (float*)dx, st);
  if (dtype == MGB_F64)
    return bwd_dx_t<double>(n, c, s, (const double*)grad, (const double*)x, (const double*)moments,
                            (const double*)sums, (const double*)gamma, eps, (double*)dx, st);
  return fail("mgb_batchnorm_bwd_dx: unsupported dtype %d", dtype);
}

int mgb_batchnorm_merge(int phase, int64_t c, void* moments, void* buf, void* stream) {
  if (phase < 0 || phase > 2) return fail("mgb_batchnorm_merge: phase %d", phase);
  if (c <= 0) return 0;
  bn_merge_kernel<<<(unsigned)((c + 127) / 128), 128, 0, as_stream(stream)>>>(phase, c, (double*)moments,
                                                                              (double*)buf);
  MGB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
