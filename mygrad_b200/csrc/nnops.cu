// "Config-5 closure" ops: the few element-wise / small ops a conv net needs
// around conv_nd and max_pool so that a whole Tensor.backward stays on the device.
// All HBM-bound (or tiny).  Reference semantics restated:
//   ReLu              src/mygrad/nnet/activations/relu.py:10-17   y = x * (x > 0), dx = g * (x > 0)
//   Add (broadcast)   src/mygrad/math/arithmetic/ops.py:28-32     out = a + b (NumPy broadcasting)
//   SoftmaxCrossEntropy src/mygrad/nnet/losses/softmax_crossentropy.py:14-50 with
//                     logsumexp src/mygrad/math/_special.py:4-68
//   transpose (for MatMul = 1x1 conv, math/misc/ops.py:91-127)
#include <math.h>

#include "common.cuh"

namespace mgb {
namespace {

constexpr int kThreads = 256;

inline int grid_for(size_t n) {
  size_t b = (n + kThreads - 1) / kThreads, cap = (size_t)sm_count() * 32;
  return (int)std::max<size_t>(1, std::min(b, cap));
}

template <typename T>
__global__ void __launch_bounds__(kThreads) relu_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, size_t n) {
  constexpr int V = 16 / sizeof(T);
  size_t nv = n / V, tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = tid; i < nv; i += stride) {
    float4 raw = __ldg(reinterpret_cast<const float4*>(x) + i);
    T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
    for (int k = 0; k < V; ++k) e[k] = e[k] * (e[k] > (T)0 ? (T)1 : (T)0);   // -x*0 = -0.0, NaN stays NaN
    reinterpret_cast<float4*>(y)[i] = raw;
  }
  for (size_t i = nv * V + tid; i < n; i += stride) y[i] = x[i] * (x[i] > (T)0 ? (T)1 : (T)0);
}

template <typename T>
__global__ void __launch_bounds__(kThreads) relu_bwd_kernel(const T* __restrict__ g, const T* __restrict__ x,
                                                            T* __restrict__ dx, size_t n) {
  constexpr int V = 16 / sizeof(T);
  size_t nv = n / V, tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = tid; i < nv; i += stride) {
    float4 gr = __ldg(reinterpret_cast<const float4*>(g) + i), xr = __ldg(reinterpret_cast<const float4*>(x) + i);
    T* ge = reinterpret_cast<T*>(&gr);
    const T* xe = reinterpret_cast<const T*>(&xr);
#pragma unroll
    for (int k = 0; k < V; ++k) ge[k] = ge[k] * (xe[k] > (T)0 ? (T)1 : (T)0);
    reinterpret_cast<float4*>(dx)[i] = gr;
  }
  for (size_t i = nv * V + tid; i < n; i += stride) dx[i] = g[i] * (x[i] > (T)0 ? (T)1 : (T)0);
}

struct BcastArgs {
  int nd;
  long long shape[MGB_MAX_DIMS];   // output shape
  long long sa[MGB_MAX_DIMS];      // element strides of a (0 on broadcast axes)
  long long sb[MGB_MAX_DIMS];
};

template <typename T>
__global__ void __launch_bounds__(kThreads) add_bcast_kernel(BcastArgs a, const T* __restrict__ x,
                                                             const T* __restrict__ y, T* __restrict__ out,
                                                             long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    long long t = i, oa = 0, ob = 0;
    for (int d = a.nd - 1; d >= 0; --d) {
      long long q = t / a.shape[d], r = t - q * a.shape[d];
      oa += r * a.sa[d];
      ob += r * a.sb[d];
      t = q;
    }
    out[i] = x[oa] + y[ob];
  }
}

// a has the output shape, b varies along one (merged) axis only: out[i] = a[i] + b[(i / inner) % len]
template <typename T>
__global__ void __launch_bounds__(kThreads) add_axis_kernel(const T* __restrict__ a, const T* __restrict__ b,
                                                            T* __restrict__ out, long long n, long long inner,
                                                            long long len) {
  constexpr int V = 16 / sizeof(T);
  const bool vec = (inner % V == 0) && (((uintptr_t)a | (uintptr_t)out) % 16 == 0);
  long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  if (vec) {
    long long nv = n / V;
    for (long long i = tid; i < nv; i += stride) {
      float4 raw = __ldg(reinterpret_cast<const float4*>(a) + i);
      T* e = reinterpret_cast<T*>(&raw);
      T bv = __ldg(b + ((i * V) / inner) % len);     // the whole vector shares one b element
#pragma unroll
      for (int k = 0; k < V; ++k) e[k] += bv;
      reinterpret_cast<float4*>(out)[i] = raw;
    }
  } else {
    for (long long i = tid; i < n; i += stride) out[i] = a[i] + __ldg(b + (i / inner) % len);
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) scale_by_kernel(const T* __restrict__ scalar, const T* __restrict__ src,
                                                            T* __restrict__ dst, size_t n) {
  const T s = __ldg(scalar);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = s * src[i];
}

template <typename T>
__global__ void __launch_bounds__(256) transpose2d_kernel(const T* __restrict__ src, T* __restrict__ dst,
                                                          long long rows, long long cols) {
  __shared__ T tile[32][33];
  long long c0 = (long long)blockIdx.x * 32, r0 = (long long)blockIdx.y * 32;
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long r = r0 + ty + 8 * i, c = c0 + tx;
    if (r < rows && c < cols) tile[ty + 8 * i][tx] = src[r * cols + c];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long c = c0 + ty + 8 * i, r = r0 + tx;
    if (r < rows && c < cols) dst[c * rows + r] = tile[tx][ty + 8 * i];
  }
}

template <typename T> __device__ __forceinline__ T texp(T v);
template <> __device__ __forceinline__ float texp(float v) { return expf(v); }
template <> __device__ __forceinline__ double texp(double v) { return exp(v); }
template <typename T> __device__ __forceinline__ T tlog(T v);
template <> __device__ __forceinline__ float tlog(float v) { return logf(v); }
template <> __device__ __forceinline__ double tlog(double v) { return log(v); }

// one warp per row: log-softmax via the reference's logsumexp (max with non-finite -> 0),
// back = (softmax - onehot) / N, row_loss = -log_softmax[label]
template <typename T>
__global__ void __launch_bounds__(kThreads) softmax_xent_rows(const T* __restrict__ x, const T* __restrict__ labels,
                                                              T* __restrict__ back, double* __restrict__ row_loss,
                                                              int N, int C, int want_back) {
  int row = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (row >= N) return;
  const T* xr = x + (long long)row * C;
  T m = -INFINITY;
  bool has_nan = false;
  for (int c = lane; c < C; c += 32) {
    T v = xr[c];
    if (v != v) has_nan = true;
    m = v > m ? v : m;
  }
  for (int o = 16; o > 0; o >>= 1) {
    T other = __shfl_xor_sync(0xffffffffu, m, o);
    m = other > m ? other : m;
    has_nan = has_nan || __shfl_xor_sync(0xffffffffu, (int)has_nan, o);
  }
  if (has_nan) m = NAN;                    // np.amax propagates NaN ...
  if (!isfinite((double)m)) m = (T)0;      // ... and logsumexp zeroes a non-finite maximum
  T s = 0;
  for (int c = lane; c < C; c += 32) s += texp(xr[c] - m);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const T lse = tlog(s) + m;
  const int label = (int)labels[row];
  if (want_back) {
    T* br = back + (long long)row * C;
    for (int c = lane; c < C; c += 32) {
      T p = texp(xr[c] - lse);
      if (c == label) p -= (T)1;
      br[c] = p / (T)N;
    }
  }
  if (lane == 0) row_loss[row] = -(double)(xr[label] - lse);
}

template <typename T>
__global__ void __launch_bounds__(kThreads) softmax_xent_finish(const double* __restrict__ row_loss, T* __restrict__ loss,
                                                                int N) {
  __shared__ double sh[kThreads];
  double acc = 0.0;
  for (int i = threadIdx.x; i < N; i += kThreads) acc += row_loss[i];   // fixed assignment: deterministic
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = (T)(sh[0] / (double)N);
}


// out[g0, g1, g2, lead, w0, w1, w2] = src[lead, g*step + w*dil]  (sliding_window_view,
// src/mygrad/nnet/layers/utils.py:198-228, materialised: device arrays have no strided views)
struct WindowGeom {
  long long lead, X[3], G[3], W[3], step[3], dil[3];
};

template <typename T>
__global__ void __launch_bounds__(kThreads) window_gather_kernel(const T* __restrict__ src, T* __restrict__ dst,
                                                                 WindowGeom g, long long total) {
  const long long wp = g.W[0] * g.W[1] * g.W[2];
  const long long xs1 = g.X[2], xs0 = g.X[1] * g.X[2], xp = g.X[0] * xs0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const long long w = t % wp; t /= wp;
    const long long l = t % g.lead; t /= g.lead;
    const long long g2 = t % g.G[2]; t /= g.G[2];
    const long long g1 = t % g.G[1];
    const long long g0 = t / g.G[1];
    const long long w2 = w % g.W[2], w1 = (w / g.W[2]) % g.W[1], w0 = w / (g.W[2] * g.W[1]);
    const long long x0 = g0 * g.step[0] + w0 * g.dil[0], x1 = g1 * g.step[1] + w1 * g.dil[1],
                    x2 = g2 * g.step[2] + w2 * g.dil[2];
    dst[i] = src[l * xp + x0 * xs0 + x1 * xs1 + x2];
  }
}

}  // namespace
}  // namespace mgb

using namespace mgb;

extern "C" int mgb_relu_fwd(int dtype, size_t n, const void* x, void* y, void* stream) {
  if (n == 0) return 0;
  if (!x || !y) return fail("mgb_relu_fwd: null");
  cudaStream_t st = as_stream(stream);
  if (dtype == MGB_F32) relu_fwd_kernel<float><<<grid_for(n / 4 + 1), kThreads, 0, st>>>((const float*)x, (float*)y, n);
  else if (dtype == MGB_F64) relu_fwd_kernel<double><<<grid_for(n / 2 + 1), kThreads, 0, st>>>((const double*)x, (double*)y, n);
  else return fail("mgb_relu_fwd: bad dtype %d", dtype);
  MGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mgb_relu_bwd(int dtype, size_t n, const void* grad, const void* x, void* dx, void* stream) {
  if (n == 0) return 0;
  if (!grad || !x || !dx) return fail("mgb_relu_bwd: null");
  cudaStream_t st = as_stream(stream);
  if (dtype == MGB_F32)
    relu_bwd_kernel<float><<<grid_for(n / 4 + 1), kThreads, 0, st>>>((const float*)grad, (const float*)x, (float*)dx, n);
  else if (dtype == MGB_F64)
    relu_bwd_kernel<double><<<grid_for(n / 2 + 1), kThreads, 0, st>>>((const double*)grad, (const double*)x, (double*)dx, n);
  else return fail("mgb_relu_bwd: bad dtype %d", dtype);
  MGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mgb_add_broadcast(int dtype, int ndim, const int64_t* out_shape, const int64_t* a_shape,
                                 const int64_t* b_shape, const void* a, const void* b, void* out, void* stream) {
  if (dtype != MGB_F32 && dtype != MGB_F64) return fail("mgb_add_broadcast: bad dtype %d", dtype);
  if (ndim < 0 || ndim > MGB_MAX_DIMS) return fail("mgb_add_broadcast: rank %d > %d", ndim, MGB_MAX_DIMS);
  BcastArgs args;
  args.nd = ndim;
  long long n = 1, na = 1, nb = 1;
  for (int d = ndim - 1; d >= 0; --d) {
    long long o = out_shape[d], sa = a_shape[d], sb = b_shape[d];
    if ((sa != o && sa != 1) || (sb != o && sb != 1)) return fail("mgb_add_broadcast: axis %d not broadcastable", d);
    args.shape[d] = o;
    args.sa[d] = sa == 1 ? 0 : na;
    args.sb[d] = sb == 1 ? 0 : nb;
    na *= sa; nb *= sb; n *= o;
  }
  if (n == 0) return 0;
  if (!a || !b || !out) return fail("mgb_add_broadcast: null");
  cudaStream_t st = as_stream(stream);
  // fast path: one operand has the output shape, the other varies along one run of axes only
  for (int swap = 0; swap < 2; ++swap) {
    const int64_t* full = swap ? b_shape : a_shape;
    const int64_t* part = swap ? a_shape : b_shape;
    bool is_full = true;
    for (int d = 0; d < ndim; ++d) is_full = is_full && full[d] == out_shape[d];
    if (!is_full) continue;
    int first = -1, last = -1;
    bool ok = true;
    for (int d = 0; d < ndim; ++d)
      if (part[d] != 1) { if (first < 0) first = d; last = d; }
    for (int d = first; d >= 0 && d <= last; ++d) ok = ok && part[d] == out_shape[d];
    if (!ok) continue;
    long long inner = 1, len = 1;
    for (int d = ndim - 1; d > last && last >= 0; --d) inner *= out_shape[d];
    for (int d = first; d >= 0 && d <= last; ++d) len *= out_shape[d];
    if (first < 0) { inner = n; len = 1; }
    const void* pa = swap ? b : a;
    const void* pb = swap ? a : b;
    if (dtype == MGB_F32)
      add_axis_kernel<float><<<grid_for((size_t)n / 4 + 1), kThreads, 0, st>>>((const float*)pa, (const float*)pb, (float*)out, n, inner, len);
    else
      add_axis_kernel<double><<<grid_for((size_t)n / 2 + 1), kThreads, 0, st>>>((const double*)pa, (const double*)pb, (double*)out, n, inner, len);
    MGB_LAUNCH_CHECK();
    return 0;
  }
  if (dtype == MGB_F32)
    add_bcast_kernel<float><<<grid_for((size_t)n), kThreads, 0, st>>>(args, (const float*)a, (const float*)b, (float*)out, n);
  else
    add_bcast_kernel<double><<<grid_for((size_t)n), kThreads, 0, st>>>(args, (const double*)a, (const double*)b, (double*)out, n);
  MGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mgb_transpose2d(int dtype, int64_t rows, int64_t cols, const void* src, void* dst, void* stream) {
  if (rows < 0 || cols < 0) return fail("mgb_transpose2d: negative extent");
  if (rows * cols == 0) return 0;
  if (!src || !dst) return fail("mgb_transpose2d: null");
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
  if (grid.y > 65535) return fail("mgb_transpose2d: too many rows (%lld)", (long long)rows);
  cudaStream_t st = as_stream(stream);
  if (dtype == MGB_F32) transpose2d_kernel<float><<<grid, 256, 0, st>>>((const float*)src, (float*)dst, rows, cols);
  else if (dtype == MGB_F64) transpose2d_kernel<double><<<grid, 256, 0, st>>>((const double*)src, (double*)dst, rows, cols);
  else return fail("mgb_transpose2d: bad dtype %d", dtype);
  MGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mgb_softmax_xent(int dtype, int64_t n, int64_t c, const void* x, const void* labels, void* loss,
                                void* back, void* stream) {
  if (dtype != MGB_F32 && dtype != MGB_F64) return fail("mgb_softmax_xent: bad dtype %d", dtype);
  if (n <= 0 || c <= 0 || n > INT32_MAX || c > INT32_MAX) return fail("mgb_softmax_xent: bad shape (%lld, %lld)", (long long)n, (long long)c);
  if (!x || !labels || !loss) return fail("mgb_softmax_xent: null");
  cudaStream_t st = as_stream(stream);
  double* row_loss = nullptr;
  MGB_CUDA(cudaMallocAsync((void**)&row_loss, sizeof(double) * (size_t)n, st));
  int blocks = (int)((n * 32 + kThreads - 1) / kThreads);
  if (dtype == MGB_F32) {
    softmax_xent_rows<float><<<blocks, kThreads, 0, st>>>((const float*)x, (const float*)labels, (float*)back, row_loss, (int)n, (int)c, back != nullptr);
    MGB_LAUNCH_CHECK();
    softmax_xent_finish<float><<<1, kThreads, 0, st>>>(row_loss, (float*)loss, (int)n);
  } else {
    softmax_xent_rows<double><<<blocks, kThreads, 0, st>>>((const double*)x, (const double*)labels, (double*)back, row_loss, (int)n, (int)c, back != nullptr);
    MGB_LAUNCH_CHECK();
    softmax_xent_finish<double><<<1, kThreads, 0, st>>>(row_loss, (double*)loss, (int)n);
  }
  MGB_LAUNCH_CHECK();
  MGB_CUDA(cudaFreeAsync(row_loss, st));
  return 0;
}

extern "C" int mgb_scale_by(int dtype, size_t n, const void* scalar, const void* src, void* dst, void* stream) {
  if (n == 0) return 0;
  if (!scalar || !src || !dst) return fail("mgb_scale_by: null");
  cudaStream_t st = as_stream(stream);
  if (dtype == MGB_F32) scale_by_kernel<float><<<grid_for(n), kThreads, 0, st>>>((const float*)scalar, (const float*)src, (float*)dst, n);
  else if (dtype == MGB_F64) scale_by_kernel<double><<<grid_for(n), kThreads, 0, st>>>((const double*)scalar, (const double*)src, (double*)dst, n);
  else return fail("mgb_scale_by: bad dtype %d", dtype);
  MGB_LAUNCH_CHECK();
  return 0;
}

// ---- skinny matmul: (n, d) @ (d, m) with a small m (classifier heads) -------------------------
// The conv GEMM tiles (128 x >=32) waste the machine on m = 10: these three kernels are
// bandwidth-bound instead (A is read once, B stays in L2 / shared memory).
namespace mgb {
namespace {

constexpr int kSkinnyMaxM = 32;

// All three kernels are templated on MP = m rounded up to 4 / 8 / 16 / 32 (a run-time `j < m` inside a 32-wide
// unrolled loop issues 32 predicated FMAs for m = 10) and accumulate runs of at most 64-128 products in the input
// type before adding them to a float64 total: float32 inputs keep float32 FMA rate, the long sums stay exact to
// ~1e-7 (the reference's sgemm accumulates in float32 throughout).

// C[n][j] = sum_d A[n][d] * B[d][j]: a block owns RB rows; thread t walks k = t, t + 256, ...: one row of B
// (m contiguous values) and RB coalesced elements of A per k for RB * m FMAs, then the block sums its 256
// partial tiles (shuffles, then warp order through shared memory: deterministic).  A is read once; B comes from L2.
template <typename T, int MP, int RB>
__global__ void __launch_bounds__(256) skinny_fwd_kernel(const T* __restrict__ A, const T* __restrict__ B,
                                                         T* __restrict__ Cout, long long n, int d, int m) {
  __shared__ double red[8][RB * MP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = (long long)blockIdx.x * RB;
  double acc[RB][MP];
#pragma unroll
  for (int r = 0; r < RB; ++r)
#pragma unroll
    for (int j = 0; j < MP; ++j) acc[r][j] = 0.0;
  if constexpr (sizeof(T) == 8) {
    for (int k = threadIdx.x; k < d; k += 256) {
      T b[MP];
#pragma unroll
      for (int j = 0; j < MP; ++j) b[j] = j < m ? __ldg(B + (long long)k * m + j) : (T)0;
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const T av = row0 + r < n ? __ldg(A + (row0 + r) * d + k) : (T)0;
#pragma unroll
        for (int j = 0; j < MP; ++j) acc[r][j] += av * b[j];
      }
    }
  } else {
    for (int k0 = 0; k0 < d; k0 += 256 * 64) {            // runs of <= 64 products in float32, totals in float64
      T part[RB][MP];
#pragma unroll
      for (int r = 0; r < RB; ++r)
#pragma unroll
        for (int j = 0; j < MP; ++j) part[r][j] = (T)0;
      for (int k = k0 + threadIdx.x; k < min(d, k0 + 256 * 64); k += 256) {
        T b[MP];
#pragma unroll
        for (int j = 0; j < MP; ++j) b[j] = j < m ? __ldg(B + (long long)k * m + j) : (T)0;
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          const T av = row0 + r < n ? __ldg(A + (row0 + r) * d + k) : (T)0;
#pragma unroll
          for (int j = 0; j < MP; ++j) part[r][j] += av * b[j];
        }
      }
#pragma unroll
      for (int r = 0; r < RB; ++r)
#pragma unroll
        for (int j = 0; j < MP; ++j) acc[r][j] += (double)part[r][j];
    }
  }
#pragma unroll
  for (int r = 0; r < RB; ++r)
#pragma unroll
    for (int j = 0; j < MP; ++j) {
      double v = acc[r][j];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[warp][r * MP + j] = v;
    }
  __syncthreads();
  for (int i = threadIdx.x; i < RB * MP; i += 256) {
    const int r = i / MP, j = i - r * MP;
    if (row0 + r < n && j < m) {
      double v = 0.0;
      for (int w = 0; w < 8; ++w) v += red[w][i];
      Cout[(row0 + r) * m + j] = (T)v;
    }
  }
}

// dA[n][k] = sum_j G[n][j] * B[k][j]: threads over k (coalesced writes), B row in registers
template <typename T, int MP>
__global__ void __launch_bounds__(256) skinny_da_kernel(const T* __restrict__ G, const T* __restrict__ B,
                                                        T* __restrict__ dA, long long n, int d, int m) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  T brow[MP];
#pragma unroll
  for (int j = 0; j < MP; ++j) brow[j] = (k < d && j < m) ? B[(long long)k * m + j] : (T)0;
  __shared__ T gs[64 * MP];
  const long long r0 = (long long)blockIdx.y * 64;
  const int rows = (int)min(64LL, n - r0);
  for (int i = threadIdx.x; i < rows * MP; i += blockDim.x) {
    const int r = i / MP, j = i - r * MP;
    gs[i] = j < m ? G[(r0 + r) * m + j] : (T)0;
  }
  __syncthreads();
  if (k >= d) return;
  for (int r = 0; r < rows; ++r) {
    T acc = (T)0;
#pragma unroll
    for (int j = 0; j < MP; ++j) acc += gs[r * MP + j] * brow[j];
    dA[(r0 + r) * d + k] = acc;
  }
}

// dB[k][j] = sum_n A[n][k] * G[n][j]: threads over k, split over n with float64 partials
template <typename T, int MP>
__global__ void __launch_bounds__(256) skinny_db_kernel(const T* __restrict__ A, const T* __restrict__ G,
                                                        double* __restrict__ partial, long long n, int d, int m,
                                                        long long rows_per_split) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const long long r0 = (long long)blockIdx.y * rows_per_split, r1 = min(n, r0 + rows_per_split);
  __shared__ T gs[64 * MP];
  double acc[MP];
#pragma unroll
  for (int j = 0; j < MP; ++j) acc[j] = 0.0;
  for (long long rb = r0; rb < r1; rb += 64) {
    const int rows = (int)min(64LL, r1 - rb);
    __syncthreads();
    for (int i = threadIdx.x; i < rows * MP; i += blockDim.x) {
      const int r = i / MP, j = i - r * MP;
      gs[i] = j < m ? G[(rb + r) * m + j] : (T)0;
    }
    __syncthreads();
    if (k < d) {
      T part[MP];
#pragma unroll
      for (int j = 0; j < MP; ++j) part[j] = (T)0;
      for (int r = 0; r < rows; ++r) {                  // <= 64 products per run
        const T av = A[(rb + r) * d + k];
#pragma unroll
        for (int j = 0; j < MP; ++j) part[j] += av * gs[r * MP + j];
      }
#pragma unroll
      for (int j = 0; j < MP; ++j) acc[j] += (double)part[j];
    }
  }
  if (k < d) {
    double* out = partial + ((long long)blockIdx.y * d + k) * m;
#pragma unroll
    for (int j = 0; j < MP; ++j)
      if (j < m) out[j] = acc[j];
  }
}

template <typename T>
__global__ void __launch_bounds__(256) skinny_db_finish(const double* __restrict__ partial, T* __restrict__ dB,
                                                        long long total, int splits) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int s = 0; s < splits; ++s) acc += partial[(long long)s * total + i];   // fixed order
    dB[i] = (T)acc;
  }
}

template <typename T, int MP>
int skinny_mp(int which, long long n, int d, int m, const T* a, const T* b, T* out, cudaStream_t st) {
  if (which == 0) {
    constexpr int RB = MP <= 8 ? 8 : MP <= 16 ? 4 : 2;
    skinny_fwd_kernel<T, MP, RB><<<(unsigned)((n + RB - 1) / RB), 256, 0, st>>>(a, b, out, n, d, m);
    MGB_LAUNCH_CHECK();
  } else if (which == 1) {
    dim3 grid((d + 255) / 256, (unsigned)((n + 63) / 64));
    skinny_da_kernel<T, MP><<<grid, 256, 0, st>>>(a, b, out, n, d, m);
    MGB_LAUNCH_CHECK();
  } else {
    int kb = (d + 255) / 256;
    int splits = (int)std::max<long long>(1, std::min<long long>((n + 63) / 64, (4LL * sm_count() + kb - 1) / kb));
    long long per = ((n + splits - 1) / splits + 63) / 64 * 64;
    splits = (int)((n + per - 1) / per);
    double* partial = nullptr;
    MGB_CUDA(cudaMallocAsync((void**)&partial, sizeof(double) * (size_t)splits * d * m, st));
    skinny_db_kernel<T, MP><<<dim3(kb, splits), 256, 0, st>>>(a, b, partial, n, d, m, per);
    MGB_LAUNCH_CHECK();
    long long total = (long long)d * m;
    skinny_db_finish<T><<<(unsigned)std::min<long long>((total + 255) / 256, 1024), 256, 0, st>>>(partial, out, total, splits);
    MGB_LAUNCH_CHECK();
    MGB_CUDA(cudaFreeAsync(partial, st));
  }
  return 0;
}

template <typename T>
int skinny_t(int which, long long n, int d, int m, const T* a, const T* b, T* out, cudaStream_t st) {
  if (m <= 4) return skinny_mp<T, 4>(which, n, d, m, a, b, out, st);
  if (m <= 8) return skinny_mp<T, 8>(which, n, d, m, a, b, out, st);
  if (m <= 16) return skinny_mp<T, 16>(which, n, d, m, a, b, out, st);
  return skinny_mp<T, 32>(which, n, d, m, a, b, out, st);
}

}  // namespace
}  // namespace mgb

// which = 0: out (n,m) = a (n,d) @ b (d,m);  1: out (n,d) = a=G (n,m) @ b (d,m)^T;
//         2: out (d,m) = a (n,d)^T @ b=G (n,m).   Requires m <= 32.
extern "C" int mgb_matmul_skinny(int dtype, int which, int64_t n, int64_t d, int64_t m, const void* a,
                                 const void* b, void* out, void* stream) {
  if (dtype != MGB_F32 && dtype != MGB_F64) return fail("mgb_matmul_skinny: bad dtype %d", dtype);
  if (which < 0 || which > 2) return fail("mgb_matmul_skinny: which=%d", which);
  if (m < 1 || m > kSkinnyMaxM || n < 0 || d < 0 || d > INT32_MAX) return fail("mgb_matmul_skinny: needs 1 <= m <= %d", kSkinnyMaxM);
  if (n == 0 || d == 0) {
    size_t esz = dtype == MGB_F32 ? 4 : 8;
    size_t cnt = which == 0 ? (size_t)(n * m) : (which == 1 ? (size_t)(n * d) : (size_t)(d * m));
    if (cnt) MGB_CUDA(cudaMemsetAsync(out, 0, cnt * esz, as_stream(stream)));
    return 0;
  }
  if (!a || !b || !out) return fail("mgb_matmul_skinny: null");
  if (dtype == MGB_F32)
    return skinny_t<float>(which, n, (int)d, (int)m, (const float*)a, (const float*)b, (float*)out, as_stream(stream));
  return skinny_t<double>(which, n, (int)d, (int)m, (const double*)a, (const double*)b, (double*)out, as_stream(stream));
}

extern "C" int mgb_window_gather(int dtype, int nd, int64_t lead, const int64_t* x, const int64_t* w,
                                 const int64_t* step, const int64_t* dil, const void* src, void* dst,
                                 void* stream) {
  if (nd < 1 || nd > 3) return fail("mgb_window_gather: %d windowed axes (1..3 supported)", nd);
  WindowGeom g;
  g.lead = lead;
  long long total = lead;
  for (int i = 0; i < 3; ++i) {              // right-align the windowed axes in 3 slots
    int j = i - (3 - nd);
    g.X[i] = j >= 0 ? x[j] : 1; g.W[i] = j >= 0 ? w[j] : 1;
    g.step[i] = j >= 0 ? step[j] : 1; g.dil[i] = j >= 0 ? dil[j] : 1;
    if (g.W[i] < 1 || g.step[i] < 1 || g.dil[i] < 1) return fail("mgb_window_gather: bad window/step/dilation");
    long long span = (g.W[i] - 1) * g.dil[i] + 1;
    if (span > g.X[i]) return fail("mgb_window_gather: window does not fit axis %d", j);
    g.G[i] = (g.X[i] - span) / g.step[i] + 1;
    total *= g.G[i] * g.W[i];
  }
  if (total == 0) return 0;
  if (!src || !dst) return fail("mgb_window_gather: null");
  cudaStream_t st = as_stream(stream);
  if (dtype == MGB_F32)
    window_gather_kernel<float><<<grid_for((size_t)total), kThreads, 0, st>>>((const float*)src, (float*)dst, g, total);
  else if (dtype == MGB_F64)
    window_gather_kernel<double><<<grid_for((size_t)total), kThreads, 0, st>>>((const double*)src, (double*)dst, g, total);
  else
    return fail("mgb_window_gather: bad dtype %d", dtype);
  MGB_LAUNCH_CHECK();
  return 0;
}
