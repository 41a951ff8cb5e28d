// Gradient bookkeeping kernels of Tensor.backward: seed fill, copy/cast,
// accumulate, scale, and the un-broadcast reduction.  All HBM-bound.
//
// Reference call sites restated:
//   np.full_like(data, 1.0)            tensor_base.py:1332        -> mgb_fill
//   np.copy / astype(var.dtype)        operation_base.py:217-224  -> mgb_copy_cast
//   var._grad += backed_grad           operation_base.py:227-228  -> mgb_axpy
//   reduce_broadcast(grad, var_shape)  _utils/__init__.py:131-168 -> mgb_reduce_broadcast
#include "common.cuh"

namespace mgb {
namespace {

constexpr int kThreads = 256;

inline int grid_for(size_t work_items) {
  size_t blocks = (work_items + kThreads - 1) / kThreads;
  size_t cap = (size_t)sm_count() * 32;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) fill_kernel(T* __restrict__ dst, size_t n, T value) {
  constexpr int V = 16 / sizeof(T);
  size_t nv = n / V;
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  float4 raw;
  T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
  for (int k = 0; k < V; ++k) e[k] = value;
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (size_t i = tid; i < nv; i += stride) d4[i] = raw;
  for (size_t i = nv * V + tid; i < n; i += stride) dst[i] = value;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) axpy_kernel(const T* __restrict__ src,
                                                        T* __restrict__ dst, size_t n) {
  constexpr int V = 16 / sizeof(T);
  size_t nv = n / V;
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (size_t i = tid; i < nv; i += stride) {
    float4 a = __ldg(s4 + i), b = d4[i];
    const T* ea = reinterpret_cast<const T*>(&a);
    T* eb = reinterpret_cast<T*>(&b);
#pragma unroll
    for (int k = 0; k < V; ++k) eb[k] += ea[k];
    d4[i] = b;
  }
  for (size_t i = nv * V + tid; i < n; i += stride) dst[i] += src[i];
}

template <typename T>
__global__ void __launch_bounds__(kThreads) scale_kernel(T* __restrict__ dst, size_t n, T alpha) {
  constexpr int V = 16 / sizeof(T);
  size_t nv = n / V;
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (size_t i = tid; i < nv; i += stride) {
    float4 raw = d4[i];
    T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
    for (int k = 0; k < V; ++k) e[k] *= alpha;
    d4[i] = raw;
  }
  for (size_t i = nv * V + tid; i < n; i += stride) dst[i] *= alpha;
}

// element-wise fallback for buffers that are not 16-byte aligned (views into a larger allocation)
template <typename S, typename D>
__global__ void __launch_bounds__(kThreads) cast_scalar_kernel(const S* __restrict__ src, D* __restrict__ dst,
                                                               size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = (D)src[i];
}

template <typename S, typename D>
__global__ void __launch_bounds__(kThreads) cast_kernel(const S* __restrict__ src,
                                                        D* __restrict__ dst, size_t n) {
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  // 4 elements per step: 16 B on the narrow side, 32 B on the wide side
  size_t n4 = n / 4;
  for (size_t i = tid; i < n4; i += stride) {
    S s[4];
    if (sizeof(S) == 4) {
      *reinterpret_cast<float4*>(s) = __ldg(reinterpret_cast<const float4*>(src) + i);
    } else {
      *reinterpret_cast<float4*>(s) = __ldg(reinterpret_cast<const float4*>(src) + 2 * i);
      *reinterpret_cast<float4*>(s + 2) = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    }
    D d[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) d[k] = (D)s[k];
    if (sizeof(D) == 4) {
      reinterpret_cast<float4*>(dst)[i] = *reinterpret_cast<float4*>(d);
    } else {
      reinterpret_cast<float4*>(dst)[2 * i] = *reinterpret_cast<float4*>(d);
      reinterpret_cast<float4*>(dst)[2 * i + 1] = *reinterpret_cast<float4*>(d + 2);
    }
  }
  for (size_t i = n4 * 4 + tid; i < n; i += stride) dst[i] = (D)src[i];
}

// ---- un-broadcast reduction -------------------------------------------------
//
// After aligning var_shape to the right of grad_shape and merging neighbouring
// axes of the same kind, the index space is an alternating list of kept (K) and
// reduced (R) groups.  Element offset = off_K(o) + off_R(j).

constexpr int kMaxGroups = MGB_MAX_DIMS;

struct RGroups {
  int nk, nr;
  long long ksize[kMaxGroups], kstride[kMaxGroups];  // kept groups, outermost first
  long long rsize[kMaxGroups], rstride[kMaxGroups];  // reduced groups, outermost first
  long long n_out, n_red;
};

__device__ __forceinline__ long long kept_offset(const RGroups& g, long long o) {
  long long off = 0;
  for (int i = g.nk - 1; i >= 0; --i) {
    long long q = o / g.ksize[i];
    off += (o - q * g.ksize[i]) * g.kstride[i];
    o = q;
  }
  return off;
}

__device__ __forceinline__ double block_sum(double v) {
  __shared__ double sh[32];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  v = (l < nw) ? sh[l] : 0.0;
  if (w == 0)
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;  // valid in warp 0
}

// Innermost group reduced and contiguous: one CTA per (output, split); warps walk
// the outer reduced index, lanes stream the contiguous inner run.
template <typename T>
__global__ void __launch_bounds__(kThreads) reduce_inner_kernel(RGroups g, const T* __restrict__ src,
                                                                double* __restrict__ partial,
                                                                int splits) {
  long long o = blockIdx.x;
  int split = blockIdx.y;
  long long base = kept_offset(g, o);
  long long inner = g.rsize[g.nr - 1];
  long long outer = g.n_red / inner;
  long long per = (outer + splits - 1) / splits;
  long long j0 = split * per, j1 = min(outer, j0 + per);
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  double acc = 0.0;
  constexpr int V = 16 / sizeof(T);
  // when a run is short, let the whole CTA cooperate on (jo, ji) pairs
  for (long long jo = j0 + warp; jo < j1; jo += nwarps) {
    long long off = base, t = jo;
    for (int i = g.nr - 2; i >= 0; --i) {
      long long q = t / g.rsize[i];
      off += (t - q * g.rsize[i]) * g.rstride[i];
      t = q;
    }
    const T* p = src + off;
    long long head = 0;
    if (inner >= 4 * V) {
      // align to 16 B, then vector body
      uintptr_t addr = (uintptr_t)p;
      head = ((16 - (addr & 15)) & 15) / sizeof(T);
      if (head > inner) head = inner;
      for (long long i = lane; i < head; i += 32) acc += (double)p[i];
      long long nv = (inner - head) / V;
      const float4* p4 = reinterpret_cast<const float4*>(p + head);
      for (long long i = lane; i < nv; i += 32) {
        float4 raw = __ldg(p4 + i);
        const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
        for (int k = 0; k < V; ++k) acc += (double)e[k];
      }
      for (long long i = head + nv * V + lane; i < inner; i += 32) acc += (double)p[i];
    } else {
      for (long long i = lane; i < inner; i += 32) acc += (double)p[i];
    }
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) partial[(long long)split * g.n_out + o] = acc;
}

// Innermost group kept: consecutive threads own consecutive outputs (coalesced),
// each loops over its slice of the reduced index space.
template <typename T>
__global__ void __launch_bounds__(kThreads) reduce_outer_kernel(RGroups g, const T* __restrict__ src,
                                                                double* __restrict__ partial,
                                                                int splits) {
  long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= g.n_out) return;
  int split = blockIdx.y;
  long long base = kept_offset(g, o);
  long long per = (g.n_red + splits - 1) / splits;
  long long j0 = split * per, j1 = min(g.n_red, j0 + per);
  double acc = 0.0;
  if (g.nr == 1) {
    long long rs = g.rstride[0];
    const T* p = src + base + j0 * rs;
    for (long long j = j0; j < j1; ++j, p += rs) acc += (double)__ldg(p);
  } else {
    for (long long j = j0; j < j1; ++j) {
      long long off = base, t = j;
      for (int i = g.nr - 1; i >= 0; --i) {
        long long q = t / g.rsize[i];
        off += (t - q * g.rsize[i]) * g.rstride[i];
        t = q;
      }
      acc += (double)__ldg(src + off);
    }
  }
  partial[(long long)split * g.n_out + o] = acc;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) reduce_finish_kernel(const double* __restrict__ partial,
                                                                 T* __restrict__ out,
                                                                 long long n_out, int splits,
                                                                 int accumulate) {
  long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_out) return;
  double acc = 0.0;
  for (int s = 0; s < splits; ++s) acc += partial[(long long)s * n_out + o];  // fixed order
  if (accumulate) acc += (double)out[o];
  out[o] = (T)acc;
}

template <typename T>
int reduce_t(const RGroups& g, bool inner_reduced, const T* src, T* out, int accumulate,
             cudaStream_t st) {
  int sms = sm_count();
  int splits = 1;
  dim3 grid;
  if (inner_reduced) {
    long long inner = g.rsize[g.nr - 1];
    long long outer = g.n_red / inner;
    // enough CTAs for ~4 waves, at most one split per 8 outer rows
    long long want = ((long long)sms * 4 + g.n_out - 1) / g.n_out;
    long long cap = (outer + 7) / 8;
    splits = (int)max(1LL, min(want, min(cap, 1024LL)));
    if (g.n_out > INT32_MAX) return fail("reduce_broadcast: too many outputs");
    grid = dim3((unsigned)g.n_out, (unsigned)splits);
  } else {
    long long blocks = (g.n_out + kThreads - 1) / kThreads;
    long long want = ((long long)sms * 4 + blocks - 1) / blocks;
    long long cap = (g.n_red + 63) / 64;
    splits = (int)max(1LL, min(want, min(cap, 1024LL)));
    grid = dim3((unsigned)blocks, (unsigned)splits);
  }
  double* partial = nullptr;
  MGB_CUDA(cudaMallocAsync((void**)&partial, sizeof(double) * (size_t)splits * (size_t)g.n_out, st));
  if (inner_reduced)
    reduce_inner_kernel<T><<<grid, kThreads, 0, st>>>(g, src, partial, splits);
  else
    reduce_outer_kernel<T><<<grid, kThreads, 0, st>>>(g, src, partial, splits);
  MGB_LAUNCH_CHECK();
  reduce_finish_kernel<T><<<(unsigned)((g.n_out + kThreads - 1) / kThreads), kThreads, 0, st>>>(
      partial, out, g.n_out, splits, accumulate);
  MGB_LAUNCH_CHECK();
  MGB_CUDA(cudaFreeAsync(partial, st));
  return 0;
}

}  // namespace
}  // namespace mgb

using namespace mgb;

extern "C" int mgb_fill(int dtype, size_t n, double value, void* dst, void* stream) {
  if (n == 0) return 0;
  if (!dst) return fail("mgb_fill: null");
  if (dtype == MGB_F32)
    fill_kernel<float><<<grid_for(n / 4 + 1), kThreads, 0, as_stream(stream)>>>((float*)dst, n, (float)value);
  else if (dtype == MGB_F64)
    fill_kernel<double><<<grid_for(n / 2 + 1), kThreads, 0, as_stream(stream)>>>((double*)dst, n, value);
  else
    return fail("mgb_fill: bad dtype %d", dtype);
  MGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mgb_axpy(int dtype, size_t n, const void* src, void* dst, void* stream) {
  if (n == 0) return 0;
  if (!dst || !src) return fail("mgb_axpy: null");
  if (dtype == MGB_F32)
    axpy_kernel<float><<<grid_for(n / 4 + 1), kThreads, 0, as_stream(stream)>>>((const float*)src, (float*)dst, n);
  else if (dtype == MGB_F64)
    axpy_kernel<double><<<grid_for(n / 2 + 1), kThreads, 0, as_stream(stream)>>>((const double*)src, (double*)dst, n);
  else
    return fail("mgb_axpy: bad dtype %d", dtype);
  MGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mgb_scale(int dtype, size_t n, double alpha, void* dst, void* stream) {
  if (n == 0) return 0;
  if (!dst) return fail("mgb_scale: null");
  if (dtype == MGB_F32)
    scale_kernel<float><<<grid_for(n / 4 + 1), kThreads, 0, as_stream(stream)>>>((float*)dst, n, (float)alpha);
  else if (dtype == MGB_F64)
    scale_kernel<double><<<grid_for(n / 2 + 1), kThreads, 0, as_stream(stream)>>>((double*)dst, n, alpha);
  else
    return fail("mgb_scale: bad dtype %d", dtype);
  MGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mgb_copy_cast(int src_dtype, int dst_dtype, size_t n, const void* src, void* dst,
                             void* stream) {
  if (n == 0) return 0;
  if (!dst || !src) return fail("mgb_copy_cast: null");
  cudaStream_t st = as_stream(stream);
  if (src_dtype == dst_dtype) {
    if (src_dtype != MGB_F32 && src_dtype != MGB_F64) return fail("mgb_copy_cast: bad dtype");
    size_t bytes = n * (src_dtype == MGB_F32 ? 4 : 8);
    MGB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
  if (!aligned && src_dtype == MGB_F32 && dst_dtype == MGB_F64)
    cast_scalar_kernel<float, double><<<grid_for(n), kThreads, 0, st>>>((const float*)src, (double*)dst, n);
  else if (!aligned && src_dtype == MGB_F64 && dst_dtype == MGB_F32)
    cast_scalar_kernel<double, float><<<grid_for(n), kThreads, 0, st>>>((const double*)src, (float*)dst, n);
  else if (src_dtype == MGB_F32 && dst_dtype == MGB_F64)
    cast_kernel<float, double><<<grid_for(n / 4 + 1), kThreads, 0, st>>>((const float*)src, (double*)dst, n);
  else if (src_dtype == MGB_F64 && dst_dtype == MGB_F32)
    cast_kernel<double, float><<<grid_for(n / 4 + 1), kThreads, 0, st>>>((const double*)src, (float*)dst, n);
  else
    return fail("mgb_copy_cast: bad dtypes %d -> %d", src_dtype, dst_dtype);
  MGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mgb_reduce_broadcast(int dtype, int g_ndim, const int64_t* g_shape, int v_ndim,
                                    const int64_t* v_shape, const void* gsrc, void* out,
                                    int accumulate, void* stream) {
  if (dtype != MGB_F32 && dtype != MGB_F64) return fail("reduce_broadcast: bad dtype %d", dtype);
  if (g_ndim < 0 || g_ndim > MGB_MAX_DIMS || v_ndim < 0 || v_ndim > g_ndim)
    return fail("reduce_broadcast: bad ranks g=%d v=%d (max %d)", g_ndim, v_ndim, MGB_MAX_DIMS);
  // classify every grad axis as kept / reduced
  long long sz[MGB_MAX_DIMS], st[MGB_MAX_DIMS];
  bool red[MGB_MAX_DIMS];
  long long n_out = 1, n_red = 1, stride = 1;
  for (int i = g_ndim - 1; i >= 0; --i) {
    sz[i] = g_shape[i];
    st[i] = stride;
    stride *= g_shape[i];
    int vi = i - (g_ndim - v_ndim);
    long long v = (vi >= 0) ? v_shape[vi] : -1;
    if (vi < 0) red[i] = true;
    else if (v == sz[i]) red[i] = false;
    else if (v == 1) red[i] = true;
    else return fail("reduce_broadcast: var axis %d (%lld) is not broadcastable to %lld", vi, v, sz[i]);
    if (red[i]) n_red *= sz[i]; else n_out *= sz[i];
  }
  cudaStream_t stm = as_stream(stream);
  size_t esz = dtype == MGB_F32 ? 4 : 8;
  if (n_out == 0) return 0;
  if (n_red == 0) {  // summing an empty axis gives zeros
    if (!accumulate) MGB_CUDA(cudaMemsetAsync(out, 0, (size_t)n_out * esz, stm));
    return 0;
  }
  if (!gsrc || !out) return fail("reduce_broadcast: null buffer");
  if (n_red == 1) {  // nothing to sum: plain copy / accumulate
    if (accumulate) return mgb_axpy(dtype, (size_t)n_out, gsrc, out, stream);
    return mgb_copy_cast(dtype, dtype, (size_t)n_out, gsrc, out, stream);
  }
  // merge neighbouring axes of the same kind (size-1 axes join either side)
  RGroups g;
  g.nk = g.nr = 0;
  g.n_out = n_out;
  g.n_red = n_red;
  int last_kind = -1;  // 0 kept, 1 reduced
  for (int i = 0; i < g_ndim; ++i) {
    if (sz[i] == 1) continue;
    int kind = red[i] ? 1 : 0;
    if (kind == last_kind) {
      if (kind) { g.rsize[g.nr - 1] *= sz[i]; g.rstride[g.nr - 1] = st[i]; }
      else      { g.ksize[g.nk - 1] *= sz[i]; g.kstride[g.nk - 1] = st[i]; }
    } else {
      if (kind) { g.rsize[g.nr] = sz[i]; g.rstride[g.nr] = st[i]; ++g.nr; }
      else      { g.ksize[g.nk] = sz[i]; g.kstride[g.nk] = st[i]; ++g.nk; }
      last_kind = kind;
    }
  }
  bool inner_reduced = (last_kind == 1);
  if (dtype == MGB_F32)
    return reduce_t<float>(g, inner_reduced, (const float*)gsrc, (float*)out, accumulate, stm);
  return reduce_t<double>(g, inner_reduced, (const double*)gsrc, (double*)out, accumulate, stm);
}
