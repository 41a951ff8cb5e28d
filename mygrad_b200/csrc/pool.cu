// N-d max / mean pooling, forward and backward (HBM-bound, no atomics).
//
// Reference semantics restated (src/mygrad/nnet/layers/pooling.py):
//   fwd  (:88-98)   out[lead, g] = max over the P-shaped window at g*stride of the
//                   trailing axes; NaN propagates (NumPy max).
//   bwd  (:100-149) argmax = FIRST maximal element in row-major window order
//                   (np.argmax on the flattened window, NaN counts as max);
//                   dx is a float64 zero buffer that receives np.add.at in
//                   increasing flat window order, then is cast to x.dtype by
//                   Operation.backward (operation_base.py:223-224).
// The backward here is a gather: every dx element is produced by exactly one
// thread which visits the windows covering it in increasing flat window order,
// so the float64 accumulation order is the reference's, bit for bit.
#include "common.cuh"

namespace mgb {
namespace {

struct PoolGeom {
  int X[3], P[3], S[3], G[3];
  long long lead;
  long long x_per_lead;  // X0*X1*X2
  long long g_per_lead;  // G0*G1*G2
};

// "takes over as the maximum": strictly greater, or the first NaN.
template <typename T>
__device__ __forceinline__ bool beats(T v, T best) {
  return (v > best) || ((v != v) && !(best != best));
}

template <typename T> struct Vec2;
template <> struct Vec2<float> { using type = float2; };
template <> struct Vec2<double> { using type = double2; };

// ---- forward ---------------------------------------------------------------

template <typename T, int MODE>
__global__ void __launch_bounds__(256) pool_fwd_generic(PoolGeom g, const T* __restrict__ x,
                                                        T* __restrict__ y, long long total) {
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total;
       o += (long long)gridDim.x * blockDim.x) {
    long long l = o / g.g_per_lead;
    int r = (int)(o - l * g.g_per_lead);
    int g2 = r % g.G[2];
    r /= g.G[2];
    int g1 = r % g.G[1];
    int g0 = r / g.G[1];
    const T* base = x + l * g.x_per_lead +
                    ((long long)(g0 * g.S[0]) * g.X[1] + g1 * g.S[1]) * g.X[2] + g2 * g.S[2];
    if (MODE == 0) {
      T best = base[0];
      for (int p0 = 0; p0 < g.P[0]; ++p0)
        for (int p1 = 0; p1 < g.P[1]; ++p1) {
          const T* row = base + ((long long)p0 * g.X[1] + p1) * g.X[2];
          for (int p2 = 0; p2 < g.P[2]; ++p2) {
            T v = row[p2];
            if (beats(v, best)) best = v;
          }
        }
      y[o] = best;
    } else {
      T acc = 0;
      for (int p0 = 0; p0 < g.P[0]; ++p0)
        for (int p1 = 0; p1 < g.P[1]; ++p1) {
          const T* row = base + ((long long)p0 * g.X[1] + p1) * g.X[2];
          for (int p2 = 0; p2 < g.P[2]; ++p2) acc += row[p2];
        }
      y[o] = acc / (T)(g.P[0] * g.P[1] * g.P[2]);
    }
  }
}

// Last pooled axis has P = S = 2: every thread reads its window rows as aligned
// 2-vectors (X2 is even, so each row pair is vector aligned).
template <typename T>
__global__ void __launch_bounds__(256) pool_max_fwd_pair(PoolGeom g, const T* __restrict__ x,
                                                         T* __restrict__ y, long long total) {
  using V = typename Vec2<T>::type;
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total;
       o += (long long)gridDim.x * blockDim.x) {
    long long l = o / g.g_per_lead;
    int r = (int)(o - l * g.g_per_lead);
    int g2 = r % g.G[2];
    r /= g.G[2];
    int g1 = r % g.G[1];
    int g0 = r / g.G[1];
    const T* base = x + l * g.x_per_lead +
                    ((long long)(g0 * g.S[0]) * g.X[1] + g1 * g.S[1]) * g.X[2] + g2 * 2;
    T best = base[0];
    for (int p0 = 0; p0 < g.P[0]; ++p0)
      for (int p1 = 0; p1 < g.P[1]; ++p1) {
        V v = __ldg(reinterpret_cast<const V*>(base + ((long long)p0 * g.X[1] + p1) * g.X[2]));
        if (beats(v.x, best)) best = v.x;
        if (beats(v.y, best)) best = v.y;
      }
    y[o] = best;
  }
}

// Same case (last axis pooled in pairs) with VW outputs per thread: every window row is
// 2*VW contiguous elements = two 16-byte loads, all issued before any compare, so a
// thread keeps 32*P0*P1 bytes in flight (HBM latency needs the memory-level parallelism).
template <typename T, int VW>
__global__ void __launch_bounds__(256) pool_max_fwd_vec(PoolGeom g, const T* __restrict__ x,
                                                        T* __restrict__ y, long long total_vec) {
  constexpr int EPV = 16 / sizeof(T);          // elements per 16-byte vector
  constexpr int NV = 2 * VW / EPV;             // vectors per window row (== 2)
  const int gv = g.G[2] / VW;                  // vector groups per output row
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total_vec;
       o += (long long)gridDim.x * blockDim.x) {
    long long t = o;
    int v2 = (int)(t % gv); t /= gv;
    int g1 = (int)(t % g.G[1]); t /= g.G[1];
    int g0 = (int)(t % g.G[0]);
    long long l = t / g.G[0];
    const T* base = x + l * g.x_per_lead +
                    ((long long)(g0 * g.S[0]) * g.X[1] + g1 * g.S[1]) * g.X[2] + v2 * (2 * VW);
    T best[VW];
    bool first = true;
    for (int p0 = 0; p0 < g.P[0]; ++p0)
      for (int p1 = 0; p1 < g.P[1]; ++p1) {
        float4 raw[NV];
        const float4* row = reinterpret_cast<const float4*>(base + ((long long)p0 * g.X[1] + p1) * g.X[2]);
#pragma unroll
        for (int k = 0; k < NV; ++k) raw[k] = __ldg(row + k);
        const T* e = reinterpret_cast<const T*>(raw);
#pragma unroll
        for (int w = 0; w < VW; ++w) {
          T a = e[2 * w], b = e[2 * w + 1];
          if (first) best[w] = a; else if (beats(a, best[w])) best[w] = a;
          if (beats(b, best[w])) best[w] = b;
        }
        first = false;
      }
    float4 out;
    T* oe = reinterpret_cast<T*>(&out);
#pragma unroll
    for (int w = 0; w < VW; ++w) oe[w] = best[w];
    T* dst = y + l * g.g_per_lead + ((long long)g0 * g.G[1] + g1) * g.G[2] + v2 * VW;
    if (VW * sizeof(T) == 16) *reinterpret_cast<float4*>(dst) = out;
    else *reinterpret_cast<float2*>(dst) = *reinterpret_cast<float2*>(&out);
  }
}

// Rows that are only 2-vector aligned (X2 = 62 floats in config 4: 248-byte rows) cannot use the 16-byte
// kernel above, and the one-window pair kernel keeps just P0*P1 eight-byte loads in flight per thread (0.53 of
// HBM).  Here a thread owns VW consecutive windows of one output row and issues all VW * P0 * P1 two-element
// loads before the first compare; the last group of a row is ragged (predicated).
template <typename T, int VW, int MAXROWS>
__global__ void __launch_bounds__(256) pool_max_fwd_pairs(PoolGeom g, const T* __restrict__ x,
                                                          T* __restrict__ y, long long total_groups) {
  using V = typename Vec2<T>::type;
  const int gv = (g.G[2] + VW - 1) / VW;
  const int rows = g.P[0] * g.P[1];            // <= MAXROWS (checked on the host)
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total_groups;
       o += (long long)gridDim.x * blockDim.x) {
    long long t = o;
    const int v2 = (int)(t % gv); t /= gv;
    const int g1 = (int)(t % g.G[1]); t /= g.G[1];
    const int g0 = (int)(t % g.G[0]);
    const long long l = t / g.G[0];
    // the thread's windows are v2, v2 + gv, v2 + 2 gv, ...: neighbouring threads touch neighbouring windows
    // in every load / store instruction (whole 32-byte sectors), not VW windows apart
    const T* base = x + l * g.x_per_lead +
                    ((long long)(g0 * g.S[0]) * g.X[1] + g1 * g.S[1]) * g.X[2];
    V raw[MAXROWS][VW];
#pragma unroll
    for (int r = 0; r < MAXROWS; ++r)
      if (r < rows) {
        const int p0 = r / g.P[1], p1 = r - p0 * g.P[1];
        const V* row = reinterpret_cast<const V*>(base + ((long long)p0 * g.X[1] + p1) * g.X[2]);
#pragma unroll
        for (int w = 0; w < VW; ++w)
          if (v2 + w * gv < g.G[2]) raw[r][w] = __ldg(row + v2 + w * gv);
      }
    T* dst = y + l * g.g_per_lead + ((long long)g0 * g.G[1] + g1) * g.G[2];
#pragma unroll
    for (int w = 0; w < VW; ++w)
      if (v2 + w * gv < g.G[2]) {
        T best = raw[0][w].x;
#pragma unroll
        for (int r = 0; r < MAXROWS; ++r)
          if (r < rows) {
            if (r > 0 && beats(raw[r][w].x, best)) best = raw[r][w].x;
            if (beats(raw[r][w].y, best)) best = raw[r][w].y;
          }
        dst[v2 + w * gv] = best;
      }
  }
}

// ---- backward --------------------------------------------------------------

// window-centric, for exact non-overlapping tilings (S == P on every axis) whose
// last axis is pooled in pairs: one thread owns one window and writes all of it.
template <typename T>
__global__ void __launch_bounds__(256) pool_max_bwd_tile_pair(PoolGeom g, const T* __restrict__ dy,
                                                              const T* __restrict__ x,
                                                              T* __restrict__ dx, long long total) {
  using V = typename Vec2<T>::type;
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total;
       o += (long long)gridDim.x * blockDim.x) {
    long long l = o / g.g_per_lead;
    int r = (int)(o - l * g.g_per_lead);
    int g2 = r % g.G[2];
    r /= g.G[2];
    int g1 = r % g.G[1];
    int g0 = r / g.G[1];
    long long off = l * g.x_per_lead +
                    ((long long)(g0 * g.P[0]) * g.X[1] + g1 * g.P[1]) * g.X[2] + g2 * 2;
    const T* base = x + off;
    T best = base[0];
    int arg = 0;
    int idx = 0;
    for (int p0 = 0; p0 < g.P[0]; ++p0)
      for (int p1 = 0; p1 < g.P[1]; ++p1) {
        V v = __ldg(reinterpret_cast<const V*>(base + ((long long)p0 * g.X[1] + p1) * g.X[2]));
        if (beats(v.x, best)) { best = v.x; arg = idx; }
        if (beats(v.y, best)) { best = v.y; arg = idx + 1; }
        idx += 2;
      }
    // float64 scatter buffer then astype(x.dtype): a single contribution is exact
    T gval = dy[o];
    idx = 0;
    for (int p0 = 0; p0 < g.P[0]; ++p0)
      for (int p1 = 0; p1 < g.P[1]; ++p1) {
        V out;
        out.x = (arg == idx) ? gval : (T)0;
        out.y = (arg == idx + 1) ? gval : (T)0;
        *reinterpret_cast<V*>(dx + off + ((long long)p0 * g.X[1] + p1) * g.X[2]) = out;
        idx += 2;
      }
  }
}

// window-centric, exact tiling, rows only 2-vector aligned: VW windows per thread, all loads first (see
// pool_max_fwd_pairs); the first maximum in row-major window order takes dy, the rest of the window zeros
template <typename T, int VW, int MAXROWS>
__global__ void __launch_bounds__(256) pool_max_bwd_tile_pairs(PoolGeom g, const T* __restrict__ dy,
                                                               const T* __restrict__ x, T* __restrict__ dx,
                                                               long long total_groups) {
  using V = typename Vec2<T>::type;
  const int gv = (g.G[2] + VW - 1) / VW;
  const int rows = g.P[0] * g.P[1];
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total_groups;
       o += (long long)gridDim.x * blockDim.x) {
    long long t = o;
    const int v2 = (int)(t % gv); t /= gv;
    const int g1 = (int)(t % g.G[1]); t /= g.G[1];
    const int g0 = (int)(t % g.G[0]);
    const long long l = t / g.G[0];
    const long long off = l * g.x_per_lead +
                          ((long long)(g0 * g.P[0]) * g.X[1] + g1 * g.P[1]) * g.X[2];
    V raw[MAXROWS][VW];
#pragma unroll
    for (int r = 0; r < MAXROWS; ++r)
      if (r < rows) {
        const int p0 = r / g.P[1], p1 = r - p0 * g.P[1];
        const V* row = reinterpret_cast<const V*>(x + off + ((long long)p0 * g.X[1] + p1) * g.X[2]);
#pragma unroll
        for (int w = 0; w < VW; ++w)
          if (v2 + w * gv < g.G[2]) raw[r][w] = __ldg(row + v2 + w * gv);
      }
    const T* gsrc = dy + l * g.g_per_lead + ((long long)g0 * g.G[1] + g1) * g.G[2];
    T gval[VW];
    int arg[VW];
#pragma unroll
    for (int w = 0; w < VW; ++w)
      if (v2 + w * gv < g.G[2]) {
        gval[w] = __ldg(gsrc + v2 + w * gv);
        T best = raw[0][w].x;
        arg[w] = 0;
#pragma unroll
        for (int r = 0; r < MAXROWS; ++r)
          if (r < rows) {
            if (r > 0 && beats(raw[r][w].x, best)) { best = raw[r][w].x; arg[w] = 2 * r; }
            if (beats(raw[r][w].y, best)) { best = raw[r][w].y; arg[w] = 2 * r + 1; }
          }
      }
#pragma unroll
    for (int r = 0; r < MAXROWS; ++r)
      if (r < rows) {
        const int p0 = r / g.P[1], p1 = r - p0 * g.P[1];
        V* row = reinterpret_cast<V*>(dx + off + ((long long)p0 * g.X[1] + p1) * g.X[2]);
#pragma unroll
        for (int w = 0; w < VW; ++w)
          if (v2 + w * gv < g.G[2]) {
            V out;
            out.x = (arg[w] == 2 * r) ? gval[w] : (T)0;
            out.y = (arg[w] == 2 * r + 1) ? gval[w] : (T)0;
            row[v2 + w * gv] = out;
          }
      }
  }
}

// window-centric, exact tiling, last axis in pairs, VW windows per thread (see pool_max_fwd_vec)
template <typename T, int VW, int MAXROWS>
__global__ void __launch_bounds__(256) pool_max_bwd_tile_vec(PoolGeom g, const T* __restrict__ dy,
                                                             const T* __restrict__ x,
                                                             T* __restrict__ dx, long long total_vec) {
  constexpr int EPV = 16 / sizeof(T);
  constexpr int NV = 2 * VW / EPV;
  const int gv = g.G[2] / VW;
  const int rows = g.P[0] * g.P[1];            // <= MAXROWS (checked on the host)
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total_vec;
       o += (long long)gridDim.x * blockDim.x) {
    long long t = o;
    int v2 = (int)(t % gv); t /= gv;
    int g1 = (int)(t % g.G[1]); t /= g.G[1];
    int g0 = (int)(t % g.G[0]);
    long long l = t / g.G[0];
    const long long off = l * g.x_per_lead +
                          ((long long)(g0 * g.P[0]) * g.X[1] + g1 * g.P[1]) * g.X[2] + v2 * (2 * VW);
    float4 raw[MAXROWS][NV];
#pragma unroll
    for (int r = 0; r < MAXROWS; ++r)
      if (r < rows) {
        int p0 = r / g.P[1], p1 = r - p0 * g.P[1];
        const float4* row = reinterpret_cast<const float4*>(x + off + ((long long)p0 * g.X[1] + p1) * g.X[2]);
#pragma unroll
        for (int k = 0; k < NV; ++k) raw[r][k] = __ldg(row + k);
      }
    float4 graw;
    const T* gsrc = dy + l * g.g_per_lead + ((long long)g0 * g.G[1] + g1) * g.G[2] + v2 * VW;
    if (VW * sizeof(T) == 16) graw = __ldg(reinterpret_cast<const float4*>(gsrc));
    else *reinterpret_cast<float2*>(&graw) = __ldg(reinterpret_cast<const float2*>(gsrc));
    const T* gv_ = reinterpret_cast<const T*>(&graw);
    T best[VW];
    int arg[VW];
#pragma unroll
    for (int r = 0; r < MAXROWS; ++r)
      if (r < rows) {
        const T* e = reinterpret_cast<const T*>(raw[r]);
#pragma unroll
        for (int w = 0; w < VW; ++w) {
          T a = e[2 * w], b = e[2 * w + 1];
          if (r == 0) { best[w] = a; arg[w] = 0; }
          else if (beats(a, best[w])) { best[w] = a; arg[w] = 2 * r; }
          if (beats(b, best[w])) { best[w] = b; arg[w] = 2 * r + 1; }
        }
      }
#pragma unroll
    for (int r = 0; r < MAXROWS; ++r)
      if (r < rows) {
        float4 out[NV];
        T* oe = reinterpret_cast<T*>(out);
#pragma unroll
        for (int w = 0; w < VW; ++w) {
          oe[2 * w] = (arg[w] == 2 * r) ? gv_[w] : (T)0;
          oe[2 * w + 1] = (arg[w] == 2 * r + 1) ? gv_[w] : (T)0;
        }
        int p0 = r / g.P[1], p1 = r - p0 * g.P[1];
        float4* row = reinterpret_cast<float4*>(dx + off + ((long long)p0 * g.X[1] + p1) * g.X[2]);
#pragma unroll
        for (int k = 0; k < NV; ++k) row[k] = out[k];
      }
  }
}

__device__ __forceinline__ int ceil_div_pos(int a, int b) {  // ceil(a/b) for b>0, any a
  return (a >= 0) ? (a + b - 1) / b : -((-a) / b);
}

// input-centric gather: any pool/stride (overlaps, gaps), max or mean.
template <typename T, int MODE>
__global__ void __launch_bounds__(256) pool_bwd_gather(PoolGeom g, const T* __restrict__ dy,
                                                       const T* __restrict__ x,
                                                       T* __restrict__ dx, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long l = i / g.x_per_lead;
    int r = (int)(i - l * g.x_per_lead);
    int c2 = r % g.X[2];
    r /= g.X[2];
    int c1 = r % g.X[1];
    int c0 = r / g.X[1];
    int lo0 = max(0, ceil_div_pos(c0 - g.P[0] + 1, g.S[0])), hi0 = min(g.G[0] - 1, c0 / g.S[0]);
    int lo1 = max(0, ceil_div_pos(c1 - g.P[1] + 1, g.S[1])), hi1 = min(g.G[1] - 1, c1 / g.S[1]);
    int lo2 = max(0, ceil_div_pos(c2 - g.P[2] + 1, g.S[2])), hi2 = min(g.G[2] - 1, c2 / g.S[2]);
    const T* xl = x + l * g.x_per_lead;
    const T* dyl = dy + l * g.g_per_lead;
    double acc = 0.0;
    for (int g0 = lo0; g0 <= hi0; ++g0)
      for (int g1 = lo1; g1 <= hi1; ++g1)
        for (int g2 = lo2; g2 <= hi2; ++g2) {
          T gval = dyl[((long long)g0 * g.G[1] + g1) * g.G[2] + g2];
          if (MODE == 1) {
            acc += (double)gval;
            continue;
          }
          const T* base = xl + ((long long)(g0 * g.S[0]) * g.X[1] + g1 * g.S[1]) * g.X[2] + g2 * g.S[2];
          T best = base[0];
          int a0 = 0, a1 = 0, a2 = 0;
          for (int p0 = 0; p0 < g.P[0]; ++p0)
            for (int p1 = 0; p1 < g.P[1]; ++p1) {
              const T* row = base + ((long long)p0 * g.X[1] + p1) * g.X[2];
              for (int p2 = 0; p2 < g.P[2]; ++p2) {
                T v = row[p2];
                if (beats(v, best)) { best = v; a0 = p0; a1 = p1; a2 = p2; }
              }
            }
          if (g0 * g.S[0] + a0 == c0 && g1 * g.S[1] + a1 == c1 && g2 * g.S[2] + a2 == c2)
            acc += (double)gval;
        }
    if (MODE == 1) acc /= (double)(g.P[0] * g.P[1] * g.P[2]);
    dx[i] = (T)acc;
  }
}

int make_geom(const mgb_pool_desc* d, PoolGeom* out) {
  if (!d) return fail("pool: null descriptor");
  if (d->nd < 1 || d->nd > MGB_MAX_SPATIAL) return fail("pool: nd=%d unsupported (1..3)", d->nd);
  if (d->dtype != MGB_F32 && d->dtype != MGB_F64) return fail("pool: bad dtype %d", d->dtype);
  if (d->mode != 0 && d->mode != 1) return fail("pool: bad mode %d", d->mode);
  PoolGeom g;
  int shift = 3 - d->nd;
  for (int i = 0; i < 3; ++i) { g.X[i] = 1; g.P[i] = 1; g.S[i] = 1; g.G[i] = 1; }
  for (int i = 0; i < d->nd; ++i) {
    int64_t X = d->X[i], P = d->P[i], S = d->stride[i];
    if (P < 1 || S < 1 || X < P || (X - P) % S != 0)
      return fail("pool: incompatible axis %d: X=%lld P=%lld stride=%lld", i, (long long)X,
                  (long long)P, (long long)S);
    if (X > INT32_MAX) return fail("pool: axis too long");
    g.X[i + shift] = (int)X; g.P[i + shift] = (int)P; g.S[i + shift] = (int)S;
    g.G[i + shift] = (int)((X - P) / S + 1);
  }
  g.lead = d->lead;
  g.x_per_lead = (long long)g.X[0] * g.X[1] * g.X[2];
  g.g_per_lead = (long long)g.G[0] * g.G[1] * g.G[2];
  if (g.x_per_lead > INT32_MAX) return fail("pool: pooled extent too large");
  *out = g;
  return 0;
}

inline int grid_for(long long total, int threads) {
  long long blocks = (total + threads - 1) / threads;
  long long cap = (long long)sm_count() * 64;  // grid-stride beyond 64 CTAs per SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

template <typename T>
int pool_fwd_t(const PoolGeom& g, int mode, const T* x, T* y, cudaStream_t st) {
  long long total = g.lead * g.g_per_lead;
  if (total == 0) return 0;
  int grid = grid_for(total, 256);
  bool pair = mode == 0 && g.P[2] == 2 && g.S[2] == 2 && (g.X[2] % 2 == 0) &&
              ((uintptr_t)x % (2 * sizeof(T)) == 0);
  constexpr int VW = 16 / sizeof(T);   // outputs per thread of the vector kernel: 32 input bytes per row
  bool vec = pair && (g.X[2] % (2 * VW) == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0);
  if (vec) {
    long long tv = total / VW;
    pool_max_fwd_vec<T, VW><<<grid_for(tv, 256), 256, 0, st>>>(g, x, y, tv);
  } else if (pair && g.P[0] * g.P[1] <= 4) {
    const long long groups = g.lead * g.G[0] * g.G[1] * ((g.G[2] + 3) / 4);
    pool_max_fwd_pairs<T, 4, 4><<<grid_for(groups, 256), 256, 0, st>>>(g, x, y, groups);
  } else if (pair)
    pool_max_fwd_pair<T><<<grid, 256, 0, st>>>(g, x, y, total);
  else if (mode == 0)
    pool_fwd_generic<T, 0><<<grid, 256, 0, st>>>(g, x, y, total);
  else
    pool_fwd_generic<T, 1><<<grid, 256, 0, st>>>(g, x, y, total);
  MGB_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int pool_bwd_t(const PoolGeom& g, int mode, const T* dy, const T* x, T* dx, cudaStream_t st) {
  long long total_x = g.lead * g.x_per_lead;
  long long total_g = g.lead * g.g_per_lead;
  if (total_x == 0) return 0;
  bool tiled = g.P[0] == g.S[0] && g.P[1] == g.S[1] && g.P[2] == g.S[2];
  bool pair = mode == 0 && tiled && g.P[2] == 2 && (g.X[2] % 2 == 0) &&
              ((uintptr_t)x % (2 * sizeof(T)) == 0) && ((uintptr_t)dx % (2 * sizeof(T)) == 0);
  constexpr int VW = 16 / sizeof(T);
  bool vec = pair && (g.X[2] % (2 * VW) == 0) && g.P[0] * g.P[1] <= 4 && ((uintptr_t)x % 16 == 0) &&
             ((uintptr_t)dx % 16 == 0) && ((uintptr_t)dy % 16 == 0);
  if (vec) {
    long long tv = total_g / VW;
    pool_max_bwd_tile_vec<T, VW, 4><<<grid_for(tv, 256), 256, 0, st>>>(g, dy, x, dx, tv);
  } else if (pair && g.P[0] * g.P[1] <= 4) {
    const long long groups = g.lead * g.G[0] * g.G[1] * ((g.G[2] + 3) / 4);
    pool_max_bwd_tile_pairs<T, 4, 4><<<grid_for(groups, 256), 256, 0, st>>>(g, dy, x, dx, groups);
  } else if (pair) {
    pool_max_bwd_tile_pair<T><<<grid_for(total_g, 256), 256, 0, st>>>(g, dy, x, dx, total_g);
  } else if (mode == 0) {
    pool_bwd_gather<T, 0><<<grid_for(total_x, 256), 256, 0, st>>>(g, dy, x, dx, total_x);
  } else {
    pool_bwd_gather<T, 1><<<grid_for(total_x, 256), 256, 0, st>>>(g, dy, x, dx, total_x);
  }
  MGB_LAUNCH_CHECK();
  return 0;
}

}  // namespace
}  // namespace mgb

using namespace mgb;

extern "C" int mgb_pool_fwd(const mgb_pool_desc* d, const void* x, void* y, void* stream) {
  PoolGeom g;
  if (make_geom(d, &g)) return 1;
  if (g.lead * g.g_per_lead > 0 && (!x || !y)) return fail("mgb_pool_fwd: null buffer");
  if (d->dtype == MGB_F32)
    return pool_fwd_t<float>(g, d->mode, (const float*)x, (float*)y, as_stream(stream));
  return pool_fwd_t<double>(g, d->mode, (const double*)x, (double*)y, as_stream(stream));
}

extern "C" int mgb_pool_bwd(const mgb_pool_desc* d, const void* dy, const void* x, void* dx,
                            void* stream) {
  PoolGeom g;
  if (make_geom(d, &g)) return 1;
  if (g.lead * g.x_per_lead > 0 && (!dy || !x || !dx)) return fail("mgb_pool_bwd: null buffer");
  if (d->dtype == MGB_F32)
    return pool_bwd_t<float>(g, d->mode, (const float*)dy, (const float*)x, (float*)dx,
                             as_stream(stream));
  return pool_bwd_t<double>(g, d->mode, (const double*)dy, (const double*)x, (double*)dx,
                            as_stream(stream));
}
