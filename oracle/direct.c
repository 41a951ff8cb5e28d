/*
 * oracle/direct.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement (direct nested loops, float64 accumulation) of the arithmetic
 * of MyGrad's conv_nd / max_pool hot path.  It is the "second opinion" beside the
 * NumPy restatement in oracle/conv_pool_oracle.py and is only ever loaded by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg — never by mygrad_b200.
 *
 * Definitions followed (reference: rsokl/MyGrad, paths relative to its checkout):
 *   conv fwd   y[n,f,g]  = sum_{c,k} w[f,c,k] * xpad[n,c,g*s+k*d]      src/mygrad/nnet/layers/conv.py:81-103
 *   conv dX    dx[n,c,i] = sum_{f,k,g: g*s+k*d=i+p} dy[n,f,g]*w[f,c,k]  conv.py:105-138
 *   conv dW    dw[f,c,k] = sum_{n,g} dy[n,f,g] * xpad[n,c,g*s+k*d]      conv.py:140-155
 *   (the brute-force formulation is the one of tests/nnet/layers/test_conv.py:100-176)
 *   max-pool fwd / argmax-scatter bwd                                  src/mygrad/nnet/layers/pooling.py:88-149
 *
 * All tensors are C-contiguous; spatial rank is normalised to 3 by the caller
 * (leading extents 1).  Inputs/outputs are float64 here; the Python wrapper casts.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int64_t N, C, F;
  int64_t X[3], W[3], S[3], P[3], D[3], G[3];
} conv_geom;

static inline int64_t xidx(const conv_geom* g, int64_t n, int64_t c, int64_t i0, int64_t i1, int64_t i2) {
  return (((n * g->C + c) * g->X[0] + i0) * g->X[1] + i1) * g->X[2] + i2;
}
static inline int64_t widx(const conv_geom* g, int64_t f, int64_t c, int64_t k0, int64_t k1, int64_t k2) {
  return (((f * g->C + c) * g->W[0] + k0) * g->W[1] + k1) * g->W[2] + k2;
}
static inline int64_t yidx(const conv_geom* g, int64_t n, int64_t f, int64_t g0, int64_t g1, int64_t g2) {
  return (((n * g->F + f) * g->G[0] + g0) * g->G[1] + g1) * g->G[2] + g2;
}

void oracle_conv_fwd(const conv_geom* g, const double* x, const double* w, double* y) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t n = 0; n < g->N; ++n)
    for (int64_t f = 0; f < g->F; ++f)
      for (int64_t g0 = 0; g0 < g->G[0]; ++g0)
        for (int64_t g1 = 0; g1 < g->G[1]; ++g1)
          for (int64_t g2 = 0; g2 < g->G[2]; ++g2) {
            double acc = 0.0;
            for (int64_t c = 0; c < g->C; ++c)
              for (int64_t k0 = 0; k0 < g->W[0]; ++k0) {
                int64_t i0 = g0 * g->S[0] + k0 * g->D[0] - g->P[0];
                if (i0 < 0 || i0 >= g->X[0]) continue;
                for (int64_t k1 = 0; k1 < g->W[1]; ++k1) {
                  int64_t i1 = g1 * g->S[1] + k1 * g->D[1] - g->P[1];
                  if (i1 < 0 || i1 >= g->X[1]) continue;
                  for (int64_t k2 = 0; k2 < g->W[2]; ++k2) {
                    int64_t i2 = g2 * g->S[2] + k2 * g->D[2] - g->P[2];
                    if (i2 < 0 || i2 >= g->X[2]) continue;
                    acc += w[widx(g, f, c, k0, k1, k2)] * x[xidx(g, n, c, i0, i1, i2)];
                  }
                }
              }
            y[yidx(g, n, f, g0, g1, g2)] = acc;
          }
}

void oracle_conv_dx(const conv_geom* g, const double* dy, const double* w, double* dx) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t n = 0; n < g->N; ++n)
    for (int64_t c = 0; c < g->C; ++c)
      for (int64_t i0 = 0; i0 < g->X[0]; ++i0)
        for (int64_t i1 = 0; i1 < g->X[1]; ++i1)
          for (int64_t i2 = 0; i2 < g->X[2]; ++i2) {
            double acc = 0.0;
            for (int64_t f = 0; f < g->F; ++f)
              for (int64_t k0 = 0; k0 < g->W[0]; ++k0) {
                int64_t t0 = i0 + g->P[0] - k0 * g->D[0];
                if (t0 < 0 || t0 % g->S[0] || t0 / g->S[0] >= g->G[0]) continue;
                for (int64_t k1 = 0; k1 < g->W[1]; ++k1) {
                  int64_t t1 = i1 + g->P[1] - k1 * g->D[1];
                  if (t1 < 0 || t1 % g->S[1] || t1 / g->S[1] >= g->G[1]) continue;
                  for (int64_t k2 = 0; k2 < g->W[2]; ++k2) {
                    int64_t t2 = i2 + g->P[2] - k2 * g->D[2];
                    if (t2 < 0 || t2 % g->S[2] || t2 / g->S[2] >= g->G[2]) continue;
                    acc += dy[yidx(g, n, f, t0 / g->S[0], t1 / g->S[1], t2 / g->S[2])] *
                           w[widx(g, f, c, k0, k1, k2)];
                  }
                }
              }
            dx[xidx(g, n, c, i0, i1, i2)] = acc;
          }
}

void oracle_conv_dw(const conv_geom* g, const double* dy, const double* x, double* dw) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t f = 0; f < g->F; ++f)
    for (int64_t c = 0; c < g->C; ++c)
      for (int64_t k0 = 0; k0 < g->W[0]; ++k0)
        for (int64_t k1 = 0; k1 < g->W[1]; ++k1)
          for (int64_t k2 = 0; k2 < g->W[2]; ++k2) {
            double acc = 0.0;
            for (int64_t n = 0; n < g->N; ++n)
              for (int64_t g0 = 0; g0 < g->G[0]; ++g0) {
                int64_t i0 = g0 * g->S[0] + k0 * g->D[0] - g->P[0];
                if (i0 < 0 || i0 >= g->X[0]) continue;
                for (int64_t g1 = 0; g1 < g->G[1]; ++g1) {
                  int64_t i1 = g1 * g->S[1] + k1 * g->D[1] - g->P[1];
                  if (i1 < 0 || i1 >= g->X[1]) continue;
                  for (int64_t g2 = 0; g2 < g->G[2]; ++g2) {
                    int64_t i2 = g2 * g->S[2] + k2 * g->D[2] - g->P[2];
                    if (i2 < 0 || i2 >= g->X[2]) continue;
                    acc += dy[yidx(g, n, f, g0, g1, g2)] * x[xidx(g, n, c, i0, i1, i2)];
                  }
                }
              }
            dw[widx(g, f, c, k0, k1, k2)] = acc;
          }
}

typedef struct {
  int64_t lead;
  int64_t X[3], P[3], S[3], G[3];
} pool_geom;

/* "v replaces best": strictly greater, or the first NaN (np.max / np.argmax semantics) */
static inline int beats(double v, double best) { return (v > best) || (isnan(v) && !isnan(best)); }

/* y = window max; arg (optional) = flat index, within the pooled extent of x, of the
 * FIRST maximal element in row-major window order (pooling.py:111-134) */
void oracle_max_pool_fwd(const pool_geom* g, const double* x, double* y, int64_t* arg) {
  int64_t xp = g->X[0] * g->X[1] * g->X[2], gp = g->G[0] * g->G[1] * g->G[2];
#pragma omp parallel for schedule(static)
  for (int64_t l = 0; l < g->lead; ++l)
    for (int64_t g0 = 0; g0 < g->G[0]; ++g0)
      for (int64_t g1 = 0; g1 < g->G[1]; ++g1)
        for (int64_t g2 = 0; g2 < g->G[2]; ++g2) {
          int64_t first = ((g0 * g->S[0]) * g->X[1] + g1 * g->S[1]) * g->X[2] + g2 * g->S[2];
          double best = x[l * xp + first];
          int64_t where = first;
          for (int64_t p0 = 0; p0 < g->P[0]; ++p0)
            for (int64_t p1 = 0; p1 < g->P[1]; ++p1)
              for (int64_t p2 = 0; p2 < g->P[2]; ++p2) {
                int64_t at = ((g0 * g->S[0] + p0) * g->X[1] + g1 * g->S[1] + p1) * g->X[2] +
                             g2 * g->S[2] + p2;
                double v = x[l * xp + at];
                if (beats(v, best)) { best = v; where = at; }
              }
          int64_t o = l * gp + (g0 * g->G[1] + g1) * g->G[2] + g2;
          y[o] = best;
          if (arg) arg[o] = where;
        }
}

/* dx (float64 buffer, zero-initialised here) += dy at the argmax, in increasing flat
 * window order — the order np.add.at uses (pooling.py:147-148) */
void oracle_max_pool_bwd(const pool_geom* g, const double* dy, const double* x, double* dx) {
  int64_t xp = g->X[0] * g->X[1] * g->X[2], gp = g->G[0] * g->G[1] * g->G[2];
  int64_t* arg = (int64_t*)malloc(sizeof(int64_t) * (size_t)(g->lead * gp > 0 ? g->lead * gp : 1));
  double* y = (double*)malloc(sizeof(double) * (size_t)(g->lead * gp > 0 ? g->lead * gp : 1));
  oracle_max_pool_fwd(g, x, y, arg);
  memset(dx, 0, sizeof(double) * (size_t)(g->lead * xp));
#pragma omp parallel for schedule(static)
  for (int64_t l = 0; l < g->lead; ++l)
    for (int64_t o = 0; o < gp; ++o) dx[l * xp + arg[l * gp + o]] += dy[l * gp + o];
  free(arg);
  free(y);
}
