// mgb_einsum: Einstein summation over strided operands.
//
// The reference hands every einsum — forward and the per-operand gradient contractions of
// EinSum.backward_var (src/mygrad/linalg/ops.py:79-243) — to numpy.einsum.  Here ONE kernel
// evaluates the general form directly:
//
//     out[o] (=|+=) sum_s  prod_k  op_k[ base_k(o) + off_k(s) ]
//
// where o runs over the OUTPUT labels and s over the SUMMED labels.  Every operand, and the output,
// is addressed through one element stride per label; the host (mygrad_b200/linalg.py) builds them:
//   * a label repeated inside one operand (trace / diagonal, "ii") gets the SUM of its axis strides;
//   * a label the operand does not carry, or carries with extent 1 (ellipsis broadcasting), gets 0;
//   * the same holds for the output, which is how a gradient is written along the diagonals of a
//     zero-filled array ("jk,k->ijkji" in the reference, ops.py:215-236) and how a summed-away
//     label is broadcast back (ops.py:168-190).
// Products and sums run in float64 (the result is rounded once), so float32 results are at least as
// accurate as NumPy's.  Small outputs with long sums use one CTA per output element.
#include "common.cuh"

namespace mgb {
namespace {

constexpr int kMaxOps = 8;
constexpr int kMaxLabels = 16;

struct EinsumArgs {
  int nops, nout, nsum;
  long long out_size[kMaxLabels], sum_size[kMaxLabels];
  long long out_stride[kMaxLabels];                 // of the output tensor, per output label
  long long op_out_stride[kMaxOps][kMaxLabels];     // of operand k, per output label
  long long op_sum_stride[kMaxOps][kMaxLabels];     // of operand k, per summed label
  const void* ops[kMaxOps];
  void* out;
  long long total_out, total_sum;
};

template <typename T>
__device__ __forceinline__ double term_sum(const EinsumArgs& a, const long long* base, long long s0, long long s1,
                                           long long step) {
  double acc = 0.0;
  for (long long s = s0; s < s1; s += step) {
    long long r = s;
    long long off[kMaxOps];
#pragma unroll
    for (int k = 0; k < kMaxOps; ++k) off[k] = k < a.nops ? base[k] : 0;
    for (int l = a.nsum - 1; l >= 0; --l) {
      const long long i = r % a.sum_size[l];
      r /= a.sum_size[l];
      for (int k = 0; k < a.nops; ++k) off[k] += i * a.op_sum_stride[k][l];
    }
    double p = 1.0;
    for (int k = 0; k < a.nops; ++k) p *= (double)static_cast<const T*>(a.ops[k])[off[k]];
    acc += p;
  }
  return acc;
}

__device__ __forceinline__ void decode_out(const EinsumArgs& a, long long o, long long* base, long long* oo) {
  for (int k = 0; k < kMaxOps; ++k) base[k] = 0;
  long long out_off = 0;
  for (int l = a.nout - 1; l >= 0; --l) {
    const long long i = o % a.out_size[l];
    o /= a.out_size[l];
    out_off += i * a.out_stride[l];
    for (int k = 0; k < a.nops; ++k) base[k] += i * a.op_out_stride[k][l];
  }
  *oo = out_off;
}

// one thread per output element
template <typename T>
__global__ void __launch_bounds__(256) einsum_thread_kernel(const EinsumArgs a) {
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < a.total_out;
       o += (long long)gridDim.x * blockDim.x) {
    long long base[kMaxOps], oo;
    decode_out(a, o, base, &oo);
    static_cast<T*>(a.out)[oo] = (T)term_sum<T>(a, base, 0, a.total_sum, 1);
  }
}

// one CTA per output element: long sums, few outputs (full reductions, traces, inner products)
template <typename T>
__global__ void __launch_bounds__(256) einsum_block_kernel(const EinsumArgs a) {
  __shared__ double part[8];
  for (long long o = blockIdx.x; o < a.total_out; o += gridDim.x) {
    long long base[kMaxOps], oo;
    decode_out(a, o, base, &oo);
    double acc = term_sum<T>(a, base, threadIdx.x, a.total_sum, blockDim.x);
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += part[w];     // fixed order: deterministic
      static_cast<T*>(a.out)[oo] = (T)t;
    }
    __syncthreads();
  }
}

}  // namespace
}  // namespace mgb

using namespace mgb;

// strides are in ELEMENTS; layout of the stride tables: [operand][label], row length = nout resp. nsum
extern "C" int mgb_einsum(int dtype, int nops, const void* const* ops, int nout, const int64_t* out_size,
                          const int64_t* out_stride, const int64_t* op_out_stride, int nsum,
                          const int64_t* sum_size, const int64_t* op_sum_stride, void* out, void* stream) {
  if (dtype != MGB_F32 && dtype != MGB_F64) return fail("mgb_einsum: bad dtype %d", dtype);
  if (nops < 1 || nops > kMaxOps) return fail("mgb_einsum: %d operands unsupported (1..%d)", nops, kMaxOps);
  if (nout < 0 || nout > kMaxLabels || nsum < 0 || nsum > kMaxLabels)
    return fail("mgb_einsum: more than %d output or summed labels", kMaxLabels);
  EinsumArgs a;
  a.nops = nops; a.nout = nout; a.nsum = nsum; a.out = out;
  a.total_out = 1; a.total_sum = 1;
  for (int l = 0; l < nout; ++l) {
    if (out_size[l] < 0) return fail("mgb_einsum: negative extent");
    a.out_size[l] = out_size[l]; a.out_stride[l] = out_stride[l]; a.total_out *= out_size[l];
  }
  for (int l = 0; l < nsum; ++l) {
    if (sum_size[l] < 0) return fail("mgb_einsum: negative extent");
    a.sum_size[l] = sum_size[l]; a.total_sum *= sum_size[l];
  }
  for (int k = 0; k < nops; ++k) {
    a.ops[k] = ops[k];
    for (int l = 0; l < nout; ++l) a.op_out_stride[k][l] = op_out_stride[(size_t)k * nout + l];
    for (int l = 0; l < nsum; ++l) a.op_sum_stride[k][l] = op_sum_stride[(size_t)k * nsum + l];
  }
  if (a.total_out == 0) return 0;
  if (!out) return fail("mgb_einsum: null output");
  for (int k = 0; k < nops; ++k)
    if (!ops[k] && a.total_sum > 0) return fail("mgb_einsum: null operand %d", k);
  cudaStream_t st = as_stream(stream);
  const bool block = a.total_out < 8LL * sm_count() && a.total_sum >= 512;
  if (block) {
    int grid = (int)std::min<long long>(a.total_out, 8LL * sm_count());
    if (dtype == MGB_F32) einsum_block_kernel<float><<<grid, 256, 0, st>>>(a);
    else einsum_block_kernel<double><<<grid, 256, 0, st>>>(a);
  } else {
    int grid = (int)std::max<long long>(1, std::min<long long>((a.total_out + 255) / 256, 16LL * sm_count()));
    if (dtype == MGB_F32) einsum_thread_kernel<float><<<grid, 256, 0, st>>>(a);
    else einsum_thread_kernel<double><<<grid, 256, 0, st>>>(a);
  }
  MGB_LAUNCH_CHECK();
  return 0;
}
