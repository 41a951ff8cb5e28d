/*
 * mgb.h — C ABI of libmgb200.so: the B200 (sm_100a) device path that sits under
 * MyGrad's `Operation` protocol for conv_nd / max_pool and the gradient
 * reductions of Tensor.backward.
 *
 * The reference (rsokl/MyGrad) has no FFI: its hot path is NumPy calls inside
 * Python `Operation` subclasses.  Each entry point below names the reference
 * code it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - every function returns int: 0 = ok, non-zero = error; the text is in
 *     mgb_last_error() (thread-local).  Nothing throws.
 *   - all pointers named x/w/y/dy/dx/dw/src/dst/ws are DEVICE pointers unless
 *     the name starts with `host_`.  Tensors are C-contiguous.
 *   - the library never allocates or retains caller buffers; scratch space is
 *     passed in as (ws, ws_bytes) and sized by the *_workspace_bytes query.
 *   - every launch is asynchronous on `stream` (a cudaStream_t passed as
 *     void*; NULL = the legacy default stream).
 *   - outputs are fully overwritten (no read-modify-write) except mgb_axpy and
 *     mgb_reduce_broadcast(accumulate=1).
 */
#ifndef MGB_H_
#define MGB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGB_MAX_SPATIAL 3
#define MGB_MAX_DIMS 8

typedef enum { MGB_F32 = 0, MGB_F64 = 1 } mgb_dtype;

/* ---- convolution ------------------------------------------------------ */

/* N-d cross-correlation descriptor: x (N,C,X...) * w (F,C,W...) -> y (N,F,G...)
 * with G = (X + 2*pad - ((W-1)*dil + 1)) / stride + 1   (conv.py:63-74).
 * nd = number of spatial dims (1..3); unused trailing entries are ignored. */
typedef struct {
  int32_t dtype;                 /* mgb_dtype */
  int32_t nd;                    /* 1..MGB_MAX_SPATIAL */
  int64_t N, C, F;
  int64_t X[MGB_MAX_SPATIAL];
  int64_t W[MGB_MAX_SPATIAL];
  int64_t stride[MGB_MAX_SPATIAL];
  int64_t pad[MGB_MAX_SPATIAL];
  int64_t dil[MGB_MAX_SPATIAL];
} mgb_conv_desc;

/* which = 0 fwd, 1 dgrad (dX), 2 wgrad (dW) */
int mgb_conv_workspace_bytes(const mgb_conv_desc* d, int which, size_t* bytes);

/* y[n,f,g] = sum_{c,k} w[f,c,k] * xpad[n,c,g*s+k*d]
 * replaces ConvND.__call__  (src/mygrad/nnet/layers/conv.py:15-103) */
int mgb_conv_fwd(const mgb_conv_desc* d, const void* x, const void* w, void* y,
                 void* ws, size_t ws_bytes, void* stream);

/* dx[n,c,i] = sum_{f,k,g : g*s+k*d = i+p} dy[n,f,g] * w[f,c,k]
 * replaces ConvND.backward_var(grad, 0)  (conv.py:105-138) */
int mgb_conv_dgrad(const mgb_conv_desc* d, const void* dy, const void* w, void* dx,
                   void* ws, size_t ws_bytes, void* stream);

/* dw[f,c,k] = sum_{n,g} dy[n,f,g] * xpad[n,c,g*s+k*d]
 * replaces ConvND.backward_var(grad, 1)  (conv.py:140-155) */
int mgb_conv_wgrad(const mgb_conv_desc* d, const void* dy, const void* x, void* dw,
                   void* ws, size_t ws_bytes, void* stream);

/* Optional operand cache of the float32 tensor-core path.  That path reads x and dy as
 * channels-last TF32 hi/lo "planes"; x planes are needed by fwd and wgrad, dy planes by dgrad
 * and wgrad.  A caller that keeps the planes between calls (as ConvND does between forward
 * and backward) saves the re-formatting passes.  tensor: 0 = x, 1 = dy.
 * mgb_conv_planes_bytes reports 0 when the descriptor does not use planes (float64, tiny
 * channel counts); the *_p entry points accept NULL planes and then behave like the plain ones. */
int mgb_conv_planes_bytes(const mgb_conv_desc* d, int tensor, size_t* bytes);
int mgb_conv_make_planes(const mgb_conv_desc* d, int tensor, const void* src, void* planes,
                         void* stream);
int mgb_conv_fwd_p(const mgb_conv_desc* d, const void* x, const void* w, void* y,
                   const void* x_planes, void* ws, size_t ws_bytes, void* stream);
int mgb_conv_dgrad_p(const mgb_conv_desc* d, const void* dy, const void* w, void* dx,
                     const void* dy_planes, void* ws, size_t ws_bytes, void* stream);
int mgb_conv_wgrad_p(const mgb_conv_desc* d, const void* dy, const void* x, void* dw,
                     const void* dy_planes, const void* x_planes, void* ws, size_t ws_bytes,
                     void* stream);

/* ---- pooling ----------------------------------------------------------- */

/* Pooling over the trailing `nd` axes of x (lead, X...) -> (lead, G...),
 * G = (X - P)/stride + 1, no padding, no dilation (pooling.py:74-86). */
typedef struct {
  int32_t dtype;                 /* mgb_dtype */
  int32_t nd;                    /* 1..MGB_MAX_SPATIAL */
  int32_t mode;                  /* 0 = max, 1 = mean */
  int32_t reserved;
  int64_t lead;                  /* product of the un-pooled leading axes */
  int64_t X[MGB_MAX_SPATIAL];
  int64_t P[MGB_MAX_SPATIAL];
  int64_t stride[MGB_MAX_SPATIAL];
} mgb_pool_desc;

/* replaces MaxPoolND.__call__  (src/mygrad/nnet/layers/pooling.py:13-98) */
int mgb_pool_fwd(const mgb_pool_desc* d, const void* x, void* y, void* stream);

/* dx[argmax-of-window(g)] += dy[g]; first max in row-major window order; NaN
 * counts as max; overlapping windows accumulate in float64 in increasing flat
 * window order, then cast to dtype.
 * replaces MaxPoolND.backward_var  (pooling.py:100-149) */
int mgb_pool_bwd(const mgb_pool_desc* d, const void* dy, const void* x, void* dx,
                 void* stream);

/* ---- conv_nd -> max_pool fused (opt-in, float32 tensor-core path) ----------
 * The reference materialises y = conv_nd(x, w) (conv.py:102-103), re-reads it in MaxPoolND.__call__
 * (pooling.py:93-98) and again in MaxPoolND.backward_var (pooling.py:109-149).  The fused pair keeps the
 * conv accumulator tile on chip: the forward writes only the pooled values and, per window, the position
 * (0..3, row-major) of its first maximum; the backward turns (dp, argmax) straight into the channels-last
 * dy planes that mgb_conv_dgrad_p / mgb_conv_wgrad_p read.  Results are bit-identical to
 * mgb_conv_fwd -> mgb_pool_fwd -> mgb_pool_bwd -> mgb_conv_make_planes when mgb_conv_fwd runs the same
 * (halo) kernel, equal to rounding otherwise.
 * Supported: 2 spatial dims, conv stride 1, >= 16 input and output channels, pool (2,2) stride (2,2)
 * over an even-sized conv output; *yes = 0 otherwise (callers then run the unfused entry points). */
int mgb_conv_pool_supported(const mgb_conv_desc* d, const mgb_pool_desc* pd, int* yes);
/* pooled: (N, F, G0/2, G1/2) float32; argmax: same shape, uint8; ws as for mgb_conv_fwd (which = 0) */
int mgb_conv_pool_fwd(const mgb_conv_desc* d, const mgb_pool_desc* pd, const void* x, const void* w,
                      void* pooled, void* argmax, const void* x_planes, void* ws, size_t ws_bytes,
                      void* stream);
/* dy_planes: mgb_conv_planes_bytes(d, 1) bytes, as made by mgb_conv_make_planes(d, 1, dy, ...) */
int mgb_pool_bwd_planes(const mgb_conv_desc* d, const mgb_pool_desc* pd, const void* dp,
                        const void* argmax, void* dy_planes, void* stream);
/* The unfused pair's transparent half of the fusion: MaxPoolND.backward_var (pooling.py:100-149) of a tensor
 * that a conv_nd produced writes dy (y.grad, as always) AND the channels-last dy planes the convolution's
 * backward reads, in one pass over y and dp — replaces mgb_pool_bwd + mgb_conv_make_planes(d, 1, ...); same
 * geometry as above.  dy is bit-identical to mgb_pool_bwd's, the planes to mgb_conv_make_planes'. */
int mgb_pool_bwd_dy_planes(const mgb_conv_desc* d, const mgb_pool_desc* pd, const void* dp, const void* y,
                           void* dy, void* dy_planes, void* stream);

/* ---- Tensor.backward reductions / element-wise --------------------------- */

/* out = (accumulate ? out : 0) + sum of g over the broadcast axes:
 * leading extra axes of g are summed away, and axes where var_shape is 1 are
 * summed with keepdims.  g_ndim >= v_ndim, both <= MGB_MAX_DIMS.
 * replaces reduce_broadcast (src/mygrad/_utils/__init__.py:131-168) as called
 * from Operation.grad_post_process_fn (src/mygrad/operation_base.py:89-105) */
int mgb_reduce_broadcast(int dtype, int g_ndim, const int64_t* g_shape, int v_ndim,
                         const int64_t* v_shape, const void* g, void* out,
                         int accumulate, void* stream);

/* dst[i] += src[i]   — `var._grad += backed_grad` (operation_base.py:227-228) */
int mgb_axpy(int dtype, size_t n, const void* src, void* dst, void* stream);
/* dst[i] = (dst_dtype) src[i] — np.copy / astype (operation_base.py:217-224) */
int mgb_copy_cast(int src_dtype, int dst_dtype, size_t n, const void* src, void* dst,
                  void* stream);
/* dst[i] = value — np.full_like(data, 1.0) seed grad (tensor_base.py:1332) */
int mgb_fill(int dtype, size_t n, double value, void* dst, void* stream);
/* dst[i] *= alpha — used by the data-parallel wrapper (1/R for mean losses) */
int mgb_scale(int dtype, size_t n, double alpha, void* dst, void* stream);

/* ---- ops around the path (what a small CNN needs so that a whole
 *      Tensor.backward stays on the device; all bandwidth-bound or tiny) ------------- */

/* y = x * (x > 0)   — ReLu.__call__ (src/mygrad/nnet/activations/relu.py:10-14) */
int mgb_relu_fwd(int dtype, size_t n, const void* x, void* y, void* stream);
/* dx = grad * (x > 0) — ReLu.backward_var (relu.py:16-17) */
int mgb_relu_bwd(int dtype, size_t n, const void* grad, const void* x, void* dx, void* stream);
/* out = a + b with NumPy broadcasting; the three shapes are given at the output rank
 * (ones prepended) — Add.__call__ (src/mygrad/math/arithmetic/ops.py:28-32) */
int mgb_add_broadcast(int dtype, int ndim, const int64_t* out_shape, const int64_t* a_shape,
                      const int64_t* b_shape, const void* a, const void* b, void* out, void* stream);
/* dst (cols, rows) = src (rows, cols)^T — operand packing for MatMul run as a 1x1 convolution
 * (src/mygrad/math/misc/ops.py:91-127) */
int mgb_transpose2d(int dtype, int64_t rows, int64_t cols, const void* src, void* dst, void* stream);
/* loss = mean_n(-log softmax(x)[n, label_n]); back = (softmax(x) - onehot)/N (may be NULL).
 * labels are class indices stored in the tensor dtype.
 * SoftmaxCrossEntropy.__call__ (src/mygrad/nnet/losses/softmax_crossentropy.py:14-50),
 * logsumexp (src/mygrad/math/_special.py:4-68) */
int mgb_softmax_xent(int dtype, int64_t n, int64_t c, const void* x, const void* labels, void* loss,
                     void* back, void* stream);
/* matmul with a small output width m <= 32 (classifier heads), bandwidth-bound:
 * which 0: out (n,m) = a (n,d) @ b (d,m); 1: out (n,d) = a (n,m) @ b (d,m)^T; 2: out (d,m) = a (n,d)^T @ b (n,m)
 * — MatMul.__call__ / backward_var (src/mygrad/math/misc/ops.py:91-127) */
int mgb_matmul_skinny(int dtype, int which, int64_t n, int64_t d, int64_t m, const void* a,
                      const void* b, void* out, void* stream);
/* dst[i] = (*scalar) * src[i], scalar on the device — `grad * self.back` (softmax_crossentropy.py:50) */
int mgb_scale_by(int dtype, size_t n, const void* scalar, const void* src, void* dst, void* stream);

/* Einstein summation  out[o] = sum_s prod_k op_k[...]  over strided operands: o runs over `nout` output
 * labels, s over `nsum` summed labels; every operand and the output are addressed through one ELEMENT
 * stride per label (0 = the operand does not vary with that label; the sum of several axis strides = a
 * repeated label, i.e. a diagonal).  Stride tables are row-major [operand][label].  Up to 8 operands and
 * 16 + 16 labels; products and sums in float64.  The output must be zero-filled by the caller when its
 * strides do not cover it (gradients written along a diagonal).
 * replaces np.einsum as called by EinSum.__call__ / backward_var (src/mygrad/linalg/ops.py:79-243) */
int mgb_einsum(int dtype, int nops, const void* const* ops, int nout, const int64_t* out_size,
               const int64_t* out_stride, const int64_t* op_out_stride, int nsum, const int64_t* sum_size,
               const int64_t* op_sum_stride, void* out, void* stream);

/* dst (G..., lead, W...) = windows of src (lead, X...) over its nd (1..3) trailing axes,
 * G = (X - ((W-1)*dil + 1)) / step + 1 — sliding_window_view (src/mygrad/nnet/layers/utils.py:7-228),
 * materialised because device arrays carry no strides.  The conv kernels never call this. */
int mgb_window_gather(int dtype, int nd, int64_t lead, const int64_t* x, const int64_t* w,
                      const int64_t* step, const int64_t* dil, const void* src, void* dst, void* stream);

/* ---- batch normalisation --------------------------------------------
 * x is (n, c, s): s = product of the axes after the channel axis.  Replaces
 * BatchNorm.__call__ / backward_var (src/mygrad/nnet/layers/batchnorm.py:26-111).
 * `moments` is 3*c float64 on the device: mean[c], population variance[c], count[c];
 * `sums` is 2*c float64: sum(grad)[c], sum(grad * x_norm)[c].  Both stay in float64 so that a
 * data-parallel caller can merge them across ranks before applying them. */
int mgb_batchnorm_workspace_bytes(int64_t n, int64_t c, size_t* bytes);
/* per-channel mean / variance over the n and s axes (batchnorm.py:56-57) */
int mgb_batchnorm_stats(int dtype, int64_t n, int64_t c, int64_t s, const void* x, void* moments,
                        void* ws, size_t ws_bytes, void* stream);
/* y = (x - mean) / sqrt(var + eps) [* gamma] [+ beta]; gamma / beta (c,) may be NULL (batchnorm.py:59-73) */
int mgb_batchnorm_fwd(int dtype, int64_t n, int64_t c, int64_t s, const void* x, const void* moments,
                      const void* gamma, const void* beta, double eps, void* y, void* stream);
/* sum(grad) and sum(grad * x_norm) per channel: dbeta and dgamma (batchnorm.py:102-107) and the
 * two means the dx formula needs */
int mgb_batchnorm_bwd_reduce(int dtype, int64_t n, int64_t c, int64_t s, const void* grad, const void* x,
                             const void* moments, double eps, void* sums, void* ws, size_t ws_bytes,
                             void* stream);
/* dx = (grad - mean(grad) - x_norm * sum(grad * x_norm)/count) / std [* gamma] (batchnorm.py:76-99) */
int mgb_batchnorm_bwd_dx(int dtype, int64_t n, int64_t c, int64_t s, const void* grad, const void* x,
                         const void* moments, const void* sums, const void* gamma, double eps, void* dx,
                         void* stream);

/* data-parallel batch norm: merges the per-rank `moments` into the full-batch ones.  The caller
 * runs phase 0, sum-all-reduces buf[0, 2c), runs phase 1, sum-all-reduces buf[2c, 3c), runs
 * phase 2 (buf: 3*c float64 on the device).  The reference has no ranks; the result to match is the
 * single-process reference on the concatenated batch. */
int mgb_batchnorm_merge(int phase, int64_t c, void* moments, void* buf, void* stream);

/* ---- runtime plumbing -------------------------------------------------- */

int mgb_device_count(int* count);
int mgb_set_device(int ordinal);
int mgb_get_device(int* ordinal);
int mgb_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem);
int mgb_malloc_async(void** ptr, size_t bytes, void* stream);
int mgb_free_async(void* ptr, void* stream);
int mgb_host_alloc(void** host_ptr, size_t bytes);      /* pinned */
int mgb_host_free(void* host_ptr);
int mgb_memcpy_h2d(void* dst, const void* host_src, size_t bytes, void* stream);
int mgb_memcpy_d2h(void* host_dst, const void* src, size_t bytes, void* stream);
int mgb_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream);
int mgb_memset(void* dst, int value, size_t bytes, void* stream);
int mgb_stream_create(void** stream);
int mgb_stream_destroy(void* stream);
int mgb_stream_synchronize(void* stream);
int mgb_device_synchronize(void);
int mgb_event_create(void** event);
int mgb_event_destroy(void* event);
int mgb_event_record(void* event, void* stream);
int mgb_event_synchronize(void* event);
int mgb_event_elapsed_ms(void* start, void* stop, float* ms);
int mgb_stream_wait_event(void* stream, void* event);
/* CUDA graphs for launch-bound steps: everything the calling thread enqueues on `stream` (not the legacy default
 * stream) between begin and end is recorded instead of executed; mgb_graph_launch replays it.  The reference has no
 * counterpart (it runs NumPy eagerly); mygrad_b200.StepGraph is the Python face. */
int mgb_graph_begin(void* stream);
int mgb_graph_end(void* stream, void** graph_exec, int* kernel_nodes);
int mgb_graph_launch(void* graph_exec, void* stream);
int mgb_graph_destroy(void* graph_exec);
/* number of kernels this library has launched in this process (bench.py's
 * gpu_launches claim is read from here) */
int mgb_launch_count(uint64_t* count);
const char* mgb_last_error(void);
const char* mgb_version(void);

/* ---- multi-GPU (new; the reference has no communication layer) ------------ */

/* NCCL communicator over the ranks of one node; id is ncclUniqueId bytes
 * (128) produced by mgb_comm_unique_id on rank 0 and shipped by the caller. */
int mgb_comm_unique_id(void* id128);
int mgb_comm_init_rank(const void* id128, int world_size, int rank);
int mgb_comm_destroy(void);
/* in-place sum all-reduce of a gradient buffer (dW, bias grads) */
int mgb_allreduce_sum(void* buf, size_t n, int dtype, void* stream);

/* Peer-memory collective (our own, over NVLink / NVSwitch): every rank allocates one region
 * (flags + two data slots of slot_bytes) and maps every other rank's region with CUDA IPC.
 *   mgb_peer_create  -> 64-byte IPC handle of this rank's region (ship it to all ranks);
 *   mgb_peer_connect <- the world_size handles in rank order (world_size * 64 bytes);
 *   mgb_peer_allreduce_sum: in-place sum over ranks of n elements (n * itemsize <= slot_bytes), one
 *     kernel: publish -> flag exchange -> peer loads, summed in rank order in float64, so every rank
 *     holds bit-identical results.  Calls are collective and must be issued in the same order on
 *     every rank.
 *   mgb_conv_wgrad_dp: mgb_conv_wgrad_p whose final split-K reduction ALSO sums over the ranks in
 *     the same kernel (the fused form of "dW = ConvND.backward_var(grad, 1)" + all-reduce; the
 *     reference has no ranks).  Without a connected peer group it is mgb_conv_wgrad_p. */
int mgb_peer_create(int rank, int world_size, size_t slot_bytes, void* ipc_handle64);
int mgb_peer_connect(const void* handles);
int mgb_peer_destroy(void);
int mgb_peer_slot_bytes(size_t* bytes);
/* *timed_out_call = ordinal of the first peer all-reduce that gave up waiting for a rank (0 = none; the
 * kernels wait at most 30 s and then raise this flag instead of hanging or trapping), *waited_for_rank =
 * the rank it was waiting for.  Once set, every later peer call fails with the same report. */
int mgb_peer_status(int* timed_out_call, int* waited_for_rank);
int mgb_peer_allreduce_sum(void* buf, size_t n, int dtype, void* stream);
int mgb_conv_wgrad_dp(const mgb_conv_desc* d, const void* dy, const void* x, void* dw,
                      const void* dy_planes, const void* x_planes, void* ws, size_t ws_bytes,
                      void* stream);

/* ---- measurement helpers (bench.py only) ---------------------------------- */

/* Register-resident MMA loops that measure the tensor-pipe ceilings the
 * roofline fractions are quoted against.  kind: 0 = FP64 DMMA m8n8k4,
 * 1 = TF32 mma.sync m16n8k8.  Returns achieved TFLOP/s. */
int mgb_probe_mma_peak(int kind, int iters, double* tflops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MGB_H_ */
