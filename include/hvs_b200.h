/* hvs_b200.h -- C ABI of libhvs_b200.so: the B200 (sm_100a) implementation of the
 * humanoid-vision-system hot path (mHC residual layer + detection decode / NMS).
 *
 * The reference (nazimurahman/humanoid-vision-system) is pure Python/PyTorch and has
 * no FFI layer: its boundary for this path is the nn.Module surface listed in
 * SURVEY.md section 8(b).  Each entry point below names the reference code it
 * replaces; INTEGRATION.md shows the ctypes binding a maintainer adds on the
 * reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - functions never allocate, never synchronise and are re-entrant across
 *     streams; the caller passes outputs and, where stated, a workspace;
 *   - return value: 0 = ok; > 0 = a cudaError_t; < 0 = HVS_ERR_* argument error.
 *     hvs_error_string() turns either into text;
 *   - there is NO CPU path: without a CUDA device the calls return the CUDA error.
 */
#ifndef HVS_B200_H
#define HVS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HVS_OK 0
#define HVS_ERR_BAD_ARG (-1)       /* null pointer / negative size                      */
#define HVS_ERR_UNSUPPORTED (-2)   /* shape outside what the kernels are built for      */
#define HVS_ERR_ALIGNMENT (-3)     /* pointer or stride not aligned as documented       */
#define HVS_ERR_WORKSPACE (-4)     /* workspace too small                               */
#define HVS_ERR_DRIVER (-5)        /* driver entry point (tensor-map encode) missing    */

int hvs_abi_version(void);
const char* hvs_error_string(int code);
/* number of kernels this library has launched since load (for bench.py's gpu_launches) */
uint64_t hvs_launch_count(void);

/* ------------------------------------------------------------------------------------
 * K1: stream mHC layer (north_star; SURVEY.md section 8(c) "K1 oracle").
 * Composed from the reference primitives RMSNorm (src/models/manifold_layers.py:449-456),
 * the sigmoid / 2*sigmoid gates (:213, :216) and the batched Sinkhorn-Knopp projection
 * (:32-93), applied per token to n residual streams of C channels.
 *
 *   x      [T, n, C] bf16, contiguous, 16-byte aligned
 *   phi    [n*C, n*n+2n] fp32   projection to (H_pre | H_post | H_res) logits
 *   bias   [n*n+2n] fp32
 *   alpha  [3] fp32             logit scale per group (pre, post, res)
 *   scale  [n*C] fp32           RMSNorm gain
 *   y      [T, n, C] bf16       y = H_res x + H_post (x) (H_pre^T x)      (may be NULL)
 *   u      [T, C] bf16          layer input H_pre^T x                      (may be NULL)
 *   coeffs [T, n*n+2n] fp32     H_pre (n) | H_post (n) | H_res (n*n, row-major) (may be NULL)
 *
 * flags: HVS_MHC_SPLIT_PHI keeps the projection operand scale*phi at fp32 accuracy instead of rounding it to bf16.
 * Supported: sk_iters in [0, 64]; n == 4, C == 512 on the tuned TMA / tensor-core kernel (the roofline path), and every
 * n in {2, 4}, C % 8 == 0, C <= 1024 -- and HVS_MHC_SPLIT_PHI for all shapes -- on a general warp-per-token kernel
 * (forward only: the training entry points below are n == 4, C == 512).
 * ---------------------------------------------------------------------------------- */
#define HVS_MHC_SPLIT_PHI 1u
/* HVS_MHC_ADAPTIVE_ITERS (opt-in; n = 4, C = 512 kernels): the per-token Sinkhorn loop stops at the first iteration that
 * changes no scaling (forward) / no entry of P (backward replay) of any token of the warp by more than 2^-20 relative.  The
 * iteration contracts geometrically, so what the remaining iterations of the reference's fixed count would still change is
 * ~2e-6 relative (coefficient tolerance: 1e-5); the fused backward differentiates the iterations actually run (gradients
 * within 1e-5 of the full sweep).  Default off: every call runs all sk_iters like the reference. */
#define HVS_MHC_ADAPTIVE_ITERS 2u

int hvs_mhc_stream_fwd(const void* x, const float* phi, const float* bias, const float* alpha,
                       const float* scale, void* y, void* u, float* coeffs, int64_t T, int n, int C,
                       int sk_iters, float eps_rms, float eps_sk, uint32_t flags, void* stream);

/* Training forward: the same kernel, additionally writing the per-token statistics the fused backward
 * needs instead of recomputing them:
 *   saved [T, HVS_MHC_SAVED_STRIDE] fp32 = raw[n*n+2n] (x . bf16(scale*phi), before the RMS factor, alpha and
 *   bias) | sum_k x_k^2 | zero pad.  112 B/token next to the 8192 B/token of x and y. */
#define HVS_MHC_SAVED_STRIDE 28
int hvs_mhc_stream_fwd_save(const void* x, const float* phi, const float* bias, const float* alpha,
                            const float* scale, void* y, void* u, float* coeffs, float* saved, int64_t T, int n,
                            int C, int sk_iters, float eps_rms, float eps_sk, uint32_t flags, void* stream);

/* y = H_res x + H_post (x) fu  with coefficients produced by hvs_mhc_stream_fwd(y=NULL);
 * fu [T, C] bf16 is the wrapped layer's output F(u). */
int hvs_mhc_stream_post(const void* x, const float* coeffs, const void* fu, void* y, int64_t T, int n,
                        int C, void* stream);

/* Backward of hvs_mhc_stream_fwd (F = identity), coefficients recomputed.
 *   dy [T,n,C] bf16 -> dx [T,n,C] bf16, dphi [n*C, n*n+2n] fp32, dbias [n*n+2n], dalpha [3],
 *   dscale [n*C]  (parameter gradients are OVERWRITTEN, not accumulated).
 * Shapes: n = 4, C = 512 on the tuned two-kernel path (TMA tiles, tensor-memory parking); every other
 * n in {2, 4}, C % 8 == 0, C <= 1024 on the general path (a per-token kernel, the tcgen05 GEMM kernel for
 * dW = x^T E with fp32-accurate two-term E, a finalize).  HVS_MHC_SPLIT_PHI is forward-only.
 * sk_iters <= 24 (HVS_ERR_UNSUPPORTED beyond; the forward alone takes up to 64).
 * workspace: hvs_mhc_stream_bwd_workspace(T, n, C) bytes, 256-byte aligned. */
size_t hvs_mhc_stream_bwd_workspace(int64_t T, int n, int C);
int hvs_mhc_stream_bwd(const void* x, const void* dy, const float* phi, const float* bias,
                       const float* alpha, const float* scale, void* dx, float* dphi, float* dbias,
                       float* dalpha, float* dscale, int64_t T, int n, int C, int sk_iters,
                       float eps_rms, float eps_sk, uint32_t flags, void* workspace,
                       size_t workspace_bytes, void* stream);

/* Backward of hvs_mhc_stream_fwd_save (F = identity) in ONE fused kernel: dx and dphi/dscale/dbias/dalpha in a
 * single pass over x and dy (12288 + 112 B/token); `saved` is the forward's [T, HVS_MHC_SAVED_STRIDE] record.
 * The Sinkhorn forward iterations are replayed from the saved logits and differentiated exactly.
 * Same outputs and conventions as hvs_mhc_stream_bwd.  sk_iters <= 24 (HVS_ERR_UNSUPPORTED beyond).
 * workspace: hvs_mhc_stream_bwd_saved_workspace(T, n, C) bytes, 256-byte aligned. */
size_t hvs_mhc_stream_bwd_saved_workspace(int64_t T, int n, int C);
int hvs_mhc_stream_bwd_saved(const void* x, const void* dy, const float* saved, const float* phi,
                             const float* bias, const float* alpha, const float* scale, void* dx, float* dphi,
                             float* dbias, float* dalpha, float* dscale, int64_t T, int n, int C, int sk_iters,
                             float eps_rms, float eps_sk, uint32_t flags, void* workspace, size_t workspace_bytes,
                             void* stream);

/* Optional per-kernel timing of the stream-mHC launches (used by bench.py for the roofline line).
 * When enabled, each launch is bracketed by CUDA events on the launching stream (a ring of 128 event
 * pairs per kernel; enabling resets it).  Nothing waits on the host while launches are recorded;
 * hvs_mhc_stream_kernel_ms synchronises the recorded events and returns the MEAN duration in ms of
 * the (last 128) launches of each kernel since profiling was enabled: out[0] forward kernel,
 * out[1] backward per-token / fused kernel, out[2] backward x^T E reduction kernel (two-kernel
 * path), out[3] backward finalize kernel (negative = not run). */
int hvs_mhc_stream_profile(int enable);
int hvs_mhc_stream_kernel_ms(float* out4_host);

/* ------------------------------------------------------------------------------------
 * Sinkhorn-Knopp projection, SinkhornKnoppProjection.forward (manifold_layers.py:32-93):
 * out = SK(in) for `batch` matrices of n x m fp32, row-major, contiguous.
 * `history` (may be NULL) receives |mean(row_sum) - 1| per iteration ([iters] fp32), the
 * reference's convergence_history buffer (:76-77).
 * Supported: n, m <= 32 for any batch ("per-token" blocks, a warp per matrix), or larger matrices with
 * n, m <= 4096 (a layer's D x D H_res_raw; `history` then only with batch == 1).
 * `out` is bitwise reproducible run to run; the `history` diagnostic of the small-block path accumulates
 * with fp32 atomics and may differ in its last bits.
 * ---------------------------------------------------------------------------------- */
int hvs_sinkhorn(const float* in, float* out, int64_t batch, int n, int m, int iters, float eps,
                 float tau, float* history, void* stream);

/* ManifoldHyperConnection.constrained_matrices (manifold_layers.py:205-221) for one layer:
 * H_pre = sigmoid(H_pre_raw) [D,nD], H_post = 2 sigmoid(H_post_raw) [nD,D],
 * H_res = SK(H_res_raw) [D,D]. */
int hvs_mhc_constrained_matrices(const float* h_pre_raw, const float* h_post_raw, const float* h_res_raw,
                                 float* h_pre, float* h_post, float* h_res, int D, int hidden, int iters,
                                 float eps, float* history, void* stream);

/* ------------------------------------------------------------------------------------
 * Row normalisations of the path.
 * RMSNorm.forward (src/models/manifold_layers.py:449-456): out = x / sqrt(mean(x^2, -1) + eps) * scale,
 * rows x dim, contiguous; x / out dtypes HVS_DTYPE_F32 or HVS_DTYPE_BF16 (statistics always fp32).
 * hvs_rmsnorm_bwd: dy -> dx (same dtype as x) and dscale [dim] fp32 (overwritten; two-stage fixed-order
 * reduction through `workspace`, hvs_rmsnorm_bwd_workspace bytes).  dim <= 6144.
 * hvs_layernorm_fwd: nn.LayerNorm over the last axis as the module applies it before / after the token path
 * (manifold_layers.py:250, :267).  out has row stride out_ld >= dim (columns [dim, out_ld) are zero-filled so
 * the row can feed a GEMM whose K is padded); x_bf16_copy (may be NULL) receives bf16(x) with row stride copy_ld.
 * ---------------------------------------------------------------------------------- */
int hvs_rmsnorm_fwd(const void* x, int x_dtype, const float* scale, void* out, int out_dtype, int64_t rows,
                    int dim, float eps, void* stream);
size_t hvs_rmsnorm_bwd_workspace(int64_t rows, int dim);
int hvs_rmsnorm_bwd(const void* x, int dtype, const float* scale, const void* dy, void* dx, float* dscale,
                    int64_t rows, int dim, float eps, void* workspace, size_t workspace_bytes, void* stream);
int hvs_layernorm_fwd(const void* x, int x_dtype, const float* weight, const float* bias, void* out, int out_dtype,
                      void* x_bf16_copy, int64_t rows, int dim, int out_ld, int copy_ld, float eps, void* stream);
/* LayerNorm backward (training of the module's norm_pre / norm_post): x and dy fp32 / bf16 independently -> dx (x's dtype), dweight,
 * dbias [dim] fp32 (overwritten; fixed-order two-stage reduction through `workspace`).  dim in {32, 64, 128, 256, 512}. */
size_t hvs_layernorm_bwd_workspace(int64_t rows, int dim);
int hvs_layernorm_bwd(const void* x, int x_dtype, const float* weight, const void* dy, int dy_dtype, void* dx, float* dweight,
                      float* dbias, int64_t rows, int dim, float eps, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * K2, the reference-literal module: static coefficients of MANY layers in one launch.
 * ManifoldHyperConnection.constrained_matrices (manifold_layers.py:205-221) for every job:
 *   H_pre = sigmoid(H_pre_raw) [D,H], H_post = 2 sigmoid(H_post_raw) [H,D], H_res = SK(H_res_raw) [D,D]
 * plus the bf16, transposed copies the token-path GEMMs take as B operands ([N,K] row-major, K padded to Dp):
 *   h_pre_t [H,Dp] = H_pre^T, h_post_t [D,H] = H_post^T, h_res_t [D,Dp] = H_res^T.
 * One cooperative kernel; a large D x D matrix is iterated by many CTAs (row slabs).  D <= 4096.
 * uv_history [(iters+1), 2, D] fp32 (scalings u_k | v_k, k = 0..iters) is what the backward consumes;
 * convergence [iters] = the reference's convergence_history buffer (:76-77).  Any output pointer except
 * h_res may be NULL.  Job arrays live in HOST memory (the call copies its tables into the workspace).
 * ---------------------------------------------------------------------------------- */
typedef struct hvs_coeff_job {
    const float* h_pre_raw;
    const float* h_post_raw;
    const float* h_res_raw;
    float* h_pre;
    float* h_post;
    float* h_res;
    void* h_pre_t;
    void* h_post_t;
    void* h_res_t;
    float* uv_history;
    float* convergence;
    int32_t D, H, Dp, reserved;
} hvs_coeff_job;

/* gradients for hvs_mhc_static_coeffs_bwd: d_h_* are dL/dH (inputs, may be NULL), d_*_raw the parameter
 * gradients (outputs, overwritten). */
typedef struct hvs_coeff_grad {
    const float* d_h_pre;
    const float* d_h_post;
    const float* d_h_res;
    float* d_h_pre_raw;
    float* d_h_post_raw;
    float* d_h_res_raw;
} hvs_coeff_grad;

size_t hvs_mhc_static_coeffs_workspace(const hvs_coeff_job* jobs_host, int num_jobs, int iters, int backward);
int hvs_mhc_static_coeffs(const hvs_coeff_job* jobs_host, int num_jobs, int iters, float eps, void* workspace,
                          size_t workspace_bytes, void* stream);
/* Backward through the gates and the Sinkhorn iterations (exact reverse sweep of the scaling form; jobs must
 * carry the uv_history written by the forward with the same raw parameters). */
int hvs_mhc_static_coeffs_bwd(const hvs_coeff_job* jobs_host, const hvs_coeff_grad* grads_host, int num_jobs,
                              int iters, float eps, void* workspace, size_t workspace_bytes, void* stream);

/* K2 token path (manifold_layers.py:248-270) as tcgen05 GEMMs with fused epilogues:
 *   out[M,N] = epilogue( A0[M,K0] B0[N,K0]^T + A1[M,K1] B1[N,K1]^T )        (second pair optional: K1 = 0)
 * A*: bf16 row-major, row stride lda* elements (multiple of 8); B*: bf16 [N,K*] row-major contiguous (the
 * nn.Linear weight layout).  K0, K1 multiples of 8 (TMA zero-fills a partial last 64-element stage); N a multiple of 32,
 * and of 256 when N > 256; bias 16-byte aligned.
 *   HVS_GEMM_EPI_NONE       out = acc (+ bias if given)                        z = LN(x) @ H_pre          (:253)
 *   HVS_GEMM_EPI_BIAS_GELU  out = gelu_erf(acc + bias[n])                      mlp Linear + GELU          (:164-168)
 *   HVS_GEMM_EPI_LAYERNORM  out = LayerNorm_N(acc) * ln_w + ln_b, N <= 512     norm_post(z@H_post + x@H_res) (:259-267)
 * out: fp32 or bf16 (out_dtype), row stride ldo elements (multiple of 8). */
#define HVS_GEMM_EPI_NONE 0
#define HVS_GEMM_EPI_BIAS_GELU 1
#define HVS_GEMM_EPI_LAYERNORM 2
int hvs_gemm_bf16(const void* a0, int64_t lda0, const void* b0, int K0, const void* a1, int64_t lda1,
                  const void* b1, int K1, const float* bias, const float* ln_w, const float* ln_b, float ln_eps,
                  void* out, int out_dtype, int64_t ldo, int64_t M, int N, int epilogue, void* stream);

/* Training form of the same kernel: the module's forward AND the ten GEMMs of its backward (the autograd of
 * manifold_layers.py:248-267: torch.matmul / nn.Linear / nn.GELU / nn.Dropout backward under bf16 autocast).
 *   out[M,N] = epilogue( op(A0) op(B0)^T (+ A1 B1^T) )
 * a_mn_major / b_mn_major = 0: the operand is [rows, K] row-major (K-major), leading dimension ld >= K.
 *                         = 1: the operand is given as its TRANSPOSE in place, [K, rows] row-major, ld >= rows
 *                              (rows = M for A, N for B).  Data gradients dX = dY W read W [out, in] as stored
 *                              (B MN-major); weight gradients dW[M = features_a, N = features_b] = A^T B contract over
 *                              K = tokens with both activations [T, features] as stored (A and B MN-major).
 * split_k > 1: out is [split_k', M, ldo] fp32 partials (split_stride elements apart, default M * ldo) of the K range cut
 *   into split_k' <= split_k parts of whole 64-element blocks; the number actually used is what hvs_gemm_choose_split
 *   returns / the value passed clamped to the number of K blocks; sum them with hvs_reduce_partials (fixed order).
 * Epilogues in addition to the three above:
 *   HVS_GEMM_EPI_BIAS_GELU_SAVE  out2 = z = bf16(acc + bias) (pre-activation, kept for the backward),
 *                                out  = dropout_p(gelu_erf(z))            nn.Linear -> nn.GELU -> nn.Dropout (:164-169)
 *   HVS_GEMM_EPI_DGELU           out  = acc * gelu_erf'(aux[m,n]) * keep(m,n) / (1 - p)      their backward
 * Dropout: element (m, n) is dropped iff 16 bits of a counter-based hash of (dropout_seed, m, n / 2) are below
 * round(p * 65536); the same (seed, p) in the forward and the backward GEMM reproduces the mask, nothing is stored.
 * dropout_p = 0 switches it off.  The second operand pair is K-major only and excludes split_k. */
#define HVS_GEMM_EPI_BIAS_GELU_SAVE 3
#define HVS_GEMM_EPI_DGELU 4
typedef struct hvs_gemm_args {
    const void* a0; int64_t lda0;
    const void* b0; int64_t ldb0;
    int K0;
    const void* a1; int64_t lda1;
    const void* b1; int64_t ldb1;
    int K1;
    int a_mn_major, b_mn_major;
    const float* bias;
    const float* ln_w; const float* ln_b; float ln_eps;
    const void* aux; int64_t ld_aux;
    void* out; int out_dtype; int64_t ldo;
    void* out2; int64_t ldo2;
    int64_t M; int N; int epilogue;
    float dropout_p; uint32_t dropout_seed;
    int split_k; int64_t split_stride;
    const uint32_t* dropout_seed_dev;   /* optional: a DEVICE word mixed into dropout_seed when the kernel runs (a step counter
                                           advanced on the device, so a captured CUDA graph draws a new mask every replay) */
    float* colsum_partials;             /* optional, HVS_GEMM_EPI_DGELU: [4 * ceil(M / 128), N] fp32, row 4 * (m / 128) + (m % 128) / 32 =
                                           the column sums of the (bf16) output over that group of 32 rows -- nn.Linear's bias gradient
                                           is hvs_colsum_f32 of it: no second pass over d z */
} hvs_gemm_args;
int hvs_gemm_bf16_ex(const hvs_gemm_args* args, void* stream);
int hvs_gemm_choose_split(int64_t M, int N, int64_t K);
/* out[i] = sum over s < splits of partials[s * split_stride + i], i < numel, in that order (numel, stride multiples of 4). */
int hvs_reduce_partials(const float* partials, int splits, int64_t split_stride, int64_t numel, float* out, void* stream);
/* out[c] = sum over rows of x[r, c] for a contiguous fp32 [rows, cols] matrix (the colsum_partials above), fixed order. */
size_t hvs_colsum_f32_workspace(int64_t rows, int cols);
int hvs_colsum_f32(const float* x, int64_t rows, int cols, float* out, void* workspace, size_t workspace_bytes, void* stream);
/* out[c] = sum over rows of x[r, c], x bf16 [rows, cols] with row stride ld (bias gradients: nn.Linear's db = sum_t dz). */
size_t hvs_colsum_bf16_workspace(int64_t rows, int cols);
int hvs_colsum_bf16(const void* x, int64_t ld, int64_t rows, int cols, float* out, void* workspace, size_t workspace_bytes, void* stream);

/* Signal-ratio monitor of ManifoldHyperConnection._monitor_stability (manifold_layers.py:295-303):
 *   dst[0] = mean_rows ||out[r, :]|| / (mean_rows ||x[r, :]|| + 1e-8),  out, x [rows, dim] contiguous fp32 / bf16.
 * dst is a device pointer (the module's signal_ratio_history slot); no host synchronisation. */
size_t hvs_signal_ratio_workspace(int64_t rows);
int hvs_signal_ratio(const void* out, int out_dtype, const void* x, int x_dtype, int64_t rows, int dim, float* dst,
                     void* workspace, size_t workspace_bytes, void* stream);

/* The whole token path (manifold_layers.py:248-267, eval mode) in ONE kernel per 128-token tile -- LayerNorm_pre in the
 * prologue, the five GEMMs chained through tensor memory and shared memory, GELU / residual / LayerNorm_post in the
 * epilogues -- for the widths whose five-launch path is bound by its HBM round trips (the backbone's first stages):
 * (D, H) = (32, 128) or (64, 256), see hvs_mhc_module_fwd_supported.
 *   x [T, D] bf16 contiguous; h_pre_t [H, D], h_post_t [D, H], h_res_t [D, D] bf16 (hvs_mhc_static_coeffs outputs);
 *   w1 [2H, H], w2 [H, 2H] bf16 (mlp.0 / mlp.3 weights), b1 [2H], b2 [H] fp32; LayerNorm weights / biases [D] fp32;
 *   out [T, D] fp32 or bf16.  All pointers 16-byte aligned. */
int hvs_mhc_module_fwd_supported(int D, int H);
int hvs_mhc_module_fwd(const void* x, const void* h_pre_t, const void* w1, const float* b1, const void* w2,
                       const float* b2, const void* h_post_t, const void* h_res_t, const float* ln_pre_w,
                       const float* ln_pre_b, float ln_pre_eps, const float* ln_post_w, const float* ln_post_b,
                       float ln_post_eps, void* out, int out_dtype, int64_t T, int D, int H, void* stream);

/* Mean duration in ms of the (last 128) launches per kernel slot since hvs_mhc_stream_profile(1):
 * 0-3 as hvs_mhc_stream_kernel_ms, 4 = K2 GEMM, 5 = decode, 6 = fused K2 module kernel, 7 = static coefficients; negative = not run. */
int hvs_profile_kernel_ms(float* out8_host);

/* ------------------------------------------------------------------------------------
 * Edges of the path.
 * hvs_grad_clip_dual: ManifoldConstrainedTrainer._apply_manifold_gradient_clipping
 * (src/training/mhc_trainer.py:342-383).  tensors_host: fp32 gradient tensors (device pointers) tagged with a group
 * (0 = mHC parameters: name contains "mhc" or "H_", 1 = others).  Per group, torch's clip_grad_norm_ rule:
 * coef = max_norm / (norm_2 + 1e-6), gradients multiplied by coef when coef < 1.  result4 (device) receives
 * {norm group 0, norm group 1, applied coef 0, applied coef 1}.  No host synchronisation; fixed-order reductions.
 * ---------------------------------------------------------------------------------- */
typedef struct hvs_grad_tensor {
    float* grad;
    int64_t numel;
    int32_t group;
    int32_t reserved;
} hvs_grad_tensor;
size_t hvs_grad_clip_dual_workspace(const hvs_grad_tensor* tensors_host, int num_tensors);
int hvs_grad_clip_dual(const hvs_grad_tensor* tensors_host, int num_tensors, float max_norm_group0,
                       float max_norm_group1, float* result4, void* workspace, size_t workspace_bytes, void* stream);

/* Squeeze-excite gate and residual add after an mHC hop of ConvMHCLayer (src/models/vision_backbone.py:125-133) as ONE pass
 * over the channels-last token view:  out[t, c] = y[t, c] * gate[image(t), c] (+ residual[t, c]);
 * y, residual, out [images * rows_per_image, channels] bf16 contiguous (may alias), gate [images, channels] bf16,
 * residual may be NULL; channels % 8 == 0; fp32 arithmetic, one rounding. */
int hvs_gate_residual_bf16(const void* y, const void* gate, const void* residual, void* out, int64_t images,
                           int64_t rows_per_image, int channels, void* stream);

/* The squeeze-excite gate of ConvMHCLayer (src/models/vision_backbone.py:77-83 `channel_attention`, applied :126-127):
 * AdaptiveAvgPool2d(1) -> 1x1 conv C -> hidden -> activation -> 1x1 conv hidden -> C -> sigmoid, in ONE launch over the
 * channels-last map (as torch ops: a reduce kernel, two small GEMMs with their bias passes, two elementwise kernels).
 *   y [images * rows_per_image, channels] bf16 contiguous; w1 [hidden, channels], b1 [hidden], w2 [channels, hidden],
 *   b2 [channels] bf16; gate [images, channels] bf16; activation as hvs_bias_act_bf16.
 * Rounding follows the bf16 autocast path (pooled mean, both 1x1 outputs, activation and sigmoid each rounded to bf16,
 * fp32 accumulation); sums are taken in a fixed order (deterministic).  channels % 8 == 0, hidden % 8 == 0,
 * channels <= 2048, hidden <= 1024, images <= 65535; y, w1, w2 16-byte aligned.
 * The workspace (hvs_se_gate_workspace bytes, 256-byte aligned) must be ZERO before its first use with this entry; the
 * kernel leaves it ready for the next call. */
size_t hvs_se_gate_workspace(int64_t images, int64_t rows_per_image, int channels);
int hvs_se_gate_bf16(const void* y, const void* w1, const void* b1, const void* w2, const void* b2, void* gate,
                     int64_t images, int64_t rows_per_image, int channels, int hidden, int activation,
                     void* workspace, size_t workspace_bytes, void* stream);

/* Bias of a BatchNorm-folded convolution + the activation that follows it (ConvMHCLayer, vision_backbone.py:100-110, eval
 * mode) in one pass over the channels-last map: out[t, c] = act(y[t, c] + bias[c]); y, out [rows, channels] bf16 (may
 * alias), bias [channels] fp32; activation 0 = identity, 1 = SiLU, 2 = ReLU, 3 = LeakyReLU(0.1); channels % 8 == 0. */
#define HVS_ACT_NONE 0
#define HVS_ACT_SILU 1
#define HVS_ACT_RELU 2
#define HVS_ACT_LEAKY_RELU_0P1 3   /* LeakyReLU(0.1): the prediction heads' conv stacks (yolo_head.py:99-112) */
int hvs_bias_act_bf16(const void* y, const float* bias, void* out, int64_t rows, int channels, int activation, void* stream);

/* hvs_preprocess_u8: ImagePreprocessor "accurate" path (src/inference/preprocessing.py:252-273, colour swap :199-203):
 * src HWC uint8 frame (device memory, 1 or 3 channels, row pitch in bytes) -> bilinear resize with cv2.INTER_LINEAR
 * sampling (half-pixel centres, edge clamp; arithmetic in fp32, without cv2's intermediate rounding to uint8) ->
 * swap_rb (BGR -> RGB) -> / 255 -> (x - mean) / std -> dst CHW [3, dst_h, dst_w] fp32 / fp16 / bf16.  mean3 / std3 are
 * HOST arrays (NULL = 0 / 1). */
int hvs_preprocess_u8(const void* src, int src_h, int src_w, int src_channels, int64_t src_pitch_bytes, void* dst,
                      int dst_dtype, int dst_h, int dst_w, int swap_rb, const float* mean3_host, const float* std3_host,
                      void* stream);

/* ------------------------------------------------------------------------------------
 * YOLODecoder.forward (src/models/yolo_head.py:220-294), one scale.
 *   pred      [B, A, H, W, 5+C] viewed through element strides pred_stride[5]
 *             (the head's permuted NCHW conv output is read in place), fp32 / fp16 / bf16
 *   anchor_wh [A, 2] fp32  anchor (w,h) / 416 (:50-51)
 *   boxes [B,A,H,W,4] fp32 xyxy normalised; class_scores [B,A,H,W] fp32;
 *   class_idx [B,A,H,W] int64; objectness [B,A,H,W] fp32 (may be NULL);
 *   scores [B,A,H,W,C] fp32 obj*cls (may be NULL).
 * ---------------------------------------------------------------------------------- */
#define HVS_DTYPE_F32 0
#define HVS_DTYPE_F16 1
#define HVS_DTYPE_BF16 2

int hvs_yolo_decode(const void* pred, int pred_dtype, const int64_t* pred_stride_host, const float* anchor_wh,
                    float* boxes, float* class_scores, int64_t* class_idx, float* objectness, float* scores,
                    int B, int A, int H, int W, int C, void* stream);

/* YOLODetectionHead's decode loop over its scales (src/models/yolo_head.py:536-555: `for scale_idx in range(self.num_scales):
 * ... self.decoder(pred, anchors, grid_size)`) as ONE launch: the coarse grids of a batch are a fraction of a wave each and would
 * run as latency-bound launches of their own.  Same outputs per scale as hvs_yolo_decode without the per-class score
 * tensor (class_scores / class_idx are the max / first argmax of obj * sigmoid(cls), bit-identical to hvs_yolo_decode).
 * All scales share B, C and the prediction dtype.  Scales that cannot take the vectorised plane-strided mapping
 * (unit W stride, W % 4 == 0, strides % 4 == 0) make the call fall back to one launch per scale.
 *   scales_host  host array of n_scales descriptors (read before the call returns) */
typedef struct {
    const void* pred;            /* [B, A, H, W, 5+C] through pred_stride (elements) */
    int64_t pred_stride[5];
    const float* anchor_wh;      /* [A, 2] fp32 */
    float* boxes;                /* [B, A, H, W, 4] */
    float* class_scores;         /* [B, A, H, W] */
    int64_t* class_idx;          /* [B, A, H, W] */
    float* objectness;           /* [B, A, H, W] or NULL */
    int A, H, W;
} hvs_decode_scale;

int hvs_yolo_decode_scales(const hvs_decode_scale* scales_host, int n_scales, int pred_dtype, int B, int C, void* stream);

/* Fused YOLOPredictionHead tail: the 1x1 prediction convolution (yolo_head.py:193-194) as a tcgen05 GEMM over the
 * head's token view, with YOLODecoder.forward (:241-285) as its epilogue -- the raw [B,3,H,W,85] predictions are never
 * written.  For the model's head: 3 anchors x (5 + 80) = 255 output channels.
 *   tokens    [B*H*W, C_in] bf16 (pixel-major = channels_last feature map), row stride ld_tokens elements
 *   weight256 [256, C_in] bf16 = pred_conv.weight[:, :, 0, 0] with one zero row appended; bias256 [256] fp32 likewise
 *   outputs as hvs_yolo_decode (boxes, class_scores, class_idx, optional objectness; no per-class score tensor).
 * Score thresholding and the order-preserving compaction of survivors (yolo_head.py:600-622) stay where the reference
 * has them, in hvs_post_process, which consumes these outputs directly. */
int hvs_head_decode_fused(const void* tokens, int64_t ld_tokens, const void* weight256, const float* bias256,
                          const float* anchor_wh, float* boxes, float* class_scores, int64_t* class_idx,
                          float* objectness, int B, int H, int W, int C_in, void* stream);

/* ------------------------------------------------------------------------------------
 * Greedy NMS over `num_problems` independent candidate sets (one CTA each).
 *   boxes   [sum N, 4] fp32; scores [sum N] fp32; classes [sum N] int64 (class-aware only)
 *   offsets [num_problems+1] int64 (device): prefix of set sizes; max_n: upper bound on any set's size
 *   (sizes the shared-memory score cache; max_n <= 49152)
 *   A candidate takes part only if score > score_thr (strict, yolo_head.py:605); pass
 *   -INFINITY to take all.  Ties in score: lower index first.
 * mode HVS_NMS_AGNOSTIC   = YOLODetectionHead.non_max_suppression (yolo_head.py:678-731):
 *                           xyxy boxes, a later box survives iff iou < iou_thr.
 * mode HVS_NMS_CLASS_AWARE= NMSFilter.apply/_standard_nms (src/inference/postprocessing.py:505-607):
 *                           cx,cy,w,h boxes converted as :540-549 (unless HVS_NMS_BOXES_XYXY),
 *                           a later box of the SAME class is suppressed iff iou > iou_thr.
 * IoU = inter / (((a1 + a2) - inter) + 1e-6) in fp32 without FMA contraction
 * (yolo_head.py:733-755, postprocessing.py:772-802).
 * Outputs per problem p: keep_count[p] <= max_det, and for r < keep_count[p]
 *   keep_idx[p*max_det + r]  index into the problem's candidates that passed score_thr,
 *                            numbered in input order (the reference's compacted list),
 *   keep_src[p*max_det + r]  index into the problem's full candidate list.
 * Both are in descending-score order.
 * ---------------------------------------------------------------------------------- */
#define HVS_NMS_AGNOSTIC 0
#define HVS_NMS_CLASS_AWARE 1
#define HVS_NMS_BOXES_XYXY 16

int hvs_nms(const float* boxes, const float* scores, const int64_t* classes, const int64_t* offsets,
            int num_problems, int64_t max_n, float score_thr, float iou_thr, int max_det, int mode, int64_t* keep_idx,
            int64_t* keep_src, int32_t* keep_count, void* stream);

/* YOLODetectionHead.post_process (yolo_head.py:571-676) for a whole batch: per scale
 * thresholded class-agnostic NMS capped at max_det, concatenation of the survivors in scale
 * order, second NMS.  Inputs are the per-scale decode outputs (num_scales <= 4):
 *   boxes[s] [B, N_s, 4], class_scores[s] [B, N_s], class_idx[s] [B, N_s]
 * Outputs: det_boxes [B, max_det, 4], det_scores [B, max_det], det_labels [B, max_det] int64,
 * det_count [B] int32.  workspace: hvs_post_process_workspace(B, num_scales, max_det) bytes. */
size_t hvs_post_process_workspace(int B, int num_scales, int max_det);
int hvs_post_process(const float* const* boxes_host, const float* const* class_scores_host,
                     const int64_t* const* class_idx_host, const int* n_per_scale_host, int num_scales,
                     int B, float conf_thr, float iou_thr, int max_det, float* det_boxes,
                     float* det_scores, int64_t* det_labels, int32_t* det_count, void* workspace,
                     size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HVS_B200_H */
