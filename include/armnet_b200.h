/*
 * armnet_b200.h -- C ABI of the B200-native ARM-Net forward hot path (libarmnet_b200.so).
 *
 * The reference (nusdbsystem/ARM-Net @ 7aeb3a4) is pure Python/PyTorch; its "FFI" for this path is
 * the chain of ATen calls made from models/layers.py, models/armnet.py, models/armnet_1h.py and
 * utils/entmax.py.  Each entry point below names the reference lines it replaces.  A maintainer binds
 * these with ctypes; armnet_b200/_capi.py is that binding.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise
 *   - tensors are dense row-major fp32 (ids: int64, or int32 when ids_i32 != 0)
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream);
 *     calls are asynchronous, re-entrant and keep no mutable global state
 *   - memory is borrowed: nothing is allocated, freed or retained across calls
 *   - return value: ARMNET_OK or a negative ARMNET_ERR_* code; never throws, never exits.
 *     armnet_last_error_string() gives the calling thread's last failure text.
 *   - out-of-range ids (reference: IndexError from nn.Embedding, layers.py:20) do not fault: the row is
 *     treated as zeros and bit 0 of *err_flag (device int, may be NULL) is set for the host to check.
 */
#ifndef ARMNET_B200_H_
#define ARMNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARMNET_B200_VERSION 100 /* 0.1.0 */

#define ARMNET_OK 0
#define ARMNET_ERR_NULL (-1)        /* required pointer is NULL */
#define ARMNET_ERR_SHAPE (-2)       /* non-positive / inconsistent sizes, alpha < 1 */
#define ARMNET_ERR_UNSUPPORTED (-3) /* no compiled kernel instance or shared-memory budget exceeded */
#define ARMNET_ERR_CUDA (-4)        /* CUDA runtime / launch error (text in armnet_last_error_string) */
#define ARMNET_ERR_ALIGN (-5)       /* pointer not aligned for its element type */

/* Threshold solver for the alpha-entmax gates (utils/entmax.py:29-68). */
#define ARMNET_SOLVER_AUTO 0   /* alpha==1 softmax; 1<alpha<2 Newton; alpha==2 Michelot; alpha>2 bisection */
#define ARMNET_SOLVER_BISECT 1 /* the reference's algorithm step by step (n_iter halvings, early exit at the
                                  fp32 fixed point), for parity validation */

int armnet_version(void);
const char *armnet_last_error_string(void);

/* Device facts the host side sizes its launches with (SM count, opt-in shared memory per block). */
int armnet_device_info(int *sm_count, int *smem_optin_bytes);

/*
 * layers.Embedding.forward (models/layers.py:15-21) fused with the value clamp of
 * ARMNetModel.forward (models/armnet.py:82, models/armnet_1h.py:81):
 *     v       = clamp ? min(max(values[b,f], clamp_lo), clamp_hi) : values[b,f]
 *     out[b,f,:] = table[ids[b,f], 0:E] * v              (one fp32 multiply: bit-exact vs the reference)
 * values is rewritten in place when clamp_inplace != 0 (the reference mutates the caller's tensor).
 * table is [V, ld] with ld >= E (row pitch in floats; ld == E for the nn.Embedding parameter itself).
 */
int armnet_embed_gather_f32(const void *ids, int ids_i32, float *values, const float *table, int64_t V,
                            int64_t ld, int64_t B, int F, int E, float *out, int clamp, float clamp_lo,
                            float clamp_hi, int clamp_inplace, int *err_flag, void *stream);

/*
 * layers.Linear.forward (models/layers.py:31-37): the 1-wide embedding of the LR term the zoo models add
 * (afm.py:40, xdfm.py:46,67):  y[b] = sum_f weight[ids[b,f]] * values[b,f] + bias[0]   (bias may be NULL).
 * weight: [V] (the nn.Embedding(nfeat, 1) parameter), y: [B].
 */
int armnet_linear_gather_f32(const void *ids, int ids_i32, const float *values, const float *weight, int64_t V,
                             int64_t B, int F, const float *bias, float *y, int *err_flag, void *stream);

/*
 * EntmaxBisect.forward / entmax_bisect / nn.Softmax(dim=-1) over the last axis
 * (utils/entmax.py:238-275, :134-175, :29-68; models/armnet.py:12-13).  x, p: [rows, F].
 */
int armnet_entmax_f32(const float *x, int64_t rows, int F, float alpha, int solver, int n_iter, float *p,
                      void *stream);

/* EntmaxBisectFunction.backward (utils/entmax.py:71-80), softmax backward when alpha == 1.
 * p, dp, dx: [rows, F]. */
int armnet_entmax_bwd_f32(const float *p, const float *dp, int64_t rows, int F, float alpha, float *dx,
                          void *stream);

/* Bytes of device scratch armnet_fused_fwd_f32 needs for the given shape (pre-contracted attention
 * parameters; batch independent). */
size_t armnet_fused_workspace_bytes(int F, int E, int K, int O);

/*
 * The fused hot path: value clamp -> embedding lookup -> attention logits -> alpha-entmax gates ->
 * gates * values -> exponential-neuron interaction.
 *   multi-head  models/armnet.py:82-87 with SparseAttLayer.forward :26-36      (w_is_linear_layout == 0)
 *       bilinear_w [K,E,D], query [K,O,D], att_values [K,O,F]
 *   one-head    models/armnet_1h.py:81-86 with SparseAttention.forward :25-34  (w_is_linear_layout != 0, K == 1)
 *       bilinear_w is the nn.Linear weight [D,E], query [O,D], att_values [O,F]
 *
 *   e[b,f,:]   = table[ids[b,f],:] * clamp(values[b,f])
 *   g[b,r,f]   = D^-0.5 * sum_x sum_y e[b,f,x] W[k,x,y] Q[k,o,y]          r = k*O + o
 *   p[b,r,:]   = entmax_alpha(g[b,r,:])            (softmax when alpha == 1)
 *   out_z[b,r,:] = exp( sum_f p[b,r,f] * att_values[r,f] * e[b,f,:] )     [B, K*O, E], the layout
 *                  arm_bn consumes (armnet.py:88 'b k o e -> b (k o) e')
 *
 * Optional epilogue (all three NULL to skip): eval-mode arm_bn = BatchNorm1d(K*O) on [B,K*O,E] (armnet.py:67,89),
 *   out_z[b,r,:] <- (out_z[b,r,:] - post_mean[r]) * post_scale[r] + post_shift[r]
 * with post_mean = running_mean, post_scale = weight / sqrt(running_var + eps), post_shift = bias   ([K*O] each).
 *
 * Optional outputs (NULL to skip): out_tau [B,K*O,2] = (threshold tau, sum of unnormalised gates) per row,
 * which is all a backward pass needs to rebuild p;  out_p [B,K*O,F] gates;  out_g [B,K*O,F] logits;
 * out_s [B,K*O,E] the pre-exp sums log(out_z) (validation only: uncoalesced stores).
 * workspace: armnet_fused_workspace_bytes(F,E,K,O) bytes, 16-byte aligned, rewritten by every call.
 * Launches two kernels on `stream` (parameter pre-contraction, fused forward).
 */
int armnet_fused_fwd_f32(const void *ids, int ids_i32, float *values, const float *table, int64_t V,
                         int64_t ld, const float *bilinear_w, const float *query, const float *att_values,
                         int w_is_linear_layout, float alpha, int solver, int n_iter, int64_t B, int F,
                         int E, int D, int K, int O, int clamp, float clamp_lo, float clamp_hi,
                         int clamp_inplace, const float *post_mean, const float *post_scale,
                         const float *post_shift, float *out_z, float *out_tau, float *out_p, float *out_g,
                         float *out_s, void *workspace, int *err_flag, void *stream);

/*
 * The same forward split in two for serving, where the attention parameters do not change between batches:
 *   armnet_fused_prepare_f32        pre-contracts (bilinear_w, query, att_values, alpha) into `workspace` (one small launch);
 *   armnet_fused_fwd_prepared_f32   runs the fused kernel on a batch with such a workspace, which it only READS -- it can be
 *                                   shared by concurrent streams.  Same arguments and semantics as armnet_fused_fwd_f32
 *                                   otherwise; alpha, F, E, D, K, O must be the ones the workspace was prepared with.
 */
int armnet_fused_prepare_f32(const float *bilinear_w, const float *query, const float *att_values,
                             int w_is_linear_layout, float alpha, int F, int E, int D, int K, int O, void *workspace,
                             void *stream);
int armnet_fused_fwd_prepared_f32(const void *ids, int ids_i32, float *values, const float *table, int64_t V, int64_t ld,
                                  float alpha, int solver, int n_iter, int64_t B, int F, int E, int D, int K, int O,
                                  int clamp, float clamp_lo, float clamp_hi, int clamp_inplace, const float *post_mean,
                                  const float *post_scale, const float *post_shift, float *out_z, float *out_tau,
                                  float *out_p, float *out_g, float *out_s, const void *workspace, int *err_flag,
                                  void *stream);

/*
 * Backward of the fused hot path w.r.t. its parameters (what autograd derives for models/armnet.py:82-87 with
 * EntmaxBisectFunction.backward, utils/entmax.py:71-80).  Per (sample, neuron) row it rebuilds the gates from the
 * (tau, sum) pairs the forward saved in out_tau and emits
 *     out_w  [B,F,K*O]  w  = p * att_values                    (armnet.py:36)
 *     out_dg [B,F,K*O]  dg = gradient w.r.t. the logits g      (entmax.py:76-79 applied to dp = (ds.e) * att_values)
 * (row index fastest: coalesced, and the layout a batched GEMM over rows wants), and accumulates over the batch
 *     acc_dvalues [K*O,F] += p * (ds . e)                      gradient of att_values
 *     acc_dm      [K*O,E] += sum_f dg_f * e_f                  gradient of the pre-contracted matrix, unscaled
 * with ds = dz * z.  The caller zeroes the accumulators; armnet_fused_bwd_finish_f32 (below) turns the partials into
 *     de[b,f,:] = sum_r out_w[b,f,r] ds[b,r,:] + d_k^-0.5 sum_r out_dg[b,f,r] M[:,r],   dT[id] += de * value,
 *     dW, dQ from acc_dm.   z: forward output WITHOUT the arm_bn epilogue; values: already clamped by the forward.
 * armnet_fused_bwd_supported() tells whether a backward kernel instance exists for (F, E).
 */
int armnet_fused_bwd_supported(int F, int E);

/* Which kernel armnet_fused_fwd_f32 / armnet_fused_fwd_prepared_f32 launch for this shape (same function either way,
 * models/armnet.py:82-89): 0 = no compiled instance, 1 = armnet_fwd_kernel (both E x F products on the FP32 pipe),
 * 2 = armnet_fwd_mma_kernel (the products as warp-level TF32 tensor-core MMAs with the 3xTF32 split; needs
 * K*O % 64 == 0, a compiled field / nemb bucket (33-40 fields, nemb 10 or 16) and a Newton-type solver; default only
 * where it measured faster on B200: nemb 16), 3 = armnet_fwd_tmem_kernel (attention logits as tcgen05.mma into tensor
 * memory, entmax with the MUFU-free pre-solve, csrc/fused_fwd_tmem.cuh; needs 39 or 40 fields, nemb <= 10, K*O a multiple
 * of 128 and <= 768, a Newton-type solver; chosen by default for a general alpha only (1 < alpha < 2, alpha != 1.5); at run time also 16-byte-aligned table rows and no validation outputs,
 * otherwise the call falls back to kind 1 / 2).  armnet_set_tuning("tmem" / "mma", 1 / 0 / -1) forces a kind on / off /
 * back to the default; the ARMNET_TMEM / ARMNET_MMA environment variables give the initial values (read once). */
int armnet_fused_fwd_kernel_kind(int F, int E, int K, int O, float alpha, int solver);
int armnet_fused_bwd_f32(const void *ids, int ids_i32, float *values, const float *table, int64_t V, int64_t ld,
                         const float *bilinear_w, const float *query, const float *att_values,
                         int w_is_linear_layout, float alpha, int64_t B, int F, int E, int D, int K, int O,
                         const float *z, const float *dz, const float *tau, float *out_w, float *out_dg,
                         float *acc_dvalues, float *acc_dm, void *workspace, int *err_flag, void *stream);

/*
 * Finishes the fused backward on the package's own kernels: from the partials of armnet_fused_bwd_f32 to the parameter
 * gradients the reference's autograd produces for the hot path (models/layers.py:12,20-21 dense embedding gradient;
 * models/armnet.py:33-34 einsum backward; models/armnet_1h.py:30-32):
 *     de[b,f,:]      = sum_r out_w[b,f,r] (dz*z)[b,r,:] + d_k^-0.5 sum_r out_dg[b,f,r] M[:,r]   (mma.sync 3xTF32, one CTA per sample)
 *     dT[ids[b,f],:] += de[b,f,:] * values[b,f]            dT [V,E]: the caller zeroes it (or accumulates across micro-batches)
 *     dW, dQ         = the two contractions of acc_dm with query / bilinear_w, times d_k^-0.5 (written, not accumulated)
 * bilinear_w / query in the layouts of armnet_fused_fwd_f32 (w_is_linear_layout: nn.Linear weight [D,E], K = 1).
 * workspace: armnet_fused_bwd_finish_workspace_bytes(E, K, O) bytes, 16-byte aligned.  F <= 64.
 */
size_t armnet_fused_bwd_finish_workspace_bytes(int E, int K, int O);
int armnet_fused_bwd_finish_f32(const void *ids, int ids_i32, const float *values, int64_t V, int64_t B, int F, int E,
                                int D, int K, int O, const float *bilinear_w, const float *query, int w_is_linear_layout,
                                const float *w, const float *dg, const float *z, const float *dz, const float *acc_dm,
                                float *dT, float *dW, float *dQ, void *workspace, void *stream);

/*
 * The trailing dense MLP in eval mode (models/layers.py:68-88, called at models/armnet.py:92, armnet_1h.py:89):
 * [Linear -> BatchNorm1d -> ReLU -> Dropout] x n -> Linear(., noutput).  Dropout is the identity in eval mode.
 *
 * First Linear (layers.py:73, the only true GEMM of the path) on the tcgen05 tensor cores with the 3xTF32 split
 * (x_hi.w_hi + x_lo.w_hi + x_hi.w_lo, fp32 accumulation in TMEM), which reproduces fp32 SGEMM to ~1e-6 relative:
 *   armnet_mlp_split_weight_f32   w [n] -> w_hi (TF32-rounded), w_lo = w - w_hi.  Once per weight version.
 *   armnet_mlp_linear_splits      recommended split-K factor for (B, K, N) on the current device (>= 1).
 *   armnet_mlp_linear_tf32x3      x [B,K] row-major, w_hi / w_lo [N,K] row-major (the nn.Linear weight layout)
 *                                 -> partials [splits, ceil(B/32), N, 32] (blocks of 32 samples, sample fastest):
 *                                 partial[s][b/32][n][b%32] = x[b, Ks] . w[n, Ks] over the s-th slice of the reduction axis; bias / BatchNorm / ReLU are applied by the tail kernel.
 *                                 Needs K % 4 == 0 and 16-byte aligned base pointers (TMA), else ARMNET_ERR_ALIGN.
 * Everything after it in ONE launch:
 *   armnet_mlp_tail_f32           h1 = relu(sum_s partial[s] * a1 + c1); n_rest further hidden layers
 *                                 h' = relu((h . W^T) * a + c); y = h . Wf^T + bf   (layers.py:73-81)
 *     packed (floats): a1[H] c1[H] | n_rest x { Wt[H][H] (transposed weight: [in][out]), a[H], c[H] } | Wf[NO][H] bf[NO]
 *     with a = bn.weight / sqrt(bn.running_var + eps), c = (linear.bias - bn.running_mean) * a + bn.bias.
 *     armnet_mlp_tail_packed_floats(H, n_rest, NO) gives the length of `packed`.  y: [B, NO].
 */
int armnet_mlp_split_weight_f32(const float *w, int64_t n, float *w_hi, float *w_lo, void *stream);
int armnet_mlp_linear_splits(int64_t B, int K, int N);
int armnet_mlp_linear_tf32x3(const float *x, int64_t B, int K, const float *w_hi, const float *w_lo, int N, int splits,
                             float *partials, void *stream);
/* The same GEMM with the product written directly, row-major: y [B,N] = x . w^T (+ bias[n] when bias != NULL), the whole
 * reduction axis in one accumulation chain (meant for short K; training uses it for dX = dY . W). */
int armnet_linear_tf32x3_dense(const float *x, int64_t B, int K, const float *w_hi, const float *w_lo, int N,
                               const float *bias, float *y, void *stream);
/* out [cols, rows] = in [rows, cols]^T (tiled through shared memory); operands of the training GEMMs (x^T, W^T, dY^T). */
int armnet_transpose_f32(const float *in, int64_t rows, int64_t cols, float *out, void *stream);
/* One hidden layer after the first on the tensor cores, fused with the activation of its input and -- for the last
 * hidden layer -- with its own BatchNorm/ReLU and the output Linear (layers.py:73-81, eval mode), 128 samples per CTA:
 *   x = relu(sum_s in_partials[s] * a_in + c_in)  [B,H_in];   acc = x . w^T  (3xTF32 on tcgen05, w_hi / w_lo [H_out,H_in])
 *   y != NULL:  y [B,NO] = relu(acc * a_out + c_out) . wf^T + bf     (wf [NO,H_out], NO <= 4)
 *   y == NULL:  out_partials [1, ceil(B/32), H_out, 32] = acc         (input of the next call)
 * H_out <= 256, in_splits <= 4, H_in % 4 == 0, else ARMNET_ERR_UNSUPPORTED / ARMNET_ERR_ALIGN (use armnet_mlp_tail_f32 then). */
int armnet_mlp_hidden_tc_f32(const float *in_partials, int in_splits, int64_t B, int H_in, const float *a_in,
                             const float *c_in, const float *w_hi, const float *w_lo, int H_out, const float *a_out,
                             const float *c_out, const float *wf, const float *bf, int NO, float *y, float *out_partials,
                             void *stream);
size_t armnet_mlp_tail_packed_floats(int H, int n_rest, int NO);
int armnet_mlp_tail_f32(const float *partials, int splits, int64_t B, int H, int n_rest, int NO, const float *packed,
                        float *y, void *stream);

/*
 * Train-mode arm_bn = nn.BatchNorm1d(K*O) on the interaction output x [B, C, L] (models/armnet.py:67,89,
 * armnet_1h.py:65,85): batch statistics per channel over (B, L), torch semantics (normalise with the biased variance,
 * running_var updated with the unbiased one, running = (1 - momentum) running + momentum batch).
 *   fwd: out = (x - mean) * invstd * weight + bias; save_mean / save_invstd [C] for the backward; running_* may be NULL.
 *   bwd: dx, dweight = sum dy * xhat, dbias = sum dy (what autograd derives for F.batch_norm(training=True)).
 * weight / bias may be NULL (affine=False).  workspace: armnet_bn_workspace_floats(B, C, L) floats.
 * Sums are taken around a per-channel pivot and combined in a fixed order in double: deterministic, and free of the
 * E[x^2] - E[x]^2 cancellation that x = exp(s) ~ 1 would cause.  Three launches each.
 */
size_t armnet_bn_workspace_floats(int64_t B, int C, int L);
int armnet_bn_train_fwd_f32(const float *x, int64_t B, int C, int L, const float *weight, const float *bias,
                            float *running_mean, float *running_var, float momentum, float eps, float *out,
                            float *save_mean, float *save_invstd, float *workspace, void *stream);
int armnet_bn_train_bwd_f32(const float *x, const float *dy, int64_t B, int C, int L, const float *weight,
                            const float *save_mean, const float *save_invstd, float *dx, float *dweight, float *dbias,
                            float *workspace, void *stream);

/*
 * Dense Adam over a flat fp32 bucket fused with gradient averaging and the [-clamp, clamp] gradient clamp of the
 * reference's per-parameter hooks (train.py:62-65: optim.Adam(lr), p.register_hook(g.clamp(-1, 1))):
 *     g = clamp(grad * grad_scale, -clamp, clamp)            (clamp <= 0: no clamp; grad_scale = 1 / world size)
 *     m = lerp(m, g, 1 - beta1);  v = beta2 v + (1 - beta2) g^2
 *     p -= lr / (1 - beta1^step) * m / (sqrt(v) / sqrt(1 - beta2^step) + eps)        (torch.optim.Adam, no amsgrad)
 * step is the 1-based update count.  One pass over (p, g, m, v); buffers 16-byte aligned.
 */
int armnet_clamp_adam_f32(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n,
                          float grad_scale, float clamp, float lr, float beta1, float beta2, float eps, int64_t step,
                          void *stream);

/*
 * Input pipeline (HOST functions, host pointers): libsvm text (`label id:val id:val ...`, data_loader.py:12-47) -> dense
 * [N, nfield] arrays.  armnet_libsvm_count_lines gives an upper bound for N; armnet_libsvm_parse fills ids (int32),
 * values, labels for every well-formed line (exactly nfield `id:val` pairs) and counts the skipped ones, like the
 * reference's per-line try/except (data_loader.py:37-44).  armnet_b200/data.py caches the result as raw binary files.
 */
int armnet_libsvm_count_lines(const char *path, int64_t *n_lines);
int armnet_libsvm_parse(const char *path, int nfield, int64_t capacity, int32_t *ids, float *values, float *labels,
                        int64_t *n_rows, int64_t *n_skipped);

/* Number of kernels the last armnet_fused_fwd_f32 call on this thread launched (bench bookkeeping). */
int armnet_last_launch_count(void);

/* Tuning / experiment switches (process-global; defaults come from the ARMNET_* environment variables, which are read
 * ONCE, at the first use, never in a launch path).  Keys: "tmem", "tmem_rows", "mma", "mma_split_rna", "mma_warps", "force_nw",
 * "force_look", "lockstep", "no_tma_gather", "no_tma_store", "gemm_1cta".  value -1 = library default.
 * Returns ARMNET_ERR_UNSUPPORTED for an unknown key.  Not part of the reference's surface: tests and A/B runs only. */
int armnet_set_tuning(const char *key, int value);

#ifdef __cplusplus
}
#endif
#endif /* ARMNET_B200_H_ */
