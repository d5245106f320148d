/*
 * libovk — C ABI of the B200-native (sm_100a) kernels behind OpenVision's ViT image tower and CLIP contrastive loss.
 *
 * The reference (zer0int/OpenVision) has no native code and no C-level plugin API: its hot path is the PyTorch module
 * surface in src/convert_upload/open_clip/{transformer,model,loss}.py, and every FLOP is an ATen/cuBLAS/cuDNN call.
 * Each entry point below therefore cites the reference call site (file:line, relative to the reference root) whose
 * ATen dispatch it replaces.  The Python host layer (openvision_b200/) binds these with ctypes; INTEGRATION.md shows
 * the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C types only: raw DEVICE pointers, explicit sizes / leading dimensions in ELEMENTS, `void* stream`
 *     (a cudaStream_t; NULL = legacy default stream).  No torch / C++ types cross the boundary.
 *   - the caller owns all memory (inputs, outputs, saved-for-backward tensors, workspaces); the library never
 *     allocates device memory and keeps no pointer after a call returns.
 *   - kernels are enqueued on `stream`; no call synchronises the device.
 *   - return value: 0 on success, negative OVK_ERR_* otherwise; ovk_last_error() gives a thread-local message.
 *   - bf16 = __nv_bfloat16 (IEEE bfloat16, 2 bytes); "f32" = float.
 *   - requires compute capability 10.x (tcgen05 / TMEM / TMA); anything else returns OVK_ERR_ARCH.
 */
#ifndef OVK_H_
#define OVK_H_

#ifdef __cplusplus
extern "C" {
#endif

#define OVK_VERSION 100 /* 0.1.0 */

enum {
  OVK_OK = 0,
  OVK_ERR_SHAPE = -1, /* unsupported / inconsistent sizes            */
  OVK_ERR_ALIGN = -2, /* pointer or leading-dimension alignment       */
  OVK_ERR_ARCH = -3,  /* device is not sm_100                         */
  OVK_ERR_CUDA = -4,  /* CUDA runtime / driver error (launch, tmap)   */
  OVK_ERR_NCCL = -5
};

/* epilogue flags of ovk_gemm_bf16 */
enum {
  OVK_EPI_NONE = 0,
  OVK_EPI_GELU_ERF = 1,   /* nn.GELU()                 transformer.py:232-236 */
  OVK_EPI_GELU_TANH = 2,  /* nn.GELU(approximate=tanh) text tower act_kwargs  */
  OVK_EPI_GELU_QUICK = 3, /* QuickGELU                 transformer.py:33-36   */
  OVK_EPI_ACT_MASK = 3,
  OVK_EPI_BIAS = 4,
  OVK_EPI_RESIDUAL = 8,    /* C = A*B^T + bias + residual (not combined with an activation) */
  OVK_EPI_SAVE_PREACT = 16 /* with an activation: also store the pre-activation A*B^T + bias (saved for backward) */
};

int ovk_version(void);
const char* ovk_last_error(void);
/* 0 when the current device can run the library (compute capability 10.x), OVK_ERR_ARCH otherwise. */
int ovk_device_supported(void);

/* ---------------------------------------------------------------------------------------------------------------
 * Dense projections: C[M,N] = epi(A[M,K] · B[N,K]^T), bf16 operands, fp32 accumulation in TMEM, bf16 output.
 * Replaces F.linear in nn.MultiheadAttention in_proj / out_proj (transformer.py:225,250-252), mlp.c_fc / c_proj
 * (transformer.py:232-236,264) and `pooled @ proj` (transformer.py:645-646).  B is the nn.Linear weight layout
 * [out_features, in_features].  lda/ldb/ldc/ldr in elements, multiples of 8; K, N multiples of 8.  `bias` is f32[N].
 * `residual` (bf16 [M, ldr]) may alias C (in-place residual stream update).
 */
int ovk_gemm_bf16(const void* A, long long lda, const void* B, long long ldb, void* C, long long ldc, int M, int N,
                  int K, const float* bias, const void* residual, long long ldr, int flags, void* stream);
/* Same, plus OVK_EPI_SAVE_PREACT: `preact` (bf16 [M, ldp]) receives A*B^T + bias before the activation — what the
 * backward of mlp.c_fc -> gelu (transformer.py:232-236) needs. */
int ovk_gemm_bf16_ex(const void* A, long long lda, const void* B, long long ldb, void* C, long long ldc, int M, int N,
                     int K, const float* bias, const void* residual, long long ldr, void* preact, long long ldp,
                     int flags, void* stream);
/* LayerNorm folded into the projection that consumes it (transformer.py:254-265: attn(ln_1(x)), mlp(ln_2(x));
 * LayerNorm itself: transformer.py:15-30):
 *   ln(x) W^T + b = rstd_i * (x Wc^T) + d,   Wc[n][k] = W[n][k] gamma[k] - mean_k(W[n][.] gamma[.]),   d = W beta + b
 * sum_k Wc[n][k] = 0 makes the product blind to the row mean, so the GEMM reads the un-normalised residual stream x
 * directly, no normalised copy is ever written, and the epilogue only scales each row by its rstd.
 *   ovk_pack_ln_linear : W [N,K] (f32 or bf16), gamma / beta f32[K], bias f32[N] or NULL -> Wc bf16 [N,K], d f32[N].
 *                        After rounding, a few entries per row are moved to the adjacent bf16 value so that the rounded
 *                        row still sums to ~0 (see csrc/ln_pack.cu).
 *   ovk_gemm_bf16_ln   : C = act(rstd_i * (A B^T) + bias) (+ residual);  B = Wc, bias = d.
 *     row_stats_in  : f32 [stats_parts_in][M][2] = partial (sum_k x, sum_k x^2) per row of A, summed in slot order
 *                     (K = the LayerNorm width, stats_parts_in <= 64); rstd is derived from them with ln_eps.
 *                     NULL: plain GEMM (rstd = 1).
 *     row_stats_out : optional f32 [ceil(N/128)][M][2]: the same statistics of the OUTPUT rows, one slot per 128 output
 *                     columns, plain stores (deterministic; needs N % 64 == 0, N > 128, linear epilogue) - this is how
 *                     out_proj / c_proj (+ residual) hand the next LayerNorm its statistics.
 *   ovk_row_stats      : the statistics (one slot) of a tensor that no GEMM of ours produced (first block). */
int ovk_pack_ln_linear(const void* W, int w_is_f32, long long ldw, const float* gamma, const float* beta,
                       const float* bias, void* Wc, long long ldwc, float* d, int N, int K, void* stream);
int ovk_gemm_bf16_ln(const void* A, long long lda, const void* B, long long ldb, void* C, long long ldc, int M, int N,
                     int K, const float* bias, const float* row_stats_in, int stats_parts_in, float ln_eps,
                     const void* residual, long long ldr, void* preact, long long ldp, float* row_stats_out, int flags,
                     void* stream);
int ovk_row_stats(const void* x, long long ldx, float* stats, int rows, int D, void* stream);
/* Patch-embedding GEMM fused with the class-token / positional-embedding add (transformer.py:610-617):
 *   C[row] = A[row] B^T + row_add[row % row_period],  row_add bf16 [row_period, N] (N % 64 == 0).
 * With A = ovk_im2col_patches(..., lead_rows = 1) (a zero row in every image's cls slot), row_period = tokens per image and
 * row_add[0] = class_embedding + pos[0], row_add[t] = pos[t], C is the finished [B, L, D] token buffer. */
int ovk_gemm_bf16_rowadd(const void* A, long long lda, const void* B, long long ldb, void* C, long long ldc, int M, int N,
                         int K, const void* row_add, int row_period, void* stream);
/* Backward GEMMs (what autograd derives from F.linear): operands are read in place, nothing is transposed in memory.
 *   ovk_gemm_bf16_nn : C[M,N] = alpha * A[M,K] * B[K,N]      (B row-major [K,N])   dX = dY * W
 *                      optional fused GELU backward: C = alpha * (A*B) (.) act'(preact), preact bf16 [M, ldp], act = OVK_EPI_GELU_*
 *   ovk_gemm_bf16_tn : C[M,N] = alpha * A[K,M]^T * B[K,N]    (A, B row-major)      dW = dY^T * X
 * C is bf16 or (c_is_f32) fp32.  lda/ldb/ldc in elements; M (tn), N multiples of 8. */
int ovk_gemm_bf16_nn(const void* A, long long lda, const void* B, long long ldb, void* C, long long ldc, int c_is_f32,
                     int M, int N, int K, float alpha, const void* preact, long long ldp, int act, void* stream);
int ovk_gemm_bf16_tn(const void* A, long long lda, const void* B, long long ldb, void* C, long long ldc, int c_is_f32,
                     int M, int N, int K, float alpha, void* stream);
/* ovk_gemm_bf16_nn with the fused GELU backward (C = alpha * (A B) (.) act'(preact), bf16) that ALSO leaves the partial column
 * sums of C — the bias gradient of the layer whose pre-activation gradient it is (autograd: db = dU.sum(0), transformer.py:233
 * c_fc) — in colsum_ws: f32 [ovk_gemm_colsum_rows(M)][N], one row per 64 rows of C, plain stores.  ovk_colsum_f32 reduces them
 * (out ACCUMULATED: zero it first).  Replaces a stand-alone pass over the [M, N] tensor. */
long long ovk_gemm_colsum_rows(int M);
int ovk_gemm_bf16_nn_dact_colsum(const void* A, long long lda, const void* B, long long ldb, void* C, long long ldc, int M, int N,
                                 int K, float alpha, const void* preact, long long ldp, int act, float* colsum_ws, void* stream);
int ovk_colsum_f32(const float* x, int rows, int cols, float* out, void* stream);
/* C[M,N] = alpha * (*alpha_dev) * op(A) * op(B), op = the storage flags below; bf16 operands, bf16 or (c_is_f32) fp32 output.
 *   a_mn = 0: A stored [M,K] (K contiguous);  a_mn = 1: A stored [K,M]
 *   b_mn = 0: B stored [N,K] (the nn.Linear layout);  b_mn = 1: B stored [K,N]           (a_mn = 1 with b_mn = 0 is not provided)
 * alpha_dev: optional DEVICE f32 scalar (NULL = 1) multiplied in by the epilogue, so a temperature that lives on the device
 * (logit_scale.exp(), model.py:250,286-293) never has to be read back by the host.  Call sites: ClipLoss.get_logits
 * (loss.py:102-118: logit_scale * image_features @ text_features.T, fp32 logits) and its autograd. */
int ovk_gemm_bf16_scaled(const void* A, long long lda, int a_mn, const void* B, long long ldb, int b_mn, void* C,
                         long long ldc, int c_is_f32, int M, int N, int K, float alpha, const float* alpha_dev, void* stream);
/* y = a + b, bf16, n elements (multiple of 8): the residual add of transformer.py:263-264 as a stand-alone call, used when
 * forward hooks sit on the branch's modules (cliptoolsoptimized.py:480-489) and the add cannot ride in a GEMM epilogue. */
int ovk_add_bf16(const void* a, const void* b, void* y, long long n, void* stream);
/* The activation as a standalone module call (nn.GELU / QuickGELU hooked by the ov-* scripts, transformer.py:33-36,
 * 232-236): y = act(x), and dx = dy * act'(x); bf16, n elements (multiple of 8), act = OVK_EPI_GELU_*. */
int ovk_act_fwd(const void* x, void* y, long long n, int act, void* stream);
int ovk_act_bwd(const void* x, const void* dy, void* dx, long long n, int act, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * LayerNorm over the last dimension (biased variance, fp32 statistics), transformer.py:15-30 (LayerNorm /
 * LayerNormFp32, eps = 1e-6 in this fork, transformer.py:458).  x, y: bf16 [rows, D] with row strides ldx / ldy
 * (elements, multiples of 8); gamma, beta: f32[D].  mean / rstd: optional f32[rows] saved for the backward pass.
 * D % 8 == 0, D <= 8192.
 */
int ovk_layernorm_fwd(const void* x, long long ldx, void* y, long long ldy, const float* gamma, const float* beta,
                      float* mean, float* rstd, int rows, int D, float eps, void* stream);
/* dx = LN'(x)·dy (+ dres) ; dgamma/dbeta (f32[D]) are ACCUMULATED atomically (zero them first).  dx may alias dy or
 * dres.  dres (bf16 [rows, lddres], nullable) is the gradient arriving through the residual connection around the
 * normalised branch (transformer.py:263-264: x = x + f(ln(x))), added in the same pass. */
int ovk_layernorm_bwd(const void* dy, long long lddy, const void* x, long long ldx, const float* gamma,
                      const float* mean, const float* rstd, const void* dres, long long lddres, void* dx,
                      long long lddx, float* dgamma, float* dbeta, int rows, int D, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Patch embedding, transformer.py:469,610-617: Conv2d(3->D, kernel = stride = P, no bias) as an im2col GEMM,
 * then x = cat([class_embedding, patches]) + positional_embedding.
 *   ovk_im2col_patches : images [B,3,H,W] (f32 when img_is_f32, else bf16, NCHW contiguous) -> cols bf16
 *                        [B*(gh*gw + lead_rows), ldc], column order (c, ph, pw) = conv1.weight.reshape(D, 3*P*P);
 *                        columns [3P², ldc) are zero-filled.  lead_rows = 1 puts one all-zero row in front of each
 *                        image's patches, so that the GEMM output already has the [B, L = N+1, D] token layout.
 *   ovk_embed_assemble : tokens[b,0,:] = cls + pos[0];  tokens[b,1+n,:] = patch[b,n,:] + pos[1+n]   (bf16 out)
 *                        patch: bf16 [B*(N + patch_lead_rows), D] (GEMM output; may alias tokens when
 *                        patch_lead_rows = 1); cls f32[D]; pos f32[L,D], L = N+1.
 */
int ovk_im2col_patches(const void* images, int img_is_f32, void* cols, long long ldc, int B, int H, int W, int P,
                       int lead_rows, void* stream);
int ovk_embed_assemble(const void* patch, int patch_lead_rows, const float* cls, const float* pos, void* tokens, int B,
                       int N, int D, void* stream);
/* conv1 + token assembly as ONE kernel (transformer.py:469,610-617; north-star kernel (1): "TMA-staged im2col GEMM fused with
 * the position-embedding add"): raw pixel rows are fetched by TMA boxes straight out of the NCHW image, converted to the bf16
 * K-major A tile in shared memory and multiplied on tcgen05; no im2col buffer exists in memory.
 *   tokens[b, 0, :]     = pos_table[0, :]                         (class_embedding + positional_embedding[0], or 0)
 *   tokens[b, 1 + p, :] = patch(b, p) . w_packed^T + pos_table[1 + p, :]
 *   images    : [B,3,H,W] NCHW contiguous, f32 (img_is_f32) or bf16
 *   w_packed  : bf16 [D, K'], K' = ovk_patch_embed_kdim(P); column ((c*PG + phg)*R + phl)*PW + pw holds
 *               conv1.weight[d, c, phg*R + phl, pw] (zero where pw >= P or phg*R + phl >= P), PW = 16 (P <= 16) or 32,
 *               R = 64 / PW, PG = ceil(P / R)
 *   pos_table : bf16 [N+1, D] or NULL (plain conv tokens with a zero class-token row: the training path, where
 *               ovk_embed_assemble's autograd owns the add)
 *   tokens    : bf16 [B, N+1, D]
 * ovk_patch_embed_supported: 1 when the geometry fits the kernel (P in {14,16,32}, image rows fit the staging ring);
 * callers use ovk_im2col_patches + ovk_gemm_bf16_rowadd otherwise. */
int ovk_patch_embed_kdim(int P);
int ovk_patch_embed_supported(int img_is_f32, int H, int W, int P, int D);
int ovk_patch_embed(const void* images, int img_is_f32, const void* w_packed, const void* pos_table, void* tokens, int B,
                    int H, int W, int P, int D, void* stream);
/* The tail of the image tower as ONE kernel (transformer.py:599-607 _global_pool, :638-640 ln_post after pooling,
 * :645-646 pooled @ proj; model.py:267 F.normalize; north-star kernel (5) "pooling head and L2-normalise"):
 *   pooled = mean(x[b, 1:, :]) (mode 0, 'avg'), x[b, 0, :] (mode 1, 'tok') or x[b, L-1, :] (mode 2, the text tower's 'last',
 *            model.py:278-284 via text_global_pool)
 *   h      = LayerNorm(pooled; gamma, beta, ln_eps)                (gamma == NULL: skipped)
 *   y      = h @ proj                 proj bf16 [D, E] row-major   (proj == NULL: y = h, E ignored)
 *   out[b] = normalize ? y / max(||y||, norm_eps) : y              (f32 when out_is_f32, else bf16)
 * x bf16 [B, L, D] contiguous, D % 8 == 0, E % 8 == 0.  fp32 between the token load and the output store. */
int ovk_pool_head(const void* x, int B, int L, int D, int mode, const float* gamma, const float* beta, float ln_eps,
                  const void* proj, int E, int normalize, float norm_eps, void* out, int out_is_f32, void* stream);
/* Backward of the patch embedding w.r.t. the input image (the ov-* scripts optimise the image through the tower):
 * dimages[b,c,y,x] = dcols[row(b,y/P,x/P), (c*P + y%P)*P + x%P]  (f32 or bf16 out), dcols = dtokens · conv1.weight. */
int ovk_col2im_patches(const void* dcols, long long ldc, void* dimages, int img_is_f32, int B, int H, int W, int P,
                       int lead_rows, void* stream);
/* out[c] += sum_r x[r,c] (bf16 [rows, ldx] -> f32[cols], ACCUMULATED): bias gradients of F.linear, and with rows = B,
 * cols = L*D the gradient of positional_embedding / class_embedding (transformer.py:615-617). */
int ovk_colsum_bf16(const void* x, long long ldx, int rows, int cols, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Multi-head self-attention core, nn.MultiheadAttention(need_weights=False, attn_mask=None) between in_proj and
 * out_proj (transformer.py:225,239-252): O = softmax(Q K^T / sqrt(hd)) V per (batch, head), flash-style with online
 * softmax (recurrence as in src/models/bpt.py:105-124).
 *   qkv : bf16 [B, L, 3, H, hd] (= in_proj output [B*L, 3*D], rows q|k|v as in in_proj_weight)
 *   out : bf16 [B, L, H*hd]
 *   lse : optional f32 [B, H, L], natural-log sum-exp of the scaled scores (saved for backward)
 * hd: 64 (Ti/B/L towers), 72 (So400m/14) or 80 (H/14); the dims past 64 ride along as a 16-wide operand block.
 */
int ovk_attention_fwd(const void* qkv, void* out, float* lse, int B, int L, int H, int hd, float scale, void* stream);
/* Backward: dqkv (bf16, same [B, L, 3, H, hd] layout) from qkv, the forward output `out`, its gradient `dout`
 * (bf16 [B, L, H*hd]) and the saved lse.  delta: f32 [B, H, L] scratch (rowsum(dout * out), written then read).
 * Scores are recomputed tile by tile; two launches (dQ, then dK/dV), no atomics. */
int ovk_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* delta,
                      int B, int L, int H, int hd, float scale, void* stream);
/* The same two with the additive CAUSAL mask of the stock text tower (transformer.py:757-763 build_causal_mask, applied
 * at model.py:276 `self.transformer(x, attn_mask=self.attn_mask)`): flags = OVK_ATT_CAUSAL masks every key j > query i
 * (-inf before the softmax, so P and dS are exactly zero there).  flags = 0 is the unmasked call above. */
#define OVK_ATT_CAUSAL 1
int ovk_attention_fwd_ex(const void* qkv, void* out, float* lse, int B, int L, int H, int hd, float scale, int flags,
                         void* stream);
int ovk_attention_bwd_ex(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* delta,
                         float* workspace, int B, int L, int H, int hd, float scale, int flags, void* stream);
/* workspace: f32[ovk_attention_bwd_workspace_floats(B, L, H, flags)] or NULL.  With it, the remainder token of L = 128 k + 1
 * (class token + power-of-two grid) is handled outside the 128-wide tiles (attention_bwd_tail_kernel + rank-1 terms in the
 * tile kernels' epilogues) instead of as a third tile row / column; the function returns 0 when the shape has no such token. */
long long ovk_attention_bwd_workspace_floats(int B, int L, int H, int flags);
/* One-pass backward (the default of the Python mirror): the score tiles are walked ONCE (key tile stationary, query tiles
 * streamed): dK / dV accumulate in TMEM as above, and the partial dQ = dS K of every (query tile, key tile) is added into an
 * fp32 [B, L, H, hd] scratch by TMA reduce-add, then converted (x scale) into the q slot of dqkv.  10 instead of 14
 * B H L^2 hd FLOPs and half the exponentials of ovk_attention_bwd_ex, same results up to the fp32 summation order of dQ.
 * workspace: f32[ovk_attention_bwd_fused_workspace_floats(B, L, H, hd, flags)], required (0 = shape not offered); it also
 * holds the remainder-token vectors of ovk_attention_bwd_workspace_floats.  delta as above.
 * Launches: attention_bwd_delta_kernel OR attention_bwd_tail_kernel (which leaves delta and the statistics as well), the tile
 * kernel, attention_bwd_dq_convert_kernel.
 * Kernel: attention_bwd_t_kernel (attention_bwd2.cu: transposed score tiles, keys on the TMEM lanes, P^T / dS^T consumed as
 * TMEM operands, half-tile software pipeline); flags | OVK_ATT_BWD_ONEPASS_V1 selects its predecessor attention_bwd_kernel<fused>. */
#define OVK_ATT_BWD_ONEPASS_V1 2
#define OVK_ATT_BWD_8_WARPS 4 /* attention_bwd_t_kernel with 8 compute warps (32 query columns each; what the Python mirror passes)
                               * instead of 16 (16 columns each; measured 7-10 % slower: twice the TMEM / barrier instructions) */
long long ovk_attention_bwd_fused_workspace_floats(int B, int L, int H, int hd, int flags);
int ovk_attention_bwd_fused(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* delta,
                            float* workspace, int B, int L, int H, int hd, float scale, int flags, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Pooling head, transformer.py:599-607,638-646.
 *   ovk_pool_tokens: mode 0 ('avg', OpenVision): pooled[b] = mean(x[b,1:,:]);  mode 1 ('tok'): pooled[b] = x[b,0,:]
 *                    x bf16 [B,L,D] -> pooled bf16 [B,D]
 *   ovk_l2_normalize: F.normalize(x, dim=-1) (model.py:267,284), y = x / max(||x||2, eps); x bf16 [rows,E] -> y f32 or bf16
 */
int ovk_pool_tokens(const void* x, void* pooled, int B, int L, int D, int mode, void* stream);
int ovk_l2_normalize(const void* x, void* y, int y_is_f32, float* norms, int rows, int E, float eps, void* stream);
/* Backward: dx[b,l,:] of the pooling (every element written), and dx = (dy - y (y·dy)) / max(||x||, eps) of F.normalize
 * (x bf16 [rows,E]; dy f32 or bf16; dx bf16). */
int ovk_pool_tokens_bwd(const void* dpooled, void* dx, int B, int L, int D, int mode, void* stream);
int ovk_l2_normalize_bwd(const void* x, const void* dy, int dy_is_f32, void* dx, int rows, int E, float eps, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * CLIP contrastive loss, loss.py:102-131 (ClipLoss.get_logits + 2x F.cross_entropy), fused: the N x N logits are
 * never written to memory in the forward pass.
 *   z = scale * A_loc · B_all^T   (rows = this rank's image features, cols = all text features);
 *   labels: row i <-> column (row_offset + i)                                   (loss.py:89-100 get_ground_truth)
 *   a_loc : bf16 [n_loc, E]    b_all : bf16 [n_all, E]
 * The temperature is read ON THE DEVICE (scale_dev: f32[1] = logit_scale.exp(), model.py:250,295-315), so no call forces a
 * host synchronisation.
 * ovk_clip_loss_fwd : partial statistics of the column window [col_offset, col_offset + n_cols) of the row block, b_rows =
 *   bf16 [n_cols, E] = the features of THOSE columns (col_offset % 256 == 0; every window but the last a multiple of 256
 *   wide).  One call with col_offset = 0, n_cols = n_all covers everything; several calls with disjoint windows tiling
 *   [0, n_all) let a rank start on its own text block while the all-gather of the others is still in flight.  Writes
 *   diag[i] = z_i,label(i) for the rows whose label falls into the window, and per-tile (max, sum) partials into `workspace`
 *   (f32 scratch of ovk_clip_loss_workspace_floats(n_loc, n_all) elements, shared by all windows).
 * ovk_clip_loss_finalize : merges the partials (all windows must have run) into, all f32, natural-log units:
 *   row_lse[n_loc]  = logsumexp_j z_ij
 *   col_max[n_all], col_sum[n_all] : THIS row block's column statistics, sum_i exp(z_ij - col_max_j) = col_sum_j,
 *                                    so that W ranks can merge their row blocks into global column LSEs.
 * ovk_clip_loss_combine : col_lse[j] = log sum_w exp(col_max[w][j]) * col_sum[w][j]   over `parts` stacked [parts, n_all] arrays
 * ovk_clip_loss_value   : out3[0] = 0.5/n_loc * (sum_i (row_lse_i - diag_i) + sum_i (col_lse[row_offset+i] - diag_i)),
 *                         out3[1], out3[2] = the two sums                             (loss.py:126-129)
 * ovk_clip_loss_grad_logits : G[i,j] = g * (w_row * exp(z_ij - row_lse_i) + w_col * exp(z_ij - col_lse_j)
 *                                     - (w_row + w_col) * [j == row_offset + i]),  bf16 [n_loc, ldg];
 *                         g = *grad_out_dev (the upstream gradient, a device scalar; NULL = 1);
 *                         d_scale_partial (f32[1], ACCUMULATED) += sum_ij G_ij * z_ij / scale.
 *   The feature gradients are then dA = scale * G · B_all and dB = scale * G^T · A_loc (ovk_gemm_bf16_scaled, alpha_dev = scale_dev).
 */
long long ovk_clip_loss_workspace_floats(int n_loc, int n_all);
int ovk_clip_loss_fwd(const void* a_loc, const void* b_rows, int n_loc, int n_cols, int col_offset, int n_all, int E,
                      int row_offset, const float* scale_dev, float* diag, float* workspace, void* stream);
int ovk_clip_loss_finalize(const float* workspace, int n_loc, int n_all, float* row_lse, float* col_max, float* col_sum,
                           void* stream);
int ovk_clip_loss_combine(const float* col_max_parts, const float* col_sum_parts, int parts, int n_all, float* col_lse,
                          void* stream);
int ovk_clip_loss_value(const float* row_lse, const float* col_lse, const float* diag, int n_loc, int row_offset,
                        float* out3, void* stream);
int ovk_clip_loss_grad_logits(const void* a_loc, const void* b_all, int n_loc, int n_all, int E, int row_offset,
                              const float* scale_dev, const float* row_lse, const float* col_lse, float w_row, float w_col,
                              const float* grad_out_dev, void* G, long long ldg, float* d_scale_partial, void* stream);

/* Optimizer step of the training recipe (src/optim/build_optax.py:188-278: clip_by_global_norm -> scale_by_adam(b1, b2,
 * mu_dtype=bf16) -> add_decayed_weights -> scale(lr) -> schedule -> -1; configs/openvision.py:265-289) on flat buffers:
 *   g' = g * gscale [* max_norm / max(||g'||, max_norm) when gnorm_sq != NULL: device scalar sum g^2 over the whole model]
 *   mu = b1 mu + (1-b1) g' (bf16 storage), nu = b2 nu + (1-b2) g'^2 (f32),
 *   p -= lr * ((mu / (1-b1^step)) / (sqrt(nu / (1-b2^step)) + eps) + wd * p);   step counts from 1.
 * ovk_sumsq ACCUMULATES sum x^2 into *out (zero it first). */
int ovk_sumsq(const void* x, int is_bf16, long long n, float* out, void* stream);
int ovk_adamw_step(void* p, int p_is_bf16, const void* g, int g_is_bf16, void* mu, float* nu, long long n, float lr,
                   float b1, float b2, float eps, float wd, int step, float gscale, const float* gnorm_sq, float max_norm,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OVK_H_ */
