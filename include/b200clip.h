/*
 * libb200clip -- C ABI of the B200 (sm_100a) CLIP dual-encoder hot path.
 *
 * The reference (zhuluntsai/Construction-CLIP) has no FFI/plugin layer: its hot path
 * sits behind the Python API of the third-party `clip` package
 *   clip.load                CLIP/predict.py:12, CLIP/train.py:105, CLIP_prefix_caption/parse_coco.py:20
 *   model(image, text)       CLIP/predict.py:46, CLIP/train.py:161, parse_coco.py:45,50
 *   model.encode_image       parse_coco.py:43, application.py:97
 *   CE(lpi)+CE(lpt) / 2      CLIP/train.py:162-166
 *   loss.backward()          CLIP/train.py:168
 * and, below that, behind torch operators (nn.Conv2d, nn.LayerNorm, nn.MultiheadAttention,
 * nn.Linear, QuickGELU, matmul, CrossEntropyLoss).  Each entry point below replaces one
 * such operator (or a fused group of them); the comment on each names what it replaces.
 * The Python package `clip/` shipped with this repository binds them through ctypes
 * (see INTEGRATION.md).
 *
 * Conventions
 *  - every function returns 0 on success, a negative B200CLIP_ERR_* otherwise;
 *    b200clip_last_error() returns a thread-local message for the last failure.
 *  - all data pointers are DEVICE pointers owned by the caller; the library never frees
 *    or retains them.  `stream` is a cudaStream_t passed as void*.
 *  - bf16 = __nv_bfloat16 storage; "rows x cols, ld" = row-major with a row pitch of `ld`
 *    elements.  Row pitches of bf16 matrices fed to the tensor-core GEMM must be multiples
 *    of 8 elements (16 bytes, a TMA requirement) and base pointers 16-byte aligned.
 *  - no function synchronises the device or allocates memory after ctx creation, so all of
 *    them can be captured into a CUDA graph.
 *  - there is no CPU fallback: without an sm_100 device every call fails with
 *    B200CLIP_ERR_DEVICE.
 */
#ifndef B200CLIP_H
#define B200CLIP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200CLIP_ABI_VERSION 1

enum {
    B200CLIP_OK = 0,
    B200CLIP_ERR_ARG = -1,     /* bad argument (shape, alignment, enum) */
    B200CLIP_ERR_DEVICE = -2,  /* no sm_100 device / wrong device */
    B200CLIP_ERR_CUDA = -3,    /* a CUDA runtime / driver call failed */
    B200CLIP_ERR_UNSUPPORTED = -4
};

/* operand storage order for the tensor-core GEMM */
enum {
    B200CLIP_MAJOR_K = 0,  /* operand stored [rows(M or N), K], K contiguous  */
    B200CLIP_MAJOR_MN = 1  /* operand stored [K, rows(M or N)], M/N contiguous */
};

/* GEMM epilogues */
enum {
    B200CLIP_EPI_NONE = 0,          /* C = acc (+bias)                                      */
    B200CLIP_EPI_QUICKGELU = 1,     /* C = quickgelu(acc+bias); optional preact = acc+bias  */
    B200CLIP_EPI_RESIDUAL = 2,      /* C = acc + bias + aux  (aux has C's dtype: bf16 or f32) */
    B200CLIP_EPI_QUICKGELU_BWD = 3, /* C = acc * quickgelu'(aux)                            */
    /* The same pair with the DERIVATIVE saved instead of the pre-activation: `preact` (forward) / `aux` (backward)
     * is a uint8 [M,N] matrix (row pitch ldc / ldaux BYTES, a multiple of 16) of codes round(210 quickgelu'(x) + 22),
     * computed from the fp32 pre-activation.  Half the bytes of the bf16 pre-activation in both directions. */
    B200CLIP_EPI_QUICKGELU_D8 = 4,     /* C = quickgelu(acc+bias); preact (required) = code of quickgelu'(acc+bias) */
    B200CLIP_EPI_QUICKGELU_BWD_D8 = 5  /* C = acc * (aux - 22) / 210                                                */
};

enum { B200CLIP_DT_BF16 = 0, B200CLIP_DT_F32 = 1, B200CLIP_DT_U8 = 2 /* im2col_patch input only */ };

typedef struct b200clip_ctx b200clip_ctx;

int b200clip_abi_version(void);
const char* b200clip_last_error(void);
/* Number of CUDA kernels this library has launched in this process (monotonic). */
uint64_t b200clip_launch_count(void);
/* Creates the per-device context (resolves the driver's tensor-map encoder, sets kernel
 * attributes).  Fails with B200CLIP_ERR_DEVICE when `device` is not compute capability 10.x. */
int b200clip_ctx_create(b200clip_ctx** out, int device);
int b200clip_ctx_destroy(b200clip_ctx* ctx);

/* ---- nn.Linear / F.linear (+ fused epilogue); its dgrad and wgrad -------------------------
 * C[M,N] = sum_k A(m,k) * B(n,k)  (+ bias[n]) -> epilogue.          tcgen05 / TMEM / TMA.
 *   forward  y = x W^T + b : A = x [M,K] K-major,  B = W [N,K] K-major
 *   dgrad    dx = dy W     : A = dy [M,N'] K-major, B = W [N',K'] as MN-major (N := K', K := N')
 *   wgrad    dW = dy^T x   : A = dy [T,N'] MN-major (M := N'), B = x [T,K'] MN-major (N := K'), K := T
 * a_major/b_major: B200CLIP_MAJOR_*.  lda/ldb: row pitch (elements) of the STORED matrix.
 * out_dtype: B200CLIP_DT_BF16 or _F32.  bias: bf16 [N] or NULL.  aux: [M,N] (ldaux) or NULL -- bf16,
 * except EPI_RESIDUAL with an fp32 C, where aux is fp32 too (the fp32 residual stream).
 * preact: bf16 [M,N] (ldc pitch) or NULL (EPI_QUICKGELU only).
 * scale: optional device fp32 scalar; acc is multiplied by it before bias (NULL = 1).
 * colsum: optional fp32 [N] (EPI_QUICKGELU_BWD only); the column sums of C (after the epilogue) are
 * ACCUMULATED into it -- the c_fc bias gradient falls out of the c_proj dgrad instead of a separate
 * reduction pass over C.
 * split_k > 1 splits the reduction over `split_k` CTAs per tile and ACCUMULATES into C with
 * fp32 atomics (requires out_dtype F32, EPI_NONE, no bias; C must be pre-zeroed or hold a
 * value to accumulate onto).  split_k = 0 lets the library choose (only when out is F32).
 * accumulate != 0 with out F32 adds to C instead of overwriting (atomic). */
int b200clip_gemm_bf16(b200clip_ctx* ctx, const void* A, int64_t lda, int a_major, const void* B, int64_t ldb,
                       int b_major, void* C, int64_t ldc, int out_dtype, const void* bias, const void* aux,
                       int64_t ldaux, void* preact, const float* scale, float* colsum, int64_t M, int64_t N,
                       int64_t K, int epilogue, int split_k, int accumulate, void* stream);

/* ---- clip.model.LayerNorm (fp32 statistics, eps, affine) ------------------------------------
 * y[r,:] = LN(x[src(r),:]) * gamma + beta ; src(r) = row_index ? row_index[r] : r.
 * A negative src(r) reads neg_row (bf16 [d]; zeros when NULL) instead of x -- the class token.
 * Optional fused add before normalisation (vision token assembly, replaces
 * cat(class_embedding, conv) + positional_embedding + ln_pre):
 *   v = x[src(r),:] + add[(r % add_period),:]   when add != NULL.
 * If pre_out != NULL the pre-normalisation value v is also stored (bf16, ldy pitch).
 * mean/rstd: fp32 [rows] or NULL.  x_dtype / y_dtype: B200CLIP_DT_BF16 or _F32 -- the residual stream
 * is kept in fp32 (ln_pre writes fp32; ln_1 / ln_2 / ln_post / ln_final read fp32 and write bf16 GEMM
 * operands); gamma, beta, neg_row, add, pre_out are always bf16. */
int b200clip_layernorm_fwd(b200clip_ctx* ctx, const void* x, int64_t ldx, const int32_t* row_index,
                           const void* neg_row, const void* add, int64_t add_period, const void* gamma,
                           const void* beta, void* y, int64_t ldy, void* pre_out, float* mean, float* rstd,
                           int64_t rows, int64_t d, float eps, int x_dtype, int y_dtype, void* stream);
/* dx[dst(r),:] = (dres ? dres[r,:] : 0) + LN'(dy[r,:]) ; dgamma/dbeta fp32 [d] ACCUMULATED (atomics).
 * x (x_dtype bf16 or f32) is read with the same src(r) mapping as the forward; dx is written at
 * dst(r) = src(r).  dy, dres, dx are bf16 (the gradient stream is bf16).  dx_colsum: optional fp32 [d],
 * ACCUMULATES the column sums of dx (the bias gradient of the Linear that produced this LayerNorm's
 * input through the residual stream). */
int b200clip_layernorm_bwd(b200clip_ctx* ctx, const void* dy, int64_t lddy, const void* x, int64_t ldx,
                           const int32_t* row_index, const void* gamma, const float* mean, const float* rstd,
                           const void* dres, int64_t lddres, void* dx, int64_t lddx, float* dgamma, float* dbeta,
                           float* dx_colsum, int64_t rows, int64_t d, int x_dtype, void* stream);

/* ---- nn.MultiheadAttention(need_weights=False) core: softmax(q k^T / 8 + mask) v, head_dim 64 ----
 * qkv: bf16 [B*S, 3*H*64] (the packed in_proj output: q | k | v, each head-major), out: bf16 [B*S, H*64].
 * causal != 0 applies upstream's build_attention_mask (-inf strictly above the diagonal; padding is
 * NOT masked).  S <= 128 runs the single-tile kernels; longer sequences (ViT-B/16: 197, ViT-L/14: 257,
 * ViT-L/14@336px: 577 tokens) stream the keys/values in 128-row blocks. */
int b200clip_attn_fwd(b200clip_ctx* ctx, const void* qkv, void* out, float* lse, int64_t B, int64_t S, int64_t H,
                      int causal, void* stream);
/* dqkv: bf16 [B*S, 3*H*64].  `out` and `lse` are the forward's results (lse: fp32 [B*H*S], the row
 * log-sum-exp in the log2 domain; pass a buffer to b200clip_attn_fwd when a backward will follow,
 * NULL otherwise); probabilities are recomputed from q,k and lse.  workspace: device scratch of at least
 * B*H*S floats, required when S > 128 (rowsum(dO o O) for the two-pass blocked backward), else may be NULL. */
int b200clip_attn_bwd(b200clip_ctx* ctx, const void* qkv, const void* out, const float* lse, const void* dout,
                      void* dqkv, void* workspace, int64_t workspace_bytes, int64_t B, int64_t S, int64_t H, int causal,
                      void* stream);
/* Packed ("varlen") form for the text tower: sample b owns rows cu[b] .. cu[b+1]-1 (int32 [B+1], at most
 * S_max <= 128 rows each) of qkv / out / dout / dqkv [total_rows, ...]; lse is fp32 [total_rows * H],
 * indexed [row * H + head].  With upstream's causal mask nothing after a caption's EOT token can reach
 * its pooled feature (clip/model.py: build_attention_mask + x[arange, text.argmax(-1)]), so those
 * positions need not exist at all: pack each caption to EOT + 1 tokens. */
/* Rows cu[B] .. total_rows-1 of out / dqkv (surplus of a static row count) are zero-filled by the kernels. */
int b200clip_attn_fwd_varlen(b200clip_ctx* ctx, const void* qkv, void* out, float* lse, const int32_t* cu, int64_t B,
                             int64_t S_max, int64_t H, int64_t total_rows, int causal, void* stream);
int b200clip_attn_bwd_varlen(b200clip_ctx* ctx, const void* qkv, const void* out, const float* lse, const void* dout,
                             void* dqkv, const int32_t* cu, int64_t B, int64_t S_max, int64_t H, int64_t total_rows,
                             int causal, void* stream);

/* ---- token_embedding(text) + positional_embedding  (clip.model.CLIP.encode_text) ---------------
 * ids int32 [B,S]; table bf16 [V,d]; pos bf16 [S,d]; out bf16 or f32 (out_dtype) [B*S,d]; eot_row int32 [B] receives
 * b*S + argmax_s ids[b,s] (first maximum) -- the row upstream pools at. */
int b200clip_embed_tokens_fwd(b200clip_ctx* ctx, const int32_t* ids, const void* table, const void* pos, void* out,
                              int out_dtype, int32_t* eot_row, int64_t B, int64_t S, int64_t d, int64_t vocab,
                              void* stream);
/* dtable fp32 [V,d] and dpos fp32 [S,d] are ACCUMULATED into (atomics). */
int b200clip_embed_tokens_bwd(b200clip_ctx* ctx, const int32_t* ids, const void* dout, float* dtable, float* dpos,
                              int64_t B, int64_t S, int64_t d, int64_t vocab, void* stream);

/* ---- packed text tower -------------------------------------------------------------------------------
 * Under upstream's causal mask (clip/model.py: build_attention_mask) nothing after a caption's EOT token can
 * reach the pooled feature x[arange, text.argmax(-1)], and those positions receive exactly-zero gradients:
 * caption b keeps len_b = argmax_s ids[b,s] + 1 rows, stored back to back.
 * text_pack_plan: cu int32 [B+1] = exclusive prefix sum of len (cu[0] = 0), eot_row int32 [B] = cu[b+1]-1 (the
 * row upstream pools at, in the packed layout).  rows_cap = rows of the caller's packed buffers (>= cu[B]
 * expected; cu is clamped to it so that a too-small buffer truncates captions instead of corrupting memory). */
int b200clip_text_pack_plan(b200clip_ctx* ctx, const int32_t* ids, int32_t* cu, int32_t* eot_row, int64_t B, int64_t S,
                            int64_t rows_cap, void* stream);
/* embed_tokens_fwd into the packed layout: out [rows_total, d]; rows cu[B] .. rows_total-1 (the surplus of a
 * static, CUDA-graph friendly row count) are zero-filled.  The packed attention kernels zero-fill the same
 * rows of their outputs, so every later row-wise kernel maps finite zeros to finite values there and the
 * backward keeps them exactly zero. */
int b200clip_embed_tokens_packed_fwd(b200clip_ctx* ctx, const int32_t* ids, const void* table, const void* pos,
                                     const int32_t* cu, void* out, int out_dtype, int64_t B, int64_t S, int64_t d,
                                     int64_t vocab, int64_t rows_total, void* stream);
/* embed_tokens_bwd reading dout bf16 [rows_total, d] in the packed layout. */
int b200clip_embed_tokens_packed_bwd(b200clip_ctx* ctx, const int32_t* ids, const void* dout, const int32_t* cu,
                                     float* dtable, float* dpos, int64_t B, int64_t S, int64_t d, int64_t vocab,
                                     void* stream);

/* ---- row gather / scatter (x[:, 0, :], x[arange, argmax] of the LAST block: only the pooled token's
 * out_proj / ln_2 / MLP is live) --------------------------------------------------------------------------
 * gather: dst[i,:] = src[idx[i],:] for i < n (rows of row_bytes, a multiple of 16; src pitch src_ld_bytes).
 * scatter: dst[idx[i],:] = src[i,:]; zero_first != 0 clears dst [dst_rows, row_bytes] before. */
int b200clip_gather_rows(b200clip_ctx* ctx, const void* src, int64_t src_ld_bytes, const int32_t* idx, void* dst,
                         int64_t n, int64_t row_bytes, void* stream);
int b200clip_scatter_rows(b200clip_ctx* ctx, const void* src, const int32_t* idx, void* dst, int64_t dst_rows,
                          int64_t n, int64_t row_bytes, int zero_first, void* stream);

/* ---- visual.conv1 (Conv2d, kernel = stride = patch, no bias) as im2col + GEMM ---------------------
 * image: [B,3,R,R] bf16 or fp32 (in_dtype) -> cols bf16 [B*g*g, ldcols], ldcols >= 3*p*p (padded
 * columns are zero-filled); column order (c, py, px) matches conv1.weight.view(width, 3*p*p).
 * in_dtype B200CLIP_DT_U8: image holds raw 8-bit pixels (the resized / centre-cropped RGB planes) and the
 * ToTensor + Normalize((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)) steps of
 * upstream's clip._transform (CLIP/train.py:56 `self.preprocess`) are applied on the fly:
 *   x = (u / 255 - mean[c]) / std[c]  -- a quarter of the host->device bytes of an fp32 batch. */
int b200clip_im2col_patch(b200clip_ctx* ctx, const void* image, int in_dtype, void* cols, int64_t ldcols, int64_t B,
                          int64_t R, int64_t patch, void* stream);

/* ---- Resize(n_px, BICUBIC) + CenterCrop(n_px) of clip._transform on the GPU (CLIP/train.py:56, CLIP/predict.py:31) ----
 * src: decoded RGB pixels, uint8 [H, W, 3] (row pitch src_pitch bytes) -> dst uint8 [3, R, R], bit-exact with Pillow's 8-bit
 * resampler (two separable passes, 22-bit fixed-point coefficients).  The tables are built by the host
 * (construction_clip_b200/data.py) for the R output columns / rows that survive the crop: xbounds / ybounds int32 [R, 2] =
 * (first source index, tap count), xcoef / ycoef int32 [R, xk] / [R, yk].  The horizontal pass runs on source rows
 * [row0, row0 + rows) -- the rows the vertical taps read -- into tmp uint8 [rows, R, 3] (caller-provided). */
int b200clip_resize_crop_u8(b200clip_ctx* ctx, const void* src, int64_t H, int64_t W, int64_t src_pitch,
                            const int32_t* xbounds, const int32_t* xcoef, int64_t xk, const int32_t* ybounds,
                            const int32_t* ycoef, int64_t yk, int64_t row0, int64_t rows, void* tmp, void* dst, int64_t R,
                            void* stream);

/* ---- small fused helpers ---------------------------------------------------------------------- */
/* colsum: out[n] (+)= sum_m x[m,n]; x bf16 [M,N] (ldx); out fp32 [N] accumulated (bias gradients). */
int b200clip_colsum(b200clip_ctx* ctx, const void* x, int64_t ldx, float* out, int64_t M, int64_t N, void* stream);
/* vision token-assembly backward: dpre bf16 [B*n, d] -> dpatch bf16 [B*(n-1), d] (rows 1..n-1 of each
 * sample), dpos fp32 [n,d] += sum_b dpre[b,t,:], dcls fp32 [d] += sum_b dpre[b,0,:]. */
int b200clip_vision_assemble_bwd(b200clip_ctx* ctx, const void* dpre, void* dpatch, float* dpos, float* dcls,
                                 int64_t B, int64_t n, int64_t d, void* stream);
/* y = x / ||x||_2 per row; x fp32 [B,E] -> y fp32 [B,E], inv_norm fp32 [B]. */
int b200clip_l2norm_fwd(b200clip_ctx* ctx, const float* x, float* y, float* inv_norm, int64_t B, int64_t E,
                        void* stream);
/* dx = (dy - y * <y,dy>) * inv_norm, output bf16 (feeds the projection dgrad/wgrad GEMMs). */
int b200clip_l2norm_bwd(b200clip_ctx* ctx, const float* dy, const float* y, const float* inv_norm, void* dx_bf16,
                        int64_t B, int64_t E, void* stream);
/* dtype conversion of flat buffers (n elements) */
int b200clip_cast_f32_to_bf16(b200clip_ctx* ctx, const float* src, void* dst, int64_t n, void* stream);
/* hi = bf16(src), lo = bf16(src - hi): lets a bf16 GEMM consume an fp32 operand at ~16 mantissa bits
 * (C = hi.W, then C += lo.W).  Used for the final feature projection (clip/model.py: `x @ self.proj`,
 * `x @ self.text_projection`), whose input rounding would otherwise dominate the logit error. */
int b200clip_split_f32_to_bf16(b200clip_ctx* ctx, const float* src, void* hi, void* lo, int64_t n, void* stream);
int b200clip_cast_bf16_to_f32(b200clip_ctx* ctx, const void* src, float* dst, int64_t n, void* stream);

/* ---- logit_scale.exp() * I @ T.t()  (CLIP.forward) ---------------------------------------------
 * img fp32 [Bi,E] and txt fp32 [Bt,E] are the L2-normalised features; logit_scale is the device
 * scalar parameter (the kernel applies exp).  logits fp32 [Bi,Bt] (= logits_per_image). */
int b200clip_logits(b200clip_ctx* ctx, const float* img, const float* txt, const float* logit_scale, float* logits,
                    int64_t Bi, int64_t Bt, int64_t E, void* stream);

/* Scratch bytes the two loss entry points below need (caller-allocated device memory, contents
 * are not preserved between calls); -1 on bad arguments. */
int64_t b200clip_clip_loss_workspace_bytes(b200clip_ctx* ctx, int64_t Bl, int64_t Bg, int64_t E);

/* ---- fused similarity + symmetric softmax cross-entropy (CLIP/train.py:162-166) -----------------
 * Local-rows formulation for data parallelism: this rank owns rows [row0, row0+Bl) of the global
 * batch Bg.  img_all / txt_all fp32 [Bg,E] are the (all-gathered) normalised features.
 * Forward: for the local rows i computes  lse_i[i] = logsumexp_j s*<img_i,txt_j>,
 * lse_t[i] = logsumexp_j s*<txt_i,img_j>, diag, and
 *   loss_sum[0] += sum_i (lse_i[i] - diag_i) ; loss_sum[1] += sum_i (lse_t[i] - diag_i),
 *   correct[0] += #{i : argmax_j logits_per_image[i,j] == i}   (CLIP/train.py:173)
 * without ever writing the Bg x Bg logits.  The global loss is (loss_sum[0]+loss_sum[1])/(2*Bg),
 * summed over ranks.  lse_i, lse_t: fp32 [Bl]. */
int b200clip_clip_loss_fwd(b200clip_ctx* ctx, const float* img_all, const float* txt_all, const float* logit_scale,
                           int64_t row0, int64_t Bl, int64_t Bg, int64_t E, float* lse_i, float* lse_t,
                           float* loss_sum, int32_t* correct, void* workspace, int64_t workspace_bytes,
                           void* stream);
/* Backward for the local rows given the GLOBAL row-LSE vectors lse_i_all / lse_t_all fp32 [Bg]
 * (all-gathered; identical to the local ones when Bl == Bg):
 *   d_img[i] = g*s/(2Bg) * sum_j (p_img[i,j] + p_txt[j,i] - 2*delta_ij) txt_j      (fp32 [Bl,E])
 *   d_txt[i] = g*s/(2Bg) * sum_j (p_txt[i,j] + p_img[j,i] - 2*delta_ij) img_j
 *   d_logit_scale[0] += the local rows' share of dL/dlogit_scale
 * where p_img[i,j] = exp(s<img_i,txt_j> - lse_i[i]), p_txt[i,j] = exp(s<txt_i,img_j> - lse_t[i]),
 * g = *grad_out (device scalar, NULL = 1).  Exact gradients for this rank's slice, no gradient
 * collective needed. */
int b200clip_clip_loss_bwd(b200clip_ctx* ctx, const float* img_all, const float* txt_all, const float* logit_scale,
                           const float* lse_i_all, const float* lse_t_all, const float* grad_out, int64_t row0,
                           int64_t Bl, int64_t Bg, int64_t E, float* d_img, float* d_txt, float* d_logit_scale,
                           void* workspace, int64_t workspace_bytes, void* stream);

/* ---- fp32 check mode (forward only; BASELINE: logits within 1e-4 of the reference) --------------------
 * Plain SIMT fp32 kernels with fp32 activations end to end; weights / biases are the bf16 values of
 * the fast path, upcast exactly.  Together with b200clip_layernorm_fwd (f32 in / f32 out),
 * b200clip_embed_tokens_fwd (f32 out), b200clip_l2norm_fwd and b200clip_logits they form a second,
 * tensor-core-free implementation of the forward used to validate the model wiring to 1e-4. */
/* C[M,N] = quickgelu?(A[M,K] . B(n,k) + bias[n]) + residual[M,N]; B bf16 [N,K] (MAJOR_K) or [K,N] (MAJOR_MN). */
int b200clip_check_gemm_f32(b200clip_ctx* ctx, const float* A, int64_t lda, const void* B_bf16, int64_t ldb,
                            int b_major, const void* bias_bf16, const float* residual, int64_t ldres, float* C,
                            int64_t ldc, int64_t M, int64_t N, int64_t K, int quickgelu, void* stream);
/* qkv fp32 [B*S, 3*H*64] -> out fp32 [B*S, H*64]; S <= 1024. */
int b200clip_check_attn_fwd_f32(b200clip_ctx* ctx, const float* qkv, float* out, int64_t B, int64_t S, int64_t H,
                                int causal, void* stream);
/* image fp32 [B,3,R,R] -> cols fp32 [B*g*g, ldcols] (columns >= 3*p*p zero filled). */
int b200clip_check_im2col_f32(b200clip_ctx* ctx, const float* image, float* cols, int64_t ldcols, int64_t B, int64_t R,
                              int64_t patch, void* stream);

/* ---- AdamW step (CLIP/train.py:143,169) on flat buffers: fp32 master weights, bf16 shadow -------
 * The update is transformers.AdamW's (the optimiser the reference constructs, correct_bias=True):
 *   p -= lr * sqrt(1 - beta2^step) / (1 - beta1^step) * m / (sqrt(v) + eps);  then  p -= lr * wd * p
 * (eps is added to the un-corrected sqrt(v); torch.optim.AdamW differs).  grad fp32 (multiplied by grad_scale).
 * hyper_dev: optional DEVICE float[3] = {lr, 1 - beta1^step, 1 - beta2^step}; when non-NULL it overrides
 * `lr` / `step`, so that a captured CUDA graph of the step can be replayed with a changing schedule. */
int b200clip_adamw(b200clip_ctx* ctx, float* master, void* param_bf16, const float* grad, float* m, float* v,
                   int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                   int64_t step, const float* hyper_dev, void* stream);

/* Same update with the gradient stored in bf16: the wire format of the sharded (ZeRO-1) gradient reduce-scatter --
 * each rank casts its flat fp32 gradient chunk to bf16, the chunk is reduce-scattered in bf16 (half the NVLink
 * bytes) and this rank's shard is consumed directly. */
int b200clip_adamw_g16(b200clip_ctx* ctx, float* master, void* param_bf16, const void* grad_bf16, float* m, float* v,
                       int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                       int64_t step, const float* hyper_dev, void* stream);

/* ---- peer-memory all-gather over NVLink / NVSwitch (csrc/peer.cu) ---------------------------------
 * Replaces the NCCL all-gather of the normalised features and of the row log-sum-exps that the data-parallel
 * InfoNCE needs between the towers and the loss (the reference itself is single-GPU:
 * CLIP/train.py:103).  One process per GPU; every rank owns a "symmetric" buffer that all peers map through
 * CUDA IPC, and ONE kernel per rank publishes its block, raises a flag in every peer, waits for the peers'
 * flags and pulls their blocks with loads through the NVLink fabric (protocol: csrc/peer.cu).
 *   b200clip_peer_buffer_bytes  size of a symmetric buffer for `world` ranks and blocks of up to slot_bytes (-1: bad args)
 *   b200clip_peer_alloc         cudaMalloc + zero + export: *ptr = device memory, handle64 = 64-byte IPC handle
 *   b200clip_peer_open          map a peer's buffer from its handle (enables peer access lazily)
 *   b200clip_peer_close / _free unmap a peer's buffer / free the own one
 *   b200clip_peer_allgather     bufs_dev: DEVICE array [world] of the buffers as mapped in this process (own one at
 *                               [rank]); 1..3 segments: dst[s] + r * bytes[s] <- rank r's src[s] (bytes[s] per rank,
 *                               a multiple of 4; sum <= slot_bytes).  Every rank must make the same sequence of calls
 *                               with the same sizes.  Stream-ordered, no host sync, CUDA-graph capturable.
 *   b200clip_peer_status        out2[0] = calls completed, out2[1] = 1 if a wait ever timed out (4 s) */
int64_t b200clip_peer_buffer_bytes(int world, int64_t slot_bytes);
int b200clip_peer_alloc(b200clip_ctx* ctx, int64_t bytes, void** ptr, void* handle64);
int b200clip_peer_open(b200clip_ctx* ctx, const void* handle64, void** ptr);
int b200clip_peer_close(b200clip_ctx* ctx, void* ptr);
int b200clip_peer_free(b200clip_ctx* ctx, void* ptr);
int b200clip_peer_allgather(b200clip_ctx* ctx, const void* const* bufs_dev, int world, int rank, int64_t slot_bytes,
                            int nseg, const void* const* src, void* const* dst, const int64_t* bytes, void* stream);
int b200clip_peer_status(b200clip_ctx* ctx, const void* own_buf, int64_t* out2);

#ifdef __cplusplus
}
#endif
#endif /* B200CLIP_H */
