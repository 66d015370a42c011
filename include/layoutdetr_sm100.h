/* liblayoutdetr_sm100 — C-ABI of the B200-native LayoutDETR hot path.
 *
 * Plain C: raw device pointers, sizes and a cudaStream_t (passed as void*).  No torch types.
 * All entry points are stream-ordered, stateless and thread-compatible; buffers are caller-owned
 * (the Python host allocates them with the PyTorch caching allocator).  Return value: 0 on
 * success, <0 for argument errors (LD_ERR_*), >0 = cudaError_t.  ld_last_error() returns a
 * human-readable message for the calling thread's last failure.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the
 * salesforce/LayoutDETR checkout).
 */
#ifndef LAYOUTDETR_SM100_H
#define LAYOUTDETR_SM100_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LD_OK               0
#define LD_ERR_INVALID_ARG -1
#define LD_ERR_UNSUPPORTED -2
#define LD_ERR_ALIGNMENT   -3
#define LD_ERR_DRIVER      -4

/* dtype codes used throughout */
#define LD_F32  0
#define LD_BF16 1

const char* ld_last_error(void);
int ld_version(void);
/* number of kernels launched by this library since load / since the last reset (bench.py gpu_launches) */
int64_t ld_launch_count(void);
void ld_launch_count_reset(void);
/* Lane scheduling: cap the grid of the persistent kernels (ld_gemm_bf16) launched on `stream` at `limit` CTAs
 * (<= 0 clears the cap; default = one CTA per SM).  The host runs independent parts of the iteration on parallel
 * streams — the frozen text encoder (training/networks_detr.py:146) next to the latency-bound DETR / ResNet chains —
 * and a capped tensor-bound lane leaves SMs for the others.  The reference has no counterpart (single stream). */
int ld_set_stream_cta_limit(void* stream, int limit);
int ld_get_stream_cta_limit(void* stream);

/* ------------------------------------------------------------------------------------------
 * Batched bf16 GEMM on tcgen05 tensor cores (TMA-staged SWIZZLE_128B tiles, TMEM accumulators)
 * with a fused epilogue.  Replaces every cuBLAS call the reference reaches through
 *   nn.Linear / F.linear            training/networks_detr.py:57-62, training/med.py:109-116,231,292,312
 *   torch.matmul (QK^T, PV)         training/med.py:183,215; nn.MultiheadAttention in
 *                                   training/detr_transformer.py:185,245-246
 *   F.conv2d as im2col GEMM         training/detr_backbone.py:105 (torchvision resnet50),
 *                                   torch_utils/ops/conv2d_gradfix.py:37,42
 * and their autograd backward GEMMs (dgrad / wgrad) via the operand "major" flags.
 *
 *   for every batch (b1, b2):
 *     acc[m, n] = sum_k A[m, k] * B[n, k]                         (fp32 accumulation)
 *     v   = acc * alpha * (alpha_dev ? *alpha_dev : 1)
 *     v   = v * col_scale[n]   (if col_scale)   + col_bias[n] (if col_bias)
 *     v  += R[m, n]            (if R)
 *     aux[m, n] = v            (if aux; pre-activation, dtype of D)
 *     v   = act(v) * post_gain
 *     D[m, n] = v | D[m, n] += v | atomicAdd(D[m, n], v)           (accumulate = 0 | 1 | 2)
 *
 * Operand storage (bf16): mn_major = 0 -> element (row, k) at ptr[row*ld + k]   ("K-major")
 *                         mn_major = 1 -> element (row, k) at ptr[k*ld + row]   ("MN-major")
 * plus batch offset b1*sb1 + b2*sb2 (elements).  ld, sb1, sb2 must be multiples of 8 elements and
 * ptr 16-byte aligned (TMA global-stride rule); the contiguous extent may be any size (TMA zero-fills
 * out-of-bounds reads, the epilogue masks out-of-range rows/columns).
 * ------------------------------------------------------------------------------------------ */
#define LD_ACT_NONE    0
#define LD_ACT_RELU    1
#define LD_ACT_GELU    2   /* exact erf GELU */
#define LD_ACT_LRELU   3   /* slope 0.2 */
#define LD_ACT_SIGMOID 4

typedef struct ld_gemm_operand {
    const void* ptr;
    int64_t ld;
    int64_t sb1, sb2;
    int32_t mn_major;
    int32_t _pad;
} ld_gemm_operand;

typedef struct ld_gemm_desc {
    int32_t M, N, K;
    int32_t nb1, nb2;            /* batch = nb1 * nb2 problems */
    int32_t act;                 /* LD_ACT_* */
    int32_t accumulate;          /* 0 store, 1 read-modify-write, 2 atomic add (required if split_k > 1) */
    int32_t split_k;             /* >= 1 */
    int32_t d_dtype, r_dtype;    /* LD_F32 / LD_BF16 */
    int32_t block_n;             /* 0 = auto, else 64, 128 or 256 */
    int32_t _pad;
    float alpha, post_gain;
    ld_gemm_operand A, B;
    void* D;        int64_t ldd, d_sb1, d_sb2;
    void* aux;      /* optional pre-activation copy, same layout/dtype as D */
    const void* R;  int64_t ldr, r_sb1, r_sb2;
    const float* col_scale; const float* col_bias;
    int64_t col_sb1, col_sb2;    /* batch strides of col_scale/col_bias (0 = shared) */
    const float* alpha_dev;      /* optional device scalar multiplied into alpha (upstream loss gradient) */
    /* fused attention-probability epilogue (requires N <= 256, bf16 D, store-only):
     *   D[m, :] = softmax_n(acc[m, n] * alpha + (key_mask[b1, n] ? mask_value : 0) + (causal && n > m ? mask_value : 0))
     * mask_value = -10000 (BERT additive mask, training/med.py:651-654) or -inf (nn.MultiheadAttention key_padding_mask). */
    int32_t softmax, causal;
    float mask_value;
    int32_t _pad2;
    const uint8_t* key_mask;     /* [nb1, N], 1 = masked, or NULL */
} ld_gemm_desc;

int ld_gemm_bf16(const ld_gemm_desc* desc, void* stream);

/* Implicit-GEMM convolution on the same kernel — replaces the cuDNN convolutions behind nn.Conv2d / F.conv2d of the ResNet-50
 * body (training/detr_backbone.py:82-95, torchvision Bottleneck conv2 3x3 and the stride-2 1x1 downsample), of input_proj's
 * neighbours and of the StyleGAN2 synthesis 3x3 layers (training/networks_stylegan2.py:30-83), forward, data gradient and weight
 * gradient, WITHOUT a patch matrix in HBM: the TMA producer loads boxes of output pixels x 64 channels straight from the NHWC
 * bf16 image at the tap's offset (out-of-image taps are zero-filled by the TMA unit; the convolution stride is the tensor map's
 * traversal stride).  `img` is [B, H, W, C] bf16 contiguous, C % 64 == 0 (mode 1 also takes C == 32: K blocks of 32 channels in
 * 64-byte-swizzled rows, weight matrix K-major, split_k = 1); Ho = (H + 2 pad - KH) / stride + 1 (same for W);
 * a box of `mode == 1 ? 128 : 64` consecutive output pixels must be a rectangle of whole rows / whole images
 * (Wo % box == 0, or box % Wo == 0 with Ho*Wo % box == 0 or box % (Ho*Wo) == 0) — LD_ERR_INVALID_ARG otherwise.
 *   mode 1 (forward, and data gradient as a convolution of dy with the flipped, transposed weights):
 *        D[M = B*Ho*Wo, N] = epilogue( patches(img)[M, K = KH*KW*C] @ desc->B[N, K]^T ),  K ordered (kh, kw, c);  desc->A is ignored
 *   mode 2 (weight gradient):
 *        D[M, N = KH*KW*C] = desc->A[K = B*Ho*Wo, M]^T (mn_major) @ patches(img)[K, N];  desc->B is ignored
 * Everything else (epilogue terms, split_k, accumulate) is as in ld_gemm_bf16; nb1 = nb2 = 1. */
typedef struct ld_conv_geom {
    const void* img;
    int32_t mode;
    int32_t B, H, W, C;
    int32_t Ho, Wo, KH, KW, stride, pad;
    int32_t _pad;
} ld_conv_geom;

int ld_conv_gemm_bf16(const ld_gemm_desc* desc, const ld_conv_geom* geom, void* stream);

/* ------------------------------------------------------------------------------------------
 * bias_act — replaces bias_act_plugin.bias_act (torch_utils/ops/bias_act.cpp:33-97, kernel bias_act.cu:24-148).
 * y = clamp(act(x + b[(i / stepB) % sizeB]) * gain); grad = 1 / 2 evaluate the first / second order
 * gradient forms from (xref, yref, dy) exactly like the reference kernel.  act = reference cuda_idx 1..9
 * (linear, relu, lrelu, tanh, sigmoid, elu, selu, softplus, swish).  dtype: LD_F32 / LD_BF16.  Any pointer
 * except x, y may be NULL.  clamp < 0 disables clamping.
 * ------------------------------------------------------------------------------------------ */
int ld_bias_act(const void* x, const void* b, const void* xref, const void* yref, const void* dy, void* y,
                int dtype, int grad, int act, float alpha, float gain, float clamp,
                int64_t sizeX, int sizeB, int64_t stepB, void* stream);

/* upfirdn2d — replaces upfirdn2d_plugin.upfirdn2d (torch_utils/ops/upfirdn2d.cpp:17-105, kernels upfirdn2d.cu:30,98).
 * Zero-insert upsample -> pad/crop -> FIR (fp32 filter [fh, fw], <= 32x32) -> decimate, strided 4-D tensors
 * (element strides for N, C, H, W; NCHW or channels-last), output size must be
 * (in*up + pad0 + pad1 - f + down) / down per axis. */
int ld_upfirdn2d(const void* x, void* y, int dtype, const float* f, int fh, int fw,
                 int N, int C, int inH, int inW, int outH, int outW,
                 const int64_t* x_strides, const int64_t* y_strides,
                 int upx, int upy, int downx, int downy, int padx0, int padx1, int pady0, int pady1,
                 int flip_filter, float gain, void* stream);

/* fma — replaces torch_utils/ops/fma.py:16 (a * b + c); b / c broadcast through element strides (0 on broadcast
 * dims) in the index space of the contiguous 4-D tensor a. */
int ld_fma_f32(const float* a, const float* b, const float* c, float* y, const int64_t* a_shape,
               const int64_t* b_strides, const int64_t* c_strides, void* stream);

/* LayerNorm (nn.LayerNorm: training/med.py:63,233,318,513; training/detr_transformer.py:191-192,252-254,41).
 * rows x C, C % 128 == 0 and C <= 1024.  fwd optionally saves mean / rstd; bwd accumulates dgamma / dbeta with
 * atomics into caller buffers (the parameters' .grad) and writes dx as bf16 and/or fp32. */
int ld_layernorm_fwd(const void* x, int x_dtype, int64_t ldx, const float* gamma, const float* beta,
                     void* y_bf16, float* y_f32, int64_t ldy, float* mean, float* rstd,
                     int rows, int C, float eps, void* stream);
/* y = LayerNorm(x + res): the residual (bf16 rows, may be NULL) is added in fp32 before the statistics — the tail of a
 * post-norm residual block (hidden = LayerNorm(dense(h) + input), training/med.py:237-242,321-325) when the dense output
 * was stored as bf16. */
int ld_layernorm_res_fwd(const void* x, int x_dtype, int64_t ldx, const void* res_bf16, int64_t ldr,
                         const float* gamma, const float* beta,
                         void* y_bf16, float* y_f32, int64_t ldy, float* mean, float* rstd,
                         int rows, int C, float eps, void* stream);
int ld_layernorm_bwd(const void* dy, int dy_dtype, int64_t lddy, const void* x, int x_dtype, int64_t ldx,
                     const float* mean, const float* rstd, const float* gamma,
                     void* dx_bf16, float* dx_f32, int64_t lddx, float* dgamma, float* dbeta,
                     int rows, int C, void* stream);

/* y = LayerNorm(dropout(x) + res) with the Philox mask of (rng_state, rng_site), group = row * C/8 + column/8 (the element order
 * of ld_dropout on the contiguous [rows, C] tensor); pre_f32 (optional) = dropout(x) + res for ld_layernorm_bwd.  Hidden-state
 * dropout of the post-norm residual blocks: training/med.py:237-242,318-325, training/detr_transformer.py:210-214,270-285.
 * bf16 x / res / y, C % 256 == 0. */
int ld_layernorm_res_dropout_fwd(const void* x_bf16, int64_t ldx, const void* res_bf16, int64_t ldr,
                                 const float* gamma, const float* beta, void* y_bf16, int64_t ldy, float* pre_f32, int64_t ldpre,
                                 float* mean, float* rstd, int rows, int C, float eps,
                                 float dropout_p, const uint32_t* rng_state, uint32_t rng_site, void* stream);
/* Dropout randomness (nn.Dropout sites of training/med.py:96,213,240,318, training/detr_transformer.py:185-194): Philox4x32-10,
 * key = 64-bit seed, counter = (group lo, group hi, site, step), eight 16-bit lanes per group, element dropped when its lane <
 * round(p * 65536).  rng_state: device uint32[4] {seed_lo, seed_hi, step, 0}.  ld_rng_advance: step += 1 (once per iteration,
 * graph-capturable).  ld_dropout: y = keep ? x / (1 - p) : 0 on a contiguous tensor (n % 8 == 0), group = index / 8; x may alias y. */
int ld_rng_advance(uint32_t* rng_state, void* stream);
int ld_dropout(const void* x, void* y, int dtype, int64_t n, float dropout_p, const uint32_t* rng_state, uint32_t rng_site, void* stream);

/* BERT embeddings: y = LayerNorm(word[ids[r]] + pos[r % T]) (training/med.py:74-97) and the scatter-add of its
 * gradient into the (fp32) embedding tables; rows with ids == pad_id get no word gradient (padding_idx). */
int ld_embed_ln_fwd(const int64_t* ids, const float* word, const float* pos, const float* gamma, const float* beta,
                    void* y_bf16, float* pre_f32, float* mean, float* rstd, int rows, int T, int C, float eps, void* stream);
int ld_embed_bwd(const int64_t* ids, const float* dpre, float* dword, float* dpos, int64_t rows, int T, int C,
                 int64_t pad_id, void* stream);

/* Masked row softmax for score matrices too wide for the fused GEMM epilogue (keys > 256), and its backward
 * dS = P * (dP - sum(dP * P)) * scale.  Mask semantics as in ld_gemm_desc.softmax. */
int ld_softmax_fwd(const float* S, int64_t lds, int64_t s_sb, void* P_bf16, int64_t ldp, int64_t p_sb,
                   int nb1, int nb2, int rows, int cols, float scale, const uint8_t* key_mask, int mask_inf, int causal,
                   void* stream);
int ld_softmax_bwd(const void* P_bf16, int64_t ldp, int64_t p_sb, const float* dP, int64_t lddp, int64_t dp_sb,
                   void* dS_bf16, int64_t ldds, int64_t ds_sb, int nb, int rows, int cols, float scale, void* stream);

/* Row-wise softmax cross-entropy with label smoothing / ignore_index, loss and d(loss)/d(logits) in one pass
 * (CrossEntropyLoss(label_smoothing=0.1) training/med.py:917-918; F.cross_entropy training/networks_detr.py:185,344,
 * training/loss.py:105,178,189).  dlogits may alias logits. */
int ld_cross_entropy(const void* logits, int dtype, int64_t ld_, const int64_t* labels, float* loss_rows,
                     void* dlogits, int g_dtype, int64_t ldg, int64_t rows, int V, float label_smoothing,
                     int64_t ignore_index, float grad_scale, const float* grad_scale_dev /* optional device scalar factor */,
                     void* stream);

/* Convolution support around the GEMM (channels-last bf16): patch gather / its gather-form adjoint, the ResNet stem
 * max-pool (torchvision resnet50 via training/detr_backbone.py:105) and NCHW <-> NHWC conversion. */
int ld_im2col_nhwc(const void* x_bf16, void* cols_bf16, int B, int H, int W, int C, int Ho, int Wo,
                   int KH, int KW, int stride, int pad, int Kp, void* stream);
int ld_col2im_nhwc(const void* cols_bf16, void* dx_bf16, int B, int H, int W, int C, int Ho, int Wo,
                   int KH, int KW, int stride, int pad, int Kp, void* stream);
int ld_maxpool3s2_fwd(const void* x_bf16, void* y_bf16, uint8_t* argmax, int B, int H, int W, int C, void* stream);
int ld_maxpool3s2_bwd(const void* dy_bf16, const uint8_t* argmax, void* dx_bf16, int B, int H, int W, int C, void* stream);
int ld_layout_convert(const void* src, int src_dtype, void* dst, int dst_dtype, int B, int C, int64_t HW, int direction, void* stream);

/* StyleGAN2 modulated-conv pieces on channels-last activations (training/networks_stylegan2.py:30-75, 307-325):
 * per-sample channel scaling, demodulation + bias + leaky-ReLU (fwd / bwd) and the style-gradient reduction. */
int ld_scale_channels(const void* x, int x_dtype, const float* s, void* y, int y_dtype, int64_t n, int64_t per_sample, int C, void* stream);
int ld_demod_bias_act_fwd(const void* x, int x_dtype, const float* d, const float* bias, void* y_bf16,
                          int B, int64_t pixels, int C, int act, float gain, void* stream);
int ld_demod_bias_act_bwd(const void* dy_bf16, const void* y_bf16, const void* x, int x_dtype, const float* d,
                          void* dx_bf16, float* dd, float* dbias, int B, int64_t pixels, int C, int act, float gain, void* stream);
int ld_channel_dot(const void* a, int a_dtype, const void* g_bf16, float* out, int B, int64_t pixels, int C, void* stream);
/* Deterministic forms of the two style-gradient reductions (bf16 rows, C % 8 == 0): fixed-order sums through a caller workspace of
 * ld_style_reduce_ws_floats(B, pixels, C) floats (twice that for ld_demod_bias_act_bwd_ws), no floating-point atomics — results are
 * bit-identical from run to run.  dd / dbias / out are accumulated into (+=), as in the atomics forms. */
int64_t ld_style_reduce_ws_floats(int B, int64_t pixels, int C);
int ld_demod_bias_act_bwd_ws(const void* dy_bf16, const void* y_bf16, const void* x_bf16, const float* d, void* dx_bf16, float* dd, float* dbias,
                             float* ws, int64_t ws_floats, int B, int64_t pixels, int C, int act, float gain, void* stream);
int ld_channel_dot_ws(const void* a_bf16, const void* g_bf16, float* out, float* ws, int64_t ws_floats, int B, int64_t pixels, int C, void* stream);

/* Small elementwise helpers used between GEMMs. */
int ld_cast_pad(const void* src, int src_dtype, int64_t lds, void* dst, int dst_dtype, int64_t ldd,
                int64_t rows, int cols_src, int cols_dst, void* stream);
int ld_axpby_bcast(const void* a, int a_dtype, const void* b, int b_dtype, void* out, int out_dtype,
                   int64_t n, int64_t period, float alpha, float beta, void* stream);
int ld_act_fwd_bf16(const void* x, void* y, int64_t n, int act, float gain, void* stream);
int ld_act_bwd(const void* dy, int dy_dtype, const void* ref, int ref_dtype, void* dx, int dx_dtype,
               int64_t n, int act, float gain, void* stream);
int ld_act_bwd_colscale(const void* dy, const void* ref, void* dx, void* dxs, const float* cs, int64_t n, int cols, int act, void* stream);
int ld_colsum_accum(const void* x, int dtype, int64_t ld_, float* out, int64_t rows, int cols, void* stream);

/* Optimizer step over flat storage (training/training_loop.py:303-328): nan_to_num + Adam + bf16 shadow refresh in one
 * pass, and the generator EMA.  n % 4 == 0, 16-byte aligned pointers; m may be NULL when beta1 == 0. */
int ld_adam_flat(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, float lr, float beta1, float beta2,
                 float eps, int step, float grad_scale,
                 const float* hyper_dev /* optional device {lr, 1-beta1^t, 1-beta2^t} overriding lr/step (CUDA-graph replay) */,
                 void* stream);
int ld_ema_flat(float* p_ema, const float* p, void* p_ema_bf16, int64_t n, float beta, void* stream);

/* Fused multi-head attention forward (QK^T -> scale + mask -> softmax -> dropout -> PV in one persistent kernel, scores /
 * probabilities stay on chip) for <= 256 keys and head_dim <= 192: BertSelfAttention.forward (training/med.py:146-228, dropout
 * :213) and nn.MultiheadAttention as called by training/detr_transformer.py:208,273,277.  q / k / v: bf16 row-major [B*L, ld]
 * buffers, head h in columns h*d .. h*d+d of the given base; o: bf16 [B*Lq, ldo].  key_mask [B, Lk] (1 = masked) adds -10000
 * (mask_inf = 0) or -inf.  For the backward pass: lse_out (optional) fp32 [B*H, Lq] = log2-domain log-sum-exp of the scaled +
 * masked scores (max2 + log2(sum 2^(s2 - max2)), s2 = s * log2 e), consumed by ld_attention_bwd.  dropout_p > 0: probabilities
 * are dropped with the Philox mask of (rng_state = device {seed_lo, seed_hi, step, 0}, rng_site), element (b, h, row, col) ->
 * group ((b*H+h)*Lq+row)*32 + col/8. */
int ld_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                     void* o, int64_t ldo, float* lse_out,
                     int B, int H, int Lq, int Lk, int d, float scale, const uint8_t* key_mask, int mask_inf, int causal,
                     float dropout_p, const uint32_t* rng_state, uint32_t rng_site, void* stream);
/* Debug aid for ld_attention_fwd: per-tile event clocks of CTA 0 are written to dev_buf (int64 [64][16]); NULL switches it off. */
int ld_debug_attention_trace(void* dev_buf);
/* Fused attention backward, the adjoint of ld_attention_fwd (same shape limits, same mask / dropout arguments): recomputes the
 * probabilities from lse (the forward's lse_out), regenerates the dropout mask and produces
 *   dq      bf16 [B*Lq, lddq] at head h columns h*d..   dQ = dS K,  dS = P o (dropout'(dO V^T) - rowsum(dO o O)) * scale
 *   pd_out  bf16 [B*H, Lq, ldp]  dropout(P)   (left operand of dV = Pd^T dO)
 *   ds_out  bf16 [B*H, Lq, ldp]  dS           (left operand of dK = dS^T Q)
 * in one kernel (scores, dO V^T and dQ accumulate in TMEM).  o / d_o: forward output and its gradient, bf16 [B*Lq, ldo].  The two
 * transposed products stay batched ld_gemm_bf16 calls.  Replaces autograd of training/med.py:183-215 and of
 * F.multi_head_attention_forward (training/detr_transformer.py:208,273,277). */
int ld_attention_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                     const void* o, const void* d_o, int64_t ldo, const float* lse,
                     void* dq, int64_t lddq, void* pd_out, void* ds_out, int64_t ldp,
                     int B, int H, int Lq, int Lk, int d, float scale, const uint8_t* key_mask, int mask_inf, int causal,
                     float dropout_p, const uint32_t* rng_state, uint32_t rng_site, void* stream);
/* Round-1 kernel (one tile per CTA, no overlap between phases); ld_attention_fwd dispatches here when LD_ATTN_V1=1 (A/B only). */
int ld_attention_fwd_v1(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                     void* o, int64_t ldo, void* p_out, int64_t ldp, int B, int H, int Lq, int Lk, int d,
                     float scale, const uint8_t* key_mask, int mask_inf, int causal, void* stream);

/* ------------------------------------------------------------------------------------------
 * Layout box losses of the generator objective, value + analytic gradient, one launch each.  Replace the eager
 * elementwise chains of metrics/metric_layoutnet.py: compute_overlap :153-179, compute_alignment :182-201,
 * generalized_iou_loss :245-275 (called from training/loss.py:117-125) and their autograd backward.
 *
 * ld_layout_losses: bbox [B, N, 4] fp32 (xc, yc, w, h), valid [B, N] bytes (1 = real element, i.e. ~padding_mask),
 *   1 <= N <= 64.  overlap[b] = sum_{i != j} area(i n j) / area(i) / n_valid (invalid boxes zeroed first);
 *   alignment[b] = sum_i -log(1 - min_{c, j != i} |X_c,i - X_c,j|) / n_valid over the six edge / centre coordinates
 *   (a minimum of exactly 1 counts as 0).  j_overlap / j_alignment (optional, [B, N, 4]) receive
 *   d overlap[b] / d bbox[b] and d alignment[b] / d bbox[b] with autograd's conventions (ties of max / min split
 *   evenly, first minimum takes the gradient, nan_to_num / masked_fill stop it).
 * ld_giou_loss: fake, real [M, 4]; loss[0] = mean_m (1 - GIoU(fake_m, real_m)); j_fake (optional) = d loss / d fake.
 * ld_rows_scale: out[i] (+)= J[i] * g[i / per] — the chain rule through a stored Jacobian (g: [B] or one scalar). */
int ld_layout_losses(const float* bbox, const uint8_t* valid, int64_t B, int N, float* overlap, float* alignment,
                     float* j_overlap, float* j_alignment, void* stream);
int ld_giou_loss(const float* fake, const float* real, int64_t M, float* loss, float* j_fake, void* stream);
int ld_rows_scale(const float* J, const float* g, float* out, int64_t n, int64_t per, int accumulate, void* stream);
/* Evaluation sweep (metrics/overlap50k_alignment50k_layoutwise_iou50k_layoutwise_docsim50k.py:36-45, replacing the per-layout
 * NumPy loop over metrics/metric_layoutnet.py compute_iou_for_layout :94-97 and compute_docsim_for_layout :224-226):
 * real, fake [B, N, 4] fp32, valid [B, N] bytes -> iou[b], docsim[b] = mean over the valid slots of IoU(real_i, fake_i)
 * (NaN -> 0) and of sqrt(min(area)) * 2^(-|d centre| - 2 |d shape|). */
int ld_layout_pair_metrics(const float* real, const float* fake, const uint8_t* valid, int64_t B, int N, float* iou,
                           float* docsim, void* stream);

/* Data-loader tail (training/dataset_layoutganpp.py:333-336): uint8 HWC images [B, H, W, 3] -> fp32 NCHW [B, 3, H, W],
 * (x / 255 - mean[c]) / std[c] with IEEE fp32 division in the reference's order (bit-identical to its NumPy expression).
 * mean3 / std3 are HOST pointers to 3 floats; H*W % 4 == 0. */
int ld_normalize_u8_image(const uint8_t* src_hwc, float* dst_nchw, int64_t B, int64_t H, int64_t W, const float* mean3,
                          const float* std3, void* stream);

/* Batched Hungarian matching — scipy.optimize.linear_sum_assignment(cost, maximize) as called by
 * metrics/metric_layoutnet.py:111,125,240 (compute_maximum_iou*, compute_maximum_docsim_for_layout).  fp64 cost
 * [problems, nr, nc] with 1 <= nr, nc <= 16; rows_out / cols_out [problems, min(nr, nc)] in scipy's order; status
 * 0 ok, -1 infeasible, -2 NaN / -inf entry.  Assignment is bit-identical to scipy (same scan order and tie-breaking). */
int ld_lsap(const double* cost, int nr, int nc, int maximize, int64_t problems, int64_t* rows_out, int64_t* cols_out,
            int* status, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LAYOUTDETR_SM100_H */
