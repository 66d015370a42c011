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
    int32_t block_n;             /* 0 = auto, else 128 or 256 */
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

#ifdef __cplusplus
}
#endif
#endif /* LAYOUTDETR_SM100_H */
