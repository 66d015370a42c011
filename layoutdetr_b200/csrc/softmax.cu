// Masked row softmax (forward / backward) for the batched-GEMM attention path.
//   P = softmax(S * scale + mask)      S fp32 [nb1*nb2, rows, lds] -> P bf16 [nb1*nb2, rows, ldp]
// mask sources (reference semantics):
//   * key padding, additive -10000  : BERT extended attention mask, training/med.py:651-654
//   * key padding, -inf             : nn.MultiheadAttention key_padding_mask, training/detr_transformer.py:208,273,277
//   * causal (col > row), -10000    : BERT decoder causal mask, training/med.py:623-640
// One warp per row, warp-shuffle max / sum.
#include "common.cuh"
#include "runtime.h"

namespace {
using namespace ld;
constexpr int SM_WARPS = 4;

__global__ void __launch_bounds__(SM_WARPS * 32)
softmax_fwd_kernel(const float* __restrict__ S, long lds, long s_sb, __nv_bfloat16* __restrict__ P, long ldp, long p_sb,
                   int nb2, int rows, int cols, float scale, const uint8_t* __restrict__ key_mask, int mask_inf, int causal,
                   long total_rows) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long gr = (long)blockIdx.x * SM_WARPS + warp;
    if (gr >= total_rows) return;
    const long b = gr / rows; const int r = (int)(gr - b * rows);
    const long b1 = b / nb2;
    const float* s = S + b * s_sb + (long)r * lds;
    __nv_bfloat16* p = P + b * p_sb + (long)r * ldp;
    const uint8_t* km = key_mask ? key_mask + b1 * cols : nullptr;
    const float neg = mask_inf ? -INFINITY : -10000.0f;

    float mx = -INFINITY;
    for (int c = lane; c < cols; c += 32) {
        float v = s[c] * scale;
        if ((km && km[c]) || (causal && c > r)) v += neg;
        mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int c = lane; c < cols; c += 32) {
        float v = s[c] * scale;
        if ((km && km[c]) || (causal && c > r)) v += neg;
        sum += __expf(v - mx);
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int c = lane; c < cols; c += 32) {
        float v = s[c] * scale;
        if ((km && km[c]) || (causal && c > r)) v += neg;
        p[c] = f32_to_bf16(__expf(v - mx) * inv);
    }
}

// dS = P * (dP - sum_c(dP * P)) * scale
__global__ void __launch_bounds__(SM_WARPS * 32)
softmax_bwd_kernel(const __nv_bfloat16* __restrict__ P, long ldp, long p_sb, const float* __restrict__ dP, long lddp, long dp_sb,
                   __nv_bfloat16* __restrict__ dS, long ldds, long ds_sb, int rows, int cols, float scale, long total_rows) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long gr = (long)blockIdx.x * SM_WARPS + warp;
    if (gr >= total_rows) return;
    const long b = gr / rows; const int r = (int)(gr - b * rows);
    const __nv_bfloat16* p = P + b * p_sb + (long)r * ldp;
    const float* dp = dP + b * dp_sb + (long)r * lddp;
    __nv_bfloat16* ds = dS + b * ds_sb + (long)r * ldds;
    float dot = 0.f;
    for (int c = lane; c < cols; c += 32) dot += bf16_to_f32(p[c]) * dp[c];
    dot = warp_sum(dot);
    for (int c = lane; c < cols; c += 32) ds[c] = f32_to_bf16(bf16_to_f32(p[c]) * (dp[c] - dot) * scale);
}
}  // namespace

extern "C" {
int ld_softmax_fwd(const float* S, int64_t lds, int64_t s_sb, void* P_bf16, int64_t ldp, int64_t p_sb,
                   int nb1, int nb2, int rows, int cols, float scale, const uint8_t* key_mask, int mask_inf, int causal,
                   void* stream) {
    LD_CHECK_ARG(S && P_bf16 && nb1 > 0 && nb2 > 0 && rows > 0 && cols > 0, "softmax_fwd: bad argument");
    const long total = (long)nb1 * nb2 * rows;
    softmax_fwd_kernel<<<ld::ceil_div(total, SM_WARPS), SM_WARPS * 32, 0, (cudaStream_t)stream>>>(
        S, lds, s_sb, (__nv_bfloat16*)P_bf16, ldp, p_sb, nb2, rows, cols, scale, key_mask, mask_inf, causal, total);
    ld::count_launch();
    LD_LAUNCH_CHECK("softmax_fwd");
    return 0;
}

int ld_softmax_bwd(const void* P_bf16, int64_t ldp, int64_t p_sb, const float* dP, int64_t lddp, int64_t dp_sb,
                   void* dS_bf16, int64_t ldds, int64_t ds_sb, int nb, int rows, int cols, float scale, void* stream) {
    LD_CHECK_ARG(P_bf16 && dP && dS_bf16 && nb > 0 && rows > 0 && cols > 0, "softmax_bwd: bad argument");
    const long total = (long)nb * rows;
    softmax_bwd_kernel<<<ld::ceil_div(total, SM_WARPS), SM_WARPS * 32, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)P_bf16, ldp, p_sb, dP, lddp, dp_sb, (__nv_bfloat16*)dS_bf16, ldds, ds_sb, rows, cols, scale, total);
    ld::count_launch();
    LD_LAUNCH_CHECK("softmax_bwd");
    return 0;
}
}
