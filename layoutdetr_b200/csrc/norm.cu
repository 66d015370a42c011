// LayerNorm forward / backward and the fused BERT embedding (+LayerNorm) kernels.
// One warp per row, row cached in registers (C % 128 == 0, C <= 1024), warp-shuffle reductions,
// 128-bit (fp32) / 64-bit (bf16) vector loads. HBM-bound: each element is read once, written once.
#include "common.cuh"
#include "runtime.h"
#include <algorithm>

namespace {
using namespace ld;

constexpr int LN_MAX_CHUNKS = 8;      // 8 * 128 = 1024 columns
constexpr int LN_WARPS = 4;

template <typename T> struct Vec4;
template <> struct Vec4<float> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
        const float4 a = *reinterpret_cast<const float4*>(p);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct Vec4<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[4]) {
        const uint2 a = *reinterpret_cast<const uint2*>(p);
        unpack_bf16x2(a.x, v[0], v[1]); unpack_bf16x2(a.y, v[2], v[3]);
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[4]) {
        uint2 a; a.x = pack_bf16x2(v[0], v[1]); a.y = pack_bf16x2(v[2], v[3]);
        *reinterpret_cast<uint2*>(p) = a;
    }
};

// y = (x - mean) * rstd * gamma + beta        (reference: nn.LayerNorm in training/med.py:63,233,318,
// training/detr_transformer.py:191-192,252-254; eps 1e-12 for BERT, 1e-5 for DETR)
template <typename TIn>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_fwd_kernel(const TIn* __restrict__ x, long ldx, const __nv_bfloat16* __restrict__ res, long ldr,
                     const float* __restrict__ gamma, const float* __restrict__ beta,
                     __nv_bfloat16* __restrict__ y16, float* __restrict__ y32, long ldy,
                     float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows, int C, float eps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * LN_WARPS + warp;
    if (row >= rows) return;
    const int chunks = C >> 7;
    float v[LN_MAX_CHUNKS][4];
    const TIn* xr = x + row * ldx;
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < LN_MAX_CHUNKS; ++j) {
        if (j < chunks) {
            Vec4<TIn>::load(xr + (j * 32 + lane) * 4, v[j]);
            if (res) {                                 // x + residual summed in fp32 (the post-norm residual block tail)
                float rr[4];
                Vec4<__nv_bfloat16>::load(res + row * ldr + (j * 32 + lane) * 4, rr);
#pragma unroll
                for (int i = 0; i < 4; ++i) v[j][i] += rr[i];
            }
            s += v[j][0] + v[j][1] + v[j][2] + v[j][3];
        }
    }
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < LN_MAX_CHUNKS; ++j) {
        if (j < chunks) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { const float d = v[j][i] - mean; q += d * d; }
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
    if (lane == 0) {
        if (mean_out) mean_out[row] = mean;
        if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int j = 0; j < LN_MAX_CHUNKS; ++j) {
        if (j < chunks) {
            const int c = (j * 32 + lane) * 4;
            float g[4], b[4], o[4];
            Vec4<float>::load(gamma + c, g);
            Vec4<float>::load(beta + c, b);
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = (v[j][i] - mean) * rstd * g[i] + b[i];
            if (y16) Vec4<__nv_bfloat16>::store(y16 + row * ldy + c, o);
            if (y32) Vec4<float>::store(y32 + row * ldy + c, o);
        }
    }
}

// bf16 in (+ bf16 residual) -> bf16 out with 128-bit accesses: C % 256 == 0, one warp per row, eight columns per lane and
// chunk.  (The generic kernel moves 8 bytes per lane and load; at [36864, 768] it reached 4.1 TB/s.)
constexpr int LN8_MAX_CHUNKS = 4;     // 4 * 256 = 1024 columns
// DROP: y = LayerNorm(dropout(x) + res) — hidden-state dropout of the post-norm residual blocks (BertSelfOutput / BertOutput
// training/med.py:237-242,318-325; DETR dropout1/2/3 training/detr_transformer.py:210-214,270-285) applied where the dense
// output is read anyway.  A lane's eight columns are one Philox group (index row * C/8 + column/8, the element order ld_dropout
// uses on the contiguous [rows, C] gradient in the backward pass).  pre32 (optional) receives dropout(x) + res for LayerNorm's backward.
struct LnDrop { const uint32_t* rng; uint32_t site, thresh16; float scale; float* pre32; long ldpre; };
template <bool DROP>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_fwd_vec8_kernel(const __nv_bfloat16* __restrict__ x, long ldx, const __nv_bfloat16* __restrict__ res, long ldr,
                          const float* __restrict__ gamma, const float* __restrict__ beta, __nv_bfloat16* __restrict__ y, long ldy,
                          float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows, int C, float eps, const LnDrop dp) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * LN_WARPS + warp;
    if (row >= rows) return;
    const int chunks = C >> 8;
    DropoutRng rng;
    rng.init(dp.rng, dp.site, DROP ? dp.thresh16 : 0u, dp.scale);
    float v[LN8_MAX_CHUNKS][8];
    uint4 xa[LN8_MAX_CHUNKS], ra[LN8_MAX_CHUNKS];
#pragma unroll
    for (int j = 0; j < LN8_MAX_CHUNKS; ++j) {
        if (j < chunks) {
            xa[j] = *reinterpret_cast<const uint4*>(x + row * ldx + (j * 32 + lane) * 8);
            if (res) ra[j] = *reinterpret_cast<const uint4*>(res + row * ldr + (j * 32 + lane) * 8);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < LN8_MAX_CHUNKS; ++j) {
        if (j < chunks) {
            const uint32_t w[4] = {xa[j].x, xa[j].y, xa[j].z, xa[j].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) unpack_bf16x2(w[k], v[j][2 * k], v[j][2 * k + 1]);
            if (DROP) {
                const uint32_t keep = rng.keep8((uint64_t)row * (uint64_t)(C >> 3) + (uint64_t)(j * 32 + lane));
#pragma unroll
                for (int k = 0; k < 8; ++k) v[j][k] = ((keep >> k) & 1u) ? v[j][k] * rng.scale : 0.0f;
            }
            if (res) {
                const uint32_t rw[4] = {ra[j].x, ra[j].y, ra[j].z, ra[j].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) { float lo, hi; unpack_bf16x2(rw[k], lo, hi); v[j][2 * k] += lo; v[j][2 * k + 1] += hi; }
            }
            if (dp.pre32) {
                float* pr = dp.pre32 + row * dp.ldpre + (j * 32 + lane) * 8;
                *reinterpret_cast<float4*>(pr) = make_float4(v[j][0], v[j][1], v[j][2], v[j][3]);
                *reinterpret_cast<float4*>(pr + 4) = make_float4(v[j][4], v[j][5], v[j][6], v[j][7]);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) s += v[j][k];
        }
    }
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < LN8_MAX_CHUNKS; ++j) {
        if (j < chunks) {
#pragma unroll
            for (int k = 0; k < 8; ++k) { const float d = v[j][k] - mean; q += d * d; }
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
    if (lane == 0) {
        if (mean_out) mean_out[row] = mean;
        if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int j = 0; j < LN8_MAX_CHUNKS; ++j) {
        if (j < chunks) {
            const int c = (j * 32 + lane) * 8;
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c) + 1);
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c) + 1);
            const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            uint4 o;
            float t[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) t[k] = (v[j][k] - mean) * rstd * g[k] + b[k];
            o.x = pack_bf16x2(t[0], t[1]); o.y = pack_bf16x2(t[2], t[3]); o.z = pack_bf16x2(t[4], t[5]); o.w = pack_bf16x2(t[6], t[7]);
            *reinterpret_cast<uint4*>(y + row * ldy + c) = o;
        }
    }
}

// dx = rstd * (dy*g - mean_c(dy*g) - xhat * mean_c(dy*g*xhat));  dgamma += sum_r dy*xhat;  dbeta += sum_r dy
constexpr int LNB_ROWS_PER_WARP = 8;
template <typename TIn, typename TDy>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_bwd_kernel(const TDy* __restrict__ dy, long lddy, const TIn* __restrict__ x, long ldx,
                     const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                     __nv_bfloat16* __restrict__ dx16, float* __restrict__ dx32, long lddx,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, int rows, int C) {
    __shared__ float red[LN_WARPS][2][1024 + 8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunks = C >> 7;
    float ag[LN_MAX_CHUNKS][4], ab[LN_MAX_CHUNKS][4];
#pragma unroll
    for (int j = 0; j < LN_MAX_CHUNKS; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) { ag[j][i] = 0.f; ab[j][i] = 0.f; }

    const long row0 = ((long)blockIdx.x * LN_WARPS + warp) * LNB_ROWS_PER_WARP;
    for (int rr = 0; rr < LNB_ROWS_PER_WARP; ++rr) {
        const long row = row0 + rr;
        if (row >= rows) break;
        const float mu = mean[row], rs = rstd[row];
        float xh[LN_MAX_CHUNKS][4], dg[LN_MAX_CHUNKS][4];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < LN_MAX_CHUNKS; ++j) {
            if (j < chunks) {
                const int c = (j * 32 + lane) * 4;
                float xv[4], dv[4], g[4];
                Vec4<TIn>::load(x + row * ldx + c, xv);
                Vec4<TDy>::load(dy + row * lddy + c, dv);
                Vec4<float>::load(gamma + c, g);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    xh[j][i] = (xv[i] - mu) * rs;
                    dg[j][i] = dv[i] * g[i];
                    s1 += dg[j][i];
                    s2 += dg[j][i] * xh[j][i];
                    ag[j][i] += dv[i] * xh[j][i];
                    ab[j][i] += dv[i];
                }
            }
        }
        s1 = warp_sum(s1) / (float)C;
        s2 = warp_sum(s2) / (float)C;
#pragma unroll
        for (int j = 0; j < LN_MAX_CHUNKS; ++j) {
            if (j < chunks) {
                const int c = (j * 32 + lane) * 4;
                float o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) o[i] = rs * (dg[j][i] - s1 - xh[j][i] * s2);
                if (dx16) Vec4<__nv_bfloat16>::store(dx16 + row * lddx + c, o);
                if (dx32) Vec4<float>::store(dx32 + row * lddx + c, o);
            }
        }
    }
    if (dgamma == nullptr && dbeta == nullptr) return;
#pragma unroll
    for (int j = 0; j < LN_MAX_CHUNKS; ++j) {
        if (j < chunks) {
            const int c = (j * 32 + lane) * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) { red[warp][0][c + i] = ag[j][i]; red[warp][1][c + i] = ab[j][i]; }
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float g = 0.f, b = 0.f;
#pragma unroll
        for (int w = 0; w < LN_WARPS; ++w) { g += red[w][0][c]; b += red[w][1][c]; }
        if (dgamma) atomicAdd(dgamma + c, g);
        if (dbeta) atomicAdd(dbeta + c, b);
    }
}

// BERT embeddings: LN(word[ids[r]] + pos[r % T]) (reference training/med.py:74-97). Tables fp32.
__global__ void __launch_bounds__(LN_WARPS * 32)
embed_ln_fwd_kernel(const int64_t* __restrict__ ids, const float* __restrict__ word, const float* __restrict__ pos,
                    const float* __restrict__ gamma, const float* __restrict__ beta,
                    __nv_bfloat16* __restrict__ y16, float* __restrict__ pre32,
                    float* __restrict__ mean_out, float* __restrict__ rstd_out,
                    int rows, int T, int C, float eps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * LN_WARPS + warp;
    if (row >= rows) return;
    const int chunks = C >> 7;
    const long id = ids[row];
    const float* wr = word + id * C;
    const float* pr = pos + (long)(row % T) * C;
    float v[LN_MAX_CHUNKS][4];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < LN_MAX_CHUNKS; ++j) {
        if (j < chunks) {
            const int c = (j * 32 + lane) * 4;
            float a[4], b[4];
            Vec4<float>::load(wr + c, a);
            Vec4<float>::load(pr + c, b);
#pragma unroll
            for (int i = 0; i < 4; ++i) { v[j][i] = a[i] + b[i]; s += v[j][i]; }
            if (pre32) Vec4<float>::store(pre32 + row * C + c, v[j]);
        }
    }
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < LN_MAX_CHUNKS; ++j)
        if (j < chunks)
#pragma unroll
            for (int i = 0; i < 4; ++i) { const float d = v[j][i] - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
    if (lane == 0) {
        if (mean_out) mean_out[row] = mean;
        if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int j = 0; j < LN_MAX_CHUNKS; ++j) {
        if (j < chunks) {
            const int c = (j * 32 + lane) * 4;
            float g[4], b[4], o[4];
            Vec4<float>::load(gamma + c, g);
            Vec4<float>::load(beta + c, b);
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = (v[j][i] - mean) * rstd * g[i] + b[i];
            Vec4<__nv_bfloat16>::store(y16 + row * C + c, o);
        }
    }
}

// scatter-add of embedding gradients: dword[ids[r]] += dpre[r], dpos[r % T] += dpre[r]
__global__ void embed_bwd_kernel(const int64_t* __restrict__ ids, const float* __restrict__ dpre,
                                 float* __restrict__ dword, float* __restrict__ dpos, long rows, int T, int C, long pad_id) {
    const long total = rows * (long)C;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long r = i / C; const int c = (int)(i - r * C);
        const float g = dpre[i];
        const long id = ids[r];
        if (dword && id != pad_id) atomicAdd(dword + id * C + c, g);   // nn.Embedding(padding_idx) gets no gradient
        if (dpos) atomicAdd(dpos + (r % T) * C + c, g);
    }
}

int check_ln(int rows, int C) {
    if (rows <= 0) { set_last_error("layernorm: rows must be > 0"); return LD_ERR_INVALID_ARG; }
    if (C % 128 != 0 || C > 128 * LN_MAX_CHUNKS || C <= 0) { set_last_error("layernorm: C=%d must be a multiple of 128 and <= 1024", C); return LD_ERR_UNSUPPORTED; }
    return 0;
}
}  // namespace

extern "C" {

int ld_layernorm_res_fwd(const void* x, int x_dtype, int64_t ldx, const void* res_bf16, int64_t ldr,
                         const float* gamma, const float* beta,
                         void* y_bf16, float* y_f32, int64_t ldy, float* mean, float* rstd,
                         int rows, int C, float eps, void* stream) {
    int e = check_ln(rows, C); if (e) return e;
    LD_CHECK_ARG(x && gamma && beta && (y_bf16 || y_f32), "layernorm_fwd: null pointer");
    LD_CHECK_ARG(ldx % 4 == 0 && ldy % 4 == 0 && (!res_bf16 || ldr % 4 == 0), "layernorm_fwd: ld must be a multiple of 4");
    const int grid = ld::ceil_div(rows, LN_WARPS);
    cudaStream_t st = (cudaStream_t)stream;
    const __nv_bfloat16* res = (const __nv_bfloat16*)res_bf16;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (x_dtype == LD_BF16 && y_bf16 && !y_f32 && C % 256 == 0 && ldx % 8 == 0 && ldy % 8 == 0 && (!res || ldr % 8 == 0) &&
        al16(x) && al16(y_bf16) && (!res || al16(res)) && al16(gamma) && al16(beta)) {
        layernorm_fwd_vec8_kernel<false><<<grid, LN_WARPS * 32, 0, st>>>((const __nv_bfloat16*)x, ldx, res, ldr, gamma, beta,
                                                                          (__nv_bfloat16*)y_bf16, ldy, mean, rstd, rows, C, eps, LnDrop{});
        ld::count_launch();
        LD_LAUNCH_CHECK("layernorm_fwd");
        return 0;
    }
    if (x_dtype == LD_F32)
        layernorm_fwd_kernel<float><<<grid, LN_WARPS * 32, 0, st>>>((const float*)x, ldx, res, ldr, gamma, beta, (__nv_bfloat16*)y_bf16, y_f32, ldy, mean, rstd, rows, C, eps);
    else
        layernorm_fwd_kernel<__nv_bfloat16><<<grid, LN_WARPS * 32, 0, st>>>((const __nv_bfloat16*)x, ldx, res, ldr, gamma, beta, (__nv_bfloat16*)y_bf16, y_f32, ldy, mean, rstd, rows, C, eps);
    ld::count_launch();
    LD_LAUNCH_CHECK("layernorm_fwd");
    return 0;
}

int ld_layernorm_res_dropout_fwd(const void* x_bf16, int64_t ldx, const void* res_bf16, int64_t ldr,
                                 const float* gamma, const float* beta, void* y_bf16, int64_t ldy, float* pre_f32, int64_t ldpre,
                                 float* mean, float* rstd, int rows, int C, float eps,
                                 float dropout_p, const uint32_t* rng_state, uint32_t rng_site, void* stream) {
    int e = check_ln(rows, C); if (e) return e;
    LD_CHECK_ARG(x_bf16 && gamma && beta && y_bf16, "layernorm_res_dropout_fwd: null pointer");
    LD_CHECK_ARG(dropout_p >= 0.0f && dropout_p < 1.0f && (dropout_p == 0.0f || rng_state), "layernorm_res_dropout_fwd: dropout arguments");
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    LD_CHECK_ARG(C % 256 == 0 && ldx % 8 == 0 && ldy % 8 == 0 && (!res_bf16 || ldr % 8 == 0) && al16(x_bf16) && al16(y_bf16) &&
                 (!res_bf16 || al16(res_bf16)) && al16(gamma) && al16(beta) && (!pre_f32 || (al16(pre_f32) && ldpre % 4 == 0)),
                 "layernorm_res_dropout_fwd: needs C %% 256 == 0 and 16-byte aligned rows");
    LnDrop dp{};
    dp.rng = rng_state; dp.site = rng_site;
    dp.thresh16 = dropout_p > 0.0f ? (uint32_t)(dropout_p * 65536.0f + 0.5f) : 0u;
    dp.scale = dropout_p > 0.0f ? 65536.0f / (65536.0f - (float)dp.thresh16) : 1.0f;
    dp.pre32 = pre_f32; dp.ldpre = ldpre;
    const int grid = ld::ceil_div(rows, LN_WARPS);
    cudaStream_t st = (cudaStream_t)stream;
    if (dp.thresh16)
        layernorm_fwd_vec8_kernel<true><<<grid, LN_WARPS * 32, 0, st>>>((const __nv_bfloat16*)x_bf16, ldx, (const __nv_bfloat16*)res_bf16, ldr,
                                                                         gamma, beta, (__nv_bfloat16*)y_bf16, ldy, mean, rstd, rows, C, eps, dp);
    else
        layernorm_fwd_vec8_kernel<false><<<grid, LN_WARPS * 32, 0, st>>>((const __nv_bfloat16*)x_bf16, ldx, (const __nv_bfloat16*)res_bf16, ldr,
                                                                          gamma, beta, (__nv_bfloat16*)y_bf16, ldy, mean, rstd, rows, C, eps, dp);
    ld::count_launch();
    LD_LAUNCH_CHECK("layernorm_res_dropout_fwd");
    return 0;
}

int ld_layernorm_fwd(const void* x, int x_dtype, int64_t ldx, const float* gamma, const float* beta,
                     void* y_bf16, float* y_f32, int64_t ldy, float* mean, float* rstd,
                     int rows, int C, float eps, void* stream) {
    return ld_layernorm_res_fwd(x, x_dtype, ldx, nullptr, 0, gamma, beta, y_bf16, y_f32, ldy, mean, rstd, rows, C, eps, stream);
}

int ld_layernorm_bwd(const void* dy, int dy_dtype, int64_t lddy, const void* x, int x_dtype, int64_t ldx,
                     const float* mean, const float* rstd, const float* gamma,
                     void* dx_bf16, float* dx_f32, int64_t lddx, float* dgamma, float* dbeta,
                     int rows, int C, void* stream) {
    int e = check_ln(rows, C); if (e) return e;
    LD_CHECK_ARG(dy && x && mean && rstd && gamma && (dx_bf16 || dx_f32), "layernorm_bwd: null pointer");
    const int grid = ld::ceil_div(rows, LN_WARPS * LNB_ROWS_PER_WARP);
    cudaStream_t st = (cudaStream_t)stream;
#define LNB(TI, TD) layernorm_bwd_kernel<TI, TD><<<grid, LN_WARPS * 32, 0, st>>>((const TD*)dy, lddy, (const TI*)x, ldx, mean, rstd, gamma, (__nv_bfloat16*)dx_bf16, dx_f32, lddx, dgamma, dbeta, rows, C)
    if (x_dtype == LD_F32 && dy_dtype == LD_F32) LNB(float, float);
    else if (x_dtype == LD_F32) LNB(float, __nv_bfloat16);
    else if (dy_dtype == LD_F32) LNB(__nv_bfloat16, float);
    else LNB(__nv_bfloat16, __nv_bfloat16);
#undef LNB
    ld::count_launch();
    LD_LAUNCH_CHECK("layernorm_bwd");
    return 0;
}

int ld_embed_ln_fwd(const int64_t* ids, const float* word, const float* pos, const float* gamma, const float* beta,
                    void* y_bf16, float* pre_f32, float* mean, float* rstd, int rows, int T, int C, float eps, void* stream) {
    int e = check_ln(rows, C); if (e) return e;
    LD_CHECK_ARG(ids && word && pos && gamma && beta && y_bf16 && T > 0, "embed_ln_fwd: bad argument");
    embed_ln_fwd_kernel<<<ld::ceil_div(rows, LN_WARPS), LN_WARPS * 32, 0, (cudaStream_t)stream>>>(
        ids, word, pos, gamma, beta, (__nv_bfloat16*)y_bf16, pre_f32, mean, rstd, rows, T, C, eps);
    ld::count_launch();
    LD_LAUNCH_CHECK("embed_ln_fwd");
    return 0;
}

int ld_embed_bwd(const int64_t* ids, const float* dpre, float* dword, float* dpos, int64_t rows, int T, int C,
                 int64_t pad_id, void* stream) {
    LD_CHECK_ARG(ids && dpre && rows > 0 && T > 0 && C > 0, "embed_bwd: bad argument");
    const long total = rows * (long)C;
    const int grid = (int)std::min<long>((total + 255) / 256, (long)ld::sm_count() * 16);
    embed_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(ids, dpre, dword, dpos, rows, T, C, pad_id);
    ld::count_launch();
    LD_LAUNCH_CHECK("embed_bwd");
    return 0;
}

}  // extern "C"
