// Convolution support kernels around the tcgen05 GEMM (channels-last bf16 activations):
//   im2col  : NHWC -> [B*Ho*Wo, pad8(KH*KW*C)] patch matrix (K order = kh, kw, c), zero padding
//   col2im  : gather-form inverse (conv dgrad, and the forward of stride-2 transposed conv)
//   maxpool : 3x3 / stride 2 / pad 1 forward (+argmax) and backward (ResNet stem)
//   layout  : NCHW fp32 <-> NHWC bf16
// The FLOPs live in ld_gemm_bf16; these are HBM-bound gathers with 128-bit accesses along C.
#include "common.cuh"
#include "runtime.h"
#include <algorithm>

namespace {
using namespace ld;

inline int cv_grid(long n_items, int threads) {
    const long blocks = (n_items + threads - 1) / threads;
    return (int)std::max<long>(1, std::min(blocks, (long)sm_count() * 16));
}

struct ConvGeom {
    int B, H, W, C;        // input  (NHWC)
    int Ho, Wo;            // output spatial
    int KH, KW, stride, pad;
    int Kp;                // padded row length of the patch matrix
};

// vectorised: C % 8 == 0.  32-bit index arithmetic (the host checks the vector count fits): the 64-bit divisions of the
// first version made this copy kernel instruction-bound (ncu: 1.9 TB/s of stores).
__global__ void __launch_bounds__(256) im2col_vec8_kernel(const uint4* __restrict__ x, uint4* __restrict__ cols, ConvGeom g) {
    const uint32_t C8 = (uint32_t)g.C >> 3;
    const uint32_t per_row = (uint32_t)(g.KH * g.KW) * C8;
    const uint32_t total = (uint32_t)g.B * g.Ho * g.Wo * per_row;
    const uint32_t Kp8 = (uint32_t)g.Kp >> 3;
    const uint32_t KW = g.KW, Wo = g.Wo, Ho = g.Ho;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t m = i / per_row, r = i - m * per_row;
        const uint32_t tap = r / C8, c8 = r - tap * C8;
        const uint32_t kh = tap / KW, kw = tap - kh * KW;
        const uint32_t t = m / Wo, ox = m - t * Wo;
        const uint32_t b = t / Ho, oy = t - b * Ho;
        const int iy = (int)(oy * g.stride + kh) - g.pad, ix = (int)(ox * g.stride + kw) - g.pad;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W)
            v = __ldg(x + ((size_t)(b * g.H + iy) * g.W + ix) * C8 + c8);
        cols[(size_t)m * Kp8 + r] = v;
    }
}

// any C (stem, C = 3): one thread gathers 8 consecutive patch elements (one 16-byte store; Kp % 8 == 0) and zero-fills the
// K padding.  The (kh, kw, c) decode is done once per thread and then advanced incrementally — the first version decoded
// every element with 64-bit divisions and ran 20x below the HBM rate.
__global__ void im2col_scalar_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ cols, ConvGeom g) {
    const int kv = g.Kp >> 3;                                  // 8-element groups per patch row
    const long total = (long)g.B * g.Ho * g.Wo * kv;
    const int Kreal = g.KH * g.KW * g.C;
    const unsigned short* xs = reinterpret_cast<const unsigned short*>(x);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long m = i / kv;
        int k = (int)(i - m * kv) * 8;
        const int ox = (int)(m % g.Wo); const long t = m / g.Wo;
        const int oy = (int)(t % g.Ho); const int b = (int)(t / g.Ho);
        int tap = k / g.C, c = k - tap * g.C;
        int kh = tap / g.KW, kw = tap - kh * g.KW;
        const int iy0 = oy * g.stride - g.pad, ix0 = ox * g.stride - g.pad;
        const long img = (long)b * g.H;
        unsigned short v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            unsigned short e = 0;
            if (k < Kreal) {
                const int iy = iy0 + kh, ix = ix0 + kw;
                if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) e = __ldg(xs + ((img + iy) * g.W + ix) * g.C + c);
            }
            v[j] = e;
            ++k;
            if (++c == g.C) { c = 0; if (++kw == g.KW) { kw = 0; ++kh; } }
        }
        uint4 o;
        o.x = v[0] | ((uint32_t)v[1] << 16); o.y = v[2] | ((uint32_t)v[3] << 16);
        o.z = v[4] | ((uint32_t)v[5] << 16); o.w = v[6] | ((uint32_t)v[7] << 16);
        reinterpret_cast<uint4*>(cols)[i] = o;
    }
}

// dx[b, iy, ix, c] = sum_{kh,kw : (iy + pad - kh) % stride == 0, oy in range} cols[(b, oy, ox), (kh, kw, c)]
// optional per-(b, c) scale and residual add in the same pass.
template <bool VEC>
__global__ void col2im_kernel(const __nv_bfloat16* __restrict__ cols, __nv_bfloat16* __restrict__ dx, ConvGeom g) {
    const int CV = VEC ? (g.C >> 3) : g.C;
    const long total = (long)g.B * g.H * g.W * CV;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % CV); long t = i / CV;
        const int ix = (int)(t % g.W); t /= g.W;
        const int iy = (int)(t % g.H); const int b = (int)(t / g.H);
        float acc[VEC ? 8 : 1];
#pragma unroll
        for (int j = 0; j < (VEC ? 8 : 1); ++j) acc[j] = 0.f;
        for (int kh = 0; kh < g.KH; ++kh) {
            const int ny = iy + g.pad - kh;
            if (ny < 0 || ny % g.stride) continue;
            const int oy = ny / g.stride;
            if (oy >= g.Ho) continue;
            for (int kw = 0; kw < g.KW; ++kw) {
                const int nx = ix + g.pad - kw;
                if (nx < 0 || nx % g.stride) continue;
                const int ox = nx / g.stride;
                if (ox >= g.Wo) continue;
                const long m = ((long)b * g.Ho + oy) * g.Wo + ox;
                const long off = m * g.Kp + (long)(kh * g.KW + kw) * g.C;
                if (VEC) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4*>(cols + off) + cv);
                    float lo, hi;
                    unpack_bf16x2(v.x, lo, hi); acc[0] += lo; acc[1] += hi;
                    unpack_bf16x2(v.y, lo, hi); acc[2] += lo; acc[3] += hi;
                    unpack_bf16x2(v.z, lo, hi); acc[4] += lo; acc[5] += hi;
                    unpack_bf16x2(v.w, lo, hi); acc[6] += lo; acc[7] += hi;
                } else {
                    acc[0] += bf16_to_f32(cols[off + cv]);
                }
            }
        }
        if (VEC) {
            uint4 o;
            o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
            o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
            reinterpret_cast<uint4*>(dx)[i] = o;
        } else {
            dx[i] = f32_to_bf16(acc[0]);
        }
    }
}

// 3x3 stride-2 pad-1 max pooling, NHWC bf16 (torchvision resnet50 stem; reference training/detr_backbone.py:105)
__global__ void maxpool3s2_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, uint8_t* __restrict__ arg,
                                      int B, int H, int W, int C, int Ho, int Wo) {
    const long total = (long)B * Ho * Wo * C;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C); long t = i / C;
        const int ox = (int)(t % Wo); t /= Wo;
        const int oy = (int)(t % Ho); const int b = (int)(t / Ho);
        float best = -INFINITY; int bi = 0;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int iy = oy * 2 - 1 + kh;
            if (iy < 0 || iy >= H) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int ix = ox * 2 - 1 + kw;
                if (ix < 0 || ix >= W) continue;
                const float v = bf16_to_f32(x[(((long)b * H + iy) * W + ix) * C + c]);
                if (v > best) { best = v; bi = kh * 3 + kw; }
            }
        }
        y[i] = f32_to_bf16(best);
        if (arg) arg[i] = (uint8_t)bi;
    }
}

__global__ void maxpool3s2_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ arg, __nv_bfloat16* __restrict__ dx,
                                      int B, int H, int W, int C, int Ho, int Wo) {
    const long total = (long)B * H * W * C;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C); long t = i / C;
        const int ix = (int)(t % W); t /= W;
        const int iy = (int)(t % H); const int b = (int)(t / H);
        float acc = 0.f;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int ny = iy + 1 - kh;
            if (ny < 0 || (ny & 1)) continue;
            const int oy = ny >> 1;
            if (oy >= Ho) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int nx = ix + 1 - kw;
                if (nx < 0 || (nx & 1)) continue;
                const int ox = nx >> 1;
                if (ox >= Wo) continue;
                const long o = (((long)b * Ho + oy) * Wo + ox) * C + c;
                if (arg[o] == kh * 3 + kw) acc += bf16_to_f32(dy[o]);
            }
        }
        dx[i] = f32_to_bf16(acc);
    }
}

template <typename TS, typename TD>
__global__ void nchw_to_nhwc_kernel(const TS* __restrict__ src, TD* __restrict__ dst, int B, int C, long HW) {
    const long total = (long)B * C * HW;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C); long t = i / C;
        const long p = t % HW; const int b = (int)(t / HW);
        const float v = (float)src[((long)b * C + c) * HW + p];
        dst[i] = (TD)v;
    }
}
template <typename TS, typename TD>
__global__ void nhwc_to_nchw_kernel(const TS* __restrict__ src, TD* __restrict__ dst, int B, int C, long HW) {
    const long total = (long)B * C * HW;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long p = i % HW; long t = i / HW;
        const int c = (int)(t % C); const int b = (int)(t / C);
        const float v = (float)src[((long)b * HW + p) * C + c];
        dst[i] = (TD)v;
    }
}

int check_geom(const ConvGeom& g) {
    if (g.B <= 0 || g.H <= 0 || g.W <= 0 || g.C <= 0 || g.Ho <= 0 || g.Wo <= 0 || g.KH <= 0 || g.KW <= 0 || g.stride <= 0 || g.pad < 0) {
        set_last_error("conv geometry: non-positive dimension"); return LD_ERR_INVALID_ARG;
    }
    if (g.Kp < g.KH * g.KW * g.C || g.Kp % 8 != 0) { set_last_error("conv geometry: Kp=%d must be >= KH*KW*C and a multiple of 8", g.Kp); return LD_ERR_INVALID_ARG; }
    return 0;
}
}  // namespace

extern "C" {

int ld_im2col_nhwc(const void* x_bf16, void* cols_bf16, int B, int H, int W, int C, int Ho, int Wo,
                   int KH, int KW, int stride, int pad, int Kp, void* stream) {
    ConvGeom g{B, H, W, C, Ho, Wo, KH, KW, stride, pad, Kp};
    int e = check_geom(g); if (e) return e;
    LD_CHECK_ARG(x_bf16 && cols_bf16, "im2col: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (C % 8 == 0 && Kp == KH * KW * C && ((uintptr_t)x_bf16 & 15) == 0 && ((uintptr_t)cols_bf16 & 15) == 0) {
        const long total = (long)B * Ho * Wo * KH * KW * (C / 8);
        LD_CHECK_ARG(total < (1L << 31), "im2col: patch matrix too large for 32-bit vector indexing");
        im2col_vec8_kernel<<<cv_grid(total, 256), 256, 0, st>>>((const uint4*)x_bf16, (uint4*)cols_bf16, g);
    } else {
        LD_CHECK_ARG(((uintptr_t)cols_bf16 & 15) == 0, "im2col: the patch matrix must be 16-byte aligned");
        const long total = (long)B * Ho * Wo * (Kp / 8);
        im2col_scalar_kernel<<<cv_grid(total, 256), 256, 0, st>>>((const __nv_bfloat16*)x_bf16, (__nv_bfloat16*)cols_bf16, g);
    }
    ld::count_launch();
    LD_LAUNCH_CHECK("im2col");
    return 0;
}

int ld_col2im_nhwc(const void* cols_bf16, void* dx_bf16, int B, int H, int W, int C, int Ho, int Wo,
                   int KH, int KW, int stride, int pad, int Kp, void* stream) {
    ConvGeom g{B, H, W, C, Ho, Wo, KH, KW, stride, pad, Kp};
    int e = check_geom(g); if (e) return e;
    LD_CHECK_ARG(cols_bf16 && dx_bf16, "col2im: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (C % 8 == 0 && ((uintptr_t)cols_bf16 & 15) == 0 && ((uintptr_t)dx_bf16 & 15) == 0) {
        const long total = (long)B * H * W * (C / 8);
        col2im_kernel<true><<<cv_grid(total, 256), 256, 0, st>>>((const __nv_bfloat16*)cols_bf16, (__nv_bfloat16*)dx_bf16, g);
    } else {
        const long total = (long)B * H * W * C;
        col2im_kernel<false><<<cv_grid(total, 256), 256, 0, st>>>((const __nv_bfloat16*)cols_bf16, (__nv_bfloat16*)dx_bf16, g);
    }
    ld::count_launch();
    LD_LAUNCH_CHECK("col2im");
    return 0;
}

int ld_maxpool3s2_fwd(const void* x_bf16, void* y_bf16, uint8_t* argmax, int B, int H, int W, int C, void* stream) {
    LD_CHECK_ARG(x_bf16 && y_bf16 && B > 0 && H > 0 && W > 0 && C > 0, "maxpool_fwd: bad argument");
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const long total = (long)B * Ho * Wo * C;
    maxpool3s2_fwd_kernel<<<cv_grid(total, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x_bf16, (__nv_bfloat16*)y_bf16, argmax, B, H, W, C, Ho, Wo);
    ld::count_launch();
    LD_LAUNCH_CHECK("maxpool_fwd");
    return 0;
}

int ld_maxpool3s2_bwd(const void* dy_bf16, const uint8_t* argmax, void* dx_bf16, int B, int H, int W, int C, void* stream) {
    LD_CHECK_ARG(dy_bf16 && argmax && dx_bf16 && B > 0 && H > 0 && W > 0 && C > 0, "maxpool_bwd: bad argument");
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const long total = (long)B * H * W * C;
    maxpool3s2_bwd_kernel<<<cv_grid(total, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dy_bf16, argmax, (__nv_bfloat16*)dx_bf16, B, H, W, C, Ho, Wo);
    ld::count_launch();
    LD_LAUNCH_CHECK("maxpool_bwd");
    return 0;
}

// direction 0: NCHW -> NHWC, 1: NHWC -> NCHW; dtypes LD_F32 / LD_BF16 on either side
int ld_layout_convert(const void* src, int src_dtype, void* dst, int dst_dtype, int B, int C, int64_t HW, int direction, void* stream) {
    LD_CHECK_ARG(src && dst && B > 0 && C > 0 && HW > 0, "layout_convert: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const long total = (long)B * C * HW;
    const int grid = cv_grid(total, 256);
#define LC(KERN, TS, TD) KERN<TS, TD><<<grid, 256, 0, st>>>((const TS*)src, (TD*)dst, B, C, HW)
    if (direction == 0) {
        if (src_dtype == LD_F32 && dst_dtype == LD_BF16) LC(nchw_to_nhwc_kernel, float, __nv_bfloat16);
        else if (src_dtype == LD_BF16 && dst_dtype == LD_F32) LC(nchw_to_nhwc_kernel, __nv_bfloat16, float);
        else if (src_dtype == LD_F32) LC(nchw_to_nhwc_kernel, float, float);
        else LC(nchw_to_nhwc_kernel, __nv_bfloat16, __nv_bfloat16);
    } else {
        if (src_dtype == LD_F32 && dst_dtype == LD_BF16) LC(nhwc_to_nchw_kernel, float, __nv_bfloat16);
        else if (src_dtype == LD_BF16 && dst_dtype == LD_F32) LC(nhwc_to_nchw_kernel, __nv_bfloat16, float);
        else if (src_dtype == LD_F32) LC(nhwc_to_nchw_kernel, float, float);
        else LC(nhwc_to_nchw_kernel, __nv_bfloat16, __nv_bfloat16);
    }
#undef LC
    ld::count_launch();
    LD_LAUNCH_CHECK("layout_convert");
    return 0;
}

}  // extern "C"
