// upfirdn2d: zero-insert upsample -> pad/crop -> 2-D FIR -> decimate, one pass, any integer up/down,
// filters up to 32x32 held in shared memory, strided (NCHW or channels-last) fp32 / bf16 tensors.
//
//   y[n,c,oy,ox] = gain * sum_{fy,fx} F[fy,fx] * U[oy*downy + fy, ox*downx + fx]
//   U[uy,ux]     = x[n,c,(uy-pady0)/upy,(ux-padx0)/upx] when both divisions are exact and in range, else 0
//   F            = f flipped in both axes unless flip_filter (i.e. true convolution by default)
//
// Semantics: torch_utils/ops/upfirdn2d.py:168-214 (_upfirdn2d_ref); replaces upfirdn2d_plugin.upfirdn2d
// (torch_utils/ops/upfirdn2d.cpp:17-105).  Only taps that land on a real (non zero-inserted) sample are
// visited, so the up=2 RGB-skip upsample reads 4 instead of 16 values per output.
#include "common.cuh"
#include "runtime.h"
#include <algorithm>
#include <cstdlib>

namespace {
using namespace ld;

struct UpfirdnArgs {
    const void* x; void* y; const float* f;
    int N, C, inH, inW, outH, outW;
    long xsn, xsc, xsh, xsw, ysn, ysc, ysh, ysw;
    int fh, fw, upx, upy, downx, downy, padx0, pady0, flip;
    int channels_last;
    float gain;
};

template <typename T>
__global__ void __launch_bounds__(256) upfirdn2d_kernel(UpfirdnArgs p) {
    __shared__ float sf[32 * 32];
    for (int i = threadIdx.x; i < p.fh * p.fw; i += blockDim.x) {
        const int fy = i / p.fw, fx = i - fy * p.fw;
        sf[i] = p.flip ? p.f[i] : p.f[(p.fh - 1 - fy) * p.fw + (p.fw - 1 - fx)];
    }
    __syncthreads();
    const T* x = (const T*)p.x; T* y = (T*)p.y;
    const long total = (long)p.N * p.C * p.outH * p.outW;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int n, c, oy, ox;
        long t = i;
        if (p.channels_last) { c = (int)(t % p.C); t /= p.C; ox = (int)(t % p.outW); t /= p.outW; oy = (int)(t % p.outH); n = (int)(t / p.outH); }
        else { ox = (int)(t % p.outW); t /= p.outW; oy = (int)(t % p.outH); t /= p.outH; c = (int)(t % p.C); n = (int)(t / p.C); }
        // first tap row/col that hits a real sample: (o*down + f - pad0) % up == 0
        const int by = oy * p.downy - p.pady0, bx = ox * p.downx - p.padx0;
        int fy0 = ((-by) % p.upy + p.upy) % p.upy;
        int fx0 = ((-bx) % p.upx + p.upx) % p.upx;
        const T* xb = x + n * p.xsn + c * p.xsc;
        float acc = 0.f;
        for (int fy = fy0; fy < p.fh; fy += p.upy) {
            const int iy = (by + fy) / p.upy;
            if (by + fy < 0 || iy >= p.inH) continue;
            for (int fx = fx0; fx < p.fw; fx += p.upx) {
                const int ix = (bx + fx) / p.upx;
                if (bx + fx < 0 || ix >= p.inW) continue;
                acc += sf[fy * p.fw + fx] * (float)xb[iy * p.xsh + ix * p.xsw];
            }
        }
        y[n * p.ysn + c * p.ysc + oy * p.ysh + ox * p.ysw] = (T)(acc * p.gain);
    }
}

// Channels-last bf16 fast path (C % 8 == 0): one thread = one output pixel x 8 channels, 128-bit loads/stores along C.
__global__ void __launch_bounds__(256) upfirdn2d_nhwc_bf16x8_kernel(UpfirdnArgs p) {
    __shared__ float sf[32 * 32];
    for (int i = threadIdx.x; i < p.fh * p.fw; i += blockDim.x) {
        const int fy = i / p.fw, fx = i - fy * p.fw;
        sf[i] = (p.flip ? p.f[i] : p.f[(p.fh - 1 - fy) * p.fw + (p.fw - 1 - fx)]) * p.gain;
    }
    __syncthreads();
    const __nv_bfloat16* x = (const __nv_bfloat16*)p.x; __nv_bfloat16* y = (__nv_bfloat16*)p.y;
    const int C8 = p.C >> 3;
    const long total = (long)p.N * p.outH * p.outW * C8;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % C8); long t = i / C8;
        const int ox = (int)(t % p.outW); t /= p.outW;
        const int oy = (int)(t % p.outH); const int n = (int)(t / p.outH);
        const int by = oy * p.downy - p.pady0, bx = ox * p.downx - p.padx0;
        const int fy0 = ((-by) % p.upy + p.upy) % p.upy;
        const int fx0 = ((-bx) % p.upx + p.upx) % p.upx;
        const __nv_bfloat16* xb = x + n * p.xsn + c8 * 8;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int fy = fy0; fy < p.fh; fy += p.upy) {
            const int iy = (by + fy) / p.upy;
            if (by + fy < 0 || iy >= p.inH) continue;
            for (int fx = fx0; fx < p.fw; fx += p.upx) {
                const int ix = (bx + fx) / p.upx;
                if (bx + fx < 0 || ix >= p.inW) continue;
                const float w = sf[fy * p.fw + fx];
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(xb + iy * p.xsh + ix * p.xsw));
                float lo, hi;
                unpack_bf16x2(v.x, lo, hi); acc[0] = fmaf(w, lo, acc[0]); acc[1] = fmaf(w, hi, acc[1]);
                unpack_bf16x2(v.y, lo, hi); acc[2] = fmaf(w, lo, acc[2]); acc[3] = fmaf(w, hi, acc[3]);
                unpack_bf16x2(v.z, lo, hi); acc[4] = fmaf(w, lo, acc[4]); acc[5] = fmaf(w, hi, acc[5]);
                unpack_bf16x2(v.w, lo, hi); acc[6] = fmaf(w, lo, acc[6]); acc[7] = fmaf(w, hi, acc[7]);
            }
        }
        uint4 o;
        o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
        o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
        *reinterpret_cast<uint4*>(y + n * p.ysn + oy * p.ysh + ox * p.ysw + c8 * 8) = o;
    }
}
// Shared-memory tiled FIR for the channels-last bf16 path with up = down = 1 and filters up to 4 x 4 (the resampling filter after
// every stride-2 transposed convolution of the StyleGAN2 synthesis network, training/networks_stylegan2.py:307-325 ->
// torch_utils/ops/conv2d_resample.py:118-123): a block owns an 8 x 8 output tile x 32 channels.  The (8 + fh - 1) x (8 + fw - 1)
// input tile is loaded ONCE with 128-bit loads (64 contiguous bytes per pixel), converted to fp32 ONCE and staged in shared memory;
// a rank-1 filter (setup_filter's outer product: the only kind the path uses) is applied as a horizontal then a vertical pass,
// 4 + 4 taps per output instead of 16; any other filter takes the 2-D loop over the same tile.  Round 1's kernel read 16 global
// values and unpacked 16 x 8 bf16 per output (issue-bound at a quarter of the HBM peak).
constexpr int UT = 8, UT_IN = UT + 3, UT_C = 32;
__global__ void __launch_bounds__(256) upfirdn2d_nhwc_fir_tiled_kernel(UpfirdnArgs p, int tiles_x) {
    __shared__ __align__(16) float tile[UT_IN * UT_IN * UT_C];      // [py][px][32 channels]
    __shared__ __align__(16) float hbuf[UT_IN * UT * UT_C];         // [py][ox][32 channels]
    __shared__ float sf[16], wx[4], wy[4];
    __shared__ int sep;
    const int t = threadIdx.x;
    if (t < p.fh * p.fw) {
        const int fy = t / p.fw, fx = t - fy * p.fw;
        sf[t] = (p.flip ? p.f[t] : p.f[(p.fh - 1 - fy) * p.fw + (p.fw - 1 - fx)]) * p.gain;
    }
    __syncthreads();
    if (t == 0) {
        int ok = sf[0] != 0.f;
        float mx = 0.f;
        for (int i = 0; i < p.fh * p.fw; ++i) mx = fmaxf(mx, fabsf(sf[i]));
        for (int i = 0; i < p.fh; ++i) wy[i] = sf[i * p.fw];
        for (int j = 0; j < p.fw; ++j) wx[j] = ok ? sf[j] / sf[0] : 0.f;
        for (int i = 0; i < p.fh && ok; ++i)
            for (int j = 0; j < p.fw; ++j) if (fabsf(wy[i] * wx[j] - sf[i * p.fw + j]) > 1e-6f * mx) ok = 0;
        sep = ok;
    }
    const int tile_id = blockIdx.x;
    const int ty = tile_id / tiles_x, tx = tile_id - ty * tiles_x;
    const int oy0 = ty * UT, ox0 = tx * UT;
    const int n = blockIdx.y, c0 = blockIdx.z * UT_C;
    const int ih = UT + p.fh - 1, iw = UT + p.fw - 1;
    const __nv_bfloat16* x = (const __nv_bfloat16*)p.x + (long)n * p.xsn + c0;
    // ---- input tile -> fp32 shared memory (zero outside the image: the padding of the op)
    for (int it = t; it < ih * iw * 4; it += 256) {
        const int c8 = it & 3, pix = it >> 2;
        const int py = pix / iw, px = pix - py * iw;
        const int gy = oy0 - p.pady0 + py, gx = ox0 - p.padx0 + px;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (gy >= 0 && gy < p.inH && gx >= 0 && gx < p.inW) {
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(x + (long)gy * p.xsh + (long)gx * p.xsw + c8 * 8));
            unpack_bf16x2(a.x, v[0], v[1]); unpack_bf16x2(a.y, v[2], v[3]); unpack_bf16x2(a.z, v[4], v[5]); unpack_bf16x2(a.w, v[6], v[7]);
        }
        float4* dst = reinterpret_cast<float4*>(&tile[(py * UT_IN + px) * UT_C + c8 * 8]);
        dst[0] = make_float4(v[0], v[1], v[2], v[3]);
        dst[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    __syncthreads();
    const int c8 = t & 3, ox = (t >> 2) & 7, oy = t >> 5;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (sep) {
        // horizontal pass over every input row of the tile
        for (int it = t; it < ih * UT * 4; it += 256) {
            const int hc8 = it & 3, hox = (it >> 2) & 7, py = it >> 5;
            float h[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int fx = 0; fx < p.fw; ++fx) {
                const float w = wx[fx];
                const float4* src = reinterpret_cast<const float4*>(&tile[(py * UT_IN + hox + fx) * UT_C + hc8 * 8]);
                const float4 a = src[0], b = src[1];
                h[0] = fmaf(w, a.x, h[0]); h[1] = fmaf(w, a.y, h[1]); h[2] = fmaf(w, a.z, h[2]); h[3] = fmaf(w, a.w, h[3]);
                h[4] = fmaf(w, b.x, h[4]); h[5] = fmaf(w, b.y, h[5]); h[6] = fmaf(w, b.z, h[6]); h[7] = fmaf(w, b.w, h[7]);
            }
            float4* dst = reinterpret_cast<float4*>(&hbuf[(py * UT + hox) * UT_C + hc8 * 8]);
            dst[0] = make_float4(h[0], h[1], h[2], h[3]);
            dst[1] = make_float4(h[4], h[5], h[6], h[7]);
        }
        __syncthreads();
        for (int fy = 0; fy < p.fh; ++fy) {
            const float w = wy[fy];
            const float4* src = reinterpret_cast<const float4*>(&hbuf[((oy + fy) * UT + ox) * UT_C + c8 * 8]);
            const float4 a = src[0], b = src[1];
            acc[0] = fmaf(w, a.x, acc[0]); acc[1] = fmaf(w, a.y, acc[1]); acc[2] = fmaf(w, a.z, acc[2]); acc[3] = fmaf(w, a.w, acc[3]);
            acc[4] = fmaf(w, b.x, acc[4]); acc[5] = fmaf(w, b.y, acc[5]); acc[6] = fmaf(w, b.z, acc[6]); acc[7] = fmaf(w, b.w, acc[7]);
        }
    } else {
        for (int fy = 0; fy < p.fh; ++fy)
            for (int fx = 0; fx < p.fw; ++fx) {
                const float w = sf[fy * p.fw + fx];
                const float4* src = reinterpret_cast<const float4*>(&tile[((oy + fy) * UT_IN + ox + fx) * UT_C + c8 * 8]);
                const float4 a = src[0], b = src[1];
                acc[0] = fmaf(w, a.x, acc[0]); acc[1] = fmaf(w, a.y, acc[1]); acc[2] = fmaf(w, a.z, acc[2]); acc[3] = fmaf(w, a.w, acc[3]);
                acc[4] = fmaf(w, b.x, acc[4]); acc[5] = fmaf(w, b.y, acc[5]); acc[6] = fmaf(w, b.z, acc[6]); acc[7] = fmaf(w, b.w, acc[7]);
            }
    }
    const int gy = oy0 + oy, gx = ox0 + ox;
    if (gy < p.outH && gx < p.outW) {
        uint4 o;
        o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
        o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
        __nv_bfloat16* y = (__nv_bfloat16*)p.y + (long)n * p.ysn + c0;
        *reinterpret_cast<uint4*>(y + (long)gy * p.ysh + (long)gx * p.ysw + c8 * 8) = o;
    }
}
}  // namespace

extern "C" int ld_upfirdn2d(const void* x, void* y, int dtype, const float* f, int fh, int fw,
                            int N, int C, int inH, int inW, int outH, int outW,
                            const int64_t* x_strides, const int64_t* y_strides,
                            int upx, int upy, int downx, int downy, int padx0, int padx1, int pady0, int pady1,
                            int flip_filter, float gain, void* stream) {
    LD_CHECK_ARG(x && y && f && x_strides && y_strides, "upfirdn2d: null pointer");
    LD_CHECK_ARG(N > 0 && C > 0 && inH > 0 && inW > 0, "upfirdn2d: empty input");
    LD_CHECK_ARG(fh >= 1 && fw >= 1 && fh <= 32 && fw <= 32, "upfirdn2d: filter must be 1..32 taps per axis (got %dx%d)", fh, fw);
    LD_CHECK_ARG(upx >= 1 && upy >= 1 && downx >= 1 && downy >= 1, "upfirdn2d: up/down factors must be >= 1");
    const int eH = (inH * upy + pady0 + pady1 - fh + downy) / downy;
    const int eW = (inW * upx + padx0 + padx1 - fw + downx) / downx;
    LD_CHECK_ARG(eH >= 1 && eW >= 1, "upfirdn2d: upsampled+padded input smaller than the filter");
    LD_CHECK_ARG(eH == outH && eW == outW, "upfirdn2d: output size %dx%d does not match expected %dx%d", outH, outW, eH, eW);
    LD_CHECK_ARG(dtype == LD_F32 || dtype == LD_BF16, "upfirdn2d: dtype must be f32 or bf16");
    UpfirdnArgs p;
    p.x = x; p.y = y; p.f = f; p.N = N; p.C = C; p.inH = inH; p.inW = inW; p.outH = outH; p.outW = outW;
    p.xsn = x_strides[0]; p.xsc = x_strides[1]; p.xsh = x_strides[2]; p.xsw = x_strides[3];
    p.ysn = y_strides[0]; p.ysc = y_strides[1]; p.ysh = y_strides[2]; p.ysw = y_strides[3];
    p.fh = fh; p.fw = fw; p.upx = upx; p.upy = upy; p.downx = downx; p.downy = downy; p.padx0 = padx0; p.pady0 = pady0;
    p.flip = flip_filter ? 1 : 0; p.gain = gain;
    p.channels_last = (y_strides[1] == 1 && C > 1) ? 1 : 0;
    const long total = (long)N * C * outH * outW;
    const int grid = (int)std::max<long>(1, std::min<long>((total + 255) / 256, (long)ld::sm_count() * 16));
    const bool vec8 = dtype == LD_BF16 && p.channels_last && x_strides[1] == 1 && C % 8 == 0 &&
                      (((uintptr_t)x | (uintptr_t)y) & 15) == 0 &&
                      x_strides[0] % 8 == 0 && x_strides[2] % 8 == 0 && x_strides[3] % 8 == 0 &&
                      y_strides[0] % 8 == 0 && y_strides[2] % 8 == 0 && y_strides[3] % 8 == 0;
    static const int env_tiled = [] { const char* e = getenv("LD_UPFIRDN_TILED"); return e ? atoi(e) : 1; }();
    if (vec8 && env_tiled && upx == 1 && upy == 1 && downx == 1 && downy == 1 && fh <= 4 && fw <= 4 && C % UT_C == 0 && N <= 65535 && C / UT_C <= 65535) {
        const int tiles_x = (outW + UT - 1) / UT, tiles_y = (outH + UT - 1) / UT;
        upfirdn2d_nhwc_fir_tiled_kernel<<<dim3((unsigned)(tiles_x * tiles_y), (unsigned)N, (unsigned)(C / UT_C)), 256, 0, (cudaStream_t)stream>>>(p, tiles_x);
    } else if (vec8) {
        const long tv = total / 8;
        const int gv = (int)std::max<long>(1, std::min<long>((tv + 255) / 256, (long)ld::sm_count() * 16));
        upfirdn2d_nhwc_bf16x8_kernel<<<gv, 256, 0, (cudaStream_t)stream>>>(p);
    } else if (dtype == LD_F32) upfirdn2d_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    else upfirdn2d_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    ld::count_launch();
    LD_LAUNCH_CHECK("upfirdn2d");
    return 0;
}
