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
    if (vec8) {
        const long tv = total / 8;
        const int gv = (int)std::max<long>(1, std::min<long>((tv + 255) / 256, (long)ld::sm_count() * 16));
        upfirdn2d_nhwc_bf16x8_kernel<<<gv, 256, 0, (cudaStream_t)stream>>>(p);
    } else if (dtype == LD_F32) upfirdn2d_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    else upfirdn2d_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    ld::count_launch();
    LD_LAUNCH_CHECK("upfirdn2d");
    return 0;
}
