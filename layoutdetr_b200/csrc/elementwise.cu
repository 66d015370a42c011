// Bytes-bound elementwise kernels: bias_act (StyleGAN2 fused bias + activation + gain + clamp with
// first/second-order gradient modes), fma, dtype casts with row padding, broadcast add, activation
// backward for GEMM-fused activations, per-sample channel modulation.  All are grid-stride with
// 128-bit vector accesses where alignment allows; grids are sized in multiples of the SM count.
#include "common.cuh"
#include "runtime.h"
#include <algorithm>

namespace {
using namespace ld;

inline int ew_grid(long n_items, int threads) {
    const long blocks = (n_items + threads - 1) / threads;
    const long cap = (long)sm_count() * 8;
    return (int)std::max<long>(1, std::min(blocks, cap));
}

template <typename T> __device__ __forceinline__ float ldf(const T* p, long i);
template <> __device__ __forceinline__ float ldf<float>(const float* p, long i) { return p[i]; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p, long i) { return bf16_to_f32(p[i]); }
template <typename T> __device__ __forceinline__ void stf(T* p, long i, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, long i, float v) { p[i] = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, long i, float v) { p[i] = f32_to_bf16(v); }

// ---------------------------------------------------------------------------------------------
// bias_act.  Semantics follow torch_utils/ops/bias_act.py:92-125 (_bias_act_ref) for the forward
// and torch_utils/ops/bias_act.cu:24-148 for the gradient modes:
//   grad = 0: y = clamp(act(x + b) * gain)
//   grad = 1: x is dy; returns d/dx of the forward, evaluated from xref (+b) / yref
//   grad = 2: x is d_dx; returns the second-order term (dy supplied), for act with has_2nd_grad
// Activation ids are the reference's cuda_idx (1 linear .. 9 swish).
// ---------------------------------------------------------------------------------------------
struct BiasActArgs {
    const void* x; const void* b; const void* xref; const void* yref; const void* dy; void* y;
    int grad, act; float alpha, gain, clamp;
    long sizeX; int sizeB; long stepB;
};

template <int A>
__device__ __forceinline__ float bias_act_eval(int G, float x, float xref, float yy, float alpha, float& yref, float gain) {
    const float expRange = 80.f, halfExpRange = 40.f;
    const float seluScale = 1.0507009873554804934193349852946f, seluAlpha = 1.6732632423543772848170429916717f;
    float y = 0.f;
    if (A == 1) { if (G <= 1) y = x; }
    else if (A == 2) { y = (G == 0) ? fmaxf(x, 0.f) : (G == 1 ? (yy > 0.f ? x : 0.f) : 0.f); }
    else if (A == 3) { y = (G == 0) ? (x > 0.f ? x : x * alpha) : (G == 1 ? (yy > 0.f ? x : x * alpha) : 0.f); }
    else if (A == 4) {
        if (G == 0) { const float c = expf(x), d = 1.f / c; y = (x < -expRange) ? -1.f : (x > expRange) ? 1.f : (c - d) / (c + d); }
        else if (G == 1) y = x * (1.f - yy * yy);
        else y = x * (1.f - yy * yy) * (-2.f * yy);
    } else if (A == 5) {
        if (G == 0) y = (x < -expRange) ? 0.f : 1.f / (expf(-x) + 1.f);
        else if (G == 1) y = x * yy * (1.f - yy);
        else y = x * yy * (1.f - yy) * (1.f - 2.f * yy);
    } else if (A == 6) {
        if (G == 0) y = (x >= 0.f) ? x : expf(x) - 1.f;
        else if (G == 1) y = (yy >= 0.f) ? x : x * (yy + 1.f);
        else y = (yy >= 0.f) ? 0.f : x * (yy + 1.f);
    } else if (A == 7) {
        if (G == 0) y = (x >= 0.f) ? seluScale * x : (seluScale * seluAlpha) * (expf(x) - 1.f);
        else if (G == 1) y = (yy >= 0.f) ? x * seluScale : x * (yy + seluScale * seluAlpha);
        else y = (yy >= 0.f) ? 0.f : x * (yy + seluScale * seluAlpha);
    } else if (A == 8) {
        if (G == 0) y = (x > expRange) ? x : logf(expf(x) + 1.f);
        else if (G == 1) y = x * (1.f - expf(-yy));
        else { const float c = expf(-yy); y = x * c * (1.f - c); }
    } else if (A == 9) {
        if (G == 0) y = (x < -expRange) ? 0.f : x / (expf(-x) + 1.f);
        else {
            const float c = expf(xref), d = c + 1.f;
            if (G == 1) y = (xref > halfExpRange) ? x : x * c * (xref + d) / (d * d);
            else y = (xref > halfExpRange) ? 0.f : x * c * (xref * (2.f - d) + 2.f * d) / (d * d * d);
            yref = (xref < -expRange) ? 0.f : xref / (expf(-xref) + 1.f) * gain;
        }
    }
    return y;
}

template <typename T, int A>
__global__ void __launch_bounds__(256) bias_act_kernel(BiasActArgs p) {
    const T* x = (const T*)p.x; const T* b = (const T*)p.b; const T* xr = (const T*)p.xref;
    const T* yr = (const T*)p.yref; const T* dyp = (const T*)p.dy; T* y = (T*)p.y;
    const int G = p.grad;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < p.sizeX; i += (long)gridDim.x * blockDim.x) {
        float xv = ldf<T>(x, i);
        const float bv = b ? ldf<T>(b, (i / p.stepB) % p.sizeB) : 0.f;
        float xref = xr ? ldf<T>(xr, i) : 0.f;
        float yref = yr ? ldf<T>(yr, i) : 0.f;
        const float dy = dyp ? ldf<T>(dyp, i) : 1.f;
        const float yy = (p.gain != 0.f) ? yref / p.gain : 0.f;
        if (G == 0) xv += bv; else xref += bv;
        float out = bias_act_eval<A>(G, xv, xref, yy, p.alpha, yref, p.gain);
        out *= p.gain * dy;
        if (p.clamp >= 0.f) {
            if (G == 0) out = (out > -p.clamp && out < p.clamp) ? out : (out >= 0.f ? p.clamp : -p.clamp);
            else out = (yref > -p.clamp && yref < p.clamp) ? out : 0.f;
        }
        stf<T>(y, i, out);
    }
}

// fp32, four consecutive elements per thread (128-bit accesses): usable when the bias is constant over groups of four
// (stepB % 4 == 0, i.e. NCHW with H*W % 4 == 0, or no bias) and every pointer is 16-byte aligned.
template <int A>
__global__ void __launch_bounds__(256) bias_act_vec4_kernel(BiasActArgs p) {
    const float4* x = (const float4*)p.x; const float* b = (const float*)p.b; const float4* xr = (const float4*)p.xref;
    const float4* yr = (const float4*)p.yref; const float4* dyp = (const float4*)p.dy; float4* y = (float4*)p.y;
    const int G = p.grad;
    const long n4 = p.sizeX >> 2;
    const float inv_gain = (p.gain != 0.f) ? 1.f / p.gain : 0.f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        const float4 xa = x[i];
        const float bv = b ? __ldg(b + ((i * 4) / p.stepB) % p.sizeB) : 0.f;
        const float4 xra = xr ? xr[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 yra = yr ? yr[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 dya = dyp ? dyp[i] : make_float4(1.f, 1.f, 1.f, 1.f);
        const float xs[4] = {xa.x, xa.y, xa.z, xa.w}, xrs[4] = {xra.x, xra.y, xra.z, xra.w};
        const float yrs[4] = {yra.x, yra.y, yra.z, yra.w}, dys[4] = {dya.x, dya.y, dya.z, dya.w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float xv = xs[j], xref = xrs[j], yref = yrs[j];
            const float yy = (p.gain != 0.f) ? yref / p.gain : 0.f;
            (void)inv_gain;
            if (G == 0) xv += bv; else xref += bv;
            float out = bias_act_eval<A>(G, xv, xref, yy, p.alpha, yref, p.gain);
            out *= p.gain * dys[j];
            if (p.clamp >= 0.f) {
                if (G == 0) out = (out > -p.clamp && out < p.clamp) ? out : (out >= 0.f ? p.clamp : -p.clamp);
                else out = (yref > -p.clamp && yref < p.clamp) ? out : 0.f;
            }
            o[j] = out;
        }
        y[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
}

template <typename T> struct BiasActVec { static bool launch(const BiasActArgs&, cudaStream_t) { return false; } };
template <> struct BiasActVec<float> {
    static bool launch(const BiasActArgs& p, cudaStream_t st) {
        auto al = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
        if (p.sizeX % 4 != 0 || (p.b && p.stepB % 4 != 0) || !al(p.x) || !al(p.xref) || !al(p.yref) || !al(p.dy) || !al(p.y)) return false;
        const int grid = ew_grid(p.sizeX / 4, 256);
        switch (p.act) {
            case 1: bias_act_vec4_kernel<1><<<grid, 256, 0, st>>>(p); break;
            case 2: bias_act_vec4_kernel<2><<<grid, 256, 0, st>>>(p); break;
            case 3: bias_act_vec4_kernel<3><<<grid, 256, 0, st>>>(p); break;
            case 4: bias_act_vec4_kernel<4><<<grid, 256, 0, st>>>(p); break;
            case 5: bias_act_vec4_kernel<5><<<grid, 256, 0, st>>>(p); break;
            case 6: bias_act_vec4_kernel<6><<<grid, 256, 0, st>>>(p); break;
            case 7: bias_act_vec4_kernel<7><<<grid, 256, 0, st>>>(p); break;
            case 8: bias_act_vec4_kernel<8><<<grid, 256, 0, st>>>(p); break;
            case 9: bias_act_vec4_kernel<9><<<grid, 256, 0, st>>>(p); break;
            default: return false;
        }
        return true;
    }
};

template <typename T>
int launch_bias_act(const BiasActArgs& p, cudaStream_t st) {
    if (BiasActVec<T>::launch(p, st)) return 0;
    const int grid = ew_grid(p.sizeX, 256);
    switch (p.act) {
        case 1: bias_act_kernel<T, 1><<<grid, 256, 0, st>>>(p); break;
        case 2: bias_act_kernel<T, 2><<<grid, 256, 0, st>>>(p); break;
        case 3: bias_act_kernel<T, 3><<<grid, 256, 0, st>>>(p); break;
        case 4: bias_act_kernel<T, 4><<<grid, 256, 0, st>>>(p); break;
        case 5: bias_act_kernel<T, 5><<<grid, 256, 0, st>>>(p); break;
        case 6: bias_act_kernel<T, 6><<<grid, 256, 0, st>>>(p); break;
        case 7: bias_act_kernel<T, 7><<<grid, 256, 0, st>>>(p); break;
        case 8: bias_act_kernel<T, 8><<<grid, 256, 0, st>>>(p); break;
        case 9: bias_act_kernel<T, 9><<<grid, 256, 0, st>>>(p); break;
        default: set_last_error("bias_act: unknown activation id %d", p.act); return LD_ERR_INVALID_ARG;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// generic helpers
// ---------------------------------------------------------------------------------------------
// dst[r, 0:cols_dst] = (c < cols_src) ? src[r, c] : 0   with dtype conversion
template <typename TS, typename TD>
__global__ void cast_pad_kernel(const TS* __restrict__ src, long lds, TD* __restrict__ dst, long ldd, long rows, int cols_src, int cols_dst) {
    const long total = rows * (long)cols_dst;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long r = i / cols_dst; const int c = (int)(i - r * cols_dst);
        stf<TD>(dst, r * ldd + c, c < cols_src ? ldf<TS>(src, r * lds + c) : 0.f);
    }
}

// contiguous vectorised cast (n % 4 == 0, 16-byte aligned)
__global__ void cast_f32_bf16_vec_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, long n4) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        const float4 a = src[i];
        uint2 o; o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w);
        dst[i] = o;
    }
}

// out[i] = a[i] * alpha + b[i % period] * beta
template <typename TA, typename TB, typename TO>
__global__ void axpby_bcast_kernel(const TA* __restrict__ a, const TB* __restrict__ b, TO* __restrict__ out, long n, long period, float alpha, float beta) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        stf<TO>(out, i, ldf<TA>(a, i) * alpha + ldf<TB>(b, i % period) * beta);
}

// dx = dy * act'(.)  for activations fused into the GEMM epilogue.
//   relu / lrelu / sigmoid use the activation OUTPUT y (ref = y); gelu uses the PRE-activation (ref = x).
template <typename TG, typename TR, typename TO>
__global__ void act_bwd_kernel(const TG* __restrict__ dy, const TR* __restrict__ ref, TO* __restrict__ dx, long n, int act, float gain) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const float g = ldf<TG>(dy, i) * gain; const float r = ldf<TR>(ref, i);
        float d;
        switch (act) {
            case LD_ACT_RELU:    d = r > 0.f ? g : 0.f; break;
            case LD_ACT_LRELU:   d = r > 0.f ? g : 0.2f * g; break;
            case LD_ACT_GELU:    d = g * gelu_erf_grad(r); break;
            case LD_ACT_SIGMOID: { const float y = r / gain; d = g * y * (1.f - y); } break;
            default:             d = g; break;
        }
        stf<TO>(dx, i, d);
    }
}

// bf16 x 8 per thread (n % 8 == 0, 16-byte aligned pointers)
__global__ void __launch_bounds__(256) act_bwd_vec_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ ref, uint4* __restrict__ dx,
                                                          long n8, int act, float gain) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
        const uint4 ga = dy[i], ra = ref[i];
        const uint32_t gw[4] = {ga.x, ga.y, ga.z, ga.w}, rw[4] = {ra.x, ra.y, ra.z, ra.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float g[2], r[2], d[2];
            unpack_bf16x2(gw[j], g[0], g[1]); unpack_bf16x2(rw[j], r[0], r[1]);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const float gg = g[k] * gain;
                switch (act) {
                    case LD_ACT_RELU:    d[k] = r[k] > 0.f ? gg : 0.f; break;
                    case LD_ACT_LRELU:   d[k] = r[k] > 0.f ? gg : 0.2f * gg; break;
                    case LD_ACT_GELU:    d[k] = gg * gelu_erf_grad(r[k]); break;
                    case LD_ACT_SIGMOID: { const float y = r[k] / gain; d[k] = gg * y * (1.f - y); } break;
                    default:             d[k] = gg; break;
                }
            }
            o[j] = pack_bf16x2(d[0], d[1]);
        }
        dx[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// y = act(x) * gain for activations that could not be fused into a GEMM epilogue (training-mode GELU keeps the
// pre-activation as the GEMM output and applies the activation here)
__global__ void act_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long n8, int act, float gain) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
        const uint4 a = reinterpret_cast<const uint4*>(x)[i];
        const uint32_t w[4] = {a.x, a.y, a.z, a.w};
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float lo, hi;
            unpack_bf16x2(w[k], lo, hi);
            if (act == LD_ACT_GELU) { lo = gelu_erf(lo); hi = gelu_erf(hi); }
            else if (act == LD_ACT_RELU) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
            else if (act == LD_ACT_LRELU) { lo = lo > 0.f ? lo : 0.2f * lo; hi = hi > 0.f ? hi : 0.2f * hi; }
            o[k] = pack_bf16x2(lo * gain, hi * gain);
        }
        reinterpret_cast<uint4*>(y)[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// dx = dy * act'(ref) (bf16) and, in the same pass, dxs = dx * cs[col] — the conv + FrozenBN + (residual) + ReLU backward
// needs both: dx feeds the residual branch, dxs the convolution (scale = w * rsqrt(var + eps), detr_backbone.py:55-65).
__global__ void act_bwd_colscale_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ ref, uint4* __restrict__ dx,
                                        uint4* __restrict__ dxs, const float* __restrict__ cs, long n8, int cols, int act) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
        const uint4 g = dy[i]; const uint4 r = ref[i];
        const int c0 = (int)((i * 8) % cols);
        const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, rw[4] = {r.x, r.y, r.z, r.w};
        uint32_t o[4], os[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float g0, g1, r0, r1;
            unpack_bf16x2(gw[k], g0, g1); unpack_bf16x2(rw[k], r0, r1);
            if (act == LD_ACT_RELU) { g0 = r0 > 0.f ? g0 : 0.f; g1 = r1 > 0.f ? g1 : 0.f; }
            else if (act == LD_ACT_LRELU) { g0 = r0 > 0.f ? g0 : 0.2f * g0; g1 = r1 > 0.f ? g1 : 0.2f * g1; }
            o[k] = pack_bf16x2(g0, g1);
            os[k] = pack_bf16x2(g0 * cs[c0 + 2 * k], g1 * cs[c0 + 2 * k + 1]);
        }
        if (dx) dx[i] = make_uint4(o[0], o[1], o[2], o[3]);
        dxs[i] = make_uint4(os[0], os[1], os[2], os[3]);
    }
}

// column sums of a [rows, cols] matrix accumulated into out[cols] (bias gradients)
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, long ld, float* __restrict__ out, long rows, int cols, int rows_per_block) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    const long r0 = (long)blockIdx.y * rows_per_block;
    const long r1 = min(rows, r0 + rows_per_block);
    float s = 0.f;
    for (long r = r0; r < r1; ++r) s += ldf<T>(x, r * ld + c);
    atomicAdd(out + c, s);
}

// bf16, eight columns per thread (128-bit loads), 8 row phases per block reduced through shared memory, 4 rows in flight
// per thread: the scalar version issued one dependent 2-byte load per row per thread.
__global__ void __launch_bounds__(256) colsum_vec8_kernel(const uint4* __restrict__ x, long ld8, float* __restrict__ out, long rows, int cols8,
                                                          int rows_per_block) {
    __shared__ float red[8][32][9];
    const int cg = blockIdx.x * 32 + threadIdx.x;                 // column group (8 columns)
    const long r0 = (long)blockIdx.y * rows_per_block;
    const long r1 = min(rows, r0 + rows_per_block);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (cg < cols8) {
        long r = r0 + threadIdx.y;
        for (; r + 24 < r1; r += 32) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = __ldg(x + (r + 8 * u) * ld8 + cg);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) { float lo, hi; unpack_bf16x2(w[j], lo, hi); acc[2 * j] += lo; acc[2 * j + 1] += hi; }
            }
        }
        for (; r < r1; r += 8) {
            const uint4 v = __ldg(x + r * ld8 + cg);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) { float lo, hi; unpack_bf16x2(w[j], lo, hi); acc[2 * j] += lo; acc[2 * j + 1] += hi; }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[threadIdx.y][threadIdx.x][j] = acc[j];
    __syncthreads();
    const int t = threadIdx.y * 32 + threadIdx.x;                 // 256 threads -> 32 column groups x 8 columns
    const int g = t >> 3, j = t & 7;
    if (blockIdx.x * 32 + g < cols8) {
        float s = 0.f;
#pragma unroll
        for (int y = 0; y < 8; ++y) s += red[y][g][j];
        atomicAdd(out + (long)(blockIdx.x * 32 + g) * 8 + j, s);
    }
}

// y[b, i, c] = x[b, i, c] * s[b, c]   (channels-last per-sample modulation; inner = C)
template <typename TX, typename TO>
__global__ void scale_channels_kernel(const TX* __restrict__ x, const float* __restrict__ s, TO* __restrict__ y, long n, long per_sample, int C) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const long b = i / per_sample; const int c = (int)(i % C);
        stf<TO>(y, i, ldf<TX>(x, i) * s[b * C + c]);
    }
}

// bf16 -> bf16, eight channels per thread (C % 8 == 0, 16-byte aligned): 128-bit accesses, one index decode per vector
__global__ void __launch_bounds__(256) scale_channels_vec_kernel(const uint4* __restrict__ x, const float* __restrict__ s, uint4* __restrict__ y,
                                                                 long n8, long per_sample8, int C8) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
        const long b = i / per_sample8; const int c8 = (int)(i % C8);
        const float4* sp = reinterpret_cast<const float4*>(s + (b * C8 + c8) * 8);
        const float4 s0 = __ldg(sp), s1 = __ldg(sp + 1);
        const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        const uint4 a = x[i];
        const uint32_t w[4] = {a.x, a.y, a.z, a.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float lo, hi; unpack_bf16x2(w[j], lo, hi);
            o[j] = pack_bf16x2(lo * sc[2 * j], hi * sc[2 * j + 1]);
        }
        y[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

}  // namespace

extern "C" {

// Mirrors bias_act_plugin.bias_act (torch_utils/ops/bias_act.cpp:33-97): caller-owned output.
int ld_bias_act(const void* x, const void* b, const void* xref, const void* yref, const void* dy, void* y,
                int dtype, int grad, int act, float alpha, float gain, float clamp,
                int64_t sizeX, int sizeB, int64_t stepB, void* stream) {
    LD_CHECK_ARG(x && y && sizeX > 0, "bias_act: null tensor or empty size");
    LD_CHECK_ARG(grad >= 0 && grad <= 2, "bias_act: grad must be 0, 1 or 2");
    LD_CHECK_ARG(b == nullptr || (sizeB > 0 && stepB > 0), "bias_act: bias given but sizeB/stepB invalid");
    LD_CHECK_ARG(dtype == LD_F32 || dtype == LD_BF16, "bias_act: dtype must be f32 or bf16");
    BiasActArgs p{x, b, xref, yref, dy, y, grad, act, alpha, gain, clamp, (long)sizeX, sizeB > 0 ? sizeB : 1, stepB > 0 ? (long)stepB : 1};
    int e = (dtype == LD_F32) ? launch_bias_act<float>(p, (cudaStream_t)stream) : launch_bias_act<__nv_bfloat16>(p, (cudaStream_t)stream);
    if (e) return e;
    ld::count_launch();
    LD_LAUNCH_CHECK("bias_act");
    return 0;
}

// fma(a, b, c) = a * b + c with numpy-style broadcasting of b and c against a contiguous 4-D `a`
// (torch_utils/ops/fma.py:16; used by modulated_conv2d's demodulate+noise branch, networks_stylegan2.py:70).
// b_strides / c_strides are element strides in a's index space, 0 on broadcast dims.
struct FmaArgs { const float* a; const float* b; const float* c; float* y; long n; int d1, d2, d3; long bs[4], cs[4]; };
__global__ void fma_kernel(FmaArgs p) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (long)gridDim.x * blockDim.x) {
        long t = i;
        const int i3 = (int)(t % p.d3); t /= p.d3;
        const int i2 = (int)(t % p.d2); t /= p.d2;
        const int i1 = (int)(t % p.d1); const long i0 = t / p.d1;
        const float bv = p.b[i0 * p.bs[0] + i1 * p.bs[1] + i2 * p.bs[2] + i3 * p.bs[3]];
        const float cv = p.c[i0 * p.cs[0] + i1 * p.cs[1] + i2 * p.cs[2] + i3 * p.cs[3]];
        p.y[i] = fmaf(p.a[i], bv, cv);
    }
}
int ld_fma_f32(const float* a, const float* b, const float* c, float* y, const int64_t* a_shape,
               const int64_t* b_strides, const int64_t* c_strides, void* stream) {
    LD_CHECK_ARG(a && b && c && y && a_shape && b_strides && c_strides, "fma: null pointer");
    FmaArgs p; p.a = a; p.b = b; p.c = c; p.y = y;
    p.n = a_shape[0] * a_shape[1] * a_shape[2] * a_shape[3];
    LD_CHECK_ARG(p.n > 0, "fma: empty tensor");
    p.d1 = (int)a_shape[1]; p.d2 = (int)a_shape[2]; p.d3 = (int)a_shape[3];
    for (int i = 0; i < 4; ++i) { p.bs[i] = b_strides[i]; p.cs[i] = c_strides[i]; }
    fma_kernel<<<ew_grid(p.n, 256), 256, 0, (cudaStream_t)stream>>>(p);
    ld::count_launch();
    LD_LAUNCH_CHECK("fma");
    return 0;
}

int ld_cast_pad(const void* src, int src_dtype, int64_t lds, void* dst, int dst_dtype, int64_t ldd,
                int64_t rows, int cols_src, int cols_dst, void* stream) {
    LD_CHECK_ARG(src && dst && rows > 0 && cols_src > 0 && cols_dst >= cols_src, "cast_pad: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const long total = rows * (long)cols_dst;
    if (src_dtype == LD_F32 && dst_dtype == LD_BF16 && cols_src == cols_dst && lds == cols_src && ldd == cols_dst &&
        total % 4 == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 7) == 0) {
        cast_f32_bf16_vec_kernel<<<ew_grid(total / 4, 256), 256, 0, st>>>((const float4*)src, (uint2*)dst, total / 4);
    } else {
        const int grid = ew_grid(total, 256);
#define CP(TS, TD) cast_pad_kernel<TS, TD><<<grid, 256, 0, st>>>((const TS*)src, lds, (TD*)dst, ldd, rows, cols_src, cols_dst)
        if (src_dtype == LD_F32 && dst_dtype == LD_BF16) CP(float, __nv_bfloat16);
        else if (src_dtype == LD_BF16 && dst_dtype == LD_F32) CP(__nv_bfloat16, float);
        else if (src_dtype == LD_F32) CP(float, float);
        else CP(__nv_bfloat16, __nv_bfloat16);
#undef CP
    }
    ld::count_launch();
    LD_LAUNCH_CHECK("cast_pad");
    return 0;
}

int ld_axpby_bcast(const void* a, int a_dtype, const void* b, int b_dtype, void* out, int out_dtype,
                   int64_t n, int64_t period, float alpha, float beta, void* stream) {
    LD_CHECK_ARG(a && b && out && n > 0 && period > 0, "axpby_bcast: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = ew_grid(n, 256);
#define AX(TA, TB, TO) axpby_bcast_kernel<TA, TB, TO><<<grid, 256, 0, st>>>((const TA*)a, (const TB*)b, (TO*)out, n, period, alpha, beta)
    const int key = a_dtype * 4 + b_dtype * 2 + out_dtype;
    switch (key) {
        case 0: AX(float, float, float); break;
        case 1: AX(float, float, __nv_bfloat16); break;
        case 2: AX(float, __nv_bfloat16, float); break;
        case 3: AX(float, __nv_bfloat16, __nv_bfloat16); break;
        case 4: AX(__nv_bfloat16, float, float); break;
        case 5: AX(__nv_bfloat16, float, __nv_bfloat16); break;
        case 6: AX(__nv_bfloat16, __nv_bfloat16, float); break;
        default: AX(__nv_bfloat16, __nv_bfloat16, __nv_bfloat16); break;
    }
#undef AX
    ld::count_launch();
    LD_LAUNCH_CHECK("axpby_bcast");
    return 0;
}

int ld_act_bwd(const void* dy, int dy_dtype, const void* ref, int ref_dtype, void* dx, int dx_dtype,
               int64_t n, int act, float gain, void* stream) {
    LD_CHECK_ARG(dy && ref && dx && n > 0, "act_bwd: bad argument");
    LD_CHECK_ARG(dy_dtype == LD_BF16 && ref_dtype == LD_BF16 && dx_dtype == LD_BF16, "act_bwd: bf16 only");
    if (n % 8 == 0 && ((((uintptr_t)dy | (uintptr_t)ref | (uintptr_t)dx) & 15) == 0))
        act_bwd_vec_kernel<<<ew_grid(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)dy, (const uint4*)ref, (uint4*)dx, n / 8, act, gain);
    else
        act_bwd_kernel<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16><<<ew_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(
            (const __nv_bfloat16*)dy, (const __nv_bfloat16*)ref, (__nv_bfloat16*)dx, n, act, gain);
    ld::count_launch();
    LD_LAUNCH_CHECK("act_bwd");
    return 0;
}

int ld_act_bwd_colscale(const void* dy, const void* ref, void* dx, void* dxs, const float* cs, int64_t n, int cols, int act, void* stream) {
    LD_CHECK_ARG(dy && dxs && cs && n > 0 && cols > 0 && n % 8 == 0 && cols % 8 == 0, "act_bwd_colscale: bad argument");
    LD_CHECK_ARG(act == LD_ACT_NONE || ref, "act_bwd_colscale: activation needs its output");
    act_bwd_colscale_kernel<<<ew_grid(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)dy, (const uint4*)(ref ? ref : dy), (uint4*)dx, (uint4*)dxs, cs, n / 8, cols, act);
    ld::count_launch();
    LD_LAUNCH_CHECK("act_bwd_colscale");
    return 0;
}

int ld_act_fwd_bf16(const void* x, void* y, int64_t n, int act, float gain, void* stream) {
    LD_CHECK_ARG(x && y && n > 0 && n % 8 == 0, "act_fwd: n must be a positive multiple of 8");
    LD_CHECK_ARG((((uintptr_t)x | (uintptr_t)y) & 15) == 0, "act_fwd: pointers must be 16-byte aligned");
    act_fwd_kernel<<<ew_grid(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, n / 8, act, gain);
    ld::count_launch();
    LD_LAUNCH_CHECK("act_fwd");
    return 0;
}

int ld_colsum_accum(const void* x, int dtype, int64_t ld_, float* out, int64_t rows, int cols, void* stream) {
    LD_CHECK_ARG(x && out && rows > 0 && cols > 0, "colsum: bad argument");
    if (dtype == LD_BF16 && cols % 8 == 0 && ld_ % 8 == 0 && ((uintptr_t)x & 15) == 0) {
        const int gx = ld::ceil_div(cols / 8, 32);
        const long want = std::max<long>(1, (long)ld::sm_count() * 4 / gx);          // ~4 blocks per SM in total
        const int rpbv = (int)std::max<long>(32, (rows + want - 1) / want);
        dim3 gridv(gx, (unsigned)ld::ceil_div((int)rows, rpbv));
        colsum_vec8_kernel<<<gridv, dim3(32, 8), 0, (cudaStream_t)stream>>>((const uint4*)x, ld_ / 8, out, rows, cols / 8, rpbv);
        ld::count_launch();
        LD_LAUNCH_CHECK("colsum");
        return 0;
    }
    const int rpb = (int)std::max<long>(32, (rows + 255) / 256);
    dim3 grid(ld::ceil_div(cols, 128), ld::ceil_div(rows, rpb));
    if (dtype == LD_F32) colsum_kernel<float><<<grid, 128, 0, (cudaStream_t)stream>>>((const float*)x, ld_, out, rows, cols, rpb);
    else colsum_kernel<__nv_bfloat16><<<grid, 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, ld_, out, rows, cols, rpb);
    ld::count_launch();
    LD_LAUNCH_CHECK("colsum");
    return 0;
}

int ld_scale_channels(const void* x, int x_dtype, const float* s, void* y, int y_dtype, int64_t n, int64_t per_sample, int C, void* stream) {
    LD_CHECK_ARG(x && s && y && n > 0 && per_sample > 0 && C > 0 && per_sample % C == 0, "scale_channels: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (x_dtype == LD_BF16 && y_dtype == LD_BF16 && C % 8 == 0 && ((((uintptr_t)x | (uintptr_t)y) & 15) == 0) && (((uintptr_t)s) & 15) == 0) {
        scale_channels_vec_kernel<<<ew_grid(n / 8, 256), 256, 0, st>>>((const uint4*)x, s, (uint4*)y, n / 8, per_sample / 8, C / 8);
        ld::count_launch();
        LD_LAUNCH_CHECK("scale_channels");
        return 0;
    }
    const int grid = ew_grid(n, 256);
    if (x_dtype == LD_BF16 && y_dtype == LD_BF16) scale_channels_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, s, (__nv_bfloat16*)y, n, per_sample, C);
    else if (x_dtype == LD_F32 && y_dtype == LD_BF16) scale_channels_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>((const float*)x, s, (__nv_bfloat16*)y, n, per_sample, C);
    else if (x_dtype == LD_BF16 && y_dtype == LD_F32) scale_channels_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, s, (float*)y, n, per_sample, C);
    else scale_channels_kernel<float, float><<<grid, 256, 0, st>>>((const float*)x, s, (float*)y, n, per_sample, C);
    ld::count_launch();
    LD_LAUNCH_CHECK("scale_channels");
    return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// StyleGAN2 modulated-conv epilogue on channels-last activations (reference
// training/networks_stylegan2.py:66-75 + SynthesisLayer.forward :307-325, non-fused modconv path):
//   y[b,p,c] = act(x[b,p,c] * d[b,c] + bias[c]) * gain          d = demodulation coefficients
// and its backward:
//   g = dy * gain * act'(y);  dx = g * d;  dd[b,c] += sum_p g * x;  dbias[c] += sum_p,b g
// act: LD_ACT_NONE or LD_ACT_LRELU (slope 0.2).
// ---------------------------------------------------------------------------------------------
namespace {
template <typename TX>
__global__ void demod_bias_act_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ d, const float* __restrict__ bias,
                                          __nv_bfloat16* __restrict__ y, long n, long per_sample, int C, int act, float gain) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const long b = i / per_sample; const int c = (int)(i % C);
        float v = ldf<TX>(x, i);
        if (d) v *= d[b * C + c];
        if (bias) v += bias[c];
        if (act == LD_ACT_LRELU) v = v > 0.f ? v : 0.2f * v;
        y[i] = f32_to_bf16(v * gain);
    }
}

// bf16 x, eight channels per thread and item (C % 8 == 0, 16-byte aligned).  A block owns 1024 consecutive items; a thread takes
// four of them 256 apart and issues its four 128-bit loads before any arithmetic (round 1: one load in flight per thread and a
// 64-bit division per item: 44 % of the HBM peak).  32-bit index arithmetic (n8 < 2^31, checked by the launcher).
__global__ void __launch_bounds__(256)
demod_bias_act_fwd_vec_kernel(const uint4* __restrict__ x, const float* __restrict__ d, const float* __restrict__ bias,
                              uint4* __restrict__ y, long n8, long per_sample8, int C8, int act, float gain) {
    const uint32_t base = blockIdx.x * 1024u + threadIdx.x;
    const uint32_t n = (uint32_t)n8, ps = (uint32_t)per_sample8;
    uint4 a[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t i = base + 256u * k;
        if (i < n) a[k] = x[i];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t i = base + 256u * k;
        if (i >= n) break;
        const uint32_t b = i / ps; const uint32_t c8 = i % (uint32_t)C8;
        float dc[8], bc[8];
        if (d) {
            const float4* dp = reinterpret_cast<const float4*>(d + ((long)b * C8 + c8) * 8);
            const float4 d0 = __ldg(dp), d1 = __ldg(dp + 1);
            dc[0] = d0.x; dc[1] = d0.y; dc[2] = d0.z; dc[3] = d0.w; dc[4] = d1.x; dc[5] = d1.y; dc[6] = d1.z; dc[7] = d1.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) dc[j] = 1.f;
        }
        if (bias) {
            const float4* bp = reinterpret_cast<const float4*>(bias + c8 * 8);
            const float4 b0 = __ldg(bp), b1 = __ldg(bp + 1);
            bc[0] = b0.x; bc[1] = b0.y; bc[2] = b0.z; bc[3] = b0.w; bc[4] = b1.x; bc[5] = b1.y; bc[6] = b1.z; bc[7] = b1.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) bc[j] = 0.f;
        }
        const uint32_t w[4] = {a[k].x, a[k].y, a[k].z, a[k].w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v[2]; unpack_bf16x2(w[j], v[0], v[1]);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                float t = fmaf(v[q], dc[2 * j + q], bc[2 * j + q]);
                if (act == LD_ACT_LRELU) t = t > 0.f ? t : 0.2f * t;
                v[q] = t * gain;
            }
            o[j] = pack_bf16x2(v[0], v[1]);
        }
        y[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// one block handles a strip of pixels of one sample; channel partial sums in registers -> atomics
template <typename TX>
__global__ void __launch_bounds__(256)
demod_bias_act_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y, const TX* __restrict__ x,
                          const float* __restrict__ d, __nv_bfloat16* __restrict__ dx, float* __restrict__ dd, float* __restrict__ dbias,
                          long pixels, int C, int act, float gain, int pix_per_block) {
    const int b = blockIdx.y;
    const long p0 = (long)blockIdx.x * pix_per_block;
    const long p1 = min(pixels, p0 + pix_per_block);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float dc = d ? d[(long)b * C + c] : 1.f;
        float sd = 0.f, sb = 0.f;
        for (long p = p0; p < p1; ++p) {
            const long i = ((long)b * pixels + p) * C + c;
            float g = bf16_to_f32(dy[i]) * gain;
            if (act == LD_ACT_LRELU && bf16_to_f32(y[i]) < 0.f) g *= 0.2f;
            sd += g * ldf<TX>(x, i);
            sb += g;
            dx[i] = f32_to_bf16(g * dc);
        }
        if (dd) atomicAdd(dd + (long)b * C + c, sd);
        if (dbias) atomicAdd(dbias + c, sb);
    }
}

// out[b,c] += sum_p a[b,p,c] * g[b,p,c]     (gradient of the per-sample style modulation x * s[b,c])
template <typename TA>
__global__ void __launch_bounds__(256)
channel_dot_kernel(const TA* __restrict__ a, const __nv_bfloat16* __restrict__ g, float* __restrict__ out, long pixels, int C, int pix_per_block) {
    const int b = blockIdx.y;
    const long p0 = (long)blockIdx.x * pix_per_block;
    const long p1 = min(pixels, p0 + pix_per_block);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (long p = p0; p < p1; ++p) {
            const long i = ((long)b * pixels + p) * C + c;
            s += ldf<TA>(a, i) * bf16_to_f32(g[i]);
        }
        atomicAdd(out + (long)b * C + c, s);
    }
}

// Vectorised variants (C % 8 == 0, bf16 x): thread = (pixel, 8 channels); per-block smem reduction of the channel sums,
// one global atomic per channel per block.
__global__ void __launch_bounds__(256)
demod_bias_act_bwd_vec_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ y, const uint4* __restrict__ x,
                              const float* __restrict__ d, uint4* __restrict__ dx, float* __restrict__ dd, float* __restrict__ dbias,
                              long pixels, int C, int act, float gain, int pix_per_block) {
    extern __shared__ float red[];                       // [2][C]
    const int C8 = C >> 3;
    const int b = blockIdx.y;
    const int cg = threadIdx.x % C8, prow = threadIdx.x / C8, prows = blockDim.x / C8;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    float dc[8], sd[8], sb[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { dc[k] = d ? d[(long)b * C + cg * 8 + k] : 1.f; sd[k] = 0.f; sb[k] = 0.f; }
    const long p0 = (long)blockIdx.x * pix_per_block;
    const long p1 = min(pixels, p0 + pix_per_block);
    if (prow < prows) {
        for (long p = p0 + prow; p < p1; p += prows) {
            const long i = ((long)b * pixels + p) * C8 + cg;
            const uint4 g4 = dy[i], y4 = y[i], x4 = x[i];
            const uint32_t gw[4] = {g4.x, g4.y, g4.z, g4.w}, yw[4] = {y4.x, y4.y, y4.z, y4.w}, xw[4] = {x4.x, x4.y, x4.z, x4.w};
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float g0, g1, y0, y1, x0, x1;
                unpack_bf16x2(gw[k], g0, g1); unpack_bf16x2(yw[k], y0, y1); unpack_bf16x2(xw[k], x0, x1);
                g0 *= gain; g1 *= gain;
                if (act == LD_ACT_LRELU) { if (y0 < 0.f) g0 *= 0.2f; if (y1 < 0.f) g1 *= 0.2f; }
                sd[2 * k] += g0 * x0; sd[2 * k + 1] += g1 * x1;
                sb[2 * k] += g0; sb[2 * k + 1] += g1;
                o[k] = pack_bf16x2(g0 * dc[2 * k], g1 * dc[2 * k + 1]);
            }
            dx[i] = make_uint4(o[0], o[1], o[2], o[3]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) { atomicAdd(&red[cg * 8 + k], sd[k]); atomicAdd(&red[C + cg * 8 + k], sb[k]); }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        if (dd) atomicAdd(dd + (long)b * C + c, red[c]);
        if (dbias) atomicAdd(dbias + c, red[C + c]);
    }
}

// Deterministic variants of the two style-gradient reductions (ld_*_ws): every thread row keeps its partial sums in its own
// shared-memory slot, the block adds them in a fixed order and writes ONE partial per (block, sample, channel) to a caller
// workspace; partial_sum_accum_kernel then adds the blocks in index order.  No floating-point atomics anywhere, so the result
// does not depend on scheduling (the atomics versions flip single bf16 roundings downstream from run to run).
__global__ void __launch_bounds__(256)
demod_bias_act_bwd_det_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ y, const uint4* __restrict__ x,
                              const float* __restrict__ d, uint4* __restrict__ dx, float* __restrict__ ws_dd, float* __restrict__ ws_db,
                              long pixels, int C, int act, float gain, int pix_per_block) {
    extern __shared__ float red[];                       // [prows][2][C]
    const int C8 = C >> 3;
    const int b = blockIdx.y;
    const int cg = threadIdx.x % C8, prow = threadIdx.x / C8, prows = blockDim.x / C8;
    float dc[8], sd[8], sb[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { dc[k] = d ? d[(long)b * C + cg * 8 + k] : 1.f; sd[k] = 0.f; sb[k] = 0.f; }
    const long p0 = (long)blockIdx.x * pix_per_block;
    const long p1 = min(pixels, p0 + pix_per_block);
    if (prow < prows) {
        for (long p = p0 + prow; p < p1; p += prows) {
            const long i = ((long)b * pixels + p) * C8 + cg;
            const uint4 g4 = dy[i], y4 = y[i], x4 = x[i];
            const uint32_t gw[4] = {g4.x, g4.y, g4.z, g4.w}, yw[4] = {y4.x, y4.y, y4.z, y4.w}, xw[4] = {x4.x, x4.y, x4.z, x4.w};
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float g0, g1, y0, y1, x0, x1;
                unpack_bf16x2(gw[k], g0, g1); unpack_bf16x2(yw[k], y0, y1); unpack_bf16x2(xw[k], x0, x1);
                g0 *= gain; g1 *= gain;
                if (act == LD_ACT_LRELU) { if (y0 < 0.f) g0 *= 0.2f; if (y1 < 0.f) g1 *= 0.2f; }
                sd[2 * k] += g0 * x0; sd[2 * k + 1] += g1 * x1;
                sb[2 * k] += g0; sb[2 * k + 1] += g1;
                o[k] = pack_bf16x2(g0 * dc[2 * k], g1 * dc[2 * k + 1]);
            }
            dx[i] = make_uint4(o[0], o[1], o[2], o[3]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) { red[(prow * 2 + 0) * C + cg * 8 + k] = sd[k]; red[(prow * 2 + 1) * C + cg * 8 + k] = sb[k]; }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = 0.f, bsum = 0.f;
        for (int r = 0; r < prows; ++r) { a += red[(r * 2 + 0) * C + c]; bsum += red[(r * 2 + 1) * C + c]; }
        const long o = ((long)blockIdx.x * gridDim.y + b) * C + c;
        ws_dd[o] = a;
        ws_db[o] = bsum;
    }
}

__global__ void __launch_bounds__(256)
channel_dot_det_kernel(const uint4* __restrict__ a, const uint4* __restrict__ g, float* __restrict__ ws, long pixels, int C, int pix_per_block) {
    extern __shared__ float red[];                       // [prows][C]
    const int C8 = C >> 3;
    const int b = blockIdx.y;
    const int cg = threadIdx.x % C8, prow = threadIdx.x / C8, prows = blockDim.x / C8;
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const long p0 = (long)blockIdx.x * pix_per_block;
    const long p1 = min(pixels, p0 + pix_per_block);
    if (prow < prows) {
        for (long p = p0 + prow; p < p1; p += prows) {
            const long i = ((long)b * pixels + p) * C8 + cg;
            const uint4 a4 = a[i], g4 = g[i];
            const uint32_t aw[4] = {a4.x, a4.y, a4.z, a4.w}, gw[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float a0, a1, g0, g1;
                unpack_bf16x2(aw[k], a0, a1); unpack_bf16x2(gw[k], g0, g1);
                s[2 * k] += a0 * g0; s[2 * k + 1] += a1 * g1;
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) red[prow * C + cg * 8 + k] = s[k];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float t = 0.f;
        for (int r = 0; r < prows; ++r) t += red[r * C + c];
        ws[((long)blockIdx.x * gridDim.y + b) * C + c] = t;
    }
}

// out[i] += sum over (block k, segment f) of ws[k * n + f * per + i], per = n / fold (fold > 1: the bias gradient also sums over the
// batch).  One warp per output element: lane l adds the terms l, l + 32, ... in order, then a fixed shuffle tree — the same
// association every run, so the result is deterministic.
__global__ void __launch_bounds__(256) partial_sum_accum_kernel(const float* __restrict__ ws, float* __restrict__ out, int nblk, long n, int fold) {
    const long per = n / fold;
    const long i = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= per) return;
    const int lane = threadIdx.x & 31;
    const int terms = nblk * fold;
    float t = 0.f;
    for (int q = lane; q < terms; q += 32) {
        const int k = q / fold, f = q - k * fold;
        t += ws[(long)k * n + (long)f * per + i];
    }
    t = warp_sum(t);
    if (lane == 0) out[i] += t;
}

__global__ void __launch_bounds__(256)
channel_dot_vec_kernel(const uint4* __restrict__ a, const uint4* __restrict__ g, float* __restrict__ out, long pixels, int C, int pix_per_block) {
    extern __shared__ float red[];                       // [C]
    const int C8 = C >> 3;
    const int b = blockIdx.y;
    const int cg = threadIdx.x % C8, prow = threadIdx.x / C8, prows = blockDim.x / C8;
    for (int i = threadIdx.x; i < C; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const long p0 = (long)blockIdx.x * pix_per_block;
    const long p1 = min(pixels, p0 + pix_per_block);
    if (prow < prows) {
        for (long p = p0 + prow; p < p1; p += prows) {
            const long i = ((long)b * pixels + p) * C8 + cg;
            const uint4 a4 = a[i], g4 = g[i];
            const uint32_t aw[4] = {a4.x, a4.y, a4.z, a4.w}, gw[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float a0, a1, g0, g1;
                unpack_bf16x2(aw[k], a0, a1); unpack_bf16x2(gw[k], g0, g1);
                s[2 * k] += a0 * g0; s[2 * k + 1] += a1 * g1;
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(&red[cg * 8 + k], s[k]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(out + (long)b * C + c, red[c]);
}
}  // namespace

extern "C" {
int ld_demod_bias_act_fwd(const void* x, int x_dtype, const float* d, const float* bias, void* y_bf16,
                          int B, int64_t pixels, int C, int act, float gain, void* stream) {
    LD_CHECK_ARG(x && y_bf16 && B > 0 && pixels > 0 && C > 0, "demod_bias_act_fwd: bad argument");
    const long n = (long)B * pixels * C;
    if (x_dtype == LD_BF16 && C % 8 == 0 && ((((uintptr_t)x | (uintptr_t)y_bf16) & 15) == 0) && n / 8 < (1L << 31) - 1024 &&
        (!d || ((uintptr_t)d & 15) == 0) && (!bias || ((uintptr_t)bias & 15) == 0)) {
        demod_bias_act_fwd_vec_kernel<<<(unsigned)((n / 8 + 1023) / 1024), 256, 0, (cudaStream_t)stream>>>(
            (const uint4*)x, d, bias, (uint4*)y_bf16, n / 8, pixels * C / 8, C / 8, act, gain);
        ld::count_launch();
        LD_LAUNCH_CHECK("demod_bias_act_fwd");
        return 0;
    }
    const int grid = ew_grid(n, 256);
    if (x_dtype == LD_F32) demod_bias_act_fwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, d, bias, (__nv_bfloat16*)y_bf16, n, pixels * C, C, act, gain);
    else demod_bias_act_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, d, bias, (__nv_bfloat16*)y_bf16, n, pixels * C, C, act, gain);
    ld::count_launch();
    LD_LAUNCH_CHECK("demod_bias_act_fwd");
    return 0;
}

int ld_demod_bias_act_bwd(const void* dy_bf16, const void* y_bf16, const void* x, int x_dtype, const float* d,
                          void* dx_bf16, float* dd, float* dbias, int B, int64_t pixels, int C, int act, float gain, void* stream) {
    LD_CHECK_ARG(dy_bf16 && y_bf16 && x && dx_bf16 && B > 0 && pixels > 0 && C > 0, "demod_bias_act_bwd: bad argument");
    if (x_dtype == LD_BF16 && C % 8 == 0 && C <= 2048 && 256 % (C / 8 > 256 ? 256 : C / 8) == 0 && C / 8 <= 256 &&
        ((((uintptr_t)dy_bf16 | (uintptr_t)y_bf16 | (uintptr_t)x | (uintptr_t)dx_bf16) & 15) == 0)) {
        const int prows = 256 / (C / 8);
        const long want_blocks = std::max<long>(1, (long)ld::sm_count() * 4 / B);
        int ppbv = (int)std::max<long>(prows, (pixels + want_blocks - 1) / want_blocks);
        dim3 gridv((unsigned)((pixels + ppbv - 1) / ppbv), (unsigned)B);
        demod_bias_act_bwd_vec_kernel<<<gridv, 256, 2 * C * sizeof(float), (cudaStream_t)stream>>>(
            (const uint4*)dy_bf16, (const uint4*)y_bf16, (const uint4*)x, d, (uint4*)dx_bf16, dd, dbias, pixels, C, act, gain, ppbv);
        ld::count_launch();
        LD_LAUNCH_CHECK("demod_bias_act_bwd");
        return 0;
    }
    const int ppb = (int)std::max<long>(1, std::min<long>(256, pixels / 4 + 1));
    dim3 grid((unsigned)((pixels + ppb - 1) / ppb), (unsigned)B);
    if (x_dtype == LD_F32) demod_bias_act_bwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dy_bf16, (const __nv_bfloat16*)y_bf16, (const float*)x, d, (__nv_bfloat16*)dx_bf16, dd, dbias, pixels, C, act, gain, ppb);
    else demod_bias_act_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dy_bf16, (const __nv_bfloat16*)y_bf16, (const __nv_bfloat16*)x, d, (__nv_bfloat16*)dx_bf16, dd, dbias, pixels, C, act, gain, ppb);
    ld::count_launch();
    LD_LAUNCH_CHECK("demod_bias_act_bwd");
    return 0;
}

// number of pixel blocks the vectorised style reductions use for (B, pixels, C): the _ws variants need nblk * B * C floats per sum
static inline void style_reduce_geom(int B, int64_t pixels, int C, int& prows, int& ppbv, unsigned& nblk) {
    prows = 256 / (C / 8);
    const long want_blocks = std::max<long>(1, (long)ld::sm_count() * 4 / B);
    ppbv = (int)std::max<long>(prows, (pixels + want_blocks - 1) / want_blocks);
    nblk = (unsigned)((pixels + ppbv - 1) / ppbv);
}
static inline bool style_vec_ok(int C) { return C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0 && C <= 2048; }

int64_t ld_style_reduce_ws_floats(int B, int64_t pixels, int C) {
    if (B <= 0 || pixels <= 0 || C <= 0 || !style_vec_ok(C)) return 0;
    int prows, ppbv; unsigned nblk;
    style_reduce_geom(B, pixels, C, prows, ppbv, nblk);
    return (int64_t)nblk * B * C;
}

int ld_demod_bias_act_bwd_ws(const void* dy_bf16, const void* y_bf16, const void* x_bf16, const float* d, void* dx_bf16, float* dd, float* dbias,
                             float* ws, int64_t ws_floats, int B, int64_t pixels, int C, int act, float gain, void* stream) {
    LD_CHECK_ARG(dy_bf16 && y_bf16 && x_bf16 && dx_bf16 && ws && B > 0 && pixels > 0 && C > 0, "demod_bias_act_bwd_ws: bad argument");
    LD_CHECK_ARG(style_vec_ok(C) && ((((uintptr_t)dy_bf16 | (uintptr_t)y_bf16 | (uintptr_t)x_bf16 | (uintptr_t)dx_bf16) & 15) == 0),
                 "demod_bias_act_bwd_ws: needs bf16 rows of C %% 8 == 0 channels, 16-byte aligned");
    int prows, ppbv; unsigned nblk;
    style_reduce_geom(B, pixels, C, prows, ppbv, nblk);
    const long per = (long)nblk * B * C;
    LD_CHECK_ARG(ws_floats >= 2 * per, "demod_bias_act_bwd_ws: workspace of %lld floats, need %ld", (long long)ws_floats, 2 * per);
    cudaStream_t st = (cudaStream_t)stream;
    dim3 gridv(nblk, (unsigned)B);
    demod_bias_act_bwd_det_kernel<<<gridv, 256, (size_t)prows * 2 * C * sizeof(float), st>>>(
        (const uint4*)dy_bf16, (const uint4*)y_bf16, (const uint4*)x_bf16, d, (uint4*)dx_bf16, ws, ws + per, pixels, C, act, gain, ppbv);
    const long n = (long)B * C;
    if (dd) partial_sum_accum_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(ws, dd, (int)nblk, n, 1);
    if (dbias) partial_sum_accum_kernel<<<(unsigned)((C + 7) / 8), 256, 0, st>>>(ws + per, dbias, (int)nblk, n, B);
    ld::count_launch(1 + (dd ? 1 : 0) + (dbias ? 1 : 0));
    LD_LAUNCH_CHECK("demod_bias_act_bwd_ws");
    return 0;
}

int ld_channel_dot_ws(const void* a_bf16, const void* g_bf16, float* out, float* ws, int64_t ws_floats, int B, int64_t pixels, int C, void* stream) {
    LD_CHECK_ARG(a_bf16 && g_bf16 && out && ws && B > 0 && pixels > 0 && C > 0, "channel_dot_ws: bad argument");
    LD_CHECK_ARG(style_vec_ok(C) && ((((uintptr_t)a_bf16 | (uintptr_t)g_bf16) & 15) == 0), "channel_dot_ws: needs bf16 rows of C %% 8 == 0 channels, 16-byte aligned");
    int prows, ppbv; unsigned nblk;
    style_reduce_geom(B, pixels, C, prows, ppbv, nblk);
    const long n = (long)B * C;
    LD_CHECK_ARG(ws_floats >= (long)nblk * n, "channel_dot_ws: workspace of %lld floats, need %ld", (long long)ws_floats, (long)nblk * n);
    cudaStream_t st = (cudaStream_t)stream;
    dim3 gridv(nblk, (unsigned)B);
    channel_dot_det_kernel<<<gridv, 256, (size_t)prows * C * sizeof(float), st>>>((const uint4*)a_bf16, (const uint4*)g_bf16, ws, pixels, C, ppbv);
    partial_sum_accum_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(ws, out, (int)nblk, n, 1);
    ld::count_launch(2);
    LD_LAUNCH_CHECK("channel_dot_ws");
    return 0;
}

int ld_channel_dot(const void* a, int a_dtype, const void* g_bf16, float* out, int B, int64_t pixels, int C, void* stream) {
    LD_CHECK_ARG(a && g_bf16 && out && B > 0 && pixels > 0 && C > 0, "channel_dot: bad argument");
    if (a_dtype == LD_BF16 && C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0 && ((((uintptr_t)a | (uintptr_t)g_bf16) & 15) == 0)) {
        const int prows = 256 / (C / 8);
        const long want_blocks = std::max<long>(1, (long)ld::sm_count() * 4 / B);
        int ppbv = (int)std::max<long>(prows, (pixels + want_blocks - 1) / want_blocks);
        dim3 gridv((unsigned)((pixels + ppbv - 1) / ppbv), (unsigned)B);
        channel_dot_vec_kernel<<<gridv, 256, C * sizeof(float), (cudaStream_t)stream>>>((const uint4*)a, (const uint4*)g_bf16, out, pixels, C, ppbv);
        ld::count_launch();
        LD_LAUNCH_CHECK("channel_dot");
        return 0;
    }
    const int ppb = (int)std::max<long>(1, std::min<long>(256, pixels / 4 + 1));
    dim3 grid((unsigned)((pixels + ppb - 1) / ppb), (unsigned)B);
    if (a_dtype == LD_F32) channel_dot_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)a, (const __nv_bfloat16*)g_bf16, out, pixels, C, ppb);
    else channel_dot_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)g_bf16, out, pixels, C, ppb);
    ld::count_launch();
    LD_LAUNCH_CHECK("channel_dot");
    return 0;
}
}  // extern "C"


// ---------------------------------------------------------------------------------------------------------------------
// Data-loader tail: uint8 HWC image batch -> fp32 NCHW normalised with per-channel mean / std, the arithmetic of the
// reference loader (training/dataset_layoutganpp.py:333-336: x.astype(float32) / 255.0 - mean) / std) in the same order
// with IEEE fp32 division, so the result is bit-identical to NumPy's while the host ships 1 byte per value instead of 4.
// ---------------------------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256)
normalize_u8_hwc_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, long pixels_per_image, long total_pixels,
                        float m0, float m1, float m2, float s0, float s1, float s2) {
    const float mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
    for (long g = (long)blockIdx.x * blockDim.x + threadIdx.x; g * 4 < total_pixels; g += (long)gridDim.x * blockDim.x) {
        const long p0 = g * 4;                                    // four consecutive pixels of one image (pixels_per_image % 4 == 0)
        const long b = p0 / pixels_per_image, q = p0 - b * pixels_per_image;
        const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src + p0 * 3);          // 12 bytes = 3 aligned words
        const uint32_t w0 = __ldg(s32), w1 = __ldg(s32 + 1), w2 = __ldg(s32 + 2);
        const uint8_t v[12] = {(uint8_t)w0, (uint8_t)(w0 >> 8), (uint8_t)(w0 >> 16), (uint8_t)(w0 >> 24),
                               (uint8_t)w1, (uint8_t)(w1 >> 8), (uint8_t)(w1 >> 16), (uint8_t)(w1 >> 24),
                               (uint8_t)w2, (uint8_t)(w2 >> 8), (uint8_t)(w2 >> 16), (uint8_t)(w2 >> 24)};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                o[k] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v[3 * k + c], 255.0f), mean[c]), stdv[c]);
            *reinterpret_cast<float4*>(dst + (b * 3 + c) * pixels_per_image + q) = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}
}  // namespace

extern "C" int ld_normalize_u8_image(const uint8_t* src_hwc, float* dst_nchw, int64_t B, int64_t H, int64_t W, const float* mean3,
                                     const float* std3, void* stream) {
    LD_CHECK_ARG(src_hwc && dst_nchw && mean3 && std3 && B > 0 && H > 0 && W > 0, "normalize_u8_image: bad argument");
    const long ppi = (long)H * W;
    LD_CHECK_ARG(ppi % 4 == 0, "normalize_u8_image: H*W = %ld must be a multiple of 4", ppi);
    LD_CHECK_ARG((reinterpret_cast<uintptr_t>(src_hwc) & 3) == 0 && (reinterpret_cast<uintptr_t>(dst_nchw) & 15) == 0,
                 "normalize_u8_image: buffers must be 4- / 16-byte aligned");
    const long groups = (long)B * ppi / 4;
    const int grid = (int)std::min<long>((groups + 255) / 256, (long)ld::sm_count() * 8);
    normalize_u8_hwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src_hwc, dst_nchw, ppi, (long)B * ppi, mean3[0], mean3[1], mean3[2],
                                                                    std3[0], std3[1], std3[2]);
    ld::count_launch();
    LD_LAUNCH_CHECK("normalize_u8_image");
    return 0;
}
