// Batched bf16 GEMM for sm_100a: TMA (SWIZZLE_128B) -> smem ring -> tcgen05.mma (cta_group::1,
// 128 x {128,256} x 16 UMMA, fp32 accumulators double-buffered in TMEM) -> fused epilogue.
//
// Persistent, warp-specialised: warps 0..7 = epilogue (two warps per TMEM lane quadrant, each taking half of the
// tile's columns), warp 9 = TMEM allocator, warp 10 = TMA producer, warp 11 = MMA issuer.  The single-thread
// producer / issuer roles sit in the HIGHEST warp ids on purpose: the sub-partition arbiter picks the eligible warp
// with the highest id first, so an arithmetic-heavy epilogue (GELU: its warps are always eligible) cannot starve
// the thread that feeds the tensor pipe.  The accumulator of tile i+1 is produced while the epilogue drains tile i.  The common
// epilogues (per-column scale/bias staged in smem, activation, residual, bf16/fp32 vector stores) are
// compile-time specialised; rare ones (aux copy, accumulate, unaligned rows) take a generic path.
//
// Operand "major" flags let one kernel serve forward (A K-major, B K-major), dgrad (B MN-major) and
// wgrad (A and B MN-major) without materialising transposes; two batch dims with arbitrary strides
// let attention run straight on the fused [B*T, 3*H*d] QKV buffer.
#include "common.cuh"
#include "runtime.h"
#include "../../include/layoutdetr_sm100.h"
#include <cstdlib>

namespace {

using namespace ld;

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;          // 16 KB
constexpr int B_STAGE_BYTES_MAX = 256 * BK * 2;     // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES_MAX;
constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;    // 192 KB operand ring (1-SM: 4 x 48 KB, CTA-pair: 6 x 32 KB)
constexpr int STAGES_2SM = 6;
constexpr int STAGE_BYTES_2SM = 32768;              // per CTA: A 128 x 64 (16 KB) + B 128 x 64 (16 KB)
constexpr int BAR_BYTES = 256;
constexpr int EPI_STAGE_BYTES = 2 * 2 * 256 * 4;    // [acc stage][scale|bias][256] fp32
constexpr int EPI_XPOSE_BYTES = 8 * 2048;           // per epilogue warp: 32 rows x 64 B transpose buffer
constexpr int SMEM_BYTES = PIPE_BYTES + BAR_BYTES + EPI_STAGE_BYTES + EPI_XPOSE_BYTES + 1024;  // +1024 alignment
constexpr int NUM_THREADS = 384;
constexpr int EPI_WARPS = 8;
constexpr int WARP_ALLOC = 9, WARP_PRODUCER = 10, WARP_MMA = 11;
constexpr int TMEM_COLS = 512;
constexpr int ACC_STRIDE = 256;                     // TMEM columns per accumulator stage

struct KParams {
    int M, N, K, nb1, nb2;
    int act, accumulate, split_k, d_dtype, r_dtype, bn;
    int a_mn, b_mn;
    int m_tiles, n_tiles, kb_total, kb_per_split, total_tiles;
    uint32_t mg_split, mg_n, mg_m, mg_b2;   // magic multipliers ceil(2^32 / d) of the tile-index divisors (0: use '/')
    int tile_m;      // 128 (one CTA per tile) or 256 (CTA pair, cta_group::2)
    int vec_ok;
    int fast;        // compile-time specialised epilogue usable (aligned rows, store-only, no aux)
    float alpha, post_gain;
    void* D; long ldd, d_sb1, d_sb2;
    void* aux;
    const void* R; long ldr, r_sb1, r_sb2;
    const float* cs; const float* cb; long col_sb1, col_sb2;
    const float* alpha_dev;
    int softmax, causal; float mask_value; const uint8_t* key_mask;
    // implicit-GEMM convolution (ld_conv_gemm_bf16): 0 off, 1 = A rows are output pixels, 2 = B (MN-major) rows are output pixels
    int cv_mode, cv_P, cv_Wo, cv_stride, cv_pad, cv_KW, cv_C;
    int k32;         // K blocks of 32 (64-byte rows, SWIZZLE_64B): implicit convolution over 32-channel images (mode 1, single-CTA kernel)
};

// First output pixel of a box -> TMA coordinates (w, h, b) of tap (kh, kw) in the NHWC image.
__device__ __forceinline__ void conv_coords(const KParams& p, int pix0, int kh, int kw, int& cw, int& chh, int& cb) {
    const int b0 = pix0 / p.cv_P;
    const int rem = pix0 - b0 * p.cv_P;
    const int y0 = rem / p.cv_Wo;
    const int x0 = rem - y0 * p.cv_Wo;
    cw = x0 * p.cv_stride + kw - p.cv_pad;
    chh = y0 * p.cv_stride + kh - p.cv_pad;
    cb = b0;
}

struct Tile {
    int b1, b2, m0, n0, kb_begin, kb_end;
};

// t / d and t % d with a host-validated multiply-high (integer division is ~25 instructions and every warp of the CTA
// decodes every tile).
__device__ __forceinline__ void divmod(int t, int d, uint32_t magic, int& q, int& r) {
    if (d == 1) { q = t; r = 0; return; }
    const int qq = magic ? (int)__umulhi((uint32_t)t, magic) : t / d;
    const int rr = t - qq * d;          // (q may alias t)
    q = qq; r = rr;
}

__device__ __forceinline__ Tile decode_tile(const KParams& p, int t) {
    Tile tl;
    int ks, nt, mt;
    divmod(t, p.split_k, p.mg_split, t, ks);
    divmod(t, p.n_tiles, p.mg_n, t, nt);
    divmod(t, p.m_tiles, p.mg_m, t, mt);
    divmod(t, p.nb2, p.mg_b2, tl.b1, tl.b2);
    tl.m0 = mt * p.tile_m;
    tl.n0 = nt * p.bn;
    tl.kb_begin = ks * p.kb_per_split;
    tl.kb_end = min(p.kb_total, tl.kb_begin + p.kb_per_split);
    return tl;
}

template <int ACT>
__device__ __forceinline__ float act_ct(float v) {
    if (ACT == LD_ACT_RELU) return fmaxf(v, 0.0f);
    if (ACT == LD_ACT_GELU) return gelu_erf(v);
    if (ACT == LD_ACT_LRELU) return v > 0.0f ? v : 0.2f * v;
    if (ACT == LD_ACT_SIGMOID) return __fdividef(1.0f, 1.0f + __expf(-v));
    return v;
}

__device__ __forceinline__ float apply_act(float v, int act) {
    switch (act) {
        case LD_ACT_RELU:    return fmaxf(v, 0.0f);
        case LD_ACT_GELU:    return gelu_erf(v);
        case LD_ACT_LRELU:   return v > 0.0f ? v : 0.2f * v;
        case LD_ACT_SIGMOID: return 1.0f / (1.0f + __expf(-v));
        default:             return v;
    }
}


// Generic (slow-path) epilogue for one 16-column chunk held in registers: every feature, scalar fallbacks.
__device__ __forceinline__ void epilogue_generic_chunk(const KParams& p, const uint32_t (&r)[16], float alpha, int n_base,
                                                       long d_off, long r_off, long c_off) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) * alpha;
    const bool full_chunk = (n_base + 16 <= p.N);
    if (p.cs) {
#pragma unroll
        for (int i = 0; i < 16; ++i) if (full_chunk || n_base + i < p.N) v[i] *= __ldg(p.cs + c_off + n_base + i);
    }
    if (p.cb) {
#pragma unroll
        for (int i = 0; i < 16; ++i) if (full_chunk || n_base + i < p.N) v[i] += __ldg(p.cb + c_off + n_base + i);
    }
    const bool vec = p.vec_ok && full_chunk;
    if (p.R) {
        if (p.r_dtype == LD_BF16) {
            const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.R) + r_off + n_base;
            if (vec) {
                const uint4 a = __ldg(reinterpret_cast<const uint4*>(rp));
                const uint4 b = __ldg(reinterpret_cast<const uint4*>(rp) + 1);
                const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) { float lo, hi; unpack_bf16x2(w[i], lo, hi); v[2 * i] += lo; v[2 * i + 1] += hi; }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) if (n_base + i < p.N) v[i] += bf16_to_f32(rp[i]);
            }
        } else {
            const float* rp = reinterpret_cast<const float*>(p.R) + r_off + n_base;
            if (vec) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(rp) + i);
        v[4 * i] += a.x; v[4 * i + 1] += a.y; v[4 * i + 2] += a.z; v[4 * i + 3] += a.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) if (n_base + i < p.N) v[i] += rp[i];
            }
        }
    }
    if (p.aux) {
        if (p.d_dtype == LD_BF16) {
            __nv_bfloat16* ap = reinterpret_cast<__nv_bfloat16*>(p.aux) + d_off + n_base;
            if (vec) {
                uint4 a, b;
                a.x = pack_bf16x2(v[0], v[1]);   a.y = pack_bf16x2(v[2], v[3]);   a.z = pack_bf16x2(v[4], v[5]);   a.w = pack_bf16x2(v[6], v[7]);
                b.x = pack_bf16x2(v[8], v[9]);   b.y = pack_bf16x2(v[10], v[11]); b.z = pack_bf16x2(v[12], v[13]); b.w = pack_bf16x2(v[14], v[15]);
                reinterpret_cast<uint4*>(ap)[0] = a; reinterpret_cast<uint4*>(ap)[1] = b;
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) if (n_base + i < p.N) ap[i] = f32_to_bf16(v[i]);
            }
        } else {
            float* ap = reinterpret_cast<float*>(p.aux) + d_off + n_base;
#pragma unroll
            for (int i = 0; i < 16; ++i) if (full_chunk || n_base + i < p.N) ap[i] = v[i];
        }
    }
    if (p.act != LD_ACT_NONE) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = apply_act(v[i], p.act);
    }
    if (p.post_gain != 1.0f) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] *= p.post_gain;
    }
    if (p.d_dtype == LD_BF16) {
        __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(p.D) + d_off + n_base;
        if (p.accumulate == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) if (full_chunk || n_base + i < p.N) v[i] += bf16_to_f32(dp[i]);
        }
        if (vec) {
            uint4 a, b;
            a.x = pack_bf16x2(v[0], v[1]);   a.y = pack_bf16x2(v[2], v[3]);   a.z = pack_bf16x2(v[4], v[5]);   a.w = pack_bf16x2(v[6], v[7]);
            b.x = pack_bf16x2(v[8], v[9]);   b.y = pack_bf16x2(v[10], v[11]); b.z = pack_bf16x2(v[12], v[13]); b.w = pack_bf16x2(v[14], v[15]);
            reinterpret_cast<uint4*>(dp)[0] = a; reinterpret_cast<uint4*>(dp)[1] = b;
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) if (n_base + i < p.N) dp[i] = f32_to_bf16(v[i]);
        }
    } else {
        float* dp = reinterpret_cast<float*>(p.D) + d_off + n_base;
        if (p.accumulate == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) if (full_chunk || n_base + i < p.N) atomicAdd(dp + i, v[i]);
        } else {
            if (p.accumulate == 1) {
#pragma unroll
                for (int i = 0; i < 16; ++i) if (full_chunk || n_base + i < p.N) v[i] += dp[i];
            }
            if (vec) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
        reinterpret_cast<float4*>(dp)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) if (n_base + i < p.N) dp[i] = v[i];
            }
        }
    }
}

// Fast-path epilogue for 32 columns of one row: v = act(acc*alpha*cs + cb (+R)) * gain -> 128-bit stores.
// cs_a / cb_a / stg_a are SHARED-space addresses (per-column scale / bias of this tile; this warp's 2 KB transpose buffer).
template <int ACT, bool OUT_BF16, int RES>   // RES: 0 none, 1 bf16, 2 fp32
__device__ __forceinline__ void epilogue_fast_chunk(const KParams& p, const uint32_t (&r)[32], float alpha, uint32_t cs_a,
                                                    uint32_t cb_a, int c0, int n_base, long d_base, int row0, long r_off,
                                                    uint32_t stg_a) {
    // alpha is folded into the staged scale; for a linear epilogue without residual so is post_gain (scale and bias)
    // (relu / lrelu are positively homogeneous: act(g v) = g act(v) for g > 0); GELU epilogues never carry a gain (host check)
    constexpr bool FOLD_GAIN = (RES == 0 && (ACT == LD_ACT_NONE || ACT == LD_ACT_RELU || ACT == LD_ACT_LRELU)) || ACT == LD_ACT_GELU;
    float v[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 s4 = lds_f4(cs_a + (c0 + 4 * j) * 4);
        const float4 b4 = lds_f4(cb_a + (c0 + 4 * j) * 4);
        // acc * scale + bias as two FFMA2 (same roundings as four FFMA, half the issue slots)
        f2_unpack(f2_fma(f2_pack(__uint_as_float(r[4 * j + 0]), __uint_as_float(r[4 * j + 1])), f2_pack(s4.x, s4.y), f2_pack(b4.x, b4.y)),
                  v[4 * j + 0], v[4 * j + 1]);
        f2_unpack(f2_fma(f2_pack(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])), f2_pack(s4.z, s4.w), f2_pack(b4.z, b4.w)),
                  v[4 * j + 2], v[4 * j + 3]);
    }
    if (RES == 1 && r_off >= 0) {
        const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.R) + r_off + n_base);
        uint4 a[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] = __ldg(rp + j);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float lo, hi;
            unpack_bf16x2(a[j].x, lo, hi); v[8 * j + 0] += lo; v[8 * j + 1] += hi;
            unpack_bf16x2(a[j].y, lo, hi); v[8 * j + 2] += lo; v[8 * j + 3] += hi;
            unpack_bf16x2(a[j].z, lo, hi); v[8 * j + 4] += lo; v[8 * j + 5] += hi;
            unpack_bf16x2(a[j].w, lo, hi); v[8 * j + 6] += lo; v[8 * j + 7] += hi;
        }
    } else if (RES == 2 && r_off >= 0) {
        const float4* rp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.R) + r_off + n_base);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 a = __ldg(rp + j);
            v[4 * j] += a.x; v[4 * j + 1] += a.y; v[4 * j + 2] += a.z; v[4 * j + 3] += a.w;
        }
    }
    if (ACT == LD_ACT_GELU) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) gelu_erf_x2(v[i], v[i + 1]);      // FFMA2 / FMUL2: the epilogue is issue-bound here
    } else if (ACT != LD_ACT_NONE) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = act_ct<ACT>(v[i]);
    }
    // Row-per-thread registers -> 32 x 64 B smem tile (XOR-swizzled 16 B pieces) -> each store instruction writes
    // eight 64 B row segments (full 32 B sectors) instead of 32 scattered 16 B pieces.
    const float gain = p.post_gain;
#define LD_G(x) (FOLD_GAIN ? (x) : (x) * gain)
    const int lane = threadIdx.x & 31;
    const int sw_w = (lane >> 1) & 3;
    const int pc = lane & 3;
    if (OUT_BF16) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack_bf16x2(LD_G(v[8 * j + 0]), LD_G(v[8 * j + 1])); o.y = pack_bf16x2(LD_G(v[8 * j + 2]), LD_G(v[8 * j + 3]));
            o.z = pack_bf16x2(LD_G(v[8 * j + 4]), LD_G(v[8 * j + 5])); o.w = pack_bf16x2(LD_G(v[8 * j + 6]), LD_G(v[8 * j + 7]));
            sts_u4(stg_a + ((lane * 4 + (j ^ sw_w)) << 4), o);
        }
        __syncwarp();
        __nv_bfloat16* dbase = reinterpret_cast<__nv_bfloat16*>(p.D) + d_base + n_base + pc * 8;
        uint4 val[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rl = (lane >> 2) + 8 * i;
            val[i] = lds_u4(stg_a + ((rl * 4 + (pc ^ ((rl >> 1) & 3))) << 4));
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rl = (lane >> 2) + 8 * i;
            if (row0 + rl < p.M) *reinterpret_cast<uint4*>(dbase + (long)(row0 + rl) * p.ldd) = val[i];
        }
        __syncwarp();
    } else {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int b = 16 * h + 4 * j;
                sts_u4(stg_a + ((lane * 4 + (j ^ sw_w)) << 4),
                       make_uint4(__float_as_uint(LD_G(v[b])), __float_as_uint(LD_G(v[b + 1])),
                                  __float_as_uint(LD_G(v[b + 2])), __float_as_uint(LD_G(v[b + 3]))));
            }
            __syncwarp();
            float* dbase = reinterpret_cast<float*>(p.D) + d_base + n_base + 16 * h + pc * 4;
            uint4 val[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int rl = (lane >> 2) + 8 * i;
                val[i] = lds_u4(stg_a + ((rl * 4 + (pc ^ ((rl >> 1) & 3))) << 4));
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int rl = (lane >> 2) + 8 * i;
                if (row0 + rl < p.M) *reinterpret_cast<uint4*>(dbase + (long)(row0 + rl) * p.ldd) = val[i];
            }
            __syncwarp();
        }
    }
}
#undef LD_G

template <int ACT, bool OUT_BF16, int RES>
__device__ __forceinline__ void epilogue_fast_tile(const KParams& p, const Tile& tl, uint32_t taddr, int col_begin, int col_end,
                                                   bool row_ok, float alpha, uint32_t cs_a, uint32_t cb_a,
                                                   long d_off, long r_off, long c_off, long d_base, int row0, uint32_t stg_a) {
    if (tl.n0 + col_end <= p.N && col_end - col_begin == 128) {
        // Common case (this warp's four 32-column chunks are all inside the matrix): the TMEM load of chunk i + 1 is in flight
        // while chunk i is converted and stored, so the warp waits for one TMEM round trip per tile instead of four.
        uint32_t ra[32], rb[32];
        tmem_ld_x32(taddr + col_begin, ra);
        tmem_ld_wait();
        tmem_ld_x32(taddr + col_begin + 32, rb);
        epilogue_fast_chunk<ACT, OUT_BF16, RES>(p, ra, alpha, cs_a, cb_a, col_begin, tl.n0 + col_begin, d_base, row0, row_ok ? r_off : -1, stg_a);
        tmem_ld_wait();
        tmem_ld_x32(taddr + col_begin + 64, ra);
        epilogue_fast_chunk<ACT, OUT_BF16, RES>(p, rb, alpha, cs_a, cb_a, col_begin + 32, tl.n0 + col_begin + 32, d_base, row0, row_ok ? r_off : -1, stg_a);
        tmem_ld_wait();
        tmem_ld_x32(taddr + col_begin + 96, rb);
        epilogue_fast_chunk<ACT, OUT_BF16, RES>(p, ra, alpha, cs_a, cb_a, col_begin + 64, tl.n0 + col_begin + 64, d_base, row0, row_ok ? r_off : -1, stg_a);
        tmem_ld_wait();
        epilogue_fast_chunk<ACT, OUT_BF16, RES>(p, rb, alpha, cs_a, cb_a, col_begin + 96, tl.n0 + col_begin + 96, d_base, row0, row_ok ? r_off : -1, stg_a);
        return;
    }
    for (int c0 = col_begin; c0 < col_end; c0 += 32) {
        const int n_base = tl.n0 + c0;
        if (n_base >= p.N) break;                       // warp-uniform
        uint32_t r[32];
        tmem_ld_x32(taddr + c0, r);
        tmem_ld_wait();
        if (n_base + 32 <= p.N) {                       // whole warp takes this branch together (transposed stores)
            epilogue_fast_chunk<ACT, OUT_BF16, RES>(p, r, alpha, cs_a, cb_a, c0, n_base, d_base, row0, row_ok ? r_off : -1, stg_a);
        } else if (row_ok) {                            // ragged last chunk of the matrix: scalar path
            uint32_t h[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) h[i] = r[i];
            epilogue_generic_chunk(p, h, alpha, n_base, d_off, r_off, c_off);
            if (n_base + 16 < p.N) {
#pragma unroll
                for (int i = 0; i < 16; ++i) h[i] = r[16 + i];
                epilogue_generic_chunk(p, h, alpha, n_base + 16, d_off, r_off, c_off);
            }
        }
    }
}


// Fused attention-probability epilogue: the whole key axis of a query row sits in this thread's TMEM lane, so
// scale + mask + softmax + bf16 cast happen in the drain and the fp32 score matrix never reaches HBM.
// Three sweeps over TMEM (max, sum, write); add_s = per-key additive mask staged in smem.
__device__ __forceinline__ void epilogue_softmax_tile(const KParams& p, const Tile& tl, uint32_t taddr, bool row_ok, int row,
                                                      float scale, uint32_t add_a, long d_off) {
    const int N = p.N;
    const int nch = (N + 31) >> 5;
    const float neg = p.mask_value;
    float mx = -INFINITY;
    for (int ch = 0; ch < nch; ++ch) {
        uint32_t r[32];
        tmem_ld_x32(taddr + ch * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int c = ch * 32 + i;
            float v = fmaf(__uint_as_float(r[i]), scale, lds_f32(add_a + c * 4));
            if (p.causal && c > row) v += neg;
            if (c < N) mx = fmaxf(mx, v);
        }
    }
    float sum = 0.f;
    for (int ch = 0; ch < nch; ++ch) {
        uint32_t r[32];
        tmem_ld_x32(taddr + ch * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int c = ch * 32 + i;
            float v = fmaf(__uint_as_float(r[i]), scale, lds_f32(add_a + c * 4));
            if (p.causal && c > row) v += neg;
            if (c < N) sum += __expf(v - mx);
        }
    }
    const float inv = __fdividef(1.0f, sum);
    __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(p.D) + d_off;
    const int n_pad = (N + 7) & ~7;
    for (int ch = 0; ch < nch; ++ch) {
        uint32_t r[32];
        tmem_ld_x32(taddr + ch * 32, r);
        tmem_ld_wait();
        float e[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int c = ch * 32 + i;
            float v = fmaf(__uint_as_float(r[i]), scale, lds_f32(add_a + c * 4));
            if (p.causal && c > row) v += neg;
            e[i] = (c < N) ? __expf(v - mx) * inv : 0.f;
        }
        if (row_ok) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = ch * 32 + 8 * j;
                if (c < n_pad) {
                    uint4 o;
                    o.x = pack_bf16x2(e[8 * j + 0], e[8 * j + 1]); o.y = pack_bf16x2(e[8 * j + 2], e[8 * j + 3]);
                    o.z = pack_bf16x2(e[8 * j + 4], e[8 * j + 5]); o.w = pack_bf16x2(e[8 * j + 6], e[8 * j + 7]);
                    *reinterpret_cast<uint4*>(dp + c) = o;
                }
            }
        }
    }
}

template <bool TWO_SM>
__device__ __forceinline__ void gemm_body(const CUtensorMap& tmA, const CUtensorMap& tmB, const KParams& p) {
    constexpr int NSTAGE = TWO_SM ? STAGES_2SM : STAGES;
    constexpr int SBYTES = TWO_SM ? STAGE_BYTES_2SM : STAGE_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar   = reinterpret_cast<uint64_t*>(smem + PIPE_BYTES);
    uint64_t* empty_bar  = full_bar + NSTAGE;
    uint64_t* tfull_bar  = empty_bar + NSTAGE;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot  = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    const uint32_t rank = TWO_SM ? cluster_ctarank() : 0u;        // 0 = leader (issues the MMAs)
    const int unit = TWO_SM ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int nunits = TWO_SM ? (int)(gridDim.x >> 1) : (int)gridDim.x;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    pdl_trigger();                       // the next kernel of this stream may start its own prologue now (no-op without the attribute)
    if (warp == WARP_PRODUCER && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == WARP_MMA && lane == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], TWO_SM ? 2 * EPI_WARPS : EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == WARP_ALLOC) {
        if (TWO_SM) { tmem_alloc_2sm(tmem_slot, TMEM_COLS); tmem_relinquish_2sm(); }
        else { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
    }
    tc_fence_before();
    if (TWO_SM) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                          // operands / residuals written by earlier kernels are complete and visible from here on

    if (warp == WARP_PRODUCER) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            const int b_rows = TWO_SM ? (p.bn >> 1) : p.bn;           // B rows this CTA stages per k-block
            const int bk = p.k32 ? 32 : BK;
            const uint32_t tx_bytes = (TWO_SM ? 2u : 1u) * (BM * bk * 2 + b_rows * bk * 2);
            for (int t = unit; t < p.total_tiles; t += nunits) {
                const Tile tl = decode_tile(p, t);
                const int am0 = tl.m0 + (int)rank * BM;
                const int bn0 = tl.n0 + (int)rank * b_rows;
                for (int kb = tl.kb_begin; kb < tl.kb_end; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
                    uint8_t* sa = smem + stage * SBYTES;
                    uint8_t* sb = sa + A_STAGE_BYTES;
                    const int k0 = kb * bk;
                    if (p.cv_mode == 1) {
                        // K block -> (tap, channel block); the A tile is a box of 128 output pixels x 64 channels of the image
                        const int tap = k0 / p.cv_C, c0 = k0 - tap * p.cv_C;
                        const int kh = tap / p.cv_KW, kw = tap - kh * p.cv_KW;
                        int cw, chh, cb;
                        conv_coords(p, am0, kh, kw, cw, chh, cb);
                        if (TWO_SM) tma_load_4d_2sm(sa, &tmA, &full_bar[stage], c0, cw, chh, cb);
                        else tma_load_4d(sa, &tmA, &full_bar[stage], c0, cw, chh, cb);
                    } else if (!p.a_mn) {
                        if (TWO_SM) tma_load_4d_2sm(sa, &tmA, &full_bar[stage], k0, am0, tl.b2, tl.b1);
                        else tma_load_4d(sa, &tmA, &full_bar[stage], k0, am0, tl.b2, tl.b1);
                    } else {
#pragma unroll
                        for (int j = 0; j < BM / 64; ++j) {
                            if (TWO_SM) tma_load_4d_2sm(sa + j * 8192, &tmA, &full_bar[stage], am0 + 64 * j, k0, tl.b2, tl.b1);
                            else tma_load_4d(sa + j * 8192, &tmA, &full_bar[stage], am0 + 64 * j, k0, tl.b2, tl.b1);
                        }
                    }
                    if (p.cv_mode == 2) {
                        // MN-major B: 64 K-rows = 64 output pixels, each 64-wide N block = 64 channels of one tap
                        for (int j = 0; j < b_rows / 64; ++j) {
                            const int n = bn0 + 64 * j;
                            const int tap = n / p.cv_C, c0 = n - tap * p.cv_C;
                            const int kh = tap / p.cv_KW, kw = tap - kh * p.cv_KW;
                            int cw, chh, cb;
                            conv_coords(p, k0, kh, kw, cw, chh, cb);
                            const int cc = n < p.N ? c0 : p.cv_C;                 // N overhang: out-of-range channel -> zero fill
                            if (TWO_SM) tma_load_4d_2sm(sb + j * 8192, &tmB, &full_bar[stage], cc, cw, chh, cb);
                            else tma_load_4d(sb + j * 8192, &tmB, &full_bar[stage], cc, cw, chh, cb);
                        }
                    } else if (!p.b_mn) {
                        if (TWO_SM) tma_load_4d_2sm(sb, &tmB, &full_bar[stage], k0, bn0, tl.b2, tl.b1);
                        else tma_load_4d(sb, &tmB, &full_bar[stage], k0, bn0, tl.b2, tl.b1);
                    } else {
                        for (int j = 0; j < b_rows / 64; ++j) {
                            if (TWO_SM) tma_load_4d_2sm(sb + j * 8192, &tmB, &full_bar[stage], bn0 + 64 * j, k0, tl.b2, tl.b1);
                            else tma_load_4d(sb + j * 8192, &tmB, &full_bar[stage], bn0 + 64 * j, k0, tl.b2, tl.b1);
                        }
                    }
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == WARP_MMA) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA only)
        if (lane == 0 && rank == 0) {
            int stage = 0; uint32_t phase = 0;
            int as = 0; uint32_t aphase = 0;
            const uint32_t idesc = make_idesc_bf16(TWO_SM ? 256 : BM, p.bn, p.a_mn, p.b_mn);
            for (int t = unit; t < p.total_tiles; t += nunits) {
                const Tile tl = decode_tile(p, t);
                mbar_wait(&tempty_bar[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * ACC_STRIDE;
                for (int kb = tl.kb_begin; kb < tl.kb_end; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * SBYTES);
                    const uint32_t sb = sa + A_STAGE_BYTES;
                    if (p.k32) {                      // 64-byte rows: two UMMA_K steps per K block
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const uint32_t acc = (kb > tl.kb_begin || k > 0) ? 1u : 0u;
                            umma_bf16_ss(tmem_d, make_smem_desc_sw64(sa + k * 32), make_smem_desc_sw64(sb + k * 32), idesc, acc);
                        }
                    } else
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // K-major: 32 B per UMMA_K step inside the 128 B swizzle row; 8-row groups 1024 B apart.
                        // MN-major: 16 K-rows (2 x 1024 B) per step; 64-wide MN blocks 8192 B apart.
                        const uint64_t da = p.a_mn ? make_smem_desc(sa + k * 2048, 8192, 1024)
                                                   : make_smem_desc(sa + k * 32, 16, 1024);
                        const uint64_t db = p.b_mn ? make_smem_desc(sb + k * 2048, 8192, 1024)
                                                   : make_smem_desc(sb + k * 32, 16, 1024);
                        const uint32_t acc = (kb > tl.kb_begin || k > 0) ? 1u : 0u;
                        if (TWO_SM) umma_bf16_ss_2sm(tmem_d, da, db, idesc, acc);
                        else umma_bf16_ss(tmem_d, da, db, idesc, acc);
                    }
                    // frees the smem stage (in both CTAs of a pair) once these MMAs retire
                    if (TWO_SM) umma_commit_2sm(&empty_bar[stage], 3); else umma_commit(&empty_bar[stage]);
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
                // accumulator complete -> epilogue warps of both CTAs
                if (TWO_SM) umma_commit_2sm(&tfull_bar[as], 3); else umma_commit(&tfull_bar[as]);
                as ^= 1; if (as == 0) aphase ^= 1;
            }
        }
    } else if (warp < EPI_WARPS) {
        // ------------------------------------------------------------------ epilogue (8 warps)
        const int e = warp;
        const int q = e & 3;                          // TMEM lane quadrant of this warp (hardware: warp_id % 4)
        const int half = e >> 2;                      // which half of the tile's columns
        const int et = threadIdx.x;                   // 0..255 within the epilogue group
        const uint32_t epi_a = smem_u32(smem + PIPE_BYTES + BAR_BYTES);
        int as = 0; uint32_t aphase = 0;
        const float alpha = p.alpha_dev ? p.alpha * __ldg(p.alpha_dev) : p.alpha;
        const float fold_gain = (!p.R && (p.act == LD_ACT_NONE || p.act == LD_ACT_RELU || p.act == LD_ACT_LRELU))
                                    ? p.post_gain : 1.0f;                                 // see epilogue_fast_chunk
        const int col_begin = half * (p.bn >> 1), col_end = col_begin + (p.bn >> 1);
        for (int t = unit; t < p.total_tiles; t += nunits) {
            Tile tl = decode_tile(p, t);
            tl.m0 += (int)rank * BM;                  // this CTA's 128 rows of the (pair) tile
            const long c_off = (long)tl.b1 * p.col_sb1 + (long)tl.b2 * p.col_sb2;
            const uint32_t cs_a = epi_a + as * 2048;  // shared-space addresses: [acc stage][scale | bias][256] fp32
            const uint32_t cb_a = cs_a + 1024;
            if (p.softmax) {                          // stage the additive key mask of this batch
                const uint8_t* km = p.key_mask ? p.key_mask + (long)tl.b1 * p.N : nullptr;
                sts_f32(cs_a + et * 4, (km && et < p.N && km[et]) ? p.mask_value : 0.0f);
                asm volatile("bar.sync 1, 256;" ::: "memory");
            } else if (p.fast) {                      // stage per-column scale / bias of this tile
                const int n = tl.n0 + et;
                const bool ok = et < p.bn && n < p.N;
                sts_f32(cs_a + et * 4, ((ok && p.cs) ? __ldg(p.cs + c_off + n) : 1.0f) * alpha * fold_gain);
                sts_f32(cb_a + et * 4, ((ok && p.cb) ? __ldg(p.cb + c_off + n) : 0.0f) * fold_gain);
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            mbar_wait(&tfull_bar[as], aphase);
            tc_fence_after();
            const int row = tl.m0 + q * 32 + lane;
            const bool row_ok = row < p.M;
            const long d_off = (long)tl.b1 * p.d_sb1 + (long)tl.b2 * p.d_sb2 + (long)row * p.ldd;
            const long r_off = (long)tl.b1 * p.r_sb1 + (long)tl.b2 * p.r_sb2 + (long)row * p.ldr;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * ACC_STRIDE;
            const long d_base = (long)tl.b1 * p.d_sb1 + (long)tl.b2 * p.d_sb2;
            const int row0 = tl.m0 + q * 32;
            const uint32_t stg_a = epi_a + EPI_STAGE_BYTES + e * 2048;
            if (p.softmax) {
                if (half == 0) epilogue_softmax_tile(p, tl, taddr, row_ok, row, alpha, cs_a, d_off);
            } else if (p.fast) {
#define LD_EPI(ACT, BF, RES) epilogue_fast_tile<ACT, BF, RES>(p, tl, taddr, col_begin, col_end, row_ok, alpha, cs_a, cb_a, d_off, r_off, c_off, d_base, row0, stg_a)
#define LD_EPI_ACT(BF, RES)                                           \
                switch (p.act) {                                      \
                    case LD_ACT_RELU:    LD_EPI(LD_ACT_RELU, BF, RES); break;    \
                    case LD_ACT_GELU:    LD_EPI(LD_ACT_GELU, BF, RES); break;    \
                    case LD_ACT_LRELU:   LD_EPI(LD_ACT_LRELU, BF, RES); break;   \
                    case LD_ACT_SIGMOID: LD_EPI(LD_ACT_SIGMOID, BF, RES); break; \
                    default:             LD_EPI(LD_ACT_NONE, BF, RES); break;    \
                }
                const int res = p.R ? (p.r_dtype == LD_BF16 ? 1 : 2) : 0;
                if (p.d_dtype == LD_BF16) {
                    if (res == 0) { LD_EPI_ACT(true, 0) } else if (res == 1) { LD_EPI_ACT(true, 1) } else { LD_EPI(LD_ACT_NONE, true, 2); }
                } else {
                    if (res == 0) { LD_EPI(LD_ACT_NONE, false, 0); } else if (res == 1) { LD_EPI(LD_ACT_NONE, false, 1); } else { LD_EPI(LD_ACT_NONE, false, 2); }
                }
#undef LD_EPI_ACT
#undef LD_EPI
            } else {
                for (int c0 = col_begin; c0 < col_end; c0 += 16) {
                    const int n_base = tl.n0 + c0;
                    if (n_base >= p.N) break;         // warp-uniform
                    uint32_t r[16];
                    tmem_ld_x16(taddr + c0, r);
                    tmem_ld_wait();
                    if (!row_ok) continue;
                    epilogue_generic_chunk(p, r, alpha, n_base, d_off, r_off, c_off);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (TWO_SM) mbar_arrive_cluster(&tempty_bar[as], 0); else mbar_arrive(&tempty_bar[as]); }
            as ^= 1; if (as == 0) aphase ^= 1;
        }
    }

    tc_fence_before();
    if (TWO_SM) cluster_sync_all(); else __syncthreads();
    if (warp == WARP_ALLOC) { if (TWO_SM) tmem_dealloc_2sm(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS); }
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const KParams p) {
    gemm_body<false>(tmA, tmB, p);
}

// CTA-pair variant: 256 x 256 tile per cluster of two CTAs (tcgen05.mma.cta_group::2); each CTA stages its own 128 rows of
// A and 128 of the 256 B rows, halving the L2->smem operand traffic per MMA.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_2sm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const KParams p) {
    gemm_body<true>(tmA, tmB, p);
}

int make_operand_map(CUtensorMap* tm, const ld_gemm_operand& op, int rows, int K, int nb1, int nb2, int box_rows, bool k32 = false) {
    if (!op.ptr) { set_last_error("gemm: null operand"); return LD_ERR_INVALID_ARG; }
    if ((reinterpret_cast<uintptr_t>(op.ptr) & 15) != 0) { set_last_error("gemm: operand pointer %p not 16-byte aligned", op.ptr); return LD_ERR_ALIGNMENT; }
    if (op.ld % 8 != 0 || (nb2 > 1 && op.sb2 % 8 != 0) || (nb1 > 1 && op.sb1 % 8 != 0)) {
        set_last_error("gemm: operand strides must be multiples of 8 elements (ld=%lld sb1=%lld sb2=%lld)",
                       (long long)op.ld, (long long)op.sb1, (long long)op.sb2);
        return LD_ERR_ALIGNMENT;
    }
    const uint64_t inner = op.mn_major ? (uint64_t)rows : (uint64_t)K;
    const uint64_t outer = op.mn_major ? (uint64_t)K : (uint64_t)rows;
    if ((uint64_t)op.ld < inner && outer > 1) { set_last_error("gemm: ld (%lld) smaller than contiguous extent (%llu)", (long long)op.ld, (unsigned long long)inner); return LD_ERR_INVALID_ARG; }
    const uint64_t dummy = 16;
    uint64_t dims[4] = {inner, outer, (uint64_t)nb2, (uint64_t)nb1};
    uint64_t strides[3] = {(uint64_t)op.ld * 2,
                           nb2 > 1 ? (uint64_t)op.sb2 * 2 : dummy,
                           nb1 > 1 ? (uint64_t)op.sb1 * 2 : dummy};
    if (outer == 1 && strides[0] == 0) strides[0] = dummy;
    for (int i = 0; i < 3; ++i) if (strides[i] == 0) { set_last_error("gemm: zero stride for a batched dimension"); return LD_ERR_INVALID_ARG; }
    const uint32_t box_outer = op.mn_major ? (uint32_t)BK : (uint32_t)box_rows;
    if (k32) return encode_tmap_bf16_4d(tm, op.ptr, dims, strides, 32, box_outer, 64);       // K-major, 32-deep K blocks
    return encode_tmap_bf16_4d(tm, op.ptr, dims, strides, 64, box_outer);
}

}  // namespace

namespace {

// Tensor map of the image operand of an implicit-GEMM convolution: dims (C, W, H, B), box = `pix` output pixels x 64 channels.
int make_conv_map(CUtensorMap* tm, const ld_conv_geom& g, int pix) {
    using namespace ld;
    if (!g.img || (reinterpret_cast<uintptr_t>(g.img) & 15) != 0) { set_last_error("conv gemm: image pointer null or not 16-byte aligned"); return LD_ERR_ALIGNMENT; }
    const bool c32 = g.C == 32 && g.mode == 1;
    if ((g.C % 64 != 0 && !c32) || g.C <= 0) { set_last_error("conv gemm: C = %d must be a positive multiple of 64 (or 32 in mode 1)", g.C); return LD_ERR_INVALID_ARG; }
    if (g.B <= 0 || g.H <= 0 || g.W <= 0 || g.KH <= 0 || g.KW <= 0 || g.stride <= 0 || g.pad < 0) { set_last_error("conv gemm: bad geometry"); return LD_ERR_INVALID_ARG; }
    if (g.Ho != (g.H + 2 * g.pad - g.KH) / g.stride + 1 || g.Wo != (g.W + 2 * g.pad - g.KW) / g.stride + 1 || g.Ho <= 0 || g.Wo <= 0) {
        set_last_error("conv gemm: Ho x Wo = %d x %d inconsistent with H x W = %d x %d, k = %d x %d, stride %d, pad %d", g.Ho, g.Wo, g.H, g.W, g.KH, g.KW, g.stride, g.pad);
        return LD_ERR_INVALID_ARG;
    }
    const int P = g.Ho * g.Wo;
    int bw, bh, bb;
    if (g.Wo >= pix) { if (g.Wo % pix) goto bad; bw = pix; bh = 1; bb = 1; }
    else {
        if (pix % g.Wo) goto bad;
        bw = g.Wo;
        if (P >= pix) { if (P % pix) goto bad; bh = pix / g.Wo; bb = 1; }
        else { if (pix % P) goto bad; bh = g.Ho; bb = pix / P; }
    }
    if (bw * g.stride > 256 || bh * g.stride > 256) goto bad;
    {
        const uint64_t dims[4] = {(uint64_t)g.C, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.B};
        const uint64_t strides[3] = {(uint64_t)g.C * 2, (uint64_t)g.W * g.C * 2, (uint64_t)g.H * g.W * g.C * 2};
        const uint32_t box[4] = {c32 ? 32u : 64u, (uint32_t)(bw * g.stride), (uint32_t)(bh * g.stride), (uint32_t)bb};
        const uint32_t estr[4] = {1u, (uint32_t)g.stride, (uint32_t)g.stride, 1u};
        return encode_tmap_bf16_4d_box(tm, g.img, dims, strides, box, estr, c32 ? 64 : 128);
    }
bad:
    set_last_error("conv gemm: a box of %d output pixels is not a rectangle of whole rows / images for Ho x Wo = %d x %d (stride %d)", pix, g.Ho, g.Wo, g.stride);
    return LD_ERR_INVALID_ARG;
}

int gemm_launch(const ld_gemm_desc* d, const ld_conv_geom* cg, void* stream);

}  // namespace

extern "C" int ld_gemm_bf16(const ld_gemm_desc* d, void* stream) { return gemm_launch(d, nullptr, stream); }

extern "C" int ld_conv_gemm_bf16(const ld_gemm_desc* d, const ld_conv_geom* g, void* stream) {
    using namespace ld;
    LD_CHECK_ARG(d != nullptr && g != nullptr, "conv gemm: null descriptor");
    LD_CHECK_ARG(g->mode == 1 || g->mode == 2, "conv gemm: mode must be 1 (forward / data gradient) or 2 (weight gradient)");
    LD_CHECK_ARG(d->nb1 == 1 && d->nb2 == 1 && !d->softmax, "conv gemm: no batch dims, no softmax epilogue");
    const long pixels = (long)g->B * g->Ho * g->Wo, taps = (long)g->KH * g->KW * g->C;
    if (g->mode == 1) LD_CHECK_ARG(d->M == pixels && d->K == taps && !d->B.mn_major, "conv gemm (mode 1): M = %d, K = %d must be B*Ho*Wo = %ld, KH*KW*C = %ld", d->M, d->K, pixels, taps);
    if (g->mode == 2) LD_CHECK_ARG(d->K == pixels && d->N == taps && d->A.mn_major, "conv gemm (mode 2): K = %d, N = %d must be B*Ho*Wo = %ld, KH*KW*C = %ld (A mn_major)", d->K, d->N, pixels, taps);
    return gemm_launch(d, g, stream);
}

namespace {
int gemm_launch(const ld_gemm_desc* d, const ld_conv_geom* cg, void* stream) {
    using namespace ld;
    LD_CHECK_ARG(d != nullptr, "gemm: null descriptor");
    LD_CHECK_ARG(d->M > 0 && d->N > 0 && d->K > 0 && d->nb1 > 0 && d->nb2 > 0, "gemm: non-positive dims M=%d N=%d K=%d nb=%dx%d", d->M, d->N, d->K, d->nb1, d->nb2);
    LD_CHECK_ARG(d->D != nullptr, "gemm: null output");
    LD_CHECK_ARG(d->split_k >= 1, "gemm: split_k must be >= 1");
    LD_CHECK_ARG(d->split_k == 1 || (d->accumulate == 2 && d->d_dtype == LD_F32 && d->act == LD_ACT_NONE && !d->aux && !d->R && !d->col_bias),
                 "gemm: split_k > 1 needs accumulate=2, fp32 output and a linear epilogue");
    LD_CHECK_ARG(d->accumulate != 2 || d->d_dtype == LD_F32, "gemm: atomic accumulation needs fp32 output");
    LD_CHECK_ARG(d->d_dtype == LD_F32 || d->d_dtype == LD_BF16, "gemm: bad d_dtype %d", d->d_dtype);
    LD_CHECK_ARG(d->block_n == 0 || d->block_n == 64 || d->block_n == 128 || d->block_n == 256, "gemm: block_n must be 0/64/128/256");
    LD_CHECK_ARG(!d->softmax || (d->N <= 256 && d->d_dtype == LD_BF16 && d->accumulate == 0 && d->split_k == 1 && !d->aux && !d->R &&
                                 !d->col_scale && !d->col_bias && d->act == LD_ACT_NONE && d->ldd % 8 == 0 && d->d_sb1 % 8 == 0 &&
                                 d->d_sb2 % 8 == 0 && ((uintptr_t)d->D & 15) == 0),
                 "gemm: fused softmax epilogue needs N <= 256, aligned bf16 store-only output and no other epilogue terms");

    KParams p{};
    p.M = d->M; p.N = d->N; p.K = d->K; p.nb1 = d->nb1; p.nb2 = d->nb2;
    p.act = d->act; p.accumulate = d->accumulate; p.split_k = d->split_k;
    p.d_dtype = d->d_dtype; p.r_dtype = d->r_dtype;
    p.a_mn = d->A.mn_major ? 1 : 0; p.b_mn = d->B.mn_major ? 1 : 0;
    p.alpha = d->alpha; p.post_gain = d->post_gain;
    p.D = d->D; p.ldd = d->ldd; p.d_sb1 = d->d_sb1; p.d_sb2 = d->d_sb2;
    p.aux = d->aux;
    p.R = d->R; p.ldr = d->ldr; p.r_sb1 = d->r_sb1; p.r_sb2 = d->r_sb2;
    p.cs = d->col_scale; p.cb = d->col_bias; p.col_sb1 = d->col_sb1; p.col_sb2 = d->col_sb2;
    p.alpha_dev = d->alpha_dev;
    p.softmax = d->softmax ? 1 : 0; p.causal = d->causal ? 1 : 0; p.mask_value = d->mask_value; p.key_mask = d->key_mask;
    if (cg) {
        p.cv_mode = cg->mode; p.cv_P = cg->Ho * cg->Wo; p.cv_Wo = cg->Wo; p.cv_stride = cg->stride; p.cv_pad = cg->pad;
        p.cv_KW = cg->KW; p.cv_C = cg->C;
        if (cg->mode == 1) p.a_mn = 0; else p.b_mn = 1;
        p.k32 = (cg->mode == 1 && cg->C == 32) ? 1 : 0;
    }

    const int sms = sm_count();
    const int cta_limit = cta_limit_for(stream);       // persistent grid cap of this stream's lane (default: every SM)
    // CTA pairs (cta_group::2) for problems with enough 256 x 256 tiles to fill the chip; LD_GEMM_2SM=0 turns them off.
    // Each CTA of a pair loads half of the B tile, so the L2 -> shared-memory fill per MMA drops by a quarter: measured on the
    // full training iteration 96.3 -> 93.5 ms per step, and 1213 -> 1249 TFLOP/s on the 36864 x 3072 x 768 GEMM alone
    // (profiles/r1_bench_n1_gemm_2sm.json vs r1_bench_n1.json, same box, back to back).
    static const int env_2sm = [] { const char* e = getenv("LD_GEMM_2SM"); return e ? atoi(e) : 1; }();
    const long nb_ = (long)d->nb1 * d->nb2;
    // threshold 3/8 of the SMs in pairs (55 of 74): the LM-decoder weight gradients (768 x 768 and 768 x 3072 outputs, split-K 8 / 2)
    // have 72 pair tiles — as 288 single-CTA 128 x 128 tiles they ran fill-bound at 480-670 TFLOP/s (profiles/r2_gemm_time_by_shape.txt)
    // (with short reductions the pair kernel's longer prologue costs 1-2 us, so the lower threshold applies from 32 K blocks per split)
    const long k_blocks_per_split = ceil_div(ceil_div(d->K, BK), d->split_k);
    const long pair_min = k_blocks_per_split >= 32 ? (sms * 3 / 8) : (sms / 2);
    const bool two_sm = env_2sm && !p.k32 && d->block_n != 128 && d->block_n != 64 && d->N > 128 && d->M >= 256 && !d->softmax &&
                        nb_ * ceil_div(d->M, 256) * ceil_div(d->N, 256) * d->split_k >= pair_min;
    p.tile_m = two_sm ? 256 : BM;
    p.m_tiles = ceil_div(p.M, p.tile_m);
    const long nb = (long)p.nb1 * p.nb2;
    int bn = d->block_n;
    if (d->softmax) bn = d->N > 128 ? 256 : 128;      // the whole key axis in one tile
    if (two_sm) bn = 256;
    if (bn == 0) {
        const long tiles256 = nb * p.m_tiles * ceil_div(p.N, 256) * p.split_k;
        bn = (p.N > 128 && tiles256 >= sms) ? 256 : 128;
        if (p.N <= 64) bn = 64;          // 64-channel layers (ResNet layer1, StyleGAN2 128^2): half the B fill and MMA time of a 128-wide tile
    }
    p.bn = bn;
    p.n_tiles = ceil_div(p.N, bn);
    p.kb_total = ceil_div(p.K, p.k32 ? 32 : BK);
    if (p.split_k > p.kb_total) p.split_k = p.kb_total;
    p.kb_per_split = ceil_div(p.kb_total, p.split_k);
    p.split_k = ceil_div(p.kb_total, p.kb_per_split);     // no empty splits
    const long total = nb * p.m_tiles * p.n_tiles * p.split_k;
    LD_CHECK_ARG(total < (1L << 30), "gemm: too many tiles");
    p.total_tiles = (int)total;
    auto magic = [total](int d) -> uint32_t {          // exact for every t < total iff t_max * (m*d - 2^32) < 2^32
        if (d <= 1) return 0u;
        const uint64_t m = ((1ull << 32) + (uint64_t)d - 1) / (uint64_t)d;
        const uint64_t e = m * (uint64_t)d - (1ull << 32);
        return (m < (1ull << 32) && (uint64_t)total * e < (1ull << 32)) ? (uint32_t)m : 0u;
    };
    p.mg_split = magic(p.split_k); p.mg_n = magic(p.n_tiles); p.mg_m = magic(p.m_tiles); p.mg_b2 = magic(p.nb2);

    // vectorised epilogue needs 16-byte aligned rows for D / aux / R
    const int d_es = p.d_dtype == LD_BF16 ? 2 : 4;
    auto aligned = [](const void* ptr, long ld, long s1, long s2, int es) {
        const long q = 16 / es;
        return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ld % q == 0 && s1 % q == 0 && s2 % q == 0;
    };
    p.vec_ok = aligned(p.D, p.ldd, p.d_sb1, p.d_sb2, d_es) ? 1 : 0;
    if (p.aux && !aligned(p.aux, p.ldd, p.d_sb1, p.d_sb2, d_es)) p.vec_ok = 0;
    if (p.R && !aligned(p.R, p.ldr, p.r_sb1, p.r_sb2, p.r_dtype == LD_BF16 ? 2 : 4)) p.vec_ok = 0;

    // fp32 read-modify-write (weight gradients accumulated straight into .grad) = the fp32-residual epilogue with R aliased to D:
    // every lane of a warp reads its row's chunk before any lane stores (the stores go through the warp's transpose buffer after a
    // __syncwarp), and warps own disjoint rows / columns — so the vectorised fast path serves it (the generic path took ~2x as long
    // on the DETR-sized weight gradients, 230 launches per iteration).
    if (p.accumulate == 1 && p.d_dtype == LD_F32 && !p.R && !p.aux && p.act == LD_ACT_NONE && p.post_gain == 1.0f && p.vec_ok) {
        p.R = p.D; p.r_dtype = LD_F32; p.ldr = p.ldd; p.r_sb1 = p.d_sb1; p.r_sb2 = p.d_sb2;
        p.accumulate = 0;
    }
    // fast epilogue: vector stores, plain store, no aux; activations only with bf16 output and non-fp32 residual
    p.fast = (p.vec_ok && p.accumulate == 0 && !p.aux &&
              (p.act == LD_ACT_NONE || (p.d_dtype == LD_BF16 && !(p.R && p.r_dtype == LD_F32)))) ? 1 : 0;
    // the fast epilogue folds post_gain into the staged scale / bias: needs gain > 0 for relu / lrelu and gain == 1 for GELU
    if ((p.act == LD_ACT_RELU || p.act == LD_ACT_LRELU) && !(p.post_gain > 0.0f)) p.fast = 0;
    if (p.act == LD_ACT_GELU && p.post_gain != 1.0f) p.fast = 0;

    alignas(64) CUtensorMap tmA, tmB;
    int e = (cg && cg->mode == 1) ? make_conv_map(&tmA, *cg, BM) : make_operand_map(&tmA, d->A, p.M, p.K, p.nb1, p.nb2, BM);
    if (e) return e;
    if (p.k32) LD_CHECK_ARG(!p.b_mn && p.split_k == 1, "conv gemm: 32-channel images need a K-major weight matrix and split_k = 1");
    e = (cg && cg->mode == 2) ? make_conv_map(&tmB, *cg, BK) : make_operand_map(&tmB, d->B, p.N, p.K, p.nb1, p.nb2, two_sm ? bn / 2 : bn, p.k32 != 0);
    if (e) return e;

    static bool attr_set = false;
    if (!attr_set) {
        int s1 = cuda_status(cudaFuncSetAttribute(gemm_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES), "gemm: set smem attr");
        if (s1) return s1;
        s1 = cuda_status(cudaFuncSetAttribute(gemm_bf16_2sm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES), "gemm: set smem attr (2sm)");
        if (s1) return s1;
        attr_set = true;
    }
    // Programmatic dependent launch, opt-in with LD_PDL=1 (the kernel's prologue overlaps the tail of the previous kernel on the stream).
    // Measured (profiles/r2_small_gemm_pdl*.txt): a chain of dependent small GEMMs goes from 5.9 to 5.2 us per launch; the training step does not
    // move (88.9 vs 89.3 ms on one box: its GEMMs alternate with kernels that carry no trigger), so it stays off by default.
    static const int env_pdl = [] { const char* e = getenv("LD_PDL"); return e ? atoi(e) : 0; }();
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = (cudaStream_t)stream;
    cfg.attrs = attr;
    cfg.numAttrs = env_pdl ? 1 : 0;
    cudaError_t le;
    if (two_sm) {
        const int cap2 = cta_limit / 2 > 0 ? cta_limit / 2 : 1;
        const int pairs = (int)(total < cap2 ? total : cap2);
        cfg.gridDim = dim3(2 * pairs);
        le = cudaLaunchKernelEx(&cfg, gemm_bf16_2sm_kernel, tmA, tmB, p);
    } else {
        const int grid = (int)(total < cta_limit ? total : cta_limit);
        cfg.gridDim = dim3(grid);
        le = cudaLaunchKernelEx(&cfg, gemm_bf16_kernel, tmA, tmB, p);
    }
    if (le != cudaSuccess) return cuda_status(le, "gemm launch");
    count_launch();
    LD_LAUNCH_CHECK("gemm launch");
    return 0;
}
}  // namespace
