// ROUND-1 KERNEL, kept only as the A/B baseline of attention_fwd_sm100.cu (LD_ATTN_V1=1): one tile per CTA, phases back to back.
// Fused multi-head attention forward for sm_100a:  O = softmax(Q K^T * scale + mask) V  per (batch, head, 128-query tile)
// with <= 256 keys and head_dim <= 192 — the shapes of the LayoutDETR path: BERT text encoder / decoder (T = 256,
// head_dim 192, additive -10000 key mask, causal for the decoder; training/med.py:146-228) and DETR self / cross
// attention (head_dim 32, 9 / 10 / 64 keys, -inf key-padding mask; training/detr_transformer.py:208,273,277).
//
// One CTA per tile, warp-specialised:
//   warp 0   TMA producer: Q tile once, then K blocks and V blocks (64 keys each) through one 4-slot smem ring
//   warp 1   MMA issuer:   S[128 x keys] = Q K^T  (tcgen05.mma, N = 64 per key block, fp32 in TMEM columns 0..255)
//                          O[128 x d]   += P_j V_j (A = un-normalised bf16 probabilities in smem, B = V block MN-major,
//                                                   fp32 in TMEM columns 256..447)
//   warp 2   TMEM allocator
//   warps 4..11  softmax + epilogue: thread = query row (TMEM lane), two warps per lane quadrant split the key axis;
//            sweep 1: row max (exchanged through smem), sweep 2: e = exp(s - max) -> bf16 into the swizzled smem A tile,
//            row sums in fp32; O is normalised by 1/sum in the epilogue (flash-attention style) and written with
//            coalesced 64-byte row segments.  Optional third sweep writes normalised P to HBM for the backward pass.
// The fp32 score matrix and (for inference) the probabilities never touch HBM.
#include <cstdlib>
#include "common.cuh"
#include "runtime.h"

namespace {
using namespace ld;

constexpr int AT_THREADS = 384;
constexpr int AT_RING = 4;
constexpr int AT_Q_BYTES = 3 * 16384;          // 128 x 192 bf16
constexpr int AT_P_BYTES = 4 * 16384;          // 128 x 256 bf16
constexpr int AT_SLOT_BYTES = 3 * 8192;        // 64 keys x 192 bf16
constexpr int AT_MASK_BYTES = 1024;            // 256 floats
constexpr int AT_XCH_BYTES = 2 * 2 * 128 * 4;  // [max|sum][half][row]
constexpr int AT_BAR_BYTES = 256;
constexpr int AT_SMEM = AT_Q_BYTES + AT_P_BYTES + AT_RING * AT_SLOT_BYTES + AT_MASK_BYTES + AT_XCH_BYTES + AT_BAR_BYTES + 1024;
constexpr uint32_t TM_S = 0, TM_O = 256;

struct AttnParams {
    int B, H, Lq, Lk, d;
    int dch, nkv, q_tiles;
    float scale, mask_value;
    int causal;
    const uint8_t* key_mask;
    __nv_bfloat16* O; long ldo;
    __nv_bfloat16* P; long ldp;
};

__global__ void __launch_bounds__(AT_THREADS, 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* q_s = smem;
    uint8_t* p_s = q_s + AT_Q_BYTES;
    uint8_t* ring = p_s + AT_P_BYTES;
    float* mask_s = reinterpret_cast<float*>(ring + AT_RING * AT_SLOT_BYTES);
    float* xch = mask_s + 256;                                   // [2 kinds][2 halves][128 rows]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xch) + AT_XCH_BYTES);
    uint64_t* full_bar = bars;            // [AT_RING]
    uint64_t* empty_bar = bars + AT_RING; // [AT_RING]
    uint64_t* q_bar = bars + 2 * AT_RING;
    uint64_t* s_bar = q_bar + 1;
    uint64_t* p_bar = q_bar + 2;
    uint64_t* o_bar = q_bar + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_bar + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int t = blockIdx.x;
    const int qt = t % p.q_tiles; t /= p.q_tiles;
    const int h = t % p.H;
    const int b = t / p.H;
    const int m0 = qt * 128;

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < AT_RING; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(q_bar, 1); mbar_init(s_bar, 1); mbar_init(p_bar, 8); mbar_init(o_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    if (warp >= 4) {                                             // additive key mask of this batch element
        const int c = threadIdx.x - 128;
        const uint8_t* km = p.key_mask ? p.key_mask + (long)b * p.Lk : nullptr;
        mask_s[c] = (c < p.Lk) ? ((km && km[c]) ? p.mask_value : 0.0f) : -INFINITY;    // keys beyond Lk never attend
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int slot_bytes = p.dch * 8192;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_bar, p.dch * 16384);
            for (int c = 0; c < p.dch; ++c) tma_load_4d(q_s + c * 16384, &tmQ, q_bar, c * 64, m0, h, b);
            int stage = 0; uint32_t phase = 0;
            for (int pass = 0; pass < 2; ++pass) {               // pass 0: K blocks, pass 1: V blocks
                const CUtensorMap* tm = pass == 0 ? &tmK : &tmV;
                for (int j = 0; j < p.nkv; ++j) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full_bar[stage], slot_bytes);
                    uint8_t* dst = ring + stage * AT_SLOT_BYTES;
                    for (int c = 0; c < p.dch; ++c) tma_load_4d(dst + c * 8192, tm, &full_bar[stage], c * 64, j * 64, h, b);
                    if (++stage == AT_RING) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            const uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);
            const uint32_t idesc_o = make_idesc_bf16(128, p.dch * 64, 0, 1);
            mbar_wait(q_bar, 0);
            tc_fence_after();
            const uint32_t qa = smem_u32(q_s);
            for (int j = 0; j < p.nkv; ++j) {                    // S[:, 64j : 64j+64] = Q K_j^T
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t kb = smem_u32(ring + stage * AT_SLOT_BYTES);
                for (int c = 0; c < p.dch; ++c) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint64_t da = make_smem_desc(qa + c * 16384 + kk * 32, 16, 1024);
                        const uint64_t db = make_smem_desc(kb + c * 8192 + kk * 32, 16, 1024);
                        umma_bf16_ss(tmem_base + TM_S + 64 * j, da, db, idesc_s, (c > 0 || kk > 0) ? 1u : 0u);
                    }
                }
                umma_commit(&empty_bar[stage]);
                if (++stage == AT_RING) { stage = 0; phase ^= 1; }
            }
            umma_commit(s_bar);                                  // scores complete -> softmax warps
            mbar_wait(p_bar, 0);                                 // probabilities are in smem
            tc_fence_after();
            const uint32_t pa = smem_u32(p_s);
            for (int j = 0; j < p.nkv; ++j) {                    // O += P[:, 64j : 64j+64] V_j
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t vb = smem_u32(ring + stage * AT_SLOT_BYTES);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const uint64_t da = make_smem_desc(pa + j * 16384 + kk * 32, 16, 1024);
                    const uint64_t db = make_smem_desc(vb + kk * 2048, 8192, 1024);
                    umma_bf16_ss(tmem_base + TM_O, da, db, idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[stage]);
                if (++stage == AT_RING) { stage = 0; phase ^= 1; }
            }
            umma_commit(o_bar);
        }
    } else if (warp >= 4) {
        const int e = warp - 4, q = e & 3, half = e >> 2;
        const int r = q * 32 + lane;                             // query row within the tile == TMEM lane
        const int row = m0 + r;
        const bool row_ok = row < p.Lq;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        // key blocks of this warp's half
        const int jb = half == 0 ? 0 : (p.nkv + 1) / 2;
        const int je = half == 0 ? (p.nkv + 1) / 2 : p.nkv;
        const uint32_t mask_a = smem_u32(mask_s);          // shared-space addresses (LDS / STS instead of generic LD / ST)
        const uint32_t xmax_a = smem_u32(xch);             // [2][128]
        const uint32_t xsum_a = xmax_a + 1024;             // [2][128]
        const uint32_t p_a = smem_u32(p_s);
        mbar_wait(s_bar, 0);
        tc_fence_after();
        // ---- sweep 1: row max over this half's keys
        float mx = -INFINITY;
        for (int j = jb; j < je; ++j) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                uint32_t v[32];
                tmem_ld_x32(lane_addr + TM_S + 64 * j + 32 * hh, v);
                tmem_ld_wait();
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                    const int c4 = 64 * j + 32 * hh + 4 * i4;
                    const float4 m4 = lds_f4(mask_a + c4 * 4);
                    const float mm[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int c = c4 + u;
                        float s = fmaf(__uint_as_float(v[4 * i4 + u]), p.scale, mm[u]);
                        if (p.causal && c > row) s += p.mask_value;
                        mx = fmaxf(mx, s);
                    }
                }
            }
        }
        sts_f32(xmax_a + (half * 128 + r) * 4, mx);
        asm volatile("bar.sync 2, 256;" ::: "memory");
        mx = fmaxf(lds_f32(xmax_a + r * 4), lds_f32(xmax_a + (128 + r) * 4));
        // ---- sweep 2: e = exp(s - max) -> bf16 into the swizzled A tile; fp32 row sum
        float sum = 0.f;
        for (int j = jb; j < je; ++j) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                uint32_t v[32];
                tmem_ld_x32(lane_addr + TM_S + 64 * j + 32 * hh, v);
                tmem_ld_wait();
                float ev[32];
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                    const int c4 = 64 * j + 32 * hh + 4 * i4;
                    const float4 m4 = lds_f4(mask_a + c4 * 4);
                    const float mm[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int c = c4 + u;
                        float s = fmaf(__uint_as_float(v[4 * i4 + u]), p.scale, mm[u]);
                        if (p.causal && c > row) s += p.mask_value;
                        ev[4 * i4 + u] = __expf(s - mx);
                        sum += ev[4 * i4 + u];
                    }
                }
                const uint32_t prow_a = p_a + j * 16384 + r * 128;
#pragma unroll
                for (int g = 0; g < 4; ++g) {                    // four 16-byte pieces (8 keys each) of this 32-key run
                    uint4 o;
                    o.x = pack_bf16x2(ev[8 * g + 0], ev[8 * g + 1]); o.y = pack_bf16x2(ev[8 * g + 2], ev[8 * g + 3]);
                    o.z = pack_bf16x2(ev[8 * g + 4], ev[8 * g + 5]); o.w = pack_bf16x2(ev[8 * g + 6], ev[8 * g + 7]);
                    const int piece = hh * 4 + g;
                    sts_u4(prow_a + ((piece ^ (r & 7)) << 4), o);
                }
            }
        }
        sts_f32(xsum_a + (half * 128 + r) * 4, sum);
        fence_proxy_async_smem();                                // generic-proxy smem writes -> visible to tcgen05.mma
        tc_fence_before();
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (lane == 0) mbar_arrive(p_bar);
        const float inv = __fdividef(1.0f, lds_f32(xsum_a + r * 4) + lds_f32(xsum_a + (128 + r) * 4));
        // ---- optional sweep 3: normalised probabilities to HBM (needed by the backward pass)
        if (p.P != nullptr) {
            __nv_bfloat16* prow_g = p.P + ((long)(b * p.H + h) * p.Lq + row) * p.ldp;
            const int n_pad = (p.Lk + 7) & ~7;
            for (int j = jb; j < je; ++j) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    uint32_t v[32];
                    tmem_ld_x32(lane_addr + TM_S + 64 * j + 32 * hh, v);
                    tmem_ld_wait();
                    if (!row_ok) continue;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int c0 = 64 * j + 32 * hh + 8 * g;
                        if (c0 >= n_pad) continue;
                        float pe[8];
                        const float4 ma = lds_f4(mask_a + c0 * 4), mb = lds_f4(mask_a + c0 * 4 + 16);
                        const float mm[8] = {ma.x, ma.y, ma.z, ma.w, mb.x, mb.y, mb.z, mb.w};
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int c = c0 + i;
                            float s = fmaf(__uint_as_float(v[8 * g + i]), p.scale, mm[i]);
                            if (p.causal && c > row) s += p.mask_value;
                            pe[i] = __expf(s - mx) * inv;
                        }
                        uint4 o;
                        o.x = pack_bf16x2(pe[0], pe[1]); o.y = pack_bf16x2(pe[2], pe[3]);
                        o.z = pack_bf16x2(pe[4], pe[5]); o.w = pack_bf16x2(pe[6], pe[7]);
                        *reinterpret_cast<uint4*>(prow_g + c0) = o;
                    }
                }
            }
        }
        // ---- epilogue: O * (1 / sum) -> bf16, coalesced 64-byte row segments through a swizzled smem tile
        mbar_wait(o_bar, 0);
        tc_fence_after();
        const uint32_t stg_a = p_a + e * 2048;                   // P tile is dead once the PV MMAs have retired
        const int sw_w = (lane >> 1) & 3, pc = lane & 3;
        const int cbeg = half * p.dch * 32, cend = cbeg + p.dch * 32;
        __nv_bfloat16* obase = p.O + (long)b * p.Lq * p.ldo + (long)h * p.d;
        for (int c0 = cbeg; c0 < cend; c0 += 32) {
            if (c0 >= p.d) break;                                // warp-uniform: padding columns of head_dim < 64
            uint32_t v[32];
            tmem_ld_x32(lane_addr + TM_O + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint4 o;
                o.x = pack_bf16x2(__uint_as_float(v[8 * g + 0]) * inv, __uint_as_float(v[8 * g + 1]) * inv);
                o.y = pack_bf16x2(__uint_as_float(v[8 * g + 2]) * inv, __uint_as_float(v[8 * g + 3]) * inv);
                o.z = pack_bf16x2(__uint_as_float(v[8 * g + 4]) * inv, __uint_as_float(v[8 * g + 5]) * inv);
                o.w = pack_bf16x2(__uint_as_float(v[8 * g + 6]) * inv, __uint_as_float(v[8 * g + 7]) * inv);
                sts_u4(stg_a + ((lane * 4 + (g ^ sw_w)) << 4), o);
            }
            __syncwarp();
            uint4 val[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int rl = (lane >> 2) + 8 * i;
                val[i] = lds_u4(stg_a + ((rl * 4 + (pc ^ ((rl >> 1) & 3))) << 4));
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int rl = (lane >> 2) + 8 * i;
                const int grow = m0 + q * 32 + rl;
                if (grow < p.Lq && c0 + pc * 8 < p.d)
                    *reinterpret_cast<uint4*>(obase + (long)grow * p.ldo + c0 + pc * 8) = val[i];
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

int make_map(CUtensorMap* tm, const void* ptr, int64_t ld_, int d, int L, int H, int B, uint32_t box_rows) {
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || ld_ % 8 != 0 || d % 8 != 0) {
        set_last_error("attention: operand base / strides must be 16-byte aligned (ld=%lld d=%d)", (long long)ld_, d);
        return LD_ERR_ALIGNMENT;
    }
    const uint64_t dims[4] = {(uint64_t)d, (uint64_t)L, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)ld_ * 2, (uint64_t)d * 2, (uint64_t)L * ld_ * 2};
    return encode_tmap_bf16_4d(tm, ptr, dims, strides, 64, box_rows);
}
}  // namespace

// q / k / v point at column 0 of head 0 inside row-major [B*L, ld] bf16 buffers (head h at columns h*d .. h*d+d).
extern "C" int ld_attention_fwd_v1(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                void* o, int64_t ldo, void* p_out, int64_t ldp, int B, int H, int Lq, int Lk, int d,
                                float scale, const uint8_t* key_mask, int mask_inf, int causal, void* stream) {
    using namespace ld;
    LD_CHECK_ARG(q && k && v && o && B > 0 && H > 0 && Lq > 0 && Lk > 0, "attention_fwd: bad argument");
    LD_CHECK_ARG(Lk <= 256 && d <= 192 && d % 8 == 0, "attention_fwd: needs <= 256 keys and head_dim <= 192 (multiple of 8); got Lk=%d d=%d", Lk, d);
    LD_CHECK_ARG(ldo % 8 == 0 && ((uintptr_t)o & 15) == 0 && (!p_out || (ldp % 8 == 0 && ((uintptr_t)p_out & 15) == 0)),
                 "attention_fwd: output alignment");
    AttnParams p{};
    p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk; p.d = d;
    p.dch = (d + 63) / 64; p.nkv = (Lk + 63) / 64; p.q_tiles = (Lq + 127) / 128;
    p.scale = scale; p.mask_value = mask_inf ? -INFINITY : -10000.0f; p.causal = causal ? 1 : 0;
    p.key_mask = key_mask;
    p.O = (__nv_bfloat16*)o; p.ldo = ldo; p.P = (__nv_bfloat16*)p_out; p.ldp = ldp;
    alignas(64) CUtensorMap tmQ, tmK, tmV;
    int e = make_map(&tmQ, q, ldq, d, Lq, H, B, 128); if (e) return e;
    e = make_map(&tmK, k, ldk, d, Lk, H, B, 64); if (e) return e;
    e = make_map(&tmV, v, ldv, d, Lk, H, B, 64); if (e) return e;
    static bool attr_set = false;
    if (!attr_set) {
        int s = cuda_status(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM), "attention: smem attr");
        if (s) return s;
        attr_set = true;
    }
    const long grid = (long)B * H * p.q_tiles;
    attention_fwd_kernel<<<(unsigned)grid, AT_THREADS, AT_SMEM, (cudaStream_t)stream>>>(tmQ, tmK, tmV, p);
    count_launch();
    LD_LAUNCH_CHECK("attention_fwd");
    return 0;
}
