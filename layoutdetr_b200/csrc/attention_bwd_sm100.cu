// Fused multi-head attention backward for sm_100a (<= 256 keys, head_dim <= 192), the adjoint of attention_fwd_sm100.cu:
//     P  = 2^(s2 - lse2)            s2 = (Q K^T) * scale * log2 e + mask2   (recomputed from the forward's per-row lse2)
//     Pd = dropout(P)               same Philox mask as the forward, regenerated from the element coordinates
//     dA = dropout'(dO V^T)
//     dS = P o (dA - D) * scale     D = rowsum(dO o O)
//     dQ = dS K
// One persistent CTA per SM, one 128-query tile at a time: S and dO V^T accumulate in TMEM (2 x 256 columns), 256 threads
// (two per query row, splitting the key blocks) turn them into Pd and dS in ONE sweep, dS goes to swizzled shared memory as the
// A operand of the dQ MMAs (dQ re-uses S's TMEM columns).  K stays resident for the tile: the same shared-memory image
// [keys x 64 values] is the K-major B operand of Q K^T and the MN-major B operand of dS K.  Pd and dS (bf16) are also written to
// HBM for the two transposed products dV = Pd^T dO and dK = dS^T Q, which remain batched tcgen05 GEMMs (ld_gemm_bf16).
// Replaces, per attention: the forward's P write, the fp32 dP GEMM, ld_softmax_bwd and the dQ GEMM.
// Reference math: autograd of BertSelfAttention.forward (training/med.py:183-215) and F.multi_head_attention_forward.
#include "common.cuh"
#include "runtime.h"

namespace {
using namespace ld;

constexpr int AB_THREADS = 384;
constexpr int AB_X_BYTES = 98304;              // Q tile (48 KB) | dO tile (48 KB); later dS (128 x 256 bf16 = 64 KB) over both
constexpr int AB_K_BYTES = 98304;              // K resident: three units of [256 keys x 64 values]; later the dQ staging tile
constexpr int AB_V_BYTES = 32768;              // one V unit [256 keys x 64 values]
constexpr int AB_MASK_BYTES = 1024;
constexpr int AB_BAR_BYTES = 256;
constexpr int AB_SMEM = AB_X_BYTES + AB_K_BYTES + AB_V_BYTES + AB_MASK_BYTES + AB_BAR_BYTES;
static_assert(AB_SMEM <= 227 * 1024, "shared memory budget");
constexpr float LOG2E = 1.4426950408889634f;
constexpr uint32_t TM_S = 0, TM_DP = 256;

struct AbParams {
    int B, H, Lq, Lk, d;
    int dch, nkb, ncols, q_tiles, total_tiles;
    float scale, scale2, mask2;
    int causal;
    const uint8_t* key_mask;
    const __nv_bfloat16* O; long ldo;          // forward output, [B*Lq, ldo], head h at columns h*d
    const float* lse;                          // [B*H, Lq] (log2 domain)
    __nv_bfloat16* Pd; long ldp;               // [B*H, Lq, ldp]
    const uint32_t* rng; uint32_t site, thresh16; float drop_scale;
};

struct TileId { int b, h, m0; };
__device__ __forceinline__ TileId decode(const AbParams& p, int t) {
    TileId id;
    const int qt = t % p.q_tiles; t /= p.q_tiles;
    id.h = t % p.H; id.b = t / p.H; id.m0 = qt * 128;
    return id;
}

template <bool CAUSAL, bool DROPOUT>
__global__ void __launch_bounds__(AB_THREADS, 1)
attention_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                     const __grid_constant__ CUtensorMap tmdQ, const __grid_constant__ CUtensorMap tmdS, const AbParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* x_s = smem;
    uint8_t* k_s = smem + AB_X_BYTES;
    uint8_t* v_s = k_s + AB_K_BYTES;
    float* mask_s = reinterpret_cast<float*>(v_s + AB_V_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(mask_s) + AB_MASK_BYTES);
    uint64_t* x_full = bars;              // Q and dO tiles landed
    uint64_t* k_full = bars + 1;          // [3] K unit ch landed
    uint64_t* v_full = bars + 4;          // V unit landed in the slot
    uint64_t* v_empty = bars + 5;         // the dP MMAs of the slot's unit retired
    uint64_t* sdp_full = bars + 6;        // S and dO V^T complete in TMEM
    uint64_t* ds_ready = bars + 7;        // dS in shared memory, S / dP columns free (8 warps arrive)
    uint64_t* dq_full = bars + 8;         // dQ complete in TMEM
    uint64_t* bufs_free = bars + 9;       // every buffer of the tile may be reloaded
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_local = (p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0u) asm volatile("trap;");
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO);
        tma_prefetch_desc(&tmdQ); tma_prefetch_desc(&tmdS);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(x_full, 1);
        for (int c = 0; c < 3; ++c) mbar_init(&k_full[c], 1);
        mbar_init(v_full, 1); mbar_init(v_empty, 1); mbar_init(sdp_full, 1); mbar_init(ds_ready, 8);
        mbar_init(dq_full, 1); mbar_init(bufs_free, 1);
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            uint32_t vcount = 0;                                  // V units issued so far (slot phase bookkeeping)
            for (int i = 0; i < n_local; ++i) {
                const TileId id = decode(p, (int)blockIdx.x + i * (int)gridDim.x);
                if (i >= 1) mbar_wait(bufs_free, (uint32_t)((i - 1) & 1));
                mbar_arrive_expect_tx(x_full, (uint32_t)p.dch * 32768u);
                for (int ch = 0; ch < p.dch; ++ch) {
                    tma_load_4d(x_s + ch * 16384, &tmQ, x_full, ch * 64, id.m0, id.h, id.b);
                    tma_load_4d(x_s + 49152 + ch * 16384, &tmdO, x_full, ch * 64, id.m0, id.h, id.b);
                }
                for (int ch = 0; ch < p.dch; ++ch) {
                    mbar_arrive_expect_tx(&k_full[ch], (uint32_t)p.ncols * 128u);
                    tma_load_4d(k_s + ch * 32768, &tmK, &k_full[ch], ch * 64, 0, id.h, id.b);
                }
                for (int ch = 0; ch < p.dch; ++ch, ++vcount) {
                    if (vcount >= 1) mbar_wait(v_empty, (vcount - 1) & 1u);
                    mbar_arrive_expect_tx(v_full, (uint32_t)p.ncols * 128u);
                    tma_load_4d(v_s, &tmV, v_full, ch * 64, 0, id.h, id.b);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc_s = make_idesc_bf16(128, p.ncols, 0, 0);
            const uint32_t idesc_q = make_idesc_bf16(128, 64, 0, 1);
            uint32_t vcount = 0;
            for (int i = 0; i < n_local; ++i) {
                const uint32_t par = (uint32_t)(i & 1);
                mbar_wait(x_full, par);
                tc_fence_after();
                const uint32_t qa = smem_u32(x_s), doa = qa + 49152, ka = smem_u32(k_s), va = smem_u32(v_s);
                for (int ch = 0; ch < p.dch; ++ch) {             // S += Q_ch K_ch^T
                    mbar_wait(&k_full[ch], par);
                    tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_bf16_ss(tmem_base + TM_S, make_smem_desc(qa + ch * 16384 + kk * 32, 16, 1024),
                                     make_smem_desc(ka + ch * 32768 + kk * 32, 16, 1024), idesc_s, (ch > 0 || kk > 0) ? 1u : 0u);
                }
                for (int ch = 0; ch < p.dch; ++ch, ++vcount) {   // dP += dO_ch V_ch^T
                    mbar_wait(v_full, vcount & 1u);
                    tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_bf16_ss(tmem_base + TM_DP, make_smem_desc(doa + ch * 16384 + kk * 32, 16, 1024),
                                     make_smem_desc(va + kk * 32, 16, 1024), idesc_s, (ch > 0 || kk > 0) ? 1u : 0u);
                    umma_commit(v_empty);
                }
                umma_commit(sdp_full);
                mbar_wait(ds_ready, par);
                tc_fence_after();
                for (int ch = 0; ch < p.dch; ++ch) {             // dQ[:, 64ch : 64ch+64] = dS K[:, 64ch : 64ch+64]
                    for (int j = 0; j < p.nkb; ++j) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma_bf16_ss(tmem_base + TM_S + 64 * ch, make_smem_desc(qa + j * 16384 + kk * 32, 16, 1024),
                                         make_smem_desc(ka + ch * 32768 + j * 8192 + kk * 2048, 8192, 1024), idesc_q,
                                         (j > 0 || kk > 0) ? 1u : 0u);
                    }
                }
                umma_commit(dq_full);
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ dS sweep + epilogue (8 warps, two threads per row)
        const int e = warp - 4, q = warp & 3, half = e >> 2;
        const int r = q * 32 + lane;
        const int et = threadIdx.x - 128;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t mask_a = smem_u32(mask_s);
        const uint32_t x_a = smem_u32(x_s), k_a = smem_u32(k_s);
        DropoutRng rng;
        rng.init(p.rng, p.site, DROPOUT ? p.thresh16 : 0u, p.drop_scale);
        const int jb = half == 0 ? 0 : (p.nkb + 1) / 2;          // key blocks of this thread
        const int je = half == 0 ? (p.nkb + 1) / 2 : p.nkb;
        for (int i = 0; i < n_local; ++i) {
            const uint32_t par = (uint32_t)(i & 1);
            const TileId id = decode(p, (int)blockIdx.x + i * (int)gridDim.x);
            const int row = id.m0 + r;
            const bool row_ok = row < p.Lq;
            const long bh = (long)id.b * p.H + id.h;
            {
                const uint8_t* km = p.key_mask ? p.key_mask + (long)id.b * p.Lk : nullptr;
                sts_f32(mask_a + et * 4, (et < p.Lk) ? ((km && km[et]) ? p.mask2 : 0.0f) : -INFINITY);
            }
            const float lse2 = row_ok ? __ldg(p.lse + bh * p.Lq + row) : 0.0f;
            mbar_wait(x_full, par);                              // dO tile is in shared memory
            // ---- D = rowsum(dO o O): dO from the swizzled smem tile, O from global (this thread's row)
            float D = 0.f;
            if (row_ok) {
                const __nv_bfloat16* orow = p.O + ((long)id.b * p.Lq + row) * p.ldo + (long)id.h * p.d;
                for (int c8 = 0; c8 < p.d; c8 += 8) {
                    const uint4 ov = __ldg(reinterpret_cast<const uint4*>(orow + c8));
                    const uint4 dv = lds_u4(x_a + 49152 + (c8 >> 6) * 16384 + r * 128 + ((((c8 >> 3) & 7) ^ (r & 7)) << 4));
                    const uint32_t ow[4] = {ov.x, ov.y, ov.z, ov.w}, dw[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        float a0, a1, b0, b1;
                        unpack_bf16x2(ow[u], a0, a1); unpack_bf16x2(dw[u], b0, b1);
                        D = fmaf(a0, b0, D); D = fmaf(a1, b1, D);
                    }
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");       // key mask staged
            mbar_wait(sdp_full, par);
            tc_fence_after();
            // ---- one sweep: P, Pd -> HBM, dS -> swizzled A tile
            const uint64_t grow = (uint64_t)(bh * p.Lq + row) * 32u;
            __nv_bfloat16* prow_g = p.Pd + (bh * p.Lq + row) * p.ldp;
            const int n_pad = (p.Lk + 7) & ~7;
            for (int j = jb; j < je; ++j) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    uint32_t sv[32], dv[32];
                    tmem_ld_x32(lane_addr + TM_S + 64 * j + 32 * hh, sv);
                    tmem_ld_x32(lane_addr + TM_DP + 64 * j + 32 * hh, dv);
                    tmem_ld_wait();
                    const uint32_t dsrow_a = x_a + j * 16384 + r * 128;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int c8 = 64 * j + 32 * hh + 8 * g;
                        const float4 ma = lds_f4(mask_a + c8 * 4), mb = lds_f4(mask_a + c8 * 4 + 16);
                        const float mm[8] = {ma.x, ma.y, ma.z, ma.w, mb.x, mb.y, mb.z, mb.w};
                        uint32_t keep = 0xFFu;
                        if (DROPOUT) keep = rng.keep8(grow + (uint32_t)(c8 >> 3));
                        float pd[8], ds[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            float s = fmaf(__uint_as_float(sv[8 * g + u]), p.scale2, mm[u]);
                            if (CAUSAL && c8 + u > row) s += p.mask2;
                            const float pr = ex2_approx(s - lse2);
                            float da = __uint_as_float(dv[8 * g + u]);
                            float pk = pr;
                            if (DROPOUT) {
                                const bool kp = (keep >> u) & 1u;
                                pk = kp ? pr * rng.scale : 0.0f;
                                da = kp ? da * rng.scale : 0.0f;
                            }
                            pd[u] = pk;
                            ds[u] = pr * (da - D) * p.scale;
                        }
                        uint4 o;
                        o.x = pack_bf16x2(ds[0], ds[1]); o.y = pack_bf16x2(ds[2], ds[3]);
                        o.z = pack_bf16x2(ds[4], ds[5]); o.w = pack_bf16x2(ds[6], ds[7]);
                        sts_u4(dsrow_a + (((hh * 4 + g) ^ (r & 7)) << 4), o);
                        if (row_ok && c8 < n_pad) {
                            uint4 w;
                            w.x = pack_bf16x2(pd[0], pd[1]); w.y = pack_bf16x2(pd[2], pd[3]);
                            w.z = pack_bf16x2(pd[4], pd[5]); w.w = pack_bf16x2(pd[6], pd[7]);
                            *reinterpret_cast<uint4*>(prow_g + c8) = w;
                        }
                    }
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(ds_ready);
            // ---- epilogue: dS tile -> HBM (TMA, once the dQ MMAs have read it); dQ -> bf16 -> smem (over K) -> TMA store
            mbar_wait(dq_full, par);
            tc_fence_after();
            if (et == 0) {
                for (int j = 0; j < p.nkb; ++j) tma_store_4d(&tmdS, x_a + j * 16384, j * 64, id.m0, (int)bh, 0);
                tma_store_commit();
            }
            const int cb = half == 0 ? 0 : (p.dch + 1) / 2, ce = half == 0 ? (p.dch + 1) / 2 : p.dch;
            for (int ch = cb; ch < ce; ++ch) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    uint32_t v[32];
                    tmem_ld_x32(lane_addr + TM_S + 64 * ch + 32 * hh, v);
                    tmem_ld_wait();
                    const uint32_t orow_a = k_a + ch * 16384 + r * 128;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint4 o;
                        o.x = pack_bf16x2(__uint_as_float(v[8 * g + 0]), __uint_as_float(v[8 * g + 1]));
                        o.y = pack_bf16x2(__uint_as_float(v[8 * g + 2]), __uint_as_float(v[8 * g + 3]));
                        o.z = pack_bf16x2(__uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5]));
                        o.w = pack_bf16x2(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7]));
                        sts_u4(orow_a + (((hh * 4 + g) ^ (r & 7)) << 4), o);
                    }
                }
            }
            tc_fence_before();
            fence_proxy_async_smem();
            asm volatile("bar.sync 1, 256;" ::: "memory");       // dQ staged by all 256 threads, TMEM drained
            if (et == 0) {
                for (int ch = 0; ch < p.dch; ++ch) tma_store_4d(&tmdQ, k_a + ch * 16384, ch * 64, id.m0, id.h, id.b);
                tma_store_commit();
                tma_store_wait_read();
                tc_fence_before();
                mbar_arrive(bufs_free);
            }
        }
        if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

int make_map(CUtensorMap* tm, const void* ptr, int64_t ld_, int d, int L, int H, int B, uint32_t box_rows) {
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || ld_ % 8 != 0 || d % 8 != 0) {
        set_last_error("attention_bwd: operand base / strides must be 16-byte aligned (ld=%lld d=%d)", (long long)ld_, d);
        return LD_ERR_ALIGNMENT;
    }
    const uint64_t dims[4] = {(uint64_t)d, (uint64_t)L, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)ld_ * 2, (uint64_t)d * 2, (uint64_t)L * ld_ * 2};
    return encode_tmap_bf16_4d(tm, ptr, dims, strides, 64, box_rows);
}

template <bool CAUSAL, bool DROPOUT>
int launch(const CUtensorMap* tm, const AbParams& p, int grid, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        int s = cuda_status(cudaFuncSetAttribute(attention_bwd_kernel<CAUSAL, DROPOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM),
                            "attention_bwd: smem attr");
        if (s) return s;
        attr_set = true;
    }
    attention_bwd_kernel<CAUSAL, DROPOUT><<<grid, AB_THREADS, AB_SMEM, stream>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], p);
    return 0;
}
}  // namespace

extern "C" int ld_attention_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                const void* o, const void* d_o, int64_t ldo, const float* lse,
                                void* dq, int64_t lddq, void* pd_out, void* ds_out, int64_t ldp,
                                int B, int H, int Lq, int Lk, int d, float scale, const uint8_t* key_mask, int mask_inf, int causal,
                                float dropout_p, const uint32_t* rng_state, uint32_t rng_site, void* stream) {
    using namespace ld;
    LD_CHECK_ARG(q && k && v && o && d_o && lse && dq && pd_out && ds_out && B > 0 && H > 0 && Lq > 0 && Lk > 0, "attention_bwd: bad argument");
    LD_CHECK_ARG(Lk <= 256 && d <= 192 && d % 8 == 0, "attention_bwd: needs <= 256 keys and head_dim <= 192 (multiple of 8); got Lk=%d d=%d", Lk, d);
    LD_CHECK_ARG(ldp % 8 == 0 && ldp >= ((Lk + 7) & ~7) && ((uintptr_t)pd_out & 15) == 0 && ((uintptr_t)ds_out & 15) == 0 &&
                 ldo % 8 == 0 && ((uintptr_t)o & 15) == 0, "attention_bwd: pd / ds / o alignment");
    LD_CHECK_ARG(dropout_p >= 0.0f && dropout_p < 1.0f && (dropout_p == 0.0f || rng_state), "attention_bwd: dropout arguments");
    AbParams p{};
    p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk; p.d = d;
    p.dch = (d + 63) / 64; p.nkb = (Lk + 63) / 64; p.ncols = p.nkb * 64; p.q_tiles = (Lq + 127) / 128;
    const long total = (long)B * H * p.q_tiles;
    LD_CHECK_ARG(total < (1L << 30) && (long)B * H < (1L << 31), "attention_bwd: too many tiles");
    p.total_tiles = (int)total;
    p.scale = scale; p.scale2 = scale * LOG2E; p.mask2 = mask_inf ? -INFINITY : -10000.0f * LOG2E; p.causal = causal ? 1 : 0;
    p.key_mask = key_mask;
    p.O = (const __nv_bfloat16*)o; p.ldo = ldo; p.lse = lse;
    p.Pd = (__nv_bfloat16*)pd_out; p.ldp = ldp;
    p.rng = rng_state; p.site = rng_site;
    p.thresh16 = dropout_p > 0.0f ? (uint32_t)(dropout_p * 65536.0f + 0.5f) : 0u;
    p.drop_scale = dropout_p > 0.0f ? 65536.0f / (65536.0f - (float)p.thresh16) : 1.0f;
    alignas(64) CUtensorMap tm[6];
    int e = make_map(&tm[0], q, ldq, d, Lq, H, B, 128); if (e) return e;
    e = make_map(&tm[1], k, ldk, d, Lk, H, B, (uint32_t)p.ncols); if (e) return e;
    e = make_map(&tm[2], v, ldv, d, Lk, H, B, (uint32_t)p.ncols); if (e) return e;
    e = make_map(&tm[3], d_o, ldo, d, Lq, H, B, 128); if (e) return e;
    e = make_map(&tm[4], dq, lddq, d, Lq, H, B, 128); if (e) return e;
    {   // dS [B*H, Lq, ldp]: dims (Lkp, Lq, B*H, 1)
        const uint64_t lkp = (uint64_t)((Lk + 7) & ~7);
        const uint64_t dims[4] = {lkp, (uint64_t)Lq, (uint64_t)B * H, 1};
        const uint64_t strides[3] = {(uint64_t)ldp * 2, (uint64_t)Lq * ldp * 2, (uint64_t)B * H * Lq * ldp * 2};
        e = encode_tmap_bf16_4d(&tm[5], ds_out, dims, strides, 64, 128); if (e) return e;
    }
    const int cap = cta_limit_for(stream);
    const int grid = (int)(total < cap ? total : cap);
    const bool drop = p.thresh16 != 0u;
    if (p.causal) e = drop ? launch<true, true>(tm, p, grid, (cudaStream_t)stream) : launch<true, false>(tm, p, grid, (cudaStream_t)stream);
    else e = drop ? launch<false, true>(tm, p, grid, (cudaStream_t)stream) : launch<false, false>(tm, p, grid, (cudaStream_t)stream);
    if (e) return e;
    count_launch();
    LD_LAUNCH_CHECK("attention_bwd");
    return 0;
}
