// Fused multi-head attention forward for sm_100a, persistent and software-pipelined:
//     O = dropout(softmax(Q K^T * scale + mask)) V      per (batch, head, 128-query tile), <= 256 keys, head_dim <= 192
// — BERT text encoder / causal LM decoder (T = 256, head_dim 192, additive -10000 key mask; reference training/med.py:146-228,
// attention-probability dropout :213) and DETR self / cross attention (head_dim 32, 9 / 10 / 64 keys, -inf key-padding mask;
// nn.MultiheadAttention as called by training/detr_transformer.py:208,273,277).
//
// The BERT launch (576 heads x 256 x 256 x 192) is HBM-bound: 208 MB of Q / K / V / O against 29 GFLOP, i.e. 32 us of traffic vs
// 13 us of MMA.  Round 1 ran one tile per CTA with load -> QK^T -> softmax -> PV -> store back to back (104 us, tensor pipe 14 %).
// Here one CTA per SM walks its tiles with TWO tiles in flight ("chains" A and B, alternating tiles), so that one chain's
// softmax / epilogue (SIMT) runs under the other chain's MMAs and under the loads of both:
//
//   warp 0        K / V producer: TMA units through a 3 x 32 KB ring, in the order the MMA warp consumes them
//                 (K unit = all keys x 64 head-dim values, V unit = 64 keys x head_dim)
//   warp 1        MMA issuer (one thread), tiles in pairs: QK^T(i), QK^T(i+1), PV(i), PV(i+1), ...   S[128 x keys] fp32 in TMEM,
//                 one 256-column region per chain; O[128 x d] re-uses the chain's S columns (S is dead once the probabilities are
//                 in smem).  (A strictly alternating order QK(i), PV(i-1) made each chain wait for the other's epilogue + Q load:
//                 13 k cycles per tile measured, 12.3 k in a timeline model of the barriers; the paired order models at 5.9 k.)
//   warp 2        TMEM allocator + Q producer + O store: X[c] holds the chain's Q tile, then its P tile, then the bf16 O tile; this
//                 one thread issues the TMA store of O(i-2) and, once the store has read X[c], the TMA load of Q(i) into it
//   warps 4..11   chain A softmax + epilogue, warps 12..19 chain B: two threads per query row (= TMEM lane), each taking half of
//                 the key blocks (first measurement: one thread per row issued at IPC 0.2 — dependent FADD / MUFU chains with
//                 two warps per scheduler — and the sweeps were 9.7 k cycles of a 13 k-cycle tile).  sweep 1 max (exchanged
//                 through the free tail of X[c]), sweep 2 e = 2^(s - max) -> (dropout) -> bf16 into the swizzled P tile, fp32 row
//                 sums (exchanged through the dead mask buffer); epilogue O * (1 / sum) -> bf16 -> swizzled X[c] -> TMA store
//                 (row-per-thread 16-byte global stores were tried: 3072 half-used sectors per tile, 7.5 k cycles of LSU time).
// Measured alternatives (B200, BERT launch, ncu; timelines from tools/attn_trace.py in profiles/r2_attn_timeline.txt):
//   one thread per row, alternating MMA order, TMA store by a softmax thread      60.6 us
//   two threads per row, direct 16-byte global stores of O                        64.2 us  (LSU-bound epilogue, 7.5 k cycles)
//   + paired MMA order, O staged in X[c] and stored by the Q-producer thread      50.6 us  <- this file
//   + both chains on the two q-tiles of one head sharing every K / V unit         67.6 us  (the chains run in lock step: the
//       SIMT phase of both, 8 k cycles, is no longer hidden under the other chain's MMAs / loads)
// What is left is the serial latency of a tile — O store 2-3 k cycles, Q load 3-4 k, QK^T 1.6 k + K-unit waits, softmax 4.6 k,
// PV 1.6 k + V-unit waits — with only two tiles in flight; a third would need the P tile out of shared memory (TMEM-resident P).
// Neither the fp32 scores nor the probabilities reach HBM; for the backward pass the kernel saves one float per row
// (log2-domain log-sum-exp), from which ld_attention_bwd recomputes P.
#include <cstdlib>
#include "common.cuh"
#include "runtime.h"

namespace {
using namespace ld;

constexpr int AF_THREADS = 640;                // 4 control warps + 2 chains x 8 softmax warps
constexpr int AF_RING = 3;
constexpr int AF_X_BYTES = 65536;              // Q tile (128 x 192 bf16 = 48 KB) during QK^T, then P (128 x 256 bf16)
constexpr int AF_XCH_OFF = 49152;              // free tail of X[c] while Q is live: row-max exchange [2 halves][128 rows] fp32
constexpr int AF_UNIT_BYTES = 32768;           // K unit: 256 keys x 64 values; V unit: 64 keys x 192 values (24 KB)
constexpr int AF_MASK_BYTES = 1024;            // 256 floats per chain (later: row-sum exchange [2 halves][128 rows])
constexpr int AF_BAR_BYTES = 256;
constexpr int AF_SMEM = 2 * AF_X_BYTES + AF_RING * AF_UNIT_BYTES + 2 * AF_MASK_BYTES + AF_BAR_BYTES;
static_assert(AF_SMEM <= 227 * 1024, "shared memory budget");
constexpr float LOG2E = 1.4426950408889634f;

struct AfParams {
    int B, H, Lq, Lk, d;
    int dch, nkb, ncols, q_tiles, total_tiles;
    float scale2, mask2;                       // scale * log2(e), mask_value * log2(e)
    int causal;
    const uint8_t* key_mask;
    __nv_bfloat16* O; long ldo;
    float* lse;                                // optional [B*H, Lq]: max2 + log2(sum) of the scaled + masked scores (log2 domain)
    const uint32_t* rng; uint32_t site, thresh16; float drop_scale;
    long long* trace;                          // debug (ld_debug_attention_trace): per-tile event clocks of CTA 0, 16 slots per tile
};

// event slots of the trace: MMA thread 0..3, first softmax thread of the tile's chain 4..8, Q-producer thread 9..11
enum { EV_QK_START = 0, EV_QK_ISSUED, EV_PV_START, EV_PV_ISSUED, EV_S_SEEN, EV_MAX_DONE, EV_P_DONE, EV_O_SEEN, EV_EPI_DONE,
       EV_STAGED_SEEN, EV_STORE_READ, EV_Q_ISSUED };
__device__ __forceinline__ void trace_ev(const AfParams& p, int tile, int ev) {
    if (p.trace != nullptr && blockIdx.x == 0 && tile < 64) p.trace[tile * 16 + ev] = clock64();
}

struct TileId { int b, h, m0; };
__device__ __forceinline__ TileId decode(const AfParams& p, int t) {
    TileId id;
    const int qt = t % p.q_tiles; t /= p.q_tiles;
    id.h = t % p.H; id.b = t / p.H; id.m0 = qt * 128;
    return id;
}

template <bool CAUSAL, bool DROPOUT>
__global__ void __launch_bounds__(AF_THREADS, 1)
attention_fwd_pipelined_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                               const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const AfParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* x_s = smem;                                          // [2][64 KB]
    uint8_t* ring = smem + 2 * AF_X_BYTES;
    float* mask_s = reinterpret_cast<float*>(ring + AF_RING * AF_UNIT_BYTES);   // [2][256]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(mask_s) + 2 * AF_MASK_BYTES);
    uint64_t* kv_full = bars;             // [3]
    uint64_t* kv_empty = bars + 3;        // [3]
    uint64_t* q_full = bars + 6;          // [2]  Q tile of the chain has landed in X[c]
    uint64_t* s_full = bars + 8;          // [2]  scores complete in TMEM
    uint64_t* p_ready = bars + 10;        // [2]  probabilities in X[c], S columns free (8 warps arrive)
    uint64_t* o_full = bars + 12;         // [2]  O complete in TMEM, P tile dead (X[c] may take the next Q tile)
    uint64_t* o_free = bars + 14;         // [2]  epilogue has drained O from TMEM (8 warps arrive)
    uint64_t* o_staged = bars + 16;       // [2]  bf16 O tile is in X[c], ready for the TMA store (8 warps arrive)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_local = (p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // >= 1 (grid <= tiles)

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0u) asm volatile("trap;");   // SWIZZLE_128B tiles need a 1024-byte aligned base
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < AF_RING; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        for (int c = 0; c < 2; ++c) {
            mbar_init(&q_full[c], 1); mbar_init(&s_full[c], 1);
            mbar_init(&p_ready[c], 8); mbar_init(&o_full[c], 1); mbar_init(&o_free[c], 8); mbar_init(&o_staged[c], 8);
        }
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ K / V producer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            auto k_units = [&](int i) {
                const TileId id = decode(p, (int)blockIdx.x + i * (int)gridDim.x);
                for (int ch = 0; ch < p.dch; ++ch) {
                    mbar_wait(&kv_empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&kv_full[stage], (uint32_t)p.ncols * 128u);
                    tma_load_4d(ring + stage * AF_UNIT_BYTES, &tmK, &kv_full[stage], ch * 64, 0, id.h, id.b);
                    if (++stage == AF_RING) { stage = 0; phase ^= 1; }
                }
            };
            auto v_units = [&](int i) {
                const TileId id = decode(p, (int)blockIdx.x + i * (int)gridDim.x);
                for (int j = 0; j < p.nkb; ++j) {
                    mbar_wait(&kv_empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&kv_full[stage], (uint32_t)p.dch * 8192u);
                    uint8_t* dst = ring + stage * AF_UNIT_BYTES;
                    for (int ch = 0; ch < p.dch; ++ch) tma_load_4d(dst + ch * 8192, &tmV, &kv_full[stage], ch * 64, j * 64, id.h, id.b);
                    if (++stage == AF_RING) { stage = 0; phase ^= 1; }
                }
            };
            for (int i = 0; i < n_local; i += 2) {       // same order as the MMA warp: K(i) K(i+1) V(i) V(i+1)
                k_units(i);
                if (i + 1 < n_local) k_units(i + 1);
                v_units(i);
                if (i + 1 < n_local) v_units(i + 1);
            }
        }
    } else if (warp == 2) {
        // ------------------------------------------------------------------ Q producer + O store
        if (lane == 0) {
            for (int i = 0; i < n_local + 2; ++i) {
                const int c = i & 1, k = i >> 1;
                const uint32_t xa = smem_u32(x_s + c * AF_X_BYTES);
                if (i >= 2) {                                    // O of the chain's previous tile: smem -> HBM
                    const TileId od = decode(p, (int)blockIdx.x + (i - 2) * (int)gridDim.x);
                    mbar_wait(&o_staged[c], (uint32_t)((k - 1) & 1));
                    trace_ev(p, i - 2, EV_STAGED_SEEN);
                    for (int ch = 0; ch < p.dch; ++ch) tma_store_4d(&tmO, xa + ch * 16384, ch * 64, od.m0, od.h, od.b);
                    tma_store_commit();
                }
                if (i < n_local) {
                    const TileId id = decode(p, (int)blockIdx.x + i * (int)gridDim.x);
                    if (i >= 2) tma_store_wait_read();           // the store has read X[c]: it may take the next Q tile
                    trace_ev(p, i, EV_STORE_READ);
                    mbar_arrive_expect_tx(&q_full[c], (uint32_t)p.dch * 16384u);
                    for (int ch = 0; ch < p.dch; ++ch)
                        tma_load_4d(x_s + c * AF_X_BYTES + ch * 16384, &tmQ, &q_full[c], ch * 64, id.m0, id.h, id.b);
                    trace_ev(p, i, EV_Q_ISSUED);
                }
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");            // stores complete before the CTA exits
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            const uint32_t idesc_s = make_idesc_bf16(128, p.ncols, 0, 0);
            const uint32_t idesc_o = make_idesc_bf16(128, p.dch * 64, 0, 1);
            auto pv = [&](int i) {                               // O = P V of local tile i
                const int c = i & 1, k = i >> 1;
                mbar_wait(&p_ready[c], (uint32_t)(k & 1));
                tc_fence_after();
                trace_ev(p, i, EV_PV_START);
                const uint32_t pa = smem_u32(x_s + c * AF_X_BYTES);
                const uint32_t tmem_o = tmem_base + (uint32_t)c * 256u;
                for (int j = 0; j < p.nkb; ++j) {
                    mbar_wait(&kv_full[stage], phase);
                    tc_fence_after();
                    const uint32_t vb = smem_u32(ring + stage * AF_UNIT_BYTES);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint64_t da = make_smem_desc(pa + j * 16384 + kk * 32, 16, 1024);
                        const uint64_t db = make_smem_desc(vb + kk * 2048, 8192, 1024);
                        umma_bf16_ss(tmem_o, da, db, idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
                    }
                    umma_commit(&kv_empty[stage]);
                    if (++stage == AF_RING) { stage = 0; phase ^= 1; }
                }
                umma_commit(&o_full[c]);
                trace_ev(p, i, EV_PV_ISSUED);
            };
            auto qk = [&](int i) {                               // S = Q K^T of local tile i
                const int c = i & 1, k = i >> 1;
                if (k >= 1) { mbar_wait(&o_free[c], (uint32_t)((k - 1) & 1)); tc_fence_after(); }   // previous O of this chain drained
                mbar_wait(&q_full[c], (uint32_t)(k & 1));
                tc_fence_after();
                trace_ev(p, i, EV_QK_START);
                const uint32_t qa = smem_u32(x_s + c * AF_X_BYTES);
                const uint32_t tmem_s = tmem_base + (uint32_t)c * 256u;
                for (int ch = 0; ch < p.dch; ++ch) {             // S += Q[:, 64ch : 64ch+64] K[:, 64ch : 64ch+64]^T
                    mbar_wait(&kv_full[stage], phase);
                    tc_fence_after();
                    const uint32_t kb = smem_u32(ring + stage * AF_UNIT_BYTES);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint64_t da = make_smem_desc(qa + ch * 16384 + kk * 32, 16, 1024);
                        const uint64_t db = make_smem_desc(kb + kk * 32, 16, 1024);
                        umma_bf16_ss(tmem_s, da, db, idesc_s, (ch > 0 || kk > 0) ? 1u : 0u);
                    }
                    umma_commit(&kv_empty[stage]);
                    if (++stage == AF_RING) { stage = 0; phase ^= 1; }
                }
                umma_commit(&s_full[c]);
                trace_ev(p, i, EV_QK_ISSUED);
            };
            for (int i = 0; i < n_local; i += 2) {
                qk(i);
                if (i + 1 < n_local) qk(i + 1);
                pv(i);
                if (i + 1 < n_local) pv(i + 1);
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ softmax + epilogue, chain c (8 warps)
        const int c = (warp - 4) >> 3, half = ((warp - 4) >> 2) & 1, q = warp & 3;   // hardware: a warp reads TMEM lanes 32 * (warp_id % 4) ..
        const int r = q * 32 + lane;                             // query row within the tile == TMEM lane
        const int ct = threadIdx.x - 128 - c * 256;              // 0..255 within the chain
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c * 256u;
        const uint32_t mask_a = smem_u32(mask_s) + c * AF_MASK_BYTES;
        const uint32_t x_a = smem_u32(x_s + c * AF_X_BYTES);
        const uint32_t xmax_a = x_a + AF_XCH_OFF;                // [2][128] fp32, valid between sweep 1 and sweep 2
        const int bar_all = 1 + c, bar_xch = 3 + c;
        const int jb = half == 0 ? 0 : (p.nkb + 1) / 2;          // key blocks of this thread
        const int je = half == 0 ? (p.nkb + 1) / 2 : p.nkb;
        DropoutRng rng;
        rng.init(p.rng, p.site, DROPOUT ? p.thresh16 : 0u, p.drop_scale);
        for (int i = c, k = 0; i < n_local; i += 2, ++k) {
            const TileId id = decode(p, (int)blockIdx.x + i * (int)gridDim.x);
            const int row = id.m0 + r;
            const bool row_ok = row < p.Lq;
            {   // additive key mask of this batch element (log2 domain); keys beyond Lk never attend
                const uint8_t* km = p.key_mask ? p.key_mask + (long)id.b * p.Lk : nullptr;
                sts_f32(mask_a + ct * 4, (ct < p.Lk) ? ((km && km[ct]) ? p.mask2 : 0.0f) : -INFINITY);
            }
            asm volatile("bar.sync %0, 256;" ::"r"(bar_all) : "memory");
            mbar_wait(&s_full[c], (uint32_t)(k & 1));
            tc_fence_after();
            if (ct == 0) trace_ev(p, i, EV_S_SEEN);
            // ---- sweep 1: row max of s2 = acc * scale2 + mask2[col] (+ causal) over this thread's key blocks
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
            for (int j = jb; j < je; ++j) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    uint32_t v[32];
                    tmem_ld_x32(lane_addr + 64 * j + 32 * hh, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i4 = 0; i4 < 8; ++i4) {
                        const int c4 = 64 * j + 32 * hh + 4 * i4;
                        const float4 m4 = lds_f4(mask_a + c4 * 4);
                        float s0 = fmaf(__uint_as_float(v[4 * i4 + 0]), p.scale2, m4.x);
                        float s1 = fmaf(__uint_as_float(v[4 * i4 + 1]), p.scale2, m4.y);
                        float s2 = fmaf(__uint_as_float(v[4 * i4 + 2]), p.scale2, m4.z);
                        float s3 = fmaf(__uint_as_float(v[4 * i4 + 3]), p.scale2, m4.w);
                        if (CAUSAL) {
                            if (c4 + 0 > row) s0 += p.mask2;
                            if (c4 + 1 > row) s1 += p.mask2;
                            if (c4 + 2 > row) s2 += p.mask2;
                            if (c4 + 3 > row) s3 += p.mask2;
                        }
                        mx0 = fmaxf(mx0, s0); mx1 = fmaxf(mx1, s1); mx2 = fmaxf(mx2, s2); mx3 = fmaxf(mx3, s3);
                    }
                }
            }
            float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
            sts_f32(xmax_a + (half * 128 + r) * 4, mx);
            asm volatile("bar.sync %0, 256;" ::"r"(bar_all) : "memory");
            mx = fmaxf(lds_f32(xmax_a + r * 4), lds_f32(xmax_a + (128 + r) * 4));
            if (ct == 0) trace_ev(p, i, EV_MAX_DONE);
            // the exchange words sit where sweep 2 of the upper half writes key block 3 of P: those writers wait (bar.sync) until
            // every thread of the chain has read its maxima (bar.arrive) — in practice never, block 3 is their last
            if (half == 0) asm volatile("bar.arrive %0, 256;" ::"r"(bar_xch) : "memory");
            // ---- sweep 2: e = 2^(s2 - max) -> bf16 (after dropout) into the swizzled A tile of the PV MMAs; fp32 row sum
            float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
            const uint64_t grow = (uint64_t)((long)(id.b * p.H + id.h) * p.Lq + row) * 32u;       // dropout group index base
            for (int j = jb; j < je; ++j) {
                if (half == 1 && j == je - 1) asm volatile("bar.sync %0, 256;" ::"r"(bar_xch) : "memory");
                const uint32_t prow_a = x_a + j * 16384 + r * 128;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    uint32_t v[32];
                    tmem_ld_x32(lane_addr + 64 * j + 32 * hh, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int g = 0; g < 4; ++g) {                // four 16-byte pieces (8 keys each) of this 32-key run
                        const int c8 = 64 * j + 32 * hh + 8 * g;
                        const float4 ma = lds_f4(mask_a + c8 * 4), mb = lds_f4(mask_a + c8 * 4 + 16);
                        const float mm[8] = {ma.x, ma.y, ma.z, ma.w, mb.x, mb.y, mb.z, mb.w};
                        float e[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            float s = fmaf(__uint_as_float(v[8 * g + u]), p.scale2, mm[u]);
                            if (CAUSAL && c8 + u > row) s += p.mask2;
                            e[u] = ex2_approx(s - mx);
                        }
                        sum0 += e[0] + e[4]; sum1 += e[1] + e[5]; sum2 += e[2] + e[6]; sum3 += e[3] + e[7];
                        if (DROPOUT) {
                            const uint32_t keep = rng.keep8(grow + (uint32_t)(c8 >> 3));
#pragma unroll
                            for (int u = 0; u < 8; ++u) e[u] = ((keep >> u) & 1u) ? e[u] * rng.scale : 0.0f;
                        }
                        uint4 o;
                        o.x = pack_bf16x2(e[0], e[1]); o.y = pack_bf16x2(e[2], e[3]);
                        o.z = pack_bf16x2(e[4], e[5]); o.w = pack_bf16x2(e[6], e[7]);
                        sts_u4(prow_a + (((hh * 4 + g) ^ (r & 7)) << 4), o);
                    }
                }
            }
            if (half == 1 && jb == je) asm volatile("bar.sync %0, 256;" ::"r"(bar_xch) : "memory");   // no key block of its own (<= 64 keys)
            // ---- P is in shared memory and this warp no longer reads S: release the PV MMAs (they write O over S's columns)
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_ready[c]);
            if (ct == 0) trace_ev(p, i, EV_P_DONE);
            // ---- row sums of the two halves through the (now dead) mask buffer
            float sum = (sum0 + sum1) + (sum2 + sum3);
            asm volatile("bar.sync %0, 256;" ::"r"(bar_all) : "memory");         // every thread of the chain is done with the mask
            sts_f32(mask_a + (half * 128 + r) * 4, sum);
            asm volatile("bar.sync %0, 256;" ::"r"(bar_all) : "memory");
            sum = lds_f32(mask_a + r * 4) + lds_f32(mask_a + (128 + r) * 4);
            const float inv = __fdividef(1.0f, sum);
            if (p.lse != nullptr && row_ok && half == 0) p.lse[(long)(id.b * p.H + id.h) * p.Lq + row] = mx + __log2f(sum);
            // ---- epilogue: O * (1 / sum) -> bf16 -> swizzled tile in X[c] (the P tile is dead), stored by the Q-producer thread.
            //      This thread: every other 32-column block of its row.
            mbar_wait(&o_full[c], (uint32_t)(k & 1));
            tc_fence_after();
            if (ct == 0) trace_ev(p, i, EV_O_SEEN);
            for (int c0 = 32 * half; c0 < p.dch * 64; c0 += 64) {
                if (c0 >= p.d) break;                            // warp-uniform: zero-padded columns of head_dim < 64 (never stored)
                uint32_t v[32];
                tmem_ld_x32(lane_addr + c0, v);
                tmem_ld_wait();
                const uint32_t orow_a = x_a + (c0 >> 6) * 16384 + r * 128;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 o;
                    o.x = pack_bf16x2(__uint_as_float(v[8 * g + 0]) * inv, __uint_as_float(v[8 * g + 1]) * inv);
                    o.y = pack_bf16x2(__uint_as_float(v[8 * g + 2]) * inv, __uint_as_float(v[8 * g + 3]) * inv);
                    o.z = pack_bf16x2(__uint_as_float(v[8 * g + 4]) * inv, __uint_as_float(v[8 * g + 5]) * inv);
                    o.w = pack_bf16x2(__uint_as_float(v[8 * g + 6]) * inv, __uint_as_float(v[8 * g + 7]) * inv);
                    sts_u4(orow_a + (((4 * half + g) ^ (r & 7)) << 4), o);
                }
            }
            fence_proxy_async_smem();                            // generic-proxy writes of the O tile -> visible to the TMA store
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&o_free[c]);                         // TMEM columns of this chain may take the next QK^T
                mbar_arrive(&o_staged[c]);                       // O tile staged: the Q-producer thread stores it and reloads X[c]
            }
            if (ct == 0) trace_ev(p, i, EV_EPI_DONE);
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

template <bool CAUSAL, bool DROPOUT>
int launch(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const CUtensorMap& tmO, const AfParams& p, int grid,
           cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        int s = cuda_status(cudaFuncSetAttribute(attention_fwd_pipelined_kernel<CAUSAL, DROPOUT>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, AF_SMEM), "attention_fwd: smem attr");
        if (s) return s;
        attr_set = true;
    }
    attention_fwd_pipelined_kernel<CAUSAL, DROPOUT><<<grid, AF_THREADS, AF_SMEM, stream>>>(tmQ, tmK, tmV, tmO, p);
    return 0;
}
long long* g_attention_trace = nullptr;
}  // namespace

// Debug aid: event clocks (clock64 of the SM running CTA 0) of the next ld_attention_fwd launches are written to `dev_buf`
// ([64 tiles][16 events] int64, see the EV_* slots above; nullptr switches it off).  tools/attn_trace.py prints the timeline.
extern "C" int ld_debug_attention_trace(void* dev_buf) { g_attention_trace = (long long*)dev_buf; return 0; }

// q / k / v point at column 0 of head 0 inside row-major [B*L, ld] bf16 buffers (head h at columns h*d .. h*d+d).
extern "C" int ld_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                void* o, int64_t ldo, float* lse_out,
                                int B, int H, int Lq, int Lk, int d, float scale, const uint8_t* key_mask, int mask_inf, int causal,
                                float dropout_p, const uint32_t* rng_state, uint32_t rng_site, void* stream) {
    using namespace ld;
    static const int env_v1 = [] { const char* e = getenv("LD_ATTN_V1"); return e ? atoi(e) : 0; }();
    if (env_v1 && dropout_p == 0.0f && lse_out == nullptr)       // round-1 kernel, kept for A/B measurements only
        return ld_attention_fwd_v1(q, ldq, k, ldk, v, ldv, o, ldo, nullptr, 0, B, H, Lq, Lk, d, scale, key_mask, mask_inf, causal, stream);
    LD_CHECK_ARG(q && k && v && o && B > 0 && H > 0 && Lq > 0 && Lk > 0, "attention_fwd: bad argument");
    LD_CHECK_ARG(Lk <= 256 && d <= 192 && d % 8 == 0, "attention_fwd: needs <= 256 keys and head_dim <= 192 (multiple of 8); got Lk=%d d=%d", Lk, d);
    LD_CHECK_ARG(ldo % 8 == 0 && ((uintptr_t)o & 15) == 0, "attention_fwd: output alignment");
    LD_CHECK_ARG(dropout_p >= 0.0f && dropout_p < 1.0f && (dropout_p == 0.0f || rng_state != nullptr), "attention_fwd: dropout arguments");
    AfParams p{};
    p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk; p.d = d;
    p.dch = (d + 63) / 64; p.nkb = (Lk + 63) / 64; p.ncols = p.nkb * 64; p.q_tiles = (Lq + 127) / 128;
    const long total = (long)B * H * p.q_tiles;
    LD_CHECK_ARG(total < (1L << 30), "attention_fwd: too many tiles");
    p.total_tiles = (int)total;
    p.scale2 = scale * LOG2E; p.mask2 = mask_inf ? -INFINITY : -10000.0f * LOG2E; p.causal = causal ? 1 : 0;
    p.key_mask = key_mask;
    p.O = (__nv_bfloat16*)o; p.ldo = ldo; p.lse = lse_out;
    p.rng = rng_state; p.site = rng_site;
    p.thresh16 = dropout_p > 0.0f ? (uint32_t)(dropout_p * 65536.0f + 0.5f) : 0u;
    p.drop_scale = dropout_p > 0.0f ? 65536.0f / (65536.0f - (float)p.thresh16) : 1.0f;
    p.trace = g_attention_trace;
    alignas(64) CUtensorMap tmQ, tmK, tmV, tmO;
    int e = make_map(&tmQ, q, ldq, d, Lq, H, B, 128); if (e) return e;
    e = make_map(&tmK, k, ldk, d, Lk, H, B, (uint32_t)p.ncols); if (e) return e;
    e = make_map(&tmV, v, ldv, d, Lk, H, B, 64); if (e) return e;
    e = make_map(&tmO, o, ldo, d, Lq, H, B, 128); if (e) return e;
    const int cap = cta_limit_for(stream);
    const int grid = (int)(total < cap ? total : cap);
    const bool drop = p.thresh16 != 0u;
    if (p.causal) e = drop ? launch<true, true>(tmQ, tmK, tmV, tmO, p, grid, (cudaStream_t)stream)
                           : launch<true, false>(tmQ, tmK, tmV, tmO, p, grid, (cudaStream_t)stream);
    else e = drop ? launch<false, true>(tmQ, tmK, tmV, tmO, p, grid, (cudaStream_t)stream)
                  : launch<false, false>(tmQ, tmK, tmV, tmO, p, grid, (cudaStream_t)stream);
    if (e) return e;
    count_launch();
    LD_LAUNCH_CHECK("attention_fwd");
    return 0;
}
