// Shared device-side helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM wrappers
// (inline PTX), warp reductions and vector load/store helpers.
//
// Everything here targets sm_100a only (B200). No multi-arch dispatch.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/layoutdetr_sm100.h"

namespace ld {

// ---------------------------------------------------------------------------------------------
// error plumbing shared by all C-ABI entry points
// ---------------------------------------------------------------------------------------------

void set_last_error(const char* fmt, ...);
int  cuda_status(cudaError_t e, const char* what);   // 0 on success, positive cudaError otherwise

#define LD_CHECK_ARG(cond, ...)                                   \
    do { if (!(cond)) { ld::set_last_error(__VA_ARGS__); return LD_ERR_INVALID_ARG; } } while (0)

#define LD_LAUNCH_CHECK(what)                                     \
    do { int _e = ld::cuda_status(cudaGetLastError(), what); if (_e) return _e; } while (0)

static inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

// number of SMs of the current device (cached)
int sm_count();
int cta_limit_for(void* stream);   // min(SM count, ld_set_stream_cta_limit of this stream)

// ---------------------------------------------------------------------------------------------
// small device utilities
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ float bf16_to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ __nv_bfloat16 f32_to_bf16(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void unpack_bf16x2(uint32_t u, float& lo, float& hi) {
    __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&u);
    lo = __low2float(h); hi = __high2float(h);
}

// erf via Abramowitz-Stegun 7.1.28: erf(x) = 1 - (1 + a1 x + ... + a6 x^6)^-16 for x >= 0 (|abs err| < 3e-7, fp32-level):
// 6 FMA + 4 squarings + ONE special-function op (rcp) and no branches — the GELU epilogue runs inside the GEMM's TMEM
// drain, where the two MUFU ops of an exp-based formula were the bottleneck.
__device__ __forceinline__ float fast_erf(float x) {
    const float ax = fabsf(x);
    float t = fmaf(0.0000430638f, ax, 0.0002765672f);
    t = fmaf(t, ax, 0.0001520143f);
    t = fmaf(t, ax, 0.0092705272f);
    t = fmaf(t, ax, 0.0422820123f);
    t = fmaf(t, ax, 0.0705230784f);
    t = fmaf(t, ax, 1.0f);
    t *= t; t *= t; t *= t; t *= t;
    return copysignf(1.0f - __fdividef(1.0f, t), x);
}

// exact (erf) GELU as used by BERT "gelu" (reference training/med.py:301, ACT2FN['gelu'])
// Same Abramowitz-Stegun form with 1/sqrt(2) and a factor 2^(1/16) folded into the coefficients (so that the 16th power
// is 2 t and the reciprocal is 0.5 (1 - erf)) and the sign handled by
//   gelu(x) = relu(x) - |x * 0.5 (1 - erf(|x| / sqrt 2))|        (x >= 0: x - q;  x < 0: q, with q = x / (2 t))
// 6 FMA + 4 MUL + MUFU.RCP + MUL + FMNMX + FADD = 14 issue slots per element (the GEMM epilogue is issue-bound here).
__device__ __forceinline__ float gelu_erf(float x) {
    const float ax = fabsf(x);
    float t = fmaf(5.6212996640e-06f, ax, 5.1055209009e-05f);
    t = fmaf(t, ax, 3.9686137011e-05f);
    t = fmaf(t, ax, 3.4227392389e-03f);
    t = fmaf(t, ax, 2.2076998457e-02f);
    t = fmaf(t, ax, 5.2075163037e-02f);
    t = fmaf(t, ax, 1.0442737824e+00f);
    t *= t; t *= t; t *= t; t *= t;
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));      // t >= 2: no range fix-up needed
    return fmaxf(x, 0.0f) - fabsf(x * r);
}
// Packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2: one issue slot, two IEEE-rn results — same roundings as the scalar ops).
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t f2_pack(float lo, float hi) { f32x2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void f2_unpack(f32x2_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2_t f2_fma(f32x2_t a, f32x2_t b, f32x2_t c) { f32x2_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2_t f2_mul(f32x2_t a, f32x2_t b) { f32x2_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2_t f2_splat(float c) { return f2_pack(c, c); }

// gelu_erf on two values at once: the same operations in the same order as the scalar version (bit-identical results),
// with the polynomial, the four squarings and the final products issued as FFMA2 / FMUL2.
__device__ __forceinline__ void gelu_erf_x2(float& x0, float& x1) {
    const f32x2_t ax = f2_pack(fabsf(x0), fabsf(x1));
    f32x2_t t = f2_fma(f2_splat(5.6212996640e-06f), ax, f2_splat(5.1055209009e-05f));
    t = f2_fma(t, ax, f2_splat(3.9686137011e-05f));
    t = f2_fma(t, ax, f2_splat(3.4227392389e-03f));
    t = f2_fma(t, ax, f2_splat(2.2076998457e-02f));
    t = f2_fma(t, ax, f2_splat(5.2075163037e-02f));
    t = f2_fma(t, ax, f2_splat(1.0442737824e+00f));
    t = f2_mul(t, t); t = f2_mul(t, t); t = f2_mul(t, t); t = f2_mul(t, t);
    float t0, t1, r0, r1;
    f2_unpack(t, t0, t1);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(t0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(t1));
    // relu(x) - |x| r  ==  fma(|x| r, -1, relu(x)): one rounding of the product, exact negation, one rounding of the sum
    const f32x2_t q = f2_mul(ax, f2_pack(r0, r1));
    const f32x2_t o = f2_fma(q, f2_splat(-1.0f), f2_pack(fmaxf(x0, 0.0f), fmaxf(x1, 0.0f)));
    f2_unpack(o, x0, x1);
}

__device__ __forceinline__ float gelu_erf_grad(float x) {
    const float cdf = 0.5f * (1.0f + fast_erf(x * 0.70710678118654752f));
    const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
    return cdf + x * pdf;
}

// ---------------------------------------------------------------------------------------------
// explicit shared-space loads / stores (addresses from smem_u32).  Pointers derived from the 1024-byte-aligned
// dynamic-smem base lose their address space in the compiler and would be accessed with GENERIC LD/ST (long-scoreboard
// latency, LG queue) — measured as the epilogue bottleneck of gemm_bf16_kernel; these compile to LDS / STS.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 lds_u4(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u4(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (launch attribute cudaLaunchAttributeProgrammaticStreamSerialization; no-ops without it):
// pdl_trigger lets the NEXT kernel of the stream start launching, pdl_wait blocks until every kernel this one depends on has
// completed and its writes are visible.  Pattern: trigger first, then the prologue that touches no global data (barrier init,
// TMEM allocation, descriptor prefetch), then wait, then everything else — the prologue and the launch latency of kernel N + 1
// run under the tail of kernel N.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) { }
}

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), rank-4 tiled loads, completion on an mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* tmap, uint64_t* bar,
                                            int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)),
          "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// TMA store of one rank-4 box from (swizzled) shared memory; rows / columns outside the tensor are clipped by the hardware
__device__ __forceinline__ void tma_store_4d(const void* tmap, uint32_t smem_src, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
        ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed bulk stores of this thread have finished READING shared memory (the source may be overwritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---------------------------------------------------------------------------------------------
// Counter-based dropout randomness: Philox4x32-10 (Salmon et al., the generator torch / cuRAND use for nn.Dropout —
// reference training/med.py:96,213,240,318, training/detr_transformer.py:185-194,210).  One call yields 128 bits = eight
// 16-bit lanes; element e of a group is dropped when lane e < thresh16 (thresh16 = round(p * 65536)).  The counter is
// (group index lo, hi, site, step) and the key the 64-bit seed, so forward and backward kernels regenerate the same mask
// from the element coordinates alone.  rng_state (device): {seed_lo, seed_hi, step, 0}.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}
struct DropoutRng {
    uint2 key; uint32_t step, site, thresh16; float scale;          // thresh16 == 0: dropout off
    __device__ __forceinline__ void init(const uint32_t* rng_state, uint32_t site_, uint32_t thresh16_, float scale_) {
        site = site_; thresh16 = thresh16_; scale = scale_;
        if (thresh16_) { key = make_uint2(__ldg(rng_state), __ldg(rng_state + 1)); step = __ldg(rng_state + 2); }
        else { key = make_uint2(0u, 0u); step = 0u; }
    }
    // keep bits (bit e set = element e of the group survives) of the 8-element group `g`
    __device__ __forceinline__ uint32_t keep8(uint64_t g) const {
        const uint4 r = philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), site, step), key);
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
        uint32_t m = 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) m |= (((w[e >> 1] >> (16 * (e & 1))) & 0xFFFFu) >= thresh16 ? 1u : 0u) << e;
        return m;
    }
};

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {   // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {         // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 inputs, fp32 accumulate. One thread issues.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                             uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// A operand from TMEM (bf16 packed), B from smem.
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b,
                                             uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 16 consecutive fp32 columns: thread i of the warp gets lane (base_lane + i)
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// CTA-pair (cta_group::2) variants: two CTAs of a cluster cooperate on one 256-row MMA tile
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes are credited to the barrier at the same smem offset in the LEADER CTA (rank 0)
__device__ __forceinline__ void tma_load_4d_2sm(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & 0xFEFFFFFFu),
          "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive (once all prior MMAs of this thread retire) on the barrier at this smem offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// arrive on the barrier at the same smem offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

// Shared-memory matrix descriptor for tcgen05.mma (SWIZZLE_128B, bf16).
//   bits [0,14)  start address >> 4          bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4     bits [46,48) descriptor version (1 on sm_100)
//   bits [61,64) layout type (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// Same for a K-major tile of 64-byte rows (32 bf16 per row, SWIZZLE_64B: layout type 4; 8-row groups 512 B apart).
__device__ __forceinline__ uint64_t make_smem_desc_sw64(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((16u >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((512u >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}

// Instruction descriptor for kind::f16 with bf16 A/B and fp32 accumulation.
//   c_format [4,6)=1 (f32)  a_format [7,10)=1 (bf16)  b_format [10,13)=1 (bf16)
//   a_major bit 15, b_major bit 16 (0 = K-major, 1 = MN-major), N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ __forceinline__ uint32_t make_idesc_bf16(int m, int n, int a_mn_major, int b_mn_major) {
    uint32_t d = 0;
    d |= 1u << 4;
    d |= 1u << 7;
    d |= 1u << 10;
    d |= (uint32_t)(a_mn_major & 1) << 15;
    d |= (uint32_t)(b_mn_major & 1) << 16;
    d |= (uint32_t)(n >> 3) << 17;
    d |= (uint32_t)(m >> 4) << 24;
    return d;
}

}  // namespace ld
