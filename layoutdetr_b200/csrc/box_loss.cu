// Layout box losses of the generator objective — generalised IoU, pairwise overlap, alignment — value and analytic
// gradient in ONE launch each, replacing ~60 eager elementwise launches per phase of the reference
// (training/loss.py:117-125 -> metrics/metric_layoutnet.py:153-201, 245-275).  The arithmetic lives in
// box_loss_math.h (shared with the host-compiled parity test).  These kernels are latency-bound: B*N*4 floats in,
// B (+ B*N*4) floats out; everything is gather-style and summed in a fixed order, so results are deterministic.
#include "common.cuh"
#include "runtime.h"
#include "box_loss_math.h"

namespace {
constexpr int BOX_MAX_SLOTS = 64;

// one CTA per layout, one thread per slot
__global__ void __launch_bounds__(BOX_MAX_SLOTS)
layout_losses_kernel(const float* __restrict__ bbox, const uint8_t* __restrict__ valid, int N,
                     float* __restrict__ overlap, float* __restrict__ alignment,
                     float* __restrict__ j_overlap, float* __restrict__ j_alignment) {
    __shared__ float sb[BOX_MAX_SLOTS * 4];
    __shared__ uint8_t sv[BOX_MAX_SLOTS];
    __shared__ float s_ov[BOX_MAX_SLOTS], s_al[BOX_MAX_SLOTS];
    const long b = blockIdx.x;
    const int i = threadIdx.x;
    if (i < N) {
        const float4 q = reinterpret_cast<const float4*>(bbox)[b * N + i];
        sb[4 * i] = q.x; sb[4 * i + 1] = q.y; sb[4 * i + 2] = q.z; sb[4 * i + 3] = q.w;
        sv[i] = valid[b * N + i] ? 1 : 0;
    }
    __syncthreads();
    int nvalid = 0;
    for (int j = 0; j < N; ++j) nvalid += sv[j];
    const float inv = 1.f / (float)nvalid;                 // 0 valid slots -> inf / NaN, as the reference's division
    if (i < N) {
        float g[4];
        s_ov[i] = ldbox::overlap_box(sb, sv, N, i, j_overlap ? g : nullptr);
        if (j_overlap) reinterpret_cast<float4*>(j_overlap)[b * N + i] = make_float4(g[0] * inv, g[1] * inv, g[2] * inv, g[3] * inv);
        s_al[i] = ldbox::alignment_box(sb, sv, N, i, j_alignment ? g : nullptr);
        if (j_alignment) reinterpret_cast<float4*>(j_alignment)[b * N + i] = make_float4(g[0] * inv, g[1] * inv, g[2] * inv, g[3] * inv);
    }
    __syncthreads();
    if (i == 0) {
        float so = 0.f, sa = 0.f;
        for (int j = 0; j < N; ++j) { so += s_ov[j]; sa += s_al[j]; }
        overlap[b] = so * inv;
        alignment[b] = sa * inv;
    }
}

// evaluation sweep: per layout, mean over the valid slots of IoU(real_i, fake_i) and of the DocSim weight
// (metrics/overlap50k_alignment50k_layoutwise_iou50k_layoutwise_docsim50k.py:36-45).  One warp per layout.
__global__ void __launch_bounds__(128)
layout_pair_metrics_kernel(const float* __restrict__ real, const float* __restrict__ fake, const uint8_t* __restrict__ valid,
                           long B, int N, float* __restrict__ iou, float* __restrict__ docsim) {
    const long b = (long)blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    float si = 0.f, sd = 0.f, n = 0.f;
    for (int i = lane; i < N; i += 32) {
        if (!valid[b * N + i]) continue;
        const float4 p = reinterpret_cast<const float4*>(real)[b * N + i], q = reinterpret_cast<const float4*>(fake)[b * N + i];
        const float pf[4] = {p.x, p.y, p.z, p.w}, qf[4] = {q.x, q.y, q.z, q.w};
        si += ldbox::iou_pair(pf, qf);
        sd += ldbox::docsim_pair(pf, qf);
        n += 1.f;
    }
    si = ld::warp_sum(si); sd = ld::warp_sum(sd); n = ld::warp_sum(n);
    if (lane == 0) { iou[b] = si / n; docsim[b] = sd / n; }
}

constexpr int GIOU_THREADS = 256;
// one CTA: mean over M index-paired rows of 1 - GIoU; J[m, :] = d mean / d fake[m, :]
__global__ void __launch_bounds__(GIOU_THREADS)
giou_loss_kernel(const float* __restrict__ fake, const float* __restrict__ real, long M, float* __restrict__ loss,
                 float* __restrict__ j_fake) {
    __shared__ float part[GIOU_THREADS];
    const float inv = 1.f / (float)M;
    float acc = 0.f;
    for (long m = threadIdx.x; m < M; m += GIOU_THREADS) {
        const float4 p = reinterpret_cast<const float4*>(fake)[m], q = reinterpret_cast<const float4*>(real)[m];
        const float pf[4] = {p.x, p.y, p.z, p.w}, qf[4] = {q.x, q.y, q.z, q.w};
        float g[4];
        acc += ldbox::giou_row(pf, qf, g);
        if (j_fake) reinterpret_cast<float4*>(j_fake)[m] = make_float4(g[0] * inv, g[1] * inv, g[2] * inv, g[3] * inv);
    }
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int s = GIOU_THREADS / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) part[threadIdx.x] += part[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) loss[0] = part[0] * inv;
}

// out[i] (+)= J[i] * g[i / per]   (chain rule of a per-layout / scalar loss through its stored Jacobian rows)
__global__ void rows_scale_kernel(const float* __restrict__ J, const float* __restrict__ g, float* __restrict__ out, long n,
                                  long per, int accumulate) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const float v = J[i] * g[i / per];
        out[i] = accumulate ? out[i] + v : v;
    }
}
}  // namespace

extern "C" int ld_layout_losses(const float* bbox, const uint8_t* valid, int64_t B, int N, float* overlap, float* alignment,
                                float* j_overlap, float* j_alignment, void* stream) {
    LD_CHECK_ARG(bbox && valid && overlap && alignment, "ld_layout_losses: null pointer");
    LD_CHECK_ARG(N >= 1 && N <= BOX_MAX_SLOTS, "ld_layout_losses: N = %d slots outside 1..%d", N, BOX_MAX_SLOTS);
    LD_CHECK_ARG((reinterpret_cast<uintptr_t>(bbox) & 15) == 0 && (!j_overlap || (reinterpret_cast<uintptr_t>(j_overlap) & 15) == 0) &&
                 (!j_alignment || (reinterpret_cast<uintptr_t>(j_alignment) & 15) == 0), "ld_layout_losses: box buffers must be 16-byte aligned");
    if (B <= 0) return 0;
    layout_losses_kernel<<<(unsigned)B, BOX_MAX_SLOTS, 0, (cudaStream_t)stream>>>(bbox, valid, N, overlap, alignment, j_overlap, j_alignment);
    ld::count_launch();
    LD_LAUNCH_CHECK("ld_layout_losses");
    return 0;
}

extern "C" int ld_giou_loss(const float* fake, const float* real, int64_t M, float* loss, float* j_fake, void* stream) {
    LD_CHECK_ARG(fake && real && loss, "ld_giou_loss: null pointer");
    LD_CHECK_ARG(M >= 1, "ld_giou_loss: M = %ld rows", (long)M);
    LD_CHECK_ARG(((reinterpret_cast<uintptr_t>(fake) | reinterpret_cast<uintptr_t>(real) | reinterpret_cast<uintptr_t>(j_fake)) & 15) == 0,
                 "ld_giou_loss: box buffers must be 16-byte aligned");
    giou_loss_kernel<<<1, GIOU_THREADS, 0, (cudaStream_t)stream>>>(fake, real, (long)M, loss, j_fake);
    ld::count_launch();
    LD_LAUNCH_CHECK("ld_giou_loss");
    return 0;
}

extern "C" int ld_rows_scale(const float* J, const float* g, float* out, int64_t n, int64_t per, int accumulate, void* stream) {
    LD_CHECK_ARG(J && g && out, "ld_rows_scale: null pointer");
    LD_CHECK_ARG(per >= 1, "ld_rows_scale: per = %ld", (long)per);
    if (n <= 0) return 0;
    const int threads = 256;
    const int grid = (int)((n + threads - 1) / threads < 1184 ? (n + threads - 1) / threads : 1184);
    rows_scale_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(J, g, out, (long)n, (long)per, accumulate);
    ld::count_launch();
    LD_LAUNCH_CHECK("ld_rows_scale");
    return 0;
}

extern "C" int ld_layout_pair_metrics(const float* real, const float* fake, const uint8_t* valid, int64_t B, int N, float* iou,
                                      float* docsim, void* stream) {
    LD_CHECK_ARG(real && fake && valid && iou && docsim, "ld_layout_pair_metrics: null pointer");
    LD_CHECK_ARG(N >= 1, "ld_layout_pair_metrics: N = %d", N);
    LD_CHECK_ARG(((reinterpret_cast<uintptr_t>(real) | reinterpret_cast<uintptr_t>(fake)) & 15) == 0,
                 "ld_layout_pair_metrics: box buffers must be 16-byte aligned");
    if (B <= 0) return 0;
    layout_pair_metrics_kernel<<<(unsigned)((B + 3) / 4), 128, 0, (cudaStream_t)stream>>>(real, fake, valid, (long)B, N, iou, docsim);
    ld::count_launch();
    LD_LAUNCH_CHECK("ld_layout_pair_metrics");
    return 0;
}
