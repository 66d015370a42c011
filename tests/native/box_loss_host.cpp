// Host build of the product's box-loss arithmetic (layoutdetr_b200/csrc/box_loss_math.h — the same header the CUDA
// kernels in box_loss.cu compile) so that the CPU test suite can compare values and analytic gradients with autograd
// of the oracle.  The loops below mirror the kernels' gather structure (one "thread" per slot, fixed summation order).
#include <stdint.h>
#include "../../layoutdetr_b200/csrc/box_loss_math.h"

extern "C" void host_layout_losses(const float* bbox, const uint8_t* valid, long B, int N, float* overlap, float* alignment,
                                   float* j_overlap, float* j_alignment) {
    for (long b = 0; b < B; ++b) {
        const float* sb = bbox + b * N * 4;
        const uint8_t* sv = valid + b * N;
        int nvalid = 0;
        for (int j = 0; j < N; ++j) nvalid += sv[j] ? 1 : 0;
        const float inv = 1.f / (float)nvalid;
        float so = 0.f, sa = 0.f;
        for (int i = 0; i < N; ++i) {
            float g[4];
            so += ldbox::overlap_box(sb, sv, N, i, g);
            for (int k = 0; k < 4; ++k) j_overlap[(b * N + i) * 4 + k] = g[k] * inv;
            sa += ldbox::alignment_box(sb, sv, N, i, g);
            for (int k = 0; k < 4; ++k) j_alignment[(b * N + i) * 4 + k] = g[k] * inv;
        }
        overlap[b] = so * inv;
        alignment[b] = sa * inv;
    }
}

extern "C" void host_giou_loss(const float* fake, const float* real, long M, float* loss, float* j_fake) {
    const float inv = 1.f / (float)M;
    double acc = 0.0;
    for (long m = 0; m < M; ++m) {
        float g[4];
        acc += ldbox::giou_row(fake + 4 * m, real + 4 * m, g);
        for (int k = 0; k < 4; ++k) j_fake[4 * m + k] = g[k] * inv;
    }
    loss[0] = (float)(acc * inv);
}

extern "C" void host_layout_pair_metrics(const float* real, const float* fake, const uint8_t* valid, long B, int N, float* iou,
                                         float* docsim) {
    for (long b = 0; b < B; ++b) {
        float si = 0.f, sd = 0.f, n = 0.f;
        for (int i = 0; i < N; ++i) {
            if (!valid[b * N + i]) continue;
            si += ldbox::iou_pair(real + (b * N + i) * 4, fake + (b * N + i) * 4);
            sd += ldbox::docsim_pair(real + (b * N + i) * 4, fake + (b * N + i) * 4);
            n += 1.f;
        }
        iou[b] = si / n;
        docsim[b] = sd / n;
    }
}
