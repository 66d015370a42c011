// Scalar arithmetic of the three layout box losses, value AND analytic gradient, written once for the device kernels
// (box_loss.cu) and for a host build (tests compile this header with g++ and compare it with autograd of the oracle).
//
// Reference: metrics/metric_layoutnet.py  generalized_iou_loss :245-275, compute_overlap :153-179,
// compute_alignment :182-201, and convert_xywh_to_ltrb util.py:62-68.  The gradients are those PyTorch autograd
// produces for these expressions: maximum / minimum split a tie evenly, torch.where passes the gradient of the taken
// branch only, nan_to_num and masked_fill stop it, `min(dim)` sends it to the first minimum.
#pragma once
#include <math.h>
#include <float.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define LD_HD __host__ __device__ __forceinline__
#else
#define LD_HD static inline
#endif

namespace ldbox {

struct Ltrb { float l, t, r, b; };

LD_HD Ltrb to_ltrb(const float* q) {            // q = (xc, yc, w, h)
    Ltrb o;
    o.l = q[0] - q[2] / 2; o.t = q[1] - q[3] / 2; o.r = q[0] + q[2] / 2; o.b = q[1] + q[3] / 2;
    return o;
}
// chain rule of to_ltrb: d/d(xc, yc, w, h) from d/d(l, t, r, b)
LD_HD void ltrb_grad_to_xywh(float dl, float dt, float dr, float db, float* g) {
    g[0] = dl + dr; g[1] = dt + db; g[2] = (dr - dl) * 0.5f; g[3] = (db - dt) * 0.5f;
}
// d max(a, b) / da  and  d min(a, b) / da  as autograd defines them
LD_HD float dmax_a(float a, float b) { return a > b ? 1.f : (a == b ? 0.5f : 0.f); }
LD_HD float dmin_a(float a, float b) { return a < b ? 1.f : (a == b ? 0.5f : 0.f); }

// ---------------------------------------------------------------------------------------------------------------
// generalised IoU of one index-paired row: returns 1 - GIoU, writes d(1 - GIoU) / d fake(xc, yc, w, h) into g[4]
// ---------------------------------------------------------------------------------------------------------------
LD_HD float giou_row(const float* fake, const float* real, float* g) {
    const Ltrb p = to_ltrb(fake), q = to_ltrb(real);
    const float a1 = (p.r - p.l) * (p.b - p.t), a2 = (q.r - q.l) * (q.b - q.t);
    const float lmax = fmaxf(p.l, q.l), rmin = fminf(p.r, q.r), tmax = fmaxf(p.t, q.t), bmin = fminf(p.b, q.b);
    const bool cond = (lmax < rmin) && (tmax < bmin);
    const float iw = rmin - lmax, ih = bmin - tmax;
    const float ai = cond ? iw * ih : 0.f;
    const float au = a1 + a2 - ai;
    const float iou = ai / au;
    const float lmin = fminf(p.l, q.l), rmax = fmaxf(p.r, q.r), tmin = fminf(p.t, q.t), bmax = fmaxf(p.b, q.b);
    const float cw = rmax - lmin, ch = bmax - tmin;
    const float ac = cw * ch;
    const float giou = iou - (ac - au) / ac;
    // loss = 1 - ai/au + (ac - au)/ac = 2 - ai/au - au/ac
    const float d_ai = -1.f / au;                                   // direct
    const float d_au = ai / (au * au) - 1.f / ac;
    const float d_ac = au / (ac * ac);
    // au = a1 + a2 - ai
    const float D_ai = d_ai - d_au;                                 // total derivative w.r.t. ai
    const float D_a1 = d_au;
    // ai = iw * ih (if cond);  iw = min(r) - max(l), ih = min(b) - max(t)
    const float d_iw = cond ? D_ai * ih : 0.f, d_ih = cond ? D_ai * iw : 0.f;
    // ac = cw * ch
    const float d_cw = d_ac * ch, d_ch = d_ac * cw;
    const float pw = p.r - p.l, ph = p.b - p.t;
    const float dl = -d_iw * dmax_a(p.l, q.l) - d_cw * dmin_a(p.l, q.l) - D_a1 * ph;
    const float dr = d_iw * dmin_a(p.r, q.r) + d_cw * dmax_a(p.r, q.r) + D_a1 * ph;
    const float dt = -d_ih * dmax_a(p.t, q.t) - d_ch * dmin_a(p.t, q.t) - D_a1 * pw;
    const float db = d_ih * dmin_a(p.b, q.b) + d_ch * dmax_a(p.b, q.b) + D_a1 * pw;
    ltrb_grad_to_xywh(dl, dt, dr, db, g);
    return 1.f - giou;
}

// ---------------------------------------------------------------------------------------------------------------
// overlap, seen from box i of one layout: value = sum_{j != i} area(i ∩ j) / area(i)  (boxes of invalid slots are
// zeroed first, as the reference does), gradient g[4] = d(sum over ALL ordered pairs of the layout) / d box i.
// Invalid slots have gradient 0 (masked_fill).
// ---------------------------------------------------------------------------------------------------------------
LD_HD float nan_to_num_f(float v, bool* pass) {
    if (v != v) { *pass = false; return 0.f; }
    if (isinf(v)) { *pass = false; return v > 0 ? FLT_MAX : -FLT_MAX; }
    *pass = true;
    return v;
}

// one ordered pair (a = the box whose area divides, b = the other box): value and the gradients w.r.t. a's and b's ltrb
LD_HD float overlap_pair(const Ltrb& a, const Ltrb& b, float* ga /*l,t,r,b*/, float* gb /*l,t,r,b*/) {
    const float aw = a.r - a.l, ah = a.b - a.t;
    const float a1 = aw * ah;
    const float lmax = fmaxf(a.l, b.l), rmin = fminf(a.r, b.r), tmax = fmaxf(a.t, b.t), bmin = fminf(a.b, b.b);
    const bool cond = (lmax < rmin) && (tmax < bmin);
    const float iw = rmin - lmax, ih = bmin - tmax;
    const float ai = cond ? iw * ih : 0.f;
    bool pass;
    const float v = nan_to_num_f(ai / a1, &pass);
    for (int k = 0; k < 4; ++k) { ga[k] = 0.f; gb[k] = 0.f; }
    if (pass) {
        const float d_ai = 1.f / a1, d_a1 = -ai / (a1 * a1);
        const float d_iw = cond ? d_ai * ih : 0.f, d_ih = cond ? d_ai * iw : 0.f;
        ga[0] = -d_iw * dmax_a(a.l, b.l) - d_a1 * ah;  gb[0] = -d_iw * dmax_a(b.l, a.l);
        ga[2] = d_iw * dmin_a(a.r, b.r) + d_a1 * ah;   gb[2] = d_iw * dmin_a(b.r, a.r);
        ga[1] = -d_ih * dmax_a(a.t, b.t) - d_a1 * aw;  gb[1] = -d_ih * dmax_a(b.t, a.t);
        ga[3] = d_ih * dmin_a(a.b, b.b) + d_a1 * aw;   gb[3] = d_ih * dmin_a(b.b, a.b);
    }
    return v;
}

LD_HD Ltrb masked_ltrb(const float* bbox, const uint8_t* valid, int j) {
    const float z[4] = {0.f, 0.f, 0.f, 0.f};
    return to_ltrb(valid[j] ? bbox + 4 * j : z);
}

LD_HD float overlap_box(const float* bbox /*[N,4]*/, const uint8_t* valid /*[N]*/, int N, int i, float* g /*[4] or null*/) {
    const Ltrb bi = masked_ltrb(bbox, valid, i);
    float sum = 0.f, gl = 0.f, gt = 0.f, gr = 0.f, gb = 0.f;
    for (int j = 0; j < N; ++j) {
        if (j == i) continue;
        const Ltrb bj = masked_ltrb(bbox, valid, j);
        float ga[4], gq[4];
        sum += overlap_pair(bi, bj, ga, gq);                 // pair (i, j): box i is the divisor
        gl += ga[0]; gt += ga[1]; gr += ga[2]; gb += ga[3];
        overlap_pair(bj, bi, ga, gq);                        // pair (j, i): box i is the other box
        gl += gq[0]; gt += gq[1]; gr += gq[2]; gb += gq[3];
    }
    if (g) {
        if (valid[i]) ltrb_grad_to_xywh(gl, gt, gr, gb, g);
        else { g[0] = g[1] = g[2] = g[3] = 0.f; }
    }
    return sum;
}

// ---------------------------------------------------------------------------------------------------------------
// alignment: for a valid box i, m_i = min over the six coordinates (l, xc, r, t, yc, b) and over all other SLOTS j
// (padded slots included, as in the reference) of |X_c,i - X_c,j|; m_i == 1 -> 0; term_i = -log(1 - m_i).
// ---------------------------------------------------------------------------------------------------------------
LD_HD float coord6(const float* q, int c) {      // q = (xc, yc, w, h); c: 0 l, 1 xc, 2 r, 3 t, 4 yc, 5 b
    switch (c) {
        case 0: return q[0] - q[2] / 2;
        case 1: return q[0];
        case 2: return q[0] + q[2] / 2;
        case 3: return q[1] - q[3] / 2;
        case 4: return q[1];
        default: return q[1] + q[3] / 2;
    }
}
// d coord6(q, c) / d q, scaled by s, accumulated into g[4]
LD_HD void coord6_grad(int c, float s, float* g) {
    switch (c) {
        case 0: g[0] += s; g[2] -= 0.5f * s; break;
        case 1: g[0] += s; break;
        case 2: g[0] += s; g[2] += 0.5f * s; break;
        case 3: g[1] += s; g[3] -= 0.5f * s; break;
        case 4: g[1] += s; break;
        default: g[1] += s; g[3] += 0.5f * s; break;
    }
}

struct AlignMin { float m; int c, j; float sgn; };   // m: minimum |delta|; (c, j): where; sgn: sign(X_ci - X_cj)

LD_HD AlignMin align_min(const float* bbox, int N, int i) {
    AlignMin r; r.m = 1.0f; r.c = -1; r.j = -1; r.sgn = 0.f;
    // reference order: min over j first (first minimum), then over c (first minimum)
    float best = INFINITY;
    for (int c = 0; c < 6; ++c) {
        const float xi = coord6(bbox + 4 * i, c);
        float mc = INFINITY; int jc = -1; float sc = 0.f;
        for (int j = 0; j < N; ++j) {
            float d, s;
            if (j == i) { d = 1.0f; s = 0.f; }               // diagonal is set to 1 (no gradient)
            else {
                const float delta = xi - coord6(bbox + 4 * j, c);
                d = fabsf(delta);
                s = delta > 0.f ? 1.f : (delta < 0.f ? -1.f : 0.f);
            }
            if (d < mc) { mc = d; jc = j; sc = s; }
        }
        if (mc < best) { best = mc; r.m = mc; r.c = c; r.j = (jc == i ? -1 : jc); r.sgn = sc; }
    }
    return r;
}

// value of box i's term and gradient of the layout's summed terms w.r.t. box i (own term + terms of the boxes whose
// minimum is attained against box i)
LD_HD float alignment_box(const float* bbox, const uint8_t* valid, int N, int i, float* g /*[4] or null*/) {
    float term = 0.f;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (valid[i]) {
        const AlignMin a = align_min(bbox, N, i);
        const float m = (a.m == 1.0f) ? 0.f : a.m;
        term = -logf(1.f - m);
        if (g && a.m != 1.0f && a.j >= 0) coord6_grad(a.c, a.sgn / (1.f - m), acc);
    }
    if (g) {
        for (int k = 0; k < N; ++k) {
            if (k == i || !valid[k]) continue;
            const AlignMin a = align_min(bbox, N, k);
            if (a.j == i && a.m != 1.0f) coord6_grad(a.c, -a.sgn / (1.f - a.m), acc);
        }
        g[0] = acc[0]; g[1] = acc[1]; g[2] = acc[2]; g[3] = acc[3];
    }
    return term;
}

// ---------------------------------------------------------------------------------------------------------------
// evaluation metrics of one index-paired box pair (metrics/metric_layoutnet.py compute_iou :65-91,
// compute_docsim_weight :204-221); no gradients
// ---------------------------------------------------------------------------------------------------------------
LD_HD float iou_pair(const float* p4, const float* q4) {
    const Ltrb p = to_ltrb(p4), q = to_ltrb(q4);
    const float a1 = (p.r - p.l) * (p.b - p.t), a2 = (q.r - q.l) * (q.b - q.t);
    const float lmax = fmaxf(p.l, q.l), rmin = fminf(p.r, q.r), tmax = fmaxf(p.t, q.t), bmin = fminf(p.b, q.b);
    const bool cond = (lmax < rmin) && (tmax < bmin);
    const float ai = cond ? (rmin - lmax) * (bmin - tmax) : 0.f;
    bool pass;
    return nan_to_num_f(ai / (a1 + a2 - ai), &pass);
}

LD_HD float docsim_pair(const float* p, const float* q) {
    const float dx = p[0] - q[0], dy = p[1] - q[1];
    const float location_difference = sqrtf(dx * dx + dy * dy);
    const float shape_difference = fabsf(p[2] - q[2]) + fabsf(p[3] - q[3]);
    const float area_factor = sqrtf(fminf(p[2] * p[3], q[2] * q[3]));
    return area_factor * exp2f(-location_difference - 2.0f * shape_difference);
}

}  // namespace ldbox
