/*
 * raster_math.h -- per-face / per-pixel arithmetic of the triangle rasterizer, shared by every
 * kernel of the forward and backward path (and by the host-side emulation used in tests/).
 *
 * The functions restate, operation by operation, what the five entry points of
 * `neural_renderer.cuda.rasterize` compute for ONE face / ONE pixel
 * (call sites: /root/reference/meshreg/neurender/rasterize.py:202,232,269,290,306).
 * They are written so that fp32 results are reproducible bit for bit:
 *   - no FMA contraction: translation units including this header are compiled with
 *     `-fmad=false` (nvcc) / `-ffp-contract=off` (host);
 *   - IEEE division (`-prec-div=true`, the nvcc default);
 *   - the one double-precision step of the reference (`1. / sum`) is kept in double.
 * Because the forward and the backward kernels call the SAME functions, the backward can
 * recompute barycentric weights and depth from `face_index_map` + `faces` instead of reading
 * `weight_map` / `face_inv_map` / sampling maps back from HBM.
 */
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define HOC_HD __host__ __device__ __forceinline__
#else
#define HOC_HD static inline
#endif

/* Back-face rule in NDC (y up): true => the face is skipped everywhere. */
HOC_HD bool hoc_face_back(const float *f)
{
    return (f[7] - f[1]) * (f[3] - f[0]) < (f[4] - f[1]) * (f[6] - f[0]);
}

/* x and y of the three vertices are finite (z may be anything). */
HOC_HD bool hoc_face_xy_finite(const float *f)
{
    const float s = (f[0] - f[0]) + (f[1] - f[1]) + (f[3] - f[3]) + (f[4] - f[4]) + (f[6] - f[6]) + (f[7] - f[7]);
    return s == 0.0f;
}

/* NDC -> pixel-index space: p = 0.5 * (x * S + S - 1). */
HOC_HD float hoc_ndc_to_pix(float v, int S)
{
    const float fs = (float)S;
    return 0.5f * (v * fs + fs - 1.0f);
}

/* Pixel-centre NDC coordinate of pixel index i: (2 i + 1 - S) / S.  The reference evaluates this
 * in double and rounds; for S <= 2048 the correctly rounded float quotient is identical
 * (checked exhaustively in tests/test_host_logic.py). */
HOC_HD float hoc_pix_centre(int i, int S)
{
    return (float)(2 * i + 1 - S) / (float)S;
}

/* Barycentric coefficient matrix (row j gives weight j as an affine function of (xi, yi, 1)). */
HOC_HD void hoc_face_inv(const float *f, int S, float *inv)
{
    const float p0x = hoc_ndc_to_pix(f[0], S), p0y = hoc_ndc_to_pix(f[1], S);
    const float p1x = hoc_ndc_to_pix(f[3], S), p1y = hoc_ndc_to_pix(f[4], S);
    const float p2x = hoc_ndc_to_pix(f[6], S), p2y = hoc_ndc_to_pix(f[7], S);
    const float den = (p2x * (p0y - p1y) + p0x * (p1y - p2y) + p1x * (p2y - p0y));
    inv[0] = (p1y - p2y) / den;
    inv[1] = (p2x - p1x) / den;
    inv[2] = (p1x * p2y - p2x * p1y) / den;
    inv[3] = (p2y - p0y) / den;
    inv[4] = (p0x - p2x) / den;
    inv[5] = (p2x * p0y - p0x * p2y) / den;
    inv[6] = (p0y - p1y) / den;
    inv[7] = (p1x - p0x) / den;
    inv[8] = (p0x * p1y - p1x * p0y) / den;
}

/* Three edge tests on the pixel centre (xp, yp) in NDC; boundary pixels are inside. */
HOC_HD bool hoc_pixel_inside(const float *f, float xp, float yp)
{
    if ((yp - f[1]) * (f[3] - f[0]) < (xp - f[0]) * (f[4] - f[1]))
        return false;
    if ((yp - f[4]) * (f[6] - f[3]) < (xp - f[3]) * (f[7] - f[4]))
        return false;
    if ((yp - f[7]) * (f[0] - f[6]) < (xp - f[6]) * (f[1] - f[7]))
        return false;
    return true;
}

/* Clamped + renormalised barycentric weights and perspective-correct depth of pixel (xi, yi)
 * for a face that passed the inside test.  Returns false when the depth is outside (near, far)
 * (NaN depths return true here and lose every `<` comparison later, like the reference). */
HOC_HD bool hoc_pixel_weights_depth(const float *f, const float *inv, int xi, int yi, float near, float far,
                                    float *w, float *zp_out)
{
    const float fx = (float)xi, fy = (float)yi;
    float w0 = inv[0] * fx + inv[1] * fy + inv[2];
    float w1 = inv[3] * fx + inv[4] * fy + inv[5];
    float w2 = inv[6] * fx + inv[7] * fy + inv[8];
    w0 = fminf(fmaxf(w0, 0.0f), 1.0f);
    w1 = fminf(fmaxf(w1, 0.0f), 1.0f);
    w2 = fminf(fmaxf(w2, 0.0f), 1.0f);
    const float w_sum = w0 + w1 + w2;
    w0 /= w_sum;
    w1 /= w_sum;
    w2 /= w_sum;
    const float s = w0 / f[2] + w1 / f[5] + w2 / f[8];
    const float zp = (float)(1.0 / (double)s);
    w[0] = w0;
    w[1] = w1;
    w[2] = w2;
    *zp_out = zp;
    if (zp <= near || far <= zp)
        return false;
    return true;
}

/* Texture coordinate along cube axis k (k-th vertex): clamp(w_k (ts-1) depth / z_k, 0, ts-1-eps). */
HOC_HD float hoc_tex_coord(float wk, float zk, float depth, int ts, float eps)
{
    float t = wk * (float)(ts - 1) * (depth / zk);
    t = fmaxf(t, 0.0f);
    t = fminf(t, (float)(ts - 1) - eps);
    return t;
}

/* Cell of the texture cube that holds coordinate t.  (int)t like the reference, but never the
 * last texel: with eps == 0 the reference's "+1" tap would fall outside the cube with weight 0;
 * anchoring the cell one texel lower gives the same value (weights 0 and 1) without the
 * out-of-bounds read. */
HOC_HD int hoc_tex_cell(float t, int ts)
{
    int c = (int)t;
    if (c > ts - 2)
        c = ts - 2;
    if (c < 0)
        c = 0;
    return c;
}

/* Order-preserving float -> uint32 map (so that atomicMin on the packed key is a z-test). */
HOC_HD uint32_t hoc_float_order(float f)
{
    union {
        float f;
        uint32_t u;
    } c;
    c.f = f;
    return (c.u & 0x80000000u) ? ~c.u : (c.u | 0x80000000u);
}

/* ---- NMR pseudo-gradient (backward_pixel_map) -------------------------------------------- */

/* One (edge, axis) pair of a face: vertices A=edge, B=edge+1, C=edge+2 in pixel space with the
 * walk coordinate first (axis 0: d0 = x, d1 = y; axis 1: d0 = y, d1 = x). */
struct HocK4Edge {
    float a0, a1, b0, b1, c0, c1;
    int dir;      /* from the inside across AB to the outside, along d1 */
    int d0_from;  /* first integer walk coordinate (inclusive) */
    int d0_to;    /* last one (inclusive); empty when d0_to < d0_from */
};

/* (ax, ay), (bx, by), (cx, cy): the three vertices in NDC, A = first vertex of the edge. */
HOC_HD void hoc_k4_edge_pts(float ax_ndc, float ay_ndc, float bx_ndc, float by_ndc, float cx_ndc, float cy_ndc, int S,
                            int axis, HocK4Edge *E)
{
    const float ax = hoc_ndc_to_pix(ax_ndc, S), ay = hoc_ndc_to_pix(ay_ndc, S);
    const float bx = hoc_ndc_to_pix(bx_ndc, S), by = hoc_ndc_to_pix(by_ndc, S);
    const float cx = hoc_ndc_to_pix(cx_ndc, S), cy = hoc_ndc_to_pix(cy_ndc, S);
    if (axis == 0) {
        E->a0 = ax; E->a1 = ay; E->b0 = bx; E->b1 = by; E->c0 = cx; E->c1 = cy;
        E->dir = (E->a0 < E->b0) ? -1 : 1;
    } else {
        E->a0 = ay; E->a1 = ax; E->b0 = by; E->b1 = bx; E->c0 = cy; E->c1 = cx;
        E->dir = (E->a0 < E->b0) ? 1 : -1;
    }
    const float lo = fmaxf(ceilf(fminf(E->a0, E->b0)), 0.0f);
    const float hi = fminf(fmaxf(E->a0, E->b0), (float)(S - 1));
    /* (int) truncation toward zero like the reference: an edge whose larger end lies in (-1, 0)
     * still visits column 0.  Clamps only keep the conversions in range. */
    E->d0_from = (int)fminf(lo, (float)S);
    E->d0_to = (int)fmaxf(hi, -2.0f);
}

HOC_HD void hoc_k4_edge(const float *f, int S, int edge, int axis, HocK4Edge *E)
{
    const int ia = edge, ib = (edge + 1) % 3, ic = (edge + 2) % 3;
    hoc_k4_edge_pts(f[3 * ia], f[3 * ia + 1], f[3 * ib], f[3 * ib + 1], f[3 * ic], f[3 * ic + 1], S, axis, E);
}

/* Column d0 of an (edge, axis): crossing of AB, the pixel just inside and just outside.
 * Returns false when the column is skipped (either pixel outside the image or the crossing is
 * not a finite number -- the reference's float->int conversion yields INT_MIN there). */
HOC_HD bool hoc_k4_column(const HocK4Edge *E, int S, int d0, float *d1_cross_out, int *d1_in_out, int *d1_out_out)
{
    const float d1_cross = (E->b1 - E->a1) / (E->b0 - E->a0) * ((float)d0 - E->a0) + E->a1;
    *d1_cross_out = d1_cross;
    if (!(d1_cross > -1.0e9f && d1_cross < 1.0e9f))
        return false;
    const int d1_in = (0 < E->dir) ? (int)floorf(d1_cross) : (int)ceilf(d1_cross);
    const int d1_out = d1_in + E->dir;
    *d1_in_out = d1_in;
    *d1_out_out = d1_out;
    if (d1_in < 0 || S <= d1_in)
        return false;
    if (d1_out < 0 || S <= d1_out)
        return false;
    return true;
}

/* Far end (inclusive, un-clamped) of the inward scan of column d0: crossing with AC or CB. */
HOC_HD int hoc_k4_inward_limit(const HocK4Edge *E, int d0)
{
    const float fd0 = (float)d0;
    float cross2;
    if ((fd0 - E->a0) * (fd0 - E->c0) < 0)
        cross2 = (E->c1 - E->a1) / (E->c0 - E->a0) * (fd0 - E->a0) + E->a1;
    else
        cross2 = (E->b1 - E->c1) / (E->b0 - E->c0) * (fd0 - E->c0) + E->c1;
    float lim = (0 < E->dir) ? ceilf(cross2) : floorf(cross2);
    if (!(lim > -2147483000.0f))
        lim = -2147483000.0f;
    if (lim > 2147483000.0f)
        lim = 2147483000.0f;
    return (int)lim;
}

/* -delta/dist contributions of one scanned pixel d1 to vertex A and vertex B of the edge. */
HOC_HD void hoc_k4_accum(const HocK4Edge *E, int S, int d0, int d1, float d1_cross, float eps, float delta,
                         float *gA, float *gB)
{
    const float fd0 = (float)d0;
    const float t = ((float)d1 - d1_cross);
    if (E->b0 != fd0) {
        float dist = (E->b0 - E->a0) / (E->b0 - fd0) * t * 2.0f / (float)S;
        dist = (0 < dist) ? dist + eps : dist - eps;
        *gA -= delta / dist;
    }
    if (E->a0 != fd0) {
        float dist = (E->b0 - E->a0) / (fd0 - E->a0) * t * 2.0f / (float)S;
        dist = (0 < dist) ? dist + eps : dist - eps;
        *gB -= delta / dist;
    }
}

/* Per-column constants of hoc_k4_accum, hoisted out of the scan loops: dist = c * (d1 - d1_cross) * 2 / S
 * with c = (b0-a0)/(b0-d0) for vertex A and (b0-a0)/(d0-a0) for vertex B (a vertex whose walk coordinate
 * equals d0 gets no contribution).  `scale` = 2 / S is applied as one multiply (<= 1 ulp from the
 * reference's `* 2. / is`; gradients carry a 1e-3 tolerance). */
struct HocK4Col {
    float cross, cA, cB, scale;
    bool hasA, hasB;
};

HOC_HD void hoc_k4_col(const HocK4Edge *E, int S, int d0, float d1_cross, HocK4Col *C)
{
    const float fd0 = (float)d0;
    C->cross = d1_cross;
    C->scale = 2.0f / (float)S;
    C->hasA = E->b0 != fd0;
    C->hasB = E->a0 != fd0;
    C->cA = C->hasA ? (E->b0 - E->a0) / (E->b0 - fd0) : 0.0f;
    C->cB = C->hasB ? (E->b0 - E->a0) / (fd0 - E->a0) : 0.0f;
}

/* delta / dist: on the device the approximate (2 ulp) divide -- the pseudo-gradient carries a 1e-3 tolerance and
 * this quotient is evaluated once per scanned pixel per vertex, which makes the IEEE divide the hot instruction. */
#ifdef __CUDA_ARCH__
#define HOC_FAST_DIV(a, b) __fdividef((a), (b))
#else
#define HOC_FAST_DIV(a, b) ((a) / (b))
#endif

HOC_HD void hoc_k4_accum_col(const HocK4Col *C, int d1, float eps, float delta, float *gA, float *gB)
{
    const float t = ((float)d1 - C->cross) * C->scale;
    if (C->hasA) {
        float dist = C->cA * t;
        dist = (0 < dist) ? dist + eps : dist - eps;
        *gA -= HOC_FAST_DIV(delta, dist);
    }
    if (C->hasB) {
        float dist = C->cB * t;
        dist = (0 < dist) ? dist + eps : dist - eps;
        *gB -= HOC_FAST_DIV(delta, dist);
    }
}
