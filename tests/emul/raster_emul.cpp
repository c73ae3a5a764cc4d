/*
 * tests/emul/raster_emul.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Serial host emulation of the ALGORITHM of csrc/raster_fwd.cu and csrc/raster_bwd.cu: same
 * per-face / per-pixel functions (it includes the product header raster_math.h), same work
 * decomposition (face-parallel z-buffer with a packed (depth, face) min key, pixel resolve,
 * face-owned gradient gather, pixel-driven inward terms, per-line outward scans clipped to the span of the
 * pixels that matter and summed with the line pass's arithmetic), with loops where the kernels have lanes.
 * It lets the CPU test-suite check the restructured algorithm against the oracle's literal
 * restatement of the reference without a GPU.  It is never loaded by the product.
 * Build: g++ -O2 -ffp-contract=off -shared -fPIC (tests/test_host_emulation.py).
 */
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../handobjectconsist_b200/csrc/raster_math.h"

static bool face_bbox(const float *f, int S, int *x0, int *y0, int *bw, int *bh)
{
    const float pxmin = hoc_ndc_to_pix(fminf(f[0], fminf(f[3], f[6])), S);
    const float pxmax = hoc_ndc_to_pix(fmaxf(f[0], fmaxf(f[3], f[6])), S);
    const float pymin = hoc_ndc_to_pix(fminf(f[1], fminf(f[4], f[7])), S);
    const float pymax = hoc_ndc_to_pix(fmaxf(f[1], fmaxf(f[4], f[7])), S);
    const float fS1 = (float)(S - 1);
    const float x_lo = fmaxf(ceilf(pxmin - 0.5f), 0.0f);
    const float x_hi = fminf(floorf(pxmax + 0.5f), fS1);
    const float y_lo = fmaxf(ceilf(pymin - 0.5f), 0.0f);
    const float y_hi = fminf(floorf(pymax + 0.5f), fS1);
    if (!(x_lo <= x_hi && y_lo <= y_hi))
        return false;
    *x0 = (int)x_lo;
    *y0 = (int)y_lo;
    *bw = (int)(x_hi - x_lo + 1.0f);
    *bh = (int)(y_hi - y_lo + 1.0f);
    return true;
}

static inline long plane_off(int layout, int S, int b, int yi, int xi)
{
    const int row = layout ? (S - 1 - yi) : yi;
    return ((long)b * S + row) * S + xi;
}
static inline long rgb_off(int layout, int S, int b, int yi, int xi, int c)
{
    if (layout)
        return (((long)b * 3 + c) * S + (S - 1 - yi)) * S + xi;
    return (((long)b * S + yi) * S + xi) * 3 + c;
}

extern "C" void emul_raster_forward(const float *faces, const float *textures, int B, int F, int S, int ts, float near_,
                                    float far_, float eps, const float *bg, int layout, float *rgb, float *alpha,
                                    float *depth, int32_t *face_index_map, float *weight_map, float *face_inv_map)
{
    std::vector<uint64_t> zbuf((size_t)B * S * S, ~0ull);
    std::vector<float> centre(S);
    for (int i = 0; i < S; i++)
        centre[i] = hoc_pix_centre(i, S);
    for (int b = 0; b < B; b++)
        for (int fi = 0; fi < F; fi++) {
            const float *f = faces + ((long)b * F + fi) * 9;
            if (!hoc_face_xy_finite(f) || hoc_face_back(f))
                continue;
            int x0, y0, bw, bh;
            if (!face_bbox(f, S, &x0, &y0, &bw, &bh))
                continue;
            float inv[9];
            hoc_face_inv(f, S, inv);
            for (int p = 0; p < bw * bh; p++) {
                const int yy = p / bw, xi = x0 + (p - yy * bw), yi = y0 + yy;
                if (!hoc_pixel_inside(f, centre[xi], centre[yi]))
                    continue;
                float w[3], zp;
                if (!hoc_pixel_weights_depth(f, inv, xi, yi, near_, far_, w, &zp))
                    continue;
                if (!(zp < far_))
                    continue;
                const uint64_t key = ((uint64_t)hoc_float_order(zp) << 32) | (uint32_t)fi;
                uint64_t &z = zbuf[((size_t)b * S + yi) * S + xi];
                z = std::min(z, key);
            }
        }
    for (int b = 0; b < B; b++)
        for (int yi = 0; yi < S; yi++)
            for (int xi = 0; xi < S; xi++) {
                const long pix = ((long)b * S + yi) * S + xi;
                const int fidx = (int)(uint32_t)(zbuf[pix] & 0xffffffffull);
                float w[3] = {0, 0, 0}, inv[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, zp = far_;
                float col[3] = {bg[0], bg[1], bg[2]};
                if (fidx >= 0) {
                    const float *f = faces + ((long)b * F + fidx) * 9;
                    hoc_face_inv(f, S, inv);
                    hoc_pixel_weights_depth(f, inv, xi, yi, near_, far_, w, &zp);
                    if (rgb) {
                        const float *tex = textures + ((long)b * F + fidx) * ts * ts * ts * 3;
                        float tf[3];
                        int ti[3];
                        for (int k = 0; k < 3; k++) {
                            const float t = hoc_tex_coord(w[k], f[3 * k + 2], zp, ts, eps);
                            ti[k] = hoc_tex_cell(t, ts);
                            tf[k] = t - (float)ti[k];
                        }
                        float acc[3] = {0, 0, 0};
                        for (int pn = 0; pn < 8; pn++) {
                            float ww = 1.0f;
                            int isc = 0;
                            for (int k = 0; k < 3; k++) {
                                if (((pn >> k) & 1) == 0) {
                                    ww *= 1.0f - tf[k];
                                    isc = isc * ts + ti[k];
                                } else {
                                    ww *= tf[k];
                                    isc = isc * ts + ti[k] + 1;
                                }
                            }
                            if (ts == 1)
                                isc = 0;
                            for (int c = 0; c < 3; c++)
                                acc[c] += ww * tex[isc * 3 + c];
                        }
                        for (int c = 0; c < 3; c++)
                            col[c] = acc[c];
                    }
                }
                face_index_map[pix] = fidx;
                if (alpha)
                    alpha[plane_off(layout, S, b, yi, xi)] = fidx >= 0 ? 1.0f : 0.0f;
                if (depth)
                    depth[plane_off(layout, S, b, yi, xi)] = zp;
                if (rgb)
                    for (int c = 0; c < 3; c++)
                        rgb[rgb_off(layout, S, b, yi, xi, c)] = col[c];
                if (weight_map)
                    for (int k = 0; k < 3; k++)
                        weight_map[pix * 3 + k] = w[k];
                if (face_inv_map)
                    for (int k = 0; k < 9; k++)
                        face_inv_map[pix * 9 + k] = inv[k];
            }
}

struct Maps {
    const int32_t *idx;
    const float *rgb, *g_rgb, *g_alpha;
    int S, layout, b;
    bool use_alpha, use_rgb;
};

static void load_I(const Maps &M, int xi, int yi, float *I)
{
    I[0] = I[1] = I[2] = I[3] = 0.0f;
    if (M.use_alpha)
        I[0] = M.idx[(long)yi * M.S + xi] >= 0 ? 1.0f : 0.0f;
    if (M.use_rgb)
        for (int k = 0; k < 3; k++)
            I[1 + k] = M.rgb[rgb_off(M.layout, M.S, M.b, yi, xi, k)];
}

static float delta_at(const Maps &M, int xi, int yi, const float *Iref)
{
    float d = 0.0f;
    if (M.use_alpha) {
        const float a = M.idx[(long)yi * M.S + xi] >= 0 ? 1.0f : 0.0f;
        d += (a - Iref[0]) * M.g_alpha[plane_off(M.layout, M.S, M.b, yi, xi)];
    }
    if (M.use_rgb)
        for (int k = 0; k < 3; k++) {
            const long o = rgb_off(M.layout, M.S, M.b, yi, xi, k);
            d += (M.rgb[o] - Iref[1 + k]) * M.g_rgb[o];
        }
    return d;
}

/* Outward scans whose pixels did NOT all have the sign of c * (d1 - cross) the line pass assumes for the whole scan
 * (it adds the reference's +-eps as one constant per scan): must stay 0. */
static long g_sign_violations = 0;
extern "C" long emul_sign_violations() { return g_sign_violations; }

extern "C" void emul_raster_backward(const float *faces, const int32_t *face_index_map, const float *rgb,
                                     const float *g_rgb, const float *g_alpha, const float *g_depth, int B, int F,
                                     int S, int ts, float near_, float far_, float eps, int layout, int use_alpha,
                                     float *grad_faces, float *grad_textures)
{
    const int tex_n = ts * ts * ts * 3;
    /* extent pre-pass */
    std::vector<int> ext((size_t)B * 4 * S);
    for (int b = 0; b < B; b++) {
        int *e = ext.data() + (size_t)b * 4 * S;
        for (int i = 0; i < S; i++) {
            e[0 * S + i] = e[2 * S + i] = 0x7f7f7f7f;
            e[1 * S + i] = e[3 * S + i] = -1;
        }
        for (int yi = 0; yi < S; yi++)
            for (int xi = 0; xi < S; xi++) {
                bool nz = false;
                if (g_rgb)
                    for (int c = 0; c < 3; c++)
                        nz = nz || !(g_rgb[rgb_off(layout, S, b, yi, xi, c)] == 0.0f);
                if (g_alpha && use_alpha)
                    nz = nz || !(g_alpha[plane_off(layout, S, b, yi, xi)] == 0.0f);
                /* the line pass looks at the covered pixels (they own the scans) and at those with a gradient */
                if (nz || face_index_map[((long)b * S + yi) * S + xi] >= 0) {
                    e[0 * S + yi] = std::min(e[0 * S + yi], xi);
                    e[1 * S + yi] = std::max(e[1 * S + yi], xi);
                    e[2 * S + xi] = std::min(e[2 * S + xi], yi);
                    e[3 * S + xi] = std::max(e[3 * S + xi], yi);
                }
            }
    }
    for (int b = 0; b < B; b++)
        for (int fi = 0; fi < F; fi++) {
            const float *f = faces + ((long)b * F + fi) * 9;
            float *gt = grad_textures ? grad_textures + ((long)b * F + fi) * tex_n : nullptr;
            float *gf = grad_faces ? grad_faces + ((long)b * F + fi) * 9 : nullptr;
            if (gf)
                for (int k = 0; k < 9; k++)
                    gf[k] = 0.0f;
            if (gt)
                for (int k = 0; k < tex_n; k++)
                    gt[k] = 0.0f;
            if (!hoc_face_xy_finite(f) || hoc_face_back(f))
                continue;
            const int32_t *idx = face_index_map + (long)b * S * S;
            float inv[9];
            hoc_face_inv(f, S, inv);
            const bool want_tex = gt && g_rgb, want_depth = gf && g_depth;
            float acc_d[3] = {0, 0, 0};
            bool any_hit = false;
            int x0, y0, bw, bh;
            if ((want_tex || want_depth) && face_bbox(f, S, &x0, &y0, &bw, &bh)) {
                for (int p = 0; p < bw * bh; p++) {
                    const int yy = p / bw, xi = x0 + (p - yy * bw), yi = y0 + yy;
                    if (idx[(long)yi * S + xi] != fi)
                        continue;
                    any_hit = true;
                    float w[3], zp;
                    hoc_pixel_weights_depth(f, inv, xi, yi, near_, far_, w, &zp);
                    if (want_depth) {
                        const float gz = g_depth[plane_off(layout, S, b, yi, xi)] * zp * zp;
                        for (int k = 0; k < 3; k++)
                            acc_d[k] += gz * w[k];
                    }
                    if (want_tex) {
                        float tf[3];
                        int ti[3];
                        for (int k = 0; k < 3; k++) {
                            const float t = hoc_tex_coord(w[k], f[3 * k + 2], zp, ts, eps);
                            ti[k] = hoc_tex_cell(t, ts);
                            tf[k] = t - (float)ti[k];
                        }
                        for (int pn = 0; pn < 8; pn++) {
                            float ww = 1.0f;
                            int isc = 0;
                            for (int k = 0; k < 3; k++) {
                                if (((pn >> k) & 1) == 0) {
                                    ww *= 1.0f - tf[k];
                                    isc = isc * ts + ti[k];
                                } else {
                                    ww *= tf[k];
                                    isc = isc * ts + ti[k] + 1;
                                }
                            }
                            if (ts == 1)
                                isc = 0;
                            for (int c = 0; c < 3; c++)
                                gt[isc * 3 + c] += ww * g_rgb[rgb_off(layout, S, b, yi, xi, c)];
                        }
                    }
                }
            }
            if (!gf)
                continue;
            if (want_depth && any_hit) {
                float tmp[2];
                for (int l = 0; l < 2; l++)
                    tmp[l] = inv[l] / f[2] + inv[3 + l] / f[5] + inv[6 + l] / f[8];
                for (int k = 0; k < 3; k++) {
                    const float zk = f[3 * k + 2];
                    gf[3 * k + 2] = acc_d[k] / (zk * zk);
                    gf[3 * k + 0] = acc_d[k] * tmp[0] * (float)S / 2.0f;
                    gf[3 * k + 1] = acc_d[k] * tmp[1] * (float)S / 2.0f;
                }
            }
        }
    if (!grad_faces)
        return;
    /* Pseudo-gradient, decomposed like the kernels: (1) from every covered pixel, its own term of the inward scan
     * of each (edge, axis) column it lies on, plus a flag byte when it is the pixel just inside the edge;
     * (2) per line, the flagged outward scans clipped to the span of non-zero incoming gradient. */
    std::vector<uint8_t> flags((size_t)B * 2 * S * S, 0);
    for (int b = 0; b < B; b++) {
        const int32_t *idx = face_index_map + (long)b * S * S;
        Maps M = {idx, rgb, g_rgb, g_alpha, S, layout, b, use_alpha && g_alpha, rgb && g_rgb};
        if (!(M.use_alpha || M.use_rgb))
            continue;
        for (int yi = 0; yi < S; yi++)
            for (int xi = 0; xi < S; xi++) {
                const int fi = idx[(long)yi * S + xi];
                if (fi < 0)
                    continue;
                const float *f = faces + ((long)b * F + fi) * 9;
                if (!hoc_face_xy_finite(f) || hoc_face_back(f))
                    continue;
                float *gf = grad_faces + ((long)b * F + fi) * 9;
                for (int combo = 0; combo < 6; combo++) {
                    const int edge = combo >> 1, axis = combo & 1;
                    HocK4Edge E;
                    hoc_k4_edge(f, S, edge, axis, &E);
                    const int d0 = axis == 0 ? xi : yi, d1p = axis == 0 ? yi : xi;
                    if (d0 < E.d0_from || d0 > E.d0_to)
                        continue;
                    float d1_cross;
                    int d1_in, d1_out;
                    if (!hoc_k4_column(&E, S, d0, &d1_cross, &d1_in, &d1_out))
                        continue;
                    if (d1_in == d1p)
                        flags[(((size_t)b * 2 + axis) * S + d0) * S + d1p] |=
                            (uint8_t)((1u << edge) | ((0 < E.dir) ? (8u << edge) : 0u));
                    const int lim = hoc_k4_inward_limit(&E, d0);
                    const int d1_from = std::max(std::min(d1_in, lim), 0);
                    const int d1_to = std::min(std::max(d1_in, lim), S - 1);
                    if (d1p < d1_from || d1p > d1_to)
                        continue;
                    float I_out[4];
                    load_I(M, axis == 0 ? d0 : d1_out, axis == 0 ? d1_out : d0, I_out);
                    const float delta = delta_at(M, xi, yi, I_out);
                    if (delta <= 0.0f)
                        continue;
                    float gA = 0.0f, gB = 0.0f;
                    hoc_k4_accum(&E, S, d0, d1p, d1_cross, eps, delta, &gA, &gB);
                    gf[edge * 3 + (1 - axis)] += gA;
                    gf[((edge + 1) % 3) * 3 + (1 - axis)] += gB;
                }
            }
        const int *e = ext.data() + (size_t)b * 4 * S;
        for (int axis = 0; axis < 2; axis++)
            for (int d0 = 0; d0 < S; d0++) {
                const int lo = axis == 0 ? e[2 * S + d0] : e[0 * S + d0];
                const int hi = axis == 0 ? e[3 * S + d0] : e[1 * S + d0];
                if (lo > hi)
                    continue;
                const uint8_t *fl = flags.data() + (((size_t)b * 2 + axis) * S + d0) * S;
                for (int i = 0; i < S; i++)
                    for (int edge = 0; edge < 3; edge++) {
                        if (!(fl[i] & (1u << edge)))
                            continue;
                        const bool pos = fl[i] & (8u << edge);
                        if (pos ? (i + 1 > hi) : (i - 1 < lo))
                            continue;
                        const int xin = axis == 0 ? d0 : i, yin = axis == 0 ? i : d0;
                        const int fi = idx[(long)yin * S + xin];
                        const float *f = faces + ((long)b * F + fi) * 9;
                        HocK4Edge E;
                        hoc_k4_edge(f, S, edge, axis, &E);
                        float d1_cross;
                        int d1_in, d1_out;
                        if (!hoc_k4_column(&E, S, d0, &d1_cross, &d1_in, &d1_out) || d1_in != i || (0 < E.dir) != pos)
                            continue; /* unreachable */
                        float I_in[4];
                        load_I(M, xin, yin, I_in);
                        const int d1_from = pos ? std::max(d1_out, lo) : lo;
                        const int d1_to = pos ? hi : std::min(d1_out, hi);
                        /* the arithmetic of hoc_raster_bwd_line_kernel's chunk loop: per-scan constants, the reference's
                         * +-eps as ONE constant per scan (sign of c * (d1_out - cross)), both distances by FMA, one
                         * reciprocal of their product for both vertices */
                        HocK4Col C;
                        hoc_k4_col(&E, S, d0, d1_cross, &C);
                        const float scale = 2.0f / (float)S, t_first = (float)d1_out - d1_cross;
                        float cA = 0.0f, cB = 0.0f, eA = 1.0f, eB = 1.0f;
                        if (C.hasA) {
                            cA = C.cA * scale;
                            eA = (0.0f < cA * t_first) ? eps : -eps;
                        }
                        if (C.hasB) {
                            cB = C.cB * scale;
                            eB = (0.0f < cB * t_first) ? eps : -eps;
                        }
                        float gA = 0.0f, gB = 0.0f;
                        for (int d1 = d1_from; d1 <= d1_to; d1++) {
                            const float u = (float)d1 - d1_cross;
                            if ((C.hasA && ((0.0f < C.cA * (u * scale)) ? eps : -eps) != eA) ||
                                (C.hasB && ((0.0f < C.cB * (u * scale)) ? eps : -eps) != eB))
                                g_sign_violations++;
                            const float delta = delta_at(M, axis == 0 ? d0 : d1, axis == 0 ? d1 : d0, I_in);
                            if (delta <= 0.0f)
                                continue;
                            const float dA = fmaf(cA, u, eA), dB = fmaf(cB, u, eB);
                            const float t = delta * (1.0f / (dA * dB));
                            if (C.hasA)
                                gA -= t * dB;
                            if (C.hasB)
                                gB -= t * dA;
                        }
                        float *gf = grad_faces + ((long)b * F + fi) * 9;
                        gf[edge * 3 + (1 - axis)] += gA;
                        gf[((edge + 1) % 3) * 3 + (1 - axis)] += gB;
                    }
            }
    }
}
