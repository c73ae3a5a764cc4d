/*
 * oracle/nmr_oracle_impl.h -- TEST INFRASTRUCTURE ONLY (never imported by the product path).
 *
 * CPU restatement of the five entry points of `neural_renderer.cuda.rasterize`
 * that /root/reference/meshreg/neurender/rasterize.py calls
 * (forward_face_index_map :202, forward_texture_sampling :232, backward_pixel_map :269,
 *  backward_textures :290, backward_depth_map :306).
 *
 * PARITY UNPINNED: the arithmetic lives in the third-party wheel `neural-renderer-pytorch`
 * (unpinned in /root/reference/environment.yml:36, upstream daniilidis-group/neural_renderer),
 * which is absent from /root/reference and not installable here.  This file restates the
 * published algorithm (Kato et al., Neural 3D Mesh Renderer, CVPR 2018, and the upstream
 * kernel structure summarised in SURVEY.md section 2.3 / Appendix C): one pass per pixel over
 * ALL faces for the forward, one serial scan per face for the pseudo-gradient.  The reference
 * holds no golden vectors for this path; analytic unit scenes in tests/ pin the conventions.
 *
 * This header is included twice by nmr_oracle.c with REAL = float / double.  The float
 * instantiation mirrors the reference's float arithmetic including its promotions of the
 * literals `2.`, `0.5`, `1.` to double (compile with -ffp-contract=off).
 */

#ifndef REAL
#error "define REAL and FN(name) before including"
#endif
/* The per-pixel / per-face bodies are plain functions (FN(fwd_pixel), FN(tex_pixel), FN(k4_face), FN(bt_pixel),
 * FN(bd_pixel)) -- one call = what ONE thread of the reference's kernels does -- driven here by host loops / a
 * pthread pool.  ORA_HD / ORA_ATOMIC_ADD / ORA_NO_HOST_DRIVERS let another driver reuse the same bodies. */
#ifndef ORA_HD
#define ORA_HD static inline
#endif
#ifndef ORA_ATOMIC_ADD
#define ORA_ATOMIC_ADD(ptr, val) (*(ptr) += (val))
#endif

/* K1: barycentric coefficient matrix of one face in pixel-index space (Appendix C1). */
ORA_HD int FN(face_inv)(const REAL *face, int is, REAL *face_inv)
{
    /* back-face rule, NDC y-up (Appendix A) */
    if ((face[7] - face[1]) * (face[3] - face[0]) < (face[4] - face[1]) * (face[6] - face[0]))
        return 0;
    REAL p[3][2];
    for (int num = 0; num < 3; num++)
        for (int dim = 0; dim < 2; dim++)
            p[num][dim] = (REAL)(0.5 * (face[3 * num + dim] * is + is - 1));
    REAL inv[9] = {
        p[1][1] - p[2][1], p[2][0] - p[1][0], p[1][0] * p[2][1] - p[2][0] * p[1][1],
        p[2][1] - p[0][1], p[0][0] - p[2][0], p[2][0] * p[0][1] - p[0][0] * p[2][1],
        p[0][1] - p[1][1], p[1][0] - p[0][0], p[0][0] * p[1][1] - p[1][0] * p[0][1]};
    REAL den = (p[2][0] * (p[0][1] - p[1][1]) + p[0][0] * (p[1][1] - p[2][1]) + p[1][0] * (p[2][1] - p[0][1]));
    for (int k = 0; k < 9; k++)
        face_inv[k] = inv[k] / den;
    return 1;
}

/* K1 + K2: rasterizer forward, rasterize.py:199-215.  Buffers must be pre-initialised by the
 * caller exactly as rasterize.py:58-85 does (face_index_map=-1, weight_map=0, depth_map=far,
 * face_inv_map=0).  The per-pixel loop over ALL faces is split over host threads by pixel range. */
struct FN(fwd_ctx) {
    const REAL *faces;
    int32_t *face_index_map;
    REAL *weight_map, *depth_map, *face_inv_map;
    const REAL *faces_inv;
    int nf, is;
    REAL near, far;
    int return_depth;
};

ORA_HD void FN(fwd_pixel)(const struct FN(fwd_ctx) *c, long i)
{
    const REAL *faces = c->faces, *faces_inv = c->faces_inv;
    const int is = c->is, nf = c->nf;
    const REAL near = c->near, far = c->far;
    {
        const int bn = (int)(i / (is * is));
        const int pn = (int)(i % (is * is));
        const int yi = pn / is;
        const int xi = pn % is;
        const REAL yp = (REAL)((2. * yi + 1 - is) / is);
        const REAL xp = (REAL)((2. * xi + 1 - is) / is);
        REAL depth_min = far;
        int face_index_min = -1;
        REAL weight_min[3] = {0, 0, 0};
        REAL face_inv_min[9] = {0};
        for (int fn = 0; fn < nf; fn++) {
            const REAL *face = &faces[((long)bn * nf + fn) * 9];
            const REAL *finv = &faces_inv[((long)bn * nf + fn) * 9];
            if ((face[7] - face[1]) * (face[3] - face[0]) < (face[4] - face[1]) * (face[6] - face[0]))
                continue;
            if (((yp - face[1]) * (face[3] - face[0]) < (xp - face[0]) * (face[4] - face[1])) ||
                ((yp - face[4]) * (face[6] - face[3]) < (xp - face[3]) * (face[7] - face[4])) ||
                ((yp - face[7]) * (face[0] - face[6]) < (xp - face[6]) * (face[1] - face[7])))
                continue;
            REAL w[3];
            w[0] = finv[0] * xi + finv[1] * yi + finv[2];
            w[1] = finv[3] * xi + finv[4] * yi + finv[5];
            w[2] = finv[6] * xi + finv[7] * yi + finv[8];
            REAL w_sum = 0;
            for (int k = 0; k < 3; k++) {
                w[k] = (REAL)fmin(fmax((double)w[k], 0.), 1.);
                w_sum += w[k];
            }
            for (int k = 0; k < 3; k++)
                w[k] /= w_sum;
            const REAL zp = (REAL)(1. / (w[0] / face[2] + w[1] / face[5] + w[2] / face[8]));
            if (zp <= near || far <= zp)
                continue;
            if (zp < depth_min) {
                depth_min = zp;
                face_index_min = fn;
                for (int k = 0; k < 3; k++)
                    weight_min[k] = w[k];
                for (int k = 0; k < 9; k++)
                    face_inv_min[k] = finv[k];
            }
        }
        if (0 <= face_index_min) {
            c->depth_map[i] = depth_min;
            c->face_index_map[i] = face_index_min;
            for (int k = 0; k < 3; k++)
                c->weight_map[3 * i + k] = weight_min[k];
            if (c->return_depth)
                for (int k = 0; k < 9; k++)
                    c->face_inv_map[9 * i + k] = face_inv_min[k];
        }
    }
}

#ifndef ORA_NO_HOST_DRIVERS
static void FN(fwd_range)(void *vc, long begin, long end)
{
    for (long i = begin; i < end; i++)
        FN(fwd_pixel)((const struct FN(fwd_ctx) *)vc, i);
}

void FN(forward_face_index_map)(const REAL *faces, int32_t *face_index_map, REAL *weight_map, REAL *depth_map,
                                REAL *face_inv_map, REAL *faces_inv, int batch_size, int num_faces, int image_size,
                                REAL near, REAL far, int return_depth)
{
    const int is = image_size, nf = num_faces;
    memset(faces_inv, 0, sizeof(REAL) * (size_t)batch_size * nf * 9);
    for (long i = 0; i < (long)batch_size * nf; i++)
        FN(face_inv)(&faces[i * 9], is, &faces_inv[i * 9]);
    struct FN(fwd_ctx) c = {faces, face_index_map, weight_map, depth_map, face_inv_map, faces_inv, nf, is, near, far,
                            return_depth};
    ora_parallel_for((long)batch_size * is * is, FN(fwd_range), &c);
}
#endif

/* K3: trilinear sampling of the per-face texture cube, rasterize.py:218-243 (Appendix C3); one pixel. */
ORA_HD void FN(tex_pixel)(const REAL *faces, const REAL *textures, const int32_t *face_index_map,
                          const REAL *weight_map, const REAL *depth_map, REAL *rgb_map, int32_t *sampling_index_map,
                          REAL *sampling_weight_map, int num_faces, int image_size, int texture_size, REAL eps, long i)
{
    const int is = image_size, nf = num_faces, ts = texture_size;
    {
        const int face_index = face_index_map[i];
        if (face_index < 0)
            return;
        const int bn = (int)(i / (is * is));
        const REAL *face = &faces[((long)bn * nf + face_index) * 9];
        const REAL *texture = &textures[((long)bn * nf + face_index) * ts * ts * ts * 3];
        const REAL *weight = &weight_map[i * 3];
        const REAL depth = depth_map[i];
        REAL tif[3];
        for (int k = 0; k < 3; k++) {
            REAL t = weight[k] * (ts - 1) * (depth / face[3 * k + 2]);
            t = (REAL)fmax((double)t, 0.);
            t = (REAL)fmin((double)t, (double)(ts - 1 - eps));
            tif[k] = t;
        }
        REAL new_pixel[3] = {0, 0, 0};
        for (int pn = 0; pn < 8; pn++) {
            REAL w = 1;
            int tii[3];
            for (int k = 0; k < 3; k++) {
                if (((pn >> k) % 2) == 0) {
                    w *= 1 - (tif[k] - (int)tif[k]);
                    tii[k] = (int)tif[k];
                } else {
                    w *= tif[k] - (int)tif[k];
                    tii[k] = (int)tif[k] + 1;
                }
            }
            const int isc = tii[0] * ts * ts + tii[1] * ts + tii[2];
            for (int k = 0; k < 3; k++)
                new_pixel[k] += w * texture[isc * 3 + k];
            if (sampling_index_map)
                sampling_index_map[i * 8 + pn] = isc;
            if (sampling_weight_map)
                sampling_weight_map[i * 8 + pn] = w;
        }
        for (int k = 0; k < 3; k++)
            rgb_map[i * 3 + k] = new_pixel[k];
    }
}

#ifndef ORA_NO_HOST_DRIVERS
void FN(forward_texture_sampling)(const REAL *faces, const REAL *textures, const int32_t *face_index_map,
                                  const REAL *weight_map, const REAL *depth_map, REAL *rgb_map,
                                  int32_t *sampling_index_map, REAL *sampling_weight_map, int batch_size,
                                  int num_faces, int image_size, int texture_size, REAL eps)
{
    for (long i = 0; i < (long)batch_size * image_size * image_size; i++)
        FN(tex_pixel)(faces, textures, face_index_map, weight_map, depth_map, rgb_map, sampling_index_map,
                      sampling_weight_map, num_faces, image_size, texture_size, eps, i);
}
#endif

/* K4: NMR pseudo-gradient of rgb/alpha w.r.t. the xy of the face vertices, rasterize.py:263-281
 * (Appendix C4).  One serial scan per face; grad_faces[b,f] is OVERWRITTEN for front faces. */
struct FN(k4_ctx) {
    const REAL *faces;
    const int32_t *face_index_map;
    const REAL *rgb_map, *alpha_map, *grad_rgb_map, *grad_alpha_map;
    REAL *grad_faces;
    int num_faces, image_size;
    REAL eps;
    int return_rgb, return_alpha;
};

ORA_HD void FN(k4_face)(const struct FN(k4_ctx) *c, long i)
{
    const REAL *faces = c->faces;
    const int32_t *face_index_map = c->face_index_map;
    const REAL *rgb_map = c->rgb_map, *alpha_map = c->alpha_map;
    const REAL *grad_rgb_map = c->grad_rgb_map, *grad_alpha_map = c->grad_alpha_map;
    REAL *grad_faces = c->grad_faces;
    const int num_faces = c->num_faces, is = c->image_size;
    const REAL eps = c->eps;
    const int return_rgb = c->return_rgb, return_alpha = c->return_alpha;
    {
        const int bn = (int)(i / num_faces);
        const int fn = (int)(i % num_faces);
        const REAL *face = &faces[i * 9];
        REAL grad_face[9] = {0};
        if ((face[7] - face[1]) * (face[3] - face[0]) < (face[4] - face[1]) * (face[6] - face[0]))
            return;
        for (int edge_num = 0; edge_num < 3; edge_num++) {
            int pi[3];
            REAL pp[3][2];
            for (int num = 0; num < 3; num++)
                pi[num] = (edge_num + num) % 3;
            for (int num = 0; num < 3; num++)
                for (int dim = 0; dim < 2; dim++)
                    pp[num][dim] = (REAL)(0.5 * (face[3 * pi[num] + dim] * is + is - 1));
            for (int axis = 0; axis < 2; axis++) {
                REAL p[3][2];
                for (int num = 0; num < 3; num++)
                    for (int dim = 0; dim < 2; dim++)
                        p[num][dim] = pp[num][(dim + axis) % 2];
                int direction;
                if (axis == 0)
                    direction = (p[0][0] < p[1][0]) ? -1 : 1;
                else
                    direction = (p[0][0] < p[1][0]) ? 1 : -1;
                const int d0_from = (int)fmax(ceil(fmin((double)p[0][0], (double)p[1][0])), 0.);
                const int d0_to = (int)fmin(fmax((double)p[0][0], (double)p[1][0]), is - 1.);
                for (int d0 = d0_from; d0 <= d0_to; d0++) {
                    int d1_in, d1_out;
                    const REAL d1_cross = (p[1][1] - p[0][1]) / (p[1][0] - p[0][0]) * (d0 - p[0][0]) + p[0][1];
                    if (0 < direction)
                        d1_in = (int)floor((double)d1_cross);
                    else
                        d1_in = (int)ceil((double)d1_cross);
                    d1_out = d1_in + direction;
                    if (d1_in < 0 || is <= d1_in)
                        continue;
                    if (d1_out < 0 || is <= d1_out)
                        continue;
                    REAL alpha_in = 0, alpha_out = 0;
                    const REAL *rgb_in = 0, *rgb_out = 0;
                    long map_index_in, map_index_out;
                    if (axis == 0) {
                        map_index_in = (long)bn * is * is + (long)d1_in * is + d0;
                        map_index_out = (long)bn * is * is + (long)d1_out * is + d0;
                    } else {
                        map_index_in = (long)bn * is * is + (long)d0 * is + d1_in;
                        map_index_out = (long)bn * is * is + (long)d0 * is + d1_out;
                    }
                    if (return_alpha) {
                        alpha_in = alpha_map[map_index_in];
                        alpha_out = alpha_map[map_index_out];
                    }
                    if (return_rgb) {
                        rgb_in = &rgb_map[map_index_in * 3];
                        rgb_out = &rgb_map[map_index_out * 3];
                    }
                    /* outward pass */
                    if (face_index_map[map_index_in] == fn) {
                        const int d1_limit = (0 < direction) ? is - 1 : 0;
                        const int d1_from = ORA_IMAX(ORA_IMIN(d1_out, d1_limit), 0);
                        const int d1_to = ORA_IMIN(ORA_IMAX(d1_out, d1_limit), is - 1);
                        for (int d1 = d1_from; d1 <= d1_to; d1++) {
                            const long mi = (axis == 0) ? (long)bn * is * is + (long)d1 * is + d0
                                                        : (long)bn * is * is + (long)d0 * is + d1;
                            REAL diff_grad = 0;
                            if (return_alpha)
                                diff_grad += (alpha_map[mi] - alpha_in) * grad_alpha_map[mi];
                            if (return_rgb)
                                for (int k = 0; k < 3; k++)
                                    diff_grad += (rgb_map[mi * 3 + k] - rgb_in[k]) * grad_rgb_map[mi * 3 + k];
                            if (diff_grad <= 0)
                                continue;
                            if (p[1][0] != d0) {
                                REAL dist = (REAL)((p[1][0] - p[0][0]) / (p[1][0] - d0) * (d1 - d1_cross) * 2. / is);
                                dist = (0 < dist) ? dist + eps : dist - eps;
                                grad_face[pi[0] * 3 + (1 - axis)] -= diff_grad / dist;
                            }
                            if (p[0][0] != d0) {
                                REAL dist = (REAL)((p[1][0] - p[0][0]) / (d0 - p[0][0]) * (d1 - d1_cross) * 2. / is);
                                dist = (0 < dist) ? dist + eps : dist - eps;
                                grad_face[pi[1] * 3 + (1 - axis)] -= diff_grad / dist;
                            }
                        }
                    }
                    /* inward pass */
                    {
                        int d1_limit;
                        REAL d0_cross2;
                        if ((d0 - p[0][0]) * (d0 - p[2][0]) < 0)
                            d0_cross2 = (p[2][1] - p[0][1]) / (p[2][0] - p[0][0]) * (d0 - p[0][0]) + p[0][1];
                        else
                            d0_cross2 = (p[1][1] - p[2][1]) / (p[1][0] - p[2][0]) * (d0 - p[2][0]) + p[2][1];
                        /* (int) of a NaN/inf crossing is implementation-defined upstream; clamp explicitly */
                        double lim = (0 < direction) ? ceil((double)d0_cross2) : floor((double)d0_cross2);
                        if (!(lim > -2147483000.0)) lim = -2147483000.0;
                        if (lim > 2147483000.0) lim = 2147483000.0;
                        d1_limit = (int)lim;
                        const int d1_from = ORA_IMAX(ORA_IMIN(d1_in, d1_limit), 0);
                        const int d1_to = ORA_IMIN(ORA_IMAX(d1_in, d1_limit), is - 1);
                        for (int d1 = d1_from; d1 <= d1_to; d1++) {
                            const long mi = (axis == 0) ? (long)bn * is * is + (long)d1 * is + d0
                                                        : (long)bn * is * is + (long)d0 * is + d1;
                            if (face_index_map[mi] != fn)
                                continue;
                            REAL diff_grad = 0;
                            if (return_alpha)
                                diff_grad += (alpha_map[mi] - alpha_out) * grad_alpha_map[mi];
                            if (return_rgb)
                                for (int k = 0; k < 3; k++)
                                    diff_grad += (rgb_map[mi * 3 + k] - rgb_out[k]) * grad_rgb_map[mi * 3 + k];
                            if (diff_grad <= 0)
                                continue;
                            if (p[1][0] != d0) {
                                REAL dist = (REAL)((p[1][0] - p[0][0]) / (p[1][0] - d0) * (d1 - d1_cross) * 2. / is);
                                dist = (0 < dist) ? dist + eps : dist - eps;
                                grad_face[pi[0] * 3 + (1 - axis)] -= diff_grad / dist;
                            }
                            if (p[0][0] != d0) {
                                REAL dist = (REAL)((p[1][0] - p[0][0]) / (d0 - p[0][0]) * (d1 - d1_cross) * 2. / is);
                                dist = (0 < dist) ? dist + eps : dist - eps;
                                grad_face[pi[1] * 3 + (1 - axis)] -= diff_grad / dist;
                            }
                        }
                    }
                }
            }
        }
        for (int k = 0; k < 9; k++)
            grad_faces[i * 9 + k] = grad_face[k];
    }
}

#ifndef ORA_NO_HOST_DRIVERS
static void FN(k4_range)(void *vc, long begin, long end)
{
    for (long i = begin; i < end; i++)
        FN(k4_face)((const struct FN(k4_ctx) *)vc, i);
}

void FN(backward_pixel_map)(const REAL *faces, const int32_t *face_index_map, const REAL *rgb_map,
                            const REAL *alpha_map, const REAL *grad_rgb_map, const REAL *grad_alpha_map,
                            REAL *grad_faces, int batch_size, int num_faces, int image_size, REAL eps,
                            int return_rgb, int return_alpha)
{
    struct FN(k4_ctx) c = {faces, face_index_map, rgb_map, alpha_map, grad_rgb_map, grad_alpha_map, grad_faces,
                           num_faces, image_size, eps, return_rgb, return_alpha};
    ora_parallel_for((long)batch_size * num_faces, FN(k4_range), &c);
}

#endif

/* K5: exact gradient w.r.t. the texture cubes, rasterize.py:284-297 (Appendix C5); one pixel.  Accumulates
 * into grad_textures (caller zero-fills, rasterize.py:151). */
ORA_HD void FN(bt_pixel)(const int32_t *face_index_map, const REAL *sampling_weight_map,
                         const int32_t *sampling_index_map, const REAL *grad_rgb_map, REAL *grad_textures,
                         int num_faces, int image_size, int texture_size, long i)
{
    const int is = image_size, nf = num_faces, ts = texture_size;
    const int face_index = face_index_map[i];
    if (face_index < 0)
        return;
    const int bn = (int)(i / (is * is));
    REAL *grad_texture = &grad_textures[((long)bn * nf + face_index) * ts * ts * ts * 3];
    for (int pn = 0; pn < 8; pn++) {
        const REAL w = sampling_weight_map[i * 8 + pn];
        const int isc = sampling_index_map[i * 8 + pn];
        for (int k = 0; k < 3; k++)
            ORA_ATOMIC_ADD(&grad_texture[isc * 3 + k], w * grad_rgb_map[i * 3 + k]);
    }
}

/* K6: analytic gradient of the interpolated depth, rasterize.py:300-315 (Appendix C6); one pixel.
 * Accumulates into grad_faces AFTER K4 stored into it. */
ORA_HD void FN(bd_pixel)(const REAL *faces, const REAL *depth_map, const int32_t *face_index_map,
                         const REAL *face_inv_map, const REAL *weight_map, const REAL *grad_depth_map,
                         REAL *grad_faces, int num_faces, int image_size, long i)
{
    const int is = image_size, nf = num_faces;
    const int fn = face_index_map[i];
    if (fn < 0)
        return;
    const int bn = (int)(i / (is * is));
    const REAL *face = &faces[((long)bn * nf + fn) * 9];
    const REAL depth = depth_map[i];
    const REAL depth2 = depth * depth;
    const REAL *face_inv = &face_inv_map[i * 9];
    const REAL *weight = &weight_map[i * 3];
    const REAL grad_depth = grad_depth_map[i];
    REAL *grad_face = &grad_faces[((long)bn * nf + fn) * 9];
    for (int k = 0; k < 3; k++) {
        const REAL z_k = face[3 * k + 2];
        ORA_ATOMIC_ADD(&grad_face[3 * k + 2], grad_depth * weight[k] * depth2 / (z_k * z_k));
    }
    REAL tmp[3] = {0, 0, 0};
    for (int k = 0; k < 3; k++)
        for (int l = 0; l < 3; l++)
            tmp[k] += -face_inv[3 * l + k] / face[3 * l + 2];
    for (int k = 0; k < 3; k++)
        for (int l = 0; l < 2; l++)
            ORA_ATOMIC_ADD(&grad_face[3 * k + l], -grad_depth * depth2 * weight[k] * tmp[l] * is / 2);
}

#ifndef ORA_NO_HOST_DRIVERS
/* serial => deterministic */
void FN(backward_textures)(const int32_t *face_index_map, const REAL *sampling_weight_map,
                           const int32_t *sampling_index_map, const REAL *grad_rgb_map, REAL *grad_textures,
                           int batch_size, int num_faces, int image_size, int texture_size)
{
    for (long i = 0; i < (long)batch_size * image_size * image_size; i++)
        FN(bt_pixel)(face_index_map, sampling_weight_map, sampling_index_map, grad_rgb_map, grad_textures, num_faces,
                     image_size, texture_size, i);
}

void FN(backward_depth_map)(const REAL *faces, const REAL *depth_map, const int32_t *face_index_map,
                            const REAL *face_inv_map, const REAL *weight_map, const REAL *grad_depth_map,
                            REAL *grad_faces, int batch_size, int num_faces, int image_size)
{
    for (long i = 0; i < (long)batch_size * image_size * image_size; i++)
        FN(bd_pixel)(faces, depth_map, face_index_map, face_inv_map, weight_map, grad_depth_map, grad_faces, num_faces,
                     image_size, i);
}
#endif
