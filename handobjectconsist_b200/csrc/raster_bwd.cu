/*
 * raster_bwd.cu -- rasterizer backward for sm_100a, one launch for all three gradients.
 *
 * Replaces backward_pixel_map, backward_textures and backward_depth_map of
 * `neural_renderer.cuda.rasterize` (bound at /root/reference/meshreg/neurender/rasterize.py:
 * 269-281, 290-297, 306-315) and the zero-fills of rasterize.py:151-181.
 *
 * The reference scatters texture / depth gradients from every pixel with float atomics and runs
 * the pseudo-gradient as one serial thread per face over whole image rows / columns.  Here a CTA owns
 * 256 faces: a thread sweeps the few pixels of its own face for the texture / depth terms (registers,
 * no atomics, no zero-fill of the outputs), the faces that own pixels are compacted and their
 * (face, edge, axis) scan tasks are spread over the CTA, and the long outward scans are clipped to the
 * span of the line where the incoming gradient is non-zero (a pixel with zero incoming gradient
 * contributes exactly nothing).  That span table is built by hoc_grad_extent_kernel, a streaming
 * pre-pass over the incoming gradients.  Weights, depth and texture taps are recomputed with the
 * forward's functions (bit-identical), so neither weight_map, face_inv_map nor the two sampling maps
 * of the reference are ever stored or read.
 */
#include "hoc_common.cuh"
#include "raster_math.h"

/* ------------------------------------------------------------------------------------------ */
/* Per-line span of non-zero incoming gradient.  ext layout: int [B][4][S] =
 * {row_lo, row_hi, col_lo, col_hi}; lo initialised to 0x7f7f7f7f, hi to -1 by memset. */
#define EXT_ROW_LO 0
#define EXT_ROW_HI 1
#define EXT_COL_LO 2
#define EXT_COL_HI 3

__global__ void __launch_bounds__(256)
hoc_grad_extent_kernel(const float *__restrict__ g_rgb, const float *__restrict__ g_alpha, int S, int layout,
                       int *__restrict__ ext)
{
    __shared__ int s_lo[8][32];
    __shared__ int s_hi[8][32];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int b = blockIdx.z;
    const int xi = blockIdx.x * 32 + tx;
    int *e = ext + (long)b * 4 * S;
    int c_lo = 0x7f7f7f7f, c_hi = -1;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int yi = blockIdx.y * 32 + r * 8 + ty;
        bool nz = false;
        if (xi < S && yi < S) {
            if (g_rgb != nullptr) {
                const float g0 = g_rgb[hoc_rgb_off(layout, S, b, yi, xi, 0)];
                const float g1 = g_rgb[hoc_rgb_off(layout, S, b, yi, xi, 1)];
                const float g2 = g_rgb[hoc_rgb_off(layout, S, b, yi, xi, 2)];
                nz = !(g0 == 0.0f) || !(g1 == 0.0f) || !(g2 == 0.0f);
            }
            if (g_alpha != nullptr)
                nz = nz || !(g_alpha[hoc_plane_off(layout, S, b, yi, xi)] == 0.0f);
        }
        const unsigned m = __ballot_sync(HOC_FULL_MASK, nz);
        if (m != 0 && tx == 0) {
            atomicMin(&e[EXT_ROW_LO * S + yi], blockIdx.x * 32 + (__ffs(m) - 1));
            atomicMax(&e[EXT_ROW_HI * S + yi], blockIdx.x * 32 + (31 - __clz(m)));
        }
        if (nz) {
            c_lo = min(c_lo, yi);
            c_hi = max(c_hi, yi);
        }
    }
    s_lo[ty][tx] = c_lo;
    s_hi[ty][tx] = c_hi;
    __syncthreads();
    if (ty == 0 && xi < S) {
#pragma unroll
        for (int r = 1; r < 8; r++) {
            c_lo = min(c_lo, s_lo[r][tx]);
            c_hi = max(c_hi, s_hi[r][tx]);
        }
        if (c_hi >= 0) {
            atomicMin(&e[EXT_COL_LO * S + xi], c_lo);
            atomicMax(&e[EXT_COL_HI * S + xi], c_hi);
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
struct HocBwdMaps {
    const int32_t *idx;   /* [S,S] of this sample, raster order */
    const float *rgb;     /* forward output, `layout`, base of the whole tensor */
    const float *g_rgb;   /* incoming gradient, `layout`, or NULL */
    const float *g_alpha; /* or NULL */
    int S, layout, b;
    bool use_alpha;       /* the alpha term exists (return_alpha and g_alpha given) */
    bool use_rgb;         /* the rgb term exists (return_rgb and g_rgb given) */
};

/* I(.) of the reference at a pixel: (alpha, r, g, b). */
__device__ __forceinline__ void hoc_load_I(const HocBwdMaps &M, int xi, int yi, float *I)
{
    I[0] = I[1] = I[2] = I[3] = 0.0f;
    if (M.use_alpha)
        I[0] = (M.idx[(long)yi * M.S + xi] >= 0) ? 1.0f : 0.0f;
    if (M.use_rgb) {
        I[1] = M.rgb[hoc_rgb_off(M.layout, M.S, M.b, yi, xi, 0)];
        I[2] = M.rgb[hoc_rgb_off(M.layout, M.S, M.b, yi, xi, 1)];
        I[3] = M.rgb[hoc_rgb_off(M.layout, M.S, M.b, yi, xi, 2)];
    }
}

/* delta = sum_ch (I(pixel) - Iref) * g(pixel), in the reference's accumulation order. */
__device__ __forceinline__ float hoc_delta(const HocBwdMaps &M, int xi, int yi, const float *Iref)
{
    float d = 0.0f;
    if (M.use_alpha) {
        const float a = (M.idx[(long)yi * M.S + xi] >= 0) ? 1.0f : 0.0f;
        d += (a - Iref[0]) * M.g_alpha[hoc_plane_off(M.layout, M.S, M.b, yi, xi)];
    }
    if (M.use_rgb) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const long o = hoc_rgb_off(M.layout, M.S, M.b, yi, xi, k);
            d += (M.rgb[o] - Iref[1 + k]) * M.g_rgb[o];
        }
    }
    return d;
}

#define BW_THREADS 256
#define BW_WARPS (BW_THREADS / 32)
#define BW_BIG 128 /* bounding boxes with more pixels than this are swept by the whole CTA */

/* Per-pixel contribution of an owned pixel to the texture (ts == 2: 8 cube corners x 3 channels) and
 * depth accumulators of its face. */
template <bool TS2>
__device__ __forceinline__ void hoc_bwd_pixel(const float *f, const float *inv, int xi, int yi, int b, int S, int ts,
                                              float near_, float far_, float eps, int layout, bool want_tex,
                                              bool want_depth, const float *__restrict__ g_rgb,
                                              const float *__restrict__ g_depth, float *acc_t, float *acc_d,
                                              float *gt, bool gt_shared)
{
    float w[3], zp;
    hoc_pixel_weights_depth(f, inv, xi, yi, near_, far_, w, &zp);
    if (want_depth) {
        const float gz = g_depth[hoc_plane_off(layout, S, b, yi, xi)] * zp * zp;
#pragma unroll
        for (int k = 0; k < 3; k++)
            acc_d[k] += gz * w[k];
    }
    if (want_tex) {
        float gr[3];
#pragma unroll
        for (int c = 0; c < 3; c++)
            gr[c] = g_rgb[hoc_rgb_off(layout, S, b, yi, xi, c)];
        float tf[3];
        int ti[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float t = hoc_tex_coord(w[k], f[3 * k + 2], zp, ts, eps);
            ti[k] = hoc_tex_cell(t, ts);
            tf[k] = t - (float)ti[k];
        }
#pragma unroll
        for (int pn = 0; pn < 8; pn++) {
            float ww = 1.0f;
            int isc = 0;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                if (((pn >> k) & 1) == 0) {
                    ww *= 1.0f - tf[k];
                    isc = isc * ts + ti[k];
                } else {
                    ww *= tf[k];
                    isc = isc * ts + ti[k] + 1;
                }
            }
            if (TS2) {
                /* ts == 2: floor(t) == 0, tap pn is cube corner (b0,b1,b2) */
                const int corner = ((pn & 1) << 2) | (pn & 2) | ((pn >> 2) & 1);
#pragma unroll
                for (int c = 0; c < 3; c++)
                    acc_t[corner * 3 + c] += ww * gr[c];
            } else {
                if (ts == 1)
                    isc = 0;
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    if (gt_shared)
                        atomicAdd(gt + isc * 3 + c, ww * gr[c]); /* CTA-cooperative sweep of a big face */
                    else
                        gt[isc * 3 + c] += ww * gr[c];           /* the face is private to this thread */
                }
            }
        }
    }
}

/* One (face, edge, axis) task of the pseudo-gradient: every integer column crossed by the edge, its
 * inward scan and (when the inside pixel is owned by the face) its outward scan clipped to the span of
 * non-zero incoming gradient.  Contributions to vertex A = edge and B = edge+1. */
__device__ __forceinline__ void hoc_k4_task(const float *f, int fi, int edge, int axis, const HocBwdMaps &M,
                                            const int *__restrict__ e, float eps, float *gA_out, float *gB_out)
{
    const int S = M.S;
    HocK4Edge E;
    hoc_k4_edge(f, S, edge, axis, &E);
    float gA = 0.0f, gB = 0.0f;
    for (int d0 = E.d0_from; d0 <= E.d0_to; d0++) {
        float d1_cross;
        int d1_in, d1_out;
        if (!hoc_k4_column(&E, S, d0, &d1_cross, &d1_in, &d1_out))
            continue;
        const int xin = axis == 0 ? d0 : d1_in, yin = axis == 0 ? d1_in : d0;
        const int xout = axis == 0 ? d0 : d1_out, yout = axis == 0 ? d1_out : d0;
        if (M.idx[(long)yin * S + xin] == fi) {
            float I_in[4];
            hoc_load_I(M, xin, yin, I_in);
            const int d1_limit = (0 < E.dir) ? S - 1 : 0;
            int d1_from = max(min(d1_out, d1_limit), 0);
            int d1_to = min(max(d1_out, d1_limit), S - 1);
            /* clip to where the incoming gradient of this line is non-zero */
            d1_from = max(d1_from, (axis == 0) ? e[EXT_COL_LO * S + d0] : e[EXT_ROW_LO * S + d0]);
            d1_to = min(d1_to, (axis == 0) ? e[EXT_COL_HI * S + d0] : e[EXT_ROW_HI * S + d0]);
            for (int d1 = d1_from; d1 <= d1_to; d1++) {
                const int xi = axis == 0 ? d0 : d1, yi = axis == 0 ? d1 : d0;
                const float delta = hoc_delta(M, xi, yi, I_in);
                if (delta <= 0.0f)
                    continue;
                hoc_k4_accum(&E, S, d0, d1, d1_cross, eps, delta, &gA, &gB);
            }
        }
        const int lim = hoc_k4_inward_limit(&E, d0);
        const int d1_from = max(min(d1_in, lim), 0);
        const int d1_to = min(max(d1_in, lim), S - 1);
        bool have_out = false;
        float I_out[4];
        for (int d1 = d1_from; d1 <= d1_to; d1++) {
            const int xi = axis == 0 ? d0 : d1, yi = axis == 0 ? d1 : d0;
            if (M.idx[(long)yi * S + xi] != fi)
                continue;
            if (!have_out) {
                hoc_load_I(M, xout, yout, I_out);
                have_out = true;
            }
            const float delta = hoc_delta(M, xi, yi, I_out);
            if (delta <= 0.0f)
                continue;
            hoc_k4_accum(&E, S, d0, d1, d1_cross, eps, delta, &gA, &gB);
        }
    }
    *gA_out = gA;
    *gB_out = gB;
}

/*
 * One CTA owns 256 consecutive faces of one sample.
 *   phase 1  thread = face: cull, then a serial sweep of the face's own clipped bounding box (faces are a
 *            few pixels large); owned pixels feed the texture / depth accumulators held in registers.
 *            Rare large faces are deferred and swept by the whole CTA.
 *   phase 2  the faces that own at least one pixel are compacted; thread = (face, edge, axis) task of the
 *            pseudo-gradient (faces that own no pixel cannot contribute: both scans require ownership).
 *   phase 3  thread = face again: sums the six task results in a fixed order, adds the depth term, stores.
 * No gradient is accumulated atomically in global memory (ts == 2); results are deterministic.
 */
template <bool TS2>
__global__ void __launch_bounds__(BW_THREADS)
hoc_raster_backward_kernel(const float *__restrict__ faces, const float *__restrict__ textures,
                           const int32_t *__restrict__ face_index_map, const float *__restrict__ rgb,
                           const float *__restrict__ g_rgb, const float *__restrict__ g_alpha,
                           const float *__restrict__ g_depth, int F, int S, int ts, float near_, float far_, float eps,
                           int layout, int use_alpha, const int *__restrict__ ext, float *__restrict__ grad_faces,
                           float *__restrict__ grad_textures)
{
    __shared__ float s_face[BW_THREADS][9];
    __shared__ float s_k4[BW_THREADS][6][2];
    __shared__ unsigned short s_vis[BW_THREADS];
    __shared__ unsigned short s_big[BW_THREADS];
    __shared__ float s_red[BW_WARPS][28];
    __shared__ int s_nvis, s_nbig;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    const int f_base = blockIdx.x * BW_THREADS;
    const int fi = f_base + tid;
    const bool valid = fi < F;
    const int tex_n = ts * ts * ts * 3;
    const int32_t *idx = face_index_map + (long)b * S * S;

    if (tid == 0) {
        s_nvis = 0;
        s_nbig = 0;
    }
    float f[9];
#pragma unroll
    for (int k = 0; k < 9; k++)
        f[k] = 0.0f;
    if (valid) {
        const float *src = faces + ((long)b * F + fi) * 9;
#pragma unroll
        for (int k = 0; k < 9; k++)
            f[k] = __ldg(src + k);
    }
#pragma unroll
    for (int k = 0; k < 9; k++)
        s_face[tid][k] = f[k];
#pragma unroll
    for (int k = 0; k < 12; k++)
        (&s_k4[tid][0][0])[k] = 0.0f;

    float *gt = (grad_textures != nullptr && valid) ? grad_textures + ((long)b * F + fi) * tex_n : nullptr;
    float *gf = (grad_faces != nullptr && valid) ? grad_faces + ((long)b * F + fi) * 9 : nullptr;
    const bool front = valid && hoc_face_xy_finite(f) && !hoc_face_back(f);

    HocBwdMaps M;
    M.idx = idx;
    M.rgb = rgb;
    M.g_rgb = g_rgb;
    M.g_alpha = g_alpha;
    M.S = S;
    M.layout = layout;
    M.b = b;
    M.use_alpha = (use_alpha != 0) && (g_alpha != nullptr);
    M.use_rgb = (rgb != nullptr) && (g_rgb != nullptr);
    const bool want_tex = (grad_textures != nullptr) && (g_rgb != nullptr);
    const bool want_depth = (grad_faces != nullptr) && (g_depth != nullptr);
    const bool want_k4 = (grad_faces != nullptr) && (M.use_alpha || M.use_rgb);

    /* generic texture size: the texture gradient block is accumulated in place, zero it first */
    if (gt != nullptr && (!TS2 || !want_tex || !front))
        for (int i = 0; i < tex_n; i++)
            gt[i] = 0.0f;
    __syncthreads(); /* counters + (for big faces) zero-fills visible CTA-wide */

    /* ---------------- phase 1: texture + depth gradients, ownership ---------------- */
    float acc_t[24];
    float acc_d[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 24; k++)
        acc_t[k] = 0.0f;
    float inv[9];
    bool hit = false;
    if (front) {
        hoc_face_inv(f, S, inv);
        const float pxmin = hoc_ndc_to_pix(fminf(f[0], fminf(f[3], f[6])), S);
        const float pxmax = hoc_ndc_to_pix(fmaxf(f[0], fmaxf(f[3], f[6])), S);
        const float pymin = hoc_ndc_to_pix(fminf(f[1], fminf(f[4], f[7])), S);
        const float pymax = hoc_ndc_to_pix(fmaxf(f[1], fmaxf(f[4], f[7])), S);
        const float fS1 = (float)(S - 1);
        const float x_lo = fmaxf(ceilf(pxmin - 0.5f), 0.0f);
        const float x_hi = fminf(floorf(pxmax + 0.5f), fS1);
        const float y_lo = fmaxf(ceilf(pymin - 0.5f), 0.0f);
        const float y_hi = fminf(floorf(pymax + 0.5f), fS1);
        if (x_lo <= x_hi && y_lo <= y_hi) {
            const int x0 = (int)x_lo, y0 = (int)y_lo, x1 = (int)x_hi, y1 = (int)y_hi;
            if ((x1 - x0 + 1) * (y1 - y0 + 1) > BW_BIG) {
                s_big[atomicAdd(&s_nbig, 1)] = (unsigned short)tid;
            } else {
                for (int yi = y0; yi <= y1; yi++)
                    for (int xi = x0; xi <= x1; xi++) {
                        if (idx[(long)yi * S + xi] != fi)
                            continue;
                        hit = true;
                        hoc_bwd_pixel<TS2>(f, inv, xi, yi, b, S, ts, near_, far_, eps, layout, want_tex, want_depth,
                                           g_rgb, g_depth, acc_t, acc_d, gt, false);
                    }
            }
        }
    }
    __syncthreads();
    /* large faces: the whole CTA sweeps the bounding box, block reduction, owner thread keeps the sums */
    const int nbig = s_nbig;
    for (int q = 0; q < nbig; q++) {
        const int owner = s_big[q];
        float bf[9], binv[9];
#pragma unroll
        for (int k = 0; k < 9; k++)
            bf[k] = s_face[owner][k];
        hoc_face_inv(bf, S, binv);
        const float fS1 = (float)(S - 1);
        const int x0 = (int)fmaxf(ceilf(hoc_ndc_to_pix(fminf(bf[0], fminf(bf[3], bf[6])), S) - 0.5f), 0.0f);
        const int x1 = (int)fminf(floorf(hoc_ndc_to_pix(fmaxf(bf[0], fmaxf(bf[3], bf[6])), S) + 0.5f), fS1);
        const int y0 = (int)fmaxf(ceilf(hoc_ndc_to_pix(fminf(bf[1], fminf(bf[4], bf[7])), S) - 0.5f), 0.0f);
        const int y1 = (int)fminf(floorf(hoc_ndc_to_pix(fmaxf(bf[1], fmaxf(bf[4], bf[7])), S) + 0.5f), fS1);
        const int bw = x1 - x0 + 1, n = bw * (y1 - y0 + 1);
        const int bfi = f_base + owner;
        float part[28];
#pragma unroll
        for (int k = 0; k < 28; k++)
            part[k] = 0.0f;
        float *bgt = (grad_textures != nullptr) ? grad_textures + ((long)b * F + bfi) * tex_n : nullptr;
        for (int p = tid; p < n; p += BW_THREADS) {
            const int yy = p / bw, xi = x0 + (p - yy * bw), yi = y0 + yy;
            if (idx[(long)yi * S + xi] != bfi)
                continue;
            part[27] = 1.0f;
            hoc_bwd_pixel<TS2>(bf, binv, xi, yi, b, S, ts, near_, far_, eps, layout, want_tex, want_depth, g_rgb,
                               g_depth, part, part + 24, bgt, true);
        }
#pragma unroll
        for (int k = 0; k < 28; k++) {
            const float v = hoc_warp_sum(part[k]);
            if (lane == 0)
                s_red[warp][k] = v;
        }
        __syncthreads();
        if (tid == owner) {
            float any = 0.0f;
            for (int wq = 0; wq < BW_WARPS; wq++) {
#pragma unroll
                for (int k = 0; k < 24; k++)
                    acc_t[k] += s_red[wq][k];
#pragma unroll
                for (int k = 0; k < 3; k++)
                    acc_d[k] += s_red[wq][24 + k];
                any += s_red[wq][27];
            }
            hit = any > 0.0f;
        }
        __syncthreads();
    }
    if (TS2 && want_tex && front) {
        float4 *dst = reinterpret_cast<float4 *>(gt); /* 96 B per face, 16 B aligned */
#pragma unroll
        for (int q = 0; q < 6; q++)
            dst[q] = make_float4(acc_t[4 * q], acc_t[4 * q + 1], acc_t[4 * q + 2], acc_t[4 * q + 3]);
    }
    if (grad_faces == nullptr)
        return;

    /* ---------------- phase 2: pseudo-gradient tasks of the faces that own pixels ---------------- */
    if (want_k4) {
        if (hit)
            s_vis[atomicAdd(&s_nvis, 1)] = (unsigned short)tid;
        __syncthreads();
        const int ntask = s_nvis * 6;
        const int *e = ext + (long)b * 4 * S;
        for (int task = tid; task < ntask; task += BW_THREADS) {
            const int lf = s_vis[task / 6];
            const int combo = task - (task / 6) * 6;
            float tf[9];
#pragma unroll
            for (int k = 0; k < 9; k++)
                tf[k] = s_face[lf][k];
            float gA, gB;
            hoc_k4_task(tf, f_base + lf, combo >> 1, combo & 1, M, e, eps, &gA, &gB);
            s_k4[lf][combo][0] = gA;
            s_k4[lf][combo][1] = gB;
        }
        __syncthreads();
    }

    /* ---------------- phase 3: assemble grad_faces ---------------- */
    if (gf == nullptr)
        return;
    float gface[9];
#pragma unroll
    for (int k = 0; k < 9; k++)
        gface[k] = 0.0f;
    if (front) {
        if (want_depth && hit) {
            float tmp[2];
#pragma unroll
            for (int l = 0; l < 2; l++)
                tmp[l] = inv[l] / f[2] + inv[3 + l] / f[5] + inv[6 + l] / f[8];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float zk = f[3 * k + 2];
                gface[3 * k + 2] = acc_d[k] / (zk * zk);
                gface[3 * k + 0] = acc_d[k] * tmp[0] * (float)S / 2.0f;
                gface[3 * k + 1] = acc_d[k] * tmp[1] * (float)S / 2.0f;
            }
        }
        if (want_k4 && hit) {
#pragma unroll
            for (int combo = 0; combo < 6; combo++) {
                const int edge = combo >> 1, axis = combo & 1;
                gface[edge * 3 + (1 - axis)] += s_k4[tid][combo][0];
                gface[((edge + 1) % 3) * 3 + (1 - axis)] += s_k4[tid][combo][1];
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 9; k++)
        gf[k] = gface[k];
}

extern "C" size_t hoc_raster_backward_workspace_bytes(int B, int F, int S)
{
    (void)F;
    if (B <= 0 || S <= 0)
        return 0;
    return (size_t)B * 4 * S * sizeof(int);
}

extern "C" int hoc_raster_backward(const float *faces, const float *textures, const int32_t *face_index_map,
                                   const float *rgb, const float *grad_rgb, const float *grad_alpha,
                                   const float *grad_depth, int B, int F, int S, int ts, float near_, float far_,
                                   float eps, int layout, int use_alpha, float *grad_faces, float *grad_textures,
                                   void *workspace, size_t workspace_bytes, void *stream)
{
    HOC_CHECK_ARG(B >= 0 && F >= 0, "hoc_raster_backward: negative batch (%d) or face count (%d)", B, F);
    HOC_CHECK_ARG(S >= 1 && S <= 2048, "hoc_raster_backward: image_size %d outside [1, 2048]", S);
    HOC_CHECK_ARG(layout == HOC_LAYOUT_RAW || layout == HOC_LAYOUT_IMAGE, "hoc_raster_backward: bad layout %d",
                  layout);
    HOC_CHECK_ARG(B <= 65535, "hoc_raster_backward: batch %d exceeds 65535", B);
    HOC_CHECK_ARG(grad_textures == nullptr || ts >= 1, "hoc_raster_backward: texture_size %d", ts);
    HOC_CHECK_ARG(grad_rgb == nullptr || rgb != nullptr, "hoc_raster_backward: grad_rgb given without rgb");
    if (B == 0 || F == 0)
        return HOC_OK;
    HOC_CHECK_ARG(faces != nullptr && face_index_map != nullptr, "hoc_raster_backward: faces / face_index_map NULL");
    if (grad_faces == nullptr && grad_textures == nullptr)
        return HOC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t need = hoc_raster_backward_workspace_bytes(B, F, S);
    int *ext = (int *)workspace;
    const bool k4 = grad_faces != nullptr && ((grad_rgb != nullptr) || (use_alpha && grad_alpha != nullptr));
    if (k4) {
        if (workspace == nullptr || workspace_bytes < need) {
            hoc_set_error("hoc_raster_backward: workspace of %zu bytes needed, %zu given", need, workspace_bytes);
            return HOC_ERR_WORKSPACE;
        }
        /* hi rows = -1 ... */
        cudaError_t e = cudaMemsetAsync(ext, 0xff, need, st);
        if (e != cudaSuccess) {
            hoc_set_error("hoc_raster_backward: memset failed: %s", cudaGetErrorString(e));
            return HOC_ERR_CUDA;
        }
        /* ... lo rows (EXT_ROW_LO, EXT_COL_LO = every second row of S ints) = 0x7f7f7f7f */
        e = cudaMemset2DAsync(ext, 2 * S * sizeof(int), 0x7f, S * sizeof(int), (size_t)B * 2, st);
        if (e != cudaSuccess) {
            hoc_set_error("hoc_raster_backward: memset2d failed: %s", cudaGetErrorString(e));
            return HOC_ERR_CUDA;
        }
        dim3 eg((S + 31) / 32, (S + 31) / 32, B);
        HOC_LAUNCH(HOC_K_GRAD_EXTENT, st,
                   (hoc_grad_extent_kernel<<<eg, dim3(32, 8), 0, st>>>(grad_rgb, (use_alpha ? grad_alpha : nullptr), S,
                                                                      layout, ext)));
        HOC_CHECK_LAUNCH("hoc_grad_extent_kernel");
    }
    dim3 grid((F + BW_THREADS - 1) / BW_THREADS, B);
    if (ts == 2)
        HOC_LAUNCH(HOC_K_RASTER_BACKWARD, st,
                   (hoc_raster_backward_kernel<true><<<grid, BW_THREADS, 0, st>>>(
                       faces, textures, face_index_map, rgb, grad_rgb, grad_alpha, grad_depth, F, S, ts, near_, far_, eps,
                       layout, use_alpha, ext, grad_faces, grad_textures)));
    else
        HOC_LAUNCH(HOC_K_RASTER_BACKWARD, st,
                   (hoc_raster_backward_kernel<false><<<grid, BW_THREADS, 0, st>>>(
                       faces, textures, face_index_map, rgb, grad_rgb, grad_alpha, grad_depth, F, S, ts, near_, far_, eps,
                       layout, use_alpha, ext, grad_faces, grad_textures)));
    HOC_CHECK_LAUNCH("hoc_raster_backward_kernel");
    return HOC_OK;
}
