/*
 * raster_bwd.cu -- rasterizer backward for sm_100a, one launch for all three gradients.
 *
 * Replaces backward_pixel_map, backward_textures and backward_depth_map of
 * `neural_renderer.cuda.rasterize` (bound at /root/reference/meshreg/neurender/rasterize.py:
 * 269-281, 290-297, 306-315) and the zero-fills of rasterize.py:151-181.
 *
 * The reference scatters texture / depth gradients from every pixel with float atomics and runs
 * the pseudo-gradient as one serial thread per face.  Here ONE WARP owns ONE FACE and produces
 * all of that face's gradients, so nothing is accumulated atomically, the result is
 * deterministic and the output buffers need no zero-fill:
 *
 *   part A (textures + depth)  the warp sweeps the face's clipped pixel bounding box, keeps the
 *           pixels whose face_index_map entry is this face, recomputes weights / depth / the
 *           eight trilinear taps with the forward's functions (bit-identical, so neither
 *           weight_map, face_inv_map nor the two sampling maps are ever stored or read) and
 *           reduces the per-lane partial sums with shuffles.
 *   part B (pseudo-gradient)   lanes take the integer columns (rows) crossed by the three
 *           edges; each lane does its short inward scan itself; the long outward scans (to the
 *           image border) are executed cooperatively, 32 pixels per step, coalesced along rows,
 *           and clipped to the span of the line where the incoming gradient is non-zero (a
 *           pixel with zero incoming gradient contributes exactly nothing).  That span table is
 *           built by hoc_grad_extent_kernel, a streaming pre-pass over the incoming gradients.
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

#define BW_WARPS 8
#define BW_THREADS (BW_WARPS * 32)

template <bool TS2>
__global__ void __launch_bounds__(BW_THREADS)
hoc_raster_backward_kernel(const float *__restrict__ faces, const float *__restrict__ textures,
                           const int32_t *__restrict__ face_index_map, const float *__restrict__ rgb,
                           const float *__restrict__ g_rgb, const float *__restrict__ g_alpha,
                           const float *__restrict__ g_depth, int F, int S, int ts, float near_, float far_, float eps,
                           int layout, int use_alpha, const int *__restrict__ ext, float *__restrict__ grad_faces,
                           float *__restrict__ grad_textures)
{
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int fi = blockIdx.x * BW_WARPS + warp;
    if (fi >= F)
        return;

    float f[9];
    {
        const float *src = faces + ((long)b * F + fi) * 9;
#pragma unroll
        for (int k = 0; k < 9; k++)
            f[k] = __ldg(src + k);
    }
    const int tex_n = ts * ts * ts * 3;
    float *gt = (grad_textures != nullptr) ? grad_textures + ((long)b * F + fi) * tex_n : nullptr;
    float *gf = (grad_faces != nullptr) ? grad_faces + ((long)b * F + fi) * 9 : nullptr;

    const bool front = hoc_face_xy_finite(f) && !hoc_face_back(f);
    if (!front) {
        if (gf != nullptr && lane < 9)
            gf[lane] = 0.0f;
        if (gt != nullptr)
            for (int i = lane; i < tex_n; i += 32)
                gt[i] = 0.0f;
        return;
    }

    const int32_t *idx = face_index_map + (long)b * S * S;

    /* ---------------- part A: texture + depth gradients over the bounding box -------------- */
    float acc_t[24];
    float acc_d[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 24; k++)
        acc_t[k] = 0.0f;
    float inv[9];
    hoc_face_inv(f, S, inv);
    const bool want_tex = (gt != nullptr) && (g_rgb != nullptr);
    const bool want_depth = (gf != nullptr) && (g_depth != nullptr);
    if (gt != nullptr && (!TS2 || !want_tex)) {
        for (int i = lane; i < tex_n; i += 32)
            gt[i] = 0.0f;
        __syncwarp();
    }
    bool any_hit = false;
    if (want_tex || want_depth) {
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
            const int x0 = (int)x_lo, y0 = (int)y_lo;
            const int bw = (int)(x_hi - x_lo + 1.0f), bh = (int)(y_hi - y_lo + 1.0f);
            const int n = bw * bh;
            for (int p = lane; p < n; p += 32) {
                const int yy = p / bw;
                const int xi = x0 + (p - yy * bw);
                const int yi = y0 + yy;
                if (idx[(long)yi * S + xi] != fi)
                    continue;
                any_hit = true;
                float w[3], zp;
                hoc_pixel_weights_depth(f, inv, xi, yi, near_, far_, w, &zp);
                if (want_depth) {
                    const float gd = g_depth[hoc_plane_off(layout, S, b, yi, xi)];
                    const float gz = gd * zp * zp;
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
                            for (int c = 0; c < 3; c++)
                                atomicAdd(gt + isc * 3 + c, ww * gr[c]);
                        }
                    }
                }
            }
        }
    }
    any_hit = __any_sync(HOC_FULL_MASK, any_hit);
    if (any_hit) {
        if (want_depth) {
#pragma unroll
            for (int k = 0; k < 3; k++)
                acc_d[k] = hoc_warp_sum(acc_d[k]);
        }
        if (TS2 && want_tex) {
#pragma unroll
            for (int k = 0; k < 24; k++)
                acc_t[k] = hoc_warp_sum(acc_t[k]);
        }
    }
    if (TS2 && want_tex && lane == 0) {
        float4 *dst = reinterpret_cast<float4 *>(gt); /* 96 B per face, 16 B aligned */
#pragma unroll
        for (int q = 0; q < 6; q++)
            dst[q] = make_float4(acc_t[4 * q], acc_t[4 * q + 1], acc_t[4 * q + 2], acc_t[4 * q + 3]);
    }
    if (gf == nullptr)
        return;

    /* depth gradient of the face (backward_depth_map), from A_k = sum gd * zp^2 * w_k */
    float gface[9];
#pragma unroll
    for (int k = 0; k < 9; k++)
        gface[k] = 0.0f;
    if (want_depth && any_hit) {
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

    /* ---------------- part B: pseudo-gradient of rgb / alpha w.r.t. vertex xy -------------- */
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

    float gsum[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}; /* slot = vertex * 2 + (0: x, 1: y) */
    if (M.use_alpha || M.use_rgb) {
        const int *e = ext + (long)b * 4 * S;
        for (int combo = 0; combo < 6; combo++) {
            const int edge = combo >> 1, axis = combo & 1;
            HocK4Edge E;
            hoc_k4_edge(f, S, edge, axis, &E);
            const int n = E.d0_to - E.d0_from + 1;
            if (n <= 0)
                continue;
            float gA = 0.0f, gB = 0.0f;
            for (int base = 0; base < n; base += 32) {
                const int t = base + lane;
                const int d0 = E.d0_from + t;
                bool pending = false;
                float d1_cross = 0.0f;
                int d1_in = 0, d1_out = 0;
                float I_in[4] = {0.f, 0.f, 0.f, 0.f};
                if (t < n && hoc_k4_column(&E, S, d0, &d1_cross, &d1_in, &d1_out)) {
                    const int xin = axis == 0 ? d0 : d1_in, yin = axis == 0 ? d1_in : d0;
                    const int xout = axis == 0 ? d0 : d1_out, yout = axis == 0 ? d1_out : d0;
                    float I_out[4];
                    hoc_load_I(M, xin, yin, I_in);
                    hoc_load_I(M, xout, yout, I_out);
                    pending = (idx[(long)yin * S + xin] == fi);
                    /* inward scan, lane-serial (bounded by the face's own extent) */
                    const int lim = hoc_k4_inward_limit(&E, d0);
                    const int d1_from = max(min(d1_in, lim), 0);
                    const int d1_to = min(max(d1_in, lim), S - 1);
                    for (int d1 = d1_from; d1 <= d1_to; d1++) {
                        const int xi = axis == 0 ? d0 : d1, yi = axis == 0 ? d1 : d0;
                        if (idx[(long)yi * S + xi] != fi)
                            continue;
                        const float delta = hoc_delta(M, xi, yi, I_out);
                        if (delta <= 0.0f)
                            continue;
                        hoc_k4_accum(&E, S, d0, d1, d1_cross, eps, delta, &gA, &gB);
                    }
                }
                /* outward scans, cooperative: 32 pixels of one scan per step */
                unsigned m = __ballot_sync(HOC_FULL_MASK, pending);
                while (m) {
                    const int j = __ffs(m) - 1;
                    m &= m - 1;
                    const int s_d0 = __shfl_sync(HOC_FULL_MASK, d0, j);
                    const int s_out = __shfl_sync(HOC_FULL_MASK, d1_out, j);
                    const float s_cross = __shfl_sync(HOC_FULL_MASK, d1_cross, j);
                    float s_I[4];
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        s_I[k] = __shfl_sync(HOC_FULL_MASK, I_in[k], j);
                    const int d1_limit = (0 < E.dir) ? S - 1 : 0;
                    int d1_from = max(min(s_out, d1_limit), 0);
                    int d1_to = min(max(s_out, d1_limit), S - 1);
                    /* clip to where the incoming gradient of this line is non-zero */
                    const int lo = (axis == 0) ? e[EXT_COL_LO * S + s_d0] : e[EXT_ROW_LO * S + s_d0];
                    const int hi = (axis == 0) ? e[EXT_COL_HI * S + s_d0] : e[EXT_ROW_HI * S + s_d0];
                    d1_from = max(d1_from, lo);
                    d1_to = min(d1_to, hi);
                    for (int d1 = d1_from + lane; d1 <= d1_to; d1 += 32) {
                        const int xi = axis == 0 ? s_d0 : d1, yi = axis == 0 ? d1 : s_d0;
                        const float delta = hoc_delta(M, xi, yi, s_I);
                        if (delta <= 0.0f)
                            continue;
                        hoc_k4_accum(&E, S, s_d0, d1, s_cross, eps, delta, &gA, &gB);
                    }
                }
            }
            /* A = edge, B = edge+1; the walk along axis updates the perpendicular coordinate */
            const int slotA = edge * 2 + (1 - axis);
            const int slotB = ((edge + 1) % 3) * 2 + (1 - axis);
#pragma unroll
            for (int s = 0; s < 6; s++) {
                if (s == slotA)
                    gsum[s] += gA;
                if (s == slotB)
                    gsum[s] += gB;
            }
        }
#pragma unroll
        for (int s = 0; s < 6; s++)
            gsum[s] = hoc_warp_sum(gsum[s]);
    }
    if (lane == 0) {
#pragma unroll
        for (int v = 0; v < 3; v++) {
            gf[3 * v + 0] = gsum[2 * v + 0] + gface[3 * v + 0];
            gf[3 * v + 1] = gsum[2 * v + 1] + gface[3 * v + 1];
            gf[3 * v + 2] = gface[3 * v + 2];
        }
    }
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
    dim3 grid((F + BW_WARPS - 1) / BW_WARPS, B);
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
