/*
 * raster_bwd.cu -- rasterizer backward for sm_100a.
 *
 * Replaces backward_pixel_map, backward_textures and backward_depth_map of
 * `neural_renderer.cuda.rasterize` (bound at /root/reference/meshreg/neurender/rasterize.py:
 * 269-281, 290-297, 306-315) and the zero-fills of rasterize.py:151-181.
 *
 * The reference runs the pseudo-gradient as ONE SERIAL THREAD PER FACE that walks whole image rows and
 * columns (O(edge length x S) per face, every pixel fetched from global memory again by every scan that
 * passes over it) and scatters texture / depth gradients with one thread per pixel.  Here the work is
 * split by what it is parallel over:
 *
 *   hoc_raster_bwd_pixel_kernel   pixel-parallel, streaming (one pass over face_index_map and the
 *           incoming gradients): per-line spans of non-zero incoming gradient (a pixel with zero incoming
 *           gradient contributes exactly nothing to any scan, so scans are clipped to the span), and for
 *           covered pixels the texture / depth gradient of the owning face plus its owned-pixel count.
 *           Weights, depth and the eight trilinear taps are recomputed with the forward's functions
 *           (bit-identical), so weight_map, face_inv_map and the reference's two sampling maps (64 B/px)
 *           are never stored or read.
 *   hoc_raster_bwd_face_kernel    face-parallel: a CTA owns 256 faces; faces that own no pixel are done
 *           (both scans of the pseudo-gradient require ownership); the others are compacted and their
 *           (face, edge, axis) tasks spread over the CTA.  A task walks the integer columns its edge
 *           crosses, does the short INWARD scan itself and hands the long OUTWARD scan to the line it
 *           runs along by appending a 4-byte record (face, edge) to that line's bucket.
 *   hoc_raster_bwd_line_kernel    line-parallel: a CTA owns one image row or column, stages the span of
 *           that line (I and dL/dI of every pixel) in shared memory ONCE and lets every lane run one
 *           outward scan out of shared memory; each scan adds its two vertex contributions to grad_faces.
 */
#include "hoc_common.cuh"
#include "raster_math.h"

#define EXT_ROW_LO 0
#define EXT_ROW_HI 1
#define EXT_COL_LO 2
#define EXT_COL_HI 3

/* Workspace carved by hoc_raster_backward (all int32 / float32, 16-byte aligned regions):
 *   ext        int [B][4][S]      {row_lo, row_hi, col_lo, col_hi}: span of non-zero incoming gradient
 *   owned      int [B][F]         pixels owned by each face
 *   acc_d      float [B][F][3]    sum over owned pixels of dL/ddepth * depth^2 * w_k
 *   line_count int [B][2][S]      outward scans queued on each line (axis 0: column x, axis 1: row y)
 *   emitters   int [B][2][S][3S]  the queue: face | edge << 29.  3S is a hard bound: a scan is keyed by its
 *                                 inside pixel on the line, which is owned by exactly one face with 3 edges */
struct HocBwdWorkspace {
    int *ext;
    int *owned;
    float *acc_d;
    int *line_count;
    int *emitters;
    size_t zero_begin, zero_bytes; /* byte range (owned .. line_count) that must be zero-filled */
    size_t total;
};

static HocBwdWorkspace hoc_bwd_workspace(void *base, int B, int F, int S)
{
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    HocBwdWorkspace w;
    size_t off = 0;
    char *p = (char *)base;
    w.ext = (int *)(p + off);
    off = up(off + sizeof(int) * 4 * (size_t)B * S);
    w.zero_begin = off;
    w.owned = (int *)(p + off);
    off = up(off + sizeof(int) * (size_t)B * F);
    w.acc_d = (float *)(p + off);
    off = up(off + sizeof(float) * 3 * (size_t)B * F);
    w.line_count = (int *)(p + off);
    off = up(off + sizeof(int) * 2 * (size_t)B * S);
    w.zero_bytes = off - w.zero_begin;
    w.emitters = (int *)(p + off);
    off = up(off + sizeof(int) * 2 * (size_t)B * S * 3 * (size_t)S);
    w.total = off;
    return w;
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

#define BW_THREADS 128
#define BW_WARPS (BW_THREADS / 32)

/*
 * Pixel pass.  Block (32, 8) covers a 32 x 32 pixel tile (4 rows per thread).
 */
template <bool TS2>
__global__ void __launch_bounds__(256)
hoc_raster_bwd_pixel_kernel(const float *__restrict__ faces, const int32_t *__restrict__ face_index_map,
                            const float *__restrict__ weight_map, const float *__restrict__ depth_map,
                            const float *__restrict__ g_rgb, const float *__restrict__ g_alpha,
                            const float *__restrict__ g_depth, int F, int S, int ts, float near_, float far_, float eps,
                            int layout, int want_ext, int tex_mode, int *__restrict__ ext, int *__restrict__ owned,
                            float *__restrict__ acc_d, float *__restrict__ grad_textures)
{
    __shared__ int s_lo[8][32];
    __shared__ int s_hi[8][32];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int b = blockIdx.z;
    const int xi = blockIdx.x * 32 + tx;
    int *e = ext + (long)b * 4 * S;
    const int tex_n = ts * ts * ts * 3;
    int c_lo = 0x7f7f7f7f, c_hi = -1;
    /* issue the loads of all four rows first (memory-level parallelism), then do the per-pixel work */
    float gr[4][3], ga[4];
    int fis[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int yi = blockIdx.y * 32 + r * 8 + ty;
        const bool in = xi < S && yi < S;
        gr[r][0] = gr[r][1] = gr[r][2] = 0.0f;
        ga[r] = 0.0f;
        fis[r] = -1;
        if (in) {
            if (g_rgb != nullptr) {
#pragma unroll
                for (int c = 0; c < 3; c++)
                    gr[r][c] = g_rgb[hoc_rgb_off(layout, S, b, yi, xi, c)];
            }
            if (g_alpha != nullptr)
                ga[r] = g_alpha[hoc_plane_off(layout, S, b, yi, xi)];
            fis[r] = face_index_map[((long)b * S + yi) * S + xi];
        }
    }
    const bool want_tex = (grad_textures != nullptr) && (g_rgb != nullptr);
    const bool want_depth = (acc_d != nullptr) && (g_depth != nullptr);
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int yi = blockIdx.y * 32 + r * 8 + ty;
        const bool nz = !(gr[r][0] == 0.0f) || !(gr[r][1] == 0.0f) || !(gr[r][2] == 0.0f) || !(ga[r] == 0.0f);
        const int fi = fis[r];
        if (fi >= 0) {
            atomicAdd(owned + (long)b * F + fi, 1);
            if ((want_tex && nz) || want_depth) {
                float f[9], w[3], zp;
                const float *src = faces + ((long)b * F + fi) * 9;
                if (weight_map != nullptr && depth_map != nullptr) {
                    /* the forward saved its weights and depth: only the three vertex depths are needed */
                    const float *wm = weight_map + (((long)b * S + yi) * S + xi) * 3;
                    w[0] = wm[0];
                    w[1] = wm[1];
                    w[2] = wm[2];
                    zp = depth_map[hoc_plane_off(layout, S, b, yi, xi)];
                    f[2] = __ldg(src + 2);
                    f[5] = __ldg(src + 5);
                    f[8] = __ldg(src + 8);
                } else { /* recompute with the forward's functions (bit-identical) */
                    float inv[9];
#pragma unroll
                    for (int k = 0; k < 9; k++)
                        f[k] = __ldg(src + k);
                    hoc_face_inv(f, S, inv);
                    hoc_pixel_weights_depth(f, inv, xi, yi, near_, far_, w, &zp);
                }
                if (want_depth) {
                    const float gz = g_depth[hoc_plane_off(layout, S, b, yi, xi)] * zp * zp;
                    if (gz != 0.0f) {
                        float *ad = acc_d + ((long)b * F + fi) * 3;
#pragma unroll
                        for (int k = 0; k < 3; k++)
                            atomicAdd(ad + k, gz * w[k]);
                    }
                }
                if (want_tex && nz && tex_mode == HOC_TEX_GRAD_VERTEX) {
                    /* textures are the multilinear extension of three vertex values (T[i,j,k] = i c0 + j c1 +
                     * k c2, ts == 2): d rgb / d c_k = t_k, so nine sums per face instead of twenty-four */
                    float *gt = grad_textures + ((long)b * F + fi) * 9;
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        const float t = hoc_tex_coord(w[k], f[3 * k + 2], zp, 2, eps);
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            const float v = t * gr[r][c];
                            if (v != 0.0f)
                                atomicAdd(gt + 3 * k + c, v);
                        }
                    }
                } else if (want_tex && nz) {
                    float *gt = grad_textures + ((long)b * F + fi) * tex_n;
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
                        if (!TS2 && ts == 1)
                            isc = 0;
                        if (ww != 0.0f) {
#pragma unroll
                            for (int c = 0; c < 3; c++) {
                                const float v = ww * gr[r][c];
                                if (v != 0.0f)
                                    atomicAdd(gt + isc * 3 + c, v);
                            }
                        }
                    }
                }
            }
        }
        if (want_ext) {
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
    }
    if (!want_ext)
        return;
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

/* One column d0 of a (face, edge, axis): queue the outward scan on the line it runs along when the inside
 * pixel is owned by the face, and do the short inward scan (the face's own pixels) here. */
__device__ __forceinline__ void hoc_k4_face_column(const HocK4Edge &E, int fi, int edge, int axis, int d0,
                                                   const HocBwdMaps &M, float eps, int *__restrict__ line_count,
                                                   int *__restrict__ emitters, float *gA_out, float *gB_out)
{
    const int S = M.S;
    float gA = 0.0f, gB = 0.0f;
    float d1_cross;
    int d1_in, d1_out;
    if (hoc_k4_column(&E, S, d0, &d1_cross, &d1_in, &d1_out)) {
        const int xin = axis == 0 ? d0 : d1_in, yin = axis == 0 ? d1_in : d0;
        if (M.idx[(long)yin * S + xin] == fi) {
            const long line = ((long)M.b * 2 + axis) * S + d0;
            const int pos = atomicAdd(line_count + line, 1);
            if (pos < 3 * S) /* cannot fail (see HocBwdWorkspace); keeps a corrupted map from overrunning */
                emitters[line * 3 * S + pos] = fi | (edge << 29);
        }
        const int lim = hoc_k4_inward_limit(&E, d0);
        const int d1_from = max(min(d1_in, lim), 0);
        const int d1_to = min(max(d1_in, lim), S - 1);
        bool have_out = false;
        float I_out[4];
        HocK4Col C;
        for (int d1 = d1_from; d1 <= d1_to; d1++) {
            const int xi = axis == 0 ? d0 : d1, yi = axis == 0 ? d1 : d0;
            if (M.idx[(long)yi * S + xi] != fi)
                continue;
            if (!have_out) {
                hoc_load_I(M, axis == 0 ? d0 : d1_out, axis == 0 ? d1_out : d0, I_out);
                hoc_k4_col(&E, S, d0, d1_cross, &C);
                have_out = true;
            }
            const float delta = hoc_delta(M, xi, yi, I_out);
            if (!(delta <= 0.0f))
                hoc_k4_accum_col(&C, d1, eps, delta, &gA, &gB);
        }
    }
    *gA_out = gA;
    *gB_out = gB;
}

/*
 * Face pass.  One CTA owns BW_THREADS consecutive faces of one sample; writes grad_faces for all of them
 * (depth term + inward scans; zeros for culled / unowned faces).  The faces that own pixels are
 * compacted and their work is flattened to one item per (face, edge, axis, column), so that every lane
 * has a short, independent task (the columns of all edges of ~35 faces, ~600 items per CTA).
 */
__global__ void __launch_bounds__(BW_THREADS, 8) /* 8 CTAs/SM: the 1152 CTAs of 16 x 9104 faces fit in one wave */
hoc_raster_bwd_face_kernel(const float *__restrict__ faces, const int32_t *__restrict__ face_index_map,
                           const float *__restrict__ rgb, const float *__restrict__ g_rgb,
                           const float *__restrict__ g_alpha, int F, int S, float eps, int layout, int use_alpha,
                           int want_depth, const int *__restrict__ owned, const float *__restrict__ acc_d,
                           int *__restrict__ line_count, int *__restrict__ emitters, float *__restrict__ grad_faces)
{
    __shared__ float s_face[BW_THREADS][9];
    __shared__ float s_k4[BW_THREADS][6][2];
    __shared__ unsigned short s_vis[BW_THREADS];
    __shared__ int s_pre[BW_THREADS * 6 + 1];
    __shared__ int s_nvis;

    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    /* faces are dealt to the CTAs round-robin (CTA c takes faces c, c + G, c + 2G, ...) so that every CTA
     * gets the same mix of large / small / hidden faces */
    const int fi = tid * gridDim.x + blockIdx.x;
    const bool valid = fi < F;
    if (tid == 0)
        s_nvis = 0;
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
    const bool front = valid && hoc_face_xy_finite(f) && !hoc_face_back(f);
    const bool hit = front && owned[(long)b * F + fi] > 0;

    HocBwdMaps M;
    M.idx = face_index_map + (long)b * S * S;
    M.rgb = rgb;
    M.g_rgb = g_rgb;
    M.g_alpha = g_alpha;
    M.S = S;
    M.layout = layout;
    M.b = b;
    M.use_alpha = (use_alpha != 0) && (g_alpha != nullptr);
    M.use_rgb = (rgb != nullptr) && (g_rgb != nullptr);
    const bool want_k4 = M.use_alpha || M.use_rgb;
    __syncthreads();

    if (want_k4) {
        if (hit)
            s_vis[atomicAdd(&s_nvis, 1)] = (unsigned short)tid;
        __syncthreads();
        const int nq = s_nvis * 6; /* (face, edge, axis) groups */
        for (int q = tid; q < nq; q += BW_THREADS) {
            const int lf = s_vis[q / 6];
            const int combo = q - (q / 6) * 6;
            HocK4Edge E;
            hoc_k4_edge(s_face[lf], S, combo >> 1, combo & 1, &E);
            s_pre[q] = max(E.d0_to - E.d0_from + 1, 0);
        }
        __syncthreads();
        if (tid < 32) { /* exclusive prefix sum of the column counts: each lane scans a contiguous chunk */
            const int per = (nq + 31) / 32;
            const int lo = min(tid * per, nq), hi = min(lo + per, nq);
            int sum = 0;
            for (int q = lo; q < hi; q++)
                sum += s_pre[q];
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(HOC_FULL_MASK, incl, o);
                if (tid >= o)
                    incl += v;
            }
            int run = incl - sum;
            for (int q = lo; q < hi; q++) {
                const int c = s_pre[q];
                s_pre[q] = run;
                run += c;
            }
            if (tid == 31)
                s_pre[nq] = incl;
        }
        __syncthreads();
        const int total = s_pre[nq];
        for (int item = tid; item < total; item += BW_THREADS) {
            int lo = 0, hi = nq; /* last q with s_pre[q] <= item */
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s_pre[mid] <= item)
                    lo = mid;
                else
                    hi = mid;
            }
            const int q = lo;
            const int lf = s_vis[q / 6];
            const int combo = q - (q / 6) * 6;
            const int edge = combo >> 1, axis = combo & 1;
            HocK4Edge E;
            hoc_k4_edge(s_face[lf], S, edge, axis, &E);
            float gA, gB;
            hoc_k4_face_column(E, lf * gridDim.x + blockIdx.x, edge, axis, E.d0_from + (item - s_pre[q]), M, eps, line_count, emitters,
                               &gA, &gB);
            if (gA != 0.0f)
                atomicAdd(&s_k4[lf][combo][0], gA);
            if (gB != 0.0f)
                atomicAdd(&s_k4[lf][combo][1], gB);
        }
        __syncthreads();
    }
    if (!valid)
        return;
    float gface[9];
#pragma unroll
    for (int k = 0; k < 9; k++)
        gface[k] = 0.0f;
    if (hit) {
        if (want_depth) {
            float inv[9], tmp[2];
            hoc_face_inv(f, S, inv);
            const float *ad = acc_d + ((long)b * F + fi) * 3;
#pragma unroll
            for (int l = 0; l < 2; l++)
                tmp[l] = inv[l] / f[2] + inv[3 + l] / f[5] + inv[6 + l] / f[8];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float zk = f[3 * k + 2];
                const float a = ad[k];
                gface[3 * k + 2] = a / (zk * zk);
                gface[3 * k + 0] = a * tmp[0] * (float)S / 2.0f;
                gface[3 * k + 1] = a * tmp[1] * (float)S / 2.0f;
            }
        }
        if (want_k4) {
#pragma unroll
            for (int combo = 0; combo < 6; combo++) {
                const int edge = combo >> 1, axis = combo & 1;
                gface[edge * 3 + (1 - axis)] += s_k4[tid][combo][0];
                gface[((edge + 1) % 3) * 3 + (1 - axis)] += s_k4[tid][combo][1];
            }
        }
    }
    float *gf = grad_faces + ((long)b * F + fi) * 9;
#pragma unroll
    for (int k = 0; k < 9; k++)
        gf[k] = gface[k];
}

/*
 * Line pass.  grid (S, 2, B): one CTA per image column (axis 0) or row (axis 1).  The span of the line
 * where the incoming gradient is non-zero is staged in shared memory once (I = (alpha, r, g, b) and
 * dL/dI per pixel); every lane then runs one queued outward scan from shared memory.
 */
#define LN_THREADS 256
__global__ void __launch_bounds__(LN_THREADS)
hoc_raster_bwd_line_kernel(const float *__restrict__ faces, const int32_t *__restrict__ face_index_map,
                           const float *__restrict__ rgb, const float *__restrict__ g_rgb,
                           const float *__restrict__ g_alpha, int F, int S, float eps, int layout, int use_alpha,
                           const int *__restrict__ ext, const int *__restrict__ line_count,
                           const int *__restrict__ emitters, float *__restrict__ grad_faces)
{
    /* per pixel of the span: float4 (P, g_r, g_g, g_b) with P = sum_ch I_ch g_ch (alpha included), then g_alpha:
     * delta = sum_ch (I_ch - Iin_ch) g_ch = P - sum_ch Iin_ch g_ch -- one 16-byte shared load and three FMAs
     * per scanned pixel (<= 1 ulp of |P| from the reference's summation order, gradients carry 1e-3) */
    extern __shared__ float4 s_line4[];
    const int d0 = blockIdx.x, axis = blockIdx.y, b = blockIdx.z;
    const long line = ((long)b * 2 + axis) * S + d0;
    const int n = min(line_count[line], 3 * S);
    if (n <= 0)
        return;
    const int *e = ext + (long)b * 4 * S;
    const int lo = (axis == 0) ? e[EXT_COL_LO * S + d0] : e[EXT_ROW_LO * S + d0];
    const int hi = (axis == 0) ? e[EXT_COL_HI * S + d0] : e[EXT_ROW_HI * S + d0];
    if (lo > hi)
        return; /* no incoming gradient anywhere on this line: every outward scan sums zeros */
    const int len = hi - lo + 1;

    HocBwdMaps M;
    M.idx = face_index_map + (long)b * S * S;
    M.rgb = rgb;
    M.g_rgb = g_rgb;
    M.g_alpha = g_alpha;
    M.S = S;
    M.layout = layout;
    M.b = b;
    M.use_alpha = (use_alpha != 0) && (g_alpha != nullptr);
    M.use_rgb = (rgb != nullptr) && (g_rgb != nullptr);

    for (int i = threadIdx.x; i < len; i += LN_THREADS) {
        const int d1 = lo + i;
        const int xi = axis == 0 ? d0 : d1, yi = axis == 0 ? d1 : d0;
        float I[4];
        hoc_load_I(M, xi, yi, I);
        float g[4] = {0.f, 0.f, 0.f, 0.f};
        if (M.use_alpha)
            g[0] = g_alpha[hoc_plane_off(layout, S, b, yi, xi)];
        if (M.use_rgb) {
#pragma unroll
            for (int k = 0; k < 3; k++)
                g[1 + k] = g_rgb[hoc_rgb_off(layout, S, b, yi, xi, k)];
        }
        s_line4[i] = make_float4(I[0] * g[0] + I[1] * g[1] + I[2] * g[2] + I[3] * g[3], g[1], g[2], g[3]);
        if (M.use_alpha)
            reinterpret_cast<float *>(s_line4 + len)[i] = g[0];
    }
    __syncthreads();
    const float *s_ga = reinterpret_cast<const float *>(s_line4 + len);

    const int *bucket = emitters + line * 3 * S;
    for (int q = threadIdx.x; q < n; q += LN_THREADS) {
        const int rec = bucket[q];
        const int fi = rec & 0x1fffffff;
        const int edge = (rec >> 29) & 3;
        float f[9];
        const float *src = faces + ((long)b * F + fi) * 9;
#pragma unroll
        for (int k = 0; k < 9; k++)
            f[k] = __ldg(src + k);
        HocK4Edge E;
        hoc_k4_edge(f, S, edge, axis, &E);
        float d1_cross;
        int d1_in, d1_out;
        if (!hoc_k4_column(&E, S, d0, &d1_cross, &d1_in, &d1_out))
            continue; /* unreachable: the face pass queued this column because it passed the same test */
        float I_in[4];
        hoc_load_I(M, axis == 0 ? d0 : d1_in, axis == 0 ? d1_in : d0, I_in);
        const int d1_limit = (0 < E.dir) ? S - 1 : 0;
        const int d1_from = max(max(min(d1_out, d1_limit), 0), lo);
        const int d1_to = min(min(max(d1_out, d1_limit), S - 1), hi);
        float gA = 0.0f, gB = 0.0f;
        HocK4Col C;
        hoc_k4_col(&E, S, d0, d1_cross, &C);
        for (int d1 = d1_from; d1 <= d1_to; d1++) {
            const int i = d1 - lo;
            const float4 pg = s_line4[i];
            float delta = pg.x - (I_in[1] * pg.y + I_in[2] * pg.z + I_in[3] * pg.w);
            if (M.use_alpha)
                delta -= I_in[0] * s_ga[i];
            if (!(delta <= 0.0f))
                hoc_k4_accum_col(&C, d1, eps, delta, &gA, &gB);
        }
        float *gf = grad_faces + ((long)b * F + fi) * 9;
        if (gA != 0.0f)
            atomicAdd(gf + edge * 3 + (1 - axis), gA);
        if (gB != 0.0f)
            atomicAdd(gf + ((edge + 1) % 3) * 3 + (1 - axis), gB);
    }
}

extern "C" size_t hoc_raster_backward_workspace_bytes(int B, int F, int S)
{
    if (B <= 0 || S <= 0 || F < 0)
        return 0;
    return hoc_bwd_workspace(nullptr, B, F, S).total;
}

extern "C" int hoc_raster_backward(const float *faces, const float *textures, const int32_t *face_index_map,
                                   const float *rgb, const float *weight_map, const float *depth,
                                   const float *grad_rgb, const float *grad_alpha,
                                   const float *grad_depth, int B, int F, int S, int ts, float near_, float far_,
                                   float eps, int layout, int use_alpha, int tex_grad_mode, float *grad_faces,
                                   float *grad_textures, void *workspace, size_t workspace_bytes, void *stream)
{
    (void)textures;
    HOC_CHECK_ARG(tex_grad_mode == HOC_TEX_GRAD_CUBE || (tex_grad_mode == HOC_TEX_GRAD_VERTEX && ts == 2),
                  "hoc_raster_backward: tex_grad_mode %d (vertex mode needs texture_size 2, got %d)", tex_grad_mode, ts);
    HOC_CHECK_ARG(B >= 0 && F >= 0, "hoc_raster_backward: negative batch (%d) or face count (%d)", B, F);
    HOC_CHECK_ARG(S >= 1 && S <= 2048, "hoc_raster_backward: image_size %d outside [1, 2048]", S);
    HOC_CHECK_ARG(layout == HOC_LAYOUT_RAW || layout == HOC_LAYOUT_IMAGE, "hoc_raster_backward: bad layout %d",
                  layout);
    HOC_CHECK_ARG(B <= 65535, "hoc_raster_backward: batch %d exceeds 65535", B);
    HOC_CHECK_ARG(F < (1 << 29), "hoc_raster_backward: face count %d exceeds 2^29", F);
    HOC_CHECK_ARG(grad_textures == nullptr || ts >= 1, "hoc_raster_backward: texture_size %d", ts);
    HOC_CHECK_ARG(grad_rgb == nullptr || rgb != nullptr, "hoc_raster_backward: grad_rgb given without rgb");
    if (B == 0 || F == 0)
        return HOC_OK;
    HOC_CHECK_ARG(faces != nullptr && face_index_map != nullptr, "hoc_raster_backward: faces / face_index_map NULL");
    if (grad_faces == nullptr && grad_textures == nullptr)
        return HOC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const HocBwdWorkspace w = hoc_bwd_workspace(workspace, B, F, S);
    if (workspace == nullptr || workspace_bytes < w.total) {
        hoc_set_error("hoc_raster_backward: workspace of %zu bytes needed, %zu given", w.total, workspace_bytes);
        return HOC_ERR_WORKSPACE;
    }
    const float *g_alpha = use_alpha ? grad_alpha : nullptr;
    const bool k4 = grad_faces != nullptr && (grad_rgb != nullptr || g_alpha != nullptr);
    const bool want_depth = grad_faces != nullptr && grad_depth != nullptr;
    const size_t tex_bytes = (tex_grad_mode == HOC_TEX_GRAD_VERTEX)
                                 ? sizeof(float) * 9 * (size_t)B * F
                                 : sizeof(float) * 3 * (size_t)ts * ts * ts * (size_t)B * F;

    cudaError_t e = cudaMemsetAsync((char *)workspace + w.zero_begin, 0, w.zero_bytes, st);
    if (e == cudaSuccess && k4) { /* hi rows = -1, lo rows (every second row of S ints) = 0x7f7f7f7f */
        e = cudaMemsetAsync(w.ext, 0xff, sizeof(int) * 4 * (size_t)B * S, st);
        if (e == cudaSuccess)
            e = cudaMemset2DAsync(w.ext, 2 * S * sizeof(int), 0x7f, S * sizeof(int), (size_t)B * 2, st);
    }
    if (e == cudaSuccess && grad_textures != nullptr)
        e = cudaMemsetAsync(grad_textures, 0, tex_bytes, st);
    if (e != cudaSuccess) {
        hoc_set_error("hoc_raster_backward: memset failed: %s", cudaGetErrorString(e));
        return HOC_ERR_CUDA;
    }
    {
        dim3 pg((S + 31) / 32, (S + 31) / 32, B);
        float *gt = (grad_rgb != nullptr) ? grad_textures : nullptr;
        if (ts == 2)
            HOC_LAUNCH(HOC_K_RASTER_BWD_PIXEL, st,
                       (hoc_raster_bwd_pixel_kernel<true><<<pg, dim3(32, 8), 0, st>>>(
                           faces, face_index_map, weight_map, depth, grad_rgb, g_alpha, grad_depth, F, S, ts, near_, far_,
                           eps, layout, k4 ? 1 : 0, tex_grad_mode, w.ext, w.owned, want_depth ? w.acc_d : nullptr, gt)));
        else
            HOC_LAUNCH(HOC_K_RASTER_BWD_PIXEL, st,
                       (hoc_raster_bwd_pixel_kernel<false><<<pg, dim3(32, 8), 0, st>>>(
                           faces, face_index_map, weight_map, depth, grad_rgb, g_alpha, grad_depth, F, S, ts, near_, far_,
                           eps, layout, k4 ? 1 : 0, tex_grad_mode, w.ext, w.owned, want_depth ? w.acc_d : nullptr, gt)));
        HOC_CHECK_LAUNCH("hoc_raster_bwd_pixel_kernel");
    }
    if (grad_faces == nullptr)
        return HOC_OK;
    {
        dim3 grid((F + BW_THREADS - 1) / BW_THREADS, B);
        HOC_LAUNCH(HOC_K_RASTER_BACKWARD, st,
                   (hoc_raster_bwd_face_kernel<<<grid, BW_THREADS, 0, st>>>(
                       faces, face_index_map, rgb, grad_rgb, g_alpha, F, S, eps, layout, use_alpha, want_depth ? 1 : 0,
                       w.owned, w.acc_d, w.line_count, w.emitters, grad_faces)));
        HOC_CHECK_LAUNCH("hoc_raster_bwd_face_kernel");
    }
    if (k4) {
        const size_t smem = sizeof(float) * 5 * (size_t)S + 16;
        if (smem > 48 * 1024) {
            e = cudaFuncSetAttribute(hoc_raster_bwd_line_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) {
                hoc_set_error("hoc_raster_backward: cannot reserve %zu bytes of shared memory: %s", smem,
                              cudaGetErrorString(e));
                return HOC_ERR_CUDA;
            }
        }
        dim3 grid(S, 2, B);
        HOC_LAUNCH(HOC_K_RASTER_BWD_LINE, st,
                   (hoc_raster_bwd_line_kernel<<<grid, LN_THREADS, smem, st>>>(
                       faces, face_index_map, rgb, grad_rgb, g_alpha, F, S, eps, layout, use_alpha, w.ext, w.line_count,
                       w.emitters, grad_faces)));
        HOC_CHECK_LAUNCH("hoc_raster_bwd_line_kernel");
    }
    return HOC_OK;
}
