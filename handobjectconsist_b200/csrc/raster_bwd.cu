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
 *   hoc_raster_bwd_pixel_kernel   pixel-parallel.  Phase 1 streams face_index_map and the incoming gradients
 *           once: per-line spans of non-zero incoming gradient (a pixel with zero incoming gradient
 *           contributes exactly nothing to any scan, so scans are clipped to the span), and for covered
 *           pixels the texture / depth gradient of the owning face.  Phase 2 runs the pseudo-gradient FROM
 *           THE PIXELS: a pixel owned by face f lies on column x and row y of f; for each of the 3 edges x
 *           2 axes it evaluates that column of the edge once and (a) adds its own term of the short INWARD
 *           scan (the reference visits exactly the owned pixels between the edge and the opposite edge),
 *           (b) if it is the pixel just inside the edge, flags the long OUTWARD scan that starts there:
 *           one byte per pixel and axis (3 "edge e starts a scan here" bits + 3 direction bits), written
 *           coalesced -- row-major for row scans, transposed for column scans.  No per-face pass, no
 *           queues, no counters: a scan is identified by (line, position, edge), the face by
 *           face_index_map at that position.
 *   hoc_raster_bwd_depth_kernel   face-parallel epilogue of backward_depth_map (only when dL/ddepth exists).
 *   hoc_raster_bwd_line_kernel    line-parallel: a CTA owns one image row or column, reads the line's flag
 *           bytes, compacts them IN POSITION ORDER into two lists (scans towards +, scans towards -, so
 *           that neighbouring lanes get scans of nearly equal length), stages the span of that line
 *           (P = sum_ch I_ch g_ch and g of every pixel) in shared memory ONCE and lets every lane run one
 *           outward scan out of shared memory; each scan adds its two vertex contributions to grad_faces.
 */
#include "hoc_common.cuh"
#include "raster_math.h"

#define EXT_ROW_LO 0
#define EXT_ROW_HI 1
#define EXT_COL_LO 2
#define EXT_COL_HI 3

/* Workspace carved by hoc_raster_backward (256-byte aligned regions):
 *   ext        int   [B][4][S]     {row_lo, row_hi, col_lo, col_hi}: span of non-zero incoming gradient
 *   cov_count  int   [B]           covered pixels listed per sample                      (zero-filled)
 *   acc_d      float [B][F][3]     sum over owned pixels of dL/ddepth * depth^2 * w_k   (zero-filled)
 *   cov_list   int   [B][S*S]      the covered pixels (yi * S + xi) that have work, in tile order
 *   flags      uint8 [B][2][S][P]  (P = S rounded up to 16) outward-scan flags; plane 0 (column scans) is stored [x][y], plane 1 (row
 *                                  scans) [y][x], so a line's bytes are contiguous.  Bit e: edge e of the face
 *                                  owning this pixel starts an outward scan here; bit 3+e: it runs towards +.
 *                                  Zero-filled by the scan pass, set by the cover pass. */
/* Row pitch of the flag planes: lines start 16-byte aligned so that the line pass can read 4 flags per load. */
__host__ __device__ __forceinline__ int hoc_flag_pitch(int S) { return (S + 15) & ~15; }

struct HocBwdWorkspace {
    int *ext;
    int *cov_count;
    float *acc_d;
    int *cov_list;
    uint8_t *flags;
    size_t count_bytes, acc_bytes;
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
    w.cov_count = (int *)(p + off);
    w.count_bytes = up(sizeof(int) * (size_t)B);
    off += w.count_bytes;
    w.acc_d = (float *)(p + off); /* directly after cov_count: one memset covers both */
    w.acc_bytes = sizeof(float) * 3 * (size_t)B * F;
    off = up(off + w.acc_bytes);
    w.cov_list = (int *)(p + off);
    off = up(off + sizeof(int) * (size_t)B * S * S);
    w.flags = (uint8_t *)(p + off);
    off = up(off + 2 * (size_t)B * S * hoc_flag_pitch(S));
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

/* 1 / x, one MUFU.RCP. */
__device__ __forceinline__ float hoc_rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

/* One (edge, axis) of the face owning pixel (xi, yi): the pixel's term of the inward scan of the column it
 * lies on (added to grad_faces) and, when it is the pixel just inside the edge, the outward-scan flag.
 * (ax..cy) are the face's vertices in NDC rotated so that A is the first vertex of the edge; gfA / gfB point
 * at the x component of vertex A / B in grad_faces.  I / g: (alpha, r, g, b) of the pixel and its incoming
 * gradient. */
__device__ __forceinline__ unsigned hoc_k4_pixel_combo(float ax, float ay, float bx, float by, float cx, float cy,
                                                       int edge, int axis, int xi, int yi, const HocBwdMaps &M,
                                                       const float *I, const float *g, float eps,
                                                       float *__restrict__ gfA, float *__restrict__ gfB)
{
    HocK4Edge E;
    hoc_k4_edge_pts(ax, ay, bx, by, cx, cy, M.S, axis, &E);
    const int d0 = axis == 0 ? xi : yi, d1p = axis == 0 ? yi : xi;
    if (d0 < E.d0_from || d0 > E.d0_to)
        return 0u;
    float d1_cross;
    int d1_in, d1_out;
    if (!hoc_k4_column(&E, M.S, d0, &d1_cross, &d1_in, &d1_out))
        return 0u;
    unsigned flag = 0u;
    if (d1_in == d1p)
        flag = (1u << edge) | ((0 < E.dir) ? (8u << edge) : 0u);
    const int lim = hoc_k4_inward_limit(&E, d0);
    const int d1_from = max(min(d1_in, lim), 0);
    const int d1_to = min(max(d1_in, lim), M.S - 1);
    if (d1_from <= d1p && d1p <= d1_to) {
        float I_out[4];
        hoc_load_I(M, axis == 0 ? d0 : d1_out, axis == 0 ? d1_out : d0, I_out);
        float delta = 0.0f; /* the reference's accumulation order: alpha, r, g, b */
        if (M.use_alpha)
            delta += (I[0] - I_out[0]) * g[0];
        if (M.use_rgb) {
#pragma unroll
            for (int k = 1; k < 4; k++)
                delta += (I[k] - I_out[k]) * g[k];
        }
        if (!(delta <= 0.0f)) {
            HocK4Col C;
            hoc_k4_col(&E, M.S, d0, d1_cross, &C);
            float gA = 0.0f, gB = 0.0f;
            hoc_k4_accum_col(&C, d1p, eps, delta, &gA, &gB);
            if (gA != 0.0f)
                atomicAdd(gfA + (1 - axis), gA);
            if (gB != 0.0f)
                atomicAdd(gfB + (1 - axis), gB);
        }
    }
    return flag;
}

/*
 * Scan pass.  Block (32, 8) covers a 32 x 32 pixel tile (4 rows per thread).  Pure streaming: reads
 * face_index_map and the incoming gradients once, writes the line spans, the list of covered pixels that have
 * work (all of them when the pseudo-gradient is wanted, else those with a texture / depth gradient) and, for the
 * pseudo-gradient, zeroes the tile's flag bytes.  One global atomic per CTA reserves the tile's list slots.
 */
template <bool K4>
__global__ void __launch_bounds__(256)
hoc_raster_bwd_scan_kernel(const int32_t *__restrict__ face_index_map, const float *__restrict__ g_rgb,
                           const float *__restrict__ g_alpha, int S, int layout, int list_all, int *__restrict__ ext,
                           int *__restrict__ cov_count, int *__restrict__ cov_list, uint8_t *__restrict__ flags)
{
    __shared__ int s_lo[8][32];
    __shared__ int s_hi[8][32];
    __shared__ int s_cnt[33]; /* per warp-row (r * 8 + ty) count, then exclusive prefix; [32] = tile base */
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int b = blockIdx.z;
    const int xi = blockIdx.x * 32 + tx;
    int *e = ext + (long)b * 4 * S;
    int c_lo = 0x7f7f7f7f, c_hi = -1;
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
    unsigned want[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int yi = blockIdx.y * 32 + r * 8 + ty;
        const bool nz = !(gr[r][0] == 0.0f) || !(gr[r][1] == 0.0f) || !(gr[r][2] == 0.0f) || !(ga[r] == 0.0f);
        want[r] = __ballot_sync(HOC_FULL_MASK, fis[r] >= 0 && (list_all || nz));
        if (tx == 0)
            s_cnt[r * 8 + ty] = __popc(want[r]);
        if (K4) {
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
    if (K4) {
        s_lo[ty][tx] = c_lo;
        s_hi[ty][tx] = c_hi;
    }
    __syncthreads();
    if (ty == 0) { /* exclusive prefix over the 32 warp-rows + one atomic for the tile */
        const int mine = s_cnt[tx];
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(HOC_FULL_MASK, incl, o);
            if (tx >= o)
                incl += t;
        }
        s_cnt[tx] = incl - mine;
        if (tx == 31)
            s_cnt[32] = (incl > 0) ? atomicAdd(cov_count + b, incl) : 0;
        if (K4 && xi < S) {
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
    __syncthreads();
    int *list = cov_list + (long)b * S * S + s_cnt[32];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int lr = r * 8 + ty;
        const int yi = blockIdx.y * 32 + lr;
        if ((want[r] >> tx) & 1u)
            list[s_cnt[lr] + __popc(want[r] & ((1u << tx) - 1u))] = yi * S + xi;
        if (K4) { /* the tile's flag bytes, 32 contiguous bytes per warp in both planes */
            const int P = hoc_flag_pitch(S);
            uint8_t *fl_col = flags + ((long)b * 2 + 0) * S * P;
            uint8_t *fl_row = flags + ((long)b * 2 + 1) * S * P;
            if (xi < S && yi < S)
                fl_row[(long)yi * P + xi] = 0;
            const int cx = blockIdx.x * 32 + lr, cy = blockIdx.y * 32 + tx;
            if (cx < S && cy < S)
                fl_col[(long)cx * P + cy] = 0;
        }
    }
}

/* Texture (backward_textures) and depth (backward_depth_map) gradient of one covered pixel. */
template <bool TS2>
__device__ __forceinline__ void hoc_cover_tex_depth(const float *__restrict__ faces,
                                                    const float *__restrict__ weight_map,
                                                    const float *__restrict__ depth_map,
                                                    const float *__restrict__ g_rgb,
                                                    const float *__restrict__ g_depth, int b, int fi, int xi, int yi,
                                                    int F, int S, int ts, float near_, float far_, float eps, int layout,
                                                    int tex_mode, float *__restrict__ acc_d,
                                                    float *__restrict__ grad_textures)
{
    const bool want_tex = (grad_textures != nullptr) && (g_rgb != nullptr);
    const bool want_depth = (acc_d != nullptr) && (g_depth != nullptr);
    float gr[3] = {0.0f, 0.0f, 0.0f};
    if (want_tex) {
#pragma unroll
        for (int c = 0; c < 3; c++)
            gr[c] = g_rgb[hoc_rgb_off(layout, S, b, yi, xi, c)];
    }
    const bool nz = !(gr[0] == 0.0f) || !(gr[1] == 0.0f) || !(gr[2] == 0.0f);
    if (!((want_tex && nz) || want_depth))
        return;
    const int tex_n = ts * ts * ts * 3;
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
        /* textures are the multilinear extension of three vertex values (T[i,j,k] = i c0 + j c1 + k c2, ts == 2):
         * d rgb / d c_k = t_k, so nine sums per face instead of twenty-four */
        float *gt = grad_textures + ((long)b * F + fi) * 9;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float t = hoc_tex_coord(w[k], f[3 * k + 2], zp, 2, eps);
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float v = t * gr[c];
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
                    const float v = ww * gr[c];
                    if (v != 0.0f)
                        atomicAdd(gt + isc * 3 + c, v);
                }
            }
        }
    }
}

/*
 * Cover pass: the work of the covered pixels, spread evenly over the GPU (the scan pass listed them).
 * K4 = false: 128 threads, one listed pixel each -> texture / depth gradient.
 * K4 = true:  224 threads work on 32 listed pixels at a time: warp c < 6 runs (edge c >> 1, axis c & 1) of the
 *             pseudo-gradient for the 32 pixels (uniform edge / axis per warp), warp 6 their texture / depth
 *             gradient; warp 0 then merges the six flag contributions and stores the two flag bytes per pixel.
 */
#define CV_THREADS_K4 224
#define CV_THREADS 128
template <bool TS2, bool K4>
__global__ void __launch_bounds__(K4 ? CV_THREADS_K4 : CV_THREADS)
hoc_raster_bwd_cover_kernel(const float *__restrict__ faces, const int32_t *__restrict__ face_index_map,
                            const float *__restrict__ rgb, const float *__restrict__ weight_map,
                            const float *__restrict__ depth_map, const float *__restrict__ g_rgb,
                            const float *__restrict__ g_alpha, const float *__restrict__ g_depth, int F, int S, int ts,
                            float near_, float far_, float eps, int layout, int use_alpha, int tex_mode,
                            const int *__restrict__ cov_count, const int *__restrict__ cov_list,
                            float *__restrict__ acc_d, uint8_t *__restrict__ flags, float *__restrict__ grad_faces,
                            float *__restrict__ grad_textures)
{
    const int b = blockIdx.y;
    const int count = min(cov_count[b], S * S);
    const int *list = cov_list + (long)b * S * S;
    const int32_t *idx = face_index_map + (long)b * S * S;
    if (!K4) {
        for (int i = blockIdx.x * CV_THREADS + threadIdx.x; i < count; i += gridDim.x * CV_THREADS) {
            const int p = list[i];
            const int yi = p / S, xi = p - yi * S;
            hoc_cover_tex_depth<TS2>(faces, weight_map, depth_map, g_rgb, g_depth, b, idx[p], xi, yi, F, S, ts, near_,
                                     far_, eps, layout, tex_mode, acc_d, grad_textures);
        }
        return;
    }
    __shared__ uint8_t s_fl[6][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
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
    for (int base = blockIdx.x * 32; base < count; base += gridDim.x * 32) {
        const int i = base + lane;
        const bool live = i < count;
        const int p = live ? list[i] : 0;
        const int yi = p / S, xi = p - yi * S;
        const int fi = live ? idx[p] : -1;
        if (wid == 6) {
            if (fi >= 0)
                hoc_cover_tex_depth<TS2>(faces, weight_map, depth_map, g_rgb, g_depth, b, fi, xi, yi, F, S, ts, near_,
                                         far_, eps, layout, tex_mode, acc_d, grad_textures);
        } else {
            unsigned fl = 0u;
            if (fi >= 0) {
                const int edge = wid >> 1, axis = wid & 1;
                const int ia = edge, ib = (edge == 2) ? 0 : edge + 1, ic = (edge == 0) ? 2 : edge - 1;
                const float *src = faces + ((long)b * F + fi) * 9;
                const float ax = __ldg(src + 3 * ia), ay = __ldg(src + 3 * ia + 1);
                const float bx = __ldg(src + 3 * ib), by = __ldg(src + 3 * ib + 1);
                const float cx = __ldg(src + 3 * ic), cy = __ldg(src + 3 * ic + 1);
                float I[4] = {1.0f, 0.0f, 0.0f, 0.0f}, g[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                if (M.use_alpha)
                    g[0] = g_alpha[hoc_plane_off(layout, S, b, yi, xi)];
                if (M.use_rgb) {
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        const long o = hoc_rgb_off(layout, S, b, yi, xi, k);
                        I[1 + k] = rgb[o];
                        g[1 + k] = g_rgb[o];
                    }
                }
                /* the owner of a pixel is front-facing with finite xy (the forward's tests); re-checked so that a
                 * corrupted map cannot produce garbage.  Same expression as hoc_face_back on the rotated vertices
                 * is NOT bit-identical, so test the face in its stored order. */
                float f[9];
                f[0] = (edge == 0) ? ax : ((edge == 1) ? cx : bx);
                f[1] = (edge == 0) ? ay : ((edge == 1) ? cy : by);
                f[3] = (edge == 0) ? bx : ((edge == 1) ? ax : cx);
                f[4] = (edge == 0) ? by : ((edge == 1) ? ay : cy);
                f[6] = (edge == 0) ? cx : ((edge == 1) ? bx : ax);
                f[7] = (edge == 0) ? cy : ((edge == 1) ? by : ay);
                f[2] = f[5] = f[8] = 0.0f;
                if (hoc_face_xy_finite(f) && !hoc_face_back(f)) {
                    float *gf = grad_faces + ((long)b * F + fi) * 9;
                    fl = hoc_k4_pixel_combo(ax, ay, bx, by, cx, cy, edge, axis, xi, yi, M, I, g, eps, gf + 3 * ia,
                                            gf + 3 * ib);
                }
            }
            s_fl[wid][lane] = (uint8_t)fl;
        }
        __syncthreads();
        if (wid == 0 && fi >= 0) {
            const unsigned fl0 = s_fl[0][lane] | s_fl[2][lane] | s_fl[4][lane];
            const unsigned fl1 = s_fl[1][lane] | s_fl[3][lane] | s_fl[5][lane];
            const int P = hoc_flag_pitch(S);
            uint8_t *fl_col = flags + ((long)b * 2 + 0) * S * P;
            uint8_t *fl_row = flags + ((long)b * 2 + 1) * S * P;
            if (fl0)
                fl_col[(long)xi * P + yi] = (uint8_t)fl0;
            if (fl1)
                fl_row[(long)yi * P + xi] = (uint8_t)fl1;
        }
        __syncthreads();
    }
}

/* backward_depth_map's per-face epilogue: grad_faces += J^T acc_d (z directly, x / y through the weights). */
__global__ void __launch_bounds__(256)
hoc_raster_bwd_depth_kernel(const float *__restrict__ faces, const float *__restrict__ acc_d, long n_faces, int S,
                            float *__restrict__ grad_faces)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_faces)
        return;
    const float *ad = acc_d + i * 3;
    const float a0 = ad[0], a1 = ad[1], a2 = ad[2];
    if (a0 == 0.0f && a1 == 0.0f && a2 == 0.0f)
        return;
    float f[9], inv[9], tmp[2];
#pragma unroll
    for (int k = 0; k < 9; k++)
        f[k] = __ldg(faces + i * 9 + k);
    if (!hoc_face_xy_finite(f) || hoc_face_back(f))
        return;
    hoc_face_inv(f, S, inv);
#pragma unroll
    for (int l = 0; l < 2; l++)
        tmp[l] = inv[l] / f[2] + inv[3 + l] / f[5] + inv[6 + l] / f[8];
    float *gf = grad_faces + i * 9;
    const float a[3] = {a0, a1, a2};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float zk = f[3 * k + 2];
        gf[3 * k + 2] += a[k] / (zk * zk);
        gf[3 * k + 0] += a[k] * tmp[0] * (float)S / 2.0f;
        gf[3 * k + 1] += a[k] * tmp[1] * (float)S / 2.0f;
    }
}

/*
 * Line pass.  grid (ceil(S / G), 2, B): one CTA per group of G adjacent image columns (axis 0) or rows
 * (axis 1) -- a single line rarely carries enough outward scans to fill the CTA's warps, G lines do, and G
 * adjacent columns are read from HBM as 4 G contiguous bytes per row.
 */
#define LN_THREADS 256
#define LN_WARPS (LN_THREADS / 32)
template <int G>
__global__ void __launch_bounds__(LN_THREADS)
hoc_raster_bwd_line_kernel(const float *__restrict__ faces, const int32_t *__restrict__ face_index_map,
                           const float *__restrict__ rgb, const float *__restrict__ g_rgb,
                           const float *__restrict__ g_alpha, int B, int F, int S, float eps, int layout,
                           int use_alpha, const int *__restrict__ ext, const uint8_t *__restrict__ flags,
                           float *__restrict__ grad_faces, int seg)
{
    /* dynamic shared memory: float4 s_line4[G][S] | ushort s_queue[G * 3 S]
     * per staged pixel: float4 (P, g_r, g_g, g_b) with P = sum_ch I_ch g_ch - g_alpha (the inside pixel of a scan
     * is covered, its alpha is 1): delta = sum_ch (I_ch - Iin_ch) g_ch = P - sum_rgb Iin_ch g_ch -- one 16-byte
     * shared load and three FMAs per scanned pixel (<= 1 ulp of |P| from the reference's summation order,
     * gradients carry 1e-3) */
    extern __shared__ float4 s_line4[];
    __shared__ int s_wtot[LN_WARPS];
    __shared__ int s_lo[G], s_hi[G];
    /* 1-D grid, sample fastest, line groups ordered from the image centre outwards: the lines that carry the
     * most scans (meshes are centred by the crop) are dispatched first, the empty border lines form the tail */
    const int ngroups = (S + G - 1) / G;
    const int b = blockIdx.x % B;
    const int rest = blockIdx.x / B;
    const int axis = rest & 1;
    const int k = rest >> 1;
    const int grp = (ngroups >> 1) + ((k & 1) ? -((k + 1) >> 1) : (k >> 1)); /* c, c-1, c+1, c-2, ...: a bijection */
    const int d0_base = grp * G;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int *e = ext + (long)b * 4 * S;
    int ulo, uhi;
    if (G == 1) { /* a line without incoming gradient: leave before anything else is computed */
        ulo = e[(axis == 0 ? EXT_COL_LO : EXT_ROW_LO) * S + d0_base];
        uhi = e[(axis == 0 ? EXT_COL_HI : EXT_ROW_HI) * S + d0_base];
        if (ulo > uhi)
            return;
        if (tid == 0) {
            s_lo[0] = ulo;
            s_hi[0] = uhi;
        }
        __syncthreads();
    } else {
        if (tid < G) {
            const int d0 = d0_base + tid;
            int lo = 0x7f7f7f7f, hi = -1;
            if (d0 < S) {
                lo = (axis == 0) ? e[EXT_COL_LO * S + d0] : e[EXT_ROW_LO * S + d0];
                hi = (axis == 0) ? e[EXT_COL_HI * S + d0] : e[EXT_ROW_HI * S + d0];
            }
            s_lo[tid] = lo;
            s_hi[tid] = hi;
        }
        __syncthreads();
        ulo = 0x7f7f7f7f;
        uhi = -1;
#pragma unroll
        for (int l = 0; l < G; l++) {
            ulo = min(ulo, s_lo[l]);
            uhi = max(uhi, s_hi[l]);
        }
        if (ulo > uhi)
            return; /* no incoming gradient anywhere on these lines: every outward scan sums zeros */
    }
    const int ulen = uhi - ulo + 1;
    unsigned short *s_queue = reinterpret_cast<unsigned short *>(s_line4 + (size_t)G * S);
    const int cap = G * 3 * S;

    /* 1. the scans of the G lines, compacted in position order (4 positions x G lines at a time): towards + from
     *    the front of the queue (length falls with position), towards - from the back (length grows with
     *    position), so that the lanes of a warp get scans of nearly equal length.  A scan that starts beyond the
     *    span of non-zero gradient of its line has nothing to sum and is dropped here.
     *    entry = position | edge << 11 | line << 13 */
    const int P = hoc_flag_pitch(S);
    const int my_l = tid % G, my_grp = tid / G;
    const int my_d0 = d0_base + my_l;
    const uint8_t *fl = flags + (((long)b * 2 + axis) * S + min(my_d0, S - 1)) * P;
    const int my_lo = s_lo[my_l], my_hi = s_hi[my_l];
    const int T = blockDim.x; /* multiple of 32 and of G, <= LN_THREADS */
    const int POS_PER_ITER = 4 * (T / G);
    int nP = 0, nN = 0;
    for (int base = 0; base < S; base += POS_PER_ITER) {
        const int i0 = base + 4 * my_grp;
        uint32_t v4 = 0u;
        if (i0 < S && my_lo <= my_hi)
            v4 = *reinterpret_cast<const uint32_t *>(fl + i0);
        unsigned mP[4], mN[4];
        int cP = 0, cN = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = i0 + k;
            const unsigned v = (i < S) ? ((v4 >> (8 * k)) & 0xffu) : 0u;
            mP[k] = v & (v >> 3) & 7u;
            mN[k] = v & ~(v >> 3) & 7u;
            if (i + 1 > my_hi)
                mP[k] = 0u;
            if (i - 1 < my_lo)
                mN[k] = 0u;
            cP += __popc(mP[k]);
            cN += __popc(mN[k]);
        }
        const int mine = cP | (cN << 16);
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(HOC_FULL_MASK, incl, o);
            if (lane >= o)
                incl += t;
        }
        if (lane == 31)
            s_wtot[wid] = incl;
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < (T >> 5); w++) {
            const int t = s_wtot[w];
            if (w < wid)
                before += t;
            total += t;
        }
        if (mine != 0) {
            const int excl = before + incl - mine;
            int pP = nP + (excl & 0xffff), pN = nN + (excl >> 16);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int rec = (i0 + k) | (my_l << 13);
#pragma unroll
                for (int ed = 0; ed < 3; ed++) {
                    if (mP[k] & (1u << ed))
                        s_queue[pP++] = (unsigned short)(rec | (ed << 11));
                    if (mN[k] & (1u << ed))
                        s_queue[cap - 1 - (pN++)] = (unsigned short)(rec | (ed << 11));
                }
            }
        }
        nP += total & 0xffff;
        nN += total >> 16;
        __syncthreads();
    }
    const int n = nP + nN;
    if (n == 0)
        return;

    const bool has_alpha = (use_alpha != 0) && (g_alpha != nullptr);
    const bool has_rgb = (rgb != nullptr) && (g_rgb != nullptr);
    const int32_t *idx = face_index_map + (long)b * S * S;

    /* 2. stage the union of the spans of the G lines (zero gradient outside a line's own span) */
    for (int t = tid; t < G * ulen; t += T) {
        /* consecutive threads read consecutive x: along the line for rows, across the lines for columns */
        const int l = (axis == 0) ? t % G : t / ulen;
        const int i = (axis == 0) ? t / G : t % ulen;
        const int d0 = d0_base + l, d1 = ulo + i;
        float4 pg = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (d0 < S) {
            const int xi = axis == 0 ? d0 : d1, yi = axis == 0 ? d1 : d0;
            if (has_rgb) {
                const long o0 = hoc_rgb_off(layout, S, b, yi, xi, 0), o1 = hoc_rgb_off(layout, S, b, yi, xi, 1),
                           o2 = hoc_rgb_off(layout, S, b, yi, xi, 2);
                pg.y = g_rgb[o0];
                pg.z = g_rgb[o1];
                pg.w = g_rgb[o2];
                pg.x = rgb[o0] * pg.y + rgb[o1] * pg.z + rgb[o2] * pg.w;
            }
            if (has_alpha) {
                const float ga = g_alpha[hoc_plane_off(layout, S, b, yi, xi)];
                const float a = (idx[(long)yi * S + xi] >= 0) ? 1.0f : 0.0f;
                pg.x += a * ga - ga;
            }
        }
        s_line4[l * S + i] = pg;
    }
    __syncthreads();

    /* 3. the scans, LN_THREADS at a time.  (a) One lane sets up one scan (edge geometry, colour of the inside
     *    pixel, range) and leaves it in shared memory; (b) the scans are cut into segments of at most `seg`
     *    pixels and every lane sums one segment, so that a warp's lanes finish together however different the
     *    scan lengths are; each segment adds its two vertex contributions to grad_faces. */
    const float scale = 2.0f / (float)S;
    __shared__ float r_cA[LN_THREADS], r_cB[LN_THREADS], r_cross[LN_THREADS], r_I1[LN_THREADS], r_I2[LN_THREADS],
        r_I3[LN_THREADS];
    __shared__ int r_from[LN_THREADS], r_to[LN_THREADS], r_gfA[LN_THREADS], r_gfB[LN_THREADS], r_row[LN_THREADS];
    __shared__ int s_pre[LN_THREADS + 1];
    for (int q0 = 0; q0 < n; q0 += T) {
        const int q = q0 + tid;
        int nseg = 0;
        if (q < n) {
            const int rec = (q < nP) ? s_queue[q] : s_queue[cap - 1 - (q - nP)];
            const int d1_in = rec & 0x7ff, edge = (rec >> 11) & 3, l = rec >> 13;
            const int d0 = d0_base + l;
            const int lo = s_lo[l], hi = s_hi[l];
            const int xin = axis == 0 ? d0 : d1_in, yin = axis == 0 ? d1_in : d0;
            const int fi = idx[(long)yin * S + xin];
            float I1 = 0.0f, I2 = 0.0f, I3 = 0.0f; /* rgb of the inside pixel (its alpha is 1: folded into P) */
            if (has_rgb) {
                I1 = rgb[hoc_rgb_off(layout, S, b, yin, xin, 0)];
                I2 = rgb[hoc_rgb_off(layout, S, b, yin, xin, 1)];
                I3 = rgb[hoc_rgb_off(layout, S, b, yin, xin, 2)];
            }
            if (fi >= 0) { /* always: the cover pass flags owned pixels only */
                const int ia = edge, ib = (edge == 2) ? 0 : edge + 1;
                const float *src = faces + ((long)b * F + fi) * 9;
                const float ax = __ldg(src + 3 * ia), ay = __ldg(src + 3 * ia + 1);
                const float bx = __ldg(src + 3 * ib), by = __ldg(src + 3 * ib + 1);
                HocK4Edge E;
                hoc_k4_edge_pts(ax, ay, bx, by, 0.0f, 0.0f, S, axis, &E);
                float d1_cross;
                int d1_chk, d1_out;
                /* always true: the cover pass flagged this column because it passed the same test */
                if (hoc_k4_column(&E, S, d0, &d1_cross, &d1_chk, &d1_out)) {
                    const int d1_from = (0 < E.dir) ? max(d1_out, lo) : lo;
                    const int d1_to = (0 < E.dir) ? hi : min(d1_out, hi);
                    HocK4Col C;
                    hoc_k4_col(&E, S, d0, d1_cross, &C);
                    /* a vertex that gets no contribution: infinite distance -> 1 / dist = 0 (d1 - cross != 0) */
                    r_cA[tid] = C.hasA ? C.cA * scale : __int_as_float(0x7f800000);
                    r_cB[tid] = C.hasB ? C.cB * scale : __int_as_float(0x7f800000);
                    r_cross[tid] = d1_cross;
                    r_I1[tid] = I1;
                    r_I2[tid] = I2;
                    r_I3[tid] = I3;
                    r_from[tid] = d1_from;
                    r_to[tid] = d1_to;
                    r_row[tid] = l * S - ulo;
                    const int gbase = (int)(((long)b * F + fi) * 9) + (1 - axis);
                    r_gfA[tid] = gbase + ia * 3;
                    r_gfB[tid] = gbase + ib * 3;
                    if (d1_to >= d1_from)
                        nseg = (d1_to - d1_from + seg) / seg;
                }
            }
        }
        /* exclusive prefix of the segment counts */
        int incl = nseg;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(HOC_FULL_MASK, incl, o);
            if (lane >= o)
                incl += t;
        }
        if (lane == 31)
            s_wtot[wid] = incl;
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < (T >> 5); w++) {
            const int t = s_wtot[w];
            if (w < wid)
                before += t;
            total += t;
        }
        s_pre[tid] = before + incl - nseg;
        if (tid == 0)
            s_pre[T] = total;
        __syncthreads();
        for (int j = tid; j < total; j += T) {
            int a = 0, c = T; /* last scan with s_pre[scan] <= j */
            while (c - a > 1) {
                const int mid = (a + c) >> 1;
                if (s_pre[mid] <= j)
                    a = mid;
                else
                    c = mid;
            }
            const int d1_from = r_from[a] + (j - s_pre[a]) * seg;
            const int d1_to = min(r_to[a], d1_from + seg - 1);
            const float cA = r_cA[a], cB = r_cB[a], I1 = r_I1[a], I2 = r_I2[a], I3 = r_I3[a];
            const float peps = eps, neps = -eps;
            float gA = 0.0f, gB = 0.0f;
            float u = (float)d1_from - r_cross[a];
            const float4 *sp = s_line4 + (r_row[a] + d1_from);
            /* branch-free body: a pixel with delta <= 0 adds 0 * (1 / dist); MUFU.RCP (2 ulp) for 1 / dist -- the
             * pseudo-gradient carries a 1e-3 tolerance and this quotient is the hot instruction of the pass */
            for (int k = d1_to - d1_from; k >= 0; k--, sp++, u += 1.0f) {
                const float4 pg = *sp;
                float delta = pg.x - __fmaf_rn(I3, pg.w, __fmaf_rn(I2, pg.z, I1 * pg.y));
                delta = (delta <= 0.0f) ? 0.0f : delta;
                float dA = cA * u, dB = cB * u;
                dA += (0.0f < dA) ? peps : neps;
                dB += (0.0f < dB) ? peps : neps;
                gA = __fmaf_rn(-delta, hoc_rcp_approx(dA), gA);
                gB = __fmaf_rn(-delta, hoc_rcp_approx(dB), gB);
            }
            if (gA != 0.0f)
                atomicAdd(grad_faces + r_gfA[a], gA);
            if (gB != 0.0f)
                atomicAdd(grad_faces + r_gfB[a], gB);
        }
        __syncthreads();
    }
}

/* Tuning knobs of the line pass (hoc_set_tuning): lines per CTA (0 = by image size), threads per CTA, segment
 * length in pixels. */
static int g_line_G = 0, g_line_threads = 128, g_line_seg = 16;

extern "C" int hoc_set_tuning(int key, int value)
{
    if (key == HOC_TUNE_LINE_GROUP && (value == 0 || value == 1 || value == 2 || value == 4 || value == 8))
        g_line_G = value;
    else if (key == HOC_TUNE_LINE_THREADS && value >= 32 && value <= LN_THREADS && value % 32 == 0)
        g_line_threads = value;
    else if (key == HOC_TUNE_LINE_SEGMENT && value >= 1 && value <= 4096)
        g_line_seg = value;
    else {
        hoc_set_error("hoc_set_tuning: bad key %d / value %d", key, value);
        return HOC_ERR_INVALID_ARG;
    }
    return HOC_OK;
}

template <int G>
static cudaError_t hoc_launch_line(const float *faces, const int32_t *face_index_map, const float *rgb,
                                   const float *grad_rgb, const float *g_alpha, int B, int F, int S, float eps,
                                   int layout, int use_alpha, const HocBwdWorkspace &w, float *grad_faces,
                                   cudaStream_t st)
{
    static size_t smem_allowed = 32 * 1024; /* static arrays of the kernel take ~13 KB of the default 48 KB */
    const size_t smem = (size_t)G * S * (sizeof(float4) + 3 * sizeof(unsigned short));
    if (smem > smem_allowed) {
        cudaError_t e = cudaFuncSetAttribute(hoc_raster_bwd_line_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess)
            return e;
        smem_allowed = smem;
    }
    const unsigned grid = (unsigned)((S + G - 1) / G) * 2u * (unsigned)B;
    HOC_LAUNCH(HOC_K_RASTER_BWD_LINE, st,
               (hoc_raster_bwd_line_kernel<G><<<grid, g_line_threads, smem, st>>>(
                   faces, face_index_map, rgb, grad_rgb, g_alpha, B, F, S, eps, layout, use_alpha, w.ext, w.flags,
                   grad_faces, g_line_seg)));
    return cudaSuccess;
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

    /* cov_count and (directly behind it) acc_d are zero-filled by one memset */
    cudaError_t e = cudaMemsetAsync(w.cov_count, 0, w.count_bytes + (want_depth ? w.acc_bytes : 0), st);
    if (e == cudaSuccess && grad_faces != nullptr) /* accumulated with atomics by the cover and line passes */
        e = cudaMemsetAsync(grad_faces, 0, sizeof(float) * 9 * (size_t)B * F, st);
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
    float *gt = (grad_rgb != nullptr) ? grad_textures : nullptr;
    {
        /* without the pseudo-gradient only pixels with a texture gradient (non-zero dL/drgb) or a depth
         * gradient have work */
        dim3 pg((S + 31) / 32, (S + 31) / 32, B);
        const int list_all = (k4 || want_depth) ? 1 : 0;
        if (k4)
            HOC_LAUNCH(HOC_K_RASTER_BWD_PIXEL, st,
                       (hoc_raster_bwd_scan_kernel<true><<<pg, dim3(32, 8), 0, st>>>(
                           face_index_map, grad_rgb, g_alpha, S, layout, list_all, w.ext, w.cov_count, w.cov_list,
                           w.flags)));
        else
            HOC_LAUNCH(HOC_K_RASTER_BWD_PIXEL, st,
                       (hoc_raster_bwd_scan_kernel<false><<<pg, dim3(32, 8), 0, st>>>(
                           face_index_map, gt != nullptr ? grad_rgb : nullptr, nullptr, S, layout, list_all, w.ext,
                           w.cov_count, w.cov_list, w.flags)));
        HOC_CHECK_LAUNCH("hoc_raster_bwd_scan_kernel");
    }
    {
        const long npix = (long)S * S;
        const int per = k4 ? 32 : CV_THREADS;
        dim3 cg((unsigned)((npix + per - 1) / per < 296 ? (npix + per - 1) / per : 296), B);
#define HOC_COVER_LAUNCH(TS2, K4)                                                                                    \
    HOC_LAUNCH(K4 ? HOC_K_RASTER_BWD_PIXEL_K4 : HOC_K_RASTER_BACKWARD_COVER, st,                                      \
               (hoc_raster_bwd_cover_kernel<TS2, K4><<<cg, K4 ? CV_THREADS_K4 : CV_THREADS, 0, st>>>(                 \
                   faces, face_index_map, rgb, weight_map, depth, grad_rgb, g_alpha, grad_depth, F, S, ts, near_, far_, \
                   eps, layout, use_alpha, tex_grad_mode, w.cov_count, w.cov_list, want_depth ? w.acc_d : nullptr,     \
                   w.flags, grad_faces, gt)))
        if (ts == 2 && k4)
            HOC_COVER_LAUNCH(true, true);
        else if (ts == 2)
            HOC_COVER_LAUNCH(true, false);
        else if (k4)
            HOC_COVER_LAUNCH(false, true);
        else
            HOC_COVER_LAUNCH(false, false);
#undef HOC_COVER_LAUNCH
        HOC_CHECK_LAUNCH("hoc_raster_bwd_cover_kernel");
    }
    if (grad_faces == nullptr)
        return HOC_OK;
    if (want_depth) {
        const long nf = (long)B * F;
        HOC_LAUNCH(HOC_K_RASTER_BACKWARD, st,
                   (hoc_raster_bwd_depth_kernel<<<(unsigned)((nf + 255) / 256), 256, 0, st>>>(faces, w.acc_d, nf, S,
                                                                                              grad_faces)));
        HOC_CHECK_LAUNCH("hoc_raster_bwd_depth_kernel");
    }
    if (k4) {
        /* lines per CTA: bounded by staging + queue (22 bytes per pixel of a line) */
        int G = g_line_G ? g_line_G : 1;
        while (G > 1 && (size_t)G * S * 22 > 96 * 1024)
            G >>= 1;
#define HOC_LINE(G_) hoc_launch_line<G_>(faces, face_index_map, rgb, grad_rgb, g_alpha, B, F, S, eps, layout, use_alpha, w, grad_faces, st)
        e = (G == 8) ? HOC_LINE(8) : (G == 4) ? HOC_LINE(4) : (G == 2) ? HOC_LINE(2) : HOC_LINE(1);
#undef HOC_LINE
        if (e != cudaSuccess) {
            hoc_set_error("hoc_raster_backward: cannot reserve shared memory for the line pass: %s",
                          cudaGetErrorString(e));
            return HOC_ERR_CUDA;
        }
        HOC_CHECK_LAUNCH("hoc_raster_bwd_line_kernel");
    }
    return HOC_OK;
}
