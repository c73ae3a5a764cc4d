/*
 * raster_fwd.cu -- rasterizer forward for sm_100a.
 *
 * Replaces forward_face_index_map + forward_texture_sampling of `neural_renderer.cuda.rasterize`
 * (bound at /root/reference/meshreg/neurender/rasterize.py:202-215,232-243) and the wrapper ops
 * around them (buffer fills :58-85, background :252-260, alpha :246-249, clones :118-124, and
 * the permute + row-flip gathers of rasterize_rgbad :417-428).
 *
 * The reference walks ALL faces from every pixel (B*S*S*F triangle tests).  Here the work is
 * proportional to what is actually covered:
 *
 *   pass 1  hoc_raster_zbuf_kernel      face-parallel.  A warp owns 32 faces: each lane sets up
 *           its own face (cull, barycentric matrix, clipped pixel bounding box).  Faces of a few
 *           pixels (the common case for a 9k-triangle mesh at 256x256) are swept by their own lane;
 *           larger ones are staged in shared memory and swept by the whole warp, one face
 *           at a time with 32 pixels in flight.  Every pixel that passes the three edge tests
 *           and the near/far test does ONE 64-bit atomicMin on a packed (depth, face) key:
 *           the minimum is the nearest depth and, among exact ties, the lowest face index --
 *           the result of the reference's in-order strict `<` loop.  The key buffer (8 B/px)
 *           stays in L2.
 *   pass 2  hoc_raster_resolve_kernel   pixel-parallel, one thread per pixel, streaming: reads
 *           the key, recomputes weights/depth of the winning face with the same functions
 *           (bit-identical), samples the texture cube, blends the background and writes every
 *           requested map once, coalesced, directly in the layout the caller returns.
 */
#include "hoc_common.cuh"
#include "raster_math.h"

#ifndef ZB_WARPS
#define ZB_WARPS 2 /* 64-thread CTAs: 2 304 of them per 16-sample render balance the uneven faces better than 128-thread ones (23.6 -> 20.7 us) */
#endif
#define ZB_THREADS (ZB_WARPS * 32)
#define ZB_FACES_PER_WARP 32 /* one face per lane before culling */
#define ZB_REC 24            /* floats per surviving face: 9 coordinates, 9 inverse, x0, y0, width, 1/width, id */

/* clipped pixel bounding box of a face: inside => pmin <= xi <= pmax in exact arithmetic; 1/16 pixel of slack absorbs
 * the fp32 rounding of the NDC -> pixel map (~1e-4 px at S = 2048) and of the edge tests (a pixel centre further than
 * ~1e-4 px outside the exact triangle cannot pass them).  Half a pixel of slack, as in round 1, made a sub-pixel face
 * -- the common case for a 9k-triangle mesh at 256 x 256 -- test 4-9 pixels where 0-2 can be inside.  false when empty. */
#define ZB_SLACK 0.0625f
__device__ __forceinline__ bool hoc_face_bbox(const float *f, int S, int *x0, int *y0, int *x1, int *y1)
{
    const float pxmin = hoc_ndc_to_pix(fminf(f[0], fminf(f[3], f[6])), S);
    const float pxmax = hoc_ndc_to_pix(fmaxf(f[0], fmaxf(f[3], f[6])), S);
    const float pymin = hoc_ndc_to_pix(fminf(f[1], fminf(f[4], f[7])), S);
    const float pymax = hoc_ndc_to_pix(fmaxf(f[1], fmaxf(f[4], f[7])), S);
    const float fS1 = (float)(S - 1);
    const float x_lo = fmaxf(ceilf(pxmin - ZB_SLACK), 0.0f);
    const float x_hi = fminf(floorf(pxmax + ZB_SLACK), fS1);
    const float y_lo = fmaxf(ceilf(pymin - ZB_SLACK), 0.0f);
    const float y_hi = fminf(floorf(pymax + ZB_SLACK), fS1);
    if (!(x_lo <= x_hi && y_lo <= y_hi))
        return false;
    *x0 = (int)x_lo;
    *y0 = (int)y_lo;
    *x1 = (int)x_hi;
    *y1 = (int)y_hi;
    return true;
}

/*
 * A warp culls 64 faces, compacts the survivors (about half of a closed, back-filled mesh) into shared
 * memory together with their barycentric matrices and pixel bounding boxes, and then walks the
 * CONCATENATION of all bounding boxes with one pixel per lane: every lane always has a pixel to test,
 * whatever the sizes of the individual faces (a 9k-triangle mesh at 256x256 has faces of 2-30 pixels,
 * a silhouette test has two triangles of 30 000).
 */
#ifdef ZB_MINB /* (occupancy experiments: minimum resident CTAs per SM) */
#define ZB_BOUNDS __launch_bounds__(ZB_THREADS, ZB_MINB)
#else
#define ZB_BOUNDS __launch_bounds__(ZB_THREADS)
#endif
__global__ void ZB_BOUNDS
hoc_raster_zbuf_kernel(const float *__restrict__ faces, unsigned long long *__restrict__ zbuf, int F, int S,
                       float near_, float far_, const int *__restrict__ row_lo)
{
    hoc_pdl_sync(); /* programmatic dependent launch: see hoc_common.cuh */
    extern __shared__ float s_centre[]; /* [S] pixel-centre NDC coordinate of index i */
    __shared__ float s_rec[ZB_WARPS][ZB_FACES_PER_WARP][ZB_REC];
    __shared__ int s_pre[ZB_WARPS][ZB_FACES_PER_WARP + 1];
    __shared__ int s_queue[ZB_WARPS][64]; /* pixels that passed the edge tests: slot | x << 8 | y << 20 */

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int y_first = (row_lo != nullptr) ? row_lo[b] : 0; /* raster rows below the sample's window are not drawn */

    for (int i = threadIdx.x; i < S; i += ZB_THREADS)
        s_centre[i] = hoc_pix_centre(i, S);

    /* faces are dealt to the warps round-robin (warp w takes faces w, w + W, w + 2W, ...): neighbouring
     * faces of a mesh have similar screen size, so contiguous chunks would give some warps (the hand, close
     * to the camera) several times the pixels of others */
    const int n_warps = gridDim.x * ZB_WARPS;
    const int gw = blockIdx.x * ZB_WARPS + warp;
    int n_surv = 0, n_pix = 0;
#pragma unroll
    for (int h = 0; h < ZB_FACES_PER_WARP / 32; h++) {
        const int fi = (h * 32 + lane) * n_warps + gw;
        float f[9];
        int x0 = 0, y0 = 0, x1 = -1, y1 = -1;
        bool keep = false;
        if (fi < F) {
            const float *src = faces + ((long)b * F + fi) * 9;
#pragma unroll
            for (int k = 0; k < 9; k++)
                f[k] = __ldg(src + k);
            keep = hoc_face_xy_finite(f) && !hoc_face_back(f) && hoc_face_bbox(f, S, &x0, &y0, &x1, &y1);
            if (keep && y0 < y_first) {
                y0 = y_first;
                keep = y0 <= y1;
            }
        }
        const unsigned m = __ballot_sync(HOC_FULL_MASK, keep);
        const int cnt = keep ? (x1 - x0 + 1) * (y1 - y0 + 1) : 0;
        /* inclusive scan of the pixel counts over the lanes */
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(HOC_FULL_MASK, incl, o);
            if (lane >= o)
                incl += v;
        }
        if (keep) {
            const int slot = n_surv + __popc(m & ((1u << lane) - 1u));
            float *rec = s_rec[warp][slot];
            float inv[9];
            hoc_face_inv(f, S, inv);
#pragma unroll
            for (int k = 0; k < 9; k++) {
                rec[k] = f[k];
                rec[9 + k] = inv[k];
            }
            rec[18] = __int_as_float(x0);
            rec[19] = __int_as_float(y0);
            rec[20] = __int_as_float(x1 - x0 + 1);
            rec[21] = 1.0f / (float)(x1 - x0 + 1);
            rec[22] = __int_as_float(fi);
            s_pre[warp][slot] = n_pix + incl - cnt;
        }
        n_surv += __popc(m);
        n_pix += __shfl_sync(HOC_FULL_MASK, incl, 31);
    }
    static_assert(ZB_FACES_PER_WARP == 32, "one prefix-sum entry per lane");
    if (lane >= n_surv)
        s_pre[warp][lane] = (lane == n_surv) ? n_pix : 0x7fffffff; /* sentinels: the slot search below needs no bound */
    if (lane == 0)
        s_pre[warp][ZB_FACES_PER_WARP] = (n_surv == ZB_FACES_PER_WARP) ? n_pix : 0x7fffffff;
    __syncthreads(); /* s_centre ready; orders the record writes */

    unsigned long long *zb = zbuf + (long)b * S * S;
    const int *pre = s_pre[warp];
    /* Two-stage loop.  Stage 1 (cheap, every lane busy): one bounding-box pixel per lane, three edge tests.
     * Pixels that pass (~15 %) are queued in shared memory; stage 2 (the IEEE divisions of the clamped
     * barycentric weights and of the perspective depth, then the atomicMin) runs whenever 32 of them are
     * waiting, again with every lane busy. */
    int *queue = s_queue[warp];
    int qn = 0; /* warp-uniform */
    for (int base = 0; base < n_pix || qn > 0; base += 32) {
        const int item = base + lane;
        bool hit = false;
        int packed = 0;
        if (item < n_pix) {
            int lo = 0; /* last slot with pre[slot] <= item: branch-free descent over the padded prefix sums */
#pragma unroll
            for (int st = ZB_FACES_PER_WARP / 2; st >= 1; st >>= 1)
                if (pre[lo + st] <= item)
                    lo += st;
            const float *rec = s_rec[warp][lo];
            const int p = item - pre[lo];
            const int bw = __float_as_int(rec[20]);
            const int yy = (int)(((float)p + 0.5f) * rec[21]); /* p / bw for p < 2^22 */
            const int xi = __float_as_int(rec[18]) + (p - yy * bw);
            const int yi = __float_as_int(rec[19]) + yy;
            hit = hoc_pixel_inside(rec, s_centre[xi], s_centre[yi]);
            packed = lo | (xi << 8) | (yi << 20);
        }
        const unsigned m = __ballot_sync(HOC_FULL_MASK, hit);
        if (hit)
            queue[qn + __popc(m & ((1u << lane) - 1u))] = packed;
        qn += __popc(m);
        __syncwarp();
        const bool flush = base + 32 >= n_pix; /* last pass over the boxes: drain what is left */
        while (qn >= 32 || (flush && qn > 0)) {
            const int take = min(qn, 32);
            if (lane < take) {
                const int pk = queue[qn - take + lane];
                const float *rec = s_rec[warp][pk & 0xff];
                const int xi = (pk >> 8) & 0xfff, yi = (pk >> 20) & 0xfff;
                float w[3], zp;
                if (hoc_pixel_weights_depth(rec, rec + 9, xi, yi, near_, far_, w, &zp) && zp < far_) {
                    /* (a NaN depth fails `zp < far` and never wins a `<` comparison in the reference either) */
                    const unsigned long long key =
                        ((unsigned long long)hoc_float_order(zp) << 32) | (unsigned)__float_as_int(rec[22]);
                    atomicMin(zb + (long)yi * S + xi, key);
                }
            }
            qn -= take;
            __syncwarp();
        }
    }
}

#ifndef RS_THREADS
#define RS_THREADS 256
#endif

/* Stage `n_per` floats per thread in shared memory and store them as one contiguous, coalesced
 * run of RS_THREADS*n_per floats starting at dst (bounded by `limit` floats). */
template <int N_PER>
__device__ __forceinline__ void hoc_store_interleaved(float *smem, const float *vals, float *dst, long limit)
{
#pragma unroll
    for (int k = 0; k < N_PER; k++)
        smem[threadIdx.x * N_PER + k] = vals[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < N_PER; k++) {
        const int o = k * RS_THREADS + threadIdx.x;
        if (o < limit)
            dst[o] = smem[o];
    }
    __syncthreads();
}

/* The work of one covered pixel: barycentric matrix, weights and depth of the winning face with the functions the
 * z-buffer pass used (bit-identical), then the texture sample.  cube mode: [ts,ts,ts,3] texels per face.  vertex mode
 * (ts == 2): three vertex values c0, c1, c2 per face; the texel of corner (i, j, k) is i c0 + j c1 + k c2, evaluated with
 * hoc_mesh_gather's expression so that the sample is bit-identical to the one taken from the materialised cube. */
__device__ __forceinline__ void hoc_resolve_covered(const float *__restrict__ faces, const float *__restrict__ textures,
                                                    int b, int F, int S, int ts, float near_, float far_, float eps,
                                                    int tex_vertex, bool want_rgb, int fidx, int xi, int yi, float *w,
                                                    float *inv, float *zp, float *col)
{
    float f[9];
    const float *src = faces + ((long)b * F + fidx) * 9;
#pragma unroll
    for (int k = 0; k < 9; k++)
        f[k] = __ldg(src + k);
    hoc_face_inv(f, S, inv);
    hoc_pixel_weights_depth(f, inv, xi, yi, near_, far_, w, zp);
    if (!want_rgb)
        return;
    const float *tex = textures + ((long)b * F + fidx) * (tex_vertex ? 9 : ts * ts * ts * 3);
    float cv[3][3];
    if (tex_vertex) {
#pragma unroll
        for (int k = 0; k < 9; k++)
            cv[k / 3][k % 3] = __ldg(tex + k);
    }
    float tf[3];
    int ti[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float t = hoc_tex_coord(w[k], f[3 * k + 2], *zp, ts, eps);
        ti[k] = hoc_tex_cell(t, ts);
        tf[k] = t - (float)ti[k];
    }
    float acc[3] = {0.f, 0.f, 0.f};
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
        /* a tap with zero weight may point one past the cube when ts == 1 */
        if (ts == 1)
            isc = 0;
        if (tex_vertex) {
            const float wi = (float)((isc >> 2) & 1), wj = (float)((isc >> 1) & 1), wk = (float)(isc & 1);
#pragma unroll
            for (int c = 0; c < 3; c++)
                acc[c] += ww * (wi * cv[0][c] + wj * cv[1][c] + wk * cv[2][c]);
        } else {
#pragma unroll
            for (int c = 0; c < 3; c++)
                acc[c] += ww * __ldg(tex + isc * 3 + c);
        }
    }
    col[0] = acc[0];
    col[1] = acc[1];
    col[2] = acc[2];
}

__global__ void __launch_bounds__(RS_THREADS)
hoc_raster_resolve_kernel(const float *__restrict__ faces, const float *__restrict__ textures,
                          const unsigned long long *__restrict__ zbuf, int F, int S, int ts, float near_, float far_,
                          float eps, float bg0, float bg1, float bg2, const float *__restrict__ bg_dev, int layout,
                          int tex_vertex, int sparse_saved, float *__restrict__ rgb, float *__restrict__ alpha,
                          float *__restrict__ depth, int32_t *__restrict__ face_index_map,
                          float *__restrict__ weight_map, float *__restrict__ face_inv_map)
{
    __shared__ float s_stage[RS_THREADS * 9];

    const int b = blockIdx.y;
    const long npix = (long)S * S;
    const long pix0 = (long)blockIdx.x * RS_THREADS;
    const long pix = pix0 + threadIdx.x;
    const bool in_range = pix < npix;

    int fidx = -1;
    int yi = 0, xi = 0;
    float w[3] = {0.f, 0.f, 0.f};
    float inv[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float zp = far_;
    float col[3];
    if (bg_dev != nullptr) {
        col[0] = __ldg(bg_dev + b * 3 + 0);
        col[1] = __ldg(bg_dev + b * 3 + 1);
        col[2] = __ldg(bg_dev + b * 3 + 2);
    } else {
        col[0] = bg0;
        col[1] = bg1;
        col[2] = bg2;
    }

    if (in_range) {
        yi = (int)((unsigned)pix / (unsigned)S);
        xi = (int)pix - yi * S;
        const unsigned long long key = zbuf[(long)b * npix + pix];
        fidx = (int)(unsigned)(key & 0xffffffffull);
        if (fidx >= 0)
            hoc_resolve_covered(faces, textures, b, F, S, ts, near_, far_, eps, tex_vertex, rgb != nullptr, fidx, xi, yi, w,
                                inv, &zp, col);
    }

    /* ---- stores ---- */
    if (in_range) {
        face_index_map[(long)b * npix + pix] = fidx;
        const long po = hoc_plane_off(layout, S, b, yi, xi);
        if (alpha != nullptr)
            alpha[po] = (fidx >= 0) ? 1.0f : 0.0f;
        if (depth != nullptr && (!sparse_saved || fidx >= 0))
            depth[po] = zp;
        if (sparse_saved && weight_map != nullptr && fidx >= 0) { /* covered pixels only: 7 % of 12 bytes per pixel */
            float *wd = weight_map + ((long)b * npix + pix) * 3;
            wd[0] = w[0];
            wd[1] = w[1];
            wd[2] = w[2];
        }
        if (rgb != nullptr && layout == HOC_LAYOUT_IMAGE) {
#pragma unroll
            for (int c = 0; c < 3; c++)
                rgb[hoc_rgb_off(layout, S, b, yi, xi, c)] = col[c];
        }
    }
    const long left = npix - pix0; /* pixels of this block that exist */
    if (rgb != nullptr && layout == HOC_LAYOUT_RAW)
        hoc_store_interleaved<3>(s_stage, col, rgb + ((long)b * npix + pix0) * 3, left * 3);
    if (weight_map != nullptr && !sparse_saved)
        hoc_store_interleaved<3>(s_stage, w, weight_map + ((long)b * npix + pix0) * 3, left * 3);
    if (face_inv_map != nullptr)
        hoc_store_interleaved<9>(s_stage, inv, face_inv_map + ((long)b * npix + pix0) * 9, left * 9);
}

/*
 * Resolve pass, image layout, four pixels per thread (S % 4 == 0, no face_inv_map, weight_map sparse or absent): what the
 * flow path and rasterize_rgbad's default outputs need.  Two phases per CTA (4 pixels per thread, one sample).  A: every thread
 * streams its four pixels -- two 16-byte loads of the keys, 16-byte stores of face_index_map, alpha, the background
 * colour and (dense mode) the background depth -- and notes the covered ones in a shared list.  B: the covered pixels
 * (a few per cent of a frame, clustered in a few CTAs) are dealt one per thread: barycentric matrix, weights, depth and
 * texture sample of the winning face are a chain of dependent loads and IEEE divisions, and a thread that ran its own
 * four covered pixels one after the other would be the tail of the launch (measured in round 1: 19.6 us against 15.8).
 */
#ifndef RS4_THREADS
#define RS4_THREADS 128 /* (128 vs 256: 14.6 / 15.7 us) */
#endif
#ifdef RS4_MINB /* (occupancy experiments: minimum resident CTAs per SM) */
#define RS4_BOUNDS __launch_bounds__(RS4_THREADS, RS4_MINB)
#else
#define RS4_BOUNDS __launch_bounds__(RS4_THREADS)
#endif
__global__ void RS4_BOUNDS
hoc_raster_resolve4_kernel(const float *__restrict__ faces, const float *__restrict__ textures,
                           const unsigned long long *__restrict__ zbuf, int F, int S, int ts, float near_, float far_,
                           float eps, float bg0, float bg1, float bg2, const float *__restrict__ bg_dev, int tex_vertex,
                           int sparse_saved, float *__restrict__ rgb, float *__restrict__ alpha,
                           float *__restrict__ depth, int32_t *__restrict__ face_index_map,
                           float *__restrict__ weight_map, const int *__restrict__ row_lo)
{
    hoc_pdl_sync(); /* programmatic dependent launch: see hoc_common.cuh */
    __shared__ unsigned short s_list[RS4_THREADS * 4];
    __shared__ int s_n;
    const int b = blockIdx.x; /* sample fastest, chunks of rows from the image centre outwards (hoc_centre_out) */
    const int chunk = hoc_centre_out(blockIdx.y, gridDim.y);
    const int S4 = S >> 2;
    const long npix = (long)S * S;
    const int q = chunk * RS4_THREADS + threadIdx.x;
    if (threadIdx.x == 0)
        s_n = 0;
    __syncthreads();
    float col[3];
    if (bg_dev != nullptr) {
        col[0] = __ldg(bg_dev + b * 3 + 0);
        col[1] = __ldg(bg_dev + b * 3 + 1);
        col[2] = __ldg(bg_dev + b * 3 + 2);
    } else {
        col[0] = bg0;
        col[1] = bg1;
        col[2] = bg2;
    }
    const int y_first = (row_lo != nullptr) ? row_lo[b] : 0; /* rows below the window: nothing is read or written */
    int xq = 0;
    const int yq = (q < S * S4) ? hoc_div_small(q, S4, &xq) : 0;
    if (q < S * S4 && yq >= y_first) {
        const int yi = yq, x0 = xq << 2;
        const long pix = (long)yi * S + x0;
        const uint4 k01 = *reinterpret_cast<const uint4 *>(zbuf + (long)b * npix + pix);
        const uint4 k23 = *reinterpret_cast<const uint4 *>(zbuf + (long)b * npix + pix + 2);
        const int f4[4] = {(int)k01.x, (int)k01.z, (int)k23.x, (int)k23.z}; /* low words: the face (-1 = empty key) */
        *reinterpret_cast<int4 *>(face_index_map + (long)b * npix + pix) = make_int4(f4[0], f4[1], f4[2], f4[3]);
        const long po = ((long)b * S + (S - 1 - yi)) * S + x0; /* image layout: rows flipped */
        if (alpha != nullptr)
            *reinterpret_cast<float4 *>(alpha + po) = make_float4(f4[0] >= 0 ? 1.0f : 0.0f, f4[1] >= 0 ? 1.0f : 0.0f,
                                                                  f4[2] >= 0 ? 1.0f : 0.0f, f4[3] >= 0 ? 1.0f : 0.0f);
        if (depth != nullptr && !sparse_saved)
            *reinterpret_cast<float4 *>(depth + po) = make_float4(far_, far_, far_, far_);
        if (rgb != nullptr) {
#pragma unroll
            for (int c = 0; c < 3; c++)
                *reinterpret_cast<float4 *>(rgb + (((long)b * 3 + c) * S + (S - 1 - yi)) * S + x0) =
                    make_float4(col[c], col[c], col[c], col[c]);
        }
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (f4[j] >= 0)
                s_list[atomicAdd(&s_n, 1)] = (unsigned short)(threadIdx.x * 4 + j);
    }
    __syncthreads(); /* orders phase A's background stores before phase B's stores to the same addresses */
    const int n = s_n;
    for (int i = threadIdx.x; i < n; i += RS4_THREADS) {
        const int loc = s_list[i];
        const int qq = chunk * RS4_THREADS + (loc >> 2);
        int xr;
        const int yi = hoc_div_small(qq, S4, &xr), xi = (xr << 2) + (loc & 3);
        const long pix = (long)yi * S + xi;
        const int fidx = (int)(unsigned)(zbuf[(long)b * npix + pix] & 0xffffffffull);
        float w[3], inv[9], zp = far_, cc[3] = {col[0], col[1], col[2]};
        hoc_resolve_covered(faces, textures, b, F, S, ts, near_, far_, eps, tex_vertex, rgb != nullptr, fidx, xi, yi, w,
                            inv, &zp, cc);
        const long po = ((long)b * S + (S - 1 - yi)) * S + xi;
        if (depth != nullptr)
            depth[po] = zp;
        if (weight_map != nullptr) {
            float *wd = weight_map + ((long)b * npix + pix) * 3;
            wd[0] = w[0];
            wd[1] = w[1];
            wd[2] = w[2];
        }
        if (rgb != nullptr) {
#pragma unroll
            for (int c = 0; c < 3; c++)
                rgb[(((long)b * 3 + c) * S + (S - 1 - yi)) * S + xi] = cc[c];
        }
    }
}

extern "C" size_t hoc_raster_forward_workspace_bytes(int B, int F, int S)
{
    (void)F;
    if (B <= 0 || S <= 0)
        return 0;
    /* rounded up to 256 bytes: hoc_mesh_gather_clear fills the keys with 16-byte stores (odd B * S * S otherwise
     * left an 8-byte tail it had to reject) */
    return ((size_t)B * S * S * sizeof(unsigned long long) + 255) & ~(size_t)255;
}

extern "C" int hoc_raster_forward_ex(const float *faces, const float *textures, int B, int F, int S, int ts,
                                     float near_, float far_, float eps, const float *background_host,
                                     const float *background_dev, int layout, const int *row_lo, float *rgb,
                                     float *alpha, float *depth, int32_t *face_index_map, float *weight_map,
                                     float *face_inv_map, void *workspace, size_t workspace_bytes, void *stream);

extern "C" int hoc_raster_forward(const float *faces, const float *textures, int B, int F, int S, int ts,
                                  float near_, float far_, float eps, const float *background_host,
                                  const float *background_dev, int layout, float *rgb, float *alpha, float *depth,
                                  int32_t *face_index_map, float *weight_map, float *face_inv_map, void *workspace,
                                  size_t workspace_bytes, void *stream)
{
    return hoc_raster_forward_ex(faces, textures, B, F, S, ts, near_, far_, eps, background_host, background_dev, layout,
                                 nullptr, rgb, alpha, depth, face_index_map, weight_map, face_inv_map, workspace,
                                 workspace_bytes, stream);
}

/* row_lo: device int [B] or NULL.  Raster rows yi < row_lo[b] of sample b (rows count from the bottom) are outside the
 * sample's window: no face is drawn there and the output maps are NOT written there (their content is undefined).
 * Honoured by the 4-pixel image-layout resolve pass (what the frame-pair path uses); with any other output
 * configuration the rows are resolved as background. */
extern "C" int hoc_raster_forward_ex(const float *faces, const float *textures, int B, int F, int S, int ts,
                                     float near_, float far_, float eps, const float *background_host,
                                     const float *background_dev, int layout, const int *row_lo, float *rgb,
                                     float *alpha, float *depth, int32_t *face_index_map, float *weight_map,
                                     float *face_inv_map, void *workspace, size_t workspace_bytes, void *stream)
{
    HOC_CHECK_ARG(B >= 0 && F >= 0, "hoc_raster_forward: negative batch (%d) or face count (%d)", B, F);
    HOC_CHECK_ARG(S >= 1 && S <= 2048, "hoc_raster_forward: image_size %d outside [1, 2048]", S);
    const bool keys_cleared = (layout & HOC_LAYOUT_KEYS_CLEARED) != 0; /* the caller filled the workspace with 0xff */
    const int tex_vertex = (layout & HOC_LAYOUT_TEX_VERTEX) ? 1 : 0;   /* textures = [B,F,3,3] vertex values */
    const int sparse_saved = (layout & HOC_LAYOUT_SPARSE_SAVED) ? 1 : 0; /* depth / weight_map at covered pixels only */
    layout &= ~(HOC_LAYOUT_KEYS_CLEARED | HOC_LAYOUT_TEX_VERTEX | HOC_LAYOUT_SPARSE_SAVED);
    HOC_CHECK_ARG(!tex_vertex || ts == 2, "hoc_raster_forward: vertex textures need texture_size 2, got %d", ts);
    HOC_CHECK_ARG(layout == HOC_LAYOUT_RAW || layout == HOC_LAYOUT_IMAGE, "hoc_raster_forward: bad layout %d", layout);
    HOC_CHECK_ARG(face_index_map != nullptr, "hoc_raster_forward: face_index_map is required");
    HOC_CHECK_ARG(rgb == nullptr || ((textures != nullptr || F == 0) && ts >= 1),
                  "hoc_raster_forward: rgb requested without textures / texture_size");
    HOC_CHECK_ARG(B == 0 || F == 0 || faces != nullptr, "hoc_raster_forward: faces is NULL");
    HOC_CHECK_ARG(B <= 65535, "hoc_raster_forward: batch %d exceeds 65535", B);
    if (B == 0)
        return HOC_OK;
    const size_t need = hoc_raster_forward_workspace_bytes(B, F, S);
    if (workspace == nullptr || workspace_bytes < need) {
        hoc_set_error("hoc_raster_forward: workspace of %zu bytes needed, %zu given", need, workspace_bytes);
        return HOC_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long *zbuf = (unsigned long long *)workspace;
    cudaError_t e = keys_cleared ? cudaSuccess : cudaMemsetAsync(zbuf, 0xff, need, st);
    if (e != cudaSuccess) {
        hoc_set_error("hoc_raster_forward: memset failed: %s", cudaGetErrorString(e));
        return HOC_ERR_CUDA;
    }
    if (F > 0) {
        const int per_cta = ZB_WARPS * ZB_FACES_PER_WARP;
        dim3 grid((F + per_cta - 1) / per_cta, B);
        HOC_LAUNCH(HOC_K_RASTER_ZBUF, st,
                   (hoc_launch_pdl((hoc_raster_zbuf_kernel), grid, ZB_THREADS, S * sizeof(float), st, faces, zbuf, F, S, near_, far_,
                                                                                        row_lo)));
        HOC_CHECK_LAUNCH("hoc_raster_zbuf_kernel");
    }
    float bg[3] = {0.f, 0.f, 0.f};
    if (background_host != nullptr) {
        bg[0] = background_host[0];
        bg[1] = background_host[1];
        bg[2] = background_host[2];
    }
    const long npix = (long)S * S;
    /* (four pixels per thread with 16-byte stores and no staging was measured: 19.6 us against 15.8 -- the pass is
     * bound by the dependent chain of the covered pixels, key -> face -> texture, which a thread then runs four times) */
    if (layout == HOC_LAYOUT_IMAGE && (S % 4) == 0 && face_inv_map == nullptr && (weight_map == nullptr || sparse_saved) &&
        ((((uintptr_t)zbuf | (uintptr_t)rgb | (uintptr_t)alpha | (uintptr_t)depth | (uintptr_t)face_index_map) & 15) == 0)) {
        const long groups = npix / 4;
        dim3 grid4(B, (unsigned)((groups + RS4_THREADS - 1) / RS4_THREADS));
        HOC_LAUNCH(HOC_K_RASTER_RESOLVE, st,
                   (hoc_launch_pdl((hoc_raster_resolve4_kernel), grid4, RS4_THREADS, 0, st, faces, textures, zbuf, F, S, ts, near_, far_, eps,
                                                                           bg[0], bg[1], bg[2], background_dev, tex_vertex,
                                                                           sparse_saved, rgb, alpha, depth,
                                                                           face_index_map, weight_map, row_lo)));
        HOC_CHECK_LAUNCH("hoc_raster_resolve4_kernel");
        return HOC_OK;
    }
    dim3 grid2((unsigned)((npix + RS_THREADS - 1) / RS_THREADS), B);
    HOC_LAUNCH(HOC_K_RASTER_RESOLVE, st,
               (hoc_raster_resolve_kernel<<<grid2, RS_THREADS, 0, st>>>(faces, textures, zbuf, F, S, ts, near_, far_, eps,
                                                                        bg[0], bg[1], bg[2], background_dev, layout,
                                                                        tex_vertex, sparse_saved, rgb, alpha, depth,
                                                                        face_index_map, weight_map, face_inv_map)));
    HOC_CHECK_LAUNCH("hoc_raster_resolve_kernel");
    return HOC_OK;
}
