/* warp_math.cuh -- arithmetic of the flow-guided warp shared by warp_photo.cu and flow_path.cu (the fused
 * finalize + warp kernel).  Bit compatibility with the ATen CUDA kernels the reference runs on is documented at the top
 * of warp_photo.cu; compile with -fmad=false. */
#pragma once
#include "hoc_common.cuh"

struct HocTaps {
    float ix, iy;
    int x0, y0;          /* north-west tap */
    float nw, ne, sw, se;
    bool b_nw, b_ne, b_sw, b_se; /* tap inside the image */
};

/* 1 / max(size - 1, 1) as torch's div-by-scalar kernel computes it (IEEE float divide); evaluated once per
 * thread by the kernels below and passed down, not once per coordinate */
__device__ __forceinline__ float hoc_inv_extent(int size) { return __fdiv_rn(1.0f, (float)max(size - 1, 1)); }

__device__ __forceinline__ float hoc_norm_coord_inv(int p, float flow, float inv)
{
    const float v = __fadd_rn((float)p, flow);
    return __fadd_rn(__fmul_rn(__fmul_rn(2.0f, v), inv), -1.0f);
}

__device__ __forceinline__ float hoc_norm_coord(int p, float flow, int size)
{
    return hoc_norm_coord_inv(p, flow, hoc_inv_extent(size));
}

__device__ __forceinline__ float hoc_unnormalize(float coord, int size)
{
    /* ((coord + 1.f) * size - 1) / 2 with the multiply-subtract contracted */
    return __fmul_rn(__fmaf_rn(__fadd_rn(coord, 1.0f), (float)size, -1.0f), 0.5f);
}

__device__ __forceinline__ void hoc_bilinear_taps_inv(int x, int y, float fx, float fy, int H, int W, float inv_w,
                                                      float inv_h, HocTaps &T);

__device__ __forceinline__ void hoc_bilinear_taps(int x, int y, float fx, float fy, int H, int W, HocTaps &T)
{
    hoc_bilinear_taps_inv(x, y, fx, fy, H, W, hoc_inv_extent(W), hoc_inv_extent(H), T);
}

__device__ __forceinline__ void hoc_bilinear_taps_inv(int x, int y, float fx, float fy, int H, int W, float inv_w,
                                                      float inv_h, HocTaps &T)
{
    const float ix = hoc_unnormalize(hoc_norm_coord_inv(x, fx, inv_w), W);
    const float iy = hoc_unnormalize(hoc_norm_coord_inv(y, fy, inv_h), H);
    T.ix = ix;
    T.iy = iy;
    /* clamp before the conversion only to keep it defined; far-away taps are out of bounds anyway */
    const float fxn = floorf(fminf(fmaxf(ix, -4.0f), (float)W + 4.0f));
    const float fyn = floorf(fminf(fmaxf(iy, -4.0f), (float)H + 4.0f));
    const int x0 = (int)fxn, y0 = (int)fyn;
    T.x0 = x0;
    T.y0 = y0;
    const float x_nw = (float)x0, y_nw = (float)y0;
    const float x_se = (float)(x0 + 1), y_se = (float)(y0 + 1);
    T.nw = __fmul_rn(__fsub_rn(x_se, ix), __fsub_rn(y_se, iy));
    T.ne = __fmul_rn(__fsub_rn(ix, x_nw), __fsub_rn(y_se, iy));
    T.sw = __fmul_rn(__fsub_rn(x_se, ix), __fsub_rn(iy, y_nw));
    T.se = __fmul_rn(__fsub_rn(ix, x_nw), __fsub_rn(iy, y_nw));
    const bool xin0 = x0 >= 0 && x0 < W, xin1 = x0 + 1 >= 0 && x0 + 1 < W;
    const bool yin0 = y0 >= 0 && y0 < H, yin1 = y0 + 1 >= 0 && y0 + 1 < H;
    /* NaN coordinates: every comparison above is false in ATen as well -> no tap */
    const bool ok = (ix == ix) && (iy == iy);
    T.b_nw = ok && xin0 && yin0;
    T.b_ne = ok && xin1 && yin0;
    T.b_sw = ok && xin0 && yin1;
    T.b_se = ok && xin1 && yin1;
}

/* grid_sample of an all-ones image: sum of the in-bounds weights in tap order. */
__device__ __forceinline__ float hoc_ones_sample(const HocTaps &T)
{
    float acc = 0.0f;
    if (T.b_nw) acc = __fmaf_rn(1.0f, T.nw, acc);
    if (T.b_ne) acc = __fmaf_rn(1.0f, T.ne, acc);
    if (T.b_sw) acc = __fmaf_rn(1.0f, T.sw, acc);
    if (T.b_se) acc = __fmaf_rn(1.0f, T.se, acc);
    return acc;
}

/* grid_sample of one channel plane (zeros padding). */
__device__ __forceinline__ float hoc_plane_sample(const float *__restrict__ plane, int W, const HocTaps &T)
{
    float acc = 0.0f;
    const float *p = plane + (long)T.y0 * W + T.x0;

    if (T.b_nw) acc = __fmaf_rn(__ldg(p), T.nw, acc);
    if (T.b_ne) acc = __fmaf_rn(__ldg(p + 1), T.ne, acc);
    if (T.b_sw) acc = __fmaf_rn(__ldg(p + W), T.sw, acc);
    if (T.b_se) acc = __fmaf_rn(__ldg(p + W + 1), T.se, acc);
    return acc;
}

/* Branch-free form of hoc_plane_sample for the streaming kernels: out-of-bounds taps are read from a clamped
 * (valid) address and enter the same FMA chain with weight 0 -- fma(v, 0, acc) == acc for every finite v -- so
 * that all loads of all planes can be issued before the first one is consumed. */
struct HocTapsFlat {
    int o[4];   /* offsets of nw, ne, sw, se inside a plane (clamped into the image) */
    float w[4]; /* their weights, 0 for taps outside the image */
};

__device__ __forceinline__ void hoc_flatten_taps(const HocTaps &T, int H, int W, HocTapsFlat &F)
{
    const int x0 = min(max(T.x0, 0), W - 1), x1 = min(max(T.x0 + 1, 0), W - 1);
    const int y0 = min(max(T.y0, 0), H - 1), y1 = min(max(T.y0 + 1, 0), H - 1);
    F.o[0] = y0 * W + x0;
    F.o[1] = y0 * W + x1;
    F.o[2] = y1 * W + x0;
    F.o[3] = y1 * W + x1;
    F.w[0] = T.b_nw ? T.nw : 0.0f;
    F.w[1] = T.b_ne ? T.ne : 0.0f;
    F.w[2] = T.b_sw ? T.sw : 0.0f;
    F.w[3] = T.b_se ? T.se : 0.0f;
}

__device__ __forceinline__ float hoc_flat_combine(const float *v, const HocTapsFlat &F)
{
    return __fmaf_rn(v[3], F.w[3], __fmaf_rn(v[2], F.w[2], __fmaf_rn(v[1], F.w[1], __fmaf_rn(v[0], F.w[0], 0.0f))));
}

/* mask[mask < thresh] = 0; mask[mask > 0] = 1 */
__device__ __forceinline__ float hoc_threshold_mask(float m, float thresh)
{
    if (m < thresh)
        m = 0.0f;
    if (m > 0.0f)
        m = 1.0f;
    return m;
}

#define WP_THREADS 256
#define WP_MAXC 4
/* The per-sample sum of |diff| is accumulated over the CTAs with double atomics.  Every CTA's partial sum is first
 * rounded to an integer multiple of 2^-28: integer-valued doubles add exactly (below 2^53), so the total does not
 * depend on the order in which the CTAs arrive -- the loss is reproducible bit for bit.  (A partial sum >= 2^-4 has no
 * bits below 2^-28: nothing is lost; smaller ones are rounded by < 2e-9.) */
#define WP_SUM_SCALE 268435456.0
#define WP_SUM_INV (1.0 / 268435456.0)

/* one pixel of one direction: sample position -> masks -> |warp - target|; v / d / wm get the three channel values
 * when VIS.  Arithmetic identical to hoc_warp_photo_forward_kernel<3, 3> (same helpers, same order). */
template <bool VIS>
__device__ __forceinline__ bool hoc_pair_pixel(const float *__restrict__ sb, const float *__restrict__ jb,
                                               const float *tv, float jc, int x, int y, float fx, float fy, int H,
                                               int W, int npix, float inv_w, float inv_h, float thresh, float *v,
                                               float *d, float *wm, float *sum_d)
{
    HocTaps T;
    hoc_bilinear_taps_inv(x, y, fx, fy, H, W, inv_w, inv_h, T);
    const float m = hoc_threshold_mask(hoc_ones_sample(T), thresh);
    HocTapsFlat F;
    hoc_flatten_taps(T, H, W, F);
    float sv[3][4], jv[3][4];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int k = 0; k < 4; k++)
            sv[c][k] = __ldg(sb + (size_t)c * npix + F.o[k]);
    float wm0 = m;
    if (jb != nullptr) {
#pragma unroll
        for (int c = 0; c < (VIS ? 3 : 1); c++)
#pragma unroll
            for (int k = 0; k < 4; k++)
                jv[c][k] = __ldg(jb + (size_t)c * npix + F.o[k]);
#pragma unroll
        for (int c = 0; c < (VIS ? 3 : 1); c++) {
            const float wj = __fmul_rn(hoc_flat_combine(jv[c], F), m);
            const float w = __fmul_rn(m, (wj == 1.0f) ? 1.0f : 0.0f);
            if (VIS)
                wm[c] = w;
            if (c == 0)
                wm0 = w;
        }
    } else if (VIS) {
        wm[0] = wm[1] = wm[2] = m;
    }
    const bool valid = (jb != nullptr) ? ((wm0 != 0.0f) && !(fx == 0.0f) && (jc == 1.0f)) : ((m != 0.0f) && !(fx == 0.0f));
    float acc = 0.0f;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float val = __fmul_rn(hoc_flat_combine(sv[c], F), m);
        const float dd = fabsf(__fsub_rn(val, tv[c]));
        if (VIS) {
            v[c] = val;
            d[c] = dd;
        }
        acc += dd; /* channel order r, g, b like the single-direction kernel */
    }
    *sum_d = acc;
    return valid;
}


/* ---- backward of one direction of pair_consist -------------------------------------------------------------------- */
struct HocPairBwdDir {
    const float *src, *target, *flow, *mult;
    const uint8_t *valid_mask;
    const double *sums;
    float *grad_rgb;  /* [B,3,S,S] or NULL (direction skipped) */
    float *grad_flow; /* [B,H,W,2] or NULL: the flow gradient itself, for callers that want it */
    int active;       /* 0: this direction carries no loss (use_backward = False): zeros */
};


/* d loss[b] / d flow at one VALID pixel (x, y) of direction D (valid => the in-bounds mask of the sample is 1):
 * loss[b] = sum_valid |warp - target| / max(count, 1); `scale` = d L / d loss[b] / max(count, 1).  The thresholded
 * masks carry no gradient, so only the bilinear taps of the source depend on the flow (imgflowarp.py:52-53). */
__device__ __forceinline__ void hoc_pair_bwd_pixel(const HocPairBwdDir &D, int b, int x, int y, int H, int W, float inv_w,
                                                   float inv_h, float scale, float *gfx, float *gfy)
{
    const long npix = (long)H * W;
    const long pix = (long)y * W + x;
    const float2 fl = *reinterpret_cast<const float2 *>(D.flow + ((long)b * npix + pix) * 2);
    HocTaps T;
    hoc_bilinear_taps_inv(x, y, fl.x, fl.y, H, W, inv_w, inv_h, T);
    const float x_nw = (float)T.x0, y_nw = (float)T.y0, x_se = (float)(T.x0 + 1), y_se = (float)(T.y0 + 1);
    float gix = 0.0f, giy = 0.0f;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float *plane = D.src + ((long)b * 3 + c) * npix;
        const float v = hoc_plane_sample(plane, W, T);
        const float d = v - D.target[((long)b * 3 + c) * npix + pix];
        const float sgn = (d > 0.0f) ? 1.0f : ((d < 0.0f) ? -1.0f : 0.0f);
        const float go = scale * sgn;
        const float *p = plane + (long)T.y0 * W + T.x0;
        if (T.b_nw) {
            const float v0 = __ldg(p);
            gix -= v0 * (y_se - T.iy) * go;
            giy -= v0 * (x_se - T.ix) * go;
        }
        if (T.b_ne) {
            const float v1 = __ldg(p + 1);
            gix += v1 * (y_se - T.iy) * go;
            giy -= v1 * (T.ix - x_nw) * go;
        }
        if (T.b_sw) {
            const float v2 = __ldg(p + W);
            gix -= v2 * (T.iy - y_nw) * go;
            giy += v2 * (x_se - T.ix) * go;
        }
        if (T.b_se) {
            const float v3 = __ldg(p + W + 1);
            gix += v3 * (T.iy - y_nw) * go;
            giy += v3 * (T.ix - x_nw) * go;
        }
    }
    /* d ix / d x_norm = W / 2 ; d x_norm / d flow = 2 / (W - 1) */
    *gfx = (0.5f * (float)W) * gix * 2.0f / (float)max(W - 1, 1);
    *gfy = (0.5f * (float)H) * giy * 2.0f / (float)max(H - 1, 1);
}
