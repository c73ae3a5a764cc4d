/*
 * warp_photo.cu -- flow-guided image warp, masked photometric L1 and the forward-backward
 * occlusion check for sm_100a.
 *
 * Replaces, for the reference's meshreg/warping/imgflowarp.py:
 *   warp()                 :31-55   (meshgrid + flow -> normalise -> grid_sample x2 -> threshold)
 *   pair_consist()         :58-115  (4 warps, valid-mask algebra, 2x criterion)
 *   get_occlusion_mask()   :118-146 + occlusion_mask_from_warped_grid() :149-172
 * and meshreg/optim/pyramidloss.py:56-58 + lossutils.py:1-8 (|a-b|, masked mean per sample).
 *
 * The reference runs ~60 ATen launches per frame pair here (CPU-built meshgrids, eight
 * grid_sample calls, ~40 element-wise kernels).  One direction of pair_consist is ONE kernel:
 * flow -> sampling position -> 4 taps of the image and of the jitter mask -> in-bounds mask ->
 * valid mask -> |warp - target| -> per-sample (sum, count), block-reduced with warp shuffles.
 *
 * Bit compatibility.  The reference's masks contain two exact floating-point tests
 * (`mask >= 0.99999` and `warped_jitter == 1`) on sums of four bilinear weights, so the kernel
 * reproduces the arithmetic of the ATen CUDA kernels the reference runs on, operation by
 * operation (this file is compiled with -fmad=false; the FMAs ATen's nvcc build contracts are
 * written explicitly as __fmaf_rn):
 *   x_norm  = ((x + flow) * 2) * (1 / (W-1)) - 1          torch mul / div-by-scalar / sub kernels
 *   ix      = fma(x_norm + 1, W, -1) * 0.5                grid_sampler_unnormalize, align_corners=False (F6 quirk)
 *   weights = (ix_se - ix) * (iy_se - iy) ...             in tap order nw, ne, sw, se
 *   value   = fma(v_se, se, fma(v_sw, sw, fma(v_ne, ne, v_nw * nw)))  over the in-bounds taps
 */
#include "hoc_common.cuh"

#include "warp_math.cuh"

template <int CT, int CJT> /* compile-time channel counts (0 = use the run-time values, any count) */
__global__ void __launch_bounds__(WP_THREADS)
hoc_warp_photo_forward_kernel(const float *__restrict__ src, const float *__restrict__ target,
                              const float *__restrict__ flow, const float *__restrict__ jitter, int C_rt, int Cj_rt,
                              int H, int W, float inv_w, float inv_h, float thresh, float *__restrict__ warped,
                              float *__restrict__ warp_mask, uint8_t *__restrict__ valid_mask,
                              uint8_t *__restrict__ flow_mask, float *__restrict__ diff, double *__restrict__ sums)
{
    const int C = CT > 0 ? CT : C_rt;
    const int Cj = CJT > 0 ? CJT : Cj_rt;
    __shared__ float s_sum[WP_THREADS / 32];
    __shared__ float s_cnt[WP_THREADS / 32];
    const int b = blockIdx.y;
    const int npix = H * W;
    const int pix = blockIdx.x * WP_THREADS + threadIdx.x;
    float my_sum = 0.0f, my_cnt = 0.0f;
    if (pix < npix) {
        const int y = pix / W;
        const int x = pix - y * W;
        const float2 fl = *reinterpret_cast<const float2 *>(flow + ((size_t)b * npix + pix) * 2);
        HocTaps T;
        hoc_bilinear_taps_inv(x, y, fl.x, fl.y, H, W, inv_w, inv_h, T);
        const float m = hoc_threshold_mask(hoc_ones_sample(T), thresh);
        const float *sb = src + (size_t)b * C * npix;
        const float *tb = target + (size_t)b * C * npix + pix;
        const float *jb = (jitter != nullptr) ? jitter + (size_t)b * Cj * npix : nullptr;
        bool valid;
        float wm[WP_MAXC];
        if (CT > 0) {
            /* fixed channel counts: every load of every plane is issued before the first use */
            HocTapsFlat F;
            hoc_flatten_taps(T, H, W, F);
            float sv[CT > 0 ? CT : 1][4], jv[CJT > 0 ? CJT : 1][4], tv[CT > 0 ? CT : 1];
            float jc = 1.0f;
#pragma unroll
            for (int c = 0; c < CT; c++) {
#pragma unroll
                for (int k = 0; k < 4; k++)
                    sv[c][k] = __ldg(sb + (size_t)c * npix + F.o[k]);
                tv[c] = __ldg(tb + (size_t)c * npix);
            }
            if (jb != nullptr) {
#pragma unroll
                for (int c = 0; c < CJT; c++)
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        jv[c][k] = __ldg(jb + (size_t)c * npix + F.o[k]);
                jc = __ldg(jb + pix);
            }
#pragma unroll
            for (int c = 0; c < WP_MAXC; c++)
                wm[c] = m;
            if (jb != nullptr) {
#pragma unroll
                for (int c = 0; c < CJT; c++) {
                    const float wj = __fmul_rn(hoc_flat_combine(jv[c], F), m);
                    wm[c] = __fmul_rn(m, (wj == 1.0f) ? 1.0f : 0.0f);
                }
                valid = (wm[0] != 0.0f) && !(fl.x == 0.0f) && (jc == 1.0f);
            } else {
                valid = (m != 0.0f) && !(fl.x == 0.0f);
            }
#pragma unroll
            for (int c = 0; c < CT; c++) {
                const size_t o = ((size_t)b * C + c) * npix + pix;
                const float v = __fmul_rn(hoc_flat_combine(sv[c], F), m);
                const float d = fabsf(__fsub_rn(v, tv[c]));
                if (warped != nullptr)
                    warped[o] = v;
                if (diff != nullptr)
                    diff[o] = d;
                if (warp_mask != nullptr)
                    warp_mask[o] = wm[c < WP_MAXC ? c : 0];
                if (valid) {
                    my_sum += d;
                    my_cnt += 1.0f;
                }
            }
        } else {
#pragma unroll
            for (int c = 0; c < WP_MAXC; c++)
                wm[c] = m;
            if (jb != nullptr) {
                for (int c = 0; c < Cj && c < WP_MAXC; c++) {
                    const float wj = __fmul_rn(hoc_plane_sample(jb + (size_t)c * npix, W, T), m);
                    wm[c] = __fmul_rn(m, (wj == 1.0f) ? 1.0f : 0.0f);
                }
                if (Cj == 1) {
#pragma unroll
                    for (int c = 1; c < WP_MAXC; c++)
                        wm[c] = wm[0];
                }
                valid = (wm[0] != 0.0f) && !(fl.x == 0.0f) && (__ldg(jb + pix) == 1.0f);
            } else {
                valid = (m != 0.0f) && !(fl.x == 0.0f);
            }
            for (int c = 0; c < C; c++) {
                const size_t o = ((size_t)b * C + c) * npix + pix;
                const float v = __fmul_rn(hoc_plane_sample(sb + (size_t)c * npix, W, T), m);
                const float d = fabsf(__fsub_rn(v, tb[(size_t)c * npix]));
                if (warped != nullptr)
                    warped[o] = v;
                if (diff != nullptr)
                    diff[o] = d;
                if (warp_mask != nullptr)
                    warp_mask[o] = wm[c < WP_MAXC ? c : 0];
                if (valid) {
                    my_sum += d;
                    my_cnt += 1.0f;
                }
            }
        }
        if (valid_mask != nullptr)
            valid_mask[(size_t)b * npix + pix] = valid ? 1 : 0;
        if (flow_mask != nullptr) { /* ~(flow == 0), both components (imgflowarp.py:93,99) */
            uchar2 fm;
            fm.x = !(fl.x == 0.0f) ? 1 : 0;
            fm.y = !(fl.y == 0.0f) ? 1 : 0;
            *reinterpret_cast<uchar2 *>(flow_mask + ((size_t)b * npix + pix) * 2) = fm;
        }
    }
    my_sum = hoc_warp_sum(my_sum);
    my_cnt = hoc_warp_sum(my_cnt);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        s_sum[warp] = my_sum;
        s_cnt[warp] = my_cnt;
    }
    __syncthreads();
    if (warp == 0) {
        float a = (lane < WP_THREADS / 32) ? s_sum[lane] : 0.0f;
        float n = (lane < WP_THREADS / 32) ? s_cnt[lane] : 0.0f;
        a = hoc_warp_sum(a);
        n = hoc_warp_sum(n);
        if (lane == 0 && n > 0.0f) {
            atomicAdd(&sums[2 * b + 0], rint((double)a * WP_SUM_SCALE));
            atomicAdd(&sums[2 * b + 1], (double)n);
        }
    }
}

/* batch_masked_mean_loss (lossutils.py:1-8) from the per-sample (sum, count): loss = sum / max(count, 1) */
__global__ void hoc_masked_mean_kernel(const double *__restrict__ sums, int B, float *__restrict__ loss)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B)
        loss[b] = (float)(sums[2 * b] * WP_SUM_INV / fmax(sums[2 * b + 1], 1.0));
}

/* pair_consist's loss of one frame pair (imgflowarp.py:108-114): loss_bwd + loss_fwd (that order) or loss_fwd. */
__global__ void hoc_pair_loss_kernel(const double *__restrict__ sums_fwd, const double *__restrict__ sums_bwd, int B,
                                     float *__restrict__ loss)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) {
        const float lf = (float)(sums_fwd[2 * b] * WP_SUM_INV / fmax(sums_fwd[2 * b + 1], 1.0));
        loss[b] = (sums_bwd != nullptr)
                      ? (float)(sums_bwd[2 * b] * WP_SUM_INV / fmax(sums_bwd[2 * b + 1], 1.0)) + lf
                      : lf;
    }
}

/* The same plus the batch mean (warpbranch.py:88: the mean over the single pair's per-sample losses), one warp:
 * no separate reduction kernel, and its adjoint (d mean / d loss[b] = 1 / B) is folded into the backward kernel. */
__global__ void hoc_pair_loss_mean_kernel(const double *__restrict__ sums_fwd, const double *__restrict__ sums_bwd, int B,
                                          float *__restrict__ loss, float *__restrict__ mean, uint4 *__restrict__ zero,
                                          long n_zero)
{
    hoc_pdl_sync(); /* programmatic dependent launch: see hoc_common.cuh */
    if (blockIdx.x > 0) { /* the other CTAs zero-fill a buffer of the step's BACKWARD (its rasterizer's counters) */
        hoc_fill16(zero, n_zero, (long)(blockIdx.x - 1) * blockDim.x + threadIdx.x, (long)(gridDim.x - 1) * blockDim.x, 0u);
        return;
    }
    if (threadIdx.x >= 32)
        return;
    float acc = 0.0f;
    for (int b = threadIdx.x; b < B; b += 32) {
        const float lf = (float)(sums_fwd[2 * b] * WP_SUM_INV / fmax(sums_fwd[2 * b + 1], 1.0));
        const float l = (sums_bwd != nullptr)
                            ? (float)(sums_bwd[2 * b] * WP_SUM_INV / fmax(sums_bwd[2 * b + 1], 1.0)) + lf
                            : lf;
        loss[b] = l;
        acc += l;
    }
    acc = hoc_warp_sum(acc);
    if (threadIdx.x == 0)
        *mean = acc / (float)B;
}

/* d loss[b] / d flow.  loss[b] = sum_valid |warp - target| / max(count, 1); the thresholded masks
 * carry no gradient, so only the bilinear taps of `src` depend on the flow. */
__global__ void __launch_bounds__(WP_THREADS)
hoc_warp_photo_backward_kernel(const float *__restrict__ src, const float *__restrict__ target,
                               const float *__restrict__ flow, const uint8_t *__restrict__ valid_mask,
                               const double *__restrict__ sums, const float *__restrict__ grad_loss, int C, int H,
                               int W, float inv_w, float inv_h, float thresh, float *__restrict__ grad_flow)
{
    const int b = blockIdx.y;
    const long npix = (long)H * W;
    const long pix = (long)blockIdx.x * WP_THREADS + threadIdx.x;
    if (pix >= npix)
        return;
    float2 g = make_float2(0.0f, 0.0f);
    if (valid_mask[(long)b * npix + pix]) {
        const int y = (int)((unsigned)pix / (unsigned)W);
        const int x = (int)pix - y * W;
        const float2 fl = *reinterpret_cast<const float2 *>(flow + ((long)b * npix + pix) * 2);
        HocTaps T;
        hoc_bilinear_taps_inv(x, y, fl.x, fl.y, H, W, inv_w, inv_h, T);
        const float cnt = (float)sums[2 * b + 1];
        const float scale = grad_loss[b] / fmaxf(cnt, 1.0f);
        const float x_nw = (float)T.x0, y_nw = (float)T.y0, x_se = (float)(T.x0 + 1), y_se = (float)(T.y0 + 1);
        float gix = 0.0f, giy = 0.0f;
        for (int c = 0; c < C; c++) {
            const float *plane = src + ((long)b * C + c) * npix;
            const float v = hoc_plane_sample(plane, W, T); /* valid => in-bounds mask is 1 */
            const float d = v - target[((long)b * C + c) * npix + pix];
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
        g.x = (0.5f * (float)W) * gix * 2.0f / (float)max(W - 1, 1);
        g.y = (0.5f * (float)H) * giy * 2.0f / (float)max(H - 1, 1);
    }
    *reinterpret_cast<float2 *>(grad_flow + ((long)b * npix + pix) * 2) = g;
}

/* ------------------------------------------------------------------------------------------ */
/* Both directions of pair_consist in ONE launch, four pixels per thread (the frame-pair fast path).
 *
 * blockIdx.z = direction k (imgflowarp.py:80-107): k = 0 warps image_ref with flow21 against image (jitter mask of
 * frame 2), k = 1 warps image with flow12 against image_ref (jitter mask of the reference frame).  A thread owns four
 * consecutive pixels of a row: flows, targets, the centre jitter value and every dense output move as 16-byte
 * accesses; only the bilinear taps (data-dependent addresses) stay scalar.
 *
 * VIS = false (training): the visualisation returns of pair_consist (`warps`, `diffs`, `warp_mask`) are not
 * produced.  What is left per pixel is the loss term and the valid mask, and both vanish where the rendered flow is
 * zero (valid = ... & (flow_x != 0), imgflowarp.py:93-101) -- 93-95 % of a frame: such a pixel costs its 8-byte flow
 * and 3 bytes of masks, nothing else.  VIS = true produces everything the reference returns. */
struct HocPairDir {
    const float *src, *target, *flow, *jitter;
    float *warped, *warp_mask, *diff;
    uint8_t *valid_mask, *flow_mask;
    double *sums;
};

template <bool VIS>
__global__ void __launch_bounds__(WP_THREADS)
hoc_warp_photo_pair_forward_kernel(HocPairDir D0, HocPairDir D1, int H, int W, float inv_w, float inv_h, float thresh)
{
    __shared__ float s_sum[WP_THREADS / 32];
    __shared__ float s_cnt[WP_THREADS / 32];
    const HocPairDir &D = blockIdx.z ? D1 : D0;
    const int b = blockIdx.y;
    const int npix = H * W, W4 = W >> 2;
    const int q = blockIdx.x * WP_THREADS + threadIdx.x; /* group of four pixels */
    float my_sum = 0.0f, my_cnt = 0.0f;
    if (q < H * W4) {
        const int y = q / W4, x0 = (q - y * W4) << 2;
        const size_t pix = (size_t)y * W + x0;
        const float4 f01 = *reinterpret_cast<const float4 *>(D.flow + ((size_t)b * npix + pix) * 2);
        const float4 f23 = *reinterpret_cast<const float4 *>(D.flow + ((size_t)b * npix + pix) * 2 + 4);
        const float fx[4] = {f01.x, f01.z, f23.x, f23.z}, fy[4] = {f01.y, f01.w, f23.y, f23.w};
        unsigned vbits = 0;
        const bool any = VIS || !(fx[0] == 0.0f) || !(fx[1] == 0.0f) || !(fx[2] == 0.0f) || !(fx[3] == 0.0f);
        if (any) {
            const float *sb = D.src + (size_t)b * 3 * npix;
            const float *jb = (D.jitter != nullptr) ? D.jitter + (size_t)b * 3 * npix : nullptr;
            const float *tb = D.target + (size_t)b * 3 * npix + pix;
            const float4 t0 = *reinterpret_cast<const float4 *>(tb), t1 = *reinterpret_cast<const float4 *>(tb + npix),
                         t2 = *reinterpret_cast<const float4 *>(tb + 2 * (size_t)npix);
            const float4 jc4 = (jb != nullptr) ? *reinterpret_cast<const float4 *>(jb + pix)
                                               : make_float4(1.0f, 1.0f, 1.0f, 1.0f);
            const float tr[4] = {t0.x, t0.y, t0.z, t0.w}, tg[4] = {t1.x, t1.y, t1.z, t1.w},
                        tbl[4] = {t2.x, t2.y, t2.z, t2.w}, jc[4] = {jc4.x, jc4.y, jc4.z, jc4.w};
            float ov[3][4], od[3][4], om[3][4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (VIS || !(fx[j] == 0.0f)) {
                    const float tv[3] = {tr[j], tg[j], tbl[j]};
                    float v[3], d[3], wm[3], sd;
                    const bool valid = hoc_pair_pixel<VIS>(sb, jb, tv, jc[j], x0 + j, y, fx[j], fy[j], H, W, npix, inv_w,
                                                           inv_h, thresh, v, d, wm, &sd);
                    if (VIS) {
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            ov[c][j] = v[c];
                            od[c][j] = d[c];
                            om[c][j] = wm[c];
                        }
                    }
                    if (valid) {
                        vbits |= 1u << (8 * j);
                        my_sum += sd;
                        my_cnt += 3.0f;
                    }
                }
            }
            if (VIS) {
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const size_t o = ((size_t)b * 3 + c) * npix + pix;
                    if (D.warped != nullptr)
                        *reinterpret_cast<float4 *>(D.warped + o) = make_float4(ov[c][0], ov[c][1], ov[c][2], ov[c][3]);
                    if (D.diff != nullptr)
                        *reinterpret_cast<float4 *>(D.diff + o) = make_float4(od[c][0], od[c][1], od[c][2], od[c][3]);
                    if (D.warp_mask != nullptr)
                        *reinterpret_cast<float4 *>(D.warp_mask + o) = make_float4(om[c][0], om[c][1], om[c][2], om[c][3]);
                }
            }
        }
        if (D.valid_mask != nullptr)
            *reinterpret_cast<unsigned *>(D.valid_mask + (size_t)b * npix + pix) = vbits;
        if (D.flow_mask != nullptr) { /* ~(flow == 0), both components (imgflowarp.py:93,99) */
            uint2 fm;
            fm.x = (!(fx[0] == 0.0f) ? 1u : 0u) | (!(fy[0] == 0.0f) ? 0x100u : 0u) | (!(fx[1] == 0.0f) ? 0x10000u : 0u) |
                   (!(fy[1] == 0.0f) ? 0x1000000u : 0u);
            fm.y = (!(fx[2] == 0.0f) ? 1u : 0u) | (!(fy[2] == 0.0f) ? 0x100u : 0u) | (!(fx[3] == 0.0f) ? 0x10000u : 0u) |
                   (!(fy[3] == 0.0f) ? 0x1000000u : 0u);
            *reinterpret_cast<uint2 *>(D.flow_mask + ((size_t)b * npix + pix) * 2) = fm;
        }
    }
    my_sum = hoc_warp_sum(my_sum);
    my_cnt = hoc_warp_sum(my_cnt);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        s_sum[warp] = my_sum;
        s_cnt[warp] = my_cnt;
    }
    __syncthreads();
    if (warp == 0) {
        float a = (lane < WP_THREADS / 32) ? s_sum[lane] : 0.0f;
        float n = (lane < WP_THREADS / 32) ? s_cnt[lane] : 0.0f;
        a = hoc_warp_sum(a);
        n = hoc_warp_sum(n);
        if (lane == 0 && n > 0.0f) {
            atomicAdd(&D.sums[2 * b + 0], rint((double)a * WP_SUM_SCALE));
            atomicAdd(&D.sums[2 * b + 1], (double)n);
        }
    }
}

/* Backward of both directions, fused with hoc_flow_finalize_backward: d loss / d flow (hoc_warp_photo_backward_kernel's
 * arithmetic) times mult = d flow / d rgb, written straight into the incoming gradient of the two renders' rgb maps
 * ([B,3,S,S], image layout; zero outside the H x W crop and in the third channel).  blockIdx.z = direction k: k = 0
 * differentiates flow21 (render 2), k = 1 flow12 (render 1).  Four pixels of a raster row per thread. */
#ifndef WPB_THREADS
#define WPB_THREADS 128 /* (128 vs 256: 11.2 / 11.9 us) */
#endif
__global__ void __launch_bounds__(WPB_THREADS)
hoc_warp_photo_pair_backward_kernel(HocPairBwdDir D0, HocPairBwdDir D1, const float *__restrict__ grad_loss,
                                    const float *__restrict__ grad_mean, int B, int S, int H, int W, float inv_w,
                                    float inv_h, uint4 *__restrict__ zero, long n_zero,
                                    const int *__restrict__ row_lo1, const int *__restrict__ row_lo2)
{
    if (n_zero > 0) { /* zero-fill for the kernels that follow (counters of the rasterizer backward), spread over the grid */
        const long nthreads = (long)gridDim.x * gridDim.y * gridDim.z * WPB_THREADS;
        hoc_fill16(zero, n_zero,
                   (((long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * WPB_THREADS + threadIdx.x,
                   nthreads, 0u);
    }
    const HocPairBwdDir &D = blockIdx.z ? D1 : D0;
    if (D.grad_rgb == nullptr && D.grad_flow == nullptr)
        return;
    /* Two phases per CTA (4 raster pixels per thread).  A: 16-byte zero stores of the gradient planes, the valid pixels noted
     * in a shared list.  B: the listed pixels (a few per cent) one per thread: 12 taps + the gradient of the bilinear
     * weights, scalar stores over the zeros. */
    __shared__ unsigned short s_list[WPB_THREADS * 4];
    __shared__ int s_n;
    const int b = blockIdx.y;
    const int S4 = S >> 2;
    const int q = blockIdx.x * WPB_THREADS + threadIdx.x;
    const long npix = (long)H * W;
    if (threadIdx.x == 0)
        s_n = 0;
    __syncthreads();
    /* rows of the render below its raster window (raster rows count from the bottom, this layout from the top) are
     * never read by the rasterizer backward: not written */
    const int *row_lo = blockIdx.z ? row_lo1 : row_lo2; /* direction 0 -> render 2, direction 1 -> render 1 */
    const int y_last = (row_lo != nullptr) ? S - 1 - row_lo[b] : S - 1;
    if (q < S * S4 && q / S4 <= y_last) {
        const int y = q / S4, x0 = (q - y * S4) << 2;
        const bool inside = y < H && x0 < W; /* W % 4 == 0: a group is inside or outside as a whole */
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (D.grad_rgb != nullptr) {
            float *dst = D.grad_rgb + (long)b * 3 * S * S + (long)y * S + x0;
            *reinterpret_cast<float4 *>(dst) = z4;
            *reinterpret_cast<float4 *>(dst + (long)S * S) = z4;
            *reinterpret_cast<float4 *>(dst + 2l * S * S) = z4;
        }
        if (inside) {
            const long pix = (long)y * W + x0;
            if (D.grad_flow != nullptr) {
                float *dst = D.grad_flow + ((long)b * npix + pix) * 2;
                *reinterpret_cast<float4 *>(dst) = z4;
                *reinterpret_cast<float4 *>(dst + 4) = z4;
            }
            if (D.active) {
                const unsigned vb = *reinterpret_cast<const unsigned *>(D.valid_mask + (long)b * npix + pix);
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if ((vb >> (8 * j)) & 0xffu)
                        s_list[atomicAdd(&s_n, 1)] = (unsigned short)(threadIdx.x * 4 + j);
            }
        }
    }
    __syncthreads(); /* orders phase A's zero stores before phase B's stores to the same addresses */
    const int n = s_n;
    if (n == 0)
        return;
    const float cnt = (float)D.sums[2 * b + 1];
    /* d L / d loss[b]: given directly, and / or through the batch mean (d mean / d loss[b] = 1 / B) */
    const float gl = ((grad_loss != nullptr) ? grad_loss[b] : 0.0f) +
                     ((grad_mean != nullptr) ? __fdiv_rn(grad_mean[0], (float)B) : 0.0f);
    const float scale = gl / fmaxf(cnt, 1.0f);
    for (int i = threadIdx.x; i < n; i += WPB_THREADS) {
        const int loc = s_list[i];
        const int qq = blockIdx.x * WPB_THREADS + (loc >> 2);
        const int y = qq / S4, x = ((qq - y * S4) << 2) + (loc & 3);
        const long pix = (long)y * W + x;
        float gfx, gfy;
        hoc_pair_bwd_pixel(D, b, x, y, H, W, inv_w, inv_h, scale, &gfx, &gfy);
        if (D.grad_rgb != nullptr) {
            const float m = D.mult[(long)b * npix + pix]; /* hoc_flow_finalize_backward_kernel */
            float *dst = D.grad_rgb + (long)b * 3 * S * S + (long)y * S + x;
            dst[0] = gfx * m;
            dst[(long)S * S] = gfy * m;
        }
        if (D.grad_flow != nullptr)
            *reinterpret_cast<float2 *>(D.grad_flow + ((long)b * npix + pix) * 2) = make_float2(gfx, gfy);
    }
}

/* Plain warp(): out = grid_sample(x, grid(flow)) * mask.  flow is NCHW [B,2,H,W] like the
 * reference's argument.  mode 0 = bilinear, 1 = nearest. */
__global__ void __launch_bounds__(WP_THREADS)
hoc_warp_kernel(const float *__restrict__ x, const float *__restrict__ flow, int C, int H, int W, float thresh,
                int mode, float *__restrict__ out, float *__restrict__ mask)
{
    const int b = blockIdx.y;
    const long npix = (long)H * W;
    const long pix = (long)blockIdx.x * WP_THREADS + threadIdx.x;
    if (pix >= npix)
        return;
    const int py = (int)((unsigned)pix / (unsigned)W);
    const int px = (int)pix - py * W;
    const float fx = flow[((long)b * 2 + 0) * npix + pix];
    const float fy = flow[((long)b * 2 + 1) * npix + pix];
    if (mode == 0) {
        HocTaps T;
        hoc_bilinear_taps(px, py, fx, fy, H, W, T);
        const float m = hoc_threshold_mask(hoc_ones_sample(T), thresh);
        for (int c = 0; c < C; c++) {
            const long o = ((long)b * C + c) * npix + pix;
            out[o] = __fmul_rn(hoc_plane_sample(x + ((long)b * C + c) * npix, W, T), m);
            if (mask != nullptr)
                mask[o] = m;
        }
    } else {
        const float ix = hoc_unnormalize(hoc_norm_coord(px, fx, W), W);
        const float iy = hoc_unnormalize(hoc_norm_coord(py, fy, H), H);
        const float rx = nearbyintf(fminf(fmaxf(ix, -4.0f), (float)W + 4.0f));
        const float ry = nearbyintf(fminf(fmaxf(iy, -4.0f), (float)H + 4.0f));
        const int sx = (int)rx, sy = (int)ry;
        const bool inb = (ix == ix) && (iy == iy) && sx >= 0 && sx < W && sy >= 0 && sy < H;
        const float m = hoc_threshold_mask(inb ? 1.0f : 0.0f, thresh);
        for (int c = 0; c < C; c++) {
            const long o = ((long)b * C + c) * npix + pix;
            const float v = inb ? __ldg(x + ((long)b * C + c) * npix + (long)sy * W + sx) : 0.0f;
            out[o] = __fmul_rn(v, m);
            if (mask != nullptr)
                mask[o] = m;
        }
    }
}


/* Gradient of warp() (bilinear) w.r.t. its NCHW flow: grad_out is the gradient of `out * mask`. */
__global__ void __launch_bounds__(WP_THREADS)
hoc_warp_backward_kernel(const float *__restrict__ x, const float *__restrict__ flow,
                         const float *__restrict__ grad_out, int C, int H, int W, float thresh,
                         float *__restrict__ grad_flow)
{
    const int b = blockIdx.y;
    const long npix = (long)H * W;
    const long pix = (long)blockIdx.x * WP_THREADS + threadIdx.x;
    if (pix >= npix)
        return;
    const int py = (int)((unsigned)pix / (unsigned)W);
    const int px = (int)pix - py * W;
    HocTaps T;
    hoc_bilinear_taps(px, py, flow[((long)b * 2 + 0) * npix + pix], flow[((long)b * 2 + 1) * npix + pix], H, W, T);
    const float m = hoc_threshold_mask(hoc_ones_sample(T), thresh);
    float gix = 0.0f, giy = 0.0f;
    if (m != 0.0f) {
        const float x_nw = (float)T.x0, y_nw = (float)T.y0, x_se = (float)(T.x0 + 1), y_se = (float)(T.y0 + 1);
        for (int c = 0; c < C; c++) {
            const float go = grad_out[((long)b * C + c) * npix + pix] * m;
            const float *p = x + ((long)b * C + c) * npix + (long)T.y0 * W + T.x0;
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
    }
    grad_flow[((long)b * 2 + 0) * npix + pix] = (0.5f * (float)W) * gix * 2.0f / (float)max(W - 1, 1);
    grad_flow[((long)b * 2 + 1) * npix + pix] = (0.5f * (float)H) * giy * 2.0f / (float)max(H - 1, 1);
}

/* Nearest-mode source pixel of warp(., flow) at (px, py); false when it falls outside. */
__device__ __forceinline__ bool hoc_nearest_src(int px, int py, float fx, float fy, int H, int W, int *sx, int *sy)
{
    const float ix = hoc_unnormalize(hoc_norm_coord(px, fx, W), W);
    const float iy = hoc_unnormalize(hoc_norm_coord(py, fy, H), H);
    const float rx = nearbyintf(fminf(fmaxf(ix, -4.0f), (float)W + 4.0f));
    const float ry = nearbyintf(fminf(fmaxf(iy, -4.0f), (float)H + 4.0f));
    *sx = (int)rx;
    *sy = (int)ry;
    return (ix == ix) && (iy == iy) && *sx >= 0 && *sx < W && *sy >= 0 && *sy < H;
}

/* Forward-backward consistency check.  The reference warps a 4-channel tensor
 * [x/W, y/H, mask, mask] there and back with nearest sampling; the grid values are just
 * coordinates, so the double warp is two dependent gathers per pixel:
 *   r --flow_a(r)--> s --flow_b(s)--> q,   occl(r) = M * [ |(q/WH * k - r/WH) * M| < thresh ]
 * with k = m_a(r) in(s) m_b(s) in(q) and M = m_a(r) * (k * m_a(q)).  blockIdx.z selects the
 * direction (0: a = frame 1, 1: a = frame 2). */
__global__ void __launch_bounds__(WP_THREADS)
hoc_occlusion_kernel(const float *__restrict__ mask1, const float *__restrict__ mask2,
                     const float *__restrict__ flow12, const float *__restrict__ flow21, int Cf, int H, int W,
                     float distance_thresh, float *__restrict__ occl1, float *__restrict__ occl2)
{
    const int b = blockIdx.y;
    const long npix = (long)H * W;
    const long pix = (long)blockIdx.x * WP_THREADS + threadIdx.x;
    if (pix >= npix)
        return;
    const bool second = blockIdx.z != 0;
    const float *ma = (second ? mask2 : mask1) + (long)b * npix;
    const float *mb = (second ? mask1 : mask2) + (long)b * npix;
    const float *fa = (second ? flow21 : flow12) + (long)b * Cf * npix;
    const float *fb = (second ? flow12 : flow21) + (long)b * Cf * npix;
    float *out = (second ? occl2 : occl1) + (long)b * npix;

    const int ry = (int)((unsigned)pix / (unsigned)W);
    const int rx = (int)pix - ry * W;
    const float inv_w = __fdiv_rn(1.0f, (float)W), inv_h = __fdiv_rn(1.0f, (float)H);
    const float m_r = ma[pix];
    /* second warp (evaluated at r): source s in the once-warped grid */
    float w0 = 0.0f, w1 = 0.0f, w2 = 0.0f;
    int sx, sy;
    if (hoc_nearest_src(rx, ry, fa[pix], fa[npix + pix], H, W, &sx, &sy)) {
        const long sp = (long)sy * W + sx;
        /* first warp (evaluated at s): source q in the coordinate grid of frame a */
        float g0 = 0.0f, g1 = 0.0f, g2 = 0.0f;
        int qx, qy;
        if (hoc_nearest_src(sx, sy, fb[sp], fb[npix + sp], H, W, &qx, &qy)) {
            g0 = __fmul_rn((float)qx, inv_w);
            g1 = __fmul_rn((float)qy, inv_h);
            g2 = ma[(long)qy * W + qx];
        }
        const float m_s = mb[sp];
        w0 = __fmul_rn(g0, m_s);
        w1 = __fmul_rn(g1, m_s);
        w2 = __fmul_rn(g2, m_s);
    }
    w0 = __fmul_rn(w0, m_r);
    w1 = __fmul_rn(w1, m_r);
    w2 = __fmul_rn(w2, m_r);
    const float M = __fmul_rn(m_r, w2);
    const float dx = __fmul_rn(__fsub_rn(w0, __fmul_rn((float)rx, inv_w)), M);
    const float dy = __fmul_rn(__fsub_rn(w1, __fmul_rn((float)ry, inv_h)), M);
    const float displ = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    out[pix] = __fmul_rn(M, (displ < distance_thresh) ? 1.0f : 0.0f);
}

/* ------------------------------------------------------------------------------------------ */
static int hoc_warp_photo_forward_impl(const float *src, const float *target, const float *flow, const float *jitter,
                                       int B, int C, int Cj, int H, int W, float thresh, float *warped,
                                       float *warp_mask, uint8_t *valid_mask, uint8_t *flow_mask, float *diff,
                                       double *sums, float *loss, void *stream, bool zero_sums);

extern "C" int hoc_warp_photo_forward(const float *src, const float *target, const float *flow, const float *jitter,
                                      int B, int C, int Cj, int H, int W, float thresh, float *warped,
                                      float *warp_mask, uint8_t *valid_mask, uint8_t *flow_mask, float *diff,
                                      double *sums, float *loss, void *stream)
{
    return hoc_warp_photo_forward_impl(src, target, flow, jitter, B, C, Cj, H, W, thresh, warped, warp_mask, valid_mask,
                                       flow_mask, diff, sums, loss, stream, true);
}

extern "C" int hoc_warp_photo_forward_acc(const float *src, const float *target, const float *flow, const float *jitter,
                                          int B, int C, int Cj, int H, int W, float thresh, float *warped,
                                          float *warp_mask, uint8_t *valid_mask, uint8_t *flow_mask, float *diff,
                                          double *sums, float *loss, void *stream)
{
    return hoc_warp_photo_forward_impl(src, target, flow, jitter, B, C, Cj, H, W, thresh, warped, warp_mask, valid_mask,
                                       flow_mask, diff, sums, loss, stream, false);
}

static int hoc_warp_photo_forward_impl(const float *src, const float *target, const float *flow, const float *jitter,
                                       int B, int C, int Cj, int H, int W, float thresh, float *warped,
                                       float *warp_mask, uint8_t *valid_mask, uint8_t *flow_mask, float *diff,
                                       double *sums, float *loss, void *stream, bool zero_sums)
{
    HOC_CHECK_ARG(B >= 0 && C >= 1 && H >= 1 && W >= 1 && (long)H * W < (1l << 31),
                  "hoc_warp_photo_forward: bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
    HOC_CHECK_ARG(B <= 65535, "hoc_warp_photo_forward: batch %d exceeds 65535", B);
    HOC_CHECK_ARG(jitter == nullptr || Cj == 1 || Cj == C, "hoc_warp_photo_forward: jitter channels %d vs %d", Cj, C);
    HOC_CHECK_ARG(warp_mask == nullptr || C <= WP_MAXC, "hoc_warp_photo_forward: warp_mask supports C <= %d", WP_MAXC);
    HOC_CHECK_ARG(sums != nullptr, "hoc_warp_photo_forward: sums is required");
    if (B == 0)
        return HOC_OK;
    HOC_CHECK_ARG(src && target && flow, "hoc_warp_photo_forward: NULL input");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = zero_sums ? cudaMemsetAsync(sums, 0, sizeof(double) * 2 * B, st) : cudaSuccess;
    if (e != cudaSuccess) {
        hoc_set_error("hoc_warp_photo_forward: memset failed: %s", cudaGetErrorString(e));
        return HOC_ERR_CUDA;
    }
    const long npix = (long)H * W;
    /* torch's `tensor / python_int` multiplies by this IEEE-float reciprocal */
    const float inv_w = 1.0f / (float)(W - 1 > 1 ? W - 1 : 1), inv_h = 1.0f / (float)(H - 1 > 1 ? H - 1 : 1);
    dim3 grid((unsigned)((npix + WP_THREADS - 1) / WP_THREADS), B);
    if (C == 3 && (jitter == nullptr || Cj == 3)) /* the reference's case: RGB images, 3-channel jitter masks */
        HOC_LAUNCH(HOC_K_WARP_PHOTO_FWD, st,
                   (hoc_warp_photo_forward_kernel<3, 3><<<grid, WP_THREADS, 0, st>>>(
                       src, target, flow, jitter, C, Cj, H, W, inv_w, inv_h, thresh, warped, warp_mask, valid_mask, flow_mask, diff, sums)));
    else
        HOC_LAUNCH(HOC_K_WARP_PHOTO_FWD, st,
                   (hoc_warp_photo_forward_kernel<0, 0><<<grid, WP_THREADS, 0, st>>>(
                       src, target, flow, jitter, C, Cj, H, W, inv_w, inv_h, thresh, warped, warp_mask, valid_mask, flow_mask, diff, sums)));
    HOC_CHECK_LAUNCH("hoc_warp_photo_forward_kernel");
    if (loss != nullptr) {
        hoc_masked_mean_kernel<<<(B + 127) / 128, 128, 0, st>>>(sums, B, loss);
        HOC_CHECK_LAUNCH("hoc_masked_mean_kernel");
    }
    return HOC_OK;
}

extern "C" int hoc_pair_loss(const double *sums_fwd, const double *sums_bwd, int B, float *loss, void *stream)
{
    HOC_CHECK_ARG(B >= 0, "hoc_pair_loss: negative batch %d", B);
    if (B == 0)
        return HOC_OK;
    HOC_CHECK_ARG(sums_fwd && loss, "hoc_pair_loss: NULL argument");
    HOC_LAUNCH(HOC_K_PAIR_LOSS, (cudaStream_t)stream,
               (hoc_pair_loss_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums_fwd, sums_bwd, B, loss)));
    HOC_CHECK_LAUNCH("hoc_pair_loss_kernel");
    return HOC_OK;
}

extern "C" int hoc_pair_loss_mean(const double *sums_fwd, const double *sums_bwd, int B, float *loss, float *mean,
                                  void *zero, size_t zero_bytes, void *stream)
{
    HOC_CHECK_ARG(B >= 1, "hoc_pair_loss_mean: batch %d", B);
    HOC_CHECK_ARG(sums_fwd && loss && mean, "hoc_pair_loss_mean: NULL argument");
    HOC_CHECK_ARG(zero == nullptr || (zero_bytes % 16 == 0 && ((uintptr_t)zero & 15) == 0),
                  "hoc_pair_loss_mean: zero buffer must be 16-byte aligned with a size multiple of 16");
    const long n_zero = zero ? (long)(zero_bytes / 16) : 0;
    const unsigned extra = n_zero ? (unsigned)((n_zero + 1023) / 1024 < 64 ? (n_zero + 1023) / 1024 : 64) : 0;
    HOC_LAUNCH(HOC_K_PAIR_LOSS, (cudaStream_t)stream,
               (hoc_launch_pdl((hoc_pair_loss_mean_kernel), 1 + extra, 128, 0, (cudaStream_t)stream, sums_fwd, sums_bwd, B, loss, mean,
                                                                                        (uint4 *)zero, n_zero)));
    HOC_CHECK_LAUNCH("hoc_pair_loss_mean_kernel");
    return HOC_OK;
}

extern "C" int hoc_warp_photo_backward(const float *src, const float *target, const float *flow,
                                       const uint8_t *valid_mask, const double *sums, const float *grad_loss, int B,
                                       int C, int H, int W, float thresh, float *grad_flow, void *stream)
{
    HOC_CHECK_ARG(B >= 0 && C >= 1 && H >= 1 && W >= 1 && (long)H * W < (1l << 31), "hoc_warp_photo_backward: bad shape B=%d C=%d H=%d W=%d", B,
                  C, H, W);
    HOC_CHECK_ARG(B <= 65535, "hoc_warp_photo_backward: batch %d exceeds 65535", B);
    if (B == 0)
        return HOC_OK;
    HOC_CHECK_ARG(src && target && flow && valid_mask && sums && grad_loss && grad_flow,
                  "hoc_warp_photo_backward: NULL argument");
    const long npix = (long)H * W;
    dim3 grid((unsigned)((npix + WP_THREADS - 1) / WP_THREADS), B);
    HOC_LAUNCH(HOC_K_WARP_PHOTO_BWD, (cudaStream_t)stream,
               (hoc_warp_photo_backward_kernel<<<grid, WP_THREADS, 0, (cudaStream_t)stream>>>(
                   src, target, flow, valid_mask, sums, grad_loss, C, H, W,
                   1.0f / (float)(W - 1 > 1 ? W - 1 : 1), 1.0f / (float)(H - 1 > 1 ? H - 1 : 1), thresh, grad_flow)));
    HOC_CHECK_LAUNCH("hoc_warp_photo_backward_kernel");
    return HOC_OK;
}

extern "C" int hoc_warp(const float *x, const float *flow_nchw, int B, int C, int H, int W, float thresh, int mode,
                        float *out, float *mask, void *stream)
{
    HOC_CHECK_ARG(B >= 0 && C >= 1 && H >= 1 && W >= 1 && (long)H * W < (1l << 31), "hoc_warp: bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
    HOC_CHECK_ARG(mode == 0 || mode == 1, "hoc_warp: mode %d (0 bilinear, 1 nearest)", mode);
    HOC_CHECK_ARG(B <= 65535, "hoc_warp: batch %d exceeds 65535", B);
    if (B == 0)
        return HOC_OK;
    HOC_CHECK_ARG(x && flow_nchw && out, "hoc_warp: NULL argument");
    const long npix = (long)H * W;
    dim3 grid((unsigned)((npix + WP_THREADS - 1) / WP_THREADS), B);
    HOC_LAUNCH(HOC_K_WARP, (cudaStream_t)stream,
               (hoc_warp_kernel<<<grid, WP_THREADS, 0, (cudaStream_t)stream>>>(x, flow_nchw, C, H, W, thresh, mode, out,
                                                                               mask)));
    HOC_CHECK_LAUNCH("hoc_warp_kernel");
    return HOC_OK;
}

extern "C" int hoc_warp_backward(const float *x, const float *flow_nchw, const float *grad_out, int B, int C, int H,
                                 int W, float thresh, float *grad_flow_nchw, void *stream)
{
    HOC_CHECK_ARG(B >= 0 && C >= 1 && H >= 1 && W >= 1 && (long)H * W < (1l << 31), "hoc_warp_backward: bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
    HOC_CHECK_ARG(B <= 65535, "hoc_warp_backward: batch %d exceeds 65535", B);
    if (B == 0)
        return HOC_OK;
    HOC_CHECK_ARG(x && flow_nchw && grad_out && grad_flow_nchw, "hoc_warp_backward: NULL argument");
    const long npix = (long)H * W;
    dim3 grid((unsigned)((npix + WP_THREADS - 1) / WP_THREADS), B);
    HOC_LAUNCH(HOC_K_WARP_BWD, (cudaStream_t)stream,
               (hoc_warp_backward_kernel<<<grid, WP_THREADS, 0, (cudaStream_t)stream>>>(x, flow_nchw, grad_out, C, H, W,
                                                                                        thresh, grad_flow_nchw)));
    HOC_CHECK_LAUNCH("hoc_warp_backward_kernel");
    return HOC_OK;
}


/* ---- frame-pair entry points (both directions, one launch each way) ----------------------------------------- */
static bool hoc_aligned16(const void *p) { return (((uintptr_t)p) & 15) == 0; }

extern "C" int hoc_warp_photo_forward_pair(const float *image_ref, const float *image, const float *flow12,
                                           const float *flow21, const float *jitter_ref, const float *jitter, int B,
                                           int H, int W, float thresh, int visuals, float *const *warped,
                                           float *const *warp_mask, float *const *diff, uint8_t *const *valid_mask,
                                           uint8_t *const *flow_mask, double *sums, void *stream)
{
    HOC_CHECK_ARG(B >= 0 && H >= 1 && W >= 4 && (W % 4) == 0 && (long)H * W < (1l << 29),
                  "hoc_warp_photo_forward_pair: bad shape B=%d H=%d W=%d (W must be a multiple of 4)", B, H, W);
    HOC_CHECK_ARG(B <= 65535, "hoc_warp_photo_forward_pair: batch %d exceeds 65535", B);
    HOC_CHECK_ARG(sums != nullptr && valid_mask != nullptr, "hoc_warp_photo_forward_pair: sums / valid_mask are required");
    if (B == 0)
        return HOC_OK;
    HOC_CHECK_ARG(image_ref && image && flow12 && flow21 && valid_mask[0] && valid_mask[1],
                  "hoc_warp_photo_forward_pair: NULL input");
    HOC_CHECK_ARG((jitter_ref == nullptr) == (jitter == nullptr), "hoc_warp_photo_forward_pair: one jitter mask missing");
    HocPairDir D[2];
    for (int k = 0; k < 2; k++) {
        D[k].src = k == 0 ? image_ref : image;
        D[k].target = k == 0 ? image : image_ref;
        D[k].flow = k == 0 ? flow21 : flow12;
        D[k].jitter = k == 0 ? jitter : jitter_ref;
        D[k].warped = (visuals && warped) ? warped[k] : nullptr;
        D[k].warp_mask = (visuals && warp_mask) ? warp_mask[k] : nullptr;
        D[k].diff = (visuals && diff) ? diff[k] : nullptr;
        D[k].valid_mask = valid_mask[k];
        D[k].flow_mask = flow_mask ? flow_mask[k] : nullptr;
        D[k].sums = sums + 2 * (size_t)B * k;
        HOC_CHECK_ARG(hoc_aligned16(D[k].src) && hoc_aligned16(D[k].target) && hoc_aligned16(D[k].flow) &&
                          hoc_aligned16(D[k].jitter) && hoc_aligned16(D[k].warped) && hoc_aligned16(D[k].warp_mask) &&
                          hoc_aligned16(D[k].diff) && (((uintptr_t)D[k].valid_mask) & 3) == 0 &&
                          (((uintptr_t)D[k].flow_mask) & 7) == 0,
                      "hoc_warp_photo_forward_pair: tensors must be 16-byte aligned");
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(sums, 0, sizeof(double) * 4 * (size_t)B, st) != cudaSuccess) {
        hoc_set_error("hoc_warp_photo_forward_pair: memset failed");
        return HOC_ERR_CUDA;
    }
    const float inv_w = 1.0f / (float)(W - 1 > 1 ? W - 1 : 1), inv_h = 1.0f / (float)(H - 1 > 1 ? H - 1 : 1);
    const long groups = (long)H * (W / 4);
    dim3 grid((unsigned)((groups + WP_THREADS - 1) / WP_THREADS), B, 2);
    if (visuals)
        HOC_LAUNCH(HOC_K_WARP_PHOTO_FWD, st,
                   (hoc_warp_photo_pair_forward_kernel<true><<<grid, WP_THREADS, 0, st>>>(D[0], D[1], H, W, inv_w, inv_h, thresh)));
    else
        HOC_LAUNCH(HOC_K_WARP_PHOTO_FWD, st,
                   (hoc_warp_photo_pair_forward_kernel<false><<<grid, WP_THREADS, 0, st>>>(D[0], D[1], H, W, inv_w, inv_h, thresh)));
    HOC_CHECK_LAUNCH("hoc_warp_photo_pair_forward_kernel");
    return HOC_OK;
}

extern "C" int hoc_warp_photo_backward_pair(const float *image_ref, const float *image, const float *flow12,
                                            const float *flow21, const uint8_t *const *valid_mask, const double *sums,
                                            const float *mult1, const float *mult2, const float *grad_loss,
                                            const float *grad_mean, int B, int S, int H, int W, int use_backward,
                                            float *grad_rgb1, float *grad_rgb2, float *grad_flow12, float *grad_flow21,
                                            void *zero, size_t zero_bytes, const int *row_lo1, const int *row_lo2,
                                            void *stream)
{
    HOC_CHECK_ARG(B >= 0 && S >= 4 && (S % 4) == 0 && H >= 1 && H <= S && W >= 4 && W <= S && (W % 4) == 0,
                  "hoc_warp_photo_backward_pair: bad shape B=%d S=%d H=%d W=%d (S, W multiples of 4)", B, S, H, W);
    HOC_CHECK_ARG(B <= 65535, "hoc_warp_photo_backward_pair: batch %d exceeds 65535", B);
    if (B == 0)
        return HOC_OK;
    HOC_CHECK_ARG(image_ref && image && flow12 && flow21 && valid_mask && valid_mask[0] && valid_mask[1] && sums &&
                      (grad_loss || grad_mean),
                  "hoc_warp_photo_backward_pair: NULL argument");
    HOC_CHECK_ARG(zero == nullptr || (zero_bytes % 16 == 0 && ((uintptr_t)zero & 15) == 0),
                  "hoc_warp_photo_backward_pair: zero buffer must be 16-byte aligned with a size multiple of 16");
    HOC_CHECK_ARG((grad_rgb1 == nullptr || mult1 != nullptr) && (grad_rgb2 == nullptr || mult2 != nullptr),
                  "hoc_warp_photo_backward_pair: grad_rgb requested without mult");
    HocPairBwdDir D[2];
    /* direction 0: warp(image_ref, flow21) vs image -> d / d flow21 -> render 2; direction 1: the reverse */
    D[0].src = image_ref; D[0].target = image; D[0].flow = flow21; D[0].mult = mult2; D[0].valid_mask = valid_mask[0];
    D[0].sums = sums; D[0].grad_rgb = grad_rgb2; D[0].grad_flow = grad_flow21; D[0].active = 1;
    D[1].src = image; D[1].target = image_ref; D[1].flow = flow12; D[1].mult = mult1; D[1].valid_mask = valid_mask[1];
    D[1].sums = sums + 2 * (size_t)B; D[1].grad_rgb = grad_rgb1; D[1].grad_flow = grad_flow12;
    D[1].active = use_backward ? 1 : 0;
    for (int k = 0; k < 2; k++)
        HOC_CHECK_ARG(hoc_aligned16(D[k].flow) && hoc_aligned16(D[k].mult) && hoc_aligned16(D[k].grad_rgb) &&
                          hoc_aligned16(D[k].grad_flow) && (((uintptr_t)D[k].valid_mask) & 3) == 0,
                      "hoc_warp_photo_backward_pair: tensors must be 16-byte aligned");
    const long groups = (long)S * (S / 4);
    dim3 grid((unsigned)((groups + WPB_THREADS - 1) / WPB_THREADS), B, 2);
    HOC_LAUNCH(HOC_K_WARP_PHOTO_BWD, (cudaStream_t)stream,
               (hoc_warp_photo_pair_backward_kernel<<<grid, WPB_THREADS, 0, (cudaStream_t)stream>>>(
                   D[0], D[1], grad_loss, grad_mean, B, S, H, W, 1.0f / (float)(W - 1 > 1 ? W - 1 : 1),
                   1.0f / (float)(H - 1 > 1 ? H - 1 : 1), (uint4 *)zero, zero ? (long)(zero_bytes / 16) : 0l, row_lo1,
                   row_lo2)));
    HOC_CHECK_LAUNCH("hoc_warp_photo_pair_backward_kernel");
    return HOC_OK;
}

/* uint8 -> float, dst = src / div - sub with the two IEEE operations of `to_tensor` + `normalize` (x / 255 - 0.5 for
 * images; x / 255 for the jitter masks): 16 bytes in, 64 bytes out per thread. */
__global__ void __launch_bounds__(256)
hoc_unpack_u8_kernel(const uint8_t *__restrict__ src, float *__restrict__ dst, long n, float div, float sub)
{
    const long i16 = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (i16 + 16 <= n && (((uintptr_t)(src + i16) | (uintptr_t)(dst + i16)) & 15) == 0) {
        const uint4 v = *reinterpret_cast<const uint4 *>(src + i16);
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            float4 o;
            o.x = __fsub_rn(__fdiv_rn((float)(w[q] & 0xffu), div), sub);
            o.y = __fsub_rn(__fdiv_rn((float)((w[q] >> 8) & 0xffu), div), sub);
            o.z = __fsub_rn(__fdiv_rn((float)((w[q] >> 16) & 0xffu), div), sub);
            o.w = __fsub_rn(__fdiv_rn((float)(w[q] >> 24), div), sub);
            *reinterpret_cast<float4 *>(dst + i16 + 4 * q) = o;
        }
    } else {
        for (long i = i16; i < n && i < i16 + 16; i++)
            dst[i] = __fsub_rn(__fdiv_rn((float)src[i], div), sub);
    }
}

extern "C" int hoc_unpack_u8(const uint8_t *src, float *dst, long long n, float div, float sub, void *stream)
{
    HOC_CHECK_ARG(n >= 0 && div != 0.0f, "hoc_unpack_u8: n = %lld, div = %g", n, (double)div);
    if (n == 0)
        return HOC_OK;
    HOC_CHECK_ARG(src && dst, "hoc_unpack_u8: NULL argument");
    const long nthreads = ((long)n + 15) / 16;
    HOC_LAUNCH(HOC_K_UNPACK_U8, (cudaStream_t)stream,
               (hoc_unpack_u8_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
                   src, dst, (long)n, div, sub)));
    HOC_CHECK_LAUNCH("hoc_unpack_u8_kernel");
    return HOC_OK;
}

extern "C" int hoc_occlusion_mask(const float *mask1, const float *mask2, const float *flow12, const float *flow21,
                                  int B, int Cf, int H, int W, float distance_thresh, float *occl1, float *occl2,
                                  void *stream)
{
    HOC_CHECK_ARG(B >= 0 && Cf >= 2 && H >= 1 && W >= 1 && (long)H * W < (1l << 31), "hoc_occlusion_mask: bad shape B=%d Cf=%d H=%d W=%d", B, Cf,
                  H, W);
    HOC_CHECK_ARG(B <= 65535, "hoc_occlusion_mask: batch %d exceeds 65535", B);
    if (B == 0)
        return HOC_OK;
    HOC_CHECK_ARG(mask1 && mask2 && flow12 && flow21 && occl1 && occl2, "hoc_occlusion_mask: NULL argument");
    const long npix = (long)H * W;
    dim3 grid((unsigned)((npix + WP_THREADS - 1) / WP_THREADS), B, 2);
    HOC_LAUNCH(HOC_K_OCCLUSION, (cudaStream_t)stream,
               (hoc_occlusion_kernel<<<grid, WP_THREADS, 0, (cudaStream_t)stream>>>(mask1, mask2, flow12, flow21, Cf, H, W,
                                                                                    distance_thresh, occl1, occl2)));
    HOC_CHECK_LAUNCH("hoc_occlusion_kernel");
    return HOC_OK;
}
