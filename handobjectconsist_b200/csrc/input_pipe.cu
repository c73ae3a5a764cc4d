/*
 * input_pipe.cu -- the image side of the frame-pair input pipeline on the GPU (sm_100a), SURVEY.md section 8(f) row f3.
 *
 * Replaces, for the two frames of a pair, what HandObjSet.get_sample does per sample on the dataset workers' CPUs with
 * PIL (/root/reference/meshreg/datasets/handobjset.py:336-379, through libyana's transform_img / apply_jitter):
 *   colour jitter of the source frame   torchvision adjust_brightness / _saturation / _hue / _contrast on PIL images,
 *                                       in the (shuffled) order the caller passes
 *   affine crop + rotation to inp_res   Image.transform(res, Image.AFFINE, inv(affinetrans)), nearest sampling
 *   to_tensor + normalize(0.5, 1)       x / 255 - 0.5, CHW float32
 *   jitter mask                         a white image through the same transform: 1 where a source pixel exists
 * The decoded uint8 frames cross PCIe once (a quarter of the bytes of the float tensors the reference's workers ship)
 * and one launch writes both frames' images and masks straight into the buffers the consistency step reads.
 *
 * Bit compatibility with PIL (oracle/inputpipe.py restates every step and is pinned against PIL / torchvision):
 *   - AFFINE + NEAREST is PIL's 16.16 fixed-point walk (Geometry.c affine_fixed): the six integer coefficients are
 *     computed on the host exactly as PIL computes them; source = (c + x a + y b) >> 16;
 *   - brightness / saturation / contrast are Image.blend: float32 deg + f * (img - deg), clipped, truncated -- the
 *     pointwise adjustments commute with nearest sampling, so they run on the sampled pixel; the one global quantity,
 *     the rounded grey mean of the (already adjusted) source frame that contrast blends with, comes from a reduction
 *     pass over the source (integer sums: order-independent);
 *   - hue follows Convert.c's rgb2hsv_row / hsv2rgb_row (float / double mix) with the wrapping uint8 shift.
 * Compiled with -fmad=false like the rest of the library.  The Gaussian blur of handobjset.py:338-339 is NOT done here.
 */
#include "hoc_common.cuh"

#define IP_THREADS 256
#define IP_OP_BRIGHTNESS 0
#define IP_OP_SATURATION 1
#define IP_OP_HUE 2
#define IP_OP_CONTRAST 3

struct HocU8x3 {
    int r, g, b;
};

__device__ __forceinline__ int hoc_ip_gray(const HocU8x3 &p)
{
    return (p.r * 19595 + p.g * 38470 + p.b * 7471 + 0x8000) >> 16;
}

/* Image.blend(degenerate, image, factor), one channel */
__device__ __forceinline__ int hoc_ip_blend(int deg, int img, float factor)
{
    const float d = (float)deg;
    float v = __fadd_rn(d, __fmul_rn(factor, __fsub_rn((float)img, d)));
    v = fminf(fmaxf(v, 0.0f), 255.0f);
    return (int)v;
}

__device__ __forceinline__ HocU8x3 hoc_ip_hue(const HocU8x3 &p, int shift)
{
    /* rgb2hsv_row */
    const int maxc = max(p.r, max(p.g, p.b)), minc = min(p.r, min(p.g, p.b));
    int uh = 0, us = 0;
    const int uv = maxc;
    if (minc != maxc) {
        const float cr = (float)(maxc - minc);
        const float s = __fdiv_rn(cr, (float)maxc);
        const float rc = __fdiv_rn((float)(maxc - p.r), cr), gc = __fdiv_rn((float)(maxc - p.g), cr),
                    bc = __fdiv_rn((float)(maxc - p.b), cr);
        float h;
        if (p.r == maxc)
            h = __fsub_rn(bc, gc);
        else if (p.g == maxc)
            h = (float)(2.0 + (double)rc - (double)bc);
        else
            h = (float)(4.0 + (double)gc - (double)rc);
        h = (float)fmod((double)h / 6.0 + 1.0, 1.0);
        uh = min(max((int)((double)h * 255.0), 0), 255);
        us = min(max((int)((double)s * 255.0), 0), 255);
    }
    uh = (uh + shift) & 0xff;
    /* hsv2rgb_row */
    HocU8x3 o;
    if (us == 0) {
        o.r = o.g = o.b = uv;
        return o;
    }
    const float fh = __fdiv_rn(__fmul_rn((float)uh, 6.0f), 255.0f);
    const float fs = __fdiv_rn((float)us, 255.0f);
    const int i = (int)floorf(fh);
    const float f = __fsub_rn(fh, (float)i);
    const float vf = (float)uv;
    const int pp = (int)floorf(__fadd_rn(__fmul_rn(vf, __fsub_rn(1.0f, fs)), 0.5f));
    const int q = (int)floorf(__fadd_rn(__fmul_rn(vf, __fsub_rn(1.0f, __fmul_rn(fs, f))), 0.5f));
    const int t = (int)floorf(__fadd_rn(__fmul_rn(vf, __fsub_rn(1.0f, __fmul_rn(fs, __fsub_rn(1.0f, f)))), 0.5f));
    switch (i % 6) {
    case 0: o.r = uv; o.g = t; o.b = pp; break;
    case 1: o.r = q; o.g = uv; o.b = pp; break;
    case 2: o.r = pp; o.g = uv; o.b = t; break;
    case 3: o.r = pp; o.g = q; o.b = uv; break;
    case 4: o.r = t; o.g = pp; o.b = uv; break;
    default: o.r = uv; o.g = pp; o.b = q; break;
    }
    o.r = min(max(o.r, 0), 255);
    o.g = min(max(o.g, 0), 255);
    o.b = min(max(o.b, 0), 255);
    return o;
}

/* the adjustments of `order` (4 ids, -1 = none) up to (not including) position `stop` */
__device__ __forceinline__ HocU8x3 hoc_ip_jitter(HocU8x3 p, const int *order, int stop, float brightness, float saturation,
                                                float contrast, int hue_shift, int gray_mean)
{
    for (int k = 0; k < stop; k++) {
        const int op = order[k];
        if (op == IP_OP_BRIGHTNESS) {
            p.r = hoc_ip_blend(0, p.r, brightness);
            p.g = hoc_ip_blend(0, p.g, brightness);
            p.b = hoc_ip_blend(0, p.b, brightness);
        } else if (op == IP_OP_SATURATION) {
            const int l = hoc_ip_gray(p);
            p.r = hoc_ip_blend(l, p.r, saturation);
            p.g = hoc_ip_blend(l, p.g, saturation);
            p.b = hoc_ip_blend(l, p.b, saturation);
        } else if (op == IP_OP_HUE) {
            p = hoc_ip_hue(p, hue_shift);
        } else if (op == IP_OP_CONTRAST) {
            p.r = hoc_ip_blend(gray_mean, p.r, contrast);
            p.g = hoc_ip_blend(gray_mean, p.g, contrast);
            p.b = hoc_ip_blend(gray_mean, p.b, contrast);
        }
    }
    return p;
}

__device__ __forceinline__ int hoc_ip_contrast_pos(const int *order)
{
    for (int k = 0; k < 4; k++)
        if (order[k] == IP_OP_CONTRAST)
            return k;
    return -1;
}

/* Pass 1: per (sample, frame) the sum of the grey values of the source frame after the adjustments that precede the
 * contrast adjustment (ImageStat.Stat(img.convert("L")).mean).  grid (chunks, B, 2).  Integer sums. */
__global__ void __launch_bounds__(IP_THREADS)
hoc_augment_gray_sum_kernel(const uint8_t *__restrict__ frame0, const uint8_t *__restrict__ frame1, int Hs, int Ws,
                            const float *__restrict__ color, const int *__restrict__ hue_shift,
                            const int *__restrict__ order, unsigned long long *__restrict__ sums)
{
    const int b = blockIdx.y, fr = blockIdx.z;
    const int *ord = order + ((long)b * 2 + fr) * 4;
    const int stop = hoc_ip_contrast_pos(ord);
    if (stop < 0)
        return; /* no contrast adjustment for this frame: the mean is not needed */
    const uint8_t *src = (fr ? frame1 : frame0) + (long)b * Hs * Ws * 3;
    const float br = color[b * 3 + 0], sa = color[b * 3 + 1];
    const int hs = hue_shift[b];
    const long n = (long)Hs * Ws;
    unsigned long long acc = 0;
    for (long i = (long)blockIdx.x * IP_THREADS + threadIdx.x; i < n; i += (long)gridDim.x * IP_THREADS) {
        HocU8x3 p;
        p.r = src[i * 3];
        p.g = src[i * 3 + 1];
        p.b = src[i * 3 + 2];
        p = hoc_ip_jitter(p, ord, stop, br, sa, 1.0f, hs, 0);
        acc += (unsigned)hoc_ip_gray(p);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        acc += __shfl_xor_sync(HOC_FULL_MASK, acc, o);
    __shared__ unsigned long long s_acc[IP_THREADS / 32];
    if ((threadIdx.x & 31) == 0)
        s_acc[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < IP_THREADS / 32; w++)
            t += s_acc[w];
        atomicAdd(sums + b * 2 + fr, t);
    }
}

/* Pass 2: the crop.  grid (ceil(H * W / 4 / threads), B, 2 frames); four output pixels of a row per thread (16-byte
 * stores of the six output planes); the source pixel of each is a 3-byte gather. */
__global__ void __launch_bounds__(IP_THREADS)
hoc_augment_frames_kernel(const uint8_t *__restrict__ frame0, const uint8_t *__restrict__ frame1, int Hs, int Ws,
                          const int *__restrict__ coef, const float *__restrict__ color,
                          const int *__restrict__ hue_shift, const int *__restrict__ order,
                          const unsigned long long *__restrict__ sums, int H, int W, float *__restrict__ image0,
                          float *__restrict__ image1, float *__restrict__ mask0, float *__restrict__ mask1)
{
    const int b = blockIdx.y, fr = blockIdx.z;
    const int W4 = W >> 2;
    const int q = blockIdx.x * IP_THREADS + threadIdx.x;
    if (q >= H * W4)
        return;
    const int y = q / W4, x0 = (q - y * W4) << 2;
    const uint8_t *src = (fr ? frame1 : frame0) + (long)b * Hs * Ws * 3;
    const int *cf = coef + b * 6;
    const long a0 = cf[0], a1 = cf[1], a2 = cf[2], a3 = cf[3], a4 = cf[4], a5 = cf[5];
    const int *ord = (order != nullptr) ? order + ((long)b * 2 + fr) * 4 : nullptr;
    float br = 1.0f, sa = 1.0f, co = 1.0f;
    int hs = 0, mean = 0;
    if (ord != nullptr) {
        br = color[b * 3 + 0];
        sa = color[b * 3 + 1];
        co = color[b * 3 + 2];
        hs = hue_shift[b];
        if (hoc_ip_contrast_pos(ord) >= 0) /* int(mean + 0.5) of the grey image */
            mean = (int)((double)sums[b * 2 + fr] / (double)((long)Hs * Ws) + 0.5);
    }
    float v[3][4], m[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const long xin = (a2 + (long)y * a1 + (long)(x0 + j) * a0) >> 16;
        const long yin = (a5 + (long)y * a4 + (long)(x0 + j) * a3) >> 16;
        const bool inside = xin >= 0 && xin < Ws && yin >= 0 && yin < Hs;
        HocU8x3 p = {0, 0, 0};
        if (inside) {
            const uint8_t *s = src + (yin * Ws + xin) * 3;
            p.r = s[0];
            p.g = s[1];
            p.b = s[2];
            if (ord != nullptr)
                p = hoc_ip_jitter(p, ord, 4, br, sa, co, hs, mean);
        }
        /* to_tensor (x / 255) and normalize (x - 0.5) / 1 */
        v[0][j] = __fsub_rn(__fdiv_rn((float)p.r, 255.0f), 0.5f);
        v[1][j] = __fsub_rn(__fdiv_rn((float)p.g, 255.0f), 0.5f);
        v[2][j] = __fsub_rn(__fdiv_rn((float)p.b, 255.0f), 0.5f);
        m[j] = inside ? 1.0f : 0.0f;
    }
    float *img = fr ? image1 : image0, *msk = fr ? mask1 : mask0;
    const long npix = (long)H * W;
    const long o = (long)b * 3 * npix + (long)y * W + x0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        *reinterpret_cast<float4 *>(img + o + c * npix) = make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
        if (msk != nullptr)
            *reinterpret_cast<float4 *>(msk + o + c * npix) = make_float4(m[0], m[1], m[2], m[3]);
    }
}

extern "C" size_t hoc_augment_frame_pair_workspace_bytes(int B)
{
    return B > 0 ? sizeof(unsigned long long) * 2 * (size_t)B : 0;
}

extern "C" int hoc_augment_frame_pair(const uint8_t *frame0, const uint8_t *frame1, int B, int Hs, int Ws,
                                      const int *coef_fix16, const float *color, const int *hue_shift,
                                      const int *order, int H, int W, float *image0, float *image1, float *mask0,
                                      float *mask1, void *workspace, size_t workspace_bytes, void *stream)
{
    HOC_CHECK_ARG(B >= 0 && Hs >= 1 && Ws >= 1 && H >= 1 && W >= 4 && (W % 4) == 0 && B <= 65535,
                  "hoc_augment_frame_pair: bad shape B=%d source %dx%d crop %dx%d (W must be a multiple of 4)", B, Ws, Hs,
                  W, H);
    if (B == 0)
        return HOC_OK;
    HOC_CHECK_ARG(frame0 && frame1 && coef_fix16 && image0 && image1, "hoc_augment_frame_pair: NULL argument");
    HOC_CHECK_ARG((mask0 == nullptr) == (mask1 == nullptr), "hoc_augment_frame_pair: one jitter mask missing");
    HOC_CHECK_ARG(order == nullptr || (color != nullptr && hue_shift != nullptr),
                  "hoc_augment_frame_pair: colour jitter order given without its parameters");
    HOC_CHECK_ARG(((((uintptr_t)image0 | (uintptr_t)image1 | (uintptr_t)mask0 | (uintptr_t)mask1)) & 15) == 0,
                  "hoc_augment_frame_pair: outputs must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long *sums = (unsigned long long *)workspace;
    if (order != nullptr) {
        const size_t need = hoc_augment_frame_pair_workspace_bytes(B);
        if (workspace == nullptr || workspace_bytes < need) {
            hoc_set_error("hoc_augment_frame_pair: workspace of %zu bytes needed, %zu given", need, workspace_bytes);
            return HOC_ERR_WORKSPACE;
        }
        if (cudaMemsetAsync(sums, 0, need, st) != cudaSuccess) {
            hoc_set_error("hoc_augment_frame_pair: memset failed");
            return HOC_ERR_CUDA;
        }
        const long n = (long)Hs * Ws;
        dim3 g1((unsigned)((n + IP_THREADS * 8 - 1) / (IP_THREADS * 8) < 64 ? (n + IP_THREADS * 8 - 1) / (IP_THREADS * 8) : 64),
                B, 2);
        HOC_LAUNCH(HOC_K_AUGMENT_STATS, st,
                   (hoc_augment_gray_sum_kernel<<<g1, IP_THREADS, 0, st>>>(frame0, frame1, Hs, Ws, color, hue_shift, order,
                                                                           sums)));
        HOC_CHECK_LAUNCH("hoc_augment_gray_sum_kernel");
    }
    const long groups = (long)H * (W / 4);
    dim3 g2((unsigned)((groups + IP_THREADS - 1) / IP_THREADS), B, 2);
    HOC_LAUNCH(HOC_K_AUGMENT_FRAMES, st,
               (hoc_augment_frames_kernel<<<g2, IP_THREADS, 0, st>>>(frame0, frame1, Hs, Ws, coef_fix16, color, hue_shift,
                                                                     order, sums, H, W, image0, image1, mask0, mask1)));
    HOC_CHECK_LAUNCH("hoc_augment_frames_kernel");
    return HOC_OK;
}
