/*
 * flow_path.cu -- the glue of the mesh-flow path as kernels (sm_100a).
 *
 * get_opticalflow (/root/reference/meshreg/warping/opticalflow.py:98-154) wraps each of its two renders
 * in ~40 small tensor ops: batch_vertex_textures (gather + zero-fill of [B,F,2,2,2,3]), fill_back
 * (two torch.cat and a permute, /root/reference/meshreg/neurender/renderer.py:250-252),
 * vertices_to_faces (renderer.py:282), and after the render the alpha threshold, the ignore-face mask
 * (a [B,S,S,14] temporary), row flips, the occlusion check (eight grid_sample calls), mask products,
 * permute, channel slice and crop (opticalflow.py:109-154).  Their autograd adjoints double the count.
 * Here:
 *   hoc_mesh_gather      verts (NDC) + int faces + per-vertex attributes -> the rasterizer's inputs
 *                        faces [B,F',3,3] and textures [B,F',2,2,2,3] with both windings (F' = 2F)
 *   hoc_mesh_scatter     adjoint: grad_faces / grad_textures -> grad_verts, grad_attrs (atomics into
 *                        [B,V,3]; the reference's index_put(accumulate) does the same)
 *   hoc_flow_finalize    two renders -> masks -> forward-backward occlusion check -> final flows
 *                        [B,H,W,2] (cropped), one launch for both directions
 *   hoc_flow_finalize_backward   grad of the flows -> grad of the rendered rgb maps
 */
#include "hoc_common.cuh"
#include "hoc_det.cuh"
#include "warp_math.cuh"

#define FP_THREADS 256

/* ------------------------------------------------------------------------------------------ */
#ifndef GA_THREADS
#define GA_THREADS 256 /* (128: 13.6 us against 11.5 -- the key fill wants the threads) */
#endif
__global__ void __launch_bounds__(GA_THREADS)
hoc_mesh_gather_kernel(const float *__restrict__ verts, const float *__restrict__ attrs,
                       const long long *__restrict__ faces_idx, int V, int F, int fill_back,
                       float *__restrict__ faces_out, float *__restrict__ tex_out, int tex_vertex,
                       uint4 *__restrict__ clear, long n_clear)
{
    if (clear != nullptr) { /* 0xff fill of the z-buffer keys of the forward that follows, spread over the grid */
        const long nthreads = (long)gridDim.x * gridDim.y * GA_THREADS;
        hoc_fill16(clear, n_clear, ((long)blockIdx.y * gridDim.x + blockIdx.x) * GA_THREADS + threadIdx.x, nthreads,
                   0xffffffffu);
    }
    /* per-face records are staged in shared memory and written out as contiguous, coalesced runs */
    __shared__ __align__(16) float s_tex[GA_THREADS * 24];
    __shared__ float s_face[GA_THREADS * 9];
    const int Fo = fill_back ? 2 * F : F;
    const int fo0 = blockIdx.x * GA_THREADS;
    const int fo = fo0 + threadIdx.x;
    const int b = blockIdx.y;
    if (fo < Fo) {
        const int f = fo >= F ? fo - F : fo;
        const long long *fi = faces_idx + ((long)b * F + f) * 3;
        long long i0 = fi[0], i1 = fi[1], i2 = fi[2];
        if (fo >= F) { /* reversed winding: (v2, v1, v0) */
            const long long t = i0;
            i0 = i2;
            i2 = t;
        }
        /* precondition 0 <= index < V (include/hoc_b200.h); a bad table must not read out of bounds */
        const long long iv[3] = {(i0 >= 0 && i0 < V) ? i0 : 0, (i1 >= 0 && i1 < V) ? i1 : 0,
                                 (i2 >= 0 && i2 < V) ? i2 : 0};
        float c[3][3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float *vs = verts + ((long)b * V + iv[k]) * 3;
            s_face[threadIdx.x * 9 + 3 * k + 0] = vs[0];
            s_face[threadIdx.x * 9 + 3 * k + 1] = vs[1];
            s_face[threadIdx.x * 9 + 3 * k + 2] = vs[2];
            if (tex_out != nullptr) {
                const float *as = attrs + ((long)b * V + iv[k]) * 3;
                c[k][0] = as[0];
                c[k][1] = as[1];
                c[k][2] = as[2];
            }
        }
        if (tex_out != nullptr && tex_vertex) {
            /* vertex mode: the three vertex values; the rasterizer evaluates the cube's texels on the fly */
#pragma unroll
            for (int k = 0; k < 3; k++)
#pragma unroll
                for (int ch = 0; ch < 3; ch++)
                    s_tex[threadIdx.x * 9 + 3 * k + ch] = c[k][ch];
        } else if (tex_out != nullptr) {
            /* cube whose trilinear sample at the barycentric coordinates is b0 c0 + b1 c1 + b2 c2:
             * T[i,j,k] = i c0 + j c1 + k c2 (the reversed copy is the permute(0,1,4,3,2,5) of the original) */
#pragma unroll
            for (int corner = 0; corner < 8; corner++) {
                const float wi = (float)((corner >> 2) & 1), wj = (float)((corner >> 1) & 1), wk = (float)(corner & 1);
#pragma unroll
                for (int ch = 0; ch < 3; ch++)
                    s_tex[threadIdx.x * 24 + corner * 3 + ch] = wi * c[0][ch] + wj * c[1][ch] + wk * c[2][ch];
            }
        }
    }
    __syncthreads();
    const int nf = min(GA_THREADS, Fo - fo0);
    float *fd = faces_out + ((long)b * Fo + fo0) * 9;
    for (int i = threadIdx.x; i < nf * 9; i += GA_THREADS)
        fd[i] = s_face[i];
    if (tex_out != nullptr && tex_vertex) {
        float *td = tex_out + ((long)b * Fo + fo0) * 9;
        for (int i = threadIdx.x; i < nf * 9; i += GA_THREADS)
            td[i] = s_tex[i];
    } else if (tex_out != nullptr) {
        float4 *td = reinterpret_cast<float4 *>(tex_out + ((long)b * Fo + fo0) * 24);
        const float4 *ts4 = reinterpret_cast<const float4 *>(s_tex);
        for (int i = threadIdx.x; i < nf * 6; i += GA_THREADS)
            td[i] = ts4[i];
    }
}

__global__ void __launch_bounds__(FP_THREADS)
hoc_mesh_scatter_kernel(const float *__restrict__ grad_faces, const float *__restrict__ grad_tex,
                        const long long *__restrict__ faces_idx, int V, int F, int fill_back, int tex_grad_mode,
                        float *__restrict__ grad_verts, float *__restrict__ grad_attrs,
                        unsigned long long *__restrict__ det_v, unsigned long long *__restrict__ det_a)
{
    hoc_pdl_sync(); /* programmatic dependent launch: see hoc_common.cuh */
    const int Fo = fill_back ? 2 * F : F;
    const int fo = blockIdx.x * FP_THREADS + threadIdx.x;
    const int b = blockIdx.y;
    if (fo >= Fo)
        return;
    const int f = fo >= F ? fo - F : fo;
    float gf[9], gc[3][3];
    bool any = false;
    if (grad_faces != nullptr && grad_verts != nullptr) {
        const float *src = grad_faces + ((long)b * Fo + fo) * 9;
#pragma unroll
        for (int k = 0; k < 9; k++) {
            gf[k] = src[k];
            any = any || (gf[k] != 0.0f);
        }
    }
    bool any_t = false;
    if (grad_tex != nullptr && grad_attrs != nullptr && tex_grad_mode == HOC_TEX_GRAD_VERTEX) {
        const float *src = grad_tex + ((long)b * Fo + fo) * 9;
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                gc[k][ch] = src[3 * k + ch];
                any_t = any_t || (gc[k][ch] != 0.0f);
            }
    } else if (grad_tex != nullptr && grad_attrs != nullptr) {
        const float4 *src = reinterpret_cast<const float4 *>(grad_tex + ((long)b * Fo + fo) * 24);
        float g[24];
#pragma unroll
        for (int q = 0; q < 6; q++) {
            const float4 v = src[q];
            g[4 * q] = v.x;
            g[4 * q + 1] = v.y;
            g[4 * q + 2] = v.z;
            g[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            /* d T[i,j,k] / d c0 = i, / d c1 = j, / d c2 = k */
            gc[0][ch] = g[4 * 3 + ch] + g[5 * 3 + ch] + g[6 * 3 + ch] + g[7 * 3 + ch];
            gc[1][ch] = g[2 * 3 + ch] + g[3 * 3 + ch] + g[6 * 3 + ch] + g[7 * 3 + ch];
            gc[2][ch] = g[1 * 3 + ch] + g[3 * 3 + ch] + g[5 * 3 + ch] + g[7 * 3 + ch];
        }
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
            for (int ch = 0; ch < 3; ch++)
                any_t = any_t || (gc[k][ch] != 0.0f);
    }
    if (!any && !any_t)
        return;
    const long long *fi = faces_idx + ((long)b * F + f) * 3;
    long long i0 = fi[0], i1 = fi[1], i2 = fi[2];
    if (fo >= F) {
        const long long t = i0;
        i0 = i2;
        i2 = t;
    }
    const long long iv[3] = {i0, i1, i2};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (iv[k] < 0 || iv[k] >= V)
            continue; /* precondition violated: skip instead of writing out of bounds */
        const long dst = ((long)b * V + iv[k]) * 3;
        if (any) {
#pragma unroll
            for (int d = 0; d < 3; d++)
                if (gf[3 * k + d] != 0.0f)
                    hoc_accum(grad_verts, dst + d, gf[3 * k + d], det_v);
        }
        if (any_t) {
#pragma unroll
            for (int ch = 0; ch < 3; ch++)
                if (gc[k][ch] != 0.0f)
                    hoc_accum(grad_attrs, dst + ch, gc[k][ch], det_a);
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Nearest-mode source pixel of the reference's warp() (imgflowarp.py:31-55) -- same arithmetic as
 * warp_photo.cu (ATen CUDA op order, see there). */
__device__ __forceinline__ float hoc_fp_norm(int p, float flow, int size)
{
    const float inv = __fdiv_rn(1.0f, (float)max(size - 1, 1));
    return __fadd_rn(__fmul_rn(__fmul_rn(2.0f, __fadd_rn((float)p, flow)), inv), -1.0f);
}
__device__ __forceinline__ bool hoc_fp_nearest(int px, int py, float fx, float fy, int S, int *sx, int *sy)
{
    const float ix = __fmul_rn(__fmaf_rn(__fadd_rn(hoc_fp_norm(px, fx, S), 1.0f), (float)S, -1.0f), 0.5f);
    const float iy = __fmul_rn(__fmaf_rn(__fadd_rn(hoc_fp_norm(py, fy, S), 1.0f), (float)S, -1.0f), 0.5f);
    const float rx = nearbyintf(fminf(fmaxf(ix, -4.0f), (float)S + 4.0f));
    const float ry = nearbyintf(fminf(fmaxf(iy, -4.0f), (float)S + 4.0f));
    *sx = (int)rx;
    *sy = (int)ry;
    return (ix == ix) && (iy == iy) && *sx >= 0 && *sx < S && *sy >= 0 && *sy < S;
}

struct HocRender {
    const float *rgb;      /* [B,3,S,S] image layout */
    const float *alpha;    /* [B,S,S]   image layout */
    const int32_t *idx;    /* [B,S,S]   raster order (rows NOT flipped) */
};

/* thresholded alpha x keep-mask of the ignored faces at image pixel (x, y) (opticalflow.py:109-116) */
/* The ignored faces: list + the range [lo, hi] its entries lie in (a face outside the range is kept without a walk
 * over the list: three mask evaluations per covered pixel times 14 ignored faces were 15 % of the finalize pass). */
struct HocIgnore {
    const int *list;
    int n, lo, hi;
};

__device__ __forceinline__ bool hoc_fp_keep(const HocIgnore &G, int fidx)
{
    bool keep = true;
    if (fidx >= G.lo && fidx <= G.hi)
        for (int k = 0; k < G.n; k++)
            keep = keep && (fidx != G.list[k]);
    return keep;
}

__device__ __forceinline__ float hoc_fp_mask(const HocRender &R, int b, int S, int x, int y, const HocIgnore &G,
                                             float *alpha_out)
{
    const long po = ((long)b * S + y) * S + x;
    const float a = R.alpha[po];
    *alpha_out = a;
    float m = (a > 0.99999f) ? 1.0f : 0.0f;
    if (G.n > 0 && m != 0.0f) /* the keep-mask only matters where alpha passed the threshold */
        m = __fmul_rn(m, hoc_fp_keep(G, R.idx[((long)b * S + (S - 1 - y)) * S + x]) ? 1.0f : 0.0f);
    return m;
}

/* pred_flow = rgb * mask at image pixel (x, y), first two channels */
__device__ __forceinline__ void hoc_fp_flow(const HocRender &R, int b, int S, int x, int y, float m, float *fx,
                                            float *fy)
{
    const long o = (((long)b * 3) * S + y) * S + x;
    *fx = __fmul_rn(R.rgb[o], m);
    *fy = __fmul_rn(R.rgb[o + (long)S * S], m);
}

/*
 * Both directions in one launch (blockIdx.z).  Direction a -> b at pixel r of render a:
 *   mask_a(r) = [alpha_a > 0.99999] * keep_a                      (opticalflow.py:109-116)
 *   pf_a(r)   = rgb_a(r) * mask_a(r)                              (:118)
 *   occlusion check (imgflowarp.py:118-172): r --pf_a--> s --pf_b--> q with nearest sampling,
 *   occl_a(r) = M * [ |(grid_a(q) k - grid_a(r)) M| < 0.03 ],  k = m_a(r) in(s) m_b(s) in(q), M = m_a(r) k m_a(q)
 *   where, as in the reference, the mask used for render 2 inside the check is its raw alpha (sic, :139)
 *   mask_a' = mask_a * occl_a  (render 2: alpha * occl_2);   flow_a = pf_a * mask_a'   (:146-150)
 * Output: flow [B,H,W,2] (cropped to H x W) and mult [B,H,W] = d flow / d rgb = mask_a * mask_a'.
 */
/* One pixel r = (rx, ry) of direction a -> b (see the comment above): `alpha_r`, `rgb0`, `rgb1` are render a's alpha and
 * first two colour channels at r (loaded by the caller, scalar or vectorised).  Returns the final flow and
 * mult = d flow / d rgb. */
__device__ __forceinline__ void hoc_finalize_pixel(const HocRender &Ra, const HocRender &Rb, bool second, int b, int S,
                                                   int rx, int ry, float alpha_r, float rgb0, float rgb1,
                                                   const HocIgnore &G, int mask_occlusions,
                                                   float distance_thresh, float *fx_out, float *fy_out, float *mult_out)
{
    float mt_r = (alpha_r > 0.99999f) ? 1.0f : 0.0f; /* thresholded alpha x keep-mask (hoc_fp_mask on preloaded alpha) */
    if (G.n > 0 && mt_r != 0.0f)
        mt_r = __fmul_rn(mt_r, hoc_fp_keep(G, Ra.idx[((long)b * S + (S - 1 - ry)) * S + rx]) ? 1.0f : 0.0f);
    float fx = __fmul_rn(rgb0, mt_r), fy = __fmul_rn(rgb1, mt_r);
    float mfinal = mt_r;
    if (mask_occlusions && (second ? alpha_r : mt_r) == 0.0f && fx == fx && fy == fy) {
        /* the pixel's own mask is zero (93 % of a typical frame): every product below is zero */
        *fx_out = 0.0f;
        *fy_out = 0.0f;
        *mult_out = 0.0f;
        return;
    }
    if (mask_occlusions) {
        /* masks that enter the check: render 1 -> thresholded*keep, render 2 -> raw alpha (sic) */
        const float m_r = second ? alpha_r : mt_r;
        const float inv_s = __fdiv_rn(1.0f, (float)S);
        float w0 = 0.0f, w1 = 0.0f, w2 = 0.0f;
        int sx, sy;
        if (hoc_fp_nearest(rx, ry, fx, fy, S, &sx, &sy)) {
            float alpha_s;
            const float mt_s = hoc_fp_mask(Rb, b, S, sx, sy, G, &alpha_s);
            float sfx, sfy;
            hoc_fp_flow(Rb, b, S, sx, sy, mt_s, &sfx, &sfy);
            const float m_s = second ? mt_s : alpha_s; /* mask of the OTHER render in the check */
            float g0 = 0.0f, g1 = 0.0f, g2 = 0.0f;
            int qx, qy;
            if (hoc_fp_nearest(sx, sy, sfx, sfy, S, &qx, &qy)) {
                float alpha_q;
                const float mt_q = hoc_fp_mask(Ra, b, S, qx, qy, G, &alpha_q);
                g0 = __fmul_rn((float)qx, inv_s);
                g1 = __fmul_rn((float)qy, inv_s);
                g2 = second ? alpha_q : mt_q;
            }
            w0 = __fmul_rn(g0, m_s);
            w1 = __fmul_rn(g1, m_s);
            w2 = __fmul_rn(g2, m_s);
        }
        w0 = __fmul_rn(w0, m_r);
        w1 = __fmul_rn(w1, m_r);
        w2 = __fmul_rn(w2, m_r);
        const float M = __fmul_rn(m_r, w2);
        const float dx = __fmul_rn(__fsub_rn(w0, __fmul_rn((float)rx, inv_s)), M);
        const float dy = __fmul_rn(__fsub_rn(w1, __fmul_rn((float)ry, inv_s)), M);
        const float displ = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
        const float occl = __fmul_rn(M, (displ < distance_thresh) ? 1.0f : 0.0f);
        mfinal = __fmul_rn(m_r, occl); /* mask_flow * occl_mask (render 2: alpha * occl) */
        fx = __fmul_rn(fx, mfinal);
        fy = __fmul_rn(fy, mfinal);
    }
    *fx_out = fx;
    *fy_out = fy;
    *mult_out = mask_occlusions ? __fmul_rn(mt_r, mfinal) : mt_r;
}

__global__ void __launch_bounds__(FP_THREADS)
hoc_flow_finalize_kernel(HocRender R1, HocRender R2, int S, int H, int W, const int *__restrict__ ignore, int n_ignore,
                         int mask_occlusions, float distance_thresh, float *__restrict__ flow12,
                         float *__restrict__ flow21, float *__restrict__ mult1, float *__restrict__ mult2)
{
    const int b = blockIdx.y;
    const bool second = blockIdx.z != 0;
    const long pix = (long)blockIdx.x * FP_THREADS + threadIdx.x;
    if (pix >= (long)H * W)
        return;
    const int ry = (int)((unsigned)pix / (unsigned)W), rx = (int)pix - ry * W;
    const HocRender &Ra = second ? R2 : R1;
    const HocRender &Rb = second ? R1 : R2;
    const long po = ((long)b * S + ry) * S + rx, co = (((long)b * 3) * S + ry) * S + rx;
    float fx, fy, mu;
    HocIgnore G; /* (this one-pixel-per-thread kernel walks the list for every covered pixel) */
    G.list = ignore; G.n = n_ignore; G.lo = -0x7fffffff - 1; G.hi = 0x7fffffff;
    hoc_finalize_pixel(Ra, Rb, second, b, S, rx, ry, Ra.alpha[po], Ra.rgb[co], Ra.rgb[co + (long)S * S], G,
                       mask_occlusions, distance_thresh, &fx, &fy, &mu);
    const long o = ((long)b * H + ry) * W + rx;
    *reinterpret_cast<float2 *>((second ? flow21 : flow12) + o * 2) = make_float2(fx, fy);
    (second ? mult2 : mult1)[o] = mu;
}

/*
 * hoc_flow_finalize + the training half of hoc_warp_photo_forward_pair in one pass (frame-pair path without the
 * visualisation returns): the pixel that has just produced its flow vector is the pixel whose warp sample that flow
 * drives (pair_consist warps at p with flow(p), imgflowarp.py:80-101), so the flow never makes a round trip through
 * memory before the warp and the two dense passes become one.  Four consecutive pixels per thread: alpha, the two
 * colour planes, flows, mult and the masks move as 16-byte (8- / 4-byte for the byte masks) accesses; the occlusion
 * gathers and the bilinear taps run only on the few per cent of pixels a mesh covers.
 * blockIdx.z = 0: flow12 (render 1) -> pair_consist direction 1 (warp image against image_ref, jitter_mask_ref);
 * blockIdx.z = 1: flow21 (render 2) -> direction 0 (warp image_ref against image, jitter_mask).
 */
struct HocFinWarpDir {
    float *flow, *mult;
    const float *src, *target, *jitter;
    uint8_t *valid_mask, *flow_mask;
    double *sums;
};

#ifndef FW_THREADS
#define FW_THREADS 128 /* (128 vs 256 vs 64: 20.6 / 21.3 / 24.6 us at 16 pairs of 256 x 256) */
#endif
#ifndef FW_MINB
#define FW_MINB 8 /* 64 registers, no spill: 16.5 -> 16.0 us against the compiler's own choice of 72 */
#endif
#define FW_BOUNDS __launch_bounds__(FW_THREADS, FW_MINB)
__global__ void FW_BOUNDS
hoc_flow_finalize_warp_kernel(HocRender R1, HocRender R2, HocFinWarpDir D0, HocFinWarpDir D1, int S, int H, int W,
                              const int *__restrict__ ignore, int n_ignore, float distance_thresh, float inv_w,
                              float inv_h, float thresh, int sparse_outputs)
{
    hoc_pdl_sync(); /* programmatic dependent launch: see hoc_common.cuh */
    /* Two phases per CTA (4 pixels per thread).  A: every thread streams its four pixels -- 16-byte loads of alpha and the two
     * colour planes, 16-byte stores of the (zero) outputs -- and notes the pixels a mesh covers in a shared list.
     * B: the listed pixels (a few per cent, clustered in a few CTAs) are dealt ONE PER THREAD: each carries a chain of
     * dependent gathers (ignore table, two occlusion look-ups, 24 bilinear taps), and a thread that ran its own four
     * covered pixels one after the other would be the tail of the whole launch. */
    __shared__ unsigned short s_list[FW_THREADS * 4];
    __shared__ int s_n, s_ig_lo, s_ig_hi;
    const int b = blockIdx.x >> 1; /* (sample, direction) fastest, chunks of rows from the image centre outwards */
    const bool second = (blockIdx.x & 1) != 0;
    const int chunk = hoc_centre_out(blockIdx.y, gridDim.y);
    const HocRender &Ra = second ? R2 : R1;
    const HocRender &Rb = second ? R1 : R2;
    const HocFinWarpDir &D = second ? D1 : D0;
    const int W4 = W >> 2, npix = H * W;
    const int q = chunk * FW_THREADS + threadIdx.x;
    if (threadIdx.x == 0) {
        s_n = 0;
        s_ig_lo = 0x7fffffff;
        s_ig_hi = -0x7fffffff - 1;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n_ignore; k += FW_THREADS) { /* range of the ignored faces, for phase B */
        const int f = ignore[k];
        atomicMin(&s_ig_lo, f);
        atomicMax(&s_ig_hi, f);
    }
    if (q < H * W4) {
        int xq;
        const int ry = hoc_div_small(q, W4, &xq), x0 = xq << 2;
        const long po = ((long)b * S + ry) * S + x0, co = (((long)b * 3) * S + ry) * S + x0;
        const float4 a4 = *reinterpret_cast<const float4 *>(Ra.alpha + po);
        const float4 r4 = *reinterpret_cast<const float4 *>(Ra.rgb + co);
        const float4 g4 = *reinterpret_cast<const float4 *>(Ra.rgb + co + (long)S * S);
        const float al[4] = {a4.x, a4.y, a4.z, a4.w}, c0[4] = {r4.x, r4.y, r4.z, r4.w}, c1[4] = {g4.x, g4.y, g4.z, g4.w};
        const long o = (long)b * npix + (long)ry * W + x0;
        const float4 z4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (!sparse_outputs) { /* (sparse: flow / mult are only defined where valid_mask is set -- all the backward reads) */
            *reinterpret_cast<float4 *>(D.flow + o * 2) = z4;
            *reinterpret_cast<float4 *>(D.flow + o * 2 + 4) = z4;
            *reinterpret_cast<float4 *>(D.mult + o) = z4;
        }
        *reinterpret_cast<unsigned *>(D.valid_mask + o) = 0u;
        if (D.flow_mask != nullptr)
            *reinterpret_cast<uint2 *>(D.flow_mask + o * 2) = make_uint2(0u, 0u);
#pragma unroll
        for (int j = 0; j < 4; j++) /* (alpha 0 and finite colour: every output of the pixel is zero) */
            if (al[j] != 0.0f || !(c0[j] == c0[j]) || !(c1[j] == c1[j]))
                s_list[atomicAdd(&s_n, 1)] = (unsigned short)(threadIdx.x * 4 + j);
    }
    __syncthreads(); /* orders phase A's zero stores before phase B's stores to the same addresses */
    const int n = s_n;
    if (n == 0)
        return;
    /* The listed pixels reach the threads in the order of the shared-memory atomics, which changes from run to run:
     * a float sum per thread would make the loss depend on that order.  Every pixel's term is therefore rounded to a
     * multiple of 2^-28 FIRST and the terms are added as integers (associative): same bits every run. */
    unsigned long long my_fix = 0ull;
    int my_cnt = 0;
    HocIgnore G;
    G.list = ignore; G.n = n_ignore; G.lo = s_ig_lo; G.hi = s_ig_hi;
    for (int i = threadIdx.x; i < n; i += FW_THREADS) {
        const int loc = s_list[i];
        const int qq = chunk * FW_THREADS + (loc >> 2);
        int xq;
        const int ry = hoc_div_small(qq, W4, &xq), rx = (xq << 2) + (loc & 3);
        const long po = ((long)b * S + ry) * S + rx, co = (((long)b * 3) * S + ry) * S + rx;
        float fx, fy, mu;
        hoc_finalize_pixel(Ra, Rb, second, b, S, rx, ry, Ra.alpha[po], Ra.rgb[co], Ra.rgb[co + (long)S * S], G, 1,
                           distance_thresh, &fx, &fy, &mu);
        const long pix = (long)ry * W + rx;
        const long o = (long)b * npix + pix;
        *reinterpret_cast<float2 *>(D.flow + o * 2) = make_float2(fx, fy);
        D.mult[o] = mu;
        if (D.flow_mask != nullptr) {
            uchar2 fm;
            fm.x = !(fx == 0.0f) ? 1 : 0;
            fm.y = !(fy == 0.0f) ? 1 : 0;
            *reinterpret_cast<uchar2 *>(D.flow_mask + o * 2) = fm;
        }
        if (fx == 0.0f)
            continue; /* valid = ... & (flow_x != 0): nothing of this pixel reaches the loss (NaN flows go on) */
        const float *sb = D.src + (size_t)b * 3 * npix;
        const float *jb = (D.jitter != nullptr) ? D.jitter + (size_t)b * 3 * npix : nullptr;
        const float *tb = D.target + (size_t)b * 3 * npix + pix;
        const float tv[3] = {__ldg(tb), __ldg(tb + npix), __ldg(tb + 2 * (size_t)npix)};
        const float jc = (jb != nullptr) ? __ldg(jb + pix) : 1.0f;
        float v[3], d[3], wm[3], sd;
        if (hoc_pair_pixel<false>(sb, jb, tv, jc, rx, ry, fx, fy, H, W, npix, inv_w, inv_h, thresh, v, d, wm, &sd)) {
            D.valid_mask[o] = 1;
            my_fix += (unsigned long long)__double2ll_rn((double)sd * WP_SUM_SCALE); /* sd >= 0 */
            my_cnt += 3;
        }
    }
    /* per-sample (sum, count) */
    if (!__syncthreads_or(my_cnt > 0))
        return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_fix += __shfl_xor_sync(HOC_FULL_MASK, my_fix, o);
        my_cnt += __shfl_xor_sync(HOC_FULL_MASK, my_cnt, o);
    }
    __shared__ unsigned long long s_fix[FW_THREADS / 32];
    __shared__ int s_cn[FW_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        s_fix[warp] = my_fix;
        s_cn[warp] = my_cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long a = 0ull;
        int nn = 0;
        for (int w = 0; w < FW_THREADS / 32; w++) {
            a += s_fix[w];
            nn += s_cn[w];
        }
        if (nn > 0) { /* integer-valued doubles: exact below 2^53, so the atomics commute */
            atomicAdd(&D.sums[2 * b + 0], (double)a);
            atomicAdd(&D.sums[2 * b + 1], (double)nn);
        }
    }
}

/* grad_rgb [B,3,S,S] (image layout) = grad_flow [B,H,W,2] * mult inside the crop, 0 elsewhere / channel 2 */
__global__ void __launch_bounds__(FP_THREADS)
hoc_flow_finalize_backward_kernel(const float *__restrict__ grad_flow, const float *__restrict__ mult, int S, int H,
                                  int W, float *__restrict__ grad_rgb, const float *__restrict__ grad_flow_b,
                                  const float *__restrict__ mult_b, float *__restrict__ grad_rgb_b)
{
    if (blockIdx.z == 1) { /* the second direction of a pair, same launch */
        grad_flow = grad_flow_b;
        mult = mult_b;
        grad_rgb = grad_rgb_b;
    }
    const int b = blockIdx.y;
    const long pix = (long)blockIdx.x * FP_THREADS + threadIdx.x;
    const long npix = (long)S * S;
    if (pix >= npix)
        return;
    const int y = (int)((unsigned)pix / (unsigned)S), x = (int)pix - y * S;
    float gx = 0.0f, gy = 0.0f;
    if (y < H && x < W) {
        const long o = ((long)b * H + y) * W + x;
        const float2 g = *reinterpret_cast<const float2 *>(grad_flow + o * 2);
        const float m = mult[o];
        gx = g.x * m;
        gy = g.y * m;
    }
    float *dst = grad_rgb + (long)b * 3 * npix + pix;
    dst[0] = gx;
    dst[npix] = gy;
    dst[2 * npix] = 0.0f;
}


/* ------------------------------------------------------------------------------------------ */
/* Per-vertex front end of get_opticalflow for one frame pair (opticalflow.py:98-102,121-122 and
 * nr.projection as called from renderer.py:187): pixel locations of both frames (batch_proj2d),
 * their displacements as vertex attributes [dx, dy, 1], and the NDC coordinates of both meshes.
 * Camera tensors may be shared by the batch (leading dimension 1): *_bs is their batch stride. */
struct HocCam {
    const float *K1, *K2, *R, *t, *dist;
    int K1_bs, K2_bs, R_bs, t_bs, dist_bs; /* 0 when broadcast over the batch */
    float orig_size;
};

__device__ __forceinline__ void hoc_proj2d(const float *K, const float *v, float *u, float *w, float *hz)
{
    /* accumulate like a GEMM k-loop (the reference's bmm): acc = fma(a_k, b_k, acc) */
    const float h0 = fmaf(K[2], v[2], fmaf(K[1], v[1], K[0] * v[0]));
    const float h1 = fmaf(K[5], v[2], fmaf(K[4], v[1], K[3] * v[0]));
    const float h2 = fmaf(K[8], v[2], fmaf(K[7], v[1], K[6] * v[0]));
    *u = h0 / h2;
    *w = h1 / h2;
    *hz = h2;
}

/* nr.projection of one vertex; also returns the intermediates the backward needs. */
struct HocProj {
    float xc, yc, zc;   /* camera space after R, t */
    float x_, y_;       /* divided by z + eps */
    float ndc[3];
};
__device__ __forceinline__ void hoc_ndc_project(const float *K, const float *R, const float *t, const float *d,
                                                float orig, const float *v, HocProj *P)
{
    const float xc = fmaf(v[2], R[2], fmaf(v[1], R[1], v[0] * R[0])) + t[0];
    const float yc = fmaf(v[2], R[5], fmaf(v[1], R[4], v[0] * R[3])) + t[1];
    const float zc = fmaf(v[2], R[8], fmaf(v[1], R[7], v[0] * R[6])) + t[2];
    const float x_ = xc / (zc + 1e-9f), y_ = yc / (zc + 1e-9f);
    const float r = sqrtf(x_ * x_ + y_ * y_);
    const float r2 = r * r;
    const float radial = 1.0f + d[0] * r2 + d[1] * (r2 * r2) + d[4] * (r2 * r2 * r2);
    const float x__ = x_ * radial + 2.0f * d[2] * x_ * y_ + d[3] * (r2 + 2.0f * x_ * x_);
    const float y__ = y_ * radial + d[2] * (r2 + 2.0f * y_ * y_) + 2.0f * d[3] * x_ * y_;
    float u = fmaf(1.0f, K[2], fmaf(y__, K[1], x__ * K[0]));
    float w = fmaf(1.0f, K[5], fmaf(y__, K[4], x__ * K[3]));
    w = orig - w;
    u = 2.0f * (u - orig / 2.0f) / orig;
    w = 2.0f * (w - orig / 2.0f) / orig;
    P->xc = xc; P->yc = yc; P->zc = zc; P->x_ = x_; P->y_ = y_;
    P->ndc[0] = u; P->ndc[1] = w; P->ndc[2] = zc;
}

__global__ void __launch_bounds__(FP_THREADS)
hoc_flow_vertices_kernel(const float *__restrict__ verts1, const float *__restrict__ verts2, HocCam C, int V,
                         float *__restrict__ ndc1, float *__restrict__ ndc2, float *__restrict__ attrs12,
                         float *__restrict__ attrs21)
{
    const int b = blockIdx.y;
    const int vi = blockIdx.x * FP_THREADS + threadIdx.x;
    if (vi >= V)
        return;
    const long o = ((long)b * V + vi) * 3;
    const float v1[3] = {verts1[o], verts1[o + 1], verts1[o + 2]};
    const float v2[3] = {verts2[o], verts2[o + 1], verts2[o + 2]};
    const float *K1 = C.K1 + (long)b * C.K1_bs, *K2 = C.K2 + (long)b * C.K2_bs;
    const float *R = C.R + (long)b * C.R_bs, *t = C.t + (long)b * C.t_bs, *d = C.dist + (long)b * C.dist_bs;
    float u1, w1, u2, w2, hz;
    hoc_proj2d(K1, v1, &u1, &w1, &hz);
    hoc_proj2d(K2, v2, &u2, &w2, &hz);
    attrs12[o] = u2 - u1;
    attrs12[o + 1] = w2 - w1;
    attrs12[o + 2] = 1.0f;
    attrs21[o] = u1 - u2;
    attrs21[o + 1] = w1 - w2;
    attrs21[o + 2] = 1.0f;
    HocProj P;
    hoc_ndc_project(K1, R, t, d, C.orig_size, v1, &P);
    ndc1[o] = P.ndc[0]; ndc1[o + 1] = P.ndc[1]; ndc1[o + 2] = P.ndc[2];
    hoc_ndc_project(K2, R, t, d, C.orig_size, v2, &P);
    ndc2[o] = P.ndc[0]; ndc2[o + 1] = P.ndc[1]; ndc2[o + 2] = P.ndc[2];
}

/* adjoint of batch_proj2d: (gu, gw) -> grad v (accumulated) */
__device__ __forceinline__ void hoc_proj2d_bwd(const float *K, const float *v, float gu, float gw, float *gv)
{
    float u, w, h2;
    hoc_proj2d(K, v, &u, &w, &h2);
    const float gh0 = gu / h2, gh1 = gw / h2;
    const float gh2 = -(gu * u + gw * w) / h2;
#pragma unroll
    for (int j = 0; j < 3; j++)
        gv[j] += gh0 * K[j] + gh1 * K[3 + j] + gh2 * K[6 + j];
}

/* adjoint of nr.projection: grad ndc -> grad v (accumulated) */
__device__ __forceinline__ void hoc_ndc_project_bwd(const float *K, const float *R, const float *t, const float *d,
                                                    float orig, const float *v, const float *g, float *gv)
{
    HocProj P;
    hoc_ndc_project(K, R, t, d, orig, v, &P);
    const float x_ = P.x_, y_ = P.y_;
    const float r2 = x_ * x_ + y_ * y_;
    const float radial = 1.0f + d[0] * r2 + d[1] * r2 * r2 + d[4] * r2 * r2 * r2;
    const float dradial = d[0] + 2.0f * d[1] * r2 + 3.0f * d[4] * r2 * r2; /* d radial / d r2 */
    /* u_ndc = 2 (u_px - o/2) / o ; w_ndc = 2 ((o - w_px) - o/2) / o */
    const float gu = g[0] * 2.0f / orig, gw = -g[1] * 2.0f / orig;
    const float gx__ = gu * K[0] + gw * K[3];
    const float gy__ = gu * K[1] + gw * K[4];
    /* x__ = x_ radial + 2 p1 x_ y_ + p2 (r2 + 2 x_^2);  y__ = y_ radial + p1 (r2 + 2 y_^2) + 2 p2 x_ y_ */
    const float p1 = d[2], p2 = d[3];
    const float dxx = radial + x_ * dradial * 2.0f * x_ + 2.0f * p1 * y_ + p2 * (2.0f * x_ + 4.0f * x_);
    const float dxy = x_ * dradial * 2.0f * y_ + 2.0f * p1 * x_ + p2 * 2.0f * y_;
    const float dyx = y_ * dradial * 2.0f * x_ + p1 * 2.0f * x_ + 2.0f * p2 * y_;
    const float dyy = radial + y_ * dradial * 2.0f * y_ + p1 * (2.0f * y_ + 4.0f * y_) + 2.0f * p2 * x_;
    const float gx_ = gx__ * dxx + gy__ * dyx;
    const float gy_ = gx__ * dxy + gy__ * dyy;
    const float zi = 1.0f / (P.zc + 1e-9f);
    const float gxc = gx_ * zi, gyc = gy_ * zi;
    const float gzc = g[2] - (gx_ * x_ + gy_ * y_) * zi;
#pragma unroll
    for (int j = 0; j < 3; j++)
        gv[j] += gxc * R[j] + gyc * R[3 + j] + gzc * R[6 + j];
}

__global__ void __launch_bounds__(FP_THREADS)
hoc_flow_vertices_backward_kernel(const float *__restrict__ verts1, const float *__restrict__ verts2, HocCam C, int V,
                                  const float *__restrict__ g_ndc1, const float *__restrict__ g_ndc2,
                                  const float *__restrict__ g_a12, const float *__restrict__ g_a21,
                                  float *__restrict__ grad_v1, float *__restrict__ grad_v2)
{
    const int b = blockIdx.y;
    const int vi = blockIdx.x * FP_THREADS + threadIdx.x;
    if (vi >= V)
        return;
    const long o = ((long)b * V + vi) * 3;
    const float v1[3] = {verts1[o], verts1[o + 1], verts1[o + 2]};
    const float v2[3] = {verts2[o], verts2[o + 1], verts2[o + 2]};
    const float *K1 = C.K1 + (long)b * C.K1_bs, *K2 = C.K2 + (long)b * C.K2_bs;
    const float *R = C.R + (long)b * C.R_bs, *t = C.t + (long)b * C.t_bs, *d = C.dist + (long)b * C.dist_bs;
    /* attrs12 = loc2 - loc1, attrs21 = loc1 - loc2 */
    float gu1 = 0.f, gw1 = 0.f;
    if (g_a12 != nullptr) {
        gu1 -= g_a12[o];
        gw1 -= g_a12[o + 1];
    }
    if (g_a21 != nullptr) {
        gu1 += g_a21[o];
        gw1 += g_a21[o + 1];
    }
    float gv1[3] = {0.f, 0.f, 0.f}, gv2[3] = {0.f, 0.f, 0.f};
    if (grad_v1 != nullptr) {
        hoc_proj2d_bwd(K1, v1, gu1, gw1, gv1);
        if (g_ndc1 != nullptr) {
            const float g[3] = {g_ndc1[o], g_ndc1[o + 1], g_ndc1[o + 2]};
            hoc_ndc_project_bwd(K1, R, t, d, C.orig_size, v1, g, gv1);
        }
        grad_v1[o] = gv1[0]; grad_v1[o + 1] = gv1[1]; grad_v1[o + 2] = gv1[2];
    }
    if (grad_v2 != nullptr) {
        hoc_proj2d_bwd(K2, v2, -gu1, -gw1, gv2);
        if (g_ndc2 != nullptr) {
            const float g[3] = {g_ndc2[o], g_ndc2[o + 1], g_ndc2[o + 2]};
            hoc_ndc_project_bwd(K2, R, t, d, C.orig_size, v2, g, gv2);
        }
        grad_v2[o] = gv2[0]; grad_v2[o + 1] = gv2[1]; grad_v2[o + 2] = gv2[2];
    }
}

/* ------------------------------------------------------------------------------------------ */
/* batch_cat_meshes for a hand + object pair, both frames of a pair in one launch: verts = cat(hand, obj) along the
 * vertex axis, faces = cat(hand_faces, obj_faces + Vh) along the face axis (the hand's table may be shared by the
 * whole batch). */
__global__ void __launch_bounds__(FP_THREADS)
hoc_cat_meshes_kernel(const float *__restrict__ hand_a, const float *__restrict__ obj_a,
                      const float *__restrict__ hand_b, const float *__restrict__ obj_b,
                      const long long *__restrict__ hand_faces, int hand_faces_batched,
                      const long long *__restrict__ obj_faces, int B, int Vh, int Vo, int Fh, int Fo,
                      float *__restrict__ verts_a, float *__restrict__ verts_b, long long *__restrict__ faces)
{
    const long nv = (long)B * (Vh + Vo) * 3, nf = (faces != nullptr) ? (long)B * (Fh + Fo) * 3 : 0;
    const long total = nv * (verts_b != nullptr ? 2 : 1) + nf;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        if (i >= total - nf) { /* the face table */
            const long j = i - (total - nf);
            const int per = (Fh + Fo) * 3;
            const int b = (int)(j / per), r = (int)(j - (long)b * per);
            faces[j] = (r < Fh * 3) ? hand_faces[(hand_faces_batched ? (long)b * Fh * 3 : 0) + r]
                                    : obj_faces[(long)b * Fo * 3 + (r - Fh * 3)] + Vh;
        } else {
            const bool second = i >= nv;
            const long j = second ? i - nv : i;
            const int per = (Vh + Vo) * 3;
            const int b = (int)(j / per), r = (int)(j - (long)b * per);
            const float *hand = second ? hand_b : hand_a, *obj = second ? obj_b : obj_a;
            (second ? verts_b : verts_a)[j] = (r < Vh * 3) ? hand[(long)b * Vh * 3 + r] : obj[(long)b * Vo * 3 + (r - Vh * 3)];
        }
    }
}

extern "C" int hoc_cat_meshes(const float *hand_a, const float *obj_a, const float *hand_b, const float *obj_b,
                              const long long *hand_faces, int hand_faces_batched, const long long *obj_faces, int B,
                              int Vh, int Vo, int Fh, int Fo, float *verts_a, float *verts_b, long long *faces,
                              void *stream)
{
    HOC_CHECK_ARG(B >= 0 && Vh >= 0 && Vo >= 0 && Fh >= 0 && Fo >= 0, "hoc_cat_meshes: bad shape");
    if (B == 0)
        return HOC_OK;
    HOC_CHECK_ARG(hand_a && obj_a && verts_a, "hoc_cat_meshes: NULL vertices");
    HOC_CHECK_ARG((verts_b == nullptr) || (hand_b && obj_b), "hoc_cat_meshes: second frame requested without inputs");
    HOC_CHECK_ARG((faces == nullptr) || (hand_faces && obj_faces), "hoc_cat_meshes: faces requested without inputs");
    const long total = (long)B * (Vh + Vo) * 3 * (verts_b ? 2 : 1) + (faces ? (long)B * (Fh + Fo) * 3 : 0);
    const unsigned grid = (unsigned)((total + FP_THREADS - 1) / FP_THREADS < 1184 ? (total + FP_THREADS - 1) / FP_THREADS
                                                                                   : 1184);
    HOC_LAUNCH(HOC_K_CAT_MESHES, (cudaStream_t)stream,
               (hoc_cat_meshes_kernel<<<grid, FP_THREADS, 0, (cudaStream_t)stream>>>(
                   hand_a, obj_a, hand_b, obj_b, hand_faces, hand_faces_batched, obj_faces, B, Vh, Vo, Fh, Fo, verts_a,
                   verts_b, faces)));
    HOC_CHECK_LAUNCH("hoc_cat_meshes_kernel");
    return HOC_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* Frame-pair front end in ONE launch: what warpbranch.forward + get_opticalflow + Renderer.render do between the
 * network's vertices and the rasterizer (warpbranch.py:50-52 batch_cat_meshes; opticalflow.py:98-103,121-123
 * batch_proj2d x2, displacements, batch_vertex_textures; renderer.py:250-252,282 fill_back, nr.projection,
 * vertices_to_faces), for BOTH renders of the pair, plus the 0xff fill of the z-buffer keys.  Face-parallel: a thread
 * owns one face of one sample, projects its three vertices in both frames (the per-vertex arithmetic is a few dozen
 * flops, recomputing it per corner is cheaper than a round trip of four [B,V,3] arrays through L2 and two more
 * launches) and emits the face's records of both renders, both windings:
 *   faces [2B,F',3,3]  rows 0..B-1: mesh 1 in NDC (render 1),  rows B..2B-1: mesh 2 (render 2)
 *   tex   [2B,F',3,3]  vertex values [dx, dy, 1]: locs2 - locs1 for render 1, locs1 - locs2 for render 2
 * Same device functions as hoc_flow_vertices / hoc_mesh_gather: bit-identical records.  Also writes the concatenated
 * face table [2B,F,3] (hand first, object indices offset by Vh; rows B..2B-1 repeat rows 0..B-1) that the adjoint
 * scatter of the stacked batch walks. */
#define PF_THREADS 128
#ifdef PF_MINB /* (occupancy experiments: minimum resident CTAs per SM) */
#define PF_BOUNDS __launch_bounds__(PF_THREADS, PF_MINB)
#else
#define PF_BOUNDS __launch_bounds__(PF_THREADS)
#endif
__global__ void PF_BOUNDS
hoc_pair_front_kernel(const float *__restrict__ hand1, const float *__restrict__ obj1, const float *__restrict__ hand2,
                      const float *__restrict__ obj2, const long long *__restrict__ hand_faces, int hand_faces_batched,
                      const long long *__restrict__ obj_faces, HocCam C, int B, int Vh, int Vo, int Fh, int Fo,
                      int fill_back, float *__restrict__ faces_out, float *__restrict__ tex_out,
                      long long *__restrict__ face_table, uint4 *__restrict__ clear, long n_clear,
                      uint4 *__restrict__ zero, long n_zero, int *__restrict__ row_lo, int S, int crop_h,
                      int geom_window)
{
    hoc_pdl_sync(); /* programmatic dependent launch: see hoc_common.cuh */
    {
        const long nthreads = (long)gridDim.x * gridDim.y * PF_THREADS;
        const long t0 = ((long)blockIdx.y * gridDim.x + blockIdx.x) * PF_THREADS + threadIdx.x;
        hoc_fill16(clear, n_clear, t0, nthreads, 0xffffffffu); /* z-buffer keys of the forward that follows */
        hoc_fill16(zero, n_zero, t0, nthreads, 0u);           /* small accumulators of later kernels (the loss sums) */
    }
    if (blockIdx.x == gridDim.x - 1 && row_lo != nullptr) {
        /* The last CTA of every sample computes the pair's RASTER ROW WINDOW (the reference rasterises the
         * square that contains the frame and crops afterwards; 44 % of the pixels of a 480 x 270 frame are thrown
         * away).  Rows yi < row_lo (raster rows count from the bottom; the crop keeps the top crop_h rows) are skipped
         * by every pixel pass.  The window is EXACT, not a heuristic: the only readers outside the crop are
         *   (a) the forward-backward occlusion check, which looks one flow vector away and then another one
         *       (imgflowarp.py:118-146).  A rendered flow value is a convex combination of the vertex displacements of
         *       its face, so |flow_y| <= D = max_v |dy_v|; the nearest-sample position of warp() is
         *       (y + f) S / (S - 1) - 0.5 rounded, at most |f| S / (S - 1) + 1 rows away: two hops reach
         *       2 (D S / (S - 1) + 1) rows below the crop;
         *   (b) the pseudo-gradient (only when the geometry gradient is wanted), which needs every covered pixel of the
         *       mesh: the window then starts below the lowest vertex of both meshes. */
        __shared__ float s_red[2][PF_THREADS / 32];
        const int b = blockIdx.y, V = Vh + Vo;
        const float *K1 = C.K1 + (long)b * C.K1_bs, *K2 = C.K2 + (long)b * C.K2_bs;
        const float *R = C.R + (long)b * C.R_bs, *t = C.t + (long)b * C.t_bs, *d = C.dist + (long)b * C.dist_bs;
        float dmax = 0.0f, ymin = 3.0e38f;
        for (int i = threadIdx.x; i < V; i += PF_THREADS) {
            const float *p1 = (i < Vh) ? hand1 + ((long)b * Vh + i) * 3 : obj1 + ((long)b * Vo + (i - Vh)) * 3;
            const float *p2 = (i < Vh) ? hand2 + ((long)b * Vh + i) * 3 : obj2 + ((long)b * Vo + (i - Vh)) * 3;
            const float v1[3] = {p1[0], p1[1], p1[2]}, v2[3] = {p2[0], p2[1], p2[2]};
            float u1, w1, u2, w2, hz;
            hoc_proj2d(K1, v1, &u1, &w1, &hz);
            hoc_proj2d(K2, v2, &u2, &w2, &hz);
            dmax = fmaxf(dmax, fabsf(w2 - w1)); /* (fmaxf / fminf drop NaNs) */
            if (geom_window) {
                HocProj P;
                hoc_ndc_project(K1, R, t, d, C.orig_size, v1, &P);
                ymin = fminf(ymin, 0.5f * (P.ndc[1] * (float)S + (float)S - 1.0f));
                hoc_ndc_project(K2, R, t, d, C.orig_size, v2, &P);
                ymin = fminf(ymin, 0.5f * (P.ndc[1] * (float)S + (float)S - 1.0f));
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            dmax = fmaxf(dmax, __shfl_xor_sync(HOC_FULL_MASK, dmax, o));
            ymin = fminf(ymin, __shfl_xor_sync(HOC_FULL_MASK, ymin, o));
        }
        if ((threadIdx.x & 31) == 0) {
            s_red[0][threadIdx.x >> 5] = dmax;
            s_red[1][threadIdx.x >> 5] = ymin;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < PF_THREADS / 32; w++) {
                dmax = fmaxf(dmax, s_red[0][w]);
                ymin = fminf(ymin, s_red[1][w]);
            }
            const float hop = ceilf(dmax * (float)S / (float)max(S - 1, 1)) + 1.0f;
            float lo = (float)(S - crop_h) - 2.0f * hop - 2.0f; /* (two rows of slack) */
            if (geom_window)
                lo = fminf(lo, floorf(ymin) - 2.0f);
            const int r = (lo > 0.0f && lo == lo) ? min((int)lo, S - crop_h) : 0; /* huge / NaN displacement: no window */
            row_lo[b] = max(r, 0);
            row_lo[B + b] = max(r, 0);
        }
        return;
    }
    /* staged records of this CTA's faces: [kind][thread][9], kind = faces1, tex1, faces2, tex2 */
    /* records of (faces 1, textures 1, faces 2, textures 2), in stored winding [0..3] and reversed winding [4..7]: every
     * output run is then a plain copy with 16-byte stores (a shared copy loop that un-permuted the reversed records
     * with a division by 9 per float was 45 % of this kernel's instructions) */
    __shared__ __align__(16) float s_rec[8][PF_THREADS * 9];
    const int F = Fh + Fo, V = Vh + Vo;
    const int Fout = fill_back ? 2 * F : F;
    const int f0 = blockIdx.x * PF_THREADS;
    const int f = f0 + threadIdx.x;
    const int b = blockIdx.y;
    if (f0 >= F)
        return; /* (the window CTA when no window is wanted) */
    if (f < F) {
        long long iv[3];
        if (f < Fh) {
            const long long *fi = hand_faces + (hand_faces_batched ? (long)b * Fh * 3 : 0) + (long)f * 3;
            iv[0] = fi[0]; iv[1] = fi[1]; iv[2] = fi[2];
        } else {
            const long long *fi = obj_faces + ((long)b * Fo + (f - Fh)) * 3;
            iv[0] = fi[0] + Vh; iv[1] = fi[1] + Vh; iv[2] = fi[2] + Vh;
        }
        if (face_table != nullptr) { /* [2B,F,3]: the same table for both renders of sample b */
            long long *ft = face_table + ((long)b * F + f) * 3;
            ft[0] = iv[0]; ft[1] = iv[1]; ft[2] = iv[2];
            ft += (long)B * F * 3;
            ft[0] = iv[0]; ft[1] = iv[1]; ft[2] = iv[2];
        }
        const float *K1 = C.K1 + (long)b * C.K1_bs, *K2 = C.K2 + (long)b * C.K2_bs;
        const float *R = C.R + (long)b * C.R_bs, *t = C.t + (long)b * C.t_bs, *d = C.dist + (long)b * C.dist_bs;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const long long i = (iv[k] >= 0 && iv[k] < V) ? iv[k] : 0; /* precondition 0 <= index < V */
            const float *p1 = (i < Vh) ? hand1 + ((long)b * Vh + i) * 3 : obj1 + ((long)b * Vo + (i - Vh)) * 3;
            const float *p2 = (i < Vh) ? hand2 + ((long)b * Vh + i) * 3 : obj2 + ((long)b * Vo + (i - Vh)) * 3;
            const float v1[3] = {p1[0], p1[1], p1[2]}, v2[3] = {p2[0], p2[1], p2[2]};
            float u1, w1, u2, w2, hz;
            hoc_proj2d(K1, v1, &u1, &w1, &hz);
            hoc_proj2d(K2, v2, &u2, &w2, &hz);
            HocProj P;
            hoc_ndc_project(K1, R, t, d, C.orig_size, v1, &P);
            const int at = threadIdx.x * 9 + 3 * k, rt = threadIdx.x * 9 + 3 * (2 - k);
            const float t12[3] = {u2 - u1, w2 - w1, 1.0f}, t21[3] = {u1 - u2, w1 - w2, 1.0f};
#pragma unroll
            for (int c = 0; c < 3; c++) {
                s_rec[0][at + c] = s_rec[4][rt + c] = P.ndc[c];
                s_rec[1][at + c] = s_rec[5][rt + c] = t12[c];
                s_rec[3][at + c] = s_rec[7][rt + c] = t21[c];
            }
            hoc_ndc_project(K2, R, t, d, C.orig_size, v2, &P);
#pragma unroll
            for (int c = 0; c < 3; c++)
                s_rec[2][at + c] = s_rec[6][rt + c] = P.ndc[c];
        }
    }
    __syncthreads();
    const int nf = min(PF_THREADS, F - f0);
    if (nf <= 0)
        return;
    const int nfl = nf * 9;
#pragma unroll
    for (int kind = 0; kind < 8; kind++) {
        if (kind >= 4 && !fill_back)
            break;
        float *base = (kind & 1) ? tex_out : faces_out;
        const long row = (kind & 2) ? (long)B + b : (long)b; /* render 2 lives in the second half of the batch */
        /* reversed winding (v2, v1, v0) in the second half of the faces */
        float *dst = base + (row * Fout + (kind >= 4 ? F : 0) + f0) * 9;
        const float *src = s_rec[kind];
        if ((((uintptr_t)dst) & 15) == 0) {
            for (int i = threadIdx.x; i < (nfl >> 2); i += PF_THREADS)
                reinterpret_cast<float4 *>(dst)[i] = reinterpret_cast<const float4 *>(src)[i];
            if ((int)threadIdx.x < (nfl & 3))
                dst[(nfl & ~3) + threadIdx.x] = src[(nfl & ~3) + threadIdx.x];
        } else {
            for (int i = threadIdx.x; i < nfl; i += PF_THREADS)
                dst[i] = src[i];
        }
    }
}

/* Adjoint of the per-vertex part of hoc_pair_front: gradients of the NDC vertices / vertex attributes of both renders
 * (what hoc_mesh_scatter produces from grad_faces / grad_textures of the stacked [2B] batch) -> gradients of the
 * camera-space vertices of frame 1 (and of frame 2 when asked), written as [B,Vh+Vo,3] (hand first). */
__global__ void __launch_bounds__(FP_THREADS)
hoc_pair_back_kernel(const float *__restrict__ hand1, const float *__restrict__ obj1, const float *__restrict__ hand2,
                     const float *__restrict__ obj2, HocCam C, int B, int Vh, int Vo,
                     const float *__restrict__ g_ndc, const float *__restrict__ g_attr, int has_ndc1, int has_ndc2,
                     int has_a12, int has_a21, float *__restrict__ grad_v1, float *__restrict__ grad_v2)
{
    hoc_pdl_sync(); /* programmatic dependent launch: see hoc_common.cuh */
    const int b = blockIdx.y;
    const int vi = blockIdx.x * FP_THREADS + threadIdx.x;
    const int V = Vh + Vo;
    if (vi >= V)
        return;
    const float *p1 = (vi < Vh) ? hand1 + ((long)b * Vh + vi) * 3 : obj1 + ((long)b * Vo + (vi - Vh)) * 3;
    const float *p2 = (vi < Vh) ? hand2 + ((long)b * Vh + vi) * 3 : obj2 + ((long)b * Vo + (vi - Vh)) * 3;
    const float v1[3] = {p1[0], p1[1], p1[2]}, v2[3] = {p2[0], p2[1], p2[2]};
    const float *K1 = C.K1 + (long)b * C.K1_bs, *K2 = C.K2 + (long)b * C.K2_bs;
    const float *R = C.R + (long)b * C.R_bs, *t = C.t + (long)b * C.t_bs, *d = C.dist + (long)b * C.dist_bs;
    const long o1 = ((long)b * V + vi) * 3, o2 = (((long)B + b) * V + vi) * 3;
    /* attrs12 = loc2 - loc1 (render 1, rows 0..B-1), attrs21 = loc1 - loc2 (render 2, rows B..2B-1) */
    float gu1 = 0.f, gw1 = 0.f;
    if (has_a12) {
        gu1 -= g_attr[o1];
        gw1 -= g_attr[o1 + 1];
    }
    if (has_a21) {
        gu1 += g_attr[o2];
        gw1 += g_attr[o2 + 1];
    }
    if (grad_v1 != nullptr) {
        float gv[3] = {0.f, 0.f, 0.f};
        hoc_proj2d_bwd(K1, v1, gu1, gw1, gv);
        if (has_ndc1) {
            const float g[3] = {g_ndc[o1], g_ndc[o1 + 1], g_ndc[o1 + 2]};
            hoc_ndc_project_bwd(K1, R, t, d, C.orig_size, v1, g, gv);
        }
        grad_v1[o1] = gv[0]; grad_v1[o1 + 1] = gv[1]; grad_v1[o1 + 2] = gv[2];
    }
    if (grad_v2 != nullptr) {
        float gv[3] = {0.f, 0.f, 0.f};
        hoc_proj2d_bwd(K2, v2, -gu1, -gw1, gv);
        if (has_ndc2) {
            const float g[3] = {g_ndc[o2], g_ndc[o2 + 1], g_ndc[o2 + 2]};
            hoc_ndc_project_bwd(K2, R, t, d, C.orig_size, v2, g, gv);
        }
        grad_v2[o1] = gv[0]; grad_v2[o1 + 1] = gv[1]; grad_v2[o1 + 2] = gv[2];
    }
}

/* ------------------------------------------------------------------------------------------ */
extern "C" int hoc_mesh_gather_clear(const float *verts, const float *attrs, const long long *faces_idx, int B, int V,
                                     int F, int fill_back, int tex_mode, float *faces_out, float *textures_out,
                                     void *clear, size_t clear_bytes, void *stream);

extern "C" int hoc_mesh_gather(const float *verts, const float *attrs, const long long *faces_idx, int B, int V, int F,
                               int fill_back, float *faces_out, float *textures_out, void *stream)
{
    return hoc_mesh_gather_clear(verts, attrs, faces_idx, B, V, F, fill_back, HOC_TEX_GRAD_CUBE, faces_out, textures_out,
                                 nullptr, 0, stream);
}

extern "C" int hoc_mesh_gather_clear(const float *verts, const float *attrs, const long long *faces_idx, int B, int V,
                                     int F, int fill_back, int tex_mode, float *faces_out, float *textures_out,
                                     void *clear, size_t clear_bytes, void *stream)
{
    HOC_CHECK_ARG(tex_mode == HOC_TEX_GRAD_CUBE || tex_mode == HOC_TEX_GRAD_VERTEX, "hoc_mesh_gather_clear: tex_mode %d",
                  tex_mode);
    HOC_CHECK_ARG(clear == nullptr || (clear_bytes % 16 == 0 && ((uintptr_t)clear & 15) == 0),
                  "hoc_mesh_gather_clear: clear buffer must be 16-byte aligned with a size multiple of 16");
    if (clear != nullptr && (B == 0 || F == 0)) { /* nothing to gather: still honour the fill */
        if (cudaMemsetAsync(clear, 0xff, clear_bytes, (cudaStream_t)stream) != cudaSuccess) {
            hoc_set_error("hoc_mesh_gather_clear: memset failed");
            return HOC_ERR_CUDA;
        }
    }
    HOC_CHECK_ARG(B >= 0 && V >= 0 && F >= 0, "hoc_mesh_gather: bad shape B=%d V=%d F=%d", B, V, F);
    HOC_CHECK_ARG(B <= 65535, "hoc_mesh_gather: batch %d exceeds 65535", B);
    if (B == 0 || F == 0)
        return HOC_OK;
    HOC_CHECK_ARG(verts && faces_idx && faces_out, "hoc_mesh_gather: NULL argument");
    HOC_CHECK_ARG(textures_out == nullptr || attrs != nullptr, "hoc_mesh_gather: textures requested without attributes");
    const int Fo = fill_back ? 2 * F : F;
    dim3 grid((Fo + GA_THREADS - 1) / GA_THREADS, B);
    HOC_LAUNCH(HOC_K_MESH_GATHER, (cudaStream_t)stream,
               (hoc_mesh_gather_kernel<<<grid, GA_THREADS, 0, (cudaStream_t)stream>>>(
                   verts, attrs, faces_idx, V, F, fill_back, faces_out, textures_out,
                   tex_mode == HOC_TEX_GRAD_VERTEX ? 1 : 0, (uint4 *)clear, (long)(clear_bytes / 16))));
    HOC_CHECK_LAUNCH("hoc_mesh_gather_kernel");
    return HOC_OK;
}

extern "C" size_t hoc_mesh_scatter_workspace_bytes(int B, int V)
{
    return (g_hoc_deterministic != 0 && B > 0 && V > 0) ? 2 * 16 * 3 * (size_t)B * V : 0;
}

extern "C" int hoc_mesh_scatter_ws(const float *grad_faces, const float *grad_textures, const long long *faces_idx,
                                   int B, int V, int F, int fill_back, int tex_grad_mode, float *grad_verts,
                                   float *grad_attrs, int outputs_zeroed, void *workspace, size_t workspace_bytes,
                                   void *stream);

extern "C" int hoc_mesh_scatter(const float *grad_faces, const float *grad_textures, const long long *faces_idx, int B,
                                int V, int F, int fill_back, int tex_grad_mode, float *grad_verts, float *grad_attrs,
                                void *stream)
{
    return hoc_mesh_scatter_ws(grad_faces, grad_textures, faces_idx, B, V, F, fill_back, tex_grad_mode, grad_verts,
                               grad_attrs, 0, nullptr, 0, stream);
}

extern "C" int hoc_mesh_scatter_ws(const float *grad_faces, const float *grad_textures, const long long *faces_idx,
                                   int B, int V, int F, int fill_back, int tex_grad_mode, float *grad_verts,
                                   float *grad_attrs, int outputs_zeroed, void *workspace, size_t workspace_bytes,
                                   void *stream)
{
    HOC_CHECK_ARG(B >= 0 && V >= 0 && F >= 0, "hoc_mesh_scatter: bad shape B=%d V=%d F=%d", B, V, F);
    HOC_CHECK_ARG(B <= 65535, "hoc_mesh_scatter: batch %d exceeds 65535", B);
    cudaStream_t st = (cudaStream_t)stream;
    if (B == 0)
        return HOC_OK;
    cudaError_t e = cudaSuccess;
    const size_t nbytes = sizeof(float) * 3 * (size_t)B * V;
    if (outputs_zeroed) {
        /* an earlier kernel of the caller's sequence filled both outputs with zeros (one graph node less) */
    } else if (grad_verts != nullptr && grad_attrs == grad_verts + 3 * (size_t)B * V) {
        e = cudaMemsetAsync(grad_verts, 0, 2 * nbytes, st); /* adjacent outputs: one fill */
    } else {
        if (grad_verts != nullptr)
            e = cudaMemsetAsync(grad_verts, 0, nbytes, st);
        if (e == cudaSuccess && grad_attrs != nullptr)
            e = cudaMemsetAsync(grad_attrs, 0, nbytes, st);
    }
    if (e != cudaSuccess) {
        hoc_set_error("hoc_mesh_scatter: memset failed: %s", cudaGetErrorString(e));
        return HOC_ERR_CUDA;
    }
    if (F == 0)
        return HOC_OK;
    HOC_CHECK_ARG(faces_idx != nullptr, "hoc_mesh_scatter: faces_idx NULL");
    unsigned long long *det_v = nullptr, *det_a = nullptr;
    const size_t need = hoc_mesh_scatter_workspace_bytes(B, V);
    if (need > 0) { /* reproducible mode */
        if (workspace == nullptr || workspace_bytes < need) {
            hoc_set_error("hoc_mesh_scatter: reproducible mode needs a workspace of %zu bytes (hoc_mesh_scatter_ws), %zu given",
                          need, workspace_bytes);
            return HOC_ERR_WORKSPACE;
        }
        if (cudaMemsetAsync(workspace, 0, need, st) != cudaSuccess) {
            hoc_set_error("hoc_mesh_scatter: memset of the accumulators failed");
            return HOC_ERR_CUDA;
        }
        det_v = (unsigned long long *)workspace;
        det_a = det_v + 2 * 3 * (size_t)B * V;
    }
    const int Fo = fill_back ? 2 * F : F;
    dim3 grid((Fo + FP_THREADS - 1) / FP_THREADS, B);
    HOC_LAUNCH(HOC_K_MESH_SCATTER, st,
               (hoc_launch_pdl((hoc_mesh_scatter_kernel), grid, FP_THREADS, 0, st, grad_faces, grad_textures, faces_idx, V, F,
                                                                     fill_back, tex_grad_mode, grad_verts, grad_attrs,
                                                                     det_v, det_a)));
    HOC_CHECK_LAUNCH("hoc_mesh_scatter_kernel");
    if (need > 0) {
        const long n = 3l * B * V;
        if ((grad_verts != nullptr && hoc_det_flush(det_v, n, grad_verts, 0, st) != cudaSuccess) ||
            (grad_attrs != nullptr && hoc_det_flush(det_a, n, grad_attrs, 0, st) != cudaSuccess)) {
            hoc_set_error("hoc_mesh_scatter: flush of the accumulators failed");
            return HOC_ERR_CUDA;
        }
    }
    return HOC_OK;
}

extern "C" int hoc_flow_finalize(const float *rgb1, const float *alpha1, const int32_t *idx1, const float *rgb2,
                                 const float *alpha2, const int32_t *idx2, int B, int S, int H, int W,
                                 const int *ignore_faces, int n_ignore, int mask_occlusions, float distance_thresh,
                                 float *flow12, float *flow21, float *mult1, float *mult2, void *stream)
{
    HOC_CHECK_ARG(B >= 0 && S >= 1 && H >= 1 && W >= 1 && H <= S && W <= S,
                  "hoc_flow_finalize: bad shape B=%d S=%d H=%d W=%d", B, S, H, W);
    HOC_CHECK_ARG(B <= 65535, "hoc_flow_finalize: batch %d exceeds 65535", B);
    HOC_CHECK_ARG(n_ignore >= 0 && n_ignore <= 64, "hoc_flow_finalize: at most 64 ignored faces (got %d)", n_ignore);
    if (B == 0)
        return HOC_OK;
    HOC_CHECK_ARG(rgb1 && alpha1 && idx1 && rgb2 && alpha2 && idx2 && flow12 && flow21 && mult1 && mult2,
                  "hoc_flow_finalize: NULL argument");
    HOC_CHECK_ARG(n_ignore == 0 || ignore_faces != nullptr, "hoc_flow_finalize: ignore_faces NULL");
    HocRender R1 = {rgb1, alpha1, idx1}, R2 = {rgb2, alpha2, idx2};
    const long npix = (long)H * W;
    dim3 grid((unsigned)((npix + FP_THREADS - 1) / FP_THREADS), B, 2);
    HOC_LAUNCH(HOC_K_FLOW_FINALIZE, (cudaStream_t)stream,
               (hoc_flow_finalize_kernel<<<grid, FP_THREADS, 0, (cudaStream_t)stream>>>(
                   R1, R2, S, H, W, ignore_faces, n_ignore, mask_occlusions, distance_thresh, flow12, flow21, mult1,
                   mult2)));
    HOC_CHECK_LAUNCH("hoc_flow_finalize_kernel");
    return HOC_OK;
}

extern "C" int hoc_flow_finalize_backward(const float *grad_flow, const float *mult, int B, int S, int H, int W,
                                          float *grad_rgb, void *stream)
{
    HOC_CHECK_ARG(B >= 0 && S >= 1 && H >= 1 && W >= 1 && H <= S && W <= S,
                  "hoc_flow_finalize_backward: bad shape B=%d S=%d H=%d W=%d", B, S, H, W);
    HOC_CHECK_ARG(B <= 65535, "hoc_flow_finalize_backward: batch %d exceeds 65535", B);
    if (B == 0)
        return HOC_OK;
    HOC_CHECK_ARG(grad_flow && mult && grad_rgb, "hoc_flow_finalize_backward: NULL argument");
    const long npix = (long)S * S;
    dim3 grid((unsigned)((npix + FP_THREADS - 1) / FP_THREADS), B);
    HOC_LAUNCH(HOC_K_FLOW_FINALIZE_BWD, (cudaStream_t)stream,
               (hoc_flow_finalize_backward_kernel<<<grid, FP_THREADS, 0, (cudaStream_t)stream>>>(
                   grad_flow, mult, S, H, W, grad_rgb, nullptr, nullptr, nullptr)));
    HOC_CHECK_LAUNCH("hoc_flow_finalize_backward_kernel");
    return HOC_OK;
}

extern "C" int hoc_flow_finalize_backward_pair(const float *grad_flow12, const float *mult1, const float *grad_flow21,
                                               const float *mult2, int B, int S, int H, int W, float *grad_rgb1,
                                               float *grad_rgb2, void *stream)
{
    HOC_CHECK_ARG(B >= 0 && S >= 1 && H >= 1 && W >= 1 && H <= S && W <= S,
                  "hoc_flow_finalize_backward_pair: bad shape B=%d S=%d H=%d W=%d", B, S, H, W);
    HOC_CHECK_ARG(B <= 65535, "hoc_flow_finalize_backward_pair: batch %d exceeds 65535", B);
    if (B == 0)
        return HOC_OK;
    HOC_CHECK_ARG(grad_flow12 && mult1 && grad_rgb1 && grad_flow21 && mult2 && grad_rgb2,
                  "hoc_flow_finalize_backward_pair: NULL argument");
    const long npix = (long)S * S;
    dim3 grid((unsigned)((npix + FP_THREADS - 1) / FP_THREADS), B, 2);
    HOC_LAUNCH(HOC_K_FLOW_FINALIZE_BWD, (cudaStream_t)stream,
               (hoc_flow_finalize_backward_kernel<<<grid, FP_THREADS, 0, (cudaStream_t)stream>>>(
                   grad_flow12, mult1, S, H, W, grad_rgb1, grad_flow21, mult2, grad_rgb2)));
    HOC_CHECK_LAUNCH("hoc_flow_finalize_backward_kernel");
    return HOC_OK;
}

static HocCam hoc_make_cam(const float *K1, int K1_batched, const float *K2, int K2_batched, const float *R,
                           int R_batched, const float *t, int t_batched, const float *dist, int dist_batched,
                           float orig_size)
{
    HocCam C;
    C.K1 = K1; C.K2 = K2; C.R = R; C.t = t; C.dist = dist;
    C.K1_bs = K1_batched ? 9 : 0;
    C.K2_bs = K2_batched ? 9 : 0;
    C.R_bs = R_batched ? 9 : 0;
    C.t_bs = t_batched ? 3 : 0;
    C.dist_bs = dist_batched ? 5 : 0;
    C.orig_size = orig_size;
    return C;
}

extern "C" int hoc_flow_vertices(const float *verts1, const float *verts2, const float *K1, int K1_batched,
                                 const float *K2, int K2_batched, const float *R, int R_batched, const float *t,
                                 int t_batched, const float *dist_coeffs, int dist_batched, float orig_size, int B,
                                 int V, float *ndc1, float *ndc2, float *attrs12, float *attrs21, void *stream)
{
    HOC_CHECK_ARG(B >= 0 && V >= 0 && B <= 65535, "hoc_flow_vertices: bad shape B=%d V=%d", B, V);
    if (B == 0 || V == 0)
        return HOC_OK;
    HOC_CHECK_ARG(verts1 && verts2 && K1 && K2 && R && t && dist_coeffs && ndc1 && ndc2 && attrs12 && attrs21,
                  "hoc_flow_vertices: NULL argument");
    HocCam C = hoc_make_cam(K1, K1_batched, K2, K2_batched, R, R_batched, t, t_batched, dist_coeffs, dist_batched,
                            orig_size);
    dim3 grid((V + FP_THREADS - 1) / FP_THREADS, B);
    HOC_LAUNCH(HOC_K_FLOW_VERTICES, (cudaStream_t)stream,
               (hoc_flow_vertices_kernel<<<grid, FP_THREADS, 0, (cudaStream_t)stream>>>(verts1, verts2, C, V, ndc1,
                                                                                        ndc2, attrs12, attrs21)));
    HOC_CHECK_LAUNCH("hoc_flow_vertices_kernel");
    return HOC_OK;
}

extern "C" int hoc_flow_vertices_backward(const float *verts1, const float *verts2, const float *K1, int K1_batched,
                                          const float *K2, int K2_batched, const float *R, int R_batched,
                                          const float *t, int t_batched, const float *dist_coeffs, int dist_batched,
                                          float orig_size, int B, int V, const float *grad_ndc1,
                                          const float *grad_ndc2, const float *grad_attrs12,
                                          const float *grad_attrs21, float *grad_verts1, float *grad_verts2,
                                          void *stream)
{
    HOC_CHECK_ARG(B >= 0 && V >= 0 && B <= 65535, "hoc_flow_vertices_backward: bad shape B=%d V=%d", B, V);
    if (B == 0 || V == 0 || (grad_verts1 == nullptr && grad_verts2 == nullptr))
        return HOC_OK;
    HOC_CHECK_ARG(verts1 && verts2 && K1 && K2 && R && t && dist_coeffs, "hoc_flow_vertices_backward: NULL argument");
    HocCam C = hoc_make_cam(K1, K1_batched, K2, K2_batched, R, R_batched, t, t_batched, dist_coeffs, dist_batched,
                            orig_size);
    dim3 grid((V + FP_THREADS - 1) / FP_THREADS, B);
    HOC_LAUNCH(HOC_K_FLOW_VERTICES_BWD, (cudaStream_t)stream,
               (hoc_flow_vertices_backward_kernel<<<grid, FP_THREADS, 0, (cudaStream_t)stream>>>(
                   verts1, verts2, C, V, grad_ndc1, grad_ndc2, grad_attrs12, grad_attrs21, grad_verts1, grad_verts2)));
    HOC_CHECK_LAUNCH("hoc_flow_vertices_backward_kernel");
    return HOC_OK;
}

extern "C" int hoc_pair_front(const float *hand1, const float *obj1, const float *hand2, const float *obj2,
                              const long long *hand_faces, int hand_faces_batched, const long long *obj_faces,
                              const float *K1, int K1_batched, const float *K2, int K2_batched, const float *R,
                              int R_batched, const float *t, int t_batched, const float *dist_coeffs, int dist_batched,
                              float orig_size, int B, int Vh, int Vo, int Fh, int Fo, int fill_back, float *faces_out,
                              float *textures_out, long long *face_table, void *clear, size_t clear_bytes, void *zero,
                              size_t zero_bytes, int *row_lo, int S, int crop_h, int geom_window, void *stream)
{
    HOC_CHECK_ARG(row_lo == nullptr || (S >= 1 && crop_h >= 1 && crop_h <= S),
                  "hoc_pair_front: row window needs 1 <= crop_h <= S (got %d, %d)", crop_h, S);
    HOC_CHECK_ARG(B >= 0 && Vh >= 0 && Vo >= 0 && Fh >= 0 && Fo >= 0 && B <= 32767, "hoc_pair_front: bad shape");
    HOC_CHECK_ARG(clear == nullptr || (clear_bytes % 16 == 0 && ((uintptr_t)clear & 15) == 0),
                  "hoc_pair_front: clear buffer must be 16-byte aligned with a size multiple of 16");
    HOC_CHECK_ARG(zero == nullptr || (zero_bytes % 16 == 0 && ((uintptr_t)zero & 15) == 0),
                  "hoc_pair_front: zero buffer must be 16-byte aligned with a size multiple of 16");
    cudaStream_t st = (cudaStream_t)stream;
    if (clear == nullptr)
        clear_bytes = 0;
    if (zero == nullptr)
        zero_bytes = 0;
    if (B == 0 || Fh + Fo == 0) {
        if ((clear_bytes && cudaMemsetAsync(clear, 0xff, clear_bytes, st) != cudaSuccess) ||
            (zero_bytes && cudaMemsetAsync(zero, 0, zero_bytes, st) != cudaSuccess)) {
            hoc_set_error("hoc_pair_front: memset failed");
            return HOC_ERR_CUDA;
        }
        return HOC_OK;
    }
    HOC_CHECK_ARG((Vh == 0 || (hand1 && hand2)) && (Vo == 0 || (obj1 && obj2)), "hoc_pair_front: NULL vertices");
    HOC_CHECK_ARG((Fh == 0 || hand_faces) && (Fo == 0 || obj_faces), "hoc_pair_front: NULL face table");
    HOC_CHECK_ARG(K1 && K2 && R && t && dist_coeffs && faces_out && textures_out, "hoc_pair_front: NULL argument");
    HocCam C = hoc_make_cam(K1, K1_batched, K2, K2_batched, R, R_batched, t, t_batched, dist_coeffs, dist_batched,
                            orig_size);
    dim3 grid((Fh + Fo + PF_THREADS - 1) / PF_THREADS + 1, B); /* + the row-window CTA of every sample */
    HOC_LAUNCH(HOC_K_PAIR_FRONT, st,
               (hoc_launch_pdl((hoc_pair_front_kernel), grid, PF_THREADS, 0, st, hand1, obj1, hand2, obj2, hand_faces,
                                                                   hand_faces_batched, obj_faces, C, B, Vh, Vo, Fh, Fo,
                                                                   fill_back, faces_out, textures_out, face_table,
                                                                   (uint4 *)clear, (long)(clear_bytes / 16),
                                                                   (uint4 *)zero, (long)(zero_bytes / 16), row_lo, S,
                                                                   crop_h, geom_window)));
    HOC_CHECK_LAUNCH("hoc_pair_front_kernel");
    return HOC_OK;
}

extern "C" int hoc_pair_back(const float *hand1, const float *obj1, const float *hand2, const float *obj2,
                             const float *K1, int K1_batched, const float *K2, int K2_batched, const float *R,
                             int R_batched, const float *t, int t_batched, const float *dist_coeffs, int dist_batched,
                             float orig_size, int B, int Vh, int Vo, const float *grad_ndc, const float *grad_attrs,
                             int has_ndc1, int has_ndc2, int has_attrs12, int has_attrs21, float *grad_verts1,
                             float *grad_verts2, void *stream)
{
    HOC_CHECK_ARG(B >= 0 && Vh >= 0 && Vo >= 0 && B <= 32767, "hoc_pair_back: bad shape");
    if (B == 0 || Vh + Vo == 0 || (grad_verts1 == nullptr && grad_verts2 == nullptr))
        return HOC_OK;
    HOC_CHECK_ARG((Vh == 0 || (hand1 && hand2)) && (Vo == 0 || (obj1 && obj2)) && K1 && K2 && R && t && dist_coeffs,
                  "hoc_pair_back: NULL argument");
    HOC_CHECK_ARG((!(has_ndc1 || has_ndc2) || grad_ndc) && (!(has_attrs12 || has_attrs21) || grad_attrs),
                  "hoc_pair_back: gradient flagged but NULL");
    HocCam C = hoc_make_cam(K1, K1_batched, K2, K2_batched, R, R_batched, t, t_batched, dist_coeffs, dist_batched,
                            orig_size);
    dim3 grid((Vh + Vo + FP_THREADS - 1) / FP_THREADS, B);
    HOC_LAUNCH(HOC_K_PAIR_BACK, (cudaStream_t)stream,
               (hoc_launch_pdl((hoc_pair_back_kernel), grid, FP_THREADS, 0, (cudaStream_t)stream, 
                   hand1, obj1, hand2, obj2, C, B, Vh, Vo, grad_ndc, grad_attrs, has_ndc1, has_ndc2, has_attrs12,
                   has_attrs21, grad_verts1, grad_verts2)));
    HOC_CHECK_LAUNCH("hoc_pair_back_kernel");
    return HOC_OK;
}

extern "C" int hoc_flow_finalize_warp_ex(const float *rgb1, const float *alpha1, const int32_t *idx1, const float *rgb2,
                                         const float *alpha2, const int32_t *idx2, const float *image_ref,
                                         const float *image, const float *jitter_ref, const float *jitter, int B, int S,
                                         int H, int W, const int *ignore_faces, int n_ignore, float distance_thresh,
                                         float thresh, float *flow12, float *flow21, float *mult1, float *mult2,
                                         uint8_t *const *valid_mask, uint8_t *const *flow_mask, double *sums,
                                         int sparse_outputs, void *stream)
{
    HOC_CHECK_ARG(B >= 0 && S >= 4 && (S % 4) == 0 && H >= 1 && W >= 4 && (W % 4) == 0 && H <= S && W <= S,
                  "hoc_flow_finalize_warp: bad shape B=%d S=%d H=%d W=%d (S, W multiples of 4)", B, S, H, W);
    HOC_CHECK_ARG(B <= 65535, "hoc_flow_finalize_warp: batch %d exceeds 65535", B);
    HOC_CHECK_ARG(n_ignore >= 0 && n_ignore <= 64, "hoc_flow_finalize_warp: at most 64 ignored faces (got %d)", n_ignore);
    if (B == 0)
        return HOC_OK;
    HOC_CHECK_ARG(rgb1 && alpha1 && idx1 && rgb2 && alpha2 && idx2 && image_ref && image && flow12 && flow21 && mult1 &&
                      mult2 && valid_mask && valid_mask[0] && valid_mask[1] && sums,
                  "hoc_flow_finalize_warp: NULL argument");
    HOC_CHECK_ARG((jitter_ref == nullptr) == (jitter == nullptr), "hoc_flow_finalize_warp: one jitter mask missing");
    HOC_CHECK_ARG(n_ignore == 0 || ignore_faces != nullptr, "hoc_flow_finalize_warp: ignore_faces NULL");
    HocRender R1 = {rgb1, alpha1, idx1}, R2 = {rgb2, alpha2, idx2};
    /* finalize direction 0 produces flow12, which pair_consist's direction 1 consumes (and vice versa) */
    HocFinWarpDir D0 = {flow12, mult1, image, image_ref, jitter_ref, valid_mask[1], flow_mask ? flow_mask[1] : nullptr,
                        sums + 2 * (size_t)B};
    HocFinWarpDir D1 = {flow21, mult2, image_ref, image, jitter, valid_mask[0], flow_mask ? flow_mask[0] : nullptr, sums};
    const uintptr_t al16 = (uintptr_t)rgb1 | (uintptr_t)alpha1 | (uintptr_t)rgb2 | (uintptr_t)alpha2 | (uintptr_t)flow12 |
                           (uintptr_t)flow21 | (uintptr_t)mult1 | (uintptr_t)mult2;
    HOC_CHECK_ARG((al16 & 15) == 0 && (((uintptr_t)valid_mask[0] | (uintptr_t)valid_mask[1]) & 3) == 0 &&
                      (flow_mask == nullptr || (((uintptr_t)flow_mask[0] | (uintptr_t)flow_mask[1]) & 7) == 0),
                  "hoc_flow_finalize_warp: tensors must be 16-byte aligned");
    const long groups = (long)H * (W / 4);
    dim3 grid(2 * B, (unsigned)((groups + FW_THREADS - 1) / FW_THREADS));
    HOC_LAUNCH(HOC_K_FLOW_FINALIZE, (cudaStream_t)stream,
               (hoc_launch_pdl((hoc_flow_finalize_warp_kernel), grid, FW_THREADS, 0, (cudaStream_t)stream, 
                   R1, R2, D0, D1, S, H, W, ignore_faces, n_ignore, distance_thresh,
                   1.0f / (float)(W - 1 > 1 ? W - 1 : 1), 1.0f / (float)(H - 1 > 1 ? H - 1 : 1), thresh,
                   sparse_outputs ? 1 : 0)));
    HOC_CHECK_LAUNCH("hoc_flow_finalize_warp_kernel");
    return HOC_OK;
}

extern "C" int hoc_flow_finalize_warp(const float *rgb1, const float *alpha1, const int32_t *idx1, const float *rgb2,
                                      const float *alpha2, const int32_t *idx2, const float *image_ref,
                                      const float *image, const float *jitter_ref, const float *jitter, int B, int S,
                                      int H, int W, const int *ignore_faces, int n_ignore, float distance_thresh,
                                      float thresh, float *flow12, float *flow21, float *mult1, float *mult2,
                                      uint8_t *const *valid_mask, uint8_t *const *flow_mask, double *sums, void *stream)
{
    return hoc_flow_finalize_warp_ex(rgb1, alpha1, idx1, rgb2, alpha2, idx2, image_ref, image, jitter_ref, jitter, B, S, H,
                                     W, ignore_faces, n_ignore, distance_thresh, thresh, flow12, flow21, mult1, mult2,
                                     valid_mask, flow_mask, sums, 0, stream);
}
