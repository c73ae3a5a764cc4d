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

#define FP_THREADS 256

/* ------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(FP_THREADS)
hoc_mesh_gather_kernel(const float *__restrict__ verts, const float *__restrict__ attrs,
                       const long long *__restrict__ faces_idx, int V, int F, int fill_back,
                       float *__restrict__ faces_out, float *__restrict__ tex_out)
{
    const int Fo = fill_back ? 2 * F : F;
    const int fo = blockIdx.x * FP_THREADS + threadIdx.x;
    const int b = blockIdx.y;
    if (fo >= Fo)
        return;
    const int f = fo >= F ? fo - F : fo;
    const long long *fi = faces_idx + ((long)b * F + f) * 3;
    long long i0 = fi[0], i1 = fi[1], i2 = fi[2];
    if (fo >= F) { /* reversed winding: (v2, v1, v0) */
        const long long t = i0;
        i0 = i2;
        i2 = t;
    }
    const long long iv[3] = {i0, i1, i2};
    float c[3][3];
    float *fd = faces_out + ((long)b * Fo + fo) * 9;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float *vs = verts + ((long)b * V + iv[k]) * 3;
        fd[3 * k + 0] = vs[0];
        fd[3 * k + 1] = vs[1];
        fd[3 * k + 2] = vs[2];
        if (tex_out != nullptr) {
            const float *as = attrs + ((long)b * V + iv[k]) * 3;
            c[k][0] = as[0];
            c[k][1] = as[1];
            c[k][2] = as[2];
        }
    }
    if (tex_out != nullptr) {
        /* cube whose trilinear sample at the barycentric coordinates is b0 c0 + b1 c1 + b2 c2:
         * T[i,j,k] = i c0 + j c1 + k c2 (the reversed copy is the permute(0,1,4,3,2,5) of the original) */
        float *td = tex_out + ((long)b * Fo + fo) * 24;
#pragma unroll
        for (int corner = 0; corner < 8; corner++) {
            const float wi = (float)((corner >> 2) & 1), wj = (float)((corner >> 1) & 1), wk = (float)(corner & 1);
#pragma unroll
            for (int ch = 0; ch < 3; ch++)
                td[corner * 3 + ch] = wi * c[0][ch] + wj * c[1][ch] + wk * c[2][ch];
        }
    }
}

__global__ void __launch_bounds__(FP_THREADS)
hoc_mesh_scatter_kernel(const float *__restrict__ grad_faces, const float *__restrict__ grad_tex,
                        const long long *__restrict__ faces_idx, int V, int F, int fill_back,
                        float *__restrict__ grad_verts, float *__restrict__ grad_attrs)
{
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
    if (grad_tex != nullptr && grad_attrs != nullptr) {
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
        if (any) {
            float *dst = grad_verts + ((long)b * V + iv[k]) * 3;
#pragma unroll
            for (int d = 0; d < 3; d++)
                if (gf[3 * k + d] != 0.0f)
                    atomicAdd(dst + d, gf[3 * k + d]);
        }
        if (any_t) {
            float *dst = grad_attrs + ((long)b * V + iv[k]) * 3;
#pragma unroll
            for (int ch = 0; ch < 3; ch++)
                if (gc[k][ch] != 0.0f)
                    atomicAdd(dst + ch, gc[k][ch]);
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
__device__ __forceinline__ float hoc_fp_mask(const HocRender &R, int b, int S, int x, int y,
                                             const int *__restrict__ ignore, int n_ignore, float *alpha_out)
{
    const long po = ((long)b * S + y) * S + x;
    const float a = R.alpha[po];
    *alpha_out = a;
    float m = (a > 0.99999f) ? 1.0f : 0.0f;
    if (n_ignore > 0) {
        const int fidx = R.idx[((long)b * S + (S - 1 - y)) * S + x];
        bool keep = true;
        for (int k = 0; k < n_ignore; k++)
            keep = keep && (fidx != ignore[k]);
        m = __fmul_rn(m, keep ? 1.0f : 0.0f);
    }
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
    const int ry = (int)(pix / W), rx = (int)(pix - (long)ry * W);
    const HocRender &Ra = second ? R2 : R1;
    const HocRender &Rb = second ? R1 : R2;

    float alpha_r;
    const float mt_r = hoc_fp_mask(Ra, b, S, rx, ry, ignore, n_ignore, &alpha_r); /* thresholded * keep */
    float fx, fy;
    hoc_fp_flow(Ra, b, S, rx, ry, mt_r, &fx, &fy);
    float mfinal = mt_r;
    if (mask_occlusions) {
        /* masks that enter the check: render 1 -> thresholded*keep, render 2 -> raw alpha (sic) */
        const float m_r = second ? alpha_r : mt_r;
        const float inv_s = __fdiv_rn(1.0f, (float)S);
        float w0 = 0.0f, w1 = 0.0f, w2 = 0.0f;
        int sx, sy;
        if (hoc_fp_nearest(rx, ry, fx, fy, S, &sx, &sy)) {
            float alpha_s;
            const float mt_s = hoc_fp_mask(Rb, b, S, sx, sy, ignore, n_ignore, &alpha_s);
            float sfx, sfy;
            hoc_fp_flow(Rb, b, S, sx, sy, mt_s, &sfx, &sfy);
            const float m_s = second ? mt_s : alpha_s; /* mask of the OTHER render in the check */
            float g0 = 0.0f, g1 = 0.0f, g2 = 0.0f;
            int qx, qy;
            if (hoc_fp_nearest(sx, sy, sfx, sfy, S, &qx, &qy)) {
                float alpha_q;
                const float mt_q = hoc_fp_mask(Ra, b, S, qx, qy, ignore, n_ignore, &alpha_q);
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
    float *flow = second ? flow21 : flow12;
    float *mult = second ? mult2 : mult1;
    const long o = ((long)b * H + ry) * W + rx;
    *reinterpret_cast<float2 *>(flow + o * 2) = make_float2(fx, fy);
    mult[o] = mask_occlusions ? __fmul_rn(mt_r, mfinal) : mt_r;
}

/* grad_rgb [B,3,S,S] (image layout) = grad_flow [B,H,W,2] * mult inside the crop, 0 elsewhere / channel 2 */
__global__ void __launch_bounds__(FP_THREADS)
hoc_flow_finalize_backward_kernel(const float *__restrict__ grad_flow, const float *__restrict__ mult, int S, int H,
                                  int W, float *__restrict__ grad_rgb)
{
    const int b = blockIdx.y;
    const long pix = (long)blockIdx.x * FP_THREADS + threadIdx.x;
    const long npix = (long)S * S;
    if (pix >= npix)
        return;
    const int y = (int)(pix / S), x = (int)(pix - (long)y * S);
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
extern "C" int hoc_mesh_gather(const float *verts, const float *attrs, const long long *faces_idx, int B, int V, int F,
                               int fill_back, float *faces_out, float *textures_out, void *stream)
{
    HOC_CHECK_ARG(B >= 0 && V >= 0 && F >= 0, "hoc_mesh_gather: bad shape B=%d V=%d F=%d", B, V, F);
    HOC_CHECK_ARG(B <= 65535, "hoc_mesh_gather: batch %d exceeds 65535", B);
    if (B == 0 || F == 0)
        return HOC_OK;
    HOC_CHECK_ARG(verts && faces_idx && faces_out, "hoc_mesh_gather: NULL argument");
    HOC_CHECK_ARG(textures_out == nullptr || attrs != nullptr, "hoc_mesh_gather: textures requested without attributes");
    const int Fo = fill_back ? 2 * F : F;
    dim3 grid((Fo + FP_THREADS - 1) / FP_THREADS, B);
    HOC_LAUNCH(HOC_K_MESH_GATHER, (cudaStream_t)stream,
               (hoc_mesh_gather_kernel<<<grid, FP_THREADS, 0, (cudaStream_t)stream>>>(verts, attrs, faces_idx, V, F,
                                                                                      fill_back, faces_out,
                                                                                      textures_out)));
    HOC_CHECK_LAUNCH("hoc_mesh_gather_kernel");
    return HOC_OK;
}

extern "C" int hoc_mesh_scatter(const float *grad_faces, const float *grad_textures, const long long *faces_idx, int B,
                                int V, int F, int fill_back, float *grad_verts, float *grad_attrs, void *stream)
{
    HOC_CHECK_ARG(B >= 0 && V >= 0 && F >= 0, "hoc_mesh_scatter: bad shape B=%d V=%d F=%d", B, V, F);
    HOC_CHECK_ARG(B <= 65535, "hoc_mesh_scatter: batch %d exceeds 65535", B);
    cudaStream_t st = (cudaStream_t)stream;
    if (B == 0)
        return HOC_OK;
    cudaError_t e = cudaSuccess;
    if (grad_verts != nullptr)
        e = cudaMemsetAsync(grad_verts, 0, sizeof(float) * 3 * (size_t)B * V, st);
    if (e == cudaSuccess && grad_attrs != nullptr)
        e = cudaMemsetAsync(grad_attrs, 0, sizeof(float) * 3 * (size_t)B * V, st);
    if (e != cudaSuccess) {
        hoc_set_error("hoc_mesh_scatter: memset failed: %s", cudaGetErrorString(e));
        return HOC_ERR_CUDA;
    }
    if (F == 0)
        return HOC_OK;
    HOC_CHECK_ARG(faces_idx != nullptr, "hoc_mesh_scatter: faces_idx NULL");
    const int Fo = fill_back ? 2 * F : F;
    dim3 grid((Fo + FP_THREADS - 1) / FP_THREADS, B);
    HOC_LAUNCH(HOC_K_MESH_SCATTER, st,
               (hoc_mesh_scatter_kernel<<<grid, FP_THREADS, 0, st>>>(grad_faces, grad_textures, faces_idx, V, F,
                                                                     fill_back, grad_verts, grad_attrs)));
    HOC_CHECK_LAUNCH("hoc_mesh_scatter_kernel");
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
               (hoc_flow_finalize_backward_kernel<<<grid, FP_THREADS, 0, (cudaStream_t)stream>>>(grad_flow, mult, S, H,
                                                                                                W, grad_rgb)));
    HOC_CHECK_LAUNCH("hoc_flow_finalize_backward_kernel");
    return HOC_OK;
}
