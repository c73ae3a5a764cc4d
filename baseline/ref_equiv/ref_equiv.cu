/*
 * baseline/ref_equiv/ref_equiv.cu -- BENCHMARK BASELINE ONLY (SURVEY.md section 8d "GPU reference-equivalent
 * baseline"); never imported by the product package.
 *
 * The reference's rasterizer kernels live in the wheel `neural-renderer-pytorch`, which is absent and not
 * installable here, so the >= 10x claim of BASELINE.json needs a stand-in for "the reference on the same B200".
 * This file runs the CPU oracle's per-item bodies (oracle/nmr_oracle_impl.h: one call = what ONE thread of the
 * reference's kernels does) with the reference's LAUNCH STRUCTURE: one thread per face for the barycentric
 * matrices, one thread per pixel walking ALL faces for the forward, one thread per pixel for texture sampling,
 * ONE SERIAL THREAD PER FACE for the pseudo-gradient, one thread per pixel with float atomics for the texture and
 * depth gradients -- the five entry points bound at /root/reference/meshreg/neurender/rasterize.py:202,232,269,
 * 290,306.  Clearly a RESTATEMENT: it is as fast as a straightforward CUDA build of that structure, not tuned.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define ORA_HD static __device__ __forceinline__
#define ORA_ATOMIC_ADD(ptr, val) atomicAdd((ptr), (val))
#define ORA_NO_HOST_DRIVERS
#define ORA_IMIN(a, b) ((a) < (b) ? (a) : (b))
#define ORA_IMAX(a, b) ((a) > (b) ? (a) : (b))
#define REAL float
#define FN(name) refeq_##name
#include "../../oracle/nmr_oracle_impl.h"

#define REFEQ_THREADS 512 /* upstream launches 512-thread blocks */

__global__ void refeq_face_inv_kernel(const float *faces, float *faces_inv, long n, int is)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        refeq_face_inv(faces + i * 9, is, faces_inv + i * 9);
}

__global__ void refeq_fwd_kernel(struct refeq_fwd_ctx c, long n)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        refeq_fwd_pixel(&c, i);
}

__global__ void refeq_tex_kernel(const float *faces, const float *textures, const int32_t *face_index_map,
                                 const float *weight_map, const float *depth_map, float *rgb_map,
                                 int32_t *sampling_index_map, float *sampling_weight_map, int nf, int is, int ts,
                                 float eps, long n)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        refeq_tex_pixel(faces, textures, face_index_map, weight_map, depth_map, rgb_map, sampling_index_map,
                        sampling_weight_map, nf, is, ts, eps, i);
}

__global__ void refeq_k4_kernel(struct refeq_k4_ctx c, long n)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        refeq_k4_face(&c, i);
}

__global__ void refeq_bt_kernel(const int32_t *face_index_map, const float *sampling_weight_map,
                                const int32_t *sampling_index_map, const float *grad_rgb_map, float *grad_textures,
                                int nf, int is, int ts, long n)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        refeq_bt_pixel(face_index_map, sampling_weight_map, sampling_index_map, grad_rgb_map, grad_textures, nf, is, ts,
                       i);
}

__global__ void refeq_bd_kernel(const float *faces, const float *depth_map, const int32_t *face_index_map,
                                const float *face_inv_map, const float *weight_map, const float *grad_depth_map,
                                float *grad_faces, int nf, int is, long n)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        refeq_bd_pixel(faces, depth_map, face_index_map, face_inv_map, weight_map, grad_depth_map, grad_faces, nf, is,
                       i);
}

static unsigned blocks(long n) { return (unsigned)((n + REFEQ_THREADS - 1) / REFEQ_THREADS); }

extern "C" {

int refeq_forward_face_index_map(const float *faces, int32_t *face_index_map, float *weight_map, float *depth_map,
                                 float *face_inv_map, float *faces_inv, int B, int nf, int is, float near_, float far_,
                                 int return_depth, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    const long nfaces = (long)B * nf, npix = (long)B * is * is;
    if (nfaces == 0 || npix == 0)
        return 0;
    refeq_face_inv_kernel<<<blocks(nfaces), REFEQ_THREADS, 0, st>>>(faces, faces_inv, nfaces, is);
    struct refeq_fwd_ctx c = {faces, face_index_map, weight_map, depth_map, face_inv_map, faces_inv, nf, is, near_, far_,
                              return_depth};
    refeq_fwd_kernel<<<blocks(npix), REFEQ_THREADS, 0, st>>>(c, npix);
    return (int)cudaGetLastError();
}

int refeq_forward_texture_sampling(const float *faces, const float *textures, const int32_t *face_index_map,
                                   const float *weight_map, const float *depth_map, float *rgb_map,
                                   int32_t *sampling_index_map, float *sampling_weight_map, int B, int nf, int is, int ts,
                                   float eps, void *stream)
{
    const long npix = (long)B * is * is;
    if (npix == 0)
        return 0;
    refeq_tex_kernel<<<blocks(npix), REFEQ_THREADS, 0, (cudaStream_t)stream>>>(
        faces, textures, face_index_map, weight_map, depth_map, rgb_map, sampling_index_map, sampling_weight_map, nf, is,
        ts, eps, npix);
    return (int)cudaGetLastError();
}

int refeq_backward_pixel_map(const float *faces, const int32_t *face_index_map, const float *rgb_map,
                             const float *alpha_map, const float *grad_rgb_map, const float *grad_alpha_map,
                             float *grad_faces, int B, int nf, int is, float eps, int return_rgb, int return_alpha,
                             void *stream)
{
    const long nfaces = (long)B * nf;
    if (nfaces == 0)
        return 0;
    struct refeq_k4_ctx c = {faces, face_index_map, rgb_map, alpha_map, grad_rgb_map, grad_alpha_map, grad_faces, nf, is,
                             eps, return_rgb, return_alpha};
    refeq_k4_kernel<<<blocks(nfaces), REFEQ_THREADS, 0, (cudaStream_t)stream>>>(c, nfaces);
    return (int)cudaGetLastError();
}

int refeq_backward_textures(const int32_t *face_index_map, const float *sampling_weight_map,
                            const int32_t *sampling_index_map, const float *grad_rgb_map, float *grad_textures, int B,
                            int nf, int is, int ts, void *stream)
{
    const long npix = (long)B * is * is;
    if (npix == 0)
        return 0;
    refeq_bt_kernel<<<blocks(npix), REFEQ_THREADS, 0, (cudaStream_t)stream>>>(
        face_index_map, sampling_weight_map, sampling_index_map, grad_rgb_map, grad_textures, nf, is, ts, npix);
    return (int)cudaGetLastError();
}

int refeq_backward_depth_map(const float *faces, const float *depth_map, const int32_t *face_index_map,
                             const float *face_inv_map, const float *weight_map, const float *grad_depth_map,
                             float *grad_faces, int B, int nf, int is, void *stream)
{
    const long npix = (long)B * is * is;
    if (npix == 0)
        return 0;
    refeq_bd_kernel<<<blocks(npix), REFEQ_THREADS, 0, (cudaStream_t)stream>>>(
        faces, depth_map, face_index_map, face_inv_map, weight_map, grad_depth_map, grad_faces, nf, is, npix);
    return (int)cudaGetLastError();
}
}
