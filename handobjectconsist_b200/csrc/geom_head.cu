/* geom_head.cu -- the "geometry head" in front of the render path: what
 * MeshRegNet.recover_mano / ObjBranch.forward do between the network outputs and the meshes that
 * warpbranch.forward concatenates and renders.
 *
 *   hand   (/root/reference/meshreg/models/meshregnet.py:191-229)
 *          ManoAdaptor (a bias-free Linear 778 -> 21 on the vertices, meshregnet.py:23-51), centring of joints
 *          and vertices on the adapted joint `center_idx`, recover_3d_proj (project.py:5-23: scale / 2-D
 *          translation in pixel space -> camera-space centre), recov_joints3d / recov_handverts3d and both
 *          batch_proj2d -- ONE launch each way instead of ~25 ATen kernels.
 *   points (/root/reference/meshreg/models/objbranch.py:46-77, project.py:5-23)
 *          batch_rodrigues of the predicted axis-angle, rotation of the canonical object vertices,
 *          recover_3d_proj, batch_proj2d -- ONE launch each way.  Without a rotation it is recover_3d_proj itself.
 *
 * A sample is a few thousand points (9-18 KB): one CTA per sample (per point slice in the rotation-only forward),
 * per-sample constants (camera, centre, rotation) built once per CTA in shared memory, all sums reduced inside the
 * CTA with shuffles -- deterministic, no atomics, no workspace.  Latency-bound by construction; what these kernels
 * buy is launches (the step in front of the captured graph), not bandwidth.
 */
#include "hoc_common.cuh"

#define GH_THREADS 256
#define GH_WARPS (GH_THREADS / 32)
#define GH_MAXV 1024 /* hand vertices staged in shared memory (MANO: 778) */
#define GH_MAXJ 32   /* adapted joints (21) */

namespace {

/* ---- forward-mode dual numbers for the derivative of Rodrigues (3 seeds) ------------------- */
struct GDual {
    float v, d;
};
__device__ __forceinline__ GDual gmk(float v, float d = 0.0f)
{
    GDual r;
    r.v = v;
    r.d = d;
    return r;
}
__device__ __forceinline__ GDual operator+(GDual a, GDual b) { return gmk(a.v + b.v, a.d + b.d); }
__device__ __forceinline__ GDual operator-(GDual a, GDual b) { return gmk(a.v - b.v, a.d - b.d); }
__device__ __forceinline__ GDual operator*(GDual a, GDual b) { return gmk(a.v * b.v, a.d * b.v + a.v * b.d); }
__device__ __forceinline__ GDual operator/(GDual a, GDual b)
{
    const float q = a.v / b.v;
    return gmk(q, (a.d - q * b.d) / b.v);
}
__device__ __forceinline__ GDual operator*(float a, GDual b) { return gmk(a * b.v, a * b.d); }
__device__ __forceinline__ GDual operator+(GDual a, float b) { return gmk(a.v + b, a.d); }
__device__ __forceinline__ GDual gsqrt(GDual a)
{
    const float s = sqrtf(a.v);
    return gmk(s, a.d / (2.0f * s));
}
__device__ __forceinline__ GDual gsin(GDual a) { return gmk(sinf(a.v), cosf(a.v) * a.d); }
__device__ __forceinline__ GDual gcos(GDual a) { return gmk(cosf(a.v), -sinf(a.v) * a.d); }
__device__ __forceinline__ float gsqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ float gsin(float a) { return sinf(a); }
__device__ __forceinline__ float gcos(float a) { return cosf(a); }

/* manopth rodrigues_layer.batch_rodrigues (objbranch.py:46): angle = |r + 1e-8|, quaternion of the half angle,
 * normalised, quat2mat; row-major R. */
template <typename T>
__device__ __forceinline__ void gh_rodrigues(const T *aa, T *R)
{
    const T ex = aa[0] + 1e-8f, ey = aa[1] + 1e-8f, ez = aa[2] + 1e-8f;
    const T angle = gsqrt(ex * ex + ey * ey + ez * ez);
    const T nx = aa[0] / angle, ny = aa[1] / angle, nz = aa[2] / angle;
    const T half = 0.5f * angle;
    const T c = gcos(half), s = gsin(half);
    T w = c, x = s * nx, y = s * ny, z = s * nz;
    const T qn = gsqrt(w * w + x * x + y * y + z * z);
    w = w / qn;
    x = x / qn;
    y = y / qn;
    z = z / qn;
    const T w2 = w * w, x2 = x * x, y2 = y * y, z2 = z * z;
    const T wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
    R[0] = w2 + x2 - y2 - z2;
    R[1] = 2.0f * xy - 2.0f * wz;
    R[2] = 2.0f * wy + 2.0f * xz;
    R[3] = 2.0f * wz + 2.0f * xy;
    R[4] = w2 - x2 + y2 - z2;
    R[5] = 2.0f * yz - 2.0f * wx;
    R[6] = 2.0f * xz - 2.0f * wy;
    R[7] = 2.0f * wx + 2.0f * yz;
    R[8] = w2 - x2 - y2 + z2;
}

/* Per-sample camera state of recover_3d_proj (project.py:5-23) in shared memory. */
struct GhCam {
    float K[9];
    float C[3];   /* est_c3d = (X0, Y0, Z0) */
    float f;      /* camintr[b,0,0] */
    float dx, dy; /* est_trans * trans_factor + input_res / 2 - camintr[b,:2,2] */
};

struct GhCamArgs {
    const float *camintr; /* [B or 1,3,3] */
    const float *scale;   /* [B] */
    const float *trans;   /* [B,2] */
    int camintr_batched;
    float scale_factor, trans_factor, off_z, res_w, res_h;
};

__device__ __forceinline__ void gh_camera(const GhCamArgs &a, int b, GhCam &cam)
{
    const float *K = a.camintr + (a.camintr_batched ? (long)b * 9 : 0);
#pragma unroll
    for (int i = 0; i < 9; i++)
        cam.K[i] = K[i];
    const float f = K[0];
    const float s = a.scale[b] * a.scale_factor;
    const float tx = a.trans[2 * b] * a.trans_factor, ty = a.trans[2 * b + 1] * a.trans_factor;
    const float Z0 = f * s + a.off_z;
    cam.f = f;
    cam.dx = (tx + a.res_w / 2.0f) - K[2];
    cam.dy = (ty + a.res_h / 2.0f) - K[5];
    cam.C[0] = cam.dx * Z0 / f;
    cam.C[1] = cam.dy * Z0 / f;
    cam.C[2] = Z0;
}

/* batch_proj2d of one point (libyana.camutils.project; meshregnet.py:228-229, objbranch.py:58) */
__device__ __forceinline__ void gh_project(const float *K, float x, float y, float z, float &u, float &v)
{
    const float h0 = K[0] * x + K[1] * y + K[2] * z;
    const float h1 = K[3] * x + K[4] * y + K[5] * z;
    const float h2 = K[6] * x + K[7] * y + K[8] * z;
    u = h0 / h2;
    v = h1 / h2;
}

/* adjoint of gh_project: (gu, gv) -> added to (gx, gy, gz) */
__device__ __forceinline__ void gh_project_adjoint(const float *K, float x, float y, float z, float gu, float gv,
                                                   float &gx, float &gy, float &gz)
{
    const float h0 = K[0] * x + K[1] * y + K[2] * z;
    const float h1 = K[3] * x + K[4] * y + K[5] * z;
    const float h2 = K[6] * x + K[7] * y + K[8] * z;
    const float a0 = gu / h2, a1 = gv / h2;
    const float a2 = -(gu * (h0 / h2) + gv * (h1 / h2)) / h2;
    gx += K[0] * a0 + K[3] * a1 + K[6] * a2;
    gy += K[1] * a0 + K[4] * a1 + K[7] * a2;
    gz += K[2] * a0 + K[5] * a1 + K[8] * a2;
}

/* d est_c3d -> d scale, d trans (adjoint of gh_camera); one thread */
__device__ __forceinline__ void gh_camera_adjoint(const GhCamArgs &a, const GhCam &cam, int b, float gCx, float gCy,
                                                  float gCz, float *grad_scale, float *grad_trans)
{
    const float Z0 = cam.C[2];
    const float gZ0 = gCz + (gCx * cam.dx + gCy * cam.dy) / cam.f;
    if (grad_scale != nullptr)
        grad_scale[b] = gZ0 * cam.f * a.scale_factor;
    if (grad_trans != nullptr) {
        grad_trans[2 * b] = gCx * Z0 / cam.f * a.trans_factor;
        grad_trans[2 * b + 1] = gCy * Z0 / cam.f * a.trans_factor;
    }
}

/* Sum N per-thread values over the CTA; the totals land in s_out[0..N) (valid after the trailing barrier). */
template <int N>
__device__ __forceinline__ void gh_block_sum(const float *vals, float *s_part /* [GH_WARPS][N] */, float *s_out)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < N; k++) {
        const float s = hoc_warp_sum(vals[k]);
        if (lane == 0)
            s_part[warp * N + k] = s;
    }
    __syncthreads();
    if (threadIdx.x < N) {
        float t = 0.0f;
        for (int w = 0; w < GH_WARPS; w++)
            t += s_part[w * N + threadIdx.x];
        s_out[threadIdx.x] = t;
    }
    __syncthreads();
}

/* ================================ hand head ================================================== */
struct GhHandOut {
    float *joints3d, *verts3d, *recov_joints3d, *recov_verts3d, *joints2d, *verts2d, *center3d;
};

__global__ void __launch_bounds__(GH_THREADS)
hoc_hand_head_forward_kernel(const float *__restrict__ verts, const float *__restrict__ joints_in,
                             const float *__restrict__ adaptor, int V, int J, int center_idx, GhCamArgs ca,
                             GhHandOut out)
{
    __shared__ float s_v[GH_MAXV * 3];
    __shared__ float s_a[GH_MAXJ * 3];
    __shared__ GhCam s_cam;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *vb = verts + (long)b * V * 3;
    for (int i = tid; i < V * 3; i += GH_THREADS)
        s_v[i] = vb[i];
    if (tid == 0)
        gh_camera(ca, b, s_cam);
    if (adaptor == nullptr)
        for (int i = tid; i < J * 3; i += GH_THREADS)
            s_a[i] = joints_in[(long)b * J * 3 + i];
    __syncthreads();
    if (adaptor != nullptr) {
        /* adapted joints = W . verts: a warp per joint, lanes stride the vertices (the row of W is read once,
         * coalesced, and feeds three accumulators; s_v with a stride of 3 words: conflict-free) */
        for (int j = warp; j < J; j += GH_WARPS) {
            const float *w = adaptor + (long)j * V;
            float ax = 0.0f, ay = 0.0f, az = 0.0f;
#pragma unroll 5
            for (int v = lane; v < V; v += 32) {
                const float wv = __ldg(w + v);
                ax = __fmaf_rn(wv, s_v[v * 3], ax);
                ay = __fmaf_rn(wv, s_v[v * 3 + 1], ay);
                az = __fmaf_rn(wv, s_v[v * 3 + 2], az);
            }
            ax = hoc_warp_sum(ax);
            ay = hoc_warp_sum(ay);
            az = hoc_warp_sum(az);
            if (lane == 0) {
                s_a[j * 3] = ax;
                s_a[j * 3 + 1] = ay;
                s_a[j * 3 + 2] = az;
            }
        }
        __syncthreads();
    }
    float cx = 0.0f, cy = 0.0f, cz = 0.0f;
    if (center_idx >= 0) {
        cx = s_a[center_idx * 3];
        cy = s_a[center_idx * 3 + 1];
        cz = s_a[center_idx * 3 + 2];
    }
    const float Cx = s_cam.C[0], Cy = s_cam.C[1], Cz = s_cam.C[2];
    if (tid < 3 && out.center3d != nullptr)
        out.center3d[b * 3 + tid] = s_cam.C[tid];
    /* coordinate-parallel, coalesced: centred and recovered positions */
    for (int i = tid; i < V * 3; i += GH_THREADS) {
        const int c = i % 3;
        const float cen = (c == 0) ? cx : ((c == 1) ? cy : cz);
        const float C = (c == 0) ? Cx : ((c == 1) ? Cy : Cz);
        const float p = s_v[i] - cen;
        if (out.verts3d != nullptr)
            out.verts3d[(long)b * V * 3 + i] = p;
        if (out.recov_verts3d != nullptr)
            out.recov_verts3d[(long)b * V * 3 + i] = p + C;
    }
    for (int i = tid; i < J * 3; i += GH_THREADS) {
        const int c = i % 3;
        const float cen = (c == 0) ? cx : ((c == 1) ? cy : cz);
        const float C = (c == 0) ? Cx : ((c == 1) ? Cy : Cz);
        const float p = s_a[i] - cen;
        if (out.joints3d != nullptr)
            out.joints3d[(long)b * J * 3 + i] = p;
        if (out.recov_joints3d != nullptr)
            out.recov_joints3d[(long)b * J * 3 + i] = p + C;
    }
    /* point-parallel: projections (8-byte stores, coalesced) */
    if (out.verts2d != nullptr)
        for (int v = tid; v < V; v += GH_THREADS) {
            float u, w;
            gh_project(s_cam.K, (s_v[v * 3] - cx) + Cx, (s_v[v * 3 + 1] - cy) + Cy, (s_v[v * 3 + 2] - cz) + Cz, u, w);
            reinterpret_cast<float2 *>(out.verts2d)[(long)b * V + v] = make_float2(u, w);
        }
    if (out.joints2d != nullptr && tid < J) {
        float u, w;
        gh_project(s_cam.K, (s_a[tid * 3] - cx) + Cx, (s_a[tid * 3 + 1] - cy) + Cy, (s_a[tid * 3 + 2] - cz) + Cz, u, w);
        reinterpret_cast<float2 *>(out.joints2d)[(long)b * J + tid] = make_float2(u, w);
    }
}

struct GhHandGrad {
    const float *joints3d, *verts3d, *recov_joints3d, *recov_verts3d, *joints2d, *verts2d, *center3d;
};

/* recov_verts3d / recov_joints3d are the forward's outputs (the points the projections were taken at). */
__global__ void __launch_bounds__(GH_THREADS)
hoc_hand_head_backward_kernel(const float *__restrict__ recov_verts, const float *__restrict__ recov_joints,
                              const float *__restrict__ adaptor, int V, int J, int center_idx, GhCamArgs ca,
                              GhHandGrad g, float *__restrict__ grad_verts, float *__restrict__ grad_joints_in,
                              float *__restrict__ grad_adapt, float *__restrict__ grad_scale,
                              float *__restrict__ grad_trans)
{
    __shared__ float s_g[GH_MAXV * 3];  /* d L / d verts3d (centred vertices), total */
    __shared__ float s_ga[GH_MAXJ * 3]; /* d L / d adapted joints, total */
    __shared__ float s_part[GH_WARPS * 6];
    __shared__ float s_sum[12];
    __shared__ GhCam s_cam;
    const int b = blockIdx.x, tid = threadIdx.x;
    if (tid == 0)
        gh_camera(ca, b, s_cam);
    __syncthreads();

    /* vertices: G_rv = g_recov + proj^T g_2d ;  G_v3d = g_verts3d + G_rv */
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}; /* sum G_rv (3), sum G_v3d (3) */
    for (int v = tid; v < V; v += GH_THREADS) {
        const long o = ((long)b * V + v) * 3;
        float rx = 0.f, ry = 0.f, rz = 0.f;
        if (g.recov_verts3d != nullptr) {
            rx = g.recov_verts3d[o];
            ry = g.recov_verts3d[o + 1];
            rz = g.recov_verts3d[o + 2];
        }
        if (g.verts2d != nullptr) {
            const float2 g2 = reinterpret_cast<const float2 *>(g.verts2d)[(long)b * V + v];
            gh_project_adjoint(s_cam.K, recov_verts[o], recov_verts[o + 1], recov_verts[o + 2], g2.x, g2.y, rx, ry, rz);
        }
        float tx = rx, ty = ry, tz = rz;
        if (g.verts3d != nullptr) {
            tx += g.verts3d[o];
            ty += g.verts3d[o + 1];
            tz += g.verts3d[o + 2];
        }
        s_g[v * 3] = tx;
        s_g[v * 3 + 1] = ty;
        s_g[v * 3 + 2] = tz;
        acc[0] += rx;
        acc[1] += ry;
        acc[2] += rz;
        acc[3] += tx;
        acc[4] += ty;
        acc[5] += tz;
    }
    gh_block_sum<6>(acc, s_part, s_sum);
    /* joints: same, on the first warp */
    float jac[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (tid < J) {
        const long o = ((long)b * J + tid) * 3;
        float rx = 0.f, ry = 0.f, rz = 0.f;
        if (g.recov_joints3d != nullptr) {
            rx = g.recov_joints3d[o];
            ry = g.recov_joints3d[o + 1];
            rz = g.recov_joints3d[o + 2];
        }
        if (g.joints2d != nullptr) {
            const float2 g2 = reinterpret_cast<const float2 *>(g.joints2d)[(long)b * J + tid];
            gh_project_adjoint(s_cam.K, recov_joints[o], recov_joints[o + 1], recov_joints[o + 2], g2.x, g2.y, rx, ry,
                               rz);
        }
        float tx = rx, ty = ry, tz = rz;
        if (g.joints3d != nullptr) {
            tx += g.joints3d[o];
            ty += g.joints3d[o + 1];
            tz += g.joints3d[o + 2];
        }
        s_ga[tid * 3] = tx;
        s_ga[tid * 3 + 1] = ty;
        s_ga[tid * 3 + 2] = tz;
        jac[0] = rx;
        jac[1] = ry;
        jac[2] = rz;
        jac[3] = tx;
        jac[4] = ty;
        jac[5] = tz;
    }
    if (tid < 32) {
#pragma unroll
        for (int k = 0; k < 6; k++)
            jac[k] = hoc_warp_sum(jac[k]);
        if (tid == 0)
            for (int k = 0; k < 6; k++)
                s_sum[6 + k] = jac[k];
    }
    __syncthreads();
    if (tid == 0) {
        /* d est_c3d = sum of everything that was translated by it */
        float gC[3];
        for (int c = 0; c < 3; c++)
            gC[c] = s_sum[c] + s_sum[6 + c] + (g.center3d != nullptr ? g.center3d[b * 3 + c] : 0.0f);
        gh_camera_adjoint(ca, s_cam, b, gC[0], gC[1], gC[2], grad_scale, grad_trans);
        /* the centre joint was subtracted from every centred joint and vertex */
        if (center_idx >= 0)
            for (int c = 0; c < 3; c++)
                s_ga[center_idx * 3 + c] -= s_sum[3 + c] + s_sum[9 + c];
    }
    __syncthreads();
    if (adaptor == nullptr) {
        if (grad_joints_in != nullptr)
            for (int i = tid; i < J * 3; i += GH_THREADS)
                grad_joints_in[(long)b * J * 3 + i] = s_ga[i];
    } else if (grad_adapt != nullptr) {
        for (int i = tid; i < J * 3; i += GH_THREADS)
            grad_adapt[(long)b * J * 3 + i] = s_ga[i];
    }
    if (grad_verts != nullptr) {
        /* d verts = G_v3d + W^T d adapted: thread per vertex (W columns: coalesced over v) */
        for (int v = tid; v < V; v += GH_THREADS) {
            float ax = 0.f, ay = 0.f, az = 0.f;
            if (adaptor != nullptr)
#pragma unroll 7
                for (int j = 0; j < J; j++) {
                    const float w = __ldg(adaptor + (long)j * V + v);
                    ax = __fmaf_rn(w, s_ga[j * 3], ax);
                    ay = __fmaf_rn(w, s_ga[j * 3 + 1], ay);
                    az = __fmaf_rn(w, s_ga[j * 3 + 2], az);
                }
            s_g[v * 3] += ax;
            s_g[v * 3 + 1] += ay;
            s_g[v * 3 + 2] += az;
        }
        __syncthreads();
        for (int i = tid; i < V * 3; i += GH_THREADS)
            grad_verts[(long)b * V * 3 + i] = s_g[i];
    }
}

/* ================================ rotated / recovered points ================================ */
__global__ void __launch_bounds__(GH_THREADS)
hoc_recover_points_forward_kernel(const float *__restrict__ points, const float *__restrict__ rot, int N,
                                  GhCamArgs ca, float *__restrict__ rot_points, float *__restrict__ recov_points,
                                  float *__restrict__ points2d, float *__restrict__ center3d)
{
    __shared__ GhCam s_cam;
    __shared__ float s_R[9];
    const int b = blockIdx.y, tid = threadIdx.x;
    if (tid == 0) {
        gh_camera(ca, b, s_cam);
        if (rot != nullptr) {
            float aa[3] = {rot[b * 3], rot[b * 3 + 1], rot[b * 3 + 2]};
            gh_rodrigues<float>(aa, s_R);
        }
    }
    __syncthreads();
    if (blockIdx.x == 0 && tid < 3 && center3d != nullptr)
        center3d[b * 3 + tid] = s_cam.C[tid];
    const int n = blockIdx.x * GH_THREADS + tid;
    if (n >= N)
        return;
    const long o = ((long)b * N + n) * 3;
    float x = points[o], y = points[o + 1], z = points[o + 2];
    if (rot != nullptr) {
        const float rx = s_R[0] * x + s_R[1] * y + s_R[2] * z;
        const float ry = s_R[3] * x + s_R[4] * y + s_R[5] * z;
        const float rz = s_R[6] * x + s_R[7] * y + s_R[8] * z;
        x = rx;
        y = ry;
        z = rz;
        if (rot_points != nullptr) {
            rot_points[o] = x;
            rot_points[o + 1] = y;
            rot_points[o + 2] = z;
        }
    }
    x += s_cam.C[0];
    y += s_cam.C[1];
    z += s_cam.C[2];
    if (recov_points != nullptr) {
        recov_points[o] = x;
        recov_points[o + 1] = y;
        recov_points[o + 2] = z;
    }
    if (points2d != nullptr) {
        float u, v;
        gh_project(s_cam.K, x, y, z, u, v);
        reinterpret_cast<float2 *>(points2d)[(long)b * N + n] = make_float2(u, v);
    }
}

__global__ void __launch_bounds__(GH_THREADS)
hoc_recover_points_backward_kernel(const float *__restrict__ points, const float *__restrict__ rot, int N,
                                   GhCamArgs ca, const float *__restrict__ g_rot_points,
                                   const float *__restrict__ g_recov, const float *__restrict__ g_points2d,
                                   const float *__restrict__ g_center3d, float *__restrict__ grad_points,
                                   float *__restrict__ grad_rot, float *__restrict__ grad_scale,
                                   float *__restrict__ grad_trans)
{
    __shared__ GhCam s_cam;
    __shared__ float s_R[9];
    __shared__ float s_part[GH_WARPS * 12];
    __shared__ float s_sum[12];
    __shared__ float s_dR[27];
    const int b = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) {
        gh_camera(ca, b, s_cam);
        if (rot != nullptr) {
            float aa[3] = {rot[b * 3], rot[b * 3 + 1], rot[b * 3 + 2]};
            gh_rodrigues<float>(aa, s_R);
        }
    } else if (rot != nullptr && tid >= 32 && tid < 35) {
        /* d R / d rot[k]: the same code on dual numbers, seed k */
        const int k = tid - 32;
        GDual aa[3], dR[9];
        for (int c = 0; c < 3; c++)
            aa[c] = gmk(rot[b * 3 + c], c == k ? 1.0f : 0.0f);
        gh_rodrigues<GDual>(aa, dR);
        for (int i = 0; i < 9; i++)
            s_dR[k * 9 + i] = dR[i].d;
    }
    __syncthreads();
    /* acc[0..3): sum of d recovered points (= d est_c3d);  acc[3..12): G_R[i][k] = sum G_rot[i] * x[k] */
    float acc[12];
#pragma unroll
    for (int k = 0; k < 12; k++)
        acc[k] = 0.0f;
    for (int n = tid; n < N; n += GH_THREADS) {
        const long o = ((long)b * N + n) * 3;
        const float x = points[o], y = points[o + 1], z = points[o + 2];
        float px = x, py = y, pz = z;
        if (rot != nullptr) {
            px = s_R[0] * x + s_R[1] * y + s_R[2] * z;
            py = s_R[3] * x + s_R[4] * y + s_R[5] * z;
            pz = s_R[6] * x + s_R[7] * y + s_R[8] * z;
        }
        float rx = 0.f, ry = 0.f, rz = 0.f;
        if (g_recov != nullptr) {
            rx = g_recov[o];
            ry = g_recov[o + 1];
            rz = g_recov[o + 2];
        }
        if (g_points2d != nullptr) {
            const float2 g2 = reinterpret_cast<const float2 *>(g_points2d)[(long)b * N + n];
            gh_project_adjoint(s_cam.K, px + s_cam.C[0], py + s_cam.C[1], pz + s_cam.C[2], g2.x, g2.y, rx, ry, rz);
        }
        acc[0] += rx;
        acc[1] += ry;
        acc[2] += rz;
        if (rot != nullptr && g_rot_points != nullptr) {
            rx += g_rot_points[o];
            ry += g_rot_points[o + 1];
            rz += g_rot_points[o + 2];
        }
        if (rot != nullptr) {
            acc[3] += rx * x;
            acc[4] += rx * y;
            acc[5] += rx * z;
            acc[6] += ry * x;
            acc[7] += ry * y;
            acc[8] += ry * z;
            acc[9] += rz * x;
            acc[10] += rz * y;
            acc[11] += rz * z;
        }
        if (grad_points != nullptr) {
            float gx = rx, gy = ry, gz = rz;
            if (rot != nullptr) { /* R^T G */
                gx = s_R[0] * rx + s_R[3] * ry + s_R[6] * rz;
                gy = s_R[1] * rx + s_R[4] * ry + s_R[7] * rz;
                gz = s_R[2] * rx + s_R[5] * ry + s_R[8] * rz;
            }
            grad_points[o] = gx;
            grad_points[o + 1] = gy;
            grad_points[o + 2] = gz;
        }
    }
    gh_block_sum<12>(acc, s_part, s_sum);
    if (tid == 0) {
        float gC[3];
        for (int c = 0; c < 3; c++)
            gC[c] = s_sum[c] + (g_center3d != nullptr ? g_center3d[b * 3 + c] : 0.0f);
        gh_camera_adjoint(ca, s_cam, b, gC[0], gC[1], gC[2], grad_scale, grad_trans);
    }
    if (rot != nullptr && grad_rot != nullptr && tid < 3) {
        float t = 0.0f;
        for (int i = 0; i < 9; i++)
            t += s_sum[3 + i] * s_dR[tid * 9 + i];
        grad_rot[b * 3 + tid] = t;
    }
}

int gh_check_cam(const char *who, const float *camintr, const float *scale, const float *trans, int B, int N)
{
    HOC_CHECK_ARG(B >= 0 && B <= 65535, "%s: batch %d outside [0, 65535]", who, B);
    HOC_CHECK_ARG(N >= 0, "%s: negative point count %d", who, N);
    if (B > 0)
        HOC_CHECK_ARG(camintr && scale && trans, "%s: camintr / scale / trans is NULL", who);
    return HOC_OK;
}

GhCamArgs gh_cam_args(const float *camintr, int camintr_batched, const float *scale, const float *trans,
                      float scale_factor, float trans_factor, float off_z, float res_w, float res_h)
{
    GhCamArgs a;
    a.camintr = camintr;
    a.scale = scale;
    a.trans = trans;
    a.camintr_batched = camintr_batched;
    a.scale_factor = scale_factor;
    a.trans_factor = trans_factor;
    a.off_z = off_z;
    a.res_w = res_w;
    a.res_h = res_h;
    return a;
}

} // namespace

extern "C" int hoc_hand_head_forward(const float *verts, const float *joints_in, const float *adaptor, int B, int V,
                                     int J, int center_idx, const float *camintr, int camintr_batched,
                                     const float *scale, const float *trans, float scale_factor, float trans_factor,
                                     float off_z, float res_w, float res_h, float *joints3d, float *verts3d,
                                     float *recov_joints3d, float *recov_verts3d, float *joints2d, float *verts2d,
                                     float *center3d, void *stream)
{
    const int rc = gh_check_cam("hoc_hand_head_forward", camintr, scale, trans, B, V);
    if (rc != HOC_OK)
        return rc;
    HOC_CHECK_ARG(V <= GH_MAXV, "hoc_hand_head_forward: %d vertices, at most %d", V, GH_MAXV);
    HOC_CHECK_ARG(J >= 0 && J <= GH_MAXJ, "hoc_hand_head_forward: %d joints, at most %d", J, GH_MAXJ);
    HOC_CHECK_ARG(center_idx >= -1 && center_idx < J, "hoc_hand_head_forward: center_idx %d with %d joints", center_idx,
                  J);
    if (B == 0)
        return HOC_OK;
    HOC_CHECK_ARG(verts != nullptr, "hoc_hand_head_forward: verts is NULL");
    HOC_CHECK_ARG(adaptor != nullptr || joints_in != nullptr || J == 0,
                  "hoc_hand_head_forward: neither an adaptor nor joints given");
    HOC_CHECK_ARG((((uintptr_t)joints2d | (uintptr_t)verts2d) & 7) == 0,
                  "hoc_hand_head_forward: 2-D outputs must be 8-byte aligned");
    GhHandOut out = {joints3d, verts3d, recov_joints3d, recov_verts3d, joints2d, verts2d, center3d};
    cudaStream_t st = (cudaStream_t)stream;
    HOC_LAUNCH(HOC_K_HAND_HEAD_FWD, st,
               (hoc_hand_head_forward_kernel<<<B, GH_THREADS, 0, st>>>(
                   verts, joints_in, adaptor, V, J, center_idx,
                   gh_cam_args(camintr, camintr_batched, scale, trans, scale_factor, trans_factor, off_z, res_w, res_h),
                   out)));
    HOC_CHECK_LAUNCH("hoc_hand_head_forward_kernel");
    return HOC_OK;
}

extern "C" int hoc_hand_head_backward(const float *recov_verts3d, const float *recov_joints3d, const float *adaptor,
                                      int B, int V, int J, int center_idx, const float *camintr, int camintr_batched,
                                      const float *scale, const float *trans, float scale_factor, float trans_factor,
                                      float off_z, float res_w, float res_h, const float *g_joints3d,
                                      const float *g_verts3d, const float *g_recov_joints3d,
                                      const float *g_recov_verts3d, const float *g_joints2d, const float *g_verts2d,
                                      const float *g_center3d, float *grad_verts, float *grad_joints_in,
                                      float *grad_adapt, float *grad_scale, float *grad_trans, void *stream)
{
    const int rc = gh_check_cam("hoc_hand_head_backward", camintr, scale, trans, B, V);
    if (rc != HOC_OK)
        return rc;
    HOC_CHECK_ARG(V <= GH_MAXV, "hoc_hand_head_backward: %d vertices, at most %d", V, GH_MAXV);
    HOC_CHECK_ARG(J >= 0 && J <= GH_MAXJ, "hoc_hand_head_backward: %d joints, at most %d", J, GH_MAXJ);
    HOC_CHECK_ARG(center_idx >= -1 && center_idx < J, "hoc_hand_head_backward: center_idx %d with %d joints",
                  center_idx, J);
    if (B == 0)
        return HOC_OK;
    HOC_CHECK_ARG(g_verts2d == nullptr || recov_verts3d != nullptr,
                  "hoc_hand_head_backward: a gradient of verts2d needs recov_verts3d");
    HOC_CHECK_ARG(g_joints2d == nullptr || recov_joints3d != nullptr,
                  "hoc_hand_head_backward: a gradient of joints2d needs recov_joints3d");
    HOC_CHECK_ARG((((uintptr_t)g_joints2d | (uintptr_t)g_verts2d) & 7) == 0,
                  "hoc_hand_head_backward: 2-D gradients must be 8-byte aligned");
    GhHandGrad g = {g_joints3d, g_verts3d, g_recov_joints3d, g_recov_verts3d, g_joints2d, g_verts2d, g_center3d};
    cudaStream_t st = (cudaStream_t)stream;
    HOC_LAUNCH(HOC_K_HAND_HEAD_BWD, st,
               (hoc_hand_head_backward_kernel<<<B, GH_THREADS, 0, st>>>(
                   recov_verts3d, recov_joints3d, adaptor, V, J, center_idx,
                   gh_cam_args(camintr, camintr_batched, scale, trans, scale_factor, trans_factor, off_z, res_w, res_h),
                   g, grad_verts, grad_joints_in, grad_adapt, grad_scale, grad_trans)));
    HOC_CHECK_LAUNCH("hoc_hand_head_backward_kernel");
    return HOC_OK;
}

extern "C" int hoc_recover_points_forward(const float *points, const float *rotaxisang, int B, int N,
                                          const float *camintr, int camintr_batched, const float *scale,
                                          const float *trans, float scale_factor, float trans_factor, float off_z,
                                          float res_w, float res_h, float *rot_points, float *recov_points,
                                          float *points2d, float *center3d, void *stream)
{
    const int rc = gh_check_cam("hoc_recover_points_forward", camintr, scale, trans, B, N);
    if (rc != HOC_OK)
        return rc;
    if (B == 0)
        return HOC_OK;
    HOC_CHECK_ARG(points != nullptr || N == 0, "hoc_recover_points_forward: points is NULL");
    HOC_CHECK_ARG(((uintptr_t)points2d & 7) == 0, "hoc_recover_points_forward: points2d must be 8-byte aligned");
    dim3 grid(N > 0 ? (N + GH_THREADS - 1) / GH_THREADS : 1, B);
    cudaStream_t st = (cudaStream_t)stream;
    HOC_LAUNCH(HOC_K_RECOVER_POINTS_FWD, st,
               (hoc_recover_points_forward_kernel<<<grid, GH_THREADS, 0, st>>>(
                   points, rotaxisang, N,
                   gh_cam_args(camintr, camintr_batched, scale, trans, scale_factor, trans_factor, off_z, res_w, res_h),
                   rot_points, recov_points, points2d, center3d)));
    HOC_CHECK_LAUNCH("hoc_recover_points_forward_kernel");
    return HOC_OK;
}

extern "C" int hoc_recover_points_backward(const float *points, const float *rotaxisang, int B, int N,
                                           const float *camintr, int camintr_batched, const float *scale,
                                           const float *trans, float scale_factor, float trans_factor, float off_z,
                                           float res_w, float res_h, const float *g_rot_points,
                                           const float *g_recov_points, const float *g_points2d,
                                           const float *g_center3d, float *grad_points, float *grad_rot,
                                           float *grad_scale, float *grad_trans, void *stream)
{
    const int rc = gh_check_cam("hoc_recover_points_backward", camintr, scale, trans, B, N);
    if (rc != HOC_OK)
        return rc;
    if (B == 0)
        return HOC_OK;
    HOC_CHECK_ARG(points != nullptr || N == 0, "hoc_recover_points_backward: points is NULL");
    HOC_CHECK_ARG(((uintptr_t)g_points2d & 7) == 0, "hoc_recover_points_backward: g_points2d must be 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    HOC_LAUNCH(HOC_K_RECOVER_POINTS_BWD, st,
               (hoc_recover_points_backward_kernel<<<B, GH_THREADS, 0, st>>>(
                   points, rotaxisang, N,
                   gh_cam_args(camintr, camintr_batched, scale, trans, scale_factor, trans_factor, off_z, res_w, res_h),
                   g_rot_points, g_recov_points, g_points2d, g_center3d, grad_points, grad_rot, grad_scale,
                   grad_trans)));
    HOC_CHECK_LAUNCH("hoc_recover_points_backward_kernel");
    return HOC_OK;
}
