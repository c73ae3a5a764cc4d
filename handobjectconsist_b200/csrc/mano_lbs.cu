/*
 * mano_lbs.cu -- MANO linear-blend skinning, forward and backward, for sm_100a.
 *
 * Replaces `manopth.manolayer.ManoLayer.forward` (absent third-party dependency, called from
 * /root/reference/meshreg/models/manobranch.py:70-85,139-145) for axis-angle / PCA pose input:
 *   PCA -> full pose (48) -> Rodrigues x16 (quaternion form, `+1e-8` in the norm) -> pose_map = R[1:] - I
 *   v_shaped = template + shapedirs . betas;  J = J_regressor . v_shaped
 *   v_posed = v_shaped + posedirs . pose_map
 *   forward kinematics (root + 5 fingers x 3 levels), rest joints removed
 *   per-vertex blend of the 16 transforms, fingertip vertices appended to the joints, joint reorder,
 *   centring on `center_idx` (or + trans), millimetres.
 * In torch this is ~60 tiny launches (batched 4x4 matmuls, cats, index selects).  Here the forward is ONE launch with
 * a CTA per (64-vertex slice, sample): every CTA rebuilds the sample's small state (16 rotations, joints, kinematic
 * chain: a few hundred flops) and produces its slice -- the joint regression is folded into two model constants
 * (J = j_template + j_shapedirs . betas), so no CTA needs the whole shaped mesh, and the pose blend shapes are read
 * through a transposed copy, coalesced.  (One CTA per sample, the first version, left 116 of 148 SMs idle at B = 32
 * and took 265 us; this one 20 us.)  The backward is two launches: the vertex-parallel part per (slice, sample) -- the big linear maps
 * (posedirs, shapedirs, skinning weights) reversed by hand, reduced in the CTA and added to per-sample accumulators --
 * and a small per-sample kernel that differentiates the 16-joint kinematic chain and the Rodrigues formula with
 * forward-mode dual numbers (96 seeds: 48 pose + 48 joint coordinates, one per thread) through the SAME templated
 * code the forward runs.
 */
#include "hoc_common.cuh"

#define MN_THREADS 256
#define MN_MAXV 1024 /* vertices (MANO: 778) */

struct Dual {
    float v, d;
};
__device__ __forceinline__ Dual mk(float v, float d = 0.0f)
{
    Dual r;
    r.v = v;
    r.d = d;
    return r;
}
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { return mk(a.v + b.v, a.d + b.d); }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { return mk(a.v - b.v, a.d - b.d); }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) { return mk(a.v * b.v, a.d * b.v + a.v * b.d); }
__device__ __forceinline__ Dual operator/(Dual a, Dual b)
{
    const float q = a.v / b.v;
    return mk(q, (a.d - q * b.d) / b.v);
}
__device__ __forceinline__ Dual operator*(float a, Dual b) { return mk(a * b.v, a * b.d); }
__device__ __forceinline__ Dual operator+(Dual a, float b) { return mk(a.v + b, a.d); }
__device__ __forceinline__ Dual msqrt(Dual a)
{
    const float s = sqrtf(a.v);
    return mk(s, a.d / (2.0f * s));
}
__device__ __forceinline__ Dual msin(Dual a) { return mk(sinf(a.v), cosf(a.v) * a.d); }
__device__ __forceinline__ Dual mcos(Dual a) { return mk(cosf(a.v), -sinf(a.v) * a.d); }
__device__ __forceinline__ float msqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ float msin(float a) { return sinf(a); }
__device__ __forceinline__ float mcos(float a) { return cosf(a); }
__device__ __forceinline__ float val(float a) { return a; }
__device__ __forceinline__ float val(Dual a) { return a.v; }

/* manopth rodrigues_layer.batch_rodrigues + quat2mat for one joint */
template <typename T>
__device__ __forceinline__ void mano_rodrigues(const T *aa, T *R)
{
    const T ex = aa[0] + 1e-8f, ey = aa[1] + 1e-8f, ez = aa[2] + 1e-8f;
    const T angle = msqrt(ex * ex + ey * ey + ez * ez);
    const T nx = aa[0] / angle, ny = aa[1] / angle, nz = aa[2] / angle;
    const T half = 0.5f * angle;
    const T c = mcos(half), s = msin(half);
    T w = c, x = s * nx, y = s * ny, z = s * nz;
    const T qn = msqrt(w * w + x * x + y * y + z * z);
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

/* G_child = G_parent o [R | j_rel]:  GR' = GR R,  Gt' = GR j_rel + Gt */
template <typename T>
__device__ __forceinline__ void mano_compose(const T *GR, const T *Gt, const T *R, const T *jrel, T *oR, T *ot)
{
#pragma unroll
    for (int r = 0; r < 3; r++) {
#pragma unroll
        for (int c = 0; c < 3; c++)
            oR[3 * r + c] = GR[3 * r] * R[c] + GR[3 * r + 1] * R[3 + c] + GR[3 * r + 2] * R[6 + c];
        ot[r] = GR[3 * r] * jrel[0] + GR[3 * r + 1] * jrel[1] + GR[3 * r + 2] * jrel[2] + Gt[r];
    }
}

/* Visits the 16 joints in kinematic order (root, then each finger base -> tip) and calls
 * sink(j, GR, Gt) with the global transform of joint j.  getR(j, R9) / getJ(j, J3) supply the inputs. */
template <typename T, typename GetR, typename GetJ, typename Sink>
__device__ __forceinline__ void mano_fk(GetR getR, GetJ getJ, Sink sink)
{
    T R0[9], J0[3];
    getR(0, R0);
    getJ(0, J0);
    sink(0, R0, J0);
    for (int f = 0; f < 5; f++) {
        T pR[9], pt[3], pJ[3];
#pragma unroll
        for (int k = 0; k < 9; k++)
            pR[k] = R0[k];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            pt[k] = J0[k];
            pJ[k] = J0[k];
        }
        for (int lev = 0; lev < 3; lev++) {
            const int j = 1 + 3 * f + lev; /* MANO order: joints 3f+1, 3f+2, 3f+3 form one finger */
            T R[9], J[3], jrel[3], nR[9], nt[3];
            getR(j, R);
            getJ(j, J);
#pragma unroll
            for (int k = 0; k < 3; k++)
                jrel[k] = J[k] - pJ[k];
            mano_compose(pR, pt, R, jrel, nR, nt);
            sink(j, nR, nt);
#pragma unroll
            for (int k = 0; k < 9; k++)
                pR[k] = nR[k];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                pt[k] = nt[k];
                pJ[k] = J[k];
            }
        }
    }
}

__constant__ int c_reorder_joints[21] = {0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20};

/* Per-sample state every CTA that works on the sample rebuilds for itself (a few hundred flops). */
struct ManoPose {
    float full_pose[48];
    float R[16][9];
    float pose_map[135];
    float betas[10];
    float J[16][3];
    float AR[16][9]; /* skinning transforms: rotation */
    float At[16][3]; /*                       translation (rest joint removed) */
    float Gt[16][3]; /* global joint positions */
    float tipvp[5][3]; /* posed rest position of the five fingertip vertices */
    float tip[5][3];   /* their skinned position */
    float centre[3];
};

#define MN_VS 64              /* vertices per CTA */
#define MN_CS (3 * MN_VS)     /* coordinates per CTA */
#define MN_ACC 352            /* per-sample accumulators of the backward: gA [16][12], gpm [135], gbetas [10], pad */

/* pose -> full pose -> 16 rotations, pose_map; betas -> J (J = j_template + j_shapedirs . betas: the joint
 * regression of the shaped template is linear in betas, so it is folded into two model constants). */
__device__ __forceinline__ void mano_pose_setup(const hoc_mano_model &M, const float *pose, const float *betas, int b,
                                                ManoPose &P)
{
    const int tid = threadIdx.x;
    const int npose = 3 + M.ncomps;
    if (tid < 48) {
        float p;
        if (tid < 3) {
            p = pose[(long)b * npose + tid];
        } else {
            const int t = tid - 3;
            p = M.hands_mean[t];
            if (M.use_pca) {
                for (int c = 0; c < M.ncomps; c++)
                    p += pose[(long)b * npose + 3 + c] * M.hands_components[c * 45 + t];
            } else {
                p += pose[(long)b * npose + 3 + t];
            }
        }
        P.full_pose[tid] = p;
    }
    if (tid >= 64 && tid < 74)
        P.betas[tid - 64] = (betas != nullptr) ? betas[(long)b * 10 + tid - 64] : 0.0f;
    __syncthreads();
    if (tid < 16) {
        float R[9];
        mano_rodrigues<float>(&P.full_pose[3 * tid], R);
#pragma unroll
        for (int k = 0; k < 9; k++) {
            P.R[tid][k] = R[k];
            if (tid > 0)
                P.pose_map[(tid - 1) * 9 + k] = R[k] - ((k == 0 || k == 4 || k == 8) ? 1.0f : 0.0f);
        }
    }
    if (tid >= 64 && tid < 112) {
        const int e = tid - 64; /* joint e / 3, coordinate e % 3 */
        float a = M.j_template[e];
#pragma unroll
        for (int k = 0; k < 10; k++)
            a += M.j_shapedirs[e * 10 + k] * P.betas[k];
        (&P.J[0][0])[e] = a;
    }
    __syncthreads();
}

/* forward kinematics: 16 small transforms, one thread */
__device__ __forceinline__ void mano_pose_fk(ManoPose &P)
{
    mano_fk<float>([&](int j, float *R) {
#pragma unroll
        for (int k = 0; k < 9; k++)
            R[k] = P.R[j][k]; },
                   [&](int j, float *J) {
#pragma unroll
        for (int k = 0; k < 3; k++)
            J[k] = P.J[j][k]; },
                   [&](int j, const float *GR, const float *Gt) {
#pragma unroll
        for (int k = 0; k < 9; k++)
            P.AR[j][k] = GR[k];
#pragma unroll
        for (int r = 0; r < 3; r++) {
            P.Gt[j][r] = Gt[r];
            P.At[j][r] = Gt[r] - (GR[3 * r] * P.J[j][0] + GR[3 * r + 1] * P.J[j][1] + GR[3 * r + 2] * P.J[j][2]);
        } });
}

/* v_posed coordinate i = template + shapedirs . betas + posedirs . pose_map, one thread, coalesced over i through the
 * transposed pose blend shapes [135][3V] */
__device__ __forceinline__ float mano_vposed_coord(const hoc_mano_model &M, const ManoPose &P, int i)
{
    const int n = 3 * M.num_verts;
    float a = M.v_template[i];
    const float *sd = M.shapedirs + (long)i * 10;
#pragma unroll
    for (int k = 0; k < 10; k++)
        a += sd[k] * P.betas[k];
    const float *pt = M.posedirs_t + i;
    float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    /* 9 batches of 15 independent loads (the reads hit L2: latency, not bandwidth, is what this loop waits for) */
#pragma unroll 1
    for (int k = 0; k < 135; k += 15) {
        float v[15];
#pragma unroll
        for (int u = 0; u < 15; u++)
            v[u] = __ldg(pt + (long)(k + u) * n);
#pragma unroll
        for (int u = 0; u < 15; u++)
            acc[u % 5] += v[u] * P.pose_map[k + u];
    }
    return a + ((acc[0] + acc[1]) + (acc[2] + acc[3]) + acc[4]);
}

/* blended transform of vertex v applied to its posed rest position (x, y, z) */
__device__ __forceinline__ void mano_skin(const hoc_mano_model &M, const ManoPose &P, int v, float x, float y, float z,
                                          float *out, float *TR)
{
    float T[12];
#pragma unroll
    for (int k = 0; k < 12; k++)
        T[k] = 0.0f;
    const float4 *w4 = reinterpret_cast<const float4 *>(M.weights + (long)v * 16);
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const float4 wq = __ldg(w4 + q);
        const float wv[4] = {wq.x, wq.y, wq.z, wq.w};
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int j = 4 * q + u;
#pragma unroll
            for (int k = 0; k < 9; k++)
                T[k] += wv[u] * P.AR[j][k];
#pragma unroll
            for (int k = 0; k < 3; k++)
                T[9 + k] += wv[u] * P.At[j][k];
        }
    }
#pragma unroll
    for (int r = 0; r < 3; r++)
        out[r] = T[3 * r] * x + T[3 * r + 1] * y + T[3 * r + 2] * z + T[9 + r];
    if (TR != nullptr) {
#pragma unroll
        for (int k = 0; k < 9; k++)
            TR[k] = T[k];
    }
}

/* Steps shared by the forward and the first backward kernel: pose set-up, then -- concurrently -- forward kinematics
 * (thread 0), the slice's posed rest positions (threads 32 .. 32 + MN_CS) and the fingertips' (15 threads of warp 7);
 * then the skinned fingertips and the centre. */
__device__ __forceinline__ void mano_slice_setup(const hoc_mano_model &M, const float *pose, const float *betas,
                                                 const float *trans, int b, int v0, ManoPose &P, float *s_vp)
{
    const int tid = threadIdx.x;
    const int V = M.num_verts;
    mano_pose_setup(M, pose, betas, b, P);
    if (tid == 0)
        mano_pose_fk(P);
    if (tid >= 32 && tid < 32 + MN_CS) {
        const int il = tid - 32, i = 3 * v0 + il;
        s_vp[il] = (i < 3 * V) ? mano_vposed_coord(M, P, i) : 0.0f;
    }
    if (tid >= 224 && tid < 224 + 15) { /* the fingertips' 15 coordinates, like any other coordinate */
        const int e = tid - 224;
        P.tipvp[e / 3][e % 3] = mano_vposed_coord(M, P, 3 * M.tip_ids[e / 3] + e % 3);
    }
    __syncthreads();
    if (tid < 5)
        mano_skin(M, P, M.tip_ids[tid], P.tipvp[tid][0], P.tipvp[tid][1], P.tipvp[tid][2], P.tip[tid], nullptr);
    __syncthreads();
    if (tid == 0) {
        /* centre: -trans, or reordered joint `center_idx` = original joint (< 16) or fingertip vertex */
        float c[3] = {0.f, 0.f, 0.f};
        if (trans != nullptr) {
#pragma unroll
            for (int k = 0; k < 3; k++)
                c[k] = -trans[(long)b * 3 + k];
        } else if (M.center_idx >= 0) {
            const int src = c_reorder_joints[M.center_idx];
#pragma unroll
            for (int k = 0; k < 3; k++)
                c[k] = (src < 16) ? P.Gt[src][k] : P.tip[src - 16][k];
        }
#pragma unroll
        for (int k = 0; k < 3; k++)
            P.centre[k] = c[k];
    }
    __syncthreads();
}

/* grid (ceil(V / MN_VS), B): CTA (s, b) produces vertices [s MN_VS, (s + 1) MN_VS) of sample b; CTA (0, b) also the
 * 21 joints. */
__global__ void __launch_bounds__(MN_THREADS)
hoc_mano_forward_kernel(hoc_mano_model M, const float *__restrict__ pose, const float *__restrict__ betas,
                        const float *__restrict__ trans, float *__restrict__ verts, float *__restrict__ joints)
{
    __shared__ ManoPose P;
    __shared__ float s_vp[MN_CS];
    const int b = blockIdx.y, tid = threadIdx.x;
    const int V = M.num_verts;
    const int v0 = blockIdx.x * MN_VS;
    mano_slice_setup(M, pose, betas, trans, b, v0, P, s_vp);
    if (tid < MN_VS && v0 + tid < V) {
        const int v = v0 + tid;
        float o[3];
        mano_skin(M, P, v, s_vp[3 * tid], s_vp[3 * tid + 1], s_vp[3 * tid + 2], o, nullptr);
#pragma unroll
        for (int k = 0; k < 3; k++)
            verts[((long)b * V + v) * 3 + k] = (o[k] - P.centre[k]) * 1000.0f;
    }
    if (blockIdx.x == 0 && tid >= 64 && tid < 85) {
        const int q = tid - 64, src = c_reorder_joints[q];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float o = (src < 16) ? P.Gt[src][k] : P.tip[src - 16][k];
            joints[((long)b * 21 + q) * 3 + k] = (o - P.centre[k]) * 1000.0f;
        }
    }
}

/* sum over all outputs of the incoming gradient (x 1000): every output had the same centre subtracted */
__device__ __forceinline__ void mano_grad_centre(const float *g_verts, const float *g_joints, int b, int V, float *s_red,
                                                 float *out3)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    float a[3] = {0.f, 0.f, 0.f};
    if (g_verts != nullptr)
        for (int i = tid; i < V * 3; i += blockDim.x)
            a[i % 3] += g_verts[(long)b * V * 3 + i];
    if (g_joints != nullptr && tid < 63)
        a[tid % 3] += g_joints[(long)b * 63 + tid];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        a[k] = hoc_warp_sum(a[k]);
        if (lane == 0)
            s_red[warp * 3 + k] = a[k];
    }
    __syncthreads();
    if (tid < 3) {
        float t = 0.0f;
        for (int w = 0; w < nw; w++)
            t += s_red[w * 3 + tid];
        out3[tid] = -1000.0f * t;
    }
    __syncthreads();
}

/*
 * Backward, kernel A.  grid (ceil(V / MN_VS), B): the vertex-parallel part for one slice of one sample -- skinning
 * adjoint (dL/dA_j, 192 sums), dL/d v_posed = T_v^T g_v, and its products with the pose / shape blend shapes
 * (dL/d pose_map, 135 sums; direct part of dL/d betas, 10 sums) -- reduced over the slice in the CTA and stored in
 * the slice's own row of partial sums [B][slices][MN_ACC]; kernel B adds the rows in slice order.  No atomics: the
 * MANO gradient is reproducible bit for bit.
 */
__global__ void __launch_bounds__(MN_THREADS)
hoc_mano_backward_verts_kernel(hoc_mano_model M, const float *__restrict__ pose, const float *__restrict__ betas,
                               const float *__restrict__ trans, const float *__restrict__ g_verts,
                               const float *__restrict__ g_joints, float *__restrict__ acc)
{
    __shared__ ManoPose P;
    __shared__ float s_vp[MN_CS], s_g[MN_CS], s_gvp[MN_CS];
    __shared__ float s_w[MN_VS * 16];
    __shared__ float s_red[(MN_THREADS / 32) * 3], s_gc[3];
    const int b = blockIdx.y, tid = threadIdx.x;
    const int V = M.num_verts;
    const int v0 = blockIdx.x * MN_VS;
    const int nv = min(MN_VS, V - v0);
    mano_slice_setup(M, pose, betas, trans, b, v0, P, s_vp);
    mano_grad_centre(g_verts, g_joints, b, V, s_red, s_gc);
    for (int i = tid; i < MN_VS * 16; i += MN_THREADS)
        s_w[i] = (i < nv * 16) ? M.weights[(long)v0 * 16 + i] : 0.0f;
    /* incoming gradient of the slice's vertices: x 1000, fingertip joints routed to their vertices, centre coupling */
    if (tid < MN_VS) {
        float g[3] = {0.f, 0.f, 0.f};
        if (tid < nv) {
            const int v = v0 + tid;
            if (g_verts != nullptr)
#pragma unroll
                for (int k = 0; k < 3; k++)
                    g[k] = 1000.0f * g_verts[((long)b * V + v) * 3 + k];
            const int csrc = (trans == nullptr && M.center_idx >= 0) ? c_reorder_joints[M.center_idx] : -1;
            for (int q = 0; q < 21; q++) {
                const int src = c_reorder_joints[q];
                if (src >= 16 && M.tip_ids[src - 16] == v && g_joints != nullptr)
#pragma unroll
                    for (int k = 0; k < 3; k++)
                        g[k] += 1000.0f * g_joints[((long)b * 21 + q) * 3 + k];
            }
            if (csrc >= 16 && M.tip_ids[csrc - 16] == v)
#pragma unroll
                for (int k = 0; k < 3; k++)
                    g[k] += s_gc[k];
            float o[3], TR[9];
            mano_skin(M, P, v, s_vp[3 * tid], s_vp[3 * tid + 1], s_vp[3 * tid + 2], o, TR);
#pragma unroll
            for (int c = 0; c < 3; c++)
                s_gvp[3 * tid + c] = TR[c] * g[0] + TR[3 + c] * g[1] + TR[6 + c] * g[2];
        } else {
#pragma unroll
            for (int c = 0; c < 3; c++)
                s_gvp[3 * tid + c] = 0.0f;
        }
#pragma unroll
        for (int k = 0; k < 3; k++)
            s_g[3 * tid + k] = g[k];
    }
    __syncthreads();
    float *A = acc + ((long)b * gridDim.x + blockIdx.x) * MN_ACC;
    if (tid < 192) { /* dL/dA_j = sum_v w_vj g_v (x) [v_posed; 1] */
        const int j = tid / 12, e = tid % 12; /* e < 9: rotation entry (r, c); e >= 9: translation r */
        const int r = e < 9 ? e / 3 : e - 9, c = e < 9 ? e % 3 : -1;
        float a = 0.0f;
        for (int v = 0; v < nv; v++)
            a += s_w[v * 16 + j] * s_g[3 * v + r] * (c >= 0 ? s_vp[3 * v + c] : 1.0f);
        A[tid] = a;
    } else if (tid < 192 + 10) { /* direct part of dL/d betas = shapedirs^T dL/d v_posed */
        const int k = tid - 192;
        float a = 0.0f, a2 = 0.0f;
        const float *sd = M.shapedirs + (long)3 * v0 * 10 + k;
        int i = 0;
        for (; i + 7 < 3 * nv; i += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; u++)
                v[u] = __ldg(sd + (long)(i + u) * 10);
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
                a += v[u] * s_gvp[i + u];
                a2 += v[u + 1] * s_gvp[i + u + 1];
            }
        }
        for (; i < 3 * nv; i++)
            a += __ldg(sd + (long)i * 10) * s_gvp[i];
        a += a2;
        A[192 + 135 + k] = a;
    }
    if (tid < 135) { /* dL/d pose_map[k] = sum_i posedirs[i][k] dL/d v_posed[i]  (coalesced across threads) */
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
        const float *pd = M.posedirs + (long)3 * v0 * 135 + tid;
        int i = 0;
        for (; i + 11 < 3 * nv; i += 12) { /* 12 independent loads per batch (L2 latency bound) */
            float v[12];
#pragma unroll
            for (int u = 0; u < 12; u++)
                v[u] = __ldg(pd + (long)(i + u) * 135);
#pragma unroll
            for (int u = 0; u < 12; u += 4) {
                a0 += v[u] * s_gvp[i + u];
                a1 += v[u + 1] * s_gvp[i + u + 1];
                a2 += v[u + 2] * s_gvp[i + u + 2];
                a3 += v[u + 3] * s_gvp[i + u + 3];
            }
        }
        for (; i < 3 * nv; i++)
            a0 += __ldg(pd + (long)i * 135) * s_gvp[i];
        a0 = (a0 + a1) + (a2 + a3);
        A[192 + tid] = a0;
    }
}

/*
 * Backward, kernel B.  grid (B), 128 threads: the per-sample part -- forward-mode duals through Rodrigues and the
 * kinematic chain (96 seeds: 48 pose + 48 joint coordinates, one per thread, through the SAME templated code as the
 * forward), then the PCA / shape maps.  Objective: <gA, A> + <gGt, Gt> + <gpm, R[1:]>.
 */
#define MN_THREADS_B 128
__global__ void __launch_bounds__(MN_THREADS_B)
hoc_mano_backward_pose_kernel(hoc_mano_model M, const float *__restrict__ pose, const float *__restrict__ betas,
                              const float *__restrict__ trans, const float *__restrict__ g_verts,
                              const float *__restrict__ g_joints, const float *__restrict__ acc, int n_slices,
                              float *__restrict__ g_pose, float *__restrict__ g_betas, float *__restrict__ g_trans)
{
    __shared__ ManoPose P;
    __shared__ float gAR[16][9], gAt[16][3], gGt[16][3], gpm[135], gfull[48], gJ[48], s_A[MN_ACC];
    __shared__ float s_red[(MN_THREADS_B / 32) * 3], s_gc[3];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int npose = 3 + M.ncomps;
    mano_pose_setup(M, pose, betas, b, P);
    for (int i = tid; i < 192 + 135 + 10; i += MN_THREADS_B) { /* the slices' partial sums, in slice order */
        float a = 0.0f;
        for (int sl = 0; sl < n_slices; sl++)
            a += acc[((long)b * n_slices + sl) * MN_ACC + i];
        s_A[i] = a;
    }
    __syncthreads();
    const float *A = s_A;
    for (int i = tid; i < 192; i += MN_THREADS_B) {
        const int j = i / 12, e = i % 12;
        if (e < 9)
            gAR[j][e] = A[i];
        else
            gAt[j][e - 9] = A[i];
    }
    for (int i = tid; i < 135; i += MN_THREADS_B)
        gpm[i] = A[192 + i];
    if (tid < 48)
        (&gGt[0][0])[tid] = 0.0f;
    mano_grad_centre(g_verts, g_joints, b, M.num_verts, s_red, s_gc); /* (syncs) */
    if (tid == 0) {
        if (g_joints != nullptr)
            for (int q = 0; q < 21; q++) {
                const int src = c_reorder_joints[q];
                if (src < 16)
                    for (int k = 0; k < 3; k++)
                        gGt[src][k] += 1000.0f * g_joints[((long)b * 21 + q) * 3 + k];
            }
        if (trans != nullptr) {
            if (g_trans != nullptr)
                for (int k = 0; k < 3; k++)
                    g_trans[(long)b * 3 + k] = -s_gc[k]; /* outputs = (x + trans) * 1000 */
        } else if (M.center_idx >= 0) {
            const int src = c_reorder_joints[M.center_idx];
            if (src < 16)
                for (int k = 0; k < 3; k++)
                    gGt[src][k] += s_gc[k];
        }
    }
    __syncthreads();
    if (tid < 96) {
        const int s = tid;
        const int jd = s < 48 ? s / 3 : -1;
        Dual Rd[9];
        if (jd >= 0) {
            Dual aa[3];
#pragma unroll
            for (int k = 0; k < 3; k++)
                aa[k] = mk(P.full_pose[3 * jd + k], (3 * jd + k == s) ? 1.0f : 0.0f);
            mano_rodrigues<Dual>(aa, Rd);
        }
        float a = 0.0f;
        if (jd >= 1) {
#pragma unroll
            for (int k = 0; k < 9; k++)
                a += gpm[(jd - 1) * 9 + k] * Rd[k].d;
        }
        mano_fk<Dual>([&](int j, Dual *R) {
#pragma unroll
            for (int k = 0; k < 9; k++)
                R[k] = (j == jd) ? Rd[k] : mk(P.R[j][k]); },
                      [&](int j, Dual *J) {
#pragma unroll
            for (int k = 0; k < 3; k++)
                J[k] = mk(P.J[j][k], (s >= 48 && s - 48 == 3 * j + k) ? 1.0f : 0.0f); },
                      [&](int j, const Dual *GR, const Dual *Gt) {
#pragma unroll
            for (int k = 0; k < 9; k++)
                a += gAR[j][k] * GR[k].d;
#pragma unroll
            for (int r = 0; r < 3; r++) {
                a += gGt[j][r] * Gt[r].d;
                /* At = Gt - GR J_j */
                Dual at = Gt[r];
#pragma unroll
                for (int c = 0; c < 3; c++)
                    at = at - GR[3 * r + c] * mk(P.J[j][c], (s >= 48 && s - 48 == 3 * j + c) ? 1.0f : 0.0f);
                a += gAt[j][r] * at.d;
            } });
        if (s < 48)
            gfull[s] = a;
        else
            gJ[s - 48] = a;
    }
    __syncthreads();
    /* dL/d betas = direct part (kernel A) + j_shapedirs^T dL/dJ */
    if (g_betas != nullptr && tid < 10) {
        float a = A[192 + 135 + tid];
        for (int e = 0; e < 48; e++)
            a += M.j_shapedirs[e * 10 + tid] * gJ[e];
        g_betas[(long)b * 10 + tid] = a;
    }
    /* dL/d pose */
    if (g_pose != nullptr && tid >= 32 && tid < 32 + npose) {
        const int t = tid - 32;
        float a;
        if (t < 3) {
            a = gfull[t];
        } else if (M.use_pca) {
            a = 0.0f;
            for (int u = 0; u < 45; u++)
                a += M.hands_components[(t - 3) * 45 + u] * gfull[3 + u];
        } else {
            a = gfull[t];
        }
        g_pose[(long)b * npose + t] = a;
    }
}

static int hoc_mano_check(const hoc_mano_model *m, int B, const char *who)
{
    HOC_CHECK_ARG(m != nullptr, "%s: model is NULL", who);
    HOC_CHECK_ARG(m->num_verts >= 1 && m->num_verts <= MN_MAXV, "%s: num_verts %d outside [1, %d]", who, m->num_verts,
                  MN_MAXV);
    HOC_CHECK_ARG(m->ncomps >= 1 && m->ncomps <= 45, "%s: ncomps %d outside [1, 45]", who, m->ncomps);
    HOC_CHECK_ARG(m->use_pca || m->ncomps == 45, "%s: axis-angle input needs ncomps == 45", who);
    HOC_CHECK_ARG(m->center_idx >= -1 && m->center_idx < 21, "%s: center_idx %d", who, m->center_idx);
    HOC_CHECK_ARG(m->v_template && m->shapedirs && m->posedirs && m->posedirs_t && m->j_template && m->j_shapedirs &&
                      m->weights && m->hands_mean && (m->hands_components || !m->use_pca),
                  "%s: model tensor is NULL", who);
    HOC_CHECK_ARG(((uintptr_t)m->weights & 15) == 0, "%s: weights must be 16-byte aligned", who);
    for (int k = 0; k < 5; k++)
        HOC_CHECK_ARG(m->tip_ids[k] >= 0 && m->tip_ids[k] < m->num_verts, "%s: tip vertex %d out of range", who,
                      m->tip_ids[k]);
    HOC_CHECK_ARG(B >= 0 && B <= 65535, "%s: batch %d", who, B);
    return HOC_OK;
}

extern "C" int hoc_mano_forward(const hoc_mano_model *model, const float *pose, const float *betas, const float *trans,
                                int B, float *verts, float *joints, void *stream)
{
    const int rc = hoc_mano_check(model, B, "hoc_mano_forward");
    if (rc != HOC_OK)
        return rc;
    if (B == 0)
        return HOC_OK;
    HOC_CHECK_ARG(pose && verts && joints, "hoc_mano_forward: NULL argument");
    dim3 grid((model->num_verts + MN_VS - 1) / MN_VS, B);
    HOC_LAUNCH(HOC_K_MANO_FWD, (cudaStream_t)stream,
               (hoc_mano_forward_kernel<<<grid, MN_THREADS, 0, (cudaStream_t)stream>>>(*model, pose, betas, trans, verts,
                                                                                       joints)));
    HOC_CHECK_LAUNCH("hoc_mano_forward_kernel");
    return HOC_OK;
}

extern "C" size_t hoc_mano_backward_workspace_bytes(int B)
{
    /* one row of partial sums per (sample, 64-vertex slice); sized for the largest model the kernels accept */
    return B > 0 ? sizeof(float) * MN_ACC * (size_t)B * ((MN_MAXV + MN_VS - 1) / MN_VS) : 0;
}

extern "C" int hoc_mano_backward(const hoc_mano_model *model, const float *pose, const float *betas,
                                 const float *trans, const float *grad_verts, const float *grad_joints, int B,
                                 float *grad_pose, float *grad_betas, float *grad_trans, void *workspace,
                                 size_t workspace_bytes, void *stream)
{
    const int rc = hoc_mano_check(model, B, "hoc_mano_backward");
    if (rc != HOC_OK)
        return rc;
    if (B == 0)
        return HOC_OK;
    HOC_CHECK_ARG(pose != nullptr, "hoc_mano_backward: pose is NULL");
    const size_t need = hoc_mano_backward_workspace_bytes(B);
    if (workspace == nullptr || workspace_bytes < need) {
        hoc_set_error("hoc_mano_backward: workspace of %zu bytes needed, %zu given", need, workspace_bytes);
        return HOC_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    float *acc = (float *)workspace; /* every (sample, slice) row is fully written by kernel A: no zero-fill */
    dim3 grid((model->num_verts + MN_VS - 1) / MN_VS, B);
    HOC_LAUNCH(HOC_K_MANO_BWD, st,
               (hoc_mano_backward_verts_kernel<<<grid, MN_THREADS, 0, st>>>(*model, pose, betas, trans, grad_verts,
                                                                            grad_joints, acc)));
    HOC_CHECK_LAUNCH("hoc_mano_backward_verts_kernel");
    HOC_LAUNCH(HOC_K_MANO_BWD, st,
               (hoc_mano_backward_pose_kernel<<<B, MN_THREADS_B, 0, st>>>(*model, pose, betas, trans, grad_verts,
                                                                          grad_joints, acc, (int)grid.x, grad_pose,
                                                                          grad_betas, grad_trans)));
    HOC_CHECK_LAUNCH("hoc_mano_backward_pose_kernel");
    return HOC_OK;
}
