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
 * In torch this is ~60 tiny launches (batched 4x4 matmuls, cats, index selects); here ONE CTA per sample
 * does the whole forward, and one CTA per sample the whole backward: the big linear maps (posedirs,
 * shapedirs, J_regressor, skinning weights) are reversed by hand, the 16-joint kinematic chain and the
 * Rodrigues formula are differentiated with forward-mode dual numbers (96 seeds: 48 pose + 48 joint
 * coordinates, one per thread) through the SAME templated code the forward runs.
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

struct ManoShared {
    float full_pose[48];
    float R[16][9];
    float pose_map[135];
    float betas[10];
    float J[16][3];
    float AR[16][9]; /* skinning transforms: rotation */
    float At[16][3]; /*                       translation (rest joint removed) */
    float Gt[16][3]; /* global joint positions */
    float centre[3];
    float vs[MN_MAXV * 3]; /* v_shaped, then v_posed */
};

__constant__ int c_reorder_joints[21] = {0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20};

/* Steps shared by forward and backward: everything up to the skinning transforms, in shared memory. */
__device__ void mano_prepare(const hoc_mano_model &M, const float *pose, const float *betas, int b, ManoShared &S)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int V = M.num_verts;
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
        S.full_pose[tid] = p;
    }
    if (tid >= 64 && tid < 74)
        S.betas[tid - 64] = (betas != nullptr) ? betas[(long)b * 10 + tid - 64] : 0.0f;
    __syncthreads();
    if (tid < 16) {
        float R[9];
        mano_rodrigues<float>(&S.full_pose[3 * tid], R);
#pragma unroll
        for (int k = 0; k < 9; k++) {
            S.R[tid][k] = R[k];
            if (tid > 0)
                S.pose_map[(tid - 1) * 9 + k] = R[k] - ((k == 0 || k == 4 || k == 8) ? 1.0f : 0.0f);
        }
    }
    /* v_shaped */
    for (int i = tid; i < V * 3; i += MN_THREADS) {
        float a = M.v_template[i];
        const float *sd = M.shapedirs + (long)i * 10;
#pragma unroll
        for (int k = 0; k < 10; k++)
            a += sd[k] * S.betas[k];
        S.vs[i] = a;
    }
    __syncthreads();
    /* J = J_regressor . v_shaped : warp w handles joints w, w + 8 */
    for (int j = warp; j < 16; j += MN_THREADS / 32) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (int v = lane; v < V; v += 32) {
            const float r = M.j_regressor[(long)j * V + v];
            a0 += r * S.vs[3 * v];
            a1 += r * S.vs[3 * v + 1];
            a2 += r * S.vs[3 * v + 2];
        }
        a0 = hoc_warp_sum(a0);
        a1 = hoc_warp_sum(a1);
        a2 = hoc_warp_sum(a2);
        if (lane == 0) {
            S.J[j][0] = a0;
            S.J[j][1] = a1;
            S.J[j][2] = a2;
        }
    }
    __syncthreads();
    /* v_posed = v_shaped + posedirs . pose_map : one warp per output coordinate, coalesced over the 135 */
    for (int i = warp; i < V * 3; i += MN_THREADS / 32) {
        const float *pd = M.posedirs + (long)i * 135;
        float a = 0.f;
        for (int k = lane; k < 135; k += 32)
            a += pd[k] * S.pose_map[k];
        a = hoc_warp_sum(a);
        if (lane == 0)
            S.vs[i] += a;
    }
    /* forward kinematics (thread 0; 16 small transforms) */
    if (tid == 0) {
        mano_fk<float>([&](int j, float *R) {
#pragma unroll
            for (int k = 0; k < 9; k++)
                R[k] = S.R[j][k]; },
                       [&](int j, float *J) {
#pragma unroll
            for (int k = 0; k < 3; k++)
                J[k] = S.J[j][k]; },
                       [&](int j, const float *GR, const float *Gt) {
#pragma unroll
            for (int k = 0; k < 9; k++)
                S.AR[j][k] = GR[k];
#pragma unroll
            for (int r = 0; r < 3; r++) {
                S.Gt[j][r] = Gt[r];
                S.At[j][r] = Gt[r] - (GR[3 * r] * S.J[j][0] + GR[3 * r + 1] * S.J[j][1] + GR[3 * r + 2] * S.J[j][2]);
            } });
    }
    __syncthreads();
}

/* blended transform of vertex v applied to its posed rest position */
__device__ __forceinline__ void mano_skin(const hoc_mano_model &M, const ManoShared &S, int v, float *out, float *TR)
{
    float T[12];
#pragma unroll
    for (int k = 0; k < 12; k++)
        T[k] = 0.0f;
    const float *w = M.weights + (long)v * 16;
    for (int j = 0; j < 16; j++) {
        const float wj = w[j];
#pragma unroll
        for (int k = 0; k < 9; k++)
            T[k] += wj * S.AR[j][k];
#pragma unroll
        for (int k = 0; k < 3; k++)
            T[9 + k] += wj * S.At[j][k];
    }
    const float x = S.vs[3 * v], y = S.vs[3 * v + 1], z = S.vs[3 * v + 2];
#pragma unroll
    for (int r = 0; r < 3; r++)
        out[r] = T[3 * r] * x + T[3 * r + 1] * y + T[3 * r + 2] * z + T[9 + r];
    if (TR != nullptr) {
#pragma unroll
        for (int k = 0; k < 9; k++)
            TR[k] = T[k];
    }
}

__global__ void __launch_bounds__(MN_THREADS)
hoc_mano_forward_kernel(hoc_mano_model M, const float *__restrict__ pose, const float *__restrict__ betas,
                        const float *__restrict__ trans, float *__restrict__ verts, float *__restrict__ joints)
{
    extern __shared__ unsigned char smem_raw[];
    ManoShared &S = *reinterpret_cast<ManoShared *>(smem_raw);
    const int b = blockIdx.x, tid = threadIdx.x;
    const int V = M.num_verts;
    mano_prepare(M, pose, betas, b, S);
    /* centre: reordered joint `center_idx` = original joint (< 16) or fingertip vertex */
    if (tid == 0) {
        float c[3] = {0.f, 0.f, 0.f};
        if (trans != nullptr) {
#pragma unroll
            for (int k = 0; k < 3; k++)
                c[k] = -trans[(long)b * 3 + k];
        } else if (M.center_idx >= 0) {
            const int src = c_reorder_joints[M.center_idx];
            if (src < 16) {
#pragma unroll
                for (int k = 0; k < 3; k++)
                    c[k] = S.Gt[src][k];
            } else {
                mano_skin(M, S, M.tip_ids[src - 16], c, nullptr);
            }
        }
#pragma unroll
        for (int k = 0; k < 3; k++)
            S.centre[k] = c[k];
    }
    __syncthreads();
    for (int v = tid; v < V; v += MN_THREADS) {
        float o[3];
        mano_skin(M, S, v, o, nullptr);
#pragma unroll
        for (int k = 0; k < 3; k++)
            verts[((long)b * V + v) * 3 + k] = (o[k] - S.centre[k]) * 1000.0f;
    }
    if (tid < 21) {
        const int src = c_reorder_joints[tid];
        float o[3];
        if (src < 16) {
#pragma unroll
            for (int k = 0; k < 3; k++)
                o[k] = S.Gt[src][k];
        } else {
            mano_skin(M, S, M.tip_ids[src - 16], o, nullptr);
        }
#pragma unroll
        for (int k = 0; k < 3; k++)
            joints[((long)b * 21 + tid) * 3 + k] = (o[k] - S.centre[k]) * 1000.0f;
    }
}

struct ManoBwdShared {
    float gv[MN_MAXV * 3];  /* dL/d(skinned vertex), then dL/d v_posed */
    float gAR[16][9], gAt[16][3], gGt[16][3];
    float gpm[135];         /* dL/d pose_map */
    float gfull[48];        /* dL/d full_pose */
    float gJ[16][3];
    float gcentre[3];
    float gbetas[10];
};

__global__ void __launch_bounds__(MN_THREADS)
hoc_mano_backward_kernel(hoc_mano_model M, const float *__restrict__ pose, const float *__restrict__ betas,
                         const float *__restrict__ trans, const float *__restrict__ g_verts,
                         const float *__restrict__ g_joints, float *__restrict__ g_pose, float *__restrict__ g_betas,
                         float *__restrict__ g_trans)
{
    extern __shared__ unsigned char smem_raw[];
    ManoShared &S = *reinterpret_cast<ManoShared *>(smem_raw);
    ManoBwdShared &G = *reinterpret_cast<ManoBwdShared *>(smem_raw + ((sizeof(ManoShared) + 15) & ~(size_t)15));
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int V = M.num_verts;
    const int npose = 3 + M.ncomps;
    mano_prepare(M, pose, betas, b, S);

    /* B1: incoming gradients (x1000, minus-centre coupling), joints routed to transforms / tip vertices */
    for (int i = tid; i < V * 3; i += MN_THREADS)
        G.gv[i] = (g_verts != nullptr) ? 1000.0f * g_verts[(long)b * V * 3 + i] : 0.0f;
    if (tid < 48) {
        (&G.gGt[0][0])[tid] = 0.0f;
        (&G.gJ[0][0])[tid] = 0.0f;
    }
    if (tid < 3)
        G.gcentre[tid] = 0.0f;
    __syncthreads();
    if (tid == 0 && g_joints != nullptr) {
        for (int q = 0; q < 21; q++) {
            const int src = c_reorder_joints[q];
            for (int k = 0; k < 3; k++) {
                const float g = 1000.0f * g_joints[((long)b * 21 + q) * 3 + k];
                if (src < 16)
                    G.gGt[src][k] += g;
                else
                    G.gv[3 * M.tip_ids[src - 16] + k] += g;
                G.gcentre[k] -= g;
            }
        }
    }
    __syncthreads();
    {   /* dL/d centre -= sum of vertex gradients (the raw incoming ones, before the tip routing matters not:
           every output had the same centre subtracted) */
        float a[3] = {0.f, 0.f, 0.f};
        if (g_verts != nullptr)
            for (int v = tid; v < V; v += MN_THREADS)
#pragma unroll
                for (int k = 0; k < 3; k++)
                    a[k] += 1000.0f * g_verts[((long)b * V + v) * 3 + k];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            a[k] = hoc_warp_sum(a[k]);
            if (lane == 0 && a[k] != 0.0f)
                atomicAdd(&G.gcentre[k], -a[k]);
        }
    }
    __syncthreads();
    if (tid == 0) {
        if (trans != nullptr) {
            if (g_trans != nullptr)
#pragma unroll
                for (int k = 0; k < 3; k++)
                    g_trans[(long)b * 3 + k] = -G.gcentre[k]; /* outputs = (x + trans) * 1000 */
        } else if (M.center_idx >= 0) {
            const int src = c_reorder_joints[M.center_idx];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                if (src < 16)
                    G.gGt[src][k] += G.gcentre[k];
                else
                    G.gv[3 * M.tip_ids[src - 16] + k] += G.gcentre[k];
            }
        }
    }
    __syncthreads();

    /* B2: skinning adjoint.  dL/dA_j = sum_v w_vj g_v (x) [v_posed; 1]  (192 sums, one thread each) */
    if (tid < 192) {
        const int j = tid / 12, e = tid % 12; /* e < 9: rotation entry (r, c); e >= 9: translation r */
        const int r = e < 9 ? e / 3 : e - 9, c = e < 9 ? e % 3 : -1;
        float a = 0.0f;
        for (int v = 0; v < V; v++) {
            const float w = M.weights[(long)v * 16 + j];
            a += w * G.gv[3 * v + r] * (c >= 0 ? S.vs[3 * v + c] : 1.0f);
        }
        if (e < 9)
            G.gAR[j][e] = a;
        else
            G.gAt[j][r] = a;
    }
    __syncthreads();
    /* dL/d v_posed = T_v^T g_v (in place) */
    for (int v = tid; v < V; v += MN_THREADS) {
        float o[3], TR[9];
        mano_skin(M, S, v, o, TR);
        const float g0 = G.gv[3 * v], g1 = G.gv[3 * v + 1], g2 = G.gv[3 * v + 2];
#pragma unroll
        for (int c = 0; c < 3; c++)
            G.gv[3 * v + c] = TR[c] * g0 + TR[3 + c] * g1 + TR[6 + c] * g2;
    }
    __syncthreads();
    /* B3a: dL/d pose_map[k] = sum_i posedirs[i][k] g_vposed[i]  (thread k, coalesced across threads) */
    if (tid < 135) {
        float a = 0.0f;
        for (int i = 0; i < V * 3; i++)
            a += M.posedirs[(long)i * 135 + tid] * G.gv[i];
        G.gpm[tid] = a;
    }
    __syncthreads();

    /* B4/B5: forward-mode duals through Rodrigues + kinematic chain.  Seed s < 48: full_pose[s];
     * s >= 48: J[(s-48)/3][(s-48)%3].  Objective: <gA, A> + <gGt, Gt> + <gpm, R[1:]>. */
    if (tid < 96) {
        const int s = tid;
        const int jd = s < 48 ? s / 3 : -1;
        Dual Rd[9];
        if (jd >= 0) {
            Dual aa[3];
#pragma unroll
            for (int k = 0; k < 3; k++)
                aa[k] = mk(S.full_pose[3 * jd + k], (3 * jd + k == s) ? 1.0f : 0.0f);
            mano_rodrigues<Dual>(aa, Rd);
        }
        float acc = 0.0f;
        if (jd >= 1) {
#pragma unroll
            for (int k = 0; k < 9; k++)
                acc += G.gpm[(jd - 1) * 9 + k] * Rd[k].d;
        }
        mano_fk<Dual>([&](int j, Dual *R) {
#pragma unroll
            for (int k = 0; k < 9; k++)
                R[k] = (j == jd) ? Rd[k] : mk(S.R[j][k]); },
                      [&](int j, Dual *J) {
#pragma unroll
            for (int k = 0; k < 3; k++)
                J[k] = mk(S.J[j][k], (s >= 48 && s - 48 == 3 * j + k) ? 1.0f : 0.0f); },
                      [&](int j, const Dual *GR, const Dual *Gt) {
#pragma unroll
            for (int k = 0; k < 9; k++)
                acc += G.gAR[j][k] * GR[k].d;
#pragma unroll
            for (int r = 0; r < 3; r++) {
                acc += G.gGt[j][r] * Gt[r].d;
                /* At = Gt - GR J_j */
                Dual at = Gt[r];
#pragma unroll
                for (int c = 0; c < 3; c++)
                    at = at - GR[3 * r + c] * mk(S.J[j][c], (s >= 48 && s - 48 == 3 * j + c) ? 1.0f : 0.0f);
                acc += G.gAt[j][r] * at.d;
            } });
        if (s < 48)
            G.gfull[s] = acc;
        else
            (&G.gJ[0][0])[s - 48] = acc;
    }
    __syncthreads();

    /* B3b: dL/d v_shaped = dL/d v_posed + J_regressor^T dL/dJ ; dL/d betas = shapedirs^T dL/d v_shaped */
    if (g_betas != nullptr) {
        float a[10];
#pragma unroll
        for (int k = 0; k < 10; k++)
            a[k] = 0.0f;
        for (int i = tid; i < V * 3; i += MN_THREADS) {
            const int v = i / 3, c = i - 3 * v;
            float g = G.gv[i];
            for (int j = 0; j < 16; j++)
                g += M.j_regressor[(long)j * V + v] * G.gJ[j][c];
            const float *sd = M.shapedirs + (long)i * 10;
#pragma unroll
            for (int k = 0; k < 10; k++)
                a[k] += sd[k] * g;
        }
        if (tid < 10)
            G.gbetas[tid] = 0.0f;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 10; k++) {
            a[k] = hoc_warp_sum(a[k]);
            if (lane == 0)
                atomicAdd(&G.gbetas[k], a[k]);
        }
        __syncthreads();
        if (tid < 10)
            g_betas[(long)b * 10 + tid] = G.gbetas[tid];
    }
    /* B6: dL/d pose */
    if (g_pose != nullptr && tid < npose) {
        float a;
        if (tid < 3) {
            a = G.gfull[tid];
        } else if (M.use_pca) {
            a = 0.0f;
            for (int t = 0; t < 45; t++)
                a += M.hands_components[(tid - 3) * 45 + t] * G.gfull[3 + t];
        } else {
            a = G.gfull[tid];
        }
        g_pose[(long)b * npose + tid] = a;
    }
    (void)warp;
}

static size_t hoc_mano_smem(bool backward)
{
    size_t n = (sizeof(ManoShared) + 15) & ~(size_t)15;
    if (backward)
        n += sizeof(ManoBwdShared);
    return n;
}

static int hoc_mano_check(const hoc_mano_model *m, int B, const char *who)
{
    HOC_CHECK_ARG(m != nullptr, "%s: model is NULL", who);
    HOC_CHECK_ARG(m->num_verts >= 1 && m->num_verts <= MN_MAXV, "%s: num_verts %d outside [1, %d]", who, m->num_verts,
                  MN_MAXV);
    HOC_CHECK_ARG(m->ncomps >= 1 && m->ncomps <= 45, "%s: ncomps %d outside [1, 45]", who, m->ncomps);
    HOC_CHECK_ARG(m->use_pca || m->ncomps == 45, "%s: axis-angle input needs ncomps == 45", who);
    HOC_CHECK_ARG(m->center_idx >= -1 && m->center_idx < 21, "%s: center_idx %d", who, m->center_idx);
    HOC_CHECK_ARG(m->v_template && m->shapedirs && m->posedirs && m->j_regressor && m->weights && m->hands_mean &&
                      (m->hands_components || !m->use_pca),
                  "%s: model tensor is NULL", who);
    for (int k = 0; k < 5; k++)
        HOC_CHECK_ARG(m->tip_ids[k] >= 0 && m->tip_ids[k] < m->num_verts, "%s: tip vertex %d out of range", who,
                      m->tip_ids[k]);
    HOC_CHECK_ARG(B >= 0 && B <= 1000000, "%s: batch %d", who, B);
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
    const size_t smem = hoc_mano_smem(false); /* ~15 KB, under the 48 KB default limit */
    HOC_LAUNCH(HOC_K_MANO_FWD, (cudaStream_t)stream,
               (hoc_mano_forward_kernel<<<B, MN_THREADS, smem, (cudaStream_t)stream>>>(*model, pose, betas, trans, verts,
                                                                                       joints)));
    HOC_CHECK_LAUNCH("hoc_mano_forward_kernel");
    return HOC_OK;
}

extern "C" int hoc_mano_backward(const hoc_mano_model *model, const float *pose, const float *betas,
                                 const float *trans, const float *grad_verts, const float *grad_joints, int B,
                                 float *grad_pose, float *grad_betas, float *grad_trans, void *stream)
{
    const int rc = hoc_mano_check(model, B, "hoc_mano_backward");
    if (rc != HOC_OK)
        return rc;
    if (B == 0)
        return HOC_OK;
    HOC_CHECK_ARG(pose != nullptr, "hoc_mano_backward: pose is NULL");
    const size_t smem = hoc_mano_smem(true); /* ~29 KB */
    HOC_LAUNCH(HOC_K_MANO_BWD, (cudaStream_t)stream,
               (hoc_mano_backward_kernel<<<B, MN_THREADS, smem, (cudaStream_t)stream>>>(
                   *model, pose, betas, trans, grad_verts, grad_joints, grad_pose, grad_betas, grad_trans)));
    HOC_CHECK_LAUNCH("hoc_mano_backward_kernel");
    return HOC_OK;
}
