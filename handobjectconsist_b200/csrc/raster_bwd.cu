/*
 * raster_bwd.cu -- rasterizer backward for sm_100a.
 *
 * Replaces backward_pixel_map, backward_textures and backward_depth_map of
 * `neural_renderer.cuda.rasterize` (bound at /root/reference/meshreg/neurender/rasterize.py:
 * 269-281, 290-297, 306-315) and the zero-fills of rasterize.py:151-181.
 *
 * The reference runs the pseudo-gradient as ONE SERIAL THREAD PER FACE that walks whole image rows and
 * columns (O(edge length x S) per face, every pixel fetched from global memory again by every scan that
 * passes over it) and scatters texture / depth gradients with one thread per pixel.  Here the work is
 * split by what it is parallel over:
 *
 *   hoc_raster_bwd_scan_kernel    pixel-parallel, streaming (one pass over face_index_map and the incoming gradients;
 *   (_scan4_, _scan_pair_)        in the frame-pair path the pass COMPUTES the incoming gradient: the backward of
 *           pair_consist is fused in): per-line spans of the pixels that matter -- covered ones, which own the scans,
 *           and those with a non-zero incoming gradient, the only ones a scan gets a term from (scans are clipped to
 *           the span: exact) --, the list of pixels with a texture / depth gradient when a cover pass follows, and the
 *           zero-fill of the accumulated outputs.
 *   hoc_raster_bwd_line_kernel    line-parallel, the pseudo-gradient run FROM THE PIXELS: every term of
 *           backward_pixel_map lives on one image column or row, so a CTA stages its line once (owning face, colour,
 *           incoming gradient) and runs one candidate per (covered pixel, edge of its face): that column of the edge
 *           is evaluated once, the pixel adds its own term of the short INWARD scan (the reference visits exactly the
 *           owned pixels between the edge and the opposite edge) and, if it is the pixel just inside the edge, the
 *           lane holds the long OUTWARD scan that starts there; the warp's scans are cut into 16-pixel chunks summed
 *           out of shared memory one per lane.  There is no per-face pass: a face that owns no pixel has no term in
 *           either scan.  Its row CTAs also run backward_textures when the textures are three vertex values per
 *           face and the forward saved weights and depth (the frame-pair path).
 *   hoc_raster_bwd_cover_kernel   otherwise: texture / depth gradient of the listed pixels, spread evenly over the GPU
 *           -- weights and depth come from the forward's maps or are recomputed with the forward's functions
 *           (bit-identical); the reference's two sampling maps (64 B/px) are never stored.
 *   hoc_raster_bwd_depth_kernel   face-parallel epilogue of backward_depth_map (only when dL/ddepth exists).
 *
 * Measured progression on bench.py's workload (16 renders of 9104 faces at 256 x 256 with the pseudo-gradient + 16
 * with the texture gradient only, B200, in-graph): 548 us (one warp per face, round 1) -> 245 -> 121 -> 93 (pixel /
 * face / line passes) -> 55 + 25 us (scan, cover with scan queues, queued line pass; the two renders on two streams)
 * -> 60 us for both renders in one batch -> 48 us (scan with the pair_consist backward inside + this line pass).
 */
#include "hoc_common.cuh"
#include "hoc_det.cuh"
#include "raster_math.h"
#include "warp_math.cuh"

#define HOC_SCAN_SPAN_ALL 1
#define HOC_SCAN_NO_LIST 2
#define HOC_SCAN_TWO_CHANNELS 4 /* frame-pair path: the third gradient plane (always zero) is neither written nor read */
#define EXT_ROW_LO 0
#define EXT_ROW_HI 1
#define EXT_COL_LO 2
#define EXT_COL_HI 3

/* Workspace carved by hoc_raster_backward (256-byte aligned regions):
 *   ext        int    [B][4][S]      {row_nlo, row_hi, col_nlo, col_hi}: span of the pixels of a line that matter, stored
 *                                    as S - lo and hi + 1 so that both ends are max-reduced from 0    (zero-filled)
 *   cov_count  int    [B]            pixels listed per sample                                         (zero-filled)
 *   acc_d      float  [B][F][3]      sum over owned pixels of dL/ddepth * depth^2 * w_k   (zero-filled, depth gradient only)
 *   cov_list   int2   [B][S*S]       the pixels with a texture / depth gradient, in tile order: (yi * S + xi, owning
 *                                    face) -- the face rides along so that the cover pass starts one dependent load
 *                                    earlier; only the used part is ever touched */
struct HocBwdWorkspace {
    int *ext;
    int *cov_count;
    float *acc_d;
    int2 *cov_list;
    /* reproducible mode only (hoc_det.cuh): fixed-point accumulators of grad_faces / grad_textures / acc_d,
     * two 64-bit words per float, one contiguous zero-fill */
    unsigned long long *det_gf, *det_gt, *det_ad;
    size_t det_bytes;
    size_t count_bytes, acc_bytes; /* ext + cov_count, then acc_d: one contiguous zero-fill */
    size_t total;
};

static HocBwdWorkspace hoc_bwd_workspace(void *base, int B, int F, int S, int tex_n = 0, bool det = false)
{
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    HocBwdWorkspace w;
    size_t off = 0;
    char *p = (char *)base;
    const size_t zero_begin = off;
    w.ext = (int *)(p + off);
    off = up(off + sizeof(int) * 4 * (size_t)B * S);
    w.cov_count = (int *)(p + off);
    off = up(off + sizeof(int) * (size_t)B);
    w.count_bytes = off - zero_begin;
    w.acc_d = (float *)(p + off); /* directly after the counters: one memset covers both */
    w.acc_bytes = sizeof(float) * 3 * (size_t)B * F;
    off = up(off + w.acc_bytes);
    w.cov_list = (int2 *)(p + off);
    off = up(off + sizeof(int2) * (size_t)B * S * S);
    w.det_gf = w.det_gt = w.det_ad = nullptr;
    w.det_bytes = 0;
    if (det) {
        const size_t nf = (size_t)B * F, begin = off;
        w.det_gf = (unsigned long long *)(p + off);
        off += 16 * 9 * nf;
        w.det_gt = (unsigned long long *)(p + off);
        off += 16 * (size_t)tex_n * nf;
        w.det_ad = (unsigned long long *)(p + off);
        off = up(off + 16 * 3 * nf);
        w.det_bytes = off - begin;
    }
    w.total = off;
    return w;
}

/* ------------------------------------------------------------------------------------------ */
struct HocBwdMaps {
    const int32_t *idx;   /* [S,S] of this sample, raster order */
    const float *rgb;     /* forward output, `layout`, base of the whole tensor */
    const float *g_rgb;   /* incoming gradient, `layout`, or NULL */
    const float *g_alpha; /* or NULL */
    int S, layout, b;
    bool use_alpha;       /* the alpha term exists (return_alpha and g_alpha given) */
    bool use_rgb;         /* the rgb term exists (return_rgb and g_rgb given) */
};

/* I(.) of the reference at a pixel: (alpha, r, g, b). */
__device__ __forceinline__ void hoc_load_I(const HocBwdMaps &M, int xi, int yi, float *I)
{
    I[0] = I[1] = I[2] = I[3] = 0.0f;
    if (M.use_alpha)
        I[0] = (M.idx[(long)yi * M.S + xi] >= 0) ? 1.0f : 0.0f;
    if (M.use_rgb) {
        I[1] = M.rgb[hoc_rgb_off(M.layout, M.S, M.b, yi, xi, 0)];
        I[2] = M.rgb[hoc_rgb_off(M.layout, M.S, M.b, yi, xi, 1)];
        I[3] = M.rgb[hoc_rgb_off(M.layout, M.S, M.b, yi, xi, 2)];
    }
}

/* 1 / x, one MUFU.RCP. */
__device__ __forceinline__ float hoc_rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

/* One (edge, axis, column) of the face owning a pixel: what the line pass has evaluated for it (hoc_k4_edge_pts,
 * hoc_k4_column) when it adds the pixel's own term of the short INWARD scan of that column (the reference visits
 * exactly the owned pixels between the edge and the opposite edge): delta from the pixel and the pixel just outside the
 * edge, -delta / dist to the two vertices of the edge.  gfA / gfB index the x component of vertex A / B in grad_faces;
 * I / g: (alpha, r, g, b) of the pixel and its incoming gradient. */
struct HocK4Stage {
    HocK4Edge E;
    float d1_cross;
    int d0, d1p;
};

__device__ __forceinline__ void hoc_k4_stage_b(const HocK4Stage &T, int axis, const HocBwdMaps &M, const float *I,
                                               const float *I_out, const float *g, float eps,
                                               float *__restrict__ grad_faces, long gfA, long gfB,
                                               unsigned long long *__restrict__ det_gf)
{
    float delta = 0.0f; /* the reference's accumulation order: alpha, r, g, b */
    if (M.use_alpha)
        delta += (I[0] - I_out[0]) * g[0];
    if (M.use_rgb) {
#pragma unroll
        for (int k = 1; k < 4; k++)
            delta += (I[k] - I_out[k]) * g[k];
    }
    if (!(delta <= 0.0f)) {
        HocK4Col C;
        hoc_k4_col(&T.E, M.S, T.d0, T.d1_cross, &C);
        float gA = 0.0f, gB = 0.0f;
        hoc_k4_accum_col(&C, T.d1p, eps, delta, &gA, &gB);
        if (gA != 0.0f)
            hoc_accum(grad_faces, gfA + (1 - axis), gA, det_gf);
        if (gB != 0.0f)
            hoc_accum(grad_faces, gfB + (1 - axis), gB, det_gf);
    }
}

/*
 * Scan pass.  Block (32, 8) covers a 32 x 32 pixel tile (4 rows per thread).  Pure streaming: reads
 * face_index_map and the incoming gradients once, writes the line spans and -- unless nobody reads it -- the list of
 * the pixels with a texture / depth gradient (`list_all`: every covered pixel).  One global atomic per CTA reserves
 * the tile's list slots.
 */
__global__ void __launch_bounds__(256)
hoc_raster_bwd_scan_kernel(const int32_t *__restrict__ face_index_map, const float *__restrict__ g_rgb_k4,
                           const float *__restrict__ g_rgb_rest, const float *__restrict__ g_alpha_k4, int S, int layout,
                           int k4_samples, int list_all, int scan_flags, int *__restrict__ ext,
                           int *__restrict__ cov_count,
                           int2 *__restrict__ cov_list, float *__restrict__ zero_a, long n_a,
                           float *__restrict__ zero_b, long n_b, float *__restrict__ zero_c, long n_c,
                           const int *__restrict__ row_lo)
{
    /* samples [0, k4_samples) get the pseudo-gradient (spans + every covered pixel listed); the others only list the
     * pixels that have a texture (non-zero dL/drgb) or depth gradient */
    const bool K4 = (int)blockIdx.z < k4_samples;
    const float *g_rgb = K4 ? g_rgb_k4 : g_rgb_rest;
    const float *g_alpha = K4 ? g_alpha_k4 : nullptr;
    /* scan_flags: HOC_SCAN_SPAN_ALL = line spans for every sample (the fused line pass also runs the texture gradient of
     * the samples without pseudo-gradient, row by row), HOC_SCAN_NO_LIST = nobody reads the list of covered pixels */
    const bool SPAN = K4 || (scan_flags & HOC_SCAN_SPAN_ALL);
    const bool NO_LIST = (scan_flags & HOC_SCAN_NO_LIST) != 0;
    { /* zero-fill of the two gradient outputs (accumulated with atomics by the later passes), spread over the grid */
        const long nthreads = (long)gridDim.x * gridDim.y * gridDim.z * 256;
        const long t0 = (((long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 256 + threadIdx.y * 32 +
                        threadIdx.x;
        hoc_zero_floats(zero_a, n_a, t0, nthreads);
        hoc_zero_floats(zero_b, n_b, t0, nthreads);
        hoc_zero_floats(zero_c, n_c, t0, nthreads); /* (a buffer of the caller's NEXT kernel: hoc_mesh_scatter's outputs) */
    }
    __shared__ int s_lo[8][32];
    __shared__ int s_hi[8][32];
    __shared__ int s_cnt[33]; /* per warp-row (r * 8 + ty) count, then exclusive prefix; [32] = tile base */
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int b = blockIdx.z;
    const int xi = blockIdx.x * 32 + tx;
    const int y_first = (row_lo != nullptr) ? row_lo[b] : 0; /* rows below the sample's window hold nothing (undefined maps) */
    int *e = ext + (long)b * 4 * S;
    int c_lo = 0x7f7f7f7f, c_hi = -1;
    float gr[4][3], ga[4];
    int fis[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int yi = blockIdx.y * 32 + r * 8 + ty;
        const bool in = xi < S && yi < S && yi >= y_first;
        gr[r][0] = gr[r][1] = gr[r][2] = 0.0f;
        ga[r] = 0.0f;
        fis[r] = -1;
        if (in) {
            if (g_rgb != nullptr) {
#pragma unroll
                for (int c = 0; c < 3; c++)
                    gr[r][c] = g_rgb[hoc_rgb_off(layout, S, b, yi, xi, c)];
            }
            if (g_alpha != nullptr)
                ga[r] = g_alpha[hoc_plane_off(layout, S, b, yi, xi)];
            fis[r] = face_index_map[((long)b * S + yi) * S + xi];
        }
    }
    unsigned want[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int yi = blockIdx.y * 32 + r * 8 + ty;
        const bool nz = !(gr[r][0] == 0.0f) || !(gr[r][1] == 0.0f) || !(gr[r][2] == 0.0f) || !(ga[r] == 0.0f);
        want[r] = __ballot_sync(HOC_FULL_MASK, !NO_LIST && fis[r] >= 0 && (list_all || nz));
        if (tx == 0)
            s_cnt[r * 8 + ty] = __popc(want[r]);
        if (SPAN) {
            /* span of the pixels the line pass has to look at: covered ones (they own the scans) and those with an
             * incoming gradient (the only ones a scan gets a term from) */
            const bool sp = nz || (K4 && fis[r] >= 0);
            const unsigned m = __ballot_sync(HOC_FULL_MASK, sp);
            if (m != 0 && tx == 0) {
                atomicMax(&e[EXT_ROW_LO * S + yi], S - (blockIdx.x * 32 + (__ffs(m) - 1)));
                atomicMax(&e[EXT_ROW_HI * S + yi], blockIdx.x * 32 + (31 - __clz(m)) + 1);
            }
            if (sp) {
                c_lo = min(c_lo, yi);
                c_hi = max(c_hi, yi);
            }
        }
    }
    if (SPAN) {
        s_lo[ty][tx] = c_lo;
        s_hi[ty][tx] = c_hi;
    }
    __syncthreads();
    if (ty == 0) { /* exclusive prefix over the 32 warp-rows + one atomic for the tile */
        const int mine = s_cnt[tx];
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(HOC_FULL_MASK, incl, o);
            if (tx >= o)
                incl += t;
        }
        s_cnt[tx] = incl - mine;
        if (tx == 31)
            s_cnt[32] = (incl > 0) ? atomicAdd(cov_count + b, incl) : 0;
        if (SPAN && xi < S) {
#pragma unroll
            for (int r = 1; r < 8; r++) {
                c_lo = min(c_lo, s_lo[r][tx]);
                c_hi = max(c_hi, s_hi[r][tx]);
            }
            if (c_hi >= 0) {
                atomicMax(&e[EXT_COL_LO * S + xi], S - c_lo);
                atomicMax(&e[EXT_COL_HI * S + xi], c_hi + 1);
            }
        }
    }
    __syncthreads();
    int2 *list = cov_list + (long)b * S * S + s_cnt[32];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int lr = r * 8 + ty;
        const int yi = blockIdx.y * 32 + lr;
        if ((want[r] >> tx) & 1u)
            list[s_cnt[lr] + __popc(want[r] & ((1u << tx) - 1u))] = make_int2(yi * S + xi, fis[r]);
    }
}

/*
 * The same pass for the image layout with S % 4 == 0 (what the frame-pair path and rasterize_rgbad use): four pixels of
 * a row per thread, so face_index_map and the gradient planes move as 16-byte loads -- the scalar kernel above spends
 * ~200 instructions per pixel on address arithmetic (ncu: 12.9 M warp instructions for 2 M pixels, issue-bound).
 * CTA = 256 threads = a tile of 8 image rows x 128 columns; a warp owns one row of the tile.
 */
__global__ void __launch_bounds__(256)
hoc_raster_bwd_scan4_kernel(const int32_t *__restrict__ face_index_map, const float *__restrict__ g_rgb_k4,
                            const float *__restrict__ g_rgb_rest, const float *__restrict__ g_alpha_k4, int S,
                            int k4_samples, int list_all, int scan_flags, int *__restrict__ ext,
                           int *__restrict__ cov_count,
                            int2 *__restrict__ cov_list, float *__restrict__ zero_a, long n_a,
                            float *__restrict__ zero_b, long n_b, float *__restrict__ zero_c, long n_c,
                            const int *__restrict__ row_lo)
{
    {
        const long nthreads = (long)gridDim.x * gridDim.y * gridDim.z * 256;
        const long t0 = (((long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 256 + threadIdx.x;
        hoc_zero_floats(zero_a, n_a, t0, nthreads);
        hoc_zero_floats(zero_b, n_b, t0, nthreads);
        hoc_zero_floats(zero_c, n_c, t0, nthreads); /* (a buffer of the caller's NEXT kernel: hoc_mesh_scatter's outputs) */
    }
    __shared__ int s_clo[128], s_chi[128];
    __shared__ int s_wcnt[8], s_base;
    const int b = blockIdx.z;
    const bool K4 = b < k4_samples;
    const float *g_rgb = K4 ? g_rgb_k4 : g_rgb_rest;
    const float *g_alpha = K4 ? g_alpha_k4 : nullptr;
    /* scan_flags: HOC_SCAN_SPAN_ALL = line spans for every sample (the fused line pass also runs the texture gradient of
     * the samples without pseudo-gradient, row by row), HOC_SCAN_NO_LIST = nobody reads the list of covered pixels */
    const bool SPAN = K4 || (scan_flags & HOC_SCAN_SPAN_ALL);
    const bool NO_LIST = (scan_flags & HOC_SCAN_NO_LIST) != 0;
    const int lane = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int y_img = blockIdx.y * 8 + ty;  /* image row (rows flipped): raster row yi = S - 1 - y_img */
    const int yi = S - 1 - y_img;
    const int x0 = blockIdx.x * 128 + lane * 4;
    const int y_first = (row_lo != nullptr) ? row_lo[b] : 0;
    const bool in = y_img < S && x0 < S && yi >= y_first;
    if (threadIdx.x < 128) {
        s_clo[threadIdx.x] = 0x7fffffff;
        s_chi[threadIdx.x] = -1;
    }
    int fis[4] = {-1, -1, -1, -1};
    unsigned nzm = 0;
    if (in) {
        const int4 f4 = *reinterpret_cast<const int4 *>(face_index_map + ((long)b * S + yi) * S + x0);
        fis[0] = f4.x; fis[1] = f4.y; fis[2] = f4.z; fis[3] = f4.w;
        if (g_rgb != nullptr) {
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float4 g = *reinterpret_cast<const float4 *>(g_rgb + (((long)b * 3 + c) * S + y_img) * S + x0);
                nzm |= (!(g.x == 0.0f) ? 1u : 0u) | (!(g.y == 0.0f) ? 2u : 0u) | (!(g.z == 0.0f) ? 4u : 0u) |
                       (!(g.w == 0.0f) ? 8u : 0u);
            }
        }
        if (g_alpha != nullptr) {
            const float4 g = *reinterpret_cast<const float4 *>(g_alpha + ((long)b * S + y_img) * S + x0);
            nzm |= (!(g.x == 0.0f) ? 1u : 0u) | (!(g.y == 0.0f) ? 2u : 0u) | (!(g.z == 0.0f) ? 4u : 0u) |
                   (!(g.w == 0.0f) ? 8u : 0u);
        }
    }
    unsigned want = 0;
#pragma unroll
    for (int j = 0; j < 4; j++)
        if (!NO_LIST && fis[j] >= 0 && (list_all || ((nzm >> j) & 1u)))
            want |= 1u << j;
    __syncthreads(); /* s_clo / s_chi initialised */
    if (SPAN) {
        /* span of this row (the warp's) and of the tile's columns */
        /* (pixels the line pass has to look at: those with an incoming gradient -- the only ones a scan gets a term
         * from -- and the covered ones, which own the scans) */
        const unsigned spm = nzm | (K4 ? ((fis[0] >= 0 ? 1u : 0u) | (fis[1] >= 0 ? 2u : 0u) | (fis[2] >= 0 ? 4u : 0u) |
                                          (fis[3] >= 0 ? 8u : 0u))
                                       : 0u);
        int lo = spm ? x0 + (__ffs(spm) - 1) : 0x7fffffff, hi = spm ? x0 + (31 - __clz(spm)) : -1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(HOC_FULL_MASK, lo, o));
            hi = max(hi, __shfl_xor_sync(HOC_FULL_MASK, hi, o));
        }
        int *e = ext + (long)b * 4 * S;
        if (lane == 0 && hi >= 0) {
            atomicMax(&e[EXT_ROW_LO * S + yi], S - lo);
            atomicMax(&e[EXT_ROW_HI * S + yi], hi + 1);
        }
#pragma unroll
        for (int j = 0; j < 4; j++)
            if ((spm >> j) & 1u) {
                atomicMin(&s_clo[lane * 4 + j], yi);
                atomicMax(&s_chi[lane * 4 + j], yi);
            }
    }
    /* list slots: exclusive prefix of the per-thread counts over the CTA, one global atomic per tile */
    const int mine = __popc(want);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(HOC_FULL_MASK, incl, o);
        if (lane >= o)
            incl += t;
    }
    if (lane == 31)
        s_wcnt[ty] = incl;
    __syncthreads(); /* warp totals and the column spans are complete */
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < 8; w++) {
            const int c = s_wcnt[w];
            s_wcnt[w] = tot;
            tot += c;
        }
        s_base = (tot > 0) ? atomicAdd(cov_count + b, tot) : 0;
    }
    if (SPAN && threadIdx.x < 128) {
        const int xi = blockIdx.x * 128 + threadIdx.x;
        if (xi < S && s_chi[threadIdx.x] >= 0) {
            int *e = ext + (long)b * 4 * S;
            atomicMax(&e[EXT_COL_LO * S + xi], S - s_clo[threadIdx.x]);
            atomicMax(&e[EXT_COL_HI * S + xi], s_chi[threadIdx.x] + 1);
        }
    }
    __syncthreads();
    if (want) {
        int2 *list = cov_list + (long)b * S * S + s_base + s_wcnt[ty] + (incl - mine);
        int k = 0;
#pragma unroll
        for (int j = 0; j < 4; j++)
            if ((want >> j) & 1u)
                list[k++] = make_int2(yi * S + x0 + j, fis[j]);
    }
}

/*
 * Frame-pair path: the scan pass FUSED with the backward of pair_consist (hoc_warp_photo_backward_pair).  The incoming
 * gradient of a render's rgb map is d loss / d flow x mult at the few valid pixels of the crop and zero everywhere
 * else, so instead of one pass that writes it (12 B/px) and another that reads it back, this pass reads the valid mask
 * (1 B/px), computes the gradient of the valid pixels (phase B of the two-phase pattern: 12 bilinear taps each, one
 * pixel per thread) into a shared tile, and goes on as hoc_raster_bwd_scan4_kernel from there -- spans, covered-pixel
 * list, zero-fills -- writing the gradient planes once for the line pass.  One launch and 12 B/px less.
 * Stacked row b of the launch is render (b + row_offset) / pairs of pair (b + row_offset) % pairs; render 1's flow is
 * consumed by pair_consist's direction 1, render 2's by direction 0.
 */
struct HocPairGradSrc {
    HocPairBwdDir dir[2]; /* by pair_consist direction (warp_math.cuh); grad_rgb / grad_flow unused here */
    const float *grad_loss, *grad_mean;
    int pairs, H, W, row_offset;
    float inv_w, inv_h;
};

__global__ void __launch_bounds__(256)
hoc_raster_bwd_scan_pair_kernel(const int32_t *__restrict__ face_index_map, HocPairGradSrc G, float *__restrict__ g_rgb,
                                int S, int k4_samples, int list_all, int scan_flags,
                                int *__restrict__ ext,
                                int *__restrict__ cov_count, int2 *__restrict__ cov_list, float *__restrict__ zero_a,
                                long n_a, float *__restrict__ zero_b, long n_b, float *__restrict__ zero_c, long n_c,
                                const int *__restrict__ row_lo)
{
    hoc_pdl_sync(); /* programmatic dependent launch: see hoc_common.cuh */
    {
        const long nthreads = (long)gridDim.x * gridDim.y * gridDim.z * 256;
        const long t0 = (((long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 256 + threadIdx.x;
        /* (grid = (samples, column tiles, row tiles): sample fastest, row tiles from the image centre outwards) */
        hoc_zero_floats(zero_a, n_a, t0, nthreads);
        hoc_zero_floats(zero_b, n_b, t0, nthreads);
        hoc_zero_floats(zero_c, n_c, t0, nthreads); /* (a buffer of the caller's NEXT kernel: hoc_mesh_scatter's outputs) */
    }
    __shared__ int s_clo[128], s_chi[128];
    __shared__ int s_wcnt[8], s_base, s_n;
    __shared__ unsigned short s_list[1024];
    __shared__ __align__(16) float s_gx[1024];
    __shared__ __align__(16) float s_gy[1024];
    const int b = blockIdx.x;
    const int tile_x = blockIdx.y, tile_y = hoc_centre_out(blockIdx.z, gridDim.z);
    const bool K4 = b < k4_samples;
    /* scan_flags: HOC_SCAN_SPAN_ALL = line spans for every sample (the fused line pass also runs the texture gradient of
     * the samples without pseudo-gradient, row by row), HOC_SCAN_NO_LIST = nobody reads the list of covered pixels */
    const bool SPAN = K4 || (scan_flags & HOC_SCAN_SPAN_ALL);
    const bool NO_LIST = (scan_flags & HOC_SCAN_NO_LIST) != 0;
    const int r = b + G.row_offset;
    const int bp = r < G.pairs ? r : r - G.pairs;     /* the pair (r < 2 pairs) */
    const HocPairBwdDir &D = G.dir[r < G.pairs ? 1 : 0]; /* render 1 <- direction 1, render 2 <- direction 0 */
    const int H = G.H, W = G.W;
    const int lane = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int y_img = tile_y * 8 + ty; /* image row (rows flipped): raster row yi = S - 1 - y_img */
    const int yi = S - 1 - y_img;
    const int x0 = tile_x * 128 + lane * 4;
    const int y_first = (row_lo != nullptr) ? row_lo[b] : 0;
    const bool in = y_img < S && x0 < S && yi >= y_first;
    if (threadIdx.x < 128) {
        s_clo[threadIdx.x] = 0x7fffffff;
        s_chi[threadIdx.x] = -1;
    }
    if (threadIdx.x == 0)
        s_n = 0;
    *reinterpret_cast<float4 *>(s_gx + threadIdx.x * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4 *>(s_gy + threadIdx.x * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    /* phase A: face indices of the four pixels; the valid pixels of the crop are noted in the shared list */
    int fis[4] = {-1, -1, -1, -1};
    unsigned vb = 0u;
    const long npix = (long)H * W;
    if (in) {
        const int4 f4 = *reinterpret_cast<const int4 *>(face_index_map + ((long)b * S + yi) * S + x0);
        fis[0] = f4.x; fis[1] = f4.y; fis[2] = f4.z; fis[3] = f4.w;
        if (D.active && y_img < H && x0 < W) {
            vb = *reinterpret_cast<const unsigned *>(D.valid_mask + (long)bp * npix + (long)y_img * W + x0);
#pragma unroll
            for (int j = 0; j < 4; j++)
                if ((vb >> (8 * j)) & 0xffu)
                    s_list[atomicAdd(&s_n, 1)] = (unsigned short)(threadIdx.x * 4 + j);
        }
    }
    /* a tile no mesh covers (most of them) has nothing to list, no span and no gradient: zero planes and out */
    if (!__syncthreads_or((fis[0] & fis[1] & fis[2] & fis[3]) >= 0 || vb != 0u)) {
        if (in) {
            float *dst = g_rgb + (((long)b * 3) * S + y_img) * S + x0;
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4 *>(dst) = z4;
            *reinterpret_cast<float4 *>(dst + (long)S * S) = z4;
            if (!(scan_flags & HOC_SCAN_TWO_CHANNELS))
                *reinterpret_cast<float4 *>(dst + 2l * S * S) = z4;
        }
        return;
    }
    /* phase B: d loss / d flow x mult of the listed pixels, one per thread (hoc_warp_photo_pair_backward_kernel) */
    const int n_valid = s_n;
    if (n_valid > 0) {
        const float cnt = (float)D.sums[2 * bp + 1];
        const float gl = ((G.grad_loss != nullptr) ? G.grad_loss[bp] : 0.0f) +
                         ((G.grad_mean != nullptr) ? __fdiv_rn(G.grad_mean[0], (float)G.pairs) : 0.0f);
        const float scale = gl / fmaxf(cnt, 1.0f);
        for (int i = threadIdx.x; i < n_valid; i += 256) {
            const int loc = s_list[i];
            const int t = loc >> 2, j = loc & 3;
            const int y = tile_y * 8 + (t >> 5), x = tile_x * 128 + (t & 31) * 4 + j;
            float gfx, gfy;
            hoc_pair_bwd_pixel(D, bp, x, y, H, W, G.inv_w, G.inv_h, scale, &gfx, &gfy);
            const float m = D.mult[(long)bp * npix + (long)y * W + x]; /* hoc_flow_finalize_backward_kernel */
            s_gx[loc] = gfx * m;
            s_gy[loc] = gfy * m;
        }
    }
    __syncthreads();
    /* phase C: the gradient planes (written once, for the line pass) and the scan pass proper */
    unsigned nzm = 0;
    if (in) {
        const float4 gx = *reinterpret_cast<const float4 *>(s_gx + threadIdx.x * 4);
        const float4 gy = *reinterpret_cast<const float4 *>(s_gy + threadIdx.x * 4);
        float *dst = g_rgb + (((long)b * 3) * S + y_img) * S + x0;
        *reinterpret_cast<float4 *>(dst) = gx;
        *reinterpret_cast<float4 *>(dst + (long)S * S) = gy;
        if (!(scan_flags & HOC_SCAN_TWO_CHANNELS))
            *reinterpret_cast<float4 *>(dst + 2l * S * S) = make_float4(0.f, 0.f, 0.f, 0.f);
        nzm = ((!(gx.x == 0.0f) || !(gy.x == 0.0f)) ? 1u : 0u) | ((!(gx.y == 0.0f) || !(gy.y == 0.0f)) ? 2u : 0u) |
              ((!(gx.z == 0.0f) || !(gy.z == 0.0f)) ? 4u : 0u) | ((!(gx.w == 0.0f) || !(gy.w == 0.0f)) ? 8u : 0u);
    }
    unsigned want = 0;
#pragma unroll
    for (int j = 0; j < 4; j++)
        if (!NO_LIST && fis[j] >= 0 && (list_all || ((nzm >> j) & 1u)))
            want |= 1u << j;
    if (SPAN) {
        /* (pixels the line pass has to look at: those with an incoming gradient -- the only ones a scan gets a term
         * from -- and the covered ones, which own the scans) */
        const unsigned spm = nzm | (K4 ? ((fis[0] >= 0 ? 1u : 0u) | (fis[1] >= 0 ? 2u : 0u) | (fis[2] >= 0 ? 4u : 0u) |
                                          (fis[3] >= 0 ? 8u : 0u))
                                       : 0u);
        int lo = spm ? x0 + (__ffs(spm) - 1) : 0x7fffffff, hi = spm ? x0 + (31 - __clz(spm)) : -1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(HOC_FULL_MASK, lo, o));
            hi = max(hi, __shfl_xor_sync(HOC_FULL_MASK, hi, o));
        }
        int *e = ext + (long)b * 4 * S;
        if (lane == 0 && hi >= 0) {
            atomicMax(&e[EXT_ROW_LO * S + yi], S - lo);
            atomicMax(&e[EXT_ROW_HI * S + yi], hi + 1);
        }
#pragma unroll
        for (int j = 0; j < 4; j++)
            if ((spm >> j) & 1u) {
                atomicMin(&s_clo[lane * 4 + j], yi);
                atomicMax(&s_chi[lane * 4 + j], yi);
            }
    }
    const int mine = __popc(want);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(HOC_FULL_MASK, incl, o);
        if (lane >= o)
            incl += t;
    }
    if (lane == 31)
        s_wcnt[ty] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < 8; w++) {
            const int c = s_wcnt[w];
            s_wcnt[w] = tot;
            tot += c;
        }
        s_base = (tot > 0) ? atomicAdd(cov_count + b, tot) : 0;
    }
    if (SPAN && threadIdx.x < 128) {
        const int xi = tile_x * 128 + threadIdx.x;
        if (xi < S && s_chi[threadIdx.x] >= 0) {
            int *e = ext + (long)b * 4 * S;
            atomicMax(&e[EXT_COL_LO * S + xi], S - s_clo[threadIdx.x]);
            atomicMax(&e[EXT_COL_HI * S + xi], s_chi[threadIdx.x] + 1);
        }
    }
    __syncthreads();
    if (want) {
        int2 *list = cov_list + (long)b * S * S + s_base + s_wcnt[ty] + (incl - mine);
        int k = 0;
#pragma unroll
        for (int j = 0; j < 4; j++)
            if ((want >> j) & 1u)
                list[k++] = make_int2(yi * S + x0 + j, fis[j]);
    }
}

/* Texture (backward_textures) and depth (backward_depth_map) gradient of one covered pixel. */
template <bool TS2>
__device__ __forceinline__ void hoc_cover_tex_depth(const float *__restrict__ faces,
                                                    const float *__restrict__ weight_map,
                                                    const float *__restrict__ depth_map,
                                                    const float *__restrict__ g_rgb,
                                                    const float *__restrict__ g_depth, int b, int fi, int xi, int yi,
                                                    int F, int S, int ts, float near_, float far_, float eps, int layout,
                                                    int tex_mode, float *__restrict__ acc_d,
                                                    float *__restrict__ grad_textures,
                                                    unsigned long long *__restrict__ det_ad,
                                                    unsigned long long *__restrict__ det_gt)
{
    const bool want_tex = (grad_textures != nullptr) && (g_rgb != nullptr);
    const bool want_depth = (acc_d != nullptr) && (g_depth != nullptr);
    float gr[3] = {0.0f, 0.0f, 0.0f};
    if (want_tex) {
#pragma unroll
        for (int c = 0; c < 3; c++)
            gr[c] = g_rgb[hoc_rgb_off(layout, S, b, yi, xi, c)];
    }
    const bool nz = !(gr[0] == 0.0f) || !(gr[1] == 0.0f) || !(gr[2] == 0.0f);
    if (!((want_tex && nz) || want_depth))
        return;
    const int tex_n = ts * ts * ts * 3;
    float f[9], w[3], zp;
    const float *src = faces + ((long)b * F + fi) * 9;
    if (weight_map != nullptr && depth_map != nullptr) {
        /* the forward saved its weights and depth: only the three vertex depths are needed */
        const float *wm = weight_map + (((long)b * S + yi) * S + xi) * 3;
        w[0] = wm[0];
        w[1] = wm[1];
        w[2] = wm[2];
        zp = depth_map[hoc_plane_off(layout, S, b, yi, xi)];
        f[2] = __ldg(src + 2);
        f[5] = __ldg(src + 5);
        f[8] = __ldg(src + 8);
    } else { /* recompute with the forward's functions (bit-identical) */
        float inv[9];
#pragma unroll
        for (int k = 0; k < 9; k++)
            f[k] = __ldg(src + k);
        hoc_face_inv(f, S, inv);
        hoc_pixel_weights_depth(f, inv, xi, yi, near_, far_, w, &zp);
    }
    if (want_depth) {
        const float gz = g_depth[hoc_plane_off(layout, S, b, yi, xi)] * zp * zp;
        if (gz != 0.0f) {
            const long ad = ((long)b * F + fi) * 3;
#pragma unroll
            for (int k = 0; k < 3; k++)
                hoc_accum(acc_d, ad + k, gz * w[k], det_ad);
        }
    }
    if (want_tex && nz && tex_mode == HOC_TEX_GRAD_VERTEX) {
        /* textures are the multilinear extension of three vertex values (T[i,j,k] = i c0 + j c1 + k c2, ts == 2):
         * d rgb / d c_k = t_k, so nine sums per face instead of twenty-four */
        const long gt = ((long)b * F + fi) * 9;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float t = hoc_tex_coord(w[k], f[3 * k + 2], zp, 2, eps);
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float v = t * gr[c];
                if (v != 0.0f)
                    hoc_accum(grad_textures, gt + 3 * k + c, v, det_gt);
            }
        }
    } else if (want_tex && nz) {
        const long gt = ((long)b * F + fi) * tex_n;
        float tf[3];
        int ti[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float t = hoc_tex_coord(w[k], f[3 * k + 2], zp, ts, eps);
            ti[k] = hoc_tex_cell(t, ts);
            tf[k] = t - (float)ti[k];
        }
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
            if (!TS2 && ts == 1)
                isc = 0;
            if (ww != 0.0f) {
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const float v = ww * gr[c];
                    if (v != 0.0f)
                        hoc_accum(grad_textures, gt + isc * 3 + c, v, det_gt);
                }
            }
        }
    }
}

/*
 * Cover pass: the texture / depth gradient of the pixels the scan pass listed (those with such a gradient), one pixel
 * per thread, spread evenly over the GPU (covered pixels cluster in a few tiles, the list un-clusters them).  Not
 * launched when the line pass runs the texture gradient itself (vertex-value textures, saved weights: the frame-pair
 * path).
 */
#define CV_THREADS 128
template <bool TS2>
#ifdef CV_MINB /* (occupancy experiments: minimum resident CTAs per SM) */
#define CV_BOUNDS __launch_bounds__(CV_THREADS, CV_MINB)
#else
#define CV_BOUNDS __launch_bounds__(CV_THREADS)
#endif
__global__ void CV_BOUNDS
hoc_raster_bwd_cover_kernel(const float *__restrict__ faces, const float *__restrict__ weight_map,
                            const float *__restrict__ depth_map, const float *__restrict__ g_rgb,
                            const float *__restrict__ g_depth, int F, int S, int ts, float near_, float far_, float eps,
                            int layout, int tex_mode, const int *__restrict__ cov_count,
                            const int2 *__restrict__ cov_list, float *__restrict__ acc_d,
                            float *__restrict__ grad_textures, unsigned long long *__restrict__ det_gt,
                            unsigned long long *__restrict__ det_ad)
{
    hoc_pdl_sync(); /* programmatic dependent launch: see hoc_common.cuh */
    const int b = blockIdx.y;
    const int count = min(cov_count[b], S * S);
    const int2 *list = cov_list + (long)b * S * S;
    for (int i = blockIdx.x * CV_THREADS + threadIdx.x; i < count; i += gridDim.x * CV_THREADS) {
        const int2 e = list[i];
        const int yi = e.x / S, xi = e.x - yi * S;
        hoc_cover_tex_depth<TS2>(faces, weight_map, depth_map, g_rgb, g_depth, b, e.y, xi, yi, F, S, ts, near_, far_, eps,
                                 layout, tex_mode, acc_d, grad_textures, det_ad, det_gt);
    }
}

/* backward_depth_map's per-face epilogue: grad_faces += J^T acc_d (z directly, x / y through the weights). */
__global__ void __launch_bounds__(256)
hoc_raster_bwd_depth_kernel(const float *__restrict__ faces, const float *__restrict__ acc_d, long n_faces, int S,
                            float *__restrict__ grad_faces)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_faces)
        return;
    const float *ad = acc_d + i * 3;
    const float a0 = ad[0], a1 = ad[1], a2 = ad[2];
    if (a0 == 0.0f && a1 == 0.0f && a2 == 0.0f)
        return;
    float f[9], inv[9], tmp[2];
#pragma unroll
    for (int k = 0; k < 9; k++)
        f[k] = __ldg(faces + i * 9 + k);
    if (!hoc_face_xy_finite(f) || hoc_face_back(f))
        return;
    hoc_face_inv(f, S, inv);
#pragma unroll
    for (int l = 0; l < 2; l++)
        tmp[l] = inv[l] / f[2] + inv[3 + l] / f[5] + inv[6 + l] / f[8];
    float *gf = grad_faces + i * 9;
    const float a[3] = {a0, a1, a2};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float zk = f[3 * k + 2];
        gf[3 * k + 2] += a[k] / (zk * zk);
        gf[3 * k + 0] += a[k] * tmp[0] * (float)S / 2.0f;
        gf[3 * k + 1] += a[k] * tmp[1] * (float)S / 2.0f;
    }
}

/* One outward scan, held by one lane: everything the chunk loop needs. */
struct HocLineScan {
    float cA, cB;     /* dist to vertex A / B = c * (d1 - cross) + e: c already times 2 / S; 0 for a vertex without term */
    float eA, eB;     /* +-eps with the sign of c * (d1 - cross), which is the same for every pixel of an outward scan
                       * (they all lie strictly on the outside of the crossing); 1 for a vertex without term */
    float cross;
    float I1, I2, I3; /* colour of the inside pixel */
    int from, to;     /* range on the line, clipped to the line's span */
    int gfA, gfB;     /* element of grad_faces, -1 for a vertex without term */
    int nchunk;
};

#define LN_THREADS 256
#define LN_MAX_LINES 8 /* lines per CTA (HOC_TUNE_LINE_LINES) */
#define LN_PAD 32 /* staged entries past the span that the unrolled chunk loop may read (never used): the longest chunk */
#ifdef LN_MINB /* (occupancy experiments: minimum resident CTAs per SM) */
#define LN_BOUNDS __launch_bounds__(LN_THREADS, LN_MINB)
#else
#define LN_BOUNDS __launch_bounds__(LN_THREADS)
#endif

/*
 * Line pass: the pseudo-gradient (backward_pixel_map) of one image column / row of one sample per CTA -- grid ((sample,
 * axis), S), sample fastest and lines from the image centre outwards (the lines that carry the most work -- meshes are centred
 * by the crop -- are dispatched first, the empty border lines last; an empty line costs its CTA ~50 instructions).  On
 * rasters above 320 a CTA runs three lines one after the other (grid ((sample, axis), S / 3); see g_line_lines).
 * Every term of backward_pixel_map lives on one line: for (face f, edge, axis, d0) the inside pixel, the outside pixel, the inward scan and the outward scan all have
 * walk coordinate d0.  So the CTA of line (axis, d0) stages the line's span once -- owning face, colour, incoming
 * gradient of every pixel -- and then runs, from shared memory, one CANDIDATE per (covered pixel of the line, edge of its
 * face): the face's vertices are loaded (three lanes share a face), that column of the edge is evaluated once
 * (hoc_k4_column), the pixel adds its own term of the inward scan (colour of the outside pixel: from the staged line),
 * and if it is the pixel just inside the edge the lane holds the outward scan that starts there in registers (range
 * clipped to the span, per-column constants, colour of the inside pixel).  The warp then cuts its <= 32 scans into
 * chunks of CH pixels and every lane sums ONE chunk out of shared memory -- it finds its scan with a 5-step shuffle
 * search over the warp's prefix sums and fetches the scan's constants with shuffles -- so that the lanes finish
 * together however different the scan lengths are.  The chunk loop is fully unrolled and branch-free: per pixel one
 * 16-byte shared load, delta (3 FMA), both distances (one FMA each: c * kk + (c * u0 + e); the reference's +-eps has a
 * constant sign along an outward scan) and ONE MUFU.RCP of their product (1 / dA = dB * r, 1 / dB = dA * r; 3 ulp: the
 * pseudo-gradient carries a 1e-3 tolerance).  A pixel beyond the end of the scan, or with delta <= 0, adds nothing.
 * Chain of dependent loads per CTA: span -> line -> faces.  (Rounds 1-2 ran this as two passes -- a cover pass doing the
 * per-pixel edge work and queueing the outward scans on their lines, a line pass draining the queues: six dependent
 * loads, a 25 MB queue workspace and 40 us against 31.)
 */
template <int CH, bool MULTI>
__global__ void LN_BOUNDS
hoc_raster_bwd_line_kernel(const float *__restrict__ faces, const int32_t *__restrict__ face_index_map,
                            const float *__restrict__ rgb, const float *__restrict__ g_rgb,
                            const float *__restrict__ g_alpha, int F, int S, float eps, int layout, int use_alpha,
                            const int *__restrict__ ext, float scale, float *__restrict__ grad_faces,
                            unsigned long long *__restrict__ det_gf, int k4_samples, int g_channels,
                            const float *__restrict__ weight_map, const float *__restrict__ depth_map,
                            float *__restrict__ grad_textures, unsigned long long *__restrict__ det_gt, int lines,
                            int fold)
{
    hoc_pdl_sync(); /* programmatic dependent launch: see hoc_common.cuh */
    /* dynamic shared memory, per staged pixel: float4 (P, g_r, g_g, g_b) with P = sum_ch I_ch g_ch - g_alpha (see the
     * chunk loop below), float4 (I_r, I_g, I_b, g_alpha), int owning face */
    extern __shared__ float4 s_line4[];
    __shared__ int s_span[LN_MAX_LINES][3]; /* (lo, hi, d0) of the CTA's lines */
    float4 *s_pg = s_line4, *s_ia = s_line4 + (S + LN_PAD);
    int *s_fi = reinterpret_cast<int *>(s_line4 + 2 * (S + LN_PAD));
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int T = blockDim.x; /* multiple of 32, <= LN_THREADS */
    const bool has_alpha = (use_alpha != 0) && (g_alpha != nullptr);
    const bool has_rgb = (rgb != nullptr) && (g_rgb != nullptr);
    /* grid.x: (sample, axis) -- both axes of the samples with pseudo-gradient, then the rows of the others */
    const int bx = blockIdx.x;
    const int b = bx < 2 * k4_samples ? (bx >> 1) : bx - k4_samples;
    const int axis = bx < 2 * k4_samples ? (bx & 1) : 1;
    /* samples [0, k4_samples) get the pseudo-gradient; with grad_textures given the ROW CTAs of every sample also run
     * backward_textures for the pixels of their row (three vertex values per face, weights and depth saved by the
     * forward): they have staged each pixel's owning face and incoming gradient anyway */
    const bool K4 = b < k4_samples;
    const bool TEX = grad_textures != nullptr && axis == 1 && has_rgb;
    if (!K4 && !TEX)
        return;
    /* level 1: the span of the pixels that matter on a line (covered or with an incoming gradient: the scan pass).
     * MULTI: a CTA runs `lines` lines (HOC_TUNE_LINE_LINES), their spans fetched at once: line k of CTA y has centre-out
     * rank k * G + y (G = grid.y) -- or, folded, k * G + (G - 1 - y) for odd k, which pairs a heavy central line with a
     * light outer one -- so that an empty border line costs a loop iteration instead of a CTA.  (One line per CTA is its
     * own instantiation: the shared-memory round trip of the spans and the barrier cost the 10 000 empty CTAs of a
     * 256^2 step 2 us.) */
    if (MULTI && tid < lines) {
        const int G = gridDim.y;
        const int li = tid * G + ((fold && (tid & 1)) ? G - 1 - (int)blockIdx.y : (int)blockIdx.y);
        int lo_ = 1, hi_ = 0, d0_ = 0;
        if (li < S) {
            const int *e = ext + (long)b * 4 * S;
            d0_ = hoc_centre_out(li, S);
            lo_ = S - e[(axis == 0 ? EXT_COL_LO : EXT_ROW_LO) * S + d0_];
            hi_ = e[(axis == 0 ? EXT_COL_HI : EXT_ROW_HI) * S + d0_] - 1;
        }
        s_span[tid][0] = lo_;
        s_span[tid][1] = hi_;
        s_span[tid][2] = d0_;
    }
    const int32_t *idx = face_index_map + (long)b * S * S;
    for (int ln = 0; ln < (MULTI ? lines : 1); ln++) {
    int lo, hi, d0;
    if (MULTI) {
        __syncthreads(); /* the spans are written; every warp is done with the previous line's staged pixels */
        lo = s_span[ln][0];
        hi = s_span[ln][1];
        d0 = s_span[ln][2];
        if (lo > hi)
            continue;
    } else {
        const int *e = ext + (long)b * 4 * S;
        d0 = hoc_centre_out(blockIdx.y, S);
        lo = S - e[(axis == 0 ? EXT_COL_LO : EXT_ROW_LO) * S + d0];
        hi = e[(axis == 0 ? EXT_COL_HI : EXT_ROW_HI) * S + d0] - 1;
        if (lo > hi)
            return;
    }
    const int len = hi - lo + 1;

    /* level 2: the line */
    for (int i = tid; i < len; i += T) {
        const int d1 = lo + i;
        const int xi = axis == 0 ? d0 : d1, yi = axis == 0 ? d1 : d0;
        float4 pg = make_float4(0.0f, 0.0f, 0.0f, 0.0f), ia = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        const int fi = idx[(long)yi * S + xi];
        if (has_rgb) {
            const long o0 = hoc_rgb_off(layout, S, b, yi, xi, 0), o1 = hoc_rgb_off(layout, S, b, yi, xi, 1),
                       o2 = hoc_rgb_off(layout, S, b, yi, xi, 2);
            ia.x = rgb[o0];
            ia.y = rgb[o1];
            pg.y = g_rgb[o0];
            pg.z = g_rgb[o1];
            if (g_channels > 2) { /* (frame-pair path: the third channel carries no gradient -- the plane does not exist) */
                ia.z = rgb[o2];
                pg.w = g_rgb[o2];
            }
            pg.x = ia.x * pg.y + ia.y * pg.z + ia.z * pg.w;
        }
        if (has_alpha) {
            const float ga = g_alpha[hoc_plane_off(layout, S, b, yi, xi)];
            const float a = (fi >= 0) ? 1.0f : 0.0f;
            pg.x += a * ga - ga;
            ia.w = ga;
        }
        s_pg[i] = pg;
        s_ia[i] = ia;
        s_fi[i] = fi;
    }
    __syncthreads();

    if (TEX) { /* d rgb / d c_k = t_k for the three vertex values of the owning face (hoc_cover_tex_depth, vertex mode) */
        for (int i = tid; i < len; i += T) {
            const float4 pg = s_pg[i];
            const int fi = s_fi[i];
            if (fi < 0 || fi >= F || (pg.y == 0.0f && pg.z == 0.0f && pg.w == 0.0f))
                continue;
            const int xi = lo + i, yi = d0; /* (axis 1: the line is raster row d0) */
            const float *wm = weight_map + (((long)b * S + yi) * S + xi) * 3;
            const float w3[3] = {wm[0], wm[1], wm[2]};
            const float zp = depth_map[hoc_plane_off(layout, S, b, yi, xi)];
            const float *src = faces + ((long)b * F + fi) * 9;
            const long gt = ((long)b * F + fi) * 9;
            const float gr[3] = {pg.y, pg.z, pg.w};
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float t = hoc_tex_coord(w3[k], __ldg(src + 3 * k + 2), zp, 2, eps);
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const float v = t * gr[c];
                    if (v != 0.0f)
                        hoc_accum(grad_textures, gt + 3 * k + c, v, det_gt);
                }
            }
        }
        if (!K4)
            continue;
    }

    HocBwdMaps M; /* (for the outside pixel of an inward term that lies beyond the staged span) */
    M.idx = idx;
    M.rgb = rgb;
    M.g_rgb = g_rgb;
    M.g_alpha = g_alpha;
    M.S = S;
    M.layout = layout;
    M.b = b;
    M.use_alpha = has_alpha;
    M.use_rgb = has_rgb;
    /* level 3: the candidates, 32 per warp at a time; the warps of the CTA no longer synchronise */
    const int ncand = 3 * len;
    for (int c0 = wid * 32; c0 < ncand; c0 += T) {
        const int c = c0 + lane;
        HocLineScan sc;
        sc.cA = sc.cB = 0.0f;
        sc.eA = sc.eB = 1.0f;
        sc.cross = 0.0f;
        sc.I1 = sc.I2 = sc.I3 = 0.0f;
        sc.from = 0;
        sc.to = -1;
        sc.gfA = sc.gfB = -1;
        sc.nchunk = 0;
        const int i = (int)(((unsigned)min(c, ncand - 1) * 43691u) >> 17); /* c / 3 (exact below 98 304) */
        const int edge = min(c, ncand - 1) - 3 * i;
        const int fi = (c < ncand) ? s_fi[i] : -1;
        if (fi >= 0 && fi < F) {
            const int d1p = lo + i;
            const int ia_ = edge, ib_ = (edge == 2) ? 0 : edge + 1, ic_ = (edge == 0) ? 2 : edge - 1;
            const float *src = faces + ((long)b * F + fi) * 9;
            const float ax = __ldg(src + 3 * ia_), ay = __ldg(src + 3 * ia_ + 1);
            const float bx = __ldg(src + 3 * ib_), by = __ldg(src + 3 * ib_ + 1);
            const float cx = __ldg(src + 3 * ic_), cy = __ldg(src + 3 * ic_ + 1);
            /* the owner of a pixel is front-facing with finite xy (the forward's tests); re-checked, in the face's
             * stored vertex order (hoc_face_back on the rotated vertices is NOT bit-identical) */
            float f[9];
            f[0] = (edge == 0) ? ax : ((edge == 1) ? cx : bx);
            f[1] = (edge == 0) ? ay : ((edge == 1) ? cy : by);
            f[3] = (edge == 0) ? bx : ((edge == 1) ? ax : cx);
            f[4] = (edge == 0) ? by : ((edge == 1) ? ay : cy);
            f[6] = (edge == 0) ? cx : ((edge == 1) ? bx : ax);
            f[7] = (edge == 0) ? cy : ((edge == 1) ? by : ay);
            f[2] = f[5] = f[8] = 0.0f;
            HocK4Stage K;
            hoc_k4_edge_pts(ax, ay, bx, by, cx, cy, S, axis, &K.E);
            int d1_in = 0, d1_out = 0;
            if (hoc_face_xy_finite(f) && !hoc_face_back(f) && d0 >= K.E.d0_from && d0 <= K.E.d0_to &&
                hoc_k4_column(&K.E, S, d0, &K.d1_cross, &d1_in, &d1_out)) {
                const float4 pa = s_ia[i], pp = s_pg[i];
                const long gf = ((long)b * F + fi) * 9;
                /* (a) the pixel's own term of the inward scan (the owned pixels between the edge and the opposite edge) */
                const int lim = hoc_k4_inward_limit(&K.E, d0);
                if (max(min(d1_in, lim), 0) <= d1p && d1p <= min(max(d1_in, lim), S - 1)) {
                    const float I[4] = {1.0f, pa.x, pa.y, pa.z}, g[4] = {pa.w, pp.y, pp.z, pp.w};
                    float O[4];
                    const int j = d1_out - lo;
                    if (j >= 0 && j < len) {
                        const float4 oa = s_ia[j];
                        O[0] = (has_alpha && s_fi[j] >= 0) ? 1.0f : 0.0f;
                        O[1] = oa.x;
                        O[2] = oa.y;
                        O[3] = oa.z;
                    } else {
                        hoc_load_I(M, axis == 0 ? d0 : d1_out, axis == 0 ? d1_out : d0, O);
                    }
                    K.d0 = d0;
                    K.d1p = d1p;
                    hoc_k4_stage_b(K, axis, M, I, O, g, eps, grad_faces, gf + 3 * ia_, gf + 3 * ib_, det_gf);
                }
                /* (b) the pixel just inside the edge: the outward scan starts here, clipped to the span */
                if (d1_in == d1p) {
                    sc.from = (0 < K.E.dir) ? max(d1_out, lo) : lo;
                    sc.to = (0 < K.E.dir) ? hi : min(d1_out, hi);
                    if (sc.to >= sc.from) {
                        HocK4Col C;
                        hoc_k4_col(&K.E, S, d0, K.d1_cross, &C);
                        const float t_first = (float)d1_out - K.d1_cross; /* sign of (d1 - cross) on the whole scan */
                        const int gbase = (int)gf + (1 - axis);
                        sc.cross = K.d1_cross;
                        if (C.hasA) {
                            sc.cA = C.cA * scale;
                            sc.eA = (0.0f < sc.cA * t_first) ? eps : -eps;
                            sc.gfA = gbase + ia_ * 3;
                        }
                        if (C.hasB) {
                            sc.cB = C.cB * scale;
                            sc.eB = (0.0f < sc.cB * t_first) ? eps : -eps;
                            sc.gfB = gbase + ib_ * 3;
                        }
                        sc.I1 = pa.x;
                        sc.I2 = pa.y;
                        sc.I3 = pa.z;
                        sc.nchunk = (sc.to - sc.from + CH) / CH;
                    }
                }
            }
        }
        /* the warp's scans, cut into chunks of CH pixels, one chunk per lane */
        int incl = sc.nchunk;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(HOC_FULL_MASK, incl, o);
            if (lane >= o)
                incl += t;
        }
        const int pre = incl - sc.nchunk;
        const int total = __shfl_sync(HOC_FULL_MASK, incl, 31);
        for (int j0 = 0; j0 < total; j0 += 32) {
            const int j = min(j0 + lane, total - 1);
            const bool live = j0 + lane < total;
            int a = 0; /* last scan with pre <= j */
#pragma unroll
            for (int st = 16; st >= 1; st >>= 1) {
                const int pv = __shfl_sync(HOC_FULL_MASK, pre, (a + st) & 31);
                if (a + st < 32 && pv <= j)
                    a += st;
            }
            const int cc = (j - __shfl_sync(HOC_FULL_MASK, pre, a)) * CH;
            const int d1_from = __shfl_sync(HOC_FULL_MASK, sc.from, a) + cc;
            const int to_a = __shfl_sync(HOC_FULL_MASK, sc.to, a);
            const int left = live ? to_a - d1_from : -1; /* pixels beyond the first */
            const float cA = __shfl_sync(HOC_FULL_MASK, sc.cA, a), cB = __shfl_sync(HOC_FULL_MASK, sc.cB, a);
            const float eA = __shfl_sync(HOC_FULL_MASK, sc.eA, a), eB = __shfl_sync(HOC_FULL_MASK, sc.eB, a);
            const float I1 = __shfl_sync(HOC_FULL_MASK, sc.I1, a), I2 = __shfl_sync(HOC_FULL_MASK, sc.I2, a),
                        I3 = __shfl_sync(HOC_FULL_MASK, sc.I3, a);
            const float u0 = (float)d1_from - __shfl_sync(HOC_FULL_MASK, sc.cross, a);
            const int gfA = __shfl_sync(HOC_FULL_MASK, sc.gfA, a), gfB = __shfl_sync(HOC_FULL_MASK, sc.gfB, a);
            const float dA0 = __fmaf_rn(cA, u0, eA), dB0 = __fmaf_rn(cB, u0, eB);
            const float4 *sp = s_pg + (d1_from - lo);
            float gA = 0.0f, gB = 0.0f;
#pragma unroll
            for (int kk = 0; kk < CH; kk++) {
                const float4 pg = sp[kk]; /* at most CH - 1 entries past the staged span: the buffer is padded */
                const float delta = __fmaf_rn(-I3, pg.w, __fmaf_rn(-I2, pg.z, __fmaf_rn(-I1, pg.y, pg.x)));
                const float dA = __fmaf_rn(cA, (float)kk, dA0), dB = __fmaf_rn(cB, (float)kk, dB0);
                float t = delta * hoc_rcp_approx(dA * dB);
                t = (delta <= 0.0f || kk > left) ? 0.0f : t;
                gA = __fmaf_rn(-t, dB, gA);
                gB = __fmaf_rn(-t, dA, gB);
            }
            if (gA != 0.0f && gfA >= 0)
                hoc_accum(grad_faces, gfA, gA, det_gf);
            if (gB != 0.0f && gfB >= 0)
                hoc_accum(grad_faces, gfB, gB, det_gf);
        }
    }
    } /* (the CTA's next line) */
}

/* Tuning knobs of the line pass (hoc_set_tuning): threads per CTA, chunk length in pixels (8 or 16). */
static int g_cover_ctas = 296;
static int g_tex_in_line = 1; /* the line pass's row CTAs also run the (vertex-value) texture gradient (HOC_TUNE_TEX_IN_LINE) */
/* lines per CTA of the line pass and the way they are dealt (HOC_TUNE_LINE_LINES / _FOLD).  0: by raster size -- one line
 * per CTA up to 320 (measured at 16 pairs of 256^2, all through the MULTI instantiation: 119.8 us per step with 1,
 * 120.2 / 120.5 / 121.4 with 2 / 3 / 4 folded; unfolded 121.3 / 125.2 / 139.2 with 2 / 4 / 8; the MULTI = false
 * instantiation is another 2 us faster), three folded lines above (32 pairs of 480 x 270: 494.4 -> 483.5 us) */
static int g_line_lines = 0, g_line_fold = 1;
static int g_line_threads = 128, g_line_seg = 0; /* seg 0: by raster size (16 pixels up to 320, 32 above: 176 -> 163 us at 32 x 480^2) */

extern "C" int hoc_set_tuning(int key, int value)
{
    if (key == HOC_TUNE_LINE_THREADS && value >= 32 && value <= LN_THREADS && value % 32 == 0)
        g_line_threads = value;
    else if (key == HOC_TUNE_LINE_SEGMENT && (value == 0 || value == 8 || value == 16 || value == 32))
        g_line_seg = value;
    else if (key == HOC_TUNE_DETERMINISTIC && (value == 0 || value == 1))
        g_hoc_deterministic = value;
    else if (key == HOC_TUNE_PDL && (value == 0 || value == 1))
        g_hoc_pdl = value;
    else if (key == HOC_TUNE_COVER_CTAS && value >= 1 && value <= 65535)
        g_cover_ctas = value;
    else if (key == HOC_TUNE_TEX_IN_LINE && (value == 0 || value == 1))
        g_tex_in_line = value;
    else if (key == HOC_TUNE_LINE_LINES && value >= 0 && value <= LN_MAX_LINES)
        g_line_lines = value;
    else if (key == HOC_TUNE_LINE_FOLD && (value == 0 || value == 1))
        g_line_fold = value;
    else {
        hoc_set_error("hoc_set_tuning: bad key %d / value %d", key, value);
        return HOC_ERR_INVALID_ARG;
    }
    return HOC_OK;
}

template <int CH>
static cudaError_t hoc_launch_line(const float *faces, const int32_t *face_index_map, const float *rgb,
                                    const float *grad_rgb, const float *g_alpha, int B, int k4_samples, int F, int S,
                                    float eps, int layout, int use_alpha, const HocBwdWorkspace &w, float *grad_faces,
                                    const float *weight_map, const float *depth_map, float *grad_textures,
                                    int g_channels, cudaStream_t st)
{
    /* two float4 and one int per staged pixel, + padding for the unrolled chunk loop: 9.8 KB at S = 256, 74 KB at 2048 */
    const size_t smem = ((size_t)S + LN_PAD) * (2 * sizeof(float4) + sizeof(int));
    const int lines = g_line_lines > 0 ? g_line_lines : (S > 320 ? 3 : 1);
    auto kernel = lines > 1 ? hoc_raster_bwd_line_kernel<CH, true> : hoc_raster_bwd_line_kernel<CH, false>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess)
            return e;
    }
    HOC_LAUNCH(HOC_K_RASTER_BWD_LINE, st,
               (hoc_launch_pdl(kernel, dim3(B + k4_samples, (S + lines - 1) / lines), g_line_threads, smem, st, faces,
                               face_index_map, rgb, grad_rgb, g_alpha, F, S, eps, layout, use_alpha, w.ext,
                               2.0f / (float)S, grad_faces, w.det_gf, k4_samples, g_channels, weight_map, depth_map,
                               grad_textures, w.det_gt, lines, g_line_fold)));
    return cudaSuccess;
}

extern "C" size_t hoc_raster_backward_workspace_bytes(int B, int F, int S)
{
    if (B <= 0 || S <= 0 || F < 0)
        return 0;
    return hoc_bwd_workspace(nullptr, B, F, S).total;
}

/* Leading bytes of the workspace that must be zero when hoc_raster_backward_ex is told HOC_BWD_WORKSPACE_ZEROED
 * (line spans, counters, per-face depth sums). */
extern "C" size_t hoc_raster_backward_zero_bytes(int B, int F, int S)
{
    if (B <= 0 || S <= 0 || F < 0)
        return 0;
    const HocBwdWorkspace w = hoc_bwd_workspace(nullptr, B, F, S);
    return (w.count_bytes + w.acc_bytes + 15) & ~(size_t)15;
}

/* The same for hoc_pair_backward_raster, which never has a depth gradient: the line spans and counters only. */
extern "C" size_t hoc_pair_backward_zero_bytes(int n, int F, int S)
{
    if (n <= 0 || S <= 0 || F < 0)
        return 0;
    return (hoc_bwd_workspace(nullptr, n, F, S).count_bytes + 15) & ~(size_t)15;
}

/* Workspace of hoc_raster_backward for a given texture size: in the reproducible mode (HOC_TUNE_DETERMINISTIC) the
 * fixed-point accumulators of grad_textures ([B,F,ts^3,3] or, in HOC_TEX_GRAD_VERTEX mode, [B,F,3,3]) live in it. */
extern "C" size_t hoc_raster_backward_workspace_bytes_ex(int B, int F, int S, int ts, int tex_grad_mode)
{
    if (B <= 0 || S <= 0 || F < 0)
        return 0;
    const int tex_n = (tex_grad_mode == HOC_TEX_GRAD_VERTEX) ? 9 : 3 * ts * ts * ts;
    return hoc_bwd_workspace(nullptr, B, F, S, tex_n, g_hoc_deterministic != 0).total;
}

extern "C" int hoc_raster_backward_ex(const float *faces, const float *textures, const int32_t *face_index_map,
                                      const float *rgb, const float *weight_map, const float *depth,
                                      const float *grad_rgb, const float *grad_alpha, const float *grad_depth, int B,
                                      int F, int S, int ts, float near_, float far_, float eps, int layout,
                                      int use_alpha, int tex_grad_mode, int geom_samples, int flags, void *extra_zero,
                                      size_t extra_zero_bytes, const int *row_lo, float *grad_faces,
                                      float *grad_textures, void *workspace, size_t workspace_bytes, void *stream);

extern "C" int hoc_raster_backward(const float *faces, const float *textures, const int32_t *face_index_map,
                                   const float *rgb, const float *weight_map, const float *depth,
                                   const float *grad_rgb, const float *grad_alpha,
                                   const float *grad_depth, int B, int F, int S, int ts, float near_, float far_,
                                   float eps, int layout, int use_alpha, int tex_grad_mode, float *grad_faces,
                                   float *grad_textures, void *workspace, size_t workspace_bytes, void *stream)
{
    return hoc_raster_backward_ex(faces, textures, face_index_map, rgb, weight_map, depth, grad_rgb, grad_alpha,
                                  grad_depth, B, F, S, ts, near_, far_, eps, layout, use_alpha, tex_grad_mode, B, 0,
                                  nullptr, 0, nullptr, grad_faces, grad_textures, workspace, workspace_bytes, stream);
}

/* geom_samples: the pseudo-gradient (backward_pixel_map) is computed for samples [0, geom_samples) only; the rows of
 * grad_faces of the other samples receive the depth gradient alone (zero without grad_depth).  The frame-pair path
 * stacks both renders of a pair in one batch and needs the geometry gradient of the first one only. */
static int hoc_raster_backward_impl(const float *faces, const float *textures, const int32_t *face_index_map,
                                    const float *rgb, const float *weight_map, const float *depth,
                                    const float *grad_rgb, const float *grad_alpha, const float *grad_depth, int B,
                                    int F, int S, int ts, float near_, float far_, float eps, int layout,
                                    int use_alpha, int tex_grad_mode, int geom_samples, int flags, void *extra_zero,
                                    size_t extra_zero_bytes, const int *row_lo, float *grad_faces,
                                    float *grad_textures, void *workspace, size_t workspace_bytes, void *stream,
                                    const HocPairGradSrc *pair_src, float *grad_rgb_out);

extern "C" int hoc_raster_backward_ex(const float *faces, const float *textures, const int32_t *face_index_map,
                                      const float *rgb, const float *weight_map, const float *depth,
                                      const float *grad_rgb, const float *grad_alpha, const float *grad_depth, int B,
                                      int F, int S, int ts, float near_, float far_, float eps, int layout,
                                      int use_alpha, int tex_grad_mode, int geom_samples, int flags, void *extra_zero,
                                      size_t extra_zero_bytes, const int *row_lo, float *grad_faces,
                                      float *grad_textures, void *workspace, size_t workspace_bytes, void *stream)
{
    /* (bench.py can time the whole call with one event pair: HOC_K_RASTER_BWD_GROUP) */
    hoc_note_launch(HOC_K_RASTER_BWD_GROUP, (cudaStream_t)stream, 0);
    const int rc = hoc_raster_backward_impl(faces, textures, face_index_map, rgb, weight_map, depth, grad_rgb, grad_alpha,
                                            grad_depth, B, F, S, ts, near_, far_, eps, layout, use_alpha, tex_grad_mode,
                                            geom_samples, flags, extra_zero, extra_zero_bytes, row_lo, grad_faces,
                                            grad_textures, workspace, workspace_bytes, stream, nullptr, nullptr);
    hoc_note_launch(HOC_K_RASTER_BWD_GROUP, (cudaStream_t)stream, 1);
    return rc;
}

/* The backward of the frame-pair step from the loss to grad_faces / grad_textures in one call: the backward of
 * pair_consist (hoc_warp_photo_backward_pair) fused into the rasterizer backward's scan pass
 * (hoc_raster_bwd_scan_pair_kernel), then the line pass.  grad_rgb [n,3,S,S] is scratch of the call (two planes used).
 * The stacked batch holds rows [row_offset, row_offset + n) of (render 1 of every pair, render 2 of every pair). */
extern "C" int hoc_pair_backward_raster(const float *image_ref, const float *image, const float *flow12,
                                        const float *flow21, const uint8_t *const *valid_mask, const double *sums,
                                        const float *mult1, const float *mult2, const float *grad_loss,
                                        const float *grad_mean, int pairs, int H, int W, int use_backward,
                                        int row_offset, const float *faces, const int32_t *face_index_map,
                                        const float *rgb, const float *weight_map, const float *depth, float *grad_rgb,
                                        int n, int F, int S, float near_, float far_, float eps, int geom_samples,
                                        int flags, void *extra_zero, size_t extra_zero_bytes, const int *row_lo,
                                        float *grad_faces, float *grad_textures, void *workspace,
                                        size_t workspace_bytes, void *stream)
{
    HOC_CHECK_ARG(pairs >= 1 && n >= 1 && row_offset >= 0 && row_offset + n <= 2 * pairs,
                  "hoc_pair_backward_raster: rows [%d, %d) outside the stacked batch of %d pairs", row_offset,
                  row_offset + n, pairs);
    HOC_CHECK_ARG(S >= 4 && (S % 4) == 0 && H >= 1 && H <= S && W >= 4 && W <= S && (W % 4) == 0,
                  "hoc_pair_backward_raster: bad shape S=%d H=%d W=%d (S, W multiples of 4)", S, H, W);
    HOC_CHECK_ARG(image_ref && image && flow12 && flow21 && valid_mask && valid_mask[0] && valid_mask[1] && sums && mult1 &&
                      mult2 && (grad_loss || grad_mean) && grad_rgb && rgb,
                  "hoc_pair_backward_raster: NULL argument");
    HOC_CHECK_ARG((((uintptr_t)grad_rgb | (uintptr_t)face_index_map) & 15) == 0 &&
                      ((((uintptr_t)valid_mask[0] | (uintptr_t)valid_mask[1]) & 3) == 0),
                  "hoc_pair_backward_raster: tensors must be 16-byte aligned");
    HocPairGradSrc G;
    /* direction 0: warp(image_ref, flow21) vs image -> d / d flow21 -> render 2; direction 1: the reverse */
    G.dir[0].src = image_ref; G.dir[0].target = image; G.dir[0].flow = flow21; G.dir[0].mult = mult2;
    G.dir[0].valid_mask = valid_mask[0]; G.dir[0].sums = sums; G.dir[0].grad_rgb = nullptr; G.dir[0].grad_flow = nullptr;
    G.dir[0].active = 1;
    G.dir[1].src = image; G.dir[1].target = image_ref; G.dir[1].flow = flow12; G.dir[1].mult = mult1;
    G.dir[1].valid_mask = valid_mask[1]; G.dir[1].sums = sums + 2 * (size_t)pairs; G.dir[1].grad_rgb = nullptr;
    G.dir[1].grad_flow = nullptr; G.dir[1].active = use_backward ? 1 : 0;
    G.grad_loss = grad_loss; G.grad_mean = grad_mean;
    G.pairs = pairs; G.H = H; G.W = W; G.row_offset = row_offset;
    G.inv_w = 1.0f / (float)(W - 1 > 1 ? W - 1 : 1);
    G.inv_h = 1.0f / (float)(H - 1 > 1 ? H - 1 : 1);
    hoc_note_launch(HOC_K_RASTER_BWD_GROUP, (cudaStream_t)stream, 0);
    const int rc = hoc_raster_backward_impl(faces, nullptr, face_index_map, rgb, weight_map, depth, grad_rgb, nullptr,
                                            nullptr, n, F, S, 2, near_, far_, eps, HOC_LAYOUT_IMAGE, 1, HOC_TEX_GRAD_VERTEX,
                                            geom_samples, flags, extra_zero, extra_zero_bytes, row_lo, grad_faces,
                                            grad_textures, workspace, workspace_bytes, stream, &G, grad_rgb);
    hoc_note_launch(HOC_K_RASTER_BWD_GROUP, (cudaStream_t)stream, 1);
    return rc;
}

static int hoc_raster_backward_impl(const float *faces, const float *textures, const int32_t *face_index_map,
                                    const float *rgb, const float *weight_map, const float *depth,
                                    const float *grad_rgb, const float *grad_alpha, const float *grad_depth, int B,
                                    int F, int S, int ts, float near_, float far_, float eps, int layout,
                                    int use_alpha, int tex_grad_mode, int geom_samples, int flags, void *extra_zero,
                                    size_t extra_zero_bytes, const int *row_lo, float *grad_faces,
                                    float *grad_textures, void *workspace, size_t workspace_bytes, void *stream,
                                    const HocPairGradSrc *pair_src, float *grad_rgb_out)
{
    (void)textures;
    HOC_CHECK_ARG(extra_zero == nullptr || (extra_zero_bytes % 4 == 0 && ((uintptr_t)extra_zero & 3) == 0),
                  "hoc_raster_backward: extra_zero must be a float buffer");
    if (extra_zero == nullptr)
        extra_zero_bytes = 0;
    HOC_CHECK_ARG(geom_samples >= 0 && geom_samples <= B, "hoc_raster_backward: geom_samples %d outside [0, %d]",
                  geom_samples, B);
    HOC_CHECK_ARG(tex_grad_mode == HOC_TEX_GRAD_CUBE || (tex_grad_mode == HOC_TEX_GRAD_VERTEX && ts == 2),
                  "hoc_raster_backward: tex_grad_mode %d (vertex mode needs texture_size 2, got %d)", tex_grad_mode, ts);
    HOC_CHECK_ARG(B >= 0 && F >= 0, "hoc_raster_backward: negative batch (%d) or face count (%d)", B, F);
    HOC_CHECK_ARG(S >= 1 && S <= 2048, "hoc_raster_backward: image_size %d outside [1, 2048]", S);
    HOC_CHECK_ARG(layout == HOC_LAYOUT_RAW || layout == HOC_LAYOUT_IMAGE, "hoc_raster_backward: bad layout %d",
                  layout);
    HOC_CHECK_ARG(B <= 65535, "hoc_raster_backward: batch %d exceeds 65535", B); /* grid.y / grid.z limits */
    HOC_CHECK_ARG(F < (1 << 29), "hoc_raster_backward: face count %d exceeds 2^29", F);
    HOC_CHECK_ARG(grad_textures == nullptr || ts >= 1, "hoc_raster_backward: texture_size %d", ts);
    HOC_CHECK_ARG(grad_rgb == nullptr || rgb != nullptr, "hoc_raster_backward: grad_rgb given without rgb");
    if (B == 0 || F == 0)
        return HOC_OK;
    HOC_CHECK_ARG(faces != nullptr && face_index_map != nullptr, "hoc_raster_backward: faces / face_index_map NULL");
    if (grad_faces == nullptr && grad_textures == nullptr)
        return HOC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool det = g_hoc_deterministic != 0;
    const int tex_n = (tex_grad_mode == HOC_TEX_GRAD_VERTEX) ? 9 : 3 * ts * ts * ts;
    const HocBwdWorkspace w = hoc_bwd_workspace(workspace, B, F, S, tex_n, det);
    if (workspace == nullptr || workspace_bytes < w.total) {
        hoc_set_error("hoc_raster_backward: workspace of %zu bytes needed%s, %zu given", w.total,
                      det ? " (reproducible mode: query hoc_raster_backward_workspace_bytes_ex)" : "", workspace_bytes);
        return HOC_ERR_WORKSPACE;
    }
    const float *g_alpha = use_alpha ? grad_alpha : nullptr;
    const int k4_samples = (grad_faces != nullptr && (grad_rgb != nullptr || g_alpha != nullptr)) ? geom_samples : 0;
    const bool k4 = k4_samples > 0;
    const bool want_depth = grad_faces != nullptr && grad_depth != nullptr;
    /* When the textures are three vertex values per face and the forward saved its weights and depth (the frame-pair
     * path), the ROW CTAs of the line pass -- they stage every pixel's gradient anyway -- run backward_textures too,
     * for every sample: no cover pass at all, nobody reads the list of covered pixels */
    const bool tex_in_line = k4 && !want_depth && grad_rgb != nullptr && grad_textures != nullptr &&
                             tex_grad_mode == HOC_TEX_GRAD_VERTEX && weight_map != nullptr && depth != nullptr &&
                             g_tex_in_line;
    /* frame-pair path: the rendered image is a 2-channel flow, the incoming gradient of the third channel is zero; when
     * the line pass is the only reader of the gradient planes the third one is neither written nor read */
    const int g_channels = (pair_src != nullptr && tex_in_line) ? 2 : 3;
    const int scan_flags = (tex_in_line ? (HOC_SCAN_SPAN_ALL | HOC_SCAN_NO_LIST) : 0) |
                           (g_channels == 2 ? HOC_SCAN_TWO_CHANNELS : 0);
    const size_t tex_bytes = (tex_grad_mode == HOC_TEX_GRAD_VERTEX)
                                 ? sizeof(float) * 9 * (size_t)B * F
                                 : sizeof(float) * 3 * (size_t)ts * ts * ts * (size_t)B * F;

    /* spans, counters and (directly behind them) acc_d are zero-filled by ONE memset; the gradient outputs are
     * zero-filled by the scan pass */
    /* (HOC_BWD_WORKSPACE_ZEROED: an earlier kernel of the caller's sequence did it -- one graph node less) */
    cudaError_t e = (flags & HOC_BWD_WORKSPACE_ZEROED)
                        ? cudaSuccess
                        : cudaMemsetAsync(w.ext, 0, w.count_bytes + (want_depth ? w.acc_bytes : 0), st);
    if (e == cudaSuccess && det)
        e = cudaMemsetAsync(w.det_gf, 0, w.det_bytes, st);
    if (e != cudaSuccess) {
        hoc_set_error("hoc_raster_backward: memset failed: %s", cudaGetErrorString(e));
        return HOC_ERR_CUDA;
    }
    const long n_gf = (grad_faces != nullptr) ? 9l * B * F : 0;
    const long n_gt = (grad_textures != nullptr) ? (long)(tex_bytes / sizeof(float)) : 0;
    float *gt = (grad_rgb != nullptr) ? grad_textures : nullptr;
    {
        /* without the pseudo-gradient only pixels with a texture gradient (non-zero dL/drgb) or a depth
         * gradient have work */
        const uintptr_t al = (uintptr_t)face_index_map | (uintptr_t)grad_rgb | (uintptr_t)g_alpha;
        if (pair_src != nullptr) { /* frame-pair path: the incoming gradient is computed by the scan pass itself */
            dim3 pg4(B, (S + 127) / 128, (S + 7) / 8);
            /* the three zero-filled outputs as one region when the caller laid them out back to back */
            float *za = grad_faces, *zb = grad_textures, *zc = (float *)extra_zero;
            long na = n_gf, nb = n_gt, nc = (long)(extra_zero_bytes / sizeof(float));
            if (zb != nullptr && zc == zb + nb) {
                nb += nc;
                nc = 0;
            }
            if (za != nullptr && zb == za + na) {
                na += nb;
                nb = 0;
            }
            HOC_LAUNCH(HOC_K_RASTER_BWD_PIXEL, st,
                       (hoc_launch_pdl((hoc_raster_bwd_scan_pair_kernel), pg4, 256, 0, st, 
                           face_index_map, *pair_src, grad_rgb_out, S, k4_samples, want_depth ? 1 : 0, scan_flags, w.ext,
                           w.cov_count, w.cov_list, za, na, zb, nb, zc, nc, row_lo)));
            HOC_CHECK_LAUNCH("hoc_raster_bwd_scan_pair_kernel");
        } else if (layout == HOC_LAYOUT_IMAGE && (S % 4) == 0 && (al & 15) == 0) {
            dim3 pg4((S + 127) / 128, (S + 7) / 8, B);
            HOC_LAUNCH(HOC_K_RASTER_BWD_PIXEL, st,
                       (hoc_raster_bwd_scan4_kernel<<<pg4, 256, 0, st>>>(
                           face_index_map, grad_rgb, gt != nullptr ? grad_rgb : nullptr, g_alpha, S, k4_samples,
                           want_depth ? 1 : 0, scan_flags, w.ext, w.cov_count, w.cov_list, grad_faces, n_gf, grad_textures, n_gt,
                           (float *)extra_zero, (long)(extra_zero_bytes / sizeof(float)), row_lo)));
            HOC_CHECK_LAUNCH("hoc_raster_bwd_scan4_kernel");
        } else {
        dim3 pg((S + 31) / 32, (S + 31) / 32, B);
        HOC_LAUNCH(HOC_K_RASTER_BWD_PIXEL, st,
                   (hoc_raster_bwd_scan_kernel<<<pg, dim3(32, 8), 0, st>>>(
                       face_index_map, grad_rgb, gt != nullptr ? grad_rgb : nullptr, g_alpha, S, layout, k4_samples,
                       want_depth ? 1 : 0, scan_flags, w.ext, w.cov_count, w.cov_list, grad_faces, n_gf, grad_textures, n_gt,
                       (float *)extra_zero, (long)(extra_zero_bytes / sizeof(float)), row_lo)));
        HOC_CHECK_LAUNCH("hoc_raster_bwd_scan_kernel");
        }
    }
    if (!tex_in_line && (gt != nullptr || want_depth)) { /* cover pass: texture / depth gradient of the listed pixels */
        const long npix = (long)S * S;
        dim3 cg((unsigned)((npix + CV_THREADS - 1) / CV_THREADS < g_cover_ctas ? (npix + CV_THREADS - 1) / CV_THREADS
                                                                              : g_cover_ctas),
                B);
#define HOC_COVER_LAUNCH(TS2)                                                                                        \
    HOC_LAUNCH(HOC_K_RASTER_BACKWARD_COVER, st,                                                                      \
               (hoc_launch_pdl((hoc_raster_bwd_cover_kernel<TS2>), cg, CV_THREADS, 0, st, faces, weight_map, depth,  \
                               grad_rgb, grad_depth, F, S, ts, near_, far_, eps, layout, tex_grad_mode, w.cov_count, \
                               w.cov_list, want_depth ? w.acc_d : nullptr, gt, w.det_gt, w.det_ad)))
        if (ts == 2)
            HOC_COVER_LAUNCH(true);
        else
            HOC_COVER_LAUNCH(false);
#undef HOC_COVER_LAUNCH
        HOC_CHECK_LAUNCH("hoc_raster_bwd_cover_kernel");
    }
    const long nfaces = (long)B * F;
    if (det && gt != nullptr && !tex_in_line && hoc_det_flush(w.det_gt, nfaces * tex_n, gt, 0, st) != cudaSuccess) {
        hoc_set_error("hoc_raster_backward: flush of the texture accumulators failed");
        return HOC_ERR_CUDA;
    }
    if (grad_faces == nullptr)
        return HOC_OK;
    if (det && want_depth && hoc_det_flush(w.det_ad, nfaces * 3, w.acc_d, 0, st) != cudaSuccess) {
        hoc_set_error("hoc_raster_backward: flush of the depth accumulators failed");
        return HOC_ERR_CUDA;
    }
    /* reproducible mode: the per-face depth epilogue (a plain `+=`) runs after the flush of grad_faces below */
    if (want_depth && !det) {
        const long nf = (long)B * F;
        HOC_LAUNCH(HOC_K_RASTER_BACKWARD, st,
                   (hoc_raster_bwd_depth_kernel<<<(unsigned)((nf + 255) / 256), 256, 0, st>>>(faces, w.acc_d, nf, S,
                                                                                              grad_faces)));
        HOC_CHECK_LAUNCH("hoc_raster_bwd_depth_kernel");
    }
    if (k4) {
        const int n_line = tex_in_line ? B : k4_samples; /* (texture gradient: row CTAs of every sample) */
        const int seg = g_line_seg > 0 ? g_line_seg : (S > 320 ? 32 : 16);
        if (seg >= 32)
            e = hoc_launch_line<32>(faces, face_index_map, rgb, grad_rgb, g_alpha, n_line, k4_samples, F, S, eps, layout,
                                    use_alpha, w, grad_faces, tex_in_line ? weight_map : nullptr, depth,
                                    tex_in_line ? gt : nullptr, g_channels, st);
        else if (seg >= 16)
            e = hoc_launch_line<16>(faces, face_index_map, rgb, grad_rgb, g_alpha, n_line, k4_samples, F, S, eps, layout,
                                    use_alpha, w, grad_faces, tex_in_line ? weight_map : nullptr, depth,
                                    tex_in_line ? gt : nullptr, g_channels, st);
        else
            e = hoc_launch_line<8>(faces, face_index_map, rgb, grad_rgb, g_alpha, n_line, k4_samples, F, S, eps, layout,
                                   use_alpha, w, grad_faces, tex_in_line ? weight_map : nullptr, depth,
                                   tex_in_line ? gt : nullptr, g_channels, st);
        if (e != cudaSuccess) {
            hoc_set_error("hoc_raster_backward: line pass: %s", cudaGetErrorString(e));
            return HOC_ERR_CUDA;
        }
        HOC_CHECK_LAUNCH("hoc_raster_bwd_line_kernel");
    }
    if (det && tex_in_line && hoc_det_flush(w.det_gt, nfaces * tex_n, gt, 0, st) != cudaSuccess) {
        hoc_set_error("hoc_raster_backward: flush of the texture accumulators failed");
        return HOC_ERR_CUDA;
    }
    if (det) {
        if (k4 && hoc_det_flush(w.det_gf, nfaces * 9, grad_faces, 0, st) != cudaSuccess) {
            hoc_set_error("hoc_raster_backward: flush of the face accumulators failed");
            return HOC_ERR_CUDA;
        }
        if (want_depth) {
            HOC_LAUNCH(HOC_K_RASTER_BACKWARD, st,
                       (hoc_raster_bwd_depth_kernel<<<(unsigned)((nfaces + 255) / 256), 256, 0, st>>>(faces, w.acc_d, nfaces,
                                                                                                  S, grad_faces)));
            HOC_CHECK_LAUNCH("hoc_raster_bwd_depth_kernel");
        }
    }
    return HOC_OK;
}
