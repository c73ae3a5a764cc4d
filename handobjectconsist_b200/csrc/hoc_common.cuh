/* hoc_common.cuh -- small helpers shared by the sm_100a kernels of libhoc_b200.so. */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/hoc_b200.h"

#define HOC_FULL_MASK 0xffffffffu

/* Raised by the C-ABI entry points; text is fetched with hoc_last_error(). */
void hoc_set_error(const char *fmt, ...);

/* Counts the launch and, when bench.py armed the timer for this kernel, brackets it with events
 * on the launching stream (phase 0 before the launch, 1 after). */
void hoc_note_launch(int kernel_id, cudaStream_t st, int phase);

#define HOC_LAUNCH(kernel_id, st, ...)    \
    do {                                  \
        hoc_note_launch(kernel_id, st, 0); \
        __VA_ARGS__;                      \
        hoc_note_launch(kernel_id, st, 1); \
    } while (0)

/* Programmatic dependent launch (sm_90+): the kernels of the frame-pair step are launched with the "programmatic
 * stream serialization" attribute and start with hoc_pdl_sync() -- `griddepcontrol.launch_dependents` (the NEXT kernel
 * of the stream may be scheduled as soon as every CTA of this one has started) then `griddepcontrol.wait` (blocks until
 * the PREVIOUS kernel has completed and its writes are visible).  No kernel touches global memory before its wait, so
 * the stream's semantics are unchanged; what overlaps is the launch, the CTA scheduling and the prologue of kernel
 * N + 1 with the tail of kernel N (a captured graph keeps the edges as programmatic dependencies).  Without the
 * attribute both instructions are no-ops.  hoc_set_tuning(HOC_TUNE_PDL, 0) switches the attribute off. */
extern int g_hoc_pdl;

#ifdef __CUDACC__
__device__ __forceinline__ void hoc_pdl_sync()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

template <typename... KArgs, typename... Args>
static inline cudaError_t hoc_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                         Args &&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = g_hoc_pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

#define HOC_CHECK_ARG(cond, ...)        \
    do {                                \
        if (!(cond)) {                  \
            hoc_set_error(__VA_ARGS__); \
            return HOC_ERR_INVALID_ARG; \
        }                               \
    } while (0)

#define HOC_CHECK_LAUNCH(name)                                                    \
    do {                                                                          \
        cudaError_t e__ = cudaGetLastError();                                     \
        if (e__ != cudaSuccess) {                                                 \
            hoc_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
            return HOC_ERR_CUDA;                                                  \
        }                                                                         \
    } while (0)

/* Address helpers for the two output layouts (see include/hoc_b200.h):
 *   HOC_LAYOUT_RAW      rgb [B,S,S,3], planes [B,S,S], raster row order (row 0 = bottom)
 *   HOC_LAYOUT_IMAGE    rgb [B,3,S,S], planes [B,S,S], rows flipped (row 0 = top)            */
__device__ __forceinline__ long hoc_plane_off(int layout, int S, int b, int yi, int xi)
{
    const int row = (layout == HOC_LAYOUT_IMAGE) ? (S - 1 - yi) : yi;
    return ((long)b * S + row) * S + xi;
}

__device__ __forceinline__ long hoc_rgb_off(int layout, int S, int b, int yi, int xi, int c)
{
    if (layout == HOC_LAYOUT_IMAGE)
        return (((long)b * 3 + c) * S + (S - 1 - yi)) * S + xi;
    return (((long)b * S + yi) * S + xi) * 3 + c;
}

/* k-th element of the centre-out order of [0, n): c, c-1, c+1, c-2, ... (a bijection).  Pixel passes hand their tiles
 * to CTAs in this order, sample fastest: the tiles that hold the meshes (centred by the crop) carry the heavy work and
 * start first, the empty border tiles drain last -- instead of the last sample's mesh tiles being the tail. */
__device__ __forceinline__ int hoc_centre_out(int k, int n)
{
    return (n >> 1) + ((k & 1) ? -((k + 1) >> 1) : (k >> 1));
}

/* q / d (remainder in *rem) for 0 <= q < 2^22, 1 <= d <= 2^11 without the ~20-instruction integer division: a float
 * estimate (one MUFU.RCP) and one correction step (the estimate is off by at most one). */
__device__ __forceinline__ int hoc_div_small(int q, int d, int *rem)
{
    float inv;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"((float)d));
    int k = (int)(((float)q + 0.5f) * inv);
    int r = q - k * d;
    if (r < 0) {
        k--;
        r += d;
    } else if (r >= d) {
        k++;
        r -= d;
    }
    *rem = r;
    return k;
}

/* Grid-wide fill of n 16-byte words with the 32-bit pattern v, spread over nthreads threads (t0: this thread's linear
 * id).  The loop is controlled in 32 bits whenever the sizes allow: the 64-bit form of these loops (compare chains of
 * the unrolled body) was 35 % of the scan pass's instructions. */
__device__ __forceinline__ void hoc_fill16(uint4 *__restrict__ p, long n, long t0, long nthreads, unsigned v)
{
    const uint4 w = make_uint4(v, v, v, v);
    if (n < 0x7fffffffl && nthreads < 0x3fffffffl) {
        const unsigned n32 = (unsigned)(n > 0 ? n : 0), step = (unsigned)nthreads;
#pragma unroll 1
        for (unsigned i = (unsigned)t0; i < n32; i += step)
            p[i] = w;
    } else {
#pragma unroll 1
        for (long i = t0; i < n; i += nthreads)
            p[i] = w;
    }
}

/* Grid-wide zero-fill of n floats (any 4-byte alignment): scalar head and tail, 16-byte stores in between. */
__device__ __forceinline__ void hoc_zero_floats(float *__restrict__ p, long n, long t0, long nthreads)
{
    if (n <= 0)
        return;
    long head = (long)(((16u - (unsigned)((uintptr_t)p & 15u)) & 15u) >> 2);
    head = head < n ? head : n;
    if (t0 < head)
        p[t0] = 0.0f;
    float *q = p + head;
    const long m = n - head, m4 = m >> 2;
    hoc_fill16(reinterpret_cast<uint4 *>(q), m4, t0, nthreads, 0u);
    if (t0 < (m & 3))
        q[(m4 << 2) + t0] = 0.0f;
}

__device__ __forceinline__ float hoc_warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(HOC_FULL_MASK, v, o);
    return v;
}
