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

__device__ __forceinline__ float hoc_warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(HOC_FULL_MASK, v, o);
    return v;
}
