/* hoc_det.cuh -- order-independent accumulation for the reproducible mode (hoc_set_tuning(HOC_TUNE_DETERMINISTIC, 1)).
 *
 * The gradient sums of this library (per-face pseudo-gradient, texture / depth sums, per-vertex scatter) are
 * accumulated with float atomics in production, like the reference's own kernels (backward_textures,
 * backward_depth_map, index_put(accumulate=True)): the order of the additions changes from run to run and with it
 * the rounding.  In the reproducible mode every term is converted to a 128-bit fixed-point number (Q63.64: unit
 * 2^-64, range +-2^63) and added with two 64-bit INTEGER atomics.  Integer addition is associative, so the sum does
 * not depend on the order: two runs give the same bits, and so do a captured graph and the eager path.  A float
 * term of magnitude >= 2^-40 is represented exactly; smaller terms are truncated towards zero at 2^-64 (absolute
 * error < 5.5e-20 per term).  A second small kernel converts the accumulators to float (one rounding per sum).
 *
 * Cost: two 64-bit atomics per term instead of one 32-bit one plus the flush pass -- a mode for tests and for
 * debugging, not the production path.
 */
#pragma once
#include <cuda_runtime.h>

extern int g_hoc_deterministic; /* hoc_abi.cu; set by hoc_set_tuning */

__device__ __forceinline__ void hoc_fix128_add(unsigned long long *acc, float x)
{
    const unsigned bits = __float_as_uint(x);
    const int ex = (int)((bits >> 23) & 0xffu);
    unsigned long long m = (unsigned long long)((bits & 0x7fffffu) | (ex ? 0x800000u : 0u));
    if (m == 0ull || ex == 0xff)
        return; /* zero; inf / nan are dropped (the float path would poison the sum: the tests check finiteness there) */
    /* x = m * 2^(E - 150) with E = max(ex, 1); fixed = x * 2^64 = m * 2^(E - 86) */
    int sh = (ex ? ex : 1) - 86;
    unsigned long long lo, hi;
    if (sh > 102)
        sh = 102; /* saturate far beyond any gradient (|x| >= 2^62) */
    if (sh >= 64) {
        hi = m << (sh - 64);
        lo = 0ull;
    } else if (sh > 0) {
        lo = m << sh;
        hi = m >> (64 - sh);
    } else if (sh == 0) {
        lo = m;
        hi = 0ull;
    } else {
        lo = (sh > -24) ? (m >> (-sh)) : 0ull;
        hi = 0ull;
    }
    if (bits >> 31) { /* two's complement of the 128-bit magnitude */
        lo = ~lo + 1ull;
        hi = ~hi + (lo == 0ull ? 1ull : 0ull);
    }
    if (lo != 0ull) {
        const unsigned long long old = atomicAdd(acc, lo);
        if (old + lo < old)
            hi += 1ull; /* carry into the high word */
    }
    if (hi != 0ull)
        atomicAdd(acc + 1, hi);
}

__device__ __forceinline__ float hoc_fix128_to_float(unsigned long long lo, unsigned long long hi)
{
    const bool neg = (long long)hi < 0;
    if (neg) {
        lo = ~lo + 1ull;
        hi = ~hi + (lo == 0ull ? 1ull : 0ull);
    }
    const double v = ((double)hi * 18446744073709551616.0 + (double)lo) * 5.421010862427522170037e-20; /* 2^-64 */
    const float f = (float)v;
    return neg ? -f : f;
}

/* dst[i] += v: float atomic in production, fixed-point accumulator `det` (2 words per element of dst) otherwise */
__device__ __forceinline__ void hoc_accum(float *__restrict__ dst, long i, float v, unsigned long long *__restrict__ det)
{
    if (det != nullptr)
        hoc_fix128_add(det + 2 * i, v);
    else
        atomicAdd(dst + i, v);
}

/* dst[i] (+)= float(accumulator i) for i < n */
__global__ void __launch_bounds__(256)
hoc_det_flush_kernel(const unsigned long long *__restrict__ det, long n, float *__restrict__ dst, int add);

/* host: launch the flush on `st` */
cudaError_t hoc_det_flush(const unsigned long long *det, long n, float *dst, int add, cudaStream_t st);
