/* hoc_abi.cu -- version / error reporting / launch accounting of the C ABI (include/hoc_b200.h). */
#include <stdarg.h>
#include <stdio.h>

#include <atomic>
#include <mutex>

#include "hoc_common.cuh"
#include "hoc_det.cuh"

/* reproducible accumulation (hoc_det.cuh): set by hoc_set_tuning(HOC_TUNE_DETERMINISTIC, 0 / 1) */
int g_hoc_deterministic = 0;
int g_hoc_pdl = 0; /* programmatic dependent launch of the frame-pair kernels (HOC_TUNE_PDL) */

__global__ void __launch_bounds__(256)
hoc_det_flush_kernel(const unsigned long long *__restrict__ det, long n, float *__restrict__ dst, int add)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(det + 2 * i);
    const float v = hoc_fix128_to_float(a.x, a.y);
    dst[i] = add ? dst[i] + v : v;
}

cudaError_t hoc_det_flush(const unsigned long long *det, long n, float *dst, int add, cudaStream_t st)
{
    if (n <= 0)
        return cudaSuccess;
    hoc_det_flush_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(det, n, dst, add);
    return cudaGetLastError();
}

static thread_local char g_hoc_error[512] = "";

void hoc_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_hoc_error, sizeof(g_hoc_error), fmt, ap);
    va_end(ap);
}

extern "C" int hoc_abi_version(void) { return HOC_ABI_VERSION; }

extern "C" const char *hoc_last_error(void) { return g_hoc_error; }

/* ---- launch accounting + per-kernel device timing (used by bench.py) ---------------------- */
#define HOC_TIMER_CAP 8192
/* launches come from the Python thread (forward), the autograd engine thread (backward) and, with several devices, one
 * thread per device: the counters are atomics, the timer table is guarded by a mutex (taken only while a timer is armed) */
static std::atomic<unsigned long long> g_launches[HOC_KERNEL_COUNT];
static std::atomic<unsigned long long> g_timer_mask{0};
static std::mutex g_timer_mutex;
static thread_local int g_timer_slot[2] = {-1, -1}; /* slot reserved by phase 0 of this thread's launch in flight
                                                       ([1]: a group id bracketing several launches, nested around [0]) */
static int g_timer_n = 0;
static cudaEvent_t g_timer_ev[HOC_TIMER_CAP][2];
static int g_timer_id[HOC_TIMER_CAP];
static int g_timer_created = 0;

/* Inside a stream capture the record becomes an EXTERNAL event node of the graph: it is re-recorded by every
 * replay and cudaEventElapsedTime on it is valid, so kernels can be timed where they run in production. */
static void hoc_timer_record(cudaEvent_t ev, cudaStream_t st)
{
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusActive)
        cudaEventRecordWithFlags(ev, st, cudaEventRecordExternal);
    else
        cudaEventRecord(ev, st);
}

void hoc_note_launch(int kernel_id, cudaStream_t st, int phase)
{
    const int lvl = kernel_id >= HOC_K_RASTER_BWD_GROUP ? 1 : 0;
    if (phase == 0 && lvl == 0)
        g_launches[kernel_id].fetch_add(1, std::memory_order_relaxed);
    if (phase == 0) {
        g_timer_slot[lvl] = -1;
        if (!((g_timer_mask.load(std::memory_order_relaxed) >> kernel_id) & 1ull))
            return;
        std::lock_guard<std::mutex> lock(g_timer_mutex);
        if (g_timer_n >= HOC_TIMER_CAP)
            return;
        const int slot = g_timer_n++; /* reserved: a concurrent launch from another thread gets the next one */
        while (g_timer_created <= slot) {
            cudaEventCreate(&g_timer_ev[g_timer_created][0]);
            cudaEventCreate(&g_timer_ev[g_timer_created][1]);
            g_timer_created++;
        }
        g_timer_id[slot] = kernel_id;
        g_timer_slot[lvl] = slot;
        hoc_timer_record(g_timer_ev[slot][0], st);
    } else if (g_timer_slot[lvl] >= 0) {
        hoc_timer_record(g_timer_ev[g_timer_slot[lvl]][1], st);
        g_timer_slot[lvl] = -1;
    }
}

extern "C" unsigned long long hoc_launch_count(int kernel_id)
{
    if (kernel_id < 0) {
        unsigned long long t = 0;
        for (int k = 0; k < HOC_KERNEL_COUNT; k++)
            t += g_launches[k].load(std::memory_order_relaxed);
        return t;
    }
    return kernel_id < HOC_KERNEL_COUNT ? g_launches[kernel_id].load(std::memory_order_relaxed) : 0;
}

extern "C" int hoc_timer_begin(unsigned long long kernel_mask)
{
    std::lock_guard<std::mutex> lock(g_timer_mutex);
    g_timer_mask = kernel_mask;
    g_timer_n = 0;
    return HOC_OK;
}

extern "C" int hoc_timer_pause(void)
{
    g_timer_mask = 0; /* stop bracketing new launches, keep the recorded event pairs */
    return g_timer_n;
}

extern "C" int hoc_timer_peek(float *ms_host, int *kernel_ids_host, int capacity)
{
    const int n = g_timer_n < capacity ? g_timer_n : capacity;
    for (int i = 0; i < n; i++) {
        cudaEventSynchronize(g_timer_ev[i][1]);
        if (cudaEventElapsedTime(&ms_host[i], g_timer_ev[i][0], g_timer_ev[i][1]) != cudaSuccess)
            ms_host[i] = -1.0f;
        if (kernel_ids_host != nullptr)
            kernel_ids_host[i] = g_timer_id[i];
    }
    return n;
}

extern "C" int hoc_timer_end(float *ms_host, int *kernel_ids_host, int capacity)
{
    const int n = g_timer_n < capacity ? g_timer_n : capacity;
    for (int i = 0; i < n; i++) {
        cudaEventSynchronize(g_timer_ev[i][1]);
        if (cudaEventElapsedTime(&ms_host[i], g_timer_ev[i][0], g_timer_ev[i][1]) != cudaSuccess)
            ms_host[i] = -1.0f;
        if (kernel_ids_host != nullptr)
            kernel_ids_host[i] = g_timer_id[i];
    }
    g_timer_mask = 0;
    g_timer_n = 0;
    return n;
}
