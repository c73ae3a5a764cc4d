/* hoc_abi.cu -- version / error reporting / launch accounting of the C ABI (include/hoc_b200.h). */
#include <stdarg.h>
#include <stdio.h>

#include "hoc_common.cuh"

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
static unsigned long long g_launches[HOC_KERNEL_COUNT];
static unsigned long long g_timer_mask = 0;
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
    if (phase == 0)
        g_launches[kernel_id]++;
    if (!((g_timer_mask >> kernel_id) & 1ull) || g_timer_n >= HOC_TIMER_CAP)
        return;
    if (phase == 0) {
        if (g_timer_n >= g_timer_created) {
            cudaEventCreate(&g_timer_ev[g_timer_n][0]);
            cudaEventCreate(&g_timer_ev[g_timer_n][1]);
            g_timer_created = g_timer_n + 1;
        }
        g_timer_id[g_timer_n] = kernel_id;
        hoc_timer_record(g_timer_ev[g_timer_n][0], st);
    } else {
        hoc_timer_record(g_timer_ev[g_timer_n][1], st);
        g_timer_n++;
    }
}

extern "C" unsigned long long hoc_launch_count(int kernel_id)
{
    if (kernel_id < 0) {
        unsigned long long t = 0;
        for (int k = 0; k < HOC_KERNEL_COUNT; k++)
            t += g_launches[k];
        return t;
    }
    return kernel_id < HOC_KERNEL_COUNT ? g_launches[kernel_id] : 0;
}

extern "C" int hoc_timer_begin(unsigned long long kernel_mask)
{
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
