/*
 * oracle/nmr_oracle.c -- TEST INFRASTRUCTURE ONLY.
 * Instantiates the CPU restatement of the Neural-Mesh-Renderer rasterizer (see
 * nmr_oracle_impl.h for provenance and the "parity unpinned" note) for float and double.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load the resulting library.  Build: see oracle/Makefile (gcc -O2 -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>

/* ---- tiny host thread pool: splits [0, n) into contiguous chunks (no OpenMP runtime in this image) ---- */
static int g_ora_threads = 1;
void nmr_oracle_set_threads(int n) { g_ora_threads = n < 1 ? 1 : (n > 256 ? 256 : n); }
int nmr_oracle_get_threads(void) { return g_ora_threads; }

typedef void (*ora_range_fn)(void *ctx, long begin, long end);
struct ora_worker_arg {
    ora_range_fn fn;
    void *ctx;
    long n, nchunks;
    int tid, nt;
};
/* worker t takes chunks t, t+nt, ...: many small chunks keep the load even (covered pixels cluster) */
static void *ora_worker(void *p)
{
    struct ora_worker_arg *w = (struct ora_worker_arg *)p;
    for (long c = w->tid; c < w->nchunks; c += w->nt) {
        const long b = w->n * c / w->nchunks, e = w->n * (c + 1) / w->nchunks;
        if (b < e)
            w->fn(w->ctx, b, e);
    }
    return 0;
}
static void ora_parallel_for(long n, ora_range_fn fn, void *ctx)
{
    const int nt = g_ora_threads;
    if (nt <= 1 || n < 2 * nt) {
        fn(ctx, 0, n);
        return;
    }
    pthread_t th[256];
    struct ora_worker_arg arg[256];
    for (int t = 0; t < nt; t++) {
        arg[t].fn = fn;
        arg[t].ctx = ctx;
        arg[t].n = n;
        arg[t].nchunks = (long)nt * 16;
        arg[t].tid = t;
        arg[t].nt = nt;
        pthread_create(&th[t], 0, ora_worker, &arg[t]);
    }
    for (int t = 0; t < nt; t++)
        pthread_join(th[t], 0);
}

#define ORA_IMIN(a, b) ((a) < (b) ? (a) : (b))
#define ORA_IMAX(a, b) ((a) > (b) ? (a) : (b))

#define REAL float
#define FN(name) nmr_##name##_f32
#include "nmr_oracle_impl.h"
#undef REAL
#undef FN

#define REAL double
#define FN(name) nmr_##name##_f64
#include "nmr_oracle_impl.h"
#undef REAL
#undef FN

int nmr_oracle_abi_version(void) { return 1; }
