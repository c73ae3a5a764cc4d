/* hoc_abi.cu -- version / error reporting of the C ABI (include/hoc_b200.h). */
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
