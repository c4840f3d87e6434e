// capi.cu -- error plumbing of the C-ABI.
#include "common.cuh"
#include <stdarg.h>

static thread_local char g_err[512] = "";

void dagr_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *dagr_last_error(void) { return g_err; }
extern "C" int dagr_abi_version(void) { return DAGR_ABI_VERSION; }
