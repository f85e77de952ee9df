/* error.c -- per-thread error message + version string. */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include "common.h"
#ifdef _OPENMP
#include <omp.h>
#endif

static __thread char g_err[512] = "";

int ehyb_fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

const char *ehyb_last_error(void) { return g_err; }
const char *ehyb_version(void) { return "ehyb-b200 0.1 (sm_100a)"; }

void ehyb_die(const char *where)
{
    fprintf(stderr, "%s: %s\n", where, g_err[0] ? g_err : "unknown error");
    abort();
}

void ehyb_free_host(void *p) { free(p); }

/* Threads of the host-side OpenMP code (format build, reorder, generators).  Launchers export
 * OMP_NUM_THREADS=1 to their workers (torchrun does); a caller that knows how many cores its
 * rank may use says so here.  n <= 0: leave the runtime's setting. */
void ehyb_set_host_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int ehyb_get_host_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
