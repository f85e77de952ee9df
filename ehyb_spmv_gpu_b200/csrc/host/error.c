/* error.c -- per-thread error message + version string. */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include "common.h"

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
