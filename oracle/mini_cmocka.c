/* TEST INFRASTRUCTURE (oracle): runner behind oracle/shim/cmocka.h */
#include "shim/cmocka.h"
#include <string.h>
/* KSN_CM_SKIP / KSN_CM_ONLY: comma-separated test names to leave out / to run exclusively */
static int listed(const char *env, const char *name)
{
    const char *v = getenv(env);
    const size_t n = strlen(name);
    for (const char *p = v; p && *p; ) {
        const char *e = strchr(p, ',');
        const size_t len = e ? (size_t) (e - p) : strlen(p);
        if (len == n && !strncmp(p, name, n)) return 1;
        p = e ? e + 1 : NULL;
    }
    return 0;
}
jmp_buf ksn_cm_jmp;
int ksn_cm_failed;
int ksn_cm_run(const struct CMUnitTest *t, size_t n, int (*setup)(void **), int (*teardown)(void **))
{
    void *state = NULL;
    int nfail = 0;
    if (setup && setup(&state)) { printf("[  ERROR   ] group setup failed\n"); return 255; }
    for (size_t i = 0; i < n; i++) {
        ksn_cm_failed = 0;
        if (listed("KSN_CM_SKIP", t[i].name) || (getenv("KSN_CM_ONLY") && !listed("KSN_CM_ONLY", t[i].name))) {
            printf("[  SKIPPED ] %s\n", t[i].name);
            continue;
        }
        printf("[ RUN      ] %s\n", t[i].name);
        fflush(stdout);
        if (!setjmp(ksn_cm_jmp)) t[i].fn(&state);
        printf(ksn_cm_failed ? "[  FAILED  ] %s\n" : "[       OK ] %s\n", t[i].name);
        nfail += ksn_cm_failed;
    }
    if (teardown) teardown(&state);
    printf("[==========] %zu test(s) run, %d failed\n", n, nfail);
    return nfail;
}
