/* TEST INFRASTRUCTURE (oracle): runner behind oracle/shim/cmocka.h */
#include "shim/cmocka.h"
jmp_buf ksn_cm_jmp;
int ksn_cm_failed;
int ksn_cm_run(const struct CMUnitTest *t, size_t n, int (*setup)(void **), int (*teardown)(void **))
{
    void *state = NULL;
    int nfail = 0;
    if (setup && setup(&state)) { printf("[  ERROR   ] group setup failed\n"); return 255; }
    for (size_t i = 0; i < n; i++) {
        ksn_cm_failed = 0;
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
