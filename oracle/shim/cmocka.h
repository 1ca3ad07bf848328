/* TEST INFRASTRUCTURE (oracle): the subset of cmocka the reference's *_test.c files use. */
#ifndef KSN_ORACLE_CMOCKA_SHIM_H
#define KSN_ORACLE_CMOCKA_SHIM_H
#include <stdio.h>
#include <stdlib.h>
#include <setjmp.h>
struct CMUnitTest { const char *name; void (*fn)(void **); };
#define cmocka_unit_test(f) { #f, f }
extern jmp_buf ksn_cm_jmp;
extern int ksn_cm_failed;
#define assert_true(c) do { if (!(c)) { printf("  ASSERT FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); ksn_cm_failed = 1; longjmp(ksn_cm_jmp, 1); } } while (0)
#define assert_false(c) assert_true(!(c))
#define assert_int_equal(a, b) assert_true((long long) (a) == (long long) (b))
int ksn_cm_run(const struct CMUnitTest *t, size_t n, int (*setup)(void **), int (*teardown)(void **));
#define cmocka_run_group_tests(tests, setup, teardown) ksn_cm_run(tests, sizeof(tests) / sizeof(tests[0]), setup, teardown)
#endif
