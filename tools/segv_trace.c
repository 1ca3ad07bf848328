/* debugging aid: print a native backtrace on SIGSEGV/SIGABRT (load with ctypes, call segv_trace_install()) */
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
static void handler(int sig)
{
    void *frames[64];
    int n = backtrace(frames, 64);
    backtrace_symbols_fd(frames, n, 2);
    signal(sig, SIG_DFL);
    raise(sig);
}
void segv_trace_install(void) { signal(SIGSEGV, handler); signal(SIGABRT, handler); signal(SIGBUS, handler); }
