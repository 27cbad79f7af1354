/* LD_PRELOAD helper: print a backtrace on SIGSEGV / SIGABRT (debugging aid for the C++ client programs on the
 * GPU box, where no debugger is installed).  gcc -shared -fPIC -o segv_backtrace.so segv_backtrace.c */
#include <execinfo.h>
#include <signal.h>
#include <stdio.h>
#include <unistd.h>
static void handler(int sig) {
  void *frames[64];
  int n = backtrace(frames, 64);
  fprintf(stderr, "signal %d, backtrace:\n", sig);
  backtrace_symbols_fd(frames, n, 2);
  _exit(128 + sig);
}
__attribute__((constructor)) static void install(void) {
  signal(SIGSEGV, handler);
  signal(SIGABRT, handler);
  signal(SIGBUS, handler);
}
