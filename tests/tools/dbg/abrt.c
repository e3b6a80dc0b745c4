#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
static void h(int s){ void* bt[64]; int n = backtrace(bt,64); backtrace_symbols_fd(bt,n,2); _exit(134);}
__attribute__((constructor)) static void init(){ signal(SIGABRT,h); signal(SIGSEGV,h); }
