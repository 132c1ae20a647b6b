// The CUDA-core cross-check library (scripts/libgcc_b200_check.so) is built apart from the product library; these
// are the two runtime symbols common.cuh's launch macro expects.
#include <stdio.h>
unsigned long long g_gcc_launches = 0;
static thread_local char g_err[512] = "";
void gcc_set_error(const char* file, int line, const char* msg) { snprintf(g_err, sizeof(g_err), "%s:%d: %s", file, line, msg); }
extern "C" const char* gcc_check_last_error(void) { return g_err; }
