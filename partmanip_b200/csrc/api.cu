// Library-level entry points of libpartmanip_b200.so.
#include "common.cuh"

char g_pm_err[512] = {0};

extern "C" {
const char* pm_last_error(void) { return g_pm_err; }
int pm_version(void) { return 100; }
}
