// Library-wide state of the t2v_b200 C-ABI: last-error string, launch counter, version.
#include "t2v_common.cuh"
#include <stdarg.h>
#include <stdlib.h>

static thread_local char g_err[1024] = "";
unsigned long long g_t2v_launches = 0;

void t2v_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool t2v_pdl_enabled() {
  static const bool on = !(getenv("T2V_PDL") && getenv("T2V_PDL")[0] == '0');
  return on;
}

T2V_API const char* t2v_last_error(void) { return g_err; }
T2V_API int t2v_version(void) { return 100; }
T2V_API unsigned long long t2v_launch_count(void) { return g_t2v_launches; }
T2V_API void t2v_reset_launch_count(void) { g_t2v_launches = 0; }
T2V_API void t2v_add_launch_count(unsigned long long n) { g_t2v_launches += n; }   /* launches replayed from a CUDA graph */
