#include <stdarg.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"

namespace mtgseg {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

static thread_local bool t_pdl_scope = true;
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("MTGSEG_PDL"); return !(e && e[0] == '0'); }();
  return on && t_pdl_scope;
}
PdlScope::PdlScope(bool allow) : prev_(t_pdl_scope) { t_pdl_scope = allow; }
PdlScope::~PdlScope() { t_pdl_scope = prev_; }

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
unsigned long long launch_count() { return g_launches.load(std::memory_order_relaxed); }
}  // namespace mtgseg
