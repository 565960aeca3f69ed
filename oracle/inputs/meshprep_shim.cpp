// TEST / BENCH INFRASTRUCTURE. The error plumbing the product's host-only mesh sources expect from the CUDA library
// (csrc/sb_api.cu: sb::set_error, sb_last_error), so that csrc/sb_mesh_host.cpp and csrc/sb_part_host.cpp can be built
// into oracle/libsb_meshprep.so without any device code: bench.py's reference arm (`--impl reference`) prepares its
// inputs -- the synthetic mesh of BASELINE.json configs[1], shuffled and RCM-renumbered exactly like the product arm's --
// through this library, so the process that times the reference's CPU solver never maps libstormb200.so.
#include <cstdarg>
#include <cstdio>

#include "../../include/stormb200.h"

namespace sb {
static thread_local char g_error[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  std::vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
} // namespace sb

extern "C" SB_API const char* sb_last_error(void) { return sb::g_error; }
