// Launch helper shared by the persistent (grid-synchronising) kernels: cooperative by default, plain on request.
// See kernels.h (coop_launch_enabled) for why the two-deep pipeline asks for plain launches.
#include <cstdlib>

#include "kernels.h"

namespace rnnt {
namespace {
thread_local bool tl_coop = true;
}

bool coop_launch_enabled() {
  // RNNT_NO_COOP=1: profilers cannot replay a cooperative launch (only use on an otherwise idle GPU)
  static const bool env_off = [] { const char* e = getenv("RNNT_NO_COOP"); return e && e[0] == '1'; }();
  return tl_coop && !env_off;
}
void set_coop_launch(bool on) { tl_coop = on; }

cudaError_t launch_persistent(const void* fn, dim3 grid, dim3 block, void** kargs, size_t smem, cudaStream_t st) {
  if (coop_launch_enabled()) return cudaLaunchCooperativeKernel(fn, grid, block, kargs, smem, st);
  return cudaLaunchKernel(fn, grid, block, kargs, smem, st);
}

}  // namespace rnnt
