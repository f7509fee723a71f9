// fit_tile_kernel<..., MODE = 3, ...>: trace statistics G_t, b_t on the fused kernel's tiles (two-sub-tile
// layouts with the verified fast division only)
#include "dnmf_fit.cuh"

namespace dnmf {
int launch_fit_mode3(int nwx, int nwy, int sub, bool fd, const FitParams& p, int B, size_t smem, cudaStream_t st) {
  if (sub != 2 || !fd) return fail("dispatch_stats: layout without a fused statistics kernel");
  if (nwx == 1 && nwy == 1) return launch_fit<1, 1, 2, 3, true>(p, B, smem, st);
  if (nwx == 2 && nwy == 1) return launch_fit<2, 1, 2, 3, true>(p, B, smem, st);
  if (nwx == 2 && nwy == 2) return launch_fit<2, 2, 2, 3, true>(p, B, smem, st);
  return fail("dispatch_stats: layout without a fused statistics kernel");
}
}  // namespace dnmf
