// fit_tile_kernel<..., MODE = 0, ...>: fit (loss + gradient)
#include "dnmf_fit.cuh"

namespace dnmf {
int launch_fit_mode0(int nwx, int nwy, int sub, bool fd, const FitParams& p, int B, size_t smem, cudaStream_t st) {
#if DNMF_AFFINE_BODIES
  // affine fits (FitParams::skip_quad): the instantiation whose specialised main loops drop the z^2 terms
  if (p.skip_quad && sub == 2 && fd && p.nwz == 1) {
    if (nwx == 1 && nwy == 1) return launch_fit<1, 1, 2, 0, true, true>(p, B, smem, st);
    if (nwx == 2 && nwy == 1) return launch_fit<2, 1, 2, 0, true, true>(p, B, smem, st);
    if (nwx == 2 && nwy == 2) return launch_fit<2, 2, 2, 0, true, true>(p, B, smem, st);
  }
#endif
  DNMF_FIT_DISPATCH(0);
}
}  // namespace dnmf
