// fit_tile_kernel<..., MODE = 0, ...>: fit (loss + gradient)
#include "dnmf_fit.cuh"

namespace dnmf {
#if DNMF_SKEW_KERNELS
// the two-sub-tile layouts with the verified fast division: affine or general main loops, plain or rotated z order
template <int NWX, int NWY>
static int launch_fast(const FitParams& p, int B, size_t smem, cudaStream_t st) {
  const bool aff = DNMF_AFFINE_BODIES && p.skip_quad, skew = p.z_skew != 0;
  if (aff) return skew ? launch_fit<NWX, NWY, 2, 0, true, true, 1, true>(p, B, smem, st)
                       : launch_fit<NWX, NWY, 2, 0, true, true, 1, false>(p, B, smem, st);
  return skew ? launch_fit<NWX, NWY, 2, 0, true, false, 1, true>(p, B, smem, st)
              : launch_fit<NWX, NWY, 2, 0, true, false, 1, false>(p, B, smem, st);
}
#endif

int launch_fit_mode0(int nwx, int nwy, int sub, bool fd, const FitParams& p, int B, size_t smem, cudaStream_t st) {
#if DNMF_SKEW_KERNELS
  if (sub == 2 && fd && p.nwz == 1) {
    if (nwx == 1 && nwy == 1) return launch_fast<1, 1>(p, B, smem, st);
    if (nwx == 2 && nwy == 1) return launch_fast<2, 1>(p, B, smem, st);
    if (nwx == 2 && nwy == 2) return launch_fast<2, 2>(p, B, smem, st);
  }
#elif DNMF_AFFINE_BODIES
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
