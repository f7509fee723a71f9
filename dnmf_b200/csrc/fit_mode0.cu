// fit_tile_kernel<..., MODE = 0, ...>: fit (loss + gradient)
#include "dnmf_fit.cuh"

namespace dnmf {
int launch_fit_mode0(int nwx, int nwy, int sub, bool fd, const FitParams& p, int B, size_t smem, cudaStream_t st) {
  DNMF_FIT_DISPATCH(0);
}
}  // namespace dnmf
