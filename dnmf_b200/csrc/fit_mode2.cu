// fit_tile_kernel<..., MODE = 2, ...>: fit with scalar background, residual written (extension)
#include "dnmf_fit.cuh"

namespace dnmf {
int launch_fit_mode2(int nwx, int nwy, int sub, bool fd, const FitParams& p, int B, size_t smem, cudaStream_t st) {
  DNMF_FIT_DISPATCH(2);
}
}  // namespace dnmf
