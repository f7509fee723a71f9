// fit_tile_kernel<..., MODE = 1, ...>: forward only (writes Yhat)
#include "dnmf_fit.cuh"

namespace dnmf {
int launch_fit_mode1(int nwx, int nwy, int sub, bool fd, const FitParams& p, int B, size_t smem, cudaStream_t st) {
  DNMF_FIT_DISPATCH(1);
}
}  // namespace dnmf
