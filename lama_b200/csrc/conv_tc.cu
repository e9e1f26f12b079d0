// tcgen05 arm of ffcb_conv() — placeholder until the TMA/tcgen05 kernel lands.
#include "common.cuh"
namespace ffcb {
int conv_tc(const ffcb_conv_desc*, cudaStream_t) {
  set_error("conv: FFCB_MATH_BF16X3 not available in this build");
  return FFCB_EINVAL;
}
}  // namespace ffcb
