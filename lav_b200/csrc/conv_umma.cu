// tcgen05 implicit-GEMM 3x3 convolution — placeholder until the kernel lands (returns an error, never computes).
#include "common.cuh"
extern "C" int lavb_conv3x3_umma(const void*, int, int, int, int, const void*, int, const float*, const float*, int, void*, int,
                                 int, void*) {
  lavb::set_error("conv3x3_umma: not built yet");
  return 4;
}
