// EXPERIMENTAL (round-2 work item, not on the default path; enabled by lav_b200.layers.USE_EPI16):
// conv_umma.cu compiled a second time with 8 epilogue warps AND two CTAs per SM for the narrow layers (cout <= 128).
// Why: those layers are bound by the epilogue — 8 epilogue warps per SM (2 CTAs x 4) leave each SM sub-partition two warps
// to hide tcgen05.ld / LDS / store latency (IPC ~0.4 of 4); 2 CTAs x 8 warps doubles that.  Under
// __launch_bounds__(320, 2) ptxas fits the 8-warp epilogue variants in <= 96 registers without spills (checked), so
// 2 x 320 threads x 96 registers fit the register file; shared memory and TMEM budgets are those of the existing 2-CTA mode.
// Everything else (kernel body, host logic, descriptor) is conv_umma.cu itself.
#define LAVB_UMMA_EW8_MINBLOCKS 2
#define LAVB_UMMA_NARROW_EW 8
#define conv_umma_kernel conv_umma16_kernel
#define lavb_conv_umma lavb_conv_umma16_impl
#include "conv_umma.cu"
#undef lavb_conv_umma

// Declines (returns 4) for wide layers and the depth-to-space epilogue: callers fall back to lavb_conv_umma.
extern "C" int lavb_conv_umma16(const lavb_conv_desc* d, void* stream) {
  LAVB_CHECK_ARG(d != nullptr, "conv_umma16: null descriptor");
  if ((d->cout + 31) / 32 * 32 > 128 || d->d2s_nout) {
    lavb::set_error("conv_umma16: only layers with cout <= 128 and no depth-to-space epilogue");
    return 4;
  }
  return lavb_conv_umma16_impl(d, stream);
}
