// Rotated bilinear window gather: F.affine_grid(theta, align_corners=True) + F.grid_sample(bilinear, zeros,
// align_corners=True) of UniPlanner.crop_feature (team_code_v2/models/uniplanner.py:303-340) as one kernel on the
// channels-last feature map.  Crop k reads frame frame_idx[k] directly (the reference's `features.expand(N,...)`
// view, without materialising it for batched frames).  One warp per output pixel, 16 B per lane per step.
#include "common.cuh"

namespace lavb {

template <typename T, int VEC>
__global__ void __launch_bounds__(256) crop_kernel(const T* __restrict__ feat, int B, int H, int W, int C,
                                                   const int* __restrict__ frame_idx, const float* __restrict__ theta,
                                                   int K, int S, T* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int total = K * S * S;
  if (warp >= total) return;
  const int k = warp / (S * S), r = warp - k * S * S, j = r / S, i = r - j * S;
  // torch.linspace(-1, 1, S): start + step*idx for the first half, end - step*(S-1-idx) for the second
  const float step = 2.f / (float)(S - 1);
  const float xb = (i < S / 2) ? (-1.f + step * (float)i) : (1.f - step * (float)(S - 1 - i));
  const float yb = (j < S / 2) ? (-1.f + step * (float)j) : (1.f - step * (float)(S - 1 - j));
  const float* th = theta + k * 6;
  const float gx = fmaf(th[0], xb, fmaf(th[1], yb, th[2]));
  const float gy = fmaf(th[3], xb, fmaf(th[4], yb, th[5]));
  const float ix = (gx + 1.f) * 0.5f * (float)(W - 1);
  const float iy = (gy + 1.f) * 0.5f * (float)(H - 1);
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy;
  const float ax = ix - fx, ay = iy - fy;
  const float w00 = (1.f - ax) * (1.f - ay), w01 = ax * (1.f - ay), w10 = (1.f - ax) * ay, w11 = ax * ay;
  const bool vx0 = x0 >= 0 && x0 < W, vx1 = x0 + 1 >= 0 && x0 + 1 < W, vy0 = y0 >= 0 && y0 < H, vy1 = y0 + 1 >= 0 && y0 + 1 < H;
  int b = __ldg(frame_idx + k);
  b = b < 0 ? 0 : (b >= B ? B - 1 : b);
  const T* base = feat + (long long)b * H * W * C;
  const T* p00 = base + ((long long)y0 * W + x0) * C;
  const T* p01 = p00 + C;
  const T* p10 = p00 + (long long)W * C;
  const T* p11 = p10 + C;
  T* o = out + (long long)warp * C;
  for (int c = lane * VEC; c < C; c += 32 * VEC) {
    float acc[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
    auto add = [&](const T* p, float wgt, bool ok) {
      if (!ok) return;
#pragma unroll
      for (int e = 0; e < VEC; e += 4) {
        const float4 v = load4<T>(p + c + e);
        acc[e] = fmaf(wgt, v.x, acc[e]); acc[e + 1] = fmaf(wgt, v.y, acc[e + 1]);
        acc[e + 2] = fmaf(wgt, v.z, acc[e + 2]); acc[e + 3] = fmaf(wgt, v.w, acc[e + 3]);
      }
    };
    add(p00, w00, vx0 && vy0); add(p01, w01, vx1 && vy0); add(p10, w10, vx0 && vy1); add(p11, w11, vx1 && vy1);
#pragma unroll
    for (int e = 0; e < VEC; e += 4) store4<T>(o + c + e, make_float4(acc[e], acc[e + 1], acc[e + 2], acc[e + 3]));
  }
}

}  // namespace lavb

using namespace lavb;

extern "C" int lavb_crop_bilinear(const void* d_feat, int dtype, int b, int h, int w, int c, const int* d_frame_idx,
                                  const float* d_theta, int k, int crop, void* d_out, void* stream) {
  LAVB_CHECK_ARG(c % 8 == 0, "crop_bilinear: channels must be a multiple of 8 (got %d)", c);
  LAVB_CHECK_ARG(crop >= 2 && b >= 1, "crop_bilinear: bad crop size / batch");
  if (k == 0) return 0;
  const long long warps = (long long)k * crop * crop;
  const int blocks = ceil_div(warps * 32, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == LAVB_F32)
    crop_kernel<float, 4><<<blocks, 256, 0, st>>>((const float*)d_feat, b, h, w, c, d_frame_idx, d_theta, k, crop, (float*)d_out);
  else if (dtype == LAVB_BF16)
    crop_kernel<__nv_bfloat16, 8><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)d_feat, b, h, w, c, d_frame_idx, d_theta, k, crop,
                                                           (__nv_bfloat16*)d_out);
  else LAVB_CHECK_ARG(false, "crop_bilinear: bad dtype");
  LAVB_LAUNCH_OK();
  return 0;
}
