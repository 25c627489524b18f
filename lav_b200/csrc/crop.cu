// Rotated bilinear window gather: F.affine_grid(theta, align_corners=True) + F.grid_sample(bilinear, zeros,
// align_corners=True) of UniPlanner.crop_feature (team_code_v2/models/uniplanner.py:303-340) as one kernel on the
// channels-last feature map.  Crop k reads frame frame_idx[k] directly (the reference's `features.expand(N,...)`
// view, without materialising it for batched frames).
#include "common.cuh"

namespace lavb {

// 16 B of channels per thread (4 fp32 / 8 h16): C*sizeof(T)/16 threads per output pixel, four 16 B loads + one 16 B store
// each — one warp per pixel left a third of the lanes idle at C = 384 and issued twice as many (8 B) loads.
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 r = __ldg(reinterpret_cast<const float4*>(p)); v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct Vec16<h16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const h16* p, float (&v)[8]) {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t u[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = h1622float2(*reinterpret_cast<const h162*>(&u[e]));
      v[2 * e] = f.x; v[2 * e + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(h16* p, const float (&v)[8]) {
    uint32_t u[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const h162 h = floats2h162(v[2 * e], v[2 * e + 1]);
      u[e] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(u[0], u[1], u[2], u[3]);
  }
};

// block = one 8x8 patch of output pixels x one 64 B channel slice (4 threads of 16 B per pixel): neighbouring output
// pixels sample overlapping 2x2 input neighbourhoods, so a compact patch lets L1 serve the ~4x re-reads that a row-major
// pixel order sent to L2.  grid = (patches per crop, channel slices, crops).
template <typename T>
__global__ void __launch_bounds__(256) crop_kernel(const T* __restrict__ feat, int B, int H, int W, int C,
                                                   const int* __restrict__ frame_idx, const float* __restrict__ theta,
                                                   int K, int S, T* __restrict__ out) {
  constexpr int VEC = Vec16<T>::N;
  const int pw = (S + 7) >> 3;
  const int pj = blockIdx.x / pw, pi = blockIdx.x - pj * pw;
  const int k = blockIdx.z;
  const int j = pj * 8 + (threadIdx.x >> 5), i = pi * 8 + ((threadIdx.x >> 2) & 7);
  const int c = (blockIdx.y * 4 + (threadIdx.x & 3)) * VEC;
  if (i >= S || j >= S || c >= C) return;
  const long long pix = ((long long)k * S + j) * S + i;
  // torch.linspace(-1, 1, S): start + step*idx for the first half, end - step*(S-1-idx) for the second
  const float step = 2.f / (float)(S - 1);
  const float xb = (i < S / 2) ? (-1.f + step * (float)i) : (1.f - step * (float)(S - 1 - i));
  const float yb = (j < S / 2) ? (-1.f + step * (float)j) : (1.f - step * (float)(S - 1 - j));
  const float* th = theta + k * 6;
  const float gx = fmaf(__ldg(th + 0), xb, fmaf(__ldg(th + 1), yb, __ldg(th + 2)));
  const float gy = fmaf(__ldg(th + 3), xb, fmaf(__ldg(th + 4), yb, __ldg(th + 5)));
  const float ix = (gx + 1.f) * 0.5f * (float)(W - 1);
  const float iy = (gy + 1.f) * 0.5f * (float)(H - 1);
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy;
  const float ax = ix - fx, ay = iy - fy;
  const float w00 = (1.f - ax) * (1.f - ay), w01 = ax * (1.f - ay), w10 = (1.f - ax) * ay, w11 = ax * ay;
  const bool vx0 = x0 >= 0 && x0 < W, vx1 = x0 + 1 >= 0 && x0 + 1 < W, vy0 = y0 >= 0 && y0 < H, vy1 = y0 + 1 >= 0 && y0 + 1 < H;
  int b = __ldg(frame_idx + k);
  b = b < 0 ? 0 : (b >= B ? B - 1 : b);
  const T* p00 = feat + (long long)b * H * W * C + ((long long)y0 * W + x0) * C + c;
  float acc[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
  auto add = [&](const T* p, float wgt, bool ok) {
    if (!ok) return;
    float v[VEC];
    Vec16<T>::load(p, v);
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e] = fmaf(wgt, v[e], acc[e]);
  };
  add(p00, w00, vx0 && vy0); add(p00 + C, w01, vx1 && vy0);
  add(p00 + (long long)W * C, w10, vx0 && vy1); add(p00 + (long long)W * C + C, w11, vx1 && vy1);
  Vec16<T>::store(out + pix * C + c, acc);
}

}  // namespace lavb

using namespace lavb;

extern "C" int lavb_crop_bilinear(const void* d_feat, int dtype, int b, int h, int w, int c, const int* d_frame_idx,
                                  const float* d_theta, int k, int crop, void* d_out, void* stream) {
  LAVB_CHECK_ARG(c % 8 == 0, "crop_bilinear: channels must be a multiple of 8 (got %d)", c);
  LAVB_CHECK_ARG(crop >= 2 && b >= 1, "crop_bilinear: bad crop size / batch");
  if (k == 0) return 0;
  LAVB_CHECK_ARG(k <= 65535, "crop_bilinear: at most 65535 crops per call");
  const int pw = (crop + 7) / 8, vec = dtype == LAVB_F32 ? 4 : 8;
  const dim3 blocks(pw * pw, ceil_div(c, 4 * vec), k);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == LAVB_F32)
    crop_kernel<float><<<blocks, 256, 0, st>>>((const float*)d_feat, b, h, w, c, d_frame_idx, d_theta, k, crop, (float*)d_out);
  else if (dtype == LAVB_H16)
    crop_kernel<h16><<<blocks, 256, 0, st>>>((const h16*)d_feat, b, h, w, c, d_frame_idx, d_theta, k, crop,
                                                           (h16*)d_out);
  else LAVB_CHECK_ARG(false, "crop_bilinear: bad dtype");
  LAVB_LAUNCH_OK();
  return 0;
}
