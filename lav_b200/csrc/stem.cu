// 7x7 stride-2 stem convolution of the brake predictor's ResNet-18 on raw uint8 camera frames
// (lav/models/resnet.py:178,235-238 conv1+bn1+relu; team_code_v2/models/rgb.py:66-70 normalisation), as an implicit GEMM on
// the tensor cores with M = output pixels, N = 64, K = 7*7*3 = 147 (padded to 160): cuDNN pads the 3 input channels of a
// channels-last bf16 tensor and spends ~40 us per frame here; this kernel gathers the window bytes itself.
//   A fragment element (pixel, k) = bf16((u8 - 255 mean_c) / (255 std_c)) for k = (ky*7 + kx)*3 + c, 0 outside the image
//   (zero padding acts on the NORMALISED image, as in the reference), built in registers — no im2col buffer;
//   B = BatchNorm-folded weights [64][160] bf16 staged once per block in shared memory; fp32 accumulate (mma.sync
//   m16n8k16); epilogue bias + ReLU -> bf16 NHWC.  The "wide" image of the brake model is three cameras side by side
//   (lav_agent_fast.py:257): `ncam`/`cam_w` index the (B, ncam, H, cam_w, 3) camera tensor directly.
#include "common.cuh"

namespace lavb {

constexpr int kStemK = 160, kStemPitch = 168;   // bf16 elements per weight row in smem (pitch chosen bank-conflict free)

struct StemArgs {
  const unsigned char* img; int batch, ncam, h, cam_w;     // logical image: h x (ncam*cam_w) x 3
  const __nv_bfloat16* w; const float* bias;               // w [64][160]
  float na[3], nb[3];                                       // normalised = u8 * na[c] + nb[c]
  __nv_bfloat16* out; int ho, wo;                           // NHWC (batch, ho, wo, 64)
};

__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(128) stem7x7_u8_kernel(const __grid_constant__ StemArgs a) {
  __shared__ __align__(16) __nv_bfloat16 ws[64 * kStemPitch];
  for (int i = threadIdx.x; i < 64 * kStemK; i += blockDim.x) ws[(i / kStemK) * kStemPitch + i % kStemK] = a.w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, gq = lane >> 2, tq = lane & 3;
  // grid: x = (image, output row), y = 128-pixel segment of the row; warp = 32 consecutive output pixels (2 m-tiles)
  const int b = blockIdx.x / a.ho, oy = blockIdx.x - b * a.ho;
  const int ox0 = blockIdx.y * 128 + (threadIdx.x >> 5) * 32;
  if (ox0 >= a.wo) return;
  const int W = a.ncam * a.cam_w;
  int ox[4]; bool pv[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) { ox[r] = ox0 + (r >> 1) * 16 + (r & 1) * 8 + gq; pv[r] = ox[r] < a.wo; }
  float acc[2][8][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nn = 0; nn < 8; ++nn) acc[mt][nn][0] = acc[mt][nn][1] = acc[mt][nn][2] = acc[mt][nn][3] = 0.f;
  const long long img_b = (long long)b * a.ncam * a.h * a.cam_w * 3;
#pragma unroll 1
  for (int kk = 0; kk < kStemK / 16; ++kk) {
    // the four k indices of this lane in this K16 step: 16kk + 2tq + {0, 1, 8, 9}
    int dyk[4], dxk[4], ck[4]; bool kv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = kk * 16 + 2 * tq + (j & 1) + (j >> 1) * 8;
      kv[j] = k < 147;
      const int tap = k / 3;
      ck[j] = k - tap * 3;
      dyk[j] = tap / 7;
      dxk[j] = tap - dyk[j] * 7;
    }
    uint32_t af[4][2];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int iy = 2 * oy - 3 + dyk[j], ix = 2 * ox[r] - 3 + dxk[j];
        float x = 0.f;
        if (kv[j] && pv[r] && iy >= 0 && iy < a.h && ix >= 0 && ix < W) {
          const int cam = ix / a.cam_w, xc = ix - cam * a.cam_w;
          const unsigned char u = __ldg(a.img + img_b + (((long long)cam * a.h + iy) * a.cam_w + xc) * 3 + ck[j]);
          x = fmaf((float)u, a.na[ck[j]], a.nb[ck[j]]);
        }
        v[j] = x;
      }
      const __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
      af[r][0] = *reinterpret_cast<const uint32_t*>(&lo);      // k = 2tq, 2tq+1
      af[r][1] = *reinterpret_cast<const uint32_t*>(&hi);      // k = 2tq+8, 2tq+9
    }
#pragma unroll
    for (int nn = 0; nn < 8; ++nn) {
      const __nv_bfloat16* wp = ws + (nn * 8 + gq) * kStemPitch + kk * 16 + 2 * tq;
      const uint32_t b0 = *reinterpret_cast<const uint32_t*>(wp), b1 = *reinterpret_cast<const uint32_t*>(wp + 8);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) mma_bf16(acc[mt][nn], af[mt * 2][0], af[mt * 2 + 1][0], af[mt * 2][1], af[mt * 2 + 1][1], b0, b1);
    }
  }
  __nv_bfloat16* orow = a.out + ((long long)b * a.ho + oy) * a.wo * 64;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    if (!pv[r]) continue;
    __nv_bfloat16* op = orow + (long long)ox[r] * 64;
#pragma unroll
    for (int nn = 0; nn < 8; ++nn) {
      const int c = nn * 8 + 2 * tq;
      const float x0 = fmaxf(acc[r >> 1][nn][(r & 1) * 2] + __ldg(a.bias + c), 0.f);
      const float x1 = fmaxf(acc[r >> 1][nn][(r & 1) * 2 + 1] + __ldg(a.bias + c + 1), 0.f);
      store2<__nv_bfloat16>(op + c, x0, x1);
    }
  }
}

}  // namespace lavb

using namespace lavb;

extern "C" int lavb_stem7x7s2_u8(const void* d_img, int batch, int ncam, int h, int cam_w, const void* d_w, const float* d_bias,
                                 const float* h_mean, const float* h_std, void* d_out, void* stream) {
  LAVB_CHECK_ARG(batch >= 0 && ncam >= 1 && h >= 7 && cam_w >= 7, "stem7x7s2_u8: bad shape");
  if (batch == 0) return 0;
  StemArgs a;
  a.img = reinterpret_cast<const unsigned char*>(d_img); a.batch = batch; a.ncam = ncam; a.h = h; a.cam_w = cam_w;
  a.w = reinterpret_cast<const __nv_bfloat16*>(d_w); a.bias = d_bias;
  for (int c = 0; c < 3; ++c) { a.na[c] = 1.f / (255.f * h_std[c]); a.nb[c] = -h_mean[c] / h_std[c]; }
  a.out = reinterpret_cast<__nv_bfloat16*>(d_out);
  a.ho = (h + 6 - 7) / 2 + 1; a.wo = (ncam * cam_w + 6 - 7) / 2 + 1;
  dim3 grid(batch * a.ho, ceil_div(a.wo, 128));
  stem7x7_u8_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(a);
  LAVB_LAUNCH_OK();
  return 0;
}
