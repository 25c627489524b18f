// Point painting, sweep stacking and small ingest kernels (HBM-bound, one thread per point / element).
//
// paint: restates CoordConverter.forward + InferModel.point_painting
// (team_code_v2/model_inference.py:75-93,280-297) as ONE kernel: 3 camera projections in the
// reference's fp32 operation order (k-sequential FMA chains, IEEE division, truncation toward
// zero, bounds test on the truncated integers, later camera wins), then one gather.
#include "common.cuh"

namespace lavb {

struct CamSet {
  float m[4][41];  // K(9) | lidar_to_world(16) | world_to_cam(16)
  int ncam;
};

// row-vector dot in the order a BLAS sgemm with a k-loop produces: ((a0*b0 + a1*b1) + a2*b2) + a3*b3 with FMA
__device__ __forceinline__ float dot4(const float* r, float x, float y, float z, float w) {
  float acc = __fmul_rn(r[0], x);
  acc = __fmaf_rn(r[1], y, acc);
  acc = __fmaf_rn(r[2], z, acc);
  acc = __fmaf_rn(r[3], w, acc);
  return acc;
}
__device__ __forceinline__ float dot3(const float* r, float x, float y, float z) {
  float acc = __fmul_rn(r[0], x);
  acc = __fmaf_rn(r[1], y, acc);
  acc = __fmaf_rn(r[2], z, acc);
  return acc;
}

// float -> int64 the way x86 cvttss2si does for the cases that matter: NaN / inf / |v| >= 2^63 give the
// "indefinite" INT64_MIN (which then fails every >= 0 test in the reference).
__device__ __forceinline__ long long trunc_i64(float v) {
  if (!(fabsf(v) < 9.2233720368547758e18f)) return (long long)0x8000000000000000ull;
  return (long long)v;  // cvt.rzi
}

template <int MODE>
__global__ void __launch_bounds__(256) paint_kernel(const float* __restrict__ pts, int n, int pt_stride,
                                                    const float* __restrict__ sem, int c_in, int H, int W,
                                                    long long s_cam, long long s_c, long long s_y, long long s_x,
                                                    const __grid_constant__ CamSet cams, float* __restrict__ out,
                                                    int out_stride, int out_col0, int copy_cols,
                                                    long long pts_frame_stride, long long sem_frame_stride,
                                                    long long out_frame_stride) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // blockIdx.y = frame of a batch of independent agents (each with its own sweep and semantic maps)
  pts += blockIdx.y * pts_frame_stride;
  sem += blockIdx.y * sem_frame_stride;
  out += blockIdx.y * out_frame_stride;
  const float* p = pts + (size_t)i * pt_stride;
  float x, y, z;
  float4 p4;
  const bool vec = (pt_stride == 4);
  if (vec) {
    p4 = __ldg(reinterpret_cast<const float4*>(p));
    x = p4.x; y = p4.y; z = p4.z;
  } else {
    x = __ldg(p); y = __ldg(p + 1); z = __ldg(p + 2);
  }
  int hit_cam = -1, hit_u = 0, hit_v = 0;
#pragma unroll 1
  for (int c = 0; c < cams.ncam; ++c) {
    const float* K = cams.m[c];
    const float* L = cams.m[c] + 9;
    const float* Wc = cams.m[c] + 25;
    // world = lidar_to_world @ [x,y,z,1]
    const float w0 = dot4(L + 0, x, y, z, 1.f), w1 = dot4(L + 4, x, y, z, 1.f), w2 = dot4(L + 8, x, y, z, 1.f),
                w3 = dot4(L + 12, x, y, z, 1.f);
    // cam = world_to_cam @ world ; re-axis (cam_y, -cam_z, cam_x)
    const float c0 = dot4(Wc + 0, w0, w1, w2, w3), c1 = dot4(Wc + 4, w0, w1, w2, w3), c2 = dot4(Wc + 8, w0, w1, w2, w3);
    const float a0 = c1, a1 = -c2, a2 = c0;
    // cam_2d = K @ cam
    const float q0 = dot3(K + 0, a0, a1, a2), q1 = dot3(K + 3, a0, a1, a2), q2 = dot3(K + 6, a0, a1, a2);
    const float den = __fadd_rn(1e-5f, q2);
    const long long u = trunc_i64(__fdiv_rn(q0, den));
    const long long v = trunc_i64(__fdiv_rn(q1, den));
    const long long zi = trunc_i64(q2);
    if (zi >= 0 && u >= 0 && u < W && v >= 0 && v < H) {
      hit_cam = c; hit_u = (int)u; hit_v = (int)v;
    }
  }
  float* o = out + (size_t)i * out_stride;
  if (copy_cols == 4 && vec && out_col0 >= 4 && (out_stride % 4) == 0) {
    *reinterpret_cast<float4*>(o) = p4;
  } else {
    for (int k = 0; k < copy_cols; ++k) o[k] = __ldg(p + k);
  }
  const int c_out = (MODE == 0) ? c_in : c_in - 1;
  o += out_col0;
  if (hit_cam < 0) {
    for (int k = 0; k < c_out; ++k) o[k] = 0.f;
    return;
  }
  const float* s = sem + hit_cam * s_cam + hit_v * s_y + hit_u * s_x;
  if (MODE == 0) {
    for (int k = 0; k < c_out; ++k) o[k] = __ldg(s + k * s_c);
  } else {
    float pr[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) pr[k] = (k < c_in) ? __ldg(s + k * s_c) : 0.f;
    if (MODE == 2) {  // softmax over c_in logits (torch.softmax: exp(x-max)/sum)
      float mx = pr[0];
#pragma unroll
      for (int k = 1; k < 8; ++k) if (k < c_in) mx = fmaxf(mx, pr[k]);
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) if (k < c_in) { pr[k] = expf(pr[k] - mx); sum += pr[k]; }
#pragma unroll
      for (int k = 0; k < 8; ++k) if (k < c_in) pr[k] = __fdiv_rn(pr[k], sum);
    }
    const float bg = __fsub_rn(1.f, pr[0]);  // pred_sem[:,1:] * (1 - pred_sem[:,:1]), model_inference.py:45
#pragma unroll
    for (int k = 1; k < 8; ++k) if (k < c_in) o[k - 1] = __fmul_rn(pr[k], bg);
  }
}

struct StackParams { float R[9]; float dx, dy; };

__global__ void __launch_bounds__(256) stack_kernel(const float* __restrict__ src, int n, int src_cols,
                                                    const __grid_constant__ StackParams P, int time_idx, int n_time,
                                                    int roof_filter, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* s = src + (size_t)i * src_cols;
  const int dcols = src_cols + n_time;
  float* d = dst + (size_t)i * dcols;
  const float x = __ldg(s), y = __ldg(s + 1), z = __ldg(s + 2);
  // lidar @ R  (row vector times matrix, k-sequential), then += dloc: lav_agent_fast.py:555-563
  float nx = __fmaf_rn(z, P.R[6], __fmaf_rn(y, P.R[3], __fmul_rn(x, P.R[0])));
  float ny = __fmaf_rn(z, P.R[7], __fmaf_rn(y, P.R[4], __fmul_rn(x, P.R[1])));
  float nz = __fmaf_rn(z, P.R[8], __fmaf_rn(y, P.R[5], __fmul_rn(x, P.R[2])));
  nx = __fadd_rn(nx, P.dx);
  ny = __fadd_rn(ny, P.dy);
  if (roof_filter) {  // LAVAgent.preprocess, lav_agent.py:450 — on the sensor-frame coordinates
    if (x > -2.4f && x < 0.f && y > -0.8f && y < 0.8f && z > -1.5f && z < -1.f) nx = __int_as_float(0x7fc00000);
  }
  d[0] = nx; d[1] = ny; d[2] = nz;
  for (int k = 3; k < src_cols; ++k) d[k] = __ldg(s + k);
  for (int k = 0; k < n_time; ++k) d[src_cols + k] = (k == time_idx) ? 1.f : 0.f;
}

template <typename TOut>
__global__ void __launch_bounds__(256) rgb_norm_kernel(const void* __restrict__ rgb, int src_u8_nhwc, int n, int h, int w,
                                                       TOut* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // pixel index
  const long long npix = (long long)n * h * w;
  if (i >= npix) return;
  float r, g, b;
  if (src_u8_nhwc) {
    const unsigned char* p = reinterpret_cast<const unsigned char*>(rgb) + i * 3;
    r = p[0]; g = p[1]; b = p[2];
  } else {
    const long long hw = (long long)h * w;
    const long long img = i / hw, pix = i - img * hw;
    const float* p = reinterpret_cast<const float*>(rgb) + img * 3 * hw + pix;
    r = __ldg(p); g = __ldg(p + hw); b = __ldg(p + 2 * hw);
  }
  // (x/255. - .5)*2 , rgb.py:41 — same three fp32 roundings
  float4 v;
  v.x = __fmul_rn(__fsub_rn(__fdiv_rn(r, 255.f), .5f), 2.f);
  v.y = __fmul_rn(__fsub_rn(__fdiv_rn(g, 255.f), .5f), 2.f);
  v.z = __fmul_rn(__fsub_rn(__fdiv_rn(b, 255.f), .5f), 2.f);
  v.w = 0.f;
  store4<TOut>(out + i * 4, v);
}

template <typename TS, typename TD>
__global__ void __launch_bounds__(256) convert_kernel(const TS* __restrict__ s, TD* __restrict__ d, long long count) {
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < count) {
    store4<TD>(d + i, load4<TS>(s + i));
  } else {
    for (; i < count; ++i) d[i] = from_f32<TD>(to_f32<TS>(s[i]));
  }
}

// fp32 -> (hi, lo) h16 pair per element, laid out [row][hi(C) | lo(C)]: hi = h16(x), lo = h16(x - hi).  Feeding both
// halves to a tensor-core conv whose weights are duplicated along cin recovers ~16 mantissa bits of the input.
__global__ void __launch_bounds__(256) split_h16_kernel(const float* __restrict__ src, h16* __restrict__ dst,
                                                         long long rows, int c) {
  const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= rows * c) return;
  const long long r = i4 / c;
  const int ch = (int)(i4 - r * c);
  const float4 v = __ldg(reinterpret_cast<const float4*>(src + i4));
  const float x[4] = {v.x, v.y, v.z, v.w};
  float hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    hi[e] = h162float(float2h16(x[e]));
    lo[e] = x[e] - hi[e];
  }
  h16* d = dst + r * 2 * c + ch;
  store4<h16>(d, make_float4(hi[0], hi[1], hi[2], hi[3]));
  store4<h16>(d + c, make_float4(lo[0], lo[1], lo[2], lo[3]));
}

}  // namespace lavb

using namespace lavb;

extern "C" int lavb_split_h16(const float* d_src, void* d_dst, long long rows, int c, void* stream) {
  LAVB_CHECK_ARG(c % 4 == 0 && c > 0, "split_h16: channels must be a multiple of 4");
  if (rows == 0) return 0;
  split_h16_kernel<<<ceil_div(rows * c / 4, 256), 256, 0, (cudaStream_t)stream>>>(d_src, (h16*)d_dst, rows, c);
  LAVB_LAUNCH_OK();
  return 0;
}

static int paint_impl(const float* d_pts, int n, int pt_stride, const float* d_sem, int ncam, int c_in, int h, int w,
                      long long s_cam, long long s_c, long long s_y, long long s_x, const float* h_cams, int mode,
                      float* d_out, int out_stride, int out_col0, int copy_cols, int frames, long long pts_fs, long long sem_fs,
                      long long out_fs, void* stream) {
  LAVB_CHECK_ARG(n >= 0 && pt_stride >= 3, "paint: bad n/pt_stride");
  LAVB_CHECK_ARG(ncam >= 1 && ncam <= 4, "paint: ncam must be 1..4 (got %d)", ncam);
  LAVB_CHECK_ARG(mode >= 0 && mode <= 2, "paint: mode must be 0..2");
  LAVB_CHECK_ARG(c_in >= (mode ? 2 : 1) && c_in <= 8, "paint: c_in out of range (got %d)", c_in);
  const int c_out = mode ? c_in - 1 : c_in;
  LAVB_CHECK_ARG(copy_cols >= 0 && copy_cols <= pt_stride && out_col0 >= copy_cols && out_col0 + c_out <= out_stride,
                 "paint: output row layout inconsistent");
  if (n == 0 || frames == 0) return 0;
  CamSet cs;
  memcpy(cs.m, h_cams, sizeof(float) * 41 * ncam);
  cs.ncam = ncam;
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 blocks(ceil_div(n, 256), frames);
#define LAUNCH(M) paint_kernel<M><<<blocks, 256, 0, st>>>(d_pts, n, pt_stride, d_sem, c_in, h, w, s_cam, s_c, s_y, s_x, cs, \
                                                           d_out, out_stride, out_col0, copy_cols, pts_fs, sem_fs, out_fs)
  if (mode == 0) LAUNCH(0); else if (mode == 1) LAUNCH(1); else LAUNCH(2);
#undef LAUNCH
  LAVB_LAUNCH_OK();
  return 0;
}

extern "C" int lavb_paint(const float* d_pts, int n, int pt_stride, const float* d_sem, int ncam, int c_in, int h, int w,
                          long long s_cam, long long s_c, long long s_y, long long s_x, const float* h_cams, int mode,
                          float* d_out, int out_stride, int out_col0, int copy_cols, void* stream) {
  return paint_impl(d_pts, n, pt_stride, d_sem, ncam, c_in, h, w, s_cam, s_c, s_y, s_x, h_cams, mode, d_out, out_stride, out_col0,
                    copy_cols, 1, 0, 0, 0, stream);
}

extern "C" int lavb_paint_batched(const float* d_pts, int frames, int n, int pt_stride, long long pts_frame_stride,
                                  const float* d_sem, int ncam, int c_in, int h, int w, long long s_frame, long long s_cam,
                                  long long s_c, long long s_y, long long s_x, const float* h_cams, int mode, float* d_out,
                                  int out_stride, long long out_frame_stride, int out_col0, int copy_cols, void* stream) {
  LAVB_CHECK_ARG(frames >= 0 && frames <= 65535, "paint_batched: frames out of range");
  return paint_impl(d_pts, n, pt_stride, d_sem, ncam, c_in, h, w, s_cam, s_c, s_y, s_x, h_cams, mode, d_out, out_stride, out_col0,
                    copy_cols, frames, pts_frame_stride, s_frame, out_frame_stride, stream);
}

// ---- painting straight from the ERFNet decoder's last feature map ------------------------------------------------------------
// The frame path consumes the segmentation logits ONLY through the point-painting gather (<= 120 k samples of 221 k pixels per
// frame).  So the last ERFNet layer — output_conv = ConvTranspose2d(16, C, 2, stride 2) (lav/models/erfnet.py:122-124,132) — is
// evaluated inside the gather, for the hit pixel only: logits[k](v,u) = bias[k] + sum_c feat[v/2, u/2, c] * W[c][k][v%2][u%2],
// followed by softmax and the background suppression of model_inference.py:45.  The (H x W x C) fp32 logit maps (141 MB per 32
// frames) are never written, and the four launches of the transposed conv disappear.
struct DeconvW { float w[2][2][16][8]; float bias[8]; };      // [v%2][u%2][c_in][k]  (k < c_cls <= 8)

template <typename TF>
__global__ void __launch_bounds__(256) paint_deconv_kernel(const float* __restrict__ pts, int n, int pt_stride,
                                                           const TF* __restrict__ feat, int c_cls, int H, int W,
                                                           const __grid_constant__ CamSet cams, const DeconvW* __restrict__ dw,
                                                           float* __restrict__ out, int out_stride, int out_col0, int copy_cols,
                                                           long long pts_frame_stride, long long out_frame_stride) {
  __shared__ DeconvW sw;
  for (int i = threadIdx.x; i < (int)(sizeof(DeconvW) / 4); i += blockDim.x) reinterpret_cast<float*>(&sw)[i] = __ldg(reinterpret_cast<const float*>(dw) + i);
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  pts += blockIdx.y * pts_frame_stride;
  out += blockIdx.y * out_frame_stride;
  const float* p = pts + (size_t)i * pt_stride;
  float x, y, z;
  float4 p4;
  const bool vec = (pt_stride == 4);
  if (vec) { p4 = __ldg(reinterpret_cast<const float4*>(p)); x = p4.x; y = p4.y; z = p4.z; }
  else { x = __ldg(p); y = __ldg(p + 1); z = __ldg(p + 2); }
  int hit_cam = -1, hit_u = 0, hit_v = 0;
#pragma unroll 1
  for (int c = 0; c < cams.ncam; ++c) {       // identical projection arithmetic to paint_kernel
    const float* K = cams.m[c];
    const float* L = cams.m[c] + 9;
    const float* Wc = cams.m[c] + 25;
    const float w0 = dot4(L + 0, x, y, z, 1.f), w1 = dot4(L + 4, x, y, z, 1.f), w2 = dot4(L + 8, x, y, z, 1.f),
                w3 = dot4(L + 12, x, y, z, 1.f);
    const float c0 = dot4(Wc + 0, w0, w1, w2, w3), c1 = dot4(Wc + 4, w0, w1, w2, w3), c2 = dot4(Wc + 8, w0, w1, w2, w3);
    const float a0 = c1, a1 = -c2, a2 = c0;
    const float q0 = dot3(K + 0, a0, a1, a2), q1 = dot3(K + 3, a0, a1, a2), q2 = dot3(K + 6, a0, a1, a2);
    const float den = __fadd_rn(1e-5f, q2);
    const long long u = trunc_i64(__fdiv_rn(q0, den));
    const long long v = trunc_i64(__fdiv_rn(q1, den));
    const long long zi = trunc_i64(q2);
    if (zi >= 0 && u >= 0 && u < W && v >= 0 && v < H) { hit_cam = c; hit_u = (int)u; hit_v = (int)v; }
  }
  float* o = out + (size_t)i * out_stride;
  if (copy_cols == 4 && vec && out_col0 >= 4 && (out_stride % 4) == 0) *reinterpret_cast<float4*>(o) = p4;
  else for (int k = 0; k < copy_cols; ++k) o[k] = __ldg(p + k);
  o += out_col0;
  if (hit_cam < 0) {
    for (int k = 0; k < c_cls - 1; ++k) o[k] = 0.f;
    return;
  }
  const int hh = H >> 1, wh = W >> 1;
  const TF* f = feat + ((((long long)blockIdx.y * cams.ncam + hit_cam) * hh + (hit_v >> 1)) * wh + (hit_u >> 1)) * 16;
  float fv[16];
#pragma unroll
  for (int q = 0; q < 4; ++q) { const float4 t = load4<TF>(f + 4 * q); fv[4 * q] = t.x; fv[4 * q + 1] = t.y; fv[4 * q + 2] = t.z; fv[4 * q + 3] = t.w; }
  const float (*wk)[8] = sw.w[hit_v & 1][hit_u & 1];
  float pr[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float acc = sw.bias[k];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc = fmaf(fv[c], wk[c][k], acc);
    pr[k] = acc;
  }
  float mx = pr[0];
#pragma unroll
  for (int k = 1; k < 8; ++k) if (k < c_cls) mx = fmaxf(mx, pr[k]);
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) if (k < c_cls) { pr[k] = expf(pr[k] - mx); sum += pr[k]; }
#pragma unroll
  for (int k = 0; k < 8; ++k) if (k < c_cls) pr[k] = __fdiv_rn(pr[k], sum);
  const float bg = __fsub_rn(1.f, pr[0]);
#pragma unroll
  for (int k = 1; k < 8; ++k) if (k < c_cls) o[k - 1] = __fmul_rn(pr[k], bg);
}

extern "C" int lavb_paint_deconv_batched(const float* d_pts, int frames, int n, int pt_stride, long long pts_frame_stride,
                                         const void* d_feat, int feat_dtype, int ncam, int c_cls, int h, int w,
                                         const float* d_deconv, const float* h_cams, float* d_out, int out_stride,
                                         long long out_frame_stride, int out_col0, int copy_cols, void* stream) {
  LAVB_CHECK_ARG(frames >= 0 && frames <= 65535 && n >= 0 && pt_stride >= 3, "paint_deconv: bad sizes");
  LAVB_CHECK_ARG(ncam >= 1 && ncam <= 4 && c_cls >= 2 && c_cls <= 8, "paint_deconv: ncam 1..4, classes 2..8");
  LAVB_CHECK_ARG(h % 2 == 0 && w % 2 == 0, "paint_deconv: image size must be even (the feature map is h/2 x w/2)");
  LAVB_CHECK_ARG(copy_cols >= 0 && copy_cols <= pt_stride && out_col0 >= copy_cols && out_col0 + c_cls - 1 <= out_stride,
                 "paint_deconv: output row layout inconsistent");
  if (n == 0 || frames == 0) return 0;
  CamSet cs;
  memcpy(cs.m, h_cams, sizeof(float) * 41 * ncam);
  cs.ncam = ncam;
  const dim3 blocks(ceil_div(n, 256), frames);
  cudaStream_t st = (cudaStream_t)stream;
  const DeconvW* dw = reinterpret_cast<const DeconvW*>(d_deconv);
  if (feat_dtype == LAVB_F32)
    paint_deconv_kernel<float><<<blocks, 256, 0, st>>>(d_pts, n, pt_stride, (const float*)d_feat, c_cls, h, w, cs, dw, d_out, out_stride,
                                                       out_col0, copy_cols, pts_frame_stride, out_frame_stride);
  else if (feat_dtype == LAVB_H16)
    paint_deconv_kernel<h16><<<blocks, 256, 0, st>>>(d_pts, n, pt_stride, (const h16*)d_feat, c_cls, h, w, cs, dw, d_out, out_stride,
                                                     out_col0, copy_cols, pts_frame_stride, out_frame_stride);
  else LAVB_CHECK_ARG(false, "paint_deconv: feature dtype must be fp32 or h16");
  LAVB_LAUNCH_OK();
  return 0;
}

// ---- table-driven sweep stacking: every (frame, sweep) job of a batch in one launch; the job table lives in DEVICE
// memory so a captured CUDA graph replays with new poses / ring-buffer slots after a small H2D table update.
struct StackJob {            // 72 bytes
  const float* src; float* dst; int n; int time_idx; float R[9]; float dx, dy; int pad;
};
static_assert(sizeof(StackJob) == 72, "StackJob layout is part of the ABI (lav_b200.h)");

// Rows are 8 floats in, 8 + n_time (= 11) floats out: a thread-per-row store pattern scatters every store instruction over 32
// rows (44-byte stride, 11x the sector transactions).  The block therefore builds its 256 output rows in shared memory and copies
// them out as one contiguous, fully coalesced run.
constexpr int kStackMaxCols = 16;
__global__ void __launch_bounds__(256) stack_jobs_kernel(const StackJob* __restrict__ jobs, int src_cols, int n_time,
                                                         int roof_filter) {
  __shared__ float rows[256 * kStackMaxCols];
  const StackJob j = jobs[blockIdx.y];
  const int dcols = src_cols + n_time;
  for (int i0 = blockIdx.x * 256; i0 < j.n; i0 += gridDim.x * 256) {
    const int i = i0 + threadIdx.x;
    if (i < j.n) {
      const float* s = j.src + (size_t)i * src_cols;
      float* d = rows + threadIdx.x * dcols;
      float x, y, z;
      if (src_cols == 8) {            // fused sweeps: two aligned 16-byte loads per row
        const float4 a = __ldg(reinterpret_cast<const float4*>(s)), b = __ldg(reinterpret_cast<const float4*>(s) + 1);
        x = a.x; y = a.y; z = a.z;
        d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
      } else {
        x = __ldg(s); y = __ldg(s + 1); z = __ldg(s + 2);
        for (int k = 3; k < src_cols; ++k) d[k] = __ldg(s + k);
      }
      float nx = __fmaf_rn(z, j.R[6], __fmaf_rn(y, j.R[3], __fmul_rn(x, j.R[0])));
      float ny = __fmaf_rn(z, j.R[7], __fmaf_rn(y, j.R[4], __fmul_rn(x, j.R[1])));
      const float nz = __fmaf_rn(z, j.R[8], __fmaf_rn(y, j.R[5], __fmul_rn(x, j.R[2])));
      nx = __fadd_rn(nx, j.dx);
      ny = __fadd_rn(ny, j.dy);
      if (roof_filter && x > -2.4f && x < 0.f && y > -0.8f && y < 0.8f && z > -1.5f && z < -1.f) nx = __int_as_float(0x7fc00000);
      d[0] = nx; d[1] = ny; d[2] = nz;
      for (int k = 0; k < n_time; ++k) d[src_cols + k] = (k == j.time_idx) ? 1.f : 0.f;
    }
    __syncthreads();
    const int nrow = min(256, j.n - i0);
    float* out = j.dst + (size_t)i0 * dcols;
    for (int e = threadIdx.x; e < nrow * dcols; e += 256) out[e] = rows[e];
    __syncthreads();
  }
}

extern "C" int lavb_stack_jobs(const void* d_jobs, int n_jobs, int max_n, int src_cols, int n_time, int roof_filter, void* stream) {
  LAVB_CHECK_ARG(n_jobs >= 0 && n_jobs <= 65535 && src_cols >= 3 && n_time >= 0, "stack_jobs: bad arguments");
  LAVB_CHECK_ARG(src_cols + n_time <= kStackMaxCols, "stack_jobs: rows wider than %d floats", kStackMaxCols);
  if (n_jobs == 0 || max_n == 0) return 0;
  dim3 grid(min(ceil_div(max_n, 256), 64), n_jobs);
  stack_jobs_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const StackJob*>(d_jobs), src_cols, n_time, roof_filter);
  LAVB_LAUNCH_OK();
  return 0;
}

// Ego-roof filter as the reference applies it: an ORDER-PRESERVING drop (np.delete) of the points inside the roof box, on the
// raw sensor sweep before painting (LAVAgent.preprocess, team_code_v2/lav_agent.py:448-457, call sites :236 and
// lav_agent_fast.py:247).  One block per sweep walks it in chunks of 1024 rows: ballot + block scan give every kept row its
// output slot, so the order is the input order.  Rows past the kept count are filled with NaN (the fixed-shape pipeline's padding,
// which every downstream kernel drops).
__global__ void __launch_bounds__(1024) roof_filter_kernel(const float* __restrict__ src, int n, int cols, long long src_frame_stride,
                                                           float* __restrict__ dst, long long dst_frame_stride,
                                                           int* __restrict__ counts, int pad_nan) {
  __shared__ int warp_cnt[32];
  __shared__ int base_s;
  const float* s0 = src + (long long)blockIdx.x * src_frame_stride;
  float* d0 = dst + (long long)blockIdx.x * dst_frame_stride;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  for (int c0 = 0; c0 < n; c0 += 1024) {
    const int i = c0 + threadIdx.x;
    bool keep = false;
    if (i < n) {
      const float x = __ldg(s0 + (size_t)i * cols), y = __ldg(s0 + (size_t)i * cols + 1), z = __ldg(s0 + (size_t)i * cols + 2);
      keep = !(x > -2.4f && x < 0.f && y > -0.8f && y < 0.8f && z > -1.5f && z < -1.f);
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_cnt[warp] = __popc(m);
    __syncthreads();
    const int base = base_s;
    int woff = 0, total = 0;
    for (int w = 0; w < 32; ++w) { const int c = warp_cnt[w]; if (w < warp) woff += c; total += c; }
    if (keep) {
      const int row = base + woff + __popc(m & ((1u << lane) - 1u));
      for (int k = 0; k < cols; ++k) d0[(size_t)row * cols + k] = __ldg(s0 + (size_t)i * cols + k);
    }
    __syncthreads();
    if (threadIdx.x == 0) base_s = base + total;
    __syncthreads();
  }
  const int kept = base_s;
  if (threadIdx.x == 0 && counts) counts[blockIdx.x] = kept;
  if (pad_nan) {
    const float nanv = __int_as_float(0x7fc00000);
    for (long long e = (long long)kept * cols + threadIdx.x; e < (long long)n * cols; e += 1024) d0[e] = nanv;
  }
}

extern "C" int lavb_roof_filter(const float* d_src, int frames, int n, int cols, long long src_frame_stride, float* d_dst,
                                long long dst_frame_stride, int* d_counts, int pad_nan, void* stream) {
  LAVB_CHECK_ARG(frames >= 0 && n >= 0 && cols >= 3, "roof_filter: bad arguments");
  LAVB_CHECK_ARG(d_src != d_dst, "roof_filter: in-place compaction is not supported");
  if (frames == 0) return 0;
  roof_filter_kernel<<<frames, 1024, 0, (cudaStream_t)stream>>>(d_src, n, cols, src_frame_stride, d_dst, dst_frame_stride, d_counts,
                                                                  pad_nan);
  LAVB_LAUNCH_OK();
  return 0;
}

extern "C" int lavb_stack_sweep(const float* d_src, int n, int src_cols, const float* h_R, float dx, float dy, int time_idx,
                                int n_time, int roof_filter, float* d_dst, void* stream) {
  LAVB_CHECK_ARG(n >= 0 && src_cols >= 3 && n_time >= 0 && time_idx >= 0 && (n_time == 0 || time_idx < n_time),
                 "stack_sweep: bad arguments");
  if (n == 0) return 0;
  StackParams P;
  memcpy(P.R, h_R, sizeof(float) * 9);
  P.dx = dx; P.dy = dy;
  stack_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(d_src, n, src_cols, P, time_idx, n_time, roof_filter, d_dst);
  LAVB_LAUNCH_OK();
  return 0;
}

extern "C" int lavb_rgb_normalize(const void* d_rgb, int src_is_u8_nhwc, int n, int h, int w, void* d_out, int out_dtype,
                                  void* stream) {
  const long long npix = (long long)n * h * w;
  if (npix == 0) return 0;
  const int blocks = ceil_div(npix, 256);
  if (out_dtype == LAVB_F32)
    rgb_norm_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(d_rgb, src_is_u8_nhwc, n, h, w, (float*)d_out);
  else if (out_dtype == LAVB_H16)
    rgb_norm_kernel<h16><<<blocks, 256, 0, (cudaStream_t)stream>>>(d_rgb, src_is_u8_nhwc, n, h, w, (h16*)d_out);
  else LAVB_CHECK_ARG(false, "rgb_normalize: bad dtype");
  LAVB_LAUNCH_OK();
  return 0;
}

extern "C" int lavb_convert(const void* d_src, int src_dtype, void* d_dst, int dst_dtype, long long count, void* stream) {
  if (count == 0) return 0;
  const int blocks = ceil_div(ceil_div(count, 4), 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (src_dtype == LAVB_F32 && dst_dtype == LAVB_H16)
    convert_kernel<float, h16><<<blocks, 256, 0, st>>>((const float*)d_src, (h16*)d_dst, count);
  else if (src_dtype == LAVB_H16 && dst_dtype == LAVB_F32)
    convert_kernel<h16, float><<<blocks, 256, 0, st>>>((const h16*)d_src, (float*)d_dst, count);
  else LAVB_CHECK_ARG(false, "convert: unsupported dtype pair");
  LAVB_LAUNCH_OK();
  return 0;
}
