/* lav_b200 — C ABI of the B200-native LAV frame-path kernels (sm_100a).
 *
 * The reference (dotchen/LAV) has no FFI: its boundary is Python (torch modules).  A
 * reference-side maintainer binds these entry points with ctypes;
 * lav_b200/capi.py is that binding.  Every pointer named d_* is a DEVICE pointer,
 * h_* is a HOST pointer read synchronously during the call.  `stream` is a
 * cudaStream_t passed as void* (0 = legacy default stream).  All functions return 0
 * on success and a non-zero code otherwise; lavb_last_error() gives the message
 * (thread-local).  No global mutable state; safe from several host threads
 * (nn.DataParallel-style) as long as each uses its own stream/workspace.
 *
 * Layouts: activations are NHWC ("channels last"); `*_cstride` is the number of
 * channels of the underlying buffer (pixel stride in elements) and `*_coff` the first
 * channel this call reads/writes, so concatenations are written in place.
 */
#ifndef LAV_B200_H
#define LAV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LAVB_ABI_VERSION 3

int lavb_abi_version(void);
const char* lavb_last_error(void);
/* compute capability major*10+minor of the current device, or <0 when no device */
int lavb_device_cc(void);

/* ---------------------------------------------------------------- element types */
enum { LAVB_F32 = 0, LAVB_BF16 = 1, LAVB_F16 = 2 };
/* The 16-bit storage type of the tensor-core path ("h16" below) is fixed when the library is built: IEEE half (LAVB_F16) by
 * default — fp32 accumulation everywhere, every fp32 -> half conversion saturates at +-65504 — or bfloat16 with
 * -DLAVB_H16_BF16.  Entry points that take a dtype accept LAVB_F32 and this value only. */
int lavb_h16_dtype(void);

/* ---------------------------------------------------------------- point painting
 * replaces: InferModel.point_painting / forward_paint (team_code_v2/model_inference.py:44-50,75-93),
 *           CoordConverter.forward (model_inference.py:280-297),
 *           point_painting() (lav/utils/point_painting.py:46-66).
 * h_cams: ncam x 41 floats = K(3x3 row-major) | lidar_to_world(4x4) | world_to_cam(4x4).
 * sem element (cam,c,v,u) lives at d_sem[cam*s_cam + c*s_c + v*s_y + u*s_x].
 * mode 0: gather c_in channels as they are           -> c_in outputs
 * mode 1: d_sem holds softmax probabilities (c_in>=2) -> c_in-1 outputs p[1+j]*(1-p[0])
 * mode 2: d_sem holds logits; softmax over c_in then as mode 1
 * Row i of the output receives `copy_cols` leading columns of point i, then the painted
 * channels at column `out_col0`; later cameras overwrite earlier ones; unseen points get 0. */
int lavb_paint(const float* d_pts, int n, int pt_stride,
               const float* d_sem, int ncam, int c_in, int h, int w,
               long long s_cam, long long s_c, long long s_y, long long s_x,
               const float* h_cams, int mode,
               float* d_out, int out_stride, int out_col0, int copy_cols, void* stream);

/* the same for `frames` independent agents in one launch: frame f reads points at d_pts + f*pts_frame_stride (floats),
 * semantic maps at d_sem + f*s_frame, writes rows at d_out + f*out_frame_stride (floats). */
int lavb_paint_batched(const float* d_pts, int frames, int n, int pt_stride, long long pts_frame_stride,
                       const float* d_sem, int ncam, int c_in, int h, int w, long long s_frame, long long s_cam,
                       long long s_c, long long s_y, long long s_x, const float* h_cams, int mode, float* d_out,
                       int out_stride, long long out_frame_stride, int out_col0, int copy_cols, void* stream);

/* painting straight from the ERFNet decoder's last 16-channel feature map: the segmentation head's final layer
 * replaces: Decoder.output_conv = ConvTranspose2d(16, c_cls, 2, stride 2) (lav/models/erfnet.py:122-124,132) + torch.softmax
 *           (lav_agent_fast.py:264) + the background suppression and gather of forward_paint (model_inference.py:44-50,75-93),
 * evaluated for the hit pixel only, so the (h x w x c_cls) logit maps are never materialised.
 * d_feat: NHWC (frames*ncam, h/2, w/2, 16) fp32 or h16 = input of output_conv; d_deconv: 2*2*16*8 + 8 = 520 floats =
 * w[v%2][u%2][c_in][k] (k >= c_cls zero) | bias[8].  Output row as lavb_paint mode 2: copy_cols point columns, then c_cls-1
 * painted channels at out_col0. */
int lavb_paint_deconv_batched(const float* d_pts, int frames, int n, int pt_stride, long long pts_frame_stride,
                              const void* d_feat, int feat_dtype, int ncam, int c_cls, int h, int w,
                              const float* d_deconv, const float* h_cams, float* d_out, int out_stride,
                              long long out_frame_stride, int out_col0, int copy_cols, void* stream);

/* ---------------------------------------------------------------- sweep stacking
 * replaces: LAVAgent.get_stacked_lidar + move_lidar_points (team_code_v2/lav_agent_fast.py:363-383,547-565)
 * and the ego-roof filter LAVAgent.preprocess (lav_agent.py:448-457, roof_filter!=0 marks dropped rows x=NaN).
 * dst row = [xyz @ R + (dx,dy,0) | src cols 3..src_cols | one_hot(time_idx, n_time)];  h_R is 3x3 row-major. */
int lavb_stack_sweep(const float* d_src, int n, int src_cols, const float* h_R, float dx, float dy,
                     int time_idx, int n_time, int roof_filter, float* d_dst, void* stream);

/* Ego-roof filter as an order-preserving drop (np.delete semantics).
 * replaces: LAVAgent.preprocess (team_code_v2/lav_agent.py:448-457; applied to the raw sweep before painting at :236 and
 *           lav_agent_fast.py:247): rows with x in (-2.4,0), y in (-0.8,0.8), z in (-1.5,-1) are removed, the others keep their order.
 * `frames` independent sweeps of n rows x cols floats (frame f at d_src + f*src_frame_stride); the kept rows of frame f are written
 * to d_dst + f*dst_frame_stride, their number to d_counts[f] (may be NULL); with pad_nan the remaining rows [count, n) are filled
 * with NaN (the fixed-shape pipeline's padding — every kernel drops NaN rows).  d_dst must not alias d_src. */
int lavb_roof_filter(const float* d_src, int frames, int n, int cols, long long src_frame_stride, float* d_dst,
                     long long dst_frame_stride, int* d_counts, int pad_nan, void* stream);

/* table-driven variant: d_jobs is a DEVICE array of n_jobs 72-byte records
 *   { const float* src; float* dst; int n; int time_idx; float R[9]; float dx, dy; int pad; }
 * (one per (agent, sweep)); all jobs run in one launch and the table can be rewritten between replays of a captured
 * CUDA graph (new poses, new ring-buffer slots) without touching kernel arguments. max_n = largest job. */
int lavb_stack_jobs(const void* d_jobs, int n_jobs, int max_n, int src_cols, int n_time, int roof_filter, void* stream);

/* ---------------------------------------------------------------- PointPillars voxeliser + pillar encoder
 * replaces: PointPillarNet.forward (lav/models/point_pillar.py:92-116) incl. grid_locations :70-79,
 *           pillar_generation/decorate :55-68,81-85, DynamicPointNet.forward :28-35 (torch_scatter
 *           scatter_mean/scatter_max), scatter_points :87-90.
 * Clouds: cloud b = rows [h_cloud_start[b], +h_cloud_count[b]) of d_pts (row stride pt_stride floats,
 * first `d` columns used).  MLP: Linear(d+5,h1) -> affine(s1,t1) -> ReLU -> Linear(h1,h2) -> affine -> ReLU
 * with d_w1 [h1][d+5], d_w2 [h2][h1] row-major (nn.Linear layout); the affine is BatchNorm1d expressed as
 * y*s+t (bias folded into t).  Output canvas NHWC [B][ny][nx][h2] (row = ny-1-xi, col = yi), fully written.
 * Workspace: lavb_pillar_workspace_bytes(B, nx, ny) bytes, contents irrelevant on entry. */
size_t lavb_pillar_workspace_bytes(int batch, int nx, int ny);
int lavb_pillar_forward(const float* d_pts, int pt_stride, int d,
                        const long long* h_cloud_start, const int* h_cloud_count, int batch,
                        float min_x, float max_x, float min_y, float max_y, float ppm, int nx, int ny,
                        const float* d_w1, const float* d_s1, const float* d_t1, int h1,
                        const float* d_w2, const float* d_s2, const float* d_t2, int h2,
                        void* d_canvas, int canvas_dtype, void* d_workspace, void* stream);

/* Sorted, atomic-free variant for the tensor-core pipeline (the product encoder): counting sort of the points by canvas cell,
 * layer 1 on hi/lo-split h16 operands (~ fp32), layer 2 on h16 operands with fp32 accumulation (mma.sync), one canvas row written
 * per pillar and the rows of empty cells zero-filled by the scan pass.  out_mode 0: fp32 canvas [B][ny][nx][h2]; 1: h16 canvas
 * [B][ny][nx][hi(h2) | lo(h2)] (the error-free split lavb_split_h16 produces); 2: h16 canvas [B][ny][nx][h2].  Same semantics
 * otherwise. */
size_t lavb_pillar_sorted_workspace_bytes(int batch, int nx, int ny, long long total_points);
int lavb_pillar_forward_sorted(const float* d_pts, int pt_stride, int d,
                               const long long* h_cloud_start, const int* h_cloud_count, int batch,
                               float min_x, float max_x, float min_y, float max_y, float ppm, int nx, int ny,
                               const float* d_w1, const float* d_s1, const float* d_t1, int h1,
                               const float* d_w2, const float* d_s2, const float* d_t2, int h2,
                               void* d_canvas, int out_mode, void* d_workspace, void* stream);

/* Tile-binned variant (the 16-bit pipeline's encoder): the points are binned by canvas tile (8 x 16 cells) into 48-byte records and
 * one CTA produces each tile start to finish in shared memory — per-pillar centroids, decorate, layer 1 (hi/lo-split MMAs ~ fp32),
 * layer 2 (h16 MMAs), max-pool by shared-memory atomics — and writes it, zeros included, as full 256-byte cell rows.  Same
 * arguments, semantics and out_mode as lavb_pillar_forward_sorted; no per-cell global arrays, no canvas zero-fill pass. */
size_t lavb_pillar_tiled_workspace_bytes(int batch, int nx, int ny, long long total_points);
int lavb_pillar_forward_tiled(const float* d_pts, int pt_stride, int d,
                              const long long* h_cloud_start, const int* h_cloud_count, int batch,
                              float min_x, float max_x, float min_y, float max_y, float ppm, int nx, int ny,
                              const float* d_w1, const float* d_s1, const float* d_t1, int h1,
                              const float* d_w2, const float* d_s2, const float* d_t2, int h2,
                              void* d_canvas, int out_mode, void* d_workspace, void* stream);

/* training-mode pieces (BatchNorm1d batch statistics over all in-window points, arg-routed backward).
 * stage 0: voxelise + decorate -> d_feat [M][d+5] (M = number of in-window points, returned in *h_m),
 *          d_cell [M] int32 canvas cell id (b*ny*nx + row*nx + col), -1 never appears.
 * The Linear/BN1d/ReLU stack then runs on d_feat with autograd; stage 1 max-pools rows into the canvas and
 * records the arg-max row per (cell, channel) for the backward. */
int lavb_pillar_decorate(const float* d_pts, int pt_stride, int d,
                         const long long* h_cloud_start, const int* h_cloud_count, int batch,
                         float min_x, float max_x, float min_y, float max_y, float ppm, int nx, int ny,
                         float* d_feat, int* d_cell, int* h_m, void* d_workspace, void* stream);
int lavb_pillar_scatter_max(const float* d_h, const int* d_cell, int m, int c, long long n_cells,
                            float* d_canvas, int* d_argmax, void* stream);
int lavb_pillar_scatter_max_bwd(const float* d_gcanvas, const int* d_argmax, const int* d_cell, int m, int c,
                                float* d_gh, void* stream);

/* ---------------------------------------------------------------- generic tap-list convolution (CUDA cores)
 * replaces: every nn.Conv2d / nn.ConvTranspose2d (+ the ReLU / BatchNorm / residual that follows it) of
 *           ERFNet (lav/models/erfnet.py:12-134), ConvBackbone and Head (lav/models/lidar.py:48-164).
 * out[n, oy*out_sy+out_oy, ox*out_sx+out_ox, out_coff+co] = epi( sum_t sum_ci
 *        in[n, oy*in_sy+dy[t], ox*in_sx+dx[t], in_coff+ci] * w[t][ci][co] )   (zero outside the input)
 * epi(a): a += bias[co]; if pre_relu a=max(a,0); a = a*scale[co]+shift[co]; a += res[...]; if post_relu
 *         a=max(a,0); if sigmoid a=1/(1+exp(-a)).   Null pointers skip a step.
 * d_w: [ntaps][cin][cout_pad] fp32 with cout_pad = cout rounded up to 16.
 * A strided Conv2d uses in_s=stride, dy=ky*dil-pad; a ConvTranspose2d is issued once per output phase with
 * in_s=1, out_s=stride (lav_b200/layers.py builds the tap lists). */
typedef struct {
  const void* in; int in_dtype; int n, hin, win, cin, in_cstride, in_coff;
  void* out; int out_dtype; int hout, wout, cout, out_cstride, out_coff;
  int hog, wog;                 /* output-grid extent visited by this call */
  int in_sy, in_sx, out_sy, out_sx, out_oy, out_ox;
  int ntaps; int dy[16]; int dx[16];
  const float* w; const float* bias; const float* scale; const float* shift;
  const void* res; int res_dtype; int res_cstride, res_coff;
  int pre_relu, post_relu, sigmoid;
  /* lavb_conv_umma only: depth-to-space epilogue.  d2s_nout > 0 means the (<= 32) GEMM columns are 4 output positions
   * x d2s_nout channels: column j = pos*d2s_nout + k goes to pixel (2*gy + pos/2, 2*gx + pos%2), channel k of an fp32
   * NHWC tensor [n][hout][wout][d2s_nout] (out_s must be 2).  This is how a ConvTranspose2d(k3,s2,p1,op1) with a few
   * output channels runs as a 2x2-tap GEMM (the four detection heads' last layer). */
  int d2s_nout;
} lavb_conv_desc;
int lavb_conv_taps(const lavb_conv_desc* h_desc, void* stream);

/* grouped ConvTranspose2d(k3,s2,p1,op1) with <=4 output channels per group: the 4 head output layers in one launch.
 * replaces: Head.net[3] x4 (lav/models/lidar.py:155,159-164) incl. bias and the seg head's sigmoid.
 * d_in NHWC [n][h][w][in_cstride] (group g reads channels [g*cin_g,(g+1)*cin_g)); d_w fp32 [g][cin_g][9][4]
 * (tap = ky*3+kx, cout padded to 4); d_bias [g][4]; output g: fp32 NHWC [n][2h][2w][n_out[g]]. */
int lavb_deconv3x3s2_small(const void* d_in, int dtype, int n, int h, int w, int in_cstride, int groups, int cin_g,
                           const float* d_w, const float* d_bias, const int* h_n_out, const int* h_sigmoid,
                           float* const* h_out_ptrs, void* stream);

/* 2x2/2 max-pool -> y*scale[c]+shift[c] -> ReLU into a channel slice (ERFNet DownsamplerBlock, erfnet.py:20-23) */
int lavb_pool2_affine_relu(const void* d_in, int dtype, int n, int hin, int win, int c, int in_cstride, int in_coff,
                           const float* d_scale, const float* d_shift,
                           void* d_out, int out_cstride, int out_coff, void* stream);

/* RGB ingest: uint8 NHWC (n,h,w,3) or float NCHW (n,3,h,w) in 0..255 -> (x/255-.5)*2 as NHWC with 4 channels
 * (4th = 0).  replaces RGBSegmentationModel.normalize (lav/models/rgb.py:41). */
int lavb_rgb_normalize(const void* d_rgb, int src_is_u8_nhwc, int n, int h, int w, void* d_out, int out_dtype,
                       void* stream);

/* ---------------------------------------------------------------- brake-model stem on raw camera bytes
 * replaces: Normalize + ResNet conv1(7x7,s2,p3,3->64) + bn1 + ReLU of RGBBrakePredictionModel (team_code_v2/models/rgb.py:66-70,
 * lav/models/resnet.py:178,235-238).  d_img: uint8 (batch, ncam, h, cam_w, 3) — the logical image is the ncam cameras side by
 * side (h x ncam*cam_w), ncam <= 4, cam_w % 4 == 0; d_w: BatchNorm-folded weights h16 [64][160] with
 * k = ky*22 + kx*3 + c (slot 21 of every window row and k >= 154 are zero); d_bias [64]; h_mean/h_std: the 3 ImageNet
 * constants; d_out: h16 NHWC (batch, h/2, ncam*cam_w/2, 64). */
int lavb_stem7x7s2_u8(const void* d_img, int batch, int ncam, int h, int cam_w, const void* d_w, const float* d_bias,
                      const float* h_mean, const float* h_std, void* d_out, void* stream);
/* replaces: ResNet.maxpool = MaxPool2d(3, 2, 1) (lav/models/resnet.py:181,238) on h16 NHWC (n, h, w, c), c % 8 == 0
 * -> (n, (h-1)/2+1, (w-1)/2+1, c). */
int lavb_maxpool3x3s2_nhwc(const void* d_in, int n, int h, int w, int c, void* d_out, void* stream);

/* ---------------------------------------------------------------- detection decode (device part)
 * replaces: extract_peak (team_code_v2/model_inference.py:189-202: sigmoid, 7x7 max-pool NMS, top-k) and the per-peak
 * map reads of det_inference (:100-112).  d_center: heat-map LOGITS, d_box / d_ori: size / orientation maps, all fp32
 * NHWC [batch][h][w][2] (ncls = 2 classes in d_center).  Only local maxima with sigmoid > min_score are kept (anything
 * else is dropped by the reference's host filter anyway); per (frame, class) the max_det best go to
 * d_packed [batch][7][ncls*max_det] = score | flat index | w | h | cos | sin | W, -1e5 scores padding the rest. */
size_t lavb_det_peaks_workspace_bytes(int batch, int ncls);
int lavb_det_peaks(const float* d_center, const float* d_box, const float* d_ori, int batch, int h, int w, int ncls,
                   float min_score, int max_det, float* d_packed, void* d_workspace, void* stream);

/* ---------------------------------------------------------------- rotated bilinear crop
 * replaces: UniPlanner.crop_feature (team_code_v2/models/uniplanner.py:303-340; model_inference.py:204-238) =
 *           F.affine_grid(theta, align_corners=True) + F.grid_sample(bilinear, zeros padding, align_corners=True).
 * d_feat NHWC [b][h][w][c]; crop k samples frame d_frame_idx[k] with the 2x3 affine d_theta[k]; out NHWC
 * [k][crop][crop][c] in the feature dtype. */
int lavb_crop_bilinear(const void* d_feat, int dtype, int b, int h, int w, int c, const int* d_frame_idx,
                       const float* d_theta, int k, int crop, void* d_out, void* stream);
/* gradient of the above with respect to d_feat (fp32 NHWC; training, lav/models/uniplanner.py:56-151 through F.grid_sample):
 * replaces cudnn_grid_sampler_backward + the index_put of `features[frame]`.  A gather over the crops of each frame — no
 * atomics, fixed summation order, EVERY element of d_gfeat [b][h][w][c] is written (zeros where no crop samples). */
int lavb_crop_bilinear_bwd(const float* d_gout, int b, int h, int w, int c, const int* d_frame_idx, const float* d_theta,
                           int k, int crop, float* d_gfeat, void* stream);

/* dtype / layout helpers */
/* fp32 [rows][c] -> h16 [rows][hi(c) | lo(c)] with hi = h16(x), lo = h16(x - hi) (error-free split of the canvas so the
 * first tensor-core conv sees ~fp32 input precision; its weights are duplicated along cin by the host). */
int lavb_split_h16(const float* d_src, void* d_dst, long long rows, int c, void* stream);
int lavb_convert(const void* d_src, int src_dtype, void* d_dst, int dst_dtype, long long count, void* stream);

/* ---------------------------------------------------------------- tcgen05 implicit-GEMM tap-list convolution
 * replaces (h16 path): the Conv2d -> ReLU -> BatchNorm2d layers of ConvBackbone (lidar.py:57-131), the fused 4-head
 *           384->256 conv (lidar.py:152-154) and the 64/128-channel factorised convs of ERFNet (erfnet.py:31-61).
 * Same descriptor and epilogue semantics as lavb_conv_taps, with these differences: input is h16 NHWC, cin % 64 == 0,
 * cout % 32 == 0 and <= 256, channel offsets/strides multiples of 8; d->w points to h16 weights laid out
 * [ntaps][cout][cin] (K contiguous).  Tiles are 8 x 16 output-grid pixels; operands are fetched by TMA
 * (cuTensorMapEncodeTiled through cudaGetDriverEntryPoint), accumulators live in TMEM. */
int lavb_conv_umma(const lavb_conv_desc* h_desc, void* stream);

/* ---------------------------------------------------------------- fused (3x1 -> 1x3) convolution pair
 * replaces: conv3x1_k -> ReLU -> conv1x3_k -> bn_k [-> + input] -> ReLU of non_bottleneck_1d (lav/models/erfnet.py:37-63) in
 * one tcgen05 kernel; the intermediate activation stays in shared memory.
 *   mid = relu(conv3x1_dil(in) + bias1);  out = [relu](conv1x3_dil(mid) + shift2 [+ res])
 * The BatchNorm affine (conv + b2) * s + t is folded by the caller: w2 <- w2 * s per output channel, shift2 <- b2 * s + t.
 * bias1 / shift2 (fp32 [c]) are pre-loaded into the TMEM accumulators, so the epilogues only pack, clamp and add the residual
 * (16-bit packed arithmetic: the residual add rounds once more than an fp32 add would).
 * in / out / res: h16 NHWC (n, h, w, c) contiguous, c in {64, 128}, w in {32, 64, 128}; w1 / w2: h16 [3 taps][c out][c in];
 * res may be NULL. */
typedef struct lavb_conv_pair_desc {
  const void* in; void* out; const void* res;
  int n, h, w, c, dil, post_relu;
  const void* w1; const float* bias1;
  const void* w2; const float* shift2;
} lavb_conv_pair_desc;
int lavb_conv_pair_umma(const lavb_conv_pair_desc* h_desc, void* stream);
/* profiling aid: while d_buf is non-NULL every lavb_conv_pair_umma launch writes clock64 stamps of its pipeline events,
 * d_buf[cta][tile iteration < tiles_per_cta][8] int64 (conv_pair_umma.cu: PAIR_STAMP).  NULL switches it off (the default). */
int lavb_conv_pair_set_trace(void* d_buf, int tiles_per_cta);

/* ---------------------------------------------------------------- fused ERFNet entry block on the raw camera bytes
 * replaces: RGBSegmentationModel.normalize ((x/255 - .5) * 2, lav/models/rgb.py:41-45) + Encoder.initial_block =
 * DownsamplerBlock(3, 16): relu(bn(cat[conv3x3 s2 p1 (13 ch), maxpool2x2 (3 ch)])) (lav/models/erfnet.py:12-23,67).
 * d_rgb_u8: uint8 NHWC (n, h, w, 3); h_w27x16: host fp32 [(ky*3+kx)*3+c][16] conv weights (columns >= 13 zero); h_scale16 /
 * h_shift16: host fp32 [16], epi(a) = relu(a * scale + shift) with the conv bias folded into the first 13 shifts; d_out: NHWC
 * (n, h/2, w/2, 16) fp32 or h16. */
int lavb_erf_stem(const void* d_rgb_u8, int n, int h, int w, const float* h_w27x16, const float* h_scale16,
                  const float* h_shift16, void* d_out, int out_dtype, void* stream);

/* ---------------------------------------------------------------- fused DownsamplerBlock(16, 64)
 * replaces: Encoder.layers[0] = DownsamplerBlock(16, 64) (lav/models/erfnet.py:12-23,71): relu(bn(cat[conv3x3 s2 p1 (16 -> 48),
 * maxpool2x2 (16)])).  d_in: h16 NHWC (n, h, w, 16), h and w even, w <= 128; d_out: h16 NHWC (n, h/2, w/2, 64); d_w9: fp32
 * [9 taps (ky*3+kx)][16 cin][48 cout]; d_st: fp32 [64][2] = (scale, shift), epi(a) = relu(a * scale + shift) with the conv bias
 * folded into the first 48 shifts (the last 16 apply to the pooled channels, after the max). */
int lavb_erf_down16(const void* d_in, void* d_out, int n, int h, int w, const float* d_w9, const float* d_st, void* stream);

/* ---------------------------------------------------------------- fused 16-channel non_bottleneck_1d block
 * replaces: non_bottleneck_1d(16, dropprob, dilated=1) of the ERFNet decoder (lav/models/erfnet.py:37-63, Decoder layers 4 and 5) —
 * conv3x1 -> ReLU -> conv1x3 -> bn1 -> ReLU -> conv3x1 -> ReLU -> conv1x3 -> bn2 -> (+ input) -> ReLU — in one kernel, all four
 * intermediates in shared memory.  d_in / d_out: h16 NHWC (n, h, w, 16), distinct buffers, w % 16 == 0; d_w4: fp32
 * [4 convs][3 taps][16 cin][16 cout]; d_st: fp32 [4 convs][16 cout][2] = (scale, shift) with epi(a) = relu(a * scale + shift)
 * (conv bias folded into shift; scale = 1 for the two convs that have no BatchNorm). */
int lavb_erf_nb16(const void* d_in, void* d_out, int n, int h, int w, const float* d_w4, const float* d_st, void* stream);

/* ---------------------------------------------------------------- cluster-persistent GRU roll-out
 * replaces: one call of plan_gru = nn.GRU(4, 512, batch_first=True) (team_code_v2/models/uniplanner.py:45,247-259;
 * lav/models/bev_planner_v2.py) over `steps` time steps for `nseq` sequences.
 * d_u (nseq, steps, 4) fp32; d_h0 (nseq, 512) fp32; d_whh = weight_hh_l0 (1536, 512) fp32; d_wih = weight_ih_l0
 * (1536, 4), d_bih / d_bhh (1536,) fp32; d_out (nseq, steps, 512) fp32 = the GRU's output sequence.
 * fp32-class arithmetic: the recurrent product runs on the 16-bit tensor cores with both operands split into hi + lo parts. */
int lavb_gru_h512(const float* d_u, const float* d_h0, const float* d_whh, const float* d_wih, const float* d_bih,
                  const float* d_bhh, float* d_out, int nseq, int steps, void* stream);

/* ---------------------------------------------------------------- motion-forecast ("cast") heads in one launch
 * replaces: UniPlanner.cast / BEVPlanner.cast (team_code_v2/models/uniplanner.py:286-301, lav/models/bev_planner_v2.py:226-236):
 * for each of ncmd branches, nn.GRU(512, 64, batch_first=True) over the embedding repeated `steps` times, nn.Linear(64, 2) and the
 * cumulative sum over the steps — 6 x (GRU + Linear + cumsum) calls in the reference.  fp32 FFMA.
 * d_embd (n, 512); d_wih_t (ncmd, 512, 192) = weight_ih_l0 TRANSPOSED; d_whh_t (ncmd, 64, 192) = weight_hh_l0 transposed;
 * d_bih / d_bhh (ncmd, 192); d_wmlp (ncmd, 2, 64); d_bmlp (ncmd, 2); d_out (n, ncmd, steps, 2) fp32. */
int lavb_cast_gru(const float* d_embd, int n, const float* d_wih_t, const float* d_whh_t, const float* d_bih, const float* d_bhh,
                  const float* d_wmlp, const float* d_bmlp, int ncmd, int steps, float* d_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LAV_B200_H */
