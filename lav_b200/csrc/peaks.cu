// Detection decode, device part (extract_peak of team_code_v2/model_inference.py:189-202 + the map reads of
// det_inference :100-112): sigmoid -> 7x7 max-pool NMS -> top-k -> gather size / orientation at the peaks.
// The reference builds the full NMS map and runs a 102400-wide top-k per class; only local maxima above the score
// threshold can survive the host filter (`s > min_score`), so here every pixel above the threshold checks its own 7x7
// window (rare), survivors are appended to a short candidate list, and one warp per (frame, class) selects the
// max_det best.  Output layout = InferModel.pack_peaks: [B][7][ncls*max_det] = score | flat index | w | h | cos | sin | W.
#include "common.cuh"

namespace lavb {

constexpr int kCandCap = 8192;   // candidates kept per (frame, class); heat maps with more local maxima above the threshold than this lose the surplus

__device__ __forceinline__ float sigmoidf_ref(float x) { return 1.f / (1.f + expf(-x)); }   // torch.sigmoid, fp32

// Tile kernel: a 32x32 pixel tile (+3 halo) of one class is turned into sigmoid values in shared memory, the 7x7
// window maximum is built separably (7-wide row max, then 7-high column max) and a pixel is a peak when nothing in its
// window is larger — exactly max_pool2d(heat, 7, 1, 3) followed by `max_cls > heat` (ties of saturated values all count).
constexpr int kPT = 32, kPH = 3, kPW = kPT + 2 * kPH;      // 38

__global__ void __launch_bounds__(256) peak_candidates_kernel(const float* __restrict__ center, int B, int H, int W, int ncls,
                                                              float min_score, int* __restrict__ counts,
                                                              float2* __restrict__ cand) {
  __shared__ float sv[kPW][kPW + 1];      // sigmoid values, -inf outside the map (max_pool2d pads with -inf)
  __shared__ float rm[kPW][kPT + 1];      // row-wise 7-max for the 32 centre columns
  const int tiles_x = (W + kPT - 1) / kPT, tiles_y = (H + kPT - 1) / kPT;
  int t = blockIdx.x;
  const int c = t % ncls; t /= ncls;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y; const int b = t / tiles_y;
  const int x0 = tx * kPT - kPH, y0 = ty * kPT - kPH;
  for (int i = threadIdx.x; i < kPW * kPW; i += blockDim.x) {
    const int ly = i / kPW, lx = i - ly * kPW, y = y0 + ly, x = x0 + lx;
    sv[ly][lx] = (y >= 0 && y < H && x >= 0 && x < W) ? sigmoidf_ref(__ldg(center + (((long long)b * H + y) * W + x) * ncls + c)) : -INFINITY;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kPW * kPT; i += blockDim.x) {
    const int ly = i / kPT, lx = i - ly * kPT;
    float m = sv[ly][lx];
#pragma unroll
    for (int d = 1; d < 7; ++d) m = fmaxf(m, sv[ly][lx + d]);
    rm[ly][lx] = m;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kPT * kPT; i += blockDim.x) {
    const int ly = i / kPT, lx = i - ly * kPT, y = y0 + kPH + ly, x = x0 + kPH + lx;
    if (y >= H || x >= W) continue;
    const float s = sv[ly + kPH][lx + kPH];
    if (!((double)s > (double)min_score)) continue;
    float m = rm[ly][lx];
#pragma unroll
    for (int d = 1; d < 7; ++d) m = fmaxf(m, rm[ly + d][lx]);
    if (m > s) continue;                                     // something in the 7x7 window is larger: not a peak
    const int slot = atomicAdd(&counts[b * ncls + c], 1);
    if (slot < kCandCap) cand[(long long)(b * ncls + c) * kCandCap + slot] = make_float2(s, __int_as_float(y * W + x));
  }
}

// one block per (frame, class): max_det rounds of arg-max over the candidate list (ties -> smaller flat index)
__global__ void __launch_bounds__(256) peak_select_kernel(const float* __restrict__ box, const float* __restrict__ ori, int H, int W,
                                                          int ncls, int max_det, const int* __restrict__ counts,
                                                          float2* __restrict__ cand, float* __restrict__ packed) {
  __shared__ float sb[8]; __shared__ int sl[8], si[8];
  const int bc = blockIdx.x, b = bc / ncls, c = bc % ncls, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = min(counts[bc], kCandCap);
  float2* list = cand + (long long)bc * kCandCap;
  const int cols = ncls * max_det;
  float* out = packed + (long long)b * 7 * cols + c * max_det;
  for (int k = 0; k < max_det; ++k) {
    float best = -INFINITY; int best_loc = 0x7fffffff, best_i = -1;
    for (int j = tid; j < n; j += 256) {
      const float2 e = list[j];
      const int loc = __float_as_int(e.y);
      if (e.x > best || (e.x == best && loc < best_loc)) { best = e.x; best_loc = loc; best_i = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int ol = __shfl_xor_sync(0xffffffffu, best_loc, o), oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ob > best || (ob == best && ol < best_loc)) { best = ob; best_loc = ol; best_i = oi; }
    }
    if (lane == 0) { sb[warp] = best; sl[warp] = best_loc; si[warp] = best_i; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < 8; ++w)
        if (sb[w] > best || (sb[w] == best && sl[w] < best_loc)) { best = sb[w]; best_loc = sl[w]; best_i = si[w]; }
      if (best_i >= 0) {
        list[best_i].x = -INFINITY;                       // consumed
        const long long px = ((long long)b * H * W + best_loc) * 2;
        out[0 * cols + k] = best; out[1 * cols + k] = (float)best_loc;
        out[2 * cols + k] = __ldg(box + px); out[3 * cols + k] = __ldg(box + px + 1);
        out[4 * cols + k] = __ldg(ori + px); out[5 * cols + k] = __ldg(ori + px + 1);
      } else {                                            // fewer than max_det peaks: an entry the host filter drops
        out[0 * cols + k] = -1e5f; out[1 * cols + k] = 0.f;
        out[2 * cols + k] = out[3 * cols + k] = out[4 * cols + k] = out[5 * cols + k] = 0.f;
      }
      out[6 * cols + k] = (float)W;
    }
    __syncthreads();
  }
}

}  // namespace lavb

using namespace lavb;

extern "C" size_t lavb_det_peaks_workspace_bytes(int batch, int ncls) {
  return (size_t)batch * ncls * (sizeof(int) + kCandCap * sizeof(float2)) + 256;
}

extern "C" int lavb_det_peaks(const float* d_center, const float* d_box, const float* d_ori, int batch, int h, int w, int ncls,
                              float min_score, int max_det, float* d_packed, void* d_workspace, void* stream) {
  LAVB_CHECK_ARG(ncls >= 1 && ncls <= 8 && max_det >= 1 && max_det <= 64, "det_peaks: bad ncls / max_det");
  if (batch == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int* counts = reinterpret_cast<int*>(d_workspace);
  float2* cand = reinterpret_cast<float2*>(reinterpret_cast<char*>(d_workspace) + ((size_t)batch * ncls * sizeof(int) + 255) / 256 * 256);
  LAVB_CUDA_OK(cudaMemsetAsync(counts, 0, (size_t)batch * ncls * sizeof(int), st));
  const int tiles = batch * ceil_div(h, kPT) * ceil_div(w, kPT) * ncls;
  peak_candidates_kernel<<<tiles, 256, 0, st>>>(d_center, batch, h, w, ncls, min_score, counts, cand);
  LAVB_LAUNCH_OK();
  peak_select_kernel<<<batch * ncls, 256, 0, st>>>(d_box, d_ori, h, w, ncls, max_det, counts, cand, d_packed);
  LAVB_LAUNCH_OK();
  return 0;
}
