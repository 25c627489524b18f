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

__global__ void __launch_bounds__(256) peak_candidates_kernel(const float* __restrict__ center, int B, int H, int W, int ncls,
                                                              float min_score, int* __restrict__ counts,
                                                              float2* __restrict__ cand) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * H * W) return;
  const int x = (int)(i % W), y = (int)((i / W) % H), b = (int)(i / ((long long)W * H));
  for (int c = 0; c < ncls; ++c) {
    const float xl = __ldg(center + i * ncls + c);
    const float s = sigmoidf_ref(xl);
    if (!((double)s > (double)min_score)) continue;
    bool peak = true;
    for (int dy = -3; dy <= 3 && peak; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= H) continue;
      for (int dx = -3; dx <= 3; ++dx) {
        const int xx = x + dx;
        if (xx < 0 || xx >= W) continue;
        // sigmoid is monotonic: only a neighbour with a larger LOGIT can have a larger sigmoid; the exact fp32 sigmoid
        // comparison (ties of saturated values count as peaks, like max_pool2d on the sigmoid map) runs only then
        const float xn = __ldg(center + (((long long)b * H + yy) * W + xx) * ncls + c);
        if (xn > xl && sigmoidf_ref(xn) > s) { peak = false; break; }
      }
    }
    if (!peak) continue;
    const int slot = atomicAdd(&counts[b * ncls + c], 1);
    if (slot < kCandCap) cand[(long long)(b * ncls + c) * kCandCap + slot] = make_float2(s, __int_as_float(y * W + x));
  }
}

// one warp per (frame, class): max_det rounds of arg-max over the candidate list (ties -> smaller flat index)
__global__ void __launch_bounds__(32) peak_select_kernel(const float* __restrict__ box, const float* __restrict__ ori, int H, int W,
                                                         int ncls, int max_det, const int* __restrict__ counts,
                                                         float2* __restrict__ cand, float* __restrict__ packed) {
  const int bc = blockIdx.x, b = bc / ncls, c = bc % ncls, lane = threadIdx.x;
  const int n = min(counts[bc], kCandCap);
  float2* list = cand + (long long)bc * kCandCap;
  const int cols = ncls * max_det;
  float* out = packed + (long long)b * 7 * cols + c * max_det;
  for (int k = 0; k < max_det; ++k) {
    float best = -INFINITY; int best_loc = 0x7fffffff, best_i = -1;
    for (int j = lane; j < n; j += 32) {
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
    if (lane == 0) {
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
    __syncwarp();
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
  peak_candidates_kernel<<<ceil_div((long long)batch * h * w, 256), 256, 0, st>>>(d_center, batch, h, w, ncls, min_score, counts, cand);
  LAVB_LAUNCH_OK();
  peak_select_kernel<<<batch * ncls, 32, 0, st>>>(d_box, d_ori, h, w, ncls, max_det, counts, cand, d_packed);
  LAVB_LAUNCH_OK();
  return 0;
}
