// PointPillars dynamic voxeliser + pillar encoder (lav/models/point_pillar.py:55-116) without sort/unique:
// a pillar is addressed directly by (b, xi, yi); pass 1 accumulates the per-pillar centroid sums, pass 2
// decorates each point, runs the 2-layer point MLP and max-pools into the NHWC canvas.
#include <stdlib.h>
#include "common.cuh"

namespace lavb {

constexpr int kMaxBatch = 128;

struct Clouds {
  long long start[kMaxBatch];  // first row of cloud b in the point buffer
  int cum[kMaxBatch + 1];      // exclusive prefix of cloud sizes
  int batch;
};

struct Grid {
  float min_x, max_x, min_y, max_y, ppm;
  int nx, ny;
};

__device__ __forceinline__ int find_cloud(const Clouds& c, int i) {
  int lo = 0, hi = c.batch - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (c.cum[mid] <= i) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// grid_locations, point_pillar.py:70-79: half-open window test on the raw fp32 coordinates, then
// trunc((v - min) * ppm) in fp32.  Returns false for dropped points (NaN fails every comparison).
__device__ __forceinline__ bool locate(const Grid& g, float x, float y, int& xi, int& yi) {
  if (!(x >= g.min_x && x < g.max_x && y >= g.min_y && y < g.max_y)) return false;
  xi = (int)__fmul_rn(__fsub_rn(x, g.min_x), g.ppm);
  yi = (int)__fmul_rn(__fsub_rn(y, g.min_y), g.ppm);
  return true;
}

// pillar key space: xi in [0,nx], yi in [0,ny]  (index == n can occur by rounding)
__device__ __forceinline__ long long pillar_key(const Grid& g, int b, int xi, int yi) {
  return ((long long)b * (g.nx + 1) + xi) * (g.ny + 1) + yi;
}
// scatter_points, point_pillar.py:87-90: row = clamp(ny-1-xi), col = clamp(yi)
__device__ __forceinline__ long long canvas_cell(const Grid& g, int b, int xi, int yi) {
  int row = g.ny - 1 - xi; row = row < 0 ? 0 : (row > g.ny - 1 ? g.ny - 1 : row);
  int col = yi < 0 ? 0 : (yi > g.nx - 1 ? g.nx - 1 : yi);
  return ((long long)b * g.ny + row) * g.nx + col;
}

__global__ void __launch_bounds__(256) pillar_stats_kernel(const float* __restrict__ pts, int pt_stride,
                                                           const __grid_constant__ Clouds clouds,
                                                           const __grid_constant__ Grid g, float4* __restrict__ stats) {
  const int total = clouds.cum[clouds.batch];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = find_cloud(clouds, i);
    const float* p = pts + (clouds.start[b] + (i - clouds.cum[b])) * pt_stride;
    const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
    int xi, yi;
    if (!locate(g, x, y, xi, yi)) continue;
    atomicAdd(&stats[pillar_key(g, b, xi, yi)], make_float4(x, y, z, 1.f));  // red.global.add.v4.f32 (sm_90+)
  }
}

// decorate, point_pillar.py:55-68: [pt(D) | xyz - centroid | x - (yi/ppm + min_x) | y - (xi/ppm + min_y)]
// (the cell-origin terms use the OTHER axis' index and no +0.5 — replicated on purpose).
template <int D>
__device__ __forceinline__ void decorate(const Grid& g, const float* __restrict__ p, int xi, int yi, float4 st,
                                         float* f) {
#pragma unroll
  for (int k = 0; k < D; ++k) f[k] = __ldg(p + k);
  f[D + 0] = __fsub_rn(f[0], __fdiv_rn(st.x, st.w));
  f[D + 1] = __fsub_rn(f[1], __fdiv_rn(st.y, st.w));
  f[D + 2] = __fsub_rn(f[2], __fdiv_rn(st.z, st.w));
  f[D + 3] = __fsub_rn(f[0], __fadd_rn(__fdiv_rn((float)yi, g.ppm), g.min_x));
  f[D + 4] = __fsub_rn(f[1], __fadd_rn(__fdiv_rn((float)xi, g.ppm), g.min_y));
}

// One thread per point; both weight matrices live in shared memory k-major so that a thread reads the
// weights of 4 consecutive output channels with one broadcast LDS.128.
template <int D, int H1, int H2>
__global__ void __launch_bounds__(128) pillar_encode_kernel(const float* __restrict__ pts, int pt_stride,
                                                            const __grid_constant__ Clouds clouds,
                                                            const __grid_constant__ Grid g,
                                                            const float4* __restrict__ stats,
                                                            const float* __restrict__ w1, const float* __restrict__ s1,
                                                            const float* __restrict__ t1, const float* __restrict__ w2,
                                                            const float* __restrict__ s2, const float* __restrict__ t2,
                                                            float* __restrict__ canvas) {
  constexpr int F = D + 5;
  __shared__ __align__(16) float w1s[F][H1];
  __shared__ __align__(16) float w2s[H1][H2];
  __shared__ float s1s[H1], t1s[H1], s2s[H2], t2s[H2];
  for (int i = threadIdx.x; i < F * H1; i += blockDim.x) w1s[i % F][i / F] = __ldg(w1 + i);   // w1 is [H1][F]
  for (int i = threadIdx.x; i < H1 * H2; i += blockDim.x) w2s[i % H1][i / H1] = __ldg(w2 + i); // w2 is [H2][H1]
  for (int i = threadIdx.x; i < H1; i += blockDim.x) { s1s[i] = __ldg(s1 + i); t1s[i] = __ldg(t1 + i); }
  for (int i = threadIdx.x; i < H2; i += blockDim.x) { s2s[i] = __ldg(s2 + i); t2s[i] = __ldg(t2 + i); }
  __syncthreads();
  const int total = clouds.cum[clouds.batch];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = find_cloud(clouds, i);
    const float* p = pts + (clouds.start[b] + (i - clouds.cum[b])) * pt_stride;
    int xi, yi;
    if (!locate(g, __ldg(p), __ldg(p + 1), xi, yi)) continue;
    const float4 st = __ldg(&stats[pillar_key(g, b, xi, yi)]);
    float f[F];
    decorate<D>(g, p, xi, yi, st, f);
    float h[H1];
#pragma unroll
    for (int j = 0; j < H1; ++j) h[j] = 0.f;
#pragma unroll
    for (int k = 0; k < F; ++k) {
#pragma unroll
      for (int j = 0; j < H1; j += 4) {
        const float4 w = *reinterpret_cast<const float4*>(&w1s[k][j]);
        h[j] = fmaf(f[k], w.x, h[j]); h[j + 1] = fmaf(f[k], w.y, h[j + 1]);
        h[j + 2] = fmaf(f[k], w.z, h[j + 2]); h[j + 3] = fmaf(f[k], w.w, h[j + 3]);
      }
    }
#pragma unroll
    for (int j = 0; j < H1; ++j) { const float v = fmaf(h[j], s1s[j], t1s[j]); h[j] = v > 0.f ? v : 0.f; }
    float* cell = canvas + canvas_cell(g, b, xi, yi) * H2;
#pragma unroll 1
    for (int j0 = 0; j0 < H2; j0 += 16) {
      float o[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = 0.f;
#pragma unroll
      for (int k = 0; k < H1; ++k) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 w = *reinterpret_cast<const float4*>(&w2s[k][j0 + j]);
          o[j] = fmaf(h[k], w.x, o[j]); o[j + 1] = fmaf(h[k], w.y, o[j + 1]);
          o[j + 2] = fmaf(h[k], w.z, o[j + 2]); o[j + 3] = fmaf(h[k], w.w, o[j + 3]);
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float v = fmaf(o[j], s2s[j0 + j], t2s[j0 + j]);
        // post-ReLU values are >= +0, so float max == signed-int max on the bit pattern and the zero-filled
        // canvas is both the identity of the max and the value of empty cells.
        if (v > 0.f) atomicMax(reinterpret_cast<int*>(cell + j0 + j), __float_as_int(v));
      }
    }
  }
}

// ---------------------------------------------------------------- training-mode pieces
// order-preserving compaction of the in-window points: per-block counts -> single-block scan -> write
constexpr int kCompactBlock = 1024;

__global__ void __launch_bounds__(kCompactBlock) keep_count_kernel(const float* __restrict__ pts, int pt_stride,
                                                                   const __grid_constant__ Clouds clouds,
                                                                   const __grid_constant__ Grid g, int* __restrict__ block_count) {
  const int total = clouds.cum[clouds.batch];
  const int i = blockIdx.x * kCompactBlock + threadIdx.x;
  int keep = 0;
  if (i < total) {
    const int b = find_cloud(clouds, i);
    const float* p = pts + (clouds.start[b] + (i - clouds.cum[b])) * pt_stride;
    int xi, yi;
    keep = locate(g, __ldg(p), __ldg(p + 1), xi, yi) ? 1 : 0;
  }
  const int c = __syncthreads_count(keep);
  if (threadIdx.x == 0) block_count[blockIdx.x] = c;
}

__global__ void __launch_bounds__(1024) scan_blocks_kernel(int* __restrict__ block_count, int nblocks, int* __restrict__ total_out) {
  __shared__ int warp_sum[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nblocks ? block_count[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
    if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = warp_sum[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += y; }
      warp_sum[threadIdx.x] = w;
    }
    __syncthreads();
    const int warp_off = (threadIdx.x >> 5) ? warp_sum[(threadIdx.x >> 5) - 1] : 0;
    const int carry = carry_s;
    if (i < nblocks) block_count[i] = carry + warp_off + x - v;  // exclusive
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + warp_off + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = carry_s;
}

template <int D>
__global__ void __launch_bounds__(kCompactBlock) decorate_write_kernel(const float* __restrict__ pts, int pt_stride,
                                                                       const __grid_constant__ Clouds clouds,
                                                                       const __grid_constant__ Grid g,
                                                                       const float4* __restrict__ stats,
                                                                       const int* __restrict__ block_off,
                                                                       float* __restrict__ feat, int* __restrict__ cell) {
  constexpr int F = D + 5;
  __shared__ int warp_cnt[32];
  const int total = clouds.cum[clouds.batch];
  const int i = blockIdx.x * kCompactBlock + threadIdx.x;
  int keep = 0, xi = 0, yi = 0, b = 0;
  const float* p = nullptr;
  if (i < total) {
    b = find_cloud(clouds, i);
    p = pts + (clouds.start[b] + (i - clouds.cum[b])) * pt_stride;
    keep = locate(g, __ldg(p), __ldg(p + 1), xi, yi) ? 1 : 0;
  }
  const unsigned m = __ballot_sync(0xffffffffu, keep);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_cnt[warp] = __popc(m);
  __syncthreads();
  if (warp == 0) {
    int w = warp_cnt[lane], x = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    warp_cnt[lane] = x - w;
  }
  __syncthreads();
  if (!keep) return;
  const int row = block_off[blockIdx.x] + warp_cnt[warp] + __popc(m & ((1u << lane) - 1u));
  float f[F];
  decorate<D>(g, p, xi, yi, __ldg(&stats[pillar_key(g, b, xi, yi)]), f);
#pragma unroll
  for (int k = 0; k < F; ++k) feat[(size_t)row * F + k] = f[k];
  cell[row] = (int)canvas_cell(g, b, xi, yi);
}

__global__ void __launch_bounds__(256) scatter_max_kernel(const float* __restrict__ h, const int* __restrict__ cell,
                                                          long long mc, int c, float* __restrict__ canvas) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= mc) return;
  const int row = (int)(i / c), ch = (int)(i - (long long)row * c);
  const float v = __ldg(h + i);
  if (v > 0.f) atomicMax(reinterpret_cast<int*>(canvas + (long long)__ldg(cell + row) * c + ch), __float_as_int(v));
}
__global__ void __launch_bounds__(256) scatter_arg_kernel(const float* __restrict__ h, const int* __restrict__ cell,
                                                          long long mc, int c, const float* __restrict__ canvas,
                                                          int* __restrict__ argmax) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= mc) return;
  const int row = (int)(i / c), ch = (int)(i - (long long)row * c);
  const long long o = (long long)__ldg(cell + row) * c + ch;
  const float v = __ldg(h + i);
  const float top = canvas[o];
  if (v == top || (!(v > 0.f) && top == 0.f)) atomicMin(argmax + o, row);  // ties -> smallest row (deterministic)
}
__global__ void __launch_bounds__(256) scatter_bwd_kernel(const float* __restrict__ gcanvas, const int* __restrict__ argmax,
                                                          const int* __restrict__ cell, long long mc, int c,
                                                          float* __restrict__ gh) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= mc) return;
  const int row = (int)(i / c), ch = (int)(i - (long long)row * c);
  const long long o = (long long)__ldg(cell + row) * c + ch;
  gh[i] = (__ldg(argmax + o) == row) ? __ldg(gcanvas + o) : 0.f;
}

static int fill_clouds(Clouds& c, const long long* start, const int* count, int batch) {
  if (batch < 1 || batch > kMaxBatch) { set_error("pillar: batch must be 1..%d (got %d)", kMaxBatch, batch); return 1; }
  c.batch = batch;
  long long cum = 0;
  for (int b = 0; b < batch; ++b) {
    if (count[b] < 0 || start[b] < 0) { set_error("pillar: negative cloud start/count"); return 1; }
    c.start[b] = start[b];
    c.cum[b] = (int)cum;
    cum += count[b];
    if (cum > 0x7fffffffLL) { set_error("pillar: more than 2^31 points"); return 1; }
  }
  c.cum[batch] = (int)cum;
  return 0;
}

static size_t stats_bytes(int batch, int nx, int ny) { return (size_t)batch * (nx + 1) * (ny + 1) * sizeof(float4); }

// =====================================================================================================================
// Sorted (atomic-free canvas) pillar encoder for the tensor-core pipeline.
//   K1 count   : per point -> centroid sums (one vector RED) + per-canvas-cell point count
//   K2 offsets : exclusive scan of the counts (block sums -> single-block scan -> per-cell offsets); the same pass
//                zero-fills the canvas rows of EMPTY cells (so every canvas byte is written exactly once overall)
//                and records for every 128-slot window the first segment head at/after it (tile_start)
//   K3 fill    : counting-sort scatter of point indices into cell order
//   K4 encode  : block i owns the whole pillars whose first point lies in slots [128 i, 128 (i+1)); per 128-row chunk:
//                decorate + layer 1 (fp32 FFMA) -> h16 hidden tile in swizzled smem -> layer 2 on the tensor cores
//                (mma.sync m16n8k16 h16, fp32 accumulate; the 64x64 weight fragments live in registers) ->
//                BN affine + ReLU -> fp32 tile in smem -> 64 channel-threads walk the rows and emit one canvas row per
//                pillar (running max), as fp32 or as the [hi | lo] h16 split conv1 consumes.
// The first layer stays fp32 because its inputs are raw metric coordinates (h16 would quantise x to 0.25 m).
// =====================================================================================================================
constexpr int kCellsPerBlock = 1024;
constexpr int kRows = 128;            // slots per encode chunk
constexpr int kOsPitch = 68;          // fp32 output tile pitch (floats): conflict-free fragment stores

__global__ void __launch_bounds__(256) pillar_count_kernel(const float* __restrict__ pts, int pt_stride,
                                                           const __grid_constant__ Clouds clouds,
                                                           const __grid_constant__ Grid g, float4* __restrict__ stats,
                                                           int* __restrict__ count) {
  const int total = clouds.cum[clouds.batch];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = find_cloud(clouds, i);
    const float* p = pts + (clouds.start[b] + (i - clouds.cum[b])) * pt_stride;
    const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
    int xi, yi;
    if (!locate(g, x, y, xi, yi)) continue;
    atomicAdd(&stats[pillar_key(g, b, xi, yi)], make_float4(x, y, z, 1.f));
    atomicAdd(&count[canvas_cell(g, b, xi, yi)], 1);
  }
}

__global__ void __launch_bounds__(kCellsPerBlock) cell_block_sum_kernel(const int* __restrict__ count, long long ncells,
                                                                        int* __restrict__ block_sum) {
  const long long c = (long long)blockIdx.x * kCellsPerBlock + threadIdx.x;
  int v = c < ncells ? count[c] : 0;
  __shared__ int ws[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    int w = ws[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    if (threadIdx.x == 0) block_sum[blockIdx.x] = w;
  }
}

// per-cell exclusive offsets + zero-fill of empty canvas rows + tile_start table
__global__ void __launch_bounds__(kCellsPerBlock) cell_offsets_kernel(const int* __restrict__ count, long long ncells,
                                                                      const int* __restrict__ block_off,
                                                                      int* __restrict__ offsets, int* __restrict__ tile_start,
                                                                      const int* __restrict__ total_kept,
                                                                      uint4* __restrict__ canvas16, int row_vec16) {
  __shared__ int ws[32];
  const long long c = (long long)blockIdx.x * kCellsPerBlock + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cnt = c < ncells ? count[c] : 0;
  int x = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  if (lane == 31) ws[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int w = ws[lane], s = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
    ws[lane] = s - w;
  }
  __syncthreads();
  const int off = block_off[blockIdx.x] + ws[warp] + x - cnt;
  if (c < ncells) offsets[c] = off;
  if (c == ncells - 1) { offsets[ncells] = off + cnt; }
  if (cnt > 0) {   // window boundaries 128*i inside [off, off+cnt): first head at/after the boundary
    for (int i = (off + kRows - 1) / kRows; i * kRows < off + cnt; ++i) tile_start[i] = (i * kRows == off) ? off : off + cnt;
  }
  if (c == 0) { const int tk = *total_kept; tile_start[(tk + kRows - 1) / kRows] = tk; }
  // zero-fill: the warp's 32 cells are one contiguous canvas span; lanes sweep it 16 B at a time, skipping occupied rows
  const unsigned occ = __ballot_sync(0xffffffffu, cnt > 0 || c >= ncells);
  const long long c0 = c - lane;
  const int per_iter = 32 / row_vec16 > 0 ? 32 / row_vec16 : 1;      // rows covered per sweep step (row_vec16 = 16 B pieces / row)
  for (int r0 = 0; r0 < 32; r0 += per_iter) {
    if (row_vec16 <= 32) {
      const int r = r0 + lane / row_vec16, v = lane % row_vec16;
      if (r < 32 && !((occ >> r) & 1u)) canvas16[(c0 + r) * row_vec16 + v] = make_uint4(0u, 0u, 0u, 0u);
    } else {
      if (!((occ >> r0) & 1u))
        for (int v = lane; v < row_vec16; v += 32) canvas16[(c0 + r0) * row_vec16 + v] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}

__global__ void __launch_bounds__(256) pillar_fill_kernel(const float* __restrict__ pts, int pt_stride,
                                                          const __grid_constant__ Clouds clouds, const __grid_constant__ Grid g,
                                                          const int* __restrict__ offsets, int* __restrict__ cursor,
                                                          int* __restrict__ order, int* __restrict__ ocell) {
  const int total = clouds.cum[clouds.batch];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = find_cloud(clouds, i);
    const float* p = pts + (clouds.start[b] + (i - clouds.cum[b])) * pt_stride;
    int xi, yi;
    if (!locate(g, __ldg(p), __ldg(p + 1), xi, yi)) continue;
    const int cell = (int)canvas_cell(g, b, xi, yi);
    const int slot = __ldg(offsets + cell) + atomicAdd(cursor + cell, 1);
    order[slot] = i;
    ocell[slot] = cell;
  }
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_h16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32." LAVB_H16_PTX "." LAVB_H16_PTX ".f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

constexpr int kFsPitch = 20;          // fp32 decorated-feature tile pitch (floats)
constexpr int kW1Pitch = 24;          // h16 layer-1 weight row pitch: conflict-free B-fragment loads

struct EncSmem {                      // shared-memory plan of pillar_encode_sorted_kernel (bytes from the base)
  static constexpr int fs = 0;                                             // [128][kFsPitch] fp32
  static constexpr int os = fs + kRows * kFsPitch * 4;                     // [128][kOsPitch] fp32
  static constexpr int w1h = os + kRows * kOsPitch * 4;                    // [64][kW1Pitch] h16 (hi)
  static constexpr int w1l = w1h + 64 * kW1Pitch * 2;                      // (lo)
  static constexpr int aff = w1l + 64 * kW1Pitch * 2;                      // s1 | t1 | s2 | t2
  static constexpr int cells = aff + 4 * 64 * 4;                           // [128] int
  static constexpr int pmax = cells + kRows * 4;                           // [2 parity][first|last][4 quarters][64] fp32
  static constexpr int pcell = pmax + 2 * 2 * 4 * 64 * 4;                  // [2 parity][4 quarters][first, last, n, pad] int
  static constexpr int total = pcell + 2 * 4 * 4 * 4;
};

template <int kOutMode>      // 0: fp32 [64]; 1: h16 [hi 64 | lo 64]; 2: h16 [64]
__device__ __forceinline__ void emit_pair(void* canvas, int cell, int c, float m0, float m1) {     // channels c, c+1 (c even)
  if (kOutMode == 2) {
    *reinterpret_cast<uint32_t*>(reinterpret_cast<h16*>(canvas) + (long long)cell * 64 + c) = pack_h16(m0, m1);
  } else if (kOutMode == 1) {
    h16* o = reinterpret_cast<h16*>(canvas) + (long long)cell * 128;
    const h162 hi = floats2h162(m0, m1);
    const float2 hf = h1622float2(hi);
    const h162 lo = floats2h162(m0 - hf.x, m1 - hf.y);
    *reinterpret_cast<h162*>(o + c) = hi;
    *reinterpret_cast<h162*>(o + 64 + c) = lo;
  } else {
    *reinterpret_cast<float2*>(reinterpret_cast<float*>(canvas) + (long long)cell * 64 + c) = make_float2(m0, m1);
  }
}
template <int kOutMode>
__device__ __forceinline__ void emit_one(void* canvas, int cell, int c, float m) {
  if (kOutMode == 2) {
    reinterpret_cast<h16*>(canvas)[(long long)cell * 64 + c] = float2h16(m);
  } else if (kOutMode == 1) {
    h16* o = reinterpret_cast<h16*>(canvas) + (long long)cell * 128;
    const h16 hi = float2h16(m);
    o[c] = hi; o[64 + c] = float2h16(m - h162float(hi));
  } else {
    reinterpret_cast<float*>(canvas)[(long long)cell * 64 + c] = m;
  }
}
// fp32 pair -> h16 hi pair + h16 residual pair (error-free split to ~2^-16 relative)
__device__ __forceinline__ void split_pair(float2 f, uint32_t& hi, uint32_t& lo) {
  const h162 h = floats2h162(f.x, f.y);
  const float2 hf = h1622float2(h);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = pack_h16(f.x - hf.x, f.y - hf.y);
}

// Pillar encoder over the cell-sorted point order (both MLP layers on the tensor cores, no canvas atomics).
// A block walks "windows" of ~128 sorted rows that begin and end on pillar boundaries (tile_start), 128 rows per batch,
// warp w owning rows [32w, 32w+32) of the batch end to end — one __syncthreads per batch:
//   (1) gather + decorate: one thread per row -> 16 fp32 features in shared memory;
//   (2) layer 1 (16 -> 64) as mma.sync m16n8k16 with the features AND weights split into h16 hi + lo (3 MMAs per tile:
//       hi*hi + lo*hi + hi*lo, fp32 accumulate ~ fp32 accuracy); BN affine + ReLU on the accumulator fragments, which are
//       then re-packed IN REGISTERS as the h16 A fragments of layer 2 (64 -> 64; C-fragment layout == A-fragment layout);
//       BN affine + ReLU -> fp32 tile in shared memory;
//   (3) segmented max: each warp walks its own 32 rows, one lane per channel pair; runs that start and end inside the
//       quarter go straight to the canvas, the first / last run's partial maxima to a small table;
//   (4) after the barrier, 64 threads stitch the quarter tables with the run carried from the previous batch.
// Every canvas row of an occupied cell is written exactly once; empty cells were zero-filled by cell_offsets_kernel.
template <int D, int kOutMode>
__global__ void __launch_bounds__(kRows, 3) pillar_encode_sorted_kernel(
    const float* __restrict__ pts, int pt_stride, const __grid_constant__ Clouds clouds, const __grid_constant__ Grid g,
    const float4* __restrict__ stats, const int* __restrict__ order, const int* __restrict__ ocell,
    const int* __restrict__ tile_start, const int* __restrict__ total_kept, const float* __restrict__ w1,
    const float* __restrict__ s1, const float* __restrict__ t1, const float* __restrict__ w2, const float* __restrict__ s2,
    const float* __restrict__ t2, void* __restrict__ canvas) {
  constexpr int F = D + 5, H = 64;
  static_assert(F == 16, "layer 1 is one k16 MMA step");
  extern __shared__ __align__(128) uint8_t sm[];
  float* fs = reinterpret_cast<float*>(sm + EncSmem::fs);
  float* Os = reinterpret_cast<float*>(sm + EncSmem::os);
  h16* w1h = reinterpret_cast<h16*>(sm + EncSmem::w1h);
  h16* w1l = reinterpret_cast<h16*>(sm + EncSmem::w1l);
  float* aff = reinterpret_cast<float*>(sm + EncSmem::aff);
  int* cells = reinterpret_cast<int*>(sm + EncSmem::cells);
  float* pmax = reinterpret_cast<float*>(sm + EncSmem::pmax);
  int* pcell = reinterpret_cast<int*>(sm + EncSmem::pcell);

  const int nt = (*total_kept + kRows - 1) / kRows;
  if ((int)blockIdx.x >= nt) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gq = lane >> 2, tq = lane & 3;
  for (int i = tid; i < H * F; i += kRows) {                                      // w1 is [64 n][16 k]
    const float v = __ldg(w1 + i);
    const h16 hi = float2h16(v);
    w1h[(i / F) * kW1Pitch + i % F] = hi;
    w1l[(i / F) * kW1Pitch + i % F] = float2h16(v - h162float(hi));
  }
  for (int i = tid; i < H; i += kRows) { aff[i] = __ldg(s1 + i); aff[H + i] = __ldg(t1 + i); aff[2 * H + i] = __ldg(s2 + i); aff[3 * H + i] = __ldg(t2 + i); }
  // layer-2 weight fragments (B operand, "col" layout = rows of w2 [64 n][64 k]): b0 (k = 16kk + 2tq.., n = 8nn + gq), b1 (k + 8)
  uint32_t bfrag[4][8][2];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk)
#pragma unroll
    for (int nn = 0; nn < 8; ++nn) {
      const float* wp = w2 + (nn * 8 + gq) * H + kk * 16 + tq * 2;
      const float2 lo = __ldg(reinterpret_cast<const float2*>(wp)), hi = __ldg(reinterpret_cast<const float2*>(wp + 8));
      bfrag[kk][nn][0] = pack_h16(lo.x, lo.y);
      bfrag[kk][nn][1] = pack_h16(hi.x, hi.y);
    }
  __syncthreads();
  int it = 0;               // batch counter: parity of the quarter tables
  // persistent: the block walks windows blockIdx.x, +gridDim.x, ...
  for (int win = blockIdx.x; win < nt; win += gridDim.x) {
    const int row_begin = tile_start[win], row_end = tile_start[win + 1];
    if (row_begin >= row_end) continue;
    float run_max = 0.f;      // stitch state of channel `tid` (threads 0..63)
    int run_cell = -1;
    for (int base = row_begin; base < row_end; base += kRows, ++it) {
      const int rows = min(kRows, row_end - base);
      // ---- (1) gather + decorate (one thread per row; the row belongs to this thread's own warp)
      {
        float f[F];
#pragma unroll
        for (int k = 0; k < F; ++k) f[k] = 0.f;
        int cell = -1;
        if (tid < rows) {
          const int i = __ldg(order + base + tid);
          cell = __ldg(ocell + base + tid);
          const int b = find_cloud(clouds, i);
          const float* p = pts + (clouds.start[b] + (i - clouds.cum[b])) * pt_stride;
          int xi, yi;
          locate(g, __ldg(p), __ldg(p + 1), xi, yi);
          decorate<D>(g, p, xi, yi, __ldg(&stats[pillar_key(g, b, xi, yi)]), f);
        }
        cells[tid] = cell;
#pragma unroll
        for (int k = 0; k < F; k += 4) *reinterpret_cast<float4*>(&fs[tid * kFsPitch + k]) = make_float4(f[k], f[k + 1], f[k + 2], f[k + 3]);
      }
      __syncwarp();
      // ---- (2) both MLP layers on the tensor cores: warp w -> rows [32w, 32w+32) as two m16 tiles
#pragma unroll 1
      for (int mt = 0; mt < 2; ++mt) {
        const int row0 = warp * 32 + mt * 16;
        uint32_t ah[4], al[4];
        split_pair(*reinterpret_cast<const float2*>(&fs[(row0 + gq) * kFsPitch + 2 * tq]), ah[0], al[0]);
        split_pair(*reinterpret_cast<const float2*>(&fs[(row0 + gq + 8) * kFsPitch + 2 * tq]), ah[1], al[1]);
        split_pair(*reinterpret_cast<const float2*>(&fs[(row0 + gq) * kFsPitch + 2 * tq + 8]), ah[2], al[2]);
        split_pair(*reinterpret_cast<const float2*>(&fs[(row0 + gq + 8) * kFsPitch + 2 * tq + 8]), ah[3], al[3]);
        uint32_t a2[4][4];        // layer-2 A fragments (h16 h), one k16 step per pair of layer-1 n-tiles
#pragma unroll
        for (int nn = 0; nn < 8; ++nn) {
          const h16* wh = w1h + (nn * 8 + gq) * kW1Pitch + 2 * tq;
          const h16* wl = w1l + (nn * 8 + gq) * kW1Pitch + 2 * tq;
          const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(wh), bh1 = *reinterpret_cast<const uint32_t*>(wh + 8);
          const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(wl), bl1 = *reinterpret_cast<const uint32_t*>(wl + 8);
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
          mma_h16_16816(acc, al[0], al[1], al[2], al[3], bh0, bh1);       // small terms first
          mma_h16_16816(acc, ah[0], ah[1], ah[2], ah[3], bl0, bl1);
          mma_h16_16816(acc, ah[0], ah[1], ah[2], ah[3], bh0, bh1);
          const int col = nn * 8 + 2 * tq;
          const float2 sc = *reinterpret_cast<const float2*>(&aff[col]), sh = *reinterpret_cast<const float2*>(&aff[H + col]);
          const float sc0 = sc.x, sc1 = sc.y, sh0 = sh.x, sh1 = sh.y;
          a2[nn >> 1][(nn & 1) * 2] = pack_h16(fmaxf(fmaf(acc[0], sc0, sh0), 0.f), fmaxf(fmaf(acc[1], sc1, sh1), 0.f));       // row gq
          a2[nn >> 1][(nn & 1) * 2 + 1] = pack_h16(fmaxf(fmaf(acc[2], sc0, sh0), 0.f), fmaxf(fmaf(acc[3], sc1, sh1), 0.f));   // row gq + 8
        }
#pragma unroll
        for (int nn = 0; nn < 8; ++nn) {
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) mma_h16_16816(acc, a2[kk][0], a2[kk][1], a2[kk][2], a2[kk][3], bfrag[kk][nn][0], bfrag[kk][nn][1]);
          const int col = nn * 8 + 2 * tq;
          const float2 sc = *reinterpret_cast<const float2*>(&aff[2 * H + col]), sh = *reinterpret_cast<const float2*>(&aff[3 * H + col]);
          const float sc0 = sc.x, sc1 = sc.y, sh0 = sh.x, sh1 = sh.y;
          *reinterpret_cast<float2*>(&Os[(row0 + gq) * kOsPitch + col]) = make_float2(fmaxf(fmaf(acc[0], sc0, sh0), 0.f), fmaxf(fmaf(acc[1], sc1, sh1), 0.f));
          *reinterpret_cast<float2*>(&Os[(row0 + gq + 8) * kOsPitch + col]) = make_float2(fmaxf(fmaf(acc[2], sc0, sh0), 0.f), fmaxf(fmaf(acc[3], sc1, sh1), 0.f));
        }
      }
      __syncwarp();
      // ---- (3) quarter walk: lane = channel pair (2 lane, 2 lane + 1) over the warp's own rows
      float* pm = pmax + (it & 1) * (2 * 4 * 64);
      int* pc = pcell + (it & 1) * 16 + warp * 4;
      {
        const int n = max(0, min(32, rows - warp * 32));
        if (n > 0) {
          const int* cq = cells + warp * 32;
          const float* oq = Os + warp * 32 * kOsPitch + 2 * lane;
          const int fc = cq[0];
          int cur = fc;
          float m0 = 0.f, m1 = 0.f;
          bool first = true;
          for (int r = 0; r < n; ++r) {
            const int c = cq[r];
            if (c != cur) {
              if (first) { *reinterpret_cast<float2*>(pm + warp * 64 + 2 * lane) = make_float2(m0, m1); first = false; }
              else emit_pair<kOutMode>(canvas, cur, 2 * lane, m0, m1);
              cur = c; m0 = 0.f; m1 = 0.f;
            }
            const float2 v = *reinterpret_cast<const float2*>(oq + r * kOsPitch);
            m0 = fmaxf(m0, v.x); m1 = fmaxf(m1, v.y);
          }
          *reinterpret_cast<float2*>(pm + (first ? 0 : 4 * 64) + warp * 64 + 2 * lane) = make_float2(m0, m1);
          if (lane == 0) { pc[0] = fc; pc[1] = cur; }
        }
        if (lane == 0) pc[2] = n;
      }
      __syncthreads();
      // ---- (4) stitch the four quarters with the carried run (thread = channel)
      if (tid < H) {
        const int* pcb = pcell + (it & 1) * 16;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (pcb[q * 4 + 2] == 0) continue;
          const int fc = pcb[q * 4], lc = pcb[q * 4 + 1];
          const float fm = pm[q * 64 + tid];
          if (fc != run_cell) {
            if (run_cell >= 0) emit_one<kOutMode>(canvas, run_cell, tid, run_max);
            run_cell = fc; run_max = fm;
          } else {
            run_max = fmaxf(run_max, fm);
          }
          if (lc != fc) {
            emit_one<kOutMode>(canvas, run_cell, tid, run_max);
            run_cell = lc; run_max = pm[4 * 64 + q * 64 + tid];
          }
        }
      }
    }
    if (tid < H && run_cell >= 0) emit_one<kOutMode>(canvas, run_cell, tid, run_max);
  }   // window loop
}

struct SortedWs {
  float4* stats; int* count; int* offsets; int* cursor; int* block_sum; int* total; int* tile_start; int* order; int* ocell;
  size_t bytes;
};
static SortedWs carve_sorted(void* base, int batch, int nx, int ny, long long total_pts) {
  SortedWs w;
  const long long ncells = (long long)batch * nx * ny;
  const long long nblk = (ncells + kCellsPerBlock - 1) / kCellsPerBlock;
  char* p = reinterpret_cast<char*>(base);
  auto take = [&](size_t n) { char* r = p; p += (n + 255) / 256 * 256; return r; };
  w.stats = reinterpret_cast<float4*>(take(stats_bytes(batch, nx, ny)));
  w.count = reinterpret_cast<int*>(take(ncells * 4));
  w.cursor = reinterpret_cast<int*>(take(ncells * 4));
  w.offsets = reinterpret_cast<int*>(take((ncells + 1) * 4));
  w.block_sum = reinterpret_cast<int*>(take((nblk + 1) * 4));
  w.total = reinterpret_cast<int*>(take(256));
  w.tile_start = reinterpret_cast<int*>(take((total_pts / kRows + 2) * 4));
  w.order = reinterpret_cast<int*>(take((size_t)total_pts * 4));
  w.ocell = reinterpret_cast<int*>(take((size_t)total_pts * 4));
  w.bytes = (size_t)(p - reinterpret_cast<char*>(base));
  return w;
}


// =====================================================================================================================
// Tile-binned pillar encoder (the 16-bit pipeline's encoder): the canvas is cut into tiles of kTileR x kTileC cells and
// every tile is produced start to finish by one CTA, in shared memory:
//   K1 tile_count   : per point -> canvas tile id; per-block shared-memory histogram -> one global atomic per (block, tile)
//   K2 tile_scatter : every block re-derives its frame's exclusive tile offsets from the counts (800 entries: cheaper than a
//                     launch), ranks its points inside the block through shared-memory atomics, reserves one range per
//                     (block, tile) and writes the in-window points as 48-byte records [pt(11) | packed local cell] in tile
//                     order — out-of-window points stop here
//   K3 tile_encode  : CTA = one tile at a time (static interleave over the grid): per-pillar centroid sums in shared memory
//                     (pass A over the tile's records), then decorate + layer 1 (hi/lo-split MMAs ~ fp32) + layer 2 (h16 MMAs)
//                     exactly as the sorted kernel did, and the max-pool as shared-memory atomicMax on the fp32 tile
//                     (post-ReLU values are >= +0: int max on the bit pattern, zero = identity = empty cell); the finished
//                     tile — zeros included — leaves as full 256-byte cell rows, kTileC cells (4 KB) contiguous.
// No per-cell global arrays (stats / count / offsets / cursor), no canvas memset or zero-fill pass, no index indirection:
// global traffic = points read twice (K1 touches x,y), records written + read once, canvas written once.
// =====================================================================================================================
constexpr int kTileR = 8, kTileC = 16, kTileCells = kTileR * kTileC;     // 128 cells: a 2 m x 4 m patch at 4 px/m
constexpr int kRecF = 12;                                                // floats per binned record (48 B)
constexpr int kBinPts = 2048;                                            // points per block in K1 / K2
constexpr int kMaxTiles = 2048;                                          // tiles per frame the shared-memory histograms cover

struct TileGrid { int tiles_r, tiles_c, tiles; };

// canvas row / col BEFORE the clamp of scatter_points: row in [-1, ny-1], col in [0, nx] (index == n can occur by rounding)
__device__ __forceinline__ void raw_row_col(const Grid& g, int xi, int yi, int& row, int& col) { row = g.ny - 1 - xi; col = yi; }

__device__ __forceinline__ int tile_of(const Grid& g, const TileGrid& tg, int xi, int yi, int& packed) {
  int row, col;
  raw_row_col(g, xi, yi, row, col);
  const int crow = row < 0 ? 0 : row, ccol = col > g.nx - 1 ? g.nx - 1 : col;
  const int tr = crow / kTileR, tc = ccol / kTileC;
  // local pillar slot: (row - r0 + 1) in [0, kTileR], (col - c0) in [0, kTileC] — the overflow row/col keeps its own centroid
  packed = ((row - tr * kTileR + 1) << 8) | (col - tc * kTileC);
  return tr * tg.tiles_c + tc;
}

__global__ void __launch_bounds__(256) tile_count_kernel(const float* __restrict__ pts, int pt_stride,
                                                         const __grid_constant__ Clouds clouds, const __grid_constant__ Grid g,
                                                         const __grid_constant__ TileGrid tg, int* __restrict__ tile_count) {
  __shared__ int hist[kMaxTiles];
  const int b = blockIdx.y;
  const int n = clouds.cum[b + 1] - clouds.cum[b];
  const int i0 = blockIdx.x * kBinPts;
  if (i0 >= n) return;
  for (int t = threadIdx.x; t < tg.tiles; t += 256) hist[t] = 0;
  __syncthreads();
  const float* base = pts + clouds.start[b] * pt_stride;
  for (int i = i0 + threadIdx.x; i < min(n, i0 + kBinPts); i += 256) {
    const float* p = base + (size_t)i * pt_stride;
    int xi, yi, packed;
    if (!locate(g, __ldg(p), __ldg(p + 1), xi, yi)) continue;
    atomicAdd(&hist[tile_of(g, tg, xi, yi, packed)], 1);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < tg.tiles; t += 256)
    if (hist[t]) atomicAdd(&tile_count[b * tg.tiles + t], hist[t]);
}

template <int D>
__global__ void __launch_bounds__(256) tile_scatter_kernel(const float* __restrict__ pts, int pt_stride,
                                                           const __grid_constant__ Clouds clouds, const __grid_constant__ Grid g,
                                                           const __grid_constant__ TileGrid tg, const int* __restrict__ tile_count,
                                                           int* __restrict__ tile_cursor, int* __restrict__ tile_off,
                                                           float4* __restrict__ recs) {
  __shared__ int off[kMaxTiles];      // exclusive offsets of this frame's tiles (relative to the frame's first record)
  __shared__ int hist[kMaxTiles];     // per-block count -> then the block's base slot per tile
  __shared__ int wsum[8];
  const int b = blockIdx.y;
  const int n = clouds.cum[b + 1] - clouds.cum[b];
  const int i0 = blockIdx.x * kBinPts;
  if (i0 >= n && blockIdx.x != 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // block-wide exclusive scan of the frame's tile counts, 256 entries per round
  int carry = 0;
  for (int t0 = 0; t0 < tg.tiles; t0 += 256) {
    const int t = t0 + threadIdx.x;
    const int c = t < tg.tiles ? __ldg(tile_count + b * tg.tiles + t) : 0;
    int x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    int woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { const int v = wsum[w]; if (w < warp) woff += v; tot += v; }
    if (t < tg.tiles) { off[t] = carry + woff + x - c; hist[t] = 0; }
    carry += tot;
    __syncthreads();
  }
  if (blockIdx.x == 0) {   // one block per frame publishes the offsets for the encode kernel
    for (int t = threadIdx.x; t < tg.tiles; t += 256) tile_off[b * tg.tiles + t] = off[t];
    if (threadIdx.x == 0) tile_off[clouds.batch * tg.tiles + b] = carry;      // records of frame b (tail of the table)
  }
  if (i0 >= n) return;
  // rank inside the block
  const float* base = pts + clouds.start[b] * pt_stride;
  int my_tile[kBinPts / 256], my_rank[kBinPts / 256], my_packed[kBinPts / 256];
#pragma unroll
  for (int k = 0; k < kBinPts / 256; ++k) {
    const int i = i0 + k * 256 + threadIdx.x;
    my_tile[k] = -1;
    if (i < n) {
      const float* p = base + (size_t)i * pt_stride;
      int xi, yi;
      if (locate(g, __ldg(p), __ldg(p + 1), xi, yi)) {
        my_tile[k] = tile_of(g, tg, xi, yi, my_packed[k]);
        my_rank[k] = atomicAdd(&hist[my_tile[k]], 1);
      }
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < tg.tiles; t += 256) {
    const int c = hist[t];
    if (c) hist[t] = clouds.cum[b] + off[t] + atomicAdd(&tile_cursor[b * tg.tiles + t], c);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kBinPts / 256; ++k) {
    if (my_tile[k] < 0) continue;
    const int i = i0 + k * 256 + threadIdx.x;
    const float* p = base + (size_t)i * pt_stride;
    float f[kRecF];
#pragma unroll
    for (int e = 0; e < D; ++e) f[e] = __ldg(p + e);
#pragma unroll
    for (int e = D; e < kRecF - 1; ++e) f[e] = 0.f;
    f[kRecF - 1] = __int_as_float(my_packed[k]);
    float4* r = recs + (size_t)(hist[my_tile[k]] + my_rank[k]) * (kRecF / 4);
    r[0] = make_float4(f[0], f[1], f[2], f[3]);
    r[1] = make_float4(f[4], f[5], f[6], f[7]);
    r[2] = make_float4(f[8], f[9], f[10], f[11]);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// tcgen05 version of the tile encoder: the point MLP runs on the 5th-generation tensor cores, one row per thread.
// A chunk = 128 records of the tile = the 128 rows of an M=128 UMMA:
//   thread r decorates its record and writes row r of the A operand in shared memory (K-major, SWIZZLE_128B, written by hand:
//   16-byte piece j of row r lives at r*128 + ((j ^ (r & 7)) << 4)) as [hi(16) | lo(16) | hi(16) | 0] h16 — the error-free
//   split of the fp32 features — against B1 = [W1_hi ; W1_hi ; W1_lo ; 0]: three K16 steps give hi*Wh + lo*Wh + hi*Wl ~ fp32;
//   D1 (128 x 64 fp32) lands in TMEM; each thread reads ITS row back (tcgen05.ld 32x32b), applies BN1 + ReLU and writes the
//   h16 hidden row in place as the A operand of layer 2 (K = 64 against B2 = W2); D2 -> BN2; the max-pool is one shared-memory
//   atomicMax per (row, positive channel) into the fp32 tile.
// ~450 thread-instructions per point instead of ~1800 with mma.sync fragments.
// ---------------------------------------------------------------------------------------------------------------------
namespace tc {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t sw128_desc(uint32_t saddr) {      // K-major SWIZZLE_128B, SBO = 1024 B, version 1
  const uint32_t lo = (saddr & 0x3FFFFu) >> 4;
  const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  return (uint64_t)lo | ((uint64_t)hi << 32);
}
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
}  // namespace tc

struct TcSmem {                        // byte offsets from the 1024-aligned base
  static constexpr int a = 0;                                   // [128 rows][128 B] A operand of both layers
  static constexpr int b1 = a + 128 * 128;                      // [64 n][128 B]  W1 as [hi | hi | lo | 0]
  static constexpr int b2 = b1 + 64 * 128;                      // [64 n][128 B]  W2
  static constexpr int tile = b2 + 64 * 128;                    // [128 cells][64] fp32, columns XOR (cell & 7) << 2
  static constexpr int stats = tile + kTileCells * 64 * 4;      // [(kTileR+1)*(kTileC+1)] float4
  static constexpr int aff = stats + 2560;                      // s1 | t1 | s2 | t2
  static constexpr int cells = aff + 4 * 64 * 4;                // [128] int
  static constexpr int bars = cells + kRows * 4;                // 2 mbarriers + tmem slot
  static constexpr int total = bars + 64 + 1024;                // + alignment slack
};

// out_mode 0: fp32 [64] per cell; 1: h16 [hi 64 | lo 64]; 2: h16 [64]
template <int D, int kOutMode>
__global__ void __launch_bounds__(kRows, 3) pillar_tile_encode_tc_kernel(
    const float4* __restrict__ recs, const __grid_constant__ Clouds clouds, const __grid_constant__ Grid g,
    const __grid_constant__ TileGrid tg, const int* __restrict__ tile_count, const int* __restrict__ tile_off,
    const float* __restrict__ w1, const float* __restrict__ s1, const float* __restrict__ t1, const float* __restrict__ w2,
    const float* __restrict__ s2, const float* __restrict__ t2, void* __restrict__ canvas) {
  constexpr int F = D + 5, H = 64;
  static_assert(F == 16, "layer 1 is one k16 block per split term");
  extern __shared__ uint8_t tc_raw[];
  const uint32_t base = (tc::smem_u32(tc_raw) + 1023u) & ~1023u;
  uint8_t* gen = tc_raw + (base - tc::smem_u32(tc_raw));
  uint8_t* As = gen + TcSmem::a;
  float* tile = reinterpret_cast<float*>(gen + TcSmem::tile);
  float* stats = reinterpret_cast<float*>(gen + TcSmem::stats);
  float* aff = reinterpret_cast<float*>(gen + TcSmem::aff);
  int* cells = reinterpret_cast<int*>(gen + TcSmem::cells);
  const uint32_t bar1 = base + TcSmem::bars, bar2 = bar1 + 8, slot = bar1 + 16;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- one-time set-up: barriers, TMEM (D1 = columns [0,64), D2 = [64,128)), both weight operands, BN affines
  if (tid == 0) {
    tc::mbar_init(bar1, 1); tc::mbar_init(bar2, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int e = tid; e < 64 * 8; e += kRows) {                  // 16-byte piece jj of weight row n
    const int n = e >> 3, jj = e & 7;
    uint32_t q1[4] = {0u, 0u, 0u, 0u}, q2[4];
    if (jj < 6) {
      const float* src = w1 + n * F + (jj & 1) * 8;            // w1 is [64 n][16 k]
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const float va = __ldg(src + 2 * h), vb = __ldg(src + 2 * h + 1);
        const uint32_t hi = pack_h16(va, vb);
        if (jj < 4) q1[h] = hi;
        else { const float2 hf = unpack_h16(hi); q1[h] = pack_h16(va - hf.x, vb - hf.y); }
      }
    }
    const float* s2p = w2 + n * H + jj * 8;                     // w2 is [64 n][64 k]
#pragma unroll
    for (int h = 0; h < 4; ++h) q2[h] = pack_h16(__ldg(s2p + 2 * h), __ldg(s2p + 2 * h + 1));
    const int off = n * 128 + ((jj ^ (n & 7)) << 4);
    *reinterpret_cast<uint4*>(gen + TcSmem::b1 + off) = make_uint4(q1[0], q1[1], q1[2], q1[3]);
    *reinterpret_cast<uint4*>(gen + TcSmem::b2 + off) = make_uint4(q2[0], q2[1], q2[2], q2[3]);
  }
  for (int i = tid; i < H; i += kRows) { aff[i] = __ldg(s1 + i); aff[H + i] = __ldg(t1 + i); aff[2 * H + i] = __ldg(s2 + i); aff[3 * H + i] = __ldg(t2 + i); }
  tc::proxy_fence();
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + TcSmem::bars + 16);
  const uint32_t my_tmem = tmem + ((uint32_t)(warp * 32) << 16);
  const uint64_t a_desc = tc::sw128_desc(base + TcSmem::a), b1_desc = tc::sw128_desc(base + TcSmem::b1), b2_desc = tc::sw128_desc(base + TcSmem::b2);
  const uint32_t idesc = (1u << 4) | (kH16Fmt << 7) | (kH16Fmt << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  uint32_t phase = 0;

  const int total_tiles = clouds.batch * tg.tiles;
  constexpr int kRowBytes = kOutMode == 2 ? 128 : 256;
  for (int work = blockIdx.x; work < total_tiles; work += gridDim.x) {
    const int b = work / tg.tiles, t = work - b * tg.tiles;
    const int tr = t / tg.tiles_c, tcx = t - tr * tg.tiles_c;
    const int r0 = tr * kTileR, c0 = tcx * kTileC;
    const int n_t = __ldg(tile_count + work);
    uint8_t* cbase = reinterpret_cast<uint8_t*>(canvas) + ((size_t)b * g.ny * g.nx) * kRowBytes;
    constexpr int kVecPerCell = kRowBytes / 16;
    if (n_t == 0) {
      for (int e = tid; e < kTileCells * kVecPerCell; e += kRows) {
        const int cell = e / kVecPerCell, lr = cell / kTileC, lc = cell - lr * kTileC;
        if (r0 + lr < g.ny && c0 + lc < g.nx)
          __stcs(reinterpret_cast<uint4*>(cbase + ((size_t)(r0 + lr) * g.nx + c0 + lc) * kRowBytes) + (e % kVecPerCell), make_uint4(0u, 0u, 0u, 0u));
      }
      continue;
    }
    __syncthreads();
    for (int e = tid; e < kTileCells * 16; e += kRows) reinterpret_cast<float4*>(tile)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = tid; e < (kTileR + 1) * (kTileC + 1); e += kRows) reinterpret_cast<float4*>(stats)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const float4* rec = recs + ((size_t)clouds.cum[b] + __ldg(tile_off + work)) * (kRecF / 4);
    for (int i = tid; i < n_t; i += kRows) {                   // pass A: per-pillar centroid sums
      const float4 p0 = __ldg(rec + (size_t)i * 3);
      const int packed = __float_as_int(__ldg(reinterpret_cast<const float*>(rec + (size_t)i * 3 + 2) + 3));
      float* st = stats + ((packed >> 8) * (kTileC + 1) + (packed & 255)) * 4;
      atomicAdd(st, p0.x); atomicAdd(st + 1, p0.y); atomicAdd(st + 2, p0.z); atomicAdd(st + 3, 1.f);
    }
    __syncthreads();
    for (int cbeg = 0; cbeg < n_t; cbeg += kRows) {            // pass B: 128 records per round
      const int i = cbeg + tid;
      int cell = -1;
      {
        float f[F];
#pragma unroll
        for (int k = 0; k < F; ++k) f[k] = 0.f;
        if (i < n_t) {
          const float4 p0 = __ldg(rec + (size_t)i * 3), p1 = __ldg(rec + (size_t)i * 3 + 1), p2 = __ldg(rec + (size_t)i * 3 + 2);
          const int packed = __float_as_int(p2.w);
          const int lr1 = packed >> 8, lc = packed & 255;
          const float4 st = *reinterpret_cast<const float4*>(stats + (lr1 * (kTileC + 1) + lc) * 4);
          const int xi = g.ny - 1 - (r0 + lr1 - 1), yi = c0 + lc;
          f[0] = p0.x; f[1] = p0.y; f[2] = p0.z; f[3] = p0.w; f[4] = p1.x; f[5] = p1.y; f[6] = p1.z; f[7] = p1.w;
          f[8] = p2.x; f[9] = p2.y; f[10] = p2.z;
          f[D + 0] = __fsub_rn(f[0], __fdiv_rn(st.x, st.w));
          f[D + 1] = __fsub_rn(f[1], __fdiv_rn(st.y, st.w));
          f[D + 2] = __fsub_rn(f[2], __fdiv_rn(st.z, st.w));
          f[D + 3] = __fsub_rn(f[0], __fadd_rn(__fdiv_rn((float)yi, g.ppm), g.min_x));
          f[D + 4] = __fsub_rn(f[1], __fadd_rn(__fdiv_rn((float)xi, g.ppm), g.min_y));
          const int lr = lr1 > 0 ? lr1 - 1 : 0, lcc = lc < kTileC ? lc : kTileC - 1;
          cell = lr * kTileC + (c0 + lcc > g.nx - 1 ? g.nx - 1 - c0 : lcc);
        }
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int h = 0; h < 8; ++h) {
          hi[h] = pack_h16(f[2 * h], f[2 * h + 1]);
          const float2 hf = unpack_h16(hi[h]);
          lo[h] = pack_h16(f[2 * h] - hf.x, f[2 * h + 1] - hf.y);
        }
        uint8_t* row = As + tid * 128;
        const int sw = tid & 7;
        *reinterpret_cast<uint4*>(row + ((0 ^ sw) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(row + ((1 ^ sw) << 4)) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
        *reinterpret_cast<uint4*>(row + ((2 ^ sw) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<uint4*>(row + ((3 ^ sw) << 4)) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
        *reinterpret_cast<uint4*>(row + ((4 ^ sw) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(row + ((5 ^ sw) << 4)) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
      }
      cells[tid] = cell;
      tc::proxy_fence();                         // generic-proxy stores -> visible to the tensor core's async-proxy reads
      tc::fence_before();                        // orders this thread's TMEM reads of the previous round before the sync
      __syncthreads();
      if (tid == 0) {
        tc::fence_after();
#pragma unroll
        for (int k = 0; k < 3; ++k) tc::umma(tmem, a_desc + (uint64_t)(2 * k), b1_desc + (uint64_t)(2 * k), idesc, k ? 1u : 0u);
        tc::commit(bar1);
      }
      const bool live_warp = cbeg + warp * 32 < n_t;           // warps whose 32 rows are all past the end skip the epilogues
      if (live_warp) {
        tc::mbar_wait(bar1, phase);
        __syncwarp();
        tc::fence_after();
        uint8_t* row = As + tid * 128;
        const int sw = tid & 7;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t v[32];
          tc::tmem_ld32(my_tmem + (uint32_t)(32 * half), v);
          uint32_t w[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int c = 32 * half + 2 * j;
            w[j] = pack_h16(fmaxf(fmaf(__uint_as_float(v[2 * j]), aff[c], aff[H + c]), 0.f),
                            fmaxf(fmaf(__uint_as_float(v[2 * j + 1]), aff[c + 1], aff[H + c + 1]), 0.f));
          }
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(row + (((4 * half + q) ^ sw) << 4)) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
        }
      }
      tc::proxy_fence();
      tc::fence_before();
      __syncthreads();
      if (tid == 0) {
        tc::fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) tc::umma(tmem + 64u, a_desc + (uint64_t)(2 * k), b2_desc + (uint64_t)(2 * k), idesc, k ? 1u : 0u);
        tc::commit(bar2);
      }
      if (live_warp) {
        tc::mbar_wait(bar2, phase);
        __syncwarp();
        tc::fence_after();
        // max-pool: one shared-memory atomicMax per (row, positive channel).  Lanes = 32 different rows; the columns of a cell
        // are rotated by (cell & 7) * 4 so rows of different cells mostly hit different banks (rows of the SAME cell hit the
        // same word and are serialised by the LSU — a `__reduce_max_sync` over per-cell lane groups was tried first: with
        // non-uniform member masks it compiles to a divergent software loop and ran 10x slower).
        int* trow = reinterpret_cast<int*>(tile) + (cell < 0 ? 0 : cell) * 64;
        const int sw = (cell & 7) << 2;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t v[32];
          tc::tmem_ld32(my_tmem + (uint32_t)(64 + 32 * half), v);
          if (cell >= 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int c = 32 * half + j;
              const float val = fmaf(__uint_as_float(v[j]), aff[2 * H + c], aff[3 * H + c]);
              if (val > 0.f) atomicMax(trow + (c ^ sw), __float_as_int(val));
            }
          }
        }
      }
      phase ^= 1u;
    }
    tc::fence_before();
    __syncthreads();
    // ---- write-out, zeros included
    for (int e = tid; e < kTileCells * 16; e += kRows) {
      const int cell = e >> 4, q = e & 15, lr = cell / kTileC, lc = cell - lr * kTileC;
      if (r0 + lr >= g.ny || c0 + lc >= g.nx) continue;
      const float4 v = *reinterpret_cast<const float4*>(tile + cell * 64 + ((4 * q) ^ ((cell & 7) << 2)));
      uint8_t* row = cbase + ((size_t)(r0 + lr) * g.nx + c0 + lc) * kRowBytes;
      if (kOutMode == 0) {
        __stcs(reinterpret_cast<float4*>(row) + q, v);
      } else {
        const uint32_t h0 = pack_h16(v.x, v.y), h1 = pack_h16(v.z, v.w);
        __stcs(reinterpret_cast<uint2*>(row) + q, make_uint2(h0, h1));
        if (kOutMode == 1) {
          const float2 f0 = unpack_h16(h0), f1 = unpack_h16(h1);
          __stcs(reinterpret_cast<uint2*>(row + 128) + q, make_uint2(pack_h16(v.x - f0.x, v.y - f0.y), pack_h16(v.z - f1.x, v.w - f1.y)));
        }
      }
    }
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 0) {
    tc::fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
  }
}

struct TiledWs { int* tile_count; int* tile_cursor; int* tile_off; float4* recs; size_t bytes; };
static TiledWs carve_tiled(void* base, int batch, int tiles, long long total_pts) {
  TiledWs w;
  char* p = reinterpret_cast<char*>(base);
  auto take = [&](size_t n) { char* r = p; p += (n + 255) / 256 * 256; return r; };
  w.tile_count = reinterpret_cast<int*>(take((size_t)batch * tiles * 4));
  w.tile_cursor = reinterpret_cast<int*>(take((size_t)batch * tiles * 4));      // adjacent to tile_count: one memset
  w.tile_off = reinterpret_cast<int*>(take(((size_t)batch * tiles + batch) * 4));
  w.recs = reinterpret_cast<float4*>(take((size_t)total_pts * kRecF * 4));
  w.bytes = (size_t)(p - reinterpret_cast<char*>(base));
  return w;
}
static TileGrid make_tile_grid(int nx, int ny) {
  TileGrid tg;
  tg.tiles_r = ceil_div(ny, kTileR); tg.tiles_c = ceil_div(nx, kTileC); tg.tiles = tg.tiles_r * tg.tiles_c;
  return tg;
}

}  // namespace lavb

using namespace lavb;

extern "C" size_t lavb_pillar_workspace_bytes(int batch, int nx, int ny) {
  // centroid sums + compaction scratch (block counts for up to 2^31 points / 1024) + total
  return stats_bytes(batch, nx, ny) + ((size_t)(1 << 21) + 16) * sizeof(int);
}

extern "C" int lavb_pillar_forward(const float* d_pts, int pt_stride, int d, const long long* h_cloud_start,
                                   const int* h_cloud_count, int batch, float min_x, float max_x, float min_y, float max_y,
                                   float ppm, int nx, int ny, const float* d_w1, const float* d_s1, const float* d_t1, int h1,
                                   const float* d_w2, const float* d_s2, const float* d_t2, int h2, void* d_canvas,
                                   int canvas_dtype, void* d_workspace, void* stream) {
  Clouds clouds;
  if (fill_clouds(clouds, h_cloud_start, h_cloud_count, batch)) return 1;
  LAVB_CHECK_ARG(d == 11 && h1 == 64 && h2 == 64, "pillar_forward: only the v2 configuration (D=11, features [64,64]) is built (got D=%d [%d,%d])", d, h1, h2);
  LAVB_CHECK_ARG(canvas_dtype == LAVB_F32, "pillar_forward: canvas must be fp32");
  LAVB_CHECK_ARG(pt_stride >= d, "pillar_forward: pt_stride < d");
  cudaStream_t st = (cudaStream_t)stream;
  Grid g{min_x, max_x, min_y, max_y, ppm, nx, ny};
  float4* stats = reinterpret_cast<float4*>(d_workspace);
  LAVB_CUDA_OK(cudaMemsetAsync(stats, 0, stats_bytes(batch, nx, ny), st));
  LAVB_CUDA_OK(cudaMemsetAsync(d_canvas, 0, (size_t)batch * nx * ny * h2 * sizeof(float), st));
  const int total = clouds.cum[batch];
  if (total == 0) return 0;
  const int blocks1 = min(ceil_div(total, 256), kNumSMs * 8);
  pillar_stats_kernel<<<blocks1, 256, 0, st>>>(d_pts, pt_stride, clouds, g, stats);
  LAVB_LAUNCH_OK();
  const int blocks2 = min(ceil_div(total, 128), kNumSMs * 4);
  pillar_encode_kernel<11, 64, 64><<<blocks2, 128, 0, st>>>(d_pts, pt_stride, clouds, g, stats, d_w1, d_s1, d_t1, d_w2, d_s2,
                                                              d_t2, reinterpret_cast<float*>(d_canvas));
  LAVB_LAUNCH_OK();
  return 0;
}

extern "C" int lavb_pillar_decorate(const float* d_pts, int pt_stride, int d, const long long* h_cloud_start,
                                    const int* h_cloud_count, int batch, float min_x, float max_x, float min_y, float max_y,
                                    float ppm, int nx, int ny, float* d_feat, int* d_cell, int* h_m, void* d_workspace,
                                    void* stream) {
  Clouds clouds;
  if (fill_clouds(clouds, h_cloud_start, h_cloud_count, batch)) return 1;
  LAVB_CHECK_ARG(d == 11, "pillar_decorate: only D=11 is built (got %d)", d);
  cudaStream_t st = (cudaStream_t)stream;
  Grid g{min_x, max_x, min_y, max_y, ppm, nx, ny};
  float4* stats = reinterpret_cast<float4*>(d_workspace);
  int* scratch = reinterpret_cast<int*>(reinterpret_cast<char*>(d_workspace) + stats_bytes(batch, nx, ny));
  const int total = clouds.cum[batch];
  *h_m = 0;
  if (total == 0) return 0;
  const int nblk = ceil_div(total, kCompactBlock);
  LAVB_CHECK_ARG(nblk <= (1 << 21), "pillar_decorate: too many points");
  int* d_total = scratch + (1 << 21);
  LAVB_CUDA_OK(cudaMemsetAsync(stats, 0, stats_bytes(batch, nx, ny), st));
  pillar_stats_kernel<<<min(ceil_div(total, 256), kNumSMs * 8), 256, 0, st>>>(d_pts, pt_stride, clouds, g, stats);
  LAVB_LAUNCH_OK();
  keep_count_kernel<<<nblk, kCompactBlock, 0, st>>>(d_pts, pt_stride, clouds, g, scratch);
  LAVB_LAUNCH_OK();
  scan_blocks_kernel<<<1, 1024, 0, st>>>(scratch, nblk, d_total);
  LAVB_LAUNCH_OK();
  if (d_feat != nullptr) {
    decorate_write_kernel<11><<<nblk, kCompactBlock, 0, st>>>(d_pts, pt_stride, clouds, g, stats, scratch, d_feat, d_cell);
    LAVB_LAUNCH_OK();
  }
  LAVB_CUDA_OK(cudaMemcpyAsync(h_m, d_total, sizeof(int), cudaMemcpyDeviceToHost, st));
  LAVB_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int lavb_pillar_scatter_max(const float* d_h, const int* d_cell, int m, int c, long long n_cells, float* d_canvas,
                                       int* d_argmax, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  LAVB_CUDA_OK(cudaMemsetAsync(d_canvas, 0, (size_t)n_cells * c * sizeof(float), st));
  if (d_argmax) LAVB_CUDA_OK(cudaMemsetAsync(d_argmax, 0x7f, (size_t)n_cells * c * sizeof(int), st));
  const long long mc = (long long)m * c;
  if (mc == 0) return 0;
  scatter_max_kernel<<<ceil_div(mc, 256), 256, 0, st>>>(d_h, d_cell, mc, c, d_canvas);
  LAVB_LAUNCH_OK();
  if (d_argmax) {
    scatter_arg_kernel<<<ceil_div(mc, 256), 256, 0, st>>>(d_h, d_cell, mc, c, d_canvas, d_argmax);
    LAVB_LAUNCH_OK();
  }
  return 0;
}

extern "C" int lavb_pillar_scatter_max_bwd(const float* d_gcanvas, const int* d_argmax, const int* d_cell, int m, int c,
                                                 float* d_gh, void* stream) {
  const long long mc = (long long)m * c;
  if (mc == 0) return 0;
  scatter_bwd_kernel<<<ceil_div(mc, 256), 256, 0, (cudaStream_t)stream>>>(d_gcanvas, d_argmax, d_cell, mc, c, d_gh);
  LAVB_LAUNCH_OK();
  return 0;
}

extern "C" size_t lavb_pillar_sorted_workspace_bytes(int batch, int nx, int ny, long long total_points) {
  return carve_sorted(nullptr, batch, nx, ny, total_points).bytes;
}

extern "C" int lavb_pillar_forward_sorted(const float* d_pts, int pt_stride, int d, const long long* h_cloud_start,
                                          const int* h_cloud_count, int batch, float min_x, float max_x, float min_y,
                                          float max_y, float ppm, int nx, int ny, const float* d_w1, const float* d_s1,
                                          const float* d_t1, int h1, const float* d_w2, const float* d_s2, const float* d_t2,
                                          int h2, void* d_canvas, int out_mode, void* d_workspace, void* stream) {
  Clouds clouds;
  if (fill_clouds(clouds, h_cloud_start, h_cloud_count, batch)) return 1;
  LAVB_CHECK_ARG(d == 11 && h1 == 64 && h2 == 64, "pillar_forward_sorted: only the v2 configuration (D=11, features [64,64]) is built");
  LAVB_CHECK_ARG(out_mode >= 0 && out_mode <= 2, "pillar_forward_sorted: out_mode 0 (fp32), 1 (h16 hi|lo split) or 2 (h16)");
  LAVB_CHECK_ARG(pt_stride >= d, "pillar_forward_sorted: pt_stride < d");
  cudaStream_t st = (cudaStream_t)stream;
  Grid g{min_x, max_x, min_y, max_y, ppm, nx, ny};
  const int total = clouds.cum[batch];
  const long long ncells = (long long)batch * nx * ny;
  LAVB_CHECK_ARG(ncells < (1LL << 31), "pillar_forward_sorted: too many cells");
  const SortedWs w = carve_sorted(d_workspace, batch, nx, ny, total);
  const int row_bytes = h2 * (out_mode == 2 ? 2 : 4);        // fp32: 64*4; split: 128*2; h16: 64*2
  LAVB_CUDA_OK(cudaMemsetAsync(w.stats, 0, stats_bytes(batch, nx, ny), st));
  LAVB_CUDA_OK(cudaMemsetAsync(w.count, 0, (size_t)ncells * 8 + 512, st));       // count + cursor (adjacent, 256 B padded)
  LAVB_CUDA_OK(cudaMemsetAsync(w.tile_start, 0, ((size_t)total / kRows + 2) * 4, st));
  if (total > 0) {
    pillar_count_kernel<<<min(ceil_div(total, 256), kNumSMs * 8), 256, 0, st>>>(d_pts, pt_stride, clouds, g, w.stats, w.count);
    LAVB_LAUNCH_OK();
  }
  const int nblk = ceil_div(ncells, kCellsPerBlock);
  cell_block_sum_kernel<<<nblk, kCellsPerBlock, 0, st>>>(w.count, ncells, w.block_sum);
  LAVB_LAUNCH_OK();
  scan_blocks_kernel<<<1, 1024, 0, st>>>(w.block_sum, nblk, w.total);
  LAVB_LAUNCH_OK();
  cell_offsets_kernel<<<nblk, kCellsPerBlock, 0, st>>>(w.count, ncells, w.block_sum, w.offsets, w.tile_start, w.total,
                                                       reinterpret_cast<uint4*>(d_canvas), row_bytes / 16);
  LAVB_LAUNCH_OK();
  if (total == 0) return 0;
  pillar_fill_kernel<<<min(ceil_div(total, 256), kNumSMs * 8), 256, 0, st>>>(d_pts, pt_stride, clouds, g, w.offsets, w.cursor,
                                                                             w.order, w.ocell);
  LAVB_LAUNCH_OK();
  const size_t smem = EncSmem::total;
  LAVB_CUDA_OK(ensure_dyn_smem((const void*)pillar_encode_sorted_kernel<11, 0>, 72 * 1024));
  LAVB_CUDA_OK(ensure_dyn_smem((const void*)pillar_encode_sorted_kernel<11, 1>, 72 * 1024));
  LAVB_CUDA_OK(ensure_dyn_smem((const void*)pillar_encode_sorted_kernel<11, 2>, 72 * 1024));
  const int ntiles = min(ceil_div(total, kRows), kNumSMs * 3);     // persistent: 3 resident blocks per SM
#define LAVB_SORTED_LAUNCH(M)                                                                                                        \
  pillar_encode_sorted_kernel<11, M><<<ntiles, kRows, smem, st>>>(d_pts, pt_stride, clouds, g, w.stats, w.order, w.ocell, w.tile_start, \
                                                                  w.total, d_w1, d_s1, d_t1, d_w2, d_s2, d_t2, d_canvas)
  if (out_mode == 0) LAVB_SORTED_LAUNCH(0); else if (out_mode == 1) LAVB_SORTED_LAUNCH(1); else LAVB_SORTED_LAUNCH(2);
#undef LAVB_SORTED_LAUNCH
  LAVB_LAUNCH_OK();
  return 0;
}

extern "C" size_t lavb_pillar_tiled_workspace_bytes(int batch, int nx, int ny, long long total_points) {
  return carve_tiled(nullptr, batch, make_tile_grid(nx, ny).tiles, total_points).bytes;
}

extern "C" int lavb_pillar_forward_tiled(const float* d_pts, int pt_stride, int d, const long long* h_cloud_start,
                                         const int* h_cloud_count, int batch, float min_x, float max_x, float min_y,
                                         float max_y, float ppm, int nx, int ny, const float* d_w1, const float* d_s1,
                                         const float* d_t1, int h1, const float* d_w2, const float* d_s2, const float* d_t2,
                                         int h2, void* d_canvas, int out_mode, void* d_workspace, void* stream) {
  Clouds clouds;
  if (fill_clouds(clouds, h_cloud_start, h_cloud_count, batch)) return 1;
  LAVB_CHECK_ARG(d == 11 && h1 == 64 && h2 == 64, "pillar_forward_tiled: only the v2 configuration (D=11, features [64,64]) is built");
  LAVB_CHECK_ARG(out_mode >= 0 && out_mode <= 2, "pillar_forward_tiled: out_mode 0 (fp32), 1 (h16 hi|lo split) or 2 (h16)");
  LAVB_CHECK_ARG(pt_stride >= d, "pillar_forward_tiled: pt_stride < d");
  const TileGrid tg = make_tile_grid(nx, ny);
  LAVB_CHECK_ARG(tg.tiles <= kMaxTiles, "pillar_forward_tiled: grid of %d x %d cells has more than %d tiles", nx, ny, kMaxTiles);
  LAVB_CHECK_ARG((long long)batch * tg.tiles < (1LL << 31), "pillar_forward_tiled: too many tiles");
  cudaStream_t st = (cudaStream_t)stream;
  Grid g{min_x, max_x, min_y, max_y, ppm, nx, ny};
  const int total = clouds.cum[batch];
  const TiledWs w = carve_tiled(d_workspace, batch, tg.tiles, total);
  LAVB_CUDA_OK(cudaMemsetAsync(w.tile_count, 0, (size_t)(reinterpret_cast<char*>(w.tile_off) - reinterpret_cast<char*>(w.tile_count)), st));
  int max_n = 0;
  for (int b = 0; b < batch; ++b) max_n = max(max_n, h_cloud_count[b]);
  if (max_n > 0) {
    dim3 grid(ceil_div(max_n, kBinPts), batch);
    tile_count_kernel<<<grid, 256, 0, st>>>(d_pts, pt_stride, clouds, g, tg, w.tile_count);
    LAVB_LAUNCH_OK();
    tile_scatter_kernel<11><<<grid, 256, 0, st>>>(d_pts, pt_stride, clouds, g, tg, w.tile_count, w.tile_cursor, w.tile_off, w.recs);
    LAVB_LAUNCH_OK();
  } else {
    LAVB_CUDA_OK(cudaMemsetAsync(w.tile_off, 0, ((size_t)batch * tg.tiles + batch) * 4, st));
  }
  const int grid4 = min(batch * tg.tiles, kNumSMs * 3);              // persistent: 3 resident CTAs per SM
#define LAVB_TC_LAUNCH(M)                                                                                                          \
  {                                                                                                                               \
    LAVB_CUDA_OK(ensure_dyn_smem((const void*)pillar_tile_encode_tc_kernel<11, M>, TcSmem::total));                               \
    pillar_tile_encode_tc_kernel<11, M><<<grid4, kRows, TcSmem::total, st>>>(w.recs, clouds, g, tg, w.tile_count, w.tile_off, d_w1,  \
                                                                             d_s1, d_t1, d_w2, d_s2, d_t2, d_canvas);             \
  }
  if (out_mode == 0) LAVB_TC_LAUNCH(0) else if (out_mode == 1) LAVB_TC_LAUNCH(1) else LAVB_TC_LAUNCH(2)
#undef LAVB_TC_LAUNCH
  LAVB_LAUNCH_OK();
  return 0;
}
