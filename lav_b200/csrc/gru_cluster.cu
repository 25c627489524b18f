// Cluster-persistent plan GRU (lav_b200.heads.GRU_KERNEL): the plan GRU roll-out of UniPlanner / BEVPlanner
// (team_code_v2/models/uniplanner.py:227-259: nn.GRU(4, 512), 20 steps, 6*B sequences, called 5 times per tick) as ONE
// cluster-persistent kernel per call instead of 20 x (cuDNN GEMM + cell kernel): those 100 sequential launch pairs per tick are
// latency-bound (~10 us each whatever the batch: 1.0 ms per tick, 31 us/frame at 32 frames).
//   * a thread-block cluster of 16 CTAs owns 32 sequences for all T steps; CTA r produces the r/z/n gates of hidden units
//     [32r, 32r+32): 96 rows of W_hh, resident for the whole roll-out;
//   * fp32-class arithmetic on the 16-bit tensor cores: both operands are split error-free into h16 hi + lo parts and
//     h.W = h_hi W_hi + h_lo W_hi + h_hi W_lo (fp32 accumulate; the dropped lo*lo term is 2^-22 relative).  The plan roll-out
//     feeds its own output back five times and is not contractive on untrained weights, so plain 16-bit operands (the first
//     version of this kernel) ended 1.7e-2..5e-2 away from the fp32 reference; this one agrees to ~1e-6 per roll-out.
//     W_hi lives in shared memory (96 KB), W_lo in REGISTERS as the warp's mma B fragments (64 registers per thread: warp w
//     owns gate columns [8w, 8w+8) for all 512 k), the hidden state as hi/lo h16 copies in shared memory (one buffer:
//     96 KB + 65 KB leave no room for a second);
//   * per step every CTA multiplies the full hidden state of its 32 sequences with its weight slice (mma.sync m16n8k16, six
//     independent accumulator chains per warp), ARRIVES at the cluster barrier ("done reading h"), applies the gate math in
//     fp32 on its 32 units (fp32 master copy of h stays local) and writes the step's output rows while the barrier completes,
//     then PUSHES the hi/lo h16 slices of h' into the buffers of all 16 CTAs through distributed shared memory and closes the
//     step with a second cluster barrier.
// PyTorch gate order and formulas (r, z, n; n = tanh(W_in x + b_in + r * (W_hn h + b_hn)); h' = (1 - z) n + z h).
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace lavb {

constexpr int kGruH = 512, kGruSeq = 32, kGruUnits = 32, kGruCluster = 16, kGruCols = 96, kGruIn = 4;
constexpr int kGruWarps = kGruCols / 8;                            // one n-tile (8 gate columns) per warp
constexpr int kGruThreads = 32 * kGruWarps;                        // 384
constexpr int kGruWPitch = kGruH + 8, kGruHPitch = kGruH + 8;     // h16 elements; +8 keeps ldmatrix / fragment loads conflict-free
constexpr int kGruGPitch = kGruCols + 4;                           // fp32 gate pre-activations [seq][96]

struct GruSmem {
  static constexpr int w = 0;                                                   // [96][kGruWPitch] h16: W_hi
  static constexpr int hb = w + kGruCols * kGruWPitch * 2;                      // [hi, lo][32][kGruHPitch] h16
  static constexpr int g = hb + 2 * kGruSeq * kGruHPitch * 2;                   // [32][kGruGPitch] fp32
  static constexpr int hm = g + kGruSeq * kGruGPitch * 4;                       // [32][32] fp32 master copy of this CTA's units
  static constexpr int wih = hm + kGruSeq * kGruUnits * 4;                      // [96][4] fp32
  static constexpr int bih = wih + kGruCols * kGruIn * 4;                       // [96]
  static constexpr int bhh = bih + kGruCols * 4;                                // [96]
  static constexpr int total = bhh + kGruCols * 4;
};

__device__ __forceinline__ void gru_ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void gru_mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32." LAVB_H16_PTX "." LAVB_H16_PTX ".f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float gru_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }
// error-free split of an fp32 pair into h16 hi and lo pairs: x = hi + lo + O(2^-22 |x|)
__device__ __forceinline__ void gru_split2(float x, float y, uint32_t& hi, uint32_t& lo) {
  const h162 h = floats2h162(x, y);
  const float2 hf = h1622float2(h);
  const h162 l = floats2h162(x - hf.x, y - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

__global__ void __launch_bounds__(kGruThreads, 1) gru_cluster_kernel(const float* __restrict__ u, const float* __restrict__ h0,
                                                                     const float* __restrict__ whh, const float* __restrict__ wih,
                                                                     const float* __restrict__ bih, const float* __restrict__ bhh,
                                                                     float* __restrict__ out, int nseq, int steps) {
  extern __shared__ __align__(16) uint8_t gsm[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();                       // which 32 hidden units this CTA owns
  const int seq0 = ((int)blockIdx.x / kGruCluster) * kGruSeq;       // first sequence of this cluster
  h16* Ws = reinterpret_cast<h16*>(gsm + GruSmem::w);
  h16* Hb = reinterpret_cast<h16*>(gsm + GruSmem::hb);              // part p (0 hi, 1 lo) at Hb + p * kPart
  float* G = reinterpret_cast<float*>(gsm + GruSmem::g);
  float* Hm = reinterpret_cast<float*>(gsm + GruSmem::hm);
  float* Wi = reinterpret_cast<float*>(gsm + GruSmem::wih);
  float* Bi = reinterpret_cast<float*>(gsm + GruSmem::bih);
  float* Bh = reinterpret_cast<float*>(gsm + GruSmem::bhh);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gq = lane >> 2, tq = lane & 3;
  constexpr int kPart = kGruSeq * kGruHPitch;                       // h16 elements of one part

  // ---- one-time staging: local row c = gate*32 + j  <->  global row gate*512 + 32*rank + j
  for (int i = tid; i < kGruCols * (kGruH / 2); i += kGruThreads) {
    const int c = i / (kGruH / 2), k2 = i - c * (kGruH / 2);
    const int grow = (c >> 5) * kGruH + rank * kGruUnits + (c & 31);
    const float2 wv = __ldg(reinterpret_cast<const float2*>(whh + (long long)grow * kGruH) + k2);
    uint32_t hi, lo;
    gru_split2(wv.x, wv.y, hi, lo);
    *reinterpret_cast<uint32_t*>(Ws + c * kGruWPitch + 2 * k2) = hi;
  }
  // W_lo of this warp's 8 columns as mma B fragments (col-major K x 8): b0 = {W[col gq][16kk + 2tq], +1}, b1 = the same + 8
  uint32_t wlo[2 * (kGruH / 16)];
  {
    const int c = warp * 8 + gq;
    const int grow = (c >> 5) * kGruH + rank * kGruUnits + (c & 31);
    const float* wr = whh + (long long)grow * kGruH;
#pragma unroll
    for (int kk = 0; kk < kGruH / 16; ++kk) {
      const float2 v0 = __ldg(reinterpret_cast<const float2*>(wr + kk * 16 + 2 * tq));
      const float2 v1 = __ldg(reinterpret_cast<const float2*>(wr + kk * 16 + 2 * tq + 8));
      uint32_t hi;
      gru_split2(v0.x, v0.y, hi, wlo[2 * kk]);
      gru_split2(v1.x, v1.y, hi, wlo[2 * kk + 1]);
    }
  }
  for (int i = tid; i < kGruCols; i += kGruThreads) {
    const int grow = (i >> 5) * kGruH + rank * kGruUnits + (i & 31);
#pragma unroll
    for (int k = 0; k < kGruIn; ++k) Wi[i * kGruIn + k] = __ldg(wih + grow * kGruIn + k);
    Bi[i] = __ldg(bih + grow);
    Bh[i] = __ldg(bhh + grow);
  }
  for (int i = tid; i < kGruSeq * (kGruH / 2); i += kGruThreads) {
    const int s = i / (kGruH / 2), k = 2 * (i - s * (kGruH / 2));
    float2 v = make_float2(0.f, 0.f);
    if (seq0 + s < nseq) v = __ldg(reinterpret_cast<const float2*>(h0 + (long long)(seq0 + s) * kGruH + k));
    uint32_t hi, lo;
    gru_split2(v.x, v.y, hi, lo);
    *reinterpret_cast<uint32_t*>(Hb + s * kGruHPitch + k) = hi;
    *reinterpret_cast<uint32_t*>(Hb + kPart + s * kGruHPitch + k) = lo;
    const int j = k - rank * kGruUnits;
    if (j >= 0 && j < kGruUnits) { Hm[s * kGruUnits + j] = v.x; Hm[s * kGruUnits + j + 1] = v.y; }
  }
  cluster.sync();                                                   // every CTA of the cluster is resident and initialised

  const int gs = tid >> 3, gu = (tid & 7) * 4;                      // gate role (threads 0..255): sequence, first of 4 hidden units
  const bool gate_thread = tid < kGruSeq * 8;
  const bool seq_ok = gate_thread && seq0 + gs < nseq;
  const uint32_t a_hi = (uint32_t)__cvta_generic_to_shared(Hb + (lane & 15) * kGruHPitch + (lane >> 4) * 8);
  const uint32_t a_lo = a_hi + kPart * 2;
  const h16* wp = Ws + (warp * 8 + gq) * kGruWPitch + 2 * tq;
  for (int t = 0; t < steps; ++t) {
    // ---- (1) G[32 seq][96] = h (32 x 512) . Wslice^T: warp w -> columns [8w, 8w+8), both 16-sequence m-tiles, three split
    //          products per k-tile into six independent accumulators
    float acc[2][3][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[m][c][0] = acc[m][c][1] = acc[m][c][2] = acc[m][c][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < kGruH / 16; ++kk) {
      const uint32_t b0 = *reinterpret_cast<const uint32_t*>(wp + kk * 16), b1 = *reinterpret_cast<const uint32_t*>(wp + kk * 16 + 8);
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        uint32_t h0_, h1_, h2_, h3_, l0_, l1_, l2_, l3_;
        gru_ldmatrix_x4(a_hi + m * 16 * kGruHPitch * 2 + kk * 32, h0_, h1_, h2_, h3_);
        gru_ldmatrix_x4(a_lo + m * 16 * kGruHPitch * 2 + kk * 32, l0_, l1_, l2_, l3_);
        gru_mma(acc[m][0], h0_, h1_, h2_, h3_, b0, b1);
        gru_mma(acc[m][1], l0_, l1_, l2_, l3_, b0, b1);
        gru_mma(acc[m][2], h0_, h1_, h2_, h3_, wlo[2 * kk], wlo[2 * kk + 1]);
      }
    }
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      const int col = warp * 8 + 2 * tq;
      *reinterpret_cast<float2*>(&G[(m * 16 + gq) * kGruGPitch + col]) =
          make_float2(acc[m][0][0] + (acc[m][1][0] + acc[m][2][0]), acc[m][0][1] + (acc[m][1][1] + acc[m][2][1]));
      *reinterpret_cast<float2*>(&G[(m * 16 + gq + 8) * kGruGPitch + col]) =
          make_float2(acc[m][0][2] + (acc[m][1][2] + acc[m][2][2]), acc[m][0][3] + (acc[m][1][3] + acc[m][2][3]));
    }
    __syncthreads();
    cluster.barrier_arrive();                                       // this CTA has finished reading h(t)
    // ---- (2) gate math in fp32 for (sequence gs, units gu..gu+3) and the output row, while the barrier completes
    float hnew[4] = {0.f, 0.f, 0.f, 0.f};
    if (gate_thread) {
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (seq_ok) x = __ldg(reinterpret_cast<const float4*>(u + ((long long)(seq0 + gs) * steps + t) * kGruIn));
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = gu + e;
        const float* wr = Wi + j * kGruIn; const float* wz = Wi + (32 + j) * kGruIn; const float* wn = Wi + (64 + j) * kGruIn;
        const float ir = fmaf(wr[0], x.x, fmaf(wr[1], x.y, fmaf(wr[2], x.z, fmaf(wr[3], x.w, Bi[j]))));
        const float iz = fmaf(wz[0], x.x, fmaf(wz[1], x.y, fmaf(wz[2], x.z, fmaf(wz[3], x.w, Bi[32 + j]))));
        const float in_ = fmaf(wn[0], x.x, fmaf(wn[1], x.y, fmaf(wn[2], x.z, fmaf(wn[3], x.w, Bi[64 + j]))));
        const float r = gru_sigmoid(ir + G[gs * kGruGPitch + j] + Bh[j]);
        const float z = gru_sigmoid(iz + G[gs * kGruGPitch + 32 + j] + Bh[32 + j]);
        const float n = tanhf(fmaf(r, G[gs * kGruGPitch + 64 + j] + Bh[64 + j], in_));
        const float hold = Hm[gs * kGruUnits + j];
        hnew[e] = fmaf(z, hold - n, n);                             // (1 - z) n + z h
        Hm[gs * kGruUnits + j] = hnew[e];
      }
      if (seq_ok)
        *reinterpret_cast<float4*>(out + ((long long)(seq0 + gs) * steps + t) * kGruH + rank * kGruUnits + gu) = make_float4(hnew[0], hnew[1], hnew[2], hnew[3]);
    }
    cluster.barrier_wait();                                         // every CTA of the cluster has finished reading h(t)
    // ---- (3) push the hi/lo h16 slices of h(t+1) to all 16 CTAs
    if (gate_thread) {
      uint2 hi, lo;
      gru_split2(hnew[0], hnew[1], hi.x, lo.x);
      gru_split2(hnew[2], hnew[3], hi.y, lo.y);
      h16* dst = Hb + gs * kGruHPitch + rank * kGruUnits + gu;      // same offset in every CTA's shared memory
#pragma unroll
      for (int peer = 0; peer < kGruCluster; ++peer) {
        h16* rp = cluster.map_shared_rank(dst, peer);
        *reinterpret_cast<uint2*>(rp) = hi;
        *reinterpret_cast<uint2*>(rp + kPart) = lo;
      }
    }
    cluster.sync();      // pushes visible everywhere; nobody still reads this step's G
  }
}

}  // namespace lavb

using namespace lavb;

extern "C" int lavb_gru_h512(const float* d_u, const float* d_h0, const float* d_whh, const float* d_wih, const float* d_bih,
                             const float* d_bhh, float* d_out, int nseq, int steps, void* stream) {
  LAVB_CHECK_ARG(nseq >= 0 && steps >= 1, "gru_h512: bad shape");
  if (nseq == 0) return 0;
  LAVB_CUDA_OK(ensure_dyn_smem((const void*)gru_cluster_kernel, GruSmem::total));
  LAVB_CUDA_OK(cudaFuncSetAttribute(gru_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(ceil_div(nseq, kGruSeq) * kGruCluster);
  cfg.blockDim = dim3(kGruThreads);
  cfg.dynamicSmemBytes = GruSmem::total;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = kGruCluster; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  LAVB_CUDA_OK(cudaLaunchKernelEx(&cfg, gru_cluster_kernel, d_u, d_h0, d_whh, d_wih, d_bih, d_bhh, d_out, nseq, steps));
  return 0;
}
