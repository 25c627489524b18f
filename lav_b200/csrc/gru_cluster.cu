// Cluster-persistent plan GRU (lav_b200.heads.GRU_KERNEL; B200: roll-out of 192 sequences 1.06 -> 0.71 ms vs 100 cuDNN launch pairs):
// the plan GRU roll-out of UniPlanner / BEVPlanner (team_code_v2/models/uniplanner.py:227-259: nn.GRU(4, 512), 20 steps,
// 6*B sequences, called 5 times per tick) as ONE cluster-persistent kernel per call instead of 20 x (cuDNN GEMM + cell kernel)
// — 100 sequential launch pairs per tick, ~8 us each, are 30 us/frame of the round-1 pipeline.
//   * a thread-block cluster of 16 CTAs owns 32 sequences for all T steps; CTA r holds the 96 rows of W_hh that produce the
//     r/z/n gates of hidden units [32r, 32r+32) in shared memory (h16, 96 KB) for the whole roll-out;
//   * per step every CTA multiplies the full hidden state of its 32 sequences (h16 copy in shared memory, double buffered)
//     with its weight slice on the tensor cores (mma.sync m16n8k16, fp32 accumulate), applies the gate math in fp32 on its
//     32 units (fp32 master copy of h stays local), writes the step's output rows and PUSHES the h16 slice of h' into the
//     next-step buffer of all 16 CTAs through distributed shared memory; one cluster barrier per step.
// PyTorch gate order and formulas (r, z, n; n = tanh(W_in x + b_in + r * (W_hn h + b_hn)); h' = (1 - z) n + z h).
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace lavb {

constexpr int kGruH = 512, kGruSeq = 32, kGruUnits = 32, kGruCluster = 16, kGruCols = 96, kGruIn = 4;
constexpr int kGruWPitch = kGruH + 8, kGruHPitch = kGruH + 8;     // h16 elements; +8 keeps ldmatrix / fragment loads conflict-free
constexpr int kGruGPitch = kGruCols + 4;                           // fp32 gate pre-activations [seq][96]

struct GruSmem {
  static constexpr int w = 0;                                                   // [96][kGruWPitch] h16
  static constexpr int hb = w + kGruCols * kGruWPitch * 2;                      // [2][32][kGruHPitch] h16
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

__global__ void __launch_bounds__(256, 1) gru_cluster_kernel(const float* __restrict__ u, const float* __restrict__ h0,
                                                             const h16* __restrict__ whh, const float* __restrict__ wih,
                                                             const float* __restrict__ bih, const float* __restrict__ bhh,
                                                             float* __restrict__ out, int nseq, int steps) {
  extern __shared__ __align__(16) uint8_t gsm[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();                       // which 32 hidden units this CTA owns
  const int seq0 = ((int)blockIdx.x / kGruCluster) * kGruSeq;       // first sequence of this cluster
  h16* Ws = reinterpret_cast<h16*>(gsm + GruSmem::w);
  h16* Hb = reinterpret_cast<h16*>(gsm + GruSmem::hb);
  float* G = reinterpret_cast<float*>(gsm + GruSmem::g);
  float* Hm = reinterpret_cast<float*>(gsm + GruSmem::hm);
  float* Wi = reinterpret_cast<float*>(gsm + GruSmem::wih);
  float* Bi = reinterpret_cast<float*>(gsm + GruSmem::bih);
  float* Bh = reinterpret_cast<float*>(gsm + GruSmem::bhh);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gq = lane >> 2, tq = lane & 3;

  // ---- one-time staging: local row c = gate*32 + j  <->  global row gate*512 + 32*rank + j
  for (int i = tid; i < kGruCols * (kGruH / 8); i += 256) {
    const int c = i / (kGruH / 8), ch = i - c * (kGruH / 8);
    const int grow = (c >> 5) * kGruH + rank * kGruUnits + (c & 31);
    *reinterpret_cast<uint4*>(Ws + c * kGruWPitch + ch * 8) = __ldg(reinterpret_cast<const uint4*>(whh + (long long)grow * kGruH) + ch);
  }
  for (int i = tid; i < kGruCols; i += 256) {
    const int grow = (i >> 5) * kGruH + rank * kGruUnits + (i & 31);
#pragma unroll
    for (int k = 0; k < kGruIn; ++k) Wi[i * kGruIn + k] = __ldg(wih + grow * kGruIn + k);
    Bi[i] = __ldg(bih + grow);
    Bh[i] = __ldg(bhh + grow);
  }
  for (int i = tid; i < kGruSeq * kGruH; i += 256) {
    const int s = i / kGruH, k = i - s * kGruH;
    const float v = (seq0 + s < nseq) ? __ldg(h0 + (long long)(seq0 + s) * kGruH + k) : 0.f;
    Hb[s * kGruHPitch + k] = float2h16(v);
    const int j = k - rank * kGruUnits;
    if (j >= 0 && j < kGruUnits) Hm[s * kGruUnits + j] = v;
  }
  cluster.sync();                                                   // every CTA of the cluster is resident and initialised

  const int mt = warp & 1, ng = warp >> 1;                          // GEMM role: m-tile (16 sequences) x 3 n-tiles (24 gate columns)
  const int gs = tid >> 3, gu = (tid & 7) * 4;                      // gate role: sequence, first of 4 hidden units
  const bool seq_ok = seq0 + gs < nseq;
  for (int t = 0; t < steps; ++t) {
    const h16* hc = Hb + (t & 1) * kGruSeq * kGruHPitch;
    h16* hn = Hb + ((t & 1) ^ 1) * kGruSeq * kGruHPitch;
    // ---- (1) G[32 seq][96] = h (32 x 512) . Wslice^T on the tensor cores
    float acc[3][4];
#pragma unroll
    for (int j = 0; j < 3; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    const uint32_t a_base = (uint32_t)__cvta_generic_to_shared(hc + (mt * 16 + (lane & 15)) * kGruHPitch + (lane >> 4) * 8);
#pragma unroll 4
    for (int kk = 0; kk < kGruH / 16; ++kk) {
      uint32_t a0, a1, a2, a3;
      gru_ldmatrix_x4(a_base + kk * 32, a0, a1, a2, a3);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const h16* wp = Ws + ((ng * 3 + j) * 8 + gq) * kGruWPitch + kk * 16 + 2 * tq;
        gru_mma(acc[j], a0, a1, a2, a3, *reinterpret_cast<const uint32_t*>(wp), *reinterpret_cast<const uint32_t*>(wp + 8));
      }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int col = (ng * 3 + j) * 8 + 2 * tq;
      *reinterpret_cast<float2*>(&G[(mt * 16 + gq) * kGruGPitch + col]) = make_float2(acc[j][0], acc[j][1]);
      *reinterpret_cast<float2*>(&G[(mt * 16 + gq + 8) * kGruGPitch + col]) = make_float2(acc[j][2], acc[j][3]);
    }
    __syncthreads();
    // ---- (2) gate math in fp32 for (sequence gs, units gu..gu+3), output row, push of the h16 slice to all 16 CTAs
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (seq_ok) x = __ldg(reinterpret_cast<const float4*>(u + ((long long)(seq0 + gs) * steps + t) * kGruIn));
    float hnew[4];
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
      hnew[e] = fmaf(z, hold - n, n);                               // (1 - z) n + z h
      Hm[gs * kGruUnits + j] = hnew[e];
    }
    if (seq_ok)
      *reinterpret_cast<float4*>(out + ((long long)(seq0 + gs) * steps + t) * kGruH + rank * kGruUnits + gu) = make_float4(hnew[0], hnew[1], hnew[2], hnew[3]);
    {
      const h162 p0 = floats2h162(hnew[0], hnew[1]), p1 = floats2h162(hnew[2], hnew[3]);
      uint2 pk;
      pk.x = *reinterpret_cast<const uint32_t*>(&p0); pk.y = *reinterpret_cast<const uint32_t*>(&p1);
      h16* dst = hn + gs * kGruHPitch + rank * kGruUnits + gu;   // same offset in every CTA's shared memory
#pragma unroll
      for (int peer = 0; peer < kGruCluster; ++peer)
        *reinterpret_cast<uint2*>(cluster.map_shared_rank(dst, peer)) = pk;
    }
    cluster.sync();      // pushes visible everywhere; nobody still reads this step's h or G
  }
}

}  // namespace lavb

using namespace lavb;

extern "C" int lavb_gru_h512(const float* d_u, const float* d_h0, const void* d_whh_h16, const float* d_wih, const float* d_bih,
                             const float* d_bhh, float* d_out, int nseq, int steps, void* stream) {
  LAVB_CHECK_ARG(nseq >= 0 && steps >= 1, "gru_h512: bad shape");
  if (nseq == 0) return 0;
  LAVB_CUDA_OK(ensure_dyn_smem((const void*)gru_cluster_kernel, GruSmem::total));
  LAVB_CUDA_OK(cudaFuncSetAttribute(gru_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(ceil_div(nseq, kGruSeq) * kGruCluster);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = GruSmem::total;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = kGruCluster; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  LAVB_CUDA_OK(cudaLaunchKernelEx(&cfg, gru_cluster_kernel, d_u, d_h0, reinterpret_cast<const h16*>(d_whh_h16), d_wih, d_bih,
                                  d_bhh, d_out, nseq, steps));
  return 0;
}
