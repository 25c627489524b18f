// EXPERIMENTAL (round-2 work item, not on the default path; enabled by lav_b200.layers.USE_HALO):
// stride-1 tap-list convolution on tcgen05 where every input pixel is fetched ONCE per tile instead of once per tap.
//
// Why: conv_umma.cu loads one 16 KB A tile per (tap, 64-channel block); a 3x3 layer therefore pulls 9x the tile's input
// through a 4-stage smem ring and its narrow layers (N = 64 / 128) sit at ring-depth x latency, not at the tensor pipe
// (backbone 64->64 3x3: ~3 us per tile per SM for 0.6 us of MMAs, profiles/r01_kernels.md).  Here a tile is 16 rows x 8
// pixels and, per 64-channel block, ONE TMA box per distinct horizontal tap offset dx brings the column-shifted patch
// {64 ch, 8 px, 16 + 2hy rows} (hy = max |dy|).  Such a patch IS a canonical K-major SWIZZLE_128B operand of 8 x (16 + 2hy)
// rows — one 1024-byte core-matrix group per image row — so the A operand of tap (dy, dx) is the patch of that dx seen
// through a descriptor whose start is advanced by (hy + dy) groups: start addresses stay 1024-byte aligned, the group stride
// stays 1024 B, nothing but descriptor arithmetic changes between taps.  A 3x3 layer fetches 3 x 18 KB instead of 9 x 16 KB
// per block, a 3x1 layer 18 KB instead of 48 KB; 1xN layers gain nothing and are left to conv_umma.cu.
// Weights: all (tap, block) tiles resident in shared memory when they fit, otherwise streamed through their own ring.
// Epilogue, barriers and warp roles follow conv_umma.cu (8 epilogue warps, one CTA per SM, two TMEM accumulators).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>
#include "common.cuh"

namespace lavb {
namespace halo {

constexpr int kTileH = 16, kTileW = 8, kBlockM = 128, kBlockK = 64;
constexpr int kMaxStages = 8, kMaxTaps = 16, kEpiWarps = 8;

struct HaloArgs {
  int n, h, w, tiles_x, tiles_y, num_tiles;
  int cin, cout, cout_store, kchunks, ntaps, hy, ph, ndx;        // ph = 16 + 2hy patch rows; ndx distinct dx values
  int dxs[3], dxi[kMaxTaps];                                      // the distinct dx values; tap -> index into dxs
  int stages_a, stages_b, wres, tmem_cols;
  int out_cstride, out_coff, res_cstride, res_coff, pre_relu, post_relu;
  int dy[kMaxTaps], dx[kMaxTaps];
  __nv_bfloat16* out; const __nv_bfloat16* res;
  const float* bias; const float* scale; const float* shift;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (same encoding as conv_umma.cu; saddr is 1024 B aligned here)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
  const uint32_t lo = (saddr & 0x3FFFFu) >> 4;
  const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  return (uint64_t)lo | ((uint64_t)hi << 32);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
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
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}

__global__ void __launch_bounds__(64 + 32 * kEpiWarps, 1) conv_halo_umma_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                                               const __grid_constant__ CUtensorMap tmap_b,
                                                                               const __grid_constant__ HaloArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int patch1 = p.ph * 1024;                                 // one column-shifted patch: ph rows x 8 px x 128 B
  const int patch_bytes = p.ndx * patch1;                         // all patches of one 64-channel block = one ring stage
  const int b_bytes = p.cout * kBlockK * 2;                       // one (tap, block) weight tile: cout rows x 128 B
  const int nkb = p.ntaps * p.kchunks;
  const uint32_t a_ring = base;
  const uint32_t b_ring = a_ring + p.stages_a * patch_bytes;      // weight ring, or all nkb tiles when resident
  const uint32_t ctrl = b_ring + (p.wres ? nkb : p.stages_b) * b_bytes;
  const uint32_t afull = ctrl, aempty = ctrl + 8 * kMaxStages, bfull = ctrl + 16 * kMaxStages, bempty = ctrl + 24 * kMaxStages,
                 tfull = ctrl + 32 * kMaxStages, tempty = tfull + 16, wbar = tempty + 16, tmem_slot = wbar + 8;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_p = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));
  float* ep_bias = reinterpret_cast<float*>(gen + (tmem_slot - base) + 8);      // pre-activation bias (0 unless pre_relu)
  float* ep_st = ep_bias + 256;                                                 // interleaved (scale, shift')

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_b)) : "memory");
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(afull + 8 * s, 1); mbar_init(aempty + 8 * s, 1); mbar_init(bfull + 8 * s, 1); mbar_init(bempty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull + 8 * a, 1); mbar_init(tempty + 8 * a, 32 * kEpiWarps); }
    mbar_init(wbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int c = threadIdx.x; c < p.cout; c += blockDim.x) {
    // epi(a) = max(a + b_pre, lo) * s + t'.  Without the pre-ReLU the bias folds into the shift: (a + b) s + t = a s + (b s + t)
    const bool real = c < p.cout_store;
    const float b = (p.bias && real) ? __ldg(p.bias + c) : 0.f;
    const float sc = (p.scale && real) ? __ldg(p.scale + c) : 1.f;
    const float sh = (p.shift && real) ? __ldg(p.shift + c) : 0.f;
    ep_bias[c] = p.pre_relu ? b : 0.f;
    ep_st[2 * c] = sc;
    ep_st[2 * c + 1] = p.pre_relu ? sh : fmaf(b, sc, sh);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_p;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    if (lane == 0) {
      int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
      if (p.wres) {
        mbar_expect_tx(wbar, (uint32_t)(nkb * b_bytes));
        for (int kb = 0; kb < nkb; ++kb)                             // resident order: [kc][tap]
          tma_load_2d(b_ring + kb * b_bytes, &tmap_b, wbar, (kb / p.ntaps) * kBlockK, (kb % p.ntaps) * p.cout);
      }
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int img = tile / tiles_per_img, r = tile - img * tiles_per_img;
        const int y0 = (r / p.tiles_x) * kTileH - p.hy, x0 = (r % p.tiles_x) * kTileW;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(aempty + 8 * sa, pa ^ 1);
          mbar_expect_tx(afull + 8 * sa, (uint32_t)patch_bytes);
          for (int j = 0; j < p.ndx; ++j)
            tma_load_4d(a_ring + sa * patch_bytes + j * patch1, &tmap_a, afull + 8 * sa, kc * kBlockK, x0 + p.dxs[j], y0, img);
          if (++sa == p.stages_a) { sa = 0; pa ^= 1; }
          if (!p.wres)
            for (int t = 0; t < p.ntaps; ++t) {
              mbar_wait(bempty + 8 * sb, pb ^ 1);
              mbar_expect_tx(bfull + 8 * sb, (uint32_t)b_bytes);
              tma_load_2d(b_ring + sb * b_bytes, &tmap_b, bfull + 8 * sb, kc * kBlockK, t * p.cout);
              if (++sb == p.stages_b) { sb = 0; pb ^= 1; }
            }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.cout >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
      int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
      int acc = 0; uint32_t acc_phase = 0;
      if (p.wres) { mbar_wait(wbar, 0); tc_fence_after(); }
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(tempty + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.cout);
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(afull + 8 * sa, pa);
          tc_fence_after();
          const uint32_t patch = a_ring + sa * patch_bytes;
          for (int t = 0; t < p.ntaps; ++t) {
            uint32_t bt;
            if (p.wres) bt = b_ring + (kc * p.ntaps + t) * b_bytes;
            else { mbar_wait(bfull + 8 * sb, pb); tc_fence_after(); bt = b_ring + sb * b_bytes; }
            const uint32_t a_start = patch + (uint32_t)(p.dxi[t] * patch1 + (p.hy + p.dy[t]) * 1024);   // rows (hy+dy)*8 .. +128
            const uint64_t a_desc = make_sw128_desc(a_start), b_desc = make_sw128_desc(bt);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)                    // +32 B per K16 step inside the 128 B swizzle atom
              umma_bf16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kc | t | k) ? 1u : 0u);
            if (!p.wres) { umma_commit(bempty + 8 * sb); if (++sb == p.stages_b) { sb = 0; pb ^= 1; } }
          }
          umma_commit(aempty + 8 * sa);                               // the patch may be overwritten once these MMAs have read it
          if (++sa == p.stages_a) { sa = 0; pa ^= 1; }
        }
        umma_commit(tfull + 8 * acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    const int q = warp & 3;                        // TMEM lane quarter this warp may read (warp id % 4)
    const int half = (warp - 2) >> 2;              // the two warps of a quarter take alternate 32-column chunks
    const int row = q * 32 + lane;
    const int py = row / kTileW, px = row % kTileW;
    const float lo_pre = p.pre_relu ? 0.f : -INFINITY, lo_post = p.post_relu ? 0.f : -INFINITY;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int img = tile / tiles_per_img, r = tile - img * tiles_per_img;
      const int oy = (r / p.tiles_x) * kTileH + py, ox = (r % p.tiles_x) * kTileW + px;
      const bool valid = oy < p.h && ox < p.w;
      const long long pix = ((long long)img * p.h + oy) * p.w + ox;
      uint4 rr[4];
      auto load_res = [&](int c0) {
        const uint4* rp = reinterpret_cast<const uint4*>(p.res + pix * p.res_cstride + p.res_coff + c0);
#pragma unroll
        for (int j = 0; j < 4; ++j) rr[j] = __ldg(rp + j);
      };
      if (p.res && valid) load_res(half * 32);
      mbar_wait(tfull + 8 * acc, acc_phase);
      tc_fence_after();
      for (int c0 = half * 32; c0 < p.cout; c0 += 64) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.cout + c0), v);
        float f[32];
        const float4* st4 = reinterpret_cast<const float4*>(ep_st + 2 * c0);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 st = st4[j];
          f[2 * j] = fmaf(fmaxf(__uint_as_float(v[2 * j]) + ep_bias[c0 + 2 * j], lo_pre), st.x, st.y);
          f[2 * j + 1] = fmaf(fmaxf(__uint_as_float(v[2 * j + 1]) + ep_bias[c0 + 2 * j + 1], lo_pre), st.z, st.w);
        }
        if (valid) {
          if (p.res) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t wv[4] = {rr[j].x, rr[j].y, rr[j].z, rr[j].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 t2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wv[e]));
                f[j * 8 + e * 2] += t2.x; f[j * 8 + e * 2 + 1] += t2.y;
              }
            }
            if (c0 + 64 < p.cout) load_res(c0 + 64);
          }
          uint4* op = reinterpret_cast<uint4*>(p.out + pix * p.out_cstride + p.out_coff + c0);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (c0 + 8 * j < p.cout_store)
              op[j] = make_uint4(pack2(fmaxf(f[8 * j], lo_post), fmaxf(f[8 * j + 1], lo_post)), pack2(fmaxf(f[8 * j + 2], lo_post), fmaxf(f[8 * j + 3], lo_post)),
                                 pack2(fmaxf(f[8 * j + 4], lo_post), fmaxf(f[8 * j + 5], lo_post)), pack2(fmaxf(f[8 * j + 6], lo_post), fmaxf(f[8 * j + 7], lo_post)));
        }
      }
      tc_fence_before();
      mbar_arrive(tempty + 8 * acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

}  // namespace halo
}  // namespace lavb

using namespace lavb;
using namespace lavb::halo;

// Same descriptor as lavb_conv_umma, restricted to plain stride-1 convolutions with bf16 output.  Returns 4 (and sets the error
// text) when the layer is outside what this kernel covers or the halo patch would not fit / not pay: callers fall back to
// lavb_conv_umma.
extern "C" int lavb_conv_halo_umma(const lavb_conv_desc* d, void* stream) {
  LAVB_CHECK_ARG(d != nullptr, "conv_halo_umma: null descriptor");
  LAVB_CHECK_ARG(d->in_dtype == LAVB_BF16 && d->out_dtype == LAVB_BF16, "conv_halo_umma: bf16 in / bf16 out only");
  LAVB_CHECK_ARG(d->ntaps >= 1 && d->ntaps <= kMaxTaps, "conv_halo_umma: ntaps must be 1..16");
  LAVB_CHECK_ARG(d->cin % 64 == 0 && d->cin > 0, "conv_halo_umma: cin must be a multiple of 64 (got %d)", d->cin);
  LAVB_CHECK_ARG(d->cout % 8 == 0 && d->cout >= 8 && d->cout <= 256, "conv_halo_umma: cout must be 8..256, multiple of 8 (got %d)", d->cout);
  if (!(d->in_sy == 1 && d->in_sx == 1 && d->out_sy == 1 && d->out_sx == 1 && d->out_oy == 0 && d->out_ox == 0 &&
        d->hog == d->hout && d->wog == d->wout && d->hin == d->hout && d->win == d->wout && d->d2s_nout == 0 && !d->sigmoid)) {
    set_error("conv_halo_umma: plain stride-1 same-size convolutions only");
    return 4;
  }
  const int cout_mma = (d->cout + 31) / 32 * 32;
  LAVB_CHECK_ARG(d->res == nullptr || (cout_mma == d->cout && d->res_dtype == LAVB_BF16 && d->res_cstride % 8 == 0 && d->res_coff % 8 == 0),
                 "conv_halo_umma: residual must be bf16, 16 B aligned, cout %% 32 == 0");
  LAVB_CHECK_ARG(d->in_cstride % 8 == 0 && d->in_coff % 8 == 0 && d->in_coff + d->cin <= d->in_cstride, "conv_halo_umma: input slice misaligned");
  LAVB_CHECK_ARG(d->out_cstride % 8 == 0 && d->out_coff % 8 == 0 && d->out_coff + d->cout <= d->out_cstride, "conv_halo_umma: output slice misaligned");
  LAVB_CHECK_ARG((d->scale == nullptr) == (d->shift == nullptr), "conv_halo_umma: scale and shift come together");
  auto encode = get_encode();
  LAVB_CHECK_ARG(encode != nullptr, "conv_halo_umma: cuTensorMapEncodeTiled not available from the driver");

  HaloArgs a;
  memset(&a, 0, sizeof(a));
  for (int t = 0; t < d->ntaps; ++t) {
    a.dy[t] = d->dy[t]; a.dx[t] = d->dx[t];
    a.hy = max(a.hy, abs(d->dy[t]));
    int j = 0;
    while (j < a.ndx && a.dxs[j] != d->dx[t]) ++j;
    if (j == a.ndx) {
      if (a.ndx == 3) { set_error("conv_halo_umma: more than 3 distinct horizontal tap offsets"); return 4; }
      a.dxs[a.ndx++] = d->dx[t];
    }
    a.dxi[t] = j;
  }
  a.ph = kTileH + 2 * a.hy;
  const int patch_bytes = a.ndx * a.ph * 1024, b_bytes = cout_mma * kBlockK * 2;
  a.n = d->n; a.h = d->hout; a.w = d->wout;
  a.tiles_x = ceil_div(a.w, kTileW); a.tiles_y = ceil_div(a.h, kTileH);
  a.num_tiles = a.n * a.tiles_x * a.tiles_y;
  a.cin = d->cin; a.cout = cout_mma; a.cout_store = d->cout; a.kchunks = d->cin / kBlockK; a.ntaps = d->ntaps;
  const int nkb = a.ntaps * a.kchunks;
  // the patch must be a legal TMA box and cheaper than the per-tap tiles it replaces
  if (a.ph > 256 || patch_bytes * 5 > d->ntaps * 16384 * 4) {
    set_error("conv_halo_umma: %d patches of %d rows do not pay for %d taps", a.ndx, a.ph, d->ntaps);
    return 4;
  }
  const int budget = 200 * 1024;
  if ((long long)nkb * b_bytes + 2 * patch_bytes <= budget) {
    a.wres = 1; a.stages_b = 0;
    a.stages_a = min(kMaxStages, (budget - nkb * b_bytes) / patch_bytes);
  } else {
    a.wres = 0;
    a.stages_a = 2;
    a.stages_b = min(kMaxStages, (budget - 2 * patch_bytes) / b_bytes);
    if (a.stages_b < 2) { set_error("conv_halo_umma: operands do not fit in shared memory"); return 4; }
    if (budget - a.stages_b * b_bytes >= 3 * patch_bytes) a.stages_a = 3;
  }
  int cols = 32;
  while (cols < 2 * cout_mma) cols <<= 1;
  a.tmem_cols = cols;
  a.out_cstride = d->out_cstride; a.out_coff = d->out_coff; a.res_cstride = d->res_cstride; a.res_coff = d->res_coff;
  a.pre_relu = d->pre_relu; a.post_relu = d->post_relu;
  a.out = reinterpret_cast<__nv_bfloat16*>(d->out); a.res = reinterpret_cast<const __nv_bfloat16*>(d->res);
  a.bias = d->bias; a.scale = d->scale; a.shift = d->shift;
  if (a.num_tiles == 0) return 0;

  CUtensorMap tmap_a, tmap_b;
  {
    const __nv_bfloat16* in = reinterpret_cast<const __nv_bfloat16*>(d->in) + d->in_coff;
    cuuint64_t dims[4] = {(cuuint64_t)d->cin, (cuuint64_t)d->win, (cuuint64_t)d->hin, (cuuint64_t)d->n};
    cuuint64_t strides[3] = {(cuuint64_t)d->in_cstride * 2, (cuuint64_t)d->win * d->in_cstride * 2,
                             (cuuint64_t)d->hin * d->win * d->in_cstride * 2};
    cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)kTileW, (cuuint32_t)a.ph, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&tmap_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(in), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    LAVB_CHECK_ARG(r == CUDA_SUCCESS, "conv_halo_umma: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)d->cin, (cuuint64_t)d->ntaps * cout_mma};
    cuuint64_t strides[1] = {(cuuint64_t)d->cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)cout_mma};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tmap_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<float*>(d->w), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    LAVB_CHECK_ARG(r == CUDA_SUCCESS, "conv_halo_umma: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
  }
  const size_t smem = (size_t)a.stages_a * patch_bytes + (size_t)(a.wres ? nkb : a.stages_b) * b_bytes + 1024 /*align*/ +
                      32 * kMaxStages + 64 + 3 * 256 * sizeof(float);
  static bool configured = false;
  if (!configured) {
    LAVB_CUDA_OK(cudaFuncSetAttribute(conv_halo_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  const int grid = min(a.num_tiles, kNumSMs);
  conv_halo_umma_kernel<<<grid, 64 + 32 * kEpiWarps, smem, (cudaStream_t)stream>>>(tmap_a, tmap_b, a);
  LAVB_LAUNCH_OK();
  return 0;
}
