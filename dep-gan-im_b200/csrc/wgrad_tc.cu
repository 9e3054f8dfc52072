// tcgen05 weight-gradient kernel:  dW[dy][dx][ci][co] += sum_{n,h,w} x[n,h+dy-p,w+dx-p,ci] * b[n,h,w,co]
// (the wgrad half of Keras' K.gradients for Conv2D, TG:549/568/594; also the gradient-penalty wgrad on JVP
// activations).  bf16 operands, fp32 accumulation in tensor memory, one fp32 red.add per element per CTA at the end.
//
// GEMM view: K = pixels, N = co, M = (dx, ci).  Both operands are "MN-major" (the contiguous NHWC channel axis is
// the M / N axis, pixels are the K axis), which tcgen05 reads natively from the same [pixel][channel-chunk] shared
// memory tiles TMA produces:
//   * x: per stage ONE halo box [R+ks-1 rows][16+ks-1 cols][C channels], C in {16,32,64} = one swizzle span.
//     Pixel pitch == swizzle span, so in the canonical MN-major layout "next M block" can be made "next pixel"
//     (LBO = span): one M=128 MMA covers 128/C horizontally adjacent taps dx at once -- the Toeplitz structure of the
//     convolution is expressed in the descriptor, nothing is replicated in memory.  dy and the remaining dx groups
//     are further MMAs on shifted start addresses of the same tile.
//     dx groups beyond the first (C = 64, or C = 32 with ks = 5) are further MMAs on shifted start addresses.
//   * b: ONE box [R+ks-1 rows][16 cols][nb channels], nb <= 64 = one swizzle span, rows h0-pad .. (zero outside the
//     image).  The vertical taps are stacked on N the same way: "next N block" is made "next tile row" (LBO = row
//     pitch), so N = ks * nb and accumulator column block j holds tap dy = ks-1-j
//     (sum_p x[p + (dy,dx)] b[p] = sum_q x[q + dx] b[q - dy*W]).
//   * K = 16 pixels per MMA = one tile row; one row costs ceil(ks / (128/C)) MMAs of M = 128, N = ks * nb.
// Persistent CTAs split the pixel range; grid.y = input-channel chunk, grid.z = output-channel block.  Accumulators:
// ceil(ks / (128/C)) dx groups x (ks * nb) columns in TMEM.
#include <cuda.h>

#include <mutex>

#include "common.cuh"

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_w = nullptr;  // driver entry point: process-global
DgPerDevice g_dev_w;                 // shared-memory opt-in + SM count: per device
thread_local int g_sms_w = 148;      // SM count of the device of this thread's last wgrad_tc_init()

constexpr int WG_THREADS = 192;  // warp 0 producer, warp 1 MMA, warps 2..5 final flush
constexpr int R = 8;             // tile rows per stage (128 pixels)

struct WgGeom {
  int tiles_w, tiles_h;
  int C;          // channels per chunk (16/32/64)
  int nchunk0;    // chunks taken from x0 (the rest from x1)
  int tpm;        // dx taps stacked in one M=128 MMA
  int G;          // dx groups
  int nblk;       // output channels per CTA = one swizzle span (<= 64)
  int nspan;      // == nblk
  int S;          // pipeline stages
  uint32_t x_bytes, b_bytes, stage_bytes, x_tx, b_tx;
  uint32_t lay_x, lay_b;  // UMMA layout codes
  int tmem_cols;
  int cin_total, cout_total;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// KS = kernel size, C = input channels per chunk (one swizzle span): taps per MMA, tap groups and every descriptor
// offset of the MMA stream are compile-time, so the issuing warp's loop is a run of UTCHMMA with immediate offsets.
template <int KS, int C>
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX0,
                                                                 const __grid_constant__ CUtensorMap tmX1,
                                                                 const __grid_constant__ CUtensorMap tmB, float* dw,
                                                                 float alpha, int N, const WgGeom g, float* csum) {
  constexpr int PAD = KS / 2, HT = 16 + KS - 1;
  constexpr int SPAN = C * 2;       // bytes per pixel row of the x tile
  constexpr int TPM = 128 / C;      // dx taps stacked in one M=128 MMA
  constexpr int G = (KS + TPM - 1) / TPM;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bar_base = base + g.S * g.stage_bytes;
  const uint32_t full = bar_base, empty = full + 8 * g.S, accFull = empty + 8 * g.S, tmem_slot = accFull + 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  // channel sums of dy (bias gradient) ride along: the four flush warps, idle during the main loop, add up the
  // central rows of every b tile (CTAs of the first input-channel chunk only; every chunk sees the same dy)
  const bool do_cs = csum != nullptr && blockIdx.y == 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < g.S; ++i) { mbar_init(full + 8 * i, 1); mbar_init(empty + 8 * i, do_cs ? 5 : 1); }
    mbar_init(accFull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(g.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // work: half-tiles (R rows x 16 cols) of all slices, split contiguously over gridDim.x
  const int halves_per_img = g.tiles_w * g.tiles_h * 2;
  const long long items = (long long)halves_per_img * N;
  const long long per = (items + gridDim.x - 1) / gridDim.x;
  const long long it0 = (long long)blockIdx.x * per;
  const long long it1 = it0 + per < items ? it0 + per : items;
  const int chunk = blockIdx.y;
  const bool from0 = chunk < g.nchunk0;
  const int c_src = (from0 ? chunk : chunk - g.nchunk0) * C;                 // channel offset inside its source
  const int c_glob = from0 ? chunk * C : g.nchunk0 * C + (chunk - g.nchunk0) * C;  // ... inside the weight
  const int co0 = blockIdx.z * g.nblk;

  if (warp == 0) {
    int s = 0;
    uint32_t ph = 0;
    for (long long it = it0; it < it1; ++it) {
      const int half = (int)(it % 2);
      const long long t = it / 2;
      const int tw = (int)(t % g.tiles_w), th = (int)((t / g.tiles_w) % g.tiles_h), n = (int)(t / (g.tiles_w * g.tiles_h));
      const int w0 = tw * 16, h0 = th * 16 + half * R;
      mbar_wait(empty + 8 * s, ph ^ 1u);
      if (elect_one()) {
        const uint32_t st = base + s * g.stage_bytes;
        mbar_expect_tx(full + 8 * s, g.x_tx + g.b_tx);
        tma_load_4d(st, from0 ? &tmX0 : &tmX1, full + 8 * s, c_src, w0 - PAD, h0, n);
        tma_load_4d(st + g.x_bytes, &tmB, full + 8 * s, co0, w0, h0 - PAD, n);
      }
      __syncwarp();
      if (++s == g.S) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    // instruction descriptor: D=f32, A=B=bf16, A and B MN-major, M=128, N=nblk
    const uint32_t ncols = (uint32_t)(KS * g.nblk);  // N = vertical taps x output channels
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                           ((ncols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    // A: M blocks (C channels) step by one pixel (LBO = span); K groups of 8 pixels step by 8 pixels (SBO = 8*span)
    const uint32_t hiA = ((uint32_t)(8 * SPAN) >> 4) | (1u << 14) | (g.lay_x << 29);
    constexpr uint32_t loA_lbo = ((uint32_t)SPAN >> 4) << 16;
    // B: N blocks (nb channels) step by one tile row (LBO = 16 pixels); K groups of 8 pixels step by 8 pixels
    const uint32_t bspan = g.nspan * 2;
    const uint32_t hiB = ((uint32_t)(8 * bspan) >> 4) | (1u << 14) | (g.lay_b << 29);
    const uint32_t loB_lbo = ((uint32_t)(16 * bspan) >> 4) << 16;
    int s = 0;
    uint32_t ph = 0;
    uint32_t first = 1;
    uint32_t ready = 0;  // the next stage's barrier was already seen complete (probed mid-stage, see below)
    for (long long it = it0; it < it1; ++it) {
      if (!ready) mbar_wait(full + 8 * s, ph);
      tc_fence_after();
      const int sn = s + 1 == g.S ? 0 : s + 1;
      const uint32_t phn = s + 1 == g.S ? ph ^ 1u : ph;
      uint32_t probe = 0;
      const uint32_t st = base + s * g.stage_bytes;
      const uint32_t xa = ((st & 0x3FFFFu) >> 4), ba = (((st + g.x_bytes) & 0x3FFFFu) >> 4);
      const uint32_t b_row = (uint32_t)((16 * bspan) >> 4);  // one tile row of b in descriptor units
      if (elect_one()) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          // The MMA queue is only a couple of instructions deep, so a barrier round trip at the stage boundary idles
          // the tensor pipe: probe the next stage's barrier while this stage's MMAs are still being fed.
          if (r == R - 2) probe = mbar_try_wait(full + 8 * sn, phn);
          const uint32_t b_lo = (ba + (uint32_t)r * b_row) | loB_lbo;
          const uint32_t acc = r == 0 ? (first ^ 1u) : 1u;
#pragma unroll
          for (int gi = 0; gi < G; ++gi) {
            const uint32_t a_lo = (xa + (uint32_t)(((r * HT + gi * TPM) * SPAN) >> 4)) | loA_lbo;
            tc_mma(tmem_base + (uint32_t)gi * ncols, ((uint64_t)hiA << 32) | a_lo, ((uint64_t)hiB << 32) | b_lo, idesc,
                   acc);
          }
        }
        tc_commit(empty + 8 * s);
      }
      ready = __any_sync(0xffffffffu, probe != 0) ? 1u : 0u;
      first = 0;
      if (++s == g.S) { s = 0; ph ^= 1u; }
    }
    if (elect_one()) tc_commit(accFull);
    __syncwarp();
  } else {
    if (do_cs) {
      const int t = threadIdx.x - 64;                 // pixel of the R x 16 central block of the b tile
      const uint32_t bspan = g.nspan * 2, units = bspan / 16;
      const uint32_t po = (uint32_t)((PAD + (t >> 4)) * 16 + (t & 15)) * bspan;
      const uint32_t pxor = (po >> 7) & (units - 1);
      float acc[8][8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[u][k] = 0.f;
      int s = 0;
      uint32_t ph = 0;
      for (long long it = it0; it < it1; ++it) {
        mbar_wait(full + 8 * s, ph);
        const uint8_t* bt = smem_raw + (base + s * g.stage_bytes + g.x_bytes - raw);
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if ((uint32_t)u < units) {
            const uint4 q = *reinterpret_cast<const uint4*>(bt + po + ((((uint32_t)u) ^ pxor) << 4));
            const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              acc[u][2 * k] += __uint_as_float(w4[k] << 16);
              acc[u][2 * k + 1] += __uint_as_float(w4[k] & 0xFFFF0000u);
            }
          }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + 8 * s);
        if (++s == g.S) { s = 0; ph ^= 1u; }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if ((uint32_t)u < units) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float v = acc[u][k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) atomicAdd(csum + co0 + u * 8 + k, v);
          }
        }
    }
    if (it1 > it0) {
    // ===== final flush: TMEM -> red.add into dW[dy][dx][ci][co] =====
    mbar_wait(accFull, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int m = q * 32 + lane;        // accumulator row = (dx_local, ci)
    const int dxl = m / C, ci = m - dxl * C;
    for (int gi = 0; gi < G; ++gi) {
      const int dx = gi * TPM + dxl;
      for (int jb = 0; jb < KS; ++jb) {
        const int dy = KS - 1 - jb;  // column block jb pairs x row r with b row r - pad + jb
        for (int j = 0; j < g.nblk / 16; ++j) {
          float v[16];
          tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(gi * KS * g.nblk + jb * g.nblk + j * 16), v);
          if (dx < KS) {
            float* dst = dw + ((size_t)(dy * KS + dx) * g.cin_total + c_glob + ci) * g.cout_total + co0 + j * 16;
#pragma unroll
            for (int i = 0; i < 16; i += 4)  // 16-byte vector reductions (dst is 64-byte aligned)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(alpha * v[i]),
                           "f"(alpha * v[i + 1]), "f"(alpha * v[i + 2]), "f"(alpha * v[i + 3])
                           : "memory");
          }
        }
      }
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(g.tmem_cols) : "memory");
  }
}

CUtensorMapSwizzle swz(int chans) {
  return chans == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : chans == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}
uint32_t lay_code(int chans) { return chans == 64 ? 2u : chans == 32 ? 4u : 6u; }

int make_map(CUtensorMap* tm, const void* p, int C, int W, int H, int N, int box_c, int box_w, int box_h) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = g_encode_w(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p), dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swz(box_c), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    depgan_set_error("cuTensorMapEncodeTiled(wgrad) failed: " + std::to_string((int)r));
    return -1;
  }
  return 0;
}

uint32_t round1024(uint32_t x) { return (x + 1023u) & ~1023u; }

bool plan_w(const WgradArgs& a, WgGeom* g, uint32_t* smem) {
  if (a.x_dt != DT_BF16 || a.dy_dt != DT_BF16) return false;
  if (a.ks != 1 && a.ks != 3 && a.ks != 5) return false;
  if (a.H % 16 || a.W % 16 || a.H < 16 || a.W < 16) return false;
  if (a.C0 % 16 || a.C1 % 16 || a.C0 < 16 || a.Cout % 16 || a.Cout < 16) return false;
  int C = 64;
  while (C > 16 && (a.C0 % C || a.C1 % C)) C /= 2;
  const int tpm = 128 / C;
  const int G = (a.ks + tpm - 1) / tpm;
  // output-channel block per CTA: one swizzle span, with N = ks * nblk <= 256 and G * N <= 512 TMEM columns
  int nblk = 64;
  while (nblk > 16 && (a.Cout % nblk || a.ks * nblk > 256 || G * a.ks * nblk > 512)) nblk /= 2;
  if (a.Cout % nblk || a.ks * nblk > 256 || G * a.ks * nblk > 512) return false;
  const int nspan = nblk;
  const int ht = 16 + a.ks - 1;
  g->tiles_w = a.W / 16; g->tiles_h = a.H / 16;
  g->C = C; g->nchunk0 = a.C0 / C; g->tpm = tpm; g->G = G; g->nblk = nblk; g->nspan = nspan;
  g->x_tx = (uint32_t)R * ht * C * 2;
  g->b_tx = (uint32_t)(R + a.ks - 1) * 16 * nblk * 2;
  // the last dx group may read up to tpm-1 pixels past the halo row end of the last row: keep slack inside the stage
  g->x_bytes = round1024(g->x_tx + 128 * 2 * 8);
  g->b_bytes = round1024(g->b_tx);
  g->stage_bytes = g->x_bytes + g->b_bytes;
  int S = (int)((200u * 1024u) / g->stage_bytes);
  if (S > 6) S = 6;
  if (S < 2) return false;
  g->S = S;
  g->lay_x = lay_code(C); g->lay_b = lay_code(nspan);
  int cols = 32;
  while (cols < a.ks * G * nblk) cols *= 2;  // G dx groups x (ks * nblk) columns
  g->tmem_cols = cols;
  g->cin_total = a.C0 + a.C1; g->cout_total = a.Cout;
  *smem = 1024 + S * g->stage_bytes + 8 * (2 * S + 1) + 64;
  return true;
}

}  // namespace

bool wgrad_tc_supported(const WgradArgs& a) {
  WgGeom g;
  uint32_t smem;
  return plan_w(a, &g, &smem);
}

int wgrad_tc_init() {
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  if (!g_encode_w) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    DG_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) {
      depgan_set_error("cuTensorMapEncodeTiled entry point not available");
      return -1;
    }
    g_encode_w = reinterpret_cast<EncodeTiledFn>(fn);
  }
  int dev = 0;
  bool first = false;
  DG_TRY(dg_device_enter(g_dev_w, &dev, &first));
  if (first) {  // function attributes are per device
#define DG_WG_ATTR(KS_, C_) \
  DG_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<KS_, C_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    DG_WG_ATTR(1, 16) DG_WG_ATTR(1, 32) DG_WG_ATTR(1, 64)
    DG_WG_ATTR(3, 16) DG_WG_ATTR(3, 32) DG_WG_ATTR(3, 64)
    DG_WG_ATTR(5, 16) DG_WG_ATTR(5, 32) DG_WG_ATTR(5, 64)
#undef DG_WG_ATTR
    dg_device_mark(g_dev_w, dev);
  }
  if (dev < 64) g_sms_w = g_dev_w.sms[dev];
  else DG_CHECK_CUDA(cudaDeviceGetAttribute(&g_sms_w, cudaDevAttrMultiProcessorCount, dev));
  return 0;
}

int conv_wgrad_tc(const WgradArgs& a, cudaStream_t st) {
  if (a.N <= 0) return 0;
  DG_TRY(wgrad_tc_init());
  WgGeom g;
  uint32_t smem;
  DG_REQUIRE(plan_w(a, &g, &smem), "conv_wgrad_tc: unsupported shape");
  const int ht = 16 + a.ks - 1;
  CUtensorMap tmX0, tmX1, tmB;
  DG_TRY(make_map(&tmX0, a.x0, a.C0, a.W, a.H, a.N, g.C, ht, R));
  if (a.C1 > 0) DG_TRY(make_map(&tmX1, a.x1, a.C1, a.W, a.H, a.N, g.C, ht, R));
  else tmX1 = tmX0;
  DG_TRY(make_map(&tmB, a.dy, a.Cout, a.W, a.H, a.N, g.nspan, 16, R + a.ks - 1));
  const int chunks = (a.C0 + a.C1) / g.C, coblocks = a.Cout / g.nblk;
  const long long items = (long long)g.tiles_w * g.tiles_h * 2 * a.N;
  long long gx = g_sms_w / (chunks * coblocks);
  if (gx < 1) gx = 1;
  if (gx > items) gx = items;
  dim3 grid((unsigned)gx, chunks, coblocks);
#define DG_WG_CASE(KS_, C_) \
  case KS_ * 100 + C_: wgrad_tc_kernel<KS_, C_><<<grid, WG_THREADS, smem, st>>>(tmX0, tmX1, tmB, a.dw, a.alpha, a.N, g, a.csum); break;
  switch (a.ks * 100 + g.C) {
    DG_WG_CASE(1, 16) DG_WG_CASE(1, 32) DG_WG_CASE(1, 64)
    DG_WG_CASE(3, 16) DG_WG_CASE(3, 32) DG_WG_CASE(3, 64)
    DG_WG_CASE(5, 16) DG_WG_CASE(5, 32) DG_WG_CASE(5, 64)
    default: depgan_set_error("conv_wgrad_tc: no kernel for this (ks, C)"); return -2;
  }
#undef DG_WG_CASE
  DG_LAUNCH_CHECK();
  return 0;
}
