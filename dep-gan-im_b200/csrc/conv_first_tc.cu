// First-layer convolutions on the tensor cores (conv2d_dis_0a: 1 -> 16, 5x5, TG:322; conv2d_gen_0: nicg -> 32, 3x3,
// TG:397): fp32 image with 1-2 channels, bf16 output, the same epilogue as the other convolutions (folded bias/BN,
// ReLU, the ReLU-pattern mask of the JVP pass).
//
// K = taps x channels is tiny (9 .. 25), so the layer is a [pixels] x [K] x [Cout] GEMM whose A operand is an explicit
// im2col built on the fly: a CTA stages the fp32 halo tile of a 16x16 pixel tile in shared memory, every thread
// (= one pixel = one accumulator row) writes its K neighbourhood values as ONE K-major swizzled row, and four to eight
// tcgen05.mma produce the tile.  The image is NOT rounded to bf16: each value is split x = hi + lo into two bf16
// halves that occupy K and K more columns of the row (the weights are repeated), so the products carry 16 mantissa
// bits of the image like the fp32 CUDA-core kernel this replaces; the weights are bf16 as in every other layer of a
// bf16 network.  400 fp32 FMAs per pixel become ~100 instructions per pixel and the layer turns HBM-bound.
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
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

// x = hi + lo with hi = bf16(x), lo = bf16(x - hi); returned as one word (hi in the low half, lo in the high half)
__device__ __forceinline__ uint32_t split_bf16(float x) {
  const bf16 hi = __float2bfloat16_rn(x);
  const bf16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
  return (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
}

// KP = padded K of the split operand (2 * KS*KS*CIN rounded up to 32 or 64 = one 64 / 128-byte swizzle span)
template <int KS, int CIN>
struct FirstGeom {
  static constexpr int K = KS * KS * CIN;
  static constexpr int KP = 2 * K <= 32 ? 32 : 64;
  static constexpr int ROWB = KP * 2;                 // bytes per operand row
  static constexpr int UNITS = ROWB / 16;             // 16-byte units per row
  static constexpr uint32_t LAYOUT = KP == 64 ? 2u : 4u;  // UMMA layout code: SW128 / SW64
  static constexpr int HT = 16 + KS - 1;
};

template <int KS, int CIN, int COUT>
__global__ void __launch_bounds__(256) conv_first_tc_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ scale,
                                                            const float* __restrict__ shift, bf16* __restrict__ out,
                                                            const bf16* __restrict__ mask, int N, int H, int W,
                                                            int relu) {
  typedef FirstGeom<KS, CIN> GEO;
  constexpr int K = GEO::K, KP = GEO::KP, ROWB = GEO::ROWB, UNITS = GEO::UNITS, HT = GEO::HT, PAD = KS / 2;
  static_assert(2 * K <= KP, "split operand does not fit one swizzle span");
  static_assert(COUT == 16 || COUT == 32, "COUT");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  // [A: 256 rows x ROWB][B: COUT rows x ROWB, padded to 1024][halo fp32][scale, shift][barrier, tmem slot]
  constexpr uint32_t A_BYTES = 256 * ROWB, B_BYTES = ((COUT * ROWB + 1023) / 1024) * 1024;
  uint8_t* sA = gen;
  uint8_t* sB = gen + A_BYTES;
  uint32_t* s_halo = reinterpret_cast<uint32_t*>(gen + A_BYTES + B_BYTES);  // split (hi | lo << 16) image values
  float* s_sc = reinterpret_cast<float*>(s_halo + HT * HT * CIN);
  float* s_sh = s_sc + COUT;
  const uint32_t bar = base + A_BYTES + B_BYTES + (uint32_t)((HT * HT * CIN + 2 * COUT) * 4 + 15) / 16 * 16;
  const uint32_t tmem_slot = bar + 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(2 * COUT)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // B operand: row co = [w(k = 0..K-1), w(k = 0..K-1), 0 ...] in bf16, K-major with the hardware's XOR swizzle
  for (int i = tid; i < COUT * UNITS; i += 256) {
    const int co = i / UNITS, u = i - co * UNITS;
    __align__(16) bf16 v8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = u * 8 + j;
      const int kk = k < K ? k : k - K;
      v8[j] = k < 2 * K ? __float2bfloat16_rn(w[(size_t)kk * COUT + co]) : __float2bfloat16_rn(0.f);
    }
    const uint32_t off = (uint32_t)co * ROWB;
    const uint32_t phys = off + ((((uint32_t)u) ^ ((off >> 7) & (UNITS - 1))) << 4);
    *reinterpret_cast<uint4*>(sB + phys) = *reinterpret_cast<const uint4*>(v8);
  }
  if (tid < COUT) {
    s_sc[tid] = scale ? scale[tid] : 1.f;
    s_sh[tid] = shift ? shift[tid] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // this thread's pixel = accumulator row: strip (TMEM column block) warp / 4, lane quarter warp % 4
  const int py = tid >> 4, px = tid & 15;
  const uint32_t a_off = (uint32_t)tid * ROWB;
  const uint32_t a_xor = (a_off >> 7) & (UNITS - 1);
  const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * COUT);
  // instruction descriptor: D = f32, A = B = bf16, K-major, M = 128, N = COUT
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(COUT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t hi_desc = ((uint32_t)(8 * ROWB) >> 4) | (1u << 14) | (GEO::LAYOUT << 29);
  const uint32_t a_lo0 = (((base)&0x3FFFFu) >> 4) | (1u << 16);
  const uint32_t b_lo0 = (((base + A_BYTES) & 0x3FFFFu) >> 4) | (1u << 16);

  const int tiles_w = W / 16, tiles_h = H / 16;
  const int n_tiles = tiles_w * tiles_h * N;
  uint32_t phase = 0;
#ifdef DG_FIRST_DBG
  long long c_ph[6] = {0, 0, 0, 0, 0, 0}, c_t = clock64();
  int n_done = 0;
#define DG_PH(i_) do { const long long n_ = clock64(); c_ph[i_] += n_ - c_t; c_t = n_; } while (0)
#else
#define DG_PH(i_) do {} while (0)
#endif
  constexpr int NPRE = (HT * HT * CIN + 255) / 256;
  float pre[NPRE];
  auto fetch_halo = [&](int t_) {
    const int tw_ = t_ % tiles_w, th_ = (t_ / tiles_w) % tiles_h, n_ = t_ / (tiles_w * tiles_h);
#pragma unroll
    for (int j = 0; j < NPRE; ++j) {
      const int i = tid + 256 * j;
      const int r = i / (HT * CIN), cc = i - r * (HT * CIN);
      const int gy = th_ * 16 - PAD + r, gx = tw_ * 16 - PAD + cc / CIN, ci = cc % CIN;
      pre[j] = (i < HT * HT * CIN && gy >= 0 && gy < H && gx >= 0 && gx < W)
                   ? __ldg(x + (((size_t)n_ * H + gy) * W + gx) * CIN + ci) : 0.f;
    }
  };
  if ((int)blockIdx.x < n_tiles) fetch_halo(blockIdx.x);
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int tw = t % tiles_w, th = (t / tiles_w) % tiles_h, n = t / (tiles_w * tiles_h);
    const int w0 = tw * 16, h0 = th * 16;
    // ---- halo tile (fp32, zero outside the image): fetched into registers one tile ahead ----
#pragma unroll
    for (int j = 0; j < NPRE; ++j)
      if (tid + 256 * j < HT * HT * CIN) s_halo[tid + 256 * j] = split_bf16(pre[j]);
    __syncthreads();
    DG_PH(0);
    // ---- im2col row of this pixel: [hi(0..K-1), lo(0..K-1), 0 ...]; the halves were split once per halo value ----
    {
      uint32_t wv[K + 1];
#pragma unroll
      for (int dy = 0; dy < KS; ++dy)
#pragma unroll
        for (int dx = 0; dx < KS; ++dx)
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci)
            wv[(dy * KS + dx) * CIN + ci] = s_halo[((py + dy) * HT + px + dx) * CIN + ci];
      wv[K] = 0u;
      uint32_t row[KP / 2];  // bf16 pairs
#pragma unroll
      for (int j = 0; j < KP / 2; ++j) {
        // element e = 2j, 2j+1 of the row: e < K -> hi(e); K <= e < 2K -> lo(e - K); else 0
        uint32_t pr = 0u;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int e = 2 * j + h;
          uint32_t half = 0u;
          if (e < K) half = wv[e] & 0xFFFFu;
          else if (e < 2 * K) half = wv[e - K] >> 16;
          pr |= half << (16 * h);
        }
        row[j] = pr;
      }
#pragma unroll
      for (int u = 0; u < UNITS; ++u)
        *reinterpret_cast<uint4*>(sA + a_off + ((((uint32_t)u) ^ a_xor) << 4)) =
            make_uint4(row[4 * u], row[4 * u + 1], row[4 * u + 2], row[4 * u + 3]);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (t + (int)gridDim.x < n_tiles) fetch_halo(t + gridDim.x);  // lands while the MMAs and the epilogue run
    DG_PH(1);
    __syncthreads();
    DG_PH(2);
    // ---- MMAs: two strips of 128 pixels, KP/16 K-steps each ----
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int k = 0; k < KP / 16; ++k)
          tc_mma(tmem_base + (uint32_t)(s * COUT),
                 ((uint64_t)hi_desc << 32) | (a_lo0 + (uint32_t)((s * 128 * ROWB) >> 4) + 2u * k),
                 ((uint64_t)hi_desc << 32) | (b_lo0 + 2u * k), idesc, k != 0);
      tc_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1u;
    tc_fence_after();
    DG_PH(3);
    // ---- epilogue: folded bias / BN, ReLU, JVP mask; NHWC bf16 (adjacent lanes write adjacent pixels) ----
    const size_t o = (((size_t)n * H + h0 + py) * W + w0 + px) * COUT;
#pragma unroll
    for (int g = 0; g < COUT / 16; ++g) {
      float v[16];
      tc_ld16(t_row + (uint32_t)(g * 16), v);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        v[i] = fmaf(v[i], s_sc[g * 16 + i], s_sh[g * 16 + i]);
        if (relu) v[i] = fmaxf(v[i], 0.f);
      }
      if (mask) {
        const uint4 m0 = __ldg(reinterpret_cast<const uint4*>(mask + o + g * 16));
        const uint4 m1 = __ldg(reinterpret_cast<const uint4*>(mask + o + g * 16) + 1);
        const uint32_t mw[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float mv = __uint_as_float((i & 1) ? (mw[i >> 1] & 0xFFFF0000u) : (mw[i >> 1] << 16));
          v[i] = mv > 0.f ? v[i] : 0.f;
        }
      }
      uint4 q[2];
      __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(q);
#pragma unroll
      for (int i = 0; i < 8; ++i) hp[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      uint4* dst = reinterpret_cast<uint4*>(out + o + g * 16);
      dst[0] = q[0];
      dst[1] = q[1];
    }
    DG_PH(4);
    tc_fence_before();
    __syncthreads();  // accumulator and operand tiles are free again
    DG_PH(5);
#ifdef DG_FIRST_DBG
    ++n_done;
#endif
  }
#ifdef DG_FIRST_DBG
  if (blockIdx.x == 0 && (tid == 0 || tid == 100))
    printf("tid %d tiles %d grid %d: halo %lld build %lld sync %lld mma+wait %lld epi %lld endsync %lld (clk per tile)\n", tid,
           n_done, (int)gridDim.x, c_ph[0] / n_done, c_ph[1] / n_done, c_ph[2] / n_done, c_ph[3] / n_done,
           c_ph[4] / n_done, c_ph[5] / n_done);
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * COUT) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------
// Weight gradient of the same first layers: dW[tap][ci][co] += alpha * sum_p x[p + off(tap)][ci] * dy[p][co].
// GEMM with K = pixels: A = the im2col tile above read "MN-major" (a pixel's K taps are its contiguous M block, the
// pixel pitch is the swizzle span), B = the dy tile [pixel][COUT] read MN-major, one MMA (K = 16) per tile row of 16
// pixels, accumulator rows = taps (hi half and lo half of the split image land in rows k and K + k and are added by
// the flush's atomics; rows beyond 2K hold garbage of the M = 128 instruction and are ignored).  The accumulator
// stays in TMEM for the CTA's whole pixel range.
// ---------------------------------------------------------------------------------------------------------
template <int KS, int CIN, int COUT>
__global__ void __launch_bounds__(256) wgrad_first_tc_kernel(const float* __restrict__ x, const bf16* __restrict__ dy,
                                                             float* __restrict__ dw, int N, int H, int W, float alpha) {
  typedef FirstGeom<KS, CIN> GEO;
  constexpr int K = GEO::K, KP = GEO::KP, ROWB = GEO::ROWB, UNITS = GEO::UNITS, HT = GEO::HT, PAD = KS / 2;
  constexpr int BROW = COUT * 2, BUNITS = BROW / 16;     // dy: bytes / 16-byte units per pixel (one swizzle span)
  constexpr uint32_t BLAY = COUT == 32 ? 4u : 6u;        // SW64 / SW32
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  // [A: 256 rows x ROWB + one slack row group][B: 256 pixels x BROW][halo][barrier, tmem slot]
  constexpr uint32_t A_BYTES = 256 * ROWB + 1024, B_BYTES = 256 * BROW;
  uint8_t* sA = gen;
  uint8_t* sB = gen + A_BYTES;
  uint32_t* s_halo = reinterpret_cast<uint32_t*>(gen + A_BYTES + B_BYTES);
  const uint32_t bar = base + A_BYTES + B_BYTES + (uint32_t)(HT * HT * CIN * 4 + 15) / 16 * 16;
  const uint32_t tmem_slot = bar + 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < 1024 / 16; i += 256)  // slack rows read by the M = 128 instruction past the last pixel
    reinterpret_cast<uint4*>(sA + 256 * ROWB)[i] = make_uint4(0u, 0u, 0u, 0u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int py = tid >> 4, px = tid & 15;
  const uint32_t a_off = (uint32_t)tid * ROWB, a_xor = (a_off >> 7) & (UNITS - 1);
  const uint32_t b_off = (uint32_t)tid * BROW, b_xor = (b_off >> 7) & (BUNITS - 1);
  // D = f32, A = B = bf16, both MN-major, M = 128, N = COUT
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(COUT >> 3) << 17) |
                         ((uint32_t)(128 >> 4) << 24);
  // A: the M block of KP taps is one pixel row (span); "next M block" = next pixel (LBO = span); K groups of 8 pixels
  const uint32_t hiA = ((uint32_t)(8 * ROWB) >> 4) | (1u << 14) | (GEO::LAYOUT << 29);
  const uint32_t loA = (((base)&0x3FFFFu) >> 4) | (((uint32_t)ROWB >> 4) << 16);
  const uint32_t hiB = ((uint32_t)(8 * BROW) >> 4) | (1u << 14) | (BLAY << 29);
  const uint32_t loB = (((base + A_BYTES) & 0x3FFFFu) >> 4) | (((uint32_t)(16 * BROW) >> 4) << 16);

  const int tiles_w = W / 16, tiles_h = H / 16;
  const int n_tiles = tiles_w * tiles_h * N;
  constexpr int NPRE = (HT * HT * CIN + 255) / 256;
  float pre[NPRE];
  uint4 dpre[BUNITS];
  auto fetch = [&](int t_) {
    const int tw_ = t_ % tiles_w, th_ = (t_ / tiles_w) % tiles_h, n_ = t_ / (tiles_w * tiles_h);
#pragma unroll
    for (int j = 0; j < NPRE; ++j) {
      const int i = tid + 256 * j;
      const int r = i / (HT * CIN), cc = i - r * (HT * CIN);
      const int gy = th_ * 16 - PAD + r, gx = tw_ * 16 - PAD + cc / CIN, ci = cc % CIN;
      pre[j] = (i < HT * HT * CIN && gy >= 0 && gy < H && gx >= 0 && gx < W)
                   ? __ldg(x + (((size_t)n_ * H + gy) * W + gx) * CIN + ci) : 0.f;
    }
    const uint4* dp = reinterpret_cast<const uint4*>(dy + (((size_t)n_ * H + th_ * 16 + py) * W + tw_ * 16 + px) * COUT);
#pragma unroll
    for (int u = 0; u < BUNITS; ++u) dpre[u] = __ldg(dp + u);
  };
  uint32_t phase = 0, first = 1;
  if ((int)blockIdx.x < n_tiles) fetch(blockIdx.x);
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    if (!first) {  // the previous tile's MMAs have consumed the operand tiles
      mbar_wait(bar, phase);
      phase ^= 1u;
    }
#pragma unroll
    for (int j = 0; j < NPRE; ++j)
      if (tid + 256 * j < HT * HT * CIN) s_halo[tid + 256 * j] = split_bf16(pre[j]);
#pragma unroll
    for (int u = 0; u < BUNITS; ++u) *reinterpret_cast<uint4*>(sB + b_off + ((((uint32_t)u) ^ b_xor) << 4)) = dpre[u];
    __syncthreads();
    {
      uint32_t wv[K + 1];
#pragma unroll
      for (int dyy = 0; dyy < KS; ++dyy)
#pragma unroll
        for (int dx = 0; dx < KS; ++dx)
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci)
            wv[(dyy * KS + dx) * CIN + ci] = s_halo[((py + dyy) * HT + px + dx) * CIN + ci];
      wv[K] = 0u;
      uint32_t row[KP / 2];
#pragma unroll
      for (int j = 0; j < KP / 2; ++j) {
        uint32_t pr = 0u;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int e = 2 * j + h;
          uint32_t half = 0u;
          if (e < K) half = wv[e] & 0xFFFFu;
          else if (e < 2 * K) half = wv[e - K] >> 16;
          pr |= half << (16 * h);
        }
        row[j] = pr;
      }
#pragma unroll
      for (int u = 0; u < UNITS; ++u)
        *reinterpret_cast<uint4*>(sA + a_off + ((((uint32_t)u) ^ a_xor) << 4)) =
            make_uint4(row[4 * u], row[4 * u + 1], row[4 * u + 2], row[4 * u + 3]);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (t + (int)gridDim.x < n_tiles) fetch(t + gridDim.x);
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int r = 0; r < 16; ++r)  // K = 16 pixels = one tile row per MMA
        tc_mma(tmem_base, ((uint64_t)hiA << 32) | (loA + (uint32_t)((r * 16 * ROWB) >> 4)),
               ((uint64_t)hiB << 32) | (loB + (uint32_t)((r * 16 * BROW) >> 4)), idesc, (first && r == 0) ? 0u : 1u);
      tc_commit(bar);
    }
    first = 0;
  }
  // ---- flush: rows k (hi) and K + k (lo) of the accumulator both add into dW[k][co] ----
  if (!first) {
    mbar_wait(bar, phase);
    tc_fence_after();
    if (warp < 4) {
      const int m = warp * 32 + (tid & 31);
#pragma unroll
      for (int g = 0; g < COUT / 16; ++g) {
        float v[16];
        tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(g * 16), v);
        if (m < 2 * K) {
          float* dst = dw + (size_t)(m < K ? m : m - K) * COUT + g * 16;
#pragma unroll
          for (int i = 0; i < 16; ++i) atomicAdd(dst + i, alpha * v[i]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(32) : "memory");
  }
}

template <int KS, int CIN, int COUT>
int launch_wgrad_first_tc(const WgradArgs& a, cudaStream_t st) {
  typedef FirstGeom<KS, CIN> GEO;
  constexpr uint32_t smem = 1024 + 256 * GEO::ROWB + 1024 + 256 * COUT * 2 + GEO::HT * GEO::HT * CIN * 4 + 16 + 32;
  static DgPerDevice site;  // function attributes and occupancy are set up once per device
  static std::mutex mu;
  int dev = 0, sms = 148, per_sm = 1;
  {
    std::lock_guard<std::mutex> lock(mu);
    bool first = false;
    DG_TRY(dg_device_enter(site, &dev, &first));
    if (first) {
      DG_CHECK_CUDA(cudaFuncSetAttribute(wgrad_first_tc_kernel<KS, CIN, COUT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      DG_CHECK_CUDA(cudaFuncSetAttribute(wgrad_first_tc_kernel<KS, CIN, COUT>,
                                         cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      cudaFuncAttributes fa;
      DG_CHECK_CUDA(cudaFuncGetAttributes(&fa, wgrad_first_tc_kernel<KS, CIN, COUT>));
      int occ = (int)((216u * 1024u) / (smem + 1024u));
      const int by_regs = 65536 / (((fa.numRegs + 7) / 8 * 8) * 256);
      if (occ > by_regs) occ = by_regs;
      if (occ > 8) occ = 8;
      if (dev < 64) site.val[dev] = occ < 1 ? 1 : occ;
      dg_device_mark(site, dev);
    }
    if (dev < 64) { sms = site.sms[dev]; per_sm = site.val[dev]; }
    else DG_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int n_tiles = (a.W / 16) * (a.H / 16) * a.N;
  const int grid = n_tiles < sms * per_sm ? n_tiles : sms * per_sm;
  wgrad_first_tc_kernel<KS, CIN, COUT><<<grid, 256, smem, st>>>((const float*)a.x0, (const bf16*)a.dy, a.dw, a.N, a.H,
                                                                a.W, a.alpha);
  DG_LAUNCH_CHECK();
  return 0;
}

template <int KS, int CIN, int COUT>
int launch_first_tc(const ConvArgs& a, cudaStream_t st) {
  typedef FirstGeom<KS, CIN> GEO;
  constexpr uint32_t A_BYTES = 256 * GEO::ROWB, B_BYTES = ((COUT * GEO::ROWB + 1023) / 1024) * 1024;
  constexpr uint32_t smem = 1024 + A_BYTES + B_BYTES + (GEO::HT * GEO::HT * CIN + 2 * COUT) * 4 + 16 + 32;
  // several CTAs per SM overlap one tile's load / build / MMA / store phases; the grid must not exceed what is
  // resident at once (persistent loop), and TMEM gives each CTA 2*COUT of the SM's 512 columns
  static DgPerDevice site;  // function attributes and occupancy are set up once per device
  static std::mutex mu;
  int dev = 0, sms = 148, per_sm = 1;
  {
    std::lock_guard<std::mutex> lock(mu);
    bool first = false;
    DG_TRY(dg_device_enter(site, &dev, &first));
    if (first) {
      DG_CHECK_CUDA(cudaFuncSetAttribute(conv_first_tc_kernel<KS, CIN, COUT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      DG_CHECK_CUDA(cudaFuncSetAttribute(conv_first_tc_kernel<KS, CIN, COUT>,
                                         cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      cudaFuncAttributes fa;
      DG_CHECK_CUDA(cudaFuncGetAttributes(&fa, conv_first_tc_kernel<KS, CIN, COUT>));
      int occ = (int)((216u * 1024u) / (smem + 1024u));          // shared memory (1 KB reserved per CTA)
      const int by_regs = 65536 / (((fa.numRegs + 7) / 8 * 8) * 256);
      if (occ > by_regs) occ = by_regs;
      if (occ > 512 / (2 * COUT)) occ = 512 / (2 * COUT);      // tensor memory columns
      if (occ > 8) occ = 8;
      if (dev < 64) site.val[dev] = occ < 1 ? 1 : occ;
      dg_device_mark(site, dev);
    }
    if (dev < 64) { sms = site.sms[dev]; per_sm = site.val[dev]; }
    else DG_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int n_tiles = (a.W / 16) * (a.H / 16) * a.N;
  const int grid = n_tiles < sms * per_sm ? n_tiles : sms * per_sm;
  conv_first_tc_kernel<KS, CIN, COUT><<<grid, 256, smem, st>>>((const float*)a.in0, a.w, a.scale, a.shift, (bf16*)a.out,
                                                               (const bf16*)a.mask_src, a.N, a.H, a.W, a.relu);
  DG_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// Returns 1 if the call was taken by the tensor-core first-layer kernel, 0 if the shape is not one of its cases
// (the caller falls back to the CUDA-core kernel), < 0 on error.
int conv_first_tc_try(const ConvArgs& a, cudaStream_t st) {
  if (a.in_dt != DT_F32 || a.out_dt != DT_BF16 || a.C1 != 0 || a.out_pre || a.film_g || a.add_src || !a.out || a.deconv ||
      a.head_w || !a.w)
    return 0;
  {  // conv2d_dis_0a without the im2col tile (conv_first_band.cu): any height, widths that are multiples of 8
    const int rb = conv_first_band_try(a, st);
    if (rb != 0) return rb;
  }
  if (a.H % 16 || a.W % 16 || a.H < 16 || a.W < 16) return 0;
  int r = 1;
  // conv2d_dis_0a only: for conv2d_gen_0 (3x3, 9-18 taps, 32 outputs) the CUDA-core kernel is faster (0.087 vs
  // 0.109 ms at 64 slices) and keeps fp32 weights on the layer that feeds the whole generator
  if (a.ks == 5 && a.C0 == 1 && a.Cout == 16) r = launch_first_tc<5, 1, 16>(a, st);
  else if (a.ks == 3 && a.C0 <= 2 && a.Cout == 32 && getenv("DEPGAN_FIRST_TC_GEN")) {
    r = a.C0 == 1 ? launch_first_tc<3, 1, 32>(a, st) : launch_first_tc<3, 2, 32>(a, st);
  } else return 0;
  return r < 0 ? r : 1;
}

// Weight gradient of conv2d_dis_0a in a bf16 network (fp32 image, bf16 gradient): 1 taken, 0 not a case, < 0 error.
int wgrad_first_tc_try(const WgradArgs& a, cudaStream_t st) {
  if (a.x_dt != DT_F32 || a.dy_dt != DT_BF16 || a.C1 != 0 || !a.dw) return 0;
  {  // without the im2col tile (wgrad_first_band.cu): widths that are multiples of 128, any height
    const int rb = wgrad_first_band_try(a, st);
    if (rb != 0) return rb;
  }
  if (a.H % 16 || a.W % 16 || a.H < 16 || a.W < 16) return 0;
  int r = 1;
  if (a.ks == 5 && a.C0 == 1 && a.Cout == 16) r = launch_wgrad_first_tc<5, 1, 16>(a, st);
  else return 0;
  return r < 0 ? r : 1;
}
