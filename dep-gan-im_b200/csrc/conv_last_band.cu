// Data gradient of conv2d_dis_0a (TG:319 backwards: 16 -> 1 channel, 5x5; dD/dx for the gradient penalty TG:543-545 and for
// the generator's adversarial terms) on the tensor cores -- the third of the conv_first_band.cu family: the 16-channel
// gradient rows in shared memory ARE the A operand.
//
//     out[y][x] = sum_dy sum_dx sum_ci G[y + dy - 2][x + dx - 2][ci] * w[dy][dx][ci]          (w: flipped taps, fp32 -> bf16)
//
// A row of G is kept as 128-byte spans of 4 pixels x 16 channels (bf16), SW128-swizzled by address, rows separated by one
// zero span (the horizontal halo of both neighbours).  M = 128 consecutive spans = 128 blocks of 4 output pixels; for a
// block m the 8 input pixels 4m-2 .. 4m+5 are the 256 bytes starting half a span before span m, i.e. the K-major operand
//     A[m][k = (j, ci)] = G[4 m - 2 + j][ci],     K = 128 = 8 K-steps, read as shifted views of the same bytes,
// against the banded weights B_dy[i][(j, ci)] = w[dy][j - i][ci] (0 <= j - i <= 4; N = 8: rows 0..3 carry the bf16 hi halves of the
// fp32 weights, rows 4..7 the lo halves, added in the epilogue).  40 MMAs
// (M 128, N 8, K 16) per 512 pixels; a TMEM lane ends up with the 4 output pixels of its block = one aligned float4 of the
// fp32 output, so the epilogue is a single coalesced store per lane.
//
// A CTA loads a band of R output rows (+4 halo rows) once and walks it in M tiles with two accumulator stages.
#include <cuda.h>

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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {  // bounded: a protocol bug traps
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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

constexpr int LB_THREADS = 160;  // warps 0..3 epilogue (TMEM lane quarters), warp 4 MMA issuer + TMEM owner
constexpr int KS = 5, PAD = 2, CIN = 16;
constexpr uint32_t B_DY = 2048;  // one banded weight matrix: 2 K blocks x (8 rows x 128 B)

struct LBandGeom {
  int wblk;        // 4-pixel blocks per row (W / 4)
  int bpr;         // spans per row in shared memory (wblk + 1: the zero span between rows)
  int R;           // output rows per band
  int rows_alloc;  // rows of the plane (band + halo + the slack the last M tile reads)
  int bands;       // bands per image
  uint32_t plane;  // bytes of the plane (a 1024-byte multiple), including the leading zero span
};

// address-based 128-byte swizzle (Swizzle<3,4,3>): 16-byte unit index ^= 128-byte row index mod 8
__device__ __forceinline__ uint32_t swz128(uint32_t off) { return off ^ (((off >> 7) & 7u) << 4); }

__global__ void __launch_bounds__(LB_THREADS) conv_last_band_kernel(const bf16* __restrict__ in, const float* __restrict__ w,
                                                                    float* __restrict__ out, int N, int H, int W,
                                                                    const LBandGeom g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  // [plane: leading zero span, rows of bpr spans][B: 5 x 2 KB][barriers, tmem slot]
  uint8_t* sG = gen;
  uint8_t* sB = gen + g.plane;
  const uint32_t b_addr = base + g.plane;
  const uint32_t bar = b_addr + KS * B_DY;
  const uint32_t full0 = bar, empty0 = bar + 16, tmem_slot = bar + 32;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // banded weights (bf16, SW128 K-major): B_dy[i][k = j*16 + ci] = w[dy][j - i][ci]; K block kb = k / 64 is one 1024-byte atom
  for (int e = tid; e < KS * 8 * 128; e += LB_THREADS) {
    const int k = e & 127, i = (e >> 7) & 7, dy = e >> 10;
    // rows 0..3: output pixel i with the bf16 hi half of the fp32 weights; rows 4..7: the same pixel with the lo half (the
    // N = 8 instruction has the room): 16 mantissa bits of the weights reach the products, like the fp32-weight CUDA-core
    // kernel this replaces, at no extra MMA; the epilogue adds columns i and 4 + i
    const int j = k >> 4, ci = k & 15, dx = j - (i & 3);
    const float v = (dx >= 0 && dx < KS) ? w[(size_t)(dy * KS + dx) * CIN + ci] : 0.f;
    const uint32_t off = (uint32_t)dy * B_DY + (uint32_t)(k >> 6) * 1024u + (uint32_t)i * 128u + (uint32_t)(k & 63) * 2u;
    const bf16 hi = __float2bfloat16_rn(v);
    *reinterpret_cast<bf16*>(sB + swz128(off)) = i < 4 ? hi : __float2bfloat16_rn(v - __bfloat162float(hi));
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // D = f32, A = B = bf16, both K-major, M = 128, N = 8
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(8 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t row_b = (uint32_t)g.bpr * 128u;
  const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO = 1024 B, version 1, SW128
  constexpr uint32_t LBO1 = 1u << 16;

  uint32_t tcount = 0;
  const int n_items = N * g.bands;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int n = item / g.bands, y0 = (item - n * g.bands) * g.R;
    const int rows = min(g.R, H - y0);
    // ---- band load: gradient rows y0-2 .. ; row rr sits at 128 + rr * row_b: wblk spans of data, then one zero span ----
    if (tid < 8) *reinterpret_cast<uint4*>(sG + swz128((uint32_t)tid * 16u)) = make_uint4(0u, 0u, 0u, 0u);  // leading zero span
    {
      const int units = g.bpr * 8;  // 16-byte units per row, the zero span included
      for (int rr = warp; rr < g.rows_alloc; rr += LB_THREADS / 32) {
        const int gy = y0 - PAD + rr;
        const bool row_ok = rr < rows + 2 * PAD && gy >= 0 && gy < H;
        const uint4* rp = reinterpret_cast<const uint4*>(in + ((size_t)n * H + (row_ok ? gy : 0)) * W * CIN);
        const uint32_t roff = 128u + (uint32_t)rr * row_b;
        for (int u0 = lane; u0 < units; u0 += 128) {
          uint4 v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int u = u0 + 32 * j;
            v[j] = make_uint4(0u, 0u, 0u, 0u);
            if (row_ok && u < g.wblk * 8) v[j] = __ldg(rp + u);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int u = u0 + 32 * j;
            if (u < units) *reinterpret_cast<uint4*>(sG + swz128(roff + (uint32_t)u * 16u)) = v[j];
          }
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const int nblk = rows * g.bpr;
    const int ntile = (nblk + 127) >> 7;
    if (warp == 4) {
      // ===== MMA issuer =====
      for (int t = 0; t < ntile; ++t) {
        const uint32_t tc = tcount + (uint32_t)t, s = tc & 1u, use = tc >> 1;
        mbar_wait(empty0 + 8 * s, (use & 1u) ^ 1u);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t d = tmem_base + s * 8u;
          // block m of output row r reads from half a span before its own span, in the row r + dy of the plane
          const uint32_t a0 = base + 128u + (uint32_t)t * (128u * 128u) - 64u;
#pragma unroll 1
          for (int dy = 0; dy < KS; ++dy) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              const uint32_t aa = a0 + (uint32_t)dy * row_b + 32u * ks;
              const uint32_t bb = b_addr + (uint32_t)dy * B_DY + (uint32_t)(ks >> 2) * 1024u + (uint32_t)(ks & 3) * 32u;
              tc_mma(d, ((uint64_t)desc_hi << 32) | (((aa & 0x3FFFFu) >> 4) | LBO1),
                     ((uint64_t)desc_hi << 32) | (((bb & 0x3FFFFu) >> 4) | LBO1), idesc, (dy | ks) != 0);
            }
          }
          tc_commit(full0 + 8 * s);
        }
        __syncwarp();
      }
    } else {
      // ===== epilogue: lane = block of 4 output pixels = one float4 =====
      for (int t = 0; t < ntile; ++t) {
        const uint32_t tc = tcount + (uint32_t)t, s = tc & 1u, use = tc >> 1;
        const int blk = t * 128 + warp * 32 + lane;
        const int r = blk / g.bpr, b = blk - r * g.bpr;
        mbar_wait(full0 + 8 * s, use & 1u);
        tc_fence_after();
        uint32_t v0, v1, v2, v3, l0, l1, l2, l3;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3), "=r"(l0), "=r"(l1), "=r"(l2), "=r"(l3)
                     : "r"(tmem_base + ((uint32_t)(warp * 32) << 16) + s * 8u));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * s);
        if (blk < nblk && b < g.wblk)
          *reinterpret_cast<float4*>(out + ((size_t)n * H + y0 + r) * W + (size_t)b * 4) =
              make_float4(__uint_as_float(v0) + __uint_as_float(l0), __uint_as_float(v1) + __uint_as_float(l1),
                          __uint_as_float(v2) + __uint_as_float(l2), __uint_as_float(v3) + __uint_as_float(l3));
      }
    }
    tcount += (uint32_t)ntile;
    tc_fence_before();
    __syncthreads();  // the band's MMAs have completed (the epilogue waited for the last commit): the plane is free
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(32) : "memory");
  }
}

}  // namespace

// 1 = launched, 0 = not a case of this kernel, < 0 error.  Called by conv_fwd_simt before the CUDA-core kernel.
int conv_last_band_try(const ConvArgs& a, cudaStream_t st) {
  static const bool off = getenv("DEPGAN_NO_FIRST_BAND") != nullptr;  // A/B switch
  if (off) return 0;
  if (a.ks != 5 || a.C0 != 16 || a.C1 != 0 || a.Cout != 1 || a.in_dt != DT_BF16 || a.out_dt != DT_F32) return 0;
  if (a.scale || a.shift || a.out_pre || a.film_g || a.add_src || a.mask_src || a.relu || a.deconv || a.head_w || !a.out || !a.w)
    return 0;
  if (a.W % 4 || a.W < 32 || a.H < 1) return 0;
  if ((reinterpret_cast<uintptr_t>(a.in0) & 15) || (reinterpret_cast<uintptr_t>(a.out) & 15)) return 0;
  LBandGeom g;
  g.wblk = a.W / 4;
  g.bpr = g.wblk + 1;
  // rows per band: the most efficient M tiling whose plane fits two CTAs per SM
  int best = 0;
  double best_eff = 0.0;
  for (int r = 1; r <= 32 && r <= a.H; ++r) {
    const int tiles = (r * g.bpr + 127) / 128;
    const int rows_alloc = (tiles * 128 + 1 + g.bpr - 1) / g.bpr + 2 * PAD;
    const size_t plane = ((size_t)128 + (size_t)rows_alloc * g.bpr * 128 + 1023) & ~(size_t)1023;
    if (plane + KS * B_DY + 1024 + 128 > 100 * 1024) break;
    const double eff = (double)(r * g.wblk) / (tiles * 128.0) * r / (r + 2.0 * PAD);  // MMA rows used x rows loaded once
    if (eff > best_eff) { best_eff = eff; best = r; }
  }
  if (best == 0) return 0;
  g.R = best;
  const int tiles = (g.R * g.bpr + 127) / 128;
  g.rows_alloc = (tiles * 128 + 1 + g.bpr - 1) / g.bpr + 2 * PAD;
  if (g.rows_alloc < g.R + 2 * PAD) g.rows_alloc = g.R + 2 * PAD;
  g.plane = (uint32_t)(((size_t)128 + (size_t)g.rows_alloc * g.bpr * 128 + 1023) & ~(size_t)1023);
  g.bands = (a.H + g.R - 1) / g.R;
  const uint32_t smem = 1024 + g.plane + KS * B_DY + 64;
  static DgPerDevice site;
  static std::mutex mu;
  int dev = 0, sms = 148;
  {
    std::lock_guard<std::mutex> lock(mu);
    bool first = false;
    DG_TRY(dg_device_enter(site, &dev, &first));
    if (first) {
      DG_CHECK_CUDA(cudaFuncSetAttribute(conv_last_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
      DG_CHECK_CUDA(cudaFuncSetAttribute(conv_last_band_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      dg_device_mark(site, dev);
    }
    if (dev < 64) sms = site.sms[dev];
    else DG_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const long long items = (long long)a.N * g.bands;
  const int grid = items < 2LL * sms ? (int)items : 2 * sms;
  conv_last_band_kernel<<<grid, LB_THREADS, smem, st>>>((const bf16*)a.in0, a.w, (float*)a.out, a.N, a.H, a.W, g);
  DG_LAUNCH_CHECK();
  return 1;
}
