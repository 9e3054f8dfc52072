// Weight gradient of conv2d_dis_0a (TG:319: 1 -> 16 channels, 5x5, fp32 image, bf16 gradient) on the tensor cores with the
// image row itself as an operand -- the companion of conv_first_band.cu.
//
//     dW[dy][dx][co] = sum_{n, y, x} X[n][y + dy - 2][x + dx - 2] * G[n][y][x][co]
//
// Per image row y and vertical tap dy one small GEMM over the row's 8-pixel blocks b (the K dimension):
//     T_dy[(i, co)][j] += sum_b G[y][8 b + i][co] * Xp_{y+dy}[8 b + j],          i = 0..7, j = 0..15, Xp = row padded by 2
//   A (M = 128 = (i, co), MN-major): the 256 contiguous bytes of block b of the gradient row, brought in by TMA as two
//     128-byte swizzled spans (pixels 0..3 / 4..7 of every block);
//   B (N = 16 = j, MN-major, no swizzle): the padded image row in shared memory, read as B[b][j] = Xp[8 b + j] (consecutive
//     blocks are 16 bytes apart, the two 8-pixel units of a row overlap the next block's: LBO = 128 B, SBO = 16 B);
// and at the end dW[dy][dx][co] = sum_i T_dy[(i, co)][i + dx].  The image is split x = hi + lo into bf16 halves (both
// accumulate into the same T_dy), so its 16 mantissa bits reach the products as in wgrad_first_tc_kernel.
// 20 MMAs (M 128, N 16, K 16 blocks) per 256-pixel row against 16 + 16 per 256 pixels AND an im2col tile before: the
// kernel no longer builds anything in shared memory except the split image rows.
//
// A CTA walks bands of R image rows: all warps load the band's R + 4 padded image rows (fp32 -> hi / lo planes), then one
// warp streams the gradient rows through a TMA ring while another issues the MMAs; the five accumulators (80 TMEM
// columns) live for the CTA's whole range and are reduced over i in shared memory before one atomic per weight.
#include <cuda.h>

#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace {

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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {  // bounded: a protocol bug traps
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
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

constexpr int WB_THREADS = 256;  // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner; warps 0..3 flush; all load the image
constexpr int KS = 5, PAD = 2, COUT = 16;
constexpr int NST = 4;           // gradient-row stages in the TMA ring
constexpr int BAND_R = 8;        // image rows per band

struct WBandGeom {
  int P;            // pixels per padded image row (W + 8)
  int wblk;         // 8-pixel blocks per row (W / 8, a multiple of 16)
  int bands;        // bands per image
  uint32_t plane;   // bytes of one image plane ((BAND_R + 4) padded rows)
  uint32_t stage;   // bytes of one gradient-row stage (2 spans x wblk x 128 B)
};

__global__ void __launch_bounds__(WB_THREADS) wgrad_first_band_kernel(const __grid_constant__ CUtensorMap tm_g,
                                                                      const float* __restrict__ x, float* __restrict__ dw,
                                                                      int N, int H, int W, float alpha, const WBandGeom g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // the swizzled gradient stages need 1024-byte alignment
  uint8_t* gen = smem_raw + (base - raw);
  // [gradient ring: NST stages][hi plane][lo plane][reduction: 400 floats][barriers, tmem slot]
  const uint32_t ring = base;
  uint8_t* sHi = gen + NST * g.stage;
  uint8_t* sLo = sHi + g.plane;
  float* s_red = reinterpret_cast<float*>(sLo + g.plane);
  const uint32_t hi_addr = base + NST * g.stage, lo_addr = hi_addr + g.plane;
  const uint32_t bar = lo_addr + g.plane + 400 * 4;
  const uint32_t full0 = bar, empty0 = bar + 8 * NST, done = bar + 16 * NST, tmem_slot = done + 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < 400; i += WB_THREADS) s_red[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // D = f32, A = B = bf16, both MN-major, M = 128, N = 16
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(16 >> 3) << 17) |
                         ((uint32_t)(128 >> 4) << 24);
  const uint32_t row_b = (uint32_t)g.P * 2u;
  const uint32_t span_b = (uint32_t)g.wblk * 128u;  // one 64-element span of every block of a row
  const int ksteps = g.wblk >> 4;                   // K = 16 blocks per MMA

  uint32_t rcount = 0;   // gradient rows this CTA has streamed (ring position = rcount % NST)
  bool any = false;
  const int n_items = N * g.bands;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int n = item / g.bands, y0 = (item - n * g.bands) * BAND_R;
    const int rows = min(BAND_R, H - y0);
    // ---- image rows y0-2 .. y0+rows+1, padded by 2 zeros on the left, split into bf16 hi / lo planes ----
    {
      const int q4 = g.P >> 2;
      constexpr int NW = WB_THREADS / 32;
      for (int rr = warp; rr < BAND_R + 2 * PAD; rr += NW) {
        const int gy = y0 - PAD + rr;
        const bool row_ok = rr < rows + 2 * PAD && gy >= 0 && gy < H;
        const float* rp = x + ((size_t)n * H + (row_ok ? gy : 0)) * W;
        uint8_t* hrow = sHi + (uint32_t)rr * row_b;
        uint8_t* lrow = sLo + (uint32_t)rr * row_b;
        for (int q0 = lane; q0 < q4; q0 += 128) {
          float2 va[4][2];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int q = q0 + 32 * j, c0 = 4 * q - PAD;
            va[j][0] = va[j][1] = make_float2(0.f, 0.f);
            if (row_ok && q < q4) {
              if (c0 >= 0 && c0 + 1 < W) va[j][0] = __ldg(reinterpret_cast<const float2*>(rp + c0));
              if (c0 + 3 < W) va[j][1] = __ldg(reinterpret_cast<const float2*>(rp + c0 + 2));
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int q = q0 + 32 * j;
            if (q < q4) {
              uint32_t hw[2], lw[2];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                uint32_t hp2;
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hp2) : "f"(va[j][h].y), "f"(va[j][h].x));
                const float r0_ = va[j][h].x - __uint_as_float(hp2 << 16);
                const float r1_ = va[j][h].y - __uint_as_float(hp2 & 0xFFFF0000u);
                hw[h] = hp2;
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lw[h]) : "f"(r1_), "f"(r0_));
              }
              *reinterpret_cast<uint2*>(hrow + q * 8) = make_uint2(hw[0], hw[1]);
              *reinterpret_cast<uint2*>(lrow + q * 8) = make_uint2(lw[0], lw[1]);
            }
          }
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
      // ===== TMA producer: gradient row y0 + r = blocks [blk0, blk0 + wblk), two spans =====
      for (int r = 0; r < rows; ++r) {
        const uint32_t rc = rcount + (uint32_t)r, s = rc % NST, use = rc / NST;
        mbar_wait(empty0 + 8 * s, (use & 1u) ^ 1u);
        if (lane == 0) {
          const int blk0 = (int)((((size_t)n * H + y0 + r) * W) >> 3);
          mbar_expect_tx(full0 + 8 * s, g.stage);
          tma_load_3d(ring + s * g.stage, &tm_g, full0 + 8 * s, 0, 0, blk0);
          tma_load_3d(ring + s * g.stage + span_b, &tm_g, full0 + 8 * s, 0, 1, blk0);
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      // ===== MMA issuer =====
      for (int r = 0; r < rows; ++r) {
        const uint32_t rc = rcount + (uint32_t)r, s = rc % NST, use = rc / NST;
        mbar_wait(full0 + 8 * s, use & 1u);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a0 = ring + s * g.stage;
          // A: SW128 MN-major: 64-element spans (LBO = span_b apart), k rows 128 B apart, 8-row atoms 1024 B apart (SBO)
          const uint32_t a_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
          // B: no-swizzle MN-major: 8-pixel units 16 B apart (SBO), k = consecutive blocks 16 B apart, 8 blocks 128 B (LBO)
          const uint32_t b_hi = (16u >> 4) | (1u << 14);
          for (int dy = 0; dy < KS; ++dy) {
            const uint32_t d = tmem_base + (uint32_t)(dy * 16);
#pragma unroll
            for (int part = 0; part < 2; ++part) {
              const uint32_t brow = (part ? lo_addr : hi_addr) + (uint32_t)(r + dy) * row_b;
              for (int k = 0; k < ksteps; ++k) {
                const uint32_t a_lo = (((a0 + (uint32_t)k * 2048u) & 0x3FFFFu) >> 4) | ((span_b >> 4) << 16);
                const uint32_t b_lo = (((brow + (uint32_t)k * 256u) & 0x3FFFFu) >> 4) | ((128u >> 4) << 16);
                tc_mma(d, ((uint64_t)a_hi << 32) | a_lo, ((uint64_t)b_hi << 32) | b_lo, idesc,
                       (any || r != 0 || part != 0 || k != 0) ? 1u : 0u);
              }
            }
          }
          tc_commit(empty0 + 8 * s);
        }
        __syncwarp();
      }
      if (lane == 0) tc_commit(done);  // every MMA of the band has read the image planes
      __syncwarp();
    }
    rcount += (uint32_t)rows;
    any = true;
    // the planes are rewritten by the next band: wait until the band's MMAs have completed
    {
      const uint32_t bands_done = (uint32_t)((item - (int)blockIdx.x) / (int)gridDim.x);
      mbar_wait(done, bands_done & 1u);
    }
    tc_fence_after();
    __syncthreads();
  }
  // ---- flush: dW[dy][dx][co] += alpha * sum_i T_dy[(i, co)][i + dx] ----
  if (any) {
    if (warp < 4) {
      const int m = warp * 32 + lane, i = m >> 4, co = m & 15;
#pragma unroll 1
      for (int dy = 0; dy < KS; ++dy) {
        float v[16];
        tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(dy * 16), v);
#pragma unroll
        for (int dx = 0; dx < KS; ++dx) {
          float t = 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) t = (j == i + dx) ? v[j] : t;
          atomicAdd(&s_red[(dy * KS + dx) * COUT + co], t);
        }
      }
    }
    tc_fence_before();
    __syncthreads();
    for (int e = tid; e < 400; e += WB_THREADS) atomicAdd(dw + e, alpha * s_red[e]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(128) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_wb = nullptr;

}  // namespace

// 1 = launched, 0 = not a case of this kernel, < 0 error.  Called by wgrad_first_tc_try (conv_first_tc.cu).
int wgrad_first_band_try(const WgradArgs& a, cudaStream_t st) {
  static const bool off = getenv("DEPGAN_NO_FIRST_BAND") != nullptr;  // A/B switch: the im2col kernel instead
  if (off) return 0;
  if (a.ks != 5 || a.C0 != 1 || a.C1 != 0 || a.Cout != 16 || a.x_dt != DT_F32 || a.dy_dt != DT_BF16 || !a.dw) return 0;
  if (a.W % 128 || a.W < 128 || a.W > 2048 || a.H < 1) return 0;  // 16 blocks per K step; TMA box <= 256 blocks
  if ((reinterpret_cast<uintptr_t>(a.dy) & 15) != 0) return 0;
  WBandGeom g;
  g.P = a.W + 8;
  g.wblk = a.W / 8;
  g.bands = (a.H + BAND_R - 1) / BAND_R;
  g.plane = ((uint32_t)(BAND_R + 2 * PAD + 1) * g.P * 2u + 15u) & ~15u;  // one slack row: the last K step reads 8 pixels past
  g.stage = 2u * (uint32_t)g.wblk * 128u;
  const uint32_t smem = 1024 + NST * g.stage + 2 * g.plane + 400 * 4 + 16 * NST + 8 + 16 + 64;
  if (smem > 100 * 1024) return 0;
  static DgPerDevice site;
  static std::mutex mu;
  int dev = 0, sms = 148;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (!g_encode_wb) {
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult q;
      DG_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
      if (!fn || q != cudaDriverEntryPointSuccess) {
        depgan_set_error("cuTensorMapEncodeTiled entry point not available");
        return -1;
      }
      g_encode_wb = reinterpret_cast<EncodeTiledFn>(fn);
    }
    bool first = false;
    DG_TRY(dg_device_enter(site, &dev, &first));
    if (first) {
      DG_CHECK_CUDA(cudaFuncSetAttribute(wgrad_first_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      DG_CHECK_CUDA(cudaFuncSetAttribute(wgrad_first_band_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      dg_device_mark(site, dev);
    }
    if (dev < 64) sms = site.sms[dev];
    else DG_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  // the gradient as (64 elements = 4 pixels x 16 channels, 2 spans per block, blocks): one box = one span of a whole row
  CUtensorMap tm;
  {
    const cuuint64_t nblk = (cuuint64_t)a.N * a.H * a.W / 8;
    cuuint64_t dims[3] = {64, 2, nblk};
    cuuint64_t strides[2] = {128, 256};
    cuuint32_t box[3] = {64, 1, (cuuint32_t)g.wblk};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = g_encode_wb(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(a.dy), dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      depgan_set_error("cuTensorMapEncodeTiled(first-layer weight gradient) failed: " + std::to_string((int)r));
      return -1;
    }
  }
  const long long items = (long long)a.N * g.bands;
  const int per_sm = (int)((216u * 1024u) / (smem + 1024u)) < 4 ? (int)((216u * 1024u) / (smem + 1024u)) : 4;  // TMEM: 4 x 128 columns
  const long long cap = (long long)sms * (per_sm < 1 ? 1 : per_sm);
  const int grid = items < cap ? (int)items : (int)cap;
  wgrad_first_band_kernel<<<grid, WB_THREADS, smem, st>>>(tm, (const float*)a.x0, a.dw, a.N, a.H, a.W, a.alpha, g);
  DG_LAUNCH_CHECK();
  return 1;
}
