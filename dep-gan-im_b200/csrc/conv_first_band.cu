// conv2d_dis_0a (TG:319: 1 -> 16 channels, 5x5, on the fp32 image) on the tensor cores WITHOUT an im2col tile: the image
// row itself is the A operand.
//
// Why.  conv_first_tc.cu builds an explicit [256 pixels][64] im2col tile per 16x16 tile (25 shared-memory reads and a
// 128-byte row write per pixel, read again by the MMAs): ~93 KB of shared-memory traffic per 256 pixels, which is what
// that kernel runs out of (3.7x off its HBM floor).  Here a pixel row in shared memory, padded by two zeros on the left, is
// read as the K-major operand
//     A[m][k] = row[8 m + k],   m = block of 8 output pixels, k = 0..15      (no-swizzle canonical layout: the rows of a
//                                                                            core matrix are 16 bytes = 8 pixels apart,
//                                                                            LBO = 16 B, SBO = 128 B: rows overlap)
// and the 5 horizontal taps live in a banded weight matrix
//     B_dy[(i, co)][k] = w[dy][k - i][co]  for 0 <= k - i <= 4, else 0,      i = pixel inside the block, N = 8 * 16 = 128,
// so one tcgen05.mma (M = 128 blocks = 1024 pixels, N = 128, K = 16) per vertical tap dy and per half of the split image
// (x = hi + lo in bf16 halves, like conv_first_tc.cu: 16 mantissa bits of the image reach the products) produces
//     D[m][(i, co)] = sum_dy sum_k row_{y+dy}[8 m + k] * B_dy[(i, co)][k] = out[y][8 m + i][co].
// 10 MMAs per 1024 pixels, ~10 KB of operand reads per 256 pixels.  A TMEM lane holds the 8 pixels x 16 channels of one
// block = 256 contiguous bytes of the NHWC output; the epilogue (folded bias, ReLU or the JVP mask) transposes them
// through a swizzled per-warp staging slot so that every global store instruction writes 512 contiguous bytes.
//
// Work split: a CTA takes a band of R image rows (R * (W/8 + 1) <= 1024 blocks: the padded rows are one linear array, the
// block after each row is padding and its results are dropped), loads rows y0-2 .. y0+R+1 once (fp32 -> hi / lo planes),
// and walks the band in M tiles of 128 blocks with two TMEM accumulator stages (the MMAs of tile t+1 run under the epilogue
// of tile t).  Two CTAs per SM overlap one CTA's band load with the other's tiles.
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
// 16 accumulator columns of this thread's TMEM lane; issue only -- several loads are put in flight before one wait
__device__ __forceinline__ void tc_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait(uint32_t (&a)[16], uint32_t (&b)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) asm volatile("" : "+r"(a[i]), "+r"(b[i])::"memory");
}

constexpr int BAND_THREADS = 288;  // warps 0..7: two epilogue groups (group = TMEM stage, warp % 4 = lane quarter); warp 8 issuer
constexpr int ISSUER_WARP = 8;
constexpr int KS = 5, PAD = 2, COUT = 16;
constexpr uint32_t B_DY = 128 * 16 * 2;     // one banded weight matrix [N = 128][K = 16] bf16
constexpr uint32_t STAGE_WARP = 32 * 128;   // one epilogue warp's staging slot: 32 blocks x (4 pixels x 16 channels x 2 B)
constexpr int EPI_WARPS = 8;

struct BandGeom {
  int P;           // pixels per padded row (W + 8)
  int bpr;         // blocks per padded row (P / 8); the last one is padding
  int R;           // image rows per band
  int rows_alloc;  // rows of a plane in shared memory (band + halo + the slack the last M tile reads)
  int bands;       // bands per image
  uint32_t plane;  // bytes of one plane (16-byte multiple)
};

// no-swizzle K-major shared-memory descriptor: 8 x 16-byte core matrices, LBO = K-adjacent, SBO = M/N-adjacent (bytes)
__device__ __forceinline__ uint64_t desc_k_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  const uint32_t lo = ((addr & 0x3FFFFu) >> 4) | ((lbo >> 4) << 16);
  const uint32_t hi = (sbo >> 4) | (1u << 14);  // version 1 (sm_100), layout type 0 = no swizzle
  return ((uint64_t)hi << 32) | lo;
}

__global__ void __launch_bounds__(BAND_THREADS, 2) conv_first_band_kernel(const float* __restrict__ x,
                                                                       const float* __restrict__ w,
                                                                       const float* __restrict__ scale,
                                                                       const float* __restrict__ shift,
                                                                       bf16* __restrict__ out,
                                                                       const bf16* __restrict__ mask, int N, int H, int W,
                                                                       int relu, const BandGeom g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 127u) & ~127u;
  uint8_t* gen = smem_raw + (base - raw);
  // [hi plane][lo plane][B: 5 x 4 KB][staging: 4 x 8 KB][scale, shift][barriers, tmem slot]
  uint8_t* sHi = gen;
  uint8_t* sLo = gen + g.plane;
  uint8_t* sB = gen + 2 * g.plane;
  uint8_t* sSt = sB + KS * B_DY;
  float* s_sc = reinterpret_cast<float*>(sSt + EPI_WARPS * STAGE_WARP);
  float* s_sh = s_sc + COUT;
  const uint32_t bar = base + 2 * g.plane + KS * B_DY + EPI_WARPS * STAGE_WARP + 2 * COUT * 4;
  const uint32_t full0 = bar, empty0 = bar + 16, tmem_slot = bar + 32;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == ISSUER_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // banded weights: B_dy[n = i*16 + co][k] = w[dy][k - i][co] (bf16), canonical no-swizzle K-major:
  // byte (n / 8) * 256 + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2
  for (int e = tid; e < KS * 128 * 16; e += BAND_THREADS) {
    const int k = e & 15, n = (e >> 4) & 127, dy = e >> 11;
    const int i = n >> 4, co = n & 15, dx = k - i;
    // the folded BN / bias scale multiplies the weights here (before their bf16 rounding), so the epilogue only adds
    const float v = (dx >= 0 && dx < KS) ? w[(size_t)(dy * KS + dx) * COUT + co] * (scale ? scale[co] : 1.f) : 0.f;
    *reinterpret_cast<bf16*>(sB + dy * B_DY + (n >> 3) * 256 + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2) =
        __float2bfloat16_rn(v);
  }
  if (tid < COUT) {
    s_sc[tid] = scale ? scale[tid] : 1.f;
    s_sh[tid] = shift ? shift[tid] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // D = f32, A = B = bf16, both K-major, M = 128, N = 128
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t hi_addr = base, lo_addr = base + g.plane, b_addr = base + 2 * g.plane;
  const uint32_t row_b = (uint32_t)g.P * 2u;  // bytes per padded row
  const int wblk = W >> 3;                    // valid blocks per row

#ifdef DG_BAND_DBG  // timing experiment only: clocks per phase, printed by CTA 0
  long long c_ph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, c_t = clock64();
  const long long c_start = c_t;
#define DG_PH(i_) do { const long long n_ = clock64(); c_ph[i_] += n_ - c_t; c_t = n_; } while (0)
#else
#define DG_PH(i_) do {} while (0)
#endif
  uint32_t tcount = 0;  // M tiles this CTA has processed (TMEM stage = tcount & 1, use = tcount >> 1)
  const int n_items = N * g.bands;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int n = item / g.bands, y0 = (item - n * g.bands) * g.R;
    const int rows = min(g.R, H - y0);
    // ---- band load: padded rows y0-2 .. ; pixel column c of the image sits at index c + 2; everything else is zero ----
    {
      const int q4 = g.P >> 2;  // groups of 4 padded pixels per row
      // a warp takes rows warp, warp + 9, ...; its lanes walk the row's 4-pixel groups, four groups (eight loads) in flight
      constexpr int NW = BAND_THREADS / 32;
      for (int rr = warp; rr < g.rows_alloc; rr += NW) {
        const int gy = y0 - PAD + rr;
        const bool row_ok = rr < rows + 2 * PAD && gy >= 0 && gy < H;
        const float* rp = x + ((size_t)n * H + (row_ok ? gy : 0)) * W;
        uint8_t* hrow = sHi + (uint32_t)rr * row_b;
        uint8_t* lrow = sLo + (uint32_t)rr * row_b;
        for (int q0 = lane; q0 < q4; q0 += 128) {
          float2 va[4][2];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int q = q0 + 32 * j, c0 = 4 * q - PAD;  // image columns c0 .. c0+3 (c0 even)
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
    DG_PH(0);  // band load
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    DG_PH(1);  // barrier after the load
    const int nblk = rows * g.bpr;
    const int ntile = (nblk + 127) >> 7;
    if (warp == ISSUER_WARP) {
      // ===== MMA issuer =====
      for (int t = 0; t < ntile; ++t) {
        const uint32_t tc = tcount + (uint32_t)t, s = tc & 1u, use = tc >> 1;
        mbar_wait(empty0 + 8 * s, (use & 1u) ^ 1u);
        tc_fence_after();
        DG_PH(2);  // issuer: wait for the stage
        if (lane == 0) {
          const uint32_t d = tmem_base + s * 128u;
          const uint32_t a_off = (uint32_t)t * 2048u;  // 128 blocks x 16 bytes
#pragma unroll
          for (int dy = 0; dy < KS; ++dy) {
            const uint64_t db = desc_k_noswz(b_addr + dy * B_DY, 128u, 256u);
            tc_mma(d, desc_k_noswz(hi_addr + dy * row_b + a_off, 16u, 128u), db, idesc, dy != 0);
            tc_mma(d, desc_k_noswz(lo_addr + dy * row_b + a_off, 16u, 128u), db, idesc, 1u);
          }
          tc_commit(full0 + 8 * s);
        }
        __syncwarp();
        DG_PH(3);  // issuer: issue
      }
    } else {
      // ===== epilogue: group = warp / 4 takes the tiles of TMEM stage `group`; warp % 4 = TMEM lane quarter =====
      const int grp = warp >> 2, q = warp & 3;
      uint8_t* slot = sSt + warp * STAGE_WARP;
      float sh[COUT];
#pragma unroll
      for (int c = 0; c < COUT; ++c) sh[c] = s_sh[c];
      for (int t = 0; t < ntile; ++t) {
        const uint32_t tc = tcount + (uint32_t)t, s = tc & 1u, use = tc >> 1;
        if ((int)s != grp) continue;
        const int blk0 = t * 128 + q * 32;       // first block of this warp in the band's linear order
        const int r0 = blk0 / g.bpr, b0 = blk0 - r0 * g.bpr;
        // this lane's block (mask addressing) and the four blocks it writes out per step, without further divisions
        int r = r0, b = b0 + lane;
        if (g.bpr >= 32) {                       // at most one row boundary inside the warp's 32 blocks
          if (b >= g.bpr) { b -= g.bpr; ++r; }
        } else {
          const int wr = b / g.bpr;
          r += wr; b -= wr * g.bpr;
        }
        const bool valid = blk0 + lane < nblk && b < wblk;
        const size_t pix0 = ((size_t)n * H + y0 + r) * W + (size_t)b * 8;
        // the eight blocks this lane writes out per half (step k: block 4k + lane / 8 of the warp): byte offset of the
        // block from the band's first output pixel, or -1 for padding / out-of-band blocks
        int woff[8];
        {
          int rr = r0, bb = b0 + (lane >> 3);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (bb >= g.bpr) { bb -= g.bpr; ++rr; }  // bpr >= 5 (host check): one boundary per step of 4 blocks at most
            woff[k] = (blk0 + 4 * k + (lane >> 3) < nblk && bb < wblk) ? (rr * wblk + bb) * 256 : -1;
            bb += 4;
          }
        }
        DG_PH(2);  // epilogue: set-up
        mbar_wait(full0 + 8 * s, use & 1u);
        tc_fence_after();
        DG_PH(3);  // epilogue: wait for the accumulator
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + s * 128u;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {  // pixels 4h .. 4h+3 of every block
#pragma unroll
          for (int i2 = 0; i2 < 4; i2 += 2) {  // two pixels (32 accumulator columns) per TMEM round trip
          uint32_t va[2][16];
          tc_ld16_issue(taddr + (uint32_t)((4 * h + i2) * 16), va[0]);
          tc_ld16_issue(taddr + (uint32_t)((4 * h + i2 + 1) * 16), va[1]);
          tc_ld_wait(va[0], va[1]);
          DG_PH(4);  // TMEM loads
          if (h == 1 && i2 == 2) {  // the accumulator stage is in registers: hand it back to the issuer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * s);
          }
#pragma unroll
          for (int ii = 0; ii < 2; ++ii) {
            const int i = i2 + ii;
            float v[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) v[c] = __uint_as_float(va[ii][c]) + sh[c];
            if (mask && valid) {  // activation pattern of the JVP pass (TG:543)
              const uint4* mp = reinterpret_cast<const uint4*>(mask + (pix0 + 4 * h + i) * COUT);
              const uint4 m0 = __ldg(mp), m1 = __ldg(mp + 1);
              const uint32_t mw[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                const float mv = __uint_as_float((c & 1) ? (mw[c >> 1] & 0xFFFF0000u) : (mw[c >> 1] << 16));
                v[c] = mv > 0.f ? v[c] : 0.f;
              }
            }
            uint4 pk[2];
            uint32_t* hp = reinterpret_cast<uint32_t*>(pk);
            if (relu) {  // the ReLU rides on the conversion
#pragma unroll
              for (int c = 0; c < 8; ++c)
                asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(hp[c]) : "f"(v[2 * c + 1]), "f"(v[2 * c]));
            } else {
#pragma unroll
              for (int c = 0; c < 8; ++c)
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hp[c]) : "f"(v[2 * c + 1]), "f"(v[2 * c]));
            }
            // 16-byte units 2i, 2i+1 of this lane's 128-byte row, XOR-swizzled by the lane: 8 lanes hit 8 bank groups
            *reinterpret_cast<uint4*>(slot + lane * 128 + (((2 * i) ^ (lane & 7)) << 4)) = pk[0];
            *reinterpret_cast<uint4*>(slot + lane * 128 + (((2 * i + 1) ^ (lane & 7)) << 4)) = pk[1];
          }
          }
          __syncwarp();
          DG_PH(5);  // math + staging
          // write-out: step k moves blocks 4k .. 4k+3 of the warp; 8 lanes write the 128 contiguous bytes of one block half
          {
            const int u = lane & 7;
            uint8_t* band_out = reinterpret_cast<uint8_t*>(out + ((size_t)n * H + y0) * W * COUT) + h * 128 + u * 16;
            uint4 val[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int lb = 4 * k + (lane >> 3);
              val[k] = *reinterpret_cast<const uint4*>(slot + lb * 128 + ((u ^ (lb & 7)) << 4));
            }
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (woff[k] >= 0) *reinterpret_cast<uint4*>(band_out + (uint32_t)woff[k]) = val[k];
          }
          __syncwarp();  // the slot is rewritten by the next half
          DG_PH(6);  // write-out
        }
      }
    }
    tcount += (uint32_t)ntile;
    tc_fence_before();
    __syncthreads();  // every MMA of the band has completed (the epilogue waited for the last commit): planes are free
    DG_PH(7);  // end-of-band barrier
  }
#ifdef DG_BAND_DBG
  if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 4 || warp == ISSUER_WARP))
    printf("warp %d grid %d tiles %u total %lld: load %lld sync %lld | [2] %lld [3] %lld | ld %lld math %lld out %lld | end %lld\n",
           warp, (int)gridDim.x, tcount, clock64() - c_start, c_ph[0], c_ph[1], c_ph[2], c_ph[3], c_ph[4], c_ph[5], c_ph[6],
           c_ph[7]);
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == ISSUER_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
  }
}

}  // namespace

// 1 = launched, 0 = not a case of this kernel, < 0 error.  Called by conv_first_tc_try (conv_first_tc.cu).
int conv_first_band_try(const ConvArgs& a, cudaStream_t st) {
  static const bool off = getenv("DEPGAN_NO_FIRST_BAND") != nullptr;  // A/B switch: the im2col kernel instead
  if (off) return 0;
  if (a.ks != 5 || a.C0 != 1 || a.C1 != 0 || a.Cout != 16 || a.in_dt != DT_F32 || a.out_dt != DT_BF16) return 0;
  if (a.W % 8 || a.W < 32 || a.H < 1 || !a.out || !a.w) return 0;
  BandGeom g;
  g.P = a.W + 8;
  g.bpr = g.P / 8;
  if (g.bpr > 1024) return 0;
  g.R = 1024 / g.bpr;
  if (g.R > a.H) g.R = a.H;
  const int tmax = (g.R * g.bpr + 127) / 128;
  g.rows_alloc = (tmax * 1024 + 16 + g.P - 1) / g.P + 2 * PAD;
  if (g.rows_alloc < g.R + 2 * PAD) g.rows_alloc = g.R + 2 * PAD;
  g.bands = (a.H + g.R - 1) / g.R;
  g.plane = ((uint32_t)g.rows_alloc * g.P * 2u + 15u) & ~15u;
  const uint32_t smem = 128 + 2 * g.plane + KS * B_DY + EPI_WARPS * STAGE_WARP + 2 * COUT * 4 + 64;
  if (smem > 110 * 1024) return 0;
  static DgPerDevice site;
  static std::mutex mu;
  int dev = 0, sms = 148;
  {
    std::lock_guard<std::mutex> lock(mu);
    bool first = false;
    DG_TRY(dg_device_enter(site, &dev, &first));
    if (first) {
      DG_CHECK_CUDA(cudaFuncSetAttribute(conv_first_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
      DG_CHECK_CUDA(cudaFuncSetAttribute(conv_first_band_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      dg_device_mark(site, dev);
    }
    if (dev < 64) sms = site.sms[dev];
    else DG_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const long long items = (long long)a.N * g.bands;
  const int grid = items < 2LL * sms ? (int)items : 2 * sms;
  conv_first_band_kernel<<<grid, BAND_THREADS, smem, st>>>((const float*)a.in0, a.w, a.scale, a.shift, (bf16*)a.out,
                                                           (const bf16*)a.mask_src, a.N, a.H, a.W, a.relu, g);
  DG_LAUNCH_CHECK();
  return 1;
}
