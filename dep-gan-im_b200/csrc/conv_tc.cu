// tcgen05 / TMEM implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulate in tensor memory).
//
// Replaces Keras Conv2D 'same' stride-1 (+BatchNormalization +Activation, TG:285-304), the k2s2 Conv2DTranspose
// (TG:307-312, as a 1x1 GEMM with 4*Cout columns scattered to the 2x2 output parities), and -- with flipped /
// transposed weights -- the data-gradient and JVP passes of the same layers (TG:543-549).
//
// GEMM view: M = pixels, N = output channels, K = taps x input channels.
//   * One CTA = one 16x16 pixel tile of one slice = two M=128 accumulators ("strips" of 8 columns x 16 rows),
//     N = ncta <= 256 output columns each, living in TMEM (2*ncta columns).
//   * K loop = channel chunks (kc = 16/32/64 channels = one 32/64/128-byte swizzle span) x taps.  Per chunk ONE
//     TMA box load brings the (16+ks-1)^2 halo tile [rows][cols][kc] into shared memory ('same' zero padding
//     comes from TMA out-of-bound fill); every tap then reads its A operand as a *shifted view* of that halo
//     tile: rows of the canonical K-major layout are consecutive pixels (pitch = swizzle span), 8-row groups
//     are tile rows (SBO = halo row pitch), so a tap is just a different descriptor start address.  The
//     hardware swizzle is a function of the shared-memory address, so views starting at any pixel are valid
//     as long as the TMA destination is pattern (1024 B) aligned.  Activations are read from L2/HBM once per
//     tile (x1.27 halo overhead) instead of once per tap.
//   * B (weights, [tap][n][k] bf16) streams through its own TMA ring, one (tap, chunk) tile per stage.
//   * warp 0 lane 0 = TMA producer, warp 1 lane 0 = MMA issuer (tcgen05.mma.cta_group::1.kind::f16),
//     then all four warps run the fused epilogue straight out of TMEM (tcgen05.ld 32x32b.x16).
//   * 2 CTAs/SM are co-resident (<= ~110 KB smem, <= 256 TMEM columns each for ncta <= 128), so one CTA's
//     epilogue overlaps the other's MMA stream.
#include <cuda.h>

#include "common.cuh"

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

struct TcGeom {
  int tiles_w, tiles_h;
  int nchunk0, nchunk1;  // channel chunks taken from in0 / in1
  int kc;                // channels per chunk
  int ncols_total;       // weight rows per tap (Cout, or 4*Cout for the transposed conv)
  int ncta;              // output columns per CTA
  int tmem_cols;         // power of two >= 2*ncta
  int na, nb;            // ring depths
  uint32_t a_bytes, b_bytes;  // stage strides (1024-aligned)
  uint32_t a_tx, b_tx;        // TMA transaction bytes per stage
  uint32_t layout;            // UMMA smem layout type (2 = SW128, 4 = SW64, 6 = SW32)
};

// ---------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------
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
// Bounded wait: a protocol bug traps (surfacing as a launch failure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1)
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

// K-major shared-memory matrix descriptor (SM100 version 1): rows at one swizzle-span pitch, 8-row groups at SBO.
__device__ __forceinline__ uint64_t make_sdesc(uint32_t addr, uint32_t sbo_bytes, uint32_t layout) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128.
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void ld16_bf16(const void* base, size_t elem_off, float (&v)[16]) {
  const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(base) + elem_off);
  uint4 q[2] = {__ldg(p), __ldg(p + 1)};
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(q);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void st16_bf16(void* base, size_t elem_off, const float (&v)[16]) {
  uint4 q[2];
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(q);
#pragma unroll
  for (int i = 0; i < 8; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  uint4* p = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(base) + elem_off);
  p[0] = q[0];
  p[1] = q[1];
}
__device__ __forceinline__ void ld16_f32(const float* p, float (&v)[16]) {
  const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float4 t = __ldg(q + i);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}

// ---------------------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------------------
template <int KS>
__global__ void __launch_bounds__(128) conv_tc_kernel(const __grid_constant__ CUtensorMap tmA0,
                                                      const __grid_constant__ CUtensorMap tmA1,
                                                      const __grid_constant__ CUtensorMap tmB, const ConvArgs a,
                                                      const TcGeom g) {
  constexpr int PAD = KS / 2, HT = 16 + KS - 1, TAPS = KS * KS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t a_base = base;
  const uint32_t b_base = a_base + g.na * g.a_bytes;
  const uint32_t bar_base = b_base + g.nb * g.b_bytes;
  const uint32_t fullA = bar_base, emptyA = fullA + 8 * g.na;
  const uint32_t fullB = emptyA + 8 * g.na, emptyB = fullB + 8 * g.nb;
  const uint32_t accum = emptyB + 8 * g.nb;
  const uint32_t tmem_slot = accum + 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rowb = g.kc * 2;

  if (threadIdx.x == 0) {
    for (int i = 0; i < g.na; ++i) { mbar_init(fullA + 8 * i, 1); mbar_init(emptyA + 8 * i, 1); }
    for (int i = 0; i < g.nb; ++i) { mbar_init(fullB + 8 * i, 1); mbar_init(emptyB + 8 * i, 1); }
    mbar_init(accum, 1);
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

  const int t = blockIdx.x;
  const int tw = t % g.tiles_w, th = (t / g.tiles_w) % g.tiles_h, n = t / (g.tiles_w * g.tiles_h);
  const int w0 = tw * 16, h0 = th * 16;
  const int n0 = blockIdx.y * g.ncta;
  const int nchunks = g.nchunk0 + g.nchunk1;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer: halo tiles run one chunk ahead of the weight stream =====
    auto issueA = [&](int c) {
      const int s = c % g.na;
      mbar_wait(emptyA + 8 * s, ((c / g.na) & 1) ^ 1);
      mbar_expect_tx(fullA + 8 * s, g.a_tx);
      const bool first = c < g.nchunk0;
      tma_load_4d(a_base + s * g.a_bytes, first ? &tmA0 : &tmA1, fullA + 8 * s,
                  (first ? c : c - g.nchunk0) * g.kc, w0 - PAD, h0 - PAD, n);
    };
    issueA(0);
    int bi = 0;
    for (int c = 0; c < nchunks; ++c) {
      if (c + 1 < nchunks) issueA(c + 1);
      const int kglob = c < g.nchunk0 ? c * g.kc : a.C0 + (c - g.nchunk0) * g.kc;
      for (int tap = 0; tap < TAPS; ++tap, ++bi) {
        const int s = bi % g.nb;
        mbar_wait(emptyB + 8 * s, ((bi / g.nb) & 1) ^ 1);
        mbar_expect_tx(fullB + 8 * s, g.b_tx);
        tma_load_2d(b_base + s * g.b_bytes, &tmB, fullB + 8 * s, kglob, tap * g.ncols_total + n0);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer =====
    const uint32_t idesc = make_idesc(g.ncta);
    const uint32_t sbo_a = HT * rowb, sbo_b = 8 * rowb;
    const int ksteps = g.kc / 16;
    int bi = 0;
    for (int c = 0; c < nchunks; ++c) {
      const int s = c % g.na;
      mbar_wait(fullA + 8 * s, (c / g.na) & 1);
      for (int tap = 0; tap < TAPS; ++tap, ++bi) {
        const int sb = bi % g.nb;
        mbar_wait(fullB + 8 * sb, (bi / g.nb) & 1);
        tc_fence_after();
        const int dy = tap / KS, dx = tap % KS;
        const uint32_t b_addr = b_base + sb * g.b_bytes;
#pragma unroll
        for (int strip = 0; strip < 2; ++strip) {
          const uint32_t a_addr = a_base + s * g.a_bytes + (uint32_t)((dy * HT + dx + strip * 8) * rowb);
          for (int k = 0; k < ksteps; ++k) {
            tc_mma(tmem_base + strip * g.ncta, make_sdesc(a_addr + k * 32, sbo_a, g.layout),
                   make_sdesc(b_addr + k * 32, sbo_b, g.layout), idesc, (c | tap | k) != 0);
          }
        }
        tc_commit(emptyB + 8 * sb);
      }
      tc_commit(emptyA + 8 * s);
    }
    tc_commit(accum);
  }
  __syncwarp();

  // ===== fused epilogue: TMEM -> registers -> global =====
  mbar_wait(accum, 0);
  tc_fence_after();

  const int r = warp * 32 + lane;  // accumulator row = TMEM lane
  const int ty = r >> 3;
  const int Cout = a.Cout;
  float head_acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
  for (int strip = 0; strip < 2; ++strip) {
    const int tx = strip * 8 + (r & 7);
    const int h = h0 + ty, w = w0 + tx;
    if (a.head_w) head_acc[0] = head_acc[1] = head_acc[2] = head_acc[3] = 0.f;
#pragma unroll 1
    for (int j = 0; j < g.ncta / 16; ++j) {
      float v[16];
      tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(strip * g.ncta + j * 16), v);
      const int col = n0 + j * 16;
      int c0 = col;
      size_t opix = ((size_t)n * a.H + h) * a.W + w;
      if (a.deconv) {
        const int ab = col / Cout;
        c0 = col - ab * Cout;
        opix = ((size_t)n * 2 * a.H + 2 * h + (ab >> 1)) * (2 * a.W) + 2 * w + (ab & 1);
      }
      const size_t off = opix * Cout + c0;
      float t[16];
      if (a.scale) {
        ld16_f32(a.scale + c0, t);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] *= t[i];
      }
      if (a.shift) {
        ld16_f32(a.shift + c0, t);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += t[i];
      }
      if (a.out_pre) st16_bf16(a.out_pre, off, v);
      if (a.film_g) {
        float fg[16], fb[16];
        ld16_f32(a.film_g + (size_t)n * a.film_stride + c0, fg);
        ld16_f32(a.film_b + (size_t)n * a.film_stride + c0, fb);
        ld16_bf16(a.res, off, t);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(fmaf(v[i], fg[i], fb[i]), 0.f) + t[i];
      }
      if (a.add_src) {
        ld16_bf16(a.add_src, off, t);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += t[i];
      }
      if (a.mask_src) {
        ld16_bf16(a.mask_src, off, t);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = t[i] > 0.f ? v[i] : 0.f;
      }
      if (a.relu) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
      }
      if (a.out) st16_bf16(a.out, off, v);
      if (a.head_w) {
        for (int k = 0; k < a.head_nc; ++k) {
          float s = head_acc[k];
#pragma unroll
          for (int i = 0; i < 16; ++i) s = fmaf(v[i], __ldg(a.head_w + (size_t)(c0 + i) * a.head_nc + k), s);
          head_acc[k] = s;
        }
      }
    }
    if (a.head_w) {
      const size_t pix = ((size_t)n * a.H + h) * a.W + w;
      float o[4];
      for (int k = 0; k < a.head_nc; ++k) o[k] = head_acc[k] + __ldg(a.head_b + k);
      if (a.head_act == 0) {
        for (int k = 0; k < a.head_nc; ++k) o[k] = tanhf(o[k]);
      } else if (a.head_act == 1) {
        float m = o[0];
        for (int k = 1; k < a.head_nc; ++k) m = fmaxf(m, o[k]);
        float s = 0.f;
        for (int k = 0; k < a.head_nc; ++k) { o[k] = expf(o[k] - m); s += o[k]; }
        for (int k = 0; k < a.head_nc; ++k) o[k] = o[k] / s;
      }
      for (int k = 0; k < a.head_nc; ++k) a.head_out[pix * a.head_nc + k] = o[k];
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(g.tmem_cols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
CUtensorMapSwizzle swizzle_for(int kc) {
  return kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}

int make_act_map(CUtensorMap* tm, const void* p, int C, int W, int H, int N, int kc, int ht) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)ht, (cuuint32_t)ht, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(kc), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    depgan_set_error("cuTensorMapEncodeTiled(activation) failed: " + std::to_string((int)r));
    return -1;
  }
  return 0;
}

int make_w_map(CUtensorMap* tm, const void* p, int Cin, int rows, int kc, int ncta) {
  cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
  cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)ncta};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(p), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(kc), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    depgan_set_error("cuTensorMapEncodeTiled(weights) failed: " + std::to_string((int)r));
    return -1;
  }
  return 0;
}

uint32_t round1024(uint32_t x) { return (x + 1023u) & ~1023u; }

constexpr uint32_t SMEM_BUDGET = 108 * 1024;  // two CTAs per SM

bool plan(const ConvArgs& a, TcGeom* g, uint32_t* smem_bytes) {
  const int ks = a.ks, ht = 16 + ks - 1;
  const int ncols = a.deconv ? 4 * a.Cout : a.Cout;
  const int nsplit = (ncols + 255) / 256;
  if (ncols % nsplit) return false;
  const int ncta = ncols / nsplit;
  if (ncta % 16 || ncta < 16 || ncta > 256) return false;
  if (a.deconv && (ncta % a.Cout)) return false;
  int tmem_cols = 32;
  while (tmem_cols < 2 * ncta) tmem_cols *= 2;
  for (int kc = 64; kc >= 16; kc /= 2) {
    if (a.C0 % kc || a.C1 % kc) continue;
    const uint32_t a_bytes = round1024((uint32_t)ht * ht * kc * 2), b_bytes = round1024((uint32_t)ncta * kc * 2);
    const int na = 2;
    if (na * a_bytes + 3 * b_bytes + 2048 > SMEM_BUDGET && kc > 16) continue;
    int nb = (int)((SMEM_BUDGET - 2048 - na * a_bytes) / b_bytes);
    if (nb > 8) nb = 8;
    if (nb < 2) return false;
    g->tiles_w = a.W / 16; g->tiles_h = a.H / 16;
    g->nchunk0 = a.C0 / kc; g->nchunk1 = a.C1 / kc;
    g->kc = kc; g->ncols_total = ncols; g->ncta = ncta; g->tmem_cols = tmem_cols;
    g->na = na; g->nb = nb; g->a_bytes = a_bytes; g->b_bytes = b_bytes;
    g->a_tx = (uint32_t)ht * ht * kc * 2; g->b_tx = (uint32_t)ncta * kc * 2;
    g->layout = kc == 64 ? 2u : kc == 32 ? 4u : 6u;
    *smem_bytes = 1024 + na * a_bytes + nb * b_bytes + 8 * (2 * na + 2 * nb + 1) + 16;
    return true;
  }
  return false;
}

}  // namespace

int conv_tc_init() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  DG_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn || q != cudaDriverEntryPointSuccess) {
    depgan_set_error("cuTensorMapEncodeTiled entry point not available");
    return -1;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  DG_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  DG_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  DG_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  return 0;
}

bool conv_tc_supported(const ConvArgs& a) {
  if (a.in_dt != DT_BF16 || a.out_dt != DT_BF16) return false;
  if (a.ks != 1 && a.ks != 3 && a.ks != 5) return false;
  if (a.deconv && a.ks != 1) return false;
  if (a.H % 16 || a.W % 16 || a.H < 16 || a.W < 16) return false;
  if (a.C0 % 16 || a.C1 % 16 || a.C0 < 16) return false;
  if (a.C1 > 0 && !a.in1) return false;
  if (!a.w_tc) return false;
  if (a.head_w && (a.deconv || a.Cout > 256 || a.head_nc > 4)) return false;
  TcGeom g;
  uint32_t smem;
  return plan(a, &g, &smem);
}

int conv_fwd_tc(const ConvArgs& a, cudaStream_t st) {
  if (a.N <= 0) return 0;
  DG_TRY(conv_tc_init());
  TcGeom g;
  uint32_t smem;
  DG_REQUIRE(conv_tc_supported(a) && plan(a, &g, &smem), "conv_fwd_tc: unsupported shape");
  const int ht = 16 + a.ks - 1;
  CUtensorMap tmA0, tmA1, tmB;
  DG_TRY(make_act_map(&tmA0, a.in0, a.C0, a.W, a.H, a.N, g.kc, ht));
  if (a.C1 > 0) DG_TRY(make_act_map(&tmA1, a.in1, a.C1, a.W, a.H, a.N, g.kc, ht));
  else tmA1 = tmA0;
  DG_TRY(make_w_map(&tmB, a.w_tc, a.C0 + a.C1, a.ks * a.ks * g.ncols_total, g.kc, g.ncta));
  dim3 grid(g.tiles_w * g.tiles_h * a.N, g.ncols_total / g.ncta);
  switch (a.ks) {
    case 1: conv_tc_kernel<1><<<grid, 128, smem, st>>>(tmA0, tmA1, tmB, a, g); break;
    case 3: conv_tc_kernel<3><<<grid, 128, smem, st>>>(tmA0, tmA1, tmB, a, g); break;
    case 5: conv_tc_kernel<5><<<grid, 128, smem, st>>>(tmA0, tmA1, tmB, a, g); break;
  }
  DG_LAUNCH_CHECK();
  return 0;
}
