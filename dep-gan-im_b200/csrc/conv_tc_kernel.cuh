// tcgen05 / TMEM implicit-GEMM convolution kernel for sm_100a (bf16 operands, fp32 accumulate in tensor memory).
// Device code only; the host side (planning, tensor maps, dispatch) is conv_tc.cu and the per-kernel-size
// instantiation units are conv_tc_k1.cu / conv_tc_k3.cu / conv_tc_k5.cu.
//
// Replaces Keras Conv2D 'same' stride-1 (+BatchNormalization +Activation, TG:285-304), the k2s2 Conv2DTranspose
// (TG:307-312, as a 1x1 GEMM with 4*Cout columns scattered to the 2x2 output parities), and -- with flipped /
// transposed weights -- the data-gradient and JVP passes of the same layers (TG:543-549).
//
// GEMM view: M = pixels, N = output channels, K = taps x input channels.
//   * One work item = one 16x16 pixel tile of one slice = two M=128 accumulators ("strips" of 8 columns x 16 rows),
//     N = ncta <= 256 output columns each, living in TMEM (2*ncta columns per accumulator stage).
//   * K loop = channel chunks (kc = 16/32/64 channels = one 32/64/128-byte swizzle span) x taps.  Per chunk ONE
//     TMA box load brings the (16+ks-1)^2 halo tile [rows][cols][kc] into shared memory ('same' zero padding
//     comes from TMA out-of-bound fill); every tap then reads its A operand as a *shifted view* of that halo
//     tile: rows of the canonical K-major layout are consecutive pixels (pitch = swizzle span), 8-row groups
//     are tile rows (SBO = halo row pitch), so a tap is just a different descriptor start address.  The
//     hardware swizzle is a function of the shared-memory address, so views starting at any pixel are valid
//     as long as the TMA destination is pattern (1024 B) aligned.  Activations are read from L2/HBM once per
//     tile (x1.27 halo overhead) instead of once per tap.
//   * B (weights, [tap][n][k] bf16) streams through its own TMA ring, one (tap, chunk) tile per stage.
//     When every (chunk, tap) tile of the layer fits next to the activation ring (32->32 ... 96->96 3x3 layers)
//     the weights are loaded ONCE per CTA and stay resident (template RES).
//   * Persistent: grid = min(#work items, #SMs), one CTA per SM looping over (tile, n-split) items.
//     warp 0 = TMA producer, warp 1 = MMA issuer (tcgen05.mma.cta_group::1.kind::f16), warps 4..11 = fused epilogue
//     straight out of TMEM.  Two TMEM accumulator stages (when 4*ncta <= 512 columns) overlap the epilogue of
//     item i with the MMA stream of item i+1; the producer runs ahead across items.
//   * Epilogue (template EPI = which global side inputs exist: bit 0 FiLM residual, bit 1 add / mask sources):
//     the side inputs of the next PF 16-channel chunks -- across item boundaries -- are always in flight in a
//     register ring, and the tcgen05.ld of chunk j+1 is issued before chunk j is processed, so neither the L2/HBM
//     nor the TMEM read latency is paid per chunk.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace convtc {

struct TcGeom {
  int tiles_w, tiles_h;
  int nchunk0, nchunk1;  // channel chunks taken from in0 / in1
  int kc;                // channels per chunk
  int ncols_total;       // weight rows per tap (Cout, or 4*Cout for the transposed conv)
  int ncta;              // output columns per CTA
  int tmem_cols;         // power of two >= 2*ncta*acc_stages
  int na, nb;            // ring depths
  uint32_t a_bytes, b_bytes;  // stage strides (1024-aligned)
  uint32_t a_tx, b_tx;        // TMA transaction bytes per stage
  uint32_t layout;            // UMMA smem layout type (2 = SW128, 4 = SW64, 6 = SW32)
  int b_resident;             // 1: all (chunk, tap) weight tiles live in smem for the CTA's lifetime
  int acc_stages;             // TMEM accumulator stages (2 when 4*ncta <= 512)
};

// one launcher per kernel size, defined in conv_tc_k{1,3,5}.cu; epi = bit mask of the side inputs the call uses
int launch_ks1(int grid, uint32_t smem, cudaStream_t st, const CUtensorMap& tmA0, const CUtensorMap& tmA1,
               const CUtensorMap& tmB, const ConvArgs& a, const TcGeom& g, int epi);
int launch_ks3(int grid, uint32_t smem, cudaStream_t st, const CUtensorMap& tmA0, const CUtensorMap& tmA1,
               const CUtensorMap& tmB, const ConvArgs& a, const TcGeom& g, int epi);
int launch_ks5(int grid, uint32_t smem, cudaStream_t st, const CUtensorMap& tmA0, const CUtensorMap& tmA1,
               const CUtensorMap& tmB, const ConvArgs& a, const TcGeom& g, int epi);
int set_attrs_ks1();
int set_attrs_ks3();
int set_attrs_ks5();

#ifdef __CUDACC__
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
// 16 accumulator columns of this thread's TMEM lane.  Issue and wait are split so the next chunk's load can be in
// flight while the current chunk is processed; the wait names the registers as in/out operands so no use of them
// can be scheduled above it.
__device__ __forceinline__ void tc_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128.
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

struct Packed16 {  // 16 bf16 values as loaded (two 128-bit words)
  uint4 q[2];
  __device__ __forceinline__ void load(const void* base, size_t elem_off) {
    const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(base) + elem_off);
    q[0] = __ldg(p);
    q[1] = __ldg(p + 1);
  }
  __device__ __forceinline__ float get(int i) const {
    const uint32_t w = reinterpret_cast<const uint32_t*>(q)[i >> 1];
    return __uint_as_float((i & 1) ? (w & 0xFFFF0000u) : (w << 16));
  }
};
__device__ __forceinline__ void st16_bf16(void* base, size_t elem_off, const float (&v)[16]) {
  uint4 q[2];
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(q);
#pragma unroll
  for (int i = 0; i < 8; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  uint4* p = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(base) + elem_off);
  p[0] = q[0];
  p[1] = q[1];
}

constexpr int TC_THREADS = 384;  // warpgroup 0: producer, MMA issuer, 2 idle warps; warpgroups 1-2: epilogue
constexpr int EPI_PF = 4;  // side-input lookahead of the epilogue, in 16-channel chunks

struct Ring {
  int idx = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance(int n) {
    if (++idx == n) { idx = 0; phase ^= 1u; }
  }
};

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------------------------------------------------
// kernel: persistent, warp-specialised.  warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2-3 idle
// (they only complete warpgroup 0 so it can hand its registers over with setmaxnreg), warps 4..11 = epilogue
// (warp w reads TMEM lane quarter w%4 of strip (w-4)/4) running with the registers warpgroup 0 gave up.
// ---------------------------------------------------------------------------------------------------------
template <int KS, int KSTEPS, bool RES, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap tmA0,
                                                                const __grid_constant__ CUtensorMap tmA1,
                                                                const __grid_constant__ CUtensorMap tmB,
                                                                const ConvArgs a, const TcGeom g) {
  constexpr int PAD = KS / 2, HT = 16 + KS - 1, TAPS = KS * KS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t a_base = base;
  const uint32_t b_base = a_base + g.na * g.a_bytes;
  const uint32_t bar_base = b_base + g.nb * g.b_bytes;
  const uint32_t fullA = bar_base, emptyA = fullA + 8 * g.na;
  const uint32_t fullB = emptyA + 8 * g.na, emptyB = fullB + 8 * g.nb;
  const uint32_t accFull = emptyB + 8 * g.nb, accEmpty = accFull + 16;
  const uint32_t tmem_slot = accEmpty + 16;
  const uint32_t ss_off = tmem_slot + 16;  // scale / shift staging: 2 * ncols_total floats, then head weights
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  float* s_scale = reinterpret_cast<float*>(smem_raw + (ss_off - raw));
  float* s_shift = s_scale + g.ncols_total;
  float4* s_head = reinterpret_cast<float4*>(s_shift + g.ncols_total);  // [Cout] x (up to 4 head outputs)
  // FiLM folded with BN per (sample, channel): [2 slots][2][ncta] floats, rebuilt per work item by the epilogue warps
  float* s_film = reinterpret_cast<float*>(s_head + a.Cout);

  // warp index made provably warp-uniform so the role loops run on the uniform datapath
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  constexpr int rowb = KSTEPS * 32;  // bytes per pixel row of a chunk = one swizzle span (kc = 16*KSTEPS)
  const int nchunks = g.nchunk0 + g.nchunk1;
  const int nsplit = g.ncols_total / g.ncta;
  const int tiles_per_img = g.tiles_w * g.tiles_h;
  const int n_items = tiles_per_img * a.N * nsplit;

  if (threadIdx.x == 0) {
    for (int i = 0; i < g.na; ++i) { mbar_init(fullA + 8 * i, 1); mbar_init(emptyA + 8 * i, 1); }
    for (int i = 0; i < g.nb; ++i) { mbar_init(fullB + 8 * i, 1); mbar_init(emptyB + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(accFull + 8 * i, 1); mbar_init(accEmpty + 8 * i, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(g.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {  // BN scale / shift (or bias) once per CTA
    const int cmod = a.Cout;
    for (int i = threadIdx.x; i < g.ncols_total; i += TC_THREADS) {
      s_scale[i] = a.scale ? a.scale[i % cmod] : 1.f;
      s_shift[i] = a.shift ? a.shift[i % cmod] : 0.f;
    }
    if (a.head_w) {
      for (int i = threadIdx.x; i < a.Cout; i += TC_THREADS) {
        float hv[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k = 0; k < a.head_nc; ++k) hv[k] = a.head_w[(size_t)i * a.head_nc + k];
        s_head[i] = make_float4(hv[0], hv[1], hv[2], hv[3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int acc_stride = g.acc_stages == 2 ? g.tmem_cols / 2 : 0;

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
  if (warp == 0) {
    // ===== TMA producer (whole warp runs the loops; one elected lane issues) =====
    if (RES) {  // all weights of the layer stay in shared memory for the CTA's lifetime
      if (elect_one()) {
        mbar_expect_tx(fullB, g.b_tx * TAPS * nchunks);
        for (int c = 0; c < nchunks; ++c) {
          const int kglob = c < g.nchunk0 ? c * g.kc : a.C0 + (c - g.nchunk0) * g.kc;
          for (int tap = 0; tap < TAPS; ++tap)
            tma_load_2d(b_base + (c * TAPS + tap) * g.b_bytes, &tmB, fullB, kglob, tap * g.ncols_total);
        }
      }
      __syncwarp();
    }
    Ring ra, rb;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const int ns = it % nsplit, t = it / nsplit;
      const int tw = t % g.tiles_w, th = (t / g.tiles_w) % g.tiles_h, n = t / tiles_per_img;
      const int w0 = tw * 16, h0 = th * 16, n0 = ns * g.ncta;
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(emptyA + 8 * ra.idx, ra.phase ^ 1u);
        const bool first = c < g.nchunk0;
        if (elect_one()) {
          mbar_expect_tx(fullA + 8 * ra.idx, g.a_tx);
          tma_load_4d(a_base + ra.idx * g.a_bytes, first ? &tmA0 : &tmA1, fullA + 8 * ra.idx,
                      (first ? c : c - g.nchunk0) * g.kc, w0 - PAD, h0 - PAD, n);
        }
        __syncwarp();
        ra.advance(g.na);
        if (!RES) {
          const int kglob = first ? c * g.kc : a.C0 + (c - g.nchunk0) * g.kc;
          for (int tap = 0; tap < TAPS; ++tap) {
            mbar_wait(emptyB + 8 * rb.idx, rb.phase ^ 1u);
            if (elect_one()) {
              mbar_expect_tx(fullB + 8 * rb.idx, g.b_tx);
              tma_load_2d(b_base + rb.idx * g.b_bytes, &tmB, fullB + 8 * rb.idx, kglob, tap * g.ncols_total + n0);
            }
            __syncwarp();
            rb.advance(g.nb);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (whole warp converged, one elected lane issues tcgen05.mma / commit) =====
    // Descriptors: the high word (SBO, version, swizzle mode) is constant; per MMA only the 14-bit start
    // address field of the low word moves, in 16-byte units.
    const uint32_t idesc = make_idesc(g.ncta);
    const uint32_t hiA = ((uint32_t)(HT * rowb) >> 4) | (1u << 14) | (g.layout << 29);
    const uint32_t hiB = ((uint32_t)(8 * rowb) >> 4) | (1u << 14) | (g.layout << 29);
    constexpr uint32_t LBO1 = 1u << 16;
    constexpr uint32_t ROW16 = rowb / 16;  // one pixel row in descriptor units
    Ring ra, rb;
    int k_it = 0;
    if (RES) mbar_wait(fullB, 0);
    const uint32_t b_step = g.b_bytes >> 4;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++k_it) {
      const int as = g.acc_stages == 2 ? (k_it & 1) : 0;
      const uint32_t use = g.acc_stages == 2 ? (uint32_t)(k_it >> 1) : (uint32_t)k_it;
      mbar_wait(accEmpty + 8 * as, (use & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d0 = tmem_base + as * acc_stride, d1 = d0 + g.ncta;
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(fullA + 8 * ra.idx, ra.phase);
        tc_fence_after();
        const uint32_t a_lo = (((a_base + ra.idx * g.a_bytes) & 0x3FFFFu) >> 4) | LBO1;
        const uint32_t accc = (uint32_t)(c != 0);
        if (RES) {
          // weights resident: the whole chunk (TAPS x 2 strips x KSTEPS MMAs) is issued back to back
          const uint32_t b_lo0 = (((b_base + c * TAPS * g.b_bytes) & 0x3FFFFu) >> 4) | LBO1;
          if (elect_one()) {
#pragma unroll
            for (int tap = 0; tap < TAPS; ++tap) {
              const uint32_t at = a_lo + (uint32_t)(((tap / KS) * HT + (tap % KS)) * ROW16);
              const uint32_t b_lo = b_lo0 + tap * b_step;
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k) {
                const uint32_t acc = (tap | k) != 0 ? 1u : accc;
                tc_mma(d0, ((uint64_t)hiA << 32) | (at + 2 * k), ((uint64_t)hiB << 32) | (b_lo + 2 * k), idesc, acc);
                tc_mma(d1, ((uint64_t)hiA << 32) | (at + 8 * ROW16 + 2 * k), ((uint64_t)hiB << 32) | (b_lo + 2 * k),
                       idesc, acc);
              }
            }
            tc_commit(emptyA + 8 * ra.idx);
          }
          __syncwarp();
        } else {
#pragma unroll 1
          for (int tap = 0; tap < TAPS; ++tap) {
            mbar_wait(fullB + 8 * rb.idx, rb.phase);
            tc_fence_after();
            const uint32_t b_lo = (((b_base + rb.idx * g.b_bytes) & 0x3FFFFu) >> 4) | LBO1;
            const int dy = tap / KS, dx = tap - dy * KS;
            const uint32_t at = a_lo + (uint32_t)((dy * HT + dx) * ROW16);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k) {
                const uint32_t acc = k != 0 ? 1u : (tap != 0 ? 1u : accc);
                tc_mma(d0, ((uint64_t)hiA << 32) | (at + 2 * k), ((uint64_t)hiB << 32) | (b_lo + 2 * k), idesc, acc);
                tc_mma(d1, ((uint64_t)hiA << 32) | (at + 8 * ROW16 + 2 * k), ((uint64_t)hiB << 32) | (b_lo + 2 * k),
                       idesc, acc);
              }
              tc_commit(emptyB + 8 * rb.idx);
              if (tap == TAPS - 1) tc_commit(emptyA + 8 * ra.idx);
            }
            __syncwarp();
            rb.advance(g.nb);
          }
        }
        ra.advance(g.na);
      }
      if (elect_one()) tc_commit(accFull + 8 * as);
      __syncwarp();
    }
  }
  } else {
    // ===== epilogue warps: TMEM -> registers -> global =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    constexpr bool E_RES = (EPI & 1) != 0, E_AM = (EPI & 2) != 0;
    const int ew = warp - 4;
    const int strip = ew >> 2;
    const int q = warp & 3;          // TMEM lane quarter this warp may read
    const int r = q * 32 + lane;     // accumulator row
    const int ty = r >> 3, tx = strip * 8 + (r & 7);
    const int Cout = a.Cout;
    const int nj = g.ncta / 16;
    const bool has_film = E_RES && a.film_g != nullptr;
    const bool has_add = E_AM && a.add_src != nullptr;
    const bool has_mask = E_AM && a.mask_src != nullptr;
    const int step = (int)gridDim.x;

    // element offset of (item, this thread's pixel, first column of the item); side inputs never come with deconv
    auto item_base = [&](int it_) -> size_t {
      const int ns_ = it_ % nsplit, t_ = it_ / nsplit;
      const int tw_ = t_ % g.tiles_w, th_ = (t_ / g.tiles_w) % g.tiles_h, n_ = t_ / tiles_per_img;
      return (((size_t)n_ * a.H + th_ * 16 + ty) * a.W + tw_ * 16 + tx) * Cout + ns_ * g.ncta;
    };

    // ---- side-input ring: chunks (it_pf, j_pf) .. are in flight, EPI_PF ahead of the chunk being processed ----
    Packed16 ring_res[EPI_PF], ring_add[EPI_PF], ring_mk[EPI_PF];
    int it_pf = blockIdx.x, j_pf = 0;
    size_t base_pf = 0;
    if (EPI != 0 && it_pf < n_items) base_pf = item_base(it_pf);
#define DG_EPI_ISSUE(S_)                                                        \
  do {                                                                          \
    if (EPI != 0 && it_pf < n_items) {                                          \
      const size_t off_ = base_pf + (size_t)j_pf * 16;                          \
      if (has_film) ring_res[S_].load(a.res, off_);                             \
      if (has_add) ring_add[S_].load(a.add_src, off_);                          \
      if (has_mask) ring_mk[S_].load(a.mask_src, off_);                         \
      if (++j_pf == nj) {                                                       \
        j_pf = 0;                                                               \
        it_pf += step;                                                          \
        if (it_pf < n_items) base_pf = item_base(it_pf);                        \
      }                                                                         \
    }                                                                           \
  } while (0)
#pragma unroll
    for (int s = 0; s < EPI_PF; ++s) DG_EPI_ISSUE(s);

    // FiLM (gamma, beta) of the NEXT item's sample for this thread's table column, fetched one item ahead
    const int fc = ew * 32 + lane;  // column of the per-item FiLM table this thread fills (ncta <= 256)
    float fg_next = 0.f, fb_next = 0.f;
    auto film_fetch = [&](int it_) {
      if (has_film && it_ < n_items && fc < g.ncta) {
        const int ns_ = it_ % nsplit, n_ = (it_ / nsplit) / tiles_per_img;
        fg_next = __ldg(a.film_g + (size_t)n_ * a.film_stride + ns_ * g.ncta + fc);
        fb_next = __ldg(a.film_b + (size_t)n_ * a.film_stride + ns_ * g.ncta + fc);
      }
    };
    film_fetch(blockIdx.x);

    // ---- per-item state ----
    int it = blockIdx.x, j = 0, k_it = 0;
    int n = 0, h = 0, w = 0, n0 = 0, as = 0;
    size_t pix0 = 0;
    uint32_t t_row = 0;
    float head_acc[4] = {0.f, 0.f, 0.f, 0.f};
    float* sF = s_film;
    uint32_t vn[16];

    while (it < n_items) {
#pragma unroll
      for (int s = 0; s < EPI_PF; ++s) {
        if (it < n_items) {
          if (j == 0) {
            // ---- item start ----
            const int ns = it % nsplit, t = it / nsplit;
            const int tw = t % g.tiles_w, th = (t / g.tiles_w) % g.tiles_h;
            n = t / tiles_per_img;
            h = th * 16 + ty; w = tw * 16 + tx; n0 = ns * g.ncta;
            pix0 = ((size_t)n * a.H + h) * a.W + w;
            as = g.acc_stages == 2 ? (k_it & 1) : 0;
            const uint32_t use = g.acc_stages == 2 ? (uint32_t)(k_it >> 1) : (uint32_t)k_it;
            if (has_film) {
              // FiLM folded into the BN affine for this item's sample: v = relu(acc*(s*g) + (t*g + b)) + res
              sF = s_film + (k_it & 1) * 2 * g.ncta;
              if (fc < g.ncta) {
                sF[fc] = s_scale[n0 + fc] * fg_next;
                sF[g.ncta + fc] = fmaf(s_shift[n0 + fc], fg_next, fb_next);
              }
              // all 8 epilogue warps: the slot is complete, and every warp has left the item that last used it
              asm volatile("bar.sync 1, 256;" ::: "memory");
              film_fetch(it + step);
            }
            head_acc[0] = head_acc[1] = head_acc[2] = head_acc[3] = 0.f;
            mbar_wait(accFull + 8 * as, use & 1u);
            tc_fence_after();
            t_row = tmem_base + as * acc_stride + ((uint32_t)(q * 32) << 16) + (uint32_t)(strip * g.ncta);
            tc_ld16_issue(t_row, vn);
          }
          // this chunk's side inputs leave the ring; the slot is refilled EPI_PF chunks ahead
          Packed16 rs, ad, mk;
          if (has_film) rs = ring_res[s];
          if (has_add) ad = ring_add[s];
          if (has_mask) mk = ring_mk[s];
          DG_EPI_ISSUE(s);
          // accumulator chunk j has landed; start chunk j+1's TMEM read before working on j
          tc_ld_wait(vn);
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(vn[i]);
          if (j + 1 < nj) tc_ld16_issue(t_row + (uint32_t)((j + 1) * 16), vn);

          const int col = n0 + j * 16;
          int c0 = col;
          size_t opix = pix0;
          if (a.deconv) {
            const int ab = col / Cout;
            c0 = col - ab * Cout;
            opix = ((size_t)n * 2 * a.H + 2 * h + (ab >> 1)) * (2 * a.W) + 2 * w + (ab & 1);
          }
          const size_t off = opix * Cout + c0;
          if (has_film) {
            if (a.out_pre) {
              float vp[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) vp[i] = fmaf(v[i], s_scale[col + i], s_shift[col + i]);
              st16_bf16(a.out_pre, off, vp);
            }
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const float4 sc = reinterpret_cast<const float4*>(sF + j * 16)[i4];
              const float4 sh = reinterpret_cast<const float4*>(sF + g.ncta + j * 16)[i4];
              v[4 * i4 + 0] = fmaxf(fmaf(v[4 * i4 + 0], sc.x, sh.x), 0.f) + rs.get(4 * i4 + 0);
              v[4 * i4 + 1] = fmaxf(fmaf(v[4 * i4 + 1], sc.y, sh.y), 0.f) + rs.get(4 * i4 + 1);
              v[4 * i4 + 2] = fmaxf(fmaf(v[4 * i4 + 2], sc.z, sh.z), 0.f) + rs.get(4 * i4 + 2);
              v[4 * i4 + 3] = fmaxf(fmaf(v[4 * i4 + 3], sc.w, sh.w), 0.f) + rs.get(4 * i4 + 3);
            }
          } else {
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const float4 sc = reinterpret_cast<const float4*>(s_scale + col)[i4];
              const float4 sh = reinterpret_cast<const float4*>(s_shift + col)[i4];
              v[4 * i4 + 0] = fmaf(v[4 * i4 + 0], sc.x, sh.x);
              v[4 * i4 + 1] = fmaf(v[4 * i4 + 1], sc.y, sh.y);
              v[4 * i4 + 2] = fmaf(v[4 * i4 + 2], sc.z, sh.z);
              v[4 * i4 + 3] = fmaf(v[4 * i4 + 3], sc.w, sh.w);
            }
            if (a.out_pre) st16_bf16(a.out_pre, off, v);
          }
          if (has_add) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += ad.get(i);
          }
          if (has_mask) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = mk.get(i) > 0.f ? v[i] : 0.f;
          }
          if (a.relu) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          if (a.out) st16_bf16(a.out, off, v);
          if (a.head_w) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float4 hw = s_head[c0 + i];
              head_acc[0] = fmaf(v[i], hw.x, head_acc[0]);
              head_acc[1] = fmaf(v[i], hw.y, head_acc[1]);
              head_acc[2] = fmaf(v[i], hw.z, head_acc[2]);
              head_acc[3] = fmaf(v[i], hw.w, head_acc[3]);
            }
          }

          if (++j == nj) {
            // ---- item end: hand the accumulator stage back to the MMA issuer, write the fused head ----
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(accEmpty + 8 * as);
            if (a.head_w) {
              const int nc = a.head_nc;
              float o0 = head_acc[0] + __ldg(a.head_b), o1 = 0.f, o2 = 0.f, o3 = 0.f;
              if (nc > 1) o1 = head_acc[1] + __ldg(a.head_b + 1);
              if (nc > 2) o2 = head_acc[2] + __ldg(a.head_b + 2);
              if (nc > 3) o3 = head_acc[3] + __ldg(a.head_b + 3);
              if (a.head_act == 0) {
                o0 = tanhf(o0); o1 = tanhf(o1); o2 = tanhf(o2); o3 = tanhf(o3);
              } else if (a.head_act == 1) {
                float m = o0;
                if (nc > 1) m = fmaxf(m, o1);
                if (nc > 2) m = fmaxf(m, o2);
                if (nc > 3) m = fmaxf(m, o3);
                o0 = expf(o0 - m);
                o1 = nc > 1 ? expf(o1 - m) : 0.f;
                o2 = nc > 2 ? expf(o2 - m) : 0.f;
                o3 = nc > 3 ? expf(o3 - m) : 0.f;
                const float inv = 1.0f / (o0 + o1 + o2 + o3);
                o0 *= inv; o1 *= inv; o2 *= inv; o3 *= inv;
              }
              if (nc == 4) {
                *reinterpret_cast<float4*>(a.head_out + pix0 * 4) = make_float4(o0, o1, o2, o3);
              } else {
                float* op = a.head_out + pix0 * nc;
                op[0] = o0;
                if (nc > 1) op[1] = o1;
                if (nc > 2) op[2] = o2;
              }
            }
            j = 0;
            it += step;
            ++k_it;
          }
        }
      }
    }
#undef DG_EPI_ISSUE
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(g.tmem_cols) : "memory");
  }
}

// ---- instantiation helpers used by conv_tc_k{1,3,5}.cu ----
template <int KS, int KSTEPS, bool RES, int EPI>
int launch_one(int grid, uint32_t smem, cudaStream_t st, const CUtensorMap& tmA0, const CUtensorMap& tmA1,
               const CUtensorMap& tmB, const ConvArgs& a, const TcGeom& g) {
  conv_tc_kernel<KS, KSTEPS, RES, EPI><<<grid, TC_THREADS, smem, st>>>(tmA0, tmA1, tmB, a, g);
  DG_LAUNCH_CHECK();
  return 0;
}
template <int KS, int KSTEPS, bool RES, int EPI>
int set_attr_one() {
  DG_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<KS, KSTEPS, RES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024));
  return 0;
}
#endif  // __CUDACC__

}  // namespace convtc
