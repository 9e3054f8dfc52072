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
//   * B (weights, [tap][n][k] bf16): when every (chunk, tap) tile of the layer fits next to the activation ring
//     (32->32 ... 96->96 3x3 layers) the weights are loaded ONCE per CTA and stay resident (template RES); otherwise
//     they stream through their own TMA ring whose stage holds all taps of a channel chunk (or one kernel row, or
//     one tap, whichever fits), so the issuer waits once per stage.
//   * Persistent: grid = min(#work items, #SMs), one CTA per SM looping over (tile, n-split) items.
//     Control warps: TMA producer of the A/B rings, TMA producer of the epilogue side inputs, and two MMA issuers
//     (tcgen05.mma.cta_group::1.kind::f16), one per TMEM accumulator stage (4*ncta <= 512 columns), so the barrier
//     round trips between two items of one issuer are covered by the other issuer's MMA stream.
//   * Epilogue (8 warps, thread = one pixel = one TMEM lane; template EPI = which side inputs exist: 1 FiLM residual,
//     2 add / mask sources) never touches global memory with per-thread accesses: side inputs arrive as TMA tiles
//     [16][16][ch] in a 2-stage ring, results are packed to bf16 into a swizzled staging tile and leave by one TMA
//     store per ch-channel chunk (5-D map for the transposed conv's 2x2 scatter), double-buffered with bulk-group
//     waits.  The TMEM stage is handed back to the issuer as soon as its last column is in registers.  (The bf16
//     pre-FiLM copy `out_pre` of the training forward is the one per-thread global store left.)
//   * What bounds it: for N < 128 the 128 B/clk shared-memory port (A and B slices are re-read by every SS-mode MMA;
//     TMA fills and the epilogue staging share the port), not the tensor pipe -- see DESIGN.md section 4.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace convtc {

struct TcGeom {
  int tiles_w, tiles_h;
  int nchunk0, nchunk1;  // channel chunks taken from in0 / in1
  int nchunk2, nchunk3;  // split-half storage: the hi halves of in0 / in1 once more (the x_hi * w_lo term); else 0
  int kofs1, kofs2, kofs3;  // first weight column (K index) of the chunks of segments 1..3 (segment 0 starts at 0).  A
                            // segment's last chunk may run past its tensor's channels (C not a multiple of kc): TMA fills
                            // the missing activations with zeros, so whatever weights sit under them contribute nothing
  int kc;                // channels per chunk
  int ncols_total;       // weight rows per tap (Cout, or 4*Cout for the transposed conv)
  int ncta;              // output columns per CTA
  int tmem_cols;         // power of two >= 2*ncta*acc_stages
  int na, nb;            // ring depths (weights: stages of b_tps taps)
  int b_tps;             // taps per weight-ring stage when the weights stream: TAPS, KS or 1
  uint32_t a_bytes, b_bytes;  // stage strides (1024-aligned)
  uint32_t a_tx, b_tx;        // TMA transaction bytes per stage
  uint32_t layout;            // UMMA smem layout type (2 = SW128, 4 = SW64, 6 = SW32)
  int b_resident;             // 1: all (chunk, tap) weight tiles live in smem for the CTA's lifetime
  int acc_stages;             // TMEM accumulator stages (2 when 4*ncta <= 512)
  int n_issuers;              // MMA issuer warps (2: one per accumulator stage)
  int ch;                     // channels per epilogue staging chunk (16 / 32 / 64 = one 32/64/128-byte swizzle span)
  int n_side;                 // side-input tensors TMA-loaded per chunk (FiLM residual; add and/or mask sources)
  int stage_out;              // 1: `out` is written tile-wise from shared memory by TMA stores
  uint32_t slot_bytes;        // one staging tile: 256 pixels x ch x 2 bytes
  int pool;                   // 1: the 2x2 max-pooled tile is staged and stored as well (EPI bit 2)
  int split;                  // 1: split-half storage (DT_F16S): K = 3*Cin, hi / lo staging tiles, two stores per chunk
};

// tensor maps of one launch: activations (two concatenated sources), weights, output, epilogue side inputs
struct TcMaps {
  CUtensorMap a0, a1, b, out, s0, s1, pool;
  CUtensorMap out2;  // split-half storage: the lo half of the output (out = the hi half)
};

// one launcher per kernel size, defined in conv_tc_k{1,3,5}.cu; epi = bit mask of the side inputs the call uses
int launch_ks1(int grid, uint32_t smem, cudaStream_t st, const TcMaps& tm, const ConvArgs& a, const TcGeom& g, int epi);
int launch_ks3(int grid, uint32_t smem, cudaStream_t st, const TcMaps& tm, const ConvArgs& a, const TcGeom& g, int epi);
int launch_ks5(int grid, uint32_t smem, cudaStream_t st, const TcMaps& tm, const ConvArgs& a, const TcGeom& g, int epi);
// split-half storage (conv_tc_split.cu): 1x1 plain (the transposed conv) and 3x3 plain / FiLM
int launch_split(int grid, uint32_t smem, cudaStream_t st, const TcMaps& tm, const ConvArgs& a, const TcGeom& g, int epi);
int set_attrs_split();
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
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(tm), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2, int c3,
                                             int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"(tm), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
#ifdef DG_DBG_NOMMA  // timing experiment only (scripts/build_variant.sh): results are garbage
  return;
#endif
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

// waits for every tcgen05.ld in flight; names all four register groups so no use can be scheduled above it
__device__ __forceinline__ void tc_ld_wait4(uint32_t (&r)[4][16]) {
  tc_ld_wait(r[0]);
  asm volatile("" : "+r"(r[1][0]), "+r"(r[1][1]), "+r"(r[1][2]), "+r"(r[1][3]), "+r"(r[1][4]), "+r"(r[1][5]),
                    "+r"(r[1][6]), "+r"(r[1][7]), "+r"(r[1][8]), "+r"(r[1][9]), "+r"(r[1][10]), "+r"(r[1][11]),
                    "+r"(r[1][12]), "+r"(r[1][13]), "+r"(r[1][14]), "+r"(r[1][15]) :: "memory");
  asm volatile("" : "+r"(r[2][0]), "+r"(r[2][1]), "+r"(r[2][2]), "+r"(r[2][3]), "+r"(r[2][4]), "+r"(r[2][5]),
                    "+r"(r[2][6]), "+r"(r[2][7]), "+r"(r[2][8]), "+r"(r[2][9]), "+r"(r[2][10]), "+r"(r[2][11]),
                    "+r"(r[2][12]), "+r"(r[2][13]), "+r"(r[2][14]), "+r"(r[2][15]) :: "memory");
  asm volatile("" : "+r"(r[3][0]), "+r"(r[3][1]), "+r"(r[3][2]), "+r"(r[3][3]), "+r"(r[3][4]), "+r"(r[3][5]),
                    "+r"(r[3][6]), "+r"(r[3][7]), "+r"(r[3][8]), "+r"(r[3][9]), "+r"(r[3][10]), "+r"(r[3][11]),
                    "+r"(r[3][12]), "+r"(r[3][13]), "+r"(r[3][14]), "+r"(r[3][15]) :: "memory");
}

// kind::f16 instruction descriptor: D=f32, A=B=bf16 (format 1) or IEEE half (format 0), both K-major, M=128.
__device__ __forceinline__ uint32_t make_idesc(int n, bool f16 = false) {
  const uint32_t ab = f16 ? 0u : ((1u << 7) | (1u << 10));
  return (1u << 4) | ab | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

struct Packed16 {  // 16 half-precision values as loaded (two 128-bit words); get() decodes bf16, unpack() either format
  uint4 q[2];
  __device__ __forceinline__ void unpack(float (&s)[16], bool f16) const {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(q);
    if (f16) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        s[2 * i] = f.x; s[2 * i + 1] = f.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[2 * i] = __uint_as_float(w[i] << 16); s[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
      }
    }
  }
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
__device__ __forceinline__ void st16_bf16(void* base, size_t elem_off, const float (&v)[16], bool f16 = false) {
#ifdef DG_DBG_NOSTORE  // timing experiment only: keeps the values alive without a global store
  if (v[0] == 1.2345e-30f && v[7] == 5.4321e-30f) reinterpret_cast<float*>(base)[0] = v[3];
  return;
#endif
  uint4 q[2];
  uint32_t* h = reinterpret_cast<uint32_t*>(q);
#pragma unroll
  for (int i = 0; i < 8; ++i) h[i] = pack_h2(v[2 * i], v[2 * i + 1], f16);
  uint4* p = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(base) + elem_off);
  p[0] = q[0];
  p[1] = q[1];
}

#ifdef DG_DBG_TRACE  // timing experiment only (scripts/build_variant.sh): per-role event clocks of CTA 0
__device__ long long* g_dbg_trace = nullptr;  // [role 0..2][item 0..39][event 0..7], dumped from shared memory at exit
#define DG_TRACE_DECL __shared__ uint32_t s_trace[3 * 40 * 8];
#define DG_TRACE_SMEM 4096
#define DG_TRACE(role_, item_, ev_)                                                                       \
  do {                                                                                                    \
    if (blockIdx.x == 0 && (item_) < 40 && (threadIdx.x & 31) == 0)                                       \
      s_trace[((role_) * 40 + (item_)) * 8 + (ev_)] = (uint32_t)clock64();                                \
  } while (0)
#define DG_TRACE_DUMP                                                                                     \
  do {                                                                                                    \
    if (blockIdx.x == 0 && g_dbg_trace)                                                                   \
      for (int i_ = threadIdx.x; i_ < 3 * 40 * 8; i_ += blockDim.x) g_dbg_trace[i_] = s_trace[i_];        \
  } while (0)
#else
#define DG_TRACE_DECL
#define DG_TRACE_SMEM 0
#define DG_TRACE(role_, item_, ev_) do {} while (0)
#define DG_TRACE_DUMP do {} while (0)
#endif

constexpr int TC_THREADS = 384;  // two epilogue warpgroups (warps 0..7) + one control warpgroup (warps 8..11)
// The warp scheduler favours the highest warp id of a sub-partition, so the latency-critical single-warp roles sit in
// the LAST warpgroup and the instruction-heavy epilogue in warps 0..7.
constexpr int CTRL_W0 = 8;  // first control warp
constexpr int EPI_W0 = 0;   // first epilogue warp

// Work item index -> (n-split, tile column, tile row, slice), stepped by the grid size with mixed-radix carries so the
// per-item path has no integer division.
struct ItemIter {
  int it, ns, tw, th, n;
  int d_it, d_ns, d_tw, d_th, d_n;
  int r_ns, r_tw, r_th;
  __device__ __forceinline__ void init(int it0, int step, int nsplit, int tiles_w, int tiles_h) {
    r_ns = nsplit; r_tw = tiles_w; r_th = tiles_h;
    it = it0; d_it = step;
    ns = it0 % nsplit; int t = it0 / nsplit;
    tw = t % tiles_w; t /= tiles_w;
    th = t % tiles_h; n = t / tiles_h;
    d_ns = step % nsplit; t = step / nsplit;
    d_tw = t % tiles_w; t /= tiles_w;
    d_th = t % tiles_h; d_n = t / tiles_h;
  }
  __device__ __forceinline__ void advance() {
    it += d_it;
    ns += d_ns;
    int c = ns >= r_ns ? 1 : 0;
    ns -= c ? r_ns : 0;
    tw += d_tw + c;
    c = tw >= r_tw ? 1 : 0;
    tw -= c ? r_tw : 0;
    th += d_th + c;
    c = th >= r_th ? 1 : 0;
    th -= c ? r_th : 0;
    n += d_n + c;
  }
};

struct Ring {
  int idx = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance(int n) {
    if (++idx == n) { idx = 0; phase ^= 1u; }
  }
  __device__ __forceinline__ void skip(int k, int n) {  // k stages ahead
    idx += k;
    while (idx >= n) { idx -= n; phase ^= 1u; }
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
// kernel: persistent, warp-specialised, 12 warps.
//   control warpgroup (warps 8..11): 8 = TMA producer of the A / B rings, 9 = MMA issuer 0 + TMEM owner,
//     10 = TMA producer of the epilogue's side-input tiles, 11 = MMA issuer 1.
//   epilogue warps 0..7 (warp w reads TMEM lane quarter w%4 of strip w/4), running with the registers the control
//     warpgroup gave up.
// ---------------------------------------------------------------------------------------------------------
// F16: the 16-bit storage format of activations and weights is IEEE half instead of bfloat16 (generator inference handles,
// DEPGAN_PREC_F16); a template parameter, because a run-time flag doubles the pack / unpack instructions of the epilogue.
// SPLIT (implies F16): split-half storage, see DT_F16S in common.cuh.  The K loop runs over [in0 hi|lo], [in1 hi|lo],
// [in0 hi], [in1 hi] against weight rows packed in that order; the epilogue stages hi = half(v) and lo = half(v - hi) as
// two tiles and stores them to the two channel halves of the output pixel; the FiLM residual arrives as a hi and a lo tile.
template <int KS, int KSTEPS, bool RES, int EPI, bool F16 = false, bool SPLIT = false>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ TcMaps tm, const ConvArgs a,
                                                                const TcGeom g) {
  constexpr int PAD = KS / 2, HT = 16 + KS - 1, TAPS = KS * KS;
  extern __shared__ uint8_t smem_raw[];
  DG_TRACE_DECL
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t a_base = base;
  const uint32_t b_base = a_base + g.na * g.a_bytes;
  const uint32_t o_base = b_base + g.nb * g.b_tps * g.b_bytes;             // 2 output staging slots
  const uint32_t p_base = o_base + (g.stage_out ? (SPLIT ? 4u : 2u) * g.slot_bytes : 0u);  // 2 pooled staging slots (64 pixels)
  const uint32_t s_base = p_base + (g.pool ? g.slot_bytes / 2u : 0u);       // 2 side stages x n_side slots
  const uint32_t bar_base = s_base + 2u * g.n_side * g.slot_bytes;
  const uint32_t fullA = bar_base, emptyA = fullA + 8 * g.na;
  const uint32_t fullB = emptyA + 8 * g.na, emptyB = fullB + 8 * g.nb;
  const uint32_t accFull = emptyB + 8 * g.nb, accEmpty = accFull + 16;
  const uint32_t sideFull = accEmpty + 16, sideEmpty = sideFull + 16;
  const uint32_t tmem_slot = sideEmpty + 16;
  const uint32_t ss_off = tmem_slot + 16;  // scale / shift staging: 2 * ncols_total floats, then head weights
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  float* s_scale = reinterpret_cast<float*>(smem_raw + (ss_off - raw));
  float* s_shift = s_scale + g.ncols_total;
  float4* s_head = reinterpret_cast<float4*>(s_shift + g.ncols_total);  // [Cout] x (up to 4 head outputs)
  // FiLM folded with BN per (sample, channel): [3 slots][2][ncta] floats, slot k%3 belongs to the CTA's k-th item
  float* s_film = reinterpret_cast<float*>(s_head + (a.head_w ? a.Cout : 0));

  // warp index made provably warp-uniform so the role loops run on the uniform datapath
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  constexpr int rowb = KSTEPS * 32;  // bytes per pixel row of a chunk = one swizzle span (kc = 16*KSTEPS)
  const int nchunks = g.nchunk0 + g.nchunk1 + g.nchunk2 + g.nchunk3;
  const int nsplit = g.ncols_total / g.ncta;
  const int tiles_per_img = g.tiles_w * g.tiles_h;
  const int n_items = tiles_per_img * a.N * nsplit;
  const int nco = g.ncta / g.ch;  // staging chunks per item

  if (threadIdx.x == 0) {
    for (int i = 0; i < g.na; ++i) { mbar_init(fullA + 8 * i, 1); mbar_init(emptyA + 8 * i, 1); }
    for (int i = 0; i < g.nb; ++i) { mbar_init(fullB + 8 * i, 1); mbar_init(emptyB + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(accFull + 8 * i, 1); mbar_init(accEmpty + 8 * i, 8);
      mbar_init(sideFull + 8 * i, 1); mbar_init(sideEmpty + 8 * i, 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == CTRL_W0 + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(g.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // Programmatic dependent launch: the barrier / TMEM set-up above overlaps the previous kernel's tail.  Everything
  // that touches global memory comes after the wait (= the previous grid has completed and flushed).  The trigger
  // is only honoured once every CTA of this grid has issued it, i.e. is resident, so chains of launches cannot
  // occupy SMs ahead of an unfinished predecessor.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
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

  if (warp >= CTRL_W0) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
  if (warp == CTRL_W0) {
    // ===== TMA producer of the operand rings (whole warp runs the loops; one elected lane issues) =====
    if (RES) {  // all weights of the layer stay in shared memory for the CTA's lifetime
      if (elect_one()) {
        mbar_expect_tx(fullB, g.b_tx * TAPS * nchunks);
        for (int c = 0; c < nchunks; ++c) {
          int cs = c, kglob = 0;
          if (cs >= g.nchunk0) {
            cs -= g.nchunk0; kglob = g.kofs1;
            if (cs >= g.nchunk1) {
              cs -= g.nchunk1; kglob = g.kofs2;
              if (cs >= g.nchunk2) { cs -= g.nchunk2; kglob = g.kofs3; }
            }
          }
          kglob += cs * g.kc;
          for (int tap = 0; tap < TAPS; ++tap)
            tma_load_2d(b_base + (c * TAPS + tap) * g.b_bytes, &tm.b, fullB, kglob, tap * g.ncols_total);
        }
      }
      __syncwarp();
    }
    Ring ra, rb;
    int p_it = 0;
    ItemIter ii;
    ii.init(blockIdx.x, (int)gridDim.x, nsplit, g.tiles_w, g.tiles_h);
    for (; ii.it < n_items; ii.advance(), ++p_it) {
      const int n = ii.n;
      const int w0 = ii.tw * 16, h0 = ii.th * 16, n0 = ii.ns * g.ncta;
      DG_TRACE(0, p_it, 0);
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(emptyA + 8 * ra.idx, ra.phase ^ 1u);
        if (c == 0) DG_TRACE(0, p_it, 1);
        // chunk -> (source, channel offset): in0, in1, then (split-half storage) the hi halves of in0 and in1 again
        int cs = c, kglob = 0;
        bool first = true;
        if (cs >= g.nchunk0) {
          cs -= g.nchunk0; first = false; kglob = g.kofs1;
          if (cs >= g.nchunk1) {
            cs -= g.nchunk1; first = true; kglob = g.kofs2;
            if (cs >= g.nchunk2) { cs -= g.nchunk2; first = false; kglob = g.kofs3; }
          }
        }
        kglob += cs * g.kc;
        if (elect_one()) {
          mbar_expect_tx(fullA + 8 * ra.idx, g.a_tx);
          tma_load_4d(a_base + ra.idx * g.a_bytes, first ? &tm.a0 : &tm.a1, fullA + 8 * ra.idx, cs * g.kc, w0 - PAD,
                      h0 - PAD, n);
        }
        __syncwarp();
        ra.advance(g.na);
        if (!RES) {
          // weight ring: one stage = b_tps consecutive taps of this channel chunk (all taps, one kernel row, or one tap)
          for (int tap = 0; tap < TAPS; tap += g.b_tps) {
            mbar_wait(emptyB + 8 * rb.idx, rb.phase ^ 1u);
            if (elect_one()) {
              mbar_expect_tx(fullB + 8 * rb.idx, g.b_tx * g.b_tps);
              for (int t2 = 0; t2 < g.b_tps; ++t2)
                tma_load_2d(b_base + (rb.idx * g.b_tps + t2) * g.b_bytes, &tm.b, fullB + 8 * rb.idx, kglob,
                            (tap + t2) * g.ncols_total + n0);
            }
            __syncwarp();
            rb.advance(g.nb);
          }
        }
      }
      DG_TRACE(0, p_it, 2);
    }
  } else if (warp == CTRL_W0 + 2) {
    // ===== TMA producer of the epilogue's side-input tiles (FiLM residual, or add / mask sources) =====
    if (EPI != 0 && g.n_side > 0) {
      Ring rs;
      ItemIter ii;
      ii.init(blockIdx.x, (int)gridDim.x, nsplit, g.tiles_w, g.tiles_h);
      for (; ii.it < n_items; ii.advance()) {
        const int ns = ii.ns, tw = ii.tw, th = ii.th, n = ii.n;
        for (int cc = 0; cc < nco; ++cc) {
          mbar_wait(sideEmpty + 8 * rs.idx, rs.phase ^ 1u);
          if (elect_one()) {
            const uint32_t dst = s_base + (uint32_t)(rs.idx * g.n_side) * g.slot_bytes;
            mbar_expect_tx(sideFull + 8 * rs.idx, (uint32_t)g.n_side * g.slot_bytes);
            tma_load_4d(dst, &tm.s0, sideFull + 8 * rs.idx, ns * g.ncta + cc * g.ch, tw * 16, th * 16, n);
            if (g.n_side == 2)
              tma_load_4d(dst + g.slot_bytes, &tm.s1, sideFull + 8 * rs.idx, ns * g.ncta + cc * g.ch, tw * 16, th * 16, n);
          }
          __syncwarp();
          rs.advance(2);
        }
      }
    }
  } else {
    // ===== MMA issuers (whole warp converged, one elected lane issues tcgen05.mma / commit) =====
    // Issuer m takes the CTA's items m, m + n_issuers, ... and owns accumulator stage m, so while one issuer sits in
    // the barrier waits between two of its items the other keeps the tensor pipe fed.  (Two issuers need
    // na >= 2 * nchunks, see plan(): the parity wait of one issuer must never race the other issuer's fill.)
    // Descriptors: the high word (SBO, version, swizzle mode) is constant; per MMA only the 14-bit start
    // address field of the low word moves, in 16-byte units.
    const int m = warp == CTRL_W0 + 1 ? 0 : 1;
    const int nI = g.n_issuers;
    if (m < nI) {
    const uint32_t idesc = make_idesc(g.ncta, F16);
    const uint32_t hiA = ((uint32_t)(HT * rowb) >> 4) | (1u << 14) | (g.layout << 29);
    const uint32_t hiB = ((uint32_t)(8 * rowb) >> 4) | (1u << 14) | (g.layout << 29);
    constexpr uint32_t LBO1 = 1u << 16;
    constexpr uint32_t ROW16 = rowb / 16;  // one pixel row in descriptor units
    Ring ra, rb;
    ra.skip(m * nchunks, g.na);
    if (RES) mbar_wait(fullB, 0);
    const uint32_t b_step = g.b_bytes >> 4;
    const int as = g.acc_stages == 2 ? (nI == 2 ? m : 0) : 0;
    int k_it = 0;  // this issuer's item counter
    for (int it = blockIdx.x + m * gridDim.x; it < n_items; it += nI * gridDim.x, ++k_it) {
      // single issuer with two stages alternates them; otherwise the stage is fixed
      const int st = (g.acc_stages == 2 && nI == 1) ? (k_it & 1) : as;
      const uint32_t use = (g.acc_stages == 2 && nI == 1) ? (uint32_t)(k_it >> 1) : (uint32_t)k_it;
      if (m == 0) DG_TRACE(1, k_it, 0);
      mbar_wait(accEmpty + 8 * st, (use & 1u) ^ 1u);
      tc_fence_after();
      if (m == 0) DG_TRACE(1, k_it, 1);
      const uint32_t d0 = tmem_base + st * acc_stride, d1 = d0 + g.ncta;
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(fullA + 8 * ra.idx, ra.phase);
        tc_fence_after();
        if (c == 0 && m == 0) DG_TRACE(1, k_it, 2);
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
        } else if (g.b_tps == TAPS) {
          // streamed weights, one ring stage = all taps of the chunk: one wait, then the chunk's TAPS x 2 strips x KSTEPS
          // MMAs back to back with immediate descriptor offsets, like the resident case.  (The tap-by-tap loop below costs
          // the issuing thread ~330 clk of scalar work per tap -- measured with the role tracer on the 160 -> 64 layer:
          // 14.9 k clk per item for 180 MMAs -- which is more than its four MMAs take.)
          mbar_wait(fullB + 8 * rb.idx, rb.phase);
          tc_fence_after();
          const uint32_t b_lo0 = (((b_base + rb.idx * TAPS * g.b_bytes) & 0x3FFFFu) >> 4) | LBO1;
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
            tc_commit(emptyB + 8 * rb.idx);
            tc_commit(emptyA + 8 * ra.idx);
          }
          __syncwarp();
          rb.advance(g.nb);
        } else if (g.b_tps == KS && KS > 1) {
          // streamed weights, one ring stage = one kernel row: one wait and one election per row, its KS taps unrolled
#pragma unroll 1
          for (int dy = 0; dy < KS; ++dy) {
            mbar_wait(fullB + 8 * rb.idx, rb.phase);
            tc_fence_after();
            const uint32_t b_lo0 = (((b_base + rb.idx * KS * g.b_bytes) & 0x3FFFFu) >> 4) | LBO1;
            const uint32_t a_row = a_lo + (uint32_t)(dy * HT * ROW16);
            if (elect_one()) {
#pragma unroll
              for (int dx = 0; dx < KS; ++dx) {
                const uint32_t at = a_row + (uint32_t)(dx * ROW16);
                const uint32_t b_lo = b_lo0 + dx * b_step;
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k) {
                  const uint32_t acc = (dx | k) != 0 ? 1u : (dy != 0 ? 1u : accc);
                  tc_mma(d0, ((uint64_t)hiA << 32) | (at + 2 * k), ((uint64_t)hiB << 32) | (b_lo + 2 * k), idesc, acc);
                  tc_mma(d1, ((uint64_t)hiA << 32) | (at + 8 * ROW16 + 2 * k), ((uint64_t)hiB << 32) | (b_lo + 2 * k),
                         idesc, acc);
                }
              }
              tc_commit(emptyB + 8 * rb.idx);
              if (dy == KS - 1) tc_commit(emptyA + 8 * ra.idx);
            }
            __syncwarp();
            rb.advance(g.nb);
          }
        } else {
          // streamed weights: a ring stage holds b_tps consecutive taps (KS or 1); taps of a kernel row are
          // unrolled so their descriptor offsets are immediates
          uint32_t b_lo0 = 0;
#pragma unroll 1
          for (int dy = 0; dy < KS; ++dy) {
#pragma unroll
            for (int dx = 0; dx < KS; ++dx) {
              const int tap = dy * KS + dx;
              const bool s_first = g.b_tps == 1 || (g.b_tps == KS ? dx == 0 : tap == 0);
              const bool s_last = g.b_tps == 1 || (g.b_tps == KS ? dx == KS - 1 : tap == TAPS - 1);
              if (s_first) {
                mbar_wait(fullB + 8 * rb.idx, rb.phase);
                tc_fence_after();
                b_lo0 = (((b_base + rb.idx * g.b_tps * g.b_bytes) & 0x3FFFFu) >> 4) | LBO1;
              }
              const int slot = g.b_tps == 1 ? 0 : (g.b_tps == KS ? dx : tap);
              const uint32_t b_lo = b_lo0 + (uint32_t)slot * b_step;
              const uint32_t at = a_lo + (uint32_t)(dy * HT * ROW16) + (uint32_t)(dx * ROW16);
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k) {
                  const uint32_t acc = k != 0 ? 1u : (tap != 0 ? 1u : accc);
                  tc_mma(d0, ((uint64_t)hiA << 32) | (at + 2 * k), ((uint64_t)hiB << 32) | (b_lo + 2 * k), idesc, acc);
                  tc_mma(d1, ((uint64_t)hiA << 32) | (at + 8 * ROW16 + 2 * k), ((uint64_t)hiB << 32) | (b_lo + 2 * k),
                         idesc, acc);
                }
                if (s_last) tc_commit(emptyB + 8 * rb.idx);
                if (tap == TAPS - 1) tc_commit(emptyA + 8 * ra.idx);
              }
              __syncwarp();
              if (s_last) rb.advance(g.nb);
            }
          }
        }
        ra.advance(g.na);
      }
      if (elect_one()) tc_commit(accFull + 8 * st);
      __syncwarp();
      if (m == 0) DG_TRACE(1, k_it, 3);
      if (nI == 2) ra.skip(nchunks, g.na);  // the other issuer's item
    }
    }
  }
  } else {
    // ===== epilogue warps: TMEM -> registers -> (staging tile in shared memory -> TMA store) =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    constexpr bool E_RES = (EPI & 1) != 0, E_AM = (EPI & 2) != 0, E_POOL = (EPI & 4) != 0;
    const int ew = warp - EPI_W0;
    const int strip = ew >> 2;
    const int q = warp & 3;          // TMEM lane quarter this warp may read
    const int r = q * 32 + lane;     // accumulator row
    const int ty = r >> 3, tx = strip * 8 + (r & 7);
    const int Cout = a.Cout;
    const int ng = g.ncta / 16;       // 16-column groups per item
    const int gpc = g.ch / 16;        // groups per staging chunk
    const bool has_film = E_RES && a.film_g != nullptr;
    const bool has_add = E_AM && a.add_src != nullptr;
    const bool has_mask = E_AM && a.mask_src != nullptr;
    const bool side = EPI != 0 && g.n_side > 0;
    const bool stage_out = g.stage_out != 0;
    const bool t0 = threadIdx.x == EPI_W0 * 32;  // issues the TMA stores
    constexpr bool f16 = F16;                    // 16-bit storage format of the activations (bf16 otherwise)
    // this thread's pixel in a staging tile [16 rows][16 cols][ch] whose 16-byte units are XOR-swizzled like the TMA
    // (unit j of the pixel at byte offset o lives at o + ((j ^ ((o >> 7) & (units-1))) << 4)): conflict-free 128-bit
    // accesses for 8 neighbouring pixels
    const uint32_t p_off = (uint32_t)(ty * 16 + tx) * (uint32_t)(g.ch * 2);
    const uint32_t p_xor = (p_off >> 7) & (uint32_t)(g.ch / 8 - 1);
    const uint32_t add_slot = 0, mask_slot = has_add ? g.slot_bytes : 0u;
    const int head_nc = a.head_w ? a.head_nc : 0;

    ItemIter cur, nx1, nx2;  // this item, the next one and the one after (FiLM tables are built ahead)
    cur.init(blockIdx.x, (int)gridDim.x, nsplit, g.tiles_w, g.tiles_h);
    nx1 = cur; nx1.advance();
    nx2 = nx1; nx2.advance();

    // FiLM (gamma, beta) for this thread's column of the per-item table, fetched two items ahead
    const int fc = ew * 32 + lane;  // table column this thread fills (ncta <= 256)
    float fg_next = 0.f, fb_next = 0.f;
    auto film_fetch = [&](const ItemIter& i_) {
      if (has_film && i_.it < n_items && fc < g.ncta) {
        fg_next = __ldg(a.film_g + (size_t)i_.n * a.film_stride + i_.ns * g.ncta + fc);
        fb_next = __ldg(a.film_b + (size_t)i_.n * a.film_stride + i_.ns * g.ncta + fc);
      }
    };
    auto film_fill = [&](const ItemIter& i_, int slot_) {  // v = relu(acc*(s*g) + (t*g + b)) + res
      if (has_film && i_.it < n_items && fc < g.ncta) {
        const int n0_ = i_.ns * g.ncta;
        float* sF_ = s_film + slot_ * 2 * g.ncta;
        sF_[fc] = s_scale[n0_ + fc] * fg_next;
        sF_[g.ncta + fc] = fmaf(s_shift[n0_ + fc], fg_next, fb_next);
      }
    };
    if (has_film) {
      film_fetch(cur);
      film_fill(cur, 0);
      film_fetch(nx1);
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }

    Ring rs;             // side-input stages
    uint32_t oslot = 0;  // output staging slot of the next chunk
    int k_it = 0, fslot = 0;
    uint32_t vb[4][16];  // up to 64 accumulator columns in flight
    for (; cur.it < n_items; ++k_it) {
      const int h = cur.th * 16 + ty, w = cur.tw * 16 + tx, n0 = cur.ns * g.ncta;
      const size_t pix0 = ((size_t)cur.n * a.H + h) * a.W + w;
      const int as = g.acc_stages == 2 ? (k_it & 1) : 0;
      const uint32_t use = g.acc_stages == 2 ? (uint32_t)(k_it >> 1) : (uint32_t)k_it;
      const float* sF = s_film + fslot * 2 * g.ncta;
      const int fnext = fslot == 2 ? 0 : fslot + 1;
      if (has_film) {
        // table of the NEXT item: its slot was last read two items ago, and every thread passes one of this item's
        // chunk barriers between this write and the reads
        film_fill(nx1, fnext);
        film_fetch(nx2);
      }
      float head_acc[4] = {0.f, 0.f, 0.f, 0.f};
      if (ew == 0) DG_TRACE(2, k_it, 0);
      mbar_wait(accFull + 8 * as, use & 1u);
      tc_fence_after();
      if (ew == 0) DG_TRACE(2, k_it, 1);
      const uint32_t t_row = tmem_base + as * acc_stride + ((uint32_t)(q * 32) << 16) + (uint32_t)(strip * g.ncta);
      int gg = 0, cc = 0;
      // one 16-column group: BN / FiLM / add / mask / ReLU, pack to bf16 into the staging tile, fused head; the last
      // group of a staging chunk hands the chunk to the TMA store
      auto do_group = [&](const uint32_t (&vr)[16], const int gi) {
        if (gg == 0 && side) mbar_wait(sideFull + 8 * rs.idx, rs.phase);
        if (ew == 0 && gi == 0) DG_TRACE(2, k_it, 7);
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(vr[i]);
        const int col = n0 + gi * 16;
        // the two 16-byte units of this group inside the thread's staging pixel
        const uint32_t u0 = p_off + ((((uint32_t)(2 * gg)) ^ p_xor) << 4);
        const uint32_t u1 = p_off + ((((uint32_t)(2 * gg + 1)) ^ p_xor) << 4);
        const uint8_t* sgen = smem_raw + (s_base + (uint32_t)(rs.idx * g.n_side) * g.slot_bytes - raw);
        if (has_film) {
          if (a.out_pre) {
            float vp[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) vp[i] = fmaf(v[i], s_scale[col + i], s_shift[col + i]);
            st16_bf16(a.out_pre, pix0 * Cout + col, vp, f16);
          }
          Packed16 rs16;
          rs16.q[0] = *reinterpret_cast<const uint4*>(sgen + u0);
          rs16.q[1] = *reinterpret_cast<const uint4*>(sgen + u1);
          float rsv[16];
          rs16.unpack(rsv, f16);
          if (SPLIT) {  // residual = hi + lo (second side slot)
            Packed16 rl16;
            rl16.q[0] = *reinterpret_cast<const uint4*>(sgen + g.slot_bytes + u0);
            rl16.q[1] = *reinterpret_cast<const uint4*>(sgen + g.slot_bytes + u1);
            float rlv[16];
            rl16.unpack(rlv, true);
#pragma unroll
            for (int i = 0; i < 16; ++i) rsv[i] += rlv[i];
          }
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            const float4 sc = reinterpret_cast<const float4*>(sF + gi * 16)[i4];
            const float4 sh = reinterpret_cast<const float4*>(sF + g.ncta + gi * 16)[i4];
            v[4 * i4 + 0] = fmaxf(fmaf(v[4 * i4 + 0], sc.x, sh.x), 0.f) + rsv[4 * i4 + 0];
            v[4 * i4 + 1] = fmaxf(fmaf(v[4 * i4 + 1], sc.y, sh.y), 0.f) + rsv[4 * i4 + 1];
            v[4 * i4 + 2] = fmaxf(fmaf(v[4 * i4 + 2], sc.z, sh.z), 0.f) + rsv[4 * i4 + 2];
            v[4 * i4 + 3] = fmaxf(fmaf(v[4 * i4 + 3], sc.w, sh.w), 0.f) + rsv[4 * i4 + 3];
          }
        } else {
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            const float4 sc = reinterpret_cast<const float4*>(s_scale + col)[i4];
            const float4 sh = reinterpret_cast<const float4*>(s_shift + col)[i4];
            fma_f32x2(v[4 * i4 + 0], v[4 * i4 + 1], sc.x, sc.y, sh.x, sh.y);
            fma_f32x2(v[4 * i4 + 2], v[4 * i4 + 3], sc.z, sc.w, sh.z, sh.w);
          }
          if (a.out_pre) st16_bf16(a.out_pre, pix0 * Cout + col, v, f16);
        }
        if (has_add) {
          Packed16 ad;
          ad.q[0] = *reinterpret_cast<const uint4*>(sgen + add_slot + u0);
          ad.q[1] = *reinterpret_cast<const uint4*>(sgen + add_slot + u1);
          float adv[16];
          ad.unpack(adv, f16);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += adv[i];
        }
        if (has_mask) {
          Packed16 mk;
          mk.q[0] = *reinterpret_cast<const uint4*>(sgen + mask_slot + u0);
          mk.q[1] = *reinterpret_cast<const uint4*>(sgen + mask_slot + u1);
          float mkv[16];
          mk.unpack(mkv, f16);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = mkv[i] > 0.f ? v[i] : 0.f;
        }
        // the ReLU rides on the 16-bit conversion unless the fp32 values are used again (fused head)
        const bool relu_cvt = a.relu && head_nc == 0 && !SPLIT;
        if (a.relu && !relu_cvt) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        if (stage_out) {
          uint4 pk[2];
          uint32_t* hp = reinterpret_cast<uint32_t*>(pk);
          if (relu_cvt) {
#pragma unroll
            for (int i = 0; i < 8; ++i) hp[i] = pack_h2_relu(v[2 * i], v[2 * i + 1], f16);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) hp[i] = pack_h2(v[2 * i], v[2 * i + 1], f16);
          }
          uint8_t* ogen = smem_raw + (o_base + (SPLIT ? 2u * oslot : oslot) * g.slot_bytes - raw);
          *reinterpret_cast<uint4*>(ogen + u0) = pk[0];
          *reinterpret_cast<uint4*>(ogen + u1) = pk[1];
          if (SPLIT) {  // lo = half(v - hi) into the slot after the hi tile
            uint4 pl[2];
            uint32_t* lp = reinterpret_cast<uint32_t*>(pl);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float2 hf = unpack_h2(hp[i], true);
              lp[i] = pack_h2(v[2 * i] - hf.x, v[2 * i + 1] - hf.y, true);
            }
            *reinterpret_cast<uint4*>(ogen + g.slot_bytes + u0) = pl[0];
            *reinterpret_cast<uint4*>(ogen + g.slot_bytes + u1) = pl[1];
          }
          if (E_POOL) {
            // 2x2 max-pool on the packed bf16 pairs (the rounding is monotonic, so this equals pooling the stored
            // tensor): the window's pixels are lanes l, l^1 (column) and l^8 (row); the even/even lane stages it
            uint32_t* w8 = reinterpret_cast<uint32_t*>(pk);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              uint32_t o = __shfl_xor_sync(0xffffffffu, w8[i], 1);
              const uint32_t mw = max_h2(w8[i], o, f16);
              o = __shfl_xor_sync(0xffffffffu, mw, 8);
              w8[i] = max_h2(mw, o, f16);
            }
            if ((lane & 9) == 0) {
              const uint32_t pp = (uint32_t)((ty >> 1) * 8 + (tx >> 1)) * (uint32_t)(g.ch * 2);
              const uint32_t px_ = (pp >> 7) & (uint32_t)(g.ch / 8 - 1);
              uint8_t* pgen = smem_raw + (p_base + oslot * (g.slot_bytes / 4u) - raw);
              *reinterpret_cast<uint4*>(pgen + pp + ((((uint32_t)(2 * gg)) ^ px_) << 4)) = pk[0];
              *reinterpret_cast<uint4*>(pgen + pp + ((((uint32_t)(2 * gg + 1)) ^ px_) << 4)) = pk[1];
            }
          }
        }
        if (head_nc == 1) {
#pragma unroll
          for (int i = 0; i < 16; ++i) head_acc[0] = fmaf(v[i], s_head[col + i].x, head_acc[0]);
        } else if (head_nc > 1) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float4 hw = s_head[col + i];
            head_acc[0] = fmaf(v[i], hw.x, head_acc[0]);
            head_acc[1] = fmaf(v[i], hw.y, head_acc[1]);
            head_acc[2] = fmaf(v[i], hw.z, head_acc[2]);
            head_acc[3] = fmaf(v[i], hw.w, head_acc[3]);
          }
        }
        if (ew == 0 && gi == ng - 1) DG_TRACE(2, k_it, 4);
        if (++gg == gpc) {
          // ---- staging chunk cc complete ----
          if (side) {
            __syncwarp();
            if (lane == 0) mbar_arrive(sideEmpty + 8 * rs.idx);
            rs.advance(2);
          }
          if (stage_out) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            // store coordinates; the previous chunk's store has drained its slot before anyone passes the barrier
            // and refills it
            int c0 = n0 + cc * g.ch, c1 = 0, c3 = 0;
            if (t0) {
              if (a.deconv) {
                const int ab = c0 / Cout;
                c0 -= ab * Cout; c1 = ab & 1; c3 = ab >> 1;
              }
              asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            if (ew == 0 && gi == ng - 1) DG_TRACE(2, k_it, 5);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (ew == 0 && gi == ng - 1) DG_TRACE(2, k_it, 6);
            if (t0) {
              const uint32_t src = o_base + (SPLIT ? 2u * oslot : oslot) * g.slot_bytes;
              if (a.deconv) tma_store_5d(&tm.out, src, c0, c1, cur.tw * 16, c3, cur.n * a.H + cur.th * 16);
              else tma_store_4d(&tm.out, src, c0, cur.tw * 16, cur.th * 16, cur.n);
              if (SPLIT) {
                if (a.deconv) tma_store_5d(&tm.out2, src + g.slot_bytes, c0, c1, cur.tw * 16, c3, cur.n * a.H + cur.th * 16);
                else tma_store_4d(&tm.out2, src + g.slot_bytes, c0, cur.tw * 16, cur.th * 16, cur.n);
              }
              if (E_POOL)
                tma_store_4d(&tm.pool, p_base + oslot * (g.slot_bytes / 4u), c0, cur.tw * 8, cur.th * 8, cur.n);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            oslot ^= 1u;
          }
          gg = 0;
          ++cc;
        }
      };
      // TMEM reads run one 32-column batch ahead of the math (ping-pong register buffers)
      tc_ld16_issue(t_row, vb[0]);
      if (ng > 1) tc_ld16_issue(t_row + 16u, vb[1]);
      for (int gb = 0; gb < ng; gb += 4) {
        tc_ld_wait4(vb);
        if (ew == 0 && gb == 0) DG_TRACE(2, k_it, 3);
        if (gb + 2 < ng) {
          tc_ld16_issue(t_row + (uint32_t)((gb + 2) * 16), vb[2]);
          if (gb + 3 < ng) tc_ld16_issue(t_row + (uint32_t)((gb + 3) * 16), vb[3]);
        } else {
          // all of this item's accumulator is in registers: hand the TMEM stage back before the math and the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(accEmpty + 8 * as);
        }
        do_group(vb[0], gb);
        if (gb + 1 < ng) do_group(vb[1], gb + 1);
        if (gb + 2 < ng) {
          tc_ld_wait4(vb);
          if (gb + 4 < ng) {
            tc_ld16_issue(t_row + (uint32_t)((gb + 4) * 16), vb[0]);
            if (gb + 5 < ng) tc_ld16_issue(t_row + (uint32_t)((gb + 5) * 16), vb[1]);
          } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(accEmpty + 8 * as);
          }
          do_group(vb[2], gb + 2);
          if (gb + 3 < ng) do_group(vb[3], gb + 3);
        }
      }
      if (ew == 0) DG_TRACE(2, k_it, 2);
      if (head_nc) {
        const int nc = head_nc;
        float o0 = head_acc[0] + __ldg(a.head_b), o1 = 0.f, o2 = 0.f, o3 = 0.f;
        if (nc > 1) o1 = head_acc[1] + __ldg(a.head_b + 1);
        if (nc > 2) o2 = head_acc[2] + __ldg(a.head_b + 2);
        if (nc > 3) o3 = head_acc[3] + __ldg(a.head_b + 3);
        if (a.head_act == 0) {
          o0 = tanhf(o0);
          if (nc > 1) { o1 = tanhf(o1); o2 = tanhf(o2); o3 = tanhf(o3); }
        } else if (a.head_act == 1) {
          float mx = o0;
          if (nc > 1) mx = fmaxf(mx, o1);
          if (nc > 2) mx = fmaxf(mx, o2);
          if (nc > 3) mx = fmaxf(mx, o3);
          o0 = expf(o0 - mx);
          o1 = nc > 1 ? expf(o1 - mx) : 0.f;
          o2 = nc > 2 ? expf(o2 - mx) : 0.f;
          o3 = nc > 3 ? expf(o3 - mx) : 0.f;
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
      cur = nx1; nx1 = nx2; nx2.advance();
      fslot = fnext;
    }
    if (t0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all output tiles written
  }

  tc_fence_before();
  __syncthreads();
  DG_TRACE_DUMP;
  if (warp == CTRL_W0 + 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(g.tmem_cols) : "memory");
  }
}

// ---- instantiation helpers used by conv_tc_k{1,3,5}.cu ----
template <int KS, int KSTEPS, bool RES, int EPI, bool F16 = false, bool SPLIT = false>
int launch_one(int grid, uint32_t smem, cudaStream_t st, const TcMaps& tm, const ConvArgs& a, const TcGeom& g) {
  DG_CHECK_CUDA(dg_launch_pdl(conv_tc_kernel<KS, KSTEPS, RES, EPI, F16, SPLIT>, dim3(grid), dim3(TC_THREADS), smem, st, tm, a, g));
  DG_LAUNCH_CHECK();
  return 0;
}
template <int KS, int KSTEPS, bool RES, int EPI, bool F16 = false, bool SPLIT = false>
int set_attr_one() {
  DG_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<KS, KSTEPS, RES, EPI, F16, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024 - DG_TRACE_SMEM));
  return 0;
}

// One (kernel size, EPI) family = the six (KSTEPS, RES) instantiations; key = (kc/16)*10 + resident
#define DG_TC_CASES(KS_, EPI_)                                                            \
  case EPI_ * 100 + 10: return launch_one<KS_, 1, false, EPI_>(grid, smem, st, tm, a, g); \
  case EPI_ * 100 + 11: return launch_one<KS_, 1, true, EPI_>(grid, smem, st, tm, a, g);  \
  case EPI_ * 100 + 20: return launch_one<KS_, 2, false, EPI_>(grid, smem, st, tm, a, g); \
  case EPI_ * 100 + 21: return launch_one<KS_, 2, true, EPI_>(grid, smem, st, tm, a, g);  \
  case EPI_ * 100 + 40: return launch_one<KS_, 4, false, EPI_>(grid, smem, st, tm, a, g); \
  case EPI_ * 100 + 41: return launch_one<KS_, 4, true, EPI_>(grid, smem, st, tm, a, g);
#define DG_TC_ATTRS(KS_, EPI_)                        \
  DG_TRY((set_attr_one<KS_, 1, false, EPI_>()));      \
  DG_TRY((set_attr_one<KS_, 1, true, EPI_>()));       \
  DG_TRY((set_attr_one<KS_, 2, false, EPI_>()));      \
  DG_TRY((set_attr_one<KS_, 2, true, EPI_>()));       \
  DG_TRY((set_attr_one<KS_, 4, false, EPI_>()));      \
  DG_TRY((set_attr_one<KS_, 4, true, EPI_>()));
// IEEE-half instantiations (generator inference: plain and FiLM 3x3 layers, transposed conv): key = epi + 8
#define DG_TC_CASES_F16(KS_, EPI_)                                                                 \
  case (EPI_ + 8) * 100 + 10: return launch_one<KS_, 1, false, EPI_, true>(grid, smem, st, tm, a, g); \
  case (EPI_ + 8) * 100 + 11: return launch_one<KS_, 1, true, EPI_, true>(grid, smem, st, tm, a, g);  \
  case (EPI_ + 8) * 100 + 20: return launch_one<KS_, 2, false, EPI_, true>(grid, smem, st, tm, a, g); \
  case (EPI_ + 8) * 100 + 21: return launch_one<KS_, 2, true, EPI_, true>(grid, smem, st, tm, a, g);  \
  case (EPI_ + 8) * 100 + 40: return launch_one<KS_, 4, false, EPI_, true>(grid, smem, st, tm, a, g); \
  case (EPI_ + 8) * 100 + 41: return launch_one<KS_, 4, true, EPI_, true>(grid, smem, st, tm, a, g);
#define DG_TC_ATTRS_F16(KS_, EPI_)                          \
  DG_TRY((set_attr_one<KS_, 1, false, EPI_, true>()));      \
  DG_TRY((set_attr_one<KS_, 1, true, EPI_, true>()));       \
  DG_TRY((set_attr_one<KS_, 2, false, EPI_, true>()));      \
  DG_TRY((set_attr_one<KS_, 2, true, EPI_, true>()));       \
  DG_TRY((set_attr_one<KS_, 4, false, EPI_, true>()));      \
  DG_TRY((set_attr_one<KS_, 4, true, EPI_, true>()));
// split-half storage instantiations (conv_tc_split.cu): key = ks * 1000 + epi * 100 + (kc/16)*10 + resident
#define DG_TC_CASES_SPLIT(KS_, EPI_)                                                                                 \
  case KS_ * 1000 + EPI_ * 100 + 10: return launch_one<KS_, 1, false, EPI_, true, true>(grid, smem, st, tm, a, g);   \
  case KS_ * 1000 + EPI_ * 100 + 11: return launch_one<KS_, 1, true, EPI_, true, true>(grid, smem, st, tm, a, g);    \
  case KS_ * 1000 + EPI_ * 100 + 20: return launch_one<KS_, 2, false, EPI_, true, true>(grid, smem, st, tm, a, g);   \
  case KS_ * 1000 + EPI_ * 100 + 21: return launch_one<KS_, 2, true, EPI_, true, true>(grid, smem, st, tm, a, g);    \
  case KS_ * 1000 + EPI_ * 100 + 40: return launch_one<KS_, 4, false, EPI_, true, true>(grid, smem, st, tm, a, g);   \
  case KS_ * 1000 + EPI_ * 100 + 41: return launch_one<KS_, 4, true, EPI_, true, true>(grid, smem, st, tm, a, g);
#define DG_TC_ATTRS_SPLIT(KS_, EPI_)                              \
  DG_TRY((set_attr_one<KS_, 1, false, EPI_, true, true>()));      \
  DG_TRY((set_attr_one<KS_, 1, true, EPI_, true, true>()));       \
  DG_TRY((set_attr_one<KS_, 2, false, EPI_, true, true>()));      \
  DG_TRY((set_attr_one<KS_, 2, true, EPI_, true, true>()));       \
  DG_TRY((set_attr_one<KS_, 4, false, EPI_, true, true>()));      \
  DG_TRY((set_attr_one<KS_, 4, true, EPI_, true, true>()));
#endif  // __CUDACC__

}  // namespace convtc
