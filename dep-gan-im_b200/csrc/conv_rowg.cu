// Row-streaming tcgen05 convolution, generic in kernel size, channel chunk and output width (the second generation of
// conv_row.cu, which stays the 3x3 / 32-channel / fused-head specialisation).  It takes the layers whose tap-by-tap
// implicit GEMM is paced by the MMA issue rate instead of the tensor pipe:
//   * the 5x5 layers of Dis_C2D_FCN1 (TG:319-325): conv2d_dis_0b (16 -> 16 @ full resolution), conv2d_dis_1a / _1b
//     (16 -> 32, 32 -> 32 @ half resolution), forward (+ fused MaxPooling2D), data gradient and JVP (mask epilogue);
//   * the 64-output-channel 3x3 layers of Gen_UNet2D at half resolution (TG:411-421, 469-479): conv2d_gen_2 (32 -> 64),
//     conv2d_gen_noise_m2 / _p2 (FiLM), conv2d_gen_3, conv2d_gen_15.
//
// Why.  One tcgen05.mma (M = 128, K = 16) costs max(32 + N/4, N/2) clk (scripts/ubench/mma_rate.cu,
// profiles/r02_mma_rate_ubench.txt): 39 clk at N = 16, 52 at N = 80, 80 at N = 160.  conv_tc_kernel issues one
// instruction per (tap, 16 channels): 25 x 39 = 975 clk per 128 pixels for the 16 -> 16 5x5 layer.  Here the KS
// vertical taps are stacked on N (N = KS * Cout), so the same 128 pixels cost 5 x 52 = 260 clk, and the A slice
// (128 pixels x 16 channels) is read from shared memory once per KS taps.
//
// GEMM view (as conv_row.cu).  M = 128 consecutive pixels of ONE image row j, N = (dy, cout), K = (dx, cin):
//     D_j[x][(dy, co)] = sum_dx sum_ci  X[j][x + dx - P][ci] * W[dy][dx][ci][co],        P = KS / 2,
// the dx shift being a different start address of the [128 + KS - 1 pixels][CK channels] row tile in shared memory.
// Output row r = sum_dy D_{r + dy - P}[dy] is accumulated by the tensor core itself: TMEM holds a ring of NBLK blocks of
// CO columns, block (o mod NBLK) belongs to output row o, the weight rows are stacked (dy = KS-1 .. 0), so the KS * CO
// accumulator columns of input row j are the KS consecutive blocks of output rows j-KS+1+P .. j+P; every MMA
// accumulates, the epilogue hands a block back zeroed.  Rows outside the image are zero through TMA out-of-bounds
// fill; at segment edges the matching sub-range of weight rows is used; a range that wraps the ring is issued in two
// parts.
//
// One 128-pixel column block per CTA (consecutive MMAs into one accumulator run at the full rate -- measured, same
// table), so any width that is a multiple of 128 works.  The N * (W / 128) * H output rows are cut into one contiguous
// range per CTA (segments end at image borders; KS - 1 halo input rows per segment).
//
// Roles (12 warps): warps 0-3 and 4-7 are two epilogue groups that take alternate steps (a step = RP = 2 output rows,
// 1 for CO = 64); warp w owns TMEM lane quarter w % 4 (32 pixels), stages its pixels in its own swizzled slots and
// issues its own TMA stores -- no block- or group-wide barrier in the row loop.  Warp 8 TMA producer of the row tiles
// (and, once, of the resident weights), warp 9 MMA issuer + TMEM owner, warp 10 TMA producer of the epilogue's side rows
// (FiLM residual, add / mask sources), warp 11 scout (does the issuer's barrier waits and publishes "rows ready").  EPI bit 2: 2x2 max-pool of the stored values (vertical max of the row pair in
// registers, horizontal max by one shuffle), staged and stored by a second TMA store.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "conv_tc_kernel.cuh"

using namespace convtc;

namespace {

constexpr int RG_THREADS = 384;
constexpr int BWG = 128;   // pixels per M block
constexpr int NQG = 64;    // "input row done" barriers (ring)
constexpr int NREC = 32;   // row records for the issuer (ring; more than the deepest row-tile ring)
constexpr int NSIDEG = 4;  // side-row stages (one stage = one 128-pixel row of every side tensor)

#ifdef DG_ROWG_TRACE  // timing experiment only: event clocks of CTA 0 (roles: 0 producer, 1 issuer, 2 epilogue warp 0, 3 warp 4)
__device__ long long* g_rowg_trace = nullptr;  // [4 roles][64 rows][8 events], dumped from shared memory at exit
#define GTRACE_DECL __shared__ uint32_t s_gtrace[4 * 64 * 8];
#define GTRACE_SMEM 8192
#define GTRACE(role_, row_, ev_)                                                                      \
  do {                                                                                                \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (row_) < 64)                                    \
      s_gtrace[((role_) * 64 + (row_)) * 8 + (ev_)] = (uint32_t)clock64();                            \
  } while (0)
#define GTRACE_DUMP                                                                                   \
  do {                                                                                                \
    if (blockIdx.x == 0 && g_rowg_trace)                                                              \
      for (int i_ = threadIdx.x; i_ < 4 * 64 * 8; i_ += blockDim.x) g_rowg_trace[i_] = s_gtrace[i_];  \
  } while (0)
#else
#define GTRACE_DECL
#define GTRACE_SMEM 0
#define GTRACE(role_, row_, ev_) do {} while (0)
#define GTRACE_DUMP do {} while (0)
#endif

template <int KS, int CK, int CO>
struct RG {
  static constexpr int PAD = KS / 2;
  static constexpr int TILE_PX = BWG + KS - 1;
  static constexpr uint32_t SPAN = CK * 2;                                  // bytes per pixel of a row tile
  static constexpr uint32_t LAYOUT = SPAN == 128 ? 2u : SPAN == 64 ? 4u : 6u;  // UMMA swizzle mode of the operands
  static constexpr int KSTEPS = CK / 16;
  static constexpr uint32_t A_TX = TILE_PX * SPAN;
  static constexpr uint32_t A_STAGE = (A_TX + 1023u) & ~1023u;
  static constexpr uint32_t B_ROWS = KS * CO;                               // (dy, cout) rows of one (dx, chunk) tile
  static constexpr uint32_t B_TILE = B_ROWS * SPAN;
  static constexpr int NBLK = 512 / CO > 16 ? 16 : 512 / CO;                // TMEM ring (blocks of CO columns)
  static constexpr int RP = CO == 64 ? 1 : 2;                               // output rows per epilogue step
  static constexpr int NSLOT = CO == 64 ? 1 : 2;                            // staging slots per epilogue warp
  static constexpr uint32_t OSPAN = CO * 2;                                 // bytes per pixel of an output / side row
  static constexpr uint32_t WSLOT = RP * 32 * OSPAN;                        // one warp's staging slot
  static constexpr uint32_t PSLOT = 16 * OSPAN < 1024 ? 1024 : 16 * OSPAN;  // one warp's pooled staging slot
  static constexpr uint32_t SIDE_ROW = BWG * OSPAN;
  static constexpr int CW = CO < 32 ? CO : 32;                              // accumulator columns per epilogue pass
  static constexpr int NH = CO / CW;                                        // passes per row
  static constexpr int NSIDE = CO == 64 ? 3 : NSIDEG;                       // side-row stages (shared memory is tight at 64)
  // Output / side / pooled rows move through TMA as 128-byte "super pixels" (PG = 4 / 2 / 1 neighbouring pixels of
  // 16 / 32 / 64 channels): a 32- or 64-byte innermost box costs a shared-memory cycle per pixel, a 128-byte one per PG
  // pixels, and the shared-memory port is what these kernels run out of.
  static constexpr int PG = 128 / (int)OSPAN;
  static_assert(NBLK >= KS + RP + 1, "TMEM ring too small for the taps in flight");
  static_assert(B_TILE % 256 == 0 && (CO * SPAN) % 256 == 0, "weight tiles must keep the swizzle phase");
};

struct RowgGeom {
  int nchunk0, nchunk1;  // CK-channel chunks from in0 / in1
  int cblocks;           // 128-pixel column blocks per image row
  int na;                // row-tile ring depth
  int batch;             // input rows the issuer issues per wait (<= na)
  int pf;                // input rows the L2 prefetch runs ahead of the loads (0: off)
  int n_side;            // side tensors per output row (0..2)
  int pool;              // EPI bit 2 active
  long long rows_total;  // N * cblocks * H output rows (of 128 pixels)
};

struct RowgMaps {
  CUtensorMap a0, a1, b, out1, out2, s0, s1, pool;
};

template <int NCOL>
__device__ __forceinline__ void tg_ld(uint32_t taddr, uint32_t (&r)[NCOL]);
template <>
__device__ __forceinline__ void tg_ld<16>(uint32_t taddr, uint32_t (&r)[16]) { tc_ld16_issue(taddr, r); }
template <>
__device__ __forceinline__ void tg_ld<32>(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
template <int NCOL>
__device__ __forceinline__ void tg_fence(uint32_t (&r)[NCOL]) {  // ties the registers to the preceding wait
#pragma unroll
  for (int i = 0; i < NCOL; ++i) asm volatile("" : "+r"(r[i])::"memory");
}
__device__ __forceinline__ void tg_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
template <int NCOL>
__device__ __forceinline__ void tg_zero(uint32_t taddr);  // zeroes NCOL columns of this thread's TMEM lane (async)
template <>
__device__ __forceinline__ void tg_zero<16>(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
      ::"r"(taddr), "r"(z)
      : "memory");
}
template <>
__device__ __forceinline__ void tg_zero<32>(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
      "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
      ::"r"(taddr), "r"(z)
      : "memory");
}
__device__ __forceinline__ void tg_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint2 ld_acquire_shared_v2(uint32_t addr) {  // one 64-bit access: both words are consistent
  unsigned long long v;
  asm volatile("ld.acquire.cta.shared::cta.b64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
  return make_uint2((uint32_t)v, (uint32_t)(v >> 32));
}
// TMA prefetch of one tile into L2 (no shared-memory destination)
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* tm, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
               ::"l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.volatile.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_shared_v2(uint32_t addr, uint32_t x, uint32_t y) {
  const unsigned long long v = (unsigned long long)x | ((unsigned long long)y << 32);
  asm volatile("st.release.cta.shared::cta.b64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}

// The CTA's contiguous range of output rows (128 pixels wide), cut into segments at image borders.  A unit is one
// (slice, column block); every role walks the same segments.  `even`: range borders fall on even rows (fused pool).
struct SegIterG {
  long long pos, end;
  int H, cblocks;
  int n, cb, r0, cnt;  // current segment: slice, column block, first output row, number of output rows
  __device__ __forceinline__ void init(long long total, int H_, int cblocks_, bool even) {
    H = H_; cblocks = cblocks_;
    if (even) {
      const long long half = total >> 1;  // H is even when pooling, so is total
      pos = 2 * (half * (long long)blockIdx.x / (long long)gridDim.x);
      end = 2 * (half * (long long)(blockIdx.x + 1) / (long long)gridDim.x);
    } else {
      pos = total * (long long)blockIdx.x / (long long)gridDim.x;
      end = total * (long long)(blockIdx.x + 1) / (long long)gridDim.x;
    }
    decode();
  }
  __device__ __forceinline__ void decode() {
    if (pos >= end) { cnt = 0; return; }
    const long long unit = pos / H;
    r0 = (int)(pos - unit * H);
    n = (int)(unit / cblocks);
    cb = (int)(unit - (long long)n * cblocks);
    const long long left = end - pos;
    cnt = (long long)(H - r0) < left ? H - r0 : (int)left;
  }
  __device__ __forceinline__ bool valid() const { return pos < end; }
  __device__ __forceinline__ void next() { pos += cnt; decode(); }
};

// ---------------------------------------------------------------------------------------------------------
// EPI bit 0: FiLM residual side row; bit 1: add / mask side rows; bit 2: fused 2x2 max-pool.
//
// Indexing.  A segment of cnt output rows reads Q = cnt + KS - 1 input rows.  EVERY input row issues the same full-width
// MMA (N = KS * CO): input row j of the segment accumulates into the KS blocks of "slots" j .. j + KS - 1, where slot
// s = output row s - (KS - 1) of the segment.  Slots 0 .. KS-2 and cnt + KS - 1 .. cnt + 2 KS - 3 are phantoms (rows
// above / below the segment): they collect partial sums nobody reads and are recycled (zeroed) by the epilogue like any
// other block.  That keeps the issuer free of edge cases -- its per-row work is a handful of instructions, which is what
// the kernel's speed hangs on: one warp issues dependent scalar instructions at ~4 clk each, an mbarrier poll costs
// ~100 clk, and a row's MMAs are only 260 clk of tensor work for the 16 -> 16 layer (profiles/r02_rowg_role_trace.txt:
// the first version spent 1000 clk per row in the issuer's bookkeeping).  Slot s is final after input row min(s, Q - 1).
// ---------------------------------------------------------------------------------------------------------
// F16: IEEE-half storage of activations / weights instead of bfloat16 (compile time: see conv_tc_kernel.cuh).
template <int KS, int CK, int CO, int EPI, bool F16>
__global__ void __launch_bounds__(RG_THREADS, 1) conv_rowg_kernel(const __grid_constant__ RowgMaps tm, const ConvArgs a,
                                                                  const RowgGeom g) {
  typedef RG<KS, CK, CO> G;
  constexpr int NBLK = G::NBLK, RP = G::RP, NSLOT = G::NSLOT, CW = G::CW, NH = G::NH, NSD = G::NSIDE;
  constexpr bool E_RES = (EPI & 1) != 0, E_AM = (EPI & 2) != 0, E_POOL = (EPI & 4) != 0;
  static_assert(!E_POOL || RP == 2, "the fused pool works on row pairs");
  static_assert((NBLK & (NBLK - 1)) == 0 && (NQG & (NQG - 1)) == 0, "ring sizes are powers of two");
  extern __shared__ uint8_t smem_raw[];
  GTRACE_DECL
#ifdef DG_ROWG_TRACE
  for (int i_ = threadIdx.x; i_ < 4 * 64 * 8; i_ += blockDim.x) s_gtrace[i_] = 0;
#endif
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const int nchunks = g.nchunk0 + g.nchunk1;
  const uint32_t a_stage = (uint32_t)nchunks * G::A_STAGE;                              // one input row, all chunks
  const uint32_t b_base = base;                                                         // weights: [dx][chunk] tiles
  const uint32_t a_base = (b_base + (uint32_t)KS * nchunks * G::B_TILE + 1023u) & ~1023u;  // input-row ring
  const uint32_t o_base = a_base + (uint32_t)g.na * a_stage;                            // 8 warps x NSLOT slots
  const uint32_t p_base = o_base + 8u * NSLOT * G::WSLOT;                               // pooled: 8 warps x NSLOT
  const uint32_t s_base = p_base + (E_POOL ? 8u * NSLOT * G::PSLOT : 0u);               // NSD stages x n_side rows
  const uint32_t bar_base = s_base + (uint32_t)NSD * g.n_side * G::SIDE_ROW;
  const uint32_t fullA = bar_base;
  const uint32_t fullB = fullA + 8 * g.na;
  const uint32_t rowDone = fullB + 8, blkEmpty = rowDone + 8 * NQG;
  const uint32_t sideFull = blkEmpty + 8 * NBLK, sideEmpty = sideFull + 8 * NSD;
  const uint32_t tmem_slot = sideEmpty + 8 * NSD;
  const uint32_t rec_ring = tmem_slot + 16;              // NREC row records for the issuer (written by the scout warp)
  const uint32_t f_off = (rec_ring + 8u * NREC + 15u) & ~15u;  // floats: scale, shift, per-warp tables
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  float* s_scale = reinterpret_cast<float*>(smem_raw + (f_off - raw));
  float* s_shift = s_scale + CO;
  float* s_tab = s_shift + CO;  // [8 warps][2 buffers][scale CO | shift CO]

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < g.na; ++i) mbar_init(fullA + 8 * i, 1);
    mbar_init(fullB, 1);
    for (int i = 0; i < NREC; ++i) st_release_shared_v2(rec_ring + 8u * i, 0u, 0u);
    for (int i = 0; i < NQG; ++i) mbar_init(rowDone + 8 * i, 1);
    for (int i = 0; i < NBLK; ++i) mbar_init(blkEmpty + 8 * i, 4);
    for (int i = 0; i < NSD; ++i) { mbar_init(sideFull + 8 * i, 1); mbar_init(sideEmpty + 8 * i, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == CTRL_W0 + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x < CO) {
    s_scale[threadIdx.x] = a.scale ? a.scale[threadIdx.x] : 1.f;
    s_shift[threadIdx.x] = a.shift ? a.shift[threadIdx.x] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (warp < 4) {  // every block starts zeroed: warp w clears lane quarter w
    for (int b = 0; b < NBLK * NH; ++b) tg_zero<CW>(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(b * CW));
    tg_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp >= CTRL_W0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    if (warp == CTRL_W0) {
      // ===== TMA producer: resident weights once, then one ring stage per input row (all its channel chunks) =====
      if (elect_one()) {
        mbar_expect_tx(fullB, (uint32_t)KS * nchunks * G::B_TILE);
        for (int dx = 0; dx < KS; ++dx)
          for (int c = 0; c < nchunks; ++c) {
            const int kglob = c < g.nchunk0 ? c * CK : a.C0 + (c - g.nchunk0) * CK;
            for (int dy = 0; dy < KS; ++dy)  // rows stacked (dy = KS-1 .. 0): column block k of an MMA = slot j + k
              tma_load_2d(b_base + (uint32_t)(dx * nchunks + c) * G::B_TILE + (uint32_t)(KS - 1 - dy) * CO * G::SPAN,
                          &tm.b, fullB, kglob, (dy * KS + dx) * CO);
          }
      }
      __syncwarp();
      Ring ra;
      SegIterG si;
      si.init(g.rows_total, a.H, g.cblocks, E_POOL);
      uint32_t pq = 0;  // input-row sequence number
      for (; si.valid(); si.next()) {
        const int x0 = si.cb * BWG - G::PAD, r0 = si.r0 - G::PAD;
        const int nin = si.cnt + KS - 1;
        // optional L2 prefetch g.pf input rows ahead of the loads (an experiment that did not pay, see plan_g)
        auto prefetch_row = [&](int jr) {
          for (int c = 0; c < nchunks; ++c) {
            const bool first = c < g.nchunk0;
            tma_prefetch_4d(first ? &tm.a0 : &tm.a1, (first ? c : c - g.nchunk0) * CK, x0, r0 + jr, si.n);
          }
        };
        if (g.pf > 0 && elect_one())
          for (int jr = 0; jr < g.pf && jr < nin; ++jr) prefetch_row(jr);
        __syncwarp();
        for (int j = 0; j < nin; ++j, ++pq) {
          GTRACE(0, (int)pq, 0);
          // the stage is free once the row that used it last (na rows ago) has been multiplied: that row's "done"
          // barrier (one commit per row serves the epilogue and this ring)
          if (pq >= (uint32_t)g.na) {
            const uint32_t qp = pq - (uint32_t)g.na;
            mbar_wait(rowDone + 8 * (qp & (NQG - 1)), (qp / NQG) & 1u);
          }
          GTRACE(0, (int)pq, 1);
          if (elect_one()) {
            if (g.pf > 0 && j + g.pf < nin) prefetch_row(j + g.pf);
#ifdef DG_ROWG_NOLOAD  // timing experiment only: the row tiles are never loaded (results are garbage)
            mbar_arrive(fullA + 8 * ra.idx);
#else
            mbar_expect_tx(fullA + 8 * ra.idx, (uint32_t)nchunks * G::A_TX);
            for (int c = 0; c < nchunks; ++c) {
              const bool first = c < g.nchunk0;
              tma_load_4d(a_base + ra.idx * a_stage + (uint32_t)c * G::A_STAGE, first ? &tm.a0 : &tm.a1,
                          fullA + 8 * ra.idx, (first ? c : c - g.nchunk0) * CK, x0, r0 + j, si.n);
            }
#endif
          }
          __syncwarp();
          ra.advance(g.na);
        }
      }
    } else if (warp == CTRL_W0 + 1) {
      // ===== MMA issuer: per input row KS * KSTEPS * nchunks instructions of N = KS * CO into the blocks of slots
      // s .. s + KS - 1 (two instructions each where the range wraps the ring), issued back to back by one elected lane
      // with immediate descriptor offsets, then ONE commit ("row done": the epilogue's signal and the ring's stage-free
      // signal).  The issuer polls no mbarrier and keeps no geometry: the scout warp (below) does the waiting and leaves
      // one record per row (sequence tag, ring stage, first TMEM block) in shared memory; the record of the next row is
      // read while this row's instructions are still queued on the tensor pipe.  Everything between two rows' MMAs has
      // to fit the ~100 clk the two queued instructions last (profiles/r02_rowg_role_trace.txt). =====
      const uint32_t hi = ((uint32_t)(8 * G::SPAN) >> 4) | (1u << 14) | (G::LAYOUT << 29);
      constexpr uint32_t LBO1 = 1u << 16;
      constexpr uint32_t BLK16 = (CO * G::SPAN) >> 4;  // one dy block of weight rows in descriptor units
      constexpr uint32_t BT16 = G::B_TILE >> 4, AS16 = G::A_STAGE >> 4, SP16 = G::SPAN >> 4;
      constexpr bool f16_in = F16;
      const uint32_t idesc_full = make_idesc(KS * CO, f16_in);
      mbar_wait(fullB, 0);
      tc_fence_after();
      const uint32_t b_lo0 = ((b_base & 0x3FFFFu) >> 4) | LBO1;
      const uint32_t a_lo_base = ((a_base & 0x3FFFFu) >> 4) | LBO1;
      const uint32_t as16 = a_stage >> 4;
      const uint32_t dxb16 = (uint32_t)nchunks * BT16;  // weight tiles of consecutive dx are nchunks tiles apart

      uint32_t rows_in = 0;  // input rows of this CTA (every role walks the same segments)
      {
        SegIterG si;
        si.init(g.rows_total, a.H, g.cblocks, E_POOL);
        for (; si.valid(); si.next()) rows_in += (uint32_t)(si.cnt + KS - 1);
      }
      // Rows are issued in batches of up to `batch` (<= ring depth): one wait, one elect and one warp sync per batch, so
      // the per-row scalar work shrinks to what the elected lane does between two rows' instructions -- and that runs
      // while the tensor pipe works through its queue.
      const uint32_t batch = (uint32_t)g.batch;
      for (uint32_t q = 0; q < rows_in;) {
        const uint32_t nb = rows_in - q < batch ? rows_in - q : batch;
        // record of row r: .x = r + 1 once the row's tiles have landed and its blocks are free, .y = stage | block << 8.
        // The scout publishes in row order, so the batch is ready when its last record is.
        const uint32_t last_addr = rec_ring + ((q + nb - 1u) & (NREC - 1)) * 8u;
        uint2 rec_last = ld_acquire_shared_v2(last_addr);
        while (rec_last.x != q + nb) rec_last = ld_acquire_shared_v2(last_addr);
        tc_fence_after();
        GTRACE(1, (int)q, 0);
        if (elect_one()) {
          for (uint32_t r = 0; r < nb; ++r) {
            const uint32_t ry = r + 1u == nb ? rec_last.y : ld_shared_u32(rec_ring + ((q + r) & (NREC - 1)) * 8u + 4u);
            const uint32_t stage = ry & 0xFFu, blk0 = ry >> 8;
            const uint32_t d1 = tmem_base + blk0 * CO;
            const uint32_t a_lo = a_lo_base + stage * as16;
            if (blk0 + KS <= (uint32_t)NBLK) {
              for (int c = 0; c < nchunks; ++c) {
                const uint32_t ac = a_lo + (uint32_t)c * AS16, bc = b_lo0 + (uint32_t)c * BT16;
#pragma unroll
                for (int dx = 0; dx < KS; ++dx)
#pragma unroll
                  for (int k = 0; k < G::KSTEPS; ++k)
                    tc_mma(d1, ((uint64_t)hi << 32) | (ac + SP16 * dx + 2 * k),
                           ((uint64_t)hi << 32) | (bc + (uint32_t)dx * dxb16 + 2 * k), idesc_full, 1u);
              }
            } else {
              const uint32_t n1 = (uint32_t)NBLK - blk0;  // blocks before the ring wraps
              const uint32_t idesc1 = make_idesc((int)(CO * n1), f16_in), idesc2 = make_idesc((int)(CO * (KS - n1)), f16_in);
              for (int c = 0; c < nchunks; ++c) {
                const uint32_t ac = a_lo + (uint32_t)c * AS16, bc = b_lo0 + (uint32_t)c * BT16;
#pragma unroll
                for (int dx = 0; dx < KS; ++dx)
#pragma unroll
                  for (int k = 0; k < G::KSTEPS; ++k) {
                    const uint64_t da = ((uint64_t)hi << 32) | (ac + SP16 * dx + 2 * k);
                    tc_mma(d1, da, ((uint64_t)hi << 32) | (bc + (uint32_t)dx * dxb16 + 2 * k), idesc1, 1u);
                    tc_mma(tmem_base, da, ((uint64_t)hi << 32) | (bc + (uint32_t)dx * dxb16 + n1 * BLK16 + 2 * k), idesc2, 1u);
                  }
              }
            }
            tc_commit(rowDone + 8 * ((q + r) & (NQG - 1)));
            GTRACE(1, (int)(q + r), 1);
          }
        }
        __syncwarp();
        GTRACE(1, (int)q, 2);
        q += nb;
      }
    } else if (warp == CTRL_W0 + 3) {
      // ===== scout: waits, in row order, for the row's tiles (TMA) and for the blocks of its newest slots (drained and
      // zeroed by the epilogue), then publishes the row's record for the issuer.  It runs as far ahead as the rings allow
      // (at most na rows: a stage is refilled only after its row was multiplied), so NREC > na records never overlap. =====
      SegIterG si;
      si.init(g.rows_total, a.H, g.cblocks, E_POOL);
      Ring ra;
      uint32_t s0 = 0, acquired = 0, q = 0;
      for (; si.valid(); si.next()) {
        const int nin = si.cnt + KS - 1;
        for (int j = 0; j < nin; ++j, ++q) {
          mbar_wait(fullA + 8 * ra.idx, ra.phase);
          const uint32_t top = s0 + (uint32_t)(KS - 1);
          // a block is acquired the first time a row touches its slot: the previous owner (NBLK slots earlier) must have
          // been drained and zeroed by the four warps of its epilogue group
          while (acquired <= top) {
            if (acquired >= (uint32_t)NBLK) mbar_wait(blkEmpty + 8 * (acquired & (NBLK - 1)), ((acquired / NBLK) & 1u) ^ 1u);
            ++acquired;
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0)
            st_release_shared_v2(rec_ring + (q & (NREC - 1)) * 8u, q + 1u, (uint32_t)ra.idx | ((s0 & (NBLK - 1)) << 8));
          ra.advance(g.na);
          s0 += j + 1 == nin ? (uint32_t)KS : 1u;
        }
      }
    } else if (warp == CTRL_W0 + 2) {
      // ===== TMA producer of the side rows (FiLM residual, or add / mask sources): one 128-pixel row per output row =====
      if ((E_RES || E_AM) && g.n_side > 0) {
        Ring rs;
        SegIterG si;
        si.init(g.rows_total, a.H, g.cblocks, E_POOL);
        for (; si.valid(); si.next()) {
          for (int i = 0; i < si.cnt; ++i) {
            mbar_wait(sideEmpty + 8 * rs.idx, rs.phase ^ 1u);
            if (elect_one()) {
              const uint32_t dst = s_base + (uint32_t)(rs.idx * g.n_side) * G::SIDE_ROW;
              mbar_expect_tx(sideFull + 8 * rs.idx, (uint32_t)g.n_side * G::SIDE_ROW);
              tma_load_4d(dst, &tm.s0, sideFull + 8 * rs.idx, 0, si.cb * BWG / G::PG, si.r0 + i, si.n);
              if (g.n_side == 2)
                tma_load_4d(dst + G::SIDE_ROW, &tm.s1, sideFull + 8 * rs.idx, 0, si.cb * BWG / G::PG, si.r0 + i, si.n);
            }
            __syncwarp();
            rs.advance(NSD);
          }
        }
      }
    }
  } else {
    // ===== epilogue warps: two groups of four independent agents, alternating steps =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    const int grp = warp >> 2;       // epilogue group (takes the steps with the same parity)
    const int qd = warp & 3;         // TMEM lane quarter
    const int m = qd * 32 + lane;    // pixel of the 128-pixel block
    const bool has_add = E_AM && a.add_src != nullptr;
    const bool has_mask = E_AM && a.mask_src != nullptr;
    const bool side = (E_RES || E_AM) && g.n_side > 0;
    constexpr bool f16 = F16;
    float* tab = s_tab + warp * 4 * CO;  // [2 buffers][scale CO | shift CO]
    constexpr uint32_t OSPAN = G::OSPAN, UNITS = OSPAN / 16;
    // staging: this warp's NSLOT slots of [RP rows][32 pixels][OSPAN B], 16-byte units XOR-swizzled like the TMA
    const uint32_t w_o = o_base + (uint32_t)warp * NSLOT * G::WSLOT;
    const uint32_t w_p = p_base + (uint32_t)warp * NSLOT * G::PSLOT;
    // Rows of super pixels [pixels / PG][128 B] in the TMA's 128-byte swizzle: 16-byte unit j of super pixel s lives at
    // s * 128 + ((j ^ (s & 7)) << 4); pixel p of the row is units (p % PG) * UNITS .. + UNITS - 1 of super pixel p / PG.
    // x_off = byte offset of the super pixel, x_u0 = first unit of the pixel, x_xor = the super pixel's XOR term.
    constexpr uint32_t PG = (uint32_t)G::PG;
    const uint32_t p_off = ((uint32_t)lane / PG) * 128u, p_u0 = ((uint32_t)lane % PG) * UNITS, p_xor = ((uint32_t)lane / PG) & 7u;
    const uint32_t pl = (uint32_t)(lane >> 1);  // pooled pixel of the even lanes
    const uint32_t pp_off = (pl / PG) * 128u, pp_u0 = (pl % PG) * UNITS, pp_xor = (pl / PG) & 7u;
    const uint32_t sp_off = ((uint32_t)m / PG) * 128u, sp_u0 = ((uint32_t)m % PG) * UNITS, sp_xor = ((uint32_t)m / PG) & 7u;
    const uint32_t mask_off = has_add ? G::SIDE_ROW : 0u;
    const uint32_t t_lane = tmem_base + ((uint32_t)(qd * 32) << 16);

    uint32_t q0 = 0;        // input-row sequence number of the segment's first input row
    uint32_t sl0 = 0;       // slot sequence number of the segment's first slot
    uint32_t side0 = 0;     // real-output-row sequence number of the segment's first row (side-stage ring)
    uint32_t step = 0;      // epilogue steps so far (both groups count all of them)
    uint32_t nstaged = 0;   // stores this warp has issued (staging slot ring)
    int bseq = 0;
    SegIterG si;
    si.init(g.rows_total, a.H, g.cblocks, E_POOL);
    for (; si.valid(); q0 += (uint32_t)(si.cnt + KS - 1), sl0 += (uint32_t)(si.cnt + 2 * (KS - 1)),
                       side0 += (uint32_t)si.cnt, ++bseq, si.next()) {
      const int cnt = si.cnt;
      const int nslots = cnt + 2 * (KS - 1), nin = cnt + KS - 1;
      // per-segment table: v = acc * sc + sh   (BN folded; FiLM: the conditioning affine folded into the same pair)
      float* tb = tab + (bseq & 1) * 2 * CO;
      for (int cidx = lane; cidx < CO; cidx += 32) {
        float sc = s_scale[cidx], sh = s_shift[cidx];
        if (E_RES) {
          const float fg = __ldg(a.film_g + (size_t)si.n * a.film_stride + cidx);
          const float fb = __ldg(a.film_b + (size_t)si.n * a.film_stride + cidx);
          sh = fmaf(sh, fg, fb);
          sc *= fg;
        }
        tb[cidx] = sc;
        tb[CO + cidx] = sh;
      }
      __syncwarp();
      const int xw = si.cb * BWG + qd * 32;  // first pixel of this warp
      for (int s = 0; s < nslots; s += RP, ++step) {
        if ((int)(step & 1u) != grp) continue;
        const int gsz = nslots - s >= RP ? RP : 1;            // slots in this step
        const int i0 = s - (KS - 1);                          // output row of the step's first slot (may be a phantom)
        bool real[2];
        real[0] = i0 >= 0 && i0 < cnt;
        real[1] = RP == 2 && gsz == 2 && i0 + 1 >= 0 && i0 + 1 < cnt;
        const uint32_t o = sl0 + (uint32_t)s;                 // slot sequence number -> TMEM block
        const int trole = qd == 0 ? 2 + grp : 99;
        if (trole < 4) GTRACE(trole, (int)o, 0);
        // slot s is final after input row min(s, nin - 1); the step waits for its last slot (MMAs complete in order)
        const int jl = s + gsz - 1 < nin - 1 ? s + gsz - 1 : nin - 1;
        const uint32_t qd_ = q0 + (uint32_t)jl;
        mbar_wait(rowDone + 8 * (qd_ & (NQG - 1)), (qd_ / NQG) & 1u);
        tc_fence_after();
        if (trole < 4) GTRACE(trole, (int)o, 1);
#ifdef DG_ROWG_NOEPI  // timing experiment only: blocks are only zeroed and handed back (1), not even read (2)
        real[0] = real[1] = false;
#endif
        const bool any_real = real[0] || real[1];
        const uint32_t slot = nstaged % NSLOT;
        if (any_real) {
          // the TMA store that last read this staging slot was issued NSLOT stores (of this warp) ago
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NSLOT - 1) : "memory");
          __syncwarp();
        }
        if (trole < 4) GTRACE(trole, (int)o, 2);
        uint32_t pmax[E_POOL ? CO / 2 : 1];  // vertical max of the row pair, packed pairs
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          uint32_t va[RP][CW];
          const uint32_t blk[2] = {t_lane + (o & (NBLK - 1)) * CO + (uint32_t)(h * CW),
                                   t_lane + ((o + 1u) & (NBLK - 1)) * CO + (uint32_t)(h * CW)};
#if defined(DG_ROWG_NOEPI) && DG_ROWG_NOEPI == 1
          if (true) {
#else
          if (any_real) {
#endif
            tg_ld<CW>(blk[0], va[0]);
            if (RP == 2 && gsz == 2) tg_ld<CW>(blk[1], va[RP - 1]);
            tg_ld_wait();
#pragma unroll
            for (int u = 0; u < RP; ++u) tg_fence<CW>(va[u]);
          }
          if (trole < 4 && h == 0) GTRACE(trole, (int)o, 3);
          tg_zero<CW>(blk[0]);  // hand the blocks back zeroed (every MMA accumulates); waited for before the arrive
          if (RP == 2 && gsz == 2) tg_zero<CW>(blk[1]);
#pragma unroll
          for (int u = 0; u < RP; ++u) {
            if (real[u]) {
              float v[CW];
              // ---- affine (+ FiLM) ----
#pragma unroll
              for (int k4 = 0; k4 < CW / 4; ++k4) {
                const float4 sc4 = reinterpret_cast<const float4*>(tb + h * CW)[k4];
                const float4 sh4 = reinterpret_cast<const float4*>(tb + CO + h * CW)[k4];
                v[4 * k4 + 0] = __uint_as_float(va[u][4 * k4 + 0]); v[4 * k4 + 1] = __uint_as_float(va[u][4 * k4 + 1]);
                v[4 * k4 + 2] = __uint_as_float(va[u][4 * k4 + 2]); v[4 * k4 + 3] = __uint_as_float(va[u][4 * k4 + 3]);
                fma_f32x2(v[4 * k4 + 0], v[4 * k4 + 1], sc4.x, sc4.y, sh4.x, sh4.y);   // packed fp32x2: the bits of fmaf
                fma_f32x2(v[4 * k4 + 2], v[4 * k4 + 3], sc4.z, sc4.w, sh4.z, sh4.w);
              }
              if (side) {
                // side stage sequence = real output row sequence (the producer's order)
                const uint32_t sq = side0 + (uint32_t)(i0 + u);
                const uint32_t ss = sq % NSD;
                if (h == 0) mbar_wait(sideFull + 8 * ss, (sq / NSD) & 1u);
                const uint8_t* sgen = smem_raw + (s_base + (uint32_t)(ss * g.n_side) * G::SIDE_ROW - raw) + sp_off;
                if (E_RES) {
#pragma unroll
                  for (int uu = 0; uu < CW / 8; ++uu) {
                    const uint4 qv = *reinterpret_cast<const uint4*>(sgen + (((sp_u0 + (uint32_t)(h * (CW / 8) + uu)) ^ sp_xor) << 4));
                    const uint32_t wv[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                      const float2 rv = unpack_h2(wv[e], f16);
                      v[8 * uu + 2 * e] = fmaxf(v[8 * uu + 2 * e], 0.f) + rv.x;
                      v[8 * uu + 2 * e + 1] = fmaxf(v[8 * uu + 2 * e + 1], 0.f) + rv.y;
                    }
                  }
                }
                if (has_add) {
#pragma unroll
                  for (int uu = 0; uu < CW / 8; ++uu) {
                    const uint4 qv = *reinterpret_cast<const uint4*>(sgen + (((sp_u0 + (uint32_t)(h * (CW / 8) + uu)) ^ sp_xor) << 4));
                    const uint32_t wv[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                      const float2 rv = unpack_h2(wv[e], f16);
                      v[8 * uu + 2 * e] += rv.x;
                      v[8 * uu + 2 * e + 1] += rv.y;
                    }
                  }
                }
                if (has_mask) {
#pragma unroll
                  for (int uu = 0; uu < CW / 8; ++uu) {
                    const uint4 qv =
                        *reinterpret_cast<const uint4*>(sgen + mask_off + (((sp_u0 + (uint32_t)(h * (CW / 8) + uu)) ^ sp_xor) << 4));
                    const uint32_t wv[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                      // sign test on the raw 16-bit patterns: valid for bf16 and IEEE half alike (x > 0 <=> int16(x) > 0)
                      v[8 * uu + 2 * e] = (int)(wv[e] << 16) > 0 ? v[8 * uu + 2 * e] : 0.f;
                      v[8 * uu + 2 * e + 1] = (int)(wv[e] & 0xFFFF0000u) > 0 ? v[8 * uu + 2 * e + 1] : 0.f;
                    }
                  }
                }
                if (h == NH - 1) {
                  __syncwarp();
                  if (lane == 0) mbar_arrive(sideEmpty + 8 * ss);
                }
              }
              // (the ReLU rides on the 16-bit conversion below)
              // staging row: a step whose first slot is a phantom stages its one real row first
              const uint32_t srow = real[0] ? (uint32_t)u : 0u;
              uint8_t* ogen = smem_raw + (w_o + slot * G::WSLOT + srow * (32u * OSPAN) - raw) + p_off;
#pragma unroll
              for (int uu = 0; uu < CW / 8; ++uu) {
                uint4 pk;
                uint32_t* hp = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  hp[e] = a.relu ? pack_h2_relu(v[8 * uu + 2 * e], v[8 * uu + 2 * e + 1], f16)
                                 : pack_h2(v[8 * uu + 2 * e], v[8 * uu + 2 * e + 1], f16);
                *reinterpret_cast<uint4*>(ogen + (((p_u0 + (uint32_t)(h * (CW / 8) + uu)) ^ p_xor) << 4)) = pk;
                if (E_POOL) {
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const int pi = h * (CW / 2) + uu * 4 + e;
                    pmax[pi] = u == 0 ? hp[e] : max_h2(pmax[pi], hp[e], f16);
                  }
                }
              }
            }
          }
        }
        if (E_POOL && any_real) {
          // 2x2 max-pool of the stored values: the vertical max is in pmax, the horizontal neighbour is lane ^ 1; the
          // even lane stages pooled pixel lane / 2 (16 pixels per warp).  (Segments start and end on even rows.)
          uint8_t* pgen = smem_raw + (w_p + slot * G::PSLOT - raw) + pp_off;
#pragma unroll
          for (int uu = 0; uu < CO / 8; ++uu) {
            uint4 pk;
            uint32_t* hp = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const uint32_t o_ = __shfl_xor_sync(0xffffffffu, pmax[uu * 4 + e], 1);
              hp[e] = max_h2(pmax[uu * 4 + e], o_, f16);
            }
            if ((lane & 1) == 0) *reinterpret_cast<uint4*>(pgen + (((pp_u0 + (uint32_t)uu) ^ pp_xor) << 4)) = pk;
          }
        }
        if (trole < 4) GTRACE(trole, (int)o, 4);
        // the zeroed blocks go back to the issuer
        tg_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(blkEmpty + 8 * (o & (NBLK - 1)));
          if (gsz == 2) mbar_arrive(blkEmpty + 8 * ((o + 1u) & (NBLK - 1)));
        }
        if (trole < 4) GTRACE(trole, (int)o, 5);
        if (any_real) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (trole < 4) GTRACE(trole, (int)o, 6);
#ifdef DG_ROWG_NOSTORE  // timing experiment only: nothing is written
          if (false) {
#else
          if (lane == 0) {
#endif
            const int row = si.r0 + (real[0] ? i0 : i0 + 1);
            if (real[0] && real[1]) tma_store_4d(&tm.out2, w_o + slot * G::WSLOT, 0, xw / (int)PG, row, si.n);
            else tma_store_4d(&tm.out1, w_o + slot * G::WSLOT, 0, xw / (int)PG, row, si.n);
            if (E_POOL) tma_store_4d(&tm.pool, w_p + slot * G::PSLOT, 0, (xw >> 1) / (int)PG, row >> 1, si.n);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          if (trole < 4) GTRACE(trole, (int)o, 7);
          ++nstaged;
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // this warp's output rows are written
  }

  tc_fence_before();
  __syncthreads();
  GTRACE_DUMP;
  if (warp == CTRL_W0 + 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_g = nullptr;
DgPerDevice g_dev_g;
thread_local int g_sms_g = 148;
constexpr uint32_t SMEM_BUDGET_G = 226 * 1024 - GTRACE_SMEM;

CUtensorMapSwizzle swz_for(uint32_t span) {
  return span == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : span == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}

// (C, W, H, N) bf16 tensor, box (box_c channels starting anywhere, box_w pixels, box_h rows)
int make_map_g(CUtensorMap* tm, const void* p, int C, int W, int H, int N, int box_c, int box_w, int box_h,
               const char* what) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = g_encode_g(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p), dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swz_for((uint32_t)box_c * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    depgan_set_error(std::string("cuTensorMapEncodeTiled(generic row kernel, ") + what + ") failed: " + std::to_string((int)r));
    return -1;
  }
  return 0;
}

int make_w_map_g(CUtensorMap* tm, const void* p, int Cin, int rows, int box_c, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_c, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode_g(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(p), dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swz_for((uint32_t)box_c * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    depgan_set_error("cuTensorMapEncodeTiled(generic row kernel, weights) failed: " + std::to_string((int)r));
    return -1;
  }
  return 0;
}

int epi_of(const ConvArgs& a) {
  return (a.film_g ? 1 : 0) | ((a.add_src || a.mask_src) ? 2 : 0) | (a.pool_out ? 4 : 0);
}

template <int KS, int CK, int CO>
bool plan_g(const ConvArgs& a, RowgGeom* g, uint32_t* smem) {
  typedef RG<KS, CK, CO> G;
  if (a.C0 % CK || a.C1 % CK) return false;
  const int nchunks = (a.C0 + a.C1) / CK;
  g->nchunk0 = a.C0 / CK; g->nchunk1 = a.C1 / CK;
  g->cblocks = a.W / BWG;
  g->n_side = a.film_g ? 1 : (a.add_src ? 1 : 0) + (a.mask_src ? 1 : 0);
  g->pool = a.pool_out ? 1 : 0;
  g->rows_total = (long long)a.N * g->cblocks * a.H;
  const uint32_t wbytes = ((uint32_t)KS * nchunks * G::B_TILE + 1023u) & ~1023u;
  const uint32_t fixed = 1024 + wbytes + 8u * G::NSLOT * G::WSLOT + (g->pool ? 8u * G::NSLOT * G::PSLOT : 0u) +
                         (uint32_t)G::NSIDE * g->n_side * G::SIDE_ROW + 8u * (2 * 16 + 1 + NQG + G::NBLK + 2 * NSIDEG) + 64 +
                         8 * NREC + (2 * CO + 8 * 4 * CO) * 4 + 64;
  const uint32_t stage = (uint32_t)nchunks * G::A_STAGE;  // one ring stage = one input row, all its chunks
  if (fixed + 2u * stage > SMEM_BUDGET_G) return false;
  int na = (int)((SMEM_BUDGET_G - fixed) / stage);
  static const int na_cap = getenv("DEPGAN_ROWG_NA") ? atoi(getenv("DEPGAN_ROWG_NA")) : 8;  // A/B switch
  if (na > na_cap) na = na_cap;
  g->na = na;
  // measured (profiles/r02_rowg_pf_sweep.txt): L2 prefetch 8 / 16 / 32 rows ahead is 2-12 % SLOWER on every shape and a
  // 16-deep ring equals the 8-deep one, so neither the HBM latency nor the ring depth bounds this kernel; off by default
  static const int pf_env = getenv("DEPGAN_ROWG_PF") ? atoi(getenv("DEPGAN_ROWG_PF")) : 0;  // A/B switch
  g->pf = pf_env;
  g->batch = 1;  // measured: batches of 4 were slower (0.119 -> 0.135 ms on the 16 -> 16 layer): stages are then released in bursts
  *smem = fixed + (uint32_t)na * stage;
  return true;
}

template <int KS, int CK, int CO, int EPI, bool F16>
int launch_g(const ConvArgs& a, const RowgGeom& g, uint32_t smem, cudaStream_t st) {
  typedef RG<KS, CK, CO> G;
  static DgPerDevice attr_done;  // one opt-in per (instantiation, device)
  int dev = 0;
  bool first = false;
  DG_TRY(dg_device_enter(attr_done, &dev, &first));
  if (first) {
    DG_CHECK_CUDA(cudaFuncSetAttribute(conv_rowg_kernel<KS, CK, CO, EPI, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       227 * 1024 - GTRACE_SMEM));
    dg_device_mark(attr_done, dev);
  }
  RowgMaps tm;
  DG_TRY(make_map_g(&tm.a0, a.in0, a.C0, a.W, a.H, a.N, CK, G::TILE_PX, 1, "input 0"));
  if (a.C1 > 0) DG_TRY(make_map_g(&tm.a1, a.in1, a.C1, a.W, a.H, a.N, CK, G::TILE_PX, 1, "input 1"));
  else tm.a1 = tm.a0;
  DG_TRY(make_w_map_g(&tm.b, a.w_tc, a.C0 + a.C1, KS * KS * CO, CK, CO));
  // output, side and pooled tensors as rows of 128-byte super pixels: (C * PG, W / PG, H, N)
  constexpr int PG = G::PG, SC = CO * G::PG;
  DG_TRY(make_map_g(&tm.out1, a.out, SC, a.W / PG, a.H, a.N, SC, 32 / PG, 1, "output"));
  DG_TRY(make_map_g(&tm.out2, a.out, SC, a.W / PG, a.H, a.N, SC, 32 / PG, G::RP, "output (row pair)"));
  tm.s0 = tm.s1 = tm.pool = tm.out1;
  const void* side[2] = {nullptr, nullptr};
  if (a.film_g) side[0] = a.res;
  else {
    int k = 0;
    if (a.add_src) side[k++] = a.add_src;
    if (a.mask_src) side[k++] = a.mask_src;
  }
  if (side[0]) DG_TRY(make_map_g(&tm.s0, side[0], SC, a.W / PG, a.H, a.N, SC, BWG / PG, 1, "side input"));
  if (side[1]) DG_TRY(make_map_g(&tm.s1, side[1], SC, a.W / PG, a.H, a.N, SC, BWG / PG, 1, "side input"));
  if (a.pool_out)
    DG_TRY(make_map_g(&tm.pool, a.pool_out, SC, a.W / 2 / PG, a.H / 2, a.N, SC, 16 / PG, 1, "pooled output"));
  const long long want = (g.rows_total + 7) / 8;  // at least a few rows per CTA, so the halo rows stay cheap
  const int grid = want < g_sms_g ? (int)(want < 1 ? 1 : want) : g_sms_g;
  DG_CHECK_CUDA(dg_launch_pdl(conv_rowg_kernel<KS, CK, CO, EPI, F16>, dim3(grid), dim3(RG_THREADS), smem, st, tm, a, g));
  DG_LAUNCH_CHECK();
  static const bool dbg_sync = getenv("DEPGAN_DEBUG_SYNC") != nullptr;
  if (dbg_sync) {
    cudaError_t e2 = cudaStreamSynchronize(st);
    if (e2 != cudaSuccess) {
      char buf[384];
      snprintf(buf, sizeof buf, "conv_rowg_kernel<%d,%d,%d,%d> failed (%s): N=%d H=%d W=%d C0=%d C1=%d | grid=%d na=%d "
               "side=%d smem=%u", KS, CK, CO, EPI, cudaGetErrorString(e2), a.N, a.H, a.W, a.C0, a.C1, grid, g.na, g.n_side,
               smem);
      depgan_set_error(buf);
      return -1;
    }
  }
  return 0;
}

int rowg_init() {
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  if (!g_encode_g) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    DG_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) {
      depgan_set_error("cuTensorMapEncodeTiled entry point not available");
      return -1;
    }
    g_encode_g = reinterpret_cast<EncodeTiledFn>(fn);
  }
  int dev = 0;
  bool first = false;
  DG_TRY(dg_device_enter(g_dev_g, &dev, &first));
  if (first) dg_device_mark(g_dev_g, dev);
  if (dev < 64) g_sms_g = g_dev_g.sms[dev];
  else DG_CHECK_CUDA(cudaDeviceGetAttribute(&g_sms_g, cudaDevAttrMultiProcessorCount, dev));
  return 0;
}

// The instantiated (kernel size, channel chunk, output channels, epilogue) combinations.  run = false: query only.
// Returns 1 when the shape is taken (and, with run, launched), 0 when it is not a case, < 0 on error.
int dispatch_g(const ConvArgs& a, cudaStream_t st, bool run) {
  const int epi = epi_of(a);
  const int cin = a.C0 + a.C1;
  RowgGeom g;
  uint32_t smem = 0;
  const bool f16 = a.in_dt == DT_F16;
#define RG_CASE(KS_, CK_, CO_, EPI_, F16_)                                        \
  if (a.ks == KS_ && a.Cout == CO_ && cin % CK_ == 0 && epi == EPI_ && f16 == F16_) { \
    if (!plan_g<KS_, CK_, CO_>(a, &g, &smem)) return 0;                           \
    if (!run) return 1;                                                           \
    const int r_ = launch_g<KS_, CK_, CO_, EPI_, F16_>(a, g, smem, st);           \
    return r_ < 0 ? r_ : 1;                                                       \
  }
  // critic 5x5 layers: 16 -> 16 (dis_0b), 16 -> 32 (dis_1a), 32 -> 32 (dis_1b) and their data gradients (32 -> 16)
  if (a.ks == 5 && cin == 16) {
    RG_CASE(5, 16, 16, 0, false) RG_CASE(5, 16, 16, 2, false) RG_CASE(5, 16, 16, 4, false)
    RG_CASE(5, 16, 32, 0, false) RG_CASE(5, 16, 32, 2, false)
  }
  if (a.ks == 5 && cin == 32) {
    RG_CASE(5, 32, 16, 2, false) RG_CASE(5, 32, 16, 0, false)
    RG_CASE(5, 32, 32, 0, false) RG_CASE(5, 32, 32, 2, false) RG_CASE(5, 32, 32, 4, false)
  }
  // generator 3x3 32 -> 32 at full resolution with the fused max-pool (conv2d_gen_1 + maxpool2d_gen_0, TG:408-409); the
  // other epilogues of this shape belong to conv_row.cu, which is asked first
  if (a.ks == 3 && cin == 32) {
    RG_CASE(3, 32, 32, 4, false) RG_CASE(3, 32, 32, 4, true) RG_CASE(3, 32, 32, 0, false)
  }
  // generator 3x3 layers with 64 output channels at half resolution: 32 -> 64, 64 -> 64 (two 32-channel chunks)
  if (a.ks == 3 && (cin == 32 || cin == 64)) {
    RG_CASE(3, 32, 64, 0, false) RG_CASE(3, 32, 64, 1, false) RG_CASE(3, 32, 64, 2, false)
    RG_CASE(3, 32, 64, 0, true) RG_CASE(3, 32, 64, 1, true)
  }
#undef RG_CASE
  return 0;
}

}  // namespace

bool conv_rowg_supported(const ConvArgs& a) {
  static const bool off = getenv("DEPGAN_NO_ROWG") != nullptr;  // A/B switch: these layers through conv_tc_kernel
  if (off) return false;
  if (!dt_is_half(a.in_dt) || a.out_dt != a.in_dt) return false;
  if (a.deconv || a.head_w || a.out_pre || !a.out || !a.w_tc) return false;
  if (a.W % BWG || a.H < 1) return false;
  if (a.C1 > 0 && !a.in1) return false;
  if (a.film_g && (a.add_src || a.mask_src || !a.res || a.pool_out)) return false;
  if (a.pool_out && (a.add_src || a.mask_src || (a.H & 1))) return false;
  return dispatch_g(a, nullptr, false) == 1;
}

int conv_fwd_rowg(const ConvArgs& a, cudaStream_t st) {
  if (a.N <= 0) return 0;
  DG_TRY(rowg_init());
  DG_REQUIRE(conv_rowg_supported(a), "conv_fwd_rowg: unsupported shape");
  const int r = dispatch_g(a, st, true);
  if (r < 0) return r;
  DG_REQUIRE(r == 1, "conv_fwd_rowg: no instantiation for this shape");
  return 0;
}

#ifdef DG_ROWG_TRACE
extern "C" int depgan_dbg_set_rowg_trace(long long* p) {
  return (int)cudaMemcpyToSymbol(g_rowg_trace, &p, sizeof(p));
}
#endif
