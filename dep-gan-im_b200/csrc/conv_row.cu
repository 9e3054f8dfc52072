// Row-streaming tcgen05 convolution for the 32-output-channel 3x3 layers at full resolution (W a multiple of 256):
// conv2d_gen_1, conv2d_gen_noise_m1 / _p1 (FiLM), conv2d_gen_16 (96 -> 32) and conv2d_gen_17 (+ the 1x1
// gen_segmentation head) of Gen_UNet2D (TG:398-409, 482-495; TU:332-343, 411-424), and the same shapes of the
// data-gradient passes.  Same arithmetic contract as conv_tc_kernel (ConvArgs, common.cuh); only the mapping differs.
//
// Why a second mapping.  An SS-mode tcgen05.mma re-reads its A slice (128 pixels x 16 channels = 4 KB = 32 shared-
// memory wavefronts) for every instruction.  With N = Cout = 32 the tensor pipe needs 16 clk per instruction, so
// conv_tc_kernel's tap-by-tap stream is bound by the operand reads (40 wavefronts per MMA; ncu:
// l1tex__data_pipe_tc_wavefronts_mem_shared at 83-93 % of peak, tensor pipe 36-39 %).  Here the three vertical taps
// are stacked on N: one MMA of N = 96 multiplies an image-row slice by the weights of (dy = 2,1,0; dx fixed), so the
// A slice is read once per three taps (56 wavefronts per 3 taps instead of 120).
//
// GEMM view.  M = 128 consecutive pixels of ONE image row j, N = 96 = (dy, cout), K = (dx, cin).  For every input row
// j of a segment of consecutive output rows (plus one halo row above and below; rows outside the image are zero
// through TMA out-of-bounds fill) the MMAs compute
//     D_j[x][(dy, co)] = sum_dx sum_ci  X[j][x + dx - 1][ci] * W[dy][dx][ci][co],
// the dx shift being a different start address of the [130 pixels][32 channels] row tile in shared memory (64-byte
// swizzle, as in conv_tc_kernel), and output row r is  out[r] = D_{r-1}[dy=0] + D_r[dy=1] + D_{r+1}[dy=2].
// That sum is done by the tensor core: TMEM holds, per 128-pixel column block, a ring of 8 blocks of 32 columns, block
// (o mod 8) belongs to output row o, and the weight rows are stacked in the order (dy = 2, 1, 0), so the 96 accumulator
// columns of input row j ARE the three consecutive blocks of output rows j-1, j, j+1 -- every MMA accumulates (the
// epilogue hands a block back zeroed with tcgen05.st).  Rows at the segment edges use the matching 32- / 64-row part of
// the weight tile, a range that wraps around the ring is issued in two parts.  The epilogue thread of pixel x reads
// only the 32 final columns of its TMEM lane, every input row is read from L2 / HBM once per segment, and there are
// no shuffles and no halo pixels.
//
// Two column blocks side by side.  MMAs into the same accumulator serialise on the tensor pipe's latency (measured:
// 124 clk per N = 96 instruction when every instruction hits the same columns, against 56 clk of operand reads), so a
// CTA works on the left and the right 128-pixel block of a 256-pixel row pair at once and the issuer alternates between
// the two independent accumulator rings.
//
// Work split.  The N * (W / 256) * H output rows are cut into one contiguous range per CTA (no tail wave; segments
// end at image borders), each costing two halo input rows.
//
// Roles (12 warps): warps 0-3 / 4-7 epilogue of the left / right block -- eight independent agents: warp w owns TMEM
// lane quarter w % 4 (32 pixels), handles output rows two at a time (two TMEM loads in flight, one shared-memory fence
// and ONE TMA store per pair), stages its pixels in its own swizzled slots, so the row loop has no block- or group-wide
// barrier; warp 8 TMA producer of the row tiles, warp 9 MMA issuer (one elected lane) + TMEM owner, warp 10 TMA
// producer of the epilogue's side rows (FiLM residual, add / mask sources).  Weights ([dx][chunk][96][32] bf16,
// <= 54 KB) are resident for the CTA's lifetime.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "conv_tc_kernel.cuh"

using namespace convtc;

namespace {

constexpr int RW_THREADS = 384;
constexpr int BW = 128;                 // pixels per M block
constexpr int TILE_PX = BW + 2;         // row tile with one halo pixel on each side
constexpr uint32_t A_STAGE = 9216;      // 130 * 64 B = 8320, rounded to 1024
constexpr uint32_t A_TX = TILE_PX * 64;
constexpr uint32_t B_TILE = 96 * 64;    // one (dx, chunk) weight tile: 96 rows (dy, cout) x 32 channels
constexpr int NBLK = 8;                 // TMEM ring per column block: 8 blocks of 32 columns (one per output row in flight)
constexpr int NQ = 32;                  // "input row done" barriers (ring; see the ABA note at the epilogue wait)
constexpr int NSLOT = 2;                // output staging slots per epilogue warp (one slot = a pair of rows, 4 KB)
constexpr uint32_t WSLOT = 2 * 32 * 64; // one warp's staging slot: 2 rows x 32 pixels x 32 channels bf16
constexpr int NSIDE = 4;                // side-row stages (one stage = one 128-pixel row of one column block)
constexpr uint32_t SIDE_ROW = BW * 64;  // one side row: 128 pixels x 32 channels bf16

#ifdef DG_ROW_TRACE  // timing experiment only: event clocks of CTA 0 (roles: 0 producer, 1 issuer, 2 epilogue warp 0, 3 warp 4)
__device__ long long* g_row_trace = nullptr;  // [4 roles][64 rows][8 events]
#define RTRACE(role_, row_, ev_)                                                                      \
  do {                                                                                                \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (row_) < 64 && g_row_trace)                     \
      g_row_trace[((role_) * 64 + (row_)) * 8 + (ev_)] = clock64();                                   \
  } while (0)
#else
#define RTRACE(role_, row_, ev_) do {} while (0)
#endif

struct RowGeom {
  int nchunk0, nchunk1;  // 32-channel chunks from in0 / in1
  int pairs_w;           // 256-pixel column pairs per image row
  int na;                // row-tile ring depth
  int n_side;            // side tensors per output row (0..2)
  int stage_out;         // bf16 `out` is written
  long long rows_total;  // N * pairs_w * H output rows (of 256 pixels)
};

struct RowMaps {
  CUtensorMap a0, a1, b, out1, out2, s0, s1;
};

// 32 accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_ld32_fence(uint32_t (&r)[32]) {  // ties the registers to the preceding wait
#pragma unroll
  for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(r[i])::"memory");
}
__device__ __forceinline__ void tc_ld_wait_all() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// zeroes 32 columns of this thread's TMEM lane (asynchronous: tc_st_wait before the block is handed back)
__device__ __forceinline__ void tc_st32_zero(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
      "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
      ::"r"(taddr), "r"(z)
      : "memory");
}
__device__ __forceinline__ void tc_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// The CTA's contiguous range of output rows (256 pixels wide), cut into segments at image borders.  A unit is one
// (slice, column pair); every role walks the same segments.
struct SegIter {
  long long pos, end;
  int H, pairs_w;
  int n, pc, r0, cnt;  // current segment: slice, column pair, first output row, number of output rows
  __device__ __forceinline__ void init(long long total, int H_, int pairs_w_) {
    H = H_; pairs_w = pairs_w_;
    pos = total * (long long)blockIdx.x / (long long)gridDim.x;
    end = total * (long long)(blockIdx.x + 1) / (long long)gridDim.x;
    decode();
  }
  __device__ __forceinline__ void decode() {
    if (pos >= end) { cnt = 0; return; }
    const long long unit = pos / H;
    r0 = (int)(pos - unit * H);
    n = (int)(unit / pairs_w);
    pc = (int)(unit - (long long)n * pairs_w);
    const long long left = end - pos;
    cnt = (long long)(H - r0) < left ? H - r0 : (int)left;
  }
  __device__ __forceinline__ bool valid() const { return pos < end; }
  __device__ __forceinline__ void next() { pos += cnt; decode(); }
};

// ---------------------------------------------------------------------------------------------------------
// EPI bit 0: FiLM residual side row; bit 1: add / mask side rows.  HEAD: fused 1x1 head on the 32 outputs.
// ---------------------------------------------------------------------------------------------------------
// F16: IEEE-half storage of activations / weights instead of bfloat16 (compile time: see conv_tc_kernel.cuh).
template <int EPI, bool HEAD, bool F16 = false>
__global__ void __launch_bounds__(RW_THREADS, 1) conv_row_kernel(const __grid_constant__ RowMaps tm, const ConvArgs a,
                                                                 const RowGeom g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const int nchunks = g.nchunk0 + g.nchunk1;
  const uint32_t b_base = base;                                      // weights: [dx][chunk] tiles of 6 KB
  const uint32_t a_base = b_base + 3u * nchunks * B_TILE;            // row-tile ring
  const uint32_t o_base = a_base + (uint32_t)g.na * 2u * A_STAGE;    // 8 warps x NSLOT x 4 KB
  const uint32_t s_base = o_base + 8u * NSLOT * WSLOT;               // NSIDE stages x n_side rows
  const uint32_t bar_base = s_base + (uint32_t)NSIDE * g.n_side * SIDE_ROW;
  const uint32_t fullA = bar_base, emptyA = fullA + 8 * g.na;
  const uint32_t fullB = emptyA + 8 * g.na;
  const uint32_t rowDone = fullB + 8, blkEmpty = rowDone + 8 * NQ;   // blkEmpty[NBLK]: both column blocks' warps arrive
  const uint32_t sideFull = blkEmpty + 8 * NBLK, sideEmpty = sideFull + 8 * NSIDE;
  const uint32_t tmem_slot = sideEmpty + 8 * NSIDE;
  const uint32_t f_off = (tmem_slot + 16 + 15u) & ~15u;  // floats (16-byte aligned): scale, shift, head, per-warp tables
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  float* s_scale = reinterpret_cast<float*>(smem_raw + (f_off - raw));
  float* s_shift = s_scale + 32;
  float4* s_head = reinterpret_cast<float4*>(s_shift + 32);
  float* s_tab = reinterpret_cast<float*>(s_head + 32);  // [8 warps][2 buffers][2][32]

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < g.na; ++i) { mbar_init(fullA + 8 * i, 1); mbar_init(emptyA + 8 * i, 1); }
    mbar_init(fullB, 1);
    for (int i = 0; i < NQ; ++i) mbar_init(rowDone + 8 * i, 1);
    for (int i = 0; i < NBLK; ++i) mbar_init(blkEmpty + 8 * i, 8);
    for (int i = 0; i < NSIDE; ++i) { mbar_init(sideFull + 8 * i, 1); mbar_init(sideEmpty + 8 * i, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == CTRL_W0 + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x < 32) {
    s_scale[threadIdx.x] = a.scale ? a.scale[threadIdx.x] : 1.f;
    s_shift[threadIdx.x] = a.shift ? a.shift[threadIdx.x] : 0.f;
    if (HEAD) {
      float hv[4] = {0.f, 0.f, 0.f, 0.f};
      for (int k = 0; k < a.head_nc; ++k) hv[k] = a.head_w[(size_t)threadIdx.x * a.head_nc + k];
      s_head[threadIdx.x] = make_float4(hv[0], hv[1], hv[2], hv[3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (warp < 8) {  // all 16 blocks start zeroed: warp w clears lane quarter w % 4 of the 8 blocks of column block w / 4
    for (int b = 0; b < NBLK; ++b)
      tc_st32_zero(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(((warp >> 2) * NBLK + b) * 32));
    tc_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp >= CTRL_W0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    if (warp == CTRL_W0) {
      // ===== TMA producer: resident weights once, then per input row and chunk the left and the right row tile =====
      if (elect_one()) {
        mbar_expect_tx(fullB, 9u * nchunks * 32u * 64u);
        for (int dx = 0; dx < 3; ++dx)
          for (int c = 0; c < nchunks; ++c) {
            const int kglob = c < g.nchunk0 ? c * 32 : a.C0 + (c - g.nchunk0) * 32;
            for (int dy = 0; dy < 3; ++dy)  // rows stacked (dy = 2, 1, 0): column block k of an MMA = output row j - 2 + k
              tma_load_2d(b_base + (uint32_t)(dx * nchunks + c) * B_TILE + (uint32_t)(2 - dy) * 32u * 64u, &tm.b, fullB,
                          kglob, (dy * 3 + dx) * 32);
          }
      }
      __syncwarp();
      Ring ra;
      SegIter si;
      si.init(g.rows_total, a.H, g.pairs_w);
      for (; si.valid(); si.next()) {
        const int x0 = si.pc * 2 * BW - 1, r0 = si.r0 - 1;
        for (int j = 0; j < si.cnt + 2; ++j) {
          for (int c = 0; c < nchunks; ++c) {
            const bool first = c < g.nchunk0;
            const int ch = (first ? c : c - g.nchunk0) * 32;
            // one ring stage = the left and the right row tile of this (input row, chunk), one barrier for both
            if (c == 0) RTRACE(0, j, 0);
            mbar_wait(emptyA + 8 * ra.idx, ra.phase ^ 1u);
            if (c == 0) RTRACE(0, j, 1);
            if (elect_one()) {
              mbar_expect_tx(fullA + 8 * ra.idx, 2u * A_TX);
              tma_load_4d(a_base + ra.idx * 2u * A_STAGE, first ? &tm.a0 : &tm.a1, fullA + 8 * ra.idx, ch, x0, r0 + j, si.n);
              tma_load_4d(a_base + ra.idx * 2u * A_STAGE + A_STAGE, first ? &tm.a0 : &tm.a1, fullA + 8 * ra.idx, ch,
                          x0 + BW, r0 + j, si.n);
            }
            __syncwarp();
            ra.advance(g.na);
          }
        }
      }
    } else if (warp == CTRL_W0 + 1) {
      // ===== MMA issuer: per input row 6 * nchunks instructions of N = 96 (32 / 64 at the segment edges) per column
      // block, alternating between the two blocks' accumulator rings.  The MMA queue is only about two instructions
      // deep and an mbarrier poll costs ~100 clk even when the phase has already completed, so the waits of the NEXT
      // step (input row, chunk) are issued between the MMA groups of the current one: they run while the tensor pipe
      // works through the queued instructions. =====
      const uint32_t hi = ((uint32_t)(8 * 64) >> 4) | (1u << 14) | (4u << 29);  // SBO 512 B, 64-byte swizzle
      constexpr uint32_t LBO1 = 1u << 16;
      constexpr bool f16_in = F16;
      mbar_wait(fullB, 0);
      tc_fence_after();
      const uint32_t b_lo0 = ((b_base & 0x3FFFFu) >> 4) | LBO1;

      struct Step {
        uint32_t stage, phase;       // row-tile ring stage of this step
        uint32_t d1, idesc1, idesc2, n1, n2, brow0;
        uint32_t q, o_hi;
        int c;
        bool valid;
      };
      Ring ra;
      SegIter si;
      si.init(g.rows_total, a.H, g.pairs_w);
      uint32_t o0 = 0, q = 0, acquired = 0;
      int j = 0, c = 0;
      // geometry of the step (si, j, c); advances the iteration state afterwards
      Step rowc{};  // geometry of the current input row: the same for every channel chunk of the row
      auto make_step = [&]() -> Step {
        Step s{};
        s.valid = si.valid();
        if (!s.valid) return s;
        if (c == 0) {
          const int cnt = si.cnt;
          const int i_lo = j >= 2 ? j - 2 : 0, i_hi = j < cnt ? j : cnt - 1;
          const uint32_t o_lo = o0 + (uint32_t)i_lo;
          s.o_hi = o0 + (uint32_t)i_hi;
          s.brow0 = (uint32_t)(i_lo - (j - 2));
          const uint32_t blk0 = o_lo % NBLK, nblk = s.o_hi - o_lo + 1u;
          s.n1 = blk0 + nblk > (uint32_t)NBLK ? (uint32_t)NBLK - blk0 : nblk;
          s.n2 = nblk - s.n1;
          s.d1 = tmem_base + blk0 * 32u;
          s.idesc1 = make_idesc((int)(32u * s.n1), f16_in);
          s.idesc2 = make_idesc((int)(32u * (s.n2 ? s.n2 : 1u)), f16_in);
          rowc = s;
        } else {
          s = rowc;  // (the 96-channel decoder layer has three chunks per row: its issuer was paced by this set-up)
        }
        s.q = q; s.c = c;
        s.stage = ra.idx; s.phase = ra.phase;
        ra.advance(g.na);
        if (++c == nchunks) {
          c = 0; ++q;
          if (++j == si.cnt + 2) { j = 0; o0 += (uint32_t)si.cnt; si.next(); }
        }
        return s;
      };
      // a block is acquired the first time an input row touches its output row: the previous owner (NBLK output rows
      // earlier) must have been drained and zeroed by its eight epilogue warps
      auto wait_blocks = [&](const Step& s) {
        if (!s.valid) return;
        while (acquired <= s.o_hi) {
          if (acquired >= (uint32_t)NBLK) mbar_wait(blkEmpty + 8 * (acquired % NBLK), ((acquired / NBLK) & 1u) ^ 1u);
          ++acquired;
        }
      };
      auto wait_tiles = [&](const Step& s) {
        if (s.valid) mbar_wait(fullA + 8 * s.stage, s.phase);
      };
      Step cur = make_step();
      wait_blocks(cur);
      wait_tiles(cur);
      tc_fence_after();
      while (cur.valid) {
        if (cur.c == 0) RTRACE(1, (int)cur.q, 0);
        const uint32_t a_lo0 = (((a_base + cur.stage * 2u * A_STAGE) & 0x3FFFFu) >> 4) | LBO1;
        const uint32_t a_lo1 = a_lo0 + (A_STAGE >> 4);
        const uint32_t b_lo = b_lo0 + (uint32_t)cur.c * (B_TILE >> 4) + cur.brow0 * (32u * 64u >> 4);
        Step nxt{};
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          if (elect_one()) {
            const uint32_t bt = b_lo + (uint32_t)dx * (uint32_t)nchunks * (B_TILE >> 4);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const uint64_t db = ((uint64_t)hi << 32) | (bt + 2 * k);
              const uint64_t da0 = ((uint64_t)hi << 32) | (a_lo0 + 4 * dx + 2 * k);
              const uint64_t da1 = ((uint64_t)hi << 32) | (a_lo1 + 4 * dx + 2 * k);
              tc_mma(cur.d1, da0, db, cur.idesc1, 1u);
              tc_mma(cur.d1 + NBLK * 32u, da1, db, cur.idesc1, 1u);
              if (cur.n2) {
                const uint64_t db2 = ((uint64_t)hi << 32) | (bt + cur.n1 * (32u * 64u >> 4) + 2 * k);
                tc_mma(tmem_base, da0, db2, cur.idesc2, 1u);
                tc_mma(tmem_base + NBLK * 32u, da1, db2, cur.idesc2, 1u);
              }
            }
            if (dx == 2) {
              tc_commit(emptyA + 8 * cur.stage);
              if (cur.c == nchunks - 1) tc_commit(rowDone + 8 * (cur.q % NQ));
            }
          }
          __syncwarp();
          // the next step's waits ride on the queued MMAs
          if (dx == 0) { nxt = make_step(); wait_blocks(nxt); }
          if (dx == 1) { wait_tiles(nxt); tc_fence_after(); }
        }
        if (cur.c == nchunks - 1) RTRACE(1, (int)cur.q, 3);
        cur = nxt;
      }
    } else if (warp == CTRL_W0 + 2) {
      // ===== TMA producer of the side rows (FiLM residual, or add / mask sources): per output row one 128-pixel row for
      // the left and one for the right column block =====
      if (EPI != 0 && g.n_side > 0) {
        Ring rs;
        SegIter si;
        si.init(g.rows_total, a.H, g.pairs_w);
        for (; si.valid(); si.next()) {
          for (int i = 0; i < si.cnt; ++i) {
#pragma unroll
            for (int sb = 0; sb < 2; ++sb) {
              mbar_wait(sideEmpty + 8 * rs.idx, rs.phase ^ 1u);
              if (elect_one()) {
                const uint32_t dst = s_base + (uint32_t)(rs.idx * g.n_side) * SIDE_ROW;
                const int x = (si.pc * 2 + sb) * BW;
                mbar_expect_tx(sideFull + 8 * rs.idx, (uint32_t)g.n_side * SIDE_ROW);
                tma_load_4d(dst, &tm.s0, sideFull + 8 * rs.idx, 0, x, si.r0 + i, si.n);
                if (g.n_side == 2) tma_load_4d(dst + SIDE_ROW, &tm.s1, sideFull + 8 * rs.idx, 0, x, si.r0 + i, si.n);
              }
              __syncwarp();
              rs.advance(NSIDE);
            }
          }
        }
      }
    }
  } else {
    // ===== epilogue warps: eight independent agents =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    constexpr bool E_RES = (EPI & 1) != 0, E_AM = (EPI & 2) != 0;
    const int sb = warp >> 2;        // column block (0 left, 1 right) this warp belongs to
    const int qd = warp & 3;         // TMEM lane quarter
    const int m = qd * 32 + lane;    // pixel of the 128-pixel block
    const bool has_add = E_AM && a.add_src != nullptr;
    const bool has_mask = E_AM && a.mask_src != nullptr;
    const bool stage_out = g.stage_out != 0;
    const bool side = EPI != 0 && g.n_side > 0;
    constexpr bool f16 = F16;  // 16-bit storage format (bf16 otherwise)
    const int head_nc = HEAD ? a.head_nc : 0;
    float* tab = s_tab + warp * 128;  // [2 buffers][scale 32 | shift 32]
    // staging: this warp's NSLOT slots of [2 rows][32 pixels][64 B], 16-byte units XOR-swizzled like the TMA (64-byte mode)
    const uint32_t w_o = o_base + (uint32_t)warp * NSLOT * WSLOT;
    const uint32_t p_off = (uint32_t)lane * 64u;
    const uint32_t p_xor = (uint32_t)(lane >> 1) & 3u;
    // side rows are [128 pixels][64 B] with the same swizzle
    const uint32_t sp_off = (uint32_t)m * 64u;
    const uint32_t sp_xor = (uint32_t)(m >> 1) & 3u;
    const uint32_t mask_off = has_add ? SIDE_ROW : 0u;
    const uint32_t t_lane = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(sb * NBLK * 32);
    const uint32_t my_empty = blkEmpty;  // 8 arrivals per block: the four warps of each column block

    uint32_t q0 = 0;        // input-row sequence number of the segment's first input row
    uint32_t o0 = 0;        // output-row sequence number of the segment's first output row
    uint32_t npair = 0;     // row pairs this warp has staged (staging slot ring)
    int bseq = 0;
    SegIter si;
    si.init(g.rows_total, a.H, g.pairs_w);
    for (; si.valid(); q0 += (uint32_t)(si.cnt + 2), o0 += (uint32_t)si.cnt, ++bseq, si.next()) {
      const int cnt = si.cnt;
      // per-segment table: v = acc * sc + sh   (BN folded; FiLM: the conditioning affine folded into the same pair)
      float* tb = tab + (bseq & 1) * 64;
      {
        float sc = s_scale[lane], sh = s_shift[lane];
        if (E_RES) {
          const float fg = __ldg(a.film_g + (size_t)si.n * a.film_stride + lane);
          const float fb = __ldg(a.film_b + (size_t)si.n * a.film_stride + lane);
          sh = fmaf(sh, fg, fb);
          sc *= fg;
        }
        tb[lane] = sc;
        tb[32 + lane] = sh;
      }
      __syncwarp();
      const int xblk = (si.pc * 2 + sb) * BW;
      for (int i = 0; i < cnt; i += 2) {
        const int gsz = cnt - i >= 2 ? 2 : 1;   // rows in this group
        const uint32_t o = o0 + (uint32_t)i;
        const int trole = qd == 0 ? 2 + sb : 99;
        if (trole < 4) RTRACE(trole, (int)o, 0);
        // Output row i is complete when input row i + 2 of the segment has been accumulated; the pair waits for the later
        // one (MMAs complete in order).  NQ = 32 barriers: a barrier is committed again 32 input rows later, which touches
        // output rows >= o + 28 and therefore needed the block of output row o + NBLK = this row's block, i.e. this
        // warp's arrival below.
        const uint32_t qd_ = q0 + (uint32_t)(i + gsz - 1) + 2u;
        mbar_wait(rowDone + 8 * (qd_ % NQ), (qd_ / NQ) & 1u);
        tc_fence_after();
        if (trole < 4) RTRACE(trole, (int)o, 1);
        const uint32_t blkA = t_lane + (o % NBLK) * 32u, blkB = t_lane + ((o + 1u) % NBLK) * 32u;
        uint32_t va[2][32];
        tc_ld32_issue(blkA, va[0]);
        if (gsz == 2) tc_ld32_issue(blkB, va[1]);
        tc_ld_wait_all();
        tc_ld32_fence(va[0]);
        tc_ld32_fence(va[1]);
        if (trole < 4) RTRACE(trole, (int)o, 2);
        tc_st32_zero(blkA);        // hand the blocks back zeroed (every MMA accumulates); waited for after the math
        if (gsz == 2) tc_st32_zero(blkB);
        const uint32_t slot = npair % NSLOT;
        if (stage_out) {
          // the TMA store that last read this slot was issued NSLOT pairs ago: at most NSLOT - 1 newer groups may be pending
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NSLOT - 1) : "memory");
          __syncwarp();
        }
        if (trole < 4) RTRACE(trole, (int)o, 3);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (u < gsz) {
            const int row = si.r0 + i + u;
            float v[32];
            // ---- affine (+ FiLM) ----
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
              const float4 sc4 = reinterpret_cast<const float4*>(tb)[k4];
              const float4 sh4 = reinterpret_cast<const float4*>(tb + 32)[k4];
              v[4 * k4 + 0] = __uint_as_float(va[u][4 * k4 + 0]); v[4 * k4 + 1] = __uint_as_float(va[u][4 * k4 + 1]);
              v[4 * k4 + 2] = __uint_as_float(va[u][4 * k4 + 2]); v[4 * k4 + 3] = __uint_as_float(va[u][4 * k4 + 3]);
              fma_f32x2(v[4 * k4 + 0], v[4 * k4 + 1], sc4.x, sc4.y, sh4.x, sh4.y);   // packed fp32x2: the bits of fmaf
              fma_f32x2(v[4 * k4 + 2], v[4 * k4 + 3], sc4.z, sc4.w, sh4.z, sh4.w);
            }
            if (side) {
              // side stage sequence: (output row, column block) in the producer's order
              const uint32_t sq = 2u * (o + (uint32_t)u) + (uint32_t)sb;
              const uint32_t ss = sq % NSIDE;
              mbar_wait(sideFull + 8 * ss, (sq / NSIDE) & 1u);
              const uint8_t* sgen = smem_raw + (s_base + (uint32_t)(ss * g.n_side) * SIDE_ROW - raw) + sp_off;
              if (E_RES) {
#pragma unroll
                for (int uu = 0; uu < 4; ++uu) {
                  const uint4 qv = *reinterpret_cast<const uint4*>(sgen + (((uint32_t)uu ^ sp_xor) << 4));
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
                for (int uu = 0; uu < 4; ++uu) {
                  const uint4 qv = *reinterpret_cast<const uint4*>(sgen + (((uint32_t)uu ^ sp_xor) << 4));
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
                for (int uu = 0; uu < 4; ++uu) {
                  const uint4 qv = *reinterpret_cast<const uint4*>(sgen + mask_off + (((uint32_t)uu ^ sp_xor) << 4));
                  const uint32_t wv[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    // sign test on the raw 16-bit patterns: valid for bf16 and IEEE half alike (x > 0 <=> int16(x) > 0)
                    v[8 * uu + 2 * e] = (int)(wv[e] << 16) > 0 ? v[8 * uu + 2 * e] : 0.f;
                    v[8 * uu + 2 * e + 1] = (int)(wv[e] & 0xFFFF0000u) > 0 ? v[8 * uu + 2 * e + 1] : 0.f;
                  }
                }
              }
              __syncwarp();
              if (lane == 0) mbar_arrive(sideEmpty + 8 * ss);
            }
            // the ReLU rides on the 16-bit conversion unless the fp32 values feed the fused head
            const bool relu_cvt = a.relu && !HEAD;
            if (a.relu && !relu_cvt) {
#pragma unroll
              for (int k = 0; k < 32; ++k) v[k] = fmaxf(v[k], 0.f);
            }
            if (stage_out) {
              uint8_t* ogen = smem_raw + (w_o + slot * WSLOT + (uint32_t)u * (WSLOT / 2) - raw) + p_off;
#pragma unroll
              for (int uu = 0; uu < 4; ++uu) {
                uint4 pk;
                uint32_t* hp = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  hp[e] = relu_cvt ? pack_h2_relu(v[8 * uu + 2 * e], v[8 * uu + 2 * e + 1], f16)
                                   : pack_h2(v[8 * uu + 2 * e], v[8 * uu + 2 * e + 1], f16);
                *reinterpret_cast<uint4*>(ogen + (((uint32_t)uu ^ p_xor) << 4)) = pk;
              }
            }
            if (HEAD) {
              float h0 = 0.f, h1 = 0.f, h2 = 0.f, h3 = 0.f;
#pragma unroll
              for (int k = 0; k < 32; ++k) {
                const float4 hw = s_head[k];
                h0 = fmaf(v[k], hw.x, h0);
                if (head_nc > 1) { h1 = fmaf(v[k], hw.y, h1); h2 = fmaf(v[k], hw.z, h2); h3 = fmaf(v[k], hw.w, h3); }
              }
              const int nc = head_nc;
              float y0 = h0 + __ldg(a.head_b), y1 = 0.f, y2 = 0.f, y3 = 0.f;
              if (nc > 1) y1 = h1 + __ldg(a.head_b + 1);
              if (nc > 2) y2 = h2 + __ldg(a.head_b + 2);
              if (nc > 3) y3 = h3 + __ldg(a.head_b + 3);
              if (a.head_act == 0) {
                y0 = tanhf(y0);
                if (nc > 1) { y1 = tanhf(y1); y2 = tanhf(y2); y3 = tanhf(y3); }
              } else if (a.head_act == 1) {
                float mx = y0;
                if (nc > 1) mx = fmaxf(mx, y1);
                if (nc > 2) mx = fmaxf(mx, y2);
                if (nc > 3) mx = fmaxf(mx, y3);
                y0 = expf(y0 - mx);
                y1 = nc > 1 ? expf(y1 - mx) : 0.f;
                y2 = nc > 2 ? expf(y2 - mx) : 0.f;
                y3 = nc > 3 ? expf(y3 - mx) : 0.f;
                const float inv = 1.0f / (y0 + y1 + y2 + y3);
                y0 *= inv; y1 *= inv; y2 *= inv; y3 *= inv;
              }
              const size_t pix = ((size_t)si.n * a.H + row) * a.W + (size_t)(xblk + m);
              if (nc == 4) {
                *reinterpret_cast<float4*>(a.head_out + pix * 4) = make_float4(y0, y1, y2, y3);
              } else {
                float* op = a.head_out + pix * nc;
                op[0] = y0;
                if (nc > 1) op[1] = y1;
                if (nc > 2) op[2] = y2;
              }
            }
          }
        }
        if (trole < 4) RTRACE(trole, (int)o, 4);
        // the zeroed blocks go back to the issuer
        tc_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(my_empty + 8 * (o % NBLK));
          if (gsz == 2) mbar_arrive(my_empty + 8 * ((o + 1u) % NBLK));
        }
        if (trole < 4) RTRACE(trole, (int)o, 5);
        if (stage_out) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (trole < 4) RTRACE(trole, (int)o, 6);
          if (lane == 0) {
            if (gsz == 2) tma_store_4d(&tm.out2, w_o + slot * WSLOT, 0, xblk + qd * 32, si.r0 + i, si.n);
            else tma_store_4d(&tm.out1, w_o + slot * WSLOT, 0, xblk + qd * 32, si.r0 + i, si.n);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          if (trole < 4) RTRACE(trole, (int)o, 7);
          ++npair;
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // this warp's output rows are written
  }

  tc_fence_before();
  __syncthreads();
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
EncodeTiledFn g_encode_r = nullptr;
DgPerDevice g_dev_r;
thread_local int g_sms_r = 148;
constexpr uint32_t SMEM_BUDGET_R = 226 * 1024;

int make_row_map(CUtensorMap* tm, const void* p, int C, int W, int H, int N, int box_w, int box_h, const char* what) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {32, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = g_encode_r(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p), dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    depgan_set_error(std::string("cuTensorMapEncodeTiled(row kernel, ") + what + ") failed: " + std::to_string((int)r));
    return -1;
  }
  return 0;
}

int make_row_w_map(CUtensorMap* tm, const void* p, int Cin, int rows) {
  cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode_r(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(p), dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    depgan_set_error("cuTensorMapEncodeTiled(row kernel, weights) failed: " + std::to_string((int)r));
    return -1;
  }
  return 0;
}

bool plan_row(const ConvArgs& a, RowGeom* g, uint32_t* smem) {
  const int nchunks = (a.C0 + a.C1) / 32;
  g->nchunk0 = a.C0 / 32; g->nchunk1 = a.C1 / 32;
  g->pairs_w = a.W / (2 * BW);
  g->n_side = a.film_g ? 1 : (a.add_src ? 1 : 0) + (a.mask_src ? 1 : 0);
  g->stage_out = a.out ? 1 : 0;
  g->rows_total = (long long)a.N * g->pairs_w * a.H;
  const uint32_t fixed = 1024 + 3u * nchunks * B_TILE + 8u * NSLOT * WSLOT + (uint32_t)NSIDE * g->n_side * SIDE_ROW +
                         8u * (2 * 16 + 1 + NQ + 2 * NBLK + 2 * NSIDE) + 32 + (64 + 128 + 8 * 128) * 4 + 64;
  if (fixed + 2 * 2 * A_STAGE > SMEM_BUDGET_R) return false;
  int na = (int)((SMEM_BUDGET_R - fixed) / (2 * A_STAGE));  // one stage = the (left, right) pair of row tiles
  if (na > 8) na = 8;
  g->na = na;
  *smem = fixed + (uint32_t)na * 2u * A_STAGE;
  return true;
}

template <int EPI, bool HEAD, bool F16 = false>
int set_attr_row() {
  DG_CHECK_CUDA(cudaFuncSetAttribute(conv_row_kernel<EPI, HEAD, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  return 0;
}

int conv_row_init() {
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  if (!g_encode_r) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    DG_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) {
      depgan_set_error("cuTensorMapEncodeTiled entry point not available");
      return -1;
    }
    g_encode_r = reinterpret_cast<EncodeTiledFn>(fn);
  }
  int dev = 0;
  bool first = false;
  DG_TRY(dg_device_enter(g_dev_r, &dev, &first));
  if (first) {
    DG_TRY((set_attr_row<0, false>()));
    DG_TRY((set_attr_row<0, true>()));
    DG_TRY((set_attr_row<1, false>()));
    DG_TRY((set_attr_row<2, false>()));
    DG_TRY((set_attr_row<0, false, true>()));
    DG_TRY((set_attr_row<0, true, true>()));
    DG_TRY((set_attr_row<1, false, true>()));
    dg_device_mark(g_dev_r, dev);
  }
  if (dev < 64) g_sms_r = g_dev_r.sms[dev];
  else DG_CHECK_CUDA(cudaDeviceGetAttribute(&g_sms_r, cudaDevAttrMultiProcessorCount, dev));
  return 0;
}

}  // namespace

// The row-streaming kernel takes the 3x3, 32-output-channel layers whose width is a multiple of 256.
bool conv_row_supported(const ConvArgs& a) {
  static const bool off = getenv("DEPGAN_NO_ROW") != nullptr;  // A/B switch: every layer through conv_tc_kernel
  if (off) return false;
  if (!dt_is_half(a.in_dt) || a.out_dt != a.in_dt) return false;
  if (a.in_dt == DT_F16 && (a.add_src || a.mask_src)) return false;  // IEEE half: the inference epilogues only
  if (a.ks != 3 || a.deconv || a.Cout != 32) return false;
  if (a.W % (2 * BW) || a.H < 1) return false;
  if (a.C0 % 32 || a.C1 % 32 || a.C0 < 32 || (a.C0 + a.C1) > 96) return false;
  if (a.C1 > 0 && !a.in1) return false;
  if (!a.w_tc) return false;
  if (a.out_pre || a.pool_out) return false;                       // kept on conv_tc_kernel (training epilogues)
  if (a.film_g && (a.add_src || a.mask_src || !a.res || !a.out)) return false;
  if (a.head_w && (a.film_g || a.add_src || a.mask_src || a.head_nc > 4)) return false;
  if (!a.out && !a.head_w) return false;
  RowGeom g;
  uint32_t smem;
  return plan_row(a, &g, &smem);
}

int conv_fwd_row(const ConvArgs& a, cudaStream_t st) {
  if (a.N <= 0) return 0;
  DG_TRY(conv_row_init());
  RowGeom g;
  uint32_t smem;
  DG_REQUIRE(conv_row_supported(a) && plan_row(a, &g, &smem), "conv_fwd_row: unsupported shape");
  RowMaps tm;
  DG_TRY(make_row_map(&tm.a0, a.in0, a.C0, a.W, a.H, a.N, TILE_PX, 1, "input 0"));
  if (a.C1 > 0) DG_TRY(make_row_map(&tm.a1, a.in1, a.C1, a.W, a.H, a.N, TILE_PX, 1, "input 1"));
  else tm.a1 = tm.a0;
  DG_TRY(make_row_w_map(&tm.b, a.w_tc, a.C0 + a.C1, 9 * 32));
  tm.out1 = tm.out2 = tm.s0 = tm.s1 = tm.a0;
  if (a.out) {
    DG_TRY(make_row_map(&tm.out1, a.out, 32, a.W, a.H, a.N, 32, 1, "output"));
    DG_TRY(make_row_map(&tm.out2, a.out, 32, a.W, a.H, a.N, 32, 2, "output (row pair)"));
  }
  const void* side[2] = {nullptr, nullptr};
  if (a.film_g) side[0] = a.res;
  else {
    int k = 0;
    if (a.add_src) side[k++] = a.add_src;
    if (a.mask_src) side[k++] = a.mask_src;
  }
  if (side[0]) DG_TRY(make_row_map(&tm.s0, side[0], 32, a.W, a.H, a.N, BW, 1, "side input"));
  if (side[1]) DG_TRY(make_row_map(&tm.s1, side[1], 32, a.W, a.H, a.N, BW, 1, "side input"));
  const long long want = (g.rows_total + 3) / 4;  // at least a few rows per CTA, so the two halo rows stay cheap
  const int grid = want < g_sms_r ? (int)(want < 1 ? 1 : want) : g_sms_r;
  const int epi = a.film_g ? 1 : ((a.add_src || a.mask_src) ? 2 : 0);
  cudaError_t e;
  if (a.in_dt == DT_F16) {
    if (a.head_w) e = dg_launch_pdl(conv_row_kernel<0, true, true>, dim3(grid), dim3(RW_THREADS), smem, st, tm, a, g);
    else if (epi == 1) e = dg_launch_pdl(conv_row_kernel<1, false, true>, dim3(grid), dim3(RW_THREADS), smem, st, tm, a, g);
    else e = dg_launch_pdl(conv_row_kernel<0, false, true>, dim3(grid), dim3(RW_THREADS), smem, st, tm, a, g);
  } else if (a.head_w) e = dg_launch_pdl(conv_row_kernel<0, true>, dim3(grid), dim3(RW_THREADS), smem, st, tm, a, g);
  else if (epi == 1) e = dg_launch_pdl(conv_row_kernel<1, false>, dim3(grid), dim3(RW_THREADS), smem, st, tm, a, g);
  else if (epi == 2) e = dg_launch_pdl(conv_row_kernel<2, false>, dim3(grid), dim3(RW_THREADS), smem, st, tm, a, g);
  else e = dg_launch_pdl(conv_row_kernel<0, false>, dim3(grid), dim3(RW_THREADS), smem, st, tm, a, g);
  DG_CHECK_CUDA(e);
  DG_LAUNCH_CHECK();
  static const bool dbg_sync = getenv("DEPGAN_DEBUG_SYNC") != nullptr;
  if (dbg_sync) {
    cudaError_t e2 = cudaStreamSynchronize(st);
    if (e2 != cudaSuccess) {
      char buf[384];
      snprintf(buf, sizeof buf, "conv_row_kernel failed (%s): N=%d H=%d W=%d C0=%d C1=%d film=%d add=%d mask=%d out=%d "
               "head=%d | grid=%d na=%d side=%d smem=%u", cudaGetErrorString(e2), a.N, a.H, a.W, a.C0, a.C1,
               a.film_g != nullptr, a.add_src != nullptr, a.mask_src != nullptr, a.out != nullptr, a.head_w != nullptr,
               grid, g.na, g.n_side, smem);
      depgan_set_error(buf);
      return -1;
    }
  }
  return 0;
}

#ifdef DG_ROW_TRACE
extern "C" int depgan_dbg_set_row_trace(long long* p) {
  return (int)cudaMemcpyToSymbol(g_row_trace, &p, sizeof(p));
}
#endif
