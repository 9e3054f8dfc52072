// Row-streaming tcgen05 convolution for the 32-output-channel 3x3 layers at full resolution (W a multiple of 128):
// conv2d_gen_1, conv2d_gen_noise_m1 / _p1 (FiLM), conv2d_gen_16 (96 -> 32) and conv2d_gen_17 (+ the 1x1
// gen_segmentation head) of Gen_UNet2D (TG:398-409, 482-495; TU:332-343, 411-424), and the same shapes of the
// data-gradient passes.  Same arithmetic contract as conv_tc_kernel (ConvArgs, common.cuh); only the mapping differs.
//
// Why a second mapping.  An SS-mode tcgen05.mma re-reads its A slice (128 pixels x 16 channels = 4 KB = 32 shared-
// memory wavefronts) for every instruction.  With N = Cout = 32 the tensor pipe needs 16 clk per instruction, so
// conv_tc_kernel's tap-by-tap stream is bound by the operand reads (40 wavefronts per MMA; ncu:
// l1tex__data_pipe_tc_wavefronts_mem_shared at 83-93 % of peak, tensor pipe 36-39 %).  Here the three vertical taps
// are stacked on N: one MMA of N = 96 multiplies an image-row slice by the weights of (dy = 0,1,2; dx fixed), so the
// A slice is read once per three taps (56 wavefronts per 3 taps instead of 120).
//
// GEMM view.  M = 128 consecutive pixels of ONE image row j (a work item is a band of R output rows x 128 columns),
// N = 96 = (dy, cout), K = (dx, cin).  For every input row j of the band (R + 2 rows, zero rows outside the image come
// from TMA out-of-bounds fill)
//     D_j[x][(dy, co)] = sum_dx sum_ci  X[j][x + dx - 1][ci] * W[dy][dx][ci][co]
// is accumulated in its own TMEM slot (96 columns, ring of 5 slots), the dx shift being a different start address of
// the [130 pixels][32 channels] row tile in shared memory (64-byte swizzle, as in conv_tc_kernel).  Output row r is
//     out[r][x][co] = D_{r-1}[x][(0, co)] + D_r[x][(1, co)] + D_{r+1}[x][(2, co)],
// three accumulator columns of the SAME TMEM lane: the epilogue thread of pixel x adds them in registers, no
// shuffles, no halo pixels, and every input row is read from L2 / HBM exactly once per band (+2 halo rows per R).
//
// Roles (12 warps): warps 0-7 epilogue -- eight independent agents: warp w owns TMEM lane quarter w % 4 (32 pixels) of
// the output rows with parity (w / 4), stages its 32 pixels x 64 bytes in its own swizzled shared-memory slots and
// issues its own TMA store, so the row loop has no block- or group-wide barrier; warp 8 TMA producer of the row tiles,
// warp 9 MMA issuer (one elected lane) + TMEM owner, warp 10 TMA producer of the epilogue's side rows (FiLM residual,
// add / mask sources).  Weights ([dx][chunk][96][32] bf16, <= 54 KB) are resident for the CTA's lifetime.
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
constexpr int NACC = 5;                 // TMEM accumulator slots (96 columns each)
constexpr int NSLOT = 3;                // output staging slots per epilogue warp (2 KB each)
constexpr uint32_t WSLOT = 32 * 64;     // one warp's staging slot: 32 pixels x 32 channels bf16
constexpr int NSIDE = 4;                // side-row stages
constexpr uint32_t SIDE_ROW = BW * 64;  // one side row: 128 pixels x 32 channels bf16

struct RowGeom {
  int nchunk0, nchunk1;  // 32-channel chunks from in0 / in1
  int R;                 // output rows per band
  int bands_h, blocks_w; // bands per image column block, 128-pixel blocks per row
  int na;                // row-tile ring depth
  int n_side;            // side tensors per output row (0..2)
  int stage_out;         // bf16 `out` is written
};

struct RowMaps {
  CUtensorMap a0, a1, b, out, s0, s1;
};

__device__ __forceinline__ void mbar_arrive_n(uint32_t bar, uint32_t n) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n) : "memory");
}
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

// (n, column block, band) of the CTA's k-th work item
struct BandIter {
  int it, step, n_items, bands_h, blocks_w;
  int n, xb, band;
  __device__ __forceinline__ void init(int it0, int step_, int n_items_, int bands_h_, int blocks_w_) {
    it = it0; step = step_; n_items = n_items_; bands_h = bands_h_; blocks_w = blocks_w_;
    decode();
  }
  __device__ __forceinline__ void decode() {
    int t = it;
    band = t % bands_h; t /= bands_h;
    xb = t % blocks_w;
    n = t / blocks_w;
  }
  __device__ __forceinline__ bool valid() const { return it < n_items; }
  __device__ __forceinline__ void next() { it += step; decode(); }
};

// ---------------------------------------------------------------------------------------------------------
// EPI bit 0: FiLM residual side row; bit 1: add / mask side rows.  HEAD: fused 1x1 head on the 32 outputs.
// ---------------------------------------------------------------------------------------------------------
template <int EPI, bool HEAD>
__global__ void __launch_bounds__(RW_THREADS, 1) conv_row_kernel(const __grid_constant__ RowMaps tm, const ConvArgs a,
                                                                 const RowGeom g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const int nchunks = g.nchunk0 + g.nchunk1;
  const uint32_t b_base = base;                                      // weights: [dx][chunk] tiles of 6 KB
  const uint32_t a_base = b_base + 3u * nchunks * B_TILE;            // row-tile ring
  const uint32_t o_base = a_base + (uint32_t)g.na * A_STAGE;         // 8 warps x NSLOT x 2 KB
  const uint32_t s_base = o_base + 8u * NSLOT * WSLOT;               // NSIDE stages x n_side rows
  const uint32_t bar_base = s_base + (uint32_t)NSIDE * g.n_side * SIDE_ROW;
  const uint32_t fullA = bar_base, emptyA = fullA + 8 * g.na;
  const uint32_t fullB = emptyA + 8 * g.na;
  const uint32_t accFull = fullB + 8, accEmpty = accFull + 8 * NACC;
  const uint32_t sideFull = accEmpty + 8 * NACC, sideEmpty = sideFull + 8 * NSIDE;
  const uint32_t tmem_slot = sideEmpty + 8 * NSIDE;
  const uint32_t f_off = (tmem_slot + 16 + 15u) & ~15u;  // floats (16-byte aligned): scale[32], shift[32], head[32] float4, per-warp tables
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  float* s_scale = reinterpret_cast<float*>(smem_raw + (f_off - raw));
  float* s_shift = s_scale + 32;
  float4* s_head = reinterpret_cast<float4*>(s_shift + 32);
  float* s_tab = reinterpret_cast<float*>(s_head + 32);  // [8 warps][2 buffers][2][32]

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int R = g.R;
  const int n_items = a.N * g.blocks_w * g.bands_h;

  if (threadIdx.x == 0) {
    for (int i = 0; i < g.na; ++i) { mbar_init(fullA + 8 * i, 1); mbar_init(emptyA + 8 * i, 1); }
    mbar_init(fullB, 1);
    for (int i = 0; i < NACC; ++i) { mbar_init(accFull + 8 * i, 1); mbar_init(accEmpty + 8 * i, 12); }
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

  if (warp >= CTRL_W0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    if (warp == CTRL_W0) {
      // ===== TMA producer: resident weights once, then the input rows of every band =====
      if (elect_one()) {
        mbar_expect_tx(fullB, 9u * nchunks * 32u * 64u);
        for (int dx = 0; dx < 3; ++dx)
          for (int c = 0; c < nchunks; ++c) {
            const int kglob = c < g.nchunk0 ? c * 32 : a.C0 + (c - g.nchunk0) * 32;
            for (int dy = 0; dy < 3; ++dy)
              tma_load_2d(b_base + (uint32_t)(dx * nchunks + c) * B_TILE + (uint32_t)dy * 32u * 64u, &tm.b, fullB, kglob,
                          (dy * 3 + dx) * 32);
          }
      }
      __syncwarp();
      Ring ra;
      BandIter bi;
      bi.init(blockIdx.x, (int)gridDim.x, n_items, g.bands_h, g.blocks_w);
      for (; bi.valid(); bi.next()) {
        const int x0 = bi.xb * BW - 1, r0 = bi.band * R - 1;
        for (int j = 0; j < R + 2; ++j) {
          for (int c = 0; c < nchunks; ++c) {
            mbar_wait(emptyA + 8 * ra.idx, ra.phase ^ 1u);
            if (elect_one()) {
              const bool first = c < g.nchunk0;
              mbar_expect_tx(fullA + 8 * ra.idx, A_TX);
              tma_load_4d(a_base + ra.idx * A_STAGE, first ? &tm.a0 : &tm.a1, fullA + 8 * ra.idx,
                          (first ? c : c - g.nchunk0) * 32, x0, r0 + j, bi.n);
            }
            __syncwarp();
            ra.advance(g.na);
          }
        }
      }
    } else if (warp == CTRL_W0 + 1) {
      // ===== MMA issuer: per input row 6 * nchunks instructions of N = 96 into the row's accumulator slot =====
      const uint32_t idesc = make_idesc(96);
      const uint32_t hi = ((uint32_t)(8 * 64) >> 4) | (1u << 14) | (4u << 29);  // SBO 512 B, 64-byte swizzle
      constexpr uint32_t LBO1 = 1u << 16;
      Ring ra;
      mbar_wait(fullB, 0);
      tc_fence_after();
      const uint32_t b_lo0 = ((b_base & 0x3FFFFu) >> 4) | LBO1;
      uint32_t slot = 0, sphase = 0;  // accumulator slot of the current input row and its use parity
      BandIter bi;
      bi.init(blockIdx.x, (int)gridDim.x, n_items, g.bands_h, g.blocks_w);
      for (; bi.valid(); bi.next()) {
        for (int j = 0; j < R + 2; ++j) {
          mbar_wait(accEmpty + 8 * slot, sphase ^ 1u);
          tc_fence_after();
          const uint32_t d = tmem_base + slot * 96u;
          for (int c = 0; c < nchunks; ++c) {
            mbar_wait(fullA + 8 * ra.idx, ra.phase);
            tc_fence_after();
            const uint32_t a_lo = (((a_base + ra.idx * A_STAGE) & 0x3FFFFu) >> 4) | LBO1;
            const uint32_t b_lo = b_lo0 + (uint32_t)c * (B_TILE >> 4);
            if (elect_one()) {
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
                const uint32_t bt = b_lo + (uint32_t)dx * (uint32_t)nchunks * (B_TILE >> 4);
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                  const uint32_t acc = (c | dx | k) != 0 ? 1u : 0u;
                  tc_mma(d, ((uint64_t)hi << 32) | (a_lo + 4 * dx + 2 * k), ((uint64_t)hi << 32) | (bt + 2 * k), idesc,
                         acc);
                }
              }
              tc_commit(emptyA + 8 * ra.idx);
              if (c == nchunks - 1) tc_commit(accFull + 8 * slot);
            }
            __syncwarp();
            ra.advance(g.na);
          }
          if (++slot == NACC) { slot = 0; sphase ^= 1u; }
        }
      }
    } else if (warp == CTRL_W0 + 2) {
      // ===== TMA producer of the side rows (FiLM residual, or add / mask sources): one 128-pixel row per output row =====
      if (EPI != 0 && g.n_side > 0) {
        Ring rs;
        BandIter bi;
        bi.init(blockIdx.x, (int)gridDim.x, n_items, g.bands_h, g.blocks_w);
        for (; bi.valid(); bi.next()) {
          for (int i = 0; i < R; ++i) {
            mbar_wait(sideEmpty + 8 * rs.idx, rs.phase ^ 1u);
            if (elect_one()) {
              const uint32_t dst = s_base + (uint32_t)(rs.idx * g.n_side) * SIDE_ROW;
              mbar_expect_tx(sideFull + 8 * rs.idx, (uint32_t)g.n_side * SIDE_ROW);
              tma_load_4d(dst, &tm.s0, sideFull + 8 * rs.idx, 0, bi.xb * BW, bi.band * R + i, bi.n);
              if (g.n_side == 2) tma_load_4d(dst + SIDE_ROW, &tm.s1, sideFull + 8 * rs.idx, 0, bi.xb * BW, bi.band * R + i, bi.n);
            }
            __syncwarp();
            rs.advance(NSIDE);
          }
        }
      }
    }
  } else {
    // ===== epilogue warps: eight independent agents =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    constexpr bool E_RES = (EPI & 1) != 0, E_AM = (EPI & 2) != 0;
    const int wg = warp >> 2;        // parity of the output rows this warp handles
    const int qd = warp & 3;         // TMEM lane quarter
    const int m = qd * 32 + lane;    // pixel of the 128-pixel block
    const bool has_add = E_AM && a.add_src != nullptr;
    const bool has_mask = E_AM && a.mask_src != nullptr;
    const bool stage_out = g.stage_out != 0;
    const int head_nc = HEAD ? a.head_nc : 0;
    float* tab = s_tab + warp * 128;  // [2 buffers][scale 32 | shift 32]
    // staging: this warp's NSLOT slots of [32 pixels][64 B], 16-byte units XOR-swizzled like the TMA (64-byte mode)
    const uint32_t w_o = o_base + (uint32_t)warp * NSLOT * WSLOT;
    const uint32_t p_off = (uint32_t)lane * 64u;
    const uint32_t p_xor = (uint32_t)(lane >> 1) & 3u;
    // side rows are [128 pixels][64 B] with the same swizzle
    const uint32_t sp_off = (uint32_t)m * 64u;
    const uint32_t sp_xor = (uint32_t)(m >> 1) & 3u;
    const uint32_t mask_off = has_add ? SIDE_ROW : 0u;
    const uint32_t lane_sel = ((uint32_t)(qd * 32) << 16);

    uint32_t q0 = 0;        // input-row sequence number of the band's first input row
    uint32_t o0 = 0;        // output-row sequence number of the band's first output row
    uint32_t nrow = 0;      // rows this warp has staged (staging slot ring)
    int bseq = 0;
    BandIter bi;
    bi.init(blockIdx.x, (int)gridDim.x, n_items, g.bands_h, g.blocks_w);
    for (; bi.valid(); bi.next(), q0 += (uint32_t)(R + 2), o0 += (uint32_t)R, ++bseq) {
      // per-band table: v = acc * sc + sh   (BN folded; FiLM: relu(v * g + b) folded into the same affine)
      float* tb = tab + (bseq & 1) * 64;
      {
        float sc = s_scale[lane], sh = s_shift[lane];
        if (E_RES) {
          const float fg = __ldg(a.film_g + (size_t)bi.n * a.film_stride + lane);
          const float fb = __ldg(a.film_b + (size_t)bi.n * a.film_stride + lane);
          sh = fmaf(sh, fg, fb);
          sc *= fg;
        }
        tb[lane] = sc;
        tb[32 + lane] = sh;
      }
      __syncwarp();
      const int xpix = bi.xb * BW + m;
      for (int i = (int)((o0 ^ (uint32_t)wg) & 1u); i < R; i += 2) {
        const uint32_t o = o0 + (uint32_t)i;
        const uint32_t qi = q0 + (uint32_t)i;
        // accumulator slots of input rows r-1, r, r+1 (sequence numbers qi, qi+1, qi+2)
        const uint32_t sa = qi % NACC, sb = (qi + 1) % NACC, sc_ = (qi + 2) % NACC;
        mbar_wait(accFull + 8 * sc_, ((qi + 2) / NACC) & 1u);
        tc_fence_after();
        uint32_t va[32], vb[32], vc[32];
        tc_ld32_issue(tmem_base + lane_sel + sa * 96u, va);
        tc_ld32_issue(tmem_base + lane_sel + sb * 96u + 32u, vb);
        tc_ld32_issue(tmem_base + lane_sel + sc_ * 96u + 64u, vc);
        tc_ld_wait_all();
        tc_ld32_fence(va); tc_ld32_fence(vb); tc_ld32_fence(vc);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          // every accumulator use collects 12 arrivals (3 reading rows x 4 warps); the first / last rows of a band
          // arrive for the readers that do not exist
          const uint32_t top = i == 0 ? 1u : 0u, bot = i == R - 1 ? 1u : 0u;
          mbar_arrive_n(accEmpty + 8 * sa, 1u + 2u * top);
          mbar_arrive_n(accEmpty + 8 * sb, 1u + top + bot);
          mbar_arrive_n(accEmpty + 8 * sc_, 1u + 2u * bot);
        }
        float v[32];
#pragma unroll
        for (int k = 0; k < 32; ++k)
          v[k] = (__uint_as_float(va[k]) + __uint_as_float(vb[k])) + __uint_as_float(vc[k]);
        // ---- affine (+ FiLM) ----
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          const float4 sc4 = reinterpret_cast<const float4*>(tb)[k4];
          const float4 sh4 = reinterpret_cast<const float4*>(tb + 32)[k4];
          v[4 * k4 + 0] = fmaf(v[4 * k4 + 0], sc4.x, sh4.x);
          v[4 * k4 + 1] = fmaf(v[4 * k4 + 1], sc4.y, sh4.y);
          v[4 * k4 + 2] = fmaf(v[4 * k4 + 2], sc4.z, sh4.z);
          v[4 * k4 + 3] = fmaf(v[4 * k4 + 3], sc4.w, sh4.w);
        }
        if (EPI != 0 && g.n_side > 0) {
          const uint32_t ss = o % NSIDE;
          mbar_wait(sideFull + 8 * ss, (o / NSIDE) & 1u);
          const uint8_t* sgen = smem_raw + (s_base + (uint32_t)(ss * g.n_side) * SIDE_ROW - raw) + sp_off;
          if (E_RES) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const uint4 q = *reinterpret_cast<const uint4*>(sgen + (((uint32_t)u ^ sp_xor) << 4));
              const uint32_t wv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                v[8 * u + 2 * e] = fmaxf(v[8 * u + 2 * e], 0.f) + __uint_as_float(wv[e] << 16);
                v[8 * u + 2 * e + 1] = fmaxf(v[8 * u + 2 * e + 1], 0.f) + __uint_as_float(wv[e] & 0xFFFF0000u);
              }
            }
          }
          if (has_add) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const uint4 q = *reinterpret_cast<const uint4*>(sgen + (((uint32_t)u ^ sp_xor) << 4));
              const uint32_t wv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                v[8 * u + 2 * e] += __uint_as_float(wv[e] << 16);
                v[8 * u + 2 * e + 1] += __uint_as_float(wv[e] & 0xFFFF0000u);
              }
            }
          }
          if (has_mask) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const uint4 q = *reinterpret_cast<const uint4*>(sgen + mask_off + (((uint32_t)u ^ sp_xor) << 4));
              const uint32_t wv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                v[8 * u + 2 * e] = __uint_as_float(wv[e] << 16) > 0.f ? v[8 * u + 2 * e] : 0.f;
                v[8 * u + 2 * e + 1] = __uint_as_float(wv[e] & 0xFFFF0000u) > 0.f ? v[8 * u + 2 * e + 1] : 0.f;
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(sideEmpty + 8 * ss);
        }
        if (a.relu) {
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = fmaxf(v[k], 0.f);
        }
        const int row = bi.band * R + i;
        if (stage_out) {
          const uint32_t slot = nrow % NSLOT;
          // the TMA store that last read this slot was issued NSLOT rows ago: at most NSLOT - 1 newer groups may be
          // pending
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NSLOT - 1) : "memory");
          __syncwarp();
          uint8_t* ogen = smem_raw + (w_o + slot * WSLOT - raw) + p_off;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4 pk;
            __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
            for (int e = 0; e < 4; ++e) hp[e] = __floats2bfloat162_rn(v[8 * u + 2 * e], v[8 * u + 2 * e + 1]);
            *reinterpret_cast<uint4*>(ogen + (((uint32_t)u ^ p_xor) << 4)) = pk;
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&tm.out, w_o + slot * WSLOT, 0, bi.xb * BW + qd * 32, row, bi.n);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          ++nrow;
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
          float o0_ = h0 + __ldg(a.head_b), o1 = 0.f, o2 = 0.f, o3 = 0.f;
          if (nc > 1) o1 = h1 + __ldg(a.head_b + 1);
          if (nc > 2) o2 = h2 + __ldg(a.head_b + 2);
          if (nc > 3) o3 = h3 + __ldg(a.head_b + 3);
          if (a.head_act == 0) {
            o0_ = tanhf(o0_);
            if (nc > 1) { o1 = tanhf(o1); o2 = tanhf(o2); o3 = tanhf(o3); }
          } else if (a.head_act == 1) {
            float mx = o0_;
            if (nc > 1) mx = fmaxf(mx, o1);
            if (nc > 2) mx = fmaxf(mx, o2);
            if (nc > 3) mx = fmaxf(mx, o3);
            o0_ = expf(o0_ - mx);
            o1 = nc > 1 ? expf(o1 - mx) : 0.f;
            o2 = nc > 2 ? expf(o2 - mx) : 0.f;
            o3 = nc > 3 ? expf(o3 - mx) : 0.f;
            const float inv = 1.0f / (o0_ + o1 + o2 + o3);
            o0_ *= inv; o1 *= inv; o2 *= inv; o3 *= inv;
          }
          const size_t pix = ((size_t)bi.n * a.H + row) * a.W + xpix;
          if (nc == 4) {
            *reinterpret_cast<float4*>(a.head_out + pix * 4) = make_float4(o0_, o1, o2, o3);
          } else {
            float* op = a.head_out + pix * nc;
            op[0] = o0_;
            if (nc > 1) op[1] = o1;
            if (nc > 2) op[2] = o2;
          }
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

int make_row_map(CUtensorMap* tm, const void* p, int C, int W, int H, int N, int box_w, const char* what) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {32, (cuuint32_t)box_w, 1, 1};
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
  g->R = a.H % 32 == 0 ? 32 : 16;
  g->bands_h = a.H / g->R; g->blocks_w = a.W / BW;
  g->n_side = a.film_g ? 1 : (a.add_src ? 1 : 0) + (a.mask_src ? 1 : 0);
  g->stage_out = a.out ? 1 : 0;
  const uint32_t fixed = 1024 + 3u * nchunks * B_TILE + 8u * NSLOT * WSLOT + (uint32_t)NSIDE * g->n_side * SIDE_ROW +
                         8u * (2 * 16 + 1 + 2 * NACC + 2 * NSIDE) + 16 + (64 + 128 + 8 * 128) * 4 + 64;
  if (fixed + 4 * A_STAGE > SMEM_BUDGET_R) return false;
  int na = (int)((SMEM_BUDGET_R - fixed) / A_STAGE);
  if (na > 16) na = 16;
  g->na = na;
  *smem = fixed + (uint32_t)na * A_STAGE;
  return true;
}

template <int EPI, bool HEAD>
int set_attr_row() {
  DG_CHECK_CUDA(cudaFuncSetAttribute(conv_row_kernel<EPI, HEAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
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
    dg_device_mark(g_dev_r, dev);
  }
  if (dev < 64) g_sms_r = g_dev_r.sms[dev];
  else DG_CHECK_CUDA(cudaDeviceGetAttribute(&g_sms_r, cudaDevAttrMultiProcessorCount, dev));
  return 0;
}

}  // namespace

// The row-streaming kernel takes the 3x3, 32-output-channel layers whose width is a multiple of 128.
bool conv_row_supported(const ConvArgs& a) {
  static const bool off = getenv("DEPGAN_NO_ROW") != nullptr;  // A/B switch: every layer through conv_tc_kernel
  if (off) return false;
  if (a.in_dt != DT_BF16 || a.out_dt != DT_BF16) return false;
  if (a.ks != 3 || a.deconv || a.Cout != 32) return false;
  if (a.W % BW || a.H % 16 || a.H < 16) return false;
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
  DG_TRY(make_row_map(&tm.a0, a.in0, a.C0, a.W, a.H, a.N, TILE_PX, "input 0"));
  if (a.C1 > 0) DG_TRY(make_row_map(&tm.a1, a.in1, a.C1, a.W, a.H, a.N, TILE_PX, "input 1"));
  else tm.a1 = tm.a0;
  DG_TRY(make_row_w_map(&tm.b, a.w_tc, a.C0 + a.C1, 9 * 32));
  tm.out = tm.s0 = tm.s1 = tm.a0;
  if (a.out) DG_TRY(make_row_map(&tm.out, a.out, 32, a.W, a.H, a.N, 32, "output"));
  const void* side[2] = {nullptr, nullptr};
  if (a.film_g) side[0] = a.res;
  else {
    int k = 0;
    if (a.add_src) side[k++] = a.add_src;
    if (a.mask_src) side[k++] = a.mask_src;
  }
  if (side[0]) DG_TRY(make_row_map(&tm.s0, side[0], 32, a.W, a.H, a.N, BW, "side input"));
  if (side[1]) DG_TRY(make_row_map(&tm.s1, side[1], 32, a.W, a.H, a.N, BW, "side input"));
  const int n_items = a.N * g.blocks_w * g.bands_h;
  const int grid = n_items < g_sms_r ? n_items : g_sms_r;
  const int epi = a.film_g ? 1 : ((a.add_src || a.mask_src) ? 2 : 0);
  cudaError_t e;
  if (a.head_w) e = dg_launch_pdl(conv_row_kernel<0, true>, dim3(grid), dim3(RW_THREADS), smem, st, tm, a, g);
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
               "head=%d | R=%d na=%d side=%d smem=%u", cudaGetErrorString(e2), a.N, a.H, a.W, a.C0, a.C1,
               a.film_g != nullptr, a.add_src != nullptr, a.mask_src != nullptr, a.out != nullptr, a.head_w != nullptr,
               g.R, g.na, g.n_side, smem);
      depgan_set_error(buf);
      return -1;
    }
  }
  return 0;
}
