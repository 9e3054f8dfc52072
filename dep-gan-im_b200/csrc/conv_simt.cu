// fp32 CUDA-core implicit-GEMM convolution ('same', stride 1, k in {1,3,5}) with the fused epilogues of
// common.cuh, plus the matching weight-gradient kernel.  This is the FP32-accumulate / FP32-storage variant
// (BASELINE.json: "FP32-accumulate variant within 1e-4") and the layer of last resort for shapes the
// tcgen05 path does not take (Cin < 16 first layers, Cout = 1 dgrad of the critic's first layer).
// Replaces Keras Conv2D (+BatchNormalization, +Activation) of TG:285-304 and its K.gradients (TG:543-549).
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace {

constexpr int TW = 16, TH = 16;  // output tile (pixels)
constexpr int CK = 8;            // input channels per smem stage
constexpr int COB = 32;          // output channels per CTA
constexpr int NT = 128;          // threads: thread t -> pixel column t%16, rows t/16 and t/16+8

template <typename T>
__device__ __forceinline__ float ld_any(const void* p, size_t i) {
  return ldf(reinterpret_cast<const T*>(p) + i);
}

template <typename TI, typename TO, int KS>
__global__ void __launch_bounds__(NT) conv_fwd_simt_kernel(ConvArgs a) {
  constexpr int PAD = KS / 2;
  constexpr int IH = TH + KS - 1, IW = TW + KS - 1;
  constexpr int TAPS = KS * KS;
  __shared__ float s_in[CK][IH][IW + 1];
  __shared__ __align__(16) float s_w[TAPS][CK][COB];

  const int tid = threadIdx.x;
  const int tiles_w = (a.W + TW - 1) / TW;
  const int w0 = (blockIdx.x % tiles_w) * TW, h0 = (blockIdx.x / tiles_w) * TH;
  const int cob = blockIdx.y * COB;
  const int n = blockIdx.z;
  const int Cin = a.C0 + a.C1;
  const int px = tid & 15, py = tid >> 4;

  float acc0[COB], acc1[COB];
#pragma unroll
  for (int i = 0; i < COB; ++i) acc0[i] = acc1[i] = 0.f;

  for (int cb = 0; cb < Cin; cb += CK) {
    // stage input halo tile: channel fastest in the global read
    for (int idx = tid; idx < CK * IH * IW; idx += NT) {
      int ci = idx % CK;
      int c = (idx / CK) % IW;
      int r = idx / (CK * IW);
      int h = h0 + r - PAD, w = w0 + c - PAD, ch = cb + ci;
      float v = 0.f;
      if (h >= 0 && h < a.H && w >= 0 && w < a.W && ch < Cin) {
        size_t pix = ((size_t)n * a.H + h) * a.W + w;
        if (ch < a.C0) v = ld_any<TI>(a.in0, pix * a.C0 + ch);
        else v = ld_any<TI>(a.in1, pix * a.C1 + (ch - a.C0));
      }
      s_in[ci][r][c] = v;
    }
    for (int idx = tid; idx < TAPS * CK * COB; idx += NT) {
      int co = idx % COB;
      int ci = (idx / COB) % CK;
      int tap = idx / (COB * CK);
      float v = 0.f;
      if (cb + ci < Cin && cob + co < a.Cout) v = a.w[((size_t)tap * Cin + cb + ci) * a.Cout + cob + co];
      s_w[tap][ci][co] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int tap = 0; tap < TAPS; ++tap) {
      const int ty = tap / KS, tx = tap % KS;
#pragma unroll
      for (int ci = 0; ci < CK; ++ci) {
        const float v0 = s_in[ci][py + ty][px + tx];
        const float v1 = s_in[ci][py + 8 + ty][px + tx];
        const float4* wp = reinterpret_cast<const float4*>(&s_w[tap][ci][0]);
#pragma unroll
        for (int q = 0; q < COB / 4; ++q) {
          const float4 w4 = wp[q];
          acc0[4 * q + 0] = fmaf(v0, w4.x, acc0[4 * q + 0]);
          acc0[4 * q + 1] = fmaf(v0, w4.y, acc0[4 * q + 1]);
          acc0[4 * q + 2] = fmaf(v0, w4.z, acc0[4 * q + 2]);
          acc0[4 * q + 3] = fmaf(v0, w4.w, acc0[4 * q + 3]);
          acc1[4 * q + 0] = fmaf(v1, w4.x, acc1[4 * q + 0]);
          acc1[4 * q + 1] = fmaf(v1, w4.y, acc1[4 * q + 1]);
          acc1[4 * q + 2] = fmaf(v1, w4.z, acc1[4 * q + 2]);
          acc1[4 * q + 3] = fmaf(v1, w4.w, acc1[4 * q + 3]);
        }
      }
    }
    __syncthreads();
  }

  // epilogue
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int h = h0 + py + 8 * half, w = w0 + px;
    if (h >= a.H || w >= a.W) continue;
    const size_t pix = ((size_t)n * a.H + h) * a.W + w;
    TO* out = reinterpret_cast<TO*>(a.out) + pix * a.Cout;
#pragma unroll
    for (int i = 0; i < COB; ++i) {
      const int c = cob + i;
      if (c >= a.Cout) break;
      float v = half ? acc1[i] : acc0[i];
      if (a.scale) v *= a.scale[c];
      if (a.shift) v += a.shift[c];
      if (a.out_pre) stf(reinterpret_cast<TO*>(a.out_pre) + pix * a.Cout + c, v);
      if (a.film_g) {
        v = fmaf(v, a.film_g[(size_t)n * a.film_stride + c], a.film_b[(size_t)n * a.film_stride + c]);
        v = fmaxf(v, 0.f) + ld_any<TO>(a.res, pix * a.Cout + c);
      }
      if (a.add_src) v += ld_any<TO>(a.add_src, pix * a.Cout + c);
      if (a.mask_src) v = ld_any<TO>(a.mask_src, pix * a.Cout + c) > 0.f ? v : 0.f;
      if (a.relu) v = fmaxf(v, 0.f);
      if (a.out) stf(out + c, v);
    }
  }
}

template <typename TI, typename TO>
int launch_fwd(const ConvArgs& a, cudaStream_t st) {
  dim3 grid(((a.W + TW - 1) / TW) * ((a.H + TH - 1) / TH), (a.Cout + COB - 1) / COB, a.N);
  switch (a.ks) {
    case 1: conv_fwd_simt_kernel<TI, TO, 1><<<grid, NT, 0, st>>>(a); break;
    case 3: conv_fwd_simt_kernel<TI, TO, 3><<<grid, NT, 0, st>>>(a); break;
    case 5: conv_fwd_simt_kernel<TI, TO, 5><<<grid, NT, 0, st>>>(a); break;
    default: depgan_set_error("conv_fwd_simt: unsupported kernel size"); return -2;
  }
  DG_LAUNCH_CHECK();
  return 0;
}


// ------------------------------------------------------------------------------------------------------
// Edge layers (conv2d_gen_0: nicg -> 32 3x3, conv2d_dis_0a: 1 -> 16 5x5, and the 16 -> 1 data gradient of the
// latter).  K = ks*ks*Cin <= 25 (or Cout = 1) is far too small for tensor cores and the layers are HBM-bound on
// their wide side (2*Cout B per pixel), so they run on CUDA cores with everything that is reused held on chip:
//   * CTA = 8 warps side by side over a band of EDGE_ROWS rows; the narrow operand's halo tile is staged in shared
//     memory by the whole CTA (bulk coalesced loads, 2-3 CTAs per SM hide the fill).
//   * G = COUT/CPT lanes share one pixel, each owning CPT channels; a warp covers 32/G horizontally adjacent
//     pixels, so every wide-side access of a warp is ONE contiguous run of 32*CPT elements.
//   * the warp walks down its column strip; the ks x ks image window slides vertically in registers (ks new
//     shared-memory reads per pixel, rotation resolved at compile time by unrolling ks rows).
//   * forward: the TAPS*CIN*CPT weights of the lane live in registers for the whole strip; weight gradient: the
//     same number of accumulators does (the roles of weights and accumulators swap), reduced over the warp by
//     shuffles, over the CTA in shared memory and over CTAs by one atomic per element.
// ------------------------------------------------------------------------------------------------------
constexpr int EDGE_ROWS = 32;  // rows per CTA band
struct hsplit { char pad_[4]; };  // output tag: split-half storage (DT_F16S), a pixel = [COUT hi | COUT lo] IEEE halves

template <int N>
struct VecIO;  // N consecutive channels of one pixel
template <>
struct VecIO<2> {
  static __device__ __forceinline__ void load(const bf16* p, float (&v)[2]) {
    const uint32_t q = __ldg(reinterpret_cast<const uint32_t*>(p));
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q));
    v[0] = a.x; v[1] = a.y;
  }
  static __device__ __forceinline__ void load(const float* p, float (&v)[2]) {
    const float2 q = __ldg(reinterpret_cast<const float2*>(p));
    v[0] = q.x; v[1] = q.y;
  }
  static __device__ __forceinline__ void store(bf16* p, const float (&v)[2]) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(v[0], v[1]);
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[2]) {
    *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  }
  static __device__ __forceinline__ void load(const __half* p, float (&v)[2]) {
    const float2 a = unpack_h2(__ldg(reinterpret_cast<const uint32_t*>(p)), true);
    v[0] = a.x; v[1] = a.y;
  }
  static __device__ __forceinline__ void store(__half* p, const float (&v)[2]) {
    *reinterpret_cast<uint32_t*>(p) = pack_h2(v[0], v[1], true);
  }
};
template <>
struct VecIO<4> {
  static __device__ __forceinline__ void load(const bf16* p, float (&v)[4]) {
    const uint2 q = __ldg(reinterpret_cast<const uint2*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
    const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  }
  static __device__ __forceinline__ void store(bf16* p, const float (&v)[4]) {
    uint2 q;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
    h[0] = __floats2bfloat162_rn(v[0], v[1]);
    h[1] = __floats2bfloat162_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = q;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
  static __device__ __forceinline__ void load(const __half* p, float (&v)[4]) {
    const uint2 q = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 a = unpack_h2(q.x, true), b = unpack_h2(q.y, true);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
  static __device__ __forceinline__ void store(__half* p, const float (&v)[4]) {
    *reinterpret_cast<uint2*>(p) = make_uint2(pack_h2(v[0], v[1], true), pack_h2(v[2], v[3], true));
  }
};
template <>
struct VecIO<8> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    float lo[4], hi[4];
    VecIO<4>::load(p, lo);
    VecIO<4>::load(p + 4, hi);
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[i] = lo[i]; v[4 + i] = hi[i]; }
  }
  static __device__ __forceinline__ void load(const bf16* p, float (&v)[8]) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(bf16* p, const float (&v)[8]) {
    uint4 q;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = q;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    float4* q = reinterpret_cast<float4*>(p);
    q[0] = make_float4(v[0], v[1], v[2], v[3]);
    q[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
  static __device__ __forceinline__ void load(const __half* p, float (&v)[8]) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = unpack_h2(w[i], true);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(__half* p, const float (&v)[8]) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack_h2(v[0], v[1], true), pack_h2(v[2], v[3], true),
                                              pack_h2(v[4], v[5], true), pack_h2(v[6], v[7], true));
  }
};

// fp32 image halo tile of a band: rows y0-PAD .. y0+EDGE_ROWS+PAD-1, columns x0-PAD .. x0+TILE_W+PAD-1 ('same' zeros)
template <int KS, int CIN, int TILE_W>
struct EdgeTile {
  static constexpr int PAD = KS / 2, TR = EDGE_ROWS + KS - 1, TC = (TILE_W + KS - 1) * CIN;
  float v[TR][TC];
  __device__ __forceinline__ void fill(const float* __restrict__ x, int n, int y0, int x0, int H, int W) {
    for (int i = threadIdx.x; i < TR * TC; i += blockDim.x) {
      const int r = i / TC, cc = i - r * TC;
      const int gy = y0 - PAD + r, gx = x0 - PAD + cc / CIN, ci = cc % CIN;
      v[r][cc] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(x + (((size_t)n * H + gy) * W + gx) * CIN + ci) : 0.f;
    }
  }
  // the KS x CIN window row of output column `col` (tile-relative) at tile row r
  __device__ __forceinline__ void row(int r, int col, float (&out)[KS * CIN]) const {
#pragma unroll
    for (int q = 0; q < KS * CIN; ++q) out[q] = v[r][col * CIN + q];
  }
};

template <typename TO, int KS, int CIN, int COUT, int CPT>
__global__ void __launch_bounds__(256, 2) conv_first_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ scale,
                                                         const float* __restrict__ shift, TO* __restrict__ out,
                                                         const TO* __restrict__ mask, int H, int W, int relu) {
  constexpr int TAPS = KS * KS, G = COUT / CPT, PXW = 32 / G, KW = KS * CIN, TILE_W = 8 * PXW;
  __shared__ EdgeTile<KS, CIN, TILE_W> tile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, cg = lane % G, px = lane / G;
  const int x0 = blockIdx.x * TILE_W, y0 = blockIdx.y * EDGE_ROWS, n = blockIdx.z;
  tile.fill(x, n, y0, x0, H, W);
  const int col = warp * PXW + px, xc = x0 + col, c0 = cg * CPT;

  float wr[TAPS * CIN][CPT], sc[CPT], sh[CPT];
#pragma unroll
  for (int t = 0; t < TAPS * CIN; ++t)
#pragma unroll
    for (int c = 0; c < CPT; ++c) wr[t][c] = __ldg(w + (size_t)t * COUT + c0 + c);
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    sc[c] = scale ? __ldg(scale + c0 + c) : 1.f;
    sh[c] = shift ? __ldg(shift + c0 + c) : 0.f;
  }
  __syncthreads();
  if (xc >= W) return;
  // window slot of tile row t: t % KS; tile rows 0 .. KS-2 are preloaded
  float win[KS][KW];
#pragma unroll
  for (int d = 0; d < KS - 1; ++d) tile.row(d, col, win[d]);
  const int rows = min(EDGE_ROWS, H - y0);
  for (int r0 = 0; r0 < rows; r0 += KS) {
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      const int r = r0 + k;
      if (r < rows) {
        tile.row(r + KS - 1, col, win[(k + KS - 1) % KS]);
        float acc[CPT];
        if constexpr (CPT % 2 == 0) {
          // packed fp32x2 FMAs (one issue slot per two channels; per-lane IEEE fma, so the bits are those of fmaf)
          unsigned long long a2[CPT / 2];
#pragma unroll
          for (int c = 0; c < CPT / 2; ++c) a2[c] = 0ull;
#pragma unroll
          for (int dy = 0; dy < KS; ++dy)
#pragma unroll
            for (int q = 0; q < KW; ++q) {
              const float v = win[(k + dy) % KS][q];
              unsigned long long v2;
              asm("mov.b64 %0, {%1, %1};" : "=l"(v2) : "f"(v));
#pragma unroll
              for (int c = 0; c < CPT / 2; ++c) {
                unsigned long long w2;
                asm("mov.b64 %0, {%1, %2};" : "=l"(w2) : "f"(wr[dy * KW + q][2 * c]), "f"(wr[dy * KW + q][2 * c + 1]));
                asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a2[c]) : "l"(v2), "l"(w2));
              }
            }
#pragma unroll
          for (int c = 0; c < CPT / 2; ++c)
            asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[2 * c]), "=f"(acc[2 * c + 1]) : "l"(a2[c]));
        } else {
#pragma unroll
          for (int c = 0; c < CPT; ++c) acc[c] = 0.f;
#pragma unroll
          for (int dy = 0; dy < KS; ++dy)
#pragma unroll
            for (int q = 0; q < KW; ++q) {
              const float v = win[(k + dy) % KS][q];
#pragma unroll
              for (int c = 0; c < CPT; ++c) acc[c] = fmaf(v, wr[dy * KW + q][c], acc[c]);
            }
        }
        const size_t o = (((size_t)n * H + y0 + r) * W + xc) * COUT + c0;
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          acc[c] = fmaf(acc[c], sc[c], sh[c]);
          if (relu) acc[c] = fmaxf(acc[c], 0.f);
        }
        if constexpr (std::is_same<TO, hsplit>::value) {
          __half* op = reinterpret_cast<__half*>(out) + 2 * (o - c0) + c0;  // pixel pitch 2*COUT
          float lo[CPT];
#pragma unroll
          for (int c = 0; c < CPT; ++c) {
            const float hi = __half2float(__float2half_rn(fminf(fmaxf(acc[c], -65504.f), 65504.f)));
            lo[c] = acc[c] - hi;
          }
          VecIO<CPT>::store(op, acc);
          VecIO<CPT>::store(op + COUT, lo);
        } else {
        if (mask) {  // activation-pattern mask of the JVP pass (TG:543: the critic linearised at the mixed sample)
          float m[CPT];
          VecIO<CPT>::load(mask + o, m);
#pragma unroll
          for (int c = 0; c < CPT; ++c) acc[c] = m[c] > 0.f ? acc[c] : 0.f;
        }
        VecIO<CPT>::store(out + o, acc);
        }
      }
    }
  }
}

template <typename TO, int KS, int CIN, int COUT, int CPT>
int launch_first(const ConvArgs& a, cudaStream_t st) {
  constexpr int TILE_W = 8 * (32 / (COUT / CPT));
  dim3 grid((a.W + TILE_W - 1) / TILE_W, (a.H + EDGE_ROWS - 1) / EDGE_ROWS, a.N);
  conv_first_kernel<TO, KS, CIN, COUT, CPT><<<grid, 256, 0, st>>>((const float*)a.in0, a.w, a.scale, a.shift,
                                                                   (TO*)a.out, (const TO*)a.mask_src, a.H, a.W, a.relu);
  DG_LAUNCH_CHECK();
  return 0;
}

// returns 1 when the call was taken by a first-layer specialisation, 0 otherwise, <0 on error
template <typename TO>
int try_first(const ConvArgs& a, cudaStream_t st) {
  if (a.in_dt != DT_F32 || a.C1 != 0 || a.out_pre || a.film_g || a.add_src || !a.out) return 0;
  if (a.W % 8) return 0;
  int r = 1;
  if (a.ks == 3 && a.C0 == 1 && a.Cout == 32) r = launch_first<TO, 3, 1, 32, 8>(a, st);
  else if (a.ks == 3 && a.C0 == 2 && a.Cout == 32) r = launch_first<TO, 3, 2, 32, 4>(a, st);
  else if (a.ks == 5 && a.C0 == 1 && a.Cout == 16) r = launch_first<TO, 5, 1, 16, 2>(a, st);
  else return 0;
  return r < 0 ? r : 1;
}

// ------------------------------------------------------------------------------------------------------
// Single-output-channel convolution (the data gradient of conv2d_dis_0a: 16 -> 1, 5x5): dD/dx for the gradient
// penalty and for the generator's adversarial terms.  CTA = 32 columns x LAST_ROWS rows; the input halo tile is
// staged in shared memory as channel pairs; G = CIN/2 lanes share a pixel (2 input channels each, their 25*2 weights in
// registers).  The thread walks down its column one INPUT row at a time: the row's 5 x 2 values feed the 5 output
// rows it touches (accumulator ring, rotation unrolled), so each output costs 5 shared-memory reads instead of 25;
// a finished output row is reduced over the G lanes with shuffles.  fp32 output.
// ------------------------------------------------------------------------------------------------------
template <typename T> struct LastPair;  // two adjacent input channels as staged in shared memory
template <> struct LastPair<bf16> {
  typedef uint32_t type;
  static constexpr int ROWS = 32;
  static __device__ __forceinline__ uint32_t load(const bf16* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }
  static __device__ __forceinline__ float2 get(uint32_t q) {
    return make_float2(__uint_as_float(q << 16), __uint_as_float(q & 0xFFFF0000u));
  }
};
template <> struct LastPair<float> {
  typedef float2 type;
  static constexpr int ROWS = 12;
  static __device__ __forceinline__ float2 load(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
  static __device__ __forceinline__ float2 get(float2 q) { return q; }
};

template <typename TI, int KS, int CIN>
__global__ void __launch_bounds__(256) conv_last_kernel(const TI* __restrict__ in, const float* __restrict__ w,
                                                        float* __restrict__ out, int H, int W) {
  typedef LastPair<TI> LP;
  constexpr int LAST_ROWS = LP::ROWS;
  constexpr int PAD = KS / 2, TAPS = KS * KS, CPT = 2, G = CIN / CPT, PXW = 32 / G, TILE_W = 8 * PXW;
  constexpr int TR = LAST_ROWS + KS - 1, TCOLS = TILE_W + KS - 1;
  // [row][col][channel pair] (+1: lanes of one pixel hit distinct banks)
  __shared__ typename LP::type s_in[TR][TCOLS][G + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, cg = lane % G, px = lane / G;
  const int x0 = blockIdx.x * TILE_W, y0 = blockIdx.y * LAST_ROWS, n = blockIdx.z;
  for (int i = threadIdx.x; i < TR * TCOLS * G; i += 256) {
    const int g = i % G, c = (i / G) % TCOLS, r = i / (G * TCOLS);
    const int gy = y0 - PAD + r, gx = x0 - PAD + c;
    typename LP::type v{};
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = LP::load(in + (((size_t)n * H + gy) * W + gx) * CIN + 2 * g);
    s_in[r][c][g] = v;
  }
  float wr[TAPS][CPT];
#pragma unroll
  for (int tp = 0; tp < TAPS; ++tp)
#pragma unroll
    for (int c = 0; c < CPT; ++c) wr[tp][c] = __ldg(w + tp * CIN + cg * CPT + c);  // [tap][ci][0]
  __syncthreads();
  const int col = warp * PXW + px, xc = x0 + col;
  const int rows = min(LAST_ROWS, H - y0);
  // accumulator of output row o (tile-relative) lives in slot o % KS; input tile row t feeds outputs t-KS+1 .. t
  float acc[KS];
#pragma unroll
  for (int s = 0; s < KS; ++s) acc[s] = 0.f;
  for (int t0 = 0; t0 < rows + KS - 1; t0 += KS) {
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      const int t = t0 + k;  // input tile row; t % KS == k
      if (t < rows + KS - 1) {
        float2 v[KS];
#pragma unroll
        for (int c = 0; c < KS; ++c) v[c] = LP::get(s_in[t][col + c][cg]);
#pragma unroll
        for (int dy = 0; dy < KS; ++dy) {  // output row o = t - dy, slot (k - dy) mod KS
          float a = acc[(k - dy + KS) % KS];
#pragma unroll
          for (int c = 0; c < KS; ++c) {
            a = fmaf(v[c].x, wr[dy * KS + c][0], a);
            a = fmaf(v[c].y, wr[dy * KS + c][1], a);
          }
          acc[(k - dy + KS) % KS] = a;
        }
        // output row o = t - (KS-1) is complete (its slot is (k+1) % KS); emit it and recycle the slot
        const int o = t - (KS - 1);
        float done = acc[(k + 1) % KS];
        acc[(k + 1) % KS] = 0.f;
#pragma unroll
        for (int m = 1; m < G; m <<= 1) done += __shfl_xor_sync(0xffffffffu, done, m);
        if (o >= 0 && o < rows && cg == 0 && xc < W) out[((size_t)n * H + y0 + o) * W + xc] = done;
      }
    }
  }
}

template <typename TI>
int try_last(const ConvArgs& a, cudaStream_t st) {
  if (a.out_dt != DT_F32 || a.Cout != 1 || a.C1 != 0 || a.scale || a.shift || a.out_pre || a.film_g || a.add_src ||
      a.mask_src || a.relu || !a.out)
    return 0;
  if (a.ks == 5 && a.C0 == 16) {
    constexpr int TILE_W = 8 * (32 / 8);
    constexpr int LAST_ROWS = LastPair<TI>::ROWS;
    dim3 grid((a.W + TILE_W - 1) / TILE_W, (a.H + LAST_ROWS - 1) / LAST_ROWS, a.N);
    conv_last_kernel<TI, 5, 16><<<grid, 256, 0, st>>>((const TI*)a.in0, a.w, (float*)a.out, a.H, a.W);
  } else {
    return 0;
  }
  DG_LAUNCH_CHECK();
  return 1;
}

// ------------------------------------------------------------------------------------------------------
// Weight gradient of a first layer (fp32 input with CIN <= 2 channels): dw[tap][ci][co] += sum_p x[p+off] dy[p][co].
// Same CTA / lane mapping as conv_first_kernel with the roles of weights and accumulators swapped: per pixel one
// coalesced CPT-channel read of dy (fetched KS rows ahead into a register ring), ks new image values from the
// shared-memory tile, TAPS*CIN*CPT FMAs into register accumulators.  Persistent CTAs; reduction = shuffles over the
// warp's pixels, shared-memory atomics over the CTA's warps, one global atomic per element per CTA.
// ------------------------------------------------------------------------------------------------------
template <typename TD, int KS, int CIN, int COUT, int CPT>
__global__ void __launch_bounds__(256, 2) wgrad_first_kernel(const float* __restrict__ x, const TD* __restrict__ dy,
                                                          float* __restrict__ dw, int N, int H, int W, float alpha) {
  constexpr int TAPS = KS * KS, G = COUT / CPT, PXW = 32 / G, KW = KS * CIN, ITEMS = TAPS * CIN * COUT;
  constexpr int TILE_W = 8 * PXW;
  __shared__ EdgeTile<KS, CIN, TILE_W> tile;
  __shared__ float s_acc[ITEMS];
  for (int i = threadIdx.x; i < ITEMS; i += 256) s_acc[i] = 0.f;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, cg = lane % G, px = lane / G, c0 = cg * CPT;
  const int col = warp * PXW + px;
  const int tiles_w = (W + TILE_W - 1) / TILE_W, bands = (H + EDGE_ROWS - 1) / EDGE_ROWS;
  const int n_tiles = N * bands * tiles_w;
  float acc[TAPS * CIN][CPT];
#pragma unroll
  for (int t = 0; t < TAPS * CIN; ++t)
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[t][c] = 0.f;
  for (int tid = blockIdx.x; tid < n_tiles; tid += gridDim.x) {
    const int x0 = (tid % tiles_w) * TILE_W, y0 = ((tid / tiles_w) % bands) * EDGE_ROWS, n = tid / (tiles_w * bands);
    __syncthreads();  // everybody is done with the previous tile
    tile.fill(x, n, y0, x0, H, W);
    const int xc = x0 + col;
    const int rows = xc < W ? min(EDGE_ROWS, H - y0) : 0;
    const TD* dyp = dy + (((size_t)n * H + y0) * W + xc) * COUT + c0;
    const size_t dy_row = (size_t)W * COUT;
    float dring[KS][CPT];  // dy of rows r .. r+KS-1, fetched KS rows ahead of use
#pragma unroll
    for (int k = 0; k < KS; ++k)
      if (k < rows) VecIO<CPT>::load(dyp + k * dy_row, dring[k]);
    __syncthreads();
    if (rows > 0) {
      float win[KS][KW];
#pragma unroll
      for (int d = 0; d < KS - 1; ++d) tile.row(d, col, win[d]);
      for (int r0 = 0; r0 < rows; r0 += KS) {
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          const int r = r0 + k;
          if (r < rows) {
            tile.row(r + KS - 1, col, win[(k + KS - 1) % KS]);
            float d[CPT];
#pragma unroll
            for (int c = 0; c < CPT; ++c) d[c] = dring[k][c];
            if (r + KS < rows) VecIO<CPT>::load(dyp + (size_t)(r + KS) * dy_row, dring[k]);
#pragma unroll
            for (int ty = 0; ty < KS; ++ty)
#pragma unroll
              for (int q = 0; q < KW; ++q) {
                const float v = win[(k + ty) % KS][q];
#pragma unroll
                for (int c = 0; c < CPT; ++c) acc[ty * KW + q][c] = fmaf(v, d[c], acc[ty * KW + q][c]);
              }
          }
        }
      }
    }
  }
  // lanes with equal cg hold partial sums of the same (tap, ci, co) elements
#pragma unroll
  for (int t = 0; t < TAPS * CIN; ++t)
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      float v = acc[t][c];
#pragma unroll
      for (int m = G; m < 32; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
      if (px == 0) atomicAdd(&s_acc[t * COUT + c0 + c], v);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < ITEMS; i += 256) atomicAdd(dw + i, alpha * s_acc[i]);
}

template <typename TD>
int try_wgrad_first(const WgradArgs& a, cudaStream_t st) {
  if (a.x_dt != DT_F32 || a.C1 != 0 || a.W % 8) return 0;
#define DG_WF(KS_, CI_, CO_, CPT_)                                                                             \
  do {                                                                                                         \
    constexpr int TILE_W_ = 8 * (32 / (CO_ / CPT_));                                                           \
    long long grid = (long long)a.N * ((a.H + EDGE_ROWS - 1) / EDGE_ROWS) * ((a.W + TILE_W_ - 1) / TILE_W_);   \
    if (grid > 148 * 2) grid = 148 * 2;                                                                        \
    wgrad_first_kernel<TD, KS_, CI_, CO_, CPT_><<<(unsigned)grid, 256, 0, st>>>(                               \
        (const float*)a.x0, (const TD*)a.dy, a.dw, a.N, a.H, a.W, a.alpha);                                    \
  } while (0)
  if (a.ks == 5 && a.C0 == 1 && a.Cout == 16) DG_WF(5, 1, 16, 2);
  else if (a.ks == 3 && a.C0 == 1 && a.Cout == 32) DG_WF(3, 1, 32, 4);
  else if (a.ks == 3 && a.C0 == 2 && a.Cout == 32) DG_WF(3, 2, 32, 4);
  else return 0;
#undef DG_WF
  DG_LAUNCH_CHECK();
  return 1;
}

// ------------------------------------------------------------------------------------------------------
// wgrad.  Work item = (tap, ci pair, co octet); each item keeps a 2x8 fp32 accumulator and walks the pixels
// of the CTA's tiles; one atomicAdd per output element per CTA at the end.
// ------------------------------------------------------------------------------------------------------
constexpr int WG_CK = 8, WG_COB = 32;

template <typename TX, typename TD, int KS>
__global__ void conv_wgrad_simt_kernel(WgradArgs a, int tiles_per_cta) {
  constexpr int PAD = KS / 2;
  constexpr int IH = TH + KS - 1, IW = TW + KS - 1;
  constexpr int TAPS = KS * KS;
  constexpr int ITEMS = TAPS * (WG_CK / 2) * (WG_COB / 8);
  __shared__ float s_x[WG_CK][IH][IW + 1];
  __shared__ __align__(16) float s_d[TH * TW][WG_COB];

  const int tid = threadIdx.x, nthr = blockDim.x;
  const int Cin = a.C0 + a.C1;
  const int cb = blockIdx.y * WG_CK, cob = blockIdx.z * WG_COB;
  const int tiles_w = (a.W + TW - 1) / TW, tiles_h = (a.H + TH - 1) / TH;
  const int tiles_total = tiles_w * tiles_h * a.N;

  // item decode (one item per thread when tid < ITEMS)
  const bool active = tid < ITEMS;
  const int co8 = tid % (WG_COB / 8);
  const int ci2 = (tid / (WG_COB / 8)) % (WG_CK / 2);
  const int tap = tid / ((WG_COB / 8) * (WG_CK / 2));
  const int ty = active ? tap / KS : 0, tx = active ? tap % KS : 0;

  float acc[2][8];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int t_begin = blockIdx.x * tiles_per_cta;
  const int t_end = min(t_begin + tiles_per_cta, tiles_total);
  for (int t = t_begin; t < t_end; ++t) {
    const int n = t / (tiles_w * tiles_h);
    const int tt = t % (tiles_w * tiles_h);
    const int w0 = (tt % tiles_w) * TW, h0 = (tt / tiles_w) * TH;
    for (int idx = tid; idx < WG_CK * IH * IW; idx += nthr) {
      int ci = idx % WG_CK;
      int c = (idx / WG_CK) % IW;
      int r = idx / (WG_CK * IW);
      int h = h0 + r - PAD, w = w0 + c - PAD, ch = cb + ci;
      float v = 0.f;
      if (h >= 0 && h < a.H && w >= 0 && w < a.W && ch < Cin) {
        size_t pix = ((size_t)n * a.H + h) * a.W + w;
        if (ch < a.C0) v = ld_any<TX>(a.x0, pix * a.C0 + ch);
        else v = ld_any<TX>(a.x1, pix * a.C1 + (ch - a.C0));
      }
      s_x[ci][r][c] = v;
    }
    for (int idx = tid; idx < TH * TW * WG_COB; idx += nthr) {
      int co = idx % WG_COB;
      int p = idx / WG_COB;
      int h = h0 + p / TW, w = w0 + p % TW;
      float v = 0.f;
      if (h < a.H && w < a.W && cob + co < a.Cout)
        v = ld_any<TD>(a.dy, (((size_t)n * a.H + h) * a.W + w) * a.Cout + cob + co);
      s_d[p][co] = v;
    }
    __syncthreads();
    if (active) {
#pragma unroll 4
      for (int p = 0; p < TH * TW; ++p) {
        const int r = p / TW + ty, c = p % TW + tx;
        const float x0 = s_x[2 * ci2][r][c], x1 = s_x[2 * ci2 + 1][r][c];
        const float4 d0 = *reinterpret_cast<const float4*>(&s_d[p][co8 * 8]);
        const float4 d1 = *reinterpret_cast<const float4*>(&s_d[p][co8 * 8 + 4]);
        acc[0][0] = fmaf(x0, d0.x, acc[0][0]); acc[0][1] = fmaf(x0, d0.y, acc[0][1]);
        acc[0][2] = fmaf(x0, d0.z, acc[0][2]); acc[0][3] = fmaf(x0, d0.w, acc[0][3]);
        acc[0][4] = fmaf(x0, d1.x, acc[0][4]); acc[0][5] = fmaf(x0, d1.y, acc[0][5]);
        acc[0][6] = fmaf(x0, d1.z, acc[0][6]); acc[0][7] = fmaf(x0, d1.w, acc[0][7]);
        acc[1][0] = fmaf(x1, d0.x, acc[1][0]); acc[1][1] = fmaf(x1, d0.y, acc[1][1]);
        acc[1][2] = fmaf(x1, d0.z, acc[1][2]); acc[1][3] = fmaf(x1, d0.w, acc[1][3]);
        acc[1][4] = fmaf(x1, d1.x, acc[1][4]); acc[1][5] = fmaf(x1, d1.y, acc[1][5]);
        acc[1][6] = fmaf(x1, d1.z, acc[1][6]); acc[1][7] = fmaf(x1, d1.w, acc[1][7]);
      }
    }
    __syncthreads();
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int ci = cb + 2 * ci2 + i;
      if (ci >= Cin) continue;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int co = cob + co8 * 8 + j;
        if (co < a.Cout) atomicAdd(&a.dw[((size_t)tap * Cin + ci) * a.Cout + co], a.alpha * acc[i][j]);
      }
    }
  }
}

template <typename TX, typename TD>
int launch_wgrad(const WgradArgs& a, cudaStream_t st) {
  const int Cin = a.C0 + a.C1;
  const int tiles = ((a.W + TW - 1) / TW) * ((a.H + TH - 1) / TH) * a.N;
  const int gy = (Cin + WG_CK - 1) / WG_CK, gz = (a.Cout + WG_COB - 1) / WG_COB;
  int target_x = (148 * 8 + gy * gz - 1) / (gy * gz);
  if (target_x < 1) target_x = 1;
  int tiles_per_cta = (tiles + target_x - 1) / target_x;
  if (tiles_per_cta < 1) tiles_per_cta = 1;
  dim3 grid((tiles + tiles_per_cta - 1) / tiles_per_cta, gy, gz);
  switch (a.ks) {
    case 1: conv_wgrad_simt_kernel<TX, TD, 1><<<grid, 32, 0, st>>>(a, tiles_per_cta); break;
    case 3: conv_wgrad_simt_kernel<TX, TD, 3><<<grid, 160, 0, st>>>(a, tiles_per_cta); break;
    case 5: conv_wgrad_simt_kernel<TX, TD, 5><<<grid, 416, 0, st>>>(a, tiles_per_cta); break;
    default: depgan_set_error("conv_wgrad_simt: unsupported kernel size"); return -2;
  }
  DG_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int conv_fwd_simt(const ConvArgs& a, cudaStream_t st) {
  if (a.N <= 0) return 0;
  DG_REQUIRE(!a.deconv && !a.head_w, "conv_fwd_simt: deconv/head fusion are tcgen05-path epilogues");
  {
    // bf16 networks: the first layer runs on the tensor cores (split-bf16 image, conv_first_tc.cu)
    static const bool no_first_tc = getenv("DEPGAN_NO_FIRST_TC") != nullptr;  // A/B switch for measurements
    if (!no_first_tc) {
      const int r0 = conv_first_tc_try(a, st);
      if (r0 != 0) return r0 < 0 ? r0 : 0;
    }
    const int r = a.out_dt == DT_BF16 ? try_first<bf16>(a, st)
                  : a.out_dt == DT_F16 ? try_first<__half>(a, st)
                  : a.out_dt == DT_F16S ? try_first<hsplit>(a, st) : try_first<float>(a, st);
    if (r != 0) return r < 0 ? r : 0;
    {  // 16 -> 1 data gradient of conv2d_dis_0a on the tensor cores (conv_last_band.cu)
      const int rl = conv_last_band_try(a, st);
      if (rl != 0) return rl < 0 ? rl : 0;
    }
    if (a.in_dt != DT_F16) {
      const int r2 = a.in_dt == DT_BF16 ? try_last<bf16>(a, st) : try_last<float>(a, st);
      if (r2 != 0) return r2 < 0 ? r2 : 0;
    }
  }
  DG_REQUIRE(a.in_dt != DT_F16S && a.out_dt != DT_F16S,
             "conv_fwd_simt: split-half storage has no CUDA-core path beyond the first layer (H, W multiples of 128)");
  // IEEE-half handles (generator inference): first layer (fp32 image in) and the levels below 16x16
  if (a.in_dt == DT_F32 && a.out_dt == DT_F16) return launch_fwd<float, __half>(a, st);
  if (a.in_dt == DT_F16 && a.out_dt == DT_F16) return launch_fwd<__half, __half>(a, st);
  if (a.in_dt == DT_F32 && a.out_dt == DT_F32) return launch_fwd<float, float>(a, st);
  if (a.in_dt == DT_F32 && a.out_dt == DT_BF16) return launch_fwd<float, bf16>(a, st);
  if (a.in_dt == DT_BF16 && a.out_dt == DT_BF16) return launch_fwd<bf16, bf16>(a, st);
  if (a.in_dt == DT_BF16 && a.out_dt == DT_F32) return launch_fwd<bf16, float>(a, st);
  depgan_set_error("conv_fwd_simt: bad dtype combination");
  return -2;
}

int conv_wgrad_simt(const WgradArgs& a, cudaStream_t st) {
  if (a.N <= 0) return 0;
  {
    static const bool no_first_tc = getenv("DEPGAN_NO_FIRST_TC") != nullptr;  // A/B switch for measurements
    if (!no_first_tc) {
      const int r0 = wgrad_first_tc_try(a, st);
      if (r0 != 0) return r0 < 0 ? r0 : 0;
    }
    const int r = a.dy_dt == DT_BF16 ? try_wgrad_first<bf16>(a, st) : try_wgrad_first<float>(a, st);
    if (r != 0) return r < 0 ? r : 0;
  }
  if (a.x_dt == DT_F32 && a.dy_dt == DT_F32) return launch_wgrad<float, float>(a, st);
  if (a.x_dt == DT_BF16 && a.dy_dt == DT_BF16) return launch_wgrad<bf16, bf16>(a, st);
  if (a.x_dt == DT_F32 && a.dy_dt == DT_BF16) return launch_wgrad<float, bf16>(a, st);
  if (a.x_dt == DT_BF16 && a.dy_dt == DT_F32) return launch_wgrad<bf16, float>(a, st);
  depgan_set_error("conv_wgrad_simt: bad dtype combination");
  return -2;
}
