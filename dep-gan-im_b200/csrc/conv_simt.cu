// fp32 CUDA-core implicit-GEMM convolution ('same', stride 1, k in {1,3,5}) with the fused epilogues of
// common.cuh, plus the matching weight-gradient kernel.  This is the FP32-accumulate / FP32-storage variant
// (BASELINE.json: "FP32-accumulate variant within 1e-4") and the layer of last resort for shapes the
// tcgen05 path does not take (Cin < 16 first layers, Cout = 1 dgrad of the critic's first layer).
// Replaces Keras Conv2D (+BatchNormalization, +Activation) of TG:285-304 and its K.gradients (TG:543-549).
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
// First-layer direct convolution (conv2d_gen_0: nicg -> 32, conv2d_dis_0a: 1 -> 16).  K = ks*ks*Cin <= 25 is
// far too small for tensor cores; the layer is HBM-bound (4*Cin B in, 2*Cout B out per pixel), so: one thread
// per pixel, fp32 input halo tile and all weights in shared memory, Cout accumulators in registers, 128-bit
// vectorised NHWC stores.
// ------------------------------------------------------------------------------------------------------
template <typename TO, int KS, int CIN, int COUT>
__global__ void __launch_bounds__(256) conv_first_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ scale,
                                                         const float* __restrict__ shift, TO* __restrict__ out,
                                                         const TO* __restrict__ mask, int H, int W, int relu) {
  constexpr int PAD = KS / 2, IT = 16 + KS - 1, TAPS = KS * KS;
  __shared__ float s_in[IT][IT + 1][CIN];
  __shared__ __align__(16) float s_w[TAPS * CIN][COUT];
  __shared__ float s_sc[COUT], s_sh[COUT];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int tiles_w = W / 16;
  const int w0 = (blockIdx.x % tiles_w) * 16, h0 = (blockIdx.x / tiles_w) * 16, n = blockIdx.y;
  for (int i = tid; i < TAPS * CIN * COUT; i += 256) (&s_w[0][0])[i] = w[i];
  for (int i = tid; i < COUT; i += 256) {
    s_sc[i] = scale ? scale[i] : 1.f;
    s_sh[i] = shift ? shift[i] : 0.f;
  }
  for (int i = tid; i < IT * IT * CIN; i += 256) {
    const int ci = i % CIN, c = (i / CIN) % IT, r = i / (CIN * IT);
    const int h = h0 + r - PAD, ww = w0 + c - PAD;
    s_in[r][c][ci] = (h >= 0 && h < H && ww >= 0 && ww < W) ? x[(((size_t)n * H + h) * W + ww) * CIN + ci] : 0.f;
  }
  __syncthreads();
  float acc[COUT];
#pragma unroll
  for (int i = 0; i < COUT; ++i) acc[i] = 0.f;
#pragma unroll
  for (int tap = 0; tap < TAPS; ++tap) {
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      const float v = s_in[ty + tap / KS][tx + tap % KS][ci];
      const float4* wp = reinterpret_cast<const float4*>(&s_w[tap * CIN + ci][0]);
#pragma unroll
      for (int q = 0; q < COUT / 4; ++q) {
        const float4 w4 = wp[q];
        acc[4 * q + 0] = fmaf(v, w4.x, acc[4 * q + 0]);
        acc[4 * q + 1] = fmaf(v, w4.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(v, w4.z, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(v, w4.w, acc[4 * q + 3]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < COUT; ++i) {
    acc[i] = fmaf(acc[i], s_sc[i], s_sh[i]);
    if (relu) acc[i] = fmaxf(acc[i], 0.f);
  }
  const size_t obase = (((size_t)n * H + h0 + ty) * W + w0 + tx) * COUT;
  if (mask) {  // activation-pattern mask of the JVP pass (TG:543: the critic linearised at the mixed sample)
#pragma unroll
    for (int i = 0; i < COUT; ++i) acc[i] = ldf(mask + obase + i) > 0.f ? acc[i] : 0.f;
  }
  TO* o = out + obase;
  if constexpr (sizeof(TO) == 2) {
    uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
    for (int q = 0; q < COUT / 8; ++q) {
      uint4 pk;
      __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
      for (int k = 0; k < 4; ++k) hp[k] = __floats2bfloat162_rn(acc[8 * q + 2 * k], acc[8 * q + 2 * k + 1]);
      o4[q] = pk;
    }
  } else {
    float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
    for (int q = 0; q < COUT / 4; ++q) o4[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
  }
}

template <typename TO, int KS, int CIN, int COUT>
int launch_first(const ConvArgs& a, cudaStream_t st) {
  dim3 grid((a.W / 16) * (a.H / 16), a.N);
  conv_first_kernel<TO, KS, CIN, COUT><<<grid, 256, 0, st>>>((const float*)a.in0, a.w, a.scale, a.shift, (TO*)a.out,
                                                              (const TO*)a.mask_src, a.H, a.W, a.relu);
  DG_LAUNCH_CHECK();
  return 0;
}

// returns 1 when the call was taken by a first-layer specialisation, 0 otherwise, <0 on error
template <typename TO>
int try_first(const ConvArgs& a, cudaStream_t st) {
  if (a.in_dt != DT_F32 || a.C1 != 0 || a.out_pre || a.film_g || a.add_src || !a.out) return 0;
  if (a.H % 16 || a.W % 16) return 0;
  int r = 1;
  if (a.ks == 3 && a.C0 == 1 && a.Cout == 32) r = launch_first<TO, 3, 1, 32>(a, st);
  else if (a.ks == 3 && a.C0 == 2 && a.Cout == 32) r = launch_first<TO, 3, 2, 32>(a, st);
  else if (a.ks == 5 && a.C0 == 1 && a.Cout == 16) r = launch_first<TO, 5, 1, 16>(a, st);
  else return 0;
  return r < 0 ? r : 1;
}


// ------------------------------------------------------------------------------------------------------
// Single-output-channel convolution (the data gradient of conv2d_dis_0a: 16 -> 1, 5x5): dD/dx for the gradient
// penalty and for the generator's adversarial terms.  HBM-bound (2*Cin B in, 4 B out per pixel): one thread per
// pixel, the input halo tile in shared memory as fp32, weights broadcast from shared memory, fp32 output.
// ------------------------------------------------------------------------------------------------------
template <typename TI, int KS, int CIN>
__global__ void __launch_bounds__(256) conv_last_kernel(const TI* __restrict__ in, const float* __restrict__ w,
                                                        float* __restrict__ out, int H, int W) {
  constexpr int PAD = KS / 2, IT = 16 + KS - 1, TAPS = KS * KS;
  __shared__ float s_in[IT * IT][CIN + 1];
  __shared__ float s_w[TAPS * CIN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int tiles_w = W / 16;
  const int w0 = (blockIdx.x % tiles_w) * 16, h0 = (blockIdx.x / tiles_w) * 16, n = blockIdx.y;
  for (int i = tid; i < TAPS * CIN; i += 256) s_w[i] = w[i];  // [tap][ci][0]
  for (int i = tid; i < IT * IT * CIN; i += 256) {
    const int ci = i % CIN, p = i / CIN, c = p % IT, r = p / IT;
    const int h = h0 + r - PAD, ww = w0 + c - PAD;
    s_in[p][ci] = (h >= 0 && h < H && ww >= 0 && ww < W) ? ldf(in + (((size_t)n * H + h) * W + ww) * CIN + ci) : 0.f;
  }
  __syncthreads();
  float acc = 0.f;
#pragma unroll
  for (int tap = 0; tap < TAPS; ++tap) {
    const float* xp = s_in[(ty + tap / KS) * IT + tx + tap % KS];
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) acc = fmaf(xp[ci], s_w[tap * CIN + ci], acc);
  }
  out[((size_t)n * H + h0 + ty) * W + w0 + tx] = acc;
}

template <typename TI>
int try_last(const ConvArgs& a, cudaStream_t st) {
  if (a.out_dt != DT_F32 || a.Cout != 1 || a.C1 != 0 || a.scale || a.shift || a.out_pre || a.film_g || a.add_src ||
      a.mask_src || a.relu || !a.out || a.H % 16 || a.W % 16)
    return 0;
  dim3 grid((a.W / 16) * (a.H / 16), a.N);
  if (a.ks == 5 && a.C0 == 16)
    conv_last_kernel<TI, 5, 16><<<grid, 256, 0, st>>>((const TI*)a.in0, a.w, (float*)a.out, a.H, a.W);
  else
    return 0;
  DG_LAUNCH_CHECK();
  return 1;
}

// ------------------------------------------------------------------------------------------------------
// Weight gradient of a first layer (fp32 input with CIN <= 2 channels): dw[tap][ci][co] += sum_p x[p+off] dy[p][co].
// The result is tiny (<= 25*2*32 floats) and the pass is HBM-bound on dy: CTA = one 16x16 tile, thread =
// (tap, ci, co) item looping the tile's pixels out of shared memory, one atomic per item per CTA.
// ------------------------------------------------------------------------------------------------------
template <typename TD, int KS, int CIN, int COUT>
__global__ void __launch_bounds__(256) wgrad_first_kernel(const float* __restrict__ x, const TD* __restrict__ dy,
                                                          float* __restrict__ dw, int H, int W, int tiles, float alpha) {
  constexpr int PAD = KS / 2, IT = 16 + KS - 1, TAPS = KS * KS, ITEMS = TAPS * CIN * COUT;
  __shared__ float s_x[IT][IT + 1][CIN];
  __shared__ float s_d[256][COUT + 1];
  const int tid = threadIdx.x;
  const int tiles_w = W / 16, tiles_per_img = tiles_w * (H / 16);
  float acc[(ITEMS + 255) / 256];
#pragma unroll
  for (int k = 0; k < (ITEMS + 255) / 256; ++k) acc[k] = 0.f;
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int n = t / tiles_per_img, tt = t % tiles_per_img;
    const int w0 = (tt % tiles_w) * 16, h0 = (tt / tiles_w) * 16;
    __syncthreads();
    for (int i = tid; i < IT * IT * CIN; i += 256) {
      const int ci = i % CIN, c = (i / CIN) % IT, r = i / (CIN * IT);
      const int h = h0 + r - PAD, ww = w0 + c - PAD;
      s_x[r][c][ci] = (h >= 0 && h < H && ww >= 0 && ww < W) ? x[(((size_t)n * H + h) * W + ww) * CIN + ci] : 0.f;
    }
    for (int i = tid; i < 256 * COUT; i += 256) {
      const int co = i % COUT, p = i / COUT;
      s_d[p][co] = ldf(dy + (((size_t)n * H + h0 + p / 16) * W + w0 + p % 16) * COUT + co);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < (ITEMS + 255) / 256; ++k) {
      const int item = tid + 256 * k;
      if (item < ITEMS) {
        const int co = item % COUT, ci = (item / COUT) % CIN, tap = item / (COUT * CIN);
        const int dy_ = tap / KS, dx_ = tap % KS;
        float a = 0.f;
        for (int p = 0; p < 256; ++p) a = fmaf(s_x[(p >> 4) + dy_][(p & 15) + dx_][ci], s_d[p][co], a);
        acc[k] += a;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < (ITEMS + 255) / 256; ++k) {
    const int item = tid + 256 * k;
    if (item < ITEMS) atomicAdd(dw + item, alpha * acc[k]);
  }
}

template <typename TD>
int try_wgrad_first(const WgradArgs& a, cudaStream_t st) {
  if (a.x_dt != DT_F32 || a.C1 != 0 || a.H % 16 || a.W % 16) return 0;
  const int tiles = (a.W / 16) * (a.H / 16) * a.N;
  const int grid = tiles < 148 * 4 ? tiles : 148 * 4;
#define DG_WF(KS_, CI_, CO_)                                                                                   \
  wgrad_first_kernel<TD, KS_, CI_, CO_><<<grid, 256, 0, st>>>((const float*)a.x0, (const TD*)a.dy, a.dw, a.H, a.W, \
                                                              tiles, a.alpha)
  if (a.ks == 5 && a.C0 == 1 && a.Cout == 16) DG_WF(5, 1, 16);
  else if (a.ks == 3 && a.C0 == 1 && a.Cout == 32) DG_WF(3, 1, 32);
  else if (a.ks == 3 && a.C0 == 2 && a.Cout == 32) DG_WF(3, 2, 32);
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
    const int r = a.out_dt == DT_BF16 ? try_first<bf16>(a, st) : try_first<float>(a, st);
    if (r != 0) return r < 0 ? r : 0;
    const int r2 = a.in_dt == DT_BF16 ? try_last<bf16>(a, st) : try_last<float>(a, st);
    if (r2 != 0) return r2 < 0 ? r2 : 0;
  }
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
