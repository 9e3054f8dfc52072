// Forward-side non-convolution kernels: BN folding / weight packing, max-pool, k2s2 transposed conv,
// 1x1 heads, the FiLM noise MLP, critic tail, inference accumulation and bit-exact DEM post-processing.
#include <cuda_fp16.h>

#include "kernels.cuh"

namespace {

constexpr float BN_EPS = 1e-3f;  // Keras BatchNormalization default epsilon

template <typename T>
__device__ __forceinline__ float ldx(const void* p, size_t i) {
  return ldf(reinterpret_cast<const T*>(p) + i);
}
template <typename T>
__device__ __forceinline__ void stx(void* p, size_t i, float v) {
  stf(reinterpret_cast<T*>(p) + i, v);
}

__global__ void fold_bn_kernel(const float* bias, const float* gamma, const float* beta, const float* mean,
                               const float* var, float* scale, float* shift, float* inv_std, int C) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float b = bias ? bias[c] : 0.f;
  if (gamma) {
    const float is = 1.0f / sqrtf(var[c] + BN_EPS);
    const float s = gamma[c] * is;
    scale[c] = s;
    shift[c] = beta[c] + (b - mean[c]) * s;
    if (inv_std) inv_std[c] = is;
  } else {
    scale[c] = 1.f;
    shift[c] = b;
    if (inv_std) inv_std[c] = 1.f;
  }
}

__device__ __forceinline__ bf16 w16(float v, bool f16) {  // weight in the handle's 16-bit format, carried in a bf16 slot
  if (!f16) return __float2bfloat16_rn(v);
  const __half h = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
  return *reinterpret_cast<const bf16*>(&h);
}

// One 32 (ci) x 32 (co) tile of one tap of src[tap][ci][co] -> the three packed operands, through a shared-memory
// transpose so that both the read (co fastest) and the transposed writes (ci fastest) are coalesced (the element-wise
// version wrote w_tc / w_dg with a stride of Cin elements per lane: 46 us for the critic's 1.9 M weights).  256 threads.
__device__ __forceinline__ void pack_tile(const float* __restrict__ src, const float* __restrict__ scale, bf16* dst_tc,
                                          float* dst_dgrad, bf16* dst_tc_dgrad, int taps, int Cin, int Cout, bool f16,
                                          int tile, float (&sm)[32][33]) {
  const int tci = (Cin + 31) >> 5, tco = (Cout + 31) >> 5;
  const int tap = tile / (tci * tco), r = tile - tap * (tci * tco);
  const int ci0 = (r / tco) << 5, co0 = (r % tco) << 5;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;  // 8 warps
  const int ft = taps - 1 - tap;
  __syncthreads();  // the previous tile's reads of sm are done
#pragma unroll
  for (int k = 0; k < 4; ++k) {  // rows ci0 + w + 8k, columns co0 + lane
    const int ci = ci0 + w + 8 * k, co = co0 + lane;
    float v = 0.f;
    if (ci < Cin && co < Cout) {
      v = src[((size_t)tap * Cin + ci) * Cout + co];
      // data-gradient operand in the source orientation: flipped taps, BN scale of the forward output channel folded in
      if (dst_tc_dgrad) dst_tc_dgrad[((size_t)ft * Cin + ci) * Cout + co] = w16(scale ? v * scale[co] : v, f16);
    }
    sm[w + 8 * k][lane] = v;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {  // rows co0 + w + 8k, columns ci0 + lane
    const int co = co0 + w + 8 * k, ci = ci0 + lane;
    if (ci < Cin && co < Cout) {
      const float v = sm[lane][w + 8 * k];
      if (dst_tc) dst_tc[((size_t)tap * Cout + co) * Cin + ci] = w16(v, f16);
      if (dst_dgrad) dst_dgrad[((size_t)ft * Cout + co) * Cin + ci] = scale ? v * scale[co] : v;
    }
  }
}

__global__ void __launch_bounds__(256) pack_conv_kernel(const float* src, const float* scale, bf16* dst_tc,
                                                        float* dst_dgrad, bf16* dst_tc_dgrad, int taps, int Cin, int Cout,
                                                        int f16) {
  __shared__ float sm[32][33];
  const int ntile = taps * ((Cin + 31) >> 5) * ((Cout + 31) >> 5);
  for (int t = blockIdx.x; t < ntile; t += gridDim.x)
    pack_tile(src, scale, dst_tc, dst_dgrad, dst_tc_dgrad, taps, Cin, Cout, f16 != 0, t, sm);
}

// ---- whole-network versions of the two kernels above: blockIdx.y (pack) / blockIdx.x (fold) selects the layer ----
__global__ void prep_fold_all_kernel(const PrepTable t) {
  const PrepLayer& L = t.L[blockIdx.x];
  for (int c = threadIdx.x; c < L.C; c += blockDim.x) {
    const float b = L.bias ? L.bias[c] : 0.f;
    if (L.gamma) {
      const float is = 1.0f / sqrtf(L.var[c] + BN_EPS);
      const float s = L.gamma[c] * is;
      L.scale[c] = s;
      L.shift[c] = L.beta[c] + (b - L.mean[c]) * s;
      if (L.inv_std) L.inv_std[c] = is;
    } else {
      L.scale[c] = 1.f;
      L.shift[c] = b;
      if (L.inv_std) L.inv_std[c] = 1.f;
    }
  }
}
__global__ void __launch_bounds__(256) prep_pack_all_kernel(const PrepTable t) {
  __shared__ float sm[32][33];
  const PrepLayer& L = t.L[blockIdx.y];
  const int ntile = L.taps * ((L.cin + 31) >> 5) * ((L.C + 31) >> 5);
  for (int tl = blockIdx.x; tl < ntile; tl += gridDim.x)
    pack_tile(L.w, L.scale_dgrad ? L.scale : nullptr, L.w_tc, L.w_dg, L.w_dg_tc, L.taps, L.cin, L.C, false, tl, sm);
}

// ---- split-half storage (DT_F16S) ----
__device__ __forceinline__ void split_h(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
  lo = __float2half_rn(v - __half2float(hi));
}
__global__ void pack_split_kernel(const float* src, __half* dst, int taps, int Cin, int Cout, int c0, int kmajor) {
  const size_t total = (size_t)taps * Cin * Cout;
  const int c1 = Cin - c0, K = 3 * Cin;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int co, ci, tap;
    if (kmajor) { ci = i % Cin; co = (i / Cin) % Cout; tap = i / ((size_t)Cout * Cin); }
    else { co = i % Cout; ci = (i / Cout) % Cin; tap = i / ((size_t)Cout * Cin); }
    __half hi, lo;
    split_h(src[i], hi, lo);
    __half* row = dst + ((size_t)tap * Cout + co) * K;
    if (ci < c0) {
      row[ci] = hi; row[c0 + ci] = hi; row[2 * Cin + ci] = lo;
    } else {
      const int cj = ci - c0;
      row[2 * c0 + cj] = hi; row[2 * c0 + c1 + cj] = hi; row[2 * Cin + ci] = lo;
    }
  }
}
// 2x2 max-pool of (hi, lo) pairs, 8 channels per thread.  hi = half(v) is monotonic in v, so the window's maximum is the
// pair with the largest hi and, among equal hi, the largest lo.
__global__ void maxpool_fwd_split_kernel(const uint4* in, uint4* out, int N, int H, int W, int C8) {
  const int Ho = H / 2, Wo = W / 2;
  const size_t total = (size_t)N * Ho * Wo * C8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = i % C8;
    const size_t p = i / C8;
    const int wo = p % Wo, ho = (p / Wo) % Ho;
    const size_t n = p / ((size_t)Wo * Ho);
    // a pixel = 2*C8 vectors: C8 of hi, then C8 of lo
    const uint4* b = in + (((size_t)n * H + 2 * ho) * W + 2 * wo) * (2 * C8) + c;
    const size_t offs[4] = {0, (size_t)2 * C8, (size_t)W * 2 * C8, (size_t)W * 2 * C8 + 2 * C8};
    uint4 bh = __ldg(b), bl = __ldg(b + C8);
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      const uint4 qh = __ldg(b + offs[k]), ql = __ldg(b + offs[k] + C8);
      __half* rh = reinterpret_cast<__half*>(&bh);
      __half* rl = reinterpret_cast<__half*>(&bl);
      const __half* xh = reinterpret_cast<const __half*>(&qh);
      const __half* xl = reinterpret_cast<const __half*>(&ql);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float a = __half2float(rh[j]), bq = __half2float(xh[j]);
        const bool take = bq > a || (bq == a && __half2float(xl[j]) > __half2float(rl[j]));
        if (take) { rh[j] = xh[j]; rl[j] = xl[j]; }
      }
    }
    uint4* o = out + p * (2 * C8) + c;
    o[0] = bh;
    o[C8] = bl;
  }
}
__global__ void split_to_f32_kernel(const __half* src, float* dst, long long npix, int C) {
  const long long total = npix * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / C;
    const int c = (int)(i - p * C);
    dst[i] = __half2float(src[p * 2 * C + c]) + __half2float(src[p * 2 * C + C + c]);
  }
}

template <typename T>
__global__ void maxpool_fwd_kernel(const T* in, T* out, int N, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2;
  size_t total = (size_t)N * Ho * Wo * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int c = i % C;
    size_t p = i / C;
    int wo = p % Wo;
    int ho = (p / Wo) % Ho;
    int n = p / ((size_t)Wo * Ho);
    const T* b = in + (((size_t)n * H + 2 * ho) * W + 2 * wo) * C + c;
    float v = fmaxf(fmaxf(ldf(b), ldf(b + C)), fmaxf(ldf(b + (size_t)W * C), ldf(b + (size_t)W * C + C)));
    stf(out + i, v);
  }
}

// bf16 / fp16 max-pool, 8 channels (128 bits) per thread: 4 vector loads, 1 vector store
template <typename H2>
__global__ void maxpool_fwd_bf16x8_kernel(const uint4* in, uint4* out, int N, int H, int W, int C8) {
  const int Ho = H / 2, Wo = W / 2;
  const size_t total = (size_t)N * Ho * Wo * C8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = i % C8;
    const size_t p = i / C8;
    const int wo = p % Wo, ho = (p / Wo) % Ho;
    const size_t n = p / ((size_t)Wo * Ho);
    const uint4* b = in + (((size_t)n * H + 2 * ho) * W + 2 * wo) * C8 + c;
    uint4 q[4] = {__ldg(b), __ldg(b + C8), __ldg(b + (size_t)W * C8), __ldg(b + (size_t)W * C8 + C8)};
    uint4 r;
    H2* rr = reinterpret_cast<H2*>(&r);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const H2 a0 = reinterpret_cast<const H2*>(&q[0])[k];
      const H2 a1 = reinterpret_cast<const H2*>(&q[1])[k];
      const H2 a2 = reinterpret_cast<const H2*>(&q[2])[k];
      const H2 a3 = reinterpret_cast<const H2*>(&q[3])[k];
      rr[k] = __hmax2(__hmax2(a0, a1), __hmax2(a2, a3));
    }
    out[i] = r;
  }
}

// Transposed conv k2 s2: one CTA = 16 input pixels; thread o -> output (a,b,co) index, 16 pixel accumulators.
template <typename T>
__global__ void __launch_bounds__(128) deconv_fwd_kernel(const T* in, const float* w, const float* scale,
                                                         const float* shift, T* out, int N, int H, int W, int Cin,
                                                         int Cout, int relu) {
  extern __shared__ float s_x[];  // [16][Cin]
  const size_t npix = (size_t)N * H * W;
  const size_t p0 = (size_t)blockIdx.x * 16;
  for (int idx = threadIdx.x; idx < 16 * Cin; idx += blockDim.x) {
    size_t p = p0 + idx / Cin;
    s_x[idx] = p < npix ? ldf(in + p * Cin + idx % Cin) : 0.f;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < 4 * Cout; o += blockDim.x) {
    const int co = o % Cout, ab = o / Cout;
    const float* wr = w + (size_t)o * Cin;  // Keras (2,2,Cout,Cin): [(a*2+b)*Cout + co][ci]
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    for (int ci = 0; ci < Cin; ++ci) {
      const float wv = wr[ci];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(s_x[i * Cin + ci], wv, acc[i]);
    }
    const float s = scale ? scale[co] : 1.f, t = shift ? shift[co] : 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      size_t p = p0 + i;
      if (p >= npix) break;
      int wi = p % W;
      int hi = (p / W) % H;
      size_t n = p / ((size_t)W * H);
      size_t op = ((n * 2 * H + 2 * hi + (ab >> 1)) * 2 * W + 2 * wi + (ab & 1));
      const float v = fmaf(acc[i], s, t);
      stf(out + op * Cout + co, relu ? fmaxf(v, 0.f) : v);
    }
  }
}

template <typename T>
__global__ void head_fwd_kernel(const T* in, const float* w, const float* b, float* out, long long npix, int Cin,
                                int nc_out, int head) {
  long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= npix) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const T* x = in + p * Cin;
  for (int c = 0; c < Cin; ++c) {
    float v = ldf(x + c);
    for (int k = 0; k < nc_out; ++k) acc[k] = fmaf(v, w[c * nc_out + k], acc[k]);
  }
  for (int k = 0; k < nc_out; ++k) acc[k] += b[k];
  if (head == 0) {
    for (int k = 0; k < nc_out; ++k) out[p * nc_out + k] = tanhf(acc[k]);
  } else if (head == 1) {
    float m = acc[0];
    for (int k = 1; k < nc_out; ++k) m = fmaxf(m, acc[k]);
    float e[4], s = 0.f;
    for (int k = 0; k < nc_out; ++k) { e[k] = expf(acc[k] - m); s += e[k]; }
    for (int k = 0; k < nc_out; ++k) out[p * nc_out + k] = e[k] / s;
  } else {
    for (int k = 0; k < nc_out; ++k) out[p * nc_out + k] = acc[k];
  }
}

// FiLM MLP stage A: one CTA per sample, L*F threads (<=1024).
__global__ void film_stage_a_kernel(FilmMlpArgs a) {
  extern __shared__ float s_h1[];  // [L][F]
  const int n = blockIdx.x, L = a.L, Fd = a.F;
  const int l = threadIdx.x / Fd, f = threadIdx.x % Fd;
  float v = a.z[(size_t)n * L + l] * a.k0[f];
  v = fmaxf(fmaf(v, a.s0[f], a.t0[f]), 0.f);
  s_h1[l * Fd + f] = v;
  a.h1[((size_t)n * L + l) * Fd + f] = v;
  __syncthreads();
  float acc = 0.f;
  for (int k = 0; k < Fd; ++k) acc = fmaf(s_h1[l * Fd + k], a.k1[k * Fd + f], acc);
  a.h2[((size_t)n * L + l) * Fd + f] = fmaxf(fmaf(acc, a.s1[f], a.t1[f]), 0.f);
}

// FiLM MLP stage B: out[n, off_h + c] = s_h[c] * sum_k h2[n,k] W_h[k,c] + t_h[c].
// CTA = 32 columns x 8 samples; 512 threads = 32 columns x 16 k-slices.
constexpr int FILM_KS = 16;
__global__ void __launch_bounds__(32 * FILM_KS) film_stage_b_kernel(FilmMlpArgs a) {
  extern __shared__ float s_h2[];  // [8][K]
  __shared__ float s_red[FILM_KS][8][32];
  const int K = a.L * a.F;
  const int col0 = blockIdx.x * 32, n0 = blockIdx.y * 8;
  int h = 0;
  for (int i = 0; i < a.n_heads; ++i)
    if (col0 >= a.head_off[i]) h = i;
  const int C = a.head_c[h], cl0 = col0 - a.head_off[h];
  const float* W = a.head_w[h];
  for (int idx = threadIdx.x; idx < 8 * K; idx += blockDim.x) {
    int s = idx / K;
    s_h2[idx] = (n0 + s < a.N) ? a.h2[(size_t)(n0 + s) * K + idx % K] : 0.f;
  }
  __syncthreads();
  const int cx = threadIdx.x & 31, kq = threadIdx.x >> 5;
  float acc[8];
#pragma unroll
  for (int s = 0; s < 8; ++s) acc[s] = 0.f;
  const int kpart = (K + FILM_KS - 1) / FILM_KS;
  const int k_end = min(K, (kq + 1) * kpart);
  // the weight loads are the latency: keep 8 of them in flight
  int k = kq * kpart;
  for (; k + 8 <= k_end; k += 8) {
    float wv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) wv[u] = __ldg(W + (size_t)(k + u) * C + cl0 + cx);
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int s = 0; s < 8; ++s) acc[s] = fmaf(s_h2[s * K + k + u], wv[u], acc[s]);
  }
  for (; k < k_end; ++k) {
    const float wv = __ldg(W + (size_t)k * C + cl0 + cx);
#pragma unroll
    for (int s = 0; s < 8; ++s) acc[s] = fmaf(s_h2[s * K + k], wv, acc[s]);
  }
#pragma unroll
  for (int s = 0; s < 8; ++s) s_red[kq][s][cx] = acc[s];
  __syncthreads();
  for (int idx = threadIdx.x; idx < 8 * 32; idx += blockDim.x) {
    int s = idx >> 5, c = idx & 31;
    if (n0 + s >= a.N) continue;
    float v = 0.f;
#pragma unroll
    for (int kk = 0; kk < FILM_KS; ++kk) v += s_red[kk][s][c];
    a.out[(size_t)(n0 + s) * a.total_c + col0 + c] = fmaf(v, a.head_s[h][cl0 + c], a.head_t[h][cl0 + c]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) critic_head_fwd_kernel(const T* in, const float* w9, const float* b9,
                                                              const float* wd, const float* bd, float* out, int HW,
                                                              int C) {
  __shared__ float s_part[8];
  const int n = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc = 0.f;
  for (int p = warp; p < HW; p += 8) {
    const T* x = in + ((size_t)n * HW + p) * C;
    float d = 0.f;
    for (int c = lane; c < C; c += 32) d = fmaf(ldf(x + c), w9[c], d);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    acc = fmaf(d + b9[0], wd[p], acc);
  }
  if (lane == 0) s_part[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = bd[0];
    for (int i = 0; i < 8; ++i) s += s_part[i];
    out[n] = s;
  }
}

// bf16 fast path (C % 8 == 0): 16-byte loads, each lane owns 8 channels of a pixel
__global__ void __launch_bounds__(256) critic_head_fwd_vec_kernel(const bf16* in, const float* w9, const float* b9,
                                                                  const float* wd, const float* bd, float* out, int HW,
                                                                  int C) {
  __shared__ float s_part[8];
  const int n = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cv = C / 8;
  float acc = 0.f;
  for (int cb = lane; cb < cv; cb += 32) {
    float w8[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) w8[k] = w9[cb * 8 + k];
    for (int p = warp; p < HW; p += 8) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(in + ((size_t)n * HW + p) * C + cb * 8));
      const uint32_t* u = &q.x;
      float d = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        d = fmaf(__uint_as_float(u[k] << 16), w8[2 * k], d);
        d = fmaf(__uint_as_float(u[k] & 0xFFFF0000u), w8[2 * k + 1], d);
      }
      acc = fmaf(d, wd[p], acc);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) s_part[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = bd[0];
    for (int i = 0; i < 8; ++i) s += s_part[i];
    float swd = 0.f;  // the dis_9 bias reaches every pixel of the Dense layer
    for (int p = 0; p < HW; ++p) swd += wd[p];
    out[n] = s + b9[0] * swd;
  }
}

template <typename T>
__global__ void convert_in_kernel(const float* src, T* dst, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    stf(dst + i, src[i]);
}
template <typename T>
__global__ void copy_to_f32_kernel(const T* src, float* dst, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = ldf(src + i);
}

template <typename T>
__global__ void critic_inputs_kernel(const float* real2, const float* x1, int nicg, const float* dem,
                                     const float* ep, int which, T* batch3, int N, long long hw) {
  const long long total = (long long)N * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const float base = x1[i * nicg];
    const float d = dem[i];
    float real, fake;
    if (which == 0) { real = real2[i]; fake = base + d; }
    else { real = real2[i] - base; fake = d; }
    const float e = ep[i / hw];
    stf(batch3 + i, real);
    stf(batch3 + total + i, fake);
    stf(batch3 + 2 * total + i, e * real + (1.f - e) * fake);
  }
}

// ---- inference accumulation / post-processing: integer / label results must be bit-exact (EG:617-741) ----
__global__ void dem_accumulate_kernel(double* acc, const float* pred, const float* mask, long long n, int chan) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float prod = __fmul_rn(pred[i], mask[i / chan]);  // float32 multiply, as np.multiply(f32, f32)
    acc[i] = __dadd_rn(acc[i], (double)prod);               // float64 accumulate (np.zeros default dtype)
  }
}

__global__ void dem_postproc_kernel(const float* x, int nicg, const double* acc, double n_repeat, const float* mask,
                                    double thr, double* dem_out, double* fake2_out, unsigned char* labels,
                                    unsigned long long* count, long long npix) {
  const float thr32 = (float)thr;  // NumPy compares a float32 array with a Python float in float32
  unsigned long long local = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix;
       i += (long long)gridDim.x * blockDim.x) {
    const double dem = __ddiv_rn(acc[i], n_repeat);         // EG:628
    const float base = x[i * nicg];
    double f2 = __dadd_rn((double)base, dem);               // EG:675
    if (f2 < -1.0) f2 = -1.0;                               // EG:676
    if (f2 > 1.0) f2 = 1.0;                                 // EG:677
    if (dem_out) dem_out[i] = dem;
    if (fake2_out) fake2_out[i] = f2;
    if (f2 > thr && mask[i] != 0.f) ++local;                // EG:679-682 (strict >, times mask, count_nonzero)
    const bool f_ge = f2 >= thr, f_lt = f2 < thr;           // EG:723-741 (NaN: every comparison false)
    const bool b_ge = base >= thr32, b_lt = base < thr32;
    unsigned char lab = 0;
    if (f_lt && b_ge) lab = 1;
    if (f_ge && b_lt) lab = 2;
    if (f_ge && b_ge) lab = 3;
    if (labels) labels[i] = lab;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, local);
}

__global__ void uresnet_labels_kernel(const double* acc, double n_repeat, int chan, double* mean_out,
                                      unsigned char* labels, unsigned long long* count, long long npix) {
  unsigned long long local = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix;
       i += (long long)gridDim.x * blockDim.x) {
    int best = 0;
    double bv = __ddiv_rn(acc[i * chan], n_repeat);
    if (mean_out) mean_out[i * chan] = bv;
    for (int k = 1; k < chan; ++k) {
      double v = __ddiv_rn(acc[i * chan + k], n_repeat);
      if (mean_out) mean_out[i * chan + k] = v;
      if (v > bv) { bv = v; best = k; }  // np.argmax: first maximum wins (EU:180)
    }
    labels[i] = (unsigned char)best;
    if (best > 0) ++local;                // EU:597-600
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, local);
}

// 4x4 confusion counts of two label maps with values 0..3 (conf[4*real + fake]); everything the evaluation rows need
// (EG:688-807, EU:606-704) derives from these 16 integers.  Labels outside 0..3 are counted nowhere.
__global__ void __launch_bounds__(256) label_confusion_kernel(const unsigned char* fake, const unsigned char* real,
                                                              long long n, unsigned long long* conf) {
  __shared__ unsigned int s_c[16];
  if (threadIdx.x < 16) s_c[threadIdx.x] = 0;
  __syncthreads();
  unsigned int local[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) local[k] = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned int f = fake[i], r = real[i];
    if (f < 4u && r < 4u) {
      const unsigned int idx = 4u * r + f;
#pragma unroll
      for (int k = 0; k < 16; ++k) local[k] += (idx == (unsigned)k) ? 1u : 0u;
    }
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    unsigned int v = local[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_c[k], v);
  }
  __syncthreads();
  if (threadIdx.x < 16 && s_c[threadIdx.x]) atomicAdd(conf + threadIdx.x, (unsigned long long)s_c[threadIdx.x]);
}

__global__ void adam_kernel(float* p, const float* g, float* m, float* v, long long n, float lr_t, float b1, float b2,
                            float eps, float gscale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
  }
}

inline int grid_for(long long n, int block = 256, int cap = 148 * 16) {
  long long g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

}  // namespace

int k_fold_bn(const float* bias, const float* gamma, const float* beta, const float* mean, const float* var,
              float* scale, float* shift, float* inv_std, int C, cudaStream_t st) {
  fold_bn_kernel<<<(C + 127) / 128, 128, 0, st>>>(bias, gamma, beta, mean, var, scale, shift, inv_std, C);
  DG_LAUNCH_CHECK();
  return 0;
}

int k_pack_conv_weights(const float* src, const float* scale, bf16* dst_tc, float* dst_dgrad, bf16* dst_tc_dgrad,
                        int taps, int Cin, int Cout, cudaStream_t st, int f16) {
  if (!dst_tc && !dst_dgrad && !dst_tc_dgrad) return 0;
  const int ntile = taps * ((Cin + 31) / 32) * ((Cout + 31) / 32);
  pack_conv_kernel<<<ntile < 148 * 8 ? ntile : 148 * 8, 256, 0, st>>>(src, scale, dst_tc, dst_dgrad, dst_tc_dgrad, taps, Cin,
                                                                      Cout, f16);
  DG_LAUNCH_CHECK();
  return 0;
}

int k_pack_split_weights(const float* src, bf16* dst, int taps, int Cin, int Cout, int c0, int kmajor, cudaStream_t st) {
  pack_split_kernel<<<grid_for((long long)taps * Cin * Cout), 256, 0, st>>>(src, reinterpret_cast<__half*>(dst), taps, Cin,
                                                                            Cout, c0, kmajor);
  DG_LAUNCH_CHECK();
  return 0;
}
int k_split_to_f32(const void* src, float* dst, long long npix, int C, cudaStream_t st) {
  if (npix * C == 0) return 0;
  split_to_f32_kernel<<<grid_for(npix * C), 256, 0, st>>>((const __half*)src, dst, npix, C);
  DG_LAUNCH_CHECK();
  return 0;
}

int k_prepare_convs(const PrepTable& t, cudaStream_t st) {
  if (t.n <= 0) return 0;
  prep_fold_all_kernel<<<t.n, 256, 0, st>>>(t);
  DG_LAUNCH_CHECK();
  prep_pack_all_kernel<<<dim3(148, t.n), 256, 0, st>>>(t);  // 148 CTAs per layer walk its 32 x 32 tiles (576 per tap-9 256 x 256 layer)
  DG_LAUNCH_CHECK();
  return 0;
}

int k_maxpool_fwd(const void* in, void* out, int N, int H, int W, int C, int dt, cudaStream_t st) {
  long long total = (long long)N * (H / 2) * (W / 2) * C;
  if (total == 0) return 0;
  if (dt == DT_F16S) {
    DG_REQUIRE(C % 8 == 0, "maxpool (split-half storage): C must be a multiple of 8");
    maxpool_fwd_split_kernel<<<grid_for(total / 8, 256, 148 * 32), 256, 0, st>>>((const uint4*)in, (uint4*)out, N, H, W,
                                                                                  C / 8);
  } else if (dt == DT_F32)
    maxpool_fwd_kernel<float><<<grid_for(total), 256, 0, st>>>((const float*)in, (float*)out, N, H, W, C);
  else if (C % 8 == 0 && dt == DT_F16)
    maxpool_fwd_bf16x8_kernel<__half2><<<grid_for(total / 8, 256, 148 * 32), 256, 0, st>>>((const uint4*)in, (uint4*)out,
                                                                                          N, H, W, C / 8);
  else if (C % 8 == 0)
    maxpool_fwd_bf16x8_kernel<__nv_bfloat162><<<grid_for(total / 8, 256, 148 * 32), 256, 0, st>>>(
        (const uint4*)in, (uint4*)out, N, H, W, C / 8);
  else if (dt == DT_F16)
    maxpool_fwd_kernel<__half><<<grid_for(total), 256, 0, st>>>((const __half*)in, (__half*)out, N, H, W, C);
  else
    maxpool_fwd_kernel<bf16><<<grid_for(total), 256, 0, st>>>((const bf16*)in, (bf16*)out, N, H, W, C);
  DG_LAUNCH_CHECK();
  return 0;
}

int k_deconv_fwd(const void* in, const float* w, const float* scale, const float* shift, void* out, int N, int H,
                 int W, int Cin, int Cout, int dt, int relu, cudaStream_t st) {
  long long npix = (long long)N * H * W;
  if (npix == 0) return 0;
  int grid = (int)((npix + 15) / 16);
  size_t smem = 16 * Cin * sizeof(float);
  if (dt == DT_F32)
    deconv_fwd_kernel<float><<<grid, 128, smem, st>>>((const float*)in, w, scale, shift, (float*)out, N, H, W, Cin, Cout,
                                                      relu);
  else if (dt == DT_F16)
    deconv_fwd_kernel<__half><<<grid, 128, smem, st>>>((const __half*)in, w, scale, shift, (__half*)out, N, H, W, Cin,
                                                       Cout, relu);
  else
    deconv_fwd_kernel<bf16><<<grid, 128, smem, st>>>((const bf16*)in, w, scale, shift, (bf16*)out, N, H, W, Cin, Cout,
                                                     relu);
  DG_LAUNCH_CHECK();
  return 0;
}

int k_head_fwd(const void* in, const float* w, const float* b, float* out, long long npix, int Cin, int nc_out,
               int head, int dt, cudaStream_t st) {
  if (npix == 0) return 0;
  DG_REQUIRE(nc_out >= 1 && nc_out <= 4, "head: nc_out must be 1..4");
  int grid = (int)((npix + 127) / 128);
  if (dt == DT_F32)
    head_fwd_kernel<float><<<grid, 128, 0, st>>>((const float*)in, w, b, out, npix, Cin, nc_out, head);
  else if (dt == DT_F16)
    head_fwd_kernel<__half><<<grid, 128, 0, st>>>((const __half*)in, w, b, out, npix, Cin, nc_out, head);
  else
    head_fwd_kernel<bf16><<<grid, 128, 0, st>>>((const bf16*)in, w, b, out, npix, Cin, nc_out, head);
  DG_LAUNCH_CHECK();
  return 0;
}

int k_film_mlp_fwd(const FilmMlpArgs& a, cudaStream_t st) {
  if (a.N == 0) return 0;
  DG_REQUIRE(a.L * a.F <= 1024, "film mlp: L*F must be <= 1024");
  DG_REQUIRE(a.total_c % 32 == 0, "film mlp: head widths must be multiples of 32");
  film_stage_a_kernel<<<a.N, a.L * a.F, a.L * a.F * sizeof(float), st>>>(a);
  DG_LAUNCH_CHECK();
  dim3 grid(a.total_c / 32, (a.N + 7) / 8);
  film_stage_b_kernel<<<grid, 32 * FILM_KS, 8 * a.L * a.F * sizeof(float), st>>>(a);
  DG_LAUNCH_CHECK();
  return 0;
}

int k_critic_head_fwd(const void* in, const float* w9, const float* b9, const float* wd, const float* bd, float* out,
                      int N, int HW, int C, int dt, cudaStream_t st) {
  if (N == 0) return 0;
  if (dt == DT_F32)
    critic_head_fwd_kernel<float><<<N, 256, 0, st>>>((const float*)in, w9, b9, wd, bd, out, HW, C);
  else if (C % 8 == 0)
    critic_head_fwd_vec_kernel<<<N, 256, 0, st>>>((const bf16*)in, w9, b9, wd, bd, out, HW, C);
  else
    critic_head_fwd_kernel<bf16><<<N, 256, 0, st>>>((const bf16*)in, w9, b9, wd, bd, out, HW, C);
  DG_LAUNCH_CHECK();
  return 0;
}

int k_convert_in(const float* src, void* dst, long long n, int dt, cudaStream_t st) {
  if (n == 0) return 0;
  if (dt == DT_F32)
    convert_in_kernel<float><<<grid_for(n), 256, 0, st>>>(src, (float*)dst, n);
  else if (dt == DT_F16)
    convert_in_kernel<__half><<<grid_for(n), 256, 0, st>>>(src, (__half*)dst, n);
  else
    convert_in_kernel<bf16><<<grid_for(n), 256, 0, st>>>(src, (bf16*)dst, n);
  DG_LAUNCH_CHECK();
  return 0;
}

// float32 -> float16 / bfloat16, 8 elements (two 128-bit loads, one 128-bit store) per thread; the opt-in narrow
// device->host output of predict().  HBM-bound: 4 B read + 2 B written per element.
template <typename H2, typename CVT>
__global__ void cast_narrow_kernel(const float* __restrict__ src, uint4* __restrict__ dst, long long n8, CVT cvt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const float4 a = __ldg(reinterpret_cast<const float4*>(src) + 2 * i);
  const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
  H2 h[4] = {cvt(a.x, a.y), cvt(a.z, a.w), cvt(b.x, b.y), cvt(b.z, b.w)};
  dst[i] = *reinterpret_cast<uint4*>(h);
}
__global__ void cast_f16_tail_kernel(const float* src, __half* dst, long long from, long long n) {
  const long long i = from + threadIdx.x;
  if (i < n) dst[i] = __float2half_rn(src[i]);
}
__global__ void cast_bf16_tail_kernel(const float* src, bf16* dst, long long from, long long n) {
  const long long i = from + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16_rn(src[i]);
}
struct CvtF16 { __device__ __half2 operator()(float x, float y) const { return __floats2half2_rn(x, y); } };
struct CvtBf16 { __device__ __nv_bfloat162 operator()(float x, float y) const { return __floats2bfloat162_rn(x, y); } };

int k_cast_narrow(const float* src, void* dst, long long n, int to_f16, cudaStream_t st) {
  if (n <= 0) return 0;
  DG_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
             "cast_narrow: buffers must be 16-byte aligned");
  const long long n8 = n / 8;
  if (n8 > 0) {
    const unsigned grid = (unsigned)((n8 + 255) / 256);
    if (to_f16) cast_narrow_kernel<__half2><<<grid, 256, 0, st>>>(src, (uint4*)dst, n8, CvtF16());
    else cast_narrow_kernel<__nv_bfloat162><<<grid, 256, 0, st>>>(src, (uint4*)dst, n8, CvtBf16());
    DG_LAUNCH_CHECK();
  }
  if (n8 * 8 < n) {
    if (to_f16) cast_f16_tail_kernel<<<1, 8, 0, st>>>(src, (__half*)dst, n8 * 8, n);
    else cast_bf16_tail_kernel<<<1, 8, 0, st>>>(src, (bf16*)dst, n8 * 8, n);
    DG_LAUNCH_CHECK();
  }
  return 0;
}

int k_copy_to_f32(const void* src, float* dst, long long n, int dt, cudaStream_t st) {
  if (n == 0) return 0;
  if (dt == DT_F32)
    copy_to_f32_kernel<float><<<grid_for(n), 256, 0, st>>>((const float*)src, dst, n);
  else if (dt == DT_F16)
    copy_to_f32_kernel<__half><<<grid_for(n), 256, 0, st>>>((const __half*)src, dst, n);
  else
    copy_to_f32_kernel<bf16><<<grid_for(n), 256, 0, st>>>((const bf16*)src, dst, n);
  DG_LAUNCH_CHECK();
  return 0;
}

int k_critic_inputs(const float* real2, const float* x1, int nicg, const float* dem, const float* ep, int which,
                    void* batch3, int N, long long hw, int dt, cudaStream_t st) {
  long long total = (long long)N * hw;
  if (total == 0) return 0;
  if (dt == DT_F32)
    critic_inputs_kernel<float><<<grid_for(total), 256, 0, st>>>(real2, x1, nicg, dem, ep, which, (float*)batch3, N, hw);
  else
    critic_inputs_kernel<bf16><<<grid_for(total), 256, 0, st>>>(real2, x1, nicg, dem, ep, which, (bf16*)batch3, N, hw);
  DG_LAUNCH_CHECK();
  return 0;
}

int k_dem_accumulate(double* acc, const float* pred, const float* mask, long long n, int chan, cudaStream_t st) {
  if (n == 0) return 0;
  dem_accumulate_kernel<<<grid_for(n), 256, 0, st>>>(acc, pred, mask, n, chan);
  DG_LAUNCH_CHECK();
  return 0;
}

int k_dem_postproc(const float* x, int nicg, const double* acc, double n_repeat, const float* mask, double thr,
                   double* dem_out, double* fake2_out, unsigned char* labels, unsigned long long* count,
                   long long npix, cudaStream_t st) {
  DG_CHECK_CUDA(cudaMemsetAsync(count, 0, sizeof(unsigned long long), st));
  if (npix == 0) return 0;
  dem_postproc_kernel<<<grid_for(npix), 256, 0, st>>>(x, nicg, acc, n_repeat, mask, thr, dem_out, fake2_out, labels,
                                                      count, npix);
  DG_LAUNCH_CHECK();
  return 0;
}

int k_uresnet_labels(const double* acc, double n_repeat, int chan, double* mean_out, unsigned char* labels,
                     unsigned long long* count, long long npix, cudaStream_t st) {
  DG_CHECK_CUDA(cudaMemsetAsync(count, 0, sizeof(unsigned long long), st));
  if (npix == 0) return 0;
  uresnet_labels_kernel<<<grid_for(npix), 256, 0, st>>>(acc, n_repeat, chan, mean_out, labels, count, npix);
  DG_LAUNCH_CHECK();
  return 0;
}

int k_label_confusion(const unsigned char* fake, const unsigned char* real, long long n, unsigned long long* conf,
                      cudaStream_t st) {
  DG_CHECK_CUDA(cudaMemsetAsync(conf, 0, 16 * sizeof(unsigned long long), st));
  if (n == 0) return 0;
  // per-thread counters are 32-bit: bound the elements a thread sees
  long long blocks = (n + 256LL * 4096 - 1) / (256LL * 4096);
  if (blocks < 148 * 4) blocks = 148 * 4;
  if (blocks > (n + 255) / 256) blocks = (n + 255) / 256;
  label_confusion_kernel<<<(unsigned)blocks, 256, 0, st>>>(fake, real, n, conf);
  DG_LAUNCH_CHECK();
  return 0;
}

int k_adam(float* p, const float* g, float* m, float* v, long long n, float lr_t, float b1, float b2, float eps,
           float gscale, cudaStream_t st) {
  if (n == 0) return 0;
  adam_kernel<<<grid_for(n), 256, 0, st>>>(p, g, m, v, n, lr_t, b1, b2, eps, gscale);
  DG_LAUNCH_CHECK();
  return 0;
}
