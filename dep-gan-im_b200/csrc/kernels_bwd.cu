// Backward-side non-convolution kernels of the DEP-GAN train-step graphs (TG:523-598): max-pool routing, reductions
// for bias / BatchNorm / FiLM parameter gradients, the critic tail, gradient-penalty pieces, loss sums and the
// FiLM noise-MLP backward.  The heavy contractions (dgrad, JVP, wgrad) are convolution kernels
// (conv_tc.cu / conv_simt.cu); everything here is bandwidth- or latency-bound warp-level work.
#include "kernels.cuh"

namespace {

inline int grid_for(long long n, int block = 256, int cap = 148 * 16) {
  long long g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// index (0..3) of the first maximum in the 2x2 window (row-major scan), as TF/Keras MaxPoolGrad routes it
template <typename T>
__device__ __forceinline__ int first_argmax(const T* b, size_t C, size_t WC) {
  float v0 = ldf(b), v1 = ldf(b + C), v2 = ldf(b + WC), v3 = ldf(b + WC + C);
  int k = 0;
  float m = v0;
  if (v1 > m) { m = v1; k = 1; }
  if (v2 > m) { m = v2; k = 2; }
  if (v3 > m) { m = v3; k = 3; }
  return k;
}

// ---- 8-channel (16-byte) bf16 vector helpers for the bandwidth-bound passes below ----
struct Bf8 {
  uint4 q;
  __device__ __forceinline__ void load(const bf16* p) { q = __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ __forceinline__ void store(bf16* p) const { *reinterpret_cast<uint4*>(p) = q; }
  __device__ __forceinline__ float get(int i) const {
    const uint32_t w = (&q.x)[i >> 1];
    return __uint_as_float((i & 1) ? (w & 0xFFFF0000u) : (w << 16));
  }
  __device__ __forceinline__ void set(const float (&v)[8]) {
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  }
};

// bf16 fast paths of the two max-pool passes below (C % 8 == 0): one thread per pooled pixel and 8 channels
__device__ __forceinline__ void pool_window8(const bf16* x, size_t base, size_t C, size_t WC, int (&k)[8]) {
  Bf8 w[4];
  w[0].load(x + base); w[1].load(x + base + C); w[2].load(x + base + WC); w[3].load(x + base + WC + C);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float m = w[0].get(j);
    int kk = 0;
#pragma unroll
    for (int q = 1; q < 4; ++q) {
      const float v = w[q].get(j);
      if (v > m) { m = v; kk = q; }
    }
    k[j] = kk;
  }
}
__global__ void maxpool_bwd_vec_kernel(const bf16* dy, const bf16* x, const bf16* add_src, bf16* dx, int N, int H, int W,
                                       int C) {
  const int Ho = H / 2, Wo = W / 2, cv = C / 8;
  const size_t total = (size_t)N * Ho * Wo * cv;
  const size_t WC = (size_t)W * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % cv) * 8;
    const size_t p = i / cv;
    const int wo = (int)(p % Wo), ho = (int)((p / Wo) % Ho);
    const size_t n = p / ((size_t)Wo * Ho);
    const size_t base = ((n * H + 2 * ho) * W + 2 * wo) * C + c0;
    int k[8];
    pool_window8(x, base, C, WC, k);
    Bf8 g;
    g.load(dy + i * 8);
    const size_t offs[4] = {base, base + (size_t)C, base + WC, base + WC + C};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = k[j] == q ? g.get(j) : 0.f;
      if (add_src) {
        Bf8 ad;
        ad.load(add_src + offs[q]);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += ad.get(j);
      }
      Bf8 o;
      o.set(v);
      o.store(dx + offs[q]);
    }
  }
}
__global__ void maxpool_select_vec_kernel(const bf16* v, const bf16* x, bf16* out, int N, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2, cv = C / 8;
  const size_t total = (size_t)N * Ho * Wo * cv;
  const size_t WC = (size_t)W * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % cv) * 8;
    const size_t p = i / cv;
    const int wo = (int)(p % Wo), ho = (int)((p / Wo) % Ho);
    const size_t n = p / ((size_t)Wo * Ho);
    const size_t base = ((n * H + 2 * ho) * W + 2 * wo) * C + c0;
    int k[8];
    pool_window8(x, base, C, WC, k);
    Bf8 w[4];
    w[0].load(v + base); w[1].load(v + base + C); w[2].load(v + base + WC); w[3].load(v + base + WC + C);
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = k[j] == 0 ? w[0].get(j) : k[j] == 1 ? w[1].get(j) : k[j] == 2 ? w[2].get(j) : w[3].get(j);
    Bf8 o;
    o.set(r);
    o.store(out + i * 8);
  }
}

// dx (N,H,W,C) = route dy (N,H/2,W/2,C) to the first argmax of x in each window; optional add_src accumulates
// a second gradient arriving at the same tensor (skip connections of the generator).
template <typename T>
__global__ void maxpool_bwd_kernel(const T* dy, const T* x, const T* add_src, T* dx, int N, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2;
  const size_t total = (size_t)N * Ho * Wo * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = i % C;
    const size_t p = i / C;
    const int wo = p % Wo, ho = (p / Wo) % Ho;
    const size_t n = p / ((size_t)Wo * Ho);
    const size_t base = ((n * H + 2 * ho) * W + 2 * wo) * C + c;
    const size_t WC = (size_t)W * C;
    const int k = first_argmax(x + base, C, WC);
    const float g = ldf(dy + i);
    const size_t offs[4] = {base, base + C, base + WC, base + WC + C};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float v = q == k ? g : 0.f;
      if (add_src) v += ldf(add_src + offs[q]);
      stf(dx + offs[q], v);
    }
  }
}

// JVP through max-pool: v_pooled = v at the first argmax of x.
template <typename T>
__global__ void maxpool_select_kernel(const T* v, const T* x, T* out, int N, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2;
  const size_t total = (size_t)N * Ho * Wo * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = i % C;
    const size_t p = i / C;
    const int wo = p % Wo, ho = (p / Wo) % Ho;
    const size_t n = p / ((size_t)Wo * Ho);
    const size_t base = ((n * H + 2 * ho) * W + 2 * wo) * C + c;
    const size_t WC = (size_t)W * C;
    const int k = first_argmax(x + base, C, WC);
    const size_t off = base + (k & 1) * C + (k >> 1) * WC;
    stf(out + i, ldf(v + off));
  }
}

// out[c] += alpha * sum_rows src[row, c]   (rows x C row-major).  CTA = 32 channels x 8 row lanes.
// bf16 fast path of channel_sum_kernel (C % 8 == 0, C/8 divides 256): 16-byte loads, a warp reads 512 contiguous bytes
__global__ void __launch_bounds__(256) channel_sum_vec_kernel(const bf16* src, long long rows, int C, float* out,
                                                              float alpha, long long rows_per_cta) {
  extern __shared__ float s_cs[];  // [C]
  for (int i = threadIdx.x; i < C; i += blockDim.x) s_cs[i] = 0.f;
  __syncthreads();
  const int cv = C / 8;
  const int cg = threadIdx.x % cv, ro = threadIdx.x / cv, rstep = blockDim.x / cv;
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  const long long r1 = min(rows, r0 + rows_per_cta);
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  long long r = r0 + ro;
  for (; r + rstep < r1; r += 2 * rstep) {  // two independent loads in flight
    Bf8 x0, x1;
    x0.load(src + (r * cv + cg) * 8);
    x1.load(src + ((r + rstep) * cv + cg) * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += x0.get(k) + x1.get(k);
  }
  if (r < r1) {
    Bf8 x0;
    x0.load(src + (r * cv + cg) * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += x0.get(k);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) atomicAdd(&s_cs[cg * 8 + k], acc[k]);
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(out + i, alpha * s_cs[i]);
}

template <typename T>
__global__ void __launch_bounds__(256) channel_sum_kernel(const T* src, long long rows, int C, float* out, float alpha,
                                                          long long rows_per_cta) {
  __shared__ float s[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + cx;
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  const long long r1 = min(rows, r0 + rows_per_cta);
  float acc = 0.f;
  if (c < C)
    for (long long r = r0 + ry; r < r1; r += 8) acc += ldf(src + r * C + c);
  s[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s[i][cx];
    atomicAdd(out + c, alpha * t);
  }
}

// Parameter gradients of y = s*conv(x, W) + t with s = gamma*inv_std, t = beta + (bias - mean)*s, from the raw
// weight gradient G = sum_p x (x) dy (layout [K][C]) and sum_dy[c]:
//   dgamma = inv_std*(sum_k W[k,c] G[k,c] + (bias-mean)[c] sum_dy[c]); dbeta = sum_dy; dbias = s*sum_dy; dW = s*G.
// trans = 0: W and dW are [K][C] (Conv2D HWIO / Dense), G may alias dW.  trans = 1 (Conv2DTranspose): W and dW
// are Keras [4][C][Cin], G is [Cin][4*C] and sum_dy has 4*C entries (one per output parity).
__global__ void __launch_bounds__(128) param_grads_kernel(const float* W, const float* G, float* dW, const float* scale,
                                                          const float* inv_std, const float* bias, const float* mean,
                                                          const float* sum_dy, float* dgamma, float* dbeta,
                                                          float* dbias, int K, int C, int trans, int Cin) {
  const int c = blockIdx.x;
  __shared__ float red[4];
  float dot = 0.f;
  const float s = scale ? scale[c] : 1.f;
  if (!trans) {
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      const float g = G[(size_t)k * C + c];
      dot = fmaf(W[(size_t)k * C + c], g, dot);
      dW[(size_t)k * C + c] = s * g;
    }
  } else {
    for (int k = threadIdx.x; k < 4 * Cin; k += blockDim.x) {
      const int ab = k / Cin, ci = k % Cin;
      const float g = G[(size_t)ci * (4 * C) + ab * C + c];
      const size_t wi = ((size_t)ab * C + c) * Cin + ci;
      dot = fmaf(W[wi], g, dot);
      dW[wi] = s * g;
    }
  }
  dot = warp_sum(dot);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
  __syncthreads();
  if (threadIdx.x == 0) {
    dot = red[0] + red[1] + red[2] + red[3];
    float sd = trans ? sum_dy[c] + sum_dy[C + c] + sum_dy[2 * C + c] + sum_dy[3 * C + c] : sum_dy[c];
    if (dbias) dbias[c] = s * sd;
    if (dgamma) {
      dgamma[c] = inv_std[c] * (dot + (bias[c] - mean[c]) * sd);
      dbeta[c] = sd;
    }
  }
}

// Critic tail backward (dis_9 1x1 -> Flatten -> Dense(1), TG:339-342).  One CTA per sample row s:
//   dh[s,p,c] = go[s]*wd[p]*w9[c] masked by h>0            (gradient wrt conv2d_dis_8 pre-activation)
//   parameter gradients use `src` = h for rows < n_reg (weight go[s]) and src = v (JVP activations, weight 1)
//   for rows >= n_reg, which carry the gradient-penalty term.
template <typename T>
__global__ void __launch_bounds__(256) critic_head_bwd_kernel(const T* h, const T* v, const float* go, const float* w9,
                                                              const float* b9, const float* wd, T* dh, float* d_w9,
                                                              float* d_b9, float* d_wd, float* d_bd, int n_reg, int HW,
                                                              int C, int want_param_grads) {
  const int s = blockIdx.x;
  const float g = go[s];
  const bool reg = s < n_reg;
  const T* hs = h + (size_t)s * HW * C;
  const T* src = reg ? hs : (v ? v + (size_t)(s - n_reg) * HW * C : nullptr);
  const float wgt = reg ? g : 1.f;
  if (dh) {
    T* o = dh + (size_t)s * HW * C;
    for (int i = threadIdx.x; i < HW * C; i += blockDim.x) {
      const int p = i / C, c = i % C;
      stf(o + i, ldf(hs + i) > 0.f ? g * wd[p] * w9[c] : 0.f);
    }
  }
  if (!want_param_grads || !src) return;
  // d_wd[p] += wgt * (sum_c w9[c] src[p,c] + (reg ? b9 : 0));  d_w9[c] += wgt * sum_p wd[p] src[p,c]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int p = warp; p < HW; p += 8) {
    float d = 0.f;
    for (int c = lane; c < C; c += 32) d = fmaf(ldf(src + (size_t)p * C + c), w9[c], d);
    d = warp_sum(d);
    if (lane == 0) atomicAdd(d_wd + p, wgt * (d + (reg ? b9[0] : 0.f)));
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float d = 0.f;
    for (int p = 0; p < HW; ++p) d = fmaf(ldf(src + (size_t)p * C + c), wd[p], d);
    atomicAdd(d_w9 + c, wgt * d);
  }
  if (reg && threadIdx.x == 0) {
    float swd = 0.f;
    for (int p = 0; p < HW; ++p) swd += wd[p];
    atomicAdd(d_b9, g * swd);
    atomicAdd(d_bd, g);
  }
}


// bf16 fast path of critic_head_bwd_kernel (C % 8 == 0): 16-byte accesses, grid = (rows, SPLIT) over the pixels.
__global__ void __launch_bounds__(256) critic_head_bwd_vec_kernel(const bf16* h, const bf16* v, const float* go,
                                                                  const float* w9, const float* b9, const float* wd,
                                                                  bf16* dh, float* d_w9, float* d_b9, float* d_wd,
                                                                  float* d_bd, int n_reg, int HW, int C,
                                                                  int want_param_grads) {
  extern __shared__ float s_w9acc[];  // [C] partial d_w9 of this CTA
  const int s = blockIdx.x;
  const float g = go[s];
  const bool reg = s < n_reg;
  const bf16* hs = h + (size_t)s * HW * C;
  const bf16* src = reg ? hs : (v ? v + (size_t)(s - n_reg) * HW * C : nullptr);
  const float wgt = reg ? g : 1.f;
  const int cv = C / 8;                       // vectors per pixel
  const int ppc = (HW + gridDim.y - 1) / gridDim.y;
  const int p0 = blockIdx.y * ppc, p1 = min(HW, p0 + ppc);
  if (dh) {
    bf16* o = dh + (size_t)s * HW * C;
    for (int i = p0 * cv + threadIdx.x; i < p1 * cv; i += blockDim.x) {
      const int p = i / cv, c0 = (i - p * cv) * 8;
      Bf8 hv, ov;
      hv.load(hs + (size_t)i * 8);
      const float gw = g * wd[p];
      float r[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) r[k] = hv.get(k) > 0.f ? gw * w9[c0 + k] : 0.f;
      ov.set(r);
      ov.store(o + (size_t)i * 8);
    }
  }
  if (!want_param_grads || !src) return;
  for (int c = threadIdx.x; c < C; c += blockDim.x) s_w9acc[c] = 0.f;
  __syncthreads();
  // warp per pixel: d_wd[p] += wgt * (sum_c w9[c] src[p,c] + (reg ? b9 : 0)); lanes keep d_w9 partials of their channels
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int cb0 = 0; cb0 < cv; cb0 += 32) {
    const int cb = cb0 + lane;
    const bool act = cb < cv;
    float w8[8], acc9[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { w8[k] = act ? w9[cb * 8 + k] : 0.f; acc9[k] = 0.f; }
    for (int p = p0 + warp; p < p1; p += 8) {
      float d = 0.f;
      if (act) {
        Bf8 x;
        x.load(src + ((size_t)p * cv + cb) * 8);
        const float wp = wd[p];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float xv = x.get(k);
          d = fmaf(xv, w8[k], d);
          acc9[k] = fmaf(xv, wp, acc9[k]);
        }
      }
      d = warp_sum(d);
      if (lane == 0) atomicAdd(d_wd + p, wgt * (d + ((reg && cb0 == 0) ? b9[0] : 0.f)));
    }
    if (act) {
#pragma unroll
      for (int k = 0; k < 8; ++k) atomicAdd(&s_w9acc[cb * 8 + k], acc9[k]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(d_w9 + c, wgt * s_w9acc[c]);
  if (reg && blockIdx.y == 0 && threadIdx.x == 0) {
    float swd = 0.f;
    for (int p = 0; p < HW; ++p) swd += wd[p];
    atomicAdd(d_b9, g * swd);
    atomicAdd(d_bd, g);
  }
}

// Gradient penalty pieces (TG:543-545): per sample norm = sqrt(sum g^2); gp_partial += (norm-1)^2 / global_n;
// u = delta * (2/global_n) * (norm-1)/norm * g  (the input of the JVP pass).  One CTA per sample.
__global__ void __launch_bounds__(1024) gp_kernel(const float* g, float* u, float* gp_out, long long hw, float delta,
                                                  float inv_n) {
  __shared__ double red[32];
  __shared__ float coef;
  const float* gs = g + (size_t)blockIdx.x * hw;
  float* us = u + (size_t)blockIdx.x * hw;
  const bool vec = (hw & 3) == 0;  // 16-byte rows (hw = H*W is a multiple of 256 for the networks)
  double acc = 0.0;
  if (vec) {
    const float4* g4 = reinterpret_cast<const float4*>(gs);
    for (long long i = threadIdx.x; i < hw / 4; i += blockDim.x) {
      const float4 v = __ldg(g4 + i);
      acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
  } else {
    for (long long i = threadIdx.x; i < hw; i += blockDim.x) acc += (double)gs[i] * gs[i];
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    const float norm = (float)sqrt(t);
    atomicAdd(gp_out, (norm - 1.f) * (norm - 1.f) * inv_n);
    coef = norm > 0.f ? delta * 2.f * inv_n * (norm - 1.f) / norm : 0.f;
  }
  __syncthreads();
  if (vec) {
    const float4* g4 = reinterpret_cast<const float4*>(gs);
    float4* u4 = reinterpret_cast<float4*>(us);
    const float c = coef;
    for (long long i = threadIdx.x; i < hw / 4; i += blockDim.x) {
      const float4 v = __ldg(g4 + i);
      u4[i] = make_float4(c * v.x, c * v.y, c * v.z, c * v.w);
    }
  } else {
    for (long long i = threadIdx.x; i < hw; i += blockDim.x) us[i] = coef * gs[i];
  }
}

// out[k] += alpha * sum_i src[i] for k-th segment of `seg` elements (critic score means)
__global__ void segment_sum_kernel(const float* src, int seg, int nseg, float* out, float alpha) {
  const int k = blockIdx.x;
  if (k >= nseg) return;
  float acc = 0.f;
  for (int i = threadIdx.x; i < seg; i += blockDim.x) acc += src[(size_t)k * seg + i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) atomicAdd(out + k, alpha * acc);
}

// Generator loss partial sums (TG:576-589): sums[2] += sum|dem - (real2 - base)|, sums[3] += #(real2 >= thr),
// sums[4] += #(base+dem >= thr), sums[5] += #(both); optionally fake2 = base + dem (the Y2 critic's input) and
// the L1 gradient l1g = coef * sign(dem - real_dem).
__global__ void __launch_bounds__(256) gen_loss_sums_kernel(const float* dem, const float* x1, int nicg,
                                                            const float* real2, float thr, float* fake2, float* l1g,
                                                            float l1coef, double* sums, long long n) {
  double a = 0.0;
  unsigned long long cr = 0, cf = 0, cb = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float base = x1[i * nicg], d = dem[i], r2 = real2[i];
    const float f2 = base + d;
    const float diff = d - (r2 - base);
    a += (double)fabsf(diff);
    const bool wr = r2 >= thr, wf = f2 >= thr;
    cr += wr; cf += wf; cb += (wr && wf);
    if (fake2) fake2[i] = f2;
    if (l1g) l1g[i] = diff > 0.f ? l1coef : (diff < 0.f ? -l1coef : 0.f);
  }
  a = warp_sum(a);
  double dr = warp_sum((double)cr), df = warp_sum((double)cf), db = warp_sum((double)cb);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(sums + 2, a);
    atomicAdd(sums + 3, dr);
    atomicAdd(sums + 4, df);
    atomicAdd(sums + 5, db);
  }
}

__global__ void gen_loss_finalize_kernel(float* out6, const double* s) {
  // s: [0] sum D_y2(fake2), [1] sum D_dem(dem), [2] sum |dem-realdem|, [3] sum wmh_real, [4] sum wmh_fake,
  //    [5] sum wmh_real*wmh_fake, [6] global batch, [7] pixels per slice.  One block per candidate (batched evaluation).
  out6 += 6 * blockIdx.x;
  s += 8 * blockIdx.x;
  const double n = s[6], hw = s[7];
  const double lf = s[0] / n, lfd = s[1] / n;
  const double m1 = 100.0 * s[2] / (n * hw);
  const double dice = (2.0 * s[5] + 1e-7) / (s[3] + s[4] + 1e-7);
  const double m4 = 1.0 * (1.0 - dice);
  const double dv = s[3] / 1000.0 - s[4] / 1000.0;
  const double m3 = 100.0 * dv * dv;
  out6[0] = (float)(-lf - lfd + m1 + m3 + m4);
  out6[1] = (float)lf; out6[2] = (float)lfd; out6[3] = (float)m1; out6[4] = (float)m3; out6[5] = (float)m4;
}

// d_seg = (gy2 + gdem + l1g) * (1 - dem^2)  (tanh'), then the 1x1 head backward:
//   d_o[p,c] = d_seg[p] * w[c] masked by o>0;  d_w[c] += sum_p o[p,c] d_seg[p];  d_b += sum_p d_seg[p]
template <typename T>
__global__ void __launch_bounds__(256) gen_head_bwd_kernel(const float* gy2, const float* gdem, const float* l1g,
                                                           const float* dem, const T* o, const float* w, T* d_o,
                                                           float* d_w, float* d_b, long long npix, int C) {
  extern __shared__ float s_dw[];  // [C] + 1
  for (int i = threadIdx.x; i <= C; i += blockDim.x) s_dw[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float accw[8];  // lane handles channels lane, lane+32, ... (C <= 256)
#pragma unroll
  for (int k = 0; k < 8; ++k) accw[k] = 0.f;
  float accb = 0.f;
  for (long long p = warp; p < npix; p += nwarps) {
    const float d = dem[p];
    const float ds = (gy2[p] + gdem[p] + l1g[p]) * (1.f - d * d);
    if (lane == 0) accb += ds;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = lane + 32 * k;
      if (c < C) {
        const float ov = ldf(o + p * C + c);
        accw[k] = fmaf(ov, ds, accw[k]);
        stf(d_o + p * C + c, ov > 0.f ? ds * w[c] : 0.f);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (lane + 32 * k < C) atomicAdd(&s_dw[lane + 32 * k], accw[k]);
  if (lane == 0) atomicAdd(&s_dw[C], accb);
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(d_w + i, s_dw[i]);
  if (threadIdx.x == 0) atomicAdd(d_b, s_dw[C]);
}

// ResBlock FiLM backward (TG:403-407): r = relu(y*g + b) + a.  Given d_r:
//   d_t = d_r * (y*g + b > 0);  d_y = d_t * g;  dgam[n,c] += sum_hw d_t*y;  dbet[n,c] += sum_hw d_t
// CTA = (pixel chunk, n); thread c-lane loops pixels.
template <typename T>
__global__ void __launch_bounds__(256) film_bwd_kernel(const T* d_r, const T* y, const float* fg, const float* fb,
                                                       int fstride, T* d_y, float* dgam, float* dbet, int HW, int C,
                                                       int pix_per_cta) {
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
  const int cl = threadIdx.x % C, pl = threadIdx.x / C, pstep = blockDim.x / C;
  if (pl >= pstep) return;
  const float g = fg[(size_t)n * fstride + cl], b = fb[(size_t)n * fstride + cl];
  float ag = 0.f, ab = 0.f;
  for (int p = p0 + pl; p < p1; p += pstep) {
    const size_t i = ((size_t)n * HW + p) * C + cl;
    const float yv = ldf(y + i);
    const float dt = fmaf(yv, g, b) > 0.f ? ldf(d_r + i) : 0.f;
    stf(d_y + i, dt * g);
    ag = fmaf(dt, yv, ag);
    ab += dt;
  }
  atomicAdd(dgam + (size_t)n * fstride + cl, ag);
  atomicAdd(dbet + (size_t)n * fstride + cl, ab);
}

// bf16 fast path of film_bwd_kernel (C % 8 == 0, C/8 divides 256): thread = 8 channels, 16-byte accesses
__global__ void __launch_bounds__(256) film_bwd_vec_kernel(const bf16* d_r, const bf16* y, const float* fg,
                                                           const float* fb, int fstride, bf16* d_y, float* dgam,
                                                           float* dbet, int HW, int C, int pix_per_cta) {
  extern __shared__ float s_fb[];  // [2][C] partial (dgamma, dbeta) of this CTA
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_fb[i] = 0.f;
  __syncthreads();
  const int n = blockIdx.y, cv = C / 8;
  const int p0 = blockIdx.x * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
  const int cg = threadIdx.x % cv, pl = threadIdx.x / cv, pstep = blockDim.x / cv;
  float g[8], b[8], ag[8], ab[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    g[k] = fg[(size_t)n * fstride + cg * 8 + k];
    b[k] = fb[(size_t)n * fstride + cg * 8 + k];
    ag[k] = ab[k] = 0.f;
  }
  for (int p = p0 + pl; p < p1; p += pstep) {
    const size_t i = (((size_t)n * HW + p) * cv + cg) * 8;
    Bf8 yv, dv, ov;
    yv.load(y + i);
    dv.load(d_r + i);
    float r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float yk = yv.get(k);
      const float dt = fmaf(yk, g[k], b[k]) > 0.f ? dv.get(k) : 0.f;
      r[k] = dt * g[k];
      ag[k] = fmaf(dt, yk, ag[k]);
      ab[k] += dt;
    }
    ov.set(r);
    ov.store(d_y + i);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    atomicAdd(&s_fb[cg * 8 + k], ag[k]);
    atomicAdd(&s_fb[C + cg * 8 + k], ab[k]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(dgam + (size_t)n * fstride + c, s_fb[c]);
    atomicAdd(dbet + (size_t)n * fstride + c, s_fb[C + c]);
  }
}

// bf16 fast paths of add_mask_kernel / slice_add_kernel: 8 elements per thread
__global__ void add_mask_vec_kernel(const bf16* a, const bf16* b, const bf16* m, bf16* out, long long n8) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    Bf8 av, o;
    av.load(a + i * 8);
    float r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = av.get(k);
    if (b) {
      Bf8 bv;
      bv.load(b + i * 8);
#pragma unroll
      for (int k = 0; k < 8; ++k) r[k] += bv.get(k);
    }
    if (m) {
      Bf8 mv;
      mv.load(m + i * 8);
#pragma unroll
      for (int k = 0; k < 8; ++k) r[k] = mv.get(k) > 0.f ? r[k] : 0.f;
    }
    o.set(r);
    o.store(out + i * 8);
  }
}
__global__ void slice_add_vec_kernel(const bf16* src, int sstride, int off, const bf16* add, bf16* dst, long long rows,
                                     int C) {
  const int cv = C / 8;
  const long long total = rows * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cv;
    const int c0 = (int)(i - r * cv) * 8;
    Bf8 sv, o;
    sv.load(src + r * sstride + off + c0);
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = sv.get(k);
    if (add) {
      Bf8 ad;
      ad.load(add + i * 8);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += ad.get(k);
    }
    o.set(v);
    o.store(dst + i * 8);
  }
}

// out = a (+ b) masked by (m > 0) -- used to merge gradient branches and apply ReLU masks
template <typename T>
__global__ void add_mask_kernel(const T* a, const T* b, const T* m, T* out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = ldf(a + i);
    if (b) v += ldf(b + i);
    if (m) v = ldf(m + i) > 0.f ? v : 0.f;
    stf(out + i, v);
  }
}

// Space-to-depth of the transposed-conv output gradient with its ReLU mask:
//   out[n,h,w,(a*2+b)*C + c] = (up[n,2h+a,2w+b,c] > 0) ? d_up[n,2h+a,2w+b, c] : 0
// d_up may be a channel slice of a wider tensor (stride dstride, offset 0).
template <typename T>
__global__ void s2d_mask_kernel(const T* d_up, int dstride, const T* up, T* out, int N, int H, int W, int C) {
  const size_t total = (size_t)N * H * W * 4 * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = i % C;
    const int ab = (i / C) % 4;
    const size_t p = i / (4 * (size_t)C);
    const int w = p % W, h = (p / W) % H;
    const size_t n = p / ((size_t)W * H);
    const size_t op = (n * 2 * H + 2 * h + (ab >> 1)) * (2 * (size_t)W) + 2 * w + (ab & 1);
    const float m = up ? ldf(up + op * C + c) : 1.f;
    stf(out + i, m > 0.f ? ldf(d_up + op * dstride + c) : 0.f);
  }
}

// bf16 fast path (C, dstride multiples of 8): one thread per 8 channels, 16-byte accesses
__global__ void s2d_mask_vec_kernel(const bf16* d_up, int dstride, const bf16* up, bf16* out, int N, int H, int W, int C) {
  const int cv = C / 8;
  const size_t total = (size_t)N * H * W * 4 * cv;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % cv) * 8;
    const int ab = (int)((i / cv) % 4);
    const size_t p = i / (4 * (size_t)cv);
    const int w = (int)(p % W), h = (int)((p / W) % H);
    const size_t n = p / ((size_t)W * H);
    const size_t op = (n * 2 * H + 2 * h + (ab >> 1)) * (2 * (size_t)W) + 2 * w + (ab & 1);
    Bf8 d, o;
    d.load(d_up + op * dstride + c0);
    if (up) {
      Bf8 m;
      m.load(up + op * C + c0);
      float r[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) r[k] = m.get(k) > 0.f ? d.get(k) : 0.f;
      o.set(r);
    } else {
      o = d;
    }
    o.store(out + i * 8);
  }
}

// bf16 fast path of gen_head_bwd_kernel (C % 8 == 0, C <= 256): C/8 lanes share a pixel, 16-byte accesses
__global__ void __launch_bounds__(256) gen_head_bwd_vec_kernel(const float* gy2, const float* gdem, const float* l1g,
                                                               const float* dem, const bf16* o, const float* w, bf16* d_o,
                                                               float* d_w, float* d_b, long long npix, int C) {
  extern __shared__ float s_dw[];  // [C] + 1
  for (int i = threadIdx.x; i <= C; i += blockDim.x) s_dw[i] = 0.f;
  __syncthreads();
  const int cv = C / 8;                 // lanes per pixel (power of two <= 32)
  const int cl = threadIdx.x % cv;      // this thread's channel group
  const long long slot = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / cv;
  const long long nslots = ((long long)gridDim.x * blockDim.x) / cv;
  float w8[8], accw[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { w8[k] = w[cl * 8 + k]; accw[k] = 0.f; }
  float accb = 0.f;
  for (long long p = slot; p < npix; p += nslots) {
    const float d = dem[p];
    const float ds = (gy2[p] + gdem[p] + l1g[p]) * (1.f - d * d);
    if (cl == 0) accb += ds;
    Bf8 ov, dv;
    ov.load(o + p * C + cl * 8);
    float r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float x = ov.get(k);
      accw[k] = fmaf(x, ds, accw[k]);
      r[k] = x > 0.f ? ds * w8[k] : 0.f;
    }
    dv.set(r);
    dv.store(d_o + p * C + cl * 8);
  }
  // lanes with the same channel group sit cv apart: fold them inside the warp first
  for (int off = 16; off >= cv; off >>= 1) {
#pragma unroll
    for (int k = 0; k < 8; ++k) accw[k] += __shfl_down_sync(0xffffffffu, accw[k], off);
    accb += __shfl_down_sync(0xffffffffu, accb, off);
  }
  if ((threadIdx.x & 31) < cv) {
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(&s_dw[cl * 8 + k], accw[k]);
    if (cl == 0) atomicAdd(&s_dw[C], accb);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(d_w + i, s_dw[i]);
  if (threadIdx.x == 0) atomicAdd(d_b, s_dw[C]);
}

// dst[row, 0:C] = src[row, off:off+C] (+ add[row, 0:C])   -- channel slice of the concat gradient
template <typename T>
__global__ void slice_add_kernel(const T* src, int sstride, int off, const T* add, T* dst, long long rows, int C) {
  const long long total = rows * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = i % C;
    float v = ldf(src + r * sstride + off + c);
    if (add) v += ldf(add + i);
    stf(dst + i, v);
  }
}

// ---- FiLM noise-MLP backward (TG:353-395) ----
// stage B backward: d_out (N, total) -> per head: dWraw[k,c] = sum_n h2[n,k] d_out[n,off+c] (fp32, [K][C]),
// sum_d[off+c] = sum_n d_out;  d_h2[n,k] = sum_heads sum_c d_out[n,off+c]*s[c]*W[k,c]
__global__ void __launch_bounds__(256) film_heads_wgrad_kernel(FilmMlpArgs a, const float* d_out, float* const* dW,
                                                               float* sum_d) {
  // grid: (total_c/32, K/32); CTA computes a 32(k) x 32(c) block of one head's dWraw over all n
  const int K = a.L * a.F;
  const int col0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  int h = 0;
  for (int i = 0; i < a.n_heads; ++i)
    if (col0 >= a.head_off[i]) h = i;
  const int C = a.head_c[h], cl0 = col0 - a.head_off[h];
  const int cx = threadIdx.x & 31, ky = threadIdx.x >> 5;  // 8 k-rows per pass
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float sd = 0.f;
  for (int n = 0; n < a.N; ++n) {
    const float d = d_out[(size_t)n * a.total_c + col0 + cx];
    sd += d;
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] = fmaf(a.h2[(size_t)n * K + k0 + ky + 8 * q], d, acc[q]);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) dW[h][(size_t)(k0 + ky + 8 * q) * C + cl0 + cx] = acc[q];
  if (blockIdx.y == 0 && ky == 0) sum_d[col0 + cx] = sd;
}

__global__ void __launch_bounds__(256) film_heads_dgrad_kernel(FilmMlpArgs a, const float* d_out, float* d_h2) {
  // grid: (K/256, N); thread = one k; loops all columns of all heads.  d_h2 masked by h2 > 0 (ReLU of f1).
  const int K = a.L * a.F;
  const int k = blockIdx.x * 256 + threadIdx.x, n = blockIdx.y;
  extern __shared__ float s_d[];  // [total_c] = d_out[n,:] * scale
  for (int i = threadIdx.x; i < a.total_c; i += blockDim.x) {
    int h = 0;
    for (int j = 0; j < a.n_heads; ++j)
      if (i >= a.head_off[j]) h = j;
    s_d[i] = d_out[(size_t)n * a.total_c + i] * a.head_s[h][i - a.head_off[h]];
  }
  __syncthreads();
  if (k >= K) return;
  float acc = 0.f;
  for (int h = 0; h < a.n_heads; ++h) {
    const int C = a.head_c[h], off = a.head_off[h];
    const float* Wr = a.head_w[h] + (size_t)k * C;
    for (int c = 0; c < C; ++c) acc = fmaf(s_d[off + c], Wr[c], acc);
  }
  d_h2[(size_t)n * K + k] = a.h2[(size_t)n * K + k] > 0.f ? acc : 0.f;
}

// stage A backward.  d_h2 (N,L,F) is already ReLU-masked (gradient wrt BN output of f1).
//   f1: pre2 = s1*(h1 @ k1) + t1 ;  f0: h1 = relu(s0*(z*k0) + t0)
// One CTA per sample; results accumulated with atomics into dk1raw (F,F), sum_d1 (F), dk0raw (F), sum_d0 (F).
__global__ void film_stage_a_bwd_kernel(FilmMlpArgs a, const float* d_h2, float* dk1raw, float* sum_d1, float* dk0raw,
                                        float* sum_d0) {
  extern __shared__ float sm[];  // d2 [L*F], h1 [L*F], d1 [L*F]
  const int L = a.L, Fd = a.F, n = blockIdx.x, LF = L * Fd;
  float* d2 = sm;
  float* h1 = sm + LF;
  float* d1 = sm + 2 * LF;
  for (int i = threadIdx.x; i < LF; i += blockDim.x) {
    d2[i] = d_h2[(size_t)n * LF + i];
    h1[i] = a.h1[(size_t)n * LF + i];
  }
  __syncthreads();
  // dk1raw[k,f] += sum_l h1[l,k] * d2[l,f];   sum_d1[f] += sum_l d2[l,f]
  for (int i = threadIdx.x; i < Fd * Fd; i += blockDim.x) {
    const int k = i / Fd, f = i % Fd;
    float acc = 0.f;
    for (int l = 0; l < L; ++l) acc = fmaf(h1[l * Fd + k], d2[l * Fd + f], acc);
    atomicAdd(dk1raw + i, acc);
  }
  for (int f = threadIdx.x; f < Fd; f += blockDim.x) {
    float acc = 0.f;
    for (int l = 0; l < L; ++l) acc += d2[l * Fd + f];
    atomicAdd(sum_d1 + f, acc);
  }
  // d_h1[l,k] = sum_f d2[l,f]*s1[f]*k1[k,f], masked by h1 > 0
  for (int i = threadIdx.x; i < LF; i += blockDim.x) {
    const int l = i / Fd, k = i % Fd;
    float acc = 0.f;
    for (int f = 0; f < Fd; ++f) acc = fmaf(d2[l * Fd + f] * a.s1[f], a.k1[k * Fd + f], acc);
    d1[i] = h1[i] > 0.f ? acc : 0.f;
  }
  __syncthreads();
  // dk0raw[f] += sum_l z[l]*d1[l,f];  sum_d0[f] += sum_l d1[l,f]
  for (int f = threadIdx.x; f < Fd; f += blockDim.x) {
    float acc = 0.f, s = 0.f;
    for (int l = 0; l < L; ++l) {
      acc = fmaf(a.z[(size_t)n * L + l], d1[l * Fd + f], acc);
      s += d1[l * Fd + f];
    }
    atomicAdd(dk0raw + f, acc);
    atomicAdd(sum_d0 + f, s);
  }
}

}  // namespace

#define DISPATCH_DT(dt, CALL_F32, CALL_BF16) \
  do {                                       \
    if ((dt) == DT_F32) { CALL_F32; } else { CALL_BF16; } \
  } while (0)

int k_maxpool_bwd(const void* dy, const void* x, const void* add_src, void* dx, int N, int H, int W, int C, int dt,
                  cudaStream_t st) {
  const long long total = (long long)N * (H / 2) * (W / 2) * C;
  if (total == 0) return 0;
  if (dt == DT_BF16 && C % 8 == 0) {
    maxpool_bwd_vec_kernel<<<grid_for(total / 8), 256, 0, st>>>((const bf16*)dy, (const bf16*)x, (const bf16*)add_src,
                                                                (bf16*)dx, N, H, W, C);
    DG_LAUNCH_CHECK();
    return 0;
  }
  DISPATCH_DT(dt,
              (maxpool_bwd_kernel<float><<<grid_for(total), 256, 0, st>>>((const float*)dy, (const float*)x,
                                                                         (const float*)add_src, (float*)dx, N, H, W, C)),
              (maxpool_bwd_kernel<bf16><<<grid_for(total), 256, 0, st>>>((const bf16*)dy, (const bf16*)x,
                                                                        (const bf16*)add_src, (bf16*)dx, N, H, W, C)));
  DG_LAUNCH_CHECK();
  return 0;
}

int k_maxpool_select(const void* v, const void* x, void* out, int N, int H, int W, int C, int dt, cudaStream_t st) {
  const long long total = (long long)N * (H / 2) * (W / 2) * C;
  if (total == 0) return 0;
  if (dt == DT_BF16 && C % 8 == 0) {
    maxpool_select_vec_kernel<<<grid_for(total / 8), 256, 0, st>>>((const bf16*)v, (const bf16*)x, (bf16*)out, N, H, W, C);
    DG_LAUNCH_CHECK();
    return 0;
  }
  DISPATCH_DT(dt,
              (maxpool_select_kernel<float><<<grid_for(total), 256, 0, st>>>((const float*)v, (const float*)x,
                                                                            (float*)out, N, H, W, C)),
              (maxpool_select_kernel<bf16><<<grid_for(total), 256, 0, st>>>((const bf16*)v, (const bf16*)x, (bf16*)out,
                                                                           N, H, W, C)));
  DG_LAUNCH_CHECK();
  return 0;
}

int k_channel_sum(const void* src, long long rows, int C, float* out, float alpha, int dt, cudaStream_t st) {
  if (rows == 0) return 0;
  if (dt == DT_BF16 && C % 8 == 0 && 256 % (C / 8) == 0) {
    long long gv = (rows + 1023) / 1024;
    if (gv > 148 * 8) gv = 148 * 8;
    const long long rpcv = (rows + gv - 1) / gv;
    channel_sum_vec_kernel<<<(unsigned)((rows + rpcv - 1) / rpcv), 256, C * sizeof(float), st>>>((const bf16*)src, rows, C,
                                                                                                 out, alpha, rpcv);
    DG_LAUNCH_CHECK();
    return 0;
  }
  long long gx = (rows + 2047) / 2048;
  if (gx > 148 * 4) gx = 148 * 4;
  const long long rpc = (rows + gx - 1) / gx;
  dim3 grid((unsigned)((rows + rpc - 1) / rpc), (C + 31) / 32);
  DISPATCH_DT(dt, (channel_sum_kernel<float><<<grid, 256, 0, st>>>((const float*)src, rows, C, out, alpha, rpc)),
              (channel_sum_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)src, rows, C, out, alpha, rpc)));
  DG_LAUNCH_CHECK();
  return 0;
}

int k_param_grads(const float* W, const float* G, float* dW, const float* scale, const float* inv_std,
                  const float* bias, const float* mean, const float* sum_dy, float* dgamma, float* dbeta, float* dbias,
                  int K, int C, int trans, int Cin, cudaStream_t st) {
  param_grads_kernel<<<C, 128, 0, st>>>(W, G, dW, scale, inv_std, bias, mean, sum_dy, dgamma, dbeta, dbias, K, C, trans,
                                        Cin);
  DG_LAUNCH_CHECK();
  return 0;
}

int k_critic_head_bwd(const void* h, const void* v, const float* go, const float* w9, const float* b9, const float* wd,
                      void* dh, float* d_w9, float* d_b9, float* d_wd, float* d_bd, int rows, int n_reg, int HW, int C,
                      int want_param_grads, int dt, cudaStream_t st) {
  if (rows == 0) return 0;
  if (dt == DT_BF16 && C % 8 == 0) {
    const int split = HW >= 64 ? 4 : 1;
    critic_head_bwd_vec_kernel<<<dim3(rows, split), 256, C * sizeof(float), st>>>(
        (const bf16*)h, (const bf16*)v, go, w9, b9, wd, (bf16*)dh, d_w9, d_b9, d_wd, d_bd, n_reg, HW, C,
        want_param_grads);
    DG_LAUNCH_CHECK();
    return 0;
  }
  DISPATCH_DT(dt,
              (critic_head_bwd_kernel<float><<<rows, 256, 0, st>>>((const float*)h, (const float*)v, go, w9, b9, wd,
                                                                  (float*)dh, d_w9, d_b9, d_wd, d_bd, n_reg, HW, C,
                                                                  want_param_grads)),
              (critic_head_bwd_kernel<bf16><<<rows, 256, 0, st>>>((const bf16*)h, (const bf16*)v, go, w9, b9, wd,
                                                                 (bf16*)dh, d_w9, d_b9, d_wd, d_bd, n_reg, HW, C,
                                                                 want_param_grads)));
  DG_LAUNCH_CHECK();
  return 0;
}

int k_gp(const float* g, float* u, float* gp_out, int n, long long hw, float delta, float inv_n, cudaStream_t st) {
  if (n == 0) return 0;
  gp_kernel<<<n, 1024, 0, st>>>(g, u, gp_out, hw, delta, inv_n);
  DG_LAUNCH_CHECK();
  return 0;
}

int k_segment_sum(const float* src, int seg, int nseg, float* out, float alpha, cudaStream_t st) {
  if (seg == 0 || nseg == 0) return 0;
  segment_sum_kernel<<<nseg, 128, 0, st>>>(src, seg, nseg, out, alpha);
  DG_LAUNCH_CHECK();
  return 0;
}

int k_gen_loss_sums(const float* dem, const float* x1, int nicg, const float* real2, float thr, float* fake2,
                    float* l1g, float l1coef, double* sums, long long n, cudaStream_t st) {
  if (n == 0) return 0;
  gen_loss_sums_kernel<<<grid_for(n, 256, 148 * 4), 256, 0, st>>>(dem, x1, nicg, real2, thr, fake2, l1g, l1coef, sums,
                                                                  n);
  DG_LAUNCH_CHECK();
  return 0;
}

int k_gen_loss_finalize(float* out6, const double* sums, cudaStream_t st, int k) {
  gen_loss_finalize_kernel<<<k, 1, 0, st>>>(out6, sums);
  DG_LAUNCH_CHECK();
  return 0;
}

int k_gen_head_bwd(const float* gy2, const float* gdem, const float* l1g, const float* dem, const void* o,
                   const float* w, void* d_o, float* d_w, float* d_b, long long npix, int C, int dt, cudaStream_t st) {
  if (npix == 0) return 0;
  DG_REQUIRE(C <= 256, "gen_head_bwd: C must be <= 256");
  const int grid = grid_for(npix * 32, 256, 148 * 8);
  const size_t smem = (C + 1) * sizeof(float);
  if (dt == DT_BF16 && C % 8 == 0 && (C / 8 == 1 || C / 8 == 2 || C / 8 == 4 || C / 8 == 8 || C / 8 == 16 || C / 8 == 32)) {
    gen_head_bwd_vec_kernel<<<grid_for(npix * (C / 8), 256, 148 * 8), 256, smem, st>>>(
        gy2, gdem, l1g, dem, (const bf16*)o, w, (bf16*)d_o, d_w, d_b, npix, C);
    DG_LAUNCH_CHECK();
    return 0;
  }
  DISPATCH_DT(dt,
              (gen_head_bwd_kernel<float><<<grid, 256, smem, st>>>(gy2, gdem, l1g, dem, (const float*)o, w, (float*)d_o,
                                                                  d_w, d_b, npix, C)),
              (gen_head_bwd_kernel<bf16><<<grid, 256, smem, st>>>(gy2, gdem, l1g, dem, (const bf16*)o, w, (bf16*)d_o,
                                                                 d_w, d_b, npix, C)));
  DG_LAUNCH_CHECK();
  return 0;
}

int k_film_bwd(const void* d_r, const void* y, const float* fg, const float* fb, int fstride, void* d_y, float* dgam,
               float* dbet, int N, int HW, int C, int dt, cudaStream_t st) {
  if (N == 0) return 0;
  DG_REQUIRE(C <= 256 && 256 % 32 == 0, "film_bwd: C must be <= 256");
  int ppc = (HW + 63) / 64;  // <= 64 CTAs per sample
  if (ppc < 64) ppc = HW < 64 ? HW : 64;
  dim3 grid((HW + ppc - 1) / ppc, N);
  if (dt == DT_BF16 && C % 8 == 0 && 256 % (C / 8) == 0) {
    film_bwd_vec_kernel<<<grid, 256, 2 * C * sizeof(float), st>>>((const bf16*)d_r, (const bf16*)y, fg, fb, fstride,
                                                                 (bf16*)d_y, dgam, dbet, HW, C, ppc);
    DG_LAUNCH_CHECK();
    return 0;
  }
  DISPATCH_DT(dt,
              (film_bwd_kernel<float><<<grid, 256, 0, st>>>((const float*)d_r, (const float*)y, fg, fb, fstride,
                                                           (float*)d_y, dgam, dbet, HW, C, ppc)),
              (film_bwd_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)d_r, (const bf16*)y, fg, fb, fstride,
                                                          (bf16*)d_y, dgam, dbet, HW, C, ppc)));
  DG_LAUNCH_CHECK();
  return 0;
}

int k_add_mask(const void* a, const void* b, const void* m, void* out, long long n, int dt, cudaStream_t st) {
  if (n == 0) return 0;
  if (dt == DT_BF16 && n % 8 == 0 && ((uintptr_t)a | (uintptr_t)b | (uintptr_t)m | (uintptr_t)out) % 16 == 0) {
    add_mask_vec_kernel<<<grid_for(n / 8), 256, 0, st>>>((const bf16*)a, (const bf16*)b, (const bf16*)m, (bf16*)out, n / 8);
    DG_LAUNCH_CHECK();
    return 0;
  }
  DISPATCH_DT(dt,
              (add_mask_kernel<float><<<grid_for(n), 256, 0, st>>>((const float*)a, (const float*)b, (const float*)m,
                                                                  (float*)out, n)),
              (add_mask_kernel<bf16><<<grid_for(n), 256, 0, st>>>((const bf16*)a, (const bf16*)b, (const bf16*)m,
                                                                 (bf16*)out, n)));
  DG_LAUNCH_CHECK();
  return 0;
}

int k_s2d_mask(const void* d_up, int dstride, const void* up, void* out, int N, int H, int W, int C, int dt,
               cudaStream_t st) {
  const long long total = (long long)N * H * W * 4 * C;
  if (total == 0) return 0;
  if (dt == DT_BF16 && C % 8 == 0 && dstride % 8 == 0) {
    s2d_mask_vec_kernel<<<grid_for(total / 8), 256, 0, st>>>((const bf16*)d_up, dstride, (const bf16*)up, (bf16*)out, N,
                                                             H, W, C);
    DG_LAUNCH_CHECK();
    return 0;
  }
  DISPATCH_DT(dt,
              (s2d_mask_kernel<float><<<grid_for(total), 256, 0, st>>>((const float*)d_up, dstride, (const float*)up,
                                                                      (float*)out, N, H, W, C)),
              (s2d_mask_kernel<bf16><<<grid_for(total), 256, 0, st>>>((const bf16*)d_up, dstride, (const bf16*)up,
                                                                     (bf16*)out, N, H, W, C)));
  DG_LAUNCH_CHECK();
  return 0;
}

int k_slice_add(const void* src, int sstride, int off, const void* add, void* dst, long long rows, int C, int dt,
                cudaStream_t st) {
  if (rows == 0) return 0;
  if (dt == DT_BF16 && C % 8 == 0 && sstride % 8 == 0 && off % 8 == 0 &&
      ((uintptr_t)src | (uintptr_t)add | (uintptr_t)dst) % 16 == 0) {
    slice_add_vec_kernel<<<grid_for(rows * (C / 8)), 256, 0, st>>>((const bf16*)src, sstride, off, (const bf16*)add,
                                                                   (bf16*)dst, rows, C);
    DG_LAUNCH_CHECK();
    return 0;
  }
  DISPATCH_DT(dt,
              (slice_add_kernel<float><<<grid_for(rows * C), 256, 0, st>>>((const float*)src, sstride, off,
                                                                          (const float*)add, (float*)dst, rows, C)),
              (slice_add_kernel<bf16><<<grid_for(rows * C), 256, 0, st>>>((const bf16*)src, sstride, off,
                                                                         (const bf16*)add, (bf16*)dst, rows, C)));
  DG_LAUNCH_CHECK();
  return 0;
}

int k_film_mlp_bwd(const FilmMlpArgs& a, const float* d_out, float* const* dW_heads, float* sum_d, float* d_h2,
                   float* dk1raw, float* sum_d1, float* dk0raw, float* sum_d0, cudaStream_t st) {
  if (a.N == 0) return 0;
  const int K = a.L * a.F;
  DG_REQUIRE(K % 32 == 0 && a.total_c % 32 == 0, "film mlp bwd: sizes must be multiples of 32");
  dim3 g1(a.total_c / 32, K / 32);
  film_heads_wgrad_kernel<<<g1, 256, 0, st>>>(a, d_out, dW_heads, sum_d);
  DG_LAUNCH_CHECK();
  dim3 g2((K + 255) / 256, a.N);
  film_heads_dgrad_kernel<<<g2, 256, a.total_c * sizeof(float), st>>>(a, d_out, d_h2);
  DG_LAUNCH_CHECK();
  film_stage_a_bwd_kernel<<<a.N, 256, 3 * K * sizeof(float), st>>>(a, d_h2, dk1raw, sum_d1, dk0raw, sum_d0);
  DG_LAUNCH_CHECK();
  return 0;
}

namespace {
__global__ void fill_go_kernel(float* go, int n, float v_real, float v_fake, float v_mixed, int rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows) go[i] = i < n ? v_real : (i < 2 * n ? v_fake : v_mixed);
}
__global__ void critic_loss_finalize_kernel(float* out4, float delta) {
  out4[3] = out4[1] - out4[0] + delta * out4[2];  // loss_fake - loss_real + delta * GP  (TG:547 / 566)
}
}  // namespace

int k_fill_go(float* go, int n, float v_real, float v_fake, float v_mixed, int rows, cudaStream_t st) {
  if (rows == 0) return 0;
  fill_go_kernel<<<(rows + 127) / 128, 128, 0, st>>>(go, n, v_real, v_fake, v_mixed, rows);
  DG_LAUNCH_CHECK();
  return 0;
}
int k_critic_loss_finalize(float* out4, float delta, cudaStream_t st) {
  critic_loss_finalize_kernel<<<1, 1, 0, st>>>(out4, delta);
  DG_LAUNCH_CHECK();
  return 0;
}

namespace {
__global__ void scores_to_sums_kernel(const float* sy2, const float* sdem, int n, double* sums, double gn, double hw) {
  sy2 += (size_t)n * blockIdx.x;  // one block per candidate (batched evaluation): n scores and 8 sums each
  sdem += (size_t)n * blockIdx.x;
  sums += 8 * blockIdx.x;
  double a = 0.0, b = 0.0;
  for (int i = 0; i < n; ++i) { a += (double)sy2[i]; b += (double)sdem[i]; }
  sums[0] = a; sums[1] = b; sums[6] = gn; sums[7] = hw;
}
// Transposed-conv data-gradient operands from the Keras kernel (2,2,Cout,Cin) and the BN scale s[co]:
//   dst_f32 [(ab,co)][ci] = W*s (CUDA-core 1x1 conv, [K][N]);  dst_bf16 [ci][(ab,co)] = W*s (tcgen05 B operand)
__global__ void pack_deconv_dgrad_kernel(const float* src, const float* scale, float* dst_f32, bf16* dst_bf16, int Cin,
                                         int Cout) {
  const size_t total = (size_t)4 * Cout * Cin;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ci = i % Cin;
    const int k = i / Cin;  // (ab, co)
    const float v = src[i] * (scale ? scale[k % Cout] : 1.f);
    if (dst_f32) dst_f32[i] = v;
    if (dst_bf16) dst_bf16[(size_t)ci * (4 * Cout) + k] = __float2bfloat16_rn(v);
  }
}
}  // namespace

__global__ void loss_consts_kernel(double* sums, double gn, double hw) {
  sums[8 * blockIdx.x + 6] = gn;
  sums[8 * blockIdx.x + 7] = hw;
}
int k_loss_consts(double* sums, double gn, double hw, int k, cudaStream_t st) {
  loss_consts_kernel<<<k, 1, 0, st>>>(sums, gn, hw);
  DG_LAUNCH_CHECK();
  return 0;
}
int k_scores_to_sums(const float* sy2, const float* sdem, int n, double* sums, double gn, double hw, cudaStream_t st,
                     int k) {
  scores_to_sums_kernel<<<k, 1, 0, st>>>(sy2, sdem, n, sums, gn, hw);
  DG_LAUNCH_CHECK();
  return 0;
}
int k_pack_deconv_dgrad(const float* src, const float* scale, float* dst_f32, bf16* dst_bf16, int Cin, int Cout,
                        cudaStream_t st) {
  if (!dst_f32 && !dst_bf16) return 0;
  pack_deconv_dgrad_kernel<<<grid_for((long long)4 * Cin * Cout), 256, 0, st>>>(src, scale, dst_f32, dst_bf16, Cin, Cout);
  DG_LAUNCH_CHECK();
  return 0;
}
