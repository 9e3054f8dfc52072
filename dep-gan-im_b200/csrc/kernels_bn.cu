// Training-phase BatchNormalization (batch statistics), Dropout, softmax + categorical cross-entropy and the small
// dense layers of the noise MLP, for the DEP-UResNet supervised step (TU:291-428, 596-606).  All bandwidth- or
// latency-bound; the convolutions around them are the tcgen05 / CUDA-core kernels of conv_tc.cu / conv_simt.cu.
#include "kernels.cuh"

namespace {

constexpr float BN_EPS = 1e-3f;

inline int grid_for(long long n, int block = 256, int cap = 148 * 16) {
  long long g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

// sums[c] += sum x, sums[C + c] += sum x^2 over the CTA's rows (double accumulation)
template <typename T>
__global__ void __launch_bounds__(256) bn_sums_kernel(const T* x, long long rows, int C, double* sums,
                                                      long long rows_per_cta) {
  __shared__ double s1[8][33], s2[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + cx;
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  double a = 0.0, b = 0.0;
  if (c < C)
    for (long long r = r0 + ry; r < r1; r += 8) {
      const float v = ldf(x + r * C + c);
      a += v;
      b += (double)v * v;
    }
  s1[ry][cx] = a;
  s2[ry][cx] = b;
  __syncthreads();
  if (ry == 0 && c < C) {
    double t1 = 0.0, t2 = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { t1 += s1[i][cx]; t2 += s2[i][cx]; }
    atomicAdd(sums + c, t1);
    atomicAdd(sums + C + c, t2);
  }
}

// mean / inv_std from the sums; moving statistics updated in place (momentum 0.99, Bessel-corrected variance)
__global__ void bn_finalize_kernel(const double* sums, double M, int C, float* mean, float* inv_std, float* mov_mean,
                                   float* mov_var, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mu = sums[c] / M;
  double var = sums[C + c] / M - mu * mu;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)mu;
  inv_std[c] = (float)(1.0 / sqrt(var + (double)BN_EPS));
  if (mov_mean) {
    const double unb = M > 1.0 ? var * M / (M - 1.0) : var;
    mov_mean[c] = momentum * mov_mean[c] + (1.f - momentum) * (float)mu;
    mov_var[c] = momentum * mov_var[c] + (1.f - momentum) * (float)unb;
  }
}

// y = gamma*(x-mean)*inv_std + beta; optional: y_out = y (pre-FiLM); FiLM y = relu(y*fg[n,c]+fb[n,c]) + res;
// ReLU; Dropout keep mask (u8) with scale.
template <typename T>
__global__ void bn_apply_kernel(const T* x, const float* mean, const float* inv_std, const float* gamma,
                                const float* beta, T* out, T* y_out, long long total, int C, int relu,
                                const float* fg, const float* fb, int fstride, long long per_sample, const T* res,
                                const unsigned char* keep, float keep_scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = i % C;
    float v = fmaf((ldf(x + i) - mean[c]) * inv_std[c], gamma[c], beta[c]);
    if (y_out) stf(y_out + i, v);
    if (fg) {
      const long long n = i / per_sample;
      v = fmaxf(fmaf(v, fg[n * fstride + c], fb[n * fstride + c]), 0.f) + ldf(res + i);
    }
    if (relu) v = fmaxf(v, 0.f);
    if (keep) v = keep[i] ? v * keep_scale : 0.f;
    stf(out + i, v);
  }
}

// red[c] += sum dy, red[C + c] += sum dy * xhat   (xhat = (x-mean)*inv_std); dy optionally masked by out > 0
template <typename T>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const T* dy, const T* x, const T* relu_out,
                                                            const float* mean, const float* inv_std, long long rows,
                                                            int C, float* red, long long rows_per_cta) {
  __shared__ float s1[8][33], s2[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + cx;
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  float a = 0.f, b = 0.f;
  if (c < C) {
    const float mu = mean[c], is = inv_std[c];
    for (long long r = r0 + ry; r < r1; r += 8) {
      float d = ldf(dy + r * C + c);
      if (relu_out && !(ldf(relu_out + r * C + c) > 0.f)) d = 0.f;
      a += d;
      b = fmaf(d, (ldf(x + r * C + c) - mu) * is, b);
    }
  }
  s1[ry][cx] = a;
  s2[ry][cx] = b;
  __syncthreads();
  if (ry == 0 && c < C) {
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { t1 += s1[i][cx]; t2 += s2[i][cx]; }
    atomicAdd(red + c, t1);
    atomicAdd(red + C + c, t2);
  }
}

// d_x = gamma*inv_std*(dy - dbeta/M - xhat*dgamma/M); writes dgamma/dbeta to the gradient buffer too
template <typename T>
__global__ void bn_bwd_apply_kernel(const T* dy, const T* x, const T* relu_out, const float* mean,
                                    const float* inv_std, const float* gamma, const float* red, float invM, T* dx,
                                    long long total, int C) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = i % C;
    float d = ldf(dy + i);
    if (relu_out && !(ldf(relu_out + i) > 0.f)) d = 0.f;
    const float is = inv_std[c];
    const float xh = (ldf(x + i) - mean[c]) * is;
    stf(dx + i, gamma[c] * is * (d - red[c] * invM - xh * red[C + c] * invM));
  }
}

__global__ void copy2_kernel(const float* red, float* dbeta, float* dgamma, int C, float scale) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) { dbeta[c] = red[c] * scale; dgamma[c] = red[C + c] * scale; }
}

template <typename T>
__global__ void dropout_bwd_kernel(const T* dy, const unsigned char* keep, float scale, T* dx, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    stf(dx + i, keep[i] ? ldf(dy + i) * scale : 0.f);
}

// ---- dense layers of the noise MLP (tiny): Y = X W + b ; G = X^T dY ; dX = dY W^T ----
__global__ void dense_fwd_kernel(const float* X, const float* W, const float* b, float* Y, int rows, int K, int C) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)rows * C) return;
  const int r = i / C, c = i % C;
  float acc = b ? b[c] : 0.f;
  for (int k = 0; k < K; ++k) acc = fmaf(X[(size_t)r * K + k], W[(size_t)k * C + c], acc);
  Y[i] = acc;
}
__global__ void dense_bwd_w_kernel(const float* X, const float* dY, float* G, float* db, int rows, int K, int C) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)K * C) return;
  const int k = i / C, c = i % C;
  float acc = 0.f;
  for (int r = 0; r < rows; ++r) acc = fmaf(X[(size_t)r * K + k], dY[(size_t)r * C + c], acc);
  G[i] = acc;
  if (k == 0 && db) {
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += dY[(size_t)r * C + c];
    db[c] = s;
  }
}
// dX[r,k] (+)= sum_c dY[r, ystride*r.. + yoff + c] W[k,c]
__global__ void dense_bwd_x_kernel(const float* dY, int ystride, int yoff, const float* W, float* dX, int rows, int K,
                                   int C, int accumulate) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)rows * K) return;
  const int r = i / K, k = i % K;
  float acc = 0.f;
  for (int c = 0; c < C; ++c) acc = fmaf(dY[(size_t)r * ystride + yoff + c], W[(size_t)k * C + c], acc);
  dX[i] = accumulate ? dX[i] + acc : acc;
}
__global__ void strided_copy_kernel(const float* src, int sstride, int soff, float* dst, int dstride, int doff, int rows,
                                    int C) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)rows * C) return;
  const int r = i / C, c = i % C;
  dst[(size_t)r * dstride + doff + c] = src[(size_t)r * sstride + soff + c];
}

// Softmax + keras categorical_crossentropy (from_logits=False): p /= sum p; clip to [1e-7, 1-1e-7];
// loss += -sum_k t_k log p_k / (N_global*H*W);  dseg = softmax-Jacobian^T applied to dloss/dp.
__global__ void __launch_bounds__(256) softmax_cce_kernel(const float* prob, const float* target, float* dseg,
                                                          float* loss, long long npix, int nc, float inv_total) {
  float local = 0.f;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
    float q[4], t[4], g[4];
    float S = 0.f;
    for (int k = 0; k < nc; ++k) { q[k] = prob[p * nc + k]; t[k] = target[p * nc + k]; S += q[k]; }
    float tin = 0.f;
    for (int k = 0; k < nc; ++k) {
      const float o = q[k] / S;
      const bool inr = o >= 1e-7f && o <= 1.f - 1e-7f;
      const float oc = fminf(fmaxf(o, 1e-7f), 1.f - 1e-7f);
      local -= t[k] * logf(oc);
      g[k] = inr ? -t[k] / q[k] : 0.f;  // d(-t log(q/S))/dq_k (direct term)
      tin += inr ? t[k] : 0.f;
    }
    float dot = 0.f;
    for (int k = 0; k < nc; ++k) { g[k] += tin / S; dot = fmaf(g[k], q[k], dot); }
    for (int k = 0; k < nc; ++k) dseg[p * nc + k] = q[k] * (g[k] - dot) * inv_total;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(loss, local * inv_total);
}

// 1x1 multi-class head backward: d_o[p,c] = sum_k dseg[p,k] w[c,k] masked by o>0; d_w[c,k] += sum_p o[p,c] dseg[p,k];
// d_b[k] += sum_p dseg[p,k].  One warp per pixel group, lane = channel (C == 32).
template <typename T>
__global__ void __launch_bounds__(256) head_bwd_multi_kernel(const float* dseg, const T* o, const float* w, T* d_o,
                                                             float* d_w, float* d_b, long long npix, int C, int nc) {
  __shared__ float s_acc[8][32][5];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float wk[4] = {0.f, 0.f, 0.f, 0.f}, aw[4] = {0.f, 0.f, 0.f, 0.f}, ab[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k = 0; k < nc; ++k) wk[k] = lane < C ? w[lane * nc + k] : 0.f;
  for (long long p = warp; p < npix; p += nwarps) {
    float ds[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < nc; ++k) ds[k] = dseg[p * nc + k];
    if (lane < C) {
      const float ov = ldf(o + p * C + lane);
      float d = 0.f;
      for (int k = 0; k < nc; ++k) { d = fmaf(ds[k], wk[k], d); aw[k] = fmaf(ov, ds[k], aw[k]); }
      stf(d_o + p * C + lane, ov > 0.f ? d : 0.f);
    }
    if (lane == 0)
      for (int k = 0; k < nc; ++k) ab[k] += ds[k];
  }
  for (int k = 0; k < 4; ++k) s_acc[wp][lane][k] = aw[k];
  s_acc[wp][lane][4] = 0.f;
  __syncthreads();
  if (wp == 0 && lane < C)
    for (int k = 0; k < nc; ++k) {
      float s = 0.f;
      for (int i = 0; i < 8; ++i) s += s_acc[i][lane][k];
      atomicAdd(d_w + lane * nc + k, s);
    }
  if (lane == 0)
    for (int k = 0; k < nc; ++k) atomicAdd(d_b + k, ab[k]);
}

}  // namespace

#define DISPATCH_DT(dt, CALL_F32, CALL_BF16) \
  do {                                       \
    if ((dt) == DT_F32) { CALL_F32; } else { CALL_BF16; } \
  } while (0)

// ---- synchronised BatchNorm for data-parallel training: the per-channel batch sums of the forward pass
// and of the backward reduction are summed over the ranks by a caller-supplied hook before they are used, and the
// element count becomes the global one.  One process drives one GPU, so the hook is process-global.
static depgan_allreduce_fn g_sync_fn = nullptr;
static void* g_sync_user = nullptr;
static int g_sync_world = 1;
int bn_set_sync_hook(depgan_allreduce_fn fn, void* user, int world) {
  g_sync_fn = world > 1 ? fn : nullptr;
  g_sync_user = user;
  g_sync_world = (world > 1 && fn) ? world : 1;
  return 0;
}
int bn_sync_world() { return g_sync_world; }

int k_bn_stats(const void* x, long long rows, int C, double* sums_scratch, float* mean, float* inv_std, float* mov_mean,
               float* mov_var, float momentum, int dt, cudaStream_t st) {
  if (rows == 0) return 0;
  DG_CHECK_CUDA(cudaMemsetAsync(sums_scratch, 0, sizeof(double) * 2 * C, st));
  long long gx = (rows + 1023) / 1024;
  if (gx > 148 * 4) gx = 148 * 4;
  const long long rpc = (rows + gx - 1) / gx;
  dim3 grid((unsigned)((rows + rpc - 1) / rpc), (C + 31) / 32);
  DISPATCH_DT(dt, (bn_sums_kernel<float><<<grid, 256, 0, st>>>((const float*)x, rows, C, sums_scratch, rpc)),
              (bn_sums_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, rows, C, sums_scratch, rpc)));
  DG_LAUNCH_CHECK();
  if (g_sync_fn) {
    if (g_sync_fn(g_sync_user, sums_scratch, 2LL * C, 1, (void*)st) != 0) {
      depgan_set_error("bn_stats: the all-reduce hook failed");
      return -1;
    }
  }
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums_scratch, (double)rows * g_sync_world, C, mean, inv_std,
                                                      mov_mean, mov_var, momentum);
  DG_LAUNCH_CHECK();
  return 0;
}

int k_bn_apply(const void* x, const float* mean, const float* inv_std, const float* gamma, const float* beta, void* out,
               void* y_out, long long rows, int C, int relu, const float* fg, const float* fb, int fstride,
               long long rows_per_sample, const void* res, const unsigned char* keep, float keep_scale, int dt,
               cudaStream_t st) {
  const long long total = rows * C;
  if (total == 0) return 0;
  DISPATCH_DT(dt,
              (bn_apply_kernel<float><<<grid_for(total), 256, 0, st>>>(
                  (const float*)x, mean, inv_std, gamma, beta, (float*)out, (float*)y_out, total, C, relu, fg, fb,
                  fstride, rows_per_sample * C, (const float*)res, keep, keep_scale)),
              (bn_apply_kernel<bf16><<<grid_for(total), 256, 0, st>>>(
                  (const bf16*)x, mean, inv_std, gamma, beta, (bf16*)out, (bf16*)y_out, total, C, relu, fg, fb, fstride,
                  rows_per_sample * C, (const bf16*)res, keep, keep_scale)));
  DG_LAUNCH_CHECK();
  return 0;
}

int k_bn_bwd(const void* dy, const void* x, const void* relu_out, const float* mean, const float* inv_std,
             const float* gamma, float* red_scratch, void* dx, float* dgamma, float* dbeta, long long rows, int C,
             int dt, cudaStream_t st) {
  if (rows == 0) return 0;
  DG_CHECK_CUDA(cudaMemsetAsync(red_scratch, 0, sizeof(float) * 2 * C, st));
  long long gx = (rows + 1023) / 1024;
  if (gx > 148 * 4) gx = 148 * 4;
  const long long rpc = (rows + gx - 1) / gx;
  dim3 grid((unsigned)((rows + rpc - 1) / rpc), (C + 31) / 32);
  DISPATCH_DT(dt,
              (bn_bwd_reduce_kernel<float><<<grid, 256, 0, st>>>((const float*)dy, (const float*)x,
                                                                 (const float*)relu_out, mean, inv_std, rows, C,
                                                                 red_scratch, rpc)),
              (bn_bwd_reduce_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)dy, (const bf16*)x, (const bf16*)relu_out,
                                                                mean, inv_std, rows, C, red_scratch, rpc)));
  DG_LAUNCH_CHECK();
  if (g_sync_fn) {
    if (g_sync_fn(g_sync_user, red_scratch, 2LL * C, 0, (void*)st) != 0) {
      depgan_set_error("bn_bwd: the all-reduce hook failed");
      return -1;
    }
  }
  const float inv_rows = 1.0f / ((float)rows * (float)g_sync_world);
  const long long total = rows * C;
  DISPATCH_DT(dt,
              (bn_bwd_apply_kernel<float><<<grid_for(total), 256, 0, st>>>((const float*)dy, (const float*)x,
                                                                          (const float*)relu_out, mean, inv_std, gamma,
                                                                          red_scratch, inv_rows, (float*)dx,
                                                                          total, C)),
              (bn_bwd_apply_kernel<bf16><<<grid_for(total), 256, 0, st>>>((const bf16*)dy, (const bf16*)x,
                                                                         (const bf16*)relu_out, mean, inv_std, gamma,
                                                                         red_scratch, inv_rows, (bf16*)dx,
                                                                         total, C)));
  DG_LAUNCH_CHECK();
  // synchronised: red_scratch already holds the GLOBAL sums; the flat gradient bucket is summed over the ranks later,
  // so each rank contributes 1/world of them
  copy2_kernel<<<(C + 127) / 128, 128, 0, st>>>(red_scratch, dbeta, dgamma, C, 1.0f / (float)g_sync_world);
  DG_LAUNCH_CHECK();
  return 0;
}

int k_dropout_bwd(const void* dy, const unsigned char* keep, float scale, void* dx, long long total, int dt,
                  cudaStream_t st) {
  if (total == 0) return 0;
  DISPATCH_DT(dt,
              (dropout_bwd_kernel<float><<<grid_for(total), 256, 0, st>>>((const float*)dy, keep, scale, (float*)dx,
                                                                         total)),
              (dropout_bwd_kernel<bf16><<<grid_for(total), 256, 0, st>>>((const bf16*)dy, keep, scale, (bf16*)dx,
                                                                        total)));
  DG_LAUNCH_CHECK();
  return 0;
}

int k_dense_fwd(const float* X, const float* W, const float* b, float* Y, int rows, int K, int C, cudaStream_t st) {
  const long long n = (long long)rows * C;
  if (n == 0) return 0;
  dense_fwd_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(X, W, b, Y, rows, K, C);
  DG_LAUNCH_CHECK();
  return 0;
}
int k_dense_bwd_w(const float* X, const float* dY, float* G, float* db, int rows, int K, int C, cudaStream_t st) {
  const long long n = (long long)K * C;
  if (n == 0) return 0;
  dense_bwd_w_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(X, dY, G, db, rows, K, C);
  DG_LAUNCH_CHECK();
  return 0;
}
int k_dense_bwd_x(const float* dY, int ystride, int yoff, const float* W, float* dX, int rows, int K, int C,
                  int accumulate, cudaStream_t st) {
  const long long n = (long long)rows * K;
  if (n == 0) return 0;
  dense_bwd_x_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(dY, ystride, yoff, W, dX, rows, K, C, accumulate);
  DG_LAUNCH_CHECK();
  return 0;
}
int k_strided_copy(const float* src, int sstride, int soff, float* dst, int dstride, int doff, int rows, int C,
                   cudaStream_t st) {
  const long long n = (long long)rows * C;
  if (n == 0) return 0;
  strided_copy_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(src, sstride, soff, dst, dstride, doff, rows, C);
  DG_LAUNCH_CHECK();
  return 0;
}

int k_softmax_cce(const float* prob, const float* target, float* dseg, float* loss, long long npix, int nc,
                  float inv_total, cudaStream_t st) {
  if (npix == 0) return 0;
  DG_REQUIRE(nc >= 1 && nc <= 4, "softmax_cce: nc must be 1..4");
  softmax_cce_kernel<<<grid_for(npix, 256, 148 * 8), 256, 0, st>>>(prob, target, dseg, loss, npix, nc, inv_total);
  DG_LAUNCH_CHECK();
  return 0;
}

int k_head_bwd_multi(const float* dseg, const void* o, const float* w, void* d_o, float* d_w, float* d_b,
                     long long npix, int C, int nc, int dt, cudaStream_t st) {
  if (npix == 0) return 0;
  DG_REQUIRE(C <= 32 && nc <= 4, "head_bwd_multi: C <= 32, nc <= 4");
  const int grid = grid_for(npix * 32, 256, 148 * 8);
  DISPATCH_DT(dt,
              (head_bwd_multi_kernel<float><<<grid, 256, 0, st>>>(dseg, (const float*)o, w, (float*)d_o, d_w, d_b, npix,
                                                                 C, nc)),
              (head_bwd_multi_kernel<bf16><<<grid, 256, 0, st>>>(dseg, (const bf16*)o, w, (bf16*)d_o, d_w, d_b, npix, C,
                                                                nc)));
  DG_LAUNCH_CHECK();
  return 0;
}
