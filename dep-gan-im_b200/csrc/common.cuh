// Shared declarations for the depgan_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <string>

typedef __nv_bfloat16 bf16;

// ---- error plumbing (no exceptions across the C ABI) ----
void depgan_set_error(const std::string& msg);
extern long long g_launch_count;

#define DG_CHECK_CUDA(expr)                                                                              \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess) {                                                                             \
      depgan_set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " @" + __FILE__ + ":" +      \
                       std::to_string(__LINE__));                                                        \
      return -1;                                                                                         \
    }                                                                                                    \
  } while (0)

#define DG_LAUNCH_CHECK()                                                                                \
  do {                                                                                                   \
    ++g_launch_count;                                                                                    \
    cudaError_t _e = cudaGetLastError();                                                                 \
    if (_e != cudaSuccess) {                                                                             \
      depgan_set_error(std::string("kernel launch: ") + cudaGetErrorString(_e) + " @" + __FILE__ + ":" + \
                       std::to_string(__LINE__));                                                        \
      return -1;                                                                                         \
    }                                                                                                    \
  } while (0)

// Programmatic dependent launch (sm_90+): a kernel launched through dg_launch_pdl may be scheduled while its stream
// predecessor is still draining; DG_PDL_ENTER() at the top of the kernel (before any access to global memory) lets
// the next launch do the same and then waits for the predecessor to complete and flush.  Both instructions are no-ops
// for an ordinary <<<>>> launch.  DEPGAN_NO_PDL=1 turns the attribute off (A/B measurements).
#define DG_PDL_ENTER()                                                    \
  do {                                                                    \
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");       \
    asm volatile("griddepcontrol.wait;" ::: "memory");                    \
  } while (0)
#ifdef __CUDACC__
template <typename K, typename... Args>
inline cudaError_t dg_launch_pdl(K kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  static const bool pdl = getenv("DEPGAN_NO_PDL") == nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}
#endif

#define DG_REQUIRE(cond, msg)                                        \
  do {                                                               \
    if (!(cond)) {                                                   \
      depgan_set_error(std::string("requirement failed: ") + (msg)); \
      return -2;                                                     \
    }                                                                \
  } while (0)

#define DG_TRY(expr)           \
  do {                         \
    int _r = (expr);           \
    if (_r != 0) return _r;    \
  } while (0)

// Per-device one-time set-up.  cudaFuncSetAttribute (shared-memory opt-in, carve-out) and the SM count belong to a
// device, not to the process: a site keeps one DgPerDevice and runs its set-up once for every device it is called on.
struct DgPerDevice {
  unsigned long long done = 0;  // bit d: device d has been set up at this site
  int sms[64] = {};
  int val[64] = {};             // site-specific cached value (e.g. resident CTAs per SM)
};
// 0 on success; *dev = current device, *first = true when this site has not run its set-up on it yet (the caller
// then does the set-up and calls dg_device_mark).  Devices >= 64 are always treated as "first" (never cached).
inline int dg_device_enter(DgPerDevice& s, int* dev, bool* first) {
  if (cudaGetDevice(dev) != cudaSuccess) { depgan_set_error("cudaGetDevice failed"); return -1; }
  *first = *dev >= 64 || !((s.done >> *dev) & 1ull);
  if (*first) {
    int n = 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, *dev) != cudaSuccess) {
      depgan_set_error("cudaDeviceGetAttribute(multiProcessorCount) failed");
      return -1;
    }
    if (*dev < 64) s.sms[*dev] = n;
  }
  return 0;
}
inline void dg_device_mark(DgPerDevice& s, int dev) { if (dev < 64) s.done |= 1ull << dev; }

// DT_F16: IEEE half storage of the inference-only generator handles (DEPGAN_PREC_F16): same tcgen05 kind::f16 rate and
// the same kernels as bf16 (the 16-bit format is a run-time flag of the pack / unpack points), 3 more mantissa bits
// per stored activation -- DEM max-abs error ~1.5e-3 instead of ~1e-2 with un-normalised (freshly initialised) weights.
// DT_F16S: "split half" storage of the tensor-core <= 1e-4 handles (DEPGAN_PREC_F16X3): every value v is kept as the pair
// hi = half(v), lo = half(v - hi) (22 significant bits), a pixel's C channels stored as [C hi | C lo] (4 bytes per
// element, the size of fp32).  Convolutions run as three kind::f16 products x_hi*w_hi + x_lo*w_hi + x_hi*w_lo with fp32
// accumulation in TMEM (the dropped x_lo*w_lo term is 2^-22 relative), i.e. a K = 3*Cin implicit GEMM of the same kernel.
enum DType { DT_F32 = 0, DT_BF16 = 1, DT_F16 = 2, DT_F16S = 3 };
static inline size_t dt_size(int dt) { return (dt == DT_F32 || dt == DT_F16S) ? 4 : 2; }
static inline bool dt_is_half(int dt) { return dt == DT_BF16 || dt == DT_F16; }
static inline bool dt_is_tc(int dt) { return dt_is_half(dt) || dt == DT_F16S; }  // formats the tcgen05 kernels take

// ---- typed load/store helpers ----
__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(bf16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ float ldf(const __half* p) { return __half2float(*p); }
__device__ __forceinline__ void stf(__half* p, float v) { *p = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f)); }

// ---- two 16-bit floats in one word; `f16` (warp-uniform) selects IEEE half, else bfloat16.  Low half = first value. ----
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi, bool f16) {
  uint32_t r;
  if (f16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float2 unpack_h2(uint32_t w, bool f16) {
  if (f16) return __half22float2(*reinterpret_cast<const __half2*>(&w));
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u));
}
// the same conversion with the ReLU folded in (negative results become +0): one instruction instead of two max + one cvt
__device__ __forceinline__ uint32_t pack_h2_relu(float lo, float hi, bool f16) {
  uint32_t r;
  if (f16) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// (x0, x1) = (x0, x1) * (s0, s1) + (t0, t1) as one packed fp32x2 FMA (per-lane IEEE fma: the bits are those of fmaf)
__device__ __forceinline__ void fma_f32x2(float& x0, float& x1, float s0, float s1, float t0, float t1) {
  unsigned long long xv, sv, tv;
  asm("mov.b64 %0, {%1, %2};" : "=l"(xv) : "f"(x0), "f"(x1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(sv) : "f"(s0), "f"(s1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(tv) : "f"(t0), "f"(t1));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(xv) : "l"(xv), "l"(sv), "l"(tv));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(xv));
}
// element-wise max of two packed pairs (2x2 max-pool on the stored values)
__device__ __forceinline__ uint32_t max_h2(uint32_t a, uint32_t b, bool f16) {
  if (f16) {
    const __half2 m = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&m);
  }
  const __nv_bfloat162 m = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&m);
}

// ---- convolution argument block (SIMT and tcgen05 paths share it) ----
// acc = conv(in0 || in1, w)  ('same' padding, stride 1, ks in {1,3,5});  then per output element (n,h,w,c):
//   v = acc * scale[c] + shift[c]                       (folded bias + inference BatchNorm, TG:285-297)
//   if out_pre : out_pre = v                            (pre-FiLM tensor kept for backward)
//   if film_g  : v = relu(v * film_g[n,c] + film_b[n,c]) + res      (mul/add/relu/add_noiseZres, TG:403-407)
//   if add_src : v += add_src
//   if mask_src: v = mask_src > 0 ? v : 0               (ReLU / activation-pattern mask for dgrad and JVP)
//   if relu    : v = max(v, 0)
//   if out     : out = v
//   if pool_out: pool_out[n,h/2,w/2,c] = max of the 2x2 window of out            (tcgen05 path only, fused)
//   if head_w  : head_out[n,h,w,:] = act(sum_c v[c] * head_w[c,:] + head_b)   (1x1 gen_segmentation, TG:494-495)
// deconv = 1: ks must be 1; weights hold 4*Cout columns (a,b,co) and column (ab,co) of input pixel (h,w) is
//   written to out[n, 2h+a, 2w+b, co] (Conv2DTranspose k2 s2, TG:307-312); scale/shift are indexed by co.
struct ConvArgs {
  const void* in0;
  const void* in1;
  int C0, C1;
  const float* w;      // SIMT: [taps][Cin][Ncols] fp32   (deconv: [Cin][4*Cout])
  const bf16* w_tc;    // tcgen05: [taps][Ncols][Cin] bf16 (K-major B operand)
  const float* scale;  // [Cout] or nullptr (1)
  const float* shift;  // [Cout] or nullptr (0)
  void* out;
  void* out_pre;
  const float* film_g;
  const float* film_b;
  int film_stride;
  const void* res;
  const void* add_src;
  const void* mask_src;
  int relu;
  int deconv;
  const float* head_w;  // [Cout][head_nc] fp32
  const float* head_b;  // [head_nc]
  float* head_out;      // (N,H,W,head_nc) fp32
  int head_nc, head_act;  // act: 0 tanh, 1 softmax, 2 linear
  int N, H, W, Cout, ks;
  int in_dt, out_dt;  // DType of in0/in1 and of out/out_pre/res/add_src/mask_src
  void* pool_out;     // optional (N,H/2,W/2,Cout): 2x2 stride-2 max-pool of `out` (MaxPooling2D after the block, TG:409)
};

int conv_fwd_simt(const ConvArgs& a, cudaStream_t st);
int conv_fwd_tc(const ConvArgs& a, cudaStream_t st);  // tcgen05 path (bf16 in/out)
int conv_first_band_try(const ConvArgs& a, cudaStream_t st);  // conv2d_dis_0a, image row as the A operand: 1 taken, 0 not a case
int conv_last_band_try(const ConvArgs& a, cudaStream_t st);   // 16 -> 1 data gradient of conv2d_dis_0a: 1 taken, 0 not a case
int conv_first_tc_try(const ConvArgs& a, cudaStream_t st);  // tcgen05 first layer (fp32 image in): 1 taken, 0 not a case
bool conv_tc_supported(const ConvArgs& a);
bool conv_tc_plan_query(const ConvArgs& a, int* plan16);  // host-only: the planner's geometry for a supported layer
int conv_tc_init();  // resolves cuTensorMapEncodeTiled, sets smem attributes
// row-streaming tcgen05 kernel for the 3x3, 32-output-channel layers at full width (conv_row.cu); conv_fwd_tc routes to it
bool conv_row_supported(const ConvArgs& a);
int conv_fwd_row(const ConvArgs& a, cudaStream_t st);
// generic row-streaming kernel (conv_rowg.cu): 5x5 critic layers, 64-output-channel 3x3 layers; width a multiple of 128
bool conv_rowg_supported(const ConvArgs& a);
int conv_fwd_rowg(const ConvArgs& a, cudaStream_t st);

// wgrad: dw[tap][ci][co] += alpha * sum_p x[p+off(tap)][ci] * dy[p][co]  (fp32 accumulate, atomics)
struct WgradArgs {
  const void* x0;
  const void* x1;
  int C0, C1;
  const void* dy;
  float* dw;  // [taps][Cin][Cout] fp32, accumulated into (caller zeroes)
  int N, H, W, Cout, ks;
  int x_dt, dy_dt;
  float alpha;  // scale applied to the accumulated sum
  float* csum;  // optional: csum[co] += sum_p dy[p][co] (the bias / BN-shift gradient), fused when the kernel can
};
int conv_wgrad_simt(const WgradArgs& a, cudaStream_t st);
int conv_wgrad_tc(const WgradArgs& a, cudaStream_t st);  // tcgen05 path (bf16 x and dy); handles a.csum itself
bool wgrad_tc_supported(const WgradArgs& a);
int wgrad_tc_init();
int wgrad_first_band_try(const WgradArgs& a, cudaStream_t st);  // wgrad_first_band.cu: 1 taken, 0 not a case, < 0 error
int wgrad_first_tc_try(const WgradArgs& a, cudaStream_t st);  // tcgen05 first-layer wgrad (conv_first_tc.cu)
