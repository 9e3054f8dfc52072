// Launchers for the non-convolution kernels of the hot path (defined in kernels_fwd.cu / kernels_bwd.cu).
#pragma once
#include "common.cuh"

// ---- parameter preparation --------------------------------------------------------------------------
// scale[c] = gamma/sqrt(var+eps), shift[c] = beta + (bias - mean)*scale  (bn != null) ; else scale=1, shift=bias
int k_fold_bn(const float* bias, const float* gamma, const float* beta, const float* mean, const float* var,
              float* scale, float* shift, float* inv_std, int C, cudaStream_t st);
// dst_tc[tap][co][ci] (bf16) and dst_dgrad[flip(tap)][co][ci] (fp32) from src HWIO [tap][ci][co]
// (dgrad variants: taps flipped, multiplied by scale[co] when scale != nullptr)
int k_pack_conv_weights(const float* src, const float* scale, bf16* dst_tc, float* dst_dgrad, bf16* dst_tc_dgrad,
                        int taps, int Cin, int Cout, cudaStream_t st, int f16 = 0);  // f16: IEEE half in the 16-bit slots

// ---- forward -----------------------------------------------------------------------------------------
int k_maxpool_fwd(const void* in, void* out, int N, int H, int W, int C, int dt, cudaStream_t st);
// Conv2DTranspose k2 s2 'valid' + folded BN + ReLU (TG:307-312).  w: Keras (2,2,Cout,Cin) fp32.
int k_deconv_fwd(const void* in, const float* w, const float* scale, const float* shift, void* out, int N, int H,
                 int W, int Cin, int Cout, int dt, int relu, cudaStream_t st);
// 1x1 conv (Cin -> nc_out<=4) + tanh / softmax, fp32 output (TG:494-495, TU:423-424). seg_out optional (pre-act)
int k_head_fwd(const void* in, const float* w, const float* b, float* out, long long npix, int Cin, int nc_out,
               int head, int dt, cudaStream_t st);

struct FilmMlpArgs {          // noise path TG:353-395
  const float* z;             // (N, L, 1)
  const float* k0; const float* s0; const float* t0;   // dense 1->F kernel (F), folded scale/shift incl. bias (F)
  const float* k1; const float* s1; const float* t1;   // dense F->F kernel (F,F) (in,out), folded scale/shift
  const float* const* head_w; // device array [n_heads] of kernels (L*F, C_h) fp32
  const float* const* head_s; // device array [n_heads] folded scale (C_h)
  const float* const* head_t; // device array [n_heads] folded shift (C_h)
  const int* head_c;          // device array [n_heads] widths
  const int* head_off;        // device array [n_heads] column offset in `out`
  int n_heads, total_c;
  float* h1;                  // (N, L, F) post-ReLU activations (kept for backward)
  float* h2;                  // (N, L*F) flattened post-ReLU activations
  float* out;                 // (N, total_c): all heads side by side
  int N, L, F;
};
int k_film_mlp_fwd(const FilmMlpArgs& a, cudaStream_t st);

// critic tail: dis_9 (1x1, C->1) + Flatten + Dense(1) (TG:339-342). in: (N, HW, C); out (N)
int k_critic_head_fwd(const void* in, const float* w9, const float* b9, const float* wd, const float* bd, float* out,
                      int N, int HW, int C, int dt, cudaStream_t st);

// dst(T) = src(f32) ; optional channel selection (src has cs channels, take channel 0)
int k_convert_in(const float* src, void* dst, long long n, int dt, cudaStream_t st);

// ---- train-step input preparation (TG:528-538, 556-557) ------------------------------------------------
// which=0 (Y2 critic): real = real2, fake = base + dem ; which=1 (DEM critic): real = real2 - base, fake = dem
// writes batch3 = [real | fake | mixed] (3N,H,W,1) in dtype dt, mixed = ep*real + (1-ep)*fake
int k_critic_inputs(const float* real2, const float* x1, int nicg, const float* dem, const float* ep, int which,
                    void* batch3, int N, long long hw, int dt, cudaStream_t st);

// ---- inference accumulation / post-processing (EG:616-628, 673-741; EU:559-570, 597-600) ---------------
int k_dem_accumulate(double* acc, const float* pred, const float* mask, long long n, int chan, cudaStream_t st);
int k_dem_postproc(const float* x, int nicg, const double* acc, double n_repeat, const float* mask, double thr,
                   double* dem_out, double* fake2_out, unsigned char* labels, unsigned long long* count,
                   long long npix, cudaStream_t st);
int k_label_confusion(const unsigned char* fake, const unsigned char* real, long long n, unsigned long long* conf,
                      cudaStream_t st);
int k_uresnet_labels(const double* acc, double n_repeat, int chan, double* mean_out, unsigned char* labels,
                     unsigned long long* count, long long npix, cudaStream_t st);

int k_adam(float* p, const float* g, float* m, float* v, long long n, float lr_t, float b1, float b2, float eps,
           float gscale, cudaStream_t st);

int k_cast_narrow(const float* src, void* dst, long long n, int to_f16, cudaStream_t st);
// ---- split-half storage (DT_F16S: per pixel [C hi | C lo] IEEE halves, v = hi + lo) -------------------------------
// tcgen05 B operand of a split-half layer: dst[row][3*Cin] with row = tap*Cout + co, columns in the kernel's K order
// [hi(0:c0), hi(0:c0), hi(c0:Cin), hi(c0:Cin), lo(0:c0), lo(c0:Cin)] (c0 = channels of the first input tensor).
// src: Keras HWIO [tap][ci][co], or (kmajor = 1, the transposed conv) [tap][co][ci].
int k_pack_split_weights(const float* src, bf16* dst, int taps, int Cin, int Cout, int c0, int kmajor, cudaStream_t st);
int k_split_to_f32(const void* src, float* dst, long long npix, int C, cudaStream_t st);
int k_copy_to_f32(const void* src, float* dst, long long n, int dt, cudaStream_t st);

// ---- backward / loss kernels (kernels_bwd.cu) ------------------------------------------------------------
int k_maxpool_bwd(const void* dy, const void* x, const void* add_src, void* dx, int N, int H, int W, int C, int dt,
                  cudaStream_t st);
int k_maxpool_select(const void* v, const void* x, void* out, int N, int H, int W, int C, int dt, cudaStream_t st);
int k_channel_sum(const void* src, long long rows, int C, float* out, float alpha, int dt, cudaStream_t st);
int k_param_grads(const float* W, const float* G, float* dW, const float* scale, const float* inv_std,
                  const float* bias, const float* mean, const float* sum_dy, float* dgamma, float* dbeta, float* dbias,
                  int K, int C, int trans, int Cin, cudaStream_t st);
int k_critic_head_bwd(const void* h, const void* v, const float* go, const float* w9, const float* b9, const float* wd,
                      void* dh, float* d_w9, float* d_b9, float* d_wd, float* d_bd, int rows, int n_reg, int HW, int C,
                      int want_param_grads, int dt, cudaStream_t st);
int k_gp(const float* g, float* u, float* gp_out, int n, long long hw, float delta, float inv_n, cudaStream_t st);
int k_segment_sum(const float* src, int seg, int nseg, float* out, float alpha, cudaStream_t st);
int k_gen_loss_sums(const float* dem, const float* x1, int nicg, const float* real2, float thr, float* fake2,
                    float* l1g, float l1coef, double* sums, long long n, cudaStream_t st);
int k_loss_consts(double* sums, double gn, double hw, int k, cudaStream_t st);  // sums[j][6] = gn, sums[j][7] = hw
int k_gen_loss_finalize(float* out6, const double* sums, cudaStream_t st, int k = 1);  // k candidates: out6[k][6], sums[k][8]
int k_gen_head_bwd(const float* gy2, const float* gdem, const float* l1g, const float* dem, const void* o,
                   const float* w, void* d_o, float* d_w, float* d_b, long long npix, int C, int dt, cudaStream_t st);
int k_film_bwd(const void* d_r, const void* y, const float* fg, const float* fb, int fstride, void* d_y, float* dgam,
               float* dbet, int N, int HW, int C, int dt, cudaStream_t st);
int k_add_mask(const void* a, const void* b, const void* m, void* out, long long n, int dt, cudaStream_t st);
int k_s2d_mask(const void* d_up, int dstride, const void* up, void* out, int N, int H, int W, int C, int dt,
               cudaStream_t st);
int k_slice_add(const void* src, int sstride, int off, const void* add, void* dst, long long rows, int C, int dt,
                cudaStream_t st);
int k_film_mlp_bwd(const FilmMlpArgs& a, const float* d_out, float* const* dW_heads, float* sum_d, float* d_h2,
                   float* dk1raw, float* sum_d1, float* dk0raw, float* sum_d0, cudaStream_t st);
int k_fill_go(float* go, int n, float v_real, float v_fake, float v_mixed, int rows, cudaStream_t st);
int k_critic_loss_finalize(float* out4, float delta, cudaStream_t st);
int k_scores_to_sums(const float* sy2, const float* sdem, int n, double* sums, double gn, double hw, cudaStream_t st,
                     int k = 1);
int k_pack_deconv_dgrad(const float* src, const float* scale, float* dst_f32, bf16* dst_bf16, int Cin, int Cout,
                        cudaStream_t st);

// ---- training-phase BatchNorm / Dropout / softmax-CCE / dense helpers (kernels_bn.cu) ----------------------
// All conv layers of one network folded / packed by two launches (depgan_net_prepare runs after every optimiser step)
struct PrepLayer {
  const float *bias, *gamma, *beta, *mean, *var;
  float *scale, *shift, *inv_std;
  const float* w;
  bf16* w_tc;
  float* w_dg;
  bf16* w_dg_tc;
  int C, taps, cin, scale_dgrad;
};
struct PrepTable {
  int n;
  PrepLayer L[12];
};
int k_prepare_convs(const PrepTable& t, cudaStream_t st);

typedef int (*depgan_allreduce_fn)(void* user, void* dev_ptr, long long count, int is_f64, void* stream);
int bn_set_sync_hook(depgan_allreduce_fn fn, void* user, int world);
int bn_sync_world();
int k_bn_stats(const void* x, long long rows, int C, double* sums_scratch, float* mean, float* inv_std, float* mov_mean,
               float* mov_var, float momentum, int dt, cudaStream_t st);
int k_bn_apply(const void* x, const float* mean, const float* inv_std, const float* gamma, const float* beta, void* out,
               void* y_out, long long rows, int C, int relu, const float* fg, const float* fb, int fstride,
               long long rows_per_sample, const void* res, const unsigned char* keep, float keep_scale, int dt,
               cudaStream_t st);
int k_bn_bwd(const void* dy, const void* x, const void* relu_out, const float* mean, const float* inv_std,
             const float* gamma, float* red_scratch, void* dx, float* dgamma, float* dbeta, long long rows, int C,
             int dt, cudaStream_t st);
int k_dropout_bwd(const void* dy, const unsigned char* keep, float scale, void* dx, long long total, int dt,
                  cudaStream_t st);
int k_dense_fwd(const float* X, const float* W, const float* b, float* Y, int rows, int K, int C, cudaStream_t st);
int k_dense_bwd_w(const float* X, const float* dY, float* G, float* db, int rows, int K, int C, cudaStream_t st);
int k_dense_bwd_x(const float* dY, int ystride, int yoff, const float* W, float* dX, int rows, int K, int C,
                  int accumulate, cudaStream_t st);
int k_strided_copy(const float* src, int sstride, int soff, float* dst, int dstride, int doff, int rows, int C,
                   cudaStream_t st);
int k_softmax_cce(const float* prob, const float* target, float* dseg, float* loss, long long npix, int nc,
                  float inv_total, cudaStream_t st);
int k_head_bwd_multi(const float* dseg, const void* o, const float* w, void* d_o, float* d_w, float* d_b,
                     long long npix, int C, int nc, int dt, cudaStream_t st);
