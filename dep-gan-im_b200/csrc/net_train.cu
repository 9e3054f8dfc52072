// Train-step graphs of DEP-GAN (TG:523-598) as explicit forward/backward schedules over our CUDA kernels:
//   depgan_critic_grads : netD_y2_train / netD_dem_train without the optimizer update (WGAN-GP, two critics)
//   depgan_gen_eval     : netG_no_update
//   depgan_gen_grads    : netG_train without the optimizer update
// There is no autograd: the backward of every layer is written out.  Notation: b_l = dLoss/d(pre-activation of
// layer l).  BatchNorm is the learning-phase-0 affine of the reference, so
//   y = s*conv(x,W) + t,  s = gamma/sqrt(var+eps),  t = beta + (bias-mean)*s
// and with the raw weight gradient G = sum_p x (x) b:  dW = s*G, dbias = s*sum(b), dbeta = sum(b),
// dgamma = (sum_k W.G + (bias-mean)*sum(b)) / sqrt(var+eps).
//
// Gradient penalty (TG:543-545) by hand: the critic is piecewise linear, so with g = dD/dx (one dgrad sweep
// that also leaves every b_l of the mixed rows), u = dGP/dg and D_lin = the critic linearised at x (ReLU masks
// and max-pool routes frozen, biases dropped):  dGP/dW_l = wgrad(v_{l-1}, b_l)  where v = activations of the
// forward sweep of D_lin on u (a JVP).  Regular rows and penalty rows therefore share one wgrad per layer.
#include "net.h"

int train_alloc(depgan_net* h, Bump& bm) {
  Train* t = nullptr;
  if (bm.base) {
    t = new Train();
    h->tr = t;
  }
  Train dummy;
  Train& T = t ? *t : dummy;
  const depgan_cfg& c = h->cfg;
  const size_t NB = c.max_batch, es = h->es;
  const size_t HW = (size_t)c.H * c.W;
  T.sum_dy = bm.arr<float>(2048);
  T.g_raw = bm.arr<float>(128 * 512);
  if (h->model == DEPGAN_MODEL_CRITIC) {
    const size_t NM = (NB + 2) / 3 + 1;  // penalty rows of a [real|fake|mixed] batch
    for (int i = 0; i < 11; ++i) {
      const ConvL& L = h->c_conv[i];
      const size_t px = (size_t)h->lvl_h(L.lvl) * h->lvl_w(L.lvl);
      T.b[i] = bm.take(NB * px * L.cout * es);
      T.v[i] = bm.take(NM * px * L.cout * es);
      if (i == 1 || i == 3 || i == 5 || i == 7) {
        T.bp[i / 2] = bm.take(NB * px / 4 * L.cout * es);
        T.vp[i / 2] = bm.take(NM * px / 4 * L.cout * es);
      }
    }
    T.go = bm.arr<float>(NB);
    T.batch3 = bm.arr<float>(NB * HW);
    T.g_in = bm.arr<float>(NB * HW);
    T.u = bm.arr<float>(NM * HW);
  } else {
    size_t max_w = 0, max_in = 0;
    int cin = c.nicg;
    int skip_c[3] = {0, 0, 0};
    for (int bi = 0; bi < 7; ++bi) {
      const int w = FIRST_FM * GEN_MULT[bi], lvl = GEN_LVL[bi];
      if (bi >= 4) cin += skip_c[6 - bi];
      const size_t px = NB * h->lvl_h(lvl) * h->lvl_w(lvl);
      T.d_o[bi] = bm.take(px * w * es);
      if (px * w > max_w) max_w = px * w;
      if (bi > 0 && px * cin > max_in) max_in = px * cin;
      if (bi < 3) skip_c[bi] = w;
      cin = w;
    }
    T.d_r = bm.take(max_w * es);
    T.d_y = bm.take(max_w * es);
    T.b_in = bm.take(max_w * es);
    T.d_in = bm.take(max_in * es);
    T.s2d = bm.take(NB * HW * 64 * es);  // largest: (N, H/2, W/2, 4*64)
    T.d_film = bm.arr<float>(NB * 1024);
    T.d_h2 = bm.arr<float>(NB * c.noise_len * FIRST_FM);
    T.sum_d = bm.arr<float>(1024);
    T.sum_d1 = bm.arr<float>(64);
    T.sum_d0 = bm.arr<float>(64);
    T.dev_dw_heads = bm.arr<float*>(16);
    T.fake2 = bm.arr<float>(NB * HW);
    T.l1g = bm.arr<float>(NB * HW);
    T.sums = bm.arr<double>(8);
    if (c.training == 2) {  // Keras training phase (DEP-UResNet fit): pre-BN tensors and batch statistics
      auto bn = [&](BnState& s, int C) { s.mean = bm.arr<float>(C); s.inv_std = bm.arr<float>(C); };
      for (int bi = 0; bi < 7; ++bi) {
        const int w = FIRST_FM * GEN_MULT[bi], lvl = GEN_LVL[bi];
        const size_t px = NB * h->lvl_h(lvl) * h->lvl_w(lvl);
        T.raw_a[bi] = bm.take(px * w * es);
        T.raw_y[bi] = bm.take(px * w * es);
        T.raw_o[bi] = bm.take(px * w * es);
        bn(T.bn_in[bi], w); bn(T.bn_no[bi], w); bn(T.bn_out[bi], w);
        if (bi >= 3 && bi < 6) { T.raw_up[bi] = bm.take(px * 4 * w * es); bn(T.bn_dec[bi - 3], w); }
      }
      T.tmp_up = bm.take(NB * HW * 64 * es);
      const size_t LF = (size_t)c.noise_len * FIRST_FM;
      T.pre0 = bm.arr<float>(NB * LF);
      T.pre1 = bm.arr<float>(NB * LF);
      T.d_h1 = bm.arr<float>(NB * LF);
      T.d_pre = bm.arr<float>(NB * LF);
      bn(T.bn_f0, FIRST_FM); bn(T.bn_f1, FIRST_FM);
      for (int i = 0; i < 14; ++i) {
        const int C = FIRST_FM * GEN_MULT[i / 2];
        T.raw_head[i] = bm.arr<float>(NB * C);
        bn(T.bn_head[i], C);
      }
      T.tmp_c1 = bm.arr<float>(NB * 256);
      T.tmp_c2 = bm.arr<float>(NB * 256);
      T.dseg = bm.arr<float>(NB * HW * 4);
      T.bn_sums = bm.arr<double>(2 * 512);
      T.bn_red = bm.arr<float>(2 * 512);
    }
  }
  return 0;
}

namespace {

// ---- helpers --------------------------------------------------------------------------------------------
int dgrad_conv(depgan_net* h, const ConvL& L, const void* dy, void* dx, int dx_dt, const void* add_src,
               const void* mask_src, int n, cudaStream_t st) {
  ConvArgs a{};
  a.in0 = dy; a.C0 = L.cout; a.C1 = 0;
  a.w = L.w_dg; a.w_tc = L.w_dg_tc;
  a.out = dx; a.add_src = add_src; a.mask_src = mask_src;
  a.N = n; a.H = h->lvl_h(L.lvl); a.W = h->lvl_w(L.lvl); a.Cout = L.cin; a.ks = L.ks;
  a.in_dt = h->act_dt; a.out_dt = dx_dt;
  const bool tc = h->act_dt == DT_BF16 && conv_tc_supported(a);
  ProfScope prof(a, tc, st);
  if (tc) return conv_fwd_tc(a, st);
  return conv_fwd_simt(a, st);
}

// csum (optional): per-channel sums of dy over the same n rows, accumulated into csum[cout] -- fused into the tcgen05
// kernel when that path takes the call, otherwise a separate bandwidth pass.
int wgrad_conv(depgan_net* h, const ConvL& L, const void* x0, int C0, const void* x1, int C1, int x_dt, const void* dy,
               float* dw, int n, cudaStream_t st, float* csum = nullptr) {
  WgradArgs a{};
  a.x0 = x0; a.x1 = x1; a.C0 = C0; a.C1 = C1; a.dy = dy; a.dw = dw;
  a.N = n; a.H = h->lvl_h(L.lvl); a.W = h->lvl_w(L.lvl); a.Cout = L.cout; a.ks = L.ks;
  a.x_dt = x_dt; a.dy_dt = h->act_dt; a.alpha = 1.f;
  const bool tc = wgrad_tc_supported(a);
  if (tc) a.csum = csum;
  {
    ProfScope prof(a, tc, st);
    if (tc) DG_TRY(conv_wgrad_tc(a, st));
    else DG_TRY(conv_wgrad_simt(a, st));
  }
  if (csum && !tc)
    DG_TRY(k_channel_sum(dy, (long long)n * a.H * a.W, L.cout, csum, 1.f, h->act_dt, st));
  return 0;
}

size_t rows_off(depgan_net* h, const ConvL& L, int row) {  // byte offset of sample `row` in an activation of L
  return (size_t)row * h->lvl_h(L.lvl) * h->lvl_w(L.lvl) * L.cout * h->es;
}
const void* off_ptr(const void* p, size_t bytes) { return reinterpret_cast<const char*>(p) + bytes; }
void* off_ptr(void* p, size_t bytes) { return reinterpret_cast<char*>(p) + bytes; }

// Input tensor of critic conv i for sample `row`.
const void* critic_in(depgan_net* d, int i, int row) {
  const int prev = i - 1;
  const ConvL& P = d->c_conv[prev];
  const size_t px = (size_t)d->lvl_h(P.lvl) * d->lvl_w(P.lvl);
  if (prev == 1 || prev == 3 || prev == 5 || prev == 7)
    return off_ptr((const void*)d->c_pool[prev / 2], (size_t)row * px / 4 * P.cout * d->es);
  return off_ptr((const void*)d->c_act[prev], (size_t)row * px * P.cout * d->es);
}

// Backward sweep of the critic over `rows` samples whose forward activations are in d->c_act / c_pool.
//   go[rows]: dLoss/dD per sample.  Leaves b_l for every conv in T.b and, for rows [g_row0, g_row0+g_rows),
//   dD/dx (times go) as fp32 in T.g_in (row-major from g_row0).
int critic_backward(depgan_net* d, int rows, int g_row0, int g_rows, cudaStream_t st) {
  Train& T = *d->tr;
  const int hw4 = d->lvl_h(4) * d->lvl_w(4);
  DG_TRY(k_critic_head_bwd(d->c_act[10], nullptr, T.go, d->P(d->d9_k), d->P(d->d9_b), d->P(d->dd_k), T.b[10], nullptr,
                           nullptr, nullptr, nullptr, rows, rows, hw4, 256, 0, d->act_dt, st));
  for (int i = 10; i >= 1; --i) {
    const ConvL& L = d->c_conv[i];
    const int prev = i - 1;
    const bool pooled = (prev == 1 || prev == 3 || prev == 5 || prev == 7);
    if (pooled) {
      // gradient wrt the pooled tensor, masked by pooled value > 0 (its argmax position carries the ReLU mask)
      DG_TRY(dgrad_conv(d, L, T.b[i], T.bp[prev / 2], d->act_dt, nullptr, d->c_pool[prev / 2], rows, st));
      const ConvL& P = d->c_conv[prev];
      DG_TRY(k_maxpool_bwd(T.bp[prev / 2], d->c_act[prev], nullptr, T.b[prev], rows, d->lvl_h(P.lvl), d->lvl_w(P.lvl),
                           P.cout, d->act_dt, st));
    } else {
      DG_TRY(dgrad_conv(d, L, T.b[i], T.b[prev], d->act_dt, nullptr, d->c_act[prev], rows, st));
    }
  }
  if (g_rows > 0) {
    const ConvL& L0 = d->c_conv[0];
    DG_TRY(dgrad_conv(d, L0, off_ptr((const void*)T.b[0], rows_off(d, L0, g_row0)), T.g_in, DT_F32, nullptr, nullptr,
                      g_rows, st));
  }
  return 0;
}

}  // namespace

extern "C" {

static int critic_grads_impl(depgan_net* d, int nicg, int which, const float* real2_dev, const float* x1_dev,
                             const float* dem_dev, const float* ep_dev, float* out4_dev, int n, int global_n,
                             cudaStream_t st);

int depgan_critic_grads(depgan_net* d, depgan_net* g, int which, const float* real2_dev, const float* x1_dev,
                        const float* z_dev, const float* ep_dev, float* out4_dev, int n, int global_n, void* stream) {
  DG_REQUIRE(d && g && d->model == DEPGAN_MODEL_CRITIC && g->model == DEPGAN_MODEL_GEN, "critic_grads: bad handles");
  DG_REQUIRE(d->cfg.training && d->tr && d->grads, "critic_grads: the critic handle was not created for training");
  DG_REQUIRE(d->prepared && g->prepared, "critic_grads: call depgan_net_prepare first");
  DG_REQUIRE(n >= 1 && 3 * n <= d->cfg.max_batch && n <= g->cfg.max_batch, "critic_grads: batch too large");
  DG_REQUIRE(g->cfg.nc_out == 1 && g->cfg.H == d->cfg.H && g->cfg.W == d->cfg.W, "critic_grads: shape mismatch");
  DG_REQUIRE(which == 0 || which == 1, "critic_grads: which must be 0 (Y2) or 1 (DEM)");
  cudaStream_t st = (cudaStream_t)stream;
  // 1. G forward in inference mode (TG:533 / 556)
  DG_TRY(gen_forward_impl(g, x1_dev, z_dev, g->dem_f32, n, false, st));
  return critic_grads_impl(d, g->cfg.nicg, which, real2_dev, x1_dev, g->dem_f32, ep_dev, out4_dev, n, global_n, st);
}

// The same graph with the generator output given: the generator's weights do not change during the critic updates of a
// generator iteration (TG:796-829), so a trainer may run all their generator forwards as one batched pass
// (depgan_gen_forward on the concatenated batches) and feed the slices here.
int depgan_critic_grads_dem(depgan_net* d, int nicg, int which, const float* real2_dev, const float* x1_dev,
                            const float* dem_dev, const float* ep_dev, float* out4_dev, int n, int global_n,
                            void* stream) {
  DG_REQUIRE(d && d->model == DEPGAN_MODEL_CRITIC, "critic_grads_dem: bad handle");
  DG_REQUIRE(d->cfg.training && d->tr && d->grads, "critic_grads_dem: the critic handle was not created for training");
  DG_REQUIRE(d->prepared, "critic_grads_dem: call depgan_net_prepare first");
  DG_REQUIRE(n >= 1 && 3 * n <= d->cfg.max_batch && nicg >= 1, "critic_grads_dem: batch too large");
  DG_REQUIRE(which == 0 || which == 1, "critic_grads_dem: which must be 0 (Y2) or 1 (DEM)");
  DG_REQUIRE(real2_dev && x1_dev && dem_dev && ep_dev && out4_dev, "critic_grads_dem: null pointer");
  return critic_grads_impl(d, nicg, which, real2_dev, x1_dev, dem_dev, ep_dev, out4_dev, n, global_n,
                           (cudaStream_t)stream);
}

static int critic_grads_impl(depgan_net* d, int nicg, int which, const float* real2_dev, const float* x1_dev,
                             const float* dem_dev, const float* ep_dev, float* out4_dev, int n, int global_n,
                             cudaStream_t st) {
  if (global_n <= 0) global_n = n;
  Train& T = *d->tr;
  const long long hw = (long long)d->cfg.H * d->cfg.W;
  const float inv_n = 1.0f / (float)global_n;
  const float delta = 10.0f;  // TG:37

  // critic inputs [real | fake | mixed] (TG:534-538, 557)
  DG_TRY(k_critic_inputs(real2_dev, x1_dev, nicg, dem_dev, ep_dev, which, T.batch3, n, hw, DT_F32, st));
  // 2. critic forward on 3n rows
  DG_TRY(critic_forward_impl(d, T.batch3, d->c_out, 3 * n, st));
  DG_CHECK_CUDA(cudaMemsetAsync(out4_dev, 0, 4 * sizeof(float), st));
  DG_TRY(k_segment_sum(d->c_out, n, 2, out4_dev, inv_n, st));  // loss_real, loss_fake (TG:540-541)
  // 3. backward sweep: real rows -1/N, fake rows +1/N, mixed rows 1 (-> g = dD/dx and their b_l)
  DG_TRY(k_fill_go(T.go, n, -inv_n, inv_n, 1.0f, 3 * n, st));
  DG_TRY(critic_backward(d, 3 * n, 2 * n, n, st));
  // 4. gradient penalty value and u = delta * dGP/dg (TG:543-545)
  DG_TRY(k_gp(T.g_in, T.u, out4_dev + 2, n, hw, delta, inv_n, st));
  DG_TRY(k_critic_loss_finalize(out4_dev, delta, st));
  // 5. JVP sweep of the linearised critic on u (penalty rows only)
  {
    const void* vin = T.u;
    int vin_dt = DT_F32, C = 1;
    for (int i = 0; i < 11; ++i) {
      const ConvL& L = d->c_conv[i];
      ConvArgs a{};
      a.in0 = vin; a.C0 = C; a.w = d->P(L.k_off); a.w_tc = L.w_tc;
      a.out = T.v[i];
      a.mask_src = off_ptr((const void*)d->c_act[i], rows_off(d, L, 2 * n));
      a.N = n; a.H = d->lvl_h(L.lvl); a.W = d->lvl_w(L.lvl); a.Cout = L.cout; a.ks = L.ks;
      a.in_dt = vin_dt; a.out_dt = d->act_dt;
      {
        const bool tc = d->act_dt == DT_BF16 && conv_tc_supported(a);
        ProfScope prof(a, tc, st);
        if (tc) DG_TRY(conv_fwd_tc(a, st));
        else DG_TRY(conv_fwd_simt(a, st));
      }
      vin = T.v[i]; vin_dt = d->act_dt; C = L.cout;
      if (i == 1 || i == 3 || i == 5 || i == 7) {
        DG_TRY(k_maxpool_select(T.v[i], a.mask_src, T.vp[i / 2], n, a.H, a.W, C, d->act_dt, st));
        vin = T.vp[i / 2];
      }
    }
  }
  // 6. weight gradients: regular rows (x = activations) + penalty rows (x = JVP activations)
  DG_CHECK_CUDA(cudaMemsetAsync(d->grads, 0, sizeof(float) * d->man.total, st));
  for (int i = 0; i < 11; ++i) {
    const ConvL& L = d->c_conv[i];
    const void* x_reg = i == 0 ? (const void*)T.batch3 : critic_in(d, i, 0);
    const void* x_gp = i == 0 ? (const void*)T.u
                              : ((i - 1 == 1 || i - 1 == 3 || i - 1 == 5 || i - 1 == 7) ? (const void*)T.vp[(i - 1) / 2]
                                                                                         : (const void*)T.v[i - 1]);
    const int x_dt = i == 0 ? DT_F32 : d->act_dt;
    // the bias gradient (channel sums of the regular rows' dy) rides along with the first weight-gradient pass
    DG_TRY(wgrad_conv(d, L, x_reg, L.cin, nullptr, 0, x_dt, T.b[i], d->G(L.k_off), 2 * n, st, d->G(L.b_off)));
    DG_TRY(wgrad_conv(d, L, x_gp, L.cin, nullptr, 0, x_dt, off_ptr((const void*)T.b[i], rows_off(d, L, 2 * n)),
                      d->G(L.k_off), n, st));
  }
  const int hw4 = d->lvl_h(4) * d->lvl_w(4);
  DG_TRY(k_critic_head_bwd(d->c_act[10], T.v[10], T.go, d->P(d->d9_k), d->P(d->d9_b), d->P(d->dd_k), nullptr,
                           d->G(d->d9_k), d->G(d->d9_b), d->G(d->dd_k), d->G(d->dd_b), 3 * n, 2 * n, hw4, 256, 1,
                           d->act_dt, st));
  return 0;
}

// Shared forward part of netG_no_update / netG_train: G forward, both critics, loss partial sums.
static int gen_loss_forward(depgan_net* g, depgan_net* dy2, depgan_net* ddem, const float* x1, const float* real2,
                            const float* z, float thr, float* out6, double* sums_user, int n, int global_n, bool keep,
                            cudaStream_t st) {
  Train& T = *g->tr;
  const long long hw = (long long)g->cfg.H * g->cfg.W;
  DG_TRY(gen_forward_impl(g, x1, z, g->dem_f32, n, keep, st));
  DG_CHECK_CUDA(cudaMemsetAsync(T.sums, 0, 8 * sizeof(double), st));
  const float l1coef = 100.0f / ((float)global_n * (float)hw);  // d(100*mean|.|)/d dem
  DG_TRY(k_gen_loss_sums(g->dem_f32, x1, g->cfg.nicg, real2, thr, T.fake2, T.l1g, l1coef, T.sums, (long long)n * hw, st));
  DG_TRY(critic_forward_impl(dy2, T.fake2, dy2->c_out, n, st));       // D_y2(base + DEM)   TG:574-575
  DG_TRY(critic_forward_impl(ddem, g->dem_f32, ddem->c_out, n, st));  // D_dem(DEM)
  // sums[0], sums[1] = sum of critic scores (double accumulation of n floats: do it with the segment kernel in
  // fp32 into out6 scratch, then widen) -- n is tiny, a single-thread kernel is enough
  DG_TRY(k_scores_to_sums(dy2->c_out, ddem->c_out, n, T.sums, (double)global_n, (double)hw, st));
  DG_TRY(k_gen_loss_finalize(out6, T.sums, st));
  if (sums_user) DG_CHECK_CUDA(cudaMemcpyAsync(sums_user, T.sums, 8 * sizeof(double), cudaMemcpyDeviceToDevice, st));
  return 0;
}

static int check_gen_handles(depgan_net* g, depgan_net* dy2, depgan_net* ddem, int n, bool grads) {
  DG_REQUIRE(g && dy2 && ddem && g->model == DEPGAN_MODEL_GEN && dy2->model == DEPGAN_MODEL_CRITIC &&
                 ddem->model == DEPGAN_MODEL_CRITIC, "gen step: bad handles");
  DG_REQUIRE(g->cfg.training && g->tr && dy2->cfg.training && dy2->tr && ddem->cfg.training && ddem->tr,
             "gen step: handles were not created for training");
  DG_REQUIRE(g->prepared && dy2->prepared && ddem->prepared, "gen step: call depgan_net_prepare first");
  DG_REQUIRE(g->cfg.nc_out == 1, "gen step: the DEP-GAN generator has one output channel");
  DG_REQUIRE(n >= 1 && n <= g->cfg.max_batch && n <= dy2->cfg.max_batch && n <= ddem->cfg.max_batch,
             "gen step: batch too large");
  DG_REQUIRE(!grads || g->grads, "gen step: no gradient buffer");
  return 0;
}

int depgan_gen_eval(depgan_net* g, depgan_net* dy2, depgan_net* ddem, const float* x1_dev, const float* real2_dev,
                    const float* z_dev, float thr, float* out6_dev, double* sums_dev, int n, int global_n,
                    void* stream) {
  DG_TRY(check_gen_handles(g, dy2, ddem, n, false));
  if (global_n <= 0) global_n = n;
  return gen_loss_forward(g, dy2, ddem, x1_dev, real2_dev, z_dev, thr, out6_dev, sums_dev, n, global_n, false,
                          (cudaStream_t)stream);
}

int depgan_gen_loss_finalize(float* out6_dev, const double* sums_dev, void* stream) {
  DG_REQUIRE(out6_dev && sums_dev, "gen_loss_finalize: null pointer");
  return k_gen_loss_finalize(out6_dev, sums_dev, (cudaStream_t)stream);
}

// netG_no_update for k noise candidates of ONE batch in a single pass (TG:868-874: the ten evaluations of a generator
// iteration see the same x / real_2tp and differ only in the noise): the batch is replicated k times, the generator and
// the two critics run once on k * n rows, the loss sums are kept per candidate.  Any prepared handles with
// max_batch >= k * n serve (inference handles that share the training networks' parameter buffers: api.py), so the
// training handles' backward buffers are not multiplied by k.  Slices never mix, so every candidate's losses equal those
// of the one-by-one evaluation.  scratch: k * n * H * W * (nicg + 1) floats.
int depgan_gen_eval_multi(depgan_net* g, depgan_net* dy2, depgan_net* ddem, const float* x1_dev, const float* real2_dev,
                          const float* z_all_dev, float thr, float* out6_all_dev, double* sums_all_dev, float* scratch_dev,
                          int n, int k, int global_n, void* stream) {
  DG_REQUIRE(g && dy2 && ddem && g->model == DEPGAN_MODEL_GEN && dy2->model == DEPGAN_MODEL_CRITIC &&
                 ddem->model == DEPGAN_MODEL_CRITIC, "gen_eval_multi: bad handles");
  DG_REQUIRE(g->prepared && dy2->prepared && ddem->prepared, "gen_eval_multi: call depgan_net_prepare first");
  DG_REQUIRE(g->cfg.nc_out == 1 && g->cfg.H == dy2->cfg.H && g->cfg.W == dy2->cfg.W && g->cfg.H == ddem->cfg.H &&
                 g->cfg.W == ddem->cfg.W, "gen_eval_multi: shape mismatch");
  DG_REQUIRE(n >= 1 && k >= 1 && (long long)n * k <= g->cfg.max_batch && (long long)n * k <= dy2->cfg.max_batch &&
                 (long long)n * k <= ddem->cfg.max_batch, "gen_eval_multi: k * n exceeds a handle's max_batch");
  DG_REQUIRE(x1_dev && real2_dev && z_all_dev && out6_all_dev && sums_all_dev && scratch_dev, "gen_eval_multi: null pointer");
  if (global_n <= 0) global_n = n;
  cudaStream_t st = (cudaStream_t)stream;
  const long long hw = (long long)g->cfg.H * g->cfg.W;
  const int nicg = g->cfg.nicg;
  float* x_rep = scratch_dev;
  float* fake2 = scratch_dev + (size_t)k * n * hw * nicg;
  for (int j = 0; j < k; ++j)
    DG_CHECK_CUDA(cudaMemcpyAsync(x_rep + (size_t)j * n * hw * nicg, x1_dev, sizeof(float) * (size_t)n * hw * nicg,
                                  cudaMemcpyDeviceToDevice, st));
  DG_TRY(gen_forward_impl(g, x_rep, z_all_dev, g->dem_f32, n * k, false, st));
  DG_CHECK_CUDA(cudaMemsetAsync(sums_all_dev, 0, sizeof(double) * 8 * k, st));
  for (int j = 0; j < k; ++j)
    DG_TRY(k_gen_loss_sums(g->dem_f32 + (size_t)j * n * hw, x1_dev, nicg, real2_dev, thr, fake2 + (size_t)j * n * hw, nullptr,
                           0.f, sums_all_dev + 8 * j, (long long)n * hw, st));
  DG_TRY(critic_forward_impl(dy2, fake2, dy2->c_out, n * k, st));        // D_y2(base + DEM)   TG:574-575
  DG_TRY(critic_forward_impl(ddem, g->dem_f32, ddem->c_out, n * k, st));  // D_dem(DEM)
  DG_TRY(k_scores_to_sums(dy2->c_out, ddem->c_out, n, sums_all_dev, (double)global_n, (double)hw, st, k));
  return k_gen_loss_finalize(out6_all_dev, sums_all_dev, st, k);
}

// Finalises k candidates after their sums [0..5] were summed over the ranks (the constants [6], [7] are rewritten).
int depgan_gen_loss_finalize_multi(float* out6_all_dev, double* sums_all_dev, int k, int global_n, long long hw,
                                   void* stream) {
  DG_REQUIRE(out6_all_dev && sums_all_dev && k >= 1 && global_n >= 1 && hw >= 1, "gen_loss_finalize_multi: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  DG_TRY(k_loss_consts(sums_all_dev, (double)global_n, (double)hw, k, st));
  return k_gen_loss_finalize(out6_all_dev, sums_all_dev, st, k);
}

int depgan_gen_grads(depgan_net* g, depgan_net* dy2, depgan_net* ddem, const float* x1_dev, const float* real2_dev,
                     const float* z_dev, float thr, float* out6_dev, double* sums_dev, int n, int global_n,
                     void* stream) {
  DG_TRY(check_gen_handles(g, dy2, ddem, n, true));
  if (global_n <= 0) global_n = n;
  cudaStream_t st = (cudaStream_t)stream;
  Train& T = *g->tr;
  const depgan_cfg& c = g->cfg;
  const long long hw = (long long)c.H * c.W;
  const float inv_n = 1.0f / (float)global_n;
  const int f = FIRST_FM;
  DG_TRY(gen_loss_forward(g, dy2, ddem, x1_dev, real2_dev, z_dev, thr, out6_dev, sums_dev, n, global_n, true, st));

  // ---- dLoss/dDEM: -(1/N) dD_y2/dx - (1/N) dD_dem/dx + 100/(N*HW) sign(DEM - realDEM)   (TG:592) ----
  DG_TRY(k_fill_go(dy2->tr->go, n, -inv_n, -inv_n, -inv_n, n, st));
  DG_TRY(critic_backward(dy2, n, 0, n, st));
  DG_TRY(k_fill_go(ddem->tr->go, n, -inv_n, -inv_n, -inv_n, n, st));
  DG_TRY(critic_backward(ddem, n, 0, n, st));

  DG_CHECK_CUDA(cudaMemsetAsync(g->grads, 0, sizeof(float) * g->man.total, st));
  DG_CHECK_CUDA(cudaMemsetAsync(T.d_film, 0, sizeof(float) * (size_t)n * g->film_total, st));
  // tanh' and the 1x1 head (TG:494-495): leaves b of conv2d_gen_17 in d_o[6]
  DG_TRY(k_gen_head_bwd(dy2->tr->g_in, ddem->tr->g_in, T.l1g, g->dem_f32, g->act_o[6], g->P(g->g_seg.k_off), T.d_o[6],
                        g->G(g->g_seg.k_off), g->G(g->g_seg.b_off), (long long)n * hw, f, g->act_dt, st));

  auto finish_conv = [&](const ConvL& L, const void* b, long long rows) -> int {  // BN / bias / scale of one conv
    DG_CHECK_CUDA(cudaMemsetAsync(T.sum_dy, 0, sizeof(float) * L.cout, st));
    DG_TRY(k_channel_sum(b, rows, L.cout, T.sum_dy, 1.f, g->act_dt, st));
    return k_param_grads(g->P(L.k_off), g->G(L.k_off), g->G(L.k_off), L.scale, L.inv_std, g->P(L.b_off),
                         g->P(L.mu_off), T.sum_dy, g->G(L.g_off), g->G(L.be_off), g->G(L.b_off), L.ks * L.ks * L.cin,
                         L.cout, 0, 0, st);
  };

  for (int bi = 6; bi >= 0; --bi) {
    const int w = f * GEN_MULT[bi], lvl = GEN_LVL[bi];
    const int H = g->lvl_h(lvl), W = g->lvl_w(lvl);
    const long long px = (long long)n * H * W;
    const ConvL &Lin = g->g_in[bi], &Lno = g->g_noise[bi], &Lout = g->g_out[bi];
    // conv_out: b_out = d_o masked by o > 0 (the head kernel already masked block 6)
    if (bi != 6) DG_TRY(k_add_mask(T.d_o[bi], nullptr, g->act_o[bi], T.d_o[bi], px * w, g->act_dt, st));
    DG_TRY(wgrad_conv(g, Lout, g->act_r[bi], w, nullptr, 0, g->act_dt, T.d_o[bi], g->G(Lout.k_off), n, st));
    DG_TRY(finish_conv(Lout, T.d_o[bi], px));
    DG_TRY(dgrad_conv(g, Lout, T.d_o[bi], T.d_r, g->act_dt, nullptr, nullptr, n, st));
    // FiLM + residual (TG:403-407): d_y, d gamma(z), d beta(z)
    DG_TRY(k_film_bwd(T.d_r, g->act_y[bi], g->film_out + g->head_off[2 * bi], g->film_out + g->head_off[2 * bi + 1],
                      g->film_total, T.d_y, T.d_film + g->head_off[2 * bi], T.d_film + g->head_off[2 * bi + 1], n,
                      H * W, w, g->act_dt, st));
    DG_TRY(wgrad_conv(g, Lno, g->act_a[bi], w, nullptr, 0, g->act_dt, T.d_y, g->G(Lno.k_off), n, st));
    DG_TRY(finish_conv(Lno, T.d_y, px));
    // b_in = (d_r + dgrad_noise(d_y)) masked by a > 0
    DG_TRY(dgrad_conv(g, Lno, T.d_y, T.b_in, g->act_dt, T.d_r, g->act_a[bi], n, st));
    // conv_in
    const void *x0, *x1 = nullptr;
    int C0, C1 = 0, x_dt = g->act_dt;
    if (bi == 0) { x0 = x1_dev; C0 = c.nicg; x_dt = DT_F32; }
    else if (bi <= 3) { x0 = g->act_pool[bi - 1]; C0 = f * GEN_MULT[bi - 1]; }
    else { x0 = g->act_up[bi - 1]; C0 = f * GEN_MULT[bi - 1]; x1 = g->act_o[6 - bi]; C1 = f * GEN_MULT[6 - bi]; }
    DG_TRY(wgrad_conv(g, Lin, x0, C0, x1, C1, x_dt, T.b_in, g->G(Lin.k_off), n, st));
    DG_TRY(finish_conv(Lin, T.b_in, px));
    if (bi == 0) break;
    DG_TRY(dgrad_conv(g, Lin, T.b_in, T.d_in, g->act_dt, nullptr, nullptr, n, st));
    if (bi <= 3) {
      // input was maxpool(o[bi-1]); blocks 0..2 already hold their skip gradient in d_o[bi-1]
      DG_TRY(k_maxpool_bwd(T.d_in, g->act_o[bi - 1], T.d_o[bi - 1], T.d_o[bi - 1], n, 2 * H, 2 * W, C0, g->act_dt, st));
    } else {
      // concat [deconv_out (C0), skip (C1)]: skip slice -> d_o[6-bi]; up slice -> transposed-conv backward
      DG_TRY(k_slice_add(T.d_in, C0 + C1, C0, nullptr, T.d_o[6 - bi], px, C1, g->act_dt, st));
      const ConvL& Ld = g->g_dec[bi - 4];
      const int Hh = H / 2, Wh = W / 2;
      const long long pxh = (long long)n * Hh * Wh;
      DG_TRY(k_s2d_mask(T.d_in, C0 + C1, g->act_up[bi - 1], T.s2d, n, Hh, Wh, C0, g->act_dt, st));
      // raw weight gradient G[ci][(ab,co)] = sum_p o[p,ci] * s2d[p,(ab,co)]  (a 1x1 wgrad)
      DG_CHECK_CUDA(cudaMemsetAsync(T.g_raw, 0, sizeof(float) * (size_t)Ld.cin * 4 * Ld.cout, st));
      {
        WgradArgs wa{};
        wa.x0 = g->act_o[bi - 1]; wa.C0 = Ld.cin; wa.dy = T.s2d; wa.dw = T.g_raw;
        wa.N = n; wa.H = Hh; wa.W = Wh; wa.Cout = 4 * Ld.cout; wa.ks = 1;
        wa.x_dt = g->act_dt; wa.dy_dt = g->act_dt; wa.alpha = 1.f;
        const bool tc = wgrad_tc_supported(wa);
        ProfScope prof(wa, tc, st);
        if (tc) DG_TRY(conv_wgrad_tc(wa, st));
        else DG_TRY(conv_wgrad_simt(wa, st));
      }
      DG_CHECK_CUDA(cudaMemsetAsync(T.sum_dy, 0, sizeof(float) * 4 * Ld.cout, st));
      DG_TRY(k_channel_sum(T.s2d, pxh, 4 * Ld.cout, T.sum_dy, 1.f, g->act_dt, st));
      DG_TRY(k_param_grads(g->P(Ld.k_off), T.g_raw, g->G(Ld.k_off), Ld.scale, Ld.inv_std, g->P(Ld.b_off),
                           g->P(Ld.mu_off), T.sum_dy, g->G(Ld.g_off), g->G(Ld.be_off), g->G(Ld.b_off), 0, Ld.cout, 1,
                           Ld.cin, st));
      // data gradient: d_o[bi-1][p,ci] = sum_(ab,co) s2d[p,(ab,co)] * s[co] * Wd[ab][co][ci]  (a 1x1 conv)
      {
        ConvArgs a{};
        a.in0 = T.s2d; a.C0 = 4 * Ld.cout; a.w = Ld.w_dg; a.w_tc = Ld.w_dg_tc; a.out = T.d_o[bi - 1];
        a.N = n; a.H = Hh; a.W = Wh; a.Cout = Ld.cin; a.ks = 1; a.in_dt = g->act_dt; a.out_dt = g->act_dt;
        if (g->act_dt == DT_BF16 && conv_tc_supported(a)) DG_TRY(conv_fwd_tc(a, st));
        else DG_TRY(conv_fwd_simt(a, st));
      }
    }
  }

  // ---- FiLM noise MLP backward (TG:353-395) ----
  {
    FilmMlpArgs fa{};
    fa.z = z_dev;
    fa.k0 = g->P(g->d_f0.k_off); fa.s0 = g->d_f0.scale; fa.t0 = g->d_f0.shift;
    fa.k1 = g->P(g->d_f1.k_off); fa.s1 = g->d_f1.scale; fa.t1 = g->d_f1.shift;
    fa.head_w = g->dev_head_w; fa.head_s = g->dev_head_s; fa.head_t = g->dev_head_t;
    fa.head_c = g->dev_head_c; fa.head_off = g->dev_head_off;
    fa.n_heads = 14; fa.total_c = g->film_total;
    fa.h1 = g->film_h1; fa.h2 = g->film_h2; fa.out = g->film_out;
    fa.N = n; fa.L = c.noise_len; fa.F = f;
    if (!T.heads_uploaded) {
      float* hp[16] = {};
      for (int i = 0; i < 14; ++i) hp[i] = g->G(g->d_head[i].k_off);
      DG_CHECK_CUDA(cudaMemcpyAsync(T.dev_dw_heads, hp, sizeof(hp), cudaMemcpyHostToDevice, st));
      DG_CHECK_CUDA(cudaStreamSynchronize(st));
      T.heads_uploaded = true;
    }
    DG_CHECK_CUDA(cudaMemsetAsync(T.sum_d1, 0, 64 * sizeof(float), st));
    DG_CHECK_CUDA(cudaMemsetAsync(T.sum_d0, 0, 64 * sizeof(float), st));
    DG_TRY(k_film_mlp_bwd(fa, T.d_film, T.dev_dw_heads, T.sum_d, T.d_h2, g->G(g->d_f1.k_off), T.sum_d1,
                          g->G(g->d_f0.k_off), T.sum_d0, st));
    const int K = c.noise_len * f;
    for (int i = 0; i < 14; ++i) {
      const DenseL& D = g->d_head[i];
      DG_TRY(k_param_grads(g->P(D.k_off), g->G(D.k_off), g->G(D.k_off), D.scale, D.inv_std, g->P(D.b_off),
                           g->P(D.mu_off), T.sum_d + g->head_off[i], g->G(D.g_off), g->G(D.be_off), g->G(D.b_off), K,
                           D.out, 0, 0, st));
    }
    const DenseL &D1 = g->d_f1, &D0 = g->d_f0;
    DG_TRY(k_param_grads(g->P(D1.k_off), g->G(D1.k_off), g->G(D1.k_off), D1.scale, D1.inv_std, g->P(D1.b_off),
                         g->P(D1.mu_off), T.sum_d1, g->G(D1.g_off), g->G(D1.be_off), g->G(D1.b_off), f, f, 0, 0, st));
    DG_TRY(k_param_grads(g->P(D0.k_off), g->G(D0.k_off), g->G(D0.k_off), D0.scale, D0.inv_std, g->P(D0.b_off),
                         g->P(D0.mu_off), T.sum_d0, g->G(D0.g_off), g->G(D0.be_off), g->G(D0.b_off), 1, f, 0, 0, st));
  }
  return 0;
}

}  // extern "C"
