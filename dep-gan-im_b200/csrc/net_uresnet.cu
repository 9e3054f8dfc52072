// DEP-UResNet supervised train step (TU:291-428 compiled at TU:427, my_network.fit at TU:602-606) in Keras
// *training* phase: every BatchNormalization normalises with the statistics of the current batch (and updates its
// moving statistics, momentum 0.99), Dropout(0.25) `do_gen_1` follows conv2d_gen_10 (TU:388), the loss is
// categorical cross-entropy on the softmax output.  Forward and backward are written out layer by layer; the
// convolutions are the same tcgen05 / CUDA-core kernels as everywhere else (raw conv + bias, BN as separate
// bandwidth-bound passes because the statistics need the whole batch before anything can be normalised).
#include "net.h"

namespace {

constexpr float MOM = 0.99f;          // Keras BatchNormalization default momentum (TU:26 states the same value)
constexpr float KEEP_SCALE = 1.0f / 0.75f;

struct Ctx {
  depgan_net* g;
  Train* T;
  int n;
  cudaStream_t st;
};

int raw_conv(Ctx& c, const ConvL& L, const void* in0, int C0, const void* in1, int C1, int in_dt, void* raw) {
  ConvArgs a{};
  a.in0 = in0; a.in1 = in1; a.C0 = C0; a.C1 = C1;
  a.w = c.g->P(L.k_off); a.w_tc = L.w_tc; a.shift = c.g->P(L.b_off); a.out = raw;
  a.N = c.n; a.H = c.g->lvl_h(L.lvl); a.W = c.g->lvl_w(L.lvl); a.Cout = L.cout; a.ks = L.ks;
  a.in_dt = in_dt; a.out_dt = c.g->act_dt;
  const bool tc = c.g->act_dt == DT_BF16 && conv_tc_supported(a);
  ProfScope prof(a, tc, c.st);
  return tc ? conv_fwd_tc(a, c.st) : conv_fwd_simt(a, c.st);
}

int bn_fwd(Ctx& c, const ConvL& L, BnState& s, const void* raw, long long rows, void* out, void* y_out, int relu,
           const float* fg, const float* fb, long long rows_per_sample, const void* res, const unsigned char* keep) {
  depgan_net* g = c.g;
  DG_TRY(k_bn_stats(raw, rows, L.cout, c.T->bn_sums, s.mean, s.inv_std, g->P(L.mu_off), g->P(L.var_off), MOM,
                    g->act_dt, c.st));
  return k_bn_apply(raw, s.mean, s.inv_std, g->P(L.g_off), g->P(L.be_off), out, y_out, rows, L.cout, relu, fg, fb,
                    g->film_total, rows_per_sample, res, keep, KEEP_SCALE, g->act_dt, c.st);
}

int dgrad(Ctx& c, const ConvL& L, const void* dy, void* dx, const void* add_src) {
  ConvArgs a{};
  a.in0 = dy; a.C0 = L.cout; a.w = L.w_dg; a.w_tc = L.w_dg_tc; a.out = dx; a.add_src = add_src;
  a.N = c.n; a.H = c.g->lvl_h(L.lvl); a.W = c.g->lvl_w(L.lvl); a.Cout = L.cin; a.ks = L.ks;
  a.in_dt = c.g->act_dt; a.out_dt = c.g->act_dt;
  const bool tc = c.g->act_dt == DT_BF16 && conv_tc_supported(a);
  ProfScope prof(a, tc, c.st);
  return tc ? conv_fwd_tc(a, c.st) : conv_fwd_simt(a, c.st);
}

int wgrad(Ctx& c, const ConvL& L, const void* x0, int C0, const void* x1, int C1, int x_dt, const void* dy) {
  depgan_net* g = c.g;
  WgradArgs a{};
  a.x0 = x0; a.x1 = x1; a.C0 = C0; a.C1 = C1; a.dy = dy; a.dw = g->G(L.k_off);
  a.N = c.n; a.H = g->lvl_h(L.lvl); a.W = g->lvl_w(L.lvl); a.Cout = L.cout; a.ks = L.ks;
  a.x_dt = x_dt; a.dy_dt = g->act_dt; a.alpha = 1.f;
  const bool tc = wgrad_tc_supported(a);
  {
    ProfScope prof(a, tc, c.st);
    DG_TRY(tc ? conv_wgrad_tc(a, c.st) : conv_wgrad_simt(a, c.st));
  }
  return k_channel_sum(dy, (long long)c.n * a.H * a.W, L.cout, g->G(L.b_off), 1.f, g->act_dt, c.st);
}

// BN backward of one conv layer: dy (grad wrt the BN output, optionally through the ReLU of `relu_out`) -> d_raw
int bn_bwd(Ctx& c, const ConvL& L, BnState& s, const void* dy, const void* raw, const void* relu_out, long long rows,
           void* d_raw) {
  depgan_net* g = c.g;
  return k_bn_bwd(dy, raw, relu_out, s.mean, s.inv_std, g->P(L.g_off), c.T->bn_red, d_raw, g->G(L.g_off),
                  g->G(L.be_off), rows, L.cout, g->act_dt, c.st);
}

int dense_bn_fwd(Ctx& c, const DenseL& D, BnState& s, const float* X, int rows, float* pre, float* out, int relu) {
  depgan_net* g = c.g;
  DG_TRY(k_dense_fwd(X, g->P(D.k_off), g->P(D.b_off), pre, rows, D.in, D.out, c.st));
  DG_TRY(k_bn_stats(pre, rows, D.out, c.T->bn_sums, s.mean, s.inv_std, g->P(D.mu_off), g->P(D.var_off), MOM, DT_F32,
                    c.st));
  return k_bn_apply(pre, s.mean, s.inv_std, g->P(D.g_off), g->P(D.be_off), out, nullptr, rows, D.out, relu, nullptr,
                    nullptr, 0, 1, nullptr, nullptr, 1.f, DT_F32, c.st);
}

}  // namespace

extern "C" {

int depgan_set_sync_hook(depgan_allreduce_fn fn, void* user, int world) { return bn_set_sync_hook(fn, user, world); }

int depgan_cce_loss(const float* prob_dev, const float* onehot_dev, float* dseg_scratch_dev, float* loss_dev,
                    long long npix, int nc, float inv_total, void* stream) {
  DG_REQUIRE(prob_dev && onehot_dev && dseg_scratch_dev && loss_dev, "cce_loss: null pointer");
  return k_softmax_cce(prob_dev, onehot_dev, dseg_scratch_dev, loss_dev, npix, nc, inv_total, (cudaStream_t)stream);
}

int depgan_uresnet_grads(depgan_net* g, const float* x_dev, const float* z_dev, const float* onehot_dev,
                         const unsigned char* drop_keep_dev, float* loss_dev, int n, void* stream) {
  DG_REQUIRE(g && g->model == DEPGAN_MODEL_GEN && g->cfg.training == 2 && g->tr && g->grads,
             "uresnet_grads: the handle must be a generator created with training = 2");
  DG_REQUIRE(g->prepared, "uresnet_grads: call depgan_net_prepare first");
  DG_REQUIRE(n >= 2 && n <= g->cfg.max_batch, "uresnet_grads: batch must be 2..max_batch (batch statistics)");
  DG_REQUIRE(x_dev && z_dev && onehot_dev && drop_keep_dev && loss_dev, "uresnet_grads: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  Train& T = *g->tr;
  Ctx c{g, &T, n, st};
  const depgan_cfg& cfg = g->cfg;
  const int f = FIRST_FM, L = cfg.noise_len, nc = cfg.nc_out;
  const long long hw = (long long)cfg.H * cfg.W;

  // ================= forward =================
  // noise MLP (TU:353-395) with batch-statistic BN: 3-D BN over (N, L) per feature, 2-D BN over N for the heads
  DG_TRY(dense_bn_fwd(c, g->d_f0, T.bn_f0, z_dev, n * L, T.pre0, g->film_h1, 1));
  DG_TRY(dense_bn_fwd(c, g->d_f1, T.bn_f1, g->film_h1, n * L, T.pre1, g->film_h2, 1));
  for (int i = 0; i < 14; ++i) {
    const DenseL& D = g->d_head[i];
    DG_TRY(dense_bn_fwd(c, D, T.bn_head[i], g->film_h2, n, T.raw_head[i], T.tmp_c1, 0));
    DG_TRY(k_strided_copy(T.tmp_c1, D.out, 0, g->film_out, g->film_total, g->head_off[i], n, D.out, st));
  }
  const void* in0 = x_dev;
  const void* in1 = nullptr;
  int C0 = cfg.nicg, C1 = 0, in_dt = DT_F32;
  for (int bi = 0; bi < 7; ++bi) {
    const int w = f * GEN_MULT[bi], lvl = GEN_LVL[bi];
    const long long px = (long long)n * g->lvl_h(lvl) * g->lvl_w(lvl), pps = (long long)g->lvl_h(lvl) * g->lvl_w(lvl);
    DG_TRY(raw_conv(c, g->g_in[bi], in0, C0, in1, C1, in_dt, T.raw_a[bi]));
    DG_TRY(bn_fwd(c, g->g_in[bi], T.bn_in[bi], T.raw_a[bi], px, g->act_a[bi], nullptr, 1, nullptr, nullptr, pps,
                  nullptr, bi == 4 ? drop_keep_dev : nullptr));                       // do_gen_1 (TU:388)
    DG_TRY(raw_conv(c, g->g_noise[bi], g->act_a[bi], w, nullptr, 0, g->act_dt, T.raw_y[bi]));
    DG_TRY(bn_fwd(c, g->g_noise[bi], T.bn_no[bi], T.raw_y[bi], px, g->act_r[bi], g->act_y[bi], 0,
                  g->film_out + g->head_off[2 * bi], g->film_out + g->head_off[2 * bi + 1], pps, g->act_a[bi],
                  nullptr));
    DG_TRY(raw_conv(c, g->g_out[bi], g->act_r[bi], w, nullptr, 0, g->act_dt, T.raw_o[bi]));
    DG_TRY(bn_fwd(c, g->g_out[bi], T.bn_out[bi], T.raw_o[bi], px, g->act_o[bi], nullptr, 1, nullptr, nullptr, pps,
                  nullptr, nullptr));
    if (bi < 3) {
      DG_TRY(k_maxpool_fwd(g->act_o[bi], g->act_pool[bi], n, g->lvl_h(lvl), g->lvl_w(lvl), w, g->act_dt, st));
      in0 = g->act_pool[bi]; C0 = w; in1 = nullptr; C1 = 0;
    } else if (bi < 6) {
      const ConvL& Ld = g->g_dec[bi - 3];
      const int H = g->lvl_h(lvl), W = g->lvl_w(lvl);
      ConvArgs a{};
      a.in0 = g->act_o[bi]; a.C0 = w; a.w_tc = Ld.w_tc; a.shift = g->P(Ld.b_off); a.out = T.raw_up[bi];
      a.deconv = 1; a.N = n; a.H = H; a.W = W; a.Cout = w; a.ks = 1; a.in_dt = g->act_dt; a.out_dt = g->act_dt;
      if (g->act_dt == DT_BF16 && conv_tc_supported(a)) {
        ProfScope prof(a, true, st);
        DG_TRY(conv_fwd_tc(a, st));
      } else {
        DG_TRY(k_deconv_fwd(g->act_o[bi], g->P(Ld.k_off), nullptr, g->P(Ld.b_off), T.raw_up[bi], n, H, W, w, w,
                            g->act_dt, 0, st));
      }
      DG_TRY(bn_fwd(c, Ld, T.bn_dec[bi - 3], T.raw_up[bi], px * 4, g->act_up[bi], nullptr, 1, nullptr, nullptr,
                    pps * 4, nullptr, nullptr));
      in0 = g->act_up[bi]; C0 = w; in1 = g->act_o[5 - bi]; C1 = f * GEN_MULT[5 - bi];
    }
    in_dt = g->act_dt;
  }
  DG_TRY(k_head_fwd(g->act_o[6], g->P(g->g_seg.k_off), g->P(g->g_seg.b_off), g->dem_f32, (long long)n * hw, f, nc, 1,
                    g->act_dt, st));
  DG_CHECK_CUDA(cudaMemsetAsync(loss_dev, 0, sizeof(float), st));
  // data parallel: the loss and its gradient seed are scaled by the GLOBAL pixel count, so summing the ranks' losses
  // and gradient buckets gives the global-batch mean
  DG_TRY(k_softmax_cce(g->dem_f32, onehot_dev, T.dseg, loss_dev, (long long)n * hw, nc,
                       1.0f / ((float)n * (float)bn_sync_world() * (float)hw),
                       st));

  // ================= backward =================
  DG_CHECK_CUDA(cudaMemsetAsync(g->grads, 0, sizeof(float) * g->man.total, st));
  DG_CHECK_CUDA(cudaMemsetAsync(T.d_film, 0, sizeof(float) * (size_t)n * g->film_total, st));
  DG_TRY(k_head_bwd_multi(T.dseg, g->act_o[6], g->P(g->g_seg.k_off), T.d_o[6], g->G(g->g_seg.k_off),
                          g->G(g->g_seg.b_off), (long long)n * hw, f, nc, g->act_dt, st));
  for (int bi = 6; bi >= 0; --bi) {
    const int w = f * GEN_MULT[bi], lvl = GEN_LVL[bi];
    const int H = g->lvl_h(lvl), W = g->lvl_w(lvl);
    const long long px = (long long)n * H * W;
    const ConvL &Lin = g->g_in[bi], &Lno = g->g_noise[bi], &Lout = g->g_out[bi];
    // conv_out: relu + BN backward -> d_raw (d_y), weight/bias grads, data grad -> d_r
    DG_TRY(bn_bwd(c, Lout, T.bn_out[bi], T.d_o[bi], T.raw_o[bi], g->act_o[bi], px, T.d_y));
    DG_TRY(wgrad(c, Lout, g->act_r[bi], w, nullptr, 0, g->act_dt, T.d_y));
    DG_TRY(dgrad(c, Lout, T.d_y, T.d_r, nullptr));
    // FiLM + residual: d(BN output of conv_noise), d gamma(z), d beta(z)
    DG_TRY(k_film_bwd(T.d_r, g->act_y[bi], g->film_out + g->head_off[2 * bi], g->film_out + g->head_off[2 * bi + 1],
                      g->film_total, T.d_y, T.d_film + g->head_off[2 * bi], T.d_film + g->head_off[2 * bi + 1], n,
                      H * W, w, g->act_dt, st));
    DG_TRY(bn_bwd(c, Lno, T.bn_no[bi], T.d_y, T.raw_y[bi], nullptr, px, T.b_in));
    DG_TRY(wgrad(c, Lno, g->act_a[bi], w, nullptr, 0, g->act_dt, T.b_in));
    DG_TRY(dgrad(c, Lno, T.b_in, T.d_y, T.d_r));                  // d_a = d_r + dgrad_noise(.)
    if (bi == 4) DG_TRY(k_dropout_bwd(T.d_y, drop_keep_dev, KEEP_SCALE, T.d_y, px * w, g->act_dt, st));
    // conv_in (the ReLU mask of the post-dropout activation is equivalent: dropped positions carry zero gradient)
    DG_TRY(bn_bwd(c, Lin, T.bn_in[bi], T.d_y, T.raw_a[bi], g->act_a[bi], px, T.b_in));
    const void *x0, *x1 = nullptr;
    int c0, c1 = 0, x_dt = g->act_dt;
    if (bi == 0) { x0 = x_dev; c0 = cfg.nicg; x_dt = DT_F32; }
    else if (bi <= 3) { x0 = g->act_pool[bi - 1]; c0 = f * GEN_MULT[bi - 1]; }
    else { x0 = g->act_up[bi - 1]; c0 = f * GEN_MULT[bi - 1]; x1 = g->act_o[6 - bi]; c1 = f * GEN_MULT[6 - bi]; }
    DG_TRY(wgrad(c, Lin, x0, c0, x1, c1, x_dt, T.b_in));
    if (bi == 0) break;
    DG_TRY(dgrad(c, Lin, T.b_in, T.d_in, nullptr));
    if (bi <= 3) {
      DG_TRY(k_maxpool_bwd(T.d_in, g->act_o[bi - 1], T.d_o[bi - 1], T.d_o[bi - 1], n, 2 * H, 2 * W, c0, g->act_dt, st));
    } else {
      DG_TRY(k_slice_add(T.d_in, c0 + c1, c0, nullptr, T.d_o[6 - bi], px, c1, g->act_dt, st));
      const ConvL& Ld = g->g_dec[bi - 4];
      const int Hh = H / 2, Wh = W / 2;
      const long long pxh = (long long)n * Hh * Wh;
      // up = relu(bn(deconv)): slice its gradient out of the concat gradient, BN/ReLU backward at full resolution
      DG_TRY(k_slice_add(T.d_in, c0 + c1, 0, nullptr, T.tmp_up, px, c0, g->act_dt, st));
      DG_TRY(k_bn_bwd(T.tmp_up, T.raw_up[bi - 1], g->act_up[bi - 1], T.bn_dec[bi - 4].mean, T.bn_dec[bi - 4].inv_std,
                      g->P(Ld.g_off), T.bn_red, T.tmp_up, g->G(Ld.g_off), g->G(Ld.be_off), px, c0, g->act_dt, st));
      DG_TRY(k_s2d_mask(T.tmp_up, c0, nullptr, T.s2d, n, Hh, Wh, c0, g->act_dt, st));
      DG_CHECK_CUDA(cudaMemsetAsync(T.g_raw, 0, sizeof(float) * (size_t)Ld.cin * 4 * Ld.cout, st));
      {
        WgradArgs wa{};
        wa.x0 = g->act_o[bi - 1]; wa.C0 = Ld.cin; wa.dy = T.s2d; wa.dw = T.g_raw;
        wa.N = n; wa.H = Hh; wa.W = Wh; wa.Cout = 4 * Ld.cout; wa.ks = 1;
        wa.x_dt = g->act_dt; wa.dy_dt = g->act_dt; wa.alpha = 1.f;
        const bool tc = wgrad_tc_supported(wa);
        ProfScope prof(wa, tc, st);
        DG_TRY(tc ? conv_wgrad_tc(wa, st) : conv_wgrad_simt(wa, st));
      }
      DG_CHECK_CUDA(cudaMemsetAsync(T.sum_dy, 0, sizeof(float) * 4 * Ld.cout, st));
      DG_TRY(k_channel_sum(T.s2d, pxh, 4 * Ld.cout, T.sum_dy, 1.f, g->act_dt, st));
      DG_TRY(k_param_grads(g->P(Ld.k_off), T.g_raw, g->G(Ld.k_off), nullptr, nullptr, nullptr, nullptr, T.sum_dy,
                           nullptr, nullptr, g->G(Ld.b_off), 0, Ld.cout, 1, Ld.cin, st));
      ConvArgs a{};
      a.in0 = T.s2d; a.C0 = 4 * Ld.cout; a.w = Ld.w_dg; a.w_tc = Ld.w_dg_tc; a.out = T.d_o[bi - 1];
      a.N = n; a.H = Hh; a.W = Wh; a.Cout = Ld.cin; a.ks = 1; a.in_dt = g->act_dt; a.out_dt = g->act_dt;
      const bool tc = g->act_dt == DT_BF16 && conv_tc_supported(a);
      ProfScope prof(a, tc, st);
      DG_TRY(tc ? conv_fwd_tc(a, st) : conv_fwd_simt(a, st));
    }
  }

  // ---- noise MLP backward ----
  const int K = L * f;
  for (int i = 0; i < 14; ++i) {
    const DenseL& D = g->d_head[i];
    DG_TRY(k_strided_copy(T.d_film, g->film_total, g->head_off[i], T.tmp_c1, D.out, 0, n, D.out, st));
    DG_TRY(k_bn_bwd(T.tmp_c1, T.raw_head[i], nullptr, T.bn_head[i].mean, T.bn_head[i].inv_std, g->P(D.g_off), T.bn_red,
                    T.tmp_c2, g->G(D.g_off), g->G(D.be_off), n, D.out, DT_F32, st));
    DG_TRY(k_dense_bwd_w(g->film_h2, T.tmp_c2, g->G(D.k_off), g->G(D.b_off), n, K, D.out, st));
    DG_TRY(k_dense_bwd_x(T.tmp_c2, D.out, 0, g->P(D.k_off), T.d_h2, n, K, D.out, i > 0, st));
  }
  {
    const DenseL &D1 = g->d_f1, &D0 = g->d_f0;
    DG_TRY(k_bn_bwd(T.d_h2, T.pre1, g->film_h2, T.bn_f1.mean, T.bn_f1.inv_std, g->P(D1.g_off), T.bn_red, T.d_pre,
                    g->G(D1.g_off), g->G(D1.be_off), (long long)n * L, f, DT_F32, st));
    DG_TRY(k_dense_bwd_w(g->film_h1, T.d_pre, g->G(D1.k_off), g->G(D1.b_off), n * L, f, f, st));
    DG_TRY(k_dense_bwd_x(T.d_pre, f, 0, g->P(D1.k_off), T.d_h1, n * L, f, f, 0, st));
    DG_TRY(k_bn_bwd(T.d_h1, T.pre0, g->film_h1, T.bn_f0.mean, T.bn_f0.inv_std, g->P(D0.g_off), T.bn_red, T.d_pre,
                    g->G(D0.g_off), g->G(D0.be_off), (long long)n * L, f, DT_F32, st));
    DG_TRY(k_dense_bwd_w(z_dev, T.d_pre, g->G(D0.k_off), g->G(D0.b_off), n * L, 1, f, st));
  }
  // the moving statistics were refreshed in the parameter buffer: re-derive the folded inference operands
  return 0;
}

}  // extern "C"
