// Network executors (Gen_UNet2D TG:349-498 / TU:291-428, Dis_C2D_FCN1 TG:316-345) and the C ABI of
// include/depgan_b200.h.  Host-side orchestration only: every arithmetic step is one of our CUDA kernels.
#include "net.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

long long g_launch_count = 0;
static thread_local std::string g_err;
void depgan_set_error(const std::string& msg) { g_err = msg; }

depgan_net::~depgan_net() { delete tr; }  // the Train bookkeeping object of train_alloc(); device memory is the caller's

// =========================================================================================================
// manifest
// =========================================================================================================
namespace {

void add_entry(Manifest& m, const std::string& layer, const char* weight, std::initializer_list<int> shape) {
  ManifestEntry e;
  e.name = layer + "/" + weight;
  e.ndim = (int)shape.size();
  e.count = 1;
  int i = 0;
  for (int s : shape) { e.shape[i++] = s; e.count *= s; }
  for (; i < 4; ++i) e.shape[i] = 1;
  e.trainable = !(strcmp(weight, "moving_mean") == 0 || strcmp(weight, "moving_variance") == 0);
  m.total = (m.total + 3) & ~3LL;
  e.off = m.total;
  m.total += e.count;
  m.idx[e.name] = (int)m.e.size();
  m.e.push_back(e);
}
void add_bn(Manifest& m, const std::string& layer, int c) {
  add_entry(m, layer, "gamma", {c});
  add_entry(m, layer, "beta", {c});
  add_entry(m, layer, "moving_mean", {c});
  add_entry(m, layer, "moving_variance", {c});
}

}  // namespace

Manifest build_manifest(int model, const depgan_cfg& cfg) {
  Manifest m;
  const int f = FIRST_FM;
  if (model == DEPGAN_MODEL_GEN) {
    // noise / FiLM path TG:353-395 (creation order of the reference)
    add_entry(m, "dense_noise_1_add_f0", "kernel", {1, f});
    add_entry(m, "dense_noise_1_add_f0", "bias", {f});
    add_bn(m, "dense_bn_noise_1_add_f0", f);
    add_entry(m, "dense_noise_1_add_f1", "kernel", {f, f});
    add_entry(m, "dense_noise_1_add_f1", "bias", {f});
    add_bn(m, "dense_bn_noise_1_add_f1", f);
    const int flat = cfg.noise_len * f;
    static const char* const sufs[7] = {"_m3", "_m2", "_m1", "", "_p3", "_p2", "_p1"};
    static const int mults[7] = {3, 2, 1, 4, 3, 2, 1};
    for (int i = 0; i < 7; ++i)
      for (const char* kind : {"add", "mul"}) {
        std::string n = std::string("noise_2_") + kind + sufs[i];
        add_entry(m, "dense_" + n, "kernel", {flat, f * mults[i]});
        add_entry(m, "dense_" + n, "bias", {f * mults[i]});
        add_bn(m, "dense_bn_" + n, f * mults[i]);
      }
    int cin = cfg.nicg;
    int skip_c[3] = {0, 0, 0};
    for (int bi = 0; bi < 7; ++bi) {
      const int c = f * GEN_MULT[bi];
      if (bi >= 4) cin += skip_c[6 - bi];  // concat [deconv_out, skip] TG:450,465,479
      const char* names[3] = {GEN_IN[bi], GEN_NOISE[bi], GEN_OUT[bi]};
      const int cins[3] = {cin, c, c};
      for (int j = 0; j < 3; ++j) {
        add_entry(m, std::string("conv2d_") + names[j], "kernel", {3, 3, cins[j], c});
        add_entry(m, std::string("conv2d_") + names[j], "bias", {c});
        add_bn(m, std::string("bn_") + names[j], c);
      }
      if (bi < 3) skip_c[bi] = c;
      if (bi >= 3 && bi < 6) {
        add_entry(m, std::string("deconv2d_") + GEN_DEC[bi - 3], "kernel", {2, 2, c, c});
        add_entry(m, std::string("deconv2d_") + GEN_DEC[bi - 3], "bias", {c});
        add_bn(m, std::string("bn_") + GEN_DEC[bi - 3], c);
      }
      cin = c;
    }
    add_entry(m, "gen_segmentation", "kernel", {1, 1, f, cfg.nc_out});
    add_entry(m, "gen_segmentation", "bias", {cfg.nc_out});
  } else {
    static const struct { const char* n; int k, ci, co; } cc[11] = {
        {"conv2d_dis_0a", 5, 1, 16},   {"conv2d_dis_0b", 5, 16, 16},  {"conv2d_dis_1a", 5, 16, 32},
        {"conv2d_dis_1b", 5, 32, 32},  {"conv2d_dis_2", 3, 32, 64},   {"conv2d_dis_3", 3, 64, 64},
        {"conv2d_dis_4", 3, 64, 128},  {"conv2d_dis_5", 3, 128, 128}, {"conv2d_dis_6", 3, 128, 256},
        {"conv2d_dis_7", 3, 256, 256}, {"conv2d_dis_8", 3, 256, 256}};
    for (auto& c : cc) {
      add_entry(m, c.n, "kernel", {c.k, c.k, c.ci, c.co});
      add_entry(m, c.n, "bias", {c.co});
    }
    add_entry(m, "dis_9", "kernel", {1, 1, 256, 1});
    add_entry(m, "dis_9", "bias", {1});
    add_entry(m, "dense_1", "kernel", {(cfg.H / 16) * (cfg.W / 16), 1});
    add_entry(m, "dense_1", "bias", {1});
  }
  m.total = (m.total + 3) & ~3LL;
  return m;
}

// =========================================================================================================
// construction
// =========================================================================================================
namespace {

int check_cfg(int model, const depgan_cfg* cfg) {
  DG_REQUIRE(cfg != nullptr, "cfg is null");
  DG_REQUIRE(model == DEPGAN_MODEL_GEN || model == DEPGAN_MODEL_CRITIC, "unknown model id");
  DG_REQUIRE(cfg->H >= 16 && cfg->W >= 16 && cfg->H % 16 == 0 && cfg->W % 16 == 0, "H, W must be multiples of 16");
  DG_REQUIRE(cfg->max_batch >= 1, "max_batch must be >= 1");
  DG_REQUIRE(cfg->precision == DEPGAN_PREC_FP32 || cfg->precision == DEPGAN_PREC_BF16 ||
                 cfg->precision == DEPGAN_PREC_F16 || cfg->precision == DEPGAN_PREC_F16X3, "unknown precision");
  DG_REQUIRE(cfg->precision != DEPGAN_PREC_F16 || (model == DEPGAN_MODEL_GEN && cfg->training == 0),
             "DEPGAN_PREC_F16 is the inference format of the generator (training handles keep the bf16 range)");
  DG_REQUIRE(cfg->precision != DEPGAN_PREC_F16X3 ||
                 (model == DEPGAN_MODEL_GEN && cfg->training == 0 && cfg->H % 128 == 0 && cfg->W % 128 == 0),
             "DEPGAN_PREC_F16X3 is an inference format of the generator; H and W must be multiples of 128");
  if (model == DEPGAN_MODEL_GEN) {
    DG_REQUIRE(cfg->nicg >= 1 && cfg->nicg <= 8, "nicg must be 1..8");
    DG_REQUIRE(cfg->nc_out >= 1 && cfg->nc_out <= 4, "nc_out must be 1..4");
    DG_REQUIRE(cfg->noise_len >= 1 && cfg->noise_len * FIRST_FM <= 1024, "noise_len must be 1..32");
  }
  return 0;
}

void bind_conv(depgan_net* h, ConvL& L, const std::string& conv_layer, const std::string& bn_layer) {
  L.k_off = h->man.off(conv_layer + "/kernel");
  L.b_off = h->man.off(conv_layer + "/bias");
  L.has_bn = !bn_layer.empty();
  if (L.has_bn) {
    L.g_off = h->man.off(bn_layer + "/gamma");
    L.be_off = h->man.off(bn_layer + "/beta");
    L.mu_off = h->man.off(bn_layer + "/moving_mean");
    L.var_off = h->man.off(bn_layer + "/moving_variance");
  }
}
void bind_dense(depgan_net* h, DenseL& D, const std::string& n, int in, int out) {
  D.name = n;
  D.in = in;
  D.out = out;
  D.k_off = h->man.off("dense_" + n + "/kernel");
  D.b_off = h->man.off("dense_" + n + "/bias");
  D.g_off = h->man.off("dense_bn_" + n + "/gamma");
  D.be_off = h->man.off("dense_bn_" + n + "/beta");
  D.mu_off = h->man.off("dense_bn_" + n + "/moving_mean");
  D.var_off = h->man.off("dense_bn_" + n + "/moving_variance");
}

void alloc_conv_derived(depgan_net* h, ConvL& L, Bump& b) {
  L.scale = b.arr<float>(L.cout);
  L.shift = b.arr<float>(L.cout);
  const size_t nw = (size_t)L.taps() * L.cin * L.cout;
  // 16-bit slots: bf16, or IEEE half for DT_F16 handles; split-half handles keep K = 3*Cin rows (hi, hi, lo)
  if (dt_is_tc(h->act_dt)) L.w_tc = b.arr<bf16>(nw * (h->act_dt == DT_F16S ? 3 : 1));
  if (h->cfg.training) {
    L.inv_std = b.arr<float>(L.cout);
    L.w_dg = b.arr<float>(nw);
    if (h->act_dt == DT_BF16) L.w_dg_tc = b.arr<bf16>(nw);
  }
}
void alloc_dense_derived(depgan_net* h, DenseL& D, Bump& b) {
  D.scale = b.arr<float>(D.out);
  D.shift = b.arr<float>(D.out);
  if (h->cfg.training) D.inv_std = b.arr<float>(D.out);
}

// Lays out (or, with b.base == nullptr, just sizes) everything a handle keeps in the caller's workspace.
int layout_net(depgan_net* h, Bump& b) {
  const depgan_cfg& c = h->cfg;
  const size_t NB = c.max_batch, es = h->es;
  const int f = FIRST_FM;
  if (h->model == DEPGAN_MODEL_GEN) {
    int cin = c.nicg;
    int skip_c[3] = {0, 0, 0};
    for (int bi = 0; bi < 7; ++bi) {
      const int w = f * GEN_MULT[bi], lvl = GEN_LVL[bi];
      if (bi >= 4) cin += skip_c[6 - bi];
      ConvL* Ls[3] = {&h->g_in[bi], &h->g_noise[bi], &h->g_out[bi]};
      const char* names[3] = {GEN_IN[bi], GEN_NOISE[bi], GEN_OUT[bi]};
      const int cins[3] = {cin, w, w};
      for (int j = 0; j < 3; ++j) {
        ConvL& L = *Ls[j];
        L.name = names[j]; L.ks = 3; L.cin = cins[j]; L.cout = w; L.lvl = lvl;
        if (j == 0 && bi >= 4) L.split_c0 = cin - skip_c[6 - bi];  // [deconv_out, skip]
        bind_conv(h, L, std::string("conv2d_") + names[j], std::string("bn_") + names[j]);
        alloc_conv_derived(h, L, b);
      }
      const size_t px = NB * h->lvl_h(lvl) * h->lvl_w(lvl);
      h->act_a[bi] = b.take(px * w * es);
      if (c.training) h->act_y[bi] = b.take(px * w * es);
      h->act_r[bi] = b.take(px * w * es);
      h->act_o[bi] = b.take(px * w * es);
      if (bi < 3) { skip_c[bi] = w; h->act_pool[bi] = b.take(px / 4 * w * es); }
      if (bi >= 3 && bi < 6) {
        ConvL& L = h->g_dec[bi - 3];
        L.name = GEN_DEC[bi - 3]; L.ks = 1; L.cin = w; L.cout = w; L.lvl = lvl; L.deconv = true;
        bind_conv(h, L, std::string("deconv2d_") + GEN_DEC[bi - 3], std::string("bn_") + GEN_DEC[bi - 3]);
        alloc_conv_derived(h, L, b);
        h->act_up[bi] = b.take(px * 4 * w * es);
      }
      cin = w;
    }
    h->g_seg.name = "gen_segmentation"; h->g_seg.ks = 1; h->g_seg.cin = f; h->g_seg.cout = c.nc_out;
    bind_conv(h, h->g_seg, "gen_segmentation", "");
    // FiLM path
    bind_dense(h, h->d_f0, "noise_1_add_f0", 1, f);
    bind_dense(h, h->d_f1, "noise_1_add_f1", f, f);
    alloc_dense_derived(h, h->d_f0, b);
    alloc_dense_derived(h, h->d_f1, b);
    int off = 0;
    for (int bi = 0; bi < 7; ++bi)
      for (int k = 0; k < 2; ++k) {
        DenseL& D = h->d_head[2 * bi + k];
        bind_dense(h, D, std::string("noise_2_") + (k == 0 ? "mul" : "add") + GEN_SUF[bi], c.noise_len * f,
                   f * GEN_MULT[bi]);
        alloc_dense_derived(h, D, b);
        h->head_off[2 * bi + k] = off;
        off += D.out;
      }
    h->film_total = off;
    h->dev_head_w = b.arr<const float*>(14);
    h->dev_head_s = b.arr<const float*>(14);
    h->dev_head_t = b.arr<const float*>(14);
    h->dev_head_c = b.arr<int>(14);
    h->dev_head_off = b.arr<int>(14);
    h->film_h1 = b.arr<float>(NB * c.noise_len * f);
    h->film_h2 = b.arr<float>(NB * c.noise_len * f);
    h->film_out = b.arr<float>(NB * off);
    h->dem_f32 = b.arr<float>(NB * c.H * c.W * c.nc_out);
  } else {
    static const struct { const char* n; int k, ci, co, lvl; } cc[11] = {
        {"conv2d_dis_0a", 5, 1, 16, 0},   {"conv2d_dis_0b", 5, 16, 16, 0},  {"conv2d_dis_1a", 5, 16, 32, 1},
        {"conv2d_dis_1b", 5, 32, 32, 1},  {"conv2d_dis_2", 3, 32, 64, 2},   {"conv2d_dis_3", 3, 64, 64, 2},
        {"conv2d_dis_4", 3, 64, 128, 3},  {"conv2d_dis_5", 3, 128, 128, 3}, {"conv2d_dis_6", 3, 128, 256, 4},
        {"conv2d_dis_7", 3, 256, 256, 4}, {"conv2d_dis_8", 3, 256, 256, 4}};
    h->c_conv.assign(11, ConvL());
    for (int i = 0; i < 11; ++i) {
      ConvL& L = h->c_conv[i];
      L.name = cc[i].n; L.ks = cc[i].k; L.cin = cc[i].ci; L.cout = cc[i].co; L.lvl = cc[i].lvl;
      bind_conv(h, L, cc[i].n, "");
      alloc_conv_derived(h, L, b);
      const size_t px = NB * h->lvl_h(L.lvl) * h->lvl_w(L.lvl);
      h->c_act[i] = b.take(px * L.cout * es);
      if (i == 1 || i == 3 || i == 5 || i == 7) h->c_pool[i / 2] = b.take(px / 4 * L.cout * es);
    }
    h->d9_k = h->man.off("dis_9/kernel"); h->d9_b = h->man.off("dis_9/bias");
    h->dd_k = h->man.off("dense_1/kernel"); h->dd_b = h->man.off("dense_1/bias");
    h->c_out = b.arr<float>(NB);
  }
  if (c.training) DG_TRY(train_alloc(h, b));
  return 0;
}

int fold_conv(depgan_net* h, ConvL& L, cudaStream_t st) {
  DG_TRY(k_fold_bn(h->P(L.b_off), h->P(L.g_off), h->P(L.be_off), h->P(L.mu_off), h->P(L.var_off), L.scale, L.shift,
                   L.inv_std, L.cout, st));
  if (h->act_dt == DT_F16S) {  // inference only: no dgrad operands
    const int c0 = L.split_c0 ? L.split_c0 : L.cin;
    return k_pack_split_weights(h->P(L.k_off), L.w_tc, L.taps(), L.cin, L.cout, c0, L.deconv ? 1 : 0, st);
  }
  if (L.deconv) {
    if (L.w_tc) DG_TRY(k_convert_in(h->P(L.k_off), L.w_tc, (long long)4 * L.cin * L.cout, h->act_dt, st));
    DG_TRY(k_pack_deconv_dgrad(h->P(L.k_off), h->cfg.training == 2 ? nullptr : L.scale, L.w_dg, L.w_dg_tc, L.cin,
                               L.cout, st));
  } else {
    // training == 2 (Keras training phase): BN is applied after the raw convolution, so dgrad weights stay unscaled
    DG_TRY(k_pack_conv_weights(h->P(L.k_off), h->cfg.training == 2 ? nullptr : L.scale, L.w_tc, L.w_dg, L.w_dg_tc,
                               L.ks * L.ks, L.cin, L.cout, st, h->act_dt == DT_F16));
  }
  return 0;
}
int fold_dense(depgan_net* h, DenseL& D, cudaStream_t st) {
  return k_fold_bn(h->P(D.b_off), h->P(D.g_off), h->P(D.be_off), h->P(D.mu_off), h->P(D.var_off), D.scale, D.shift,
                   D.inv_std, D.out, st);
}

}  // namespace

// =========================================================================================================
// optional per-launch timing (bench.py roofline leg): CUDA events around every convolution launch
// =========================================================================================================
bool g_prof_on = false;
std::vector<ProfRec> g_prof;

ProfScope::ProfScope(const ConvArgs& a, bool tc, cudaStream_t s) : on(g_prof_on), st(s) {
  if (!on) return;
  const double px = (double)a.N * a.H * a.W, cin = a.C0 + a.C1;
  const double ncols = a.deconv ? 4.0 * a.Cout : (double)a.Cout;
  r.cls = !tc ? 3 : (a.pool_out ? 6 : (a.deconv ? 2 : (a.ks == 5 ? 1 : a.ks == 3 ? 0 : 2)));
  r.flops = 2.0 * px * a.ks * a.ks * cin * ncols;
  const double ies = dt_size(a.in_dt), oes = dt_size(a.out_dt);
  r.bytes = px * cin * ies + (a.out ? px * ncols * oes : 0.0) + (a.out_pre ? px * ncols * oes : 0.0) +
            (a.res ? px * ncols * oes : 0.0) + (a.add_src ? px * ncols * oes : 0.0) +
            (a.mask_src ? px * ncols * oes : 0.0) + (a.head_out ? px * a.head_nc * 4.0 : 0.0) +
            (a.pool_out ? 0.25 * px * ncols * oes : 0.0);
  r.ks = a.ks; r.H = a.H; r.W = a.W; r.cin = a.C0 + a.C1; r.cout = (int)ncols; r.n = a.N;
  cudaEventCreate(&r.e0);
  cudaEventCreate(&r.e1);
  cudaEventRecord(r.e0, st);
}
ProfScope::ProfScope(const WgradArgs& a, bool tc, cudaStream_t s) : on(g_prof_on), st(s) {
  if (!on) return;
  const double px = (double)a.N * a.H * a.W, cin = a.C0 + a.C1;
  r.cls = tc ? 5 : 4;
  r.flops = 2.0 * px * a.ks * a.ks * cin * a.Cout;
  r.bytes = px * cin * dt_size(a.x_dt) + px * a.Cout * dt_size(a.dy_dt);
  r.ks = a.ks; r.H = a.H; r.W = a.W; r.cin = (int)cin; r.cout = a.Cout; r.n = a.N;
  cudaEventCreate(&r.e0);
  cudaEventCreate(&r.e1);
  cudaEventRecord(r.e0, st);
}
ProfScope::~ProfScope() {
  if (!on) return;
  cudaEventRecord(r.e1, st);
  g_prof.push_back(r);
}

// =========================================================================================================
// forward executors
// =========================================================================================================
int net_conv(depgan_net* h, const ConvL& L, const void* in0, int C0, const void* in1, int C1, int in_dt, ConvArgs a,
             int n, cudaStream_t st) {
  a.in0 = in0; a.in1 = in1; a.C0 = C0; a.C1 = C1;
  a.w = h->P(L.k_off); a.w_tc = L.w_tc;
  a.scale = L.has_bn ? L.scale : nullptr;
  a.shift = L.shift;
  a.N = n; a.H = h->lvl_h(L.lvl); a.W = h->lvl_w(L.lvl); a.Cout = L.cout; a.ks = L.ks;
  a.in_dt = in_dt; a.out_dt = h->act_dt;
  if (a.pool_out && !(dt_is_tc(h->act_dt) && conv_tc_supported(a))) {  // unfused fallback: conv, then the pool
    void* pool_out = a.pool_out;
    a.pool_out = nullptr;
    DG_TRY(net_conv(h, L, in0, C0, in1, C1, in_dt, a, n, st));
    return k_maxpool_fwd(a.out, pool_out, n, a.H, a.W, a.Cout, h->act_dt, st);
  }
  const bool tc = dt_is_tc(h->act_dt) && conv_tc_supported(a);
  ProfScope prof(a, tc, st);
  if (tc) return conv_fwd_tc(a, st);
  if (a.head_w) {  // unfused fallback: conv, then the 1x1 head
    ConvArgs b = a;
    b.head_w = nullptr;
    DG_TRY(conv_fwd_simt(b, st));
    return k_head_fwd(a.out, a.head_w, a.head_b, a.head_out, (long long)n * a.H * a.W, a.Cout, a.head_nc, a.head_act,
                      a.out_dt, st);
  }
  return conv_fwd_simt(a, st);
}

static int net_deconv(depgan_net* h, const ConvL& L, const void* in, void* out, int n, cudaStream_t st) {
  const int H = h->lvl_h(L.lvl), W = h->lvl_w(L.lvl);
  if (dt_is_tc(h->act_dt)) {
    ConvArgs a{};
    a.in0 = in; a.C0 = L.cin; a.w_tc = L.w_tc; a.scale = L.scale; a.shift = L.shift; a.out = out; a.relu = 1;
    a.deconv = 1; a.N = n; a.H = H; a.W = W; a.Cout = L.cout; a.ks = 1; a.in_dt = h->act_dt; a.out_dt = h->act_dt;
    if (conv_tc_supported(a)) {
      ProfScope prof(a, true, st);
      return conv_fwd_tc(a, st);
    }
  }
  return k_deconv_fwd(in, h->P(L.k_off), L.scale, L.shift, out, n, H, W, L.cin, L.cout, h->act_dt, 1, st);
}

int gen_forward_impl(depgan_net* g, const float* x, const float* z, float* out, int n, bool keep, cudaStream_t st) {
  const depgan_cfg& c = g->cfg;
  const int f = FIRST_FM;
  {  // FiLM parameters for all 7 blocks in one pass (TG:353-395)
    FilmMlpArgs fa{};
    fa.z = z;
    fa.k0 = g->P(g->d_f0.k_off); fa.s0 = g->d_f0.scale; fa.t0 = g->d_f0.shift;
    fa.k1 = g->P(g->d_f1.k_off); fa.s1 = g->d_f1.scale; fa.t1 = g->d_f1.shift;
    fa.head_w = g->dev_head_w; fa.head_s = g->dev_head_s; fa.head_t = g->dev_head_t;
    fa.head_c = g->dev_head_c; fa.head_off = g->dev_head_off;
    fa.n_heads = 14; fa.total_c = g->film_total;
    fa.h1 = g->film_h1; fa.h2 = g->film_h2; fa.out = g->film_out;
    fa.N = n; fa.L = c.noise_len; fa.F = f;
    DG_TRY(k_film_mlp_fwd(fa, st));
  }
  const void* in0 = x;
  const void* in1 = nullptr;
  int C0 = c.nicg, C1 = 0, in_dt = DT_F32;
  for (int bi = 0; bi < 7; ++bi) {
    const int w = f * GEN_MULT[bi];
    ConvArgs e{};
    e.out = g->act_a[bi]; e.relu = 1;                                   // conv2d_bn_relu
    DG_TRY(net_conv(g, g->g_in[bi], in0, C0, in1, C1, in_dt, e, n, st));
    e = ConvArgs{};                                                      // conv2d_bn -> mul, add, relu, +a
    e.out = g->act_r[bi];
    e.out_pre = keep ? g->act_y[bi] : nullptr;
    e.film_g = g->film_out + g->head_off[2 * bi]; e.film_b = g->film_out + g->head_off[2 * bi + 1];
    e.film_stride = g->film_total; e.res = g->act_a[bi];
    DG_TRY(net_conv(g, g->g_noise[bi], g->act_a[bi], w, nullptr, 0, g->act_dt, e, n, st));
    e = ConvArgs{};                                                      // conv2d_bn_relu
    e.out = g->act_o[bi]; e.relu = 1;
    if (bi == 6) {                                                       // + gen_segmentation 1x1 + tanh/softmax
      e.head_w = g->P(g->g_seg.k_off); e.head_b = g->P(g->g_seg.b_off); e.head_out = out;
      e.head_nc = c.nc_out; e.head_act = c.nc_out == 1 ? 0 : 1;
      if (!keep && dt_is_tc(g->act_dt)) e.out = nullptr;                // inference: gen_17 never leaves the SM
    }
    // MaxPooling2D (TG:409): fused into the epilogue of the block's last conv for training handles (the per-launch
    // profile counts those launches as their own class, 6, with the pooled bytes included).  Inference handles keep
    // the separate bandwidth pass: measured at batch 64 the two are equal within noise (27.37 k fused vs 27.46 k
    // slices/s), the epilogue's extra shared-memory traffic costing what the saved pass would have.
    // Round 2: the first block's pool (32 channels at full resolution) is fused for inference handles too -- the generic
    // row kernel pools the row pair it holds in registers, which saves re-reading the 268 MB tensor (A/B switch below).
    static const bool no_infer_fuse = getenv("DEPGAN_NO_INFER_POOLFUSE") != nullptr;
    const bool fuse_pool = bi < 3 && (g->cfg.training != 0 || (bi == 0 && dt_is_half(g->act_dt) && !no_infer_fuse));
    if (fuse_pool) e.pool_out = g->act_pool[bi];
    DG_TRY(net_conv(g, g->g_out[bi], g->act_r[bi], w, nullptr, 0, g->act_dt, e, n, st));
    if (bi < 3) {
      if (!fuse_pool)
        DG_TRY(k_maxpool_fwd(g->act_o[bi], g->act_pool[bi], n, g->lvl_h(GEN_LVL[bi]), g->lvl_w(GEN_LVL[bi]), w,
                             g->act_dt, st));
      in0 = g->act_pool[bi]; C0 = w; in1 = nullptr; C1 = 0;
    } else if (bi < 6) {
      DG_TRY(net_deconv(g, g->g_dec[bi - 3], g->act_o[bi], g->act_up[bi], n, st));
      in0 = g->act_up[bi]; C0 = w; in1 = g->act_o[5 - bi]; C1 = f * GEN_MULT[5 - bi];  // [deconv, skip]
    }
    in_dt = g->act_dt;
  }
  return 0;
}

int critic_forward_impl(depgan_net* d, const float* x, float* out, int n, cudaStream_t st) {
  const void* in = x;
  int in_dt = DT_F32, C = 1;
  for (int i = 0; i < 11; ++i) {
    const ConvL& L = d->c_conv[i];
    const bool pooled = i == 1 || i == 3 || i == 5 || i == 7;
    ConvArgs e{};
    e.out = d->c_act[i]; e.relu = 1;
    if (pooled) e.pool_out = d->c_pool[i / 2];                          // MaxPooling2D fused into the conv
    DG_TRY(net_conv(d, L, in, C, nullptr, 0, in_dt, e, n, st));
    in = d->c_act[i]; C = L.cout; in_dt = d->act_dt;
    if (pooled) in = d->c_pool[i / 2];
  }
  const int hw = d->lvl_h(4) * d->lvl_w(4);
  return k_critic_head_fwd(in, d->P(d->d9_k), d->P(d->d9_b), d->P(d->dd_k), d->P(d->dd_b), out, n, hw, 256, d->act_dt,
                           st);
}

// =========================================================================================================
// C ABI
// =========================================================================================================
extern "C" {

const char* depgan_last_error(void) { return g_err.c_str(); }
int depgan_abi_version(void) { return 1; }

int depgan_manifest_count(int model, const depgan_cfg* cfg) {
  if (check_cfg(model, cfg)) return -2;
  return (int)build_manifest(model, *cfg).e.size();
}
int depgan_manifest_entry(int model, const depgan_cfg* cfg, int idx, char* name, int name_cap, int* ndim, int* shape,
                          long long* offset, int* trainable) {
  if (check_cfg(model, cfg)) return -2;
  Manifest m = build_manifest(model, *cfg);
  DG_REQUIRE(idx >= 0 && idx < (int)m.e.size(), "manifest index out of range");
  const ManifestEntry& e = m.e[idx];
  DG_REQUIRE(name && name_cap > (int)e.name.size(), "name buffer too small");
  strcpy(name, e.name.c_str());
  if (ndim) *ndim = e.ndim;
  if (shape) for (int i = 0; i < e.ndim; ++i) shape[i] = e.shape[i];
  if (offset) *offset = e.off;
  if (trainable) *trainable = e.trainable;
  return 0;
}
long long depgan_manifest_floats(int model, const depgan_cfg* cfg) {
  if (check_cfg(model, cfg)) return -2;
  return build_manifest(model, *cfg).total;
}

long long depgan_workspace_bytes(int model, const depgan_cfg* cfg) {
  if (check_cfg(model, cfg)) return -2;
  depgan_net h;
  h.model = model; h.cfg = *cfg; h.man = build_manifest(model, *cfg);
  h.act_dt = act_dt_of(cfg->precision);
  h.es = dt_size(h.act_dt);
  Bump b;
  if (layout_net(&h, b)) return -2;
  return (long long)b.used + 1024;
}

depgan_net* depgan_net_create(int model, const depgan_cfg* cfg, float* params_dev, float* grads_dev,
                              void* workspace_dev, long long workspace_bytes) {
  if (check_cfg(model, cfg)) return nullptr;
  if (!params_dev || !workspace_dev) { depgan_set_error("params_dev / workspace_dev must not be null"); return nullptr; }
  if (cfg->training && !grads_dev) { depgan_set_error("training handles need a gradient buffer"); return nullptr; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    depgan_set_error("no CUDA device: depgan_b200 has no CPU fallback");
    return nullptr;
  }
  depgan_net* h = new depgan_net();
  h->model = model; h->cfg = *cfg; h->man = build_manifest(model, *cfg);
  h->params = params_dev; h->grads = grads_dev;
  h->act_dt = act_dt_of(cfg->precision);
  h->es = dt_size(h->act_dt);
  Bump b;
  b.base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace_dev) + 255) & ~uintptr_t(255));
  b.cap = (size_t)workspace_bytes;
  if (layout_net(h, b) || (long long)b.used + 256 > workspace_bytes) {
    if (g_err.empty() || (long long)b.used + 256 > workspace_bytes) depgan_set_error("workspace too small");
    delete h;
    return nullptr;
  }
  if (dt_is_tc(h->act_dt) && conv_tc_init()) { delete h; return nullptr; }
  return h;
}

void depgan_net_destroy(depgan_net* h) { delete h; }

int depgan_net_prepare(depgan_net* h, void* stream) {
  DG_REQUIRE(h != nullptr, "null handle");
  cudaStream_t st = (cudaStream_t)stream;
  if (h->model == DEPGAN_MODEL_GEN) {
    for (int bi = 0; bi < 7; ++bi) {
      DG_TRY(fold_conv(h, h->g_in[bi], st));
      DG_TRY(fold_conv(h, h->g_noise[bi], st));
      DG_TRY(fold_conv(h, h->g_out[bi], st));
    }
    for (int i = 0; i < 3; ++i) DG_TRY(fold_conv(h, h->g_dec[i], st));
    DG_TRY(fold_dense(h, h->d_f0, st));
    DG_TRY(fold_dense(h, h->d_f1, st));
    for (int i = 0; i < 14; ++i) DG_TRY(fold_dense(h, h->d_head[i], st));
    if (!h->prepared) {
      const float *hw[14], *hs[14], *ht[14];
      int hc[14];
      for (int i = 0; i < 14; ++i) {
        hw[i] = h->P(h->d_head[i].k_off); hs[i] = h->d_head[i].scale; ht[i] = h->d_head[i].shift;
        hc[i] = h->d_head[i].out;
      }
      DG_CHECK_CUDA(cudaMemcpyAsync(h->dev_head_w, hw, sizeof(hw), cudaMemcpyHostToDevice, st));
      DG_CHECK_CUDA(cudaMemcpyAsync(h->dev_head_s, hs, sizeof(hs), cudaMemcpyHostToDevice, st));
      DG_CHECK_CUDA(cudaMemcpyAsync(h->dev_head_t, ht, sizeof(ht), cudaMemcpyHostToDevice, st));
      DG_CHECK_CUDA(cudaMemcpyAsync(h->dev_head_c, hc, sizeof(hc), cudaMemcpyHostToDevice, st));
      DG_CHECK_CUDA(cudaMemcpyAsync(h->dev_head_off, h->head_off, sizeof(h->head_off), cudaMemcpyHostToDevice, st));
      DG_CHECK_CUDA(cudaStreamSynchronize(st));  // the host arrays above are stack temporaries
    }
  } else {
    // the critics are re-prepared after each of their 10 updates per generator iteration: two launches for all layers
    PrepTable t{};
    for (auto& L : h->c_conv) {
      if (L.deconv || t.n >= 12) { t.n = -1; break; }
      PrepLayer& p = t.L[t.n++];
      p.bias = h->P(L.b_off); p.gamma = h->P(L.g_off); p.beta = h->P(L.be_off); p.mean = h->P(L.mu_off);
      p.var = h->P(L.var_off); p.scale = L.scale; p.shift = L.shift; p.inv_std = L.inv_std;
      p.w = h->P(L.k_off); p.w_tc = L.w_tc; p.w_dg = L.w_dg; p.w_dg_tc = L.w_dg_tc;
      p.C = L.cout; p.taps = L.ks * L.ks; p.cin = L.cin; p.scale_dgrad = h->cfg.training == 2 ? 0 : 1;
    }
    if (t.n > 0) DG_TRY(k_prepare_convs(t, st));
    else for (auto& L : h->c_conv) DG_TRY(fold_conv(h, L, st));
  }
  h->prepared = true;
  return 0;
}

int depgan_gen_forward(depgan_net* g, const float* x_dev, const float* z_dev, float* out_dev, int n, void* stream) {
  DG_REQUIRE(g && g->model == DEPGAN_MODEL_GEN, "gen_forward: not a generator handle");
  DG_REQUIRE(g->prepared, "gen_forward: call depgan_net_prepare after loading weights");
  DG_REQUIRE(n >= 0 && n <= g->cfg.max_batch, "gen_forward: n exceeds max_batch");
  if (n == 0) return 0;
  return gen_forward_impl(g, x_dev, z_dev, out_dev, n, false, (cudaStream_t)stream);
}

int depgan_critic_forward(depgan_net* d, const float* x_dev, float* out_dev, int n, void* stream) {
  DG_REQUIRE(d && d->model == DEPGAN_MODEL_CRITIC, "critic_forward: not a critic handle");
  DG_REQUIRE(d->prepared, "critic_forward: call depgan_net_prepare after loading weights");
  DG_REQUIRE(n >= 0 && n <= d->cfg.max_batch, "critic_forward: n exceeds max_batch");
  if (n == 0) return 0;
  return critic_forward_impl(d, x_dev, out_dev, n, (cudaStream_t)stream);
}

int depgan_adam_step(float* params_dev, const float* grads_dev, float* m_dev, float* v_dev, long long n, int t,
                     float lr, float beta1, float beta2, float eps, float grad_scale, void* stream) {
  DG_REQUIRE(t >= 1, "adam: t is 1-based");
  // lr_t in double, as Keras evaluates it in python/TF float32 graph: lr * sqrt(1-b2^t) / (1-b1^t)
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, t)) / (1.0 - pow((double)beta1, t));
  return k_adam(params_dev, grads_dev, m_dev, v_dev, n, (float)lr_t, beta1, beta2, eps, grad_scale,
                (cudaStream_t)stream);
}

int depgan_dem_accumulate(double* acc_dev, const float* pred_dev, const float* mask_dev, long long n, int chan,
                          void* stream) {
  DG_REQUIRE(chan >= 1, "dem_accumulate: chan must be >= 1");
  return k_dem_accumulate(acc_dev, pred_dev, mask_dev, n, chan, (cudaStream_t)stream);
}
int depgan_dem_postproc(const float* x_dev, int nicg, const double* acc_dev, double n_repeat, const float* mask_dev,
                        double thr, double* dem_out_dev, double* fake2_out_dev, unsigned char* labels_dev,
                        unsigned long long* count_dev, long long npix, void* stream) {
  DG_REQUIRE(nicg >= 1 && count_dev, "dem_postproc: bad arguments");
  return k_dem_postproc(x_dev, nicg, acc_dev, n_repeat, mask_dev, thr, dem_out_dev, fake2_out_dev, labels_dev,
                        count_dev, npix, (cudaStream_t)stream);
}
int depgan_label_confusion(const unsigned char* fake_labels_dev, const unsigned char* real_labels_dev, long long n,
                            unsigned long long* conf16_dev, void* stream) {
  DG_REQUIRE(n >= 0 && conf16_dev && (n == 0 || (fake_labels_dev && real_labels_dev)), "label_confusion: bad arguments");
  return k_label_confusion(fake_labels_dev, real_labels_dev, n, conf16_dev, (cudaStream_t)stream);
}

int depgan_uresnet_labels(const double* acc_dev, double n_repeat, int chan, double* mean_out_dev,
                          unsigned char* labels_dev, unsigned long long* count_dev, long long npix, void* stream) {
  DG_REQUIRE(chan >= 1 && labels_dev && count_dev, "uresnet_labels: bad arguments");
  return k_uresnet_labels(acc_dev, n_repeat, chan, mean_out_dev, labels_dev, count_dev, npix, (cudaStream_t)stream);
}

long long depgan_launch_count(void) { return g_launch_count; }

int depgan_profile_begin(void) {
  for (auto& r : g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  g_prof.clear();
  g_prof_on = true;
  return 0;
}
int depgan_profile_end(double* ms, double* flops, double* bytes, long long* launches, int ncls) {
  g_prof_on = false;
  DG_REQUIRE(ms && flops && bytes && launches && ncls >= 1, "profile_end: bad arguments");
  for (int i = 0; i < ncls; ++i) { ms[i] = flops[i] = bytes[i] = 0.0; launches[i] = 0; }
  DG_CHECK_CUDA(cudaDeviceSynchronize());
  FILE* log = nullptr;
  if (const char* path = getenv("DEPGAN_PROFILE_LOG")) log = fopen(path, "w");
  if (log) fprintf(log, "class,ks,N,H,W,cin,ncols,ms,tflops,gbs\n");
  for (auto& r : g_prof) {
    float t = 0.f;
    DG_CHECK_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
    if (log)
      fprintf(log, "%d,%d,%d,%d,%d,%d,%d,%.4f,%.1f,%.1f\n", r.cls, r.ks, r.n, r.H, r.W, r.cin, r.cout, t,
              r.flops / (t * 1e-3) / 1e12, r.bytes / (t * 1e-3) / 1e9);
    if (r.cls < ncls) { ms[r.cls] += t; flops[r.cls] += r.flops; bytes[r.cls] += r.bytes; launches[r.cls] += 1; }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  if (log) fclose(log);
  g_prof.clear();
  return 0;
}

int depgan_debug_activation(depgan_net* h, const char* name, float* out_dev, long long cap_floats, long long* n_floats,
                            int n, void* stream) {
  DG_REQUIRE(h && name, "debug_activation: bad arguments");
  const void* src = nullptr;
  long long cnt = 0;
  int ch = 1;  // channels of the activation found
  std::string nm(name);
  if (h->model == DEPGAN_MODEL_GEN) {
    for (int bi = 0; bi < 7 && !src; ++bi) {
      const long long px = (long long)n * h->lvl_h(GEN_LVL[bi]) * h->lvl_w(GEN_LVL[bi]);
      const int w = FIRST_FM * GEN_MULT[bi];
      ch = w;
      if (nm == GEN_IN[bi]) { src = h->act_a[bi]; cnt = px * w; }
      else if (nm == GEN_NOISE[bi]) { src = h->act_r[bi]; cnt = px * w; }
      else if (nm == GEN_OUT[bi]) { src = h->act_o[bi]; cnt = px * w; }
      else if (bi >= 3 && bi < 6 && nm == GEN_DEC[bi - 3]) { src = h->act_up[bi]; cnt = px * 4 * w; }
    }
    if (!src && nm == "film") {
      DG_REQUIRE((long long)n * h->film_total <= cap_floats, "debug_activation: buffer too small");
      if (n_floats) *n_floats = (long long)n * h->film_total;
      DG_CHECK_CUDA(cudaMemcpyAsync(out_dev, h->film_out, sizeof(float) * n * h->film_total, cudaMemcpyDeviceToDevice,
                                    (cudaStream_t)stream));
      return 0;
    }
  } else {
    for (int i = 0; i < 11 && !src; ++i)
      if (nm == h->c_conv[i].name) {
        src = h->c_act[i];
        cnt = (long long)n * h->lvl_h(h->c_conv[i].lvl) * h->lvl_w(h->c_conv[i].lvl) * h->c_conv[i].cout;
      }
  }
  DG_REQUIRE(src != nullptr, "debug_activation: unknown activation name");
  DG_REQUIRE(cnt <= cap_floats, "debug_activation: buffer too small");
  if (n_floats) *n_floats = cnt;
  if (h->act_dt == DT_F16S) return k_split_to_f32(src, out_dev, cnt / ch, ch, (cudaStream_t)stream);
  return k_copy_to_f32(src, out_dev, cnt, h->act_dt, (cudaStream_t)stream);
}

// Kernel-level entry point (tests / micro-benchmarks): one fused convolution on caller-owned buffers.
static ConvArgs conv_args_of(const depgan_conv_desc* d) {
  ConvArgs a{};
  a.in0 = d->in0; a.in1 = d->in1; a.C0 = d->C0; a.C1 = d->C1;
  a.w = d->w_f32; a.w_tc = (const bf16*)d->w_bf16;
  a.scale = d->scale; a.shift = d->shift; a.out = d->out; a.out_pre = d->out_pre;
  a.film_g = d->film_g; a.film_b = d->film_b; a.film_stride = d->film_stride; a.res = d->res;
  a.add_src = d->add_src; a.mask_src = d->mask_src; a.relu = d->relu; a.deconv = d->deconv;
  a.head_w = d->head_w; a.head_b = d->head_b; a.head_out = d->head_out; a.head_nc = d->head_nc;
  a.head_act = d->head_act;
  a.N = d->N; a.H = d->H; a.W = d->W; a.Cout = d->Cout; a.ks = d->ks;
  a.in_dt = d->in_bf16 ? DT_BF16 : DT_F32; a.out_dt = d->out_bf16 ? DT_BF16 : DT_F32;
  a.pool_out = d->pool_out;
  return a;
}

int depgan_op_conv_plan(const depgan_conv_desc* d, int* plan16) {
  DG_REQUIRE(d != nullptr && plan16 != nullptr, "op_conv_plan: null argument");
  return conv_tc_plan_query(conv_args_of(d), plan16) ? 1 : 0;
}

int depgan_op_conv2d(const depgan_conv_desc* d, void* stream) {
  DG_REQUIRE(d != nullptr, "op_conv2d: null descriptor");
  ConvArgs a = conv_args_of(d);
  DG_REQUIRE(!a.pool_out || d->use_tc, "op_conv2d: pool_out is an epilogue of the tcgen05 path");
  if (d->use_tc) {
    DG_REQUIRE(conv_tc_supported(a), "op_conv2d: shape not supported by the tcgen05 path");
    return conv_fwd_tc(a, (cudaStream_t)stream);
  }
  return conv_fwd_simt(a, (cudaStream_t)stream);
}

int depgan_op_wgrad(const void* x0, const void* x1, int C0, int C1, const void* dy, float* dw, int N, int H, int W,
                    int Cout, int ks, int use_tc, void* stream) {
  WgradArgs a{};
  a.x0 = x0; a.x1 = x1; a.C0 = C0; a.C1 = C1; a.dy = dy; a.dw = dw; a.N = N; a.H = H; a.W = W; a.Cout = Cout;
  a.ks = ks; a.alpha = 1.f;
  a.x_dt = a.dy_dt = use_tc == 1 ? DT_BF16 : DT_F32;
  if (use_tc == 2) a.dy_dt = DT_BF16;  // first-layer case of the bf16 networks: fp32 image, bf16 gradient
  if (use_tc == 1) {
    DG_REQUIRE(wgrad_tc_supported(a), "op_wgrad: shape not supported by the tcgen05 path");
    return conv_wgrad_tc(a, (cudaStream_t)stream);
  }
  return conv_wgrad_simt(a, (cudaStream_t)stream);
}

int depgan_op_wgrad_csum(const void* x0, const void* x1, int C0, int C1, const void* dy, float* dw, float* csum, int N,
                         int H, int W, int Cout, int ks, void* stream) {
  WgradArgs a{};
  a.x0 = x0; a.x1 = x1; a.C0 = C0; a.C1 = C1; a.dy = dy; a.dw = dw; a.N = N; a.H = H; a.W = W; a.Cout = Cout;
  a.ks = ks; a.alpha = 1.f; a.csum = csum;
  a.x_dt = a.dy_dt = DT_BF16;
  DG_REQUIRE(wgrad_tc_supported(a), "op_wgrad_csum: shape not supported by the tcgen05 path");
  return conv_wgrad_tc(a, (cudaStream_t)stream);
}

// fp32 [taps][Cin][Cout] (Keras HWIO) -> bf16 [taps][Cout][Cin] (the tcgen05 B operand); tests / benchmarks.
int depgan_op_pack_weights(const float* w_f32_dev, void* w_bf16_dev, int taps, int cin, int cout, void* stream) {
  return k_pack_conv_weights(w_f32_dev, nullptr, (bf16*)w_bf16_dev, nullptr, nullptr, taps, cin, cout,
                             (cudaStream_t)stream);
}
int depgan_op_f32_to_bf16(const float* src_dev, void* dst_dev, long long n, void* stream) {
  return k_convert_in(src_dev, dst_dev, n, DT_BF16, (cudaStream_t)stream);
}
int depgan_op_f32_to_f16(const float* src_dev, void* dst_dev, long long n, void* stream) {
  return k_cast_narrow(src_dev, dst_dev, n, 1, (cudaStream_t)stream);
}
int depgan_op_bf16_to_f32(const void* src_dev, float* dst_dev, long long n, void* stream) {
  return k_copy_to_f32(src_dev, dst_dev, n, DT_BF16, (cudaStream_t)stream);
}

}  // extern "C"
