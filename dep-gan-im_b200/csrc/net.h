// Internal network description shared by net.cu (forward, C ABI) and net_train.cu (backward / losses).
#pragma once
#include <map>
#include <string>
#include <vector>

#include "../../include/depgan_b200.h"
#include "kernels.cuh"

constexpr int FIRST_FM = 32;  // first_fm_G TG:36 (the reference never uses another width)

struct ManifestEntry {
  std::string name;  // "<keras layer>/<weight>"
  int ndim;
  int shape[4];
  long long off;    // float offset into the flat parameter buffer (16-byte aligned)
  long long count;
  int trainable;
};
struct Manifest {
  std::vector<ManifestEntry> e;
  long long total = 0;
  std::map<std::string, int> idx;
  long long off(const std::string& n) const {
    auto it = idx.find(n);
    return it == idx.end() ? -1 : e[it->second].off;
  }
};
Manifest build_manifest(int model, const depgan_cfg& cfg);

// One Conv2D / Conv2DTranspose (+ optional BatchNormalization) layer.
struct ConvL {
  std::string name;  // keras suffix, e.g. "gen_0", "de_gen_9", "conv2d_dis_0a"
  int ks = 3, cin = 0, cout = 0, lvl = 0;
  bool deconv = false, has_bn = false;
  long long k_off = -1, b_off = -1, g_off = -1, be_off = -1, mu_off = -1, var_off = -1;
  // derived (device)
  float* scale = nullptr;    // [cout] gamma/sqrt(var+eps)            (1 without BN)
  float* shift = nullptr;    // [cout] beta + (bias-mean)*scale       (bias without BN)
  float* inv_std = nullptr;  // [cout] 1/sqrt(var+eps)                (training, BN only)
  bf16* w_tc = nullptr;      // [taps][cout][cin] bf16 (deconv: [4*cout][cin])
  float* w_dg = nullptr;     // dgrad weights for the SIMT path  [taps][cout][cin] * scale[cout], taps flipped
  bf16* w_dg_tc = nullptr;   // dgrad weights for the tcgen05 path [taps][cin][cout] * scale[cout], taps flipped
  float* w_lin = nullptr;    // critic JVP: same as kernel (alias into params)
  int split_c0 = 0;          // split-half handles: channels of the first input of a two-source layer (0: one source)
  int taps() const { return deconv ? 4 : ks * ks; }
};

struct DenseL {  // Dense (+BN) of the FiLM noise path
  std::string name;
  int in = 0, out = 0;
  long long k_off = -1, b_off = -1, g_off = -1, be_off = -1, mu_off = -1, var_off = -1;
  float* scale = nullptr;
  float* shift = nullptr;
  float* inv_std = nullptr;
};

struct Bump {  // workspace bump allocator (dry run when base == nullptr)
  char* base = nullptr;
  size_t used = 0, cap = 0;
  void* take(size_t bytes) {
    used = (used + 255) & ~size_t(255);
    void* p = base ? base + used : nullptr;
    used += bytes;
    return p;
  }
  template <typename T>
  T* arr(size_t n) { return reinterpret_cast<T*>(take(n * sizeof(T))); }
};

struct depgan_net {
  int model = 0;
  depgan_cfg cfg{};
  Manifest man;
  float* params = nullptr;
  float* grads = nullptr;
  int act_dt = DT_F32;
  size_t es = 4;
  bool prepared = false;

  // ---- generator ----
  ConvL g_in[7], g_noise[7], g_out[7], g_dec[3], g_seg;
  DenseL d_f0, d_f1, d_head[14];  // heads ordered block-major: [2*bi] = mul, [2*bi+1] = add
  int head_off[14];               // column offset of each head in film_out (stride FILM_TOTAL)
  int film_total = 0;
  const float** dev_head_w = nullptr;
  const float** dev_head_s = nullptr;
  const float** dev_head_t = nullptr;
  int* dev_head_c = nullptr;
  int* dev_head_off = nullptr;
  float *film_h1 = nullptr, *film_h2 = nullptr, *film_out = nullptr;
  void *act_a[7] = {}, *act_y[7] = {}, *act_r[7] = {}, *act_o[7] = {}, *act_pool[3] = {}, *act_up[7] = {};
  float* seg_w = nullptr;  // alias into params: gen_segmentation kernel (32, nc_out), bias
  float* seg_b = nullptr;
  float* dem_f32 = nullptr;  // last generator output (training graphs keep it here), (N,H,W,nc_out)

  // ---- critic ----
  std::vector<ConvL> c_conv;  // 11 convs
  long long d9_k = -1, d9_b = -1, dd_k = -1, dd_b = -1;
  void* c_act[11] = {};   // post-ReLU conv outputs
  void* c_pool[4] = {};   // pooled outputs (after convs 1,3,5,7)
  float* c_out = nullptr;  // (max_batch) critic scores

  // ---- data-parallel update (dp.cu): an attached NCCL communicator and / or peer-memory mailbox (not owned) ----
  void* nccl_comm = nullptr;
  struct depgan_peer* peer = nullptr;
  int dp_world = 1;

  // ---- training scratch (net_train.cu); owned by the handle ----
  struct Train* tr = nullptr;
  depgan_net() = default;
  depgan_net(const depgan_net&) = delete;
  depgan_net& operator=(const depgan_net&) = delete;
  ~depgan_net();  // net.cu (Train is complete there)

  int lvl_h(int lvl) const { return cfg.H >> lvl; }
  int lvl_w(int lvl) const { return cfg.W >> lvl; }
  float* P(long long off) const { return off < 0 ? nullptr : params + off; }
  float* G(long long off) const { return (off < 0 || !grads) ? nullptr : grads + off; }
};

// level / width tables of the generator blocks (TG:398-491)
static const int GEN_LVL[7] = {0, 1, 2, 3, 2, 1, 0};
static const int GEN_MULT[7] = {1, 2, 3, 4, 3, 2, 1};
static const char* const GEN_IN[7] = {"gen_0", "gen_2", "gen_4", "gen_8", "gen_10", "gen_14", "gen_16"};
static const char* const GEN_NOISE[7] = {"gen_noise_m1", "gen_noise_m2", "gen_noise_m3", "gen_noise_p4",
                                         "gen_noise_p3", "gen_noise_p2", "gen_noise_p1"};
static const char* const GEN_OUT[7] = {"gen_1", "gen_3", "gen_5", "gen_9", "gen_11", "gen_15", "gen_17"};
static const char* const GEN_SUF[7] = {"_m1", "_m2", "_m3", "", "_p3", "_p2", "_p1"};
static const char* const GEN_DEC[3] = {"de_gen_9", "de_gen_11", "de_gen_15"};

struct BnState { float* mean = nullptr; float* inv_std = nullptr; };  // batch statistics of one BN layer

struct Train {
  // shared scratch
  float* sum_dy = nullptr;  // [2048]
  float* g_raw = nullptr;   // raw weight gradient of a transposed conv [Cin][4*Cout]
  // ---- critic ----
  void* b[11] = {};   // pre-activation gradients of every conv, all rows
  void* bp[4] = {};   // gradients at pooled resolution
  void* v[11] = {};   // JVP activations (penalty rows only)
  void* vp[4] = {};
  float* go = nullptr;      // per-row output gradient
  float* batch3 = nullptr;  // [real | fake | mixed] critic inputs, fp32 (3n,H,W,1)
  float* g_in = nullptr;    // dD/dx, fp32 (rows,H,W,1)
  float* u = nullptr;       // JVP input (penalty rows)
  // ---- generator ----
  void* d_o[7] = {};
  void *d_r = nullptr, *d_y = nullptr, *b_in = nullptr, *d_in = nullptr, *s2d = nullptr;
  float *d_film = nullptr, *d_h2 = nullptr, *sum_d = nullptr, *sum_d1 = nullptr, *sum_d0 = nullptr;
  float** dev_dw_heads = nullptr;
  float *fake2 = nullptr, *l1g = nullptr;
  double* sums = nullptr;
  bool heads_uploaded = false;
  // ---- Keras training phase (cfg.training == 2, DEP-UResNet fit) ----
  BnState bn_in[7], bn_no[7], bn_out[7], bn_dec[3], bn_f0, bn_f1, bn_head[14];
  void *raw_a[7] = {}, *raw_y[7] = {}, *raw_o[7] = {}, *raw_up[7] = {}, *tmp_up = nullptr;
  float *pre0 = nullptr, *pre1 = nullptr, *d_h1 = nullptr, *d_pre = nullptr, *raw_head[14] = {};
  float *tmp_c1 = nullptr, *tmp_c2 = nullptr, *dseg = nullptr, *bn_red = nullptr;
  double* bn_sums = nullptr;
};


// per-launch timing records (bench.py roofline leg; see depgan_profile_begin/end)
struct ProfRec { int cls; double flops, bytes; cudaEvent_t e0, e1; int ks, H, W, cin, cout, n; };
struct ProfScope {
  bool on;
  ProfRec r;
  cudaStream_t st;
  ProfScope(const ConvArgs& a, bool tc, cudaStream_t s);
  ProfScope(const WgradArgs& a, bool tc, cudaStream_t s);
  ~ProfScope();
};

inline int act_dt_of(int precision) {
  return precision == DEPGAN_PREC_BF16 ? DT_BF16 : precision == DEPGAN_PREC_F16 ? DT_F16
         : precision == DEPGAN_PREC_F16X3 ? DT_F16S : DT_F32;
}
int net_conv(depgan_net* h, const ConvL& L, const void* in0, int C0, const void* in1, int C1, int in_dt, ConvArgs extra,
             int n, cudaStream_t st);
int gen_forward_impl(depgan_net* g, const float* x, const float* z, float* out, int n, bool keep, cudaStream_t st);
int critic_forward_impl(depgan_net* d, const float* x, float* out, int n, cudaStream_t st);
int train_alloc(depgan_net* h, Bump& b);
