// Host side of the tcgen05 implicit-GEMM convolution: shape planning (shared-memory rings, TMEM columns), TMA tensor
// maps and dispatch to the kernel instantiations of conv_tc_k{1,3,5}.cu.  The kernel itself is conv_tc_kernel.cuh.
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "conv_tc_kernel.cuh"

using namespace convtc;

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;   // driver entry point: process-global
DgPerDevice g_dev;                  // shared-memory opt-in of every instantiation + SM count: per device
thread_local int g_num_sms = 148;   // SM count of the device the last conv_tc_init() of this thread ran on

CUtensorMapSwizzle swizzle_for(int kc) {  // kc bf16 channels = one swizzle span
  return kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}

// `pitch` = 16-bit elements per pixel in memory when the map covers only a channel range of the tensor (the hi or lo half
// of a split-half tensor); 0 = the tensor has exactly C channels
int make_nhwc_map(CUtensorMap* tm, const void* p, int C, int W, int H, int N, int box_c, int box_w, int box_h,
                  const char* what, int pitch = 0) {
  const cuuint64_t P = pitch ? pitch : C;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {P * 2, (cuuint64_t)W * P * 2, (cuuint64_t)H * W * P * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(box_c), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    depgan_set_error(std::string("cuTensorMapEncodeTiled(") + what + ") failed: " + std::to_string((int)r));
    return -1;
  }
  return 0;
}

int make_act_map(CUtensorMap* tm, const void* p, int C, int W, int H, int N, int kc, int ht) {
  return make_nhwc_map(tm, p, C, W, H, N, kc, ht, ht, "activation");
}

// Output of the k2 s2 transposed convolution, (N, 2H, 2W, C), seen from the input grid as (c, b, w, a, n*H + h):
// element ((n*2H + 2h + a) * 2W + 2w + b) * C + c.  One box = the [16 h][16 w] tile of one output parity (a, b).
int make_deconv_out_map(CUtensorMap* tm, const void* p, int C, int W, int H, int N, int ch, int pitch = 0) {
  const cuuint64_t P = pitch ? pitch : C;
  cuuint64_t dims[5] = {(cuuint64_t)C, 2, (cuuint64_t)W, 2, (cuuint64_t)N * H};
  cuuint64_t strides[4] = {P * 2, 2 * P * 2, (cuuint64_t)2 * W * P * 2, (cuuint64_t)4 * W * P * 2};
  cuuint32_t box[5] = {(cuuint32_t)ch, 1, 16, 1, 16};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(p), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(ch), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    depgan_set_error("cuTensorMapEncodeTiled(deconv output) failed: " + std::to_string((int)r));
    return -1;
  }
  return 0;
}

int make_w_map(CUtensorMap* tm, const void* p, int Cin, int rows, int kc, int ncta) {
  cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
  cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)ncta};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(p), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(kc), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    depgan_set_error("cuTensorMapEncodeTiled(weights) failed: " + std::to_string((int)r));
    return -1;
  }
  return 0;
}

uint32_t round1024(uint32_t x) { return (x + 1023u) & ~1023u; }

constexpr uint32_t SMEM_BUDGET = 226 * 1024 - DG_TRACE_SMEM;  // one persistent CTA per SM (227 KB opt-in limit)

uint32_t bar_bytes(int na, int nb) { return 8u * (2 * na + 2 * nb + 8) + 32u; }

bool plan(const ConvArgs& a, TcGeom* g, uint32_t* smem_bytes) {
  const int ks = a.ks, ht = 16 + ks - 1, taps = ks * ks;
  const int ncols = a.deconv ? 4 * a.Cout : a.Cout;
  // the transposed conv's epilogue is long (4*Cout columns): 128-column items keep two accumulator stages
  const int nsplit = (a.deconv && ncols % 128 == 0) ? ncols / 128 : (ncols + 255) / 256;
  if (ncols % nsplit) return false;
  const int ncta = ncols / nsplit;
  if (ncta % 16 || ncta < 16 || ncta > 256) return false;
  const int acc_stages = 4 * ncta <= 512 ? 2 : 1;
  int tmem_cols = 32;
  while (tmem_cols < 2 * ncta * acc_stages) tmem_cols *= 2;
  // epilogue staging: ch-channel chunks of the item's [16][16] tile (a transposed-conv chunk stays inside one parity)
  int ch = ncta % 32 == 0 ? 32 : 16;
  if (a.deconv) ch = (a.Cout % 64 == 0 && ncta % 64 == 0) ? 64 : (a.Cout % 32 == 0 && ncta % 32 == 0) ? 32 : 16;
  const bool split = a.in_dt == DT_F16S;  // hi / lo tiles everywhere: K = 3*Cin, two staging tiles and residual tiles
  if (split) {
    // the hi + lo tiles double the staging and residual slots: 16-channel chunks leave room for more activation stages
    // (measured at batch 64: 10.1 k slices/s against 9.2 k with 32-channel chunks; DEPGAN_SPLIT_CH = 32 / 64 to compare)
    static const int split_ch = getenv("DEPGAN_SPLIT_CH") ? atoi(getenv("DEPGAN_SPLIT_CH")) : 16;
    if ((split_ch == 16 || split_ch == 32 || split_ch == 64) && ncta % split_ch == 0 && (!a.deconv || a.Cout % split_ch == 0))
      ch = split_ch;
  }
  const int n_side = a.film_g ? (split ? 2 : 1) : (a.add_src ? 1 : 0) + (a.mask_src ? 1 : 0);
  const int stage_out = a.out ? 1 : 0;
  const uint32_t slot = 256u * ch * 2u;
  const int pool = a.pool_out ? 1 : 0;
  const uint32_t staging = (stage_out ? (split ? 4u : 2u) : 0u) * slot + 2u * n_side * slot + (pool ? slot / 2u : 0u);
  const uint32_t floats = 2u * ncols * 4 + (a.head_w ? 16u * a.Cout : 0u) + (a.film_g ? 24u * ncta : 0u) + 64u;
  // padded 64-channel chunks are OFF by default: measured on the 160 -> 64 layer they do not pay (0.249 ms against 0.202 ms
  // with exact 32-channel chunks), because a 64-channel stage of all nine taps no longer fits twice next to the
  // activation ring and the weights fall back to one ring stage per kernel row.  DEPGAN_KPAD=1 turns them on (A/B).
  static const bool no_pad = getenv("DEPGAN_KPAD") == nullptr;
  for (int kc = 64; kc >= 16; kc /= 2) {
    // (opt-in) 64-channel chunks when a source is not a multiple of 64 channels: its last chunk runs past the tensor, TMA
    // fills the missing channels with zeros, and the weights under them (the next source's, or out of bounds) multiply
    // zeros; only while the padded K stays within 25 % of the real one (160 -> 192, 224 -> 256).
    const int pc0 = (a.C0 + kc - 1) / kc, pc1 = (a.C1 + kc - 1) / kc;
    const bool exact = a.C0 % kc == 0 && a.C1 % kc == 0;
    const bool padded = !exact && !split && !no_pad && kc == 64 && a.C0 >= 64 && (a.C1 == 0 || a.C1 >= 64) &&
                        4 * (pc0 + pc1) * kc <= 5 * (a.C0 + a.C1);
    if (!exact && !padded) continue;
    const int nchunks = split ? 3 * (a.C0 + a.C1) / kc : pc0 + pc1;
    const uint32_t a_bytes = round1024((uint32_t)ht * ht * kc * 2), b_bytes = round1024((uint32_t)ncta * kc * 2);
    int na = 0, nb = 0, resident = 0, tps = 1;
    const int nb_all = taps * nchunks;
    const uint32_t w_all = (uint32_t)nb_all * b_bytes;
    const uint32_t fixed_res = 1024 + bar_bytes(6, nb_all) + floats + staging;
    if (nsplit == 1 && nb_all <= 56 && fixed_res + w_all + 2 * a_bytes <= SMEM_BUDGET) {
      resident = 1;
      nb = nb_all;
      na = (int)((SMEM_BUDGET - fixed_res - w_all) / a_bytes);
      if (na > 6) na = 6;
    } else {
      // streamed weights: ring stages of all taps of a chunk, else one kernel row, else single taps
      const uint32_t fixed_str = 1024 + bar_bytes(3, 12) + floats + staging;
      bool ok = false;
      const int tps_opts[3] = {taps, ks, 1};
      for (int oi = 0; oi < 3 && !ok; ++oi) {
        tps = tps_opts[oi];
        if (oi > 0 && tps == tps_opts[oi - 1]) continue;
        const uint32_t stage = (uint32_t)tps * b_bytes;
        const int want = tps == 1 ? 4 : (tps == taps ? 2 : 3);  // stages needed to keep the issuer fed
        for (na = 3; na >= 2 && !ok; --na) {
          if (fixed_str + na * a_bytes + want * stage > SMEM_BUDGET) continue;
          nb = (int)((SMEM_BUDGET - fixed_str - na * a_bytes) / stage);
          const int cap = tps == 1 ? 12 : (tps == taps ? 3 : 6);
          if (nb > cap) nb = cap;
          ok = true;
          break;
        }
      }
      if (!ok) {
        if (kc > 16) continue;
        return false;
      }
    }
    g->tiles_w = a.W / 16; g->tiles_h = a.H / 16;
    g->nchunk0 = split ? 2 * a.C0 / kc : pc0; g->nchunk1 = split ? 2 * a.C1 / kc : pc1;
    g->nchunk2 = split ? a.C0 / kc : 0; g->nchunk3 = split ? a.C1 / kc : 0;
    // weight columns: [in0 | in1] as packed by k_pack_conv_weights, or the split-half order [2 C0 | 2 C1 | C0 | C1]
    g->kofs1 = split ? 2 * a.C0 : a.C0;
    g->kofs2 = split ? 2 * (a.C0 + a.C1) : 0;
    g->kofs3 = split ? 2 * (a.C0 + a.C1) + a.C0 : 0;
    g->split = split ? 1 : 0;
    g->kc = kc; g->ncols_total = ncols; g->ncta = ncta; g->tmem_cols = tmem_cols;
    g->na = na; g->nb = nb; g->b_tps = tps; g->a_bytes = a_bytes; g->b_bytes = b_bytes;
    g->a_tx = (uint32_t)ht * ht * kc * 2; g->b_tx = (uint32_t)ncta * kc * 2;
    g->layout = kc == 64 ? 2u : kc == 32 ? 4u : 6u;
    g->b_resident = resident; g->acc_stages = acc_stages;
    // Two issuers share the A ring through parity waits.  A parity wait for fill f of a stage is only meaningful once
    // fill f-1 has landed; an issuer knows that for its own previous item (two items back), so a stage must not be
    // refilled more than once within two consecutive items: na >= 2 * nchunks.
    g->n_issuers = (resident && acc_stages == 2 && na >= 2 * nchunks) ? 2 : 1;
    g->ch = ch; g->n_side = n_side; g->stage_out = stage_out; g->slot_bytes = slot; g->pool = pool;
    *smem_bytes = 1024 + na * a_bytes + nb * (resident ? 1 : tps) * b_bytes + staging + bar_bytes(na, nb) + floats;
    return true;
  }
  return false;
}

}  // namespace

static bool tile_kernel_supported(const ConvArgs& a);

int conv_tc_init() {
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    DG_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) {
      depgan_set_error("cuTensorMapEncodeTiled entry point not available");
      return -1;
    }
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  int dev = 0;
  bool first = false;
  DG_TRY(dg_device_enter(g_dev, &dev, &first));
  if (first) {  // function attributes are per device: opt every instantiation in on each device we run on
    DG_TRY(set_attrs_ks1());
    DG_TRY(set_attrs_ks3());
    DG_TRY(set_attrs_ks5());
    DG_TRY(set_attrs_split());
    dg_device_mark(g_dev, dev);
  }
  if (dev < 64) g_num_sms = g_dev.sms[dev];
  else DG_CHECK_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  return 0;
}

bool conv_tc_supported(const ConvArgs& a) {
  // conv_row.cu takes the 32-channel full-width 3x3 layers, conv_rowg.cu the critics' 5x5 and the 64-channel 3x3 layers
  return conv_row_supported(a) || conv_rowg_supported(a) || tile_kernel_supported(a);
}

static bool tile_kernel_supported(const ConvArgs& a) {
  if (!dt_is_tc(a.in_dt) || a.out_dt != a.in_dt) return false;
  // split-half storage is instantiated for generator inference on the tile kernel: plain / FiLM 3x3, transposed conv
  if (a.in_dt == DT_F16S && (a.add_src || a.mask_src || a.pool_out || a.out_pre ||
                             !(a.ks == 3 || (a.ks == 1 && a.deconv && !a.film_g))))
    return false;
  // IEEE-half storage is instantiated for what generator inference runs: plain / FiLM 3x3 layers and the transposed conv
  if (a.in_dt == DT_F16 && (a.add_src || a.mask_src || a.pool_out || a.out_pre || !(a.ks == 3 || (a.ks == 1 && !a.film_g))))
    return false;
  if (a.ks != 1 && a.ks != 3 && a.ks != 5) return false;
  if (a.deconv && a.ks != 1) return false;
  if (a.H % 16 || a.W % 16 || a.H < 16 || a.W < 16) return false;
  if (a.C0 % 16 || a.C1 % 16 || a.C0 < 16) return false;
  if (a.C1 > 0 && !a.in1) return false;
  if (!a.w_tc) return false;
  if (a.head_w && (a.deconv || a.Cout > 256 || a.head_nc > 4)) return false;
  if (a.deconv && (a.film_g || a.add_src || a.mask_src)) return false;  // side inputs index the conv layout
  if (a.film_g && (a.add_src || a.mask_src || !a.out || !a.res)) return false;
  if (a.film_g && a.ks != 3) return false;  // the FiLM-residual epilogue is instantiated for the 3x3 kernels only
  if (a.deconv && (a.out_pre || !a.out)) return false;
  if (a.pool_out && (a.deconv || a.film_g || a.add_src || a.mask_src || a.head_w || a.out_pre || !a.out || a.ks == 1 ||
                     a.Cout > 256))
    return false;  // the fused pool is an epilogue of the plain conv + BN + ReLU layers only
  TcGeom g;
  uint32_t smem;
  return plan(a, &g, &smem);
}

bool conv_tc_plan_query(const ConvArgs& a, int* o) {
  TcGeom g;
  uint32_t smem;
  if (!tile_kernel_supported(a) || !plan(a, &g, &smem)) return false;
  const int v[16] = {g.kc, g.ncta, g.nchunk0 + g.nchunk1, g.na, g.nb, g.b_tps, g.b_resident, g.acc_stages, g.n_issuers,
                     g.ch, g.n_side, g.tmem_cols, (int)smem, g.pool, g.stage_out, g.ncols_total / g.ncta};
  for (int i = 0; i < 16; ++i) o[i] = v[i];
  return true;
}

int conv_fwd_tc(const ConvArgs& a, cudaStream_t st) {
  if (a.N <= 0) return 0;
  if (conv_row_supported(a)) return conv_fwd_row(a, st);
  if (conv_rowg_supported(a)) return conv_fwd_rowg(a, st);
  DG_TRY(conv_tc_init());
  TcGeom g;
  uint32_t smem;
  DG_REQUIRE(tile_kernel_supported(a) && plan(a, &g, &smem), "conv_fwd_tc: unsupported shape");
  const int ht = 16 + a.ks - 1;
  const bool split = g.split != 0;
  const int kmul = split ? 2 : 1;      // stored channels per logical input channel
  const int opitch = split ? 2 * a.Cout : 0;  // split-half outputs / side inputs: [Cout hi | Cout lo] per pixel
  TcMaps tm;
  DG_TRY(make_act_map(&tm.a0, a.in0, kmul * a.C0, a.W, a.H, a.N, g.kc, ht));
  if (a.C1 > 0) DG_TRY(make_act_map(&tm.a1, a.in1, kmul * a.C1, a.W, a.H, a.N, g.kc, ht));
  else tm.a1 = tm.a0;
  DG_TRY(make_w_map(&tm.b, a.w_tc, (split ? 3 : 1) * (a.C0 + a.C1), a.ks * a.ks * g.ncols_total, g.kc, g.ncta));
  tm.out = tm.out2 = tm.s0 = tm.s1 = tm.pool = tm.a0;  // placeholders for the maps this launch does not use
  if (a.pool_out) DG_TRY(make_nhwc_map(&tm.pool, a.pool_out, a.Cout, a.W / 2, a.H / 2, a.N, g.ch, 8, 8, "pooled output"));
  if (a.out) {
    const bf16* lo = reinterpret_cast<const bf16*>(a.out) + a.Cout;
    if (a.deconv) {
      DG_TRY(make_deconv_out_map(&tm.out, a.out, a.Cout, a.W, a.H, a.N, g.ch, opitch));
      if (split) DG_TRY(make_deconv_out_map(&tm.out2, lo, a.Cout, a.W, a.H, a.N, g.ch, opitch));
    } else {
      DG_TRY(make_nhwc_map(&tm.out, a.out, a.Cout, a.W, a.H, a.N, g.ch, 16, 16, "output", opitch));
      if (split) DG_TRY(make_nhwc_map(&tm.out2, lo, a.Cout, a.W, a.H, a.N, g.ch, 16, 16, "output (lo half)", opitch));
    }
  }
  // side inputs of the epilogue: the FiLM residual, or the add source followed by the mask source
  const void* side[2] = {nullptr, nullptr};
  if (a.film_g) side[0] = a.res;
  else {
    int k = 0;
    if (a.add_src) side[k++] = a.add_src;
    if (a.mask_src) side[k++] = a.mask_src;
  }
  if (split && a.film_g) side[1] = reinterpret_cast<const bf16*>(a.res) + a.Cout;  // the residual's lo half
  if (side[0]) DG_TRY(make_nhwc_map(&tm.s0, side[0], a.Cout, a.W, a.H, a.N, g.ch, 16, 16, "side input", opitch));
  if (side[1]) DG_TRY(make_nhwc_map(&tm.s1, side[1], a.Cout, a.W, a.H, a.N, g.ch, 16, 16, "side input", opitch));
  const int n_items = g.tiles_w * g.tiles_h * a.N * (g.ncols_total / g.ncta);
  const int grid = n_items < g_num_sms ? n_items : g_num_sms;
  // side inputs the epilogue has to stream (selects the EPI instantiation): 1 FiLM residual, 2 add / mask
  const int need = (a.film_g ? 1 : ((a.add_src || a.mask_src) ? 2 : (a.pool_out ? 4 : 0))) + (a.in_dt == DT_F16 ? 8 : 0);
  int rc;
  if (split) rc = launch_split(grid, smem, st, tm, a, g, a.film_g ? 1 : 0);
  else switch (a.ks) {
    case 1: rc = launch_ks1(grid, smem, st, tm, a, g, need); break;
    case 3: rc = launch_ks3(grid, smem, st, tm, a, g, need); break;
    case 5: rc = launch_ks5(grid, smem, st, tm, a, g, need); break;
    default: depgan_set_error("conv_fwd_tc: no kernel for this kernel size"); return -2;
  }
  static const bool dbg_sync = getenv("DEPGAN_DEBUG_SYNC") != nullptr;  // debugging aid: fail at the faulting launch
  if (rc == 0 && dbg_sync) {
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      char buf[512];
      snprintf(buf, sizeof buf,
               "conv_tc_kernel failed (%s): N=%d H=%d W=%d C0=%d C1=%d Cout=%d ks=%d deconv=%d film=%d add=%d mask=%d "
               "pre=%d out=%d head=%d relu=%d | kc=%d ncta=%d na=%d nb=%d tps=%d res=%d acc=%d iss=%d ch=%d side=%d smem=%u",
               cudaGetErrorString(e), a.N, a.H, a.W, a.C0, a.C1, a.Cout, a.ks, a.deconv, a.film_g != nullptr,
               a.add_src != nullptr, a.mask_src != nullptr, a.out_pre != nullptr, a.out != nullptr, a.head_w != nullptr,
               a.relu, g.kc, g.ncta, g.na, g.nb, g.b_tps, g.b_resident, g.acc_stages, g.n_issuers, g.ch, g.n_side, smem);
      depgan_set_error(buf);
      return -1;
    }
  }
  return rc;
}
