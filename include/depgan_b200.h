/* depgan_b200 -- C ABI of the B200-native DEP-GAN / DEP-UResNet hot path.
 *
 * The reference (febrianrachmadi/dep-gan-im) has no FFI of its own: its boundary is the Keras object surface
 * its four scripts use.  Every entry point below cites the reference call it replaces
 * (TG = DEP-GAN_PROB_IM_twoCritics_training_4fold.py, EG = DEP-GAN_testing_4fold.py,
 *  TU = DEP-UResNet-wNoises-training-4fold.py, EU = DEP-UResNet_testing_4fold.py).
 *
 * Conventions: plain pointers and sizes only; all *_dev pointers are CUDA device pointers owned by the caller;
 * every compute call is asynchronous on `stream` (a cudaStream_t passed as void*); return 0 = ok, <0 = error
 * (message via depgan_last_error()); no exceptions cross the boundary; no allocation after *_create.
 * One handle per (device, network); a handle is not thread-safe, distinct handles are.
 * Layouts: images NHWC float32; noise (N,L,1) float32; parameters = one flat float32 buffer in Keras tensor
 * layouts (Conv2D HWIO, Conv2DTranspose (kh,kw,Cout,Cin), Dense (in,out), BN gamma/beta/mean/var) at the
 * offsets reported by depgan_manifest_entry().
 */
#ifndef DEPGAN_B200_H
#define DEPGAN_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define DEPGAN_MODEL_GEN 0     /* Gen_UNet2D   TG:349-498 / TU:291-428 */
#define DEPGAN_MODEL_CRITIC 1  /* Dis_C2D_FCN1 TG:316-345 */

#define DEPGAN_PREC_FP32 0 /* fp32 activations, fp32 CUDA-core implicit GEMM (the <=1e-4 variant) */
#define DEPGAN_PREC_BF16 1 /* bf16 activations, tcgen05/TMEM implicit GEMM, fp32 accumulate (<=1e-2) */
#define DEPGAN_PREC_F16 2  /* IEEE-half activations and weights, same tcgen05 kernels and rate (kind::f16), fp32 accumulate:
                              generator inference handles only (training == 0); DEM error ~7x below the bf16 path */
#define DEPGAN_PREC_F16X3 3 /* the tensor-core <= 1e-4 variant: activations and weights as (hi, lo) IEEE-half pairs, every
                               convolution as the three tcgen05 products x_hi*w_hi + x_lo*w_hi + x_hi*w_lo with fp32 accumulate
                               (22 significant bits per stored value); generator inference handles only, H and W multiples
                               of 128 */

#define DEPGAN_HEAD_TANH 0    /* DEP-GAN generator TG:494-495 */
#define DEPGAN_HEAD_SOFTMAX 1 /* DEP-UResNet TU:423-424 */

typedef struct depgan_cfg {
  int H, W;       /* slice size (256,256 in the reference, TG:40); multiples of 16 */
  int nicg;       /* generator input channels TG:22 (critic: ignored, always 1) */
  int nc_out;     /* generator output channels: 1 = tanh head, 4 = softmax head */
  int noise_len;  /* noiseSize TG:41 (32) */
  int max_batch;  /* workspace is sized for this many slices per call */
  int precision;  /* DEPGAN_PREC_* */
  int training;   /* 0: inference; 1: DEP-GAN train graphs (learning phase 0);
                     2: Keras training phase (batch-statistic BN + Dropout) for the DEP-UResNet fit path */
} depgan_cfg;

typedef struct depgan_net depgan_net; /* opaque: one generator or one critic */

const char* depgan_last_error(void);
int depgan_abi_version(void);

/* ---- weight manifest (host only; mirrors Keras layer/weight names of the reference) ---- */
int depgan_manifest_count(int model, const depgan_cfg* cfg);
/* name: "<keras layer>/<weight>" e.g. "conv2d_gen_0/kernel"; shape has ndim (<=4) entries;
 * offset = float offset into the flat parameter buffer; trainable = 0 for BN moving statistics. */
int depgan_manifest_entry(int model, const depgan_cfg* cfg, int idx, char* name, int name_cap, int* ndim,
                          int* shape, long long* offset, int* trainable);
long long depgan_manifest_floats(int model, const depgan_cfg* cfg); /* flat buffer length incl. padding */

/* ---- network handles ---- */
long long depgan_workspace_bytes(int model, const depgan_cfg* cfg);
/* Replaces Gen_UNet2D(...) TG:520/EG:380/TU:579/EU:399 and Dis_C2D_FCN1(...) TG:513,516.
 * params_dev: flat parameters (caller-owned, may be updated in place by depgan_adam_step);
 * grads_dev: flat gradient buffer of the same length (may be NULL when cfg->training == 0). */
depgan_net* depgan_net_create(int model, const depgan_cfg* cfg, float* params_dev, float* grads_dev,
                              void* workspace_dev, long long workspace_bytes);
void depgan_net_destroy(depgan_net* h);
/* Re-derive folded BN scale/shift and packed (bf16 / dgrad) weights after the flat parameters changed.
 * Replaces the implicit variable read of every session.run; call after load_weights (EG:383) / Adam. */
int depgan_net_prepare(depgan_net* h, void* stream);

/* ---- forward: model.predict([x, z]) EG:621, EU:558, TG:848 ; critic.predict(x) TG:846-848 ---- */
int depgan_gen_forward(depgan_net* g, const float* x_dev, const float* z_dev, float* out_dev, int n, void* stream);
int depgan_critic_forward(depgan_net* d, const float* x_dev, float* out_dev, int n, void* stream);

/* ---- train-step graphs (gradients only; the optimizer is depgan_adam_step) ----
 * depgan_critic_grads: netD_y2_train (which=0, TG:532-552) / netD_dem_train (which=1, TG:554-571) without the
 *   update: runs G forward (phase 0), critic on real/fake/mixed, gradient penalty double-backward, writes
 *   d loss / d theta_D into d's gradient buffer and {loss_real, loss_fake, grad_penalty, loss} to out4_dev.
 * depgan_gen_eval: netG_no_update TG:595-596 -> {loss, loss_fake, loss_fake_dem, M1, M3, M4} (out6_dev);
 *   sums_dev (3 doubles: sum wmh_real, sum wmh_fake, sum intersection) lets data-parallel callers all-reduce
 *   the batch-global dice/volume terms; global_n = global batch for the means.
 * depgan_gen_grads: netG_train TG:597-598 without the update. */
int depgan_critic_grads(depgan_net* d, depgan_net* g, int which, const float* real2_dev, const float* x1_dev,
                        const float* z_dev, const float* ep_dev, float* out4_dev, int n, int global_n, void* stream);
int depgan_gen_eval(depgan_net* g, depgan_net* dy2, depgan_net* ddem, const float* x1_dev, const float* real2_dev,
                    const float* z_dev, float thr, float* out6_dev, double* sums_dev, int n, int global_n,
                    void* stream);
int depgan_gen_grads(depgan_net* g, depgan_net* dy2, depgan_net* ddem, const float* x1_dev, const float* real2_dev,
                     const float* z_dev, float thr, float* out6_dev, double* sums_dev, int n, int global_n,
                     void* stream);
/* Final generator-loss terms from (possibly all-reduced) partial sums: out6 = combine(partials). */
int depgan_gen_loss_finalize(float* out6_dev, const double* sums_dev, void* stream);

/* ---- DEP-UResNet supervised step: my_network.fit(...) TU:602-606 on the model compiled at TU:427 ----
 * Keras *training* phase: BatchNormalization uses (and back-propagates through) the statistics of the batch and
 * refreshes moving_mean / moving_variance in the parameter buffer (momentum 0.99, Bessel-corrected variance);
 * Dropout(0.25) `do_gen_1` after conv2d_gen_10 (TU:388) uses the caller's keep mask (n, H/4, W/4, 96) of 0/1 bytes;
 * loss = mean categorical cross-entropy of the softmax output vs onehot (n,H,W,nc_out), written to loss_dev; the
 * gradient of every trainable tensor is left in the gradient buffer (apply depgan_adam_step with Keras' defaults
 * lr 1e-4, beta_1 0.9, beta_2 0.999, then depgan_net_prepare).  The handle must be created with cfg.training = 2.
 * depgan_cce_loss: the same loss for a given softmax output (validation, TU:606); dseg_scratch: npix*nc floats. */
int depgan_uresnet_grads(depgan_net* g, const float* x_dev, const float* z_dev, const float* onehot_dev,
                         const unsigned char* drop_keep_dev, float* loss_dev, int n, void* stream);
int depgan_cce_loss(const float* prob_dev, const float* onehot_dev, float* dseg_scratch_dev, float* loss_dev,
                    long long npix, int nc, float inv_total, void* stream);

/* Keras-form Adam (keras.optimizers.Adam.get_updates; call sites TG:549,568,594; TU:427):
 * lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m=b1*m+(1-b1)*g; v=b2*v+(1-b2)*g^2; p -= lr_t*m/(sqrt(v)+eps).
 * grad_scale multiplies g first (1/world after a sum all-reduce). t is 1-based. */
int depgan_adam_step(float* params_dev, const float* grads_dev, float* m_dev, float* v_dev, long long n, int t,
                     float lr, float beta1, float beta2, float eps, float grad_scale, void* stream);

/* ---- inference accumulation + DEM post-processing (bit-exact integer/label work) ----
 * depgan_dem_accumulate: acc_f64 += (double)(pred_f32 * mask_f32)      EG:622-624 / EU:559-560
 *   (mask index = i / chan, so a (Z,H,W) mask broadcasts over nc_out channels).
 * depgan_dem_postproc: dem = acc/n_repeat; fake2 = clip(base + dem, -1, 1); count = #((fake2 > thr) * mask != 0);
 *   label 1/2/3 = shrink/grow/stay                                     EG:628, 673-686, 711-741
 *   base is channel 0 of x (stride nicg floats).  count_dev: one unsigned long long (zeroed by the call).
 * depgan_uresnet_labels: mean = acc/n_repeat (optional output, IEEE division as NumPy), argmax over the
 *   chan mean probabilities (first max wins) + count(label>0)                      EU:564, 570, 597-600 */
int depgan_dem_accumulate(double* acc_dev, const float* pred_dev, const float* mask_dev, long long n, int chan,
                          void* stream);
int depgan_dem_postproc(const float* x_dev, int nicg, const double* acc_dev, double n_repeat,
                        const float* mask_dev, double thr, double* dem_out_dev, double* fake2_out_dev,
                        unsigned char* labels_dev, unsigned long long* count_dev, long long npix, void* stream);
int depgan_uresnet_labels(const double* acc_dev, double n_repeat, int chan, double* mean_out_dev,
                          unsigned char* labels_dev, unsigned long long* count_dev, long long npix, void* stream);

/* ---- data-parallel DEP-UResNet fit: synchronised BatchNorm ----
 * With a hook installed, depgan_uresnet_grads sums every BatchNorm layer's batch statistics (forward: 2*C doubles,
 * backward reduction: 2*C floats, both in the handle's workspace) over the ranks through `fn` before using them,
 * counts world * n samples, and scales the loss / gradient seed by the global pixel count; the caller then sums the
 * flat gradient bucket and the loss over the ranks.  fn must enqueue an in-place sum all-reduce of `count` elements at
 * dev_ptr on `stream` (or order it after that stream) and return 0.  world <= 1 or fn == NULL removes the hook.
 * Process-global: one process drives one GPU. */
typedef int (*depgan_allreduce_fn)(void* user, void* dev_ptr, long long count, int is_f64, void* stream);
int depgan_set_sync_hook(depgan_allreduce_fn fn, void* user, int world);

/* ---- evaluation of a predicted label map against the ground-truth one (EG:688-807 / EU:606-704) ----
 * conf16[4*real + fake] = number of voxels with ground-truth label `real` and predicted label `fake` (labels 0..3:
 * background / shrink / grow / stay).  The six Dice scores of the reference's CSV row are ratios of sums of these
 * integers, so the row is exact given the label maps.  conf16_dev: 16 unsigned long long (zeroed by the call). */
int depgan_label_confusion(const unsigned char* fake_labels_dev, const unsigned char* real_labels_dev, long long n,
                           unsigned long long* conf16_dev, void* stream);

/* ---- introspection for tests / profiling ---- */
/* Number of kernels this library launched since load (all handles). */
long long depgan_launch_count(void);
/* Per-launch device timing of the convolution kernels (bench.py roofline leg): between begin and end every
 * convolution launch is bracketed by CUDA events on its stream.  Classes: 0 = tcgen05 3x3, 1 = tcgen05 5x5,
 * 2 = tcgen05 1x1 / transposed conv, 3 = fp32 CUDA-core conv, 4 = CUDA-core weight gradient, 5 = tcgen05 weight
 * gradient, 6 = tcgen05 conv with the fused 2x2 max-pool epilogue.  flops/bytes are the algorithmic figures of
 * DESIGN.md (2*k*k*Cin*Cout per pixel; activation bytes read + written once). */
int depgan_profile_begin(void);
int depgan_profile_end(double* ms_by_class, double* flops_by_class, double* bytes_by_class,
                       long long* launches_by_class, int ncls);
/* Copy an internal activation of the last forward over n slices (by keras layer suffix, e.g. "gen_0",
 * "gen_noise_m1" = the ResBlock sum, "de_gen_9", "conv2d_dis_3", or "film" = all FiLM vectors) to out_dev as
 * float32. */
int depgan_debug_activation(depgan_net* h, const char* name, float* out_dev, long long cap_floats,
                            long long* n_floats, int n, void* stream);

/* depgan_critic_grads with the generator output given (dem_dev (n,H,W,1) float32 = G(x1, z)): the generator's weights do
 * not change during the critic updates of a generator iteration (TG:796-829), so all their generator forwards can run as
 * one batched depgan_gen_forward on the concatenated batches and be fed here slice by slice. */
int depgan_critic_grads_dem(depgan_net* d, int nicg, int which, const float* real2_dev, const float* x1_dev,
                            const float* dem_dev, const float* ep_dev, float* out4_dev, int n, int global_n,
                            void* stream);
/* netG_no_update for k noise candidates of ONE batch in a single pass (TG:868-874: the ten evaluations of a generator
 * iteration share x / real_2tp and differ only in the noise).  z_all (k, n, L, 1); out6_all (k, 6); sums_all (k, 8);
 * scratch: k * n * H * W * (nicg + 1) floats.  Any prepared handles with max_batch >= k * n serve (e.g. inference handles
 * created on the training networks' parameter buffers).  Every candidate's six values equal those of depgan_gen_eval.
 * Data parallel: sum sums_all over the ranks (depgan_dp_allreduce_f64, 8 * k doubles), then
 * depgan_gen_loss_finalize_multi. */
int depgan_gen_eval_multi(depgan_net* g, depgan_net* dy2, depgan_net* ddem, const float* x1_dev, const float* real2_dev,
                          const float* z_all_dev, float thr, float* out6_all_dev, double* sums_all_dev, float* scratch_dev,
                          int n, int k, int global_n, void* stream);
int depgan_gen_loss_finalize_multi(float* out6_all_dev, double* sums_all_dev, int k, int global_n, long long hw,
                                   void* stream);

/* ---- data-parallel update without Python in the loop (depgan_allreduce_attach; TG:549, 568, 594 applied
 * to the global batch) ----
 * After depgan_critic_grads / depgan_gen_grads (called with the GLOBAL batch size) each rank's gradient bucket holds its
 * shard's share of the global-batch gradient.  depgan_dp_update sums the buckets over the ranks, applies Keras' Adam
 * (step count t >= 1; lr_t = lr * sqrt(1 - beta_2^t) / (1 - beta_1^t), eps added to sqrt(v)) to params / m / v and calls
 * depgan_net_prepare -- one call per update.  extra_f64 (optional, <= 32 doubles on the device): loss partial sums that
 * are summed over the ranks in the same pass, in place.  Transport = whatever is attached to the handle:
 *   depgan_peer_attach      one node: a fused reduce + Adam kernel that reads every rank's mailbox over NVLink (CUDA IPC);
 *                           the summation order is the rank order on every replica, so the replicas stay bit-identical
 *   depgan_allreduce_attach any topology: ncclAllReduce on `stream` (NCCL is resolved at run time with dlsym from the
 *                           library the process already uses, or from the path given to depgan_nccl_load), then Adam
 *   neither                 single GPU: the local bucket
 * Mailboxes: depgan_peer_create(n_floats = depgan_manifest_floats, world <= 16, rank) allocates this rank's mailbox on
 * the current device; depgan_peer_handle copies its 64-byte CUDA IPC handle out; the caller exchanges the handles by any
 * means (MPI, files, torch.distributed) and passes all `world` of them, in rank order, to depgan_peer_connect. */
typedef struct depgan_peer depgan_peer;
int depgan_nccl_load(const char* path_or_null);
int depgan_nccl_unique_id(void* id128);
int depgan_nccl_init(void** comm_out, int world, int rank, const void* id128);
int depgan_nccl_destroy(void* comm);
int depgan_allreduce_attach(depgan_net* h, void* nccl_comm, int world);
depgan_peer* depgan_peer_create(long long n_floats, int world, int rank);
int depgan_peer_handle(depgan_peer* p, void* out64);
int depgan_peer_connect(depgan_peer* p, const void* handles_world_x_64);
void depgan_peer_destroy(depgan_peer* p);
int depgan_peer_attach(depgan_net* h, depgan_peer* p);
int depgan_dp_update(depgan_net* h, float* m_dev, float* v_dev, int t, float lr, float beta_1, float beta_2, float eps,
                     double* extra_f64, int n_extra, void* stream);
/* The sum of the buckets over the ranks left in the bucket (no optimizer step); extra_f64 as above. */
int depgan_dp_allreduce_grads(depgan_net* h, double* extra_f64, int n_extra, void* stream);
/* In-place sum over the ranks of n <= 128 doubles on the device: the loss partial sums of the forward-only evaluations
 * (TG:868-877 -- every rank must see the same ten candidate losses to pick the same noise). */
int depgan_dp_allreduce_f64(depgan_net* h, double* buf_dev, int n, void* stream);

/* ---- kernel-level entry points (parity tests and micro-benchmarks of the convolution kernels) ----
 * One fused 'same' stride-1 convolution (Keras Conv2D TG:285-304 / Conv2DTranspose k2s2 TG:307-312 when
 * deconv=1) on caller-owned device buffers.  Epilogue order: v = acc*scale+shift; out_pre=v;
 * FiLM: v = relu(v*film_g[n,c]+film_b[n,c]) + res; v += add_src; v = mask_src>0 ? v : 0; relu; out = v;
 * head: head_out = act(v . head_w + head_b) (act 0 tanh, 1 softmax, 2 linear).
 * use_tc=1: tcgen05/TMEM path (bf16 in/out, w_bf16 [taps][Ncols][Cin]); use_tc=0: fp32 CUDA-core path
 * (w_f32 [taps][Cin][Cout], any dtype combination). */
typedef struct depgan_conv_desc {
  const void* in0; const void* in1; int C0, C1;
  const float* w_f32; const void* w_bf16;
  const float* scale; const float* shift;
  void* out; void* out_pre;
  const float* film_g; const float* film_b; int film_stride; const void* res;
  const void* add_src; const void* mask_src;
  int relu, deconv;
  const float* head_w; const float* head_b; float* head_out; int head_nc, head_act;
  int N, H, W, Cout, ks;
  int in_bf16, out_bf16, use_tc;
  void* pool_out; /* optional (N,H/2,W/2,Cout), tcgen05 path: 2x2 stride-2 max-pool of `out`, fused into the epilogue */
} depgan_conv_desc;
int depgan_op_conv2d(const depgan_conv_desc* d, void* stream);
/* Host-only query of the tcgen05 convolution planner for the layer `d` describes (pointers are only tested for NULL,
 * nothing is launched; works without a GPU).  Returns 1 and fills plan16 when the tcgen05 path takes the layer, 0 when
 * it does not, <0 on bad arguments.  plan16 = { kc, ncta, nchunks, na, nb, b_tps, b_resident, acc_stages, n_issuers,
 * ch, n_side, tmem_cols, smem_bytes, pool, stage_out, nsplit } (shared-memory ring / TMEM geometry). */
int depgan_op_conv_plan(const depgan_conv_desc* d, int* plan16);
/* Weight gradient of one convolution: dw[tap][Cin][Cout] (fp32, Keras HWIO order) += sum_p x[p+off(tap)] (x) dy[p].
 * use_tc=1: tcgen05 path (x, dy bf16); use_tc=0: fp32 CUDA-core path (x, dy float32); use_tc=2: CUDA-core path with
 * x float32 and dy bf16 (the first layer of a bf16 network).  The caller zeroes dw. */
int depgan_op_wgrad(const void* x0, const void* x1, int C0, int C1, const void* dy, float* dw, int N, int H, int W,
                    int Cout, int ks, int use_tc, void* stream);
/* tcgen05 weight gradient (x, dy bf16) that also accumulates the bias gradient csum[co] += sum_p dy[p][co] from the same
 * pass over dy (the caller zeroes dw and csum). */
int depgan_op_wgrad_csum(const void* x0, const void* x1, int C0, int C1, const void* dy, float* dw, float* csum, int N,
                         int H, int W, int Cout, int ks, void* stream);
int depgan_op_pack_weights(const float* w_f32_dev, void* w_bf16_dev, int taps, int cin, int cout, void* stream);
int depgan_op_f32_to_bf16(const float* src_dev, void* dst_dev, long long n, void* stream);
/* float32 -> IEEE float16 (round to nearest even), 16-byte aligned buffers: the opt-in narrow device->host output of
 * predict() (an extension; the Keras contract, EG:621 / EU:558, returns float32 and stays the default). */
int depgan_op_f32_to_f16(const float* src_dev, void* dst_dev, long long n, void* stream);
int depgan_op_bf16_to_f32(const void* src_dev, float* dst_dev, long long n, void* stream);

#ifdef __cplusplus
}
#endif
#endif
