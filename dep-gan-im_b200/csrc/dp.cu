// Data-parallel weight update behind the C ABI: after depgan_critic_grads / depgan_gen_grads every
// rank holds the gradient of ITS shard of the batch (already scaled by 1 / global batch, TG:540-547, 592) in the flat
// gradient bucket; the update a single GPU would apply to the whole batch needs the sum over the ranks, then Keras'
// Adam (TG:549, 568, 594), then the re-packing of the derived tensors.  Two transports:
//
//  * depgan_peer_*  (one node, NVLink / NVSwitch): ONE fused pass.  Every rank owns a cudaMalloc'ed mailbox that the
//    other processes map through CUDA IPC.  publish: the bucket (+ a few loss partial sums) is copied into the mailbox and
//    the last CTA raises this rank's epoch flag in every peer's memory (system-scope release).  reduce + Adam: every CTA
//    waits until all ranks' flags reached the epoch, then reads ALL mailboxes straight over NVLink, sums them in rank
//    order (so every replica computes bit-identical sums) and applies the Adam update to its own parameters in the same
//    loop -- the reduced gradient never exists in memory and there is no separate optimizer kernel.  Mailboxes are
//    double buffered by epoch parity, which makes the "everyone has finished reading" barrier implicit: a rank can only
//    publish epoch e + 2 after its own reduce of e + 1, which waited for every peer's publish of e + 1, which those peers
//    issued after their reduce of e.
//  * depgan_nccl_* / depgan_allreduce_attach  (any topology): ncclAllReduce on the caller's stream, resolved with dlsym
//    from the NCCL the process already has (libtorch's, or the one a C caller links), followed by the Adam kernel.
//
// depgan_dp_update is the one call a trainer makes per update in either mode; depgan_dp_allreduce_f64 sums the loss
// partial sums of the forward-only evaluations (TG:868-877: every rank must pick the same noise).
#include <dlfcn.h>

#include <cmath>
#include <cstring>

#include "net.h"

namespace {

constexpr int DP_MAX_WORLD = 16;
constexpr int DP_EXTRA = 32;   // doubles of side data (loss partial sums) that ride along with the bucket
constexpr int DP_SMALL = 128;  // doubles of one small all-reduce (the sums of ten candidate evaluations: 80)

// ---------------------------------------------------------------------------------------------------------
// NCCL through dlsym (no link-time dependency: the library stays loadable without NCCL)
// ---------------------------------------------------------------------------------------------------------
struct NcclId { char b[128]; };
typedef int (*nccl_get_id_t)(NcclId*);
typedef int (*nccl_init_rank_t)(void**, int, NcclId, int);
typedef int (*nccl_destroy_t)(void*);
typedef int (*nccl_allreduce_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*nccl_errstr_t)(int);
struct NcclApi {
  bool tried = false, ok = false;
  nccl_get_id_t get_id = nullptr;
  nccl_init_rank_t init_rank = nullptr;
  nccl_destroy_t destroy = nullptr;
  nccl_allreduce_t allreduce = nullptr;
  nccl_errstr_t errstr = nullptr;
} g_nccl;
constexpr int NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8, NCCL_SUM = 0;

int nccl_load(const char* path) {
  if (g_nccl.ok) return 0;
  g_nccl.tried = true;
  void* h = nullptr;
  if (path && *path) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h && dlsym(RTLD_DEFAULT, "ncclAllReduce")) h = RTLD_DEFAULT;
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // already mapped by the host application?
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    depgan_set_error("NCCL not found (dlopen libnccl.so.2); pass its path to depgan_nccl_load");
    return -1;
  }
  g_nccl.get_id = (nccl_get_id_t)dlsym(h, "ncclGetUniqueId");
  g_nccl.init_rank = (nccl_init_rank_t)dlsym(h, "ncclCommInitRank");
  g_nccl.destroy = (nccl_destroy_t)dlsym(h, "ncclCommDestroy");
  g_nccl.allreduce = (nccl_allreduce_t)dlsym(h, "ncclAllReduce");
  g_nccl.errstr = (nccl_errstr_t)dlsym(h, "ncclGetErrorString");
  if (!g_nccl.get_id || !g_nccl.init_rank || !g_nccl.destroy || !g_nccl.allreduce) {
    depgan_set_error("the NCCL library lacks ncclGetUniqueId / ncclCommInitRank / ncclCommDestroy / ncclAllReduce");
    return -1;
  }
  g_nccl.ok = true;
  return 0;
}
int nccl_check(int rc, const char* what) {
  if (rc == 0) return 0;
  depgan_set_error(std::string(what) + ": " + (g_nccl.errstr ? g_nccl.errstr(rc) : "NCCL error") + " (" + std::to_string(rc) + ")");
  return -1;
}

// ---------------------------------------------------------------------------------------------------------
// peer-memory kernels
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct PeerView {  // passed by value to the kernels
  float* box[DP_MAX_WORLD];        // mailbox of this epoch's parity on every rank (own included), as mapped here
  unsigned* flags[DP_MAX_WORLD];   // epoch flags of every rank: flags[r][s] = last epoch rank s published (bucket kind)
  int world, rank;
};

// bucket -> own mailbox (16-byte vectors; the tail and the extras by the last threads), then the LAST CTA to finish tells
// every rank that this rank's mailbox holds epoch `epoch`
__global__ void __launch_bounds__(256) dp_publish_kernel(const float* __restrict__ grads, long long n, const double* extra,
                                                         int n_extra, PeerView pv, unsigned epoch, unsigned* counter) {
  float* box = pv.box[pv.rank];
  const long long n4 = n >> 2;
  const float4* src = reinterpret_cast<const float4*>(grads);
  float4* dst = reinterpret_cast<float4*>(box);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i];
  if (blockIdx.x == 0) {
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) box[i] = grads[i];
    double* ex = reinterpret_cast<double*>(box + ((n + 3) & ~3LL));
    for (int i = threadIdx.x; i < n_extra; i += blockDim.x) ex[i] = extra[i];
  }
  __threadfence_system();  // this thread's mailbox writes are visible to the peers before the counter moves
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(counter, 1u);
    if (done == gridDim.x - 1) {
      *counter = 0;
      __threadfence_system();
      for (int r = 0; r < pv.world; ++r) st_release_sys(pv.flags[r] + pv.rank, epoch);
    }
  }
}

// waits for every rank's epoch, then g = sum over ranks (rank order) and Keras Adam on the local replica, in one pass
__global__ void __launch_bounds__(256) dp_reduce_adam_kernel(PeerView pv, unsigned epoch, float* __restrict__ p,
                                                             float* __restrict__ m, float* __restrict__ v, long long n,
                                                             float lr_t, float b1, float b2, float eps, double* extra_out,
                                                             int n_extra, float* grads_out) {
  if (threadIdx.x < pv.world) {
    const unsigned* f = pv.flags[pv.rank] + threadIdx.x;
    while ((int)(ld_acquire_sys(f) - epoch) < 0) __nanosleep(64);
  }
  __syncthreads();
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 g = __ldcv(reinterpret_cast<const float4*>(pv.box[0]) + i);  // volatile loads: never a stale cached line
    for (int r = 1; r < pv.world; ++r) {
      const float4 h = __ldcv(reinterpret_cast<const float4*>(pv.box[r]) + i);
      g.x += h.x; g.y += h.y; g.z += h.z; g.w += h.w;
    }
    if (grads_out) reinterpret_cast<float4*>(grads_out)[i] = g;
    if (p) {
      float4 mi = reinterpret_cast<float4*>(m)[i], vi = reinterpret_cast<float4*>(v)[i], pi = reinterpret_cast<float4*>(p)[i];
      mi.x = b1 * mi.x + (1.f - b1) * g.x; vi.x = b2 * vi.x + (1.f - b2) * g.x * g.x; pi.x -= lr_t * mi.x / (sqrtf(vi.x) + eps);
      mi.y = b1 * mi.y + (1.f - b1) * g.y; vi.y = b2 * vi.y + (1.f - b2) * g.y * g.y; pi.y -= lr_t * mi.y / (sqrtf(vi.y) + eps);
      mi.z = b1 * mi.z + (1.f - b1) * g.z; vi.z = b2 * vi.z + (1.f - b2) * g.z * g.z; pi.z -= lr_t * mi.z / (sqrtf(vi.z) + eps);
      mi.w = b1 * mi.w + (1.f - b1) * g.w; vi.w = b2 * vi.w + (1.f - b2) * g.w * g.w; pi.w -= lr_t * mi.w / (sqrtf(vi.w) + eps);
      reinterpret_cast<float4*>(m)[i] = mi;
      reinterpret_cast<float4*>(v)[i] = vi;
      reinterpret_cast<float4*>(p)[i] = pi;
    }
  }
  if (blockIdx.x == 0) {
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
      float g = 0.f;
      for (int r = 0; r < pv.world; ++r) g += __ldcv(pv.box[r] + i);
      if (grads_out) grads_out[i] = g;
      if (p) {
        const float mi = b1 * m[i] + (1.f - b1) * g, vi = b2 * v[i] + (1.f - b2) * g * g;
        m[i] = mi; v[i] = vi;
        p[i] -= lr_t * mi / (sqrtf(vi) + eps);
      }
    }
    const long long eo = (n + 3) & ~3LL;
    for (int i = threadIdx.x; i < n_extra; i += blockDim.x) {
      double s = 0.0;
      for (int r = 0; r < pv.world; ++r) s += __ldcv(reinterpret_cast<const double*>(pv.box[r] + eo) + i);
      extra_out[i] = s;
    }
  }
}

// all-reduce of a few doubles (loss partial sums of a forward-only evaluation): one CTA, own small mailbox + own flags
__global__ void __launch_bounds__(DP_SMALL) dp_small_allreduce_kernel(PeerView pv, unsigned epoch, double* buf, int n) {
  double* mine = reinterpret_cast<double*>(pv.box[pv.rank]);
  if ((int)threadIdx.x < n) mine[threadIdx.x] = buf[threadIdx.x];
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < pv.world) {
    st_release_sys(pv.flags[threadIdx.x] + pv.rank, epoch);
    const unsigned* f = pv.flags[pv.rank] + threadIdx.x;
    while ((int)(ld_acquire_sys(f) - epoch) < 0) __nanosleep(64);
  }
  __syncthreads();
  if ((int)threadIdx.x < n) {
    double s = 0.0;
    for (int r = 0; r < pv.world; ++r) s += __ldcv(reinterpret_cast<const double*>(pv.box[r]) + threadIdx.x);
    buf[threadIdx.x] = s;
  }
}

}  // namespace

// One rank's mailbox and its view of the peers'.  Layout of the cudaMalloc'ed block (the same on every rank):
//   [bucket mailbox parity 0][parity 1][small mailbox parity 0][parity 1][bucket flags: world u32][small flags: world u32]
//   [CTA counter]
struct depgan_peer {
  int world = 1, rank = 0, dev = 0;
  long long n = 0;
  size_t box_bytes = 0, small_bytes = 0, total = 0;
  char* base = nullptr;
  char* peer_base[DP_MAX_WORLD] = {};
  bool connected = false;
  unsigned epoch = 0, epoch_small = 0;
  cudaIpcMemHandle_t handle;
  double* extra_dev = nullptr;  // staging of the extras on this rank (inside `base`, after the counter)
  size_t off_small() const { return 2 * box_bytes; }
  size_t off_flags() const { return off_small() + 2 * small_bytes; }
  size_t off_counter() const { return off_flags() + 2 * DP_MAX_WORLD * sizeof(unsigned); }
  PeerView view(bool small, unsigned ep) const {
    PeerView pv;
    pv.world = world; pv.rank = rank;
    for (int r = 0; r < world; ++r) {
      char* b = peer_base[r];
      pv.box[r] = reinterpret_cast<float*>(small ? b + off_small() + (ep & 1u) * small_bytes : b + (ep & 1u) * box_bytes);
      pv.flags[r] = reinterpret_cast<unsigned*>(b + off_flags()) + (small ? DP_MAX_WORLD : 0);
    }
    return pv;
  }
};

extern "C" {

// ---- NCCL transport --------------------------------------------------------------------------------------
int depgan_nccl_load(const char* path) { return nccl_load(path); }

int depgan_nccl_unique_id(void* id128) {
  DG_TRY(nccl_load(nullptr));
  DG_REQUIRE(id128 != nullptr, "nccl_unique_id: null buffer");
  NcclId id;
  DG_TRY(nccl_check(g_nccl.get_id(&id), "ncclGetUniqueId"));
  memcpy(id128, id.b, 128);
  return 0;
}

int depgan_nccl_init(void** comm, int world, int rank, const void* id128) {
  DG_TRY(nccl_load(nullptr));
  DG_REQUIRE(comm && id128 && world >= 1 && rank >= 0 && rank < world, "nccl_init: bad arguments");
  NcclId id;
  memcpy(id.b, id128, 128);
  return nccl_check(g_nccl.init_rank(comm, world, id, rank), "ncclCommInitRank");
}

int depgan_nccl_destroy(void* comm) {
  if (!comm || !g_nccl.ok) return 0;
  return nccl_check(g_nccl.destroy(comm), "ncclCommDestroy");
}

int depgan_allreduce_attach(depgan_net* h, void* nccl_comm, int world) {
  DG_REQUIRE(h != nullptr && world >= 1, "allreduce_attach: bad arguments");
  if (nccl_comm) DG_TRY(nccl_load(nullptr));
  h->nccl_comm = nccl_comm;
  h->dp_world = nccl_comm ? world : 1;
  return 0;
}

// ---- peer-memory transport -------------------------------------------------------------------------------
depgan_peer* depgan_peer_create(long long n_floats, int world, int rank) {
  if (n_floats < 0 || world < 1 || world > DP_MAX_WORLD || rank < 0 || rank >= world) {
    depgan_set_error("peer_create: bad arguments (world <= 16)");
    return nullptr;
  }
  depgan_peer* p = new depgan_peer();
  p->world = world; p->rank = rank; p->n = n_floats;
  cudaGetDevice(&p->dev);
  p->box_bytes = ((((size_t)n_floats + 3) & ~(size_t)3) * sizeof(float) + DP_EXTRA * sizeof(double) + 255) & ~(size_t)255;
  p->small_bytes = DP_SMALL * sizeof(double);
  p->total = p->off_counter() + 256 + DP_EXTRA * sizeof(double);
  if (cudaMalloc(&p->base, p->total) != cudaSuccess || cudaMemset(p->base, 0, p->total) != cudaSuccess ||
      cudaIpcGetMemHandle(&p->handle, p->base) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
    depgan_set_error(std::string("peer_create: ") + cudaGetErrorString(cudaGetLastError()));
    if (p->base) cudaFree(p->base);
    delete p;
    return nullptr;
  }
  p->extra_dev = reinterpret_cast<double*>(p->base + p->off_counter() + 256);
  p->peer_base[rank] = p->base;
  p->connected = world == 1;
  return p;
}

int depgan_peer_handle(depgan_peer* p, void* out64) {
  DG_REQUIRE(p && out64, "peer_handle: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  memcpy(out64, &p->handle, 64);
  return 0;
}

int depgan_peer_connect(depgan_peer* p, const void* handles) {
  DG_REQUIRE(p && handles, "peer_connect: null pointer");
  for (int r = 0; r < p->world; ++r) {
    if (r == p->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, reinterpret_cast<const char*>(handles) + 64 * r, 64);
    void* ptr = nullptr;
    DG_CHECK_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    p->peer_base[r] = reinterpret_cast<char*>(ptr);
  }
  p->connected = true;
  return 0;
}

void depgan_peer_destroy(depgan_peer* p) {
  if (!p) return;
  for (int r = 0; r < p->world; ++r)
    if (r != p->rank && p->peer_base[r]) cudaIpcCloseMemHandle(p->peer_base[r]);
  if (p->base) cudaFree(p->base);
  delete p;
}

int depgan_peer_attach(depgan_net* h, depgan_peer* p) {
  DG_REQUIRE(h != nullptr, "peer_attach: null handle");
  DG_REQUIRE(!p || (p->connected && p->n == h->man.total), "peer_attach: mailbox not connected or of another size");
  h->peer = p;
  if (p) h->dp_world = p->world;
  return 0;
}

// ---- the update ------------------------------------------------------------------------------------------
// Sums the gradient bucket of `h` over the ranks (peer mailboxes or NCCL, whichever is attached; neither: the local
// gradient is used), applies Keras' Adam with step count t to params / m / v and re-packs the derived tensors.
// extra_f64 (optional, n_extra <= 32 doubles, device): loss partial sums reduced in the same pass, in place.
int depgan_dp_update(depgan_net* h, float* m_dev, float* v_dev, int t, float lr, float beta_1, float beta_2, float eps,
                     double* extra_f64, int n_extra, void* stream) {
  DG_REQUIRE(h && h->grads && m_dev && v_dev && t >= 1, "dp_update: bad arguments (training handle, t >= 1)");
  DG_REQUIRE(n_extra >= 0 && n_extra <= DP_EXTRA && (n_extra == 0 || extra_f64), "dp_update: at most 32 extra doubles");
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = h->man.total;
  const float lr_t = lr * sqrtf(1.0f - powf(beta_2, (float)t)) / (1.0f - powf(beta_1, (float)t));
  if (h->peer && h->peer->world > 1) {
    depgan_peer* p = h->peer;
    const unsigned ep = ++p->epoch;
    const PeerView pv = p->view(false, ep);
    const int grid = 148 * 4;
    dp_publish_kernel<<<grid, 256, 0, st>>>(h->grads, n, extra_f64, n_extra, pv, ep,
                                            reinterpret_cast<unsigned*>(p->base + p->off_counter()));
    DG_LAUNCH_CHECK();
    dp_reduce_adam_kernel<<<grid, 256, 0, st>>>(pv, ep, h->params, m_dev, v_dev, n, lr_t, beta_1, beta_2, eps, extra_f64,
                                                n_extra, nullptr);
    DG_LAUNCH_CHECK();
  } else {
    if (h->nccl_comm && h->dp_world > 1) {
      DG_TRY(nccl_check(g_nccl.allreduce(h->grads, h->grads, (size_t)n, NCCL_FLOAT32, NCCL_SUM, h->nccl_comm, st),
                        "ncclAllReduce(gradient bucket)"));
      if (n_extra)
        DG_TRY(nccl_check(g_nccl.allreduce(extra_f64, extra_f64, (size_t)n_extra, NCCL_FLOAT64, NCCL_SUM, h->nccl_comm, st),
                          "ncclAllReduce(loss sums)"));
    }
    DG_TRY(k_adam(h->params, h->grads, m_dev, v_dev, n, lr_t, beta_1, beta_2, eps, 1.0f, st));
  }
  return depgan_net_prepare(h, stream);
}

// Sum of the gradient bucket over the ranks, left in the bucket (no optimizer step): for callers that inspect the
// global gradient.  extra_f64 as in depgan_dp_update.
int depgan_dp_allreduce_grads(depgan_net* h, double* extra_f64, int n_extra, void* stream) {
  DG_REQUIRE(h && h->grads, "dp_allreduce_grads: not a training handle");
  DG_REQUIRE(n_extra >= 0 && n_extra <= DP_EXTRA && (n_extra == 0 || extra_f64), "dp_allreduce_grads: at most 32 extras");
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = h->man.total;
  if (h->peer && h->peer->world > 1) {
    depgan_peer* p = h->peer;
    const unsigned ep = ++p->epoch;
    const PeerView pv = p->view(false, ep);
    const int grid = 148 * 4;
    dp_publish_kernel<<<grid, 256, 0, st>>>(h->grads, n, extra_f64, n_extra, pv, ep,
                                            reinterpret_cast<unsigned*>(p->base + p->off_counter()));
    DG_LAUNCH_CHECK();
    dp_reduce_adam_kernel<<<grid, 256, 0, st>>>(pv, ep, nullptr, nullptr, nullptr, n, 0.f, 0.f, 0.f, 0.f, extra_f64, n_extra,
                                                h->grads);
    DG_LAUNCH_CHECK();
  } else if (h->nccl_comm && h->dp_world > 1) {
    DG_TRY(nccl_check(g_nccl.allreduce(h->grads, h->grads, (size_t)n, NCCL_FLOAT32, NCCL_SUM, h->nccl_comm, st),
                      "ncclAllReduce(gradient bucket)"));
    if (n_extra)
      DG_TRY(nccl_check(g_nccl.allreduce(extra_f64, extra_f64, (size_t)n_extra, NCCL_FLOAT64, NCCL_SUM, h->nccl_comm, st),
                        "ncclAllReduce(loss sums)"));
  }
  return 0;
}

// In-place sum over the ranks of n <= 128 doubles on the device (loss partial sums of depgan_gen_eval).
int depgan_dp_allreduce_f64(depgan_net* h, double* buf_dev, int n, void* stream) {
  DG_REQUIRE(h && buf_dev && n >= 1 && n <= DP_SMALL, "dp_allreduce_f64: 1..128 doubles");
  cudaStream_t st = (cudaStream_t)stream;
  if (h->peer && h->peer->world > 1) {
    depgan_peer* p = h->peer;
    const unsigned ep = ++p->epoch_small;
    dp_small_allreduce_kernel<<<1, DP_SMALL, 0, st>>>(p->view(true, ep), ep, buf_dev, n);
    DG_LAUNCH_CHECK();
  } else if (h->nccl_comm && h->dp_world > 1) {
    DG_TRY(nccl_check(g_nccl.allreduce(buf_dev, buf_dev, (size_t)n, NCCL_FLOAT64, NCCL_SUM, h->nccl_comm, st),
                      "ncclAllReduce(loss sums)"));
  }
  return 0;
}

}  // extern "C"
