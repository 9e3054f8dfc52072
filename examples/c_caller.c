/* c_caller -- runs generator inference (model.predict([x, z]), EG:621 / EU:558) through the C ABI of
 * include/depgan_b200.h with no Python in the loop ("a C caller must be able to run cfg 1").
 *
 *   c_caller params.bin x.bin z.bin out.bin N nicg nc_out [bf16|fp32] [H W]
 *
 * params.bin : the flat float32 parameter buffer in the manifest layout (depgan_manifest_entry() offsets; Keras
 *              tensor layouts) -- what a host program fills from the netG_*.h5 / trained_depuresnet_*.h5 file
 * x.bin      : N*H*W*nicg float32 (NHWC), z.bin : N*32 float32, out.bin : N*H*W*nc_out float32 (written)
 *
 * Build (see __graft_entry__.build()):
 *   gcc -std=c99 -O2 -Iinclude -I/usr/local/cuda/include examples/c_caller.c -o examples/c_caller \
 *       -Ldep-gan-im_b200 -ldepgan_b200 -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,'$ORIGIN/../dep-gan-im_b200'
 * Everything fails loudly: there is no CPU path behind this program.
 */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "depgan_b200.h"

#define CK_CUDA(call)                                                                          \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) {                                                                   \
      fprintf(stderr, "c_caller: %s failed: %s\n", #call, cudaGetErrorString(e_));             \
      return 2;                                                                                \
    }                                                                                          \
  } while (0)
#define CK_DG(call)                                                                            \
  do {                                                                                         \
    int r_ = (call);                                                                           \
    if (r_ != 0) {                                                                             \
      fprintf(stderr, "c_caller: %s failed (%d): %s\n", #call, r_, depgan_last_error());       \
      return 3;                                                                                \
    }                                                                                          \
  } while (0)

static float* read_floats(const char* path, long long count) {
  FILE* f = fopen(path, "rb");
  float* p;
  if (!f) {
    fprintf(stderr, "c_caller: cannot open %s\n", path);
    return NULL;
  }
  p = (float*)malloc((size_t)count * sizeof(float));
  if (!p || (long long)fread(p, sizeof(float), (size_t)count, f) != count) {
    fprintf(stderr, "c_caller: %s does not hold %lld float32 values\n", path, count);
    free(p);
    fclose(f);
    return NULL;
  }
  fclose(f);
  return p;
}

int main(int argc, char** argv) {
  depgan_cfg cfg;
  depgan_net* g;
  long long n_par, ws_bytes, n_x, n_z, n_out;
  float *h_par, *h_x, *h_z, *h_out, *d_par, *d_x, *d_z, *d_out;
  void* d_ws;
  cudaStream_t st;
  cudaEvent_t e0, e1;
  float ms = 0.f;
  int n, i, reps = 5;
  FILE* fo;

  if (argc < 8) {
    fprintf(stderr, "usage: %s params.bin x.bin z.bin out.bin N nicg nc_out [bf16|fp32] [H W]\n", argv[0]);
    return 1;
  }
  n = atoi(argv[5]);
  memset(&cfg, 0, sizeof cfg);
  cfg.H = argc > 10 ? atoi(argv[9]) : 256;
  cfg.W = argc > 10 ? atoi(argv[10]) : 256;
  cfg.nicg = atoi(argv[6]);
  cfg.nc_out = atoi(argv[7]);
  cfg.noise_len = 32;
  cfg.max_batch = n;
  cfg.precision = (argc > 8 && strcmp(argv[8], "fp32") == 0) ? DEPGAN_PREC_FP32
                  : (argc > 8 && strcmp(argv[8], "f16") == 0) ? DEPGAN_PREC_F16 : DEPGAN_PREC_BF16;
  cfg.training = 0;

  n_par = depgan_manifest_floats(DEPGAN_MODEL_GEN, &cfg);
  ws_bytes = depgan_workspace_bytes(DEPGAN_MODEL_GEN, &cfg);
  if (n_par <= 0 || ws_bytes <= 0) {
    fprintf(stderr, "c_caller: bad configuration: %s\n", depgan_last_error());
    return 3;
  }
  n_x = (long long)n * cfg.H * cfg.W * cfg.nicg;
  n_z = (long long)n * cfg.noise_len;
  n_out = (long long)n * cfg.H * cfg.W * cfg.nc_out;
  h_par = read_floats(argv[1], n_par);
  h_x = read_floats(argv[2], n_x);
  h_z = read_floats(argv[3], n_z);
  h_out = (float*)malloc((size_t)n_out * sizeof(float));
  if (!h_par || !h_x || !h_z || !h_out) return 1;

  CK_CUDA(cudaSetDevice(0));
  CK_CUDA(cudaStreamCreate(&st));
  CK_CUDA(cudaMalloc((void**)&d_par, (size_t)n_par * sizeof(float)));
  CK_CUDA(cudaMalloc(&d_ws, (size_t)ws_bytes));
  CK_CUDA(cudaMalloc((void**)&d_x, (size_t)n_x * sizeof(float)));
  CK_CUDA(cudaMalloc((void**)&d_z, (size_t)n_z * sizeof(float)));
  CK_CUDA(cudaMalloc((void**)&d_out, (size_t)n_out * sizeof(float)));
  CK_CUDA(cudaMemcpyAsync(d_par, h_par, (size_t)n_par * sizeof(float), cudaMemcpyHostToDevice, st));
  CK_CUDA(cudaMemcpyAsync(d_x, h_x, (size_t)n_x * sizeof(float), cudaMemcpyHostToDevice, st));
  CK_CUDA(cudaMemcpyAsync(d_z, h_z, (size_t)n_z * sizeof(float), cudaMemcpyHostToDevice, st));

  /* Gen_UNet2D(...) + load_weights(...)  (EG:380-383) */
  g = depgan_net_create(DEPGAN_MODEL_GEN, &cfg, d_par, NULL, d_ws, ws_bytes);
  if (!g) {
    fprintf(stderr, "c_caller: depgan_net_create failed: %s\n", depgan_last_error());
    return 3;
  }
  CK_DG(depgan_net_prepare(g, (void*)st));

  /* model.predict([x, z])  (EG:621) */
  CK_DG(depgan_gen_forward(g, d_x, d_z, d_out, n, (void*)st));
  CK_CUDA(cudaStreamSynchronize(st));
  CK_CUDA(cudaEventCreate(&e0));
  CK_CUDA(cudaEventCreate(&e1));
  CK_CUDA(cudaEventRecord(e0, st));
  for (i = 0; i < reps; ++i) CK_DG(depgan_gen_forward(g, d_x, d_z, d_out, n, (void*)st));
  CK_CUDA(cudaEventRecord(e1, st));
  CK_CUDA(cudaStreamSynchronize(st));
  CK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  CK_CUDA(cudaMemcpy(h_out, d_out, (size_t)n_out * sizeof(float), cudaMemcpyDeviceToHost));

  fo = fopen(argv[4], "wb");
  if (!fo || (long long)fwrite(h_out, sizeof(float), (size_t)n_out, fo) != n_out) {
    fprintf(stderr, "c_caller: cannot write %s\n", argv[4]);
    return 1;
  }
  fclose(fo);
  printf("c_caller: %d slices of %dx%d, %s, %.3f ms per forward (%.0f slices/s), %lld kernel launches\n", n, cfg.H,
         cfg.W, cfg.precision == DEPGAN_PREC_BF16 ? "bf16" : cfg.precision == DEPGAN_PREC_F16 ? "f16" : "fp32", ms / reps, 1e3 * n * reps / ms,
         depgan_launch_count());

  depgan_net_destroy(g);
  cudaFree(d_par); cudaFree(d_ws); cudaFree(d_x); cudaFree(d_z); cudaFree(d_out);
  cudaStreamDestroy(st);
  free(h_par); free(h_x); free(h_z); free(h_out);
  return 0;
}
