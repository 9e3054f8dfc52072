// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M = 128, K = 16, SS mode, 64-byte swizzle) as a function of N, of the
// accumulator pattern (same D, two / four alternating Ds), of the D column offset and of the accumulate flag.
// One CTA per SM, thread 0 issues REPS instructions back to back and waits for the commit.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// all parameters are compile-time so that the descriptors stay on the uniform datapath (warp-converged issue)
template <int N, int ND, int DOFF, int DSTRIDE, int ACCUM, int ASTEP, int CEVERY = 0>
__global__ void __launch_bounds__(128, 1) mma_rate(long long* out, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[2];  // CEVERY > 0: a commit to these after every CEVERY instructions (never waited for)
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t hi = ((uint32_t)(8 * 64) >> 4) | (1u << 14) | (4u << 29);
    const uint32_t a_lo = ((base & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t b_lo = (((base + 16384u) & 0x3FFFFu) >> 4) | (1u << 16);
    const long long t0 = clock64();
    for (int r = 0; r < reps; r += 8) {
      if (elect_one()) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint32_t d = tmem + (uint32_t)DOFF + (uint32_t)((u % ND) * DSTRIDE);
          const uint64_t da = ((uint64_t)hi << 32) | (a_lo + (uint32_t)((u & 3) * ASTEP));
          const uint64_t db = ((uint64_t)hi << 32) | (b_lo + (uint32_t)((u & 1) * 2));
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc),
              "r"((uint32_t)ACCUM)
              : "memory");
          if (CEVERY > 0 && (u + 1) % CEVERY == 0) {
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2[0])) : "memory");
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2[1])) : "memory");
          }
        }
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (elect_one())
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    __syncwarp();
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    const long long t2 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int N, int ND, int DOFF, int DSTRIDE, int ACCUM, int ASTEP, int CEVERY = 0>
void run(long long* d) {
  const int reps = 512;
  cudaFuncSetAttribute(mma_rate<N, ND, DOFF, DSTRIDE, ACCUM, ASTEP, CEVERY>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int rep = 0; rep < 2; ++rep) mma_rate<N, ND, DOFF, DSTRIDE, ACCUM, ASTEP, CEVERY><<<148, 128, 64 * 1024>>>(d, reps);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-6d %-4d %-6d %-8d %-6d %-6d c%-2d | %10.1f %10.1f  %s\n", N, ND, DOFF, DSTRIDE, ACCUM, ASTEP, CEVERY, (double)h[0] / reps,
         (double)h[1] / reps, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  printf("%-6s %-4s %-6s %-8s %-6s %-6s | %10s %10s\n", "N", "nd", "d_off", "d_stride", "accum", "a_step", "clk/MMA(issue)", "clk/MMA(done)");
  run<32, 1, 0, 0, 1, 0>(d);   run<32, 2, 0, 32, 1, 0>(d);   run<32, 4, 0, 32, 1, 0>(d);   run<32, 2, 0, 32, 1, 4>(d);
  run<64, 1, 0, 0, 1, 0>(d);   run<64, 2, 0, 64, 1, 0>(d);   run<64, 4, 0, 64, 1, 0>(d);
  run<96, 1, 0, 0, 1, 0>(d);   run<96, 2, 0, 96, 1, 0>(d);   run<96, 2, 0, 256, 1, 0>(d);  run<96, 4, 0, 96, 1, 0>(d);
  run<96, 1, 32, 0, 1, 0>(d);  run<96, 2, 32, 256, 1, 0>(d); run<96, 2, 0, 256, 0, 0>(d);  run<96, 2, 0, 256, 1, 4>(d);
  run<96, 2, 0, 256, 1, 64>(d); run<96, 8, 0, 32, 1, 4>(d);  run<96, 4, 32, 32, 1, 4>(d);
  run<128, 1, 0, 0, 1, 0>(d);  run<128, 2, 0, 128, 1, 0>(d); run<128, 2, 0, 256, 1, 0>(d);
  run<192, 1, 0, 0, 1, 0>(d);  run<192, 2, 0, 256, 1, 0>(d);
  run<256, 1, 0, 0, 1, 0>(d);  run<256, 2, 0, 256, 1, 0>(d);
  run<16, 1, 0, 0, 1, 0>(d);   run<16, 1, 0, 0, 1, 2>(d);    run<48, 1, 0, 0, 1, 0>(d);    run<80, 1, 0, 0, 1, 0>(d);   run<80, 1, 0, 0, 1, 2>(d);
  run<160, 1, 0, 0, 1, 0>(d);  run<160, 1, 0, 0, 1, 4>(d);   run<192, 1, 0, 0, 1, 4>(d);   run<192, 1, 0, 0, 1, 8>(d);
  // two commits after every 4 / 8 instructions (the row kernels commit "stage free" + "row done" per input row)
  run<80, 1, 0, 0, 1, 2, 4>(d);  run<80, 1, 0, 0, 1, 2, 8>(d);  run<192, 1, 0, 0, 1, 4, 4>(d);  run<192, 1, 0, 0, 1, 4, 8>(d);
  run<16, 1, 0, 0, 1, 2, 1>(d);  run<96, 1, 0, 0, 1, 4, 2>(d);
  return 0;
}
