// conv_tc_kernel instantiations for 5x5 kernels (see conv_tc_kernel.cuh).
#include "conv_tc_kernel.cuh"

namespace convtc {

// epi = epilogue variant: 0 plain, 1 FiLM residual, 2 add / mask sources, 4 plain + fused 2x2 max-pool
int launch_ks5(int grid, uint32_t smem, cudaStream_t st, const TcMaps& tm, const ConvArgs& a, const TcGeom& g, int epi) {
  switch (epi * 100 + (g.kc / 16) * 10 + (g.b_resident ? 1 : 0)) {
    DG_TC_CASES(5, 0)
    DG_TC_CASES(5, 2)
    DG_TC_CASES(5, 4)
    default: depgan_set_error("conv_fwd_tc: no kernel for this (ks, kc, epi)"); return -2;
  }
}

int set_attrs_ks5() {
  DG_TC_ATTRS(5, 0)
  DG_TC_ATTRS(5, 2)
  DG_TC_ATTRS(5, 4)
  return 0;
}

}  // namespace convtc
