// conv_tc_kernel instantiations for 1x1 kernels (see conv_tc_kernel.cuh).
#include "conv_tc_kernel.cuh"

namespace convtc {

// epi = side inputs of the call: 0 none, 1 FiLM residual, 2 add / mask sources
int launch_ks1(int grid, uint32_t smem, cudaStream_t st, const TcMaps& tm, const ConvArgs& a, const TcGeom& g, int epi) {
  switch (epi * 100 + (g.kc / 16) * 10 + (g.b_resident ? 1 : 0)) {
    DG_TC_CASES(1, 0)
    DG_TC_CASES(1, 2)
    DG_TC_CASES_F16(1, 0)
    default: depgan_set_error("conv_fwd_tc: no kernel for this (ks, kc, epi)"); return -2;
  }
}

int set_attrs_ks1() {
  DG_TC_ATTRS(1, 0)
  DG_TC_ATTRS(1, 2)
  DG_TC_ATTRS_F16(1, 0)
  return 0;
}

}  // namespace convtc
