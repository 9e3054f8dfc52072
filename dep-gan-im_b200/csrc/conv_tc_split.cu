// conv_tc_kernel instantiations for split-half storage (DT_F16S, the tensor-core <= 1e-4 variant; see conv_tc_kernel.cuh):
// the 1x1 kernel of the transposed convolution and the plain / FiLM 3x3 layers of generator inference.
#include "conv_tc_kernel.cuh"

namespace convtc {

int launch_split(int grid, uint32_t smem, cudaStream_t st, const TcMaps& tm, const ConvArgs& a, const TcGeom& g, int epi) {
  switch (a.ks * 1000 + epi * 100 + (g.kc / 16) * 10 + (g.b_resident ? 1 : 0)) {
    DG_TC_CASES_SPLIT(1, 0)
    DG_TC_CASES_SPLIT(3, 0)
    DG_TC_CASES_SPLIT(3, 1)
    default: depgan_set_error("conv_fwd_tc: no split-half kernel for this (ks, kc, epi)"); return -2;
  }
}

int set_attrs_split() {
  DG_TC_ATTRS_SPLIT(1, 0)
  DG_TC_ATTRS_SPLIT(3, 0)
  DG_TC_ATTRS_SPLIT(3, 1)
  return 0;
}

}  // namespace convtc
