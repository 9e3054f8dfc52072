// conv_tc_kernel instantiations for 3x3 kernels (see conv_tc_kernel.cuh).
#include "conv_tc_kernel.cuh"

namespace convtc {

// epi = epilogue variant: 0 plain, 1 FiLM residual, 2 add / mask sources, 4 plain + fused 2x2 max-pool; + 8 = IEEE-half
// storage (inference: plain and FiLM only)
int launch_ks3(int grid, uint32_t smem, cudaStream_t st, const TcMaps& tm, const ConvArgs& a, const TcGeom& g, int epi) {
  switch (epi * 100 + (g.kc / 16) * 10 + (g.b_resident ? 1 : 0)) {
    DG_TC_CASES(3, 0)
    DG_TC_CASES(3, 1)
    DG_TC_CASES(3, 2)
    DG_TC_CASES(3, 4)
    DG_TC_CASES_F16(3, 0)
    DG_TC_CASES_F16(3, 1)
    default: depgan_set_error("conv_fwd_tc: no kernel for this (ks, kc, epi)"); return -2;
  }
}

int set_attrs_ks3() {
  DG_TC_ATTRS(3, 0)
  DG_TC_ATTRS(3, 1)
  DG_TC_ATTRS(3, 2)
  DG_TC_ATTRS(3, 4)
  DG_TC_ATTRS_F16(3, 0)
  DG_TC_ATTRS_F16(3, 1)
  return 0;
}

}  // namespace convtc

#ifdef DG_DBG_TRACE
extern "C" int depgan_dbg_set_trace(long long* p) {
  return (int)cudaMemcpyToSymbol(convtc::g_dbg_trace, &p, sizeof(p));
}
#endif
