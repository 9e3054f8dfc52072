// conv_tc_kernel instantiations for 5x5 kernels (see conv_tc_kernel.cuh).
#include "conv_tc_kernel.cuh"

namespace convtc {

// EPI variants built for this kernel size; a call is routed to the smallest superset of its side inputs
static int pick_epi(int need) {
  return need == 0 ? 0 : (need == 2 ? 2 : 3);
}

int launch_ks5(int grid, uint32_t smem, cudaStream_t st, const CUtensorMap& tmA0, const CUtensorMap& tmA1,
               const CUtensorMap& tmB, const ConvArgs& a, const TcGeom& g, int need) {
  const int key = pick_epi(need) * 100 + (g.kc / 16) * 10 + (g.b_resident ? 1 : 0);
  switch (key) {
    case 10: return launch_one<5, 1, false, 0>(grid, smem, st, tmA0, tmA1, tmB, a, g);
    case 11: return launch_one<5, 1, true, 0>(grid, smem, st, tmA0, tmA1, tmB, a, g);
    case 20: return launch_one<5, 2, false, 0>(grid, smem, st, tmA0, tmA1, tmB, a, g);
    case 21: return launch_one<5, 2, true, 0>(grid, smem, st, tmA0, tmA1, tmB, a, g);
    case 40: return launch_one<5, 4, false, 0>(grid, smem, st, tmA0, tmA1, tmB, a, g);
    case 41: return launch_one<5, 4, true, 0>(grid, smem, st, tmA0, tmA1, tmB, a, g);
    case 210: return launch_one<5, 1, false, 2>(grid, smem, st, tmA0, tmA1, tmB, a, g);
    case 211: return launch_one<5, 1, true, 2>(grid, smem, st, tmA0, tmA1, tmB, a, g);
    case 220: return launch_one<5, 2, false, 2>(grid, smem, st, tmA0, tmA1, tmB, a, g);
    case 221: return launch_one<5, 2, true, 2>(grid, smem, st, tmA0, tmA1, tmB, a, g);
    case 240: return launch_one<5, 4, false, 2>(grid, smem, st, tmA0, tmA1, tmB, a, g);
    case 241: return launch_one<5, 4, true, 2>(grid, smem, st, tmA0, tmA1, tmB, a, g);
    case 310: return launch_one<5, 1, false, 3>(grid, smem, st, tmA0, tmA1, tmB, a, g);
    case 311: return launch_one<5, 1, true, 3>(grid, smem, st, tmA0, tmA1, tmB, a, g);
    case 320: return launch_one<5, 2, false, 3>(grid, smem, st, tmA0, tmA1, tmB, a, g);
    case 321: return launch_one<5, 2, true, 3>(grid, smem, st, tmA0, tmA1, tmB, a, g);
    case 340: return launch_one<5, 4, false, 3>(grid, smem, st, tmA0, tmA1, tmB, a, g);
    case 341: return launch_one<5, 4, true, 3>(grid, smem, st, tmA0, tmA1, tmB, a, g);
    default: depgan_set_error("conv_fwd_tc: no kernel for this (ks, kc, epi)"); return -2;
  }
}

int set_attrs_ks5() {
  DG_TRY((set_attr_one<5, 1, false, 0>()));
  DG_TRY((set_attr_one<5, 1, true, 0>()));
  DG_TRY((set_attr_one<5, 2, false, 0>()));
  DG_TRY((set_attr_one<5, 2, true, 0>()));
  DG_TRY((set_attr_one<5, 4, false, 0>()));
  DG_TRY((set_attr_one<5, 4, true, 0>()));
  DG_TRY((set_attr_one<5, 1, false, 2>()));
  DG_TRY((set_attr_one<5, 1, true, 2>()));
  DG_TRY((set_attr_one<5, 2, false, 2>()));
  DG_TRY((set_attr_one<5, 2, true, 2>()));
  DG_TRY((set_attr_one<5, 4, false, 2>()));
  DG_TRY((set_attr_one<5, 4, true, 2>()));
  DG_TRY((set_attr_one<5, 1, false, 3>()));
  DG_TRY((set_attr_one<5, 1, true, 3>()));
  DG_TRY((set_attr_one<5, 2, false, 3>()));
  DG_TRY((set_attr_one<5, 2, true, 3>()));
  DG_TRY((set_attr_one<5, 4, false, 3>()));
  DG_TRY((set_attr_one<5, 4, true, 3>()));
  return 0;
}

}  // namespace convtc
