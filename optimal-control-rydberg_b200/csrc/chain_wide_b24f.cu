// chain_wide_b24f.cu -- the complex64 instantiations of chain_wide_b24.cu (optional single-precision mode)
#include "chain_wide.cuh"
wide_kernel_t qocvl_wide_pick_b24f(int spc, int items, int threads) { return wide_pick_blocks<cplxf, WIDE_MAXB, WIDE_MAXD, false>(spc, items, threads); }
wide_kernel_t qocvl_wide_pick_b24f_fwd(int items) { return wide_pick_blocks_fwd<cplxf, WIDE_MAXB, WIDE_MAXD, false>(items); }
