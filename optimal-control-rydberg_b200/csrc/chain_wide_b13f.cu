// chain_wide_b13f.cu -- the complex64 instantiations of chain_wide_b13.cu (optional single-precision mode)
#include "chain_wide.cuh"
wide_kernel_t qocvl_wide_pick_b13f(int spc, int items, int threads) { return wide_pick_blocks<cplxf, 1, 3, true>(spc, items, threads); }
wide_kernel_t qocvl_wide_pick_b13f_fwd(int items) { return wide_pick_blocks_fwd<cplxf, 1, 3, true>(items); }
