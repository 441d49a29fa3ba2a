// chain_wide_b24.cu -- k_chain_wide, general form: <= 2 off-diagonal and <= 4 diagonal generator blocks of any family / values
#include "chain_wide.cuh"
wide_kernel_t qocvl_wide_pick_b24(int spc, int items, int threads, int c64) {
  return c64 ? qocvl_wide_pick_b24f(spc, items, threads) : wide_pick_blocks<cplx, WIDE_MAXB, WIDE_MAXD, false>(spc, items, threads);
}
wide_kernel_t qocvl_wide_pick_b24_fwd(int items, int c64) {
  return c64 ? qocvl_wide_pick_b24f_fwd(items) : wide_pick_blocks_fwd<cplx, WIDE_MAXB, WIDE_MAXD, false>(items);
}
