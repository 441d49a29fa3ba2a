// chain_wide_b13.cu -- k_chain_wide for problems with ONE off-diagonal generator block in the diagonal (u) family whose
// entries all carry the same value, and <= 3 diagonal generator blocks: the blockade chains of BASELINE.json configs 4-5.
#include "chain_wide.cuh"
wide_kernel_t qocvl_wide_pick_b13(int spc, int items, int threads, int c64) {
  return c64 ? qocvl_wide_pick_b13f(spc, items, threads) : wide_pick_blocks<cplx, 1, 3, true>(spc, items, threads);
}
wide_kernel_t qocvl_wide_pick_b13_fwd(int items, int c64) {
  return c64 ? qocvl_wide_pick_b13f_fwd(items) : wide_pick_blocks_fwd<cplx, 1, 3, true>(items);
}
