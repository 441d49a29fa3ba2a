// chain_wide_b13.cu -- k_chain_wide for problems with ONE off-diagonal generator block in the diagonal (u) family whose
// entries all carry the same value, and <= 3 diagonal generator blocks: the blockade chains of BASELINE.json configs 4-5.
#include "chain_wide.cuh"
wide_kernel_t qocvl_wide_pick_b13(int spc, int items, int threads) { return wide_pick_blocks<1, 3, true>(spc, items, threads); }
wide_kernel_t qocvl_wide_pick_b13_fwd(int items) { return wide_pick_blocks_fwd<1, 3, true>(items); }
