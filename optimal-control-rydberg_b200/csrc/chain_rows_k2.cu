// chain_rows_k2.cu -- instances of k_chain_rows for two multiplier classes per matrix entry (general noise tables).
#include "chain_rows.cuh"

rows_fn qocvl_rows_kernel_k2(int wm, bool store, bool tma) {
  if (wm > 3) {
    if (!store) return k_chain_rows<6, 2, 0, 0, 0>;
    return tma ? k_chain_rows<6, 2, 1, 1, 0> : k_chain_rows<6, 2, 1, 0, 0>;
  }
  if (!store) return k_chain_rows<3, 2, 0, 0, 0>;
  return tma ? k_chain_rows<3, 2, 1, 1, 0> : k_chain_rows<3, 2, 1, 0, 0>;
}
