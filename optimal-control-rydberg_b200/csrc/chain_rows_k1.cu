// chain_rows_k1.cu -- instances of k_chain_rows for one multiplier class per matrix entry (every reference-style ensemble)
// and rows / columns with at most 3 off-diagonal entries (the reference's two models).
#include "chain_rows.cuh"

rows_fn qocvl_rows_kernel_k1(int wm, bool store, bool tma, bool split) {
  if (wm > 3) return qocvl_rows_kernel_k1w(store, tma, split);
  if (!store) return split ? k_chain_rows<3, 1, 0, 0, 1> : k_chain_rows<3, 1, 0, 0, 0>;
  if (tma) return split ? k_chain_rows<3, 1, 1, 1, 1> : k_chain_rows<3, 1, 1, 1, 0>;
  return split ? k_chain_rows<3, 1, 1, 0, 1> : k_chain_rows<3, 1, 1, 0, 0>;
}
