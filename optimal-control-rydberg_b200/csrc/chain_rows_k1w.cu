// chain_rows_k1w.cu -- the same for systems with up to 6 off-diagonal entries per row / column.
#include "chain_rows.cuh"

rows_fn qocvl_rows_kernel_k1w(bool store, bool tma, bool split) {
  if (!store) return split ? k_chain_rows<6, 1, 0, 0, 1> : k_chain_rows<6, 1, 0, 0, 0>;
  if (tma) return split ? k_chain_rows<6, 1, 1, 1, 1> : k_chain_rows<6, 1, 1, 1, 0>;
  return split ? k_chain_rows<6, 1, 1, 0, 1> : k_chain_rows<6, 1, 1, 0, 0>;
}
