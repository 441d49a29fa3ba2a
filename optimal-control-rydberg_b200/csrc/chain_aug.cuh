// chain_aug.cuh -- argument block and per-slice entry tables of the lane-group ensemble kernel (chain_aug.cu).
#pragma once
#include "common.cuh"

__device__ __forceinline__ cplx aug_to_d(cplx v) { return v; }
__device__ __forceinline__ cplx aug_to_d(cplxf v) { return cmake((double)v.x, (double)v.y); }

struct AugArgs {
  int N, J, M, C, S;
  int w, wc;                 // off-diagonal row slots / column slots in use
  int ncr, ncc;              // contribution slots per lane: (1+w)*KC, (1+wc)*KC
  int KA, KG, ncls, nparts;
  unsigned add_row_mask, add_col_mask;
  int want_grad;
  double dt, inv_duration, w_infid, w_adiab, inv_fid_norm, inv_samples;
  cplx d_const;              // overlap contribution of the start kets that never move
  const int *row_col, *col_row, *r_cls, *c_cls, *r_gen, *c_gen, *ra_gen, *ca_gen;
  const cplx *r_val, *c_val, *ra_val, *ca_val;
  void *p_row, *p_col;       // per-slice entry tables (written by k_aug_tables, read by k_chain_aug) in the sweep's type
  const cplx* coef;          // [M][J]
  const cplx* dcoef;         // [M][J][C]
  const int* qdeg;           // [M]
  const double* mcls;        // [S][ncls]
  const cplx* add;           // [S][J]
  const cplx* starts;        // [LPG]
  const cplx* targets;       // [LPG] (symmetry / orbit weights folded in)
  const int* pair;           // [LPG] partner row in q^dag p, or -1
  const double* wq;          // [2][LPG]: weight of the row's term in q^dag p (q rows only) / in the cotangent
  void* ckpt;                // [M][total threads]  (complex of the kernel's arithmetic type, like terms / y_partial)
  double* sample_out;        // [S][3]
  void* terms;               // [M][qs][total threads]  Taylor terms t_0 .. t_{q-1} of every slice (STORE = 1)
  int qs;                    // slots per slice in `terms` (the a-priori Taylor cap)
  void* y_partial;           // [warps][M][ncc][LPG]  per-warp sums of m * Xbar
  double* gwave;             // [M][C]
};

// P_i(k) = dt * sum_{g in contribution i} coef_kg * val_ig for every row-owner and column-owner contribution:
// the sample-independent part of the slice generators, computed once per evaluation (M * (ncr+ncc) * L values).
template <typename CT>
__global__ void k_aug_tables(const AugArgs P, int L) {
  const int k = blockIdx.x;
  for (int t = threadIdx.x; t < (P.ncr + P.ncc) * L; t += blockDim.x) {
    const bool rowtab = t < P.ncr * L;
    const int u = rowtab ? t : t - P.ncr * L;
    const int c = u / L, lane = u % L;
    const int* gen = rowtab ? P.r_gen : P.c_gen;
    const cplx* val = rowtab ? P.r_val : P.c_val;
    cplx acc = cmake(0, 0);
    for (int i = 0; i < P.KG; ++i) {
      const int g = gen[((size_t)c * P.KG + i) * L + lane];
      if (g >= 0) acc = cfma(P.coef[(size_t)k * P.J + g], val[((size_t)c * P.KG + i) * L + lane], acc);
    }
    acc = cscale(acc, P.dt);
    CT v;   // stored in the sweep's own type: a conversion in k_chain_aug would wait on the load it sits behind
    v.x = acc.x; v.y = acc.y;
    if (rowtab) ((CT*)P.p_row)[((size_t)k * P.ncr + c) * L + lane] = v;
    else ((CT*)P.p_col)[((size_t)k * P.ncc + c) * L + lane] = v;
  }
}

