// chain_small.cu -- the hot kernel (n <= 16): time-ordered propagation of the columns the cost reads, through
// every slice of every noise sample, and the exact reverse sweep that yields d cost / d wave.
//
// Replaces (numerically, not structurally) ST:247-302 / TQ:460-567 plus the reverse pass
// jax.value_and_grad builds for them (TQ:592).  Formulation (oracle/vector_model.py is the CPU
// twin): the Van Loan slice generator is X_k = [[A_k, B_k],[0, A_k]] with
//   A_k = dt * sum_j c_kj G^u_j,   B_k = dt * sum_j c_kj G^w_j,   c_kj = coef_kj*mul_sj + add_sj.
// exp(X_k) is applied to vectors by a truncated Taylor series (degree q_k chosen from |X_k|_1
// so the remainder is < 2^-55): t_0 = z, t_j = X t_{j-1}/j, z' = sum t_j.  The columns needed are
//   z_v = U x_v (one per start ket that moves) and p = W psi (coupled to q = U psi through B_k).
// Reverse sweep of the same recurrence:  tb_q = lam', tb_{j-1} = lam' + X^dag tb_j / j,
//   Xbar += conj(tb_j) t_{j-1}^T / j, contracted on the fly with the generators' sparsity.
//
// Mapping (v3): ONE NOISE SAMPLE PER GROUP OF LPG LANES.  Lane r owns row r of A_k (diagonal +
// WMAX off-diagonals, in registers) and element r of ALL NV+1 vectors of its sample, so the
// NV+1 Taylor chains of a lane are independent instruction streams (ILP) that share the row
// values, and the q -> p coupling through B_k is local to the group.  Gathers of the other rows'
// elements go through a double-buffered shared-memory exchange guarded by __syncwarp only: there
// is no CTA-wide barrier inside the sweeps.  Taylor terms of the reverse sweep are parked in
// shared memory [j][v][lane]; state checkpoints z_k go to HBM, coalesced, once per slice.
#include "common.cuh"
#include "model.cuh"
#include <algorithm>

__constant__ double c_inv[QOCVL_QMAX + 2];   // 1/j

#ifndef QOCVL_MINB
#define QOCVL_MINB 1   // min CTAs of 512 threads per SM the register allocation must allow (tuning knob)
#endif

struct ChainArgs {
  int n, J, M, C, S, nv, va;   // nv = moving start kets, va = which of them is psi
  int w_u, wc_u, kc_u;         // off-diagonal row slots / col slots / contributions (u family)
  int w_w, wc_w, kc_w;         // (w family, no diagonal slot)
  int qcap, want_grad, nwarps_total;
  double dt, inv_duration, w_infid, w_adiab, inv_fid_norm, inv_samples;
  cplx d_const;                // overlap contribution of the start kets that never move
  const int *u_row_col, *u_col_row, *u_rc_gen, *u_cc_gen;
  const cplx *u_rc_val, *u_cc_val;
  const int *w_row_col, *w_col_row, *w_rc_gen, *w_cc_gen;
  const cplx *w_rc_val, *w_cc_val;
  const cplx* coef;    // [M][J]
  const cplx* dcoef;   // [M][J][C]
  const int* qdeg;     // [M]
  const double* mul;   // [S][J]
  const cplx* add;     // [S][J]
  const cplx* starts;  // [nv][LPG]
  const cplx* targets; // [nv][LPG] (symmetry weight folded in)
  cplx* ckpt;          // [M][nv+1][total threads]   (TP=2: [chunk_len][nv+1][total threads])
  double* sample_out;  // [S][3]
  double* gw_partial;  // [total warps][M][C]
  // ---- time-parallel single-sample mode (TP != 0): the M slices are cut into chunks that run concurrently
  int chunk_len, n_chunks, ncol_pad;
  cplx* tp_uc;         // [n_chunks][n][LPG]  column i of the chunk propagator block U_c
  cplx* tp_wc;         // [n_chunks][n][LPG]  column i of W_c
  const cplx* tp_zc;   // [n_chunks+1][nv+1][LPG]  states at the chunk boundaries (k_tp_combine)
  const cplx* tp_lc;   // [n_chunks+1][nv+1][LPG]  cotangents at the chunk boundaries
  double* tp_gwave;    // [M][C]  written directly (every slice belongs to exactly one chunk)
};

template <int LPG>
__device__ __forceinline__ cplx group_sum(cplx v) {
#pragma unroll
  for (int off = LPG / 2; off > 0; off >>= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, off);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, off);
  }
  return v;
}

// TLOCAL = 1 parks the Taylor terms of the reverse sweep in thread-local memory (L1-cached, write-back,
// lane-interleaved by the hardware) instead of shared memory: occupancy is then bounded by registers only.
// TP = 0: one noise sample per lane group, all M slices (the ensemble hot path).
// TP = 1: single sample, work item = (chunk, basis column): forward sweep over the chunk's slices only, result =
//         one column of the chunk's Van Loan propagator blocks (U_c e_i, W_c e_i).
// TP = 2: single sample, work item = chunk: forward + reverse sweep over the chunk's slices between the boundary
//         states / cotangents that k_tp_combine stitched together from the chunk propagators.
template <int LPG, int NV, int WMAX, int WBMAX, int TLOCAL, int TP>
__global__ void __launch_bounds__(512, QOCVL_MINB) k_chain(const ChainArgs P) {
  constexpr int NVP = NV + 1;                      // the NV columns of U plus p = W psi (index NV)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int grp = tid / LPG, r = tid % LPG, ngroups = nthreads / LPG;
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
  const int J = P.J, C = P.C, JC = P.J * P.C, va = P.va;
  const size_t gtid = (size_t)blockIdx.x * nthreads + tid, gthreads = (size_t)gridDim.x * nthreads;
  const int wid = blockIdx.x * ngroups + grp;        // work item of my lane group
  const int sample = TP ? 0 : wid;
  int chunk = 0, colid = 0;
  if (TP == 1) { chunk = wid / P.ncol_pad; colid = wid % P.ncol_pad; }
  if (TP == 2) chunk = wid;
  const bool item_ok = TP == 0 ? (wid < P.S) : (TP == 1 ? (chunk < P.n_chunks && colid < P.n) : (chunk < P.n_chunks));
  const bool live = item_ok && r < P.n;
  const int sm = (TP == 0 && wid < P.S) ? wid : (TP ? 0 : P.S - 1);
  // slice range of my work item; every group of the launch runs the same number of local slices (nloc)
  const int k0 = TP ? (chunk < P.n_chunks ? chunk : P.n_chunks - 1) * P.chunk_len : 0;
  const int k1 = TP ? (k0 + P.chunk_len < P.M ? k0 + P.chunk_len : P.M) : P.M;
  const int nloc = TP ? P.chunk_len : P.M;

  // ---- shared memory carve-up -------------------------------------------------------
  const int n_urc = (P.w_u + 1) * P.kc_u * LPG, n_ucc = (P.wc_u + 1) * P.kc_u * LPG;
  const int n_wrc = P.w_w * P.kc_w * LPG, n_wcc = P.wc_w * P.kc_w * LPG;
  cplx* carve = (cplx*)smem_raw;
  cplx* s_urcv = carve; carve += n_urc;                        // [w_u+1][kc_u][LPG]  row-owner contributions
  cplx* s_uccv = carve; carve += n_ucc;                        // [wc_u+1][kc_u][LPG] col-owner contributions
  cplx* s_wrcv = carve; carve += n_wrc;
  cplx* s_wccv = carve; carve += n_wcc;
  cplx* s_cs = carve + (size_t)grp * J; carve += (size_t)ngroups * J;          // [ngroups][J] slice coefficients
  // d coef / d wave of the current slice: one table per warp (TP = 0: its groups share the slice) or per group
  cplx* s_dco = carve + (size_t)(TP ? grp : warp) * JC; carve += (size_t)ngroups * JC;   // [ngroups][J][C]
  cplx* s_x = carve + (size_t)grp * 2 * NVP * LPG; carve += (size_t)ngroups * 2 * NVP * LPG;   // [ngroups][2][NVP][LPG]
  cplx* s_T = carve + (TLOCAL ? 0 : (size_t)grp * P.qcap * NVP * LPG);      // [ngroups][qcap][NVP][LPG] (TLOCAL=0)
  carve += TLOCAL ? 0 : (size_t)ngroups * P.qcap * NVP * LPG;
  cplx t_loc[TLOCAL ? QOCVL_QMAX * NVP : 1];                                // [QMAX][NVP] (TLOCAL=1)
  double* s_mul = (double*)carve + (size_t)grp * J;                            // [ngroups][J]
  int* s_urcg = (int*)((double*)carve + (size_t)ngroups * J);
  int* s_uccg = s_urcg + n_urc;
  int* s_wrcg = s_uccg + n_ucc;
  int* s_wccg = s_wrcg + n_wrc;
  for (int i = tid; i < n_urc; i += nthreads) { s_urcv[i] = P.u_rc_val[i]; s_urcg[i] = P.u_rc_gen[i]; }
  for (int i = tid; i < n_ucc; i += nthreads) { s_uccv[i] = P.u_cc_val[i]; s_uccg[i] = P.u_cc_gen[i]; }
  for (int i = tid; i < n_wrc; i += nthreads) { s_wrcv[i] = P.w_rc_val[i]; s_wrcg[i] = P.w_rc_gen[i]; }
  for (int i = tid; i < n_wcc; i += nthreads) { s_wccv[i] = P.w_cc_val[i]; s_wccg[i] = P.w_cc_gen[i]; }
  for (int j = r; j < J; j += LPG) s_mul[j] = P.mul[(size_t)sm * J + j];
  __syncthreads();                                             // the only CTA-wide barrier

  // static pattern indices of my row / column (empty slots point at my own row and carry a zero value)
  int rcol[WMAX], crow[WMAX];
#pragma unroll
  for (int s = 0; s < WMAX; ++s) {
    rcol[s] = (s < P.w_u) ? P.u_row_col[(1 + s) * LPG + r] : r;
    crow[s] = (s < P.wc_u) ? P.u_col_row[(1 + s) * LPG + r] : r;
  }
  int brcol[WBMAX], bcrow[WBMAX];
#pragma unroll
  for (int s = 0; s < WBMAX; ++s) {
    brcol[s] = (s < P.w_w) ? P.w_row_col[s * LPG + r] : r;
    bcrow[s] = (s < P.wc_w) ? P.w_col_row[s * LPG + r] : r;
  }
  // how many of my slots are real entries (they form a prefix): gathers of empty slots are predicated
  // off, which saves shared-memory wavefronts -- the sweeps are bound by the shared-memory pipe
  int rcnt = 0, ccnt = 0, brcnt = 0, bccnt = 0;
  for (int s = 0; s < P.w_u; ++s) rcnt += s_urcg[((1 + s) * P.kc_u) * LPG + r] >= 0;
  for (int s = 0; s < P.wc_u; ++s) ccnt += s_uccg[((1 + s) * P.kc_u) * LPG + r] >= 0;
  for (int s = 0; s < P.w_w; ++s) brcnt += s_wrcg[(s * P.kc_w) * LPG + r] >= 0;
  for (int s = 0; s < P.wc_w; ++s) bccnt += s_wccg[(s * P.kc_w) * LPG + r] >= 0;
  // my share of the per-sample noise table: lane r prepares coefficient j = r (+LPG)
  double mymul[(QOCVL_MAX_GEN + LPG - 1) / LPG];
  cplx myadd[(QOCVL_MAX_GEN + LPG - 1) / LPG];
#pragma unroll
  for (int i = 0; i < (QOCVL_MAX_GEN + LPG - 1) / LPG; ++i) {
    const int j = r + i * LPG;
    mymul[i] = (j < J) ? P.mul[(size_t)sm * J + j] : 0.0;
    myadd[i] = (j < J) ? P.add[(size_t)sm * J + j] : cmake(0, 0);
  }

  const cplx czero = cmake(0.0, 0.0);
  cplx z[NVP], ytgt[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    z[v] = live ? P.starts[v * LPG + r] : czero;
    ytgt[v] = live ? P.targets[v * LPG + r] : czero;
  }
  z[NV] = czero;
  if (TP == 1) z[0] = (live && r == colid) ? cmake(1.0, 0.0) : czero;            // basis column e_colid
  if (TP == 2) {
#pragma unroll
    for (int v = 0; v < NVP; ++v) z[v] = live ? P.tp_zc[((size_t)chunk * NVP + v) * LPG + r] : czero;
  }
  int ph = 0;                                                   // exchange phase offset: 0 or NVP*LPG

  cplx a_d, a_row[WMAX], b_row[WBMAX];

  // exchange layout per vector: re[LPG] then im[LPG] (8-byte accesses: the 16 lanes of a sample hit 16
  // distinct bank pairs whatever the gather pattern -> no shared-memory bank conflicts)
  auto xput = [&](double* xs, int v, cplx val) {
    xs[v * 2 * LPG + r] = val.x;
    xs[v * 2 * LPG + LPG + r] = val.y;
  };
  auto xget = [&](const double* xs, int v, int idx) -> cplx {
    return cmake(xs[v * 2 * LPG + idx], xs[v * 2 * LPG + LPG + idx]);
  };

  // cs[j] = coef_kj * mul_sj + add_sj for slice kk, written to the group's shared slot
  auto load_coef = [&](int kk, cplx* creg) {
#pragma unroll
    for (int i = 0; i < (QOCVL_MAX_GEN + LPG - 1) / LPG; ++i) {
      const int j = r + i * LPG;
      creg[i] = (j < J) ? P.coef[(size_t)kk * J + j] : czero;
    }
  };
  auto store_cs = [&](const cplx* creg) {
#pragma unroll
    for (int i = 0; i < (QOCVL_MAX_GEN + LPG - 1) / LPG; ++i) {
      const int j = r + i * LPG;
      if (j < J) s_cs[j] = cadd(cscale(creg[i], mymul[i]), myadd[i]);
    }
  };
  auto slot_value = [&](const int* gen, const cplx* val, int kcn, int slot) -> cplx {
    cplx acc = czero;
    for (int kc = 0; kc < kcn; ++kc) {
      const int gi = gen[(slot * kcn + kc) * LPG + r];
      if (gi >= 0) acc = cfma(s_cs[gi], val[(slot * kcn + kc) * LPG + r], acc);
    }
    return cscale(acc, P.dt);
  };
  auto assemble_rows = [&]() {
    a_d = slot_value(s_urcg, s_urcv, P.kc_u, 0);
#pragma unroll
    for (int s = 0; s < WMAX; ++s) a_row[s] = (s < P.w_u) ? slot_value(s_urcg, s_urcv, P.kc_u, 1 + s) : czero;
#pragma unroll
    for (int s = 0; s < WBMAX; ++s) b_row[s] = (s < P.w_w) ? slot_value(s_wrcg, s_wrcv, P.kc_w, s) : czero;
  };
  // one forward Taylor step on all vectors of the sample: t <- X t / j
  auto step_fwd = [&](cplx* t, double invj) {
    double* xs = (double*)(s_x + ph);
#pragma unroll
    for (int v = 0; v < NVP; ++v) xput(xs, v, t[v]);
    __syncwarp();
    cplx tq_b[WBMAX];
#pragma unroll
    for (int s = 0; s < WBMAX; ++s) tq_b[s] = (s < brcnt) ? xget(xs, va, brcol[s]) : czero;
#pragma unroll
    for (int v = 0; v < NVP; ++v) {
      // one accumulator per vector: with >= 3 resident warps per scheduler the dependent chain is hidden,
      // and the FP64 pipe (not latency) is what the sweep is bound by -- fewer instructions win
      // (the single-sample time-parallel sweeps, TP != 0, run ONE warp per SM sub-core: there the dependent FP64
      // chain is the step time, so they split it over two accumulators)
      cplx acc0 = cmul(a_d, t[v]), acc1 = czero;
#pragma unroll
      for (int s = 0; s < WMAX; ++s) {
        const cplx xv = (s < rcnt) ? xget(xs, v, rcol[s]) : czero;
        if (TP && (s & 1)) acc1 = cfma(a_row[s], xv, acc1);
        else acc0 = cfma(a_row[s], xv, acc0);
      }
      if (v == NV) {
#pragma unroll
        for (int s = 0; s < WBMAX; ++s) {
          if (TP) acc1 = cfma(b_row[s], tq_b[s], acc1);
          else acc0 = cfma(b_row[s], tq_b[s], acc0);
        }
      }
      if (TP) acc0 = cadd(acc0, acc1);
      t[v] = cscale(acc0, invj);
    }
    ph = NVP * LPG - ph;
  };

  // ---- prologue: coefficients of slice 0
  {
    cplx creg[(QOCVL_MAX_GEN + LPG - 1) / LPG];
    load_coef(k0, creg);
    store_cs(creg);
    __syncwarp();
  }
  // Taylor degree of local slice kk: in the time-parallel modes the groups of a warp sit on different slices,
  // so the warp runs the largest degree any of them needs (extra terms are below the truncation tolerance)
  auto slice_degree = [&](int k, bool act) -> int {
    int q = act ? P.qdeg[k] : 0;
    if (TP) {
#pragma unroll
      for (int off = LPG; off < 32; off <<= 1) q = max(q, __shfl_xor_sync(0xffffffffu, q, off));
    }
    return q;
  };
  auto clear_rows = [&]() {                          // a local slice beyond the end of the last chunk: X = 0
    a_d = czero;
#pragma unroll
    for (int s = 0; s < WMAX; ++s) a_row[s] = czero;
#pragma unroll
    for (int s = 0; s < WBMAX; ++s) b_row[s] = czero;
  };

  // =================================== forward sweep ===================================
  for (int kk = 0; kk < nloc; ++kk) {
    const int k = (k0 + kk < P.M) ? k0 + kk : P.M - 1;
    const bool act = (TP == 0) || (k0 + kk < k1);
    const int q = slice_degree(k, act);
    assemble_rows();
    if (TP && !act) clear_rows();
    __syncwarp();                                   // everyone has read cs of slice k
    cplx creg[(QOCVL_MAX_GEN + LPG - 1) / LPG];
    const bool more = kk + 1 < nloc;
    if (more) load_coef((k0 + kk + 1 < P.M) ? k0 + kk + 1 : P.M - 1, creg);   // in flight during this slice
    if (P.want_grad && TP != 1) {
#pragma unroll
      for (int v = 0; v < NVP; ++v) P.ckpt[((size_t)(TP ? kk : k) * NVP + v) * gthreads + gtid] = z[v];
    }
    cplx t[NVP];
#pragma unroll
    for (int v = 0; v < NVP; ++v) t[v] = z[v];
    for (int j = 1; j <= q; ++j) {
      step_fwd(t, c_inv[j]);
#pragma unroll
      for (int v = 0; v < NVP; ++v) z[v] = cadd(z[v], t[v]);
    }
    if (more) store_cs(creg);
    __syncwarp();
  }
  if (TP == 1) {                                    // one column of the chunk propagator blocks
    if (live) {
      P.tp_uc[((size_t)chunk * P.n + colid) * LPG + r] = z[0];
      P.tp_wc[((size_t)chunk * P.n + colid) * LPG + r] = z[NV];
    }
    return;
  }

  // =================================== cost ============================================
  cplx dsum = czero, qp = czero;
  if (TP == 0) {
    cplx part = czero;
#pragma unroll
    for (int v = 0; v < NV; ++v) part = cfma_conj(ytgt[v], z[v], part);     // sum_v conj(y_v[r]) z_v[r]
    cplx zq = czero;
#pragma unroll
    for (int v = 0; v < NV; ++v) if (v == va) zq = z[v];
    dsum = cadd(P.d_const, group_sum<LPG>(part));
    qp = group_sum<LPG>(cfma_conj(zq, z[NV], czero));                        // q^dag p
    if (r == 0 && sample < P.S) {
      const double infid = 1.0 - (dsum.x * dsum.x + dsum.y * dsum.y) * P.inv_fid_norm;
      const double adiab = 1.0 - qp.x * P.inv_duration;
      P.sample_out[(size_t)sample * 3 + 0] = P.w_infid * infid + P.w_adiab * adiab;
      P.sample_out[(size_t)sample * 3 + 1] = infid;
      P.sample_out[(size_t)sample * 3 + 2] = adiab;
    }
  }
  if (!P.want_grad) return;

  // terminal cotangents lam = d cost / d conj(z), scaled by 1/S_global (ensemble mean)
  cplx lam[NVP];
  {
    const double ws = live ? P.inv_samples : 0.0;
    const double ca = -0.5 * P.w_adiab * P.inv_duration * ws;
    cplx zq = czero;
#pragma unroll
    for (int v = 0; v < NV; ++v) if (v == va) zq = z[v];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      lam[v] = cscale(cmul(dsum, ytgt[v]), -P.w_infid * P.inv_fid_norm * ws);
      if (v == va) lam[v] = cadd(lam[v], cscale(z[NV], ca));                 // -(wa/2T) p
    }
    lam[NV] = cscale(zq, ca);                                                // -(wa/2T) q
  }
  if (TP == 2) {                                    // cotangents at the end of my chunk (k_tp_combine)
#pragma unroll
    for (int v = 0; v < NVP; ++v) lam[v] = live ? P.tp_lc[((size_t)(chunk + 1) * NVP + v) * LPG + r] : czero;
  }

  // ---- prologue of the reverse sweep
  __syncwarp();
  {
    cplx creg[(QOCVL_MAX_GEN + LPG - 1) / LPG];
    load_coef((k0 + nloc - 1 < P.M) ? k0 + nloc - 1 : P.M - 1, creg);
    store_cs(creg);
    __syncwarp();
  }

  // =================================== reverse sweep ===================================
  const size_t gwarp = (size_t)blockIdx.x * nwarps + warp;
  for (int kk = nloc - 1; kk >= 0; --kk) {
    const int k = (k0 + kk < P.M) ? k0 + kk : P.M - 1;
    const bool act = (TP == 0) || (k0 + kk < k1);
    const int q = slice_degree(k, act);
    assemble_rows();
    if (TP && !act) clear_rows();
    // column-owner values: A[a][r] for the rows a of my column, B[a][r] for the q <- p coupling
    cplx a_col[WMAX], b_col[WBMAX];
#pragma unroll
    for (int s = 0; s < WMAX; ++s) a_col[s] = (s < P.wc_u) ? slot_value(s_uccg, s_uccv, P.kc_u, 1 + s) : czero;
#pragma unroll
    for (int s = 0; s < WBMAX; ++s) b_col[s] = (s < P.wc_w) ? slot_value(s_wccg, s_wccv, P.kc_w, s) : czero;
    if (TP && !act) {                               // beyond the end of the last chunk: X^dag = 0 as well
#pragma unroll
      for (int s = 0; s < WMAX; ++s) a_col[s] = czero;
#pragma unroll
      for (int s = 0; s < WBMAX; ++s) b_col[s] = czero;
    }
    __syncwarp();                                   // everyone has read cs of slice k
    cplx creg[(QOCVL_MAX_GEN + LPG - 1) / LPG];
    if (kk > 0) load_coef((k0 + kk - 1 < P.M) ? k0 + kk - 1 : P.M - 1, creg);
    // d coef / d wave of this slice: needed only by the contraction at the end of the slice
    cplx dreg[3];
    if (TP == 0) {
#pragma unroll
      for (int i = 0; i < 3; ++i) dreg[i] = (lane + 32 * i < JC) ? P.dcoef[(size_t)k * JC + lane + 32 * i] : czero;
    }
    // ---- recompute forward Taylor terms t_0 .. t_{q-1}
    cplx t[NVP];
#pragma unroll
    for (int v = 0; v < NVP; ++v) {
      t[v] = P.ckpt[((size_t)(TP ? kk : k) * NVP + v) * gthreads + gtid];
      if (TLOCAL) t_loc[v] = t[v]; else s_T[v * LPG + r] = t[v];
    }
    for (int j = 1; j < q; ++j) {
      step_fwd(t, c_inv[j]);
#pragma unroll
      for (int v = 0; v < NVP; ++v) {
        if (TLOCAL) t_loc[j * NVP + v] = t[v]; else s_T[((size_t)j * NVP + v) * LPG + r] = t[v];
      }
    }
    // ---- descending adjoint recurrence with on-the-fly contraction
    cplx tb[NVP];
#pragma unroll
    for (int v = 0; v < NVP; ++v) tb[v] = lam[v];
    cplx xc_d = czero, xc[WMAX], xcb[WBMAX];
#pragma unroll
    for (int s = 0; s < WMAX; ++s) xc[s] = czero;
#pragma unroll
    for (int s = 0; s < WBMAX; ++s) xcb[s] = czero;
    for (int j = q; j >= 1; --j) {
      const double invj = c_inv[j];
      double* xs = (double*)(s_x + ph);
#pragma unroll
      for (int v = 0; v < NVP; ++v) xput(xs, v, tb[v]);
      __syncwarp();
#pragma unroll
      for (int v = 0; v < NVP; ++v) {
        const cplx tprev = cscale(TLOCAL ? t_loc[(j - 1) * NVP + v] : s_T[((size_t)(j - 1) * NVP + v) * LPG + r], invj);
        cplx acc0 = cfma_conj(a_d, tb[v], czero), acc1 = czero;
        xc_d = cfma_conj(tb[v], tprev, xc_d);
#pragma unroll
        for (int s = 0; s < WMAX; ++s) {
          const cplx ta = (s < ccnt) ? xget(xs, v, crow[s]) : czero;
          if (TP && (s & 1)) acc1 = cfma_conj(a_col[s], ta, acc1);
          else acc0 = cfma_conj(a_col[s], ta, acc0);
          xc[s] = cfma_conj(ta, tprev, xc[s]);
        }
        if (v == va) {                              // q receives B^dag tb_p; Bbar pairs tb_p with t_q
#pragma unroll
          for (int s = 0; s < WBMAX; ++s) {
            const cplx ta = (s < bccnt) ? xget(xs, NV, bcrow[s]) : czero;
            if (TP) acc1 = cfma_conj(b_col[s], ta, acc1);
            else acc0 = cfma_conj(b_col[s], ta, acc0);
            xcb[s] = cfma_conj(ta, tprev, xcb[s]);
          }
        }
        if (TP) acc0 = cadd(acc0, acc1);
        tb[v] = cadd(lam[v], cscale(acc0, invj));
      }
      ph = NVP * LPG - ph;
    }
#pragma unroll
    for (int v = 0; v < NVP; ++v) lam[v] = tb[v];
    // ---- contract Xbar with the generators: d cost / d wave[k][c]
    if (TP == 0) {
#pragma unroll
      for (int i = 0; i < 3; ++i) if (lane + 32 * i < JC) s_dco[lane + 32 * i] = dreg[i];
    } else {                                        // groups of a warp sit on different slices
      for (int i = r; i < JC; i += LPG) s_dco[i] = P.dcoef[(size_t)k * JC + i];
    }
    __syncwarp();
    double gw[QOCVL_MAX_CHAN];
#pragma unroll
    for (int c = 0; c < QOCVL_MAX_CHAN; ++c) gw[c] = 0.0;
    if (live && act) {
      const double two_dt = 2.0 * P.dt;
      for (int slot = 0; slot <= P.wc_u; ++slot) {
        cplx xval = xc_d;
#pragma unroll
        for (int s = 0; s < WMAX; ++s)
          if (slot == s + 1) xval = xc[s];
        for (int kc = 0; kc < P.kc_u; ++kc) {
          const int gi = s_uccg[(slot * P.kc_u + kc) * LPG + r];
          if (gi >= 0) {
            const cplx tmp = cscale(cmul(s_uccv[(slot * P.kc_u + kc) * LPG + r], xval), two_dt * s_mul[gi]);
#pragma unroll
            for (int c = 0; c < QOCVL_MAX_CHAN; ++c)
              if (c < C) gw[c] += tmp.x * s_dco[gi * C + c].x - tmp.y * s_dco[gi * C + c].y;
          }
        }
      }
      for (int slot = 0; slot < P.wc_w; ++slot) {
        cplx xval = czero;
#pragma unroll
        for (int s = 0; s < WBMAX; ++s)
          if (slot == s) xval = xcb[s];
        for (int kc = 0; kc < P.kc_w; ++kc) {
          const int gi = s_wccg[(slot * P.kc_w + kc) * LPG + r];
          if (gi >= 0) {
            const cplx tmp = cscale(cmul(s_wccv[(slot * P.kc_w + kc) * LPG + r], xval), two_dt * s_mul[gi]);
#pragma unroll
            for (int c = 0; c < QOCVL_MAX_CHAN; ++c)
              if (c < C) gw[c] += tmp.x * s_dco[gi * C + c].x - tmp.y * s_dco[gi * C + c].y;
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < QOCVL_MAX_CHAN; ++c) {
      double vsum = gw[c];
      if (TP == 0) {
        for (int off = 16; off > 0; off >>= 1) vsum += __shfl_xor_sync(0xffffffffu, vsum, off);
        if (lane == 0 && c < C) P.gw_partial[(gwarp * P.M + k) * C + c] = vsum;
      } else {                                      // my chunk owns slice k: reduce over my group only
#pragma unroll
        for (int off = LPG / 2; off > 0; off >>= 1) vsum += __shfl_xor_sync(0xffffffffu, vsum, off);
        if (r == 0 && c < C && act && item_ok) P.tp_gwave[(size_t)k * C + c] = vsum;
      }
    }
    if (kk > 0) store_cs(creg);
    __syncwarp();
  }
}

static size_t chain_smem_bytes(const qocvl_plan* p, int ngroups, int qcap) {   // qcap = 0: Taylor terms not in smem
  const int J = p->d.n_gen, C = p->d.n_chan, L = p->npad, nvp = p->n_live + 1;
  const int nwarps = ngroups * L / 32;
  const size_t w_w = p->lw.nnz ? p->lw.n_row_slots : 0, wc_w = p->lw.nnz ? p->lw.n_col_slots : 0;
  const size_t tab = ((size_t)(p->lu.n_row_slots + p->lu.n_col_slots) * p->lu.kc + (w_w + wc_w) * p->lw.kc) * L;
  const size_t cplx_count = tab + (size_t)ngroups * J + (size_t)ngroups * J * C + (size_t)ngroups * 2 * nvp * L +
                            (size_t)ngroups * qcap * nvp * L;
  return cplx_count * sizeof(cplx) + (size_t)ngroups * J * sizeof(double) + tab * sizeof(int);
}

typedef void (*chain_fn)(const ChainArgs);
template <int LPG, int TL, int TP>
static chain_fn pick_nv(int nv) {
  switch (nv) {
    case 1: return k_chain<LPG, 1, 4, 2, TL, TP>;
    case 2: return k_chain<LPG, 2, 4, 2, TL, TP>;
    case 3: return LPG == 16 ? k_chain<16, 3, 4, 2, TL, TP> : nullptr;
    case 4: return LPG == 16 ? k_chain<16, 4, 4, 2, TL, TP> : nullptr;
    default: return nullptr;
  }
}
static chain_fn pick_kernel(const qocvl_plan* p) {
  if (p->tlocal) return p->npad == 8 ? pick_nv<8, 1, 0>(p->n_live) : pick_nv<16, 1, 0>(p->n_live);
  return p->npad == 8 ? pick_nv<8, 0, 0>(p->n_live) : pick_nv<16, 0, 0>(p->n_live);
}
static chain_fn pick_tp_sweep(const qocvl_plan* p) {
  return p->npad == 8 ? pick_nv<8, 1, 2>(p->n_live) : pick_nv<16, 1, 2>(p->n_live);
}
static chain_fn pick_tp_prop(const qocvl_plan* p) {
  return p->npad == 8 ? k_chain<8, 1, 4, 2, 1, 1> : k_chain<16, 1, 4, 2, 1, 1>;
}

// Launch geometry and per-sample buffers of the lane-per-row kernel (called by qocvl_chain_configure
// once the live start kets and the a-priori Taylor cap are known).
int qocvl_small_configure(qocvl_plan* p) {
  const qocvl_desc& d = p->d;
  const int L = p->npad, S = p->n_samples;
  if (L == 0 || p->lu.n_row_slots - 1 > 4 || p->lu.n_col_slots - 1 > 4 || p->lw.n_row_slots > 2 ||
      p->lw.n_col_slots > 2 || d.n_gen * d.n_chan > 96 || p->n_live > 4)
    return QOCVL_ERR_UNSUPPORTED;
  {
    double inv[QOCVL_QMAX + 2];
    inv[0] = 0.0;
    for (int j = 1; j < QOCVL_QMAX + 2; ++j) inv[j] = 1.0 / (double)j;
    QCUDA(cudaMemcpyToSymbol(c_inv, inv, sizeof(inv)));
  }
  // a CTA is only a container of independent groups (no barrier inside the sweeps)
  const int gpw = 32 / L;
  p->tlocal = true;                                        // Taylor terms in thread-local memory (L1)
  if (p->has_opt("tlocal")) p->tlocal = p->opt("tlocal") != 0;
  // Samples per CTA.  The kernel is compiled for <= 128 registers (512 threads / SM); beyond ~8 resident
  // warps an SM's throughput is flat, so the makespan is the number of samples the busiest SM has to process.
  // Pick the CTA size that minimises it (wave quantisation was worth 20 % on cfg 3: 4096 samples / 148 SMs ->
  // 28 samples per CTA = 147 CTAs = one even wave, instead of 256 CTAs of 16 = 1.73 waves).
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device);
  const int gmax = 512 / L / gpw * gpw;
  int ngroups = gpw;
  {
    long best_load = -1;
    for (int g = gpw; g <= gmax; g += gpw) {
      const long n_ctas = (S + g - 1) / g;
      const long load = (n_ctas + sms - 1) / sms * g;
      if (best_load < 0 || load < best_load || (load == best_load && g <= 32 && g > ngroups)) {
        best_load = load; ngroups = g;
      }
    }
  }
  if (!p->tlocal) ngroups = std::min(ngroups, 2 * gpw);
  if (p->has_opt("spb")) ngroups = std::max(gpw, (p->opt("spb") + gpw - 1) / gpw * gpw);
  ngroups = std::min(ngroups, (S + gpw - 1) / gpw * gpw);  // small ensembles: do not launch idle warps
  const int q_smem = p->tlocal ? 0 : p->qcap;
  while (ngroups > gpw && chain_smem_bytes(p, ngroups, q_smem) > 100 * 1024) ngroups -= gpw;
  p->spb = ngroups; p->groups = ngroups; p->threads = ngroups * L;
  p->chain_smem = chain_smem_bytes(p, ngroups, q_smem);
  p->qcap_apriori = p->qcap;
  if (p->tlocal) p->qcap = QOCVL_QMAX;                    // no shared-memory limit on the degree any more
  if (p->chain_smem > 227 * 1024) return QOCVL_ERR_UNSUPPORTED;
  p->blocks = (S + ngroups - 1) / ngroups;
  chain_fn fn = pick_kernel(p);
  if (!fn) return QOCVL_ERR_UNSUPPORTED;
  QCUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->chain_smem));
  const size_t total_threads = (size_t)p->blocks * p->threads;
  p->n_parts = total_threads / 32;                         // one partial gradient per warp
  const size_t ck = (size_t)d.n_slices * (p->n_live + 1) * total_threads;
  if (p->d_ckpt) { cudaFree(p->d_ckpt); p->d_ckpt = nullptr; p->bytes -= p->ckpt_bytes; }
  QCUDA(cudaMalloc((void**)&p->d_ckpt, ck * sizeof(cplx)));
  p->ckpt_bytes = ck * sizeof(cplx);
  p->bytes += p->ckpt_bytes;
  return QOCVL_OK;
}

static void fill_common_args(qocvl_plan* p, ChainArgs& a, bool want_grad);

int qocvl_small_launch(qocvl_plan* p, bool want_grad, cudaStream_t st) {
  ChainArgs a;
  fill_common_args(p, a, want_grad);
  pick_kernel(p)<<<p->blocks, p->threads, p->chain_smem, st>>>(a);
  p->launches++;
  QCUDA(cudaGetLastError());
  return QOCVL_OK;
}

// ===================================================================================================
// Time-parallel single-sample path (cfg 1 / cfg 2: the reference's own notebook workloads).
// A single sample exposes no ensemble parallelism and the slices are sequentially dependent, so the
// plain sweep would run on one warp.  Instead:
//   A. k_chain<.., TP=1>: every (chunk, basis column) pair propagates one column through its chunk:
//      the chunk's Van Loan propagator blocks U_c, W_c (n_chunks * n independent work items);
//   B. k_tp_combine: one small CTA walks the chunks once forward (boundary states, cost, terminal
//      cotangents) and once backward (boundary cotangents) with dense n x n mat-vecs;
//   C. k_chain<.., TP=2>: every chunk runs the usual forward + reverse sweep between its boundaries.
// Sequential depth drops from M slices to ~2 chunk_len slices + n_chunks small mat-vecs.
// ===================================================================================================
struct CombineArgs {
  int n, nv, va, n_chunks, lpg;
  double inv_duration, w_infid, w_adiab, inv_fid_norm, inv_samples;
  cplx d_const;
  const cplx *uc, *wc, *starts, *targets;
  cplx *zc, *lc;
  double* sample_out;
  int want_grad;
};

__global__ void __launch_bounds__(128) k_tp_combine(const CombineArgs A) {
  // Sequential over the chunks, so latency is everything: the propagator blocks of the NEXT chunk are
  // staged global -> registers -> shared memory while the current chunk (already in shared memory) is applied.
  __shared__ cplx z[5][16], y[5][16];
  __shared__ cplx s_u[2][16 * 17], s_w[2][16 * 17];   // [buffer][column i][row r] of U_c / W_c (n, L <= 16); row stride L + 1:
                                                  // the backward walk reads columns (r * LS + a) without bank conflicts
  __shared__ cplx s_d;
  const int tid = threadIdx.x, L = A.lpg, n = A.n, NV = A.nv, NVP = A.nv + 1, va = A.va;
  const int v = tid / L, r = tid % L, nL = n * L, LS = L + 1;
  const bool mine = v < NVP;
  const cplx czero = cmake(0, 0);
  // each thread moves up to 4 elements of [U_c | W_c] per chunk; source and destination are fixed, so they are
  // worked out once (the walk executes one warp per scheduler: every instruction of the loop body counts)
  cplx stage[4];                                   // 2 * nL <= 512 elements over 128 threads
  const cplx* src[4];
  cplx* dst[4];                                    // position in buffer 0; buffer 1 is 16 * 17 elements further
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int e = tid + 128 * i;                   // element of [U_c | W_c]
    const bool is_w = e >= nL;
    const int f = is_w ? e - nL : e;
    src[i] = f < nL ? (is_w ? A.wc : A.uc) + f : nullptr;
    dst[i] = (is_w ? &s_w[0][0] : &s_u[0][0]) + (f / L) * LS + f % L;
  }
  auto fetch = [&](int c) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (src[i]) stage[i] = src[i][(size_t)c * nL];
  };
  auto commit = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (src[i]) dst[i][buf * 16 * 17] = stage[i];
  };
  if (mine) {
    z[v][r] = (v < NV && r < n) ? A.starts[v * L + r] : czero;
    A.zc[(size_t)v * L + r] = z[v][r];
  }
  fetch(0);
  commit(0);
  __syncthreads();
  for (int c = 0; c < A.n_chunks; ++c) {
    const int buf = c & 1;
    if (c + 1 < A.n_chunks) fetch(c + 1);
    // the walk is one dependent chain of mat-vecs: four independent accumulators per matrix keep the FP64
    // dependency depth of a step at n/4 complex multiply-adds instead of n
    cplx au[4] = {czero, czero, czero, czero}, aw[4] = {czero, czero, czero, czero};
    if (mine && r < n) {
      int i = 0;
      for (; i + 3 < n; i += 4) {
#pragma unroll
        for (int x = 0; x < 4; ++x) au[x] = cfma(s_u[buf][(i + x) * LS + r], z[v][i + x], au[x]);
      }
      for (; i < n; ++i) au[0] = cfma(s_u[buf][i * LS + r], z[v][i], au[0]);
      if (v == NV) {
        i = 0;
        for (; i + 3 < n; i += 4) {
#pragma unroll
          for (int x = 0; x < 4; ++x) aw[x] = cfma(s_w[buf][(i + x) * LS + r], z[va][i + x], aw[x]);
        }
        for (; i < n; ++i) aw[0] = cfma(s_w[buf][i * LS + r], z[va][i], aw[0]);
      }
    }
    const cplx acc = cadd(cadd(cadd(au[0], au[1]), cadd(au[2], au[3])), cadd(cadd(aw[0], aw[1]), cadd(aw[2], aw[3])));
    __syncthreads();                               // everyone has read z and buffer `buf`
    if (mine) {
      z[v][r] = acc;
      A.zc[((size_t)(c + 1) * NVP + v) * L + r] = acc;
    }
    if (c + 1 < A.n_chunks) commit(buf ^ 1);
    __syncthreads();
  }
  if (tid == 0) {
    cplx d = A.d_const, qp = czero;
    for (int vv = 0; vv < NV; ++vv)
      for (int i = 0; i < n; ++i) d = cfma_conj(A.targets[vv * L + i], z[vv][i], d);
    for (int i = 0; i < n; ++i) qp = cfma_conj(z[va][i], z[NV][i], qp);
    const double infid = 1.0 - (d.x * d.x + d.y * d.y) * A.inv_fid_norm;
    const double adiab = 1.0 - qp.x * A.inv_duration;
    A.sample_out[0] = A.w_infid * infid + A.w_adiab * adiab;
    A.sample_out[1] = infid;
    A.sample_out[2] = adiab;
    s_d = d;
  }
  __syncthreads();
  if (!A.want_grad) return;
  // terminal cotangents (same expressions as the TP = 0 kernel)
  if (mine) {
    const double ca = -0.5 * A.w_adiab * A.inv_duration * A.inv_samples;
    cplx lam = czero;
    if (r < n) {
      if (v == NV) lam = cscale(z[va][r], ca);
      else {
        lam = cscale(cmul(s_d, A.targets[v * L + r]), -A.w_infid * A.inv_fid_norm * A.inv_samples);
        if (v == va) lam = cadd(lam, cscale(z[NV][r], ca));
      }
    }
    y[v][r] = lam;
    A.lc[((size_t)A.n_chunks * NVP + v) * L + r] = lam;
  }
  // backward walk: lam_c = P_c^dag lam_{c+1}; thread (v, r) uses COLUMN r of U_c (and of W_c for the q vector);
  // the last chunk's blocks are still in buffer (n_chunks-1)&1
  __syncthreads();
  for (int c = A.n_chunks - 1; c >= 0; --c) {
    const int buf = c & 1;
    if (c > 0) fetch(c - 1);
    cplx au[4] = {czero, czero, czero, czero}, aw[4] = {czero, czero, czero, czero};
    if (mine && r < n) {
      int a = 0;
      for (; a + 3 < n; a += 4) {
#pragma unroll
        for (int x = 0; x < 4; ++x) au[x] = cfma_conj(s_u[buf][r * LS + a + x], y[v][a + x], au[x]);
      }
      for (; a < n; ++a) au[0] = cfma_conj(s_u[buf][r * LS + a], y[v][a], au[0]);
      if (v == va) {
        a = 0;
        for (; a + 3 < n; a += 4) {
#pragma unroll
          for (int x = 0; x < 4; ++x) aw[x] = cfma_conj(s_w[buf][r * LS + a + x], y[NV][a + x], aw[x]);
        }
        for (; a < n; ++a) aw[0] = cfma_conj(s_w[buf][r * LS + a], y[NV][a], aw[0]);
      }
    }
    const cplx acc = cadd(cadd(cadd(au[0], au[1]), cadd(au[2], au[3])), cadd(cadd(aw[0], aw[1]), cadd(aw[2], aw[3])));
    __syncthreads();
    if (mine) {
      y[v][r] = acc;
      A.lc[((size_t)c * NVP + v) * L + r] = acc;
    }
    if (c > 0) commit(buf ^ 1);
    __syncthreads();
  }
}

int qocvl_tp_configure(qocvl_plan* p) {
  const qocvl_desc& d = p->d;
  const int L = p->npad, n = d.n, M = d.n_slices, nvp = p->n_live + 1;
  // depth ~ chunk_len * 4 q Taylor steps (A + C) + 2 * n_chunks combine steps: balance the two
  int lc = (int)lround(sqrt(2.0 * (double)M / (double)std::max(4, p->qcap_apriori)));   // measured optimum on cfg 2
  if (p->has_opt("tp_chunk")) lc = p->opt("tp_chunk");
  lc = std::max(2, std::min(lc, M));
  p->tp_chunk = lc;
  p->tp_nchunks = (M + lc - 1) / lc;
  const int gpw = 32 / L;
  p->tp_ncol_pad = (n + gpw - 1) / gpw * gpw;      // columns of one chunk fill whole warps
  const size_t blk = (size_t)p->tp_nchunks * n * L, bnd = (size_t)(p->tp_nchunks + 1) * nvp * L;
  for (cplx** q : {&p->d_tp_uc, &p->d_tp_wc, &p->d_tp_zc, &p->d_tp_lc}) if (*q) { cudaFree(*q); *q = nullptr; }
  QCUDA(cudaMalloc((void**)&p->d_tp_uc, blk * sizeof(cplx)));
  QCUDA(cudaMalloc((void**)&p->d_tp_wc, blk * sizeof(cplx)));
  QCUDA(cudaMalloc((void**)&p->d_tp_zc, bnd * sizeof(cplx)));
  QCUDA(cudaMalloc((void**)&p->d_tp_lc, bnd * sizeof(cplx)));
  // launch geometry: 4 warps per CTA for both sweeps
  const int ngroups = 4 * gpw;
  p->tp_threads = ngroups * L;
  p->tp_blocks_a = (p->tp_nchunks * p->tp_ncol_pad + ngroups - 1) / ngroups;
  p->tp_blocks_c = (p->tp_nchunks + ngroups - 1) / ngroups;
  qocvl_plan tmp_geom = *p;                         // smem sizes depend on nv: A uses nv = 1, C uses n_live
  tmp_geom.n_live = 1;
  p->tp_smem_a = chain_smem_bytes(&tmp_geom, ngroups, 0);
  p->tp_smem_c = chain_smem_bytes(p, ngroups, 0);
  chain_fn fa = pick_tp_prop(p), fc = pick_tp_sweep(p);
  if (!fa || !fc) return QOCVL_ERR_UNSUPPORTED;
  QCUDA(cudaFuncSetAttribute(fa, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->tp_smem_a));
  QCUDA(cudaFuncSetAttribute(fc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->tp_smem_c));
  const size_t ck = (size_t)lc * nvp * p->tp_blocks_c * p->tp_threads;
  if (p->d_ckpt) { cudaFree(p->d_ckpt); p->d_ckpt = nullptr; p->bytes -= p->ckpt_bytes; }
  QCUDA(cudaMalloc((void**)&p->d_ckpt, ck * sizeof(cplx)));
  p->ckpt_bytes = ck * sizeof(cplx);
  p->bytes += p->ckpt_bytes + (2 * blk + 2 * bnd) * sizeof(cplx);
  p->n_parts = 1;
  return QOCVL_OK;
}

static void fill_common_args(qocvl_plan* p, ChainArgs& a, bool want_grad) {
  const qocvl_desc& d = p->d;
  a.n = d.n; a.J = d.n_gen; a.M = d.n_slices; a.C = d.n_chan; a.S = p->n_samples;
  a.nv = p->n_live; a.va = p->adiab_live;
  a.w_u = p->lu.n_row_slots - 1; a.wc_u = p->lu.n_col_slots - 1; a.kc_u = p->lu.kc;
  a.w_w = p->lw.nnz ? p->lw.n_row_slots : 0; a.wc_w = p->lw.nnz ? p->lw.n_col_slots : 0; a.kc_w = p->lw.kc;
  a.qcap = p->qcap; a.want_grad = want_grad ? 1 : 0;
  a.nwarps_total = (int)p->n_parts;
  a.dt = d.dt; a.inv_duration = 1.0 / d.duration; a.w_infid = d.w_infid; a.w_adiab = d.w_adiab;
  a.inv_fid_norm = 1.0 / d.fid_norm; a.inv_samples = 1.0 / (double)p->global_samples;
  a.d_const = p->d_const;
  a.u_row_col = p->lu.row_col; a.u_col_row = p->lu.col_row;
  a.u_rc_gen = p->lu.rc_gen; a.u_cc_gen = p->lu.cc_gen; a.u_rc_val = p->lu.rc_val; a.u_cc_val = p->lu.cc_val;
  a.w_row_col = p->lw.row_col; a.w_col_row = p->lw.col_row;
  a.w_rc_gen = p->lw.rc_gen; a.w_cc_gen = p->lw.cc_gen; a.w_rc_val = p->lw.rc_val; a.w_cc_val = p->lw.cc_val;
  a.coef = p->d_coef; a.dcoef = p->d_dcoef; a.qdeg = p->d_qdeg; a.mul = p->d_mul; a.add = p->d_add;
  a.starts = p->d_starts; a.targets = p->d_targets; a.ckpt = p->d_ckpt; a.sample_out = p->d_sample_out;
  a.gw_partial = p->d_gw_partial;
  a.chunk_len = p->tp_chunk; a.n_chunks = p->tp_nchunks; a.ncol_pad = p->tp_ncol_pad;
  a.tp_uc = p->d_tp_uc; a.tp_wc = p->d_tp_wc; a.tp_zc = p->d_tp_zc; a.tp_lc = p->d_tp_lc;
  a.tp_gwave = nullptr;
}

// gwave: [M][C] device buffer that receives d cost / d wave directly (no partial reduction needed)
int qocvl_tp_launch(qocvl_plan* p, bool want_grad, double* gwave, cudaStream_t st) {
  const qocvl_desc& d = p->d;
  ChainArgs a;
  fill_common_args(p, a, want_grad);
  ChainArgs pa = a;                                 // A: one column per work item, q = the column, p = W column
  pa.nv = 1; pa.va = 0; pa.want_grad = 0;
  pick_tp_prop(p)<<<p->tp_blocks_a, p->tp_threads, p->tp_smem_a, st>>>(pa);
  CombineArgs c;
  c.n = d.n; c.nv = p->n_live; c.va = p->adiab_live; c.n_chunks = p->tp_nchunks; c.lpg = p->npad;
  c.inv_duration = a.inv_duration; c.w_infid = a.w_infid; c.w_adiab = a.w_adiab; c.inv_fid_norm = a.inv_fid_norm;
  c.inv_samples = a.inv_samples; c.d_const = a.d_const;
  c.uc = p->d_tp_uc; c.wc = p->d_tp_wc; c.starts = p->d_starts; c.targets = p->d_targets;
  c.zc = p->d_tp_zc; c.lc = p->d_tp_lc; c.sample_out = p->d_sample_out; c.want_grad = want_grad ? 1 : 0;
  k_tp_combine<<<1, 128, 0, st>>>(c);
  p->launches += 2;
  if (want_grad) {
    a.tp_gwave = gwave;
    pick_tp_sweep(p)<<<p->tp_blocks_c, p->tp_threads, p->tp_smem_c, st>>>(a);
    p->launches++;
  }
  QCUDA(cudaGetLastError());
  return QOCVL_OK;
}
