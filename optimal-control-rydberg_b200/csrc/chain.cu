// chain.cu -- the hot kernel: time-ordered propagation of the columns the cost reads, through
// every slice of every noise sample, and the exact reverse sweep that yields d cost / d wave.
//
// Replaces (numerically, not structurally) ST:247-302 / TQ:460-567 plus the reverse pass
// jax.value_and_grad builds for them (TQ:592).  Formulation (oracle/vector_model.py is the CPU
// twin): the Van Loan slice generator is X_k = [[A_k, B_k],[0, A_k]] with
//   A_k = dt * sum_j c_kj G^u_j,   B_k = dt * sum_j c_kj G^w_j,   c_kj = coef_kj*mul_sj + add_sj.
// exp(X_k) is applied to vectors by a truncated Taylor series (degree q_k chosen from |X_k|_1
// so the remainder is < 2^-60): t_0 = z, t_j = X t_{j-1}/j, z' = sum t_j.  The columns needed are
//   z_v = U x_v (one per start ket that moves) and p = W psi (coupled to q = U psi through B_k).
// Reverse sweep of the same recurrence:  tb_q = lam', tb_{j-1} = lam' + X^dag tb_j / j,
//   Xbar += conj(tb_j) t_{j-1}^T / j, contracted on the fly with the generators' sparsity.
//
// Mapping: one vector per group of LPG lanes (row r of the vector on lane r), vectors of a
// sample in the same CTA.  Sparse rows of A_k live in registers (WMAX off-diagonals),
// gathers go through a double-buffered shared-memory exchange, Taylor terms t_j of the
// reverse sweep are parked in shared memory [j][thread].  The (q, p) pair of a sample sits in
// adjacent groups of one warp so the B_k coupling needs only __syncwarp.
// Pipeline: ONE __syncthreads per slice.  While slice k is being propagated, the rows of
// A_{k+1}, B_{k+1} are assembled into the other half of a double buffer (work split over all
// groups of the sample) and the coefficients of slice k+2 are already in flight from global memory.
#include "common.cuh"
#include <algorithm>

#define ROLE_Q 0  // U psi: couples INTO p (forward), receives from p (reverse)
#define ROLE_P 1  // W psi
#define ROLE_Z 2  // any other start ket

__constant__ double c_inv[QOCVL_QMAX + 2];   // 1/j

struct ChainArgs {
  int n, J, M, C, S, R, adiab_index, spb, groups;
  int w_u, wc_u, kc_u;   // off-diagonal row slots / col slots / contributions (u family)
  int w_w, wc_w, kc_w;   // (w family, no diagonal slot)
  int qcap, want_grad, nblocks;
  double dt, inv_duration, w_infid, w_adiab, inv_fid_norm, inv_samples;
  cplx d_const;          // overlap contribution of the start kets that never move
  const int *u_row_col, *u_col_row, *u_col_src, *u_rc_gen, *u_cc_gen;
  const cplx *u_rc_val, *u_cc_val;
  const int *w_row_col, *w_col_row, *w_col_src, *w_rc_gen, *w_cc_gen;
  const cplx *w_rc_val, *w_cc_val;
  const cplx* coef;    // [M][J]
  const cplx* dcoef;   // [M][J][C]
  const int* qdeg;     // [M]
  const double* mul;   // [S][J]
  const cplx* add;     // [S][J]
  const cplx* starts;  // [R-1][LPG]
  const cplx* targets;
  cplx* ckpt;          // [M][nblocks][threads]
  double* sample_out;  // [S][3]
  double* gw_partial;  // [nblocks][M][C]
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

template <int LPG, int WMAX, int WBMAX>
__global__ void __launch_bounds__(512) k_chain(const ChainArgs P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int g = tid / LPG, r = tid % LPG;
  const int J = P.J, C = P.C, spb = P.spb, JC = P.J * P.C;
  const int nslot_u = P.w_u + 1;             // row-owner slots incl. diagonal
  const int nslot_w = P.w_w > 0 ? P.w_w : 1;
  const int nwarps = (nthreads + 31) >> 5, warp = tid >> 5, lane = tid & 31;

  // ---- shared memory carve-up -------------------------------------------------------
  cplx* s_coef = (cplx*)smem_raw;                               // [3][J]      stage ring
  cplx* s_dcoef = s_coef + 3 * J;                               // [3][J][C]
  cplx* s_add = s_dcoef + 3 * JC;                               // [spb][J]
  cplx* s_au = s_add + spb * J;                                 // [2][spb][nslot_u][LPG]
  cplx* s_bw = s_au + 2 * spb * nslot_u * LPG;                  // [2][spb][nslot_w][LPG]
  cplx* s_x = s_bw + 2 * spb * nslot_w * LPG;                   // [2][groups][LPG]
  cplx* s_dpart = s_x + 2 * P.groups * LPG;                     // [groups]  overlap partials
  cplx* s_T = s_dpart + P.groups;                               // [qcap][nthreads]
  // generator-contribution tables (constant over the chain, shared by every sample of the CTA)
  const int n_urc = nslot_u * P.kc_u * LPG, n_ucc = (P.wc_u + 1) * P.kc_u * LPG;
  const int n_wrc = P.w_w * P.kc_w * LPG, n_wcc = P.wc_w * P.kc_w * LPG;
  cplx* s_urcv = s_T + (size_t)P.qcap * nthreads;               // [nslot_u][kc_u][LPG]
  cplx* s_uccv = s_urcv + n_urc;                                // [wc_u+1][kc_u][LPG]
  cplx* s_wrcv = s_uccv + n_ucc;                                // [w_w][kc_w][LPG]
  cplx* s_wccv = s_wrcv + n_wrc;                                // [wc_w][kc_w][LPG]
  double* s_mul = (double*)(s_wccv + n_wcc);                    // [spb][J]
  double* s_red = s_mul + spb * J;                              // [2][nwarps][C]
  int* s_urcg = (int*)(s_red + 2 * nwarps * C);
  int* s_uccg = s_urcg + n_urc;
  int* s_wrcg = s_uccg + n_ucc;
  int* s_wccg = s_wrcg + n_wrc;
  int* s_q = s_wccg + n_wcc;                                    // [3] Taylor degree ring
  for (int i = tid; i < n_urc; i += nthreads) { s_urcv[i] = P.u_rc_val[i]; s_urcg[i] = P.u_rc_gen[i]; }
  for (int i = tid; i < n_ucc; i += nthreads) { s_uccv[i] = P.u_cc_val[i]; s_uccg[i] = P.u_cc_gen[i]; }
  for (int i = tid; i < n_wrc; i += nthreads) { s_wrcv[i] = P.w_rc_val[i]; s_wrcg[i] = P.w_rc_gen[i]; }
  for (int i = tid; i < n_wcc; i += nthreads) { s_wccv[i] = P.w_cc_val[i]; s_wccg[i] = P.w_cc_gen[i]; }

  // ---- who am I ----------------------------------------------------------------------
  const int nz = P.R - 2;                    // plain start kets per sample
  int role, s_l, v, grank;
  bool pad_group = false;
  if (g < 2 * spb) {
    s_l = g >> 1;
    role = (g & 1) ? ROLE_P : ROLE_Q;
    grank = g & 1;
    v = P.adiab_index;
  } else {
    const int gz = g - 2 * spb;
    role = ROLE_Z;
    if (nz > 0 && gz < spb * nz) {
      s_l = gz / nz;
      const int zi = gz % nz;
      v = zi < P.adiab_index ? zi : zi + 1;
      grank = 2 + zi;
    } else {  // padding group (only to fill the last warp)
      s_l = 0; v = 0; grank = 0; pad_group = true;
    }
  }
  const int sample = blockIdx.x * spb + s_l;
  const bool live = !pad_group && sample < P.S && r < P.n;

  // per-sample noise tables
  for (int i = tid; i < spb * J; i += nthreads) {
    const int sl = i / J, j = i % J;
    int sm = blockIdx.x * spb + sl;
    if (sm >= P.S) sm = P.S - 1;
    s_mul[i] = P.mul[(size_t)sm * J + j];
    s_add[i] = P.add[(size_t)sm * J + j];
  }

  // static pattern indices of my row / column (empty slots point at my own row and carry a zero value)
  int rcol[WMAX], crow[WMAX], csrc[WMAX];
#pragma unroll
  for (int s = 0; s < WMAX; ++s) {
    rcol[s] = r; crow[s] = r; csrc[s] = r;
    if (s < P.w_u) rcol[s] = P.u_row_col[(1 + s) * LPG + r];
    if (s < P.wc_u) { crow[s] = P.u_col_row[(1 + s) * LPG + r]; csrc[s] = P.u_col_src[(1 + s) * LPG + r]; }
  }
  int brcol[WBMAX], bcrow[WBMAX], bcsrc[WBMAX];
#pragma unroll
  for (int s = 0; s < WBMAX; ++s) {
    brcol[s] = r; bcrow[s] = r; bcsrc[s] = r;
    if (s < P.w_w) brcol[s] = P.w_row_col[s * LPG + r];
    if (s < P.wc_w) { bcrow[s] = P.w_col_row[s * LPG + r]; bcsrc[s] = P.w_col_src[s * LPG + r]; }
  }

  const cplx czero = cmake(0.0, 0.0);
  cplx z = czero;
  if (live && role != ROLE_P) z = P.starts[v * LPG + r];
  const cplx ytgt = (live && role != ROLE_P) ? P.targets[v * LPG + r] : czero;

  // exchange buffers: [phase][group][lane]; ph is the phase offset (0 or groups*LPG)
  cplx* const xme0 = s_x + (size_t)g * LPG;
  // partner group of the (q,p) pair: q = g-1 for p, p = g+1 for q (unused for ROLE_Z)
  const int gp = role == ROLE_P ? g - 1 : (role == ROLE_Q ? g + 1 : g);
  cplx* const xpa0 = s_x + (size_t)gp * LPG;
  const int ph_stride = P.groups * LPG;
  int ph = 0;

  cplx a_d, a_row[WMAX], b_row[WBMAX];

  // Stage s of a sweep handles slice k(s); rings: coefficients s%3, assembled rows s&1.
  // Assemble the sparse rows of A, B for stage `st` (all groups of a sample share the slots).
  auto assemble = [&](int st) {
    if (pad_group) return;
    const cplx* co = s_coef + (st % 3) * J;
    const cplx* ad = s_add + s_l * J;
    const double* mu = s_mul + s_l * J;
    cplx* au = s_au + (size_t)((st & 1) * spb + s_l) * nslot_u * LPG;
    cplx* bw = s_bw + (size_t)((st & 1) * spb + s_l) * nslot_w * LPG;
    for (int slot = grank; slot < nslot_u + P.w_w; slot += P.R) {
      cplx acc = czero;
      if (slot < nslot_u) {
        for (int kc = 0; kc < P.kc_u; ++kc) {
          const int gi = s_urcg[(slot * P.kc_u + kc) * LPG + r];
          if (gi >= 0) {
            const cplx cs = cadd(cscale(co[gi], mu[gi]), ad[gi]);
            acc = cfma(cs, s_urcv[(slot * P.kc_u + kc) * LPG + r], acc);
          }
        }
        au[slot * LPG + r] = cscale(acc, P.dt);
      } else {
        const int sw = slot - nslot_u;
        for (int kc = 0; kc < P.kc_w; ++kc) {
          const int gi = s_wrcg[(sw * P.kc_w + kc) * LPG + r];
          if (gi >= 0) {
            const cplx cs = cadd(cscale(co[gi], mu[gi]), ad[gi]);
            acc = cfma(cs, s_wrcv[(sw * P.kc_w + kc) * LPG + r], acc);
          }
        }
        bw[sw * LPG + r] = cscale(acc, P.dt);
      }
    }
  };
  auto load_rows = [&](int st) {
    const cplx* au = s_au + (size_t)((st & 1) * spb + s_l) * nslot_u * LPG;
    a_d = au[r];
#pragma unroll
    for (int s = 0; s < WMAX; ++s) a_row[s] = (s < P.w_u) ? au[(1 + s) * LPG + r] : czero;
    const cplx* bw = s_bw + (size_t)((st & 1) * spb + s_l) * nslot_w * LPG;
#pragma unroll
    for (int s = 0; s < WBMAX; ++s) b_row[s] = (s < P.w_w && role == ROLE_P) ? bw[s * LPG + r] : czero;
  };
  // one forward Taylor step: returns X t / j
  auto step_fwd = [&](cplx t, double invj) -> cplx {
    xme0[ph + r] = t;
    __syncwarp();
    const cplx* xv = xme0 + ph;
    cplx acc0 = cmul(a_d, t), acc1 = czero;
#pragma unroll
    for (int s = 0; s < WMAX; s += 2) {
      acc0 = cfma(a_row[s], xv[rcol[s]], acc0);
      if (s + 1 < WMAX) acc1 = cfma(a_row[s + 1], xv[rcol[s + 1]], acc1);
    }
    if (role == ROLE_P) {
      const cplx* xq = xpa0 + ph;
#pragma unroll
      for (int s = 0; s < WBMAX; ++s) acc1 = cfma(b_row[s], xq[brcol[s]], acc1);
    }
    ph = ph_stride - ph;
    return cscale(cadd(acc0, acc1), invj);
  };

  // ---- prologue of the forward sweep: coefficients of stages 0 and 1, rows of stage 0
  if (tid < J) {
    s_coef[tid] = P.coef[tid];
    if (P.M > 1) s_coef[J + tid] = P.coef[(size_t)J + tid];
  }
  if (tid < 2) s_q[tid] = P.qdeg[tid < P.M ? tid : 0];
  __syncthreads();
  assemble(0);

  // =================================== forward sweep ===================================
  for (int k = 0; k < P.M; ++k) {
    __syncthreads();
    const int q = s_q[k % 3];
    load_rows(k);
    // coefficients of slice k+2: in flight while this slice is propagated
    cplx cnext = czero;
    int qnext = 0;
    const bool pf = (k + 2 < P.M);
    if (pf && tid < J) cnext = P.coef[(size_t)(k + 2) * J + tid];
    if (pf && tid == 0) qnext = P.qdeg[k + 2];
    if (P.want_grad) P.ckpt[((size_t)k * P.nblocks + blockIdx.x) * nthreads + tid] = z;
    cplx t = z;
    for (int j = 1; j <= q; ++j) {
      t = step_fwd(t, c_inv[j]);
      z = cadd(z, t);
    }
    if (k + 1 < P.M) assemble(k + 1);
    if (pf && tid < J) s_coef[((k + 2) % 3) * J + tid] = cnext;
    if (pf && tid == 0) s_q[(k + 2) % 3] = qnext;
  }

  // =================================== cost ============================================
  // overlap partials  y_v^dag z_v  (q and z groups), and  q^dag p  (p groups)
  xme0[ph + r] = z;
  __syncwarp();
  cplx part = czero;
  if (role == ROLE_P) part = cfma_conj(xpa0[ph + r], z, czero);   // conj(q_r) p_r
  else part = cfma_conj(ytgt, z, czero);                          // conj(y_r) z_r
  const cplx zpartner = xpa0[ph + r];                             // q sees p, p sees q
  ph = ph_stride - ph;
  part = group_sum<LPG>(part);
  if (r == 0) s_dpart[g] = part;
  __syncthreads();
  cplx dsum = cadd(P.d_const, s_dpart[2 * s_l]);                  // constant kets + the q group
  for (int i = 0; i < nz; ++i) dsum = cadd(dsum, s_dpart[2 * spb + s_l * nz + i]);
  const cplx qp = s_dpart[2 * s_l + 1];
  if (role == ROLE_Q && r == 0 && sample < P.S) {
    const double infid = 1.0 - (dsum.x * dsum.x + dsum.y * dsum.y) * P.inv_fid_norm;
    const double adiab = 1.0 - qp.x * P.inv_duration;
    P.sample_out[(size_t)sample * 3 + 0] = P.w_infid * infid + P.w_adiab * adiab;
    P.sample_out[(size_t)sample * 3 + 1] = infid;
    P.sample_out[(size_t)sample * 3 + 2] = adiab;
  }
  if (!P.want_grad) return;

  // terminal cotangents lam = d cost / d conj(z), scaled by 1/S_global (ensemble mean)
  cplx lam = czero;
  {
    const double ws = (sample < P.S && !pad_group) ? P.inv_samples : 0.0;
    const double ca = -0.5 * P.w_adiab * P.inv_duration * ws;
    if (role == ROLE_P) {
      lam = cscale(zpartner, ca);                                 // -(wa/2T) q
    } else {
      lam = cscale(cmul(dsum, ytgt), -P.w_infid * P.inv_fid_norm * ws);
      if (role == ROLE_Q) lam = cadd(lam, cscale(zpartner, ca));  // -(wa/2T) p
    }
    if (!live) lam = czero;
  }

  // ---- prologue of the reverse sweep: stage s handles slice M-1-s
  __syncthreads();
  for (int st = 0; st < 2 && st < P.M; ++st) {
    const int kk = P.M - 1 - st;
    if (tid < J) s_coef[st * J + tid] = P.coef[(size_t)kk * J + tid];
    for (int i = tid; i < JC; i += nthreads) s_dcoef[st * JC + i] = P.dcoef[(size_t)kk * JC + i];
    if (tid == 0) s_q[st] = P.qdeg[kk];
  }
  __syncthreads();
  assemble(0);

  // =================================== reverse sweep ===================================
  for (int st = 0; st < P.M; ++st) {
    const int k = P.M - 1 - st;
    __syncthreads();
    const int q = s_q[st % 3];
    // flush the previous stage's per-warp gradient sums (slice k+1)
    if (st > 0 && tid < C) {
      double acc = 0.0;
      const double* red = s_red + ((st - 1) & 1) * nwarps * C;
      for (int w = 0; w < nwarps; ++w) acc += red[w * C + tid];
      P.gw_partial[((size_t)blockIdx.x * P.M + (k + 1)) * C + tid] = acc;
    }
    load_rows(st);
    // column-owner values: A[a][r] for the rows a of my column (and B[a][r] for the q group)
    cplx a_col[WMAX], b_col[WBMAX];
    {
      const cplx* au = s_au + (size_t)((st & 1) * spb + s_l) * nslot_u * LPG;
#pragma unroll
      for (int s = 0; s < WMAX; ++s) a_col[s] = (s < P.wc_u) ? au[csrc[s]] : czero;
      const cplx* bw = s_bw + (size_t)((st & 1) * spb + s_l) * nslot_w * LPG;
#pragma unroll
      for (int s = 0; s < WBMAX; ++s) b_col[s] = (s < P.wc_w && role == ROLE_Q) ? bw[bcsrc[s]] : czero;
    }
    // coefficients of stage st+2 (slice k-2): in flight during this slice
    const bool pf = (st + 2 < P.M);
    cplx cnext = czero, dnext = czero;
    int qnext = 0;
    if (pf && tid < J) cnext = P.coef[(size_t)(k - 2) * J + tid];
    if (pf && tid < JC) dnext = P.dcoef[(size_t)(k - 2) * JC + tid];
    if (pf && tid == 0) qnext = P.qdeg[k - 2];
    // ---- recompute forward Taylor terms t_0 .. t_{q-1}
    cplx t = P.ckpt[((size_t)k * P.nblocks + blockIdx.x) * nthreads + tid];
    s_T[tid] = t;
    for (int j = 1; j < q; ++j) {
      t = step_fwd(t, c_inv[j]);
      s_T[(size_t)j * nthreads + tid] = t;
    }
    // ---- descending adjoint recurrence with on-the-fly contraction
    cplx tb = lam;
    cplx xc_d = czero, xc[WMAX], xcb[WBMAX];
#pragma unroll
    for (int s = 0; s < WMAX; ++s) xc[s] = czero;
#pragma unroll
    for (int s = 0; s < WBMAX; ++s) xcb[s] = czero;
    for (int j = q; j >= 1; --j) {
      const double invj = c_inv[j];
      xme0[ph + r] = tb;
      __syncwarp();
      const cplx tprev = cscale(s_T[(size_t)(j - 1) * nthreads + tid], invj);
      cplx acc0 = cfma_conj(a_d, tb, czero), acc1 = czero;
      xc_d = cfma_conj(tb, tprev, xc_d);
      const cplx* xv = xme0 + ph;
#pragma unroll
      for (int s = 0; s < WMAX; ++s) {
        const cplx ta = xv[crow[s]];
        if (s & 1) acc1 = cfma_conj(a_col[s], ta, acc1);
        else acc0 = cfma_conj(a_col[s], ta, acc0);
        xc[s] = cfma_conj(ta, tprev, xc[s]);
      }
      if (role == ROLE_Q) {
        const cplx* xp = xpa0 + ph;
#pragma unroll
        for (int s = 0; s < WBMAX; ++s) {
          const cplx ta = xp[bcrow[s]];
          acc1 = cfma_conj(b_col[s], ta, acc1);
          xcb[s] = cfma_conj(ta, tprev, xcb[s]);
        }
      }
      ph = ph_stride - ph;
      tb = cadd(lam, cscale(cadd(acc0, acc1), invj));
    }
    lam = tb;
    // ---- contract Xbar with the generators: d cost / d wave[k][c]
    double gw[QOCVL_MAX_CHAN];
#pragma unroll
    for (int c = 0; c < QOCVL_MAX_CHAN; ++c) gw[c] = 0.0;
    if (live) {
      const cplx* dco = s_dcoef + (st % 3) * JC;
      const double two_dt = 2.0 * P.dt;
      for (int slot = 0; slot <= P.wc_u; ++slot) {
        cplx xval = xc_d;
#pragma unroll
        for (int s = 0; s < WMAX; ++s)
          if (slot == s + 1) xval = xc[s];
        for (int kc = 0; kc < P.kc_u; ++kc) {
          const int gi = s_uccg[(slot * P.kc_u + kc) * LPG + r];
          if (gi >= 0) {
            const cplx tmp = cscale(cmul(s_uccv[(slot * P.kc_u + kc) * LPG + r], xval),
                                    two_dt * s_mul[s_l * J + gi]);
#pragma unroll
            for (int c = 0; c < QOCVL_MAX_CHAN; ++c)
              if (c < C) gw[c] += tmp.x * dco[gi * C + c].x - tmp.y * dco[gi * C + c].y;
          }
        }
      }
      if (role == ROLE_Q) {
        for (int slot = 0; slot < P.wc_w; ++slot) {
          cplx xval = czero;
#pragma unroll
          for (int s = 0; s < WBMAX; ++s)
            if (slot == s) xval = xcb[s];
          for (int kc = 0; kc < P.kc_w; ++kc) {
            const int gi = s_wccg[(slot * P.kc_w + kc) * LPG + r];
            if (gi >= 0) {
              const cplx tmp = cscale(cmul(s_wccv[(slot * P.kc_w + kc) * LPG + r], xval),
                                      two_dt * s_mul[s_l * J + gi]);
#pragma unroll
              for (int c = 0; c < QOCVL_MAX_CHAN; ++c)
                if (c < C) gw[c] += tmp.x * dco[gi * C + c].x - tmp.y * dco[gi * C + c].y;
            }
          }
        }
      }
    }
    {
      double* red = s_red + (st & 1) * nwarps * C;
#pragma unroll
      for (int c = 0; c < QOCVL_MAX_CHAN; ++c) {
        double vsum = gw[c];
        for (int off = 16; off > 0; off >>= 1) vsum += __shfl_xor_sync(0xffffffffu, vsum, off);
        if (lane == 0 && c < C) red[warp * C + c] = vsum;
      }
    }
    if (st + 1 < P.M) assemble(st + 1);
    if (pf && tid < J) s_coef[((st + 2) % 3) * J + tid] = cnext;
    if (pf && tid < JC) s_dcoef[((st + 2) % 3) * JC + tid] = dnext;
    if (pf && tid == 0) s_q[(st + 2) % 3] = qnext;
  }
  __syncthreads();
  if (tid < C) {
    double acc = 0.0;
    const double* red = s_red + ((P.M - 1) & 1) * nwarps * C;
    for (int w = 0; w < nwarps; ++w) acc += red[w * C + tid];
    P.gw_partial[((size_t)blockIdx.x * P.M + 0) * C + tid] = acc;
  }
}

// out3 = ensemble sums of (cost, infid, adiab) / global sample count; deterministic order.
__global__ void k_reduce_samples(int S, double inv_samples, const double* __restrict__ sample_out,
                                 double* __restrict__ out3) {
  __shared__ double sh[3][256];
  const int tid = threadIdx.x;
  double a0 = 0, a1 = 0, a2 = 0;
  for (int s = tid; s < S; s += 256) {
    a0 += sample_out[(size_t)s * 3 + 0];
    a1 += sample_out[(size_t)s * 3 + 1];
    a2 += sample_out[(size_t)s * 3 + 2];
  }
  sh[0][tid] = a0; sh[1][tid] = a1; sh[2][tid] = a2;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (tid < off) { sh[0][tid] += sh[0][tid + off]; sh[1][tid] += sh[1][tid + off]; sh[2][tid] += sh[2][tid + off]; }
    __syncthreads();
  }
  if (tid < 3) out3[tid] = sh[tid][0] * inv_samples;
}

// gwave[k][c] = sum over CTAs of their partial sums
__global__ void k_reduce_gwave(int nblocks, int MC, const double* __restrict__ partial,
                               double* __restrict__ gwave) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= MC) return;
  double acc = 0.0;
  for (int b = 0; b < nblocks; ++b) acc += partial[(size_t)b * MC + i];
  gwave[i] = acc;
}

static size_t chain_smem_bytes(const qocvl_plan* p, int spb, int groups, int threads, int qcap) {
  const int J = p->d.n_gen, C = p->d.n_chan, L = p->npad;
  const int nslot_u = p->lu.n_row_slots, nslot_w = p->lw.n_row_slots;
  size_t cplx_count = (size_t)3 * J + (size_t)3 * J * C + (size_t)spb * J + (size_t)2 * spb * nslot_u * L +
                      (size_t)2 * spb * nslot_w * L + (size_t)2 * groups * L + groups + (size_t)qcap * threads;
  size_t dbl_count = (size_t)spb * J + (size_t)2 * ((threads + 31) / 32) * C;
  // generator-contribution tables: value (cplx) + generator id (int) per (slot, contribution, lane)
  const size_t w_w = p->lw.nnz ? p->lw.n_row_slots : 0, wc_w = p->lw.nnz ? p->lw.n_col_slots : 0;
  const size_t tab = ((size_t)(p->lu.n_row_slots + p->lu.n_col_slots) * p->lu.kc + (w_w + wc_w) * p->lw.kc) * L;
  return (cplx_count + tab) * sizeof(cplx) + dbl_count * sizeof(double) + (tab + 4) * sizeof(int);
}

typedef void (*chain_fn)(const ChainArgs);
static chain_fn pick_kernel(const qocvl_plan* p) {
  const bool narrow = p->lu.n_row_slots - 1 <= 2 && p->lu.n_col_slots - 1 <= 2;
  if (p->npad == 8) return narrow ? k_chain<8, 2, 2> : k_chain<8, 4, 2>;
  return narrow ? k_chain<16, 2, 2> : k_chain<16, 4, 2>;
}

// Decide the launch geometry, size the a-priori Taylor cap and (re)allocate per-sample buffers.
int qocvl_chain_configure(qocvl_plan* p) {
  if (!p->have_gens || !p->have_cost) return QOCVL_ERR_STATE;
  if (p->groups > 0) return QOCVL_OK;
  const qocvl_desc& d = p->d;
  const int L = p->npad, S = p->n_samples, J = d.n_gen, n = d.n;
  if (p->lu.n_row_slots - 1 > 4 || p->lu.n_col_slots - 1 > 4 || p->lw.n_row_slots > 2 || p->lw.n_col_slots > 2)
    return QOCVL_ERR_UNSUPPORTED;
  {
    double inv[QOCVL_QMAX + 2];
    inv[0] = 0.0;
    for (int j = 1; j < QOCVL_QMAX + 2; ++j) inv[j] = 1.0 / (double)j;
    QCUDA(cudaMemcpyToSymbol(c_inv, inv, sizeof(inv)));
  }
  // ---- start kets annihilated by every generator block never move (e.g. |00> of the CZ model):
  //      drop them from the propagation and keep their overlap as a constant.
  std::vector<cplx> live_s, live_t;
  p->d_const = cmake(0, 0);
  p->n_live = 0;
  p->adiab_live = 0;
  std::vector<int> weight(d.n_vec, 1);          // how many start kets this one stands for
  std::vector<char> moves(d.n_vec, 0);
  for (int v = 0; v < d.n_vec; ++v) {
    bool mv = (v == d.adiab_index);
    for (int j = 0; j < J && !mv; ++j)
      for (int a = 0; a < n && !mv; ++a) {
        std::complex<double> acc(0, 0);
        for (int b = 0; b < n; ++b)
          acc += p->h_gu[((size_t)j * n + a) * n + b] *
                 std::complex<double>(p->h_starts[(size_t)v * L + b].x, p->h_starts[(size_t)v * L + b].y);
        mv = std::abs(acc) > 0.0;
      }
    moves[v] = mv;
  }
  // ---- symmetry hint: verify G^u invariance, then merge start kets that are images of each other
  if (!p->h_perm.empty()) {
    const std::vector<int>& pm = p->h_perm;
    for (int j = 0; j < J; ++j)
      for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b)
          if (p->h_gu[((size_t)j * n + pm[a]) * n + pm[b]] != p->h_gu[((size_t)j * n + a) * n + b])
            return QOCVL_ERR_ARG;
    auto is_image = [&](const std::vector<cplx>& kets, int v, int w) {   // kets[w] == P kets[v] ?
      for (int i = 0; i < n; ++i) {
        const cplx x = kets[(size_t)v * L + i], y = kets[(size_t)w * L + pm[i]];
        if (x.x != y.x || x.y != y.y) return false;
      }
      return true;
    };
    std::vector<int> order;                     // representatives: the adiabatic ket first
    order.push_back(d.adiab_index);
    for (int v = 0; v < d.n_vec; ++v) if (v != d.adiab_index) order.push_back(v);
    for (size_t i = 0; i < order.size(); ++i) {
      const int v = order[i];
      if (!moves[v] || weight[v] == 0) continue;
      for (size_t k2 = i + 1; k2 < order.size(); ++k2) {
        const int w = order[k2];
        if (!moves[w] || weight[w] == 0 || w == d.adiab_index) continue;
        if (is_image(p->h_starts, v, w) && is_image(p->h_targets, v, w)) {
          weight[v] += weight[w];
          weight[w] = 0;
        }
      }
    }
  }
  for (int v = 0; v < d.n_vec; ++v) {
    if (moves[v] && weight[v] > 0) {
      if (v == d.adiab_index) p->adiab_live = p->n_live;
      for (int r = 0; r < L; ++r) {
        live_s.push_back(p->h_starts[(size_t)v * L + r]);
        // the weight rides on the target ket: overlap and terminal cotangent are both linear in it
        live_t.push_back(cscale(p->h_targets[(size_t)v * L + r], (double)weight[v]));
      }
      p->n_live++;
    } else if (!moves[v]) {
      for (int r = 0; r < n; ++r) {   // y_v^dag x_v
        const cplx y = p->h_targets[(size_t)v * L + r], x = p->h_starts[(size_t)v * L + r];
        p->d_const.x += y.x * x.x + y.y * x.y;
        p->d_const.y += y.x * x.y - y.y * x.x;
      }
    }
  }
  if (p->d_starts) { cudaFree(p->d_starts); p->d_starts = nullptr; }
  if (p->d_targets) { cudaFree(p->d_targets); p->d_targets = nullptr; }
  QCUDA(cudaMalloc((void**)&p->d_starts, live_s.size() * sizeof(cplx)));
  QCUDA(cudaMalloc((void**)&p->d_targets, live_t.size() * sizeof(cplx)));
  QCUDA(cudaMemcpy(p->d_starts, live_s.data(), live_s.size() * sizeof(cplx), cudaMemcpyHostToDevice));
  QCUDA(cudaMemcpy(p->d_targets, live_t.data(), live_t.size() * sizeof(cplx), cudaMemcpyHostToDevice));
  const int R = p->n_live + 1;
  // ---- a-priori bound on |X_k|_1 (sizes the shared-memory ring of Taylor terms).
  // |c_j| <= 1 for every pulse-driven coefficient (tanh / 1-exp regularisation, filter taps >= 0 with
  // sum < 1, projector weights are ratios <= 1); 1e-3 relative margin for filter round-off.
  std::vector<double> mulmax(J), addmax(J);
  QCUDA(cudaMemcpy(mulmax.data(), p->d_mulmax, J * sizeof(double), cudaMemcpyDeviceToHost));
  QCUDA(cudaMemcpy(addmax.data(), p->d_addmax, J * sizeof(double), cudaMemcpyDeviceToHost));
  double bound = 0.0;
  for (int b = 0; b < n; ++b) {
    double s = 0.0;
    for (int j = 0; j < J; ++j) {
      const double cmax = j < 2 ? std::abs(p->h_coef_const[j]) : 1.001;
      s += (cmax * mulmax[j] + addmax[j]) * p->h_gcol[(size_t)j * n + b];
    }
    bound = std::max(bound, s);
  }
  bound *= d.dt;
  if (!(bound == bound)) return QOCVL_ERR_ARG;
  p->qcap = qocvl_taylor_degree(bound);
  // ---- samples per CTA: aim for ~128-192 threads
  int spb = 160 / (R * L);
  if (const char* e = getenv("QOCVL_SPB")) spb = atoi(e);
  if (spb < 1) spb = 1;
  if (spb > S) spb = S;
  const int gpw = 32 / L;
  for (;;) {
    int groups = ((spb * R + gpw - 1) / gpw) * gpw;
    int threads = groups * L;
    size_t smem = chain_smem_bytes(p, spb, groups, threads, p->qcap);
    if ((smem <= 200 * 1024 && threads <= 512) || spb == 1) {
      p->spb = spb; p->groups = groups; p->threads = threads; p->chain_smem = smem;
      break;
    }
    --spb;
  }
  if (p->chain_smem > 227 * 1024 || p->threads < J * d.n_chan) {
    // the coefficient prefetch uses one thread per (generator, channel) entry
    if (p->chain_smem > 227 * 1024) return QOCVL_ERR_UNSUPPORTED;
    const int need = ((J * d.n_chan + L - 1) / L + gpw - 1) / gpw * gpw;   // pad with idle groups
    p->groups = std::max(p->groups, need);
    p->threads = p->groups * L;
    p->chain_smem = chain_smem_bytes(p, p->spb, p->groups, p->threads, p->qcap);
  }
  p->blocks = (S + p->spb - 1) / p->spb;
  QCUDA(cudaFuncSetAttribute(pick_kernel(p), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->chain_smem));
  // ---- per-sample buffers
  const size_t ck = (size_t)d.n_slices * p->blocks * p->threads;
  if (p->d_ckpt) { cudaFree(p->d_ckpt); p->d_ckpt = nullptr; p->bytes -= p->ckpt_bytes; }
  QCUDA(cudaMalloc((void**)&p->d_ckpt, ck * sizeof(cplx)));
  p->ckpt_bytes = ck * sizeof(cplx);
  if (p->d_gw_partial) { cudaFree(p->d_gw_partial); p->d_gw_partial = nullptr; p->bytes -= p->gw_partial_elems * sizeof(double); }
  p->gw_partial_elems = (size_t)p->blocks * d.n_slices * d.n_chan;
  QCUDA(cudaMalloc((void**)&p->d_gw_partial, p->gw_partial_elems * sizeof(double)));
  p->bytes += p->ckpt_bytes + p->gw_partial_elems * sizeof(double);
  return QOCVL_OK;
}

int qocvl_launch_chain(qocvl_plan* p, bool want_grad, double* out3, double* gwave, cudaStream_t st) {
  int rc = qocvl_chain_configure(p);
  if (rc) return rc;
  const qocvl_desc& d = p->d;
  ChainArgs a;
  a.n = d.n; a.J = d.n_gen; a.M = d.n_slices; a.C = d.n_chan; a.S = p->n_samples; a.R = p->n_live + 1;
  a.adiab_index = p->adiab_live; a.spb = p->spb; a.groups = p->groups;
  a.w_u = p->lu.n_row_slots - 1; a.wc_u = p->lu.n_col_slots - 1; a.kc_u = p->lu.kc;
  a.w_w = p->lw.nnz ? p->lw.n_row_slots : 0; a.wc_w = p->lw.nnz ? p->lw.n_col_slots : 0; a.kc_w = p->lw.kc;
  a.qcap = p->qcap; a.want_grad = want_grad ? 1 : 0; a.nblocks = p->blocks;
  a.dt = d.dt; a.inv_duration = 1.0 / d.duration; a.w_infid = d.w_infid; a.w_adiab = d.w_adiab;
  a.inv_fid_norm = 1.0 / d.fid_norm; a.inv_samples = 1.0 / (double)p->global_samples;
  a.d_const = p->d_const;
  a.u_row_col = p->lu.row_col; a.u_col_row = p->lu.col_row; a.u_col_src = p->lu.col_src;
  a.u_rc_gen = p->lu.rc_gen; a.u_cc_gen = p->lu.cc_gen; a.u_rc_val = p->lu.rc_val; a.u_cc_val = p->lu.cc_val;
  a.w_row_col = p->lw.row_col; a.w_col_row = p->lw.col_row; a.w_col_src = p->lw.col_src;
  a.w_rc_gen = p->lw.rc_gen; a.w_cc_gen = p->lw.cc_gen; a.w_rc_val = p->lw.rc_val; a.w_cc_val = p->lw.cc_val;
  a.coef = p->d_coef; a.dcoef = p->d_dcoef; a.qdeg = p->d_qdeg; a.mul = p->d_mul; a.add = p->d_add;
  a.starts = p->d_starts; a.targets = p->d_targets; a.ckpt = p->d_ckpt; a.sample_out = p->d_sample_out;
  a.gw_partial = p->d_gw_partial;
  if (p->timing) QCUDA(cudaEventRecord(p->ev0, st));
  pick_kernel(p)<<<p->blocks, p->threads, p->chain_smem, st>>>(a);
  if (p->timing) { QCUDA(cudaEventRecord(p->ev1, st)); p->timed_once = true; }
  k_reduce_samples<<<1, 256, 0, st>>>(p->n_samples, a.inv_samples, p->d_sample_out, out3);
  p->launches += 2;
  if (want_grad) {
    const int MC = d.n_slices * d.n_chan;
    k_reduce_gwave<<<(MC + 255) / 256, 256, 0, st>>>(p->blocks, MC, p->d_gw_partial, gwave);
    p->launches++;
  }
  QCUDA(cudaGetLastError());
  return QOCVL_OK;
}
