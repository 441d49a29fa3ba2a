// chain.cu -- the hot kernel: time-ordered propagation of the columns the cost reads, through
// every slice of every noise sample, and the exact reverse sweep that yields d cost / d wave.
//
// Replaces (numerically, not structurally) ST:247-302 / TQ:460-567 plus the reverse pass
// jax.value_and_grad builds for them (TQ:592).  Formulation (oracle/vector_model.py is the CPU
// twin): the Van Loan slice generator is X_k = [[A_k, B_k],[0, A_k]] with
//   A_k = dt * sum_j c_kj G^u_j,   B_k = dt * sum_j c_kj G^w_j,   c_kj = coef_kj*mul_sj + add_sj.
// exp(X_k) is applied to vectors by a truncated Taylor series (degree q_k chosen from |X_k|_1
// so the remainder is < 2^-60): t_0 = z, t_j = X t_{j-1}/j, z' = sum t_j.  The columns needed are
//   z_v = U x_v (one per start ket) and p = W psi (coupled to q = U psi through B_k).
// Reverse sweep of the same recurrence:  tb_q = lam', tb_{j-1} = lam' + X^dag tb_j / j,
//   Xbar += conj(tb_j) t_{j-1}^T / j, contracted on the fly with the generators' sparsity.
//
// Mapping: one vector per group of LPG lanes (row r of the vector on lane r), vectors of a
// sample in the same CTA.  Sparse rows of A_k live in registers (<= WMAX off-diagonals),
// gathers go through a double-buffered shared-memory exchange, Taylor terms t_j of the
// reverse sweep are parked in shared memory [j][thread].  The (q, p) pair of a sample sits in
// adjacent groups of one warp so the B_k coupling needs only __syncwarp.
#include "common.cuh"

#define ROLE_Q 0  // U psi: couples INTO p (forward), receives from p (reverse)
#define ROLE_P 1  // W psi
#define ROLE_Z 2  // any other start ket

struct ChainArgs {
  int n, J, M, C, S, R, n_vec, adiab_index, spb, groups;
  int w_u, wc_u, kc_u;   // off-diagonal row slots / col slots / contributions (u family)
  int w_w, wc_w, kc_w;   // (w family, no diagonal slot)
  int qcap, want_grad, nblocks;
  double dt, inv_duration, w_infid, w_adiab, inv_fid_norm, inv_samples;
  const int *u_row_col, *u_col_row, *u_col_src, *u_rc_gen, *u_cc_gen;
  const cplx *u_rc_val, *u_cc_val;
  const int *w_row_col, *w_col_row, *w_col_src, *w_rc_gen, *w_cc_gen;
  const cplx *w_rc_val, *w_cc_val;
  const cplx* coef;    // [M][J]
  const cplx* dcoef;   // [M][J][C]
  const int* qdeg;     // [M]
  const double* mul;   // [S][J]
  const cplx* add;     // [S][J]
  const cplx* starts;  // [n_vec][LPG]
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
  const int J = P.J, C = P.C, spb = P.spb;
  const int nslot_u = P.w_u + 1;             // row-owner slots incl. diagonal
  const int nslot_w = P.w_w > 0 ? P.w_w : 1;

  // ---- shared memory carve-up -------------------------------------------------------
  cplx* s_coef = (cplx*)smem_raw;                               // [2][J]
  cplx* s_dcoef = s_coef + 2 * J;                               // [2][J][C]
  cplx* s_cs = s_dcoef + 2 * J * C;                             // [spb][J]
  cplx* s_add = s_cs + spb * J;                                 // [spb][J]
  cplx* s_au = s_add + spb * J;                                 // [spb][nslot_u][LPG]
  cplx* s_bw = s_au + spb * nslot_u * LPG;                      // [spb][nslot_w][LPG]
  cplx* s_x = s_bw + spb * nslot_w * LPG;                       // [2][groups][LPG]
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
  double* s_red = s_mul + spb * J;                              // [nwarps][C]
  int* s_urcg = (int*)(s_red + ((nthreads + 31) >> 5) * C);
  int* s_uccg = s_urcg + n_urc;
  int* s_wrcg = s_uccg + n_ucc;
  int* s_wccg = s_wrcg + n_wrc;
  for (int i = tid; i < n_urc; i += nthreads) { s_urcv[i] = P.u_rc_val[i]; s_urcg[i] = P.u_rc_gen[i]; }
  for (int i = tid; i < n_ucc; i += nthreads) { s_uccv[i] = P.u_cc_val[i]; s_uccg[i] = P.u_cc_gen[i]; }
  for (int i = tid; i < n_wrc; i += nthreads) { s_wrcv[i] = P.w_rc_val[i]; s_wrcg[i] = P.w_rc_gen[i]; }
  for (int i = tid; i < n_wcc; i += nthreads) { s_wccv[i] = P.w_cc_val[i]; s_wccg[i] = P.w_cc_gen[i]; }

  // ---- who am I ----------------------------------------------------------------------
  int role, s_l, v;
  if (g < 2 * spb) {
    s_l = g >> 1;
    role = (g & 1) ? ROLE_P : ROLE_Q;
    v = P.adiab_index;
  } else {
    const int gz = g - 2 * spb, nz = P.R - 2;
    if (nz > 0 && gz < spb * nz) {
      s_l = gz / nz;
      const int zi = gz % nz;
      v = zi < P.adiab_index ? zi : zi + 1;
      role = ROLE_Z;
    } else {  // padding group
      s_l = 0; v = 0; role = ROLE_Z;
    }
  }
  const bool pad_group = (g >= 2 * spb) && !(P.R > 2 && (g - 2 * spb) < spb * (P.R - 2));
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

  // static pattern indices of my row / column
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

  // exchange buffers: [phase][group][lane]; ph is the byte-free phase offset (0 or groups*LPG)
  cplx* const xme0 = s_x + (size_t)g * LPG;
  // partner group of the (q,p) pair: q = g-1 for p, p = g+1 for q (unused for ROLE_Z)
  const int gp = role == ROLE_P ? g - 1 : (role == ROLE_Q ? g + 1 : g);
  cplx* const xpa0 = s_x + (size_t)gp * LPG;
  const int ph_stride = P.groups * LPG;
  int ph = 0;

  cplx a_d, a_row[WMAX], b_row[WBMAX];

  // Assemble the sparse rows of A_k, B_k for every sample of this CTA (done by the q groups).
  auto assemble = [&](int k, int buf) {
    if (role == ROLE_Q && !pad_group) {
      const unsigned gmask = (LPG == 32) ? 0xffffffffu : (((1u << LPG) - 1u) << ((g * LPG) & 31));
      for (int j = r; j < J; j += LPG) {
        const cplx c = s_coef[buf * J + j];
        s_cs[s_l * J + j] = cadd(cscale(c, s_mul[s_l * J + j]), s_add[s_l * J + j]);
      }
      __syncwarp(gmask);
      for (int slot = 0; slot < nslot_u; ++slot) {
        cplx acc = czero;
        for (int kc = 0; kc < P.kc_u; ++kc) {
          const int gi = s_urcg[(slot * P.kc_u + kc) * LPG + r];
          if (gi >= 0) acc = cfma(s_cs[s_l * J + gi], s_urcv[(slot * P.kc_u + kc) * LPG + r], acc);
        }
        s_au[(s_l * nslot_u + slot) * LPG + r] = cscale(acc, P.dt);
      }
      for (int slot = 0; slot < P.w_w; ++slot) {
        cplx acc = czero;
        for (int kc = 0; kc < P.kc_w; ++kc) {
          const int gi = s_wrcg[(slot * P.kc_w + kc) * LPG + r];
          if (gi >= 0) acc = cfma(s_cs[s_l * J + gi], s_wrcv[(slot * P.kc_w + kc) * LPG + r], acc);
        }
        s_bw[(s_l * nslot_w + slot) * LPG + r] = cscale(acc, P.dt);
      }
    }
  };
  auto load_rows = [&]() {
    const cplx* au = s_au + (size_t)s_l * nslot_u * LPG;
    a_d = au[r];
#pragma unroll
    for (int s = 0; s < WMAX; ++s) a_row[s] = (s < P.w_u) ? au[(1 + s) * LPG + r] : czero;
    const cplx* bw = s_bw + (size_t)s_l * nslot_w * LPG;
#pragma unroll
    for (int s = 0; s < WBMAX; ++s) b_row[s] = (s < P.w_w && role == ROLE_P) ? bw[s * LPG + r] : czero;
  };
  // one forward Taylor step: returns X t / j
  auto step_fwd = [&](cplx t, double invj) -> cplx {
    xme0[ph + r] = t;
    __syncwarp();
    cplx acc = cmul(a_d, t);
    const cplx* xv = xme0 + ph;
#pragma unroll
    for (int s = 0; s < WMAX; ++s)
      if (s < P.w_u) acc = cfma(a_row[s], xv[rcol[s]], acc);
    if (role == ROLE_P) {
      const cplx* xq = xpa0 + ph;
#pragma unroll
      for (int s = 0; s < WBMAX; ++s)
        if (s < P.w_w) acc = cfma(b_row[s], xq[brcol[s]], acc);
    }
    ph = ph_stride - ph;
    return cscale(acc, invj);
  };

  __syncthreads();

  // =================================== forward sweep ===================================
  for (int k = 0; k < P.M; ++k) {
    const int buf = k & 1;
    if (tid < J) s_coef[buf * J + tid] = P.coef[(size_t)k * J + tid];
    const int q = P.qdeg[k];
    __syncthreads();
    assemble(k, buf);
    __syncthreads();
    load_rows();
    if (P.want_grad) P.ckpt[((size_t)k * P.nblocks + blockIdx.x) * nthreads + tid] = z;
    cplx t = z;
    for (int j = 1; j <= q; ++j) {
      t = step_fwd(t, 1.0 / (double)j);
      z = cadd(z, t);
    }
  }

  // =================================== cost ============================================
  // overlap partials  y_v^dag z_v  (q and z groups), and  q^dag p  (p groups)
  xme0[ph + r] = z;
  __syncwarp();
  cplx part = czero;
  if (role == ROLE_P) part = cfma_conj(xpa0[ph + r], z, czero);   // conj(q_r) p_r
  else part = cfma_conj(ytgt, z, czero);                        // conj(y_r) z_r
  const cplx zpartner = xpa0[ph + r];                             // q sees p, p sees q
  ph = ph_stride - ph;
  part = group_sum<LPG>(part);
  if (r == 0) s_dpart[g] = part;
  __syncthreads();
  cplx dsum = czero;
  {
    dsum = s_dpart[2 * s_l];                                    // the q group
    const int nz = P.R - 2;
    for (int i = 0; i < nz; ++i) dsum = cadd(dsum, s_dpart[2 * spb + s_l * nz + i]);
  }
  const cplx qp = s_dpart[2 * s_l + 1];
  if (role == ROLE_Q && r == 0 && sample < P.S && !pad_group) {
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
      lam = cscale(zpartner, ca);                               // -(wa/2T) q
    } else {
      lam = cscale(cmul(dsum, ytgt), -P.w_infid * P.inv_fid_norm * ws);
      if (role == ROLE_Q) lam = cadd(lam, cscale(zpartner, ca));  // -(wa/2T) p
    }
    if (!live) lam = czero;
  }
  __syncthreads();

  // =================================== reverse sweep ===================================
  const int warp = tid >> 5, lane = tid & 31, nwarps = (nthreads + 31) >> 5;
  for (int k = P.M - 1; k >= 0; --k) {
    const int buf = k & 1;
    if (tid < J) s_coef[buf * J + tid] = P.coef[(size_t)k * J + tid];
    for (int i = tid; i < J * C; i += nthreads) s_dcoef[buf * J * C + i] = P.dcoef[(size_t)k * J * C + i];
    const int q = P.qdeg[k];
    __syncthreads();
    // flush the previous slice's per-warp gradient sums (slice k+1)
    if (k < P.M - 1 && tid < C) {
      double acc = 0.0;
      for (int w = 0; w < nwarps; ++w) acc += s_red[w * C + tid];
      P.gw_partial[((size_t)blockIdx.x * P.M + (k + 1)) * C + tid] = acc;
    }
    assemble(k, buf);
    __syncthreads();
    load_rows();
    // column-owner values: A[a][r] for the rows a of my column (and B[a][r] for the q group)
    cplx a_col[WMAX], b_col[WBMAX];
    {
      const cplx* au = s_au + (size_t)s_l * nslot_u * LPG;
#pragma unroll
      for (int s = 0; s < WMAX; ++s) a_col[s] = (s < P.wc_u) ? au[csrc[s]] : czero;
      const cplx* bw = s_bw + (size_t)s_l * nslot_w * LPG;
#pragma unroll
      for (int s = 0; s < WBMAX; ++s) b_col[s] = (s < P.wc_w && role == ROLE_Q) ? bw[bcsrc[s]] : czero;
    }
    // ---- recompute forward Taylor terms t_0 .. t_{q-1}
    cplx t = P.ckpt[((size_t)k * P.nblocks + blockIdx.x) * nthreads + tid];
    s_T[tid] = t;
    for (int j = 1; j < q; ++j) {
      t = step_fwd(t, 1.0 / (double)j);
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
      const double invj = 1.0 / (double)j;
      xme0[ph + r] = tb;
      __syncwarp();
      const cplx tprev = cscale(s_T[(size_t)(j - 1) * nthreads + tid], invj);
      cplx acc = cfma_conj(a_d, tb, czero);
      xc_d = cfma_conj(tb, tprev, xc_d);
      const cplx* xv = xme0 + ph;
#pragma unroll
      for (int s = 0; s < WMAX; ++s)
        if (s < P.wc_u) {
          const cplx ta = xv[crow[s]];
          acc = cfma_conj(a_col[s], ta, acc);
          xc[s] = cfma_conj(ta, tprev, xc[s]);
        }
      if (role == ROLE_Q) {
        const cplx* xp = xpa0 + ph;
#pragma unroll
        for (int s = 0; s < WBMAX; ++s)
          if (s < P.wc_w) {
            const cplx ta = xp[bcrow[s]];
            acc = cfma_conj(b_col[s], ta, acc);
            xcb[s] = cfma_conj(ta, tprev, xcb[s]);
          }
      }
      ph = ph_stride - ph;
      tb = cadd(lam, cscale(acc, invj));
    }
    lam = tb;
    // ---- contract Xbar with the generators: d cost / d wave[k][c]
    double gw[QOCVL_MAX_CHAN];
#pragma unroll
    for (int c = 0; c < QOCVL_MAX_CHAN; ++c) gw[c] = 0.0;
    if (live) {
      const cplx* dco = s_dcoef + buf * J * C;
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
#pragma unroll
    for (int c = 0; c < QOCVL_MAX_CHAN; ++c) {
      double vsum = gw[c];
      for (int off = 16; off > 0; off >>= 1) vsum += __shfl_xor_sync(0xffffffffu, vsum, off);
      if (lane == 0 && c < C) s_red[warp * C + c] = vsum;
    }
  }
  __syncthreads();
  if (tid < C) {
    double acc = 0.0;
    for (int w = 0; w < nwarps; ++w) acc += s_red[w * C + tid];
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
  size_t cplx_count = (size_t)2 * J + (size_t)2 * J * C + (size_t)2 * spb * J + (size_t)spb * nslot_u * L +
                      (size_t)spb * nslot_w * L + (size_t)2 * groups * L + groups + (size_t)qcap * threads;
  size_t dbl_count = (size_t)spb * J + (size_t)((threads + 31) / 32) * C;
  // generator-contribution tables: value (cplx) + generator id (int) per (slot, contribution, lane)
  const size_t w_w = p->lw.nnz ? p->lw.n_row_slots : 0, wc_w = p->lw.nnz ? p->lw.n_col_slots : 0;
  const size_t tab = ((size_t)(p->lu.n_row_slots + p->lu.n_col_slots) * p->lu.kc + (w_w + wc_w) * p->lw.kc) * L;
  return (cplx_count + tab) * sizeof(cplx) + dbl_count * sizeof(double) + tab * sizeof(int);
}

typedef void (*chain_fn)(const ChainArgs);
static chain_fn pick_kernel(int npad) { return npad == 8 ? k_chain<8, 4, 2> : k_chain<16, 4, 2>; }

// Decide the launch geometry, size the a-priori Taylor cap and (re)allocate per-sample buffers.
int qocvl_chain_configure(qocvl_plan* p) {
  if (!p->have_gens || !p->have_cost) return QOCVL_ERR_STATE;
  if (p->groups > 0) return QOCVL_OK;
  const qocvl_desc& d = p->d;
  const int L = p->npad, R = d.n_vec + 1, S = p->n_samples, J = d.n_gen;
  if (p->lu.n_row_slots - 1 > 4 || p->lu.n_col_slots - 1 > 4 || p->lw.n_row_slots > 2 || p->lw.n_col_slots > 2)
    return QOCVL_ERR_UNSUPPORTED;
  // a-priori bound on |X_k|_1: pulse-driven coefficients are bounded by the regularisation
  // (|c| <= 1 up to filter overshoot); 1.5 is the safety margin.  Drift coefficients are known.
  std::vector<double> mulmax(J), addmax(J);
  cudaMemcpy(mulmax.data(), p->d_mulmax, J * sizeof(double), cudaMemcpyDeviceToHost);
  cudaMemcpy(addmax.data(), p->d_addmax, J * sizeof(double), cudaMemcpyDeviceToHost);
  // |c_j| <= 1 for every pulse-driven coefficient (tanh / 1-exp regularisation, filter taps >= 0 with
  // sum < 1, projector weights are ratios <= 1); 1e-3 relative margin for filter round-off.
  double bound = 0.0;
  for (int b = 0; b < d.n; ++b) {
    double s = 0.0;
    for (int j = 0; j < J; ++j) {
      const double cmax = j < 2 ? std::abs(p->h_coef_const[j]) : 1.001;
      s += (cmax * mulmax[j] + addmax[j]) * p->h_gcol[(size_t)j * d.n + b];
    }
    bound = std::max(bound, s);
  }
  bound *= d.dt;
  if (!(bound == bound)) return QOCVL_ERR_ARG;
  p->qcap = qocvl_taylor_degree(bound);
  // samples per CTA: aim for ~192 threads
  int spb = 192 / (R * L);
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
  if (p->chain_smem > 227 * 1024) return QOCVL_ERR_UNSUPPORTED;
  p->blocks = (S + p->spb - 1) / p->spb;
  QCUDA(cudaFuncSetAttribute(pick_kernel(L), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->chain_smem));
  // per-sample buffers
  const size_t ck = (size_t)d.n_slices * p->blocks * p->threads;
  if (p->d_ckpt) { cudaFree(p->d_ckpt); p->d_ckpt = nullptr; }
  QCUDA(cudaMalloc((void**)&p->d_ckpt, ck * sizeof(cplx)));
  p->ckpt_bytes = ck * sizeof(cplx);
  if (p->d_gw_partial) { cudaFree(p->d_gw_partial); p->d_gw_partial = nullptr; }
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
  a.n = d.n; a.J = d.n_gen; a.M = d.n_slices; a.C = d.n_chan; a.S = p->n_samples; a.R = d.n_vec + 1;
  a.n_vec = d.n_vec; a.adiab_index = d.adiab_index; a.spb = p->spb; a.groups = p->groups;
  a.w_u = p->lu.n_row_slots - 1; a.wc_u = p->lu.n_col_slots - 1; a.kc_u = p->lu.kc;
  a.w_w = p->lw.nnz ? p->lw.n_row_slots : 0; a.wc_w = p->lw.nnz ? p->lw.n_col_slots : 0; a.kc_w = p->lw.kc;
  a.qcap = p->qcap; a.want_grad = want_grad ? 1 : 0; a.nblocks = p->blocks;
  a.dt = d.dt; a.inv_duration = 1.0 / d.duration; a.w_infid = d.w_infid; a.w_adiab = d.w_adiab;
  a.inv_fid_norm = 1.0 / d.fid_norm; a.inv_samples = 1.0 / (double)p->global_samples;
  a.u_row_col = p->lu.row_col; a.u_col_row = p->lu.col_row; a.u_col_src = p->lu.col_src;
  a.u_rc_gen = p->lu.rc_gen; a.u_cc_gen = p->lu.cc_gen; a.u_rc_val = p->lu.rc_val; a.u_cc_val = p->lu.cc_val;
  a.w_row_col = p->lw.row_col; a.w_col_row = p->lw.col_row; a.w_col_src = p->lw.col_src;
  a.w_rc_gen = p->lw.rc_gen; a.w_cc_gen = p->lw.cc_gen; a.w_rc_val = p->lw.rc_val; a.w_cc_val = p->lw.cc_val;
  a.coef = p->d_coef; a.dcoef = p->d_dcoef; a.qdeg = p->d_qdeg; a.mul = p->d_mul; a.add = p->d_add;
  a.starts = p->d_starts; a.targets = p->d_targets; a.ckpt = p->d_ckpt; a.sample_out = p->d_sample_out;
  a.gw_partial = p->d_gw_partial;
  if (p->timing) QCUDA(cudaEventRecord(p->ev0, st));
  pick_kernel(p->npad)<<<p->blocks, p->threads, p->chain_smem, st>>>(a);
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
