// pulse.cu -- pulse map (Gaussian basis -> regularisation -> frequency filter), auxiliary
// amplitudes with their derivatives, and the reverse (VJP) of all of it.
//   forward : ST:173-245, TQ:319-455        reverse: hand-derived, mirrors oracle/vector_model.py
#include "common.cuh"
#include "model.cuh"

__device__ __forceinline__ double basis_time(int t, int nt) {
  // numpy.linspace(-1, 1, nt): start + t*step, last point forced to the stop value
  const double step = 2.0 / (double)(nt - 1);
  return (t == nt - 1) ? 1.0 : (-1.0 + (double)t * step);
}

// raw[t][c] = sum_n a[n][c] exp(-(tau_t - tanh b)^2 / tanh(c)^2)   (ST:173-202, TQ:319-353)
// reg = model regularisation (ST:204-227, TQ:355-373)
// Eight lanes per basis point, one channel per lane (C <= 5): the N-term sums of the channels run side by side, then
// lane 0 of the group applies the regularisation to the gathered channel values.  tanh(b), tanh(c) once per CTA.
__global__ void __launch_bounds__(128) k_raw_reg(int model, int nt, int N, int C, const double* __restrict__ abc,
                                                 double* __restrict__ raw, double* __restrict__ reg) {
  extern __shared__ double s_bc[];                 // [2][N][C]: tanh(b), tanh(c)
  for (int i = threadIdx.x; i < N * C; i += blockDim.x) {
    s_bc[i] = tanh(abc[(size_t)N * C + i]);
    s_bc[N * C + i] = tanh(abc[(size_t)2 * N * C + i]);
  }
  __syncthreads();
  const int t = blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3), ch = threadIdx.x & 7;
  const bool ok = t < nt;
  const double tau = basis_time(ok ? t : 0, nt);
  double acc = 0.0;
  if (ok && ch < C) {
    for (int n = 0; n < N; ++n) {
      const double bt = s_bc[n * C + ch], ct = s_bc[N * C + n * C + ch];
      const double dev = tau - bt;
      acc += exp(-(dev * dev) / (ct * ct)) * abc[n * C + ch];
    }
    raw[(size_t)t * C + ch] = acc;
  }
  double r[QOCVL_MAX_CHAN];
#pragma unroll
  for (int c = 0; c < QOCVL_MAX_CHAN; ++c) r[c] = __shfl_sync(0xffffffffu, acc, (threadIdx.x & 24) + c);
  if (!ok || ch != 0) return;
  double* o = reg + (size_t)t * C;
  if (model == QOCVL_MODEL_CHAIN) {                // Omega: 1 - exp(-x^2) in [0,1);  Delta: tanh(x) in (-1,1)
    o[0] = 1.0 - exp(-(r[0] * r[0]));
    o[1] = tanh(r[1]);
  } else if (model == QOCVL_MODEL_TWO_QUBIT) {
    o[0] = 0.0 * r[0] - 1.0;                       // TQ:364
    o[1] = 1.0 - exp(-(r[1] * r[1]));              // TQ:365
    o[2] = tanh(r[2]);
    o[3] = 1.0 - exp(-(r[3] * r[3]));
    o[4] = tanh(r[4]);
  } else {
    o[0] = tanh(r[0]);                             // ST:222
    for (int p0 = 1; p0 <= 3; p0 += 2) {           // ST:204-213
      const double rad = sqrt(r[p0] * r[p0] + r[p0 + 1] * r[p0 + 1]);
      const double sc = (rad == 0.0) ? 1.0 : tanh(rad) / rad;
      o[p0] = sc * r[p0];
      o[p0 + 1] = sc * r[p0 + 1];
    }
  }
}

// wave[k][c] = sum_t taps[(k - t - pad) mod M] reg[t][c]      (TQ:375-397 as a real convolution)
// transpose=1: out[t][c] = sum_k taps[(k - t - pad) mod M] in[k][c]
__global__ void k_filter(int M, int nt, int pad, int C, int transpose, const double* __restrict__ taps,
                         const double* __restrict__ in, double* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int n_out = transpose ? nt : M, n_in = transpose ? M : nt;
  if (warp >= n_out) return;
  double acc[QOCVL_MAX_CHAN];
#pragma unroll
  for (int c = 0; c < QOCVL_MAX_CHAN; ++c) acc[c] = 0.0;
  for (int i = lane; i < n_in; i += 32) {
    int idx = transpose ? (i - warp - pad) : (warp - i - pad);
    idx %= M;
    if (idx < 0) idx += M;
    const double h = taps[idx];
#pragma unroll
    for (int c = 0; c < QOCVL_MAX_CHAN; ++c)
      if (c < C) acc[c] = fma(h, in[(size_t)i * C + c], acc[c]);
  }
#pragma unroll
  for (int c = 0; c < QOCVL_MAX_CHAN; ++c) {
    double v = acc[c];
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0 && c < C) out[(size_t)warp * C + c] = v;
  }
}

// Per-slice generator coefficients and their derivatives with respect to the C real pulse values
// of that slice (model.cuh), plus the Taylor degree for the slice from the norm bound.
struct NormTables {   // CSC views of the two block families (plan.cu / chain_big.cu::qocvl_build_csr)
  const int *u_colptr, *u_ceid, *u_egen; const cplx* u_eval; int u_kc;
  const int *w_colptr, *w_ceid, *w_egen; const cplx* w_eval; int w_kc;
  const double* gcol;
  const int *u_crow, *w_crow, *u_rowptr, *u_col, *w_rowptr, *w_col;   // CSR + CSC views (k_role_degrees)
};

__global__ void k_coefficients(ModelParams mp, int M, const double* __restrict__ wave,
                               const cplx* __restrict__ coef_const, NormTables nt,
                               const double* __restrict__ mulmax, const double* __restrict__ addmax, int qcap,
                               double taylor_tol, cplx* __restrict__ coef, cplx* __restrict__ dcoef, int* __restrict__ qdeg,
                               int* __restrict__ flag, const unsigned char* __restrict__ colrole, int nroles,
                               int* __restrict__ qdeg_role, int do_norm) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= M) return;
  const int J = mp.J, C = mp.C;
  cplx co[QOCVL_MAX_GEN];
  cplx dc[QOCVL_MAX_GEN][QOCVL_MAX_CHAN];
  qocvl_model_coefficients(mp, k, wave + (size_t)k * C, coef_const, co, dc);
  for (int j = 0; j < J; ++j) {
    coef[(size_t)k * J + j] = co[j];
    for (int c = 0; c < C; ++c) dcoef[((size_t)k * J + j) * C + c] = dc[j][c];
  }
  if (!do_norm) return;                        // the norms and degrees come from k_slice_degrees (one warp per slice)
  double role_nrm[4] = {0.0, 0.0, 0.0, 0.0};   // (DUO_MAXR roles)
  const double nrm = qocvl_slice_norm_csr(mp, co, nt.u_colptr, nt.u_ceid, nt.u_egen, nt.u_eval, nt.u_kc,
                                          nt.w_colptr, nt.w_ceid, nt.w_egen, nt.w_eval, nt.w_kc, nt.gcol,
                                          mulmax, addmax, colrole, nroles, role_nrm);
  int q = qocvl_taylor_degree(nrm, taylor_tol);
  bool bad = false;
  if (!(nrm <= QOCVL_NORM_MAX) || q > qcap) {   // also catches NaN pulses
    atomicOr(flag, 1);
    q = 2;
    bad = true;
    // fail loudly: poison the slice so the cost and gradient come out NaN instead of silently wrong
    coef[(size_t)k * J] = cmake(nan(""), nan(""));
  }
  qdeg[k] = q;
  // Taylor degree of every role of the two-lane kernels (chain_duo.cu): each role is a diagonal block of the
  // propagated system and gets the degree ITS norm needs (CZ: the {U psi, W psi} pair has half the norm of U|11>)
  if (colrole)
    for (int r = 0; r < nroles; ++r)
      qdeg_role[(size_t)r * M + k] = bad ? 2 : min(q, qocvl_taylor_degree(role_nrm[r], taylor_tol));
}

// Per-role 2-norm bounds for the Taylor degrees of the two-lane kernels (chain_duo.cu): qdeg_role[r][k] comes in as the
// degree the role's 1-norm bound asks for (k_coefficients) and leaves as the smaller of that and the degree of its 2-norm bound:
//   |Y(s)|_2 <= | |Y(s)| |_2 <= |N|_2 = sqrt(rho(N^T N)) <= sqrt(max_i (N^T N v)_i / v_i)   for ANY positive vector v
// (entrywise absolute values only raise the 2-norm, it is monotone on non-negative matrices, and the last step is the
// Collatz-Wielandt bound), where N >= |Y(s)| entrywise for every noise sample:  N_e = |nominal entry| + the entrywise
// deviation bound used above.  N is the non-negative image of [[A, 0], [B, A]] restricted to the role's (q, p) states
// (colrole: [2][n] role masks of the q part and the p part; the states of a role are closed under the pattern, so the
// roles are diagonal blocks and one iteration on the union serves them all).  Three power steps make v the Perron vector
// and the bound tight: for the reference's models | |Y| |_2 = |Y|_2 to 3 digits and the bound is 0.78 of the 1-norm
// (CZ: mean degree 9.9 -> 9.4 and 11.9 -> 11.0).  Rigorous for every v, so the iteration count only affects tightness.
// One warp per slice, lane = basis state (n <= 32): the kernel also computes what k_coefficients computes serially for
// larger systems -- the whole-system 1-norm bound, its degree, the poison-and-flag rule -- so that the pulse-map stage
// of a single-sample evaluation is not dominated by a serial norm loop.  The magnitudes N_e go to shared memory once;
// column sums (1-norms) and t = N^T y run over the columns (CSC view), y = N v over the rows (CSR view) of the lane's state.
#define QOCVL_N2_EMAX_U 160
#define QOCVL_N2_EMAX_W 96
#define QOCVL_N2_STEPS 3
__global__ void __launch_bounds__(128) k_slice_degrees(int n, int J, double dt, int M, cplx* __restrict__ coef,
                                                       NormTables nt, const double* __restrict__ mulmax,
                                                       const double* __restrict__ addmax, int qcap, double taylor_tol,
                                                       int* __restrict__ qdeg, int* __restrict__ flag,
                                                       const unsigned char* __restrict__ colrole, int nroles, int norm2,
                                                       int* __restrict__ qdeg_role, double* __restrict__ phi) {
  __shared__ double s_mu[4][QOCVL_N2_EMAX_U], s_mw[4][QOCVL_N2_EMAX_W], s_v[4][64], s_y[4][64];
  // the index tables of the two sparse views, once per CTA (its four warps walk them in every pass of the iteration)
  __shared__ int s_ucol[QOCVL_N2_EMAX_U], s_uceid[QOCVL_N2_EMAX_U], s_ucrow[QOCVL_N2_EMAX_U];
  __shared__ int s_wcol[QOCVL_N2_EMAX_W], s_wceid[QOCVL_N2_EMAX_W], s_wcrow[QOCVL_N2_EMAX_W];
  for (int i = threadIdx.x; i < nt.u_rowptr[n]; i += blockDim.x) { s_ucol[i] = nt.u_col[i]; s_uceid[i] = nt.u_ceid[i]; s_ucrow[i] = nt.u_crow[i]; }
  for (int i = threadIdx.x; i < nt.w_rowptr[n]; i += blockDim.x) { s_wcol[i] = nt.w_col[i]; s_wceid[i] = nt.w_ceid[i]; s_wcrow[i] = nt.w_crow[i]; }
  __syncthreads();
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5, k = blockIdx.x * 4 + wl;
  if (k >= M) return;
  double *mu = s_mu[wl], *mw = s_mw[wl], *sv = s_v[wl], *sy = s_y[wl];
  cplx co[QOCVL_MAX_GEN];
  double wgt[QOCVL_MAX_GEN];
  for (int j = 0; j < J; ++j) {
    co[j] = coef[(size_t)k * J + j];
    wgt[j] = sqrt(co[j].x * co[j].x + co[j].y * co[j].y) * mulmax[J + j] + addmax[j];
  }
  for (int fam = 0; fam < 2; ++fam) {                      // N_e = (|nominal entry| + deviation bound) dt, in row order
    const int cnt = fam ? nt.w_rowptr[n] : nt.u_rowptr[n], kc = fam ? nt.w_kc : nt.u_kc;
    const int* egen = fam ? nt.w_egen : nt.u_egen;
    const cplx* eval = fam ? nt.w_eval : nt.u_eval;
    for (int e = lane; e < cnt; e += 32) {
      double re = 0, im = 0, dev = 0;
      for (int c = 0; c < kc; ++c) {
        const int g = egen[e * kc + c];
        if (g >= 0) {
          const cplx v = eval[e * kc + c];
          re += co[g].x * v.x - co[g].y * v.y; im += co[g].x * v.y + co[g].y * v.x;
          dev += wgt[g] * sqrt(v.x * v.x + v.y * v.y);
        }
      }
      (fam ? mw : mu)[e] = (sqrt(re * re + im * im) + dev) * dt;
    }
  }
  __syncwarp();
  const bool on = lane < n;
  const int ur0 = on ? nt.u_rowptr[lane] : 0, ur1 = on ? nt.u_rowptr[lane + 1] : 0;
  const int wr0 = on ? nt.w_rowptr[lane] : 0, wr1 = on ? nt.w_rowptr[lane + 1] : 0;
  const int uc0 = on ? nt.u_colptr[lane] : 0, uc1 = on ? nt.u_colptr[lane + 1] : 0;
  const int wc0 = on ? nt.w_colptr[lane] : 0, wc1 = on ? nt.w_colptr[lane + 1] : 0;
  // ---- 1-norm bounds: column n + b of [[A, B], [0, A]] sums |A_ab| + |B_ab| (qocvl_slice_norm_csr's bound, in parallel)
  double colsum = 0.0;
  for (int i = uc0; i < uc1; ++i) colsum += mu[s_uceid[i]];
  for (int i = wc0; i < wc1; ++i) colsum += mw[s_wceid[i]];
  double nrm = colsum;
  for (int off = 16; off > 0; off >>= 1) nrm = fmax(nrm, __shfl_xor_sync(0xffffffffu, nrm, off));
  const bool isnan_ = !(nrm == nrm) || !(co[0].x == co[0].x);
  int q = qocvl_taylor_degree(nrm, taylor_tol);
  const bool bad = isnan_ || !(nrm <= QOCVL_NORM_MAX) || q > qcap;   // also catches NaN pulses
  if (bad) q = 2;
  if (lane == 0) {
    if (bad) {
      atomicOr(flag, 1);
      // fail loudly: poison the slice so the cost and gradient come out NaN instead of silently wrong
      coef[(size_t)k * J] = cmake(nan(""), nan(""));
    }
    qdeg[k] = q;
  }
  if (!colrole) return;
  const unsigned rq = on ? colrole[lane] : 0u, rp = on ? colrole[n + lane] : 0u;   // roles of my state's q / p element
  // ---- phase shift per role (phi != null): the sweeps of role r run the Taylor series of  Y = X - i phi_r I  with
  //      phi_r = midpoint of the range of Im diag(X) over the role's states (exp(X) = e^{i phi} exp(Y); the phases of
  //      all slices go into the role's target kets, k_duo_phase).  From here on the diagonal lives in `dmag`.
  int ed = -1;                                             // my state's diagonal entry in A
  for (int e = ur0; e < ur1; ++e) if (s_ucol[e] == lane) ed = e;
  double dre = 0.0, dim = 0.0, ddev = 0.0;
  if (ed >= 0) {
    for (int c = 0; c < nt.u_kc; ++c) {
      const int g = nt.u_egen[ed * nt.u_kc + c];
      if (g >= 0) {
        const cplx v = nt.u_eval[ed * nt.u_kc + c];
        dre += co[g].x * v.x - co[g].y * v.y; dim += co[g].x * v.y + co[g].y * v.x;
        ddev += wgt[g] * sqrt(v.x * v.x + v.y * v.y);
      }
    }
    dre *= dt; dim *= dt; ddev *= dt;
  }
  double myphi = 0.0;
  for (int r = 0; r < nroles; ++r) {
    const bool in = ((rq | rp) >> r) & 1u;
    double hi = in ? dim : -1e300, lo = in ? -dim : -1e300;
    for (int off = 16; off > 0; off >>= 1) {
      hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, off));
      lo = fmax(lo, __shfl_xor_sync(0xffffffffu, lo, off));
    }
    const double ph = (phi && !bad && hi > -1e299) ? 0.5 * (hi - lo) : 0.0;
    if (in) myphi = ph;                                    // (the host enables the shift only if a state has ONE role)
    if (phi && lane == 0) phi[(size_t)r * M + k] = ph;
  }
  __syncwarp();
  if (ed >= 0) mu[ed] = 0.0;
  __syncwarp();
  const double dmag = sqrt(dre * dre + (dim - myphi) * (dim - myphi)) + ddev;
  colsum = dmag;                                           // column sums of the shifted generator
  for (int i = uc0; i < uc1; ++i) colsum += mu[s_uceid[i]];
  for (int i = wc0; i < wc1; ++i) colsum += mw[s_wceid[i]];
  double role_n1[4];
  for (int r = 0; r < nroles; ++r) {
    double v = (((rq | rp) >> r) & 1u) ? colsum : 0.0;
    for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
    role_n1[r] = v;
  }
  double role_n2[4] = {1e300, 1e300, 1e300, 1e300};
  if (norm2 && !bad) {
    double vq = rq ? 1.0 : 0.0, vp = rp ? 1.0 : 0.0;
    for (int it = 0; it <= QOCVL_N2_STEPS; ++it) {
      __syncwarp();
      sv[lane] = vq; sv[32 + lane] = vp;
      __syncwarp();
      double yq = dmag * vq, yp = dmag * vp;               // y = N v:  y_q = |A| v_q,  y_p = |B| v_q + |A| v_p
      for (int e = ur0; e < ur1; ++e) {
        const int c = s_ucol[e];
        yq = fma(mu[e], sv[c], yq); yp = fma(mu[e], sv[32 + c], yp);
      }
      for (int e = wr0; e < wr1; ++e) yp = fma(mw[e], sv[s_wcol[e]], yp);
      sy[lane] = rq ? yq : 0.0; sy[32 + lane] = rp ? yp : 0.0;   // rows outside every role do not exist in the propagated system
      __syncwarp();
      double tq = dmag * sy[lane], tp = dmag * sy[32 + lane];   // t = N^T y:  t_q = |A|^T y_q + |B|^T y_p,  t_p = |A|^T y_p
      for (int i = uc0; i < uc1; ++i) {
        const double m = mu[s_uceid[i]];
        const int r = s_ucrow[i];
        tq = fma(m, sy[r], tq); tp = fma(m, sy[32 + r], tp);
      }
      for (int i = wc0; i < wc1; ++i) tq = fma(mw[s_wceid[i]], sy[32 + s_wcrow[i]], tq);
      if (!rq) tq = 0.0;
      if (!rp) tp = 0.0;
      if (it == QOCVL_N2_STEPS) {
        for (int r = 0; r < nroles; ++r) {
          double best = fmax(((rq >> r) & 1u) ? tq / vq : 0.0, ((rp >> r) & 1u) ? tp / vp : 0.0);
          for (int off = 16; off > 0; off >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, off));
          const double n2 = sqrt(best) * (1.0 + 1e-12);    // (rounding of the iteration itself)
          if (n2 == n2) role_n2[r] = n2;
        }
      } else {
        double mx = fmax(tq, tp);
        for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        const double inv = mx > 0.0 ? 1.0 / mx : 0.0;
        if (rq) vq = fmax(tq * inv, 1e-3);                 // next positive vector (the bound needs v > 0)
        if (rp) vp = fmax(tp * inv, 1e-3);
      }
    }
  }
  // every role -- a diagonal block of the propagated system -- gets the degree ITS norm needs
  // (CZ: the {U psi, W psi} pair has half the norm of U|11>)
  if (lane == 0)
    for (int r = 0; r < nroles; ++r)
      qdeg_role[(size_t)r * M + k] = bad ? 2 : min(q, qocvl_taylor_degree(fmin(role_n1[r], role_n2[r]), taylor_tol));
}

// aux[k][0..J-3] = coef[k][2..J-1]   (ST:244-245, TQ:452-455)
__global__ void k_aux(int M, int J, const cplx* __restrict__ coef, cplx* __restrict__ aux) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * (J - 2)) return;
  const int k = i / (J - 2), j = i % (J - 2);
  aux[i] = coef[(size_t)k * J + 2 + j];
}

// d cost / d raw from d cost / d reg  (reverse of the regularisation)
__global__ void k_reg_vjp(int model, int nt, int C, const double* __restrict__ raw,
                          const double* __restrict__ greg, double* __restrict__ graw) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nt) return;
  const double* r = raw + (size_t)t * C;
  const double* g = greg + (size_t)t * C;
  double* o = graw + (size_t)t * C;
  if (model == QOCVL_MODEL_CHAIN) {
    o[0] = g[0] * 2.0 * r[0] * exp(-(r[0] * r[0]));
    const double th = tanh(r[1]);
    o[1] = g[1] * (1.0 - th * th);
  } else if (model == QOCVL_MODEL_TWO_QUBIT) {
    o[0] = 0.0;                                       // TQ:364 -- channel 0 is a constant
    o[1] = g[1] * 2.0 * r[1] * exp(-(r[1] * r[1]));
    o[3] = g[3] * 2.0 * r[3] * exp(-(r[3] * r[3]));
    double th = tanh(r[2]); o[2] = g[2] * (1.0 - th * th);
    th = tanh(r[4]);        o[4] = g[4] * (1.0 - th * th);
  } else {
    const double th0 = tanh(r[0]);
    o[0] = g[0] * (1.0 - th0 * th0);
    for (int p0 = 1; p0 <= 3; p0 += 2) {
      const double x = r[p0], y = r[p0 + 1];
      const double rad = sqrt(x * x + y * y);
      double sc = 1.0, dsc = 0.0;
      if (rad != 0.0) {
        const double th = tanh(rad);
        sc = th / rad;
        dsc = ((1.0 - th * th) * rad - th) / (rad * rad * rad);
      }
      const double dot = g[p0] * x + g[p0 + 1] * y;
      o[p0] = sc * g[p0] + dsc * dot * x;
      o[p0 + 1] = sc * g[p0 + 1] + dsc * dot * y;
    }
  }
}

// One warp per (n, c): ga, gb, gc from graw (reverse of the Gaussian basis, incl. the tanh clamps)
__global__ void k_basis_vjp(int nt, int N, int C, const double* __restrict__ abc,
                            const double* __restrict__ graw, double* __restrict__ grads) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= N * C) return;
  const int n = warp / C, ch = warp % C;
  const double a = abc[n * C + ch];
  const double bt = tanh(abc[(size_t)N * C + n * C + ch]);
  const double ct = tanh(abc[(size_t)2 * N * C + n * C + ch]);
  const double ic2 = 1.0 / (ct * ct), ic3 = ic2 / ct;
  double sa = 0, sb = 0, sc = 0;
  for (int t = lane; t < nt; t += 32) {
    const double dev = basis_time(t, nt) - bt;
    const double mode = exp(-(dev * dev) * ic2);
    const double g = graw[(size_t)t * C + ch] * mode;
    sa += g;
    sb += g * 2.0 * dev * ic2;
    sc += g * 2.0 * dev * dev * ic3;
  }
  for (int off = 16; off > 0; off >>= 1) {
    sa += __shfl_xor_sync(0xffffffffu, sa, off);
    sb += __shfl_xor_sync(0xffffffffu, sb, off);
    sc += __shfl_xor_sync(0xffffffffu, sc, off);
  }
  if (lane == 0) {
    grads[n * C + ch] = sa;
    grads[(size_t)N * C + n * C + ch] = a * sb * (1.0 - bt * bt);
    grads[(size_t)2 * N * C + n * C + ch] = a * sc * (1.0 - ct * ct);
  }
}

int qocvl_launch_pulse_forward(qocvl_plan* p, const double* abc, cudaStream_t st) {
  const qocvl_desc& d = p->d;
  const int nt = d.n_basis_pts, M = d.n_slices, C = d.n_chan;
  if (d.pad > 0 && !p->have_filter) return QOCVL_ERR_STATE;
  double* reg_out = p->have_filter ? p->d_reg : p->d_wave;   // no filter: wave == reg (ST)
  k_raw_reg<<<(nt + 15) / 16, 128, 2 * d.n_gauss * C * sizeof(double), st>>>(d.model, nt, d.n_gauss, C, abc, p->d_raw, reg_out);
  p->launches++;
  if (p->have_filter) {
    k_filter<<<(M * 32 + 255) / 256, 256, 0, st>>>(M, nt, d.pad, C, 0, p->d_taps, p->d_reg, p->d_wave);
    p->launches++;
  }
  QCUDA(cudaGetLastError());
  return QOCVL_OK;
}

int qocvl_launch_coefficients(qocvl_plan* p, const double* wave, cudaStream_t st) {
  const qocvl_desc& d = p->d;
  ModelParams mp = {d.model, d.n, d.n_gen, d.n_chan, d.dt, d.eps, d.scale[0], d.scale[1], d.n_slices};
  NormTables nt = {p->cu.colptr, p->cu.ceid, p->cu.egen, p->cu.eval, p->cu.kc,
                   p->cw.colptr, p->cw.ceid, p->cw.egen, p->cw.eval, p->cw.kc, p->d_gnorm,
                   p->cu.crow, p->cw.crow, p->cu.rowptr, p->cu.col, p->cw.rowptr, p->cw.col};
  const bool role_q = p->use_aug && p->use_duo && p->duo.role_q;   // per-role Taylor degrees of the two-lane kernels
  // small systems: norms, degrees and the poison rule by one warp per slice (k_slice_degrees) instead of one thread
  const bool warp_norms = d.n <= 32 && p->cu.nnz <= QOCVL_N2_EMAX_U && p->cw.nnz <= QOCVL_N2_EMAX_W;
  // large systems on the samples-minor kernel: k_wide_degrees (one CTA per slice, launched by qocvl_wide_launch) owns the
  // norms once the path is configured; the very first evaluation still runs the serial rule here
  const bool wide_norms = !warp_norms && p->groups > 0 && p->use_wide && p->wide_norms;
  k_coefficients<<<(d.n_slices + 63) / 64, 64, 0, st>>>(mp, d.n_slices, wave, p->d_coef_const, nt, p->d_mulmax,
                                                       p->d_addmax, p->qcap, p->taylor_tol(), p->d_coef, p->d_dcoef, p->d_qdeg,
                                                       p->d_flag, role_q ? p->duo.d_colrole : nullptr, role_q ? p->duo.nroles : 0,
                                                       p->duo.d_qdeg_role, (warp_norms || wide_norms) ? 0 : 1);
  p->wide_norms_fresh = wide_norms;
  p->launches++;
  if (warp_norms) {         // ("duo_roleq" = 2, the default: per-role degrees from min(1-norm, 2-norm bound); 1: 1-norm only)
    k_slice_degrees<<<(d.n_slices + 3) / 4, 128, 0, st>>>(d.n, d.n_gen, d.dt, d.n_slices, p->d_coef, nt, p->d_mulmax,
                                                         p->d_addmax, p->qcap, p->taylor_tol(), p->d_qdeg, p->d_flag,
                                                         role_q ? p->duo.d_colrole : nullptr, role_q ? p->duo.nroles : 0,
                                                         p->opt("duo_roleq", 3) >= 2 ? 1 : 0, p->duo.d_qdeg_role,
                                                         role_q && p->duo.shift_on ? p->duo.d_phi : nullptr);
    p->launches++;
  }
  QCUDA(cudaGetLastError());
  return QOCVL_OK;
}

int qocvl_launch_aux(qocvl_plan* p, double* aux, cudaStream_t st) {
  const int total = p->d.n_slices * (p->d.n_gen - 2);
  k_aux<<<(total + 255) / 256, 256, 0, st>>>(p->d.n_slices, p->d.n_gen, p->d_coef, (cplx*)aux);
  p->launches++;
  QCUDA(cudaGetLastError());
  return QOCVL_OK;
}

int qocvl_launch_pulse_vjp(qocvl_plan* p, const double* abc, const double* gwave, double* out_grads,
                           cudaStream_t st) {
  const qocvl_desc& d = p->d;
  const int nt = d.n_basis_pts, M = d.n_slices, C = d.n_chan;
  const double* greg = gwave;
  if (p->have_filter) {
    k_filter<<<(nt * 32 + 255) / 256, 256, 0, st>>>(M, nt, d.pad, C, 1, p->d_taps, gwave, p->d_greg);
    p->launches++;
    greg = p->d_greg;
  }
  k_reg_vjp<<<(nt + 127) / 128, 128, 0, st>>>(d.model, nt, C, p->d_raw, greg, p->d_graw);
  k_basis_vjp<<<(d.n_gauss * C * 32 + 127) / 128, 128, 0, st>>>(nt, d.n_gauss, C, abc, p->d_graw, out_grads);
  p->launches += 2;
  QCUDA(cudaGetLastError());
  return QOCVL_OK;
}
