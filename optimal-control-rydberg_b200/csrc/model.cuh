// model.cuh -- model-specific map from the physical amplitudes of one slice to the generator
// coefficients (and their derivatives), shared by the device kernel (pulse.cu::k_coefficients)
// and by the host's a-priori norm bound (chain.cu::qocvl_chain_configure).
//   TQ:416-476 (two-qubit), ST:236-255 (state transfer)
#pragma once
#include "common.cuh"

#define PI_D 3.14159265358979323846

struct ModelParams {
  int model, n, J, C;
  double dt, eps, s0, s1;   // s0, s1: the two Rabi scales of the model
  int M;                    // number of slices (time-dependent penalty weights of the CHAIN model)
};

// co[j]: coefficient of generator j for this slice; dc[j][c] = d co[j] / d wave[c]
__host__ __device__ inline void qocvl_model_coefficients(const ModelParams& mp, int k, const double* w,
                                                         const cplx* coef_const, cplx* co,
                                                         cplx (*dc)[QOCVL_MAX_CHAN]) {
  for (int j = 0; j < QOCVL_MAX_GEN; ++j) {
    co[j] = cmake(0, 0);
    for (int c = 0; c < QOCVL_MAX_CHAN; ++c) dc[j][c] = cmake(0, 0);
  }
  if (mp.model == QOCVL_MODEL_CHAIN) {
    // Blockade (PXP) chain, SURVEY.md 8(d) cfg 4/5 (no reference counterpart): two pulse-driven
    // Hamiltonian pieces  -i*Omega(t)/Omega_max * G0,  -i*Delta(t)/Delta_max * G1  and two penalty
    // pieces in the top-right block with the time ramp  (1 - k/M) * G2 + (k/M) * G3.  No drift.
    co[0] = cmake(0, -w[0]);  dc[0][0] = cmake(0, -1);
    co[1] = cmake(0, -w[1]);  dc[1][1] = cmake(0, -1);
    const double ramp = (double)k / (double)mp.M;
    co[2] = cmake(1.0 - ramp, 0);
    co[3] = cmake(ramp, 0);
    return;
  }
  co[0] = coef_const[0];   // drift pieces: ST:254 (+1), TQ:470 (-i)
  co[1] = coef_const[1];
  if (mp.model == QOCVL_MODEL_TWO_QUBIT) {
    const double ri = mp.s0, rr = mp.s1;
    const double si = sin(PI_D * w[2]), ci = cos(PI_D * w[2]);
    const double sr = sin(PI_D * w[4]), cr = cos(PI_D * w[4]);
    co[2] = cmake(0, -w[0]);              dc[2][0] = cmake(0, -1);
    co[3] = cmake(0, -w[1] * ci);         dc[3][1] = cmake(0, -ci); dc[3][2] = cmake(0, PI_D * w[1] * si);
    co[4] = cmake(0, -w[1] * si);         dc[4][1] = cmake(0, -si); dc[4][2] = cmake(0, -PI_D * w[1] * ci);
    co[5] = cmake(0, -w[3] * cr);         dc[5][3] = cmake(0, -cr); dc[5][4] = cmake(0, PI_D * w[3] * sr);
    co[6] = cmake(0, -w[3] * sr);         dc[6][3] = cmake(0, -sr); dc[6][4] = cmake(0, -PI_D * w[3] * cr);
    const double mi = ri * w[1], mr = rr * w[3];
    const double den_raw = mr * mr + mi * mi;
    const bool clamped = den_raw < mp.eps;   // TQ:433-434
    const double den = clamped ? mp.eps : den_raw;
    const double dd1 = clamped ? 0.0 : 2.0 * mi * ri, dd3 = clamped ? 0.0 : 2.0 * mr * rr;
    const double sp = sin(PI_D * (w[2] + w[4])), cp = cos(PI_D * (w[2] + w[4]));
    const cplx ph = cmake(cp, sp);
    const double inv = 1.0 / den, inv2 = inv * inv;
    co[7] = cmake(mr * mr * inv, 0);
    co[8] = cmake(mi * mi * inv, 0);
    co[9] = cscale(ph, -mi * mr * inv);
    co[10] = cconj(co[9]);
    dc[7][1] = cmake(-mr * mr * dd1 * inv2, 0);
    dc[7][3] = cmake(2.0 * mr * rr * inv - mr * mr * dd3 * inv2, 0);
    dc[8][1] = cmake(2.0 * mi * ri * inv - mi * mi * dd1 * inv2, 0);
    dc[8][3] = cmake(-mi * mi * dd3 * inv2, 0);
    dc[9][1] = cscale(ph, -(ri * mr * inv - mi * mr * dd1 * inv2));
    dc[9][3] = cscale(ph, -(mi * rr * inv - mi * mr * dd3 * inv2));
    dc[9][2] = cmul(cmake(0, PI_D), co[9]);
    dc[9][4] = dc[9][2];
    for (int c = 0; c < 5; ++c) dc[10][c] = cconj(dc[9][c]);
  } else {  // state transfer
    const double r1 = mp.s0, r2 = mp.s1;
    for (int c = 0; c < 5; ++c) { co[2 + c] = cmake(0, -w[c]); dc[2 + c][c] = cmake(0, -1); }
    const cplx o1 = cmake(r1 * w[1], r1 * w[2]), o2 = cmake(r2 * w[3], r2 * w[4]);
    const double n1 = o1.x * o1.x + o1.y * o1.y, n2 = o2.x * o2.x + o2.y * o2.y;
    const double den = n1 + n2, inv = 1.0 / den, inv2 = inv * inv;   // ST:242-243: no guard (quirk q7)
    const cplx o12 = cmul(o1, o2);
    co[7] = cmake(n2 * inv, 0);
    co[8] = cmake(n1 * inv, 0);
    co[9] = cscale(o12, -inv);
    co[10] = cconj(co[9]);
    const double dn1[5] = {0, 2 * r1 * r1 * w[1], 2 * r1 * r1 * w[2], 0, 0};
    const double dn2[5] = {0, 0, 0, 2 * r2 * r2 * w[3], 2 * r2 * r2 * w[4]};
    const cplx do1[5] = {cmake(0, 0), cmake(r1, 0), cmake(0, r1), cmake(0, 0), cmake(0, 0)};
    const cplx do2[5] = {cmake(0, 0), cmake(0, 0), cmake(0, 0), cmake(r2, 0), cmake(0, r2)};
    for (int c = 1; c < 5; ++c) {
      const double dden = dn1[c] + dn2[c];
      dc[7][c] = cmake(dn2[c] * inv - n2 * dden * inv2, 0);
      dc[8][c] = cmake(dn1[c] * inv - n1 * dden * inv2, 0);
      const cplx num = cadd(cmul(do1[c], o2), cmul(o1, do2[c]));
      dc[9][c] = cadd(cscale(num, -inv), cscale(o12, dden * inv2));
      dc[10][c] = cconj(dc[9][c]);
    }
  }
}

// Which physical-amplitude channels the coefficient of generator g depends on (bit c = channel c): the sparsity of dc
// above.  Used to skip gradient work that only a dead channel would consume.
__host__ __device__ inline unsigned qocvl_model_gen_channels(int model, int g) {
  if (model == QOCVL_MODEL_CHAIN) return g < 2 ? 1u << g : 0u;
  if (g < 2) return 0u;                                   // drift pieces
  if (g >= 7) return 0x1Eu;                               // projector weights: channels 1-4
  if (model == QOCVL_MODEL_TWO_QUBIT) return g == 2 ? 0x1u : (g <= 4 ? 0x6u : 0x18u);
  return 1u << (g - 2);                                   // state transfer: one channel per Hamiltonian piece
}
// Channels whose gradient the pulse map discards (abc -> wave -> cost path only): the two-qubit model forces its detuning
// channel to -1 (TQ:364, quirk q2), so d cost / d wave[:, 0] never reaches the parameters.
__host__ __device__ inline unsigned qocvl_model_dead_channels(int model) {
  return model == QOCVL_MODEL_TWO_QUBIT ? 0x1u : 0u;
}

// Sparse twin of qocvl_slice_norm below (same bound, O(nnz) instead of O(n^2 J)): column sums through
// the CSC views of the two block families.  *_colptr [n+1], *_ceid [nnz] (entry id in row order),
// *_egen / *_eval [nnz][kc].
__host__ __device__ inline double qocvl_slice_norm_csr(const ModelParams& mp, const cplx* co,
                                                      const int* u_colptr, const int* u_ceid, const int* u_egen,
                                                      const cplx* u_eval, int u_kc, const int* w_colptr,
                                                      const int* w_ceid, const int* w_egen, const cplx* w_eval,
                                                      int w_kc, const double* gcol, const double* mulmax,
                                                      const double* addmax, const unsigned char* colrole = nullptr,
                                                      int nroles = 0, double* role_nrm = nullptr) {
  const int n = mp.n, J = mp.J;
  double wgt[QOCVL_MAX_GEN];
  for (int j = 0; j < J; ++j)
    wgt[j] = sqrt(co[j].x * co[j].x + co[j].y * co[j].y) * mulmax[J + j] + addmax[j];
  double nrm = 0.0;
  for (int b = 0; b < n; ++b) {
    double s = 0.0;
    for (int j = 0; j < J; ++j) s += wgt[j] * gcol[j * n + b];
    for (int i = u_colptr[b]; i < u_colptr[b + 1]; ++i) {
      const int e = u_ceid[i];
      double re = 0, im = 0;
      for (int c = 0; c < u_kc; ++c) {
        const int g = u_egen[e * u_kc + c];
        if (g >= 0) {
          const cplx v = u_eval[e * u_kc + c];
          re += co[g].x * v.x - co[g].y * v.y; im += co[g].x * v.y + co[g].y * v.x;
        }
      }
      s += sqrt(re * re + im * im);
    }
    for (int i = w_colptr[b]; i < w_colptr[b + 1]; ++i) {
      const int e = w_ceid[i];
      double re = 0, im = 0;
      for (int c = 0; c < w_kc; ++c) {
        const int g = w_egen[e * w_kc + c];
        if (g >= 0) {
          const cplx v = w_eval[e * w_kc + c];
          re += co[g].x * v.x - co[g].y * v.y; im += co[g].x * v.y + co[g].y * v.x;
        }
      }
      s += sqrt(re * re + im * im);
    }
    nrm = fmax(nrm, s);
    // per-role bound (chain_duo.cu): the roles are the diagonal blocks of the propagated system, so the Taylor
    // remainder of a role is bounded by the columns of ITS basis states alone
    if (colrole)
      for (int r = 0; r < nroles; ++r)
        if (((colrole[b] | colrole[n + b]) >> r) & 1u) role_nrm[r] = fmax(role_nrm[r], s * mp.dt);
  }
  return nrm * mp.dt;
}

// Bound on |X|_1 of the 2n x 2n slice generator over every noise sample:
//   exact 1-norm of the nominal matrix (column b: sum_a |A_ab|; column n+b: sum_a |B_ab| + |A_ab|)
//   + entrywise bound of the per-sample deviation (|c_j| max|mul_j - 1| + max|add_j|) * colsum|G_j|.
// gdense: [J][2n][2n]; gcol: [J][n]; mulmax: [2][J] (max|mul|, max|mul-1|); addmax: [J].
__host__ __device__ inline double qocvl_slice_norm(const ModelParams& mp, const cplx* co, const cplx* gdense,
                                                  const double* gcol, const double* mulmax,
                                                  const double* addmax) {
  const int n = mp.n, J = mp.J, D = 2 * mp.n;
  double wgt[QOCVL_MAX_GEN];
  for (int j = 0; j < J; ++j)
    wgt[j] = sqrt(co[j].x * co[j].x + co[j].y * co[j].y) * mulmax[J + j] + addmax[j];
  double nrm = 0.0;
  for (int b = 0; b < n; ++b) {
    double s = 0.0;
    for (int j = 0; j < J; ++j) s += wgt[j] * gcol[j * n + b];
    for (int a = 0; a < n; ++a) {
      cplx ea = cmake(0, 0), eb = cmake(0, 0);
      for (int j = 0; j < J; ++j) {
        const cplx ga = gdense[((size_t)j * D + a) * D + b], gb = gdense[((size_t)j * D + a) * D + n + b];
        ea.x += co[j].x * ga.x - co[j].y * ga.y; ea.y += co[j].x * ga.y + co[j].y * ga.x;
        eb.x += co[j].x * gb.x - co[j].y * gb.y; eb.y += co[j].x * gb.y + co[j].y * gb.x;
      }
      s += sqrt(ea.x * ea.x + ea.y * ea.y) + sqrt(eb.x * eb.x + eb.y * eb.y);
    }
    nrm = fmax(nrm, s);
  }
  return nrm * mp.dt;
}
