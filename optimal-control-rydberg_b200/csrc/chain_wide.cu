// chain_wide.cu -- ensembles on LARGE Hilbert spaces (blockade chains, BASELINE.json configs 4-5): the same cost +
// gradient recurrences as chain_small.cu (see its header for the maths), organised for the one resource that bounds a
// sparse complex mat-vec on the SM: the 128 B / clk shared-memory pipe (every multiply-add needs a 16-byte gather).
//
//   * SPC noise samples per CTA in a SAMPLE-MINOR layout  t[vector][row][sample]  (complex, 16 B): the gather of one
//     row for the 8 samples of a CTA is one conflict-free 128-byte wavefront whatever the sparsity pattern, and the
//     column index / matrix value of an entry is loaded once for the 8 samples (one broadcast L1 access).
//   * no per-slice assembled copy of the matrix: X_k(s) = dt * sum_g c_g(k,s) G_g is applied generator by generator,
//     y = sum_g c_g (G_g t), with the STATIC generator values; the diagonal of X is folded into one register value
//     per owned row and slice.
//   * reverse sweep by generator as well: with u_g = G_g^dag tb_j gathered for the adjoint mat-vec anyway, the
//     gradient with respect to the coefficient of generator g is  sum_j <u_g , t_{j-1}> / j  -- a thread-local
//     product with the thread's OWN element of the forward term, so no per-entry Xbar accumulators and no second
//     gather.  One scalar per generator and sample is reduced per slice and contracted with d coef / d wave.
//   * forward Taylor terms t_0 .. t_{q-1} of every slice are streamed to HBM by the forward sweep and read back
//     (own elements only, straight into registers) by the reverse sweep; when the stream does not fit, only the
//     slice checkpoints are kept and the terms are recomputed per slice (QOCVL_WIDE_STORE=0 forces that).
//   * rows are sorted by their number of off-diagonal entries and dealt to the warps in groups of 32/SPC rows whose
//     entry lists are padded to a common length and stored interleaved: no intra-warp divergence, consecutive
//     addresses for the 4 rows of a warp.
// Results are deterministic (fixed reduction orders, no atomics).  The older one-CTA-per-sample kernel (chain_big.cu)
// stays as the fallback for problems this kernel does not take (several moving start kets, many generator blocks)
// and as its cross-check (QOCVL_WIDE=0).
#include "common.cuh"
#include <algorithm>
#include <numeric>

#define WIDE_THREADS 512
#define WIDE_MAXB 2     // generator blocks with off-diagonal entries
#define WIDE_MAXD 4     // generator blocks with diagonal entries
static constexpr int kWideUnroll = 4;   // entries of a row list in flight per thread
#define WIDE_NVP 2      // propagated vectors: one moving start ket q and p = W psi

struct WideArgs {
  int n, J, M, C, S, qcap, want_grad, store, ngroups, nb, nd;
  int ent_fwd, ent_rev, dreal;              // padded entries per direction; diagonals stored as real doubles
  double dt, inv_duration, w_infid, w_adiab, inv_fid_norm, inv_samples;
  cplx d_const;
  int bgen[WIDE_MAXB], bfam[WIDE_MAXB];   // generator and family (0 = diagonal blocks u, 1 = top-right block w)
  int buni[WIDE_MAXB];                    // every off-diagonal entry of the block has the same value bval
  cplx bval[WIDE_MAXB];
  int dgen[WIDE_MAXD], dfam[WIDE_MAXD];
  const int* grow;        // [ngroups * RW] row owned by (group, row slot), -1 = padding
  const int2* fwd_co;     // [nb][ngroups] {entries per row of the group, offset of the group's entry block}
  const int2* rev_co;
  const cplx* fwd_val;    // entry block of a group: [entry][RW]; row lists (col, G[row][col])
  const unsigned short* fwd_col;
  const cplx* rev_val;    // column lists (row, G[row][col]) -- conjugated in the kernel
  const unsigned short* rev_col;
  const cplx* dval;       // [nd][ngroups][RW] diagonal of each diagonal block, in owner order
  const double* dval_re;  // the same as real doubles (dreal)
  const cplx* coef;       // [M][J]
  const cplx* dcoef;      // [M][J][C]
  const int* qdeg;        // [M]
  const double* mul;      // [S][J]
  const cplx* add;        // [S][J]
  const cplx* starts;     // [1][n]
  const cplx* targets;    // [1][n]
  cplx* ring;             // [cta][(store ? M : 1) * qcap][NVP][n][SPC]
  cplx* ckpt;             // recompute mode: [cta][M][NVP][n][SPC]
  double* sample_out;     // [S][3]
  double* gw_partial;     // [cta][M][C]
};

__device__ __forceinline__ cplx ld_stream(const cplx* p) {
  const double2 v = __ldcs(reinterpret_cast<const double2*>(p));
  return v;
}
__device__ __forceinline__ void st_stream(cplx* p, cplx v) { __stcs(reinterpret_cast<double2*>(p), v); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int SPC, int ITEMS, int THREADS, int NB, int ND>
__global__ void __launch_bounds__(THREADS, 1) k_chain_wide(const WideArgs P) {
  constexpr int NVP = WIDE_NVP, NV = NVP - 1, VA = 0;   // vectors: q (index 0, also psi of the adiabaticity term), p
  constexpr int RW = 32 / SPC, NSLOT = NB + ND;   // compiled generator-block counts (off-diagonal, diagonal)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int s = lane % SPC, rw = lane / SPC;
  const int n = P.n, J = P.J, C = P.C, M = P.M;
  const int cta = blockIdx.x;
  const int nact = min(SPC, P.S - cta * SPC);            // live samples of this CTA (padding lanes repeat the last)
  const int nr = n + 1;                                  // row n of every vector stays zero (read by padding entries)
  const int VS = NVP * nr * SPC;                         // elements of one "all vectors" buffer
  const cplx czero = cmake(0, 0);

  // ---- shared memory
  cplx* carve = (cplx*)smem_raw;
  cplx* s_buf = carve; carve += 2 * VS;                  // Taylor term / cotangent term, double buffered
  cplx* s_cs = carve; carve += 2 * J * SPC;              // dt * (coef_kg * mul_sg + add_sg), double buffered by slice
  cplx* s_add = carve; carve += J * SPC;
  cplx* s_dco = carve; carve += J * C;
  cplx* s_scr = carve; carve += (size_t)(THREADS / 32) * NSLOT * SPC;
  cplx* s_ysum = carve; carve += NSLOT * SPC;
  cplx* s_dvc = carve; carve += P.dreal ? 0 : ND * P.ngroups * RW;     // diagonals of the diagonal blocks
  double* s_mul = (double*)carve;                        // [J][SPC]
  double* s_red = s_mul + J * SPC;                       // [nwarps][3][SPC] cost reductions
  double* s_dsum = s_red + (THREADS / 32) * 3 * SPC;   // [2][SPC]
  double* s_dvr = s_dsum + 2 * SPC;                      // diagonals as real doubles
  int2* s_fco = (int2*)(s_dvr + (P.dreal ? ND * P.ngroups * RW : 0));   // [NB][ngroups] {entries, offset}
  int2* s_rco = s_fco + NB * P.ngroups;
  unsigned short* s_fcol = (unsigned short*)(s_rco + NB * P.ngroups);   // column indices, [entry][RW] per group
  unsigned short* s_rcol = s_fcol + ((P.ent_fwd + 7) & ~7);
  for (int i = tid; i < P.nb * P.ngroups; i += blockDim.x) { s_fco[i] = P.fwd_co[i]; s_rco[i] = P.rev_co[i]; }
  for (int i = tid; i < P.ent_fwd; i += blockDim.x) s_fcol[i] = P.fwd_col[i];
  for (int i = tid; i < P.ent_rev; i += blockDim.x) s_rcol[i] = P.rev_col[i];
  for (int i = tid; i < P.nd * P.ngroups * RW; i += blockDim.x) {
    if (P.dreal) s_dvr[i] = P.dval_re[i]; else s_dvc[i] = P.dval[i];
  }
  auto diag_value = [&](int d, int g) {
    const int idx = (d * P.ngroups + g) * RW + rw;
    return P.dreal ? cmake(s_dvr[idx], 0.0) : s_dvc[idx];
  };

  // ---- owned rows
  int row[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int g = i * nwarps + warp;
    row[i] = (g < P.ngroups) ? P.grow[g * RW + rw] : -1;
  }
  auto at = [&](int v, int r) { return (v * nr + r) * SPC + s; };
  for (int i = tid; i < 2 * NVP * SPC; i += blockDim.x)
    s_buf[(i / (NVP * SPC)) * VS + (((i / SPC) % NVP) * nr + n) * SPC + (i % SPC)] = czero;

  for (int i = tid; i < J * SPC; i += blockDim.x) {
    const int g = i / SPC, ss = i % SPC, sm = cta * SPC + min(ss, nact - 1);
    s_mul[i] = P.mul[(size_t)sm * J + g];
    s_add[i] = P.add[(size_t)sm * J + g];
  }
  __syncthreads();
  auto load_cs = [&](int k, int b) {
    for (int i = tid; i < J * SPC; i += blockDim.x) {
      const int g = i / SPC;
      s_cs[b * J * SPC + i] = cscale(cadd(cscale(P.coef[(size_t)k * J + g], s_mul[i]), s_add[i]), P.dt);
    }
  };
  // coefficients of this thread's sample for the generator blocks of the slice, in registers
  cplx csb[NB], csd[ND];
  auto coefficients = [&](int b) {
#pragma unroll
    for (int x = 0; x < NB; ++x) csb[x] = (x < P.nb) ? s_cs[(b * J + P.bgen[x]) * SPC + s] : czero;
#pragma unroll
    for (int x = 0; x < ND; ++x) csd[x] = (x < P.nd) ? s_cs[(b * J + P.dgen[x]) * SPC + s] : czero;
  };

  cplx z[ITEMS][NVP];      // forward: state; reverse: cotangent lambda
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    z[i][0] = (row[i] >= 0) ? P.starts[row[i]] : czero;
    z[i][NV] = czero;
  }
  // one forward Taylor step of one owned row:  y = (X t)[row] / j
  auto forward_item = [&](int i, const cplx* t, double invj, cplx* y) {
    const int r = row[i], g = i * nwarps + warp;
    const cplx tq = t[at(VA, r)], tp = t[at(NV, r)];
    cplx accq = czero, accp = czero;
#pragma unroll
    for (int bb = 0; bb < NB; ++bb) {
      if (bb < P.nb) {
        const int2 co = s_fco[bb * P.ngroups + g];
        const unsigned short* ec = s_fcol + co.y + rw;
        const bool both = (P.bfam[bb] == 0);     // u blocks act on q and p, w blocks map q into p
        cplx sq = czero, sp = czero;
        if (P.buni[bb]) {                        // one common value: plain sums, scaled once
#pragma unroll kWideUnroll
          for (int e = 0; e < co.x; ++e) {
            const int col = ec[e * RW];
            sq = cadd(sq, t[at(VA, col)]);
            if (both) sp = cadd(sp, t[at(NV, col)]);
          }
          sq = cmul(P.bval[bb], sq);
          sp = cmul(P.bval[bb], sp);
        } else {
          const cplx* ev = P.fwd_val + co.y + rw;
#pragma unroll kWideUnroll
          for (int e = 0; e < co.x; ++e) {
            const cplx val = __ldg(&ev[e * RW]);
            const int col = ec[e * RW];
            sq = cfma(val, t[at(VA, col)], sq);
            if (both) sp = cfma(val, t[at(NV, col)], sp);
          }
        }
        if (both) { accq = cfma(csb[bb], sq, accq); accp = cfma(csb[bb], sp, accp); }
        else accp = cfma(csb[bb], sq, accp);
      }
    }
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      if (d < P.nd) {
        const cplx c = cmul(csd[d], diag_value(d, g));
        if (P.dfam[d] == 0) { accq = cfma(c, tq, accq); accp = cfma(c, tp, accp); }
        else accp = cfma(c, tq, accp);
      }
    }
    y[0] = cscale(accq, invj);
    y[NV] = cscale(accp, invj);
  };

  const size_t ring_slices = P.store ? (size_t)M : 1;
  cplx* ring = P.ring + (size_t)cta * ring_slices * P.qcap * VS;
  cplx* ckpt = P.store ? nullptr : P.ckpt + (size_t)cta * M * VS;

  // =================================== forward sweep ===================================
  load_cs(0, 0);
  __syncthreads();
  for (int k = 0; k < M; ++k) {
    const int q = P.qdeg[k], b = k & 1;
    if (k + 1 < M) load_cs(k + 1, b ^ 1);
    coefficients(b);
    cplx* t0 = (P.store ? ring + (size_t)k * P.qcap * VS : ckpt + (size_t)k * VS);
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      if (row[i] < 0) continue;
#pragma unroll
      for (int v = 0; v < NVP; ++v) {
        s_buf[at(v, row[i])] = z[i][v];
        if (P.want_grad) st_stream(&t0[at(v, row[i])], z[i][v]);
      }
    }
    __syncthreads();
    int cur = 0;
    for (int j = 1; j <= q; ++j) {
      const cplx* t = s_buf + cur * VS;
      cplx* y = s_buf + (cur ^ 1) * VS;
      const double invj = 1.0 / (double)j;
      cplx* rj = ring + ((size_t)k * P.qcap + j) * VS;
      const bool keep = (j < q);
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        if (row[i] < 0) continue;
        cplx yy[NVP];
        forward_item(i, t, invj, yy);
#pragma unroll
        for (int v = 0; v < NVP; ++v) {
          z[i][v] = cadd(z[i][v], yy[v]);
          if (keep) {
            y[at(v, row[i])] = yy[v];
            if (P.want_grad && P.store) st_stream(&rj[at(v, row[i])], yy[v]);
          }
        }
      }
      __syncthreads();
      cur ^= 1;
    }
  }

  // =================================== cost ============================================
  {
    double red[3] = {0, 0, 0};   // Re d, Im d, Re q^dag p
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      if (row[i] < 0) continue;
      const cplx y = P.targets[row[i]], zq = z[i][0], zp = z[i][NV];
      red[0] += y.x * zq.x + y.y * zq.y;
      red[1] += y.x * zq.y - y.y * zq.x;
      red[2] += zq.x * zp.x + zq.y * zp.y;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      for (int off = SPC; off < 32; off <<= 1) red[c] += __shfl_xor_sync(0xffffffffu, red[c], off);
      if (rw == 0) s_red[(warp * 3 + c) * SPC + s] = red[c];
    }
    __syncthreads();
    if (tid < SPC) {
      double a[3] = {0, 0, 0};
      for (int w = 0; w < nwarps; ++w)
        for (int c = 0; c < 3; ++c) a[c] += s_red[(w * 3 + c) * SPC + tid];
      const double dx = a[0] + P.d_const.x, dy = a[1] + P.d_const.y;
      s_dsum[tid] = dx; s_dsum[SPC + tid] = dy;
      if (tid < nact) {
        const double infid = 1.0 - (dx * dx + dy * dy) * P.inv_fid_norm;
        const double adiab = 1.0 - a[2] * P.inv_duration;
        double* o = P.sample_out + (size_t)(cta * SPC + tid) * 3;
        o[0] = P.w_infid * infid + P.w_adiab * adiab;
        o[1] = infid;
        o[2] = adiab;
      }
    }
    __syncthreads();
  }
  if (!P.want_grad) return;

  // terminal cotangents (scaled by 1 / global sample count):  lam_q = cf * d * y + ca * p,  lam_p = ca * q
  {
    const double ca = -0.5 * P.w_adiab * P.inv_duration * P.inv_samples;
    const double cf = -P.w_infid * P.inv_fid_norm * P.inv_samples;
    const cplx dsum = cmake(s_dsum[s], s_dsum[SPC + s]);
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      if (row[i] < 0) continue;
      const cplx zq = z[i][0], zp = z[i][NV];
      z[i][0] = cadd(cscale(cmul(dsum, P.targets[row[i]]), cf), cscale(zp, ca));
      z[i][NV] = cscale(zq, ca);
    }
  }

  // =================================== reverse sweep ===================================
  load_cs(M - 1, (M - 1) & 1);
  __syncthreads();
  for (int k = M - 1; k >= 0; --k) {
    const int q = P.qdeg[k], b = k & 1, qprev = k > 0 ? P.qdeg[k - 1] : 0;
    if (k > 0) load_cs(k - 1, b ^ 1);
    coefficients(b);
    cplx* rk = ring + (P.store ? (size_t)k * P.qcap * VS : 0);
    int cur = 0;
    if (!P.store) {
      // ---- recompute the forward terms t_0 .. t_{q-1} of this slice into the ring (own elements)
      const cplx* ck = ckpt + (size_t)k * VS;
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        if (row[i] < 0) continue;
#pragma unroll
        for (int v = 0; v < NVP; ++v) {
          const cplx zz = ld_stream(&ck[at(v, row[i])]);
          s_buf[at(v, row[i])] = zz;
          st_stream(&rk[at(v, row[i])], zz);
        }
      }
      __syncthreads();
      for (int j = 1; j < q; ++j) {
        const cplx* t = s_buf + cur * VS;
        cplx* y = s_buf + (cur ^ 1) * VS;
        cplx* rj = rk + (size_t)j * VS;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
          if (row[i] < 0) continue;
          cplx yy[NVP];
          forward_item(i, t, 1.0 / (double)j, yy);
#pragma unroll
          for (int v = 0; v < NVP; ++v) { y[at(v, row[i])] = yy[v]; st_stream(&rj[at(v, row[i])], yy[v]); }
        }
        __syncthreads();
        cur ^= 1;
      }
    }
    // ---- descending adjoint recurrence  tb_{j-1} = lam + X^dag tb_j / j
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      if (row[i] < 0) continue;
#pragma unroll
      for (int v = 0; v < NVP; ++v) s_buf[cur * VS + at(v, row[i])] = z[i][v];
    }
    __syncthreads();
    // (the previous slice's readers of s_dco / s_ysum are past this barrier)
    for (int i = tid; i < J * C; i += blockDim.x) s_dco[i] = P.dcoef[(size_t)k * J * C + i];
    cplx yb[NB], yd[ND];   // Ybar of every generator block: sum_j <G^dag tb_j , t_{j-1}> / j
#pragma unroll
    for (int bb = 0; bb < NB; ++bb) yb[bb] = czero;
#pragma unroll
    for (int d = 0; d < ND; ++d) yd[d] = czero;
    for (int j = q; j >= 1; --j) {
      const cplx* t = s_buf + cur * VS;
      cplx* y = s_buf + (cur ^ 1) * VS;
      const double invj = 1.0 / (double)j;
      const cplx* rj = rk + (size_t)(j - 1) * VS;
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        const int c = row[i], g = i * nwarps + warp;
        if (c < 0) continue;
        const cplx fq_raw = ld_stream(&rj[at(VA, c)]);             // own elements of t_{j-1}, used after the gathers
        const cplx fp_raw = ld_stream(&rj[at(NV, c)]);
        if (j > 2) {                                                // pull the terms of step j-2 from HBM into L2
          prefetch_l2(&rj[at(VA, c) - 2 * VS]);
          prefetch_l2(&rj[at(NV, c) - 2 * VS]);
        } else if (P.store && k > 0 && qprev - 3 + j >= 0) {       // ... or the top terms of the next slice to process
          const cplx* nx = rk - (size_t)P.qcap * VS + (size_t)(qprev - 3 + j) * VS;
          prefetch_l2(&nx[at(VA, c)]);
          prefetch_l2(&nx[at(NV, c)]);
        }
        const cplx bq = t[at(VA, c)], bp = t[at(NV, c)];           // own elements of tb_j
        cplx accq = czero, accp = czero;
        cplx sqb[NB], spb[NB];                                      // (G^dag tb_q)[c], (G^dag tb_p)[c] per block
#pragma unroll
        for (int bb = 0; bb < NB; ++bb) {
          sqb[bb] = czero; spb[bb] = czero;
          if (bb < P.nb) {
            const int2 co = s_rco[bb * P.ngroups + g];
            const unsigned short* ec = s_rcol + co.y + rw;
            const bool both = (P.bfam[bb] == 0);   // w blocks: (G_w^dag tb_p)[c] feeds q
            cplx sq = czero, sp = czero;
            if (P.buni[bb]) {
#pragma unroll kWideUnroll
              for (int e = 0; e < co.x; ++e) {
                const int a = ec[e * RW];
                sp = cadd(sp, t[at(NV, a)]);
                if (both) sq = cadd(sq, t[at(VA, a)]);
              }
              const cplx vc = cconj(P.bval[bb]);
              sq = cmul(vc, sq);
              sp = cmul(vc, sp);
            } else {
              const cplx* ev = P.rev_val + co.y + rw;
#pragma unroll kWideUnroll
              for (int e = 0; e < co.x; ++e) {
                const cplx val = __ldg(&ev[e * RW]);
                const int a = ec[e * RW];
                sp = cfma_conj(val, t[at(NV, a)], sp);
                if (both) sq = cfma_conj(val, t[at(VA, a)], sq);
              }
            }
            if (both) { accq = cfma_conj(csb[bb], sq, accq); accp = cfma_conj(csb[bb], sp, accp); }
            else accq = cfma_conj(csb[bb], sp, accq);
            sqb[bb] = sq; spb[bb] = sp;
          }
        }
        const cplx fq = cscale(fq_raw, invj), fp = cscale(fp_raw, invj);
#pragma unroll
        for (int bb = 0; bb < NB; ++bb) {
          if (bb < P.nb) {
            if (P.bfam[bb] == 0) yb[bb] = cfma_conj(spb[bb], fp, cfma_conj(sqb[bb], fq, yb[bb]));
            else yb[bb] = cfma_conj(spb[bb], fq, yb[bb]);
          }
        }
        // diagonal blocks: Ybar_d += dval * (conj(tb_q) f_q + conj(tb_p) f_p)  (u)   |   dval * conj(tb_p) f_q  (w)
        const cplx wsum = cfma_conj(bp, fp, cfma_conj(bq, fq, czero));
        const cplx wpq = cfma_conj(bp, fq, czero);
#pragma unroll
        for (int d = 0; d < ND; ++d) {
          if (d < P.nd) {
            const cplx dv = diag_value(d, g);
            const cplx cc = cmul(csd[d], dv);
            if (P.dfam[d] == 0) {
              accq = cfma_conj(cc, bq, accq); accp = cfma_conj(cc, bp, accp);
              yd[d] = cfma(dv, wsum, yd[d]);
            } else {
              accq = cfma_conj(cc, bp, accq);
              yd[d] = cfma(dv, wpq, yd[d]);
            }
          }
        }
        const cplx nq = cadd(z[i][0], cscale(accq, invj)), np = cadd(z[i][NV], cscale(accp, invj));
        if (j > 1) { y[at(VA, c)] = nq; y[at(NV, c)] = np; }
        else { z[i][0] = nq; z[i][NV] = np; }
      }
      __syncthreads();
      cur ^= 1;
    }
    // ---- d cost / d wave[k][c] = sum_s sum_g Re(2 * mul_sg * Ybar_g(s) * dcoef[k][g][c])
    auto warp_partial = [&](cplx v, int x) {
      for (int off = SPC; off < 32; off <<= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, off);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, off);
      }
      if (rw == 0) s_scr[(warp * NSLOT + x) * SPC + s] = v;
    };
#pragma unroll
    for (int bb = 0; bb < NB; ++bb) warp_partial(yb[bb], bb);
#pragma unroll
    for (int dd = 0; dd < ND; ++dd) warp_partial(yd[dd], NB + dd);
    __syncthreads();
    if (tid < NSLOT * SPC) {
      cplx a = czero;
      for (int w = 0; w < nwarps; ++w) a = cadd(a, s_scr[w * NSLOT * SPC + tid]);
      s_ysum[tid] = a;
    }
    __syncthreads();
    if (tid < C) {
      double gw = 0.0;
      for (int x = 0; x < NSLOT; ++x) {
        int g;
        if (x < NB) { if (x >= P.nb) continue; g = P.bgen[x]; }
        else { if (x - NB >= P.nd) continue; g = P.dgen[x - NB]; }
        const cplx dc = s_dco[g * C + tid];
        for (int ss = 0; ss < nact; ++ss) {
          const cplx yv = cscale(s_ysum[x * SPC + ss], 2.0 * P.dt * s_mul[g * SPC + ss]);
          gw += yv.x * dc.x - yv.y * dc.y;
        }
      }
      P.gw_partial[((size_t)cta * M + k) * C + tid] = gw;
    }
    // s_dco / s_scr / s_ysum are rewritten only after a barrier of the next slice
  }
}

// ---------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------
template <typename T>
static int wide_upload(T** dst, const std::vector<T>& src) {
  if (*dst) { cudaFree(*dst); *dst = nullptr; }
  QCUDA(cudaMalloc((void**)dst, std::max<size_t>(1, src.size()) * sizeof(T)));
  if (!src.empty()) QCUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
  return QOCVL_OK;
}

// dynamic shared memory of one CTA: vectors, per-slice coefficients, reduction scratch and the generator tables
static size_t wide_smem_bytes(int n, int J, int C, int spc, int threads, int nb_c, int nd_c, int ngroups,
                              size_t ent_fwd, size_t ent_rev, bool dreal) {
  const int rw = 32 / spc, nslot = nb_c + nd_c;
  const size_t vs = (size_t)WIDE_NVP * n * spc;
  size_t cplx_count = 2 * vs + 2 * (size_t)J * spc + (size_t)J * spc + (size_t)J * C +
                      (size_t)(threads / 32) * nslot * spc + (size_t)nslot * spc +
                      (dreal ? 0 : (size_t)nd_c * ngroups * rw);
  size_t dbl_count = (size_t)J * spc + (size_t)(threads / 32) * 3 * spc + 2 * (size_t)spc +
                     (dreal ? (size_t)nd_c * ngroups * rw : 0);
  return cplx_count * sizeof(cplx) + dbl_count * sizeof(double) + 2 * (size_t)nb_c * ngroups * sizeof(int2) +
         (((ent_fwd + 7) & ~(size_t)7) + ((ent_rev + 7) & ~(size_t)7)) * sizeof(unsigned short);
}

typedef void (*wide_kernel_t)(const WideArgs);
// The register file bounds rows per thread (state, diagonal and cotangent of every owned row live in registers):
// 512 threads x 128 registers up to 4 rows, 384 x 168 up to 8, 256 x 255 up to 14.
static int wide_threads_for(int spc, int ngroups) {
  if (const char* e = getenv("QOCVL_WIDE_THREADS")) {
    const int t = atoi(e);
    if (t == 256 || t == 384 || t == 512) return t;
  }
  if (spc != 8) return 512;
  if (ngroups <= 4 * 16) return 512;
  if (ngroups <= 8 * 12) return 384;
  return 256;
}
template <int NB, int ND>
static wide_kernel_t wide_pick_blocks(int spc, int items, int threads) {
  if (spc == 8 && threads == 512) {
    if (items <= 2) return k_chain_wide<8, 2, 512, NB, ND>;
    if (items <= 4) return k_chain_wide<8, 4, 512, NB, ND>;
    return nullptr;
  }
  if (spc == 8 && threads == 384) {
    if (items <= 8) return k_chain_wide<8, 8, 384, NB, ND>;
    return nullptr;
  }
  if (spc == 8 && threads == 256) {
    if (items <= 14) return k_chain_wide<8, 14, 256, NB, ND>;
    return nullptr;
  }
  if (spc == 1) {
    if (items <= 1) return k_chain_wide<1, 1, 512, NB, ND>;
    if (items <= 2) return k_chain_wide<1, 2, 512, NB, ND>;
  }
  return nullptr;
}
static wide_kernel_t wide_pick(int spc, int items, int threads, int nb, int nd) {
  if (nb <= 1 && nd <= 3) return wide_pick_blocks<1, 3>(spc, items, threads);
  if (nb <= WIDE_MAXB && nd <= WIDE_MAXD) return wide_pick_blocks<WIDE_MAXB, WIDE_MAXD>(spc, items, threads);
  return nullptr;
}

void WideLayout::release() {
  void* ptrs[] = {grow, fwd_co, rev_co, fwd_val, fwd_col, rev_val, rev_col, dval, dval_re};
  for (void* q : ptrs) if (q) cudaFree(q);
  *this = WideLayout();
}

// Build the tables of the wide kernel.  QOCVL_ERR_UNSUPPORTED when the problem does not fit it (the caller then
// falls back to the one-CTA-per-sample kernel).
int qocvl_wide_configure(qocvl_plan* p) {
  const qocvl_desc& d = p->d;
  const int n = d.n, J = d.n_gen, S = p->n_samples, M = d.n_slices;
  p->use_wide = false;
  if (p->n_live != 1 || p->adiab_live != 0) return QOCVL_ERR_UNSUPPORTED;
  WideLayout& L = p->lwide;
  L.release();
  // ---- generator blocks: off-diagonal part and diagonal part of every (generator, family)
  struct Blk { int g, fam; };
  std::vector<Blk> off_blk, diag_blk;
  auto G = [&](int fam, int g, int a, int b) {
    return (fam == 0 ? p->h_gu : p->h_gw)[((size_t)g * n + a) * n + b];
  };
  for (int fam = 0; fam < 2; ++fam)
    for (int g = 0; g < J; ++g) {
      bool has_off = false, has_diag = false;
      for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b)
          if (std::abs(G(fam, g, a, b)) > 0.0) (a == b ? has_diag : has_off) = true;
      if (has_off) off_blk.push_back({g, fam});
      if (has_diag) diag_blk.push_back({g, fam});
    }
  if ((int)off_blk.size() > WIDE_MAXB || (int)diag_blk.size() > WIDE_MAXD) return QOCVL_ERR_UNSUPPORTED;
  const int nb = (int)off_blk.size(), nd = (int)diag_blk.size();
  const int nb_c = (nb <= 1 && nd <= 3) ? 1 : WIDE_MAXB, nd_c = (nb <= 1 && nd <= 3) ? 3 : WIDE_MAXD;   // compiled counts
  // a block whose off-diagonal entries all carry one value needs no value table
  std::vector<int> buni(nb, 1);
  std::vector<std::complex<double>> bval(nb, 0.0);
  std::vector<int> degree(n, 0);
  double offdiag_macs = 0;
  for (int bi = 0; bi < nb; ++bi) {
    bool first = true;
    long long cnt = 0;
    for (int a = 0; a < n; ++a)
      for (int c = 0; c < n; ++c) {
        const std::complex<double> v = G(off_blk[bi].fam, off_blk[bi].g, a, c);
        if (a == c || std::abs(v) == 0.0) continue;
        if (first) { bval[bi] = v; first = false; }
        else if (v != bval[bi]) buni[bi] = 0;
        degree[a]++; degree[c]++; ++cnt;
      }
    offdiag_macs += (off_blk[bi].fam == 0 ? WIDE_NVP : 1) * (double)cnt;
  }
  bool dreal = true;
  for (int di = 0; di < nd; ++di)
    for (int a = 0; a < n; ++a) dreal = dreal && G(diag_blk[di].fam, diag_blk[di].g, a, a).imag() == 0.0;
  // ---- rows sorted by off-diagonal work (row + column entries), dealt to the groups in that order
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return degree[x] > degree[y]; });
  // ---- tables for a given number of samples per CTA: for every (block, group) the row lists (forward) / column
  //      lists (reverse), padded to the longest list of the group (index = own row, value 0) and stored
  //      [entry][row slot]; diagonals in owner order
  struct Tables {
    int rw = 0, ngroups = 0;
    std::vector<int> grow;
    std::vector<int2> co[2];
    std::vector<cplx> val[2];
    std::vector<unsigned short> col[2];
    std::vector<cplx> dval;
    std::vector<double> dval_re;
  };
  auto build = [&](int spc_) {
    Tables T;
    T.rw = 32 / spc_; T.ngroups = (n + T.rw - 1) / T.rw;
    T.grow.assign((size_t)T.ngroups * T.rw, -1);
    for (int i = 0; i < n; ++i) T.grow[i] = order[i];
    for (int dir = 0; dir < 2; ++dir) {
      T.co[dir].resize((size_t)std::max(1, nb) * T.ngroups);
      for (int bi = 0; bi < nb; ++bi)
        for (int g = 0; g < T.ngroups; ++g) {
          std::vector<std::vector<std::pair<int, std::complex<double>>>> lists(T.rw);
          size_t longest = 0;
          for (int r = 0; r < T.rw; ++r) {
            const int own = T.grow[(size_t)g * T.rw + r];
            if (own < 0) continue;
            for (int o = 0; o < n; ++o) {
              if (o == own) continue;
              const std::complex<double> v = dir == 0 ? G(off_blk[bi].fam, off_blk[bi].g, own, o)
                                                      : G(off_blk[bi].fam, off_blk[bi].g, o, own);
              if (std::abs(v) > 0.0) lists[r].push_back({o, v});
            }
            longest = std::max(longest, lists[r].size());
          }
          T.co[dir][(size_t)bi * T.ngroups + g] = make_int2((int)longest, (int)T.col[dir].size());
          for (size_t e = 0; e < longest; ++e)
            for (int r = 0; r < T.rw; ++r) {
              const int own = std::max(0, T.grow[(size_t)g * T.rw + r]);
              if (e < lists[r].size()) {
                T.val[dir].push_back(cmake(lists[r][e].second.real(), lists[r][e].second.imag()));
                T.col[dir].push_back((unsigned short)lists[r][e].first);
              } else {   // padding: a uniform-valued block adds the zero row n (kept zero in both buffers)
                T.val[dir].push_back(cmake(0, 0));
                T.col[dir].push_back((unsigned short)(buni[bi] ? n : own));
              }
            }
        }
    }
    T.dval.assign((size_t)std::max(1, nd) * T.ngroups * T.rw, cmake(0, 0));
    T.dval_re.assign(T.dval.size(), 0.0);
    for (int di = 0; di < nd; ++di)
      for (int i = 0; i < T.ngroups * T.rw; ++i) {
        if (T.grow[i] < 0) continue;
        const std::complex<double> v = G(diag_blk[di].fam, diag_blk[di].g, T.grow[i], T.grow[i]);
        T.dval[(size_t)di * T.ngroups * T.rw + i] = cmake(v.real(), v.imag());
        T.dval_re[(size_t)di * T.ngroups * T.rw + i] = v.real();
      }
    return T;
  };
  // ---- samples per CTA (8 when the CTA's vectors and tables fit shared memory), warps, rows per thread
  if (n + 1 > 65535) return QOCVL_ERR_UNSUPPORTED;
  int spc = (S >= 5) ? 8 : 1;
  if (const char* e = getenv("QOCVL_WIDE_SPC")) spc = (atoi(e) == 8) ? 8 : 1;
  Tables T;
  int max_threads = 0, nwarps = 0, items = 0;
  wide_kernel_t kern = nullptr;
  size_t smem = 0;
  for (;;) {
    T = build(spc);
    max_threads = wide_threads_for(spc, T.ngroups);
    nwarps = std::min(max_threads / 32, T.ngroups);
    items = (T.ngroups + nwarps - 1) / nwarps;
    nwarps = (T.ngroups + items - 1) / items;              // fewest warps that still need `items` passes
    kern = wide_pick(spc, items, max_threads, nb, nd);
    // (n + 1 rows per vector: row n is the zero row the padding entries of uniform-valued blocks read)
    smem = wide_smem_bytes(n + 1, J, d.n_chan, spc, max_threads, nb_c, nd_c, T.ngroups, T.col[0].size(),
                           T.col[1].size(), dreal);
    if (kern && smem <= 227 * 1024) break;
    if (spc == 1) return QOCVL_ERR_UNSUPPORTED;
    spc = 1;
  }
  const int ngroups = T.ngroups;
  int rc;
  if ((rc = wide_upload(&L.grow, T.grow))) return rc;
  if ((rc = wide_upload(&L.fwd_co, T.co[0]))) return rc;
  if ((rc = wide_upload(&L.rev_co, T.co[1]))) return rc;
  if ((rc = wide_upload(&L.fwd_val, T.val[0]))) return rc;
  if ((rc = wide_upload(&L.fwd_col, T.col[0]))) return rc;
  if ((rc = wide_upload(&L.rev_val, T.val[1]))) return rc;
  if ((rc = wide_upload(&L.rev_col, T.col[1]))) return rc;
  if ((rc = wide_upload(&L.dval, T.dval))) return rc;
  if ((rc = wide_upload(&L.dval_re, T.dval_re))) return rc;
  L.ent_fwd = (int)T.col[0].size(); L.ent_rev = (int)T.col[1].size(); L.dreal = dreal;
  for (int i = 0; i < nb; ++i) { L.buni[i] = buni[i]; L.bval[i] = cmake(bval[i].real(), bval[i].imag()); }
  L.macs = offdiag_macs + 3.0 * n;   // useful complex multiply-adds of one Taylor step (+ du*tq, du*tp, dw*tq)
  L.padded_entries = (long long)T.col[0].size();
  L.spc = spc; L.items = items; L.max_threads = max_threads; L.ngroups = ngroups; L.nb = nb; L.nd = nd;
  for (int i = 0; i < nb; ++i) { L.bgen[i] = off_blk[i].g; L.bfam[i] = off_blk[i].fam; }
  for (int i = 0; i < nd; ++i) { L.dgen[i] = diag_blk[i].g; L.dfam[i] = diag_blk[i].fam; }
  // ---- geometry and buffers
  const int ctas = (S + spc - 1) / spc;
  QCUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  p->chain_smem = smem;
  p->threads = nwarps * 32;
  p->blocks = ctas;
  p->spb = spc; p->groups = 1;
  p->n_parts = (size_t)ctas;
  const size_t vs = (size_t)WIDE_NVP * (n + 1) * spc * sizeof(cplx);
  const size_t stream_bytes = (size_t)ctas * M * p->qcap * vs;
  size_t free_b = 0, total_b = 0;
  QCUDA(cudaMemGetInfo(&free_b, &total_b));
  bool store = stream_bytes <= (size_t)(0.6 * (double)free_b);
  if (const char* e = getenv("QOCVL_WIDE_STORE")) store = store && atoi(e) != 0;
  if (p->d_ckpt) { cudaFree(p->d_ckpt); p->d_ckpt = nullptr; p->bytes -= p->ckpt_bytes; p->ckpt_bytes = 0; }
  if (p->d_tring) { cudaFree(p->d_tring); p->d_tring = nullptr; p->bytes -= p->tring_bytes; p->tring_bytes = 0; }
  p->tring_bytes = store ? stream_bytes : (size_t)ctas * p->qcap * vs;
  QCUDA(cudaMalloc((void**)&p->d_tring, p->tring_bytes));
  if (!store) {
    p->ckpt_bytes = (size_t)ctas * M * vs;
    QCUDA(cudaMalloc((void**)&p->d_ckpt, p->ckpt_bytes));
  }
  p->bytes += p->ckpt_bytes + p->tring_bytes;
  L.store = store;
  p->use_wide = true;
  return QOCVL_OK;
}

int qocvl_wide_launch(qocvl_plan* p, bool want_grad, cudaStream_t st) {
  const qocvl_desc& d = p->d;
  const WideLayout& L = p->lwide;
  WideArgs a;
  memset(&a, 0, sizeof(a));
  a.n = d.n; a.J = d.n_gen; a.M = d.n_slices; a.C = d.n_chan; a.S = p->n_samples;
  a.qcap = p->qcap; a.want_grad = want_grad ? 1 : 0; a.store = L.store ? 1 : 0;
  a.ngroups = L.ngroups; a.nb = L.nb; a.nd = L.nd;
  a.ent_fwd = L.ent_fwd; a.ent_rev = L.ent_rev; a.dreal = L.dreal ? 1 : 0;
  a.dt = d.dt; a.inv_duration = 1.0 / d.duration; a.w_infid = d.w_infid; a.w_adiab = d.w_adiab;
  a.inv_fid_norm = 1.0 / d.fid_norm; a.inv_samples = 1.0 / (double)p->global_samples;
  a.d_const = p->d_const;
  for (int i = 0; i < WIDE_MAXB; ++i) { a.bgen[i] = L.bgen[i]; a.bfam[i] = L.bfam[i]; a.buni[i] = L.buni[i]; a.bval[i] = L.bval[i]; }
  for (int i = 0; i < WIDE_MAXD; ++i) { a.dgen[i] = L.dgen[i]; a.dfam[i] = L.dfam[i]; }
  a.grow = L.grow; a.fwd_co = L.fwd_co; a.rev_co = L.rev_co;
  a.fwd_val = L.fwd_val; a.fwd_col = L.fwd_col; a.rev_val = L.rev_val; a.rev_col = L.rev_col; a.dval = L.dval; a.dval_re = L.dval_re;
  a.coef = p->d_coef; a.dcoef = p->d_dcoef; a.qdeg = p->d_qdeg; a.mul = p->d_mul; a.add = p->d_add;
  a.starts = p->d_starts; a.targets = p->d_targets;
  a.ring = p->d_tring; a.ckpt = p->d_ckpt;
  a.sample_out = p->d_sample_out; a.gw_partial = p->d_gw_partial;
  wide_kernel_t kern = wide_pick(L.spc, L.items, L.max_threads, L.nb, L.nd);
  kern<<<p->blocks, p->threads, p->chain_smem, st>>>(a);
  p->launches++;
  QCUDA(cudaGetLastError());
  return QOCVL_OK;
}
