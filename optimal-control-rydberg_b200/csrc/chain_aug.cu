// chain_aug.cu -- the ensemble hot kernel on the REDUCED system.
//
// Same numbers as chain_small.cu (ST:247-302 / TQ:460-567 and the reverse pass of jax.value_and_grad,
// TQ:592), but the host first shrinks the problem with two exact, engine-verified reductions:
//   1. reachability: a start ket only ever populates the basis states connected to it through the union
//      sparsity pattern of the generators (CZ: U|10> lives on {|10>,|i0>,|r0>}, U|11> on the 9 states of
//      {1,i,r} x {1,i,r}; |00> is inert).  Everything else is a structural zero that the lane-per-row kernel
//      of chain_small.cu multiplies anyway.  The adjoint only needs the same rows: Xbar is contracted on
//      the pattern restricted to the columns the forward vector populates, and that set is closed under X.
//   2. orbit folding: when the declared basis permutation (qocvl_set_symmetry, verified against every
//      generator) leaves a start ket invariant, the vector stays constant on the permutation's orbits; only
//      one representative per orbit is kept and the columns of X are summed over each orbit
//      (X' = R X E, exact because X commutes with the permutation).  CZ: the 9 states of U|11> -> 6.
// The vectors of all sectors, and p = W psi with its coupling block B, are then stacked into ONE sparse
// system of N rows (CZ: 3 + 6 + 3 = 12 instead of 3 x 16 = 48 vector elements per sample):
//     Z = [q_0 | q_1 | ... | p],   Xaug = blockdiag(A'_0, A'_1, ..., A'_p) + (p <- q_va block B').
// Kernel mapping: one noise sample per group of LPG >= N lanes, lane r owns row r of Xaug (diagonal + <= WMAX
// off-diagonals) and element r of Z.  A matrix entry of sample s is A_e(s) + sum_i m_s[class_i] * P_i(k): generators
// are grouped into classes by their per-sample multiplier column, P_i(k) comes from per-slice tables that
// k_aug_tables fills once per evaluation, A_e is the additive-noise constant.  Off-row elements are gathered through
// a double-buffered shared-memory exchange guarded by __syncwarp only (no CTA barrier in the sweeps).  The forward
// sweep streams every Taylor term to HBM as one dense stream; the reverse sweep reads the stream backwards through
// a cp.async ring and emits per-warp sums of m_s * Xbar_e, which k_aug_reduce adds up in a fixed order and contracts
// with d coef / d wave once per slice.  (STORE = 0: only checkpoints go to HBM and the terms are recomputed.)
#include "chain_aug.cuh"
#include <algorithm>
#include <cmath>
#include <map>
#include <type_traits>

typedef std::complex<double> zc;

#ifndef QOCVL_RING
#define QOCVL_RING 8   // depth (Taylor terms) of the cp.async ring that streams the terms back in STORE mode
#endif
__constant__ double c_ainv[QOCVL_QMAX + 2];   // 1/j
__constant__ float c_ainvf[QOCVL_QMAX + 2];   // 1/j for the complex64 instantiation

// Arithmetic type of the sweeps: R = double (complex128, the parity path) or float (qocvl_set_precision(QOCVL_C64)).
// The state, the Taylor terms (registers, exchange buffers and the HBM stream), the assembled matrix entries and the
// per-warp Xbar partials are R; tables, per-sample cost, terminal cotangent and k_aug_reduce's sums stay double.
template <typename R> struct AugT;
template <> struct AugT<double> {
  typedef cplx C;
  static __device__ __forceinline__ C mk(double re, double im) { return cmake(re, im); }
  static __device__ __forceinline__ double ainv(int j) { return c_ainv[j]; }
};
template <> struct AugT<float> {
  typedef cplxf C;
  static __device__ __forceinline__ C mk(double re, double im) { return cmakef((float)re, (float)im); }
  static __device__ __forceinline__ float ainv(int j) { return c_ainvf[j]; }
};

// STORE = 1: the forward sweep streams every Taylor term t_0 .. t_{q-1} of every slice to HBM and the reverse sweep
// streams them back (cp.async, eight steps ahead of their use) instead of recomputing them:
// a third of the mat-vecs and of the shared-memory exchange traffic are traded for HBM bandwidth that the kernel
// otherwise leaves idle.  STORE = 0 keeps only the checkpoints z_k and recomputes (12x less HBM capacity).
template <typename R, int LPG, int WMAX, int KC, int STORE>
__global__ void __launch_bounds__(512, 1) k_chain_aug(const AugArgs P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  typedef AugT<R> T;
  typedef typename T::C CT;                      // complex of the sweep's arithmetic type
  constexpr bool F64 = sizeof(R) == 8;
  constexpr int NSL = 1 + WMAX;                  // slots per lane (diagonal + off-diagonals)
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int grp = tid / LPG, r = tid % LPG, ngroups = nthreads / LPG;
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
  const size_t gtid = (size_t)blockIdx.x * nthreads + tid, gthreads = (size_t)gridDim.x * nthreads;
  const int sample = blockIdx.x * ngroups + grp;
  const bool item_ok = sample < P.S;
  const bool live = item_ok && r < P.N;
  const int sm = item_ok ? sample : P.S - 1;

  // ---- shared memory: exchange buffers and (only with additive noise) the per-sample constant terms
  CT* carve = (CT*)smem_raw;
  CT* s_x = carve + (size_t)grp * 2 * LPG; carve += (size_t)ngroups * 2 * LPG;     // [ngroups][2 buffers][LPG]
  CT* s_A = carve + (size_t)grp * 2 * NSL * LPG;                                   // [ngroups][2][NSL][LPG]
  carve += P.KA > 0 ? (size_t)ngroups * 2 * NSL * LPG : 0;
  CT* s_ring = carve + tid;                                                      // [RING][nthreads] (STORE)

  int rcol[WMAX], crow[WMAX];
#pragma unroll
  for (int s = 0; s < WMAX; ++s) {
    rcol[s] = (s < P.w) ? P.row_col[(1 + s) * LPG + r] : r;
    crow[s] = (s < P.wc) ? P.col_row[(1 + s) * LPG + r] : r;
  }
  // per-sample multiplier of each of my contributions (0 = no contribution)
  R m_row[NSL * KC], m_col[NSL * KC];
#pragma unroll
  for (int s = 0; s < NSL; ++s) {
#pragma unroll
    for (int i = 0; i < KC; ++i) {
      const int cr = (s <= P.w) ? P.r_cls[(s * KC + i) * LPG + r] : -1;
      const int cc = (s <= P.wc) ? P.c_cls[(s * KC + i) * LPG + r] : -1;
      m_row[s * KC + i] = (R)((cr >= 0 && live) ? P.mcls[(size_t)sm * P.ncls + cr] : 0.0);
      m_col[s * KC + i] = (R)((cc >= 0 && live) ? P.mcls[(size_t)sm * P.ncls + cc] : 0.0);
    }
  }
  const cplx dzero = cmake(0.0, 0.0);
  const CT czero = T::mk(0.0, 0.0);
  if (P.KA > 0) {   // A_e = dt * sum_g add_sg * val_eg for my row-owner and column-owner slots
    for (int s = 0; s < NSL; ++s) {
      cplx ar = dzero, ac = dzero;
      for (int i = 0; i < P.KA; ++i) {
        const int gr = (s <= P.w) ? P.ra_gen[(s * P.KA + i) * LPG + r] : -1;
        const int gc = (s <= P.wc) ? P.ca_gen[(s * P.KA + i) * LPG + r] : -1;
        if (gr >= 0 && live) ar = cfma(P.add[(size_t)sm * P.J + gr], P.ra_val[(s * P.KA + i) * LPG + r], ar);
        if (gc >= 0 && live) ac = cfma(P.add[(size_t)sm * P.J + gc], P.ca_val[(s * P.KA + i) * LPG + r], ac);
      }
      s_A[s * LPG + r] = T::mk(ar.x * P.dt, ar.y * P.dt);
      s_A[(NSL + s) * LPG + r] = T::mk(ac.x * P.dt, ac.y * P.dt);
    }
  }
  const CT A_diag = P.KA > 0 ? s_A[r] : czero;
  CT z = live ? T::mk(P.starts[r].x, P.starts[r].y) : czero;
  const cplx ytgt = live ? P.targets[r] : dzero;
  const int pr = live ? P.pair[r] : -1;
  const double w_qp = live ? P.wq[r] : 0.0, w_lam = live ? P.wq[LPG + r] : 0.0;
  CT t_loc[STORE ? 1 : QOCVL_QMAX];
  CT a_d, a_row[WMAX];
  __syncwarp();

  // exchange: two buffers per sample, each re[LPG] then im[LPG]; the 8-byte accesses of the LPG lanes of a sample
  // hit distinct banks whatever the gather pattern.  Every Taylor loop starts on buffer 0 (after a __syncwarp) and
  // is unrolled by two, so the buffer offsets are immediates; empty slots gather the lane's own element (their
  // matrix value is zero), which costs no extra shared-memory wavefront and needs no predicate.
  // complex64: one 8-byte element per row ([2 buffers][LPG] float2), a single load per gather.
  typedef typename std::conditional<F64, double, CT>::type XE;
  XE* const x_own = (XE*)s_x + r;
  const XE* x_row[WMAX];
  const XE* x_col[WMAX];
#pragma unroll
  for (int s = 0; s < WMAX; ++s) {
    x_row[s] = (const XE*)s_x + rcol[s];
    x_col[s] = (const XE*)s_x + crow[s];
  }
  auto xput = [&](int buf, CT val) {
    if constexpr (F64) { x_own[buf * 2 * LPG] = val.x; x_own[buf * 2 * LPG + LPG] = val.y; }
    else x_own[buf * LPG] = val;
  };
  auto xget = [&](const XE* src, int buf) -> CT {
    if constexpr (F64) return T::mk(src[buf * 2 * LPG], src[buf * 2 * LPG + LPG]);
    else return src[buf * LPG];
  };
  // entry of slot s: A + sum_i m_i * P_i(k).  The table values P_i(k) are fetched one phase ahead into `raw`
  // (global -> registers, L2-resident tables shared by every sample) so the load latency hides behind the sweeps.
  const CT* const p_row = (const CT*)P.p_row;
  const CT* const p_col = (const CT*)P.p_col;
  auto load_raw = [&](const CT* tab, int nslots, CT* raw) {
#pragma unroll
    for (int s = 0; s < NSL; ++s)
#pragma unroll
      for (int i = 0; i < KC; ++i) raw[s * KC + i] = (s <= nslots) ? __ldg(tab + (size_t)(s * KC + i) * LPG) : czero;
  };
  auto slot_value = [&](const CT* raw, const R* m, unsigned mask, int col, int s) -> CT {
    CT acc = czero;
#pragma unroll
    for (int i = 0; i < KC; ++i) {
      acc.x = fma(m[s * KC + i], raw[s * KC + i].x, acc.x);
      acc.y = fma(m[s * KC + i], raw[s * KC + i].y, acc.y);
    }
    if (s == 0) acc = cadd(acc, A_diag);             // the diagonal's additive term lives in registers
    else if ((mask >> s) & 1u) acc = cadd(acc, s_A[(col * NSL + s) * LPG + r]);
    return acc;
  };
  CT raw_r[NSL * KC];                                // row-owner table values of the NEXT slice to be visited
  auto finish_rows = [&]() {
    a_d = slot_value(raw_r, m_row, P.add_row_mask, 0, 0);
#pragma unroll
    for (int s = 0; s < WMAX; ++s) a_row[s] = (s < P.w) ? slot_value(raw_r, m_row, P.add_row_mask, 0, 1 + s) : czero;
  };
  // one forward Taylor step: t <- Xaug t / j
  auto step_fwd = [&](CT t, R invj, int buf) -> CT {
    xput(buf, t);
    __syncwarp();
    CT acc = cmul(a_d, t);
#pragma unroll
    for (int s = 0; s < WMAX; ++s) acc = cfma(a_row[s], xget(x_row[s], buf), acc);
    return cscale(acc, invj);
  };

  // =================================== forward sweep ===================================
  load_raw(p_row + r, P.w, raw_r);
  int q_next = P.qdeg[0];
  size_t n_stream = 0;                               // STORE: Taylor terms written so far (the stream is dense)
  for (int k = 0; k < P.M; ++k) {
    const int q = q_next;
    finish_rows();
    if (k + 1 < P.M) {
      load_raw(p_row + (size_t)(k + 1) * P.ncr * LPG + r, P.w, raw_r);
      q_next = P.qdeg[k + 1];
    }
    const bool keep = P.want_grad && live;
    CT* tout = STORE ? (CT*)P.terms + n_stream * gthreads + gtid : (CT*)P.ckpt + (size_t)k * gthreads + gtid;
    n_stream += q;
    if (keep) tout[0] = z;
    CT t = z;
    __syncwarp();                                   // the previous loop's readers are done with buffer 0
    int j = 1;
    for (; j + 1 <= q; j += 2) {                    // running store pointer: two adds per term instead of a 64-bit multiply
      t = step_fwd(t, T::ainv(j), 0);
      z = cadd(z, t);
      tout += gthreads;
      if (STORE && keep) *tout = t;
      t = step_fwd(t, T::ainv(j + 1), 1);
      z = cadd(z, t);
      tout += gthreads;
      if (STORE && keep && j + 1 < q) *tout = t;
    }
    if (j <= q) {
      t = step_fwd(t, T::ainv(j), 0);
      z = cadd(z, t);
    }
  }

  // =================================== cost ============================================
  cplx zp;                                                          // the cost itself is always evaluated in fp64
  {
    __syncwarp();
    xput(0, z);
    __syncwarp();
    zp = (pr >= 0) ? aug_to_d(xget((const XE*)s_x + pr, 0)) : dzero;
  }
  const cplx zd = aug_to_d(z);
  cplx dsum = cfma_conj(ytgt, zd, dzero);                           // sum_r conj(y_r) z_r
  cplx qp = cscale(cfma_conj(zd, zp, dzero), w_qp);                 // q^dag p
#pragma unroll
  for (int off = LPG / 2; off > 0; off >>= 1) {
    dsum.x += __shfl_xor_sync(0xffffffffu, dsum.x, off);
    dsum.y += __shfl_xor_sync(0xffffffffu, dsum.y, off);
    qp.x += __shfl_xor_sync(0xffffffffu, qp.x, off);
  }
  dsum = cadd(P.d_const, dsum);
  if (r == 0 && item_ok) {
    const double infid = 1.0 - (dsum.x * dsum.x + dsum.y * dsum.y) * P.inv_fid_norm;
    const double adiab = 1.0 - qp.x * P.inv_duration;
    P.sample_out[(size_t)sample * 3 + 0] = P.w_infid * infid + P.w_adiab * adiab;
    P.sample_out[(size_t)sample * 3 + 1] = infid;
    P.sample_out[(size_t)sample * 3 + 2] = adiab;
  }
  if (!P.want_grad) return;

  // terminal cotangent lam = d cost / d conj(Z), scaled by 1/S_global (ensemble mean)
  CT lam;
  {
    const double ws = live ? P.inv_samples : 0.0;
    const double ca = -0.5 * P.w_adiab * P.inv_duration * ws;
    cplx l = cscale(cmul(dsum, ytgt), -P.w_infid * P.inv_fid_norm * ws);
    l = cadd(l, cscale(zp, ca * w_lam));            // q rows receive -(wa/2T) p, p rows -(wa/2T) q
    lam = T::mk(l.x, l.y);
  }

  // =================================== reverse sweep ===================================
  const size_t gwarp = (size_t)blockIdx.x * nwarps + warp;
  CT lam_tb = lam, xc_d, xc[WMAX], a_col[WMAX];
  // one descending step of the adjoint recurrence; Xbar accumulated on my column's entries
  auto step_adj = [&](CT tprev, R invj, int buf) {
    xput(buf, lam_tb);
    __syncwarp();
    CT acc = cfma_conj(a_d, lam_tb, czero);
    xc_d = cfma_conj(lam_tb, tprev, xc_d);
#pragma unroll
    for (int s = 0; s < WMAX; ++s) {
      const CT ta = xget(x_col[s], buf);
      acc = cfma_conj(a_col[s], ta, acc);
      xc[s] = cfma_conj(ta, tprev, xc[s]);
    }
    lam_tb = cadd(lam, cscale(acc, invj));
  };
  auto clear_xbar = [&]() {
    xc_d = czero;
#pragma unroll
    for (int s = 0; s < WMAX; ++s) xc[s] = czero;
  };
  if constexpr (STORE != 0) {
    // The terms were written as one dense stream t_0 .. t_{q_0 - 1} | t_0 .. t_{q_1 - 1} | ... ; the reverse sweep
    // consumes it strictly backwards, across slice boundaries, so it is fetched RING steps ahead of its use with
    // cp.async into a small thread-private shared-memory ring (HBM latency under load is ~6 adjoint steps).
    constexpr int RING = QOCVL_RING;
    const CT* sp = (const CT*)P.terms + gtid + (n_stream - 1) * gthreads;
    const unsigned ring0 = (unsigned)__cvta_generic_to_shared(s_ring);
    const unsigned ring_stride = (unsigned)(nthreads * sizeof(CT));
#pragma unroll
    for (int i = 0; i < RING; ++i) s_ring[(size_t)i * nthreads] = czero;   // dead lanes read zeros for ever
    int slot = 0;                                   // the slot just consumed is the one the next request refills:
    int left = live ? (int)n_stream : 0;            // one ring index; `left` = terms this lane still has to request
    auto issue = [&]() {
      if (left > 0) {
        if constexpr (F64)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring0 + slot * ring_stride), "l"(sp) : "memory");
        else   // 8-byte copies exist only in the .ca flavour
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(ring0 + slot * ring_stride), "l"(sp) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      --left;
      sp -= gthreads;
      slot = (slot + 1) & (RING - 1);
    };
    auto fetch = [&]() -> CT {
      asm volatile("cp.async.wait_group %0;" ::"n"(RING - 1) : "memory");
      const CT v = s_ring[(size_t)slot * nthreads];
      issue();
      return v;
    };
#pragma unroll
    for (int i = 0; i < RING; ++i) issue();
    CT raw_c[NSL * KC];
    int q_cur = P.qdeg[P.M - 1];
    q_next = P.M > 1 ? P.qdeg[P.M - 2] : 0;
    load_raw(p_col + (size_t)(P.M - 1) * P.ncc * LPG + r, P.wc, raw_c);
    for (int k = P.M - 1; k >= 0; --k) {
      const int q = q_cur;
      a_d = slot_value(raw_c, m_col, P.add_col_mask, 1, 0);
#pragma unroll
      for (int s = 0; s < WMAX; ++s) a_col[s] = (s < P.wc) ? slot_value(raw_c, m_col, P.add_col_mask, 1, 1 + s) : czero;
      if (k > 0) {                                  // next slice's tables and degree in flight during this one
        load_raw(p_col + (size_t)(k - 1) * P.ncc * LPG + r, P.wc, raw_c);
        q_cur = q_next;
        q_next = k > 1 ? P.qdeg[k - 2] : 0;
      }
      lam = lam_tb;
      clear_xbar();
      __syncwarp();
      {
        int j = q;
        for (; j >= 2; j -= 2) {
          step_adj(cscale(fetch(), T::ainv(j)), T::ainv(j), 0);
          step_adj(cscale(fetch(), T::ainv(j - 1)), T::ainv(j - 1), 1);
        }
        if (j >= 1) step_adj(cscale(fetch(), T::ainv(j)), T::ainv(j), 0);
      }
      // ---- Ybar_e = sum over the warp's samples of m_s * Xbar_e: the generators' d coef / d wave is the same for
      //      every sample, so the contraction with it happens once per slice in k_aug_reduce, not here
      CT* yout = (CT*)P.y_partial + ((gwarp * P.M + k) * P.ncc) * LPG + r;
  #pragma unroll
      for (int s = 0; s < NSL; ++s) {
        if (s <= P.wc) {
          const CT xval = (s == 0) ? xc_d : xc[s > 0 ? s - 1 : 0];
  #pragma unroll
          for (int i = 0; i < KC; ++i) {
            CT y = cscale(xval, m_col[s * KC + i]);
  #pragma unroll
            for (int off = LPG; off < 32; off <<= 1) {
              y.x += __shfl_xor_sync(0xffffffffu, y.x, off);
              y.y += __shfl_xor_sync(0xffffffffu, y.y, off);
            }
            if (lane < LPG) yout[(size_t)(s * KC + i) * LPG] = y;
          }
        }
      }
    }
  } else {
    load_raw(p_row + (size_t)(P.M - 1) * P.ncr * LPG + r, P.w, raw_r);
    q_next = P.qdeg[P.M - 1];
    CT t_next = live ? ((const CT*)P.ckpt)[(size_t)(P.M - 1) * gthreads + gtid] : czero;
    for (int k = P.M - 1; k >= 0; --k) {
      const int q = q_next;
      finish_rows();
      CT raw_c[NSL * KC];                           // column-owner table values: in flight during the recompute
      load_raw(p_col + (size_t)k * P.ncc * LPG + r, P.wc, raw_c);
      // ---- recompute the forward Taylor terms, stored pre-scaled: s_j = t_{j-1} / j is what the adjoint pairs
      //      with tb_j, and s_{j+1} = X s_j / (j+1)
      CT t = t_next;
      t_loc[0] = t;
      __syncwarp();
      {
        int j = 1;
        for (; j + 1 < q; j += 2) {
          t = step_fwd(t, T::ainv(j + 1), 0);
          t_loc[j] = t;
          t = step_fwd(t, T::ainv(j + 2), 1);
          t_loc[j + 1] = t;
        }
        if (j < q) {
          t = step_fwd(t, T::ainv(j + 1), 0);
          t_loc[j] = t;
        }
      }
#pragma unroll
      for (int s = 0; s < WMAX; ++s) a_col[s] = (s < P.wc) ? slot_value(raw_c, m_col, P.add_col_mask, 1, 1 + s) : czero;
      if (k > 0) {                                  // next slice's rows, checkpoint and degree: in flight during the adjoint
        load_raw(p_row + (size_t)(k - 1) * P.ncr * LPG + r, P.w, raw_r);
        t_next = live ? ((const CT*)P.ckpt)[(size_t)(k - 1) * gthreads + gtid] : czero;
        q_next = P.qdeg[k - 1];
      }
      lam = lam_tb;
      clear_xbar();
      __syncwarp();
      {
        int j = q;
        for (; j >= 2; j -= 2) {
          step_adj(t_loc[j - 1], T::ainv(j), 0);
          step_adj(t_loc[j - 2], T::ainv(j - 1), 1);
        }
        if (j >= 1) step_adj(t_loc[j - 1], T::ainv(j), 0);
      }
      // ---- Ybar_e = sum over the warp's samples of m_s * Xbar_e: the generators' d coef / d wave is the same for
      //      every sample, so the contraction with it happens once per slice in k_aug_reduce, not here
      CT* yout = (CT*)P.y_partial + ((gwarp * P.M + k) * P.ncc) * LPG + r;
  #pragma unroll
      for (int s = 0; s < NSL; ++s) {
        if (s <= P.wc) {
          const CT xval = (s == 0) ? xc_d : xc[s > 0 ? s - 1 : 0];
  #pragma unroll
          for (int i = 0; i < KC; ++i) {
            CT y = cscale(xval, m_col[s * KC + i]);
  #pragma unroll
            for (int off = LPG; off < 32; off <<= 1) {
              y.x += __shfl_xor_sync(0xffffffffu, y.x, off);
              y.y += __shfl_xor_sync(0xffffffffu, y.y, off);
            }
            if (lane < LPG) yout[(size_t)(s * KC + i) * LPG] = y;
          }
        }
      }
    }
  }
}

// gwave[k][c] = 2 dt Re sum_{e,i} Ybar_{e,i}(k) * sum_{g in (e,i)} val_eg * dcoef_kgc, with Ybar summed over the
// per-warp partials in a fixed order (deterministic).  One block per slice, one thread per (contribution, lane).
template <typename CT>
__global__ void __launch_bounds__(1024) k_aug_reduce(const AugArgs P, int L) {
  __shared__ double red[QOCVL_MAX_CHAN][1024];
  const int k = blockIdx.x, t = threadIdx.x, nt = P.ncc * L;
  double term[QOCVL_MAX_CHAN];
#pragma unroll
  for (int c = 0; c < QOCVL_MAX_CHAN; ++c) term[c] = 0.0;
  if (t < nt) {
    cplx y0 = cmake(0, 0), y1 = cmake(0, 0), y2 = cmake(0, 0), y3 = cmake(0, 0);
    const CT* src = (const CT*)P.y_partial + (size_t)k * nt + t;
    const size_t stride = (size_t)P.M * nt;
    int wv = 0;
    for (; wv + 4 <= P.nparts; wv += 4) {          // 4 independent chains (fixed association: still deterministic)
      y0 = cadd(y0, aug_to_d(src[(size_t)wv * stride]));
      y1 = cadd(y1, aug_to_d(src[(size_t)(wv + 1) * stride]));
      y2 = cadd(y2, aug_to_d(src[(size_t)(wv + 2) * stride]));
      y3 = cadd(y3, aug_to_d(src[(size_t)(wv + 3) * stride]));
    }
    for (; wv < P.nparts; ++wv) y0 = cadd(y0, aug_to_d(src[(size_t)wv * stride]));
    const cplx y = cadd(cadd(y0, y1), cadd(y2, y3));
    const int c = t / L, lane = t % L;
    for (int i = 0; i < P.KG; ++i) {
      const int g = P.c_gen[((size_t)c * P.KG + i) * L + lane];
      if (g < 0) continue;
      const cplx yv = cmul(y, P.c_val[((size_t)c * P.KG + i) * L + lane]);
#pragma unroll
      for (int ch = 0; ch < QOCVL_MAX_CHAN; ++ch)
        if (ch < P.C) {
          const cplx d = P.dcoef[((size_t)k * P.J + g) * P.C + ch];
          term[ch] += yv.x * d.x - yv.y * d.y;
        }
    }
  }
#pragma unroll
  for (int c = 0; c < QOCVL_MAX_CHAN; ++c) red[c][t] = term[c];
  __syncthreads();
  for (int off = blockDim.x >> 1; off > 0; off >>= 1) {
    if (t < off) {
#pragma unroll
      for (int c = 0; c < QOCVL_MAX_CHAN; ++c) red[c][t] += red[c][t + off];
    }
    __syncthreads();
  }
  if (t < P.C) P.gwave[(size_t)k * P.C + t] = 2.0 * P.dt * red[t][0];
}

typedef void (*aug_fn)(const AugArgs);
template <typename R, int KC, int STORE>
static aug_fn pick_aug_kc(int lpg, int wmax) {
  if (wmax <= 3) {
    switch (lpg) {
      case 8: return k_chain_aug<R, 8, 3, KC, STORE>;
      case 16: return k_chain_aug<R, 16, 3, KC, STORE>;
      case 32: return k_chain_aug<R, 32, 3, KC, STORE>;
    }
  } else if (wmax <= 6) {
    switch (lpg) {
      case 8: return k_chain_aug<R, 8, 6, KC, STORE>;
      case 16: return k_chain_aug<R, 16, 6, KC, STORE>;
      case 32: return k_chain_aug<R, 32, 6, KC, STORE>;
    }
  }
  return nullptr;
}
template <typename R>
static aug_fn pick_aug_r(int lpg, int wmax, int kc, bool store) {
  if (kc == 1) return store ? pick_aug_kc<R, 1, 1>(lpg, wmax) : pick_aug_kc<R, 1, 0>(lpg, wmax);
  if (kc == 2) return store ? pick_aug_kc<R, 2, 1>(lpg, wmax) : pick_aug_kc<R, 2, 0>(lpg, wmax);
  return nullptr;
}
static aug_fn pick_aug(int lpg, int wmax, int kc, bool store, bool c64) {
  return c64 ? pick_aug_r<float>(lpg, wmax, kc, store) : pick_aug_r<double>(lpg, wmax, kc, store);
}
static size_t aug_elem_bytes(const qocvl_plan* p) { return p->c64 ? sizeof(cplxf) : sizeof(cplx); }

static size_t aug_smem_bytes(const qocvl_plan* p, int ngroups) {
  const int L = p->aug_lpg, nsl = 1 + (p->aug_wmax <= 3 ? 3 : 6);
  return ((size_t)ngroups * 2 * L + (p->la.KA > 0 ? (size_t)ngroups * 2 * nsl * L : 0) +
          (p->aug_store ? (size_t)QOCVL_RING * ngroups * L : 0)) * aug_elem_bytes(p);
}

// ---------------------------------------------------------------------------------------------------
// Host: build the reduced, stacked system from the generators, the live start kets and the symmetry hint.
// Returns QOCVL_ERR_UNSUPPORTED when the reduced system does not fit the kernel (N > 32 rows or more than
// 6 off-diagonals in a row / column); the caller then falls back to chain_small.cu / chain_big.cu.
// ---------------------------------------------------------------------------------------------------
namespace {
struct Sector {
  std::vector<int> rows;        // original basis index of each kept row (orbit representative)
  std::vector<int> rep_of;      // [n] kept-row index of every basis state of the support, -1 outside
  std::vector<int> orbit;       // orbit size of each kept row
};

// smallest set containing `seed` and closed under "column c populated => rows r with pat[r][c] populated"
std::vector<char> closure(const std::vector<char>& pat, int n, std::vector<char> in) {
  std::vector<int> stack;
  for (int i = 0; i < n; ++i) if (in[i]) stack.push_back(i);
  while (!stack.empty()) {
    const int c = stack.back(); stack.pop_back();
    for (int r = 0; r < n; ++r)
      if (pat[(size_t)r * n + c] && !in[r]) { in[r] = 1; stack.push_back(r); }
  }
  return in;
}

// rows of the support; with fold = true one representative per orbit of perm (the support must be closed under it)
bool make_sector(const std::vector<char>& sup, int n, const std::vector<int>& perm, bool fold, Sector* out) {
  out->rows.clear(); out->orbit.clear();
  out->rep_of.assign(n, -1);
  for (int i = 0; i < n; ++i) {
    if (!sup[i] || out->rep_of[i] >= 0) continue;
    const int idx = (int)out->rows.size();
    out->rows.push_back(i);
    out->rep_of[i] = idx;
    int cnt = 1;
    if (fold) {
      for (int j = perm[i]; j != i; j = perm[j]) {
        if (!sup[j]) return false;
        out->rep_of[j] = idx;
        ++cnt;
      }
    }
    out->orbit.push_back(cnt);
  }
  return true;
}
}  // namespace

int qocvl_aug_configure(qocvl_plan* p) {
  const qocvl_desc& d = p->d;
  const int n = d.n, J = d.n_gen, nv = p->n_live, va = p->adiab_live;
  p->use_aug = false;
  if (nv < 1 || (int)p->h_live_s.size() != nv * n) return QOCVL_ERR_UNSUPPORTED;
  auto GU = [&](int j, int a, int b) { return p->h_gu[((size_t)j * n + a) * n + b]; };
  auto GW = [&](int j, int a, int b) { return p->h_gw[((size_t)j * n + a) * n + b]; };
  std::vector<char> PU((size_t)n * n, 0), PW((size_t)n * n, 0);
  bool any_w = false;
  for (int j = 0; j < J; ++j)
    for (int a = 0; a < n; ++a)
      for (int b = 0; b < n; ++b) {
        if (std::abs(GU(j, a, b)) > 0.0) PU[(size_t)a * n + b] = 1;
        if (std::abs(GW(j, a, b)) > 0.0) { PW[(size_t)a * n + b] = 1; any_w = true; }
      }
  // the permutation hint was verified against G^u by qocvl_chain_configure; folding p needs G^w invariant too
  const std::vector<int>& pm = p->h_perm;
  const bool have_perm = !pm.empty();
  bool w_invariant = have_perm;
  for (int j = 0; j < J && w_invariant; ++j)
    for (int a = 0; a < n && w_invariant; ++a)
      for (int b = 0; b < n; ++b)
        if (GW(j, pm[a], pm[b]) != GW(j, a, b)) { w_invariant = false; break; }
  auto invariant = [&](const std::vector<cplx>& kets, int v) {
    for (int i = 0; i < n; ++i) {
      const cplx x = kets[(size_t)v * n + i], y = kets[(size_t)v * n + pm[i]];
      if (x.x != y.x || x.y != y.y) return false;
    }
    return true;
  };
  // ---- sectors of the moving start kets, then of p = W psi
  std::vector<Sector> sec(nv + 1);
  std::vector<int> off(nv + 2, 0);
  for (int v = 0; v < nv; ++v) {
    std::vector<char> seed(n, 0);
    for (int i = 0; i < n; ++i) {
      const cplx x = p->h_live_s[(size_t)v * n + i];
      seed[i] = (x.x != 0.0 || x.y != 0.0);
    }
    const std::vector<char> sup = closure(PU, n, seed);
    bool fold = have_perm && invariant(p->h_live_s, v) && (v != va || !any_w || w_invariant);
    if (p->opt("nofold")) fold = false;
    if (!make_sector(sup, n, pm, fold, &sec[v])) {
      if (!make_sector(sup, n, pm, false, &sec[v])) return QOCVL_ERR_UNSUPPORTED;
      fold = false;
    }
    if (v == va) {   // p rows: whatever B reaches from the support of q, closed under A; folded like q
      std::vector<char> seed_p(n, 0);
      bool any = false;
      for (int c = 0; c < n; ++c)
        if (sup[c])
          for (int a = 0; a < n; ++a)
            if (PW[(size_t)a * n + c]) { seed_p[a] = 1; any = true; }
      if (any) {
        const std::vector<char> sup_p = closure(PU, n, seed_p);
        if (!make_sector(sup_p, n, pm, fold, &sec[nv])) return QOCVL_ERR_UNSUPPORTED;
      }
    }
    off[v + 1] = off[v] + (int)sec[v].rows.size();
  }
  off[nv + 1] = off[nv] + (int)sec[nv].rows.size();
  const int N = off[nv + 1];
  if (N < 1 || N > 32) return QOCVL_ERR_UNSUPPORTED;
  const int L = N <= 8 ? 8 : (N <= 16 ? 16 : 32);
  // ---- stacked generators [J][N][N]
  std::vector<zc> aug((size_t)J * N * N, zc(0, 0));
  auto fill_block = [&](const Sector& rs, int roff, const Sector& cs, int coff, bool wfam) {
    for (int j = 0; j < J; ++j)
      for (size_t a = 0; a < rs.rows.size(); ++a)
        for (int c = 0; c < n; ++c) {
          const int cc = cs.rep_of[c];
          if (cc < 0) continue;
          const zc v = wfam ? GW(j, rs.rows[a], c) : GU(j, rs.rows[a], c);
          aug[((size_t)j * N + roff + a) * N + coff + cc] += v;
        }
  };
  for (int v = 0; v < nv; ++v) fill_block(sec[v], off[v], sec[v], off[v], false);
  if (!sec[nv].rows.empty()) {
    fill_block(sec[nv], off[nv], sec[nv], off[nv], false);
    fill_block(sec[nv], off[nv], sec[va], off[va], true);
  }
  // ---- kets, partner rows and weights of the stacked vector
  std::vector<cplx> st(L, cmake(0, 0)), tg(L, cmake(0, 0));
  std::vector<int> pair(L, -1);
  std::vector<double> wq(2 * (size_t)L, 0.0);
  for (int v = 0; v < nv; ++v)
    for (size_t a = 0; a < sec[v].rows.size(); ++a) {
      st[off[v] + a] = p->h_live_s[(size_t)v * n + sec[v].rows[a]];
      cplx y = cmake(0, 0);
      for (int i = 0; i < n; ++i)
        if (sec[v].rep_of[i] == (int)a) y = cadd(y, p->h_live_t[(size_t)v * n + i]);
      tg[off[v] + a] = y;
    }
  for (size_t a = 0; a < sec[va].rows.size(); ++a) {
    const int pa = sec[nv].rows.empty() ? -1 : sec[nv].rep_of[sec[va].rows[a]];
    if (pa < 0) continue;
    const int rq = off[va] + (int)a, rp = off[nv] + pa;
    if (sec[nv].rows[pa] != sec[va].rows[a] || sec[nv].orbit[pa] != sec[va].orbit[a]) return QOCVL_ERR_UNSUPPORTED;
    pair[rq] = rp; pair[rp] = rq;
    wq[rq] = (double)sec[va].orbit[a];
    wq[L + rq] = wq[L + rp] = (double)sec[va].orbit[a];
  }
  // ---- generator classes: generators whose per-sample multiplier columns are identical share a class
  //      (class 0: multiplier 1 for every sample), so their contributions to an entry merge into one table value
  const int S = p->n_samples;
  if ((int)p->h_mul.size() != S * J || (int)p->h_add.size() != S * J) return QOCVL_ERR_STATE;
  std::vector<int> cls(J, 0);
  std::vector<char> has_add(J, 0);
  int ncls = 1;
  {
    std::map<std::vector<double>, int> seen;
    for (int g = 0; g < J; ++g) {
      std::vector<double> col(S);
      bool ones = true;
      for (int s = 0; s < S; ++s) {
        col[s] = p->h_mul[(size_t)s * J + g];
        ones = ones && col[s] == 1.0;
        const cplx ad = p->h_add[(size_t)s * J + g];
        if (ad.x != 0.0 || ad.y != 0.0) has_add[g] = 1;
      }
      if (ones) continue;
      auto it = seen.find(col);
      if (it == seen.end()) { seen[col] = ncls; cls[g] = ncls++; } else cls[g] = it->second;
    }
  }
  std::vector<double> mcls((size_t)S * ncls, 1.0);
  for (int g = 0; g < J; ++g)
    if (cls[g] > 0)
      for (int s = 0; s < S; ++s) mcls[(size_t)s * ncls + cls[g]] = p->h_mul[(size_t)s * J + g];
  // ---- sparsity pattern of the stacked system: per-row and per-column off-diagonal lists
  auto AG = [&](int j, int a, int b) { return aug[((size_t)j * N + a) * N + b]; };
  std::vector<std::vector<int>> rows(N), cols(N);
  int nnz = 0;
  for (int a = 0; a < N; ++a)
    for (int b = 0; b < N; ++b) {
      bool any = false;
      for (int j = 0; j < J; ++j) any = any || std::abs(AG(j, a, b)) > 0.0;
      if (!any) continue;
      ++nnz;
      if (a != b) { rows[a].push_back(b); cols[b].push_back(a); }
    }
  int w = 0, wc = 0;
  for (int r = 0; r < N; ++r) { w = std::max(w, (int)rows[r].size()); wc = std::max(wc, (int)cols[r].size()); }
  const int wmax = std::max(w, wc);
  if (wmax > 6) return QOCVL_ERR_UNSUPPORTED;
  // contributions of entry (a, b) grouped by class, and its generators that carry an additive per-sample term
  typedef std::vector<std::pair<int, zc>> GenList;
  auto contribs = [&](int a, int b) {
    std::map<int, GenList> by_cls;
    for (int j = 0; j < J; ++j)
      if (std::abs(AG(j, a, b)) > 0.0) by_cls[cls[j]].push_back({j, AG(j, a, b)});
    return by_cls;
  };
  auto addgens = [&](int a, int b) {
    GenList l;
    for (int j = 0; j < J; ++j)
      if (has_add[j] && std::abs(AG(j, a, b)) > 0.0) l.push_back({j, AG(j, a, b)});
    return l;
  };
  int KC = 1, KG = 1, KA = 0;
  for (int a = 0; a < N; ++a)
    for (int b = 0; b < N; ++b) {
      const auto cb = contribs(a, b);
      KC = std::max(KC, (int)cb.size());
      for (const auto& kv : cb) KG = std::max(KG, (int)kv.second.size());
      KA = std::max(KA, (int)addgens(a, b).size());
    }
  if (KC > 2) return QOCVL_ERR_UNSUPPORTED;
  if (!pick_aug(L, wmax, KC, false, p->c64 != 0)) return QOCVL_ERR_UNSUPPORTED;
  const int Lk = L;                                    // lanes per sample of the kernel instance
  const int nr = 1 + w, nc = 1 + wc;
  std::vector<int> row_col((size_t)nr * Lk), col_row((size_t)nc * Lk);
  std::vector<int> r_cls((size_t)nr * KC * Lk, -1), c_cls((size_t)nc * KC * Lk, -1);
  std::vector<int> r_gen((size_t)nr * KC * KG * Lk, -1), c_gen((size_t)nc * KC * KG * Lk, -1);
  std::vector<cplx> r_val((size_t)nr * KC * KG * Lk, cmake(0, 0)), c_val((size_t)nc * KC * KG * Lk, cmake(0, 0));
  const int KAa = std::max(KA, 1);
  std::vector<int> ra_gen((size_t)nr * KAa * Lk, -1), ca_gen((size_t)nc * KAa * Lk, -1);
  std::vector<cplx> ra_val((size_t)nr * KAa * Lk, cmake(0, 0)), ca_val((size_t)nc * KAa * Lk, cmake(0, 0));
  for (int s = 0; s < nr; ++s) for (int r = 0; r < Lk; ++r) row_col[(size_t)s * Lk + r] = r;
  for (int s = 0; s < nc; ++s) for (int r = 0; r < Lk; ++r) col_row[(size_t)s * Lk + r] = r;
  unsigned add_row_mask = 0, add_col_mask = 0;
  auto fill = [&](std::vector<int>& tcls, std::vector<int>& tgen, std::vector<cplx>& tval, std::vector<int>& agen,
                  std::vector<cplx>& aval, unsigned* mask, int slot, int lane, int a, int b) {
    int i = 0;
    for (const auto& kv : contribs(a, b)) {
      tcls[((size_t)slot * KC + i) * Lk + lane] = kv.first;
      for (size_t x = 0; x < kv.second.size(); ++x) {
        const size_t idx = (((size_t)slot * KC + i) * KG + x) * Lk + lane;
        tgen[idx] = kv.second[x].first;
        tval[idx] = cmake(kv.second[x].second.real(), kv.second[x].second.imag());
      }
      ++i;
    }
    const GenList ag = addgens(a, b);
    for (size_t x = 0; x < ag.size(); ++x) {
      const size_t idx = ((size_t)slot * KAa + x) * Lk + lane;
      agen[idx] = ag[x].first;
      aval[idx] = cmake(ag[x].second.real(), ag[x].second.imag());
      *mask |= 1u << slot;
    }
  };
  for (int r = 0; r < N; ++r) {
    fill(r_cls, r_gen, r_val, ra_gen, ra_val, &add_row_mask, 0, r, r, r);
    fill(c_cls, c_gen, c_val, ca_gen, ca_val, &add_col_mask, 0, r, r, r);
    for (size_t s = 0; s < rows[r].size(); ++s) {
      row_col[(1 + s) * Lk + r] = rows[r][s];
      fill(r_cls, r_gen, r_val, ra_gen, ra_val, &add_row_mask, 1 + (int)s, r, r, rows[r][s]);
    }
    for (size_t s = 0; s < cols[r].size(); ++s) {
      col_row[(1 + s) * Lk + r] = cols[r][s];
      fill(c_cls, c_gen, c_val, ca_gen, ca_val, &add_col_mask, 1 + (int)s, r, cols[r][s], r);
    }
  }
  st.resize(Lk, cmake(0, 0)); tg.resize(Lk, cmake(0, 0)); pair.resize(Lk, -1);
  {
    std::vector<double> wq2(2 * (size_t)Lk, 0.0);
    for (int r = 0; r < L; ++r) { wq2[r] = wq[r]; wq2[Lk + r] = wq[L + r]; }
    wq = wq2;
  }
  // ---- upload
  AugLayout& la = p->la;
  la.release();
  la.N = N; la.L = Lk; la.w = w; la.wc = wc; la.KC = KC; la.KG = KG; la.KA = KA; la.ncls = ncls; la.nnz = nnz;
  la.add_row_mask = add_row_mask; la.add_col_mask = add_col_mask;
  const int M = d.n_slices;
#define AUG_UP(ptr, vec)                                                                              \
  do {                                                                                                \
    QCUDA(cudaMalloc((void**)&(ptr), std::max<size_t>(1, (vec).size()) * sizeof((vec)[0])));           \
    if (!(vec).empty())                                                                               \
      QCUDA(cudaMemcpy((ptr), (vec).data(), (vec).size() * sizeof((vec)[0]), cudaMemcpyHostToDevice)); \
  } while (0)
  AUG_UP(la.row_col, row_col); AUG_UP(la.col_row, col_row);
  AUG_UP(la.r_cls, r_cls); AUG_UP(la.c_cls, c_cls);
  AUG_UP(la.r_gen, r_gen); AUG_UP(la.c_gen, c_gen); AUG_UP(la.r_val, r_val); AUG_UP(la.c_val, c_val);
  AUG_UP(la.ra_gen, ra_gen); AUG_UP(la.ca_gen, ca_gen); AUG_UP(la.ra_val, ra_val); AUG_UP(la.ca_val, ca_val);
  AUG_UP(la.mcls, mcls);
  QCUDA(cudaMalloc((void**)&la.p_row, (size_t)M * nr * KC * Lk * sizeof(cplx)));
  QCUDA(cudaMalloc((void**)&la.p_col, (size_t)M * nc * KC * Lk * sizeof(cplx)));
  for (void** q : {(void**)&p->d_aug_starts, (void**)&p->d_aug_targets, (void**)&p->d_aug_pair, (void**)&p->d_aug_wq})
    if (*q) { cudaFree(*q); *q = nullptr; }
  AUG_UP(p->d_aug_starts, st); AUG_UP(p->d_aug_targets, tg); AUG_UP(p->d_aug_pair, pair); AUG_UP(p->d_aug_wq, wq);
#undef AUG_UP
  p->aug_n = N; p->aug_lpg = Lk; p->aug_nnz = nnz; p->aug_wmax = wmax;
  p->h_aug_pair = pair;
  {
    double inv[QOCVL_QMAX + 2];
    inv[0] = 0.0;
    for (int j = 1; j < QOCVL_QMAX + 2; ++j) inv[j] = 1.0 / (double)j;
    QCUDA(cudaMemcpyToSymbol(c_ainv, inv, sizeof(inv)));
    float invf[QOCVL_QMAX + 2];
    for (int j = 0; j < QOCVL_QMAX + 2; ++j) invf[j] = (float)inv[j];
    QCUDA(cudaMemcpyToSymbol(c_ainvf, invf, sizeof(invf)));
  }
  // ---- third-generation kernels (chain_duo.cu: two lanes per coupled sector) whenever they take the system
  p->use_duo = false;
  p->h_aug_gen = aug; p->h_aug_cls = cls; p->h_aug_hasadd = has_add; p->h_aug_mcls = mcls; p->h_aug_wq = wq;
  p->h_aug_st = st; p->h_aug_tg = tg;
  // basis states behind every stacked row (its orbit): chain_duo.cu bounds the norm of a role by their columns
  p->aug_p0 = off[nv];                                  // first stacked row of p = W psi
  p->h_aug_members.assign(N, std::vector<int>());
  for (int v = 0; v <= nv; ++v)
    for (int i = 0; i < n; ++i)
      if (sec[v].rep_of.size() == (size_t)n && sec[v].rep_of[i] >= 0) p->h_aug_members[off[v] + sec[v].rep_of[i]].push_back(i);
  if (p->opt("duo", 1) != 0) {
    const int rc = qocvl_duo_configure(p);
    if (rc == QOCVL_OK) { p->use_aug = true; return QOCVL_OK; }
    if (rc != QOCVL_ERR_UNSUPPORTED) return rc;
  }
  // ---- Taylor terms: streamed through HBM (STORE) when the a-priori degree cap and the device memory allow it
  const int gpw = 32 / Lk;
  p->aug_qs = p->qcap;                                // a-priori cap (qocvl_chain_configure); k_coefficients flags excess
  p->aug_store = (size_t)M * p->aug_qs < ((size_t)1 << 31);   // the kernel counts the stream's terms in 32 bits
  if (p->has_opt("store")) p->aug_store = p->aug_store && p->opt("store") != 0;
  if (p->d_ckpt) { cudaFree(p->d_ckpt); p->d_ckpt = nullptr; p->bytes -= p->ckpt_bytes; p->ckpt_bytes = 0; }
  if (p->d_ypart) { cudaFree(p->d_ypart); p->d_ypart = nullptr; p->bytes -= p->ypart_bytes; p->ypart_bytes = 0; }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device);
  for (int attempt = 0; attempt < 2; ++attempt) {
    // launch geometry: a CTA is only a container of independent groups (<= 256 threads, 128 registers per thread,
    // so two CTAs fit an SM).  Pick the number of samples per CTA that minimises the number of samples the busiest
    // SM has to process (wave quantisation), preferring larger CTAs on ties.
    const int gmax = 256 / Lk / gpw * gpw;
    int ngroups = gpw;
    long best_load = -1;
    for (int g = gpw; g <= gmax; g += gpw) {
      const long n_ctas = (S + g - 1) / g;
      // resident CTAs per SM (register-limited: 128 registers per thread in complex128, 96 in complex64)
      const long per_sm = std::max(1, std::min(32, (p->c64 ? 640 : 512) / (g * Lk)));
      const long waves = (n_ctas + per_sm * sms - 1) / (per_sm * sms);
      const long load = waves * per_sm * g;
      if (best_load < 0 || load <= best_load) { best_load = load; ngroups = g; }
    }
    if (p->has_opt("spb")) ngroups = std::max(gpw, std::min(gmax, (p->opt("spb") + gpw - 1) / gpw * gpw));
    ngroups = std::min(ngroups, (S + gpw - 1) / gpw * gpw);
    p->spb = ngroups; p->groups = ngroups; p->threads = ngroups * Lk;
    p->chain_smem = aug_smem_bytes(p, ngroups);
    if (p->chain_smem > 200 * 1024) return QOCVL_ERR_UNSUPPORTED;
    p->blocks = (S + ngroups - 1) / ngroups;
    const size_t total_threads = (size_t)p->blocks * p->threads;
    p->aug_parts = (int)((total_threads + 31) / 32);
    p->ckpt_bytes = (size_t)M * (p->aug_store ? p->aug_qs : 1) * total_threads * aug_elem_bytes(p);
    p->ypart_bytes = (size_t)p->aug_parts * M * nc * KC * Lk * aug_elem_bytes(p);
    if (p->aug_store) {                                // fall back to recomputation when HBM capacity is short
      size_t free_b = 0, total_b = 0;
      cudaMemGetInfo(&free_b, &total_b);
      if (p->ckpt_bytes + p->ypart_bytes > free_b / 10 * 8) { p->aug_store = false; continue; }
    }
    break;
  }
  aug_fn fn = pick_aug(Lk, wmax, KC, p->aug_store, p->c64 != 0);
  if (!fn) return QOCVL_ERR_UNSUPPORTED;
  QCUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->chain_smem));
  p->n_parts = 0;                                     // no [warps][M][C] partial-gradient buffer on this path
  QCUDA(cudaMalloc((void**)&p->d_ckpt, p->ckpt_bytes));   // checkpoints z_k, or all Taylor terms in STORE mode
  QCUDA(cudaMalloc((void**)&p->d_ypart, p->ypart_bytes));
  p->bytes += p->ckpt_bytes + p->ypart_bytes;
  p->qcap_apriori = p->qcap;
  if (!p->aug_store) p->qcap = QOCVL_QMAX;            // thread-local ring: no limit on the degree
  p->use_aug = true;
  return QOCVL_OK;
}

int qocvl_aug_launch(qocvl_plan* p, bool want_grad, double* gwave, cudaStream_t st) {
  const qocvl_desc& d = p->d;
  const AugLayout& la = p->la;
  AugArgs a;
  a.N = la.N; a.J = d.n_gen; a.M = d.n_slices; a.C = d.n_chan; a.S = p->n_samples;
  a.w = la.w; a.wc = la.wc; a.ncr = (1 + la.w) * la.KC; a.ncc = (1 + la.wc) * la.KC;
  a.KA = la.KA; a.KG = la.KG; a.ncls = la.ncls; a.nparts = p->aug_parts;
  a.add_row_mask = la.add_row_mask; a.add_col_mask = la.add_col_mask;
  a.want_grad = want_grad ? 1 : 0;
  a.dt = d.dt; a.inv_duration = 1.0 / d.duration; a.w_infid = d.w_infid; a.w_adiab = d.w_adiab;
  a.inv_fid_norm = 1.0 / d.fid_norm; a.inv_samples = 1.0 / (double)p->global_samples;
  a.d_const = p->d_const;
  a.row_col = la.row_col; a.col_row = la.col_row; a.r_cls = la.r_cls; a.c_cls = la.c_cls;
  a.r_gen = la.r_gen; a.c_gen = la.c_gen; a.r_val = la.r_val; a.c_val = la.c_val;
  a.ra_gen = la.ra_gen; a.ca_gen = la.ca_gen; a.ra_val = la.ra_val; a.ca_val = la.ca_val;
  a.p_row = la.p_row; a.p_col = la.p_col;
  a.coef = p->d_coef; a.dcoef = p->d_dcoef; a.qdeg = p->d_qdeg; a.mcls = la.mcls; a.add = p->d_add;
  a.starts = p->d_aug_starts; a.targets = p->d_aug_targets; a.pair = p->d_aug_pair; a.wq = p->d_aug_wq;
  a.ckpt = p->d_ckpt; a.terms = p->d_ckpt; a.qs = p->aug_qs;
  a.sample_out = p->d_sample_out; a.y_partial = p->d_ypart; a.gwave = gwave;
  if (p->use_duo) return qocvl_duo_launch(p, want_grad, gwave, st);
  if (p->c64) k_aug_tables<cplxf><<<d.n_slices, 256, 0, st>>>(a, la.L);
  else k_aug_tables<cplx><<<d.n_slices, 256, 0, st>>>(a, la.L);
  pick_aug(la.L, p->aug_wmax, la.KC, p->aug_store, p->c64 != 0)<<<p->blocks, p->threads, p->chain_smem, st>>>(a);
  p->launches += 2;
  if (want_grad) {
    if (!gwave) return QOCVL_ERR_ARG;
    int nt = 32;
    while (nt < a.ncc * la.L) nt <<= 1;
    if (p->c64) k_aug_reduce<cplxf><<<d.n_slices, nt, 0, st>>>(a, la.L);
    else k_aug_reduce<cplx><<<d.n_slices, nt, 0, st>>>(a, la.L);
    p->launches++;
  }
  QCUDA(cudaGetLastError());
  return QOCVL_OK;
}
