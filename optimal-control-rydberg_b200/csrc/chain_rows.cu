// chain_rows.cu -- the ensemble hot kernel, second generation: one COUPLED SECTOR of the reduced stacked system per warp.
//
// Same numbers as chain_aug.cu (ST:247-302 / TQ:460-567 and the reverse pass of jax.value_and_grad, TQ:592) on the same
// reduced stacked system Xaug (chain_aug.cu builds it: reachability + orbit folding, CZ: 12 rows, 36 non-zeros).  What
// changes is the mapping, driven by round 1's profile (FP64 pipe 45 % busy, but more than half of the issued DFMAs were
// padding: 12 live rows in 16-lane groups, 2 of 3 off-diagonal slots used):
//   * Xaug is block diagonal up to the p <- q coupling.  Its connected components ("roles"; CZ: {U|10> , W|10>} = 6 rows
//     and {U|11>} = 6 rows) never exchange data during the sweeps, so every warp holds ONE role of floor(32 / rows)
//     samples (CZ: 5 samples x 6 rows = 30 of 32 lanes) and runs the slot loop only as deep as ITS role needs (CZ: 3 and
//     2 off-diagonal slots).  Issued multiply-adds per useful one: 1.78 -> 1.24.
//   * Off-row elements are gathered with warp shuffles (no shared-memory exchange buffer, no __syncwarp round trip).
//   * The Taylor recurrence runs on UNSCALED powers u_j = X u_{j-1}, z' = sum_j u_j / j! (one DFMA pair with a constant
//     operand per row instead of a scale and an add); the adjoint recurrence on w_j = tbar_j / j!:
//     w_q = lam' / q!, w_{j-1} = lam' / (j-1)! + X^dag w_j, Xbar = sum_j w_j u_{j-1}^dag, lam = w_0 -- the exact adjoint of
//     the forward polynomial, as before.
//   * STORE = 1: the forward sweep streams u_0 .. u_{q-1} of every slice to HBM, one contiguous 512-byte slab per warp
//     and term ([warp][term][lane]); the reverse sweep streams them back with 1-D BULK TMA (cp.async.bulk.shared::cluster
//     .global, 2 KB = 4 terms per copy, issued by one lane, completion on an mbarrier ring, 4 copies in flight per warp)
//     instead of one cp.async + ring bookkeeping per thread and term.  STORE = 0 keeps only checkpoints and recomputes.
//   * The per-warp partial sums of m_s * Xbar_e are reduced over the warp's samples with shuffles before they are
//     written (384 B per warp and slice instead of 1 KB per two samples); k_rows_reduce sums them in a fixed order.
// One CTA barrier per evaluation (the cost needs the overlaps of all roles of a sample); none inside the sweeps.
#include "chain_aug.cuh"
#include <algorithm>
#include <numeric>

#define ROWS_MAX_ROLES 8
#define ROWS_TCH 4      // Taylor terms per bulk copy (4 x 512 B = 2 KB)
#define ROWS_NSLOT 4    // bulk copies in flight per warp
#define ROWS_RING 8     // depth of the per-thread cp.async ring of the non-TMA variant

__constant__ double c_ifact[QOCVL_QMAX + 2];   // 1 / j!

struct RowsArgs {
  AugArgs a;
  int L;                          // lanes per sample of chain_aug.cu's tables (their row stride)
  int nroles, G, wpc;             // roles, samples per CTA, warps per CTA
  int rows[ROWS_MAX_ROLES], spw[ROWS_MAX_ROLES], w[ROWS_MAX_ROLES], wc[ROWS_MAX_ROLES];
  int warp0[ROWS_MAX_ROLES], warps[ROWS_MAX_ROLES], parts[ROWS_MAX_ROLES];
  unsigned long long yoff[ROWS_MAX_ROLES];   // element offset of the role's region in y_partial
  const int* grow;                // [nroles][32] stacked row of (role, local row)
  const int* lrow;                // [L] local row of a stacked row inside its role
  unsigned long long cap;         // terms reserved per warp in the stream (multiple of ROWS_TCH)
  unsigned long long total_warps;
};

__device__ __forceinline__ cplx shfl_c(cplx v, int src) {
  v.x = __shfl_sync(0xffffffffu, v.x, src);
  v.y = __shfl_sync(0xffffffffu, v.y, src);
  return v;
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int WMAX, int KC, int STORE, int TMA>
__global__ void __launch_bounds__(256) k_chain_rows(const RowsArgs Q) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const AugArgs& P = Q.a;
  constexpr int NSL = 1 + WMAX;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nthreads = blockDim.x;
  int c = 0;
#pragma unroll
  for (int i = 1; i < ROWS_MAX_ROLES; ++i)
    if (i < Q.nroles && warp >= Q.warp0[i]) c = i;
  const int R = Q.rows[c], spw = Q.spw[c], wr = Q.w[c], wcr = Q.wc[c];
  const int jw = warp - Q.warp0[c];
  const int sl = lane / R, r = lane - sl * R;
  const bool lane_ok = sl < spw;
  const int s_cta = jw * spw + sl;                       // sample of this lane inside the CTA
  const int sample = blockIdx.x * Q.G + s_cta;
  const bool live = lane_ok && s_cta < Q.G && sample < P.S;
  const int L = Q.L, N = P.N;
  const int gr = lane_ok ? Q.grow[c * 32 + r] : 0;       // my row of the stacked system (index into the tables)
  const int sm = live ? sample : 0;
  const size_t gw = (size_t)blockIdx.x * Q.wpc + warp;   // global warp: owns one slab sequence of the stream

  // ---- shared memory: [bulk ring | mbarriers | cost exchange | additive off-diagonal terms]
  unsigned char* carve = smem_raw;
  cplx* ring = (cplx*)carve + (size_t)warp * ((STORE && TMA) ? ROWS_NSLOT * ROWS_TCH * 32 : (STORE ? ROWS_RING * 32 : 0));
  carve += (size_t)Q.wpc * ((STORE && TMA) ? ROWS_NSLOT * ROWS_TCH * 32 : (STORE ? ROWS_RING * 32 : 0)) * sizeof(cplx);
  unsigned long long* bars = (unsigned long long*)carve + warp * ROWS_NSLOT;
  carve += (size_t)Q.wpc * ROWS_NSLOT * sizeof(unsigned long long);
  cplx* s_d = (cplx*)carve; carve += (size_t)Q.G * N * sizeof(cplx);
  double* s_q = (double*)carve; carve += (size_t)Q.G * N * sizeof(double);
  cplx* s_A = (cplx*)carve + tid;                         // [2][NSL][nthreads] (only with additive noise)

  // ---- gather sources (lanes of my own sample inside this warp), per-sample class multipliers
  int src_row[WMAX], src_col[WMAX];
#pragma unroll
  for (int s = 0; s < WMAX; ++s) {
    const int cr = (s < wr && lane_ok) ? P.row_col[(1 + s) * L + gr] : gr;
    const int cc = (s < wcr && lane_ok) ? P.col_row[(1 + s) * L + gr] : gr;
    src_row[s] = lane_ok ? sl * R + Q.lrow[cr] : lane;
    src_col[s] = lane_ok ? sl * R + Q.lrow[cc] : lane;
  }
  double m_row[NSL * KC], m_col[NSL * KC];
#pragma unroll
  for (int s = 0; s < NSL; ++s) {
#pragma unroll
    for (int i = 0; i < KC; ++i) {
      const int cr = (s <= wr) ? P.r_cls[(s * KC + i) * L + gr] : -1;
      const int cc = (s <= wcr) ? P.c_cls[(s * KC + i) * L + gr] : -1;
      m_row[s * KC + i] = (cr >= 0 && live) ? P.mcls[(size_t)sm * P.ncls + cr] : 0.0;
      m_col[s * KC + i] = (cc >= 0 && live) ? P.mcls[(size_t)sm * P.ncls + cc] : 0.0;
    }
  }
  const cplx czero = cmake(0.0, 0.0);
  cplx A_diag = czero;
  if (P.KA > 0) {   // A_e = dt * sum_g add_sg * val_eg for my row-owner and column-owner slots
    for (int s = 0; s < NSL; ++s) {
      cplx ar = czero, ac = czero;
      for (int i = 0; i < P.KA; ++i) {
        const int g_r = (s <= wr) ? P.ra_gen[(s * P.KA + i) * L + gr] : -1;
        const int g_c = (s <= wcr) ? P.ca_gen[(s * P.KA + i) * L + gr] : -1;
        if (g_r >= 0 && live) ar = cfma(P.add[(size_t)sm * P.J + g_r], P.ra_val[(s * P.KA + i) * L + gr], ar);
        if (g_c >= 0 && live) ac = cfma(P.add[(size_t)sm * P.J + g_c], P.ca_val[(s * P.KA + i) * L + gr], ac);
      }
      if (s == 0) A_diag = cscale(ar, P.dt);
      s_A[(size_t)s * nthreads] = cscale(ar, P.dt);
      s_A[(size_t)(NSL + s) * nthreads] = cscale(ac, P.dt);
    }
  }
  // entry of slot s: A + sum_i m_i * P_i(k); the table values P_i(k) are fetched one slice ahead into `raw`
  const cplx* const p_row = (const cplx*)P.p_row + gr;
  const cplx* const p_col = (const cplx*)P.p_col + gr;
  auto load_raw = [&](const cplx* tab, int nslots, cplx* raw) {
#pragma unroll
    for (int s = 0; s < NSL; ++s)
#pragma unroll
      for (int i = 0; i < KC; ++i) raw[s * KC + i] = (s <= nslots) ? __ldg(tab + (size_t)(s * KC + i) * L) : czero;
  };
  auto slot_value = [&](const cplx* raw, const double* m, unsigned mask, int col, int s) -> cplx {
    cplx acc = czero;
#pragma unroll
    for (int i = 0; i < KC; ++i) {
      acc.x = fma(m[s * KC + i], raw[s * KC + i].x, acc.x);
      acc.y = fma(m[s * KC + i], raw[s * KC + i].y, acc.y);
    }
    if (s == 0) acc = cadd(acc, A_diag);
    else if ((mask >> s) & 1u) acc = cadd(acc, s_A[(size_t)(col * NSL + s) * nthreads]);
    return acc;
  };
  cplx a_d, a_row[WMAX], a_col[WMAX];
  cplx raw_r[NSL * KC];
  auto finish_rows = [&]() {
    a_d = slot_value(raw_r, m_row, P.add_row_mask, 0, 0);
#pragma unroll
    for (int s = 0; s < WMAX; ++s) a_row[s] = (s < wr) ? slot_value(raw_r, m_row, P.add_row_mask, 0, 1 + s) : czero;
  };
  // u <- Xaug u (unscaled Taylor power): the slot loop is as deep as this warp's role needs (warp-uniform bound)
  auto matvec = [&](cplx u) -> cplx {
    cplx acc = cmul(a_d, u);
#pragma unroll
    for (int s = 0; s < WMAX; ++s)
      if (s < wr) acc = cfma(a_row[s], shfl_c(u, src_row[s]), acc);
    return acc;
  };

  cplx z = live ? P.starts[gr] : czero;
  const cplx ytgt = live ? P.targets[gr] : czero;
  const int pr = live ? P.pair[gr] : -1;
  const double w_qp = live ? P.wq[gr] : 0.0, w_lam = live ? P.wq[L + gr] : 0.0;
  const bool keep = P.want_grad != 0;
  cplx* const stream = (cplx*)P.terms + (gw * Q.cap) * 32 + lane;          // STORE: my lane of this warp's slabs
  cplx* const ckpt = (cplx*)P.ckpt + gw * 32 + lane;                        // !STORE: [M][total warps][32]

  // =================================== forward sweep ===================================
  load_raw(p_row, wr, raw_r);
  int q_next = P.qdeg[0];
  unsigned n_stream = 0;
  cplx* tout = stream;
  for (int k = 0; k < P.M; ++k) {
    const int q = q_next;
    finish_rows();
    if (k + 1 < P.M) {
      load_raw(p_row + (size_t)(k + 1) * P.ncr * L, wr, raw_r);
      q_next = P.qdeg[k + 1];
    }
    if (keep) {
      if (STORE) { __stcs((double2*)tout, z); tout += 32; }
      else __stcs((double2*)(ckpt + (size_t)k * Q.total_warps * 32), z);
    }
    n_stream += q;
    cplx u = z;
    for (int j = 1; j <= q; ++j) {
      u = matvec(u);
      z.x = fma(u.x, c_ifact[j], z.x);
      z.y = fma(u.y, c_ifact[j], z.y);
      if (STORE && keep && j < q) { __stcs((double2*)tout, u); tout += 32; }
    }
  }

  // =================================== cost ============================================
  {
    const cplx zp = shfl_c(z, pr >= 0 ? sl * R + Q.lrow[pr] : lane);
    if (live) {
      s_d[s_cta * N + gr] = cfma_conj(ytgt, z, czero);                       // conj(y_r) z_r
      s_q[s_cta * N + gr] = pr >= 0 ? w_qp * (z.x * zp.x + z.y * zp.y) : 0.0;   // Re(conj(z) p) on the q rows
    }
    __syncthreads();                                                          // the only CTA barrier of the evaluation
    cplx dsum = P.d_const;
    double qp = 0.0;
    if (live)
      for (int g = 0; g < N; ++g) {                                           // fixed order: identical in every lane
        dsum = cadd(dsum, s_d[s_cta * N + g]);
        qp += s_q[s_cta * N + g];
      }
    if (live && c == 0 && r == 0) {
      const double infid = 1.0 - (dsum.x * dsum.x + dsum.y * dsum.y) * P.inv_fid_norm;
      const double adiab = 1.0 - qp * P.inv_duration;
      P.sample_out[(size_t)sample * 3 + 0] = P.w_infid * infid + P.w_adiab * adiab;
      P.sample_out[(size_t)sample * 3 + 1] = infid;
      P.sample_out[(size_t)sample * 3 + 2] = adiab;
    }
    if (!P.want_grad) return;
    // terminal cotangent lam = d cost / d conj(Z), scaled by 1/S_global (ensemble mean)
    const double ws = live ? P.inv_samples : 0.0;
    const double ca = -0.5 * P.w_adiab * P.inv_duration * ws;
    cplx l = cscale(cmul(dsum, ytgt), -P.w_infid * P.inv_fid_norm * ws);
    z = cadd(l, cscale(zp, ca * w_lam));              // q rows receive -(wa/2T) p, p rows -(wa/2T) q
  }
  cplx lam = z, lam_tb = z;

  // =================================== reverse sweep ===================================
  cplx xc_d, xc[WMAX];
  // one descending step of the adjoint recurrence on w_j = tbar_j / j!: Xbar accumulated on my column's entries
  auto step_adj = [&](cplx uprev, double ifact_prev) {
    cplx acc = cfma_conj(a_d, lam_tb, czero);
    xc_d = cfma_conj(lam_tb, uprev, xc_d);
#pragma unroll
    for (int s = 0; s < WMAX; ++s)
      if (s < wcr) {
        const cplx ta = shfl_c(lam_tb, src_col[s]);
        acc = cfma_conj(a_col[s], ta, acc);
        xc[s] = cfma_conj(ta, uprev, xc[s]);
      }
    lam_tb.x = fma(lam.x, ifact_prev, acc.x);
    lam_tb.y = fma(lam.y, ifact_prev, acc.y);
  };
  // per-warp sums over the samples of m_s * Xbar_e (the generators' d coef / d wave is the same for every sample, so
  // the contraction with it happens once per slice in k_rows_reduce): tree over the sample slots, fixed order
  const int ncc_c = (1 + wcr) * KC;
  cplx* const ybase = (cplx*)P.y_partial + Q.yoff[c] + ((size_t)blockIdx.x * Q.warps[c] + jw) * ncc_c * R + r;
  const size_t ystride = (size_t)Q.parts[c] * ncc_c * R;     // per slice
  int top = 1;
  while (top * 2 < spw) top *= 2;
  auto emit = [&](int k) {
    cplx* yout = ybase + (size_t)k * ystride;
#pragma unroll
    for (int s = 0; s < NSL; ++s) {
      if (s <= wcr) {
        const cplx xval = (s == 0) ? xc_d : xc[s > 0 ? s - 1 : 0];
#pragma unroll
        for (int i = 0; i < KC; ++i) {
          cplx y = cscale(xval, m_col[s * KC + i]);
          for (int o = top; o >= 1; o >>= 1) {
            const int off = o * R;
            const double yx = __shfl_down_sync(0xffffffffu, y.x, off);
            const double yy = __shfl_down_sync(0xffffffffu, y.y, off);
            if (lane + off < 32) { y.x += yx; y.y += yy; }
          }
          if (lane < R) yout[(size_t)(s * KC + i) * R] = y;
        }
      }
    }
  };
  auto clear_xbar = [&]() {
    xc_d = czero;
#pragma unroll
    for (int s = 0; s < WMAX; ++s) xc[s] = czero;
  };
  cplx raw_c[NSL * KC];
  auto finish_cols = [&]() {
    a_d = slot_value(raw_c, m_col, P.add_col_mask, 1, 0);
#pragma unroll
    for (int s = 0; s < WMAX; ++s) a_col[s] = (s < wcr) ? slot_value(raw_c, m_col, P.add_col_mask, 1, 1 + s) : czero;
  };

  if constexpr (STORE != 0) {
    // ---- the stream is consumed strictly backwards: term n lives in chunk n / TCH of this warp's slab sequence
    int n = (int)n_stream - 1;
    unsigned seq = 0;                                    // chunks consumed so far
    const int c_last = n / ROWS_TCH;
    const unsigned ring_u32 = smem_u32(ring);
    const unsigned bar_u32 = smem_u32(bars);
    const cplx* const gsrc = (const cplx*)P.terms + (gw * Q.cap) * 32;       // chunk ch starts at gsrc + ch*TCH*32
    constexpr unsigned CH_BYTES = ROWS_TCH * 32 * sizeof(cplx);
    // cp.async ring of the non-TMA variant (per-thread 16-byte copies, ROWS_RING terms ahead)
    const cplx* sp = stream + (size_t)n * 32;
    int slot = 0, left = (int)n_stream;
    int pending = -1;                                    // TMA: ring slot whose chunk was finished by the previous fetch
    auto tma_issue = [&](int ring_slot, int ch) {        // one elected lane: arm the barrier, start the bulk copy
      if (lane == 0 && ch >= 0) {
        const unsigned bar = bar_u32 + ring_slot * 8, dst = ring_u32 + ring_slot * CH_BYTES;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(CH_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(gsrc + (size_t)ch * ROWS_TCH * 32), "r"(CH_BYTES), "r"(bar) : "memory");
      }
    };
    auto issue_async = [&]() {
      if (left > 0)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring_u32 + (slot * 32 + lane) * 16), "l"(sp) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
      --left;
      sp -= 32;
      slot = (slot + 1) & (ROWS_RING - 1);
    };
    if constexpr (TMA != 0) {
      // the slabs were written with ordinary (generic-proxy) stores by the lanes of THIS warp; the bulk copies read them
      // through the async proxy: every writer fences, then the warp synchronises, then one lane issues
      asm volatile("fence.proxy.async.global;" ::: "memory");
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < ROWS_NSLOT; ++i)
          asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_u32 + i * 8) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < ROWS_NSLOT; ++i) tma_issue(i, c_last - i);
    } else {
#pragma unroll
      for (int i = 0; i < ROWS_RING; ++i) ring[i * 32 + lane] = czero;
#pragma unroll
      for (int i = 0; i < ROWS_RING; ++i) issue_async();
    }
    bool fresh = true;                                   // the next term is the first one read from its chunk
    auto fetch = [&]() -> cplx {
      if constexpr (TMA != 0) {
        const int rs = seq & (ROWS_NSLOT - 1);
        if (pending >= 0) {                              // refill the slot finished one step ago (its reads have retired)
          __syncwarp();
          tma_issue(pending, c_last - (int)(seq - 1) - ROWS_NSLOT);
          pending = -1;
        }
        if (fresh) {
          const unsigned bar = bar_u32 + rs * 8, parity = (seq / ROWS_NSLOT) & 1u;
          unsigned done = 0;
          while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(bar), "r"(parity) : "memory");
          fresh = false;
        }
        const int within = n & (ROWS_TCH - 1);
        const cplx v = ring[(rs * ROWS_TCH + within) * 32 + lane];
        if (within == 0) { pending = rs; ++seq; fresh = true; }
        --n;
        return v;
      } else {
        asm volatile("cp.async.wait_group %0;" ::"n"(ROWS_RING - 1) : "memory");
        const cplx v = ring[slot * 32 + lane];
        issue_async();
        return v;
      }
    };
    int q_cur = P.qdeg[P.M - 1];
    q_next = P.M > 1 ? P.qdeg[P.M - 2] : 0;
    load_raw(p_col + (size_t)(P.M - 1) * P.ncc * L, wcr, raw_c);
    for (int k = P.M - 1; k >= 0; --k) {
      const int q = q_cur;
      finish_cols();
      if (k > 0) {
        load_raw(p_col + (size_t)(k - 1) * P.ncc * L, wcr, raw_c);
        q_cur = q_next;
        q_next = k > 1 ? P.qdeg[k - 2] : 0;
      }
      lam = lam_tb;
      lam_tb = cscale(lam, c_ifact[q]);                 // w_q = lam' / q!
      clear_xbar();
      for (int j = q; j >= 1; --j) step_adj(fetch(), c_ifact[j - 1]);
      emit(k);
    }
  } else {
    cplx t_loc[QOCVL_QMAX];
    load_raw(p_row + (size_t)(P.M - 1) * P.ncr * L, wr, raw_r);
    q_next = P.qdeg[P.M - 1];
    cplx z_next = __ldcs((const double2*)(ckpt + (size_t)(P.M - 1) * Q.total_warps * 32));
    for (int k = P.M - 1; k >= 0; --k) {
      const int q = q_next;
      finish_rows();
      load_raw(p_col + (size_t)k * P.ncc * L, wcr, raw_c);   // in flight during the recompute
      cplx u = z_next;
      t_loc[0] = u;
      for (int j = 1; j < q; ++j) {
        u = matvec(u);
        t_loc[j] = u;
      }
      finish_cols();                                        // same diagonal entry, plus my column's entries
      if (k > 0) {
        load_raw(p_row + (size_t)(k - 1) * P.ncr * L, wr, raw_r);
        z_next = __ldcs((const double2*)(ckpt + (size_t)(k - 1) * Q.total_warps * 32));
        q_next = P.qdeg[k - 1];
      }
      lam = lam_tb;
      lam_tb = cscale(lam, c_ifact[q]);
      clear_xbar();
      for (int j = q; j >= 1; --j) step_adj(t_loc[j - 1], c_ifact[j - 1]);
      emit(k);
    }
  }
}

// gwave[k][c] = 2 dt Re sum_{e,i} Ybar_{e,i}(k) * sum_{g in (e,i)} val_eg * dcoef_kgc, with Ybar summed over the
// per-warp partials in a fixed order (deterministic).  One block per slice; thread = (part group, entry).
__global__ void __launch_bounds__(256) k_rows_reduce(const RowsArgs Q, int KC, int ngroups, int etotal) {
  extern __shared__ __align__(16) unsigned char smem_red[];
  const AugArgs& P = Q.a;
  cplx* sy = (cplx*)smem_red;                                  // [ngroups][etotal]
  double* red = (double*)(sy + (size_t)ngroups * etotal);      // [QOCVL_MAX_CHAN][256]
  const int k = blockIdx.x, t = threadIdx.x;
  const int g = t / etotal, e = t - g * etotal;
  // which role / slot / row is entry e?
  int c = 0, base = 0;
  for (int i = 0; i < Q.nroles; ++i) {
    const int ne = (1 + Q.wc[i]) * KC * Q.rows[i];
    if (e >= base + ne) { base += ne; c = i + 1; }
  }
  cplx y = cmake(0, 0);
  int contrib = 0, r = 0;
  if (g < ngroups && c < Q.nroles) {
    const int R = Q.rows[c], ncc_c = (1 + Q.wc[c]) * KC, parts = Q.parts[c];
    const int le = e - base;
    contrib = le / R; r = le - contrib * R;
    const cplx* src = (const cplx*)P.y_partial + Q.yoff[c] + (size_t)k * parts * ncc_c * R + le;
    const size_t stride = (size_t)ncc_c * R;
    cplx y0 = cmake(0, 0), y1 = cmake(0, 0);
    int pt = g;
    for (; pt + ngroups < parts; pt += 2 * ngroups) {       // two independent chains, fixed association
      y0 = cadd(y0, src[(size_t)pt * stride]);
      y1 = cadd(y1, src[(size_t)(pt + ngroups) * stride]);
    }
    if (pt < parts) y0 = cadd(y0, src[(size_t)pt * stride]);
    y = cadd(y0, y1);
    sy[(size_t)g * etotal + e] = y;
  }
  __syncthreads();
  double term[QOCVL_MAX_CHAN];
#pragma unroll
  for (int ch = 0; ch < QOCVL_MAX_CHAN; ++ch) term[ch] = 0.0;
  if (g == 0 && c < Q.nroles) {
    y = cmake(0, 0);
    for (int i = 0; i < ngroups; ++i) y = cadd(y, sy[(size_t)i * etotal + e]);
    const int gr = Q.grow[c * 32 + r], L = Q.L;
    for (int i = 0; i < P.KG; ++i) {
      const int gen = P.c_gen[((size_t)contrib * P.KG + i) * L + gr];
      if (gen < 0) continue;
      const cplx yv = cmul(y, P.c_val[((size_t)contrib * P.KG + i) * L + gr]);
#pragma unroll
      for (int ch = 0; ch < QOCVL_MAX_CHAN; ++ch)
        if (ch < P.C) {
          const cplx d = P.dcoef[((size_t)k * P.J + gen) * P.C + ch];
          term[ch] += yv.x * d.x - yv.y * d.y;
        }
    }
  }
#pragma unroll
  for (int ch = 0; ch < QOCVL_MAX_CHAN; ++ch) red[ch * 256 + t] = term[ch];
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (t < off) {
#pragma unroll
      for (int ch = 0; ch < QOCVL_MAX_CHAN; ++ch) red[ch * 256 + t] += red[ch * 256 + t + off];
    }
    __syncthreads();
  }
  if (t < P.C) P.gwave[(size_t)k * P.C + t] = 2.0 * P.dt * red[t * 256];
}

typedef void (*rows_fn)(const RowsArgs);
template <int KC, int STORE, int TMA>
static rows_fn pick_rows_w(int wmax) {
  if (wmax <= 3) return k_chain_rows<3, KC, STORE, TMA>;
  if (wmax <= 6) return k_chain_rows<6, KC, STORE, TMA>;
  return nullptr;
}
static rows_fn pick_rows(int wmax, int kc, bool store, bool tma) {
  if (kc == 1) return store ? (tma ? pick_rows_w<1, 1, 1>(wmax) : pick_rows_w<1, 1, 0>(wmax)) : pick_rows_w<1, 0, 0>(wmax);
  if (kc == 2) return store ? (tma ? pick_rows_w<2, 1, 1>(wmax) : pick_rows_w<2, 1, 0>(wmax)) : pick_rows_w<2, 0, 0>(wmax);
  return nullptr;
}

static size_t rows_smem_bytes(const qocvl_plan* p, int wmax) {
  const RowsGeom& g = p->rg;
  const int nsl = 1 + (wmax <= 3 ? 3 : 6);
  size_t b = 0;
  if (p->aug_store) b += (size_t)g.wpc * (g.tma ? ROWS_NSLOT * ROWS_TCH : ROWS_RING) * 32 * sizeof(cplx);
  b += (size_t)g.wpc * ROWS_NSLOT * sizeof(unsigned long long);
  b += (size_t)g.G * p->aug_n * (sizeof(cplx) + sizeof(double));
  if (p->la.KA > 0) b += (size_t)2 * nsl * g.wpc * 32 * sizeof(cplx);
  return b;
}

// ---------------------------------------------------------------------------------------------------
// Host: roles = connected components of the stacked pattern (rows paired in the adiabaticity term are kept together),
// launch geometry, workspaces.  Called by qocvl_aug_configure once the reduced tables exist.
// Returns QOCVL_ERR_UNSUPPORTED when the system does not fit (the lane-group kernel of chain_aug.cu then runs).
// ---------------------------------------------------------------------------------------------------
int qocvl_rows_configure(qocvl_plan* p) {
  const AugLayout& la = p->la;
  const int N = la.N, L = la.L, M = p->d.n_slices, S = p->n_samples;
  RowsGeom& g = p->rg;
  g = RowsGeom();
  p->use_rows = false;
  if (p->c64) return QOCVL_ERR_UNSUPPORTED;                 // complex64 runs on the lane-group kernel
  std::vector<int> parent(N);
  std::iota(parent.begin(), parent.end(), 0);
  auto find = [&](int x) { while (parent[x] != x) x = parent[x] = parent[parent[x]]; return x; };
  auto unite = [&](int a, int b) { a = find(a); b = find(b); if (a != b) parent[std::max(a, b)] = std::min(a, b); };
  for (int r = 0; r < N; ++r) {
    for (int s = 1; s <= la.w; ++s) unite(r, p->h_aug_row_col[(size_t)s * L + r]);
    if (p->h_aug_pair[r] >= 0) unite(r, p->h_aug_pair[r]);
  }
  std::vector<std::vector<int>> comp;
  std::vector<int> comp_of(N, -1);
  for (int r = 0; r < N; ++r) {
    const int root = find(r);
    if (comp_of[root] < 0) { comp_of[root] = (int)comp.size(); comp.emplace_back(); }
    comp[comp_of[root]].push_back(r);
  }
  // more components than role slots: merge the smallest ones (a role may hold several independent blocks)
  while ((int)comp.size() > ROWS_MAX_ROLES) {
    std::sort(comp.begin(), comp.end(), [](const std::vector<int>& a, const std::vector<int>& b) { return a.size() > b.size(); });
    std::vector<int> last = comp.back(); comp.pop_back();
    comp.back().insert(comp.back().end(), last.begin(), last.end());
  }
  g.nroles = (int)comp.size();
  std::vector<int> grow((size_t)g.nroles * 32, -1), lrow(L, 0);
  int wmax = 0;
  for (int c = 0; c < g.nroles; ++c) {
    std::sort(comp[c].begin(), comp[c].end());
    const int R = (int)comp[c].size();
    if (R > 32) return QOCVL_ERR_UNSUPPORTED;
    g.rows[c] = R;
    g.spw[c] = 32 / R;
    int w = 0, wc = 0;
    for (int i = 0; i < R; ++i) {
      const int row = comp[c][i];
      grow[(size_t)c * 32 + i] = row;
      lrow[row] = i;
      for (int s = 1; s <= la.w; ++s) if (p->h_aug_row_col[(size_t)s * L + row] != row) w = std::max(w, s);
      for (int s = 1; s <= la.wc; ++s) if (p->h_aug_col_row[(size_t)s * L + row] != row) wc = std::max(wc, s);
    }
    g.w[c] = w; g.wc[c] = wc;
    wmax = std::max(wmax, std::max(w, wc));
  }
  {
    int etotal = 0;
    for (int c = 0; c < g.nroles; ++c) etotal += (1 + g.wc[c]) * la.KC * g.rows[c];
    if (etotal > 256) return QOCVL_ERR_UNSUPPORTED;      // k_rows_reduce: one thread per (slot, class, row)
  }
  g.tma = p->opt("rows_tma", 1) != 0;
  if (!pick_rows(wmax, la.KC, false, false)) return QOCVL_ERR_UNSUPPORTED;
  // samples per CTA: least lane waste, at most 8 warps; ties -> more samples per CTA
  int bestG = 1;
  double best_eff = -1.0;
  for (int G = 1; G <= 64; ++G) {
    int warps = 0;
    for (int c = 0; c < g.nroles; ++c) warps += (G + g.spw[c] - 1) / g.spw[c];
    if (warps > 8) break;
    const double eff = (double)G * N / (32.0 * warps);
    if (eff > best_eff + 1e-12 || (eff > best_eff - 1e-12 && warps <= 4)) { best_eff = std::max(best_eff, eff); bestG = G; }
  }
  if (p->has_opt("spb")) bestG = std::max(1, std::min(64, p->opt("spb")));
  g.G = std::min(bestG, std::max(1, S));
  g.wpc = 0;
  for (int c = 0; c < g.nroles; ++c) {
    g.warp0[c] = g.wpc;
    g.warps[c] = (g.G + g.spw[c] - 1) / g.spw[c];
    g.wpc += g.warps[c];
  }
  if (g.wpc > 8) return QOCVL_ERR_UNSUPPORTED;
  p->blocks = (S + g.G - 1) / g.G;
  p->threads = g.wpc * 32;
  p->spb = g.G; p->groups = g.G;
  g.total_warps = (size_t)p->blocks * g.wpc;
  // ---- Taylor terms: streamed through HBM (STORE) when the a-priori degree cap and the device memory allow it
  p->aug_qs = p->qcap;
  p->aug_store = (size_t)M * p->aug_qs < ((size_t)1 << 31);
  if (p->has_opt("store")) p->aug_store = p->aug_store && p->opt("store") != 0;
  g.cap = ((size_t)M * p->aug_qs + ROWS_TCH - 1) / ROWS_TCH * ROWS_TCH + ROWS_TCH;
  size_t ytot = 0;
  for (int c = 0; c < g.nroles; ++c) {
    g.parts[c] = p->blocks * g.warps[c];
    g.yoff[c] = ytot;
    ytot += (size_t)M * g.parts[c] * (1 + g.wc[c]) * la.KC * g.rows[c];
  }
  if (p->d_ckpt) { cudaFree(p->d_ckpt); p->d_ckpt = nullptr; p->bytes -= p->ckpt_bytes; p->ckpt_bytes = 0; }
  if (p->d_ypart) { cudaFree(p->d_ypart); p->d_ypart = nullptr; p->bytes -= p->ypart_bytes; p->ypart_bytes = 0; }
  p->ypart_bytes = ytot * sizeof(cplx);
  for (int attempt = 0; attempt < 2; ++attempt) {
    p->ckpt_bytes = g.total_warps * 32 * sizeof(cplx) * (p->aug_store ? g.cap : (size_t)M);
    if (p->aug_store) {                                  // fall back to recomputation when HBM capacity is short
      size_t free_b = 0, total_b = 0;
      cudaMemGetInfo(&free_b, &total_b);
      if (p->ckpt_bytes + p->ypart_bytes > free_b / 10 * 8) { p->aug_store = false; continue; }
    }
    break;
  }
  p->chain_smem = rows_smem_bytes(p, wmax);
  if (p->chain_smem > 200 * 1024) return QOCVL_ERR_UNSUPPORTED;
  rows_fn fn = pick_rows(wmax, la.KC, p->aug_store, g.tma);
  if (!fn) return QOCVL_ERR_UNSUPPORTED;
  QCUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->chain_smem));
  if (g.d_grow) { cudaFree(g.d_grow); g.d_grow = nullptr; }
  if (g.d_lrow) { cudaFree(g.d_lrow); g.d_lrow = nullptr; }
  QCUDA(cudaMalloc((void**)&g.d_grow, grow.size() * sizeof(int)));
  QCUDA(cudaMalloc((void**)&g.d_lrow, lrow.size() * sizeof(int)));
  QCUDA(cudaMemcpy(g.d_grow, grow.data(), grow.size() * sizeof(int), cudaMemcpyHostToDevice));
  QCUDA(cudaMemcpy(g.d_lrow, lrow.data(), lrow.size() * sizeof(int), cudaMemcpyHostToDevice));
  QCUDA(cudaMalloc((void**)&p->d_ckpt, p->ckpt_bytes));
  QCUDA(cudaMalloc((void**)&p->d_ypart, std::max<size_t>(16, p->ypart_bytes)));
  p->bytes += p->ckpt_bytes + p->ypart_bytes;
  {
    double f[QOCVL_QMAX + 2];
    f[0] = 1.0;
    for (int j = 1; j < QOCVL_QMAX + 2; ++j) f[j] = f[j - 1] / (double)j;
    QCUDA(cudaMemcpyToSymbol(c_ifact, f, sizeof(f)));
  }
  g.wmax = wmax;
  p->aug_parts = (int)g.total_warps;
  p->n_parts = 0;
  p->qcap_apriori = p->qcap;
  if (!p->aug_store) p->qcap = QOCVL_QMAX;
  p->use_rows = true;
  return QOCVL_OK;
}

int qocvl_rows_launch(qocvl_plan* p, const AugArgs& a, bool want_grad, cudaStream_t st) {
  const RowsGeom& g = p->rg;
  const AugLayout& la = p->la;
  RowsArgs q;
  q.a = a;
  q.L = la.L; q.nroles = g.nroles; q.G = g.G; q.wpc = g.wpc;
  for (int c = 0; c < ROWS_MAX_ROLES; ++c) {
    q.rows[c] = g.rows[c]; q.spw[c] = g.spw[c]; q.w[c] = g.w[c]; q.wc[c] = g.wc[c];
    q.warp0[c] = g.warp0[c]; q.warps[c] = g.warps[c]; q.parts[c] = g.parts[c]; q.yoff[c] = g.yoff[c];
  }
  q.grow = g.d_grow; q.lrow = g.d_lrow; q.cap = g.cap; q.total_warps = g.total_warps;
  k_aug_tables<cplx><<<p->d.n_slices, 256, 0, st>>>(a, la.L);
  pick_rows(g.wmax, la.KC, p->aug_store, g.tma)<<<p->blocks, p->threads, p->chain_smem, st>>>(q);
  p->launches += 2;
  if (want_grad) {
    int etotal = 0;
    for (int c = 0; c < g.nroles; ++c) etotal += (1 + g.wc[c]) * la.KC * g.rows[c];
    const int ngroups = 256 / etotal;
    const size_t smem = (size_t)ngroups * etotal * sizeof(cplx) + (size_t)QOCVL_MAX_CHAN * 256 * sizeof(double);
    k_rows_reduce<<<p->d.n_slices, 256, smem, st>>>(q, la.KC, ngroups, etotal);
    p->launches++;
  }
  QCUDA(cudaGetLastError());
  return QOCVL_OK;
}
