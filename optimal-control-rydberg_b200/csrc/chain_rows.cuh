// chain_rows.cuh -- the ensemble hot kernel, second generation: one COUPLED SECTOR of the reduced stacked system per warp.
//
// Same numbers as chain_aug.cu (ST:247-302 / TQ:460-567 and the reverse pass of jax.value_and_grad, TQ:592) on the same
// reduced stacked system Xaug (chain_aug.cu builds it: reachability + orbit folding, CZ: 12 rows, 36 non-zeros).  What
// changes is the mapping, driven by round 1's profile (FP64 pipe 45 % busy, but more than half of the issued DFMAs were
// padding: 12 live rows in 16-lane groups, 2 of 3 off-diagonal slots used):
//   * Xaug is block diagonal up to the p <- q coupling.  Its connected components ("roles"; CZ: {U|10> , W|10>} = 6 rows
//     and {U|11>} = 6 rows) never exchange data during the sweeps, so every warp holds ONE role of floor(32 / rows)
//     samples (CZ: 5 samples x 6 rows = 30 of 32 lanes) and runs a slot loop compiled for ITS role's depth (the kernel
//     dispatches once per warp to a body instantiated for 2, 3 or 6 off-diagonal slots; CZ: 3 and 2): issued complex
//     multiply-adds per useful one 1.78 -> 1.24, and no per-slot branches.
//   * Off-row elements are gathered with 32-bit warp shuffles (no shared-memory exchange, no __syncwarp round trip).
//   * The Taylor recurrence runs on UNSCALED powers u_j = X u_{j-1}, z' = sum_j u_j / j! (one DFMA pair with a constant
//     operand per row instead of a scale and an add); the adjoint recurrence on w_j = tbar_j / j!:
//     w_q = lam' / q!, w_{j-1} = lam' / (j-1)! + X^dag w_j, Xbar = sum_j w_j u_{j-1}^dag, lam = w_0 -- the exact adjoint of
//     the forward polynomial, as before.
//   * SPLIT = 1 (chosen when the ensemble leaves fewer than ~6 warps per SM, e.g. a 512-sample shard of a strong-scaling
//     run): the row sum runs on two accumulator chains (dependent depth 5 instead of 8) at the price of one extra add.
//   * STORE = 1: the forward sweep streams u_0 .. u_{q-1} of every slice to HBM, one contiguous 512-byte slab per warp
//     and term ([warp][term][lane]); the reverse sweep streams them back.  TMA = 1: ONE 1-D bulk copy per warp and SLICE
//     (cp.async.bulk.shared::cluster.global, q_k x 512 B, issued by one lane one slice ahead of its use, completion on an
//     mbarrier, two slabs per warp in shared memory) instead of one cp.async + ring bookkeeping per thread and term.
//     TMA = 0: the per-thread cp.async ring of chain_aug.cu.  STORE = 0 keeps only checkpoints and recomputes.
//   * One role per CTA, the roles of one sample group form a THREAD-BLOCK CLUSTER: the role is a function of blockIdx
//     (so the compiler knows it is uniform: with the role derived from the warp index every shuffle was wrapped in
//     WARPSYNC / ENDCOLLECTIVE and the slot branches in BSSY / BSYNC), and the only cross-role step of an evaluation -- the
//     cost needs the overlaps of all roles of a sample -- reads the other CTAs' shared memory (DSMEM) after one
//     cluster barrier.
//   * The per-warp partial sums of m_s * Xbar_e are reduced over the warp's samples with shuffles before they are
//     written (384 B per warp and slice instead of 1 KB per two samples); k_rows_reduce sums them in a fixed order.
// One cluster barrier per evaluation; no barrier of any kind inside the sweeps.
#pragma once
#include "chain_aug.cuh"
#include <cooperative_groups.h>

#define ROWS_MAX_ROLES 8
#define ROWS_RING 8     // depth of the per-thread cp.async ring of the non-TMA variant

struct RowsArgs {
  AugArgs a;
  int L;                          // lanes per sample of chain_aug.cu's tables (their row stride)
  int nroles, G, wpc;             // roles (= CTAs per cluster), samples per cluster, warps per CTA
  int rows[ROWS_MAX_ROLES], spw[ROWS_MAX_ROLES], w[ROWS_MAX_ROLES], wc[ROWS_MAX_ROLES];
  int warp0[ROWS_MAX_ROLES], warps[ROWS_MAX_ROLES], parts[ROWS_MAX_ROLES];
  unsigned long long yoff[ROWS_MAX_ROLES];   // element offset of the role's region in y_partial
  const int* grow;                // [nroles][32] stacked row of (role, local row)
  const int* lrow;                // [L] local row of a stacked row inside its role
  const int* rrole;               // [L] role of a stacked row
  unsigned long long cap;         // terms reserved per warp in the stream
  unsigned long long total_warps;
  int qs;                         // a-priori Taylor cap: terms per slab slot of the bulk ring
  double ifact[QOCVL_QMAX + 2];   // 1 / j!
};

__device__ __forceinline__ double shfl_d(double v, int src) {
  // 32-bit halves: the CUDA double overload packs / unpacks with `asm volatile`, which pins every shuffle in program order
  const int lo = __shfl_sync(0xffffffffu, __double2loint(v), src);
  const int hi = __shfl_sync(0xffffffffu, __double2hiint(v), src);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ cplx shfl_c(cplx v, int src) { return cmake(shfl_d(v.x, src), shfl_d(v.y, src)); }
__device__ __forceinline__ double shfl_down_d(double v, int off) {
  const int lo = __shfl_down_sync(0xffffffffu, __double2loint(v), off);
  const int hi = __shfl_down_sync(0xffffffffu, __double2hiint(v), off);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// The whole evaluation of one warp, compiled for W off-diagonal slots per row / column.
template <int W, int KC, int STORE, int TMA, int SPLIT>
__device__ __forceinline__ void rows_body(const RowsArgs& Q, unsigned char* smem_raw, const int c) {
  const AugArgs& P = Q.a;
  constexpr int NSL = 1 + W;
  namespace cg = cooperative_groups;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nthreads = blockDim.x;
  const int R = Q.rows[c], spw = Q.spw[c];
  const int jw = warp;                                   // warp of this role inside the sample group
  const int sl = lane / R, r = lane - sl * R;
  const bool lane_ok = sl < spw && jw < Q.warps[c];      // (a role with few rows needs fewer warps than the CTA has)
  const int s_cta = jw * spw + sl;                       // sample of this lane inside the cluster's group
  const int group = blockIdx.x / Q.nroles;               // sample group = cluster index
  const int sample = group * Q.G + s_cta;
  const bool live = lane_ok && s_cta < Q.G && sample < P.S;
  const int L = Q.L, N = P.N;
  const int gr = lane_ok ? Q.grow[c * 32 + r] : 0;       // my row of the stacked system (index into the tables)
  const int sm = live ? sample : 0;
  const size_t gw = (size_t)blockIdx.x * Q.wpc + warp;   // global warp: owns one slab sequence of the stream

  // ---- shared memory: [bulk slabs or cp.async ring | mbarriers | cost exchange | additive off-diagonal terms]
  constexpr int NBAR = 2;
  const size_t ring_elems = STORE ? (TMA ? (size_t)2 * Q.qs * 32 : (size_t)ROWS_RING * 32) : 0;
  unsigned char* carve = smem_raw;
  cplx* ring = (cplx*)carve + (size_t)warp * ring_elems;
  carve += (size_t)Q.wpc * ring_elems * sizeof(cplx);
  unsigned long long* bars = (unsigned long long*)carve + warp * NBAR;
  carve += (size_t)Q.wpc * NBAR * sizeof(unsigned long long);
  cplx* s_d = (cplx*)carve; carve += (size_t)Q.G * N * sizeof(cplx);
  double* s_q = (double*)carve; carve += ((size_t)Q.G * N * sizeof(double) + 15) / 16 * 16;   // keep s_A 16-byte aligned
  cplx* s_A = (cplx*)carve + tid;                         // [2][NSL][nthreads] (only with additive noise)

  // ---- gather sources (lanes of my own sample inside this warp), per-sample class multipliers.  Slots the tables do
  //      not have (W may exceed the system's widest row) gather the lane's own element against a zero entry.
  int src_row[W], src_col[W];
#pragma unroll
  for (int s = 0; s < W; ++s) {
    const int cr = (s < P.w && lane_ok) ? P.row_col[(1 + s) * L + gr] : gr;
    const int cc = (s < P.wc && lane_ok) ? P.col_row[(1 + s) * L + gr] : gr;
    src_row[s] = lane_ok ? sl * R + Q.lrow[cr] : lane;
    src_col[s] = lane_ok ? sl * R + Q.lrow[cc] : lane;
  }
  double m_row[NSL * KC], m_col[NSL * KC];
#pragma unroll
  for (int s = 0; s < NSL; ++s) {
#pragma unroll
    for (int i = 0; i < KC; ++i) {
      const int cr = (s <= P.w) ? P.r_cls[(s * KC + i) * L + gr] : -1;
      const int cc = (s <= P.wc) ? P.c_cls[(s * KC + i) * L + gr] : -1;
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
        const int g_r = (s <= P.w) ? P.ra_gen[(s * P.KA + i) * L + gr] : -1;
        const int g_c = (s <= P.wc) ? P.ca_gen[(s * P.KA + i) * L + gr] : -1;
        if (g_r >= 0 && live) ar = cfma(P.add[(size_t)sm * P.J + g_r], P.ra_val[(s * P.KA + i) * L + gr], ar);
        if (g_c >= 0 && live) ac = cfma(P.add[(size_t)sm * P.J + g_c], P.ca_val[(s * P.KA + i) * L + gr], ac);
      }
      if (s == 0) A_diag = cscale(ar, P.dt);
      s_A[(size_t)s * nthreads] = cscale(ar, P.dt);
      s_A[(size_t)(NSL + s) * nthreads] = cscale(ac, P.dt);
    }
  }
  const bool add_off_row = (P.add_row_mask >> 1) != 0u, add_off_col = (P.add_col_mask >> 1) != 0u;
  // entry of slot s: A + sum_i m_i * P_i(k); the table values P_i(k) are fetched one slice ahead into `raw`
  const cplx* const p_row = (const cplx*)P.p_row + gr;
  const cplx* const p_col = (const cplx*)P.p_col + gr;
  auto load_raw = [&](const cplx* tab, int nslots, cplx* raw) {
#pragma unroll
    for (int s = 0; s < NSL; ++s)
#pragma unroll
      for (int i = 0; i < KC; ++i) raw[s * KC + i] = (s <= nslots) ? __ldg(tab + (size_t)(s * KC + i) * L) : czero;
  };
  auto slot_value = [&](const cplx* raw, const double* m, bool add_off, int col, int s) -> cplx {
    cplx acc = czero;
#pragma unroll
    for (int i = 0; i < KC; ++i) {
      acc.x = fma(m[s * KC + i], raw[s * KC + i].x, acc.x);
      acc.y = fma(m[s * KC + i], raw[s * KC + i].y, acc.y);
    }
    if (s == 0) acc = cadd(acc, A_diag);
    else if (add_off) acc = cadd(acc, s_A[(size_t)(col * NSL + s) * nthreads]);   // rare: additive noise off the diagonal
    return acc;
  };
  cplx a_d, a_row[W], a_col[W];
  cplx raw_r[NSL * KC];
  auto finish_rows = [&]() {
    a_d = slot_value(raw_r, m_row, add_off_row, 0, 0);
#pragma unroll
    for (int s = 0; s < W; ++s) a_row[s] = slot_value(raw_r, m_row, add_off_row, 0, 1 + s);
  };
  // u <- Xaug u (unscaled Taylor power)
  auto matvec = [&](cplx u) -> cplx {
    cplx x[W];
#pragma unroll
    for (int s = 0; s < W; ++s) x[s] = shfl_c(u, src_row[s]);
    if constexpr (SPLIT != 0) {
      cplx acc0 = cmul(a_d, u), acc1 = cmul(a_row[0], x[0]);
#pragma unroll
      for (int s = 1; s < W; ++s) {
        if (s & 1) acc0 = cfma(a_row[s], x[s], acc0);
        else acc1 = cfma(a_row[s], x[s], acc1);
      }
      return cadd(acc0, acc1);
    } else {
      cplx acc = cmul(a_d, u);
#pragma unroll
      for (int s = 0; s < W; ++s) acc = cfma(a_row[s], x[s], acc);
      return acc;
    }
  };

  cplx z = live ? P.starts[gr] : czero;
  const cplx ytgt = live ? P.targets[gr] : czero;
  const int pr = live ? P.pair[gr] : -1;
  const double w_qp = live ? P.wq[gr] : 0.0, w_lam = live ? P.wq[L + gr] : 0.0;
  const bool keep = P.want_grad != 0;
  cplx* const stream = (cplx*)P.terms + (gw * Q.cap) * 32 + lane;          // STORE: my lane of this warp's slabs
  cplx* const ckpt = (cplx*)P.ckpt + gw * 32 + lane;                        // !STORE: [M][total warps][32]
  const int nsr = min(W, P.w), nsc = min(W, P.wc);                         // table slots that exist for this body

  // =================================== forward sweep ===================================
  load_raw(p_row, nsr, raw_r);
  int q_next = P.qdeg[0];
  unsigned n_stream = 0;
  cplx* tout = stream;
  for (int k = 0; k < P.M; ++k) {
    const int q = q_next;
    finish_rows();
    if (k + 1 < P.M) {
      load_raw(p_row + (size_t)(k + 1) * P.ncr * L, nsr, raw_r);
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
      z.x = fma(u.x, Q.ifact[j], z.x);
      z.y = fma(u.y, Q.ifact[j], z.y);
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
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();                                                           // the only barrier of the evaluation
    cplx dsum = P.d_const;
    double qp = 0.0;
    if (live)
      for (int g = 0; g < N; ++g) {                                           // fixed order: identical in every lane
        const int owner = Q.rrole[g];                                         // CTA of the cluster that holds row g
        dsum = cadd(dsum, cluster.map_shared_rank(s_d, owner)[s_cta * N + g]);
        qp += cluster.map_shared_rank(s_q, owner)[s_cta * N + g];
      }
    cluster.sync();                                       // nobody leaves (or reuses its shared memory) while peers still read
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
  cplx xc_d, xc[W];
  // one descending step of the adjoint recurrence on w_j = tbar_j / j!: Xbar accumulated on my column's entries
  auto step_adj = [&](cplx uprev, double ifact_prev) {
    cplx ta[W];
#pragma unroll
    for (int s = 0; s < W; ++s) ta[s] = shfl_c(lam_tb, src_col[s]);
    xc_d = cfma_conj(lam_tb, uprev, xc_d);
#pragma unroll
    for (int s = 0; s < W; ++s) xc[s] = cfma_conj(ta[s], uprev, xc[s]);
    cplx acc;
    if constexpr (SPLIT != 0) {
      cplx acc0 = cfma_conj(a_d, lam_tb, czero), acc1 = cfma_conj(a_col[0], ta[0], czero);
#pragma unroll
      for (int s = 1; s < W; ++s) {
        if (s & 1) acc0 = cfma_conj(a_col[s], ta[s], acc0);
        else acc1 = cfma_conj(a_col[s], ta[s], acc1);
      }
      acc = cadd(acc0, acc1);
    } else {
      acc = cfma_conj(a_d, lam_tb, czero);
#pragma unroll
      for (int s = 0; s < W; ++s) acc = cfma_conj(a_col[s], ta[s], acc);
    }
    lam_tb.x = fma(lam.x, ifact_prev, acc.x);
    lam_tb.y = fma(lam.y, ifact_prev, acc.y);
  };
  // per-warp sums over the samples of m_s * Xbar_e (the generators' d coef / d wave is the same for every sample, so
  // the contraction with it happens once per slice in k_rows_reduce): tree over the sample slots, fixed order.
  // The role's region of y_partial holds (1 + wc_role) * KC contributions per row.
  const int wcr = Q.wc[c];
  const int ncc_c = (1 + wcr) * KC;
  cplx* const ybase = (cplx*)P.y_partial + Q.yoff[c] + ((size_t)group * Q.warps[c] + (jw < Q.warps[c] ? jw : 0)) * ncc_c * R + r;
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
            const double yx = shfl_down_d(y.x, off), yy = shfl_down_d(y.y, off);
            if (lane + off < 32) { y.x += yx; y.y += yy; }
          }
          if (lane < R && jw < Q.warps[c]) yout[(size_t)(s * KC + i) * R] = y;
        }
      }
    }
  };
  auto clear_xbar = [&]() {
    xc_d = czero;
#pragma unroll
    for (int s = 0; s < W; ++s) xc[s] = czero;
  };
  cplx raw_c[NSL * KC];
  auto finish_cols = [&]() {
    a_d = slot_value(raw_c, m_col, add_off_col, 1, 0);
#pragma unroll
    for (int s = 0; s < W; ++s) a_col[s] = slot_value(raw_c, m_col, add_off_col, 1, 1 + s);
  };

  if constexpr (STORE != 0 && TMA != 0) {
    // ---- one bulk copy per slice: slab k = terms [n_k, n_k + q_k) of this warp's stream; two slabs in shared memory,
    //      the copy of slice k-1 is in flight while slice k is consumed
    const unsigned ring_u32 = smem_u32(ring), bar_u32 = smem_u32(bars);
    const unsigned slab_bytes = (unsigned)Q.qs * 32 * sizeof(cplx);
    const cplx* const gsrc = (const cplx*)P.terms + (gw * Q.cap) * 32;
    auto tma_issue = [&](int slot, unsigned n0, int q) {   // one elected lane: arm the barrier, start the bulk copy
      if (lane == 0) {
        const unsigned bar = bar_u32 + slot * 8, dst = ring_u32 + slot * slab_bytes, bytes = (unsigned)q * 32 * sizeof(cplx);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(gsrc + (size_t)n0 * 32), "r"(bytes), "r"(bar) : "memory");
      }
    };
    // the slabs were written with ordinary (generic-proxy) stores by the lanes of THIS warp; the bulk copies read them
    // through the async proxy: every writer fences, then the warp synchronises, then one lane issues
    asm volatile("fence.proxy.async.global;" ::: "memory");
    if (lane == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_u32) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_u32 + 8) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    int q0 = P.qdeg[P.M - 1];                            // degrees of slices k, k-1, k-2
    int q1 = P.M > 1 ? P.qdeg[P.M - 2] : 0;
    int q2 = P.M > 2 ? P.qdeg[P.M - 3] : 0;
    unsigned n0 = n_stream - q0;                         // first term of slice k
    tma_issue(0, n0, q0);
    if (P.M > 1) tma_issue(1, n0 - q1, q1);
    load_raw(p_col + (size_t)(P.M - 1) * P.ncc * L, nsc, raw_c);
    for (int k = P.M - 1, it = 0; k >= 0; --k, ++it) {
      const int q = q0, slot = it & 1;
      finish_cols();
      if (k > 0) load_raw(p_col + (size_t)(k - 1) * P.ncc * L, nsc, raw_c);
      lam = lam_tb;
      lam_tb = cscale(lam, Q.ifact[q]);                 // w_q = lam' / q!
      clear_xbar();
      {
        const unsigned bar = bar_u32 + slot * 8, parity = (it >> 1) & 1u;
        unsigned done = 0;
        while (!done)
          asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }"
                       : "=r"(done) : "r"(bar), "r"(parity) : "memory");
      }
      const cplx* slab = ring + (size_t)slot * Q.qs * 32 + lane;
      for (int j = q; j >= 1; --j) step_adj(slab[(j - 1) * 32], Q.ifact[j - 1]);
      emit(k);
      // this slab has been consumed (its values went through the DFMAs above): refill it with slice k-2
      const unsigned n1 = n0 - q1;
      __syncwarp();
      if (k >= 2) tma_issue(slot, n1 - q2, q2);
      n0 = n1; q0 = q1; q1 = q2;
      q2 = k >= 3 ? P.qdeg[k - 3] : 0;
    }
  } else if constexpr (STORE != 0) {
    // ---- per-thread cp.async ring, ROWS_RING terms ahead (the stream is consumed strictly backwards)
    const unsigned ring_u32 = smem_u32(ring);
    const cplx* sp = stream + (size_t)(n_stream - 1) * 32;
    int slot = 0, left = (int)n_stream;
    auto issue_async = [&]() {
      if (left > 0)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring_u32 + (slot * 32 + lane) * 16), "l"(sp) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
      --left;
      sp -= 32;
      slot = (slot + 1) & (ROWS_RING - 1);
    };
#pragma unroll
    for (int i = 0; i < ROWS_RING; ++i) ring[i * 32 + lane] = czero;
#pragma unroll
    for (int i = 0; i < ROWS_RING; ++i) issue_async();
    auto fetch = [&]() -> cplx {
      asm volatile("cp.async.wait_group %0;" ::"n"(ROWS_RING - 1) : "memory");
      const cplx v = ring[slot * 32 + lane];
      issue_async();
      return v;
    };
    int q_cur = P.qdeg[P.M - 1];
    q_next = P.M > 1 ? P.qdeg[P.M - 2] : 0;
    load_raw(p_col + (size_t)(P.M - 1) * P.ncc * L, nsc, raw_c);
    for (int k = P.M - 1; k >= 0; --k) {
      const int q = q_cur;
      finish_cols();
      if (k > 0) {
        load_raw(p_col + (size_t)(k - 1) * P.ncc * L, nsc, raw_c);
        q_cur = q_next;
        q_next = k > 1 ? P.qdeg[k - 2] : 0;
      }
      lam = lam_tb;
      lam_tb = cscale(lam, Q.ifact[q]);
      clear_xbar();
      for (int j = q; j >= 1; --j) step_adj(fetch(), Q.ifact[j - 1]);
      emit(k);
    }
  } else {
    cplx t_loc[QOCVL_QMAX];
    load_raw(p_row + (size_t)(P.M - 1) * P.ncr * L, nsr, raw_r);
    q_next = P.qdeg[P.M - 1];
    cplx z_next = __ldcs((const double2*)(ckpt + (size_t)(P.M - 1) * Q.total_warps * 32));
    for (int k = P.M - 1; k >= 0; --k) {
      const int q = q_next;
      finish_rows();
      load_raw(p_col + (size_t)k * P.ncc * L, nsc, raw_c);   // in flight during the recompute
      cplx u = z_next;
      t_loc[0] = u;
      for (int j = 1; j < q; ++j) {
        u = matvec(u);
        t_loc[j] = u;
      }
      finish_cols();                                        // same diagonal entry, plus my column's entries
      if (k > 0) {
        load_raw(p_row + (size_t)(k - 1) * P.ncr * L, nsr, raw_r);
        z_next = __ldcs((const double2*)(ckpt + (size_t)(k - 1) * Q.total_warps * 32));
        q_next = P.qdeg[k - 1];
      }
      lam = lam_tb;
      lam_tb = cscale(lam, Q.ifact[q]);
      clear_xbar();
      for (int j = q; j >= 1; --j) step_adj(t_loc[j - 1], Q.ifact[j - 1]);
      emit(k);
    }
  }
}

// One launch for every role: the CTAs of a cluster hold the roles of one sample group (role = cluster rank, a function of
// blockIdx), and each runs the body compiled for its role's slot depth.  WM = deepest role of the plan (3 or 6).
template <int WM, int KC, int STORE, int TMA, int SPLIT>
__global__ void __launch_bounds__(256, 2) k_chain_rows(const __grid_constant__ RowsArgs Q) {
  extern __shared__ __align__(128) unsigned char smem_rows[];
  const int c = blockIdx.x % Q.nroles;
  const int depth = max(Q.w[c], Q.wc[c]);
  if (depth <= 2) rows_body<2, KC, STORE, TMA, SPLIT>(Q, smem_rows, c);
  else if (WM == 3 || depth == 3) rows_body<3, KC, STORE, TMA, SPLIT>(Q, smem_rows, c);
  else if constexpr (WM == 6) rows_body<6, KC, STORE, TMA, SPLIT>(Q, smem_rows, c);
}

typedef void (*rows_fn)(const RowsArgs);
rows_fn qocvl_rows_kernel_k1(int wm, bool store, bool tma, bool split);   // chain_rows_k1.cu / chain_rows_k1w.cu (KC = 1)
rows_fn qocvl_rows_kernel_k1w(bool store, bool tma, bool split);
rows_fn qocvl_rows_kernel_k2(int wm, bool store, bool tma);               // chain_rows_k2.cu  (KC = 2, never SPLIT)
