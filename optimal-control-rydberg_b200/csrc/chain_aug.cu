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
// off-diagonals, assembled per slice from the per-sample coefficients) and element r of Z.  Off-row elements
// are gathered through a double-buffered shared-memory exchange guarded by __syncwarp only; Taylor terms of
// the reverse sweep sit in thread-local memory (L1); checkpoints Z_k go to HBM once per slice, coalesced.
#include "common.cuh"
#include <algorithm>
#include <cmath>

typedef std::complex<double> zc;

__constant__ double c_ainv[QOCVL_QMAX + 2];   // 1/j

struct AugArgs {
  int N, J, M, C, S;
  int w, wc, kc;             // off-diagonal row slots / column slots / generator contributions per slot
  int want_grad;
  double dt, inv_duration, w_infid, w_adiab, inv_fid_norm, inv_samples;
  cplx d_const;              // overlap contribution of the start kets that never move
  const int *row_col, *col_row, *rc_gen, *cc_gen;
  const cplx *rc_val, *cc_val;
  const cplx* coef;    // [M][J]
  const cplx* dcoef;   // [M][J][C]
  const int* qdeg;     // [M]
  const double* mul;   // [S][J]
  const cplx* add;     // [S][J]
  const cplx* starts;  // [LPG]
  const cplx* targets; // [LPG] (symmetry / orbit weights folded in)
  const int* pair;     // [LPG] partner row in q^dag p, or -1
  const double* wq;    // [2][LPG]: weight of the row's term in q^dag p (q rows only) / in the cotangent (q and p rows)
  cplx* ckpt;          // [M][total threads]
  double* sample_out;  // [S][3]
  double* gw_partial;  // [total warps][M][C]
};

template <int LPG>
__device__ __forceinline__ cplx aug_group_sum(cplx v) {
#pragma unroll
  for (int off = LPG / 2; off > 0; off >>= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, off);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, off);
  }
  return v;
}

template <int LPG, int WMAX>
__global__ void __launch_bounds__(512, 1) k_chain_aug(const AugArgs P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NCO = (QOCVL_MAX_GEN + LPG - 1) / LPG;   // coefficients a lane prepares per slice
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int grp = tid / LPG, r = tid % LPG, ngroups = nthreads / LPG;
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
  const int J = P.J, C = P.C, JC = P.J * P.C;
  const size_t gtid = (size_t)blockIdx.x * nthreads + tid, gthreads = (size_t)gridDim.x * nthreads;
  const int sample = blockIdx.x * ngroups + grp;
  const bool item_ok = sample < P.S;
  const bool live = item_ok && r < P.N;
  const int sm = item_ok ? sample : P.S - 1;

  // ---- shared memory carve-up
  const int n_rc = (P.w + 1) * P.kc * LPG, n_cc = (P.wc + 1) * P.kc * LPG;
  cplx* carve = (cplx*)smem_raw;
  cplx* s_rcv = carve; carve += n_rc;                               // [w+1][kc][LPG]  row-owner contributions
  cplx* s_ccv = carve; carve += n_cc;                               // [wc+1][kc][LPG] column-owner contributions
  cplx* s_cs = carve + (size_t)grp * J; carve += (size_t)ngroups * J;      // [ngroups][J] slice coefficients
  cplx* s_dco = carve + (size_t)warp * JC; carve += (size_t)nwarps * JC;   // [nwarps][J][C] d coef / d wave
  cplx* s_x = carve + (size_t)grp * 2 * LPG; carve += (size_t)ngroups * 2 * LPG;   // [ngroups][2 buffers][LPG]
  double* s_mul = (double*)carve + (size_t)grp * J;                 // [ngroups][J]
  int* s_rcg = (int*)((double*)carve + (size_t)ngroups * J);
  int* s_ccg = s_rcg + n_rc;
  for (int i = tid; i < n_rc; i += nthreads) { s_rcv[i] = P.rc_val[i]; s_rcg[i] = P.rc_gen[i]; }
  for (int i = tid; i < n_cc; i += nthreads) { s_ccv[i] = P.cc_val[i]; s_ccg[i] = P.cc_gen[i]; }
  for (int j = r; j < J; j += LPG) s_mul[j] = P.mul[(size_t)sm * J + j];
  __syncthreads();                                                  // the only CTA-wide barrier

  int rcol[WMAX], crow[WMAX];
#pragma unroll
  for (int s = 0; s < WMAX; ++s) {
    rcol[s] = (s < P.w) ? P.row_col[(1 + s) * LPG + r] : r;
    crow[s] = (s < P.wc) ? P.col_row[(1 + s) * LPG + r] : r;
  }
  int rcnt = 0, ccnt = 0;     // real entries of my row / column (they form a prefix of the slots)
  for (int s = 0; s < P.w; ++s) rcnt += s_rcg[((1 + s) * P.kc) * LPG + r] >= 0;
  for (int s = 0; s < P.wc; ++s) ccnt += s_ccg[((1 + s) * P.kc) * LPG + r] >= 0;
  double mymul[NCO];
  cplx myadd[NCO];
#pragma unroll
  for (int i = 0; i < NCO; ++i) {
    const int j = r + i * LPG;
    mymul[i] = (j < J) ? P.mul[(size_t)sm * J + j] : 0.0;
    myadd[i] = (j < J) ? P.add[(size_t)sm * J + j] : cmake(0, 0);
  }
  const cplx czero = cmake(0.0, 0.0);
  cplx z = live ? P.starts[r] : czero;
  const cplx ytgt = live ? P.targets[r] : czero;
  const int pr = live ? P.pair[r] : -1;
  const double w_qp = live ? P.wq[r] : 0.0, w_lam = live ? P.wq[LPG + r] : 0.0;
  int ph = 0;                                                       // exchange buffer offset: 0 or LPG
  cplx t_loc[QOCVL_QMAX];
  cplx a_d, a_row[WMAX];

  // exchange layout: re[LPG] then im[LPG]; 8-byte accesses of the LPG lanes of a sample hit distinct banks
  auto xput = [&](double* xs, cplx val) { xs[r] = val.x; xs[LPG + r] = val.y; };
  auto xget = [&](const double* xs, int idx) -> cplx { return cmake(xs[idx], xs[LPG + idx]); };
  auto load_coef = [&](int kk, cplx* creg) {
#pragma unroll
    for (int i = 0; i < NCO; ++i) {
      const int j = r + i * LPG;
      creg[i] = (j < J) ? P.coef[(size_t)kk * J + j] : czero;
    }
  };
  auto store_cs = [&](const cplx* creg) {
#pragma unroll
    for (int i = 0; i < NCO; ++i) {
      const int j = r + i * LPG;
      if (j < J) s_cs[j] = cadd(cscale(creg[i], mymul[i]), myadd[i]);
    }
  };
  auto slot_value = [&](const int* gen, const cplx* val, int slot) -> cplx {
    cplx acc = czero;
    for (int kc = 0; kc < P.kc; ++kc) {
      const int gi = gen[(slot * P.kc + kc) * LPG + r];
      if (gi >= 0) acc = cfma(s_cs[gi], val[(slot * P.kc + kc) * LPG + r], acc);
    }
    return cscale(acc, P.dt);
  };
  auto assemble_rows = [&]() {
    a_d = slot_value(s_rcg, s_rcv, 0);
#pragma unroll
    for (int s = 0; s < WMAX; ++s) a_row[s] = (s < P.w) ? slot_value(s_rcg, s_rcv, 1 + s) : czero;
  };
  // one forward Taylor step: t <- Xaug t / j
  auto step_fwd = [&](cplx t, double invj) -> cplx {
    double* xs = (double*)(s_x + ph);
    xput(xs, t);
    __syncwarp();
    cplx acc = cmul(a_d, t);
#pragma unroll
    for (int s = 0; s < WMAX; ++s) {
      const cplx xv = (s < rcnt) ? xget(xs, rcol[s]) : czero;
      acc = cfma(a_row[s], xv, acc);
    }
    ph = LPG - ph;
    return cscale(acc, invj);
  };

  {
    cplx creg[NCO];
    load_coef(0, creg);
    store_cs(creg);
    __syncwarp();
  }
  // =================================== forward sweep ===================================
  for (int k = 0; k < P.M; ++k) {
    const int q = P.qdeg[k];
    assemble_rows();
    __syncwarp();                                   // everyone has read cs of slice k
    cplx creg[NCO];
    const bool more = k + 1 < P.M;
    if (more) load_coef(k + 1, creg);               // in flight during this slice
    if (P.want_grad && live) P.ckpt[(size_t)k * gthreads + gtid] = z;
    cplx t = z;
    for (int j = 1; j <= q; ++j) {
      t = step_fwd(t, c_ainv[j]);
      z = cadd(z, t);
    }
    if (more) store_cs(creg);
    __syncwarp();
  }

  // =================================== cost ============================================
  cplx zp;
  {
    double* xs = (double*)(s_x + ph);
    xput(xs, z);
    __syncwarp();
    zp = (pr >= 0) ? xget(xs, pr) : czero;
    ph = LPG - ph;
  }
  const cplx dsum = cadd(P.d_const, aug_group_sum<LPG>(cfma_conj(ytgt, z, czero)));   // sum_r conj(y_r) z_r
  {
    const cplx qp = aug_group_sum<LPG>(cscale(cfma_conj(z, zp, czero), w_qp));         // q^dag p
    if (r == 0 && item_ok) {
      const double infid = 1.0 - (dsum.x * dsum.x + dsum.y * dsum.y) * P.inv_fid_norm;
      const double adiab = 1.0 - qp.x * P.inv_duration;
      P.sample_out[(size_t)sample * 3 + 0] = P.w_infid * infid + P.w_adiab * adiab;
      P.sample_out[(size_t)sample * 3 + 1] = infid;
      P.sample_out[(size_t)sample * 3 + 2] = adiab;
    }
  }
  if (!P.want_grad) return;

  // terminal cotangent lam = d cost / d conj(Z), scaled by 1/S_global (ensemble mean)
  cplx lam;
  {
    const double ws = live ? P.inv_samples : 0.0;
    const double ca = -0.5 * P.w_adiab * P.inv_duration * ws;
    lam = cscale(cmul(dsum, ytgt), -P.w_infid * P.inv_fid_norm * ws);
    lam = cadd(lam, cscale(zp, ca * w_lam));        // q rows receive -(wa/2T) p, p rows -(wa/2T) q
  }
  __syncwarp();
  {
    cplx creg[NCO];
    load_coef(P.M - 1, creg);
    store_cs(creg);
    __syncwarp();
  }

  // =================================== reverse sweep ===================================
  const size_t gwarp = (size_t)blockIdx.x * nwarps + warp;
  for (int k = P.M - 1; k >= 0; --k) {
    const int q = P.qdeg[k];
    assemble_rows();
    cplx a_col[WMAX];                               // column-owner values Xaug[a][r] for the rows a of my column
#pragma unroll
    for (int s = 0; s < WMAX; ++s) a_col[s] = (s < P.wc) ? slot_value(s_ccg, s_ccv, 1 + s) : czero;
    __syncwarp();                                   // everyone has read cs of slice k
    cplx creg[NCO];
    if (k > 0) load_coef(k - 1, creg);
    cplx dreg[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) dreg[i] = (lane + 32 * i < JC) ? P.dcoef[(size_t)k * JC + lane + 32 * i] : czero;
    // ---- recompute the forward Taylor terms t_0 .. t_{q-1}
    cplx t = live ? P.ckpt[(size_t)k * gthreads + gtid] : czero;
    t_loc[0] = t;
    for (int j = 1; j < q; ++j) {
      t = step_fwd(t, c_ainv[j]);
      t_loc[j] = t;
    }
    // ---- descending adjoint recurrence with on-the-fly contraction
    cplx tb = lam;
    cplx xc_d = czero, xc[WMAX];
#pragma unroll
    for (int s = 0; s < WMAX; ++s) xc[s] = czero;
    for (int j = q; j >= 1; --j) {
      const double invj = c_ainv[j];
      double* xs = (double*)(s_x + ph);
      xput(xs, tb);
      __syncwarp();
      const cplx tprev = cscale(t_loc[j - 1], invj);
      cplx acc = cfma_conj(a_d, tb, czero);
      xc_d = cfma_conj(tb, tprev, xc_d);
#pragma unroll
      for (int s = 0; s < WMAX; ++s) {
        const cplx ta = (s < ccnt) ? xget(xs, crow[s]) : czero;
        acc = cfma_conj(a_col[s], ta, acc);
        xc[s] = cfma_conj(ta, tprev, xc[s]);
      }
      tb = cadd(lam, cscale(acc, invj));
      ph = LPG - ph;
    }
    lam = tb;
    // ---- contract Xbar with the generators: d cost / d wave[k][c]
#pragma unroll
    for (int i = 0; i < 3; ++i) if (lane + 32 * i < JC) s_dco[lane + 32 * i] = dreg[i];
    __syncwarp();
    double gw[QOCVL_MAX_CHAN];
#pragma unroll
    for (int c = 0; c < QOCVL_MAX_CHAN; ++c) gw[c] = 0.0;
    if (live) {
      const double two_dt = 2.0 * P.dt;
      for (int slot = 0; slot <= P.wc; ++slot) {
        cplx xval = xc_d;
#pragma unroll
        for (int s = 0; s < WMAX; ++s)
          if (slot == s + 1) xval = xc[s];
        for (int kc = 0; kc < P.kc; ++kc) {
          const int gi = s_ccg[(slot * P.kc + kc) * LPG + r];
          if (gi >= 0) {
            const cplx tmp = cscale(cmul(s_ccv[(slot * P.kc + kc) * LPG + r], xval), two_dt * s_mul[gi]);
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
      for (int off = 16; off > 0; off >>= 1) vsum += __shfl_xor_sync(0xffffffffu, vsum, off);
      if (lane == 0 && c < C) P.gw_partial[(gwarp * P.M + k) * C + c] = vsum;
    }
    if (k > 0) store_cs(creg);
    __syncwarp();
  }
}

typedef void (*aug_fn)(const AugArgs);
static aug_fn pick_aug(int lpg, int wmax) {
  if (wmax <= 3) {
    switch (lpg) {
      case 4: return k_chain_aug<4, 3>;
      case 8: return k_chain_aug<8, 3>;
      case 16: return k_chain_aug<16, 3>;
      case 32: return k_chain_aug<32, 3>;
    }
  } else if (wmax <= 6) {
    switch (lpg) {
      case 4: return k_chain_aug<4, 6>;
      case 8: return k_chain_aug<8, 6>;
      case 16: return k_chain_aug<16, 6>;
      case 32: return k_chain_aug<32, 6>;
    }
  }
  return nullptr;
}

static size_t aug_smem_bytes(const qocvl_plan* p, int ngroups) {
  const int J = p->d.n_gen, C = p->d.n_chan, L = p->aug_lpg;
  const int nwarps = std::max(1, ngroups * L / 32);
  const size_t tab = (size_t)(p->la.n_row_slots + p->la.n_col_slots) * p->la.kc * L;
  const size_t cplx_count = tab + (size_t)ngroups * J + (size_t)nwarps * J * C + (size_t)ngroups * 2 * L;
  return cplx_count * sizeof(cplx) + (size_t)ngroups * J * sizeof(double) + tab * sizeof(int);
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
    if (getenv("QOCVL_NOFOLD")) fold = false;
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
  const int L = N <= 4 ? 4 : (N <= 8 ? 8 : (N <= 16 ? 16 : 32));
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
  int rc = qocvl_build_lane_layout(p, aug, N, L, true, &p->la);
  if (rc) return rc;
  const int wmax = std::max(p->la.n_row_slots, p->la.n_col_slots) - 1;
  if (wmax > 6 || J * d.n_chan > 96) return QOCVL_ERR_UNSUPPORTED;
  p->aug_n = N; p->aug_lpg = L; p->aug_nnz = p->la.nnz; p->aug_wmax = wmax;
  aug_fn fn = pick_aug(L, wmax);
  if (!fn) return QOCVL_ERR_UNSUPPORTED;
  for (void** q : {(void**)&p->d_aug_starts, (void**)&p->d_aug_targets, (void**)&p->d_aug_pair, (void**)&p->d_aug_wq})
    if (*q) { cudaFree(*q); *q = nullptr; }
  QCUDA(cudaMalloc((void**)&p->d_aug_starts, L * sizeof(cplx)));
  QCUDA(cudaMalloc((void**)&p->d_aug_targets, L * sizeof(cplx)));
  QCUDA(cudaMalloc((void**)&p->d_aug_pair, L * sizeof(int)));
  QCUDA(cudaMalloc((void**)&p->d_aug_wq, 2 * L * sizeof(double)));
  QCUDA(cudaMemcpy(p->d_aug_starts, st.data(), L * sizeof(cplx), cudaMemcpyHostToDevice));
  QCUDA(cudaMemcpy(p->d_aug_targets, tg.data(), L * sizeof(cplx), cudaMemcpyHostToDevice));
  QCUDA(cudaMemcpy(p->d_aug_pair, pair.data(), L * sizeof(int), cudaMemcpyHostToDevice));
  QCUDA(cudaMemcpy(p->d_aug_wq, wq.data(), 2 * L * sizeof(double), cudaMemcpyHostToDevice));
  {
    double inv[QOCVL_QMAX + 2];
    inv[0] = 0.0;
    for (int j = 1; j < QOCVL_QMAX + 2; ++j) inv[j] = 1.0 / (double)j;
    QCUDA(cudaMemcpyToSymbol(c_ainv, inv, sizeof(inv)));
  }
  // ---- launch geometry: a CTA is only a container of independent groups; pick the number of samples per CTA
  //      that minimises the load of the busiest SM (wave quantisation), as in chain_small.cu
  const int S = p->n_samples, gpw = 32 / L;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device);
  const int gmax = 512 / L / gpw * gpw;
  int ngroups = gpw;
  {
    long best_load = -1;
    for (int g = gpw; g <= gmax; g += gpw) {
      const long n_ctas = (S + g - 1) / g;
      const long load = (n_ctas + sms - 1) / sms * g;
      if (best_load < 0 || load < best_load || (load == best_load && g * L <= 256 && g > ngroups)) {
        best_load = load; ngroups = g;
      }
    }
  }
  if (const char* e = getenv("QOCVL_SPB")) ngroups = std::max(gpw, std::min(gmax, (atoi(e) + gpw - 1) / gpw * gpw));
  ngroups = std::min(ngroups, (S + gpw - 1) / gpw * gpw);
  while (ngroups > gpw && aug_smem_bytes(p, ngroups) > 200 * 1024) ngroups -= gpw;
  p->spb = ngroups; p->groups = ngroups; p->threads = ngroups * L;
  p->chain_smem = aug_smem_bytes(p, ngroups);
  if (p->chain_smem > 227 * 1024) return QOCVL_ERR_UNSUPPORTED;
  p->blocks = (S + ngroups - 1) / ngroups;
  QCUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->chain_smem));
  const size_t total_threads = (size_t)p->blocks * p->threads;
  p->n_parts = (total_threads + 31) / 32;
  const size_t ck = (size_t)d.n_slices * total_threads;
  if (p->d_ckpt) { cudaFree(p->d_ckpt); p->d_ckpt = nullptr; p->bytes -= p->ckpt_bytes; }
  QCUDA(cudaMalloc((void**)&p->d_ckpt, ck * sizeof(cplx)));
  p->ckpt_bytes = ck * sizeof(cplx);
  p->bytes += p->ckpt_bytes;
  p->qcap_apriori = p->qcap;
  p->qcap = QOCVL_QMAX;
  p->use_aug = true;
  return QOCVL_OK;
}

int qocvl_aug_launch(qocvl_plan* p, bool want_grad, cudaStream_t st) {
  const qocvl_desc& d = p->d;
  AugArgs a;
  a.N = p->aug_n; a.J = d.n_gen; a.M = d.n_slices; a.C = d.n_chan; a.S = p->n_samples;
  a.w = p->la.n_row_slots - 1; a.wc = p->la.n_col_slots - 1; a.kc = p->la.kc;
  a.want_grad = want_grad ? 1 : 0;
  a.dt = d.dt; a.inv_duration = 1.0 / d.duration; a.w_infid = d.w_infid; a.w_adiab = d.w_adiab;
  a.inv_fid_norm = 1.0 / d.fid_norm; a.inv_samples = 1.0 / (double)p->global_samples;
  a.d_const = p->d_const;
  a.row_col = p->la.row_col; a.col_row = p->la.col_row;
  a.rc_gen = p->la.rc_gen; a.cc_gen = p->la.cc_gen; a.rc_val = p->la.rc_val; a.cc_val = p->la.cc_val;
  a.coef = p->d_coef; a.dcoef = p->d_dcoef; a.qdeg = p->d_qdeg; a.mul = p->d_mul; a.add = p->d_add;
  a.starts = p->d_aug_starts; a.targets = p->d_aug_targets; a.pair = p->d_aug_pair; a.wq = p->d_aug_wq;
  a.ckpt = p->d_ckpt; a.sample_out = p->d_sample_out; a.gw_partial = p->d_gw_partial;
  pick_aug(p->aug_lpg, p->aug_wmax)<<<p->blocks, p->threads, p->chain_smem, st>>>(a);
  p->launches++;
  QCUDA(cudaGetLastError());
  return QOCVL_OK;
}
