// chain_duo.cu -- host side of the two-lanes-per-sector ensemble kernels (chain_duo.cuh): the search for the cut of
// every role, the tables, the launch sequence; and the three small kernels around the sweeps
// (k_duo_tables, k_duo_terminal, k_duo_reduce).
#include "chain_duo.cuh"
#include "model.cuh"
#include <algorithm>
#include <cmath>
#include <numeric>

typedef std::complex<double> zc;

// P_u(k) = dt * sum_g coef_kg * val_ug for every (role, lane type, slot): the sample-independent part of the slice
// generators, once per evaluation.  One block per slice.
template <typename CT>
__global__ void k_duo_tables(const DuoArgs Q, const unsigned long long* __restrict__ ent_out,
                             const int* __restrict__ ent_stride) {
  const int k = blockIdx.x;
  for (int d = threadIdx.x; d < Q.ntot; d += blockDim.x) {
    cplx acc = cmake(0.0, 0.0);
    for (int i = 0; i < Q.KG; ++i) {
      const int g = Q.ent_gen[d * Q.KG + i];
      if (g >= 0) acc = cfma(Q.coef[(size_t)k * Q.J + g], Q.ent_val[d * Q.KG + i], acc);
    }
    acc = cscale(acc, Q.dt);
    if (Q.phi && Q.ent_srole[d] >= 0) acc.y -= Q.phi[(size_t)Q.ent_srole[d] * Q.M + k];   // Y = X - i phi_role I
    // (stored in the sweep's own type: a conversion inside the sweep would sit right behind the staged load)
    ((CT*)Q.tab)[ent_out[d] + (size_t)k * ent_stride[d]] = duo_from<CT>(acc);
  }
}

// Accumulated phase of the per-role shifts: the sweeps propagate z~ = e^{-i Phi_r} z on the rows of role r, Phi_r =
// sum_k phi_r(k).  The cost reads z only through <y, z> (and q^dag p inside one role: phase free), so the role's target
// kets are rotated instead, y~ = e^{-i Phi_r} y: <y~, z~> = <y, z>, and the terminal cotangent k_duo_terminal derives from
// y~ is the cotangent of z~.  One block, fixed summation order (deterministic).
__global__ void __launch_bounds__(256) k_duo_phase(int M, int N, int nroles, const double* __restrict__ phi,
                                                   const int* __restrict__ rowrole, const cplx* __restrict__ targets,
                                                   cplx* __restrict__ rot) {
  __shared__ double part[256];
  __shared__ double total[DUO_MAXR];
  for (int r = 0; r < nroles; ++r) {
    double acc = 0.0;
    for (int k = threadIdx.x; k < M; k += 256) acc += phi[(size_t)r * M + k];
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
      if (threadIdx.x < off) part[threadIdx.x] += part[threadIdx.x + off];
      __syncthreads();
    }
    if (threadIdx.x == 0) total[r] = part[0];
    __syncthreads();
  }
  for (int i = threadIdx.x; i < N; i += 256) {
    const int r = rowrole[i];
    cplx y = targets[i];
    if (r >= 0) {
      double sn, cs;
      sincos(-total[r], &sn, &cs);
      y = cmul(cmake(cs, sn), y);
    }
    rot[i] = y;
  }
}

// Per-sample cost (TQ:549-574 / ST:285-310 on the stacked vector) and terminal cotangent lam = d cost / d conj(Z),
// scaled by 1 / S_global (ensemble mean).  One thread per sample.
__global__ void k_duo_terminal(const DuoArgs Q) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= Q.S) return;
  const cplx* z = Q.zfin + (size_t)s * Q.N;
  cplx dsum = Q.d_const;
  double qp = 0.0;
  for (int r = 0; r < Q.N; ++r) {
    const cplx zr = z[r];
    dsum = cfma_conj(Q.targets[r], zr, dsum);
    const int pr = Q.pair[r];
    if (pr >= 0) {
      const cplx zp = z[pr];
      qp += Q.wq[r] * (zr.x * zp.x + zr.y * zp.y);
    }
  }
  const double infid = 1.0 - (dsum.x * dsum.x + dsum.y * dsum.y) * Q.inv_fid_norm;
  const double adiab = 1.0 - qp * Q.inv_duration;
  Q.sample_out[(size_t)s * 3 + 0] = Q.w_infid * infid + Q.w_adiab * adiab;
  Q.sample_out[(size_t)s * 3 + 1] = infid;
  Q.sample_out[(size_t)s * 3 + 2] = adiab;
  if (!Q.want_grad) return;
  const double ws = Q.inv_samples, ca = -0.5 * Q.w_adiab * Q.inv_duration * ws;
  for (int r = 0; r < Q.N; ++r) {
    cplx l = cscale(cmul(dsum, Q.targets[r]), -Q.w_infid * Q.inv_fid_norm * ws);
    const int pr = Q.pair[r];
    if (pr >= 0) l = cadd(l, cscale(z[pr], ca * Q.wq[Q.N + r]));   // q rows receive -(wa/2T) p, p rows -(wa/2T) q
    Q.lam[(size_t)s * Q.N + r] = l;
  }
}

// Time-parallel evaluation (mode 1 / 2 of the sweeps): walk the chunk propagators of one (sample, role).
//   dir = +1: z_0 = start kets, zstart[c] = z_c, z_{c+1} = P_c z_c; the last state goes to zfin (-> k_duo_terminal).
//   dir = -1: lam_C = terminal cotangent, lamend[c] = lam_{c+1}, lam_c = P_c^dag lam_{c+1}   (P_c^dag is exactly what
//             the reverse sweep applies slice by slice: the adjoint of the same truncated Taylor polynomials).
// Roles partition the stacked rows and never mix, so each (sample, role) is walked by its own group of 8 lanes (6 live:
// one local row each, the vector exchanged with shuffles); a role's propagator is a 6x6 block stored by column.
// Mirrored roles ([[U, 0], [W, U]]): only the independent triple's columns were swept; the dependent triple's own block
// is the same U.
__device__ __forceinline__ cplx duo_pfull(const DuoArgs& Q, const cplx* P, const int* colof, int role, int li, int lj) {
  if (colof[lj] >= 0) return P[colof[lj] * 6 + li];                 // swept column
  if (Q.mirror[role] && li / 3 == lj / 3) {                          // dependent triple's own block = independent one's
    const int sh = lj < 3 ? 3 : -3;
    return P[colof[lj + sh] * 6 + li + sh];
  }
  return cmake(0.0, 0.0);
}

__global__ void __launch_bounds__(128) k_duo_combine(const DuoArgs Q, int dir) {
  const int grp = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, li = threadIdx.x & 7;
  const int s = grp / Q.nroles, role = grp % Q.nroles;
  const bool on = s < Q.S && li < 6;
  const int sm = s < Q.S ? s : Q.S - 1;
  const int base = (threadIdx.x & 31) & ~7;                           // first lane of my group inside the warp
  int colof[6];                                                       // swept column of local row l, or -1
#pragma unroll
  for (int l = 0; l < 6; ++l) colof[l] = -1;
  for (int jc = 0; jc < Q.ncol[role]; ++jc) {
    const int l = Q.colrow[role * 6 + jc];
#pragma unroll
    for (int l2 = 0; l2 < 6; ++l2) if (l2 == l) colof[l2] = jc;
  }
  const int myrow = li < 6 ? Q.srow[role * 6 + li] : -1;
  cplx z = cmake(0.0, 0.0);
  if (on && myrow >= 0) z = dir > 0 ? Q.starts[myrow] : Q.lam[(size_t)sm * Q.N + myrow];
  // my six propagator entries of a chunk -- row li of P (forward) or column li (backward) -- are requested one chunk
  // ahead: the walk is a chain of `chunks` dependent steps, and with the loads inside the step every step waited for six
  // L2 round trips in a row (61 us of a 240 us single-sample evaluation for each of the two walks).
  auto fetch = [&](int cc, cplx (&pe)[6]) {
    const int c = dir > 0 ? cc : Q.chunks - 1 - cc;
    const cplx* P = Q.prop + (((size_t)sm * Q.chunks + c) * Q.nroles + role) * 36;
#pragma unroll
    for (int l = 0; l < 6; ++l)
      pe[l] = li < 6 ? (dir > 0 ? duo_pfull(Q, P, colof, role, li, l) : duo_pfull(Q, P, colof, role, l, li)) : cmake(0.0, 0.0);
  };
  cplx pe[6], pn[6];
  fetch(0, pe);
  for (int cc = 0; cc < Q.chunks; ++cc) {
    const int c = dir > 0 ? cc : Q.chunks - 1 - cc;
    if (cc + 1 < Q.chunks) fetch(cc + 1, pn);
    if (on && myrow >= 0) ((dir > 0 ? Q.zstart : Q.lamend) + ((size_t)sm * Q.chunks + c) * Q.N)[myrow] = z;
    cplx zn = cmake(0.0, 0.0);
#pragma unroll
    for (int l = 0; l < 6; ++l) {
      const cplx zl = cmake(__shfl_sync(0xffffffffu, z.x, base + l), __shfl_sync(0xffffffffu, z.y, base + l));
      if (dir > 0) zn = cfma(pe[l], zl, zn);                 // z'_li = sum_l P[li][l] z_l
      else zn = cfma_conj(pe[l], zl, zn);                    // lam'_li = sum_l conj(P[l][li]) lam_l
    }
    z = zn;
    if (cc + 1 < Q.chunks) {
#pragma unroll
      for (int l = 0; l < 6; ++l) pe[l] = pn[l];
    }
  }
  if (dir > 0 && on && myrow >= 0) Q.zfin[(size_t)sm * Q.N + myrow] = z;
}

// gwave[k][c] = 2 dt Re sum_x Ybar_x(k) * sum_{g in x} val_xg * dcoef_kgc, with Ybar summed over the per-warp partials
// in a fixed order (deterministic).  One block of 256 threads per slice: 64 descriptor slots x 4 interleaved parts.
__global__ void __launch_bounds__(256) k_duo_reduce(const DuoArgs Q, const unsigned long long* __restrict__ ent_ysrc,
                                                     const int* __restrict__ ent_ystride, const int* __restrict__ ent_ywarps) {
  __shared__ cplx part[4][64];
  __shared__ double red[QOCVL_MAX_CHAN][64];
  const int k = blockIdx.x, d = threadIdx.x & 63, pt = threadIdx.x >> 6;
  const bool on = d < Q.ntot && Q.ent_x[d] >= 0;
  cplx y0 = cmake(0.0, 0.0), y1 = cmake(0.0, 0.0);
  if (on) {
    const int stride = ent_ystride[d], nw = ent_ywarps[d];
    const cplx* src = Q.ypart + ent_ysrc[d] + (size_t)k * stride;
    const size_t wstride = (size_t)Q.M * stride;
    int wv = pt;
    for (; wv + 4 < nw; wv += 8) {
      y0 = cadd(y0, src[(size_t)wv * wstride]);
      y1 = cadd(y1, src[(size_t)(wv + 4) * wstride]);
    }
    if (wv < nw) y0 = cadd(y0, src[(size_t)wv * wstride]);
  }
  part[pt][d] = cadd(y0, y1);
  __syncthreads();
  if (pt == 0) {
    double term[QOCVL_MAX_CHAN];
#pragma unroll
    for (int c = 0; c < QOCVL_MAX_CHAN; ++c) term[c] = 0.0;
    if (on) {
      const cplx y = cadd(cadd(part[0][d], part[1][d]), cadd(part[2][d], part[3][d]));
      for (int i = 0; i < Q.KG; ++i) {
        const int g = Q.ent_gen[d * Q.KG + i];
        if (g < 0) continue;
        const cplx yv = cmul(y, Q.ent_val[d * Q.KG + i]);
#pragma unroll
        for (int c = 0; c < QOCVL_MAX_CHAN; ++c)
          if (c < Q.C) {
            const cplx dc = Q.dcoef[((size_t)k * Q.J + g) * Q.C + c];
            term[c] += yv.x * dc.x - yv.y * dc.y;
          }
      }
    }
#pragma unroll
    for (int c = 0; c < QOCVL_MAX_CHAN; ++c) red[c][d] = term[c];
  }
  __syncthreads();
  if (threadIdx.x < Q.C) {
    double acc = 0.0;
    for (int i = 0; i < 64; ++i) acc += red[threadIdx.x][i];
    Q.gwave[(size_t)k * Q.C + threadIdx.x] = 2.0 * Q.dt * acc;
  }
}

// (three CTAs per SM: the (sample group, chunk, column) items of the time-parallel form are many)
template <typename CT>
__global__ void __launch_bounds__(128, 3) k_duo_forward(const DuoArgs Q) {
  extern __shared__ __align__(128) unsigned char duo_smem[];
  const int role = blockIdx.x % Q.nroles;
  switch (Q.variant[role]) {
    case 0: case 3: duo_forward<DuoPat0, CT>(Q, role, duo_smem); break;   // (the forward sweep has no Xbar: lean = full)
    case 1: case 4: duo_forward<DuoPat1, CT>(Q, role, duo_smem); break;
    default: duo_forward<DuoPat2, CT>(Q, role, duo_smem); break;
  }
}

// chunk propagators: DUO_COLS basis columns per item (two CTAs per SM)
template <typename CT>
__global__ void __launch_bounds__(128, 2) k_duo_forward_cols(const DuoArgs Q) {
  extern __shared__ __align__(128) unsigned char duo_smem[];
  const int role = blockIdx.x % Q.nroles;
  switch (Q.variant[role]) {
    case 0: case 3: duo_forward_cols<DuoPat0, CT, DUO_COLS>(Q, role, duo_smem); break;
    case 1: case 4: duo_forward_cols<DuoPat1, CT, DUO_COLS>(Q, role, duo_smem); break;
    default: duo_forward_cols<DuoPat2, CT, DUO_COLS>(Q, role, duo_smem); break;
  }
}

template <typename CT>
__global__ void __launch_bounds__(128, 1) k_duo_reverse(const DuoArgs Q) {
  extern __shared__ __align__(128) unsigned char duo_smem[];
  const int role = blockIdx.x % Q.nroles;
  switch (Q.variant[role]) {
    case 0: duo_reverse<DuoPat0, CT>(Q, role, duo_smem); break;
    case 1: duo_reverse<DuoPat1, CT>(Q, role, duo_smem); break;
    case 3: duo_reverse<DuoPat3, CT>(Q, role, duo_smem); break;
    case 4: duo_reverse<DuoPat4, CT>(Q, role, duo_smem); break;
    default: duo_reverse<DuoPat2, CT>(Q, role, duo_smem); break;
  }
}

// ---------------------------------------------------------------------------------------------------
// Host
// ---------------------------------------------------------------------------------------------------
namespace {
struct PatInfo { unsigned tm, cm, xm; int s0, s1, nt, nc, lean; };   // lean: the pattern's twin without diagonal Xbar, or -1
const PatInfo kPat[DUO_NVAR] = {
    {(unsigned)DuoPat0::TM, (unsigned)DuoPat0::CM, (unsigned)DuoPat0::XM, DuoPat0::S0, DuoPat0::S1, DuoPat0::NT, DuoPat0::NC, 3},
    {(unsigned)DuoPat1::TM, (unsigned)DuoPat1::CM, (unsigned)DuoPat1::XM, DuoPat1::S0, DuoPat1::S1, DuoPat1::NT, DuoPat1::NC, 4},
    {(unsigned)DuoPat2::TM, (unsigned)DuoPat2::CM, (unsigned)DuoPat2::XM, DuoPat2::S0, DuoPat2::S1, DuoPat2::NT, DuoPat2::NC, -1},
    {(unsigned)DuoPat3::TM, (unsigned)DuoPat3::CM, (unsigned)DuoPat3::XM, DuoPat3::S0, DuoPat3::S1, DuoPat3::NT, DuoPat3::NC, -1},
    {(unsigned)DuoPat4::TM, (unsigned)DuoPat4::CM, (unsigned)DuoPat4::XM, DuoPat4::S0, DuoPat4::S1, DuoPat4::NT, DuoPat4::NC, -1}};
const int kSearchVariants = 3;                            // the cut search only considers the full patterns

// compact slot index of a bit inside a mask
int slot_of(unsigned mask, int bit) { return duo_popc(mask & ((1u << bit) - 1u)); }
}  // namespace

void DuoState::release() {
  for (void* q : {(void*)d_ent_x_lean, (void*)d_ent_ystride_lean, (void*)d_ent_ysrc_lean}) if (q) cudaFree(q);
  void* ptrs[] = {d_phi, d_rowrole, d_ent_srole, d_targets_rot, d_colrole, d_qdeg_role, d_tab, d_con_a, d_srow, d_ent_gen, d_ent_val, d_ent_x, d_ent_out, d_ent_stride, d_ent_ysrc,
                  d_ent_ystride, d_ent_ywarps, d_starts, d_targets, d_pair, d_wq, d_zfin, d_lam, d_colrow, d_prop,
                  d_zstart, d_lamend};
  for (void* q : ptrs) if (q) cudaFree(q);
  *this = DuoState();
}

// Called by qocvl_aug_configure once the stacked system exists on the host (p->h_aug_*).  Returns
// QOCVL_ERR_UNSUPPORTED when some role does not split into two triples coupled through <= 2 rows each (or the system
// needs two multiplier classes per entry / complex64): the older kernels then run.
int qocvl_duo_configure(qocvl_plan* p) {
  DuoState& D = p->duo;
  D.release();
  p->use_duo = false;
  const int esz = p->c64 ? (int)sizeof(cplxf) : (int)sizeof(cplx);   // element of the sweeps (qocvl_set_precision)
  const int N = p->aug_n, J = p->d.n_gen, M = p->d.n_slices, S = p->n_samples;
  const std::vector<zc>& aug = p->h_aug_gen;
  if ((int)aug.size() != J * N * N || N < 1) return QOCVL_ERR_UNSUPPORTED;
  auto AG = [&](int j, int a, int b) { return aug[((size_t)j * N + a) * N + b]; };
  std::vector<char> pat((size_t)N * N, 0);
  for (int j = 0; j < J; ++j)
    for (int a = 0; a < N; ++a)
      for (int b = 0; b < N; ++b)
        if (std::abs(AG(j, a, b)) > 0.0) pat[(size_t)a * N + b] = 1;
  // ---- roles: connected components of the pattern (undirected)
  std::vector<int> parent(N);
  std::iota(parent.begin(), parent.end(), 0);
  auto find = [&](int x) { while (parent[x] != x) x = parent[x] = parent[parent[x]]; return x; };
  for (int a = 0; a < N; ++a)
    for (int b = 0; b < N; ++b)
      if (pat[(size_t)a * N + b]) {
        const int ra = find(a), rb = find(b);
        if (ra != rb) parent[std::max(ra, rb)] = std::min(ra, rb);
      }
  std::vector<std::vector<int>> comp;
  {
    std::vector<int> comp_of(N, -1);
    for (int r = 0; r < N; ++r) {
      const int root = find(r);
      if (comp_of[root] < 0) { comp_of[root] = (int)comp.size(); comp.emplace_back(); }
      comp[comp_of[root]].push_back(r);
    }
  }
  // small components may share a role (two of <= 3 rows: one lane each, no coupling)
  std::sort(comp.begin(), comp.end(), [](const std::vector<int>& a, const std::vector<int>& b) { return a.size() > b.size(); });
  for (size_t i = 0; i < comp.size(); ++i) {
    if (comp[i].size() > 6) return QOCVL_ERR_UNSUPPORTED;
    if (comp[i].size() > 3) continue;
    for (size_t j2 = i + 1; j2 < comp.size(); ++j2)
      if (comp[j2].size() <= 3) {
        // keep them apart with padding rows so that the search below puts one component per lane
        std::vector<int> merged = comp[i];
        merged.resize(3, -1);
        merged.insert(merged.end(), comp[j2].begin(), comp[j2].end());
        comp[i] = merged;
        comp.erase(comp.begin() + j2);
        break;
      }
  }
  const int nroles = (int)comp.size();
  if (nroles > DUO_MAXR) return QOCVL_ERR_UNSUPPORTED;
  // ---- cut and order of every role: cheapest compiled pattern that covers some arrangement of its rows
  std::vector<int> srow((size_t)nroles * 6, -1);
  int variant[DUO_MAXR] = {0};
  for (int r = 0; r < nroles; ++r) {
    std::vector<int> rows = comp[r];
    rows.resize(6, -1);
    std::sort(rows.begin(), rows.end());
    auto nz = [&](int a, int b) { return a >= 0 && b >= 0 && pat[(size_t)a * N + b]; };
    int best_v = -1, best_cost = 1 << 30;
    std::vector<int> best_perm;
    std::vector<int> perm = rows;
    do {
      for (int v = 0; v < kSearchVariants; ++v) {
        const PatInfo& pi = kPat[v];
        const int cost = pi.nt + pi.nc;
        if (cost >= best_cost) continue;
        const int snd[2] = {pi.s0, pi.s1};
        bool ok = true;
        for (int t = 0; t < 2 && ok; ++t) {
          const int* mine = &perm[t * 3];
          const int* other = &perm[(1 - t) * 3];
          for (int i = 0; i < 3 && ok; ++i) {
            for (int ip = 0; ip < 3 && ok; ++ip)
              if (nz(mine[i], mine[ip]) && !((pi.tm >> (3 * i + ip)) & 1u)) ok = false;
            for (int ip = 0; ip < 3 && ok; ++ip) {
              if (!nz(mine[i], other[ip])) continue;
              int g = -1, h = -1;
              for (int x = 0; x < 2; ++x) { if (snd[x] == i) g = x; if (snd[x] == ip) h = x; }
              if (g < 0 || h < 0 || !((pi.cm >> (2 * g + h)) & 1u)) ok = false;
            }
          }
        }
        if (ok) { best_cost = cost; best_v = v; best_perm = perm; }
      }
    } while (std::next_permutation(perm.begin(), perm.end()));
    if (best_v < 0) return QOCVL_ERR_UNSUPPORTED;
    variant[r] = best_v;
    for (int i = 0; i < 6; ++i) srow[(size_t)r * 6 + i] = best_perm[i];
  }
  // ---- generator classes (as in chain_aug.cu): one class per entry is all this kernel takes
  const std::vector<int>& cls = p->h_aug_cls;
  const std::vector<char>& has_add = p->h_aug_hasadd;
  const int ncls = p->la.ncls;
  if ((int)cls.size() != J || (int)p->h_aug_mcls.size() != S * ncls) return QOCVL_ERR_STATE;
  if (ncls > 2) return QOCVL_ERR_UNSUPPORTED;           // the lane programs select between 1 and ONE per-sample multiplier
  // ---- descriptors: (role, lane type, slot) -> stacked (row, col), generators, class, additive generators
  struct Desc { int role, type, u, a, b, x; };
  std::vector<Desc> desc;
  int KG = 1;
  int nu[DUO_MAXR] = {0}, nx[DUO_MAXR] = {0}, ent0[DUO_MAXR] = {0};
  for (int r = 0; r < nroles; ++r) {
    const PatInfo& pi = kPat[variant[r]];
    nu[r] = pi.nt + 2 * pi.nc;
    nx[r] = pi.nt + pi.nc;
    ent0[r] = (int)desc.size();
    const int snd[2] = {pi.s0, pi.s1};
    for (int t = 0; t < 2; ++t) {
      const int* mine = &srow[(size_t)r * 6 + t * 3];
      const int* other = &srow[(size_t)r * 6 + (1 - t) * 3];
      std::vector<Desc> lane(nu[r]);
      for (int i = 0; i < 3; ++i)
        for (int ip = 0; ip < 3; ++ip)
          if ((pi.tm >> (3 * i + ip)) & 1u) {
            const int u = slot_of(pi.tm, 3 * i + ip);
            lane[u] = {r, t, u, mine[i], mine[ip], u};
          }
      for (int g = 0; g < 2; ++g)
        for (int h = 0; h < 2; ++h)
          if ((pi.cm >> (2 * g + h)) & 1u) {
            const int c = slot_of(pi.cm, 2 * g + h);
            lane[pi.nt + c] = {r, t, pi.nt + c, mine[snd[g]], other[snd[h]], -1};                     // row owner
            lane[pi.nt + pi.nc + c] = {r, t, pi.nt + pi.nc + c, other[snd[g]], mine[snd[h]], pi.nt + c};   // column owner
          }
      desc.insert(desc.end(), lane.begin(), lane.end());
    }
  }
  const int ntot = (int)desc.size();
  if (ntot > 64) return QOCVL_ERR_UNSUPPORTED;          // k_duo_reduce: 64 descriptor slots
  for (const Desc& d : desc) {
    if (d.a < 0 || d.b < 0) continue;
    int cnt = 0, c0 = -1;
    for (int j = 0; j < J; ++j)
      if (std::abs(AG(j, d.a, d.b)) > 0.0) {
        ++cnt;
        if (c0 < 0) c0 = cls[j];
        else if (cls[j] != c0) return QOCVL_ERR_UNSUPPORTED;   // two multiplier classes in one entry
      }
    KG = std::max(KG, cnt);
  }
  std::vector<int> ent_gen((size_t)ntot * KG, -1), ent_x(ntot, -1), ent_stride(ntot), ent_ystride(ntot), ent_ywarps(ntot);
  std::vector<cplx> ent_val((size_t)ntot * KG, cmake(0, 0));
  std::vector<unsigned long long> ent_out(ntot), ent_ysrc(ntot);
  const int warps_per_role = (S + DUO_SPW - 1) / DUO_SPW;
  size_t tab_off[DUO_MAXR], con_off[DUO_MAXR], y_off[DUO_MAXR], warp0[DUO_MAXR];
  size_t tab_total = 0, con_total = 0, y_total = 0;
  for (int r = 0; r < nroles; ++r) {
    tab_off[r] = tab_total; tab_total += (size_t)M * 2 * nu[r];
    con_off[r] = con_total; con_total += (size_t)2 * nu[r] * S;
    y_off[r] = y_total; y_total += (size_t)warps_per_role * M * 2 * nx[r];
    warp0[r] = (size_t)r * warps_per_role;
  }
  unsigned clsmask[2 * DUO_MAXR] = {0}, amask[2 * DUO_MAXR] = {0};
  std::vector<cplx> con_a(con_total, cmake(0, 0));
  for (int i = 0; i < ntot; ++i) {
    const Desc& d = desc[i];
    ent_x[i] = d.x;
    ent_out[i] = tab_off[d.role] + (size_t)d.type * nu[d.role] + d.u;
    ent_stride[i] = 2 * nu[d.role];
    ent_ysrc[i] = y_off[d.role] + (size_t)d.type * nx[d.role] + (d.x >= 0 ? d.x : 0);
    ent_ystride[i] = 2 * nx[d.role];
    ent_ywarps[i] = warps_per_role;
    if (d.a < 0 || d.b < 0) continue;
    int cnt = 0, c0 = -1;
    for (int j = 0; j < J; ++j) {
      const zc v = AG(j, d.a, d.b);
      if (std::abs(v) == 0.0) continue;
      ent_gen[(size_t)i * KG + cnt] = j;
      ent_val[(size_t)i * KG + cnt] = cmake(v.real(), v.imag());
      c0 = cls[j];
      ++cnt;
    }
    const size_t cbase = con_off[d.role] + ((size_t)d.type * nu[d.role] + d.u) * S;
    for (int s = 0; s < S; ++s) {
      zc acc(0, 0);
      for (int j = 0; j < J; ++j)
        if (has_add[j] && std::abs(AG(j, d.a, d.b)) > 0.0) {
          const cplx ad = p->h_add[(size_t)s * J + j];
          acc += zc(ad.x, ad.y) * AG(j, d.a, d.b);
        }
      con_a[cbase + s] = cmake(acc.real() * p->d.dt, acc.imag() * p->d.dt);
      if (acc != zc(0, 0)) amask[d.role * 2 + d.type] |= 1u << d.u;
    }
    if (c0 > 0) clsmask[d.role * 2 + d.type] |= 1u << d.u;
  }
  // ---- lean twin (pulse-map path): a role drops the Xbar of its own-block DIAGONAL when none of those entries has a
  //      generator whose coefficient depends on a live pulse channel (qocvl_model_gen_channels / _dead_channels)
  bool has_lean = false;
  int variant_lean[DUO_MAXR], nx_lean[DUO_MAXR];
  std::vector<int> ent_x_lean = ent_x, ent_ystride_lean = ent_ystride;
  std::vector<unsigned long long> ent_ysrc_lean = ent_ysrc;
  {
    const unsigned dead = p->opt("duo_nolean") ? 0u : qocvl_model_dead_channels(p->d.model);
    for (int r = 0; r < nroles; ++r) {
      variant_lean[r] = variant[r]; nx_lean[r] = nx[r];
      const int lv = kPat[variant[r]].lean;
      if (lv < 0 || dead == 0u) continue;
      const PatInfo &pf = kPat[variant[r]], &pl = kPat[lv];
      bool ok = true;
      for (int i = ent0[r]; i < ent0[r] + 2 * nu[r] && ok; ++i) {
        const Desc& d = desc[i];
        if (d.u >= pf.nt || d.a < 0 || d.b < 0) continue;                  // own-block entries only
        int bit = 0;                                                       // bit of entry slot d.u inside TM
        for (int b2 = 0, cnt = 0; b2 < 9; ++b2)
          if ((pf.tm >> b2) & 1u) { if (cnt == d.u) bit = b2; ++cnt; }
        if ((pl.xm >> bit) & 1u) continue;                                 // kept in the lean pattern
        for (int j = 0; j < J; ++j)
          if (std::abs(AG(j, d.a, d.b)) > 0.0 && (qocvl_model_gen_channels(p->d.model, j) & ~dead)) ok = false;
      }
      if (!ok) continue;
      has_lean = true;
      variant_lean[r] = lv;
      nx_lean[r] = duo_popc(pl.xm) + pl.nc;
      for (int i = ent0[r]; i < ent0[r] + 2 * nu[r]; ++i) {
        const Desc& d = desc[i];
        int x = -1;
        if (d.u < pf.nt) {
          int bit = 0;
          for (int b2 = 0, cnt = 0; b2 < 9; ++b2)
            if ((pf.tm >> b2) & 1u) { if (cnt == d.u) bit = b2; ++cnt; }
          if ((pl.xm >> bit) & 1u) x = slot_of(pl.xm, bit);
        } else if (d.x >= 0) {
          x = d.x - pf.nt + duo_popc(pl.xm);                               // column-owner coupling slots follow
        }
        ent_x_lean[i] = x;
        ent_ystride_lean[i] = 2 * nx_lean[r];
        ent_ysrc_lean[i] = y_off[r] + (size_t)d.type * nx_lean[r] + (x >= 0 ? x : 0);
      }
    }
  }
  // ---- geometry: warps per CTA so that the reverse kernel's shared memory fits
  int numax = 0;
  for (int r = 0; r < nroles; ++r) numax = std::max(numax, nu[r]);
  int wpc = 4;
  while (wpc > 1 && (size_t)wpc * duo_smem_layout(numax, p->qcap, true, esz).total > 220 * 1024) wpc >>= 1;
  if ((size_t)wpc * duo_smem_layout(numax, p->qcap, true, esz).total > 220 * 1024) return QOCVL_ERR_UNSUPPORTED;
  if (p->has_opt("duo_wpc")) wpc = std::max(1, std::min(wpc, p->opt("duo_wpc")));
  D.nroles = nroles; D.wpc = wpc; D.KG = KG; D.ntot = ntot; D.warps_per_role = warps_per_role;
  D.smem_fwd = (size_t)wpc * duo_smem_layout(numax, p->qcap, false, esz).total;
  D.smem_rev = (size_t)wpc * duo_smem_layout(numax, p->qcap, true, esz).total;
  D.numax = numax;
  for (int r = 0; r < nroles; ++r) {
    D.variant[r] = variant[r]; D.nu[r] = nu[r]; D.nx[r] = nx[r]; D.ent0[r] = ent0[r];
    D.tab_off[r] = tab_off[r]; D.con_off[r] = con_off[r]; D.y_off[r] = y_off[r]; D.warp0[r] = warp0[r];
  }
  p->blocks = nroles * ((warps_per_role + wpc - 1) / wpc);
  p->threads = wpc * 32;
  p->spb = wpc * DUO_SPW; p->groups = p->spb;
  // ---- time-parallel evaluation of small ensembles: the sweeps are bound by the sequential depth of one sample's
  //      slices (the evaluation takes the same 8.7 ms for 512 and for 4096 samples), so a small ensemble -- a shard of
  //      a strong-scaling run -- cuts the slice axis into chunks that run concurrently.  The price is the chunk
  //      propagators: one forward sweep per basis column of every role (CZ: 3 + 6 instead of 1 + 1).
  //      A role of the Van Loan form [[T, 0], [C, T]] (the {U psi, W psi} pair: one triple evolves on its own, both
  //      triples carry the same own block) has the propagator [[U, 0], [W, U]]: only the columns of the independent
  //      triple are swept, the other three are mirrored by k_duo_combine (mirror[r] = 1 + the independent lane type).
  std::vector<int> colrow((size_t)nroles * 6, 0);
  int ncol[DUO_MAXR] = {0}, mirror[DUO_MAXR] = {0};
  for (int r = 0; r < nroles; ++r) {
    const int* rows = &srow[(size_t)r * 6];
    for (int t = 0; t < 2 && !mirror[r]; ++t) {          // is triple t independent of triple 1 - t, with equal own blocks?
      bool ok = true;
      for (int i = 0; i < 3 && ok; ++i) {
        if (rows[t * 3 + i] < 0 || rows[(1 - t) * 3 + i] < 0) { ok = false; break; }
        for (int ip = 0; ip < 3 && ok; ++ip) {
          if (pat[(size_t)rows[t * 3 + i] * N + rows[(1 - t) * 3 + ip]]) ok = false;
          for (int j = 0; j < J && ok; ++j)
            if (AG(j, rows[t * 3 + i], rows[t * 3 + ip]) != AG(j, rows[(1 - t) * 3 + i], rows[(1 - t) * 3 + ip])) ok = false;
        }
      }
      if (ok) mirror[r] = 1 + t;
    }
    if (p->opt("duo_nomirror")) mirror[r] = 0;
    for (int l = 0; l < 6; ++l)
      if (rows[l] >= 0 && (!mirror[r] || l / 3 == mirror[r] - 1)) colrow[(size_t)r * 6 + ncol[r]++] = l;
  }
  int chunks = 1;
  if (S <= 2304 && M >= 512) {                          // measured on cfg 3 (profiles/README.md): 16 chunks win up to ~2300 samples
    while (chunks < 16 && M / (chunks * 2) >= 64) chunks *= 2;
  }
  if (S == 1 && M >= 32) chunks = std::max(1, std::min(32, M / 16));   // one sample (notebooks 01 / 02): ~16 slices per chunk
  if (p->has_opt("strong_chunks")) chunks = std::max(1, std::min(std::min(64, M / 4), p->opt("strong_chunks")));
  D.lc = (M + chunks - 1) / chunks;
  D.chunks = (M + D.lc - 1) / D.lc;
  D.G = warps_per_role;
  for (int r = 0; r < nroles; ++r) { D.ncol[r] = ncol[r]; D.mirror[r] = mirror[r]; }
  // ---- upload
#define DUO_UP(ptr, vec)                                                                               \
  do {                                                                                                 \
    QCUDA(cudaMalloc((void**)&(ptr), std::max<size_t>(1, (vec).size()) * sizeof((vec)[0])));            \
    if (!(vec).empty())                                                                                \
      QCUDA(cudaMemcpy((ptr), (vec).data(), (vec).size() * sizeof((vec)[0]), cudaMemcpyHostToDevice));  \
  } while (0)
  for (int i = 0; i < 2 * DUO_MAXR; ++i) { D.clsmask[i] = clsmask[i]; D.amask[i] = amask[i]; }
  DUO_UP(D.d_con_a, con_a); DUO_UP(D.d_srow, srow); DUO_UP(D.d_colrow, colrow);
  DUO_UP(D.d_ent_gen, ent_gen); DUO_UP(D.d_ent_val, ent_val); DUO_UP(D.d_ent_x, ent_x);
  DUO_UP(D.d_ent_out, ent_out); DUO_UP(D.d_ent_stride, ent_stride);
  DUO_UP(D.d_ent_ysrc, ent_ysrc); DUO_UP(D.d_ent_ystride, ent_ystride); DUO_UP(D.d_ent_ywarps, ent_ywarps);
  D.has_lean = has_lean;
  D.nnz_x_lean = 0;
  for (int i = 0; i < ntot; ++i)
    if (ent_x_lean[i] >= 0 && desc[i].a >= 0 && desc[i].b >= 0 && pat[(size_t)desc[i].a * N + desc[i].b]) {
      D.nnz_x_lean++;
      D.nnz_x_lean_role[desc[i].role]++;
    }
  // ---- per-role Taylor degrees: the roles are the diagonal blocks of the stacked system, so the remainder of a role's
  //      series is bounded by the norm of ITS block -- the column sums of the basis states its rows stand for
  //      (k_coefficients, model.cuh::qocvl_slice_norm_csr).  CZ: the {U psi, W psi} pair has half the norm of U|11>.
  {
    std::vector<unsigned char> colrole(2 * (size_t)p->d.n, 0);   // [q part | p part][basis state]
    bool ok = (int)p->h_aug_members.size() == N && p->opt("duo_roleq", 3) != 0;
    for (int r = 0; r < nroles && ok; ++r)
      for (int row : comp[r]) {
        if (row < 0) continue;
        if (p->h_aug_members[row].empty()) ok = false;
        for (int b : p->h_aug_members[row]) colrole[(row >= p->aug_p0 ? p->d.n : 0) + b] |= (unsigned char)(1u << r);
        for (int c2 : comp[r])
          if (c2 >= 0 && pat[(size_t)row * N + c2]) D.nnz_role[r]++;
      }
    D.role_q = ok;
    if (ok) {
      DUO_UP(D.d_colrole, colrole);
      QCUDA(cudaMalloc((void**)&D.d_qdeg_role, (size_t)DUO_MAXR * M * sizeof(int)));
      QCUDA(cudaMemset(D.d_qdeg_role, 0, (size_t)DUO_MAXR * M * sizeof(int)));
    }
    // ---- phase shift per role: needs one role per basis state (the shift of a state's diagonal is its role's) and
    //      diagonal entries without a per-sample multiplier (entry = A + m P would scale the shift with m)
    //      (complex128 only: in fp32 the rotating frame costs accuracy -- without the shift the dominant amplitudes sit
    //      on states with an exactly zero diagonal and are never multiplied; measured 10x the adiabaticity error)
    bool shift = ok && p->opt("duo_roleq", 3) >= 3 && !p->c64;
    for (int b = 0; b < p->d.n && shift; ++b) shift = duo_popc((unsigned)(colrole[b] | colrole[p->d.n + b])) <= 1;
    std::vector<int> rowrole(N, -1), ent_srole(ntot, -1);
    for (int r = 0; r < nroles; ++r)
      for (int row : comp[r]) if (row >= 0) rowrole[row] = r;
    for (int i = 0; i < ntot && shift; ++i) {
      const Desc& d = desc[i];
      if (d.a < 0 || d.a != d.b || d.u >= kPat[variant[d.role]].nt) continue;       // own-block diagonal entries only
      ent_srole[i] = d.role;
      if ((clsmask[d.role * 2 + d.type] >> d.u) & 1u) shift = false;
    }
    // every row of a role needs its diagonal slot in the pattern (all compiled patterns have the full diagonal)
    D.shift_on = shift;
    if (shift) {
      DUO_UP(D.d_rowrole, rowrole); DUO_UP(D.d_ent_srole, ent_srole);
      QCUDA(cudaMalloc((void**)&D.d_phi, (size_t)DUO_MAXR * M * sizeof(double)));
      QCUDA(cudaMemset(D.d_phi, 0, (size_t)DUO_MAXR * M * sizeof(double)));
      QCUDA(cudaMalloc((void**)&D.d_targets_rot, (size_t)N * sizeof(cplx)));
    }
  }
  for (int r = 0; r < nroles; ++r) { D.variant_lean[r] = variant_lean[r]; D.nx_lean[r] = nx_lean[r]; }
  DUO_UP(D.d_ent_x_lean, ent_x_lean); DUO_UP(D.d_ent_ysrc_lean, ent_ysrc_lean); DUO_UP(D.d_ent_ystride_lean, ent_ystride_lean);
  {
    std::vector<cplx> st(p->h_aug_st.begin(), p->h_aug_st.begin() + N), tg(p->h_aug_tg.begin(), p->h_aug_tg.begin() + N);
    std::vector<int> pr(p->h_aug_pair.begin(), p->h_aug_pair.begin() + N);
    const int Lk = (int)p->h_aug_wq.size() / 2;
    std::vector<double> wq(2 * (size_t)N);
    for (int r = 0; r < N; ++r) { wq[r] = p->h_aug_wq[r]; wq[N + r] = p->h_aug_wq[Lk + r]; }
    DUO_UP(D.d_starts, st); DUO_UP(D.d_targets, tg); DUO_UP(D.d_pair, pr); DUO_UP(D.d_wq, wq);
  }
#undef DUO_UP
  QCUDA(cudaMalloc((void**)&D.d_tab, tab_total * sizeof(cplx)));
  QCUDA(cudaMalloc((void**)&D.d_zfin, (size_t)S * N * sizeof(cplx)));
  QCUDA(cudaMalloc((void**)&D.d_lam, (size_t)S * N * sizeof(cplx)));
  if (D.chunks > 1) {
    QCUDA(cudaMalloc((void**)&D.d_prop, (size_t)S * D.chunks * nroles * 36 * sizeof(cplx)));
    QCUDA(cudaMemset(D.d_prop, 0, (size_t)S * D.chunks * nroles * 36 * sizeof(cplx)));
    QCUDA(cudaMalloc((void**)&D.d_zstart, (size_t)S * D.chunks * N * sizeof(cplx)));
    QCUDA(cudaMalloc((void**)&D.d_lamend, (size_t)S * D.chunks * N * sizeof(cplx)));
  }
  if (p->d_ckpt) { cudaFree(p->d_ckpt); p->d_ckpt = nullptr; p->bytes -= p->ckpt_bytes; p->ckpt_bytes = 0; }
  if (p->d_ypart) { cudaFree(p->d_ypart); p->d_ypart = nullptr; p->bytes -= p->ypart_bytes; p->ypart_bytes = 0; }
  p->ckpt_bytes = (size_t)nroles * warps_per_role * M * 96 * esz;
  p->ypart_bytes = y_total * sizeof(cplx);
  QCUDA(cudaMalloc((void**)&p->d_ckpt, p->ckpt_bytes));
  QCUDA(cudaMalloc((void**)&p->d_ypart, std::max<size_t>(16, p->ypart_bytes)));
  p->bytes += p->ckpt_bytes + p->ypart_bytes;
  if (p->c64) {
    QCUDA(cudaFuncSetAttribute(k_duo_forward<cplxf>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D.smem_fwd));
    QCUDA(cudaFuncSetAttribute(k_duo_forward_cols<cplxf>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D.smem_fwd));
    QCUDA(cudaFuncSetAttribute(k_duo_reverse<cplxf>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D.smem_rev));
  } else {
    QCUDA(cudaFuncSetAttribute(k_duo_forward<cplx>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D.smem_fwd));
    QCUDA(cudaFuncSetAttribute(k_duo_forward_cols<cplx>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D.smem_fwd));
    QCUDA(cudaFuncSetAttribute(k_duo_reverse<cplx>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D.smem_rev));
  }
  p->aug_store = false;
  p->aug_parts = nroles * warps_per_role;
  p->n_parts = 0;
  p->qcap_apriori = p->qcap;
  p->use_duo = true;
  return QOCVL_OK;
}

int qocvl_duo_launch(qocvl_plan* p, bool want_grad, double* gwave, cudaStream_t st) {
  const DuoState& D = p->duo;
  const qocvl_desc& d = p->d;
  if (want_grad && !gwave) return QOCVL_ERR_ARG;
  DuoArgs q;
  memset(&q, 0, sizeof(q));
  q.M = d.n_slices; q.S = p->n_samples; q.C = d.n_chan; q.J = d.n_gen; q.N = p->aug_n;
  q.nroles = D.nroles; q.wpc = D.wpc; q.qcap = p->qcap_apriori; q.want_grad = want_grad ? 1 : 0; q.KG = D.KG;
  q.ntot = D.ntot;
  for (int r = 0; r < D.nroles; ++r) {
    q.variant[r] = D.variant[r]; q.nu[r] = D.nu[r]; q.nx[r] = D.nx[r]; q.warps[r] = D.warps_per_role;
    q.tab_off[r] = D.tab_off[r]; q.con_off[r] = D.con_off[r]; q.y_off[r] = D.y_off[r]; q.warp0[r] = D.warp0[r];
    q.ent0[r] = D.ent0[r];
  }
  q.dt = d.dt; q.inv_duration = 1.0 / d.duration; q.w_infid = d.w_infid; q.w_adiab = d.w_adiab;
  q.inv_fid_norm = 1.0 / d.fid_norm; q.inv_samples = 1.0 / (double)p->global_samples;
  q.d_const = p->d_const;
  q.tab = D.d_tab; q.con_a = D.d_con_a; q.srow = D.d_srow;
  q.mcls = p->la.mcls; q.ncls = p->la.ncls;
  for (int i = 0; i < 2 * DUO_MAXR; ++i) { q.clsmask[i] = D.clsmask[i]; q.amask[i] = D.amask[i]; }
  q.ent_gen = D.d_ent_gen; q.ent_val = D.d_ent_val; q.ent_x = D.d_ent_x;
  q.coef = p->d_coef; q.dcoef = p->d_dcoef;
  q.qdeg = D.role_q ? D.d_qdeg_role : p->d_qdeg; q.qstride = D.role_q ? d.n_slices : 0;   // [role][M] or one row for all
  q.starts = D.d_starts; q.targets = D.shift_on ? D.d_targets_rot : D.d_targets; q.pair = D.d_pair; q.wq = D.d_wq;
  q.phi = D.shift_on ? D.d_phi : nullptr; q.ent_srole = D.d_ent_srole;
  q.ckpt = p->d_ckpt;                                   // (elements of the sweeps' type: the kernels cast)
  q.zfin = D.d_zfin; q.lam = D.d_lam; q.ypart = p->d_ypart;
  q.sample_out = p->d_sample_out; q.gwave = gwave;
  q.ifact[0] = 1.0;
  for (int j = 1; j < QOCVL_QMAX + 2; ++j) q.ifact[j] = q.ifact[j - 1] / (double)j;
  for (int j = 0; j < QOCVL_QMAX + 2; ++j) q.ifactf[j] = (float)q.ifact[j];
  q.chunks = D.chunks; q.lc = D.lc; q.G = D.G; q.cpi = 1;
  for (int r = 0; r < D.nroles; ++r) { q.ncol[r] = D.ncol[r]; q.mirror[r] = D.mirror[r]; }
  q.colrow = D.d_colrow; q.prop = D.d_prop; q.zstart = D.d_zstart; q.lamend = D.d_lamend;
  // pulse-map path: the reverse sweep and the reduction skip what only a dead channel would consume
  const bool lean = want_grad && p->lean_now && D.has_lean;
  if (want_grad) p->duo.last_lean = lean;
  const unsigned long long* ysrc = lean ? D.d_ent_ysrc_lean : D.d_ent_ysrc;
  const int* ystride = lean ? D.d_ent_ystride_lean : D.d_ent_ystride;
  if (lean) {
    q.ent_x = D.d_ent_x_lean;
    for (int r = 0; r < D.nroles; ++r) { q.variant[r] = D.variant_lean[r]; q.nx[r] = D.nx_lean[r]; }
  }
  // one sweep launch: `items` work items (warps) per role
  auto sweep = [&](bool reverse, int mode) {
    DuoArgs a = q;
    a.mode = mode;
    // (several basis columns per item share the entry registers, but measured SLOWER: 0.95 vs 0.87 ms at 512 samples --
    //  a third of the warps, 214 registers; kept behind "duo_cols" = 1 as a cross-check)
    const bool cols = mode == 1 && p->opt("duo_cols", 0) != 0;
    a.cpi = cols ? DUO_COLS : 1;
    int most = 0;
    for (int r = 0; r < D.nroles; ++r) {
      a.warps[r] = D.G * (mode == 0 ? 1 : D.chunks) * (mode == 1 ? (D.ncol[r] + a.cpi - 1) / a.cpi : 1);
      most = std::max(most, a.warps[r]);
    }
    const int blocks = D.nroles * ((most + D.wpc - 1) / D.wpc);
    if (p->c64) {
      if (reverse) k_duo_reverse<cplxf><<<blocks, p->threads, D.smem_rev, st>>>(a);
      else if (cols) k_duo_forward_cols<cplxf><<<blocks, p->threads, D.smem_fwd, st>>>(a);
      else k_duo_forward<cplxf><<<blocks, p->threads, D.smem_fwd, st>>>(a);
    } else {
      if (reverse) k_duo_reverse<cplx><<<blocks, p->threads, D.smem_rev, st>>>(a);
      else if (cols) k_duo_forward_cols<cplx><<<blocks, p->threads, D.smem_fwd, st>>>(a);
      else k_duo_forward<cplx><<<blocks, p->threads, D.smem_fwd, st>>>(a);
    }
    p->launches++;
  };
  const bool timed = p->timing_now && p->ev_stage[0];
  // "debug_sync" option: synchronise after every launch and name the kernel that failed (diagnostic only)
  const bool dbg = p->opt("debug_sync") != 0;
#define DUO_DBG(name)                                                                     \
  do {                                                                                    \
    if (dbg) {                                                                            \
      cudaError_t e_ = cudaStreamSynchronize(st);                                         \
      if (e_ == cudaSuccess) e_ = cudaGetLastError();                                     \
      if (e_ != cudaSuccess) {                                                            \
        g_qocvl_cuda_error = std::string(name) + ": " + cudaGetErrorString(e_);           \
        return QOCVL_ERR_CUDA;                                                            \
      }                                                                                   \
    }                                                                                     \
  } while (0)
  if (p->c64) k_duo_tables<cplxf><<<d.n_slices, 64, 0, st>>>(q, D.d_ent_out, D.d_ent_stride);
  else k_duo_tables<cplx><<<d.n_slices, 64, 0, st>>>(q, D.d_ent_out, D.d_ent_stride);
  p->launches++;
  DUO_DBG("k_duo_tables");
  if (D.shift_on) {
    k_duo_phase<<<1, 256, 0, st>>>(q.M, q.N, q.nroles, D.d_phi, D.d_rowrole, D.d_targets, D.d_targets_rot);
    p->launches++;
    DUO_DBG("k_duo_phase");
  }
  if (D.chunks > 1) {            // chunk propagators -> boundary states -> (terminal cost) -> boundary cotangents
    sweep(false, 1);
    DUO_DBG("k_duo_forward (chunk propagators)");
    k_duo_combine<<<(q.S * q.nroles * 8 + 127) / 128, 128, 0, st>>>(q, +1);
    p->launches++;
    DUO_DBG("k_duo_combine (+1)");
  } else {
    if (timed) QCUDA(cudaEventRecord(p->ev_stage[0], st));
    sweep(false, 0);
    if (timed) { QCUDA(cudaEventRecord(p->ev_stage[1], st)); p->stage_timed[0] = true; }
    DUO_DBG("k_duo_forward");
  }
  k_duo_terminal<<<(q.S + 127) / 128, 128, 0, st>>>(q);
  p->launches++;
  DUO_DBG("k_duo_terminal");
  if (want_grad) {
    if (D.chunks > 1) {
      k_duo_combine<<<(q.S * q.nroles * 8 + 127) / 128, 128, 0, st>>>(q, -1);
      p->launches++;
      if (timed) QCUDA(cudaEventRecord(p->ev_stage[0], st));
      sweep(false, 2);
      if (timed) { QCUDA(cudaEventRecord(p->ev_stage[1], st)); p->stage_timed[0] = true; }
    }
    if (timed) QCUDA(cudaEventRecord(p->ev_stage[2], st));
    DUO_DBG("k_duo_combine (-1) / k_duo_forward (chunks)");
    sweep(true, D.chunks > 1 ? 2 : 0);
    if (timed) { QCUDA(cudaEventRecord(p->ev_stage[3], st)); p->stage_timed[1] = true; }
    DUO_DBG("k_duo_reverse");
    k_duo_reduce<<<d.n_slices, 256, 0, st>>>(q, ysrc, ystride, D.d_ent_ywarps);
    p->launches++;
  }
  QCUDA(cudaGetLastError());
  return QOCVL_OK;
}
