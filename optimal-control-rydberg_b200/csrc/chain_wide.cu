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
#include "chain_wide.cuh"
#include <cooperative_groups.h>
#include <algorithm>
#include <numeric>

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
                              size_t ent_fwd, size_t ent_rev, bool dreal, size_t esz) {
  const int rw = 32 / spc, nslot = nb_c + nd_c;
  const size_t vs = (size_t)WIDE_NVP * n * spc;
  const size_t sweep_bytes = (2 * vs + 2 * (size_t)J * spc) * esz;     // vectors and per-slice coefficients: the sweeps' type
  size_t cplx_count = (size_t)J * spc + (size_t)J * C +
                      (size_t)(threads / 32) * nslot * spc + (size_t)nslot * spc +
                      (dreal ? 0 : (size_t)nd_c * ngroups * rw);
  size_t dbl_count = (size_t)J * spc + (size_t)(threads / 32) * 3 * spc + 2 * (size_t)spc +
                     (dreal ? (size_t)nd_c * ngroups * rw : 0);
  return sweep_bytes + cplx_count * sizeof(cplx) + dbl_count * sizeof(double) + 2 * (size_t)nb_c * ngroups * sizeof(int2) +
         (((ent_fwd + 7) & ~(size_t)7) + ((ent_rev + 7) & ~(size_t)7)) * sizeof(unsigned short) +
         (((size_t)ngroups * rw + 15) & ~(size_t)15);
}

// The register file bounds rows per thread (state, diagonal and cotangent of every owned row live in registers):
// 512 threads x 128 registers up to 4 rows, 384 x 168 up to 8, 256 x 255 up to 14.
static int wide_threads_for(const qocvl_plan* p, int spc, int ngroups) {
  if (p->has_opt("wide_threads")) {
    const int t = p->opt("wide_threads");
    if (t == 256 || t == 384 || t == 512) return t;
  }
  if (spc != 8) return 512;
  if (ngroups <= 4 * 16) return 512;
  if (ngroups <= 8 * 12) return 384;
  return 256;
}
static wide_kernel_t wide_pick(int c64, int spc, int items, int threads, int nb, int nd, bool simple, bool fwd = false) {
  if (fwd) {
    if (spc != 8 || threads != 384) return nullptr;
    return simple ? qocvl_wide_pick_b13_fwd(items, c64) : qocvl_wide_pick_b24_fwd(items, c64);
  }
  if (simple) return qocvl_wide_pick_b13(spc, items, threads, c64);
  if (nb <= WIDE_MAXB && nd <= WIDE_MAXD) return qocvl_wide_pick_b24(spc, items, threads, c64);
  return nullptr;
}

void WideLayout::release() {
  void* ptrs[] = {grow, fwd_co, rev_co, fwd_val, fwd_col, rev_val, rev_col, dval, dval_re, dmask};
  for (void* q : ptrs) if (q) cudaFree(q);
  *this = WideLayout();
}

// Tables and launch geometry of the wide kernel for (up to) spc_request samples per CTA.  QOCVL_ERR_UNSUPPORTED when
// the problem does not fit the kernel.
static int wide_build(qocvl_plan* p, int spc_request, WideLayout& L, bool fwd = false) {
  const qocvl_desc& d = p->d;
  const int n = d.n, J = d.n_gen;
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
  // the specialised instantiation: one uniform-valued off-diagonal block of the u family, <= 3 diagonal blocks
  const bool simple = (nb == 1 && nd <= 3 && buni[0] && off_blk[0].fam == 0);
  const int nb_c = simple ? 1 : WIDE_MAXB, nd_c = simple ? 3 : WIDE_MAXD;   // compiled block counts
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
    std::vector<unsigned char> dmask;
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
          const size_t lpad = (longest + WIDE_CHUNK - 1) / WIDE_CHUNK * WIDE_CHUNK;   // lists padded to whole chunks
          T.co[dir][(size_t)bi * T.ngroups + g] = make_int2((int)(lpad / WIDE_CHUNK), (int)T.col[dir].size());
          for (int r = 0; r < T.rw; ++r) {
            const int own = std::max(0, T.grow[(size_t)g * T.rw + r]);
            for (size_t e = 0; e < lpad; ++e) {
              if (e < lists[r].size()) {
                T.val[dir].push_back(cmake(lists[r][e].second.real(), lists[r][e].second.imag()));
                T.col[dir].push_back((unsigned short)(lists[r][e].first * WIDE_NVP * spc_));
              } else {   // padding: a uniform-valued block adds the zero row n (kept zero in both buffers)
                T.val[dir].push_back(cmake(0, 0));
                T.col[dir].push_back((unsigned short)((buni[bi] ? n : own) * WIDE_NVP * spc_));
              }
            }
          }
        }
    }
    T.dval.assign((size_t)std::max(1, nd) * T.ngroups * T.rw, cmake(0, 0));
    T.dval_re.assign(T.dval.size(), 0.0);
    T.dmask.assign((size_t)T.ngroups * T.rw, 0);
    for (int di = 0; di < nd; ++di)
      for (int i = 0; i < T.ngroups * T.rw; ++i) {
        if (T.grow[i] < 0) continue;
        const std::complex<double> v = G(diag_blk[di].fam, diag_blk[di].g, T.grow[i], T.grow[i]);
        T.dval[(size_t)di * T.ngroups * T.rw + i] = cmake(v.real(), v.imag());
        T.dval_re[(size_t)di * T.ngroups * T.rw + i] = v.real();
        if (std::abs(v) > 0.0) T.dmask[i] |= (unsigned char)(1u << di);
      }
    return T;
  };
  // ---- samples per CTA (8 when the CTA's vectors and tables fit shared memory), warps, rows per thread
  if ((n + 1) * WIDE_NVP * 8 > 65535) return QOCVL_ERR_UNSUPPORTED;   // pre-scaled 16-bit element offsets
  int spc = spc_request;
  Tables T;
  int max_threads = 0, nwarps = 0, items = 0;
  wide_kernel_t kern = nullptr;
  size_t smem = 0;
  for (;;) {
    T = build(spc);
    max_threads = fwd ? 384 : wide_threads_for(p, spc, T.ngroups);
    nwarps = std::min(max_threads / 32, T.ngroups);
    items = (T.ngroups + nwarps - 1) / nwarps;
    nwarps = (T.ngroups + items - 1) / items;              // fewest warps that still need `items` passes
    kern = wide_pick(p->c64, spc, items, max_threads, nb, nd, simple, fwd);
    // (n + 1 rows per vector: row n is the zero row the padding entries of uniform-valued blocks read)
    smem = wide_smem_bytes(n + 1, J, d.n_chan, spc, max_threads, nb_c, nd_c, T.ngroups, T.col[0].size(),
                           T.col[1].size(), dreal, p->c64 ? sizeof(cplxf) : sizeof(cplx));
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
  if ((rc = wide_upload(&L.dmask, T.dmask))) return rc;
  L.simple = simple; L.fwd = fwd;
  L.ent_fwd = (int)T.col[0].size(); L.ent_rev = (int)T.col[1].size(); L.dreal = dreal;
  for (int i = 0; i < nb; ++i) { L.buni[i] = buni[i]; L.bval[i] = cmake(bval[i].real(), bval[i].imag()); }
  L.macs = offdiag_macs + 3.0 * n;   // useful complex multiply-adds of one Taylor step (+ du*tq, du*tp, dw*tq)
  L.padded_entries = (long long)T.col[0].size();
  L.spc = spc; L.items = items; L.max_threads = max_threads; L.ngroups = ngroups; L.nb = nb; L.nd = nd;
  for (int i = 0; i < nb; ++i) { L.bgen[i] = off_blk[i].g; L.bfam[i] = off_blk[i].fam; }
  for (int i = 0; i < nd; ++i) { L.dgen[i] = diag_blk[i].g; L.dfam[i] = diag_blk[i].fam; }
  QCUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  L.smem = smem;
  L.threads = nwarps * 32;
  return QOCVL_OK;
}

// ---------------------------------------------------------------------------------------------------
// time-parallel single-sample path: stitch the chunk propagators together (one cluster of 8 CTAs).
//   forward : z_{c+1} = (U_c q_c , U_c p_c + W_c q_c), boundary states zc[c]
//   cost, terminal cotangents lc[n_chunks]
//   backward: lam_c = (U_c^dag lq + W_c^dag lp , U_c^dag lp), boundary cotangents lc[c]
// U_c, W_c are stored column by column ([column][row]: each column was written by one lane of the column sweep).
// ---------------------------------------------------------------------------------------------------
struct WideCombineArgs {
  int n, n_chunks, want_grad;
  double inv_duration, w_infid, w_adiab, inv_fid_norm, inv_samples;
  cplx d_const;
  const cplx *uc, *wc, *starts, *targets;
  cplx *zc, *lc;
  double* sample_out;
};

// A thread-block CLUSTER of WIDE_CC CTAs walks the chunks: every CTA keeps the whole current state (q, p) in its own shared
// memory, multiplies its slice of the rows (forward) / columns (backward) of the chunk's blocks -- so the 2 n^2 block
// entries of a chunk stream in through WIDE_CC SMs instead of one -- and writes its part of the next state into the shared
// memory of all WIDE_CC CTAs (distributed shared memory), one cluster barrier per chunk.  (One CTA: 1.35 ms of cfg 4's 9.3.)
#define WIDE_CC 8
__global__ void __cluster_dims__(WIDE_CC, 1, 1) __launch_bounds__(1024) k_wide_combine(const WideCombineArgs A) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = A.n, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  const int rank = (int)cluster.block_rank();
  const int per = (n + WIDE_CC - 1) / WIDE_CC, r0 = min(n, rank * per), r1 = min(n, r0 + per), nrows = r1 - r0;
  cplx* s_state = (cplx*)smem_raw;             // [2 buffers][q | p][n]: the whole state, double buffered by chunk
  cplx* s_part = s_state + 4 * n;              // [parts][2][per] partial sums of my rows
  const int parts = max(1, nt / max(1, per));  // thread (part, rr): row r0 + rr, columns b = part, part + parts, ...
  double* s_red = (double*)(s_part + (size_t)parts * 2 * per);   // [3][32]
  const cplx czero = cmake(0, 0);
  for (int r = tid; r < n; r += nt) { s_state[r] = A.starts[r]; s_state[n + r] = czero; }
  for (int r = r0 + tid; r < r1; r += nt) { A.zc[r] = A.starts[r]; A.zc[n + r] = czero; }
  cluster.sync();                              // (every CTA of the cluster is running before anyone writes into it)
  int cur = 0;
  const int part = tid / max(1, per), rr = tid % max(1, per);
  for (int c = 0; c < A.n_chunks; ++c) {
    const cplx* U = A.uc + (size_t)c * n * n;
    const cplx* W = A.wc + (size_t)c * n * n;
    const cplx *s_q = s_state + cur * 2 * n, *s_p = s_q + n;
    if (part < parts && rr < nrows) {
      cplx aq = czero, ap = czero;
      for (int b = part; b < n; b += parts) {
        const cplx u = U[(size_t)b * n + r0 + rr], w = W[(size_t)b * n + r0 + rr], qb = s_q[b], pb = s_p[b];
        aq = cfma(u, qb, aq);
        ap = cfma(w, qb, cfma(u, pb, ap));
      }
      s_part[((size_t)part * 2 + 0) * per + rr] = aq;
      s_part[((size_t)part * 2 + 1) * per + rr] = ap;
    }
    __syncthreads();
    for (int i = tid; i < 2 * nrows; i += nt) {          // i = v * nrows + rr: fixed summation order over the parts
      const int v = i / nrows, r2 = i % nrows;
      cplx a = czero;
      for (int pp = 0; pp < parts; ++pp) a = cadd(a, s_part[((size_t)pp * 2 + v) * per + r2]);
      const int dst = ((cur ^ 1) * 2 + v) * n + r0 + r2;
      for (int k = 0; k < WIDE_CC; ++k) cluster.map_shared_rank(s_state, k)[dst] = a;
      A.zc[(size_t)(c + 1) * 2 * n + v * n + r0 + r2] = a;
    }
    cluster.sync();
    cur ^= 1;
  }
  // ---- cost (every CTA holds the whole final state and evaluates it; rank 0 writes)
  const cplx *s_q = s_state + cur * 2 * n, *s_p = s_q + n;
  double red[3] = {0, 0, 0};
  for (int i = tid; i < n; i += nt) {
    const cplx y = A.targets[i], zq = s_q[i], zp = s_p[i];
    red[0] += y.x * zq.x + y.y * zq.y;
    red[1] += y.x * zq.y - y.y * zq.x;
    red[2] += zq.x * zp.x + zq.y * zp.y;
  }
  for (int k = 0; k < 3; ++k) {
    for (int off = 16; off > 0; off >>= 1) red[k] += __shfl_xor_sync(0xffffffffu, red[k], off);
    if (lane == 0) s_red[k * 32 + warp] = red[k];
  }
  __syncthreads();
  double tot[3] = {0, 0, 0};
  for (int k = 0; k < 3; ++k)
    for (int w = 0; w < nwarps; ++w) tot[k] += s_red[k * 32 + w];
  const cplx dsum = cmake(tot[0] + A.d_const.x, tot[1] + A.d_const.y);
  if (tid == 0 && rank == 0) {
    const double infid = 1.0 - (dsum.x * dsum.x + dsum.y * dsum.y) * A.inv_fid_norm;
    const double adiab = 1.0 - tot[2] * A.inv_duration;
    A.sample_out[0] = A.w_infid * infid + A.w_adiab * adiab;
    A.sample_out[1] = infid;
    A.sample_out[2] = adiab;
  }
  if (!A.want_grad) return;                    // (uniform over the cluster)
  // ---- terminal cotangents into the other state buffer (local: every CTA computes all of them)
  const double ca = -0.5 * A.w_adiab * A.inv_duration * A.inv_samples;
  const double cf = -A.w_infid * A.inv_fid_norm * A.inv_samples;
  {
    cplx* s_l = s_state + (cur ^ 1) * 2 * n;
    for (int i = tid; i < n; i += nt) {
      const cplx lq = cadd(cscale(cmul(dsum, A.targets[i]), cf), cscale(s_p[i], ca));
      const cplx lp = cscale(s_q[i], ca);
      s_l[i] = lq; s_l[n + i] = lp;
      if (i >= r0 && i < r1) {
        A.lc[(size_t)A.n_chunks * 2 * n + i] = lq;
        A.lc[(size_t)A.n_chunks * 2 * n + n + i] = lp;
      }
    }
  }
  cur ^= 1;
  cluster.sync();                              // nobody overwrites the final state while a neighbour still reads it
  // ---- backward walk: my slice of the output elements b, one warp per element, lanes over the rows of column b
  for (int c = A.n_chunks - 1; c >= 0; --c) {
    const cplx* U = A.uc + (size_t)c * n * n;
    const cplx* W = A.wc + (size_t)c * n * n;
    const cplx *s_lq = s_state + cur * 2 * n, *s_lp = s_lq + n;
    for (int b = r0 + warp; b < r1; b += nwarps) {
      cplx aq = czero, ap = czero;
      for (int r2 = lane; r2 < n; r2 += 32) {
        const cplx u = U[(size_t)b * n + r2], w = W[(size_t)b * n + r2], lq = s_lq[r2], lp = s_lp[r2];
        aq = cfma_conj(w, lp, cfma_conj(u, lq, aq));
        ap = cfma_conj(u, lp, ap);
      }
      for (int off = 16; off > 0; off >>= 1) {
        aq.x += __shfl_xor_sync(0xffffffffu, aq.x, off); aq.y += __shfl_xor_sync(0xffffffffu, aq.y, off);
        ap.x += __shfl_xor_sync(0xffffffffu, ap.x, off); ap.y += __shfl_xor_sync(0xffffffffu, ap.y, off);
      }
      if (lane < WIDE_CC) {                    // lane k delivers the element to CTA k
        cplx* dst = cluster.map_shared_rank(s_state, lane) + (cur ^ 1) * 2 * n;
        dst[b] = aq; dst[n + b] = ap;
      }
      if (lane == 0) {
        A.lc[(size_t)c * 2 * n + b] = aq;
        A.lc[(size_t)c * 2 * n + n + b] = ap;
      }
    }
    cluster.sync();
    cur ^= 1;
  }
}

// ---------------------------------------------------------------------------------------------------
// Taylor degrees of the samples-minor kernel: one CTA per slice computes, for the whole propagated system
// [[A, B], [0, A]] (q = U psi and p = W psi on every basis state),
//   * the 1-norm bound k_coefficients computes serially (column sums of |nominal entry| + entrywise noise-deviation
//     bound) and its degree; the poison-and-flag rule lives here when k_coefficients skips its norm loop;
//   * a phase shift  phi_k = midpoint of the range of Im diag(X_k):  the sweeps run the Taylor series of
//     Y_k = X_k - i phi_k I  (exp(X_k) = e^{i phi_k} exp(Y_k); the phases of all slices commute with everything and are
//     applied once, to the target kets: k_wide_phase).  For the blockade chains the diagonal is the detuning times the
//     excitation number (0 .. N/2): the shift halves its contribution to the norm;
//   * a 2-norm bound of the shifted generator:  |Y(s)|_2 <= | |Y(s)| |_2 <= |N|_2 = sqrt(rho(N^T N))
//     <= sqrt(max_i (N^T N v)_i / v_i)  for ANY positive v (Collatz-Wielandt; N >= |Y(s)| entrywise for every noise
//     sample).  A few power steps make the bound tight; it is rigorous whatever v is.
// The degree is the one min(1-norm, 2-norm bound) of Y_k asks for (cfg 5: mean degree 15.2 -> 13.1).
// ---------------------------------------------------------------------------------------------------
struct WideNormArgs {
  int n, J, M, qcap, fresh, do_shift, u_kc, w_kc;
  double dt, tol;
  cplx* coef;
  const int *u_rowptr, *u_col, *u_colptr, *u_crow, *u_ceid, *u_egen;
  const int *w_rowptr, *w_col, *w_colptr, *w_crow, *w_ceid, *w_egen;
  const cplx *u_eval, *w_eval;
  const double *mulmax, *addmax;
  int *qdeg, *flag;
  double* shift;
};
#define WIDE_NORM_THREADS 256
#define WIDE_NORM_STEPS 5

__device__ __forceinline__ double wide_block_max(double v, double* red, int tid) {
  for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  double r = red[0];
  for (int w = 1; w < WIDE_NORM_THREADS / 32; ++w) r = fmax(r, red[w]);
  return r;
}

__global__ void __launch_bounds__(WIDE_NORM_THREADS) k_wide_degrees(const WideNormArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[WIDE_NORM_THREADS / 32];
  const int n = A.n, J = A.J, k = blockIdx.x, tid = threadIdx.x, T = WIDE_NORM_THREADS;
  const int nu = A.u_rowptr[n], nw = A.w_rowptr[n];
  double* mu = (double*)smem_raw;            // [nu] off-diagonal magnitudes of A (diagonal entries: 0, see dmag)
  double* mw = mu + nu;                      // [nw]
  double* dre = mw + nw;                     // [n] Re, Im of the nominal diagonal of A (times dt), its deviation bound
  double* dim = dre + n;
  double* ddev = dim + n;
  double* dmag = ddev + n;                   // [n] |shifted diagonal| + deviation
  double* vq = dmag + n;                     // [n] each
  double *vp = vq + n, *yq = vp + n, *yp = yq + n;
  cplx co[QOCVL_MAX_GEN];
  double wgt[QOCVL_MAX_GEN];
  for (int j = 0; j < J; ++j) {
    co[j] = A.coef[(size_t)k * J + j];
    wgt[j] = sqrt(co[j].x * co[j].x + co[j].y * co[j].y) * A.mulmax[J + j] + A.addmax[j];
  }
  for (int r = tid; r < n; r += T) { dre[r] = 0.0; dim[r] = 0.0; ddev[r] = 0.0; }
  __syncthreads();
  for (int r = tid; r < n; r += T) {                       // magnitudes, row by row (the row's diagonal entry apart)
    for (int fam = 0; fam < 2; ++fam) {
      const int e0 = fam ? A.w_rowptr[r] : A.u_rowptr[r], e1 = fam ? A.w_rowptr[r + 1] : A.u_rowptr[r + 1];
      const int kc = fam ? A.w_kc : A.u_kc;
      const int *egen = fam ? A.w_egen : A.u_egen, *col = fam ? A.w_col : A.u_col;
      const cplx* eval = fam ? A.w_eval : A.u_eval;
      for (int e = e0; e < e1; ++e) {
        double re = 0, im = 0, dev = 0;
        for (int c = 0; c < kc; ++c) {
          const int g = egen[e * kc + c];
          if (g >= 0) {
            const cplx v = eval[e * kc + c];
            re += co[g].x * v.x - co[g].y * v.y; im += co[g].x * v.y + co[g].y * v.x;
            dev += wgt[g] * sqrt(v.x * v.x + v.y * v.y);
          }
        }
        if (!fam && col[e] == r) { dre[r] = re * A.dt; dim[r] = im * A.dt; ddev[r] = dev * A.dt; mu[e] = 0.0; }
        else (fam ? mw : mu)[e] = (sqrt(re * re + im * im) + dev) * A.dt;
      }
    }
  }
  __syncthreads();
  // ---- unshifted 1-norm bound (qocvl_slice_norm_csr's) and the poison rule
  double cs = 0.0, lo = 1e300, hi = -1e300;
  for (int b = tid; b < n; b += T) {
    double sum = sqrt(dre[b] * dre[b] + dim[b] * dim[b]) + ddev[b];
    for (int i = A.u_colptr[b]; i < A.u_colptr[b + 1]; ++i) sum += mu[A.u_ceid[i]];
    for (int i = A.w_colptr[b]; i < A.w_colptr[b + 1]; ++i) sum += mw[A.w_ceid[i]];
    cs = fmax(cs, sum);
    lo = fmin(lo, dim[b]); hi = fmax(hi, dim[b]);
  }
  const double nrm_u = wide_block_max(cs, red, tid);
  int q_u = qocvl_taylor_degree(nrm_u, A.tol);
  const bool bad = !(nrm_u == nrm_u) || !(co[0].x == co[0].x) || !(nrm_u <= QOCVL_NORM_MAX) || q_u > A.qcap;
  if (bad) {                                               // (uniform: every thread sees the same nrm_u)
    if (tid == 0) {
      if (A.fresh) {
        atomicOr(A.flag, 1);
        A.coef[(size_t)k * J] = cmake(nan(""), nan(""));   // fail loudly: NaN cost and gradient, never silently wrong
        A.qdeg[k] = 2;
      }
      A.shift[k] = 0.0;
    }
    return;
  }
  // ---- phase shift: midpoint of the range of Im diag
  const double dhi = wide_block_max(hi, red, tid), dlo = -wide_block_max(-lo, red, tid);
  const double phi = A.do_shift ? 0.5 * (dhi + dlo) : 0.0;
  for (int r = tid; r < n; r += T) {
    const double im = dim[r] - phi;
    dmag[r] = sqrt(dre[r] * dre[r] + im * im) + ddev[r];
    vq[r] = 1.0; vp[r] = 1.0;
  }
  __syncthreads();
  // ---- shifted 1-norm
  cs = 0.0;
  for (int b = tid; b < n; b += T) {
    double sum = dmag[b];
    for (int i = A.u_colptr[b]; i < A.u_colptr[b + 1]; ++i) sum += mu[A.u_ceid[i]];
    for (int i = A.w_colptr[b]; i < A.w_colptr[b + 1]; ++i) sum += mw[A.w_ceid[i]];
    cs = fmax(cs, sum);
  }
  const double n1 = wide_block_max(cs, red, tid);
  // ---- Collatz-Wielandt bound of | N |_2,  N = [[|A|, 0], [|B|, |A|]] acting on (v_q, v_p)
  double best = 0.0;
  for (int it = 0; it <= WIDE_NORM_STEPS; ++it) {
    for (int r = tid; r < n; r += T) {                     // y = N v:  y_q = |A| v_q,  y_p = |B| v_q + |A| v_p
      double aq = dmag[r] * vq[r], ap = dmag[r] * vp[r];
      for (int e = A.u_rowptr[r]; e < A.u_rowptr[r + 1]; ++e) {
        const int c = A.u_col[e];
        aq = fma(mu[e], vq[c], aq); ap = fma(mu[e], vp[c], ap);
      }
      for (int e = A.w_rowptr[r]; e < A.w_rowptr[r + 1]; ++e) ap = fma(mw[e], vq[A.w_col[e]], ap);
      yq[r] = aq; yp[r] = ap;
    }
    __syncthreads();
    double mx = 0.0, ratio = 0.0, tq_[4], tp_[4];          // t = N^T y  (rows b = tid, tid + T, ...: at most 4 per thread here)
    int cnt = 0;
    for (int b = tid; b < n; b += T, ++cnt) {
      double tq = dmag[b] * yq[b], tp = dmag[b] * yp[b];
      for (int i = A.u_colptr[b]; i < A.u_colptr[b + 1]; ++i) {
        const double m = mu[A.u_ceid[i]];
        const int r = A.u_crow[i];
        tq = fma(m, yq[r], tq); tp = fma(m, yp[r], tp);
      }
      for (int i = A.w_colptr[b]; i < A.w_colptr[b + 1]; ++i) tq = fma(mw[A.w_ceid[i]], yp[A.w_crow[i]], tq);
      ratio = fmax(ratio, fmax(tq / vq[b], tp / vp[b]));
      mx = fmax(mx, fmax(tq, tp));
      if (cnt < 4) { tq_[cnt] = tq; tp_[cnt] = tp; }
    }
    if (it == WIDE_NORM_STEPS) { best = wide_block_max(ratio, red, tid); break; }
    mx = wide_block_max(mx, red, tid);                     // (its barriers also separate the reads of v from the writes)
    const double inv = mx > 0.0 ? 1.0 / mx : 0.0;
    cnt = 0;
    for (int b = tid; b < n; b += T, ++cnt) {              // next positive vector (the bound needs v > 0)
      vq[b] = fmax(tq_[cnt] * inv, 1e-3); vp[b] = fmax(tp_[cnt] * inv, 1e-3);
    }
    __syncthreads();
  }
  if (tid == 0) {
    double nb = n1;
    const double n2 = sqrt(best) * (1.0 + 1e-12);
    if (n2 == n2 && n2 > 0.0) nb = fmin(nb, n2);
    int q = qocvl_taylor_degree(nb, A.tol);
    if (q > q_u) q = q_u;
    if (!A.fresh && A.qdeg[k] < q) q = A.qdeg[k];          // (k_coefficients already ran its own rule: keep a poisoned slice's 2)
    A.qdeg[k] = q;
    A.shift[k] = phi;
  }
}

// Accumulated phase of the shifts: exp(X_M) ... exp(X_1) = e^{i Phi} exp(Y_M) ... exp(Y_1), Phi = sum_k phi_k.  The sweeps
// propagate z~ = e^{-i Phi} z; the cost reads z only through <y, z> (and q^dag p, which is phase free), so the target kets
// are rotated instead, y~ = e^{-i Phi} y: <y~, z~> = <y, z>, and the terminal cotangent the kernels derive from y~ is
// the cotangent of z~.  One block, fixed summation order (deterministic).
__global__ void __launch_bounds__(256) k_wide_phase(int M, int count, const double* __restrict__ shift,
                                                    const cplx* __restrict__ targets, cplx* __restrict__ rot) {
  __shared__ double part[256];
  double acc = 0.0;
  for (int k = threadIdx.x; k < M; k += 256) acc += shift[k];
  part[threadIdx.x] = acc;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) part[threadIdx.x] += part[threadIdx.x + off];
    __syncthreads();
  }
  double sn, cs;
  sincos(-part[0], &sn, &cs);
  const cplx ph = cmake(cs, sn);
  for (int i = threadIdx.x; i < count; i += 256) rot[i] = cmul(ph, targets[i]);
}

// Build the tables of the wide kernel.  QOCVL_ERR_UNSUPPORTED when the problem does not fit it (the caller then
// falls back to the one-CTA-per-sample kernel).
int qocvl_wide_configure(qocvl_plan* p) {
  const qocvl_desc& d = p->d;
  const int n = d.n, S = p->n_samples, M = d.n_slices;
  p->use_wide = false; p->use_wide_tp = false;
  if (p->n_live != 1 || p->adiab_live != 0) return QOCVL_ERR_UNSUPPORTED;
  int spc = (S >= 5) ? 8 : 1;
  if (p->has_opt("wide_spc")) spc = (p->opt("wide_spc") == 8) ? 8 : 1;
  WideLayout& L = p->lwide;
  int rc = wide_build(p, spc, L);
  if (rc) return rc;
  spc = L.spc;
  // ---- a single sample has no ensemble parallelism: cut the slice axis into chunks (column sweeps of every chunk
  //      with 8 basis columns per CTA -> combine -> per-chunk forward + reverse sweeps)
  bool tp = (S == 1 && M >= 64);
  if (p->has_opt("tp")) tp = tp && p->opt("tp") != 0;
  if (tp) {
    rc = wide_build(p, 8, p->lwide_a, true);
    if (rc == QOCVL_ERR_UNSUPPORTED || (rc == QOCVL_OK && p->lwide_a.spc != 8)) tp = false;
    else if (rc) return rc;
  }
  int lc = 0, nchunks = 0;
  if (tp) {
    lc = (M + 63) / 64;                                // ~64 chunks: 64 * ceil(n/8) column CTAs fill the machine
    lc = std::max(lc, 8);
    if (p->has_opt("tp_chunk")) lc = std::max(1, std::min(M, p->opt("tp_chunk")));
    nchunks = (M + lc - 1) / lc;
    const size_t blk = (size_t)nchunks * n * n, bnd = (size_t)(nchunks + 1) * WIDE_NVP * n;
    for (cplx** q : {&p->d_tp_uc, &p->d_tp_wc, &p->d_tp_zc, &p->d_tp_lc}) if (*q) { cudaFree(*q); *q = nullptr; }
    QCUDA(cudaMalloc((void**)&p->d_tp_uc, blk * sizeof(cplx)));
    QCUDA(cudaMalloc((void**)&p->d_tp_wc, blk * sizeof(cplx)));
    QCUDA(cudaMalloc((void**)&p->d_tp_zc, bnd * sizeof(cplx)));
    QCUDA(cudaMalloc((void**)&p->d_tp_lc, bnd * sizeof(cplx)));
    p->bytes += (2 * blk + 2 * bnd) * sizeof(cplx);
    p->tp_chunk = lc; p->tp_nchunks = nchunks;
    p->tp_blocks_a = nchunks * ((n + 7) / 8);
    p->use_wide_tp = true;
  }
  // ---- geometry and buffers
  const int ctas = tp ? nchunks : (S + spc - 1) / spc;
  const int cta_slices = tp ? lc : M;
  p->chain_smem = L.smem;
  p->threads = L.threads;
  p->blocks = ctas;
  p->spb = spc; p->groups = 1;
  p->n_parts = (size_t)ctas;
  const size_t vs = (size_t)WIDE_NVP * (n + 1) * spc * (p->c64 ? sizeof(cplxf) : sizeof(cplx));   // elements of the sweeps' type
  const size_t stream_bytes = (size_t)ctas * cta_slices * p->qcap * vs;
  size_t free_b = 0, total_b = 0;
  QCUDA(cudaMemGetInfo(&free_b, &total_b));
  bool store = stream_bytes <= (size_t)(0.6 * (double)free_b);
  if (p->has_opt("wide_store")) store = store && p->opt("wide_store") != 0;
  if (p->d_ckpt) { cudaFree(p->d_ckpt); p->d_ckpt = nullptr; p->bytes -= p->ckpt_bytes; p->ckpt_bytes = 0; }
  if (p->d_tring) { cudaFree(p->d_tring); p->d_tring = nullptr; p->bytes -= p->tring_bytes; p->tring_bytes = 0; }
  p->tring_bytes = store ? stream_bytes : (size_t)ctas * p->qcap * vs;
  QCUDA(cudaMalloc((void**)&p->d_tring, p->tring_bytes));
  if (!store) {
    p->ckpt_bytes = (size_t)ctas * cta_slices * vs;
    QCUDA(cudaMalloc((void**)&p->d_ckpt, p->ckpt_bytes));
  }
  p->bytes += p->ckpt_bytes + p->tring_bytes;
  L.store = store;
  // ---- Taylor degrees by k_wide_degrees (phase shift + 2-norm bound) when its tables fit shared memory ("wide_norms" = 0:
  //      the serial 1-norm rule of k_coefficients, no shift)
  {
    const size_t need = ((size_t)p->cu.nnz + p->cw.nnz + 8 * (size_t)n) * sizeof(double);
    p->wide_norms = need <= 200 * 1024 && n <= 4 * WIDE_NORM_THREADS && p->opt("wide_norms", 1) != 0;
    p->wide_norm_smem = p->wide_norms ? need : 0;
    if (p->wide_norms && need > 48 * 1024)
      QCUDA(cudaFuncSetAttribute(k_wide_degrees, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    if (p->d_wide_shift) { cudaFree(p->d_wide_shift); p->d_wide_shift = nullptr; }
    if (p->d_targets_rot) { cudaFree(p->d_targets_rot); p->d_targets_rot = nullptr; }
    if (p->wide_norms) {
      QCUDA(cudaMalloc((void**)&p->d_wide_shift, (size_t)M * sizeof(double)));
      QCUDA(cudaMemset(p->d_wide_shift, 0, (size_t)M * sizeof(double)));
      QCUDA(cudaMalloc((void**)&p->d_targets_rot, (size_t)n * sizeof(cplx)));
    }
  }
  p->use_wide = true;
  return QOCVL_OK;
}

static void wide_fill_args(const qocvl_plan* p, const WideLayout& L, WideArgs& a) {
  const qocvl_desc& d = p->d;
  memset(&a, 0, sizeof(a));
  a.n = d.n; a.J = d.n_gen; a.M = d.n_slices; a.C = d.n_chan; a.S = p->n_samples;
  a.qcap = p->qcap;
  a.ngroups = L.ngroups; a.nb = L.nb; a.nd = L.nd;
  a.ent_fwd = L.ent_fwd; a.ent_rev = L.ent_rev; a.dreal = L.dreal ? 1 : 0;
  a.dt = d.dt; a.inv_duration = 1.0 / d.duration; a.w_infid = d.w_infid; a.w_adiab = d.w_adiab;
  a.inv_fid_norm = 1.0 / d.fid_norm; a.inv_samples = 1.0 / (double)p->global_samples;
  a.d_const = p->d_const;
  for (int i = 0; i < WIDE_MAXB; ++i) { a.bgen[i] = L.bgen[i]; a.bfam[i] = L.bfam[i]; a.buni[i] = L.buni[i]; a.bval[i] = L.bval[i]; }
  for (int i = 0; i < WIDE_MAXD; ++i) { a.dgen[i] = L.dgen[i]; a.dfam[i] = L.dfam[i]; }
  a.grow = L.grow; a.fwd_co = L.fwd_co; a.rev_co = L.rev_co;
  a.fwd_val = L.fwd_val; a.fwd_col = L.fwd_col; a.rev_val = L.rev_val; a.rev_col = L.rev_col;
  a.dval = L.dval; a.dval_re = L.dval_re; a.dmask = L.dmask;
  a.coef = p->d_coef; a.dcoef = p->d_dcoef; a.qdeg = p->d_qdeg; a.mul = p->d_mul; a.add = p->d_add;
  a.starts = p->d_starts; a.targets = p->wide_norms ? p->d_targets_rot : p->d_targets;
  a.shift = p->wide_norms ? p->d_wide_shift : nullptr;
  a.ring = p->d_tring; a.ckpt = p->d_ckpt;
  a.sample_out = p->d_sample_out; a.gw_partial = p->d_gw_partial;
  a.chunk_len = p->tp_chunk; a.n_chunks = p->tp_nchunks; a.ctas_per_chunk = (d.n + 7) / 8;
  a.tp_uc = p->d_tp_uc; a.tp_wc = p->d_tp_wc; a.tp_zc = p->d_tp_zc; a.tp_lc = p->d_tp_lc;
}

// gwave: [M][C] device buffer; the time-parallel path writes d cost / d wave into it directly
int qocvl_wide_launch(qocvl_plan* p, bool want_grad, double* gwave, cudaStream_t st) {
  const WideLayout& L = p->lwide;
  if (p->wide_norms) {
    WideNormArgs na;
    memset(&na, 0, sizeof(na));
    na.n = p->d.n; na.J = p->d.n_gen; na.M = p->d.n_slices; na.qcap = p->qcap;
    na.fresh = p->wide_norms_fresh ? 1 : 0;                 // k_coefficients skipped its norm loop for this evaluation
    na.do_shift = p->opt("wide_shift", 1) != 0 ? 1 : 0;
    na.u_kc = p->cu.kc; na.w_kc = p->cw.kc; na.dt = p->d.dt; na.tol = p->taylor_tol();
    na.coef = p->d_coef;
    na.u_rowptr = p->cu.rowptr; na.u_col = p->cu.col; na.u_colptr = p->cu.colptr; na.u_crow = p->cu.crow;
    na.u_ceid = p->cu.ceid; na.u_egen = p->cu.egen; na.u_eval = p->cu.eval;
    na.w_rowptr = p->cw.rowptr; na.w_col = p->cw.col; na.w_colptr = p->cw.colptr; na.w_crow = p->cw.crow;
    na.w_ceid = p->cw.ceid; na.w_egen = p->cw.egen; na.w_eval = p->cw.eval;
    na.mulmax = p->d_mulmax; na.addmax = p->d_addmax; na.qdeg = p->d_qdeg; na.flag = p->d_flag; na.shift = p->d_wide_shift;
    k_wide_degrees<<<p->d.n_slices, WIDE_NORM_THREADS, p->wide_norm_smem, st>>>(na);
    k_wide_phase<<<1, 256, 0, st>>>(p->d.n_slices, p->d.n, p->d_wide_shift, p->d_targets, p->d_targets_rot);
    p->launches += 2;
  }
  WideArgs a;
  wide_fill_args(p, L, a);
  a.want_grad = want_grad ? 1 : 0; a.store = L.store ? 1 : 0;
  wide_kernel_t kern = wide_pick(p->c64, L.spc, L.items, L.max_threads, L.nb, L.nd, L.simple);
  if (!p->use_wide_tp) {
    kern<<<p->blocks, p->threads, p->chain_smem, st>>>(a);
    p->launches++;
    QCUDA(cudaGetLastError());
    return QOCVL_OK;
  }
  // ---- time-parallel: column sweeps -> combine -> chunk sweeps
  const WideLayout& LA = p->lwide_a;
  WideArgs pa;
  wide_fill_args(p, LA, pa);
  pa.mode = 1; pa.want_grad = 0; pa.store = 0;
  wide_pick(p->c64, LA.spc, LA.items, LA.max_threads, LA.nb, LA.nd, LA.simple, LA.fwd)<<<p->tp_blocks_a, LA.threads, LA.smem, st>>>(pa);
  WideCombineArgs c;
  c.n = p->d.n; c.n_chunks = p->tp_nchunks; c.want_grad = want_grad ? 1 : 0;
  c.inv_duration = a.inv_duration; c.w_infid = a.w_infid; c.w_adiab = a.w_adiab; c.inv_fid_norm = a.inv_fid_norm;
  c.inv_samples = a.inv_samples; c.d_const = a.d_const;
  c.uc = p->d_tp_uc; c.wc = p->d_tp_wc; c.starts = p->d_starts; c.targets = a.targets;
  c.zc = p->d_tp_zc; c.lc = p->d_tp_lc; c.sample_out = p->d_sample_out;
  const int per = (p->d.n + WIDE_CC - 1) / WIDE_CC;
  const int cthreads = std::min(1024, std::max(128, (per * std::max(1, 512 / per) + 31) / 32 * 32));
  const int parts = std::max(1, cthreads / per);
  const size_t csmem = ((size_t)4 * p->d.n + (size_t)parts * 2 * per) * sizeof(cplx) + 3 * 32 * sizeof(double);
  if (csmem > 48 * 1024) QCUDA(cudaFuncSetAttribute(k_wide_combine, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
  k_wide_combine<<<WIDE_CC, cthreads, csmem, st>>>(c);
  p->launches += 2;
  if (want_grad) {
    a.mode = 2; a.tp_gwave = gwave;
    kern<<<p->tp_nchunks, p->threads, p->chain_smem, st>>>(a);
    p->launches++;
  }
  QCUDA(cudaGetLastError());
  return QOCVL_OK;
}
