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
                              size_t ent_fwd, size_t ent_rev, bool dreal) {
  const int rw = 32 / spc, nslot = nb_c + nd_c;
  const size_t vs = (size_t)WIDE_NVP * n * spc;
  size_t cplx_count = 2 * vs + 2 * (size_t)J * spc + (size_t)J * spc + (size_t)J * C +
                      (size_t)(threads / 32) * nslot * spc + (size_t)nslot * spc +
                      (dreal ? 0 : (size_t)nd_c * ngroups * rw);
  size_t dbl_count = (size_t)J * spc + (size_t)(threads / 32) * 3 * spc + 2 * (size_t)spc +
                     (dreal ? (size_t)nd_c * ngroups * rw : 0);
  return cplx_count * sizeof(cplx) + dbl_count * sizeof(double) + 2 * (size_t)nb_c * ngroups * sizeof(int2) +
         (((ent_fwd + 7) & ~(size_t)7) + ((ent_rev + 7) & ~(size_t)7)) * sizeof(unsigned short) +
         (((size_t)ngroups * rw + 15) & ~(size_t)15);
}

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
static wide_kernel_t wide_pick(int spc, int items, int threads, int nb, int nd, bool simple) {
  if (simple) return qocvl_wide_pick_b13(spc, items, threads);
  if (nb <= WIDE_MAXB && nd <= WIDE_MAXD) return qocvl_wide_pick_b24(spc, items, threads);
  return nullptr;
}

void WideLayout::release() {
  void* ptrs[] = {grow, fwd_co, rev_co, fwd_val, fwd_col, rev_val, rev_col, dval, dval_re, dmask};
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
          const size_t lpad = (longest + 3) / 4 * 4;      // lists padded to whole 4-entry chunks
          T.co[dir][(size_t)bi * T.ngroups + g] = make_int2((int)(lpad / 4), (int)T.col[dir].size());
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
    kern = wide_pick(spc, items, max_threads, nb, nd, simple);
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
  if ((rc = wide_upload(&L.dmask, T.dmask))) return rc;
  L.simple = simple;
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
  a.fwd_val = L.fwd_val; a.fwd_col = L.fwd_col; a.rev_val = L.rev_val; a.rev_col = L.rev_col; a.dval = L.dval; a.dval_re = L.dval_re; a.dmask = L.dmask;
  a.coef = p->d_coef; a.dcoef = p->d_dcoef; a.qdeg = p->d_qdeg; a.mul = p->d_mul; a.add = p->d_add;
  a.starts = p->d_starts; a.targets = p->d_targets;
  a.ring = p->d_tring; a.ckpt = p->d_ckpt;
  a.sample_out = p->d_sample_out; a.gw_partial = p->d_gw_partial;
  wide_kernel_t kern = wide_pick(L.spc, L.items, L.max_threads, L.nb, L.nd, L.simple);
  kern<<<p->blocks, p->threads, p->chain_smem, st>>>(a);
  p->launches++;
  QCUDA(cudaGetLastError());
  return QOCVL_OK;
}
