// chain_rows.cu -- host side of the sector-per-warp ensemble kernel (chain_rows.cuh): roles, launch geometry, workspaces,
// and the fixed-order reduction of the per-warp gradient partials.  The kernel instances live in chain_rows_k1.cu /
// chain_rows_k2.cu (one translation unit per class count so they compile in parallel).
#include "chain_rows.cuh"
#include <algorithm>
#include <numeric>

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

static rows_fn pick_rows(int wmax, int kc, bool store, bool tma, bool split) {
  if (wmax > 6) return nullptr;
  if (kc == 1) return qocvl_rows_kernel_k1(wmax, store, tma, split);
  if (kc == 2) return qocvl_rows_kernel_k2(wmax, store, tma);
  return nullptr;
}

static size_t rows_smem_bytes(const qocvl_plan* p, int wmax) {
  const RowsGeom& g = p->rg;
  const int nsl = 1 + (wmax <= 2 ? 2 : (wmax <= 3 ? 3 : 6));
  size_t b = 0;
  if (p->aug_store) b += (size_t)g.wpc * (g.tma ? 2 * p->aug_qs : ROWS_RING) * 32 * sizeof(cplx);
  b += (size_t)g.wpc * 2 * sizeof(unsigned long long);
  b += (size_t)g.G * p->aug_n * sizeof(cplx) + ((size_t)g.G * p->aug_n * sizeof(double) + 15) / 16 * 16;
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
  if (!pick_rows(wmax, la.KC, false, false, false)) return QOCVL_ERR_UNSUPPORTED;
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
  g.cap = (size_t)M * p->aug_qs;
  // bulk TMA keeps two whole slabs (a-priori Taylor cap x 512 B) per warp in shared memory: only when that is modest
  if ((size_t)2 * p->aug_qs * 32 * sizeof(cplx) * g.wpc > 64 * 1024) g.tma = false;
  // few warps per SM (small shards): the two-accumulator body shortens the dependent chain of every Taylor step
  {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device);
    g.split = la.KC == 1 && (double)g.total_warps / sms < 6.0;
    if (p->has_opt("rows_split")) g.split = la.KC == 1 && p->opt("rows_split") != 0;
  }
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
  rows_fn fn = pick_rows(wmax, la.KC, p->aug_store, g.tma, g.split);
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
  q.qs = p->aug_qs;
  q.ifact[0] = 1.0;
  for (int j = 1; j < QOCVL_QMAX + 2; ++j) q.ifact[j] = q.ifact[j - 1] / (double)j;
  k_aug_tables<cplx><<<p->d.n_slices, 256, 0, st>>>(a, la.L);
  pick_rows(g.wmax, la.KC, p->aug_store, g.tma, g.split)<<<p->blocks, p->threads, p->chain_smem, st>>>(q);
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
