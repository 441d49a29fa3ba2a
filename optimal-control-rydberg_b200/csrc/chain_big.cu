// chain_big.cu -- the same cost + gradient recurrences as chain_small.cu (see its header for the maths) for
// Hilbert spaces that do not fit one row per lane: ONE CTA PER NOISE SAMPLE, rows strided over the threads,
// all vectors of the sample and the assembled sparse values of A_k, B_k resident in shared memory, CSR
// (row-owner) mat-vecs forward and CSC (column-owner) mat-vecs in the reverse sweep, the Taylor-term ring in
// global memory (L2-resident: it is re-written every slice), one __syncthreads per Taylor step.
// Target: the blockade-subspace chains of BASELINE.json configs 4-5 (n = 144, 377); also runs every n <= 16
// problem (QOCVL_FORCE_BIG=1), which is how it is cross-checked against the lane-per-row kernel.
#include "common.cuh"
#include <algorithm>

struct BigArgs {
  int n, J, M, C, S, nv, va, qcap, want_grad, nnz_u, nnz_w, kc_u, kc_w;
  double dt, inv_duration, w_infid, w_adiab, inv_fid_norm, inv_samples;
  cplx d_const;
  // u family
  const int *u_rowptr, *u_col, *u_erow, *u_colptr, *u_crow, *u_ceid, *u_egen;
  const cplx* u_eval;
  // w family
  const int *w_rowptr, *w_col, *w_erow, *w_colptr, *w_crow, *w_ceid, *w_egen;
  const cplx* w_eval;
  const cplx* coef;    // [M][J]
  const cplx* dcoef;   // [M][J][C]
  const int* qdeg;     // [M]
  const double* mul;   // [S][J]
  const cplx* add;     // [S][J]
  const cplx* starts;  // [nv][n]
  const cplx* targets; // [nv][n]
  cplx* ckpt;          // [S][M][nv+1][n]
  cplx* tring;         // [S][qcap][nv+1][n]
  double* sample_out;  // [S][3]
  double* gw_partial;  // [S][M][C]
};

// block-wide sum of up to 6 doubles; result valid in every thread
__device__ __forceinline__ void block_sum(double* vals, int count, double* s_scratch) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = (blockDim.x + 31) >> 5;
  for (int i = 0; i < count; ++i) {
    double v = vals[i];
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) s_scratch[i * 32 + warp] = v;
  }
  __syncthreads();
  for (int i = 0; i < count; ++i) {
    double v = 0.0;
    for (int w = 0; w < nwarps; ++w) v += s_scratch[i * 32 + w];
    vals[i] = v;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(512) k_chain_big(const BigArgs P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int n = P.n, J = P.J, C = P.C, JC = P.J * P.C, nvp = P.nv + 1, va = P.va, NV = P.nv;
  const int sample = blockIdx.x;
  const int vn = nvp * n;                       // elements of one "all vectors" buffer
  // ---- shared memory
  cplx* carve = (cplx*)smem_raw;
  cplx* s_au = carve; carve += P.nnz_u;         // assembled dt * A_k values, entry (row) order
  cplx* s_aw = carve; carve += P.nnz_w;
  cplx* s_xu = carve; carve += P.want_grad ? P.nnz_u : 0;   // Xbar accumulators on the pattern
  cplx* s_xw = carve; carve += P.want_grad ? P.nnz_w : 0;
  cplx* s_z = carve; carve += vn;               // state / cotangent
  cplx* s_t = carve; carve += vn;               // current Taylor term
  cplx* s_y = carve; carve += vn;               // next Taylor term
  cplx* s_p = carve; carve += P.want_grad ? vn : 0;         // t_{j-1} staged for the contraction
  cplx* s_cs = carve; carve += J;
  cplx* s_dco = carve; carve += JC;
  double* s_mul = (double*)carve;               // [J]
  double* s_scr = s_mul + J;                    // [6*32] reduction scratch
  const cplx czero = cmake(0, 0);

  for (int j = tid; j < J; j += nt) s_mul[j] = P.mul[(size_t)sample * J + j];
  for (int i = tid; i < vn; i += nt) {
    const int v = i / n, r = i % n;
    s_z[i] = (v < NV) ? P.starts[v * n + r] : czero;
  }
  __syncthreads();

  // values of the slice:  cs_j = coef_kj * mul_sj + add_sj ;  A[e] = dt * sum_c cs[gen] * val
  auto assemble = [&](int k) {
    for (int j = tid; j < J; j += nt)
      s_cs[j] = cadd(cscale(P.coef[(size_t)k * J + j], s_mul[j]), P.add[(size_t)sample * J + j]);
    __syncthreads();
    for (int e = tid; e < P.nnz_u; e += nt) {
      cplx acc = czero;
      for (int c = 0; c < P.kc_u; ++c) {
        const int g = P.u_egen[e * P.kc_u + c];
        if (g >= 0) acc = cfma(s_cs[g], P.u_eval[e * P.kc_u + c], acc);
      }
      s_au[e] = cscale(acc, P.dt);
    }
    for (int e = tid; e < P.nnz_w; e += nt) {
      cplx acc = czero;
      for (int c = 0; c < P.kc_w; ++c) {
        const int g = P.w_egen[e * P.kc_w + c];
        if (g >= 0) acc = cfma(s_cs[g], P.w_eval[e * P.kc_w + c], acc);
      }
      s_aw[e] = cscale(acc, P.dt);
    }
    __syncthreads();
  };
  // y = X t / j   (rows strided over threads; the p vector, index NV, also receives B t_q)
  auto matvec = [&](const cplx* t, cplx* y, double invj) {
    for (int i = tid; i < vn; i += nt) {
      const int v = i / n, r = i % n;
      cplx acc = czero;
      for (int e = P.u_rowptr[r]; e < P.u_rowptr[r + 1]; ++e) acc = cfma(s_au[e], t[v * n + P.u_col[e]], acc);
      if (v == NV)
        for (int e = P.w_rowptr[r]; e < P.w_rowptr[r + 1]; ++e) acc = cfma(s_aw[e], t[va * n + P.w_col[e]], acc);
      y[i] = cscale(acc, invj);
    }
  };

  cplx* ckpt = P.ckpt + (size_t)sample * P.M * vn;
  // =================================== forward sweep ===================================
  for (int k = 0; k < P.M; ++k) {
    const int q = P.qdeg[k];
    assemble(k);
    for (int i = tid; i < vn; i += nt) {
      const cplx zz = s_z[i];
      if (P.want_grad) ckpt[(size_t)k * vn + i] = zz;
      s_t[i] = zz;
    }
    __syncthreads();
    cplx *t = s_t, *y = s_y;
    for (int j = 1; j <= q; ++j) {
      matvec(t, y, 1.0 / (double)j);
      for (int i = tid; i < vn; i += nt) s_z[i] = cadd(s_z[i], y[i]);
      __syncthreads();
      cplx* tmp = t; t = y; y = tmp;
    }
  }
  // =================================== cost ============================================
  double red[4] = {0, 0, 0, 0};   // Re d, Im d, Re q^dag p, -
  for (int i = tid; i < NV * n; i += nt) {
    const cplx y = P.targets[i], z = s_z[i];
    red[0] += y.x * z.x + y.y * z.y;
    red[1] += y.x * z.y - y.y * z.x;
  }
  for (int r = tid; r < n; r += nt) {
    const cplx qv = s_z[va * n + r], pv = s_z[NV * n + r];
    red[2] += qv.x * pv.x + qv.y * pv.y;
  }
  block_sum(red, 3, s_scr);
  const cplx dsum = cmake(red[0] + P.d_const.x, red[1] + P.d_const.y);
  if (tid == 0) {
    const double infid = 1.0 - (dsum.x * dsum.x + dsum.y * dsum.y) * P.inv_fid_norm;
    const double adiab = 1.0 - red[2] * P.inv_duration;
    P.sample_out[(size_t)sample * 3 + 0] = P.w_infid * infid + P.w_adiab * adiab;
    P.sample_out[(size_t)sample * 3 + 1] = infid;
    P.sample_out[(size_t)sample * 3 + 2] = adiab;
  }
  if (!P.want_grad) return;

  // terminal cotangents into s_z (lam), scaled by 1/S_global
  {
    const double ca = -0.5 * P.w_adiab * P.inv_duration * P.inv_samples;
    const double cf = -P.w_infid * P.inv_fid_norm * P.inv_samples;
    __syncthreads();
    for (int r = tid; r < n; r += nt) { s_y[r] = s_z[va * n + r]; s_y[n + r] = s_z[NV * n + r]; }   // q, p
    __syncthreads();
    for (int i = tid; i < vn; i += nt) {
      const int v = i / n, r = i % n;
      cplx lam;
      if (v == NV) lam = cscale(s_y[r], ca);                                   // -(wa/2T) q
      else {
        lam = cscale(cmul(dsum, P.targets[i]), cf);
        if (v == va) lam = cadd(lam, cscale(s_y[n + r], ca));                  // -(wa/2T) p
      }
      s_z[i] = lam;
    }
    __syncthreads();
  }

  // =================================== reverse sweep ===================================
  cplx* ring = P.tring + (size_t)sample * P.qcap * vn;
  for (int k = P.M - 1; k >= 0; --k) {
    const int q = P.qdeg[k];
    assemble(k);
    for (int i = tid; i < JC; i += nt) s_dco[i] = P.dcoef[(size_t)k * JC + i];
    for (int e = tid; e < P.nnz_u; e += nt) s_xu[e] = czero;
    for (int e = tid; e < P.nnz_w; e += nt) s_xw[e] = czero;
    // ---- recompute forward Taylor terms t_0 .. t_{q-1} into the ring
    for (int i = tid; i < vn; i += nt) {
      const cplx zz = ckpt[(size_t)k * vn + i];
      s_t[i] = zz;
      ring[i] = zz;
    }
    __syncthreads();
    cplx *t = s_t, *y = s_y;
    for (int j = 1; j < q; ++j) {
      matvec(t, y, 1.0 / (double)j);
      for (int i = tid; i < vn; i += nt) ring[(size_t)j * vn + i] = y[i];
      __syncthreads();
      cplx* tmp = t; t = y; y = tmp;
    }
    // ---- descending adjoint recurrence: tb in t, next tb in y, lam stays in s_z
    for (int i = tid; i < vn; i += nt) t[i] = s_z[i];
    __syncthreads();
    for (int j = q; j >= 1; --j) {
      const double invj = 1.0 / (double)j;
      for (int i = tid; i < vn; i += nt) s_p[i] = cscale(ring[(size_t)(j - 1) * vn + i], invj);
      __syncthreads();
      // Xbar on the pattern: sum over vectors of conj(tb[row]) * t_{j-1}[col] / j
      for (int e = tid; e < P.nnz_u; e += nt) {
        const int a = P.u_erow[e], b = P.u_col[e];
        cplx acc = s_xu[e];
        for (int v = 0; v < nvp; ++v) acc = cfma_conj(t[v * n + a], s_p[v * n + b], acc);
        s_xu[e] = acc;
      }
      for (int e = tid; e < P.nnz_w; e += nt) {     // B couples tb_p (row) with t_q (col)
        const int a = P.w_erow[e], b = P.w_col[e];
        s_xw[e] = cfma_conj(t[NV * n + a], s_p[va * n + b], s_xw[e]);
      }
      // tb_{j-1} = lam + X^dag tb_j / j   (column-owner gathers; q also receives B^dag tb_p)
      for (int i = tid; i < vn; i += nt) {
        const int v = i / n, c = i % n;
        cplx acc = czero;
        for (int x = P.u_colptr[c]; x < P.u_colptr[c + 1]; ++x)
          acc = cfma_conj(s_au[P.u_ceid[x]], t[v * n + P.u_crow[x]], acc);
        if (v == va)
          for (int x = P.w_colptr[c]; x < P.w_colptr[c + 1]; ++x)
            acc = cfma_conj(s_aw[P.w_ceid[x]], t[NV * n + P.w_crow[x]], acc);
        y[i] = cadd(s_z[i], cscale(acc, invj));
      }
      __syncthreads();
      cplx* tmp = t; t = y; y = tmp;
    }
    for (int i = tid; i < vn; i += nt) s_z[i] = t[i];
    // ---- contract Xbar with the generators: d cost / d wave[k][c]
    double gw[QOCVL_MAX_CHAN];
    for (int c = 0; c < QOCVL_MAX_CHAN; ++c) gw[c] = 0.0;
    const double two_dt = 2.0 * P.dt;
    for (int e = tid; e < P.nnz_u; e += nt)
      for (int kc = 0; kc < P.kc_u; ++kc) {
        const int g = P.u_egen[e * P.kc_u + kc];
        if (g >= 0) {
          const cplx tmp = cscale(cmul(P.u_eval[e * P.kc_u + kc], s_xu[e]), two_dt * s_mul[g]);
          for (int c = 0; c < C; ++c) gw[c] += tmp.x * s_dco[g * C + c].x - tmp.y * s_dco[g * C + c].y;
        }
      }
    for (int e = tid; e < P.nnz_w; e += nt)
      for (int kc = 0; kc < P.kc_w; ++kc) {
        const int g = P.w_egen[e * P.kc_w + kc];
        if (g >= 0) {
          const cplx tmp = cscale(cmul(P.w_eval[e * P.kc_w + kc], s_xw[e]), two_dt * s_mul[g]);
          for (int c = 0; c < C; ++c) gw[c] += tmp.x * s_dco[g * C + c].x - tmp.y * s_dco[g * C + c].y;
        }
      }
    block_sum(gw, C, s_scr);
    if (tid < C) P.gw_partial[((size_t)sample * P.M + k) * C + tid] = gw[tid];
  }
}

// ---------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------
template <typename T>
static int upload(T** dst, const std::vector<T>& src) {
  if (*dst) { cudaFree(*dst); *dst = nullptr; }
  QCUDA(cudaMalloc((void**)dst, std::max<size_t>(1, src.size()) * sizeof(T)));
  if (!src.empty()) QCUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
  return QOCVL_OK;
}

// CSR / CSC tables of the union pattern of one block family  blocks[J][n][n]
int qocvl_build_csr(qocvl_plan* p, const std::vector<std::complex<double>>& blocks, CsrLayout* out) {
  const int n = p->d.n, J = p->d.n_gen;
  auto G = [&](int j, int a, int b) { return blocks[((size_t)j * n + a) * n + b]; };
  std::vector<int> rowptr(n + 1, 0), col, erow, colptr(n + 1, 0), crow, ceid;
  int kc = 1;
  for (int a = 0; a < n; ++a) {
    for (int b = 0; b < n; ++b) {
      int cnt = 0;
      for (int j = 0; j < J; ++j) cnt += std::abs(G(j, a, b)) > 0.0;
      if (cnt) { col.push_back(b); erow.push_back(a); kc = std::max(kc, cnt); }
    }
    rowptr[a + 1] = (int)col.size();
  }
  const int nnz = (int)col.size();
  std::vector<int> egen((size_t)nnz * kc, -1);
  std::vector<cplx> eval((size_t)nnz * kc, cmake(0, 0));
  for (int e = 0; e < nnz; ++e) {
    int c = 0;
    for (int j = 0; j < J; ++j) {
      const std::complex<double> v = G(j, erow[e], col[e]);
      if (std::abs(v) > 0.0) { egen[(size_t)e * kc + c] = j; eval[(size_t)e * kc + c] = cmake(v.real(), v.imag()); ++c; }
    }
  }
  std::vector<std::vector<int>> bycol(n);
  for (int e = 0; e < nnz; ++e) bycol[col[e]].push_back(e);
  for (int b = 0; b < n; ++b) {
    for (int e : bycol[b]) { crow.push_back(erow[e]); ceid.push_back(e); }
    colptr[b + 1] = (int)crow.size();
  }
  out->n = n; out->nnz = nnz; out->kc = kc;
  out->h_colptr = colptr; out->h_ceid = ceid; out->h_egen = egen; out->h_eval = eval;
  int rc;
  if ((rc = upload(&out->rowptr, rowptr))) return rc;
  if ((rc = upload(&out->col, col))) return rc;
  if ((rc = upload(&out->erow, erow))) return rc;
  if ((rc = upload(&out->colptr, colptr))) return rc;
  if ((rc = upload(&out->crow, crow))) return rc;
  if ((rc = upload(&out->ceid, ceid))) return rc;
  if ((rc = upload(&out->egen, egen))) return rc;
  if ((rc = upload(&out->eval, eval))) return rc;
  return QOCVL_OK;
}

static size_t big_smem_bytes(const qocvl_plan* p, bool want_grad) {
  const int n = p->d.n, J = p->d.n_gen, C = p->d.n_chan, vn = (p->n_live + 1) * n;
  size_t cplx_count = (size_t)(p->cu.nnz + p->cw.nnz) * (want_grad ? 2 : 1) + (size_t)vn * (want_grad ? 4 : 3) + J + J * C;
  return cplx_count * sizeof(cplx) + (size_t)(J + 6 * 32) * sizeof(double);
}

int qocvl_big_configure(qocvl_plan* p) {
  const qocvl_desc& d = p->d;
  const int S = p->n_samples, vn = (p->n_live + 1) * d.n;
  const size_t smem = big_smem_bytes(p, true);
  if (smem > 227 * 1024) return QOCVL_ERR_UNSUPPORTED;
  QCUDA(cudaFuncSetAttribute(k_chain_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  p->chain_smem = smem;
  p->threads = std::min(512, std::max(64, (vn + 31) / 32 * 32));   // one (vector, row) output per thread when it fits
  p->blocks = S;
  p->spb = 1; p->groups = 1;
  p->n_parts = (size_t)S;
  const size_t ck = (size_t)S * d.n_slices * vn;
  if (p->d_ckpt) { cudaFree(p->d_ckpt); p->d_ckpt = nullptr; p->bytes -= p->ckpt_bytes; }
  QCUDA(cudaMalloc((void**)&p->d_ckpt, ck * sizeof(cplx)));
  p->ckpt_bytes = ck * sizeof(cplx);
  if (p->d_tring) { cudaFree(p->d_tring); p->d_tring = nullptr; p->bytes -= p->tring_bytes; }
  p->tring_bytes = (size_t)S * p->qcap * vn * sizeof(cplx);
  QCUDA(cudaMalloc((void**)&p->d_tring, p->tring_bytes));
  p->bytes += p->ckpt_bytes + p->tring_bytes;
  return QOCVL_OK;
}

int qocvl_big_launch(qocvl_plan* p, bool want_grad, cudaStream_t st) {
  const qocvl_desc& d = p->d;
  BigArgs a;
  a.n = d.n; a.J = d.n_gen; a.M = d.n_slices; a.C = d.n_chan; a.S = p->n_samples;
  a.nv = p->n_live; a.va = p->adiab_live; a.qcap = p->qcap; a.want_grad = want_grad ? 1 : 0;
  a.nnz_u = p->cu.nnz; a.nnz_w = p->cw.nnz; a.kc_u = p->cu.kc; a.kc_w = p->cw.kc;
  a.dt = d.dt; a.inv_duration = 1.0 / d.duration; a.w_infid = d.w_infid; a.w_adiab = d.w_adiab;
  a.inv_fid_norm = 1.0 / d.fid_norm; a.inv_samples = 1.0 / (double)p->global_samples;
  a.d_const = p->d_const;
  a.u_rowptr = p->cu.rowptr; a.u_col = p->cu.col; a.u_erow = p->cu.erow; a.u_colptr = p->cu.colptr;
  a.u_crow = p->cu.crow; a.u_ceid = p->cu.ceid; a.u_egen = p->cu.egen; a.u_eval = p->cu.eval;
  a.w_rowptr = p->cw.rowptr; a.w_col = p->cw.col; a.w_erow = p->cw.erow; a.w_colptr = p->cw.colptr;
  a.w_crow = p->cw.crow; a.w_ceid = p->cw.ceid; a.w_egen = p->cw.egen; a.w_eval = p->cw.eval;
  a.coef = p->d_coef; a.dcoef = p->d_dcoef; a.qdeg = p->d_qdeg; a.mul = p->d_mul; a.add = p->d_add;
  a.starts = p->d_starts; a.targets = p->d_targets; a.ckpt = p->d_ckpt; a.tring = p->d_tring;
  a.sample_out = p->d_sample_out; a.gw_partial = p->d_gw_partial;
  k_chain_big<<<p->blocks, p->threads, big_smem_bytes(p, want_grad), st>>>(a);
  p->launches++;
  QCUDA(cudaGetLastError());
  return QOCVL_OK;
}
