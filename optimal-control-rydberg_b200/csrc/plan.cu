// plan.cu -- plan lifetime, generator sparsity extraction, constant uploads (host side of the C-ABI).
#include "common.cuh"
#include <algorithm>
#include <cmath>

thread_local std::string g_qocvl_cuda_error;

typedef std::complex<double> zc;

template <typename T>
static int dev_alloc(qocvl_plan* p, T** ptr, size_t count) {
  if (*ptr) { cudaFree(*ptr); *ptr = nullptr; }
  if (count == 0) count = 1;
  QCUDA(cudaMalloc((void**)ptr, count * sizeof(T)));
  p->bytes += count * sizeof(T);
  return QOCVL_OK;
}

template <typename T>
static int dev_upload(qocvl_plan* p, T** ptr, const std::vector<T>& host) {
  int rc = dev_alloc(p, ptr, host.size());
  if (rc) return rc;
  if (!host.empty())
    QCUDA(cudaMemcpy(*ptr, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
  return QOCVL_OK;
}

static void free_layout(BlockLayout& l) {
  cudaFree(l.row_col); cudaFree(l.col_row); cudaFree(l.col_src);
  cudaFree(l.rc_gen); cudaFree(l.rc_val); cudaFree(l.cc_gen); cudaFree(l.cc_val);
  l = BlockLayout();
}

// Build the lane-per-row sparse tables for one block family.
//   blocks: [J][n][n], np_ lanes per vector;  diag_slot: reserve slot 0 for the diagonal entry (u family).
int qocvl_build_lane_layout(qocvl_plan* p, const std::vector<zc>& blocks, int n, int np_, bool diag_slot,
                            BlockLayout* out) {
  const int J = p->d.n_gen;
  auto G = [&](int j, int a, int b) { return blocks[((size_t)j * n + a) * n + b]; };
  std::vector<char> pat((size_t)n * n, 0);
  int nnz = 0;
  for (int a = 0; a < n; ++a)
    for (int b = 0; b < n; ++b) {
      bool any = false;
      for (int j = 0; j < J; ++j) any = any || (std::abs(G(j, a, b)) > 0.0);
      pat[(size_t)a * n + b] = any;
      nnz += any;
    }
  // per-row and per-column entry lists
  std::vector<std::vector<int>> rows(n), cols(n);
  for (int a = 0; a < n; ++a)
    for (int b = 0; b < n; ++b) {
      if (!pat[(size_t)a * n + b]) continue;
      if (diag_slot && a == b) continue;
      rows[a].push_back(b);
      cols[b].push_back(a);
    }
  int w_row = 0, w_col = 0, kc = 1;
  for (int r = 0; r < n; ++r) {
    w_row = std::max(w_row, (int)rows[r].size());
    w_col = std::max(w_col, (int)cols[r].size());
  }
  for (int a = 0; a < n; ++a)
    for (int b = 0; b < n; ++b) {
      int c = 0;
      for (int j = 0; j < J; ++j) c += (std::abs(G(j, a, b)) > 0.0);
      kc = std::max(kc, c);
    }
  const int base = diag_slot ? 1 : 0;
  const int nrs = std::max(1, base + w_row), ncs = std::max(1, base + w_col);
  std::vector<int> row_col((size_t)nrs * np_), col_row((size_t)ncs * np_), col_src((size_t)ncs * np_, 0);
  std::vector<int> rc_gen((size_t)nrs * kc * np_, -1), cc_gen((size_t)ncs * kc * np_, -1);
  std::vector<cplx> rc_val((size_t)nrs * kc * np_, cmake(0, 0)), cc_val((size_t)ncs * kc * np_, cmake(0, 0));
  for (int s = 0; s < nrs; ++s) for (int r = 0; r < np_; ++r) row_col[(size_t)s * np_ + r] = r;
  for (int s = 0; s < ncs; ++s) for (int r = 0; r < np_; ++r) col_row[(size_t)s * np_ + r] = r;
  auto fill_contrib = [&](std::vector<int>& gen, std::vector<cplx>& val, int slot, int lane, int a, int b) {
    int c = 0;
    for (int j = 0; j < J; ++j) {
      zc v = G(j, a, b);
      if (std::abs(v) > 0.0) {
        size_t idx = ((size_t)slot * kc + c) * np_ + lane;
        gen[idx] = j;
        val[idx] = cmake(v.real(), v.imag());
        ++c;
      }
    }
  };
  for (int r = 0; r < n; ++r) {
    if (diag_slot) {
      fill_contrib(rc_gen, rc_val, 0, r, r, r);
      fill_contrib(cc_gen, cc_val, 0, r, r, r);
      col_src[r] = r;  // slot 0 * npad + r
    }
    for (size_t s = 0; s < rows[r].size(); ++s) {
      row_col[(base + s) * np_ + r] = rows[r][s];
      fill_contrib(rc_gen, rc_val, base + (int)s, r, r, rows[r][s]);
    }
    for (size_t s = 0; s < cols[r].size(); ++s) {
      const int a = cols[r][s];  // entry (a, r)
      col_row[(base + s) * np_ + r] = a;
      fill_contrib(cc_gen, cc_val, base + (int)s, r, a, r);
      // where does (a, r) live in row a's slot list?
      int pos = (int)(std::find(rows[a].begin(), rows[a].end(), r) - rows[a].begin());
      col_src[(base + s) * np_ + r] = (base + pos) * np_ + a;
    }
  }
  free_layout(*out);
  out->n_row_slots = nrs; out->n_col_slots = ncs; out->kc = kc; out->nnz = nnz;
  int rc;
  if ((rc = dev_upload(p, &out->row_col, row_col))) return rc;
  if ((rc = dev_upload(p, &out->col_row, col_row))) return rc;
  if ((rc = dev_upload(p, &out->col_src, col_src))) return rc;
  if ((rc = dev_upload(p, &out->rc_gen, rc_gen))) return rc;
  if ((rc = dev_upload(p, &out->rc_val, rc_val))) return rc;
  if ((rc = dev_upload(p, &out->cc_gen, cc_gen))) return rc;
  if ((rc = dev_upload(p, &out->cc_val, cc_val))) return rc;
  return QOCVL_OK;
}

extern "C" int qocvl_plan_create(qocvl_plan** out, const qocvl_desc* desc, int device) {
  if (!out || !desc) return QOCVL_ERR_ARG;
  const qocvl_desc& d = *desc;
  if (d.n < 1 || d.n_gen < 1 || d.n_gen > QOCVL_MAX_GEN || d.n_chan < 1 || d.n_chan > QOCVL_MAX_CHAN ||
      d.n_slices < 1 || d.n_basis_pts < 2 || d.n_gauss < 1 || d.n_vec < 1 || d.pad < 0 ||
      d.adiab_index < 0 || d.adiab_index >= d.n_vec || d.n_slices != d.n_basis_pts + 2 * d.pad ||
      !(d.dt > 0) || !(d.duration > 0) || !(d.fid_norm > 0))
    return QOCVL_ERR_ARG;
  if (d.model != QOCVL_MODEL_STATE_TRANSFER && d.model != QOCVL_MODEL_TWO_QUBIT && d.model != QOCVL_MODEL_CHAIN)
    return QOCVL_ERR_UNSUPPORTED;
  if ((d.model == QOCVL_MODEL_STATE_TRANSFER || d.model == QOCVL_MODEL_TWO_QUBIT) &&
      (d.n_gen != 11 || d.n_chan != 5))
    return QOCVL_ERR_ARG;
  if (d.model == QOCVL_MODEL_CHAIN && (d.n_gen != 4 || d.n_chan != 2 || d.pad != 0)) return QOCVL_ERR_ARG;
  if (d.n > 1024) return QOCVL_ERR_UNSUPPORTED;
  QCUDA(cudaSetDevice(device));
  qocvl_plan* p = new qocvl_plan();
  p->d = d;
  p->device = device;
  p->npad = d.n <= 8 ? 8 : (d.n <= 16 ? 16 : 0);   // 0: no lane-per-row tables, big-n kernel only
  const int M = d.n_slices, nt = d.n_basis_pts, C = d.n_chan, J = d.n_gen;
  int rc = 0;
  rc |= dev_alloc(p, &p->d_raw, (size_t)nt * C);
  rc |= dev_alloc(p, &p->d_reg, (size_t)nt * C);
  rc |= dev_alloc(p, &p->d_wave, (size_t)M * C);
  rc |= dev_alloc(p, &p->d_coef, (size_t)M * J);
  rc |= dev_alloc(p, &p->d_dcoef, (size_t)M * J * C);
  rc |= dev_alloc(p, &p->d_qdeg, (size_t)M);
  rc |= dev_alloc(p, &p->d_flag, 1);
  rc |= dev_alloc(p, &p->d_gwave, (size_t)M * C);
  rc |= dev_alloc(p, &p->d_greg, (size_t)nt * C);
  rc |= dev_alloc(p, &p->d_graw, (size_t)nt * C);
  rc |= dev_alloc(p, &p->d_out3, 4);
  if (rc) { qocvl_plan_destroy(p); return QOCVL_ERR_CUDA; }
  cudaMemset(p->d_flag, 0, sizeof(int));
  // default: one nominal sample
  rc = qocvl_set_noise(p, 1, nullptr, nullptr);
  if (rc) { qocvl_plan_destroy(p); return rc; }
  *out = p;
  return QOCVL_OK;
}

extern "C" int qocvl_plan_destroy(qocvl_plan* p) {
  if (!p) return QOCVL_OK;
  cudaSetDevice(p->device);
  free_layout(p->lu); free_layout(p->lw); p->la.release(); p->lwide.release(); p->lwide_a.release(); p->duo.release();
  for (CsrLayout* c : {&p->cu, &p->cw}) {
    cudaFree(c->rowptr); cudaFree(c->col); cudaFree(c->erow); cudaFree(c->colptr);
    cudaFree(c->crow); cudaFree(c->ceid); cudaFree(c->egen); cudaFree(c->eval);
  }
  if (p->d_tring) cudaFree(p->d_tring);
  for (void* q : {(void*)p->d_gblk_u, (void*)p->d_gblk_w, (void*)p->d_blk_work}) if (q) cudaFree(q);
  for (auto& row : p->graphs) for (auto& g : row) if (g.exec) cudaGraphExecDestroy(g.exec);
  if (p->cap_stream) cudaStreamDestroy(p->cap_stream);
  if (p->ev0) cudaEventDestroy(p->ev0);
  if (p->ev1) cudaEventDestroy(p->ev1);
  for (cudaEvent_t e : p->ev_stage) if (e) cudaEventDestroy(e);
  void* ptrs[] = {p->d_gdense, p->d_coef_const, p->d_gnorm, p->d_starts, p->d_targets, p->d_mul, p->d_add,
                  p->d_mulmax, p->d_addmax, p->d_taps, p->d_raw, p->d_reg, p->d_wave, p->d_coef, p->d_dcoef,
                  p->d_qdeg, p->d_flag, p->d_wide_shift, p->d_targets_rot, p->d_ckpt, p->d_sample_out, p->d_gw_partial, p->d_gwave, p->d_greg,
                  p->d_graw, p->d_out3, p->d_dense_a, p->d_dense_b, p->d_aug_starts, p->d_aug_targets,
                  p->d_aug_pair, p->d_aug_wq, p->d_ypart, p->d_tp_uc, p->d_tp_wc, p->d_tp_zc, p->d_tp_lc};
  for (void* q : ptrs) if (q) cudaFree(q);
  delete p;
  return QOCVL_OK;
}

extern "C" int qocvl_set_generators(qocvl_plan* p, const double* gens_u, const double* gens_w,
                                    const double* coef_const) {
  if (!p || !gens_u || !gens_w || !coef_const) return QOCVL_ERR_ARG;
  QCUDA(cudaSetDevice(p->device));
  const int n = p->d.n, J = p->d.n_gen, D = 2 * n;
  const size_t blk = (size_t)J * n * n;
  for (cplx** q : {&p->d_gblk_u, &p->d_gblk_w}) if (*q) { cudaFree(*q); *q = nullptr; }   // dense blocks: rebuilt on use
  p->h_gu.resize(blk); p->h_gw.resize(blk); p->h_coef_const.resize(J);
  for (size_t i = 0; i < blk; ++i) {
    p->h_gu[i] = zc(gens_u[2 * i], gens_u[2 * i + 1]);
    p->h_gw[i] = zc(gens_w[2 * i], gens_w[2 * i + 1]);
  }
  std::vector<cplx> cc(J);
  for (int j = 0; j < J; ++j) {
    p->h_coef_const[j] = zc(coef_const[2 * j], coef_const[2 * j + 1]);
    cc[j] = cmake(coef_const[2 * j], coef_const[2 * j + 1]);
  }
  int rc;
  if ((rc = qocvl_build_csr(p, p->h_gu, &p->cu))) return rc;
  if ((rc = qocvl_build_csr(p, p->h_gw, &p->cw))) return rc;
  if (p->npad > 0) {
    if ((rc = qocvl_build_lane_layout(p, p->h_gu, p->d.n, p->npad, true, &p->lu))) return rc;
    if ((rc = qocvl_build_lane_layout(p, p->h_gw, p->d.n, p->npad, false, &p->lw))) return rc;
  }
  // column sums of |G_u| + |G_w| per generator: |X|_1 <= dt * max_b sum_j |c_j| * gcol[j][b]
  // (entrywise triangle inequality; the 2n x 2n block matrix has column sums |A| and |B|+|A|)
  std::vector<double> gnorm((size_t)J * n, 0.0);
  for (int j = 0; j < J; ++j)
    for (int b = 0; b < n; ++b) {
      double s = 0;
      for (int a = 0; a < n; ++a)
        s += std::abs(p->h_gu[((size_t)j * n + a) * n + b]) + std::abs(p->h_gw[((size_t)j * n + a) * n + b]);
      gnorm[(size_t)j * n + b] = s;
    }
  p->h_gcol = gnorm;
  if ((rc = dev_upload(p, &p->d_gnorm, gnorm))) return rc;
  if ((rc = dev_upload(p, &p->d_coef_const, cc))) return rc;
  // dense 2n x 2n copies for exponentials()/propagate() (shared-memory dense path: 2n <= 32 only)
  if (n <= 16) {
    std::vector<cplx> dense((size_t)J * D * D, cmake(0, 0));
    for (int j = 0; j < J; ++j)
      for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b) {
          zc u = p->h_gu[((size_t)j * n + a) * n + b], w = p->h_gw[((size_t)j * n + a) * n + b];
          dense[((size_t)j * D + a) * D + b] = cmake(u.real(), u.imag());
          dense[((size_t)j * D + n + a) * D + n + b] = cmake(u.real(), u.imag());
          dense[((size_t)j * D + a) * D + n + b] = cmake(w.real(), w.imag());
        }
    if ((rc = dev_upload(p, &p->d_gdense, dense))) return rc;
  }
  p->have_gens = true;
  p->groups = 0;  // force re-configuration of the chain launch
  return QOCVL_OK;
}

extern "C" int qocvl_set_filter(qocvl_plan* p, const double* taps) {
  if (!p) return QOCVL_ERR_ARG;
  QCUDA(cudaSetDevice(p->device));
  if (!taps) {
    if (p->d.pad != 0) return QOCVL_ERR_ARG;
    p->have_filter = false;
    return QOCVL_OK;
  }
  std::vector<double> h(taps, taps + p->d.n_slices);
  int rc = dev_upload(p, &p->d_taps, h);
  if (rc) return rc;
  p->cfg_epoch++;
  p->have_filter = true;
  return QOCVL_OK;
}

extern "C" int qocvl_set_cost(qocvl_plan* p, const double* starts, const double* targets) {
  if (!p || !starts || !targets) return QOCVL_ERR_ARG;
  QCUDA(cudaSetDevice(p->device));
  const int n = p->d.n, V = p->d.n_vec, np_ = n;   // host copies are unpadded [n_vec][n]
  std::vector<cplx> s((size_t)V * np_, cmake(0, 0)), t((size_t)V * np_, cmake(0, 0));
  for (int v = 0; v < V; ++v)
    for (int r = 0; r < n; ++r) {
      s[(size_t)v * np_ + r] = cmake(starts[2 * ((size_t)v * n + r)], starts[2 * ((size_t)v * n + r) + 1]);
      t[(size_t)v * np_ + r] = cmake(targets[2 * ((size_t)v * n + r)], targets[2 * ((size_t)v * n + r) + 1]);
    }
  p->h_starts = s;     // the chain configuration compacts these to the kets that actually move
  p->h_targets = t;
  p->groups = 0;
  p->have_cost = true;
  return QOCVL_OK;
}

extern "C" int qocvl_set_noise(qocvl_plan* p, int n_samples, const double* mul, const double* add) {
  if (!p || n_samples < 1) return QOCVL_ERR_ARG;
  QCUDA(cudaSetDevice(p->device));
  const int J = p->d.n_gen;
  // mulmax = [ max_s |mul_sj| (J entries) , max_s |mul_sj - 1| (J entries) ]
  std::vector<double> hm((size_t)n_samples * J, 1.0), mulmax(2 * (size_t)J, 0.0), addmax(J, 0.0);
  std::vector<cplx> ha((size_t)n_samples * J, cmake(0, 0));
  for (size_t i = 0; i < hm.size(); ++i) {
    if (mul) hm[i] = mul[i];
    if (add) ha[i] = cmake(add[2 * i], add[2 * i + 1]);
    const int j = (int)(i % J);
    mulmax[j] = std::max(mulmax[j], std::fabs(hm[i]));
    mulmax[J + j] = std::max(mulmax[J + j], std::fabs(hm[i] - 1.0));
    addmax[j] = std::max(addmax[j], std::hypot(ha[i].x, ha[i].y));
  }
  int rc;
  p->h_mul = hm; p->h_add = ha;
  if ((rc = dev_upload(p, &p->d_mul, hm))) return rc;
  if ((rc = dev_upload(p, &p->d_add, ha))) return rc;
  if ((rc = dev_upload(p, &p->d_mulmax, mulmax))) return rc;
  if ((rc = dev_upload(p, &p->d_addmax, addmax))) return rc;
  if ((rc = dev_alloc(p, &p->d_sample_out, (size_t)n_samples * 3))) return rc;
  p->n_samples = n_samples;
  p->global_samples = n_samples;
  p->groups = 0;
  return QOCVL_OK;
}

extern "C" int qocvl_set_symmetry(qocvl_plan* p, const int* perm) {
  if (!p) return QOCVL_ERR_ARG;
  p->groups = 0;
  if (!perm) { p->h_perm.clear(); return QOCVL_OK; }
  const int n = p->d.n;
  std::vector<int> seen(n, 0);
  for (int i = 0; i < n; ++i) {
    if (perm[i] < 0 || perm[i] >= n || seen[perm[i]]++) return QOCVL_ERR_ARG;
  }
  p->h_perm.assign(perm, perm + n);
  return QOCVL_OK;
}

extern "C" int qocvl_set_shard(qocvl_plan* p, int global_samples) {
  if (!p || global_samples < p->n_samples) return QOCVL_ERR_ARG;
  p->global_samples = global_samples;
  p->cfg_epoch++;
  return QOCVL_OK;
}

// Explicit per-plan switches (tuning knobs and the cross-check paths the tests compare against each other).
// The library reads no environment variables: everything that changes a plan's behaviour is set on the plan.
extern "C" int qocvl_set_option(qocvl_plan* p, const char* name, int value) {
  static const char* known[] = {"aug", "store", "spb", "nofold", "force_big", "qcap", "tp", "tp_chunk", "tlocal",
                                "wide", "wide_store", "wide_spc", "wide_threads", "strong_chunks", "duo", "duo_wpc", "duo_nomirror", "duo_nolean", "duo_cols", "duo_roleq", "wide_norms", "wide_shift", "taylor_tol_exp", "graph", "debug_sync"};
  if (!p || !name) return QOCVL_ERR_ARG;
  bool ok = false;
  for (const char* k : known) ok = ok || strcmp(k, name) == 0;
  if (!ok) return QOCVL_ERR_ARG;
  if (value < 0) p->opts.erase(name);   // negative = back to automatic
  else p->opts[name] = value;
  p->groups = 0;                        // re-configure at the next evaluation
  return QOCVL_OK;
}

extern "C" int qocvl_set_precision(qocvl_plan* p, int dtype) {
  if (!p || (dtype != QOCVL_C128 && dtype != QOCVL_C64)) return QOCVL_ERR_ARG;
  if (p->c64 != dtype) { p->c64 = dtype; p->groups = 0; }
  return QOCVL_OK;
}
extern "C" int qocvl_precision(const qocvl_plan* p) { return p ? p->c64 : 0; }

extern "C" size_t qocvl_workspace_bytes(const qocvl_plan* p) { return p ? p->bytes : 0; }
extern "C" long long qocvl_launch_count(const qocvl_plan* p) { return p ? p->launches : 0; }
extern "C" int qocvl_last_taylor_degree(const qocvl_plan* p) { return p ? p->qcap : 0; }
extern "C" int qocvl_vectors_propagated(const qocvl_plan* p) { return (p && p->groups > 0) ? p->n_live + 1 : 0; }
extern "C" int qocvl_reduced_rows(const qocvl_plan* p) { return (p && p->groups > 0 && p->use_aug) ? p->aug_n : 0; }
extern "C" const char* qocvl_last_cuda_error(void) { return g_qocvl_cuda_error.c_str(); }
extern "C" int qocvl_version(void) { return 100; }

extern "C" const char* qocvl_strerror(int status) {
  switch (status) {
    case QOCVL_OK: return "ok";
    case QOCVL_ERR_ARG: return "invalid argument";
    case QOCVL_ERR_CUDA: return "CUDA runtime error";
    case QOCVL_ERR_STATE: return "plan not fully set up (generators / cost / filter missing)";
    case QOCVL_ERR_NORM: return "slice generator norm too large for the Taylor propagator (refine delta_t)";
    case QOCVL_ERR_UNSUPPORTED: return "configuration not supported by this build";
    default: return "unknown status";
  }
}
