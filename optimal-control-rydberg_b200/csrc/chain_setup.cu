// chain_setup.cu -- host logic shared by the two chain kernels (chain_small.cu: one row per lane, n <= 16;
// chain_big.cu: rows strided over a CTA, any n): which start kets have to be propagated, the a-priori
// Taylor cap, kernel dispatch, and the deterministic reductions that finish an evaluation.
#include "common.cuh"
#include "model.cuh"
#include <algorithm>

// out3 = ensemble sums of (cost, infid, adiab) / global sample count; deterministic order.
__global__ void k_reduce_samples(int S, double inv_samples, const double* __restrict__ sample_out,
                                 double* __restrict__ out3) {
  __shared__ double sh[3][256];
  const int tid = threadIdx.x;
  double a0 = 0, a1 = 0, a2 = 0;
  for (int s = tid; s < S; s += 256) {
    a0 += sample_out[(size_t)s * 3 + 0];
    a1 += sample_out[(size_t)s * 3 + 1];
    a2 += sample_out[(size_t)s * 3 + 2];
  }
  sh[0][tid] = a0; sh[1][tid] = a1; sh[2][tid] = a2;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (tid < off) { sh[0][tid] += sh[0][tid + off]; sh[1][tid] += sh[1][tid + off]; sh[2][tid] += sh[2][tid + off]; }
    __syncthreads();
  }
  if (tid < 3) out3[tid] = sh[tid][0] * inv_samples;
}

// gwave[k][c] = sum over warps / CTAs of their partial sums (fixed order: deterministic)
__global__ void k_reduce_gwave(int nparts, int MC, const double* __restrict__ partial,
                               double* __restrict__ gwave) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= MC) return;
  double acc = 0.0;
  for (int b = 0; b < nparts; ++b) acc += partial[(size_t)b * MC + i];
  gwave[i] = acc;
}

// Host-side a-priori bound on |X_k|_1: maximum of the exact slice norm over a grid of admissible
// physical amplitudes (the regularisation keeps every amplitude in [-1, 1], ST:204-227 / TQ:355-373,
// and the filter taps are >= 0 with sum < 1), with a 5 % margin for the grid spacing.  It only sizes
// the rings of Taylor terms: the device computes the per-slice degree from the actual amplitudes and
// poisons + flags a slice that would exceed this cap (never silently wrong).
static double apriori_norm_bound(const qocvl_plan* p, const std::vector<double>& mulmax,
                                 const std::vector<double>& addmax) {
  const qocvl_desc& d = p->d;
  ModelParams mp = {d.model, d.n, d.n_gen, d.n_chan, d.dt, d.eps, d.scale[0], d.scale[1], d.n_slices};
  const int J = d.n_gen;
  std::vector<cplx> cc(J);
  for (int j = 0; j < J; ++j) cc[j] = cmake(p->h_coef_const[j].real(), p->h_coef_const[j].imag());
  cplx co[QOCVL_MAX_GEN];
  cplx dc[QOCVL_MAX_GEN][QOCVL_MAX_CHAN];
  double worst = 0.0;
  auto probe = [&](int k, const double* w) {
    qocvl_model_coefficients(mp, k, w, cc.data(), co, dc);
    worst = std::max(worst, qocvl_slice_norm_csr(
        mp, co, p->cu.h_colptr.data(), p->cu.h_ceid.data(), p->cu.h_egen.data(), p->cu.h_eval.data(), p->cu.kc,
        p->cw.h_colptr.data(), p->cw.h_ceid.data(), p->cw.h_egen.data(), p->cw.h_eval.data(), p->cw.kc,
        p->h_gcol.data(), mulmax.data(), addmax.data()));
  };
  if (d.model == QOCVL_MODEL_CHAIN) {
    for (int k = 0; k < d.n_slices; k += std::max(1, d.n_slices - 1))
      for (int i0 = 0; i0 <= 2; ++i0)
        for (int i1 = -1; i1 <= 1; ++i1) {
          double w[QOCVL_MAX_CHAN] = {0.5 * i0, (double)i1, 0, 0, 0};
          probe(k, w);
        }
    return 1.05 * worst;
  }
  const int NA = 8;   // angles
  for (int i0 = 0; i0 < 2; ++i0)
    for (int r1 = 1; r1 <= 2; ++r1)
      for (int a1 = 0; a1 < NA; ++a1)
        for (int r3 = 1; r3 <= 2; ++r3)
          for (int a3 = 0; a3 < NA; ++a3) {
            double w[QOCVL_MAX_CHAN];
            if (d.model == QOCVL_MODEL_TWO_QUBIT) {   // (magnitude, phase) pairs
              w[0] = i0 ? -1.0 : 0.0;
              w[1] = 0.5 * r1; w[2] = -1.0 + 2.0 * a1 / NA;
              w[3] = 0.5 * r3; w[4] = -1.0 + 2.0 * a3 / NA;
            } else {                                  // (re, im) pairs inside the unit disk
              w[0] = i0 ? -1.0 : 1.0;
              w[1] = 0.5 * r1 * cos(2 * PI_D * a1 / NA); w[2] = 0.5 * r1 * sin(2 * PI_D * a1 / NA);
              w[3] = 0.5 * r3 * cos(2 * PI_D * a3 / NA); w[4] = 0.5 * r3 * sin(2 * PI_D * a3 / NA);
            }
            probe(0, w);
          }
  return 1.05 * worst;
}

// Decide which start kets are propagated, the Taylor cap, which kernel runs, and (re)allocate buffers.
int qocvl_chain_configure(qocvl_plan* p) {
  if (!p->have_gens || !p->have_cost) return QOCVL_ERR_STATE;
  if (p->groups > 0) return QOCVL_OK;
  p->cfg_epoch++;                                          // cached evaluation graphs (api.cu) are stale from here on
  const qocvl_desc& d = p->d;
  const int J = d.n_gen, n = d.n;
  p->use_big = (p->npad == 0);
  if (p->opt("force_big")) p->use_big = true;
  // the lane-per-row kernel holds <= 4 off-diagonals per row/column (<= 2 for the top-right block)
  if (!p->use_big && (p->lu.n_row_slots - 1 > 4 || p->lu.n_col_slots - 1 > 4 || p->lw.n_row_slots > 2 ||
                      p->lw.n_col_slots > 2 || d.n_vec > 4))
    p->use_big = true;
  // ---- start kets annihilated by every generator block never move (e.g. |00> of the CZ model):
  //      drop them from the propagation and keep their overlap as a constant.
  p->d_const = cmake(0, 0);
  p->n_live = 0;
  p->adiab_live = 0;
  std::vector<int> weight(d.n_vec, 1);          // how many start kets this one stands for
  std::vector<char> moves(d.n_vec, 0);
  auto ket = [&](const std::vector<cplx>& kets, int v, int i) { return kets[(size_t)v * n + i]; };
  for (int v = 0; v < d.n_vec; ++v) {
    bool mv = (v == d.adiab_index);
    for (int j = 0; j < J && !mv; ++j)
      for (int a = 0; a < n && !mv; ++a) {
        std::complex<double> acc(0, 0);
        for (int b = 0; b < n; ++b)
          acc += p->h_gu[((size_t)j * n + a) * n + b] *
                 std::complex<double>(ket(p->h_starts, v, b).x, ket(p->h_starts, v, b).y);
        mv = std::abs(acc) > 0.0;
      }
    moves[v] = mv;
  }
  // ---- symmetry hint: verify G^u invariance, then merge start kets that are images of each other
  if (!p->h_perm.empty()) {
    const std::vector<int>& pm = p->h_perm;
    for (int j = 0; j < J; ++j)
      for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b)
          if (p->h_gu[((size_t)j * n + pm[a]) * n + pm[b]] != p->h_gu[((size_t)j * n + a) * n + b])
            return QOCVL_ERR_ARG;
    auto is_image = [&](const std::vector<cplx>& kets, int v, int w) {   // kets[w] == P kets[v] ?
      for (int i = 0; i < n; ++i) {
        const cplx x = ket(kets, v, i), y = ket(kets, w, pm[i]);
        if (x.x != y.x || x.y != y.y) return false;
      }
      return true;
    };
    std::vector<int> order;                     // representatives: the adiabatic ket first
    order.push_back(d.adiab_index);
    for (int v = 0; v < d.n_vec; ++v) if (v != d.adiab_index) order.push_back(v);
    for (size_t i = 0; i < order.size(); ++i) {
      const int v = order[i];
      if (!moves[v] || weight[v] == 0) continue;
      for (size_t k2 = i + 1; k2 < order.size(); ++k2) {
        const int w = order[k2];
        if (!moves[w] || weight[w] == 0 || w == d.adiab_index) continue;
        if (is_image(p->h_starts, v, w) && is_image(p->h_targets, v, w)) {
          weight[v] += weight[w];
          weight[w] = 0;
        }
      }
    }
  }
  const int stride = p->use_big ? n : p->npad;  // device layout of the live kets: [n_live][stride]
  std::vector<cplx> live_s, live_t;
  p->h_live_s.clear(); p->h_live_t.clear();
  for (int v = 0; v < d.n_vec; ++v) {
    if (moves[v] && weight[v] > 0) {
      if (v == d.adiab_index) p->adiab_live = p->n_live;
      for (int r = 0; r < n; ++r) {          // unpadded host copies for the reduced-system builder (chain_aug.cu)
        p->h_live_s.push_back(ket(p->h_starts, v, r));
        p->h_live_t.push_back(cscale(ket(p->h_targets, v, r), (double)weight[v]));
      }
      for (int r = 0; r < stride; ++r) {
        live_s.push_back(r < n ? ket(p->h_starts, v, r) : cmake(0, 0));
        // the weight rides on the target ket: overlap and terminal cotangent are both linear in it
        live_t.push_back(r < n ? cscale(ket(p->h_targets, v, r), (double)weight[v]) : cmake(0, 0));
      }
      p->n_live++;
    } else if (!moves[v]) {
      for (int r = 0; r < n; ++r) {   // y_v^dag x_v
        const cplx y = ket(p->h_targets, v, r), x = ket(p->h_starts, v, r);
        p->d_const.x += y.x * x.x + y.y * x.y;
        p->d_const.y += y.x * x.y - y.y * x.x;
      }
    }
  }
  if (p->d_starts) { cudaFree(p->d_starts); p->d_starts = nullptr; }
  if (p->d_targets) { cudaFree(p->d_targets); p->d_targets = nullptr; }
  QCUDA(cudaMalloc((void**)&p->d_starts, live_s.size() * sizeof(cplx)));
  QCUDA(cudaMalloc((void**)&p->d_targets, live_t.size() * sizeof(cplx)));
  QCUDA(cudaMemcpy(p->d_starts, live_s.data(), live_s.size() * sizeof(cplx), cudaMemcpyHostToDevice));
  QCUDA(cudaMemcpy(p->d_targets, live_t.data(), live_t.size() * sizeof(cplx), cudaMemcpyHostToDevice));
  // ---- a-priori Taylor cap
  std::vector<double> mulmax(2 * (size_t)J), addmax(J);
  QCUDA(cudaMemcpy(mulmax.data(), p->d_mulmax, 2 * J * sizeof(double), cudaMemcpyDeviceToHost));
  QCUDA(cudaMemcpy(addmax.data(), p->d_addmax, J * sizeof(double), cudaMemcpyDeviceToHost));
  const double bound = apriori_norm_bound(p, mulmax, addmax);
  if (!(bound == bound)) return QOCVL_ERR_ARG;
  p->qcap = qocvl_taylor_degree(bound, p->taylor_tol());
  if (p->has_opt("qcap")) p->qcap = std::max(2, std::min(QOCVL_QMAX, p->opt("qcap")));
  // ---- kernel-specific geometry and buffers
  // a single sample has no ensemble parallelism: cut the slice axis into concurrent chunks instead
  p->use_tp = !p->use_big && p->n_samples == 1 && d.n_slices >= 32 && !p->c64;
  if (p->has_opt("tp")) p->use_tp = p->use_tp && p->opt("tp") != 0;
  // ... unless the two-lane kernels take the system: their time-parallel form (chain_duo.cu, chunks of ~16 slices) is
  // 15-25 % faster on the reference's notebook problems.  chain_small.cu's path stays for systems they do not take and
  // whenever it is asked for (tp = 1 / tp_chunk).
  const bool tp_asked = p->has_opt("tp_chunk") || (p->has_opt("tp") && p->opt("tp") == 1);
  const bool single_on_duo = p->use_tp && !tp_asked && p->opt("duo", 1) != 0 && p->opt("aug", 1) != 0 && !p->opt("force_big");
  // ensembles: the reduced stacked system (chain_aug.cu) whenever it fits its kernel
  p->use_aug = false;
  int rc = QOCVL_ERR_UNSUPPORTED;
  bool try_aug = (!p->use_tp || single_on_duo) && !p->opt("force_big");
  if (p->has_opt("aug")) try_aug = try_aug && p->opt("aug") != 0;
  if (try_aug) {
    rc = qocvl_aug_configure(p);
    if (rc != QOCVL_OK && rc != QOCVL_ERR_UNSUPPORTED) { p->groups = 0; return rc; }
    if (single_on_duo) {
      if (p->use_aug && p->use_duo) p->use_tp = false;
      else p->use_aug = false;                    // (the lane-group kernel is no match for one sample)
    }
  }
  // complex64 exists on the reduced stacked system and on the samples-minor kernel of the large Hilbert spaces
  // (chain_wide.cu): anything else fails loudly instead of running in fp64
  if (p->c64 && !p->use_aug && !p->use_big) { p->groups = 0; return QOCVL_ERR_UNSUPPORTED; }
  p->use_wide = false;
  if (!p->use_aug && p->use_big) {          // large Hilbert spaces: the samples-minor kernel when it takes the problem
    bool try_wide = true;
    if (p->has_opt("wide")) try_wide = p->opt("wide") != 0;
    if (try_wide) {
      rc = qocvl_wide_configure(p);
      if (rc != QOCVL_OK && rc != QOCVL_ERR_UNSUPPORTED) { p->groups = 0; return rc; }
    }
  }
  if (p->c64 && !p->use_aug && !p->use_wide) { p->groups = 0; return QOCVL_ERR_UNSUPPORTED; }
  if (!p->use_aug && !p->use_wide) {
    rc = p->use_big ? qocvl_big_configure(p) : qocvl_small_configure(p);
    if (rc) { p->groups = 0; return rc; }
  }
  if (p->use_tp && (rc = qocvl_tp_configure(p))) { p->groups = 0; return rc; }
  if (p->d_gw_partial) { cudaFree(p->d_gw_partial); p->d_gw_partial = nullptr; p->bytes -= p->gw_partial_elems * sizeof(double); }
  p->gw_partial_elems = std::max<size_t>(1, p->n_parts * d.n_slices * d.n_chan);
  QCUDA(cudaMalloc((void**)&p->d_gw_partial, p->gw_partial_elems * sizeof(double)));
  p->bytes += p->gw_partial_elems * sizeof(double);
  return QOCVL_OK;
}

int qocvl_launch_chain(qocvl_plan* p, bool want_grad, double* out3, double* gwave, cudaStream_t st) {
  int rc = qocvl_chain_configure(p);
  if (rc) return rc;
  const qocvl_desc& d = p->d;
  // event timing is skipped while the stream is being captured into a CUDA graph (the recorded events would not be
  // valid for cudaEventElapsedTime)
  bool timing = p->timing;
  if (timing) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) timing = false;
  }
  p->timing_now = timing;
  p->stage_timed[0] = p->stage_timed[1] = false;
  if (timing) QCUDA(cudaEventRecord(p->ev0, st));
  if (p->use_tp) rc = qocvl_tp_launch(p, want_grad, gwave, st);
  else if (p->use_aug) rc = qocvl_aug_launch(p, want_grad, gwave, st);
  else if (p->use_wide) rc = qocvl_wide_launch(p, want_grad, gwave, st);
  else rc = p->use_big ? qocvl_big_launch(p, want_grad, st) : qocvl_small_launch(p, want_grad, st);
  if (rc) return rc;
  if (timing) { QCUDA(cudaEventRecord(p->ev1, st)); p->timed_once = true; }
  k_reduce_samples<<<1, 256, 0, st>>>(p->n_samples, 1.0 / (double)p->global_samples, p->d_sample_out, out3);
  p->launches++;
  if (want_grad && !p->use_tp && !p->use_aug && !p->use_wide_tp) {
    const int MC = d.n_slices * d.n_chan;
    k_reduce_gwave<<<(MC + 255) / 256, 256, 0, st>>>((int)p->n_parts, MC, p->d_gw_partial, gwave);
    p->launches++;
  }
  QCUDA(cudaGetLastError());
  return QOCVL_OK;
}
