// common.cuh -- shared definitions for the qocvl sm_100a kernels (plan layout, complex helpers).
#pragma once
#include <cuda_runtime.h>
#include <cmath>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <complex>
#include <string>
#include <map>

#include "../../include/qocvl.h"

typedef double2 cplx;  // (re, im)

#define QOCVL_MAX_GEN 16
#define QOCVL_MAX_CHAN 5
#define QOCVL_QMAX 40        // hard cap of the Taylor degree
#define QOCVL_NORM_MAX 4.0   // |X_k|_1 above this -> QOCVL_ERR_NORM

__host__ __device__ __forceinline__ cplx cmake(double re, double im) { cplx r; r.x = re; r.y = im; return r; }
__host__ __device__ __forceinline__ cplx cadd(cplx a, cplx b) { return cmake(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cplx csub(cplx a, cplx b) { return cmake(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cplx cmul(cplx a, cplx b) {
  return cmake(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ cplx cconj(cplx a) { return cmake(a.x, -a.y); }
__host__ __device__ __forceinline__ cplx cscale(cplx a, double s) { return cmake(a.x * s, a.y * s); }
// acc + a*b
__device__ __forceinline__ cplx cfma(cplx a, cplx b, cplx acc) {
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
  return acc;
}
// acc + conj(a)*b
__device__ __forceinline__ cplx cfma_conj(cplx a, cplx b, cplx acc) {
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(-a.y, b.x, acc.y);
  return acc;
}

// complex64 twins of the helpers above (the optional single-precision mode of chain_aug.cu)
typedef float2 cplxf;
__host__ __device__ __forceinline__ cplxf cmakef(float re, float im) { cplxf r; r.x = re; r.y = im; return r; }
__host__ __device__ __forceinline__ cplxf cadd(cplxf a, cplxf b) { return cmakef(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cplxf cmul(cplxf a, cplxf b) {
  return cmakef(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ cplxf cscale(cplxf a, float s) { return cmakef(a.x * s, a.y * s); }
__host__ __device__ __forceinline__ cplxf cconj(cplxf a) { return cmakef(a.x, -a.y); }
__device__ __forceinline__ cplxf cfma(cplxf a, cplxf b, cplxf acc) {
  acc.x = fmaf(a.x, b.x, acc.x); acc.x = fmaf(-a.y, b.y, acc.x);
  acc.y = fmaf(a.x, b.y, acc.y); acc.y = fmaf(a.y, b.x, acc.y);
  return acc;
}
__device__ __forceinline__ cplxf cfma_conj(cplxf a, cplxf b, cplxf acc) {
  acc.x = fmaf(a.x, b.x, acc.x); acc.x = fmaf(a.y, b.y, acc.x);
  acc.y = fmaf(a.x, b.y, acc.y); acc.y = fmaf(-a.y, b.x, acc.y);
  return acc;
}

// Smallest q with  nrm^(q+1)/(q+1)! <= tol.  complex128: tol = 2^-55 = u/4 (u = fp64 unit round-off: a smaller
// remainder is not representable next to the O(1) state; the oracle's twin rule,
// oracle/vector_model.py::taylor_degree, uses the stricter 2^-60).  complex64: 2^-26 = u_fp32/4
// (measured with the tight per-role bounds: the actual remainder is ~1e-11 per slice, profiles/c64_shift_probe.py).
#define QOCVL_TAYLOR_TOL_C128 2.7755575615628914e-17  // 2^-55
#define QOCVL_TAYLOR_TOL_C64 1.4901161193847656e-08   // 2^-26
__host__ __device__ inline int qocvl_taylor_degree(double nrm, double tol = QOCVL_TAYLOR_TOL_C128) {
  double term = 1.0;
  int q = 0;
  while (q < QOCVL_QMAX) {
    ++q;
    term *= nrm / q;
    if (q >= 2 && term * nrm / (q + 1) <= tol) return q;
  }
  return QOCVL_QMAX;
}

extern thread_local std::string g_qocvl_cuda_error;

#define QCUDA(call)                                                                      \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      g_qocvl_cuda_error = std::string(#call) + ": " + cudaGetErrorString(e_);           \
      return QOCVL_ERR_CUDA;                                                             \
    }                                                                                    \
  } while (0)

// Sparse layout of one generator block family (u = diagonal blocks, w = top-right blocks),
// rows distributed one per lane.  "row-owner" slots feed y = A x, "col-owner" slots feed
// y = A^dag x and the gradient contraction.  Slot 0 of the u family is the diagonal.
struct BlockLayout {
  int n_row_slots = 0;  // per-row slots (u: 1 + off-diagonals; w: all entries)
  int n_col_slots = 0;
  int kc = 0;           // max generator contributions per slot
  int nnz = 0;          // true non-zeros of the union pattern
  // device tables, all [slot][npad] (contribution tables [slot][kc][npad])
  int* row_col = nullptr;   // column read by row-owner slot (own row when empty)
  int* col_row = nullptr;   // row read by col-owner slot
  int* col_src = nullptr;   // index (slot*npad + row) of that entry in the row-owner value table
  int* rc_gen = nullptr;    // generator id or -1
  cplx* rc_val = nullptr;
  int* cc_gen = nullptr;
  cplx* cc_val = nullptr;
};

// CSR (row-owner) and CSC (column-owner) view of the union sparsity pattern of one block family,
// with up to kc generator contributions per entry:  value_e = dt * sum_c coef[gen[e][c]] * val[e][c].
struct CsrLayout {
  int n = 0, nnz = 0, kc = 1;
  // device
  int* rowptr = nullptr;   // [n+1]
  int* col = nullptr;      // [nnz]   column of entry e
  int* erow = nullptr;     // [nnz]   row of entry e
  int* colptr = nullptr;   // [n+1]
  int* crow = nullptr;     // [nnz]   row of the i-th entry in column order
  int* ceid = nullptr;     // [nnz]   its entry id e (row order)
  int* egen = nullptr;     // [nnz][kc] generator id or -1
  cplx* eval = nullptr;    // [nnz][kc]
  // host copies (a-priori norm bound)
  std::vector<int> h_colptr, h_ceid, h_egen;
  std::vector<cplx> h_eval;
};

// Device tables of the reduced stacked system (chain_aug.cu).  Generators are grouped into "classes" by their
// per-sample multiplier column (class 0: multiplier 1 for every sample), so that a matrix entry of sample s is
//   X_e(k, s) = A_e(s) + sum_i m_s[cls_i] * P_i(k),   P_i(k) = dt * sum_{g in class i of e} coef_kg * val_eg
// with P sample-independent (one small table per slice, k_aug_tables) and A the additive-noise part.
struct AugLayout {
  int N = 0, L = 0, w = 0, wc = 0, KC = 1, KG = 1, KA = 0, ncls = 1, nnz = 0;
  unsigned add_row_mask = 0, add_col_mask = 0;   // slots that carry an additive per-sample term
  int *row_col = nullptr, *col_row = nullptr;    // [1+w][L], [1+wc][L]  (slot 0 = diagonal)
  int *r_cls = nullptr, *c_cls = nullptr;        // [(1+w)*KC][L], [(1+wc)*KC][L]  class id or -1
  int *r_gen = nullptr, *c_gen = nullptr;        // [(1+w)*KC][KG][L] generator ids of the contribution (-1 = none)
  cplx *r_val = nullptr, *c_val = nullptr;
  int *ra_gen = nullptr, *ca_gen = nullptr;      // [1+w][KA][L] generators with an additive per-sample coefficient
  cplx *ra_val = nullptr, *ca_val = nullptr;
  double* mcls = nullptr;                        // [S][ncls]
  cplx *p_row = nullptr, *p_col = nullptr;       // [M][(1+w)*KC][L], [M][(1+wc)*KC][L]
  void release() {
    void* ptrs[] = {row_col, col_row, r_cls, c_cls, r_gen, c_gen, r_val, c_val, ra_gen, ca_gen, ra_val, ca_val,
                    mcls, p_row, p_col};
    for (void* q : ptrs) if (q) cudaFree(q);
    *this = AugLayout();
  }
};

// Device tables of the wide ensemble kernel (chain_wide.cu): rows sorted by off-diagonal work and dealt to the warps
// in groups; per (generator block, group) an interleaved, padded entry block for the row lists (forward sweep) and
// the column lists (reverse sweep); the diagonal of every generator block as a dense vector.
struct WideLayout {
  int spc = 0, items = 0, max_threads = 0, ngroups = 0, nb = 0, nd = 0;
  int bgen[8] = {0}, bfam[8] = {0}, dgen[8] = {0}, dfam[8] = {0}, buni[8] = {0};
  cplx bval[8] = {};
  int ent_fwd = 0, ent_rev = 0;
  bool store = false, dreal = false, simple = false, fwd = false;
  size_t smem = 0;
  int threads = 0;
  double macs = 0.0;                         // useful complex multiply-adds of one Taylor step of one sample
  long long padded_entries = 0, true_entries = 0;
  int* grow = nullptr;
  int2 *fwd_co = nullptr, *rev_co = nullptr;
  cplx *fwd_val = nullptr, *rev_val = nullptr, *dval = nullptr;
  double* dval_re = nullptr;
  unsigned char* dmask = nullptr;
  unsigned short *fwd_col = nullptr, *rev_col = nullptr;
  void release();
};

// Host state of the two-lanes-per-sector ensemble kernels (chain_duo.cu).
struct DuoState {
  int nroles = 0, wpc = 0, KG = 1, ntot = 0, warps_per_role = 0, numax = 0;
  int chunks = 1, lc = 0, G = 0, ncol[4] = {0}, mirror[4] = {0};   // time-parallel evaluation of small ensembles (slice-axis chunks)
  int variant[4] = {0}, nu[4] = {0}, nx[4] = {0}, ent0[4] = {0};
  // "lean" twin of the reverse sweep: no Xbar on own-block entries that only a dead pulse channel would consume
  bool has_lean = false, last_lean = false;
  int nnz_x_lean = 0;        // entries of the stacked system whose Xbar the lean reverse sweep accumulates
  int variant_lean[4] = {0}, nx_lean[4] = {0};
  int *d_ent_x_lean = nullptr, *d_ent_ystride_lean = nullptr;
  unsigned long long* d_ent_ysrc_lean = nullptr;
  size_t tab_off[4] = {0}, con_off[4] = {0}, y_off[4] = {0}, warp0[4] = {0};
  size_t smem_fwd = 0, smem_rev = 0;
  cplx *d_tab = nullptr, *d_con_a = nullptr, *d_ent_val = nullptr, *d_starts = nullptr, *d_targets = nullptr,
       *d_zfin = nullptr, *d_lam = nullptr, *d_prop = nullptr, *d_zstart = nullptr, *d_lamend = nullptr;
  double* d_wq = nullptr;
  unsigned clsmask[8] = {0}, amask[8] = {0};   // per (role, lane type): entries of class 1 / with an additive-noise part
  int *d_srow = nullptr, *d_ent_gen = nullptr, *d_ent_x = nullptr, *d_ent_stride = nullptr, *d_ent_ystride = nullptr,
      *d_ent_ywarps = nullptr, *d_pair = nullptr, *d_colrow = nullptr;
  unsigned long long *d_ent_out = nullptr, *d_ent_ysrc = nullptr;
  // per-role Taylor degrees: colrole[b] = bit mask of the roles whose sectors contain basis state b; k_coefficients
  // writes qdeg_role[role][M] from the column sums of those states ("duo_roleq" = 0: every role runs the global degree)
  unsigned char* d_colrole = nullptr;
  int* d_qdeg_role = nullptr;
  bool role_q = false;
  int nnz_role[4] = {0}, nnz_x_lean_role[4] = {0};   // non-zeros / lean Xbar entries of every role (flop model)
  // per-role phase shift ("duo_roleq" = 3, the default): k_slice_degrees writes phi[role][M], k_duo_tables subtracts
  // i phi from the role's diagonal entries, k_duo_phase rotates the role's target kets by the accumulated phase
  bool shift_on = false;
  double* d_phi = nullptr;
  int *d_rowrole = nullptr, *d_ent_srole = nullptr;   // [N] role of a stacked row; [ntot] role if the descriptor is a diagonal entry, else -1
  cplx* d_targets_rot = nullptr;
  void release();
};

struct qocvl_plan {
  qocvl_desc d;
  // qocvl_set_option: explicit per-plan tuning / cross-check switches (the library reads no environment variables)
  std::map<std::string, int> opts;
  bool has_opt(const char* name) const { return opts.count(name) != 0; }
  int opt(const char* name, int dflt = 0) const { auto it = opts.find(name); return it == opts.end() ? dflt : it->second; }
  int device = 0;
  int npad = 0;            // lanes per vector (8, 16 or 32)
  int n_samples = 1;
  int global_samples = 1;
  bool have_gens = false, have_cost = false, have_filter = false;
  long long launches = 0;
  size_t bytes = 0;
  int qcap = 0;            // a-priori Taylor-degree cap used to size shared memory
  int qcap_apriori = 0;    // the same before the thread-local ring lifts it to QOCVL_QMAX
  // generators
  std::vector<std::complex<double>> h_gu, h_gw, h_coef_const;  // host copies
  std::vector<double> h_gcol;   // [J][n] column sums of |G_u| + |G_w| (entrywise-abs norm bound)
  BlockLayout lu, lw;      // lane-per-row tables (n <= 16 only)
  CsrLayout cu, cw;        // CSR / CSC tables (any n): big-n kernel, norm bounds
  // samples-minor kernel (chain_wide.cu): per-slice phase shift phi_k of the generator (X_k - i phi_k I is what the
  // Taylor series sees; the accumulated phase goes into the rotated target kets) and its degree kernel's shared memory
  double* d_wide_shift = nullptr;      // [M]
  cplx* d_targets_rot = nullptr;       // [n_live][n]
  size_t wide_norm_smem = 0;           // 0: k_wide_degrees not available for this system (serial norms of k_coefficients)
  bool wide_norms = false, wide_norms_fresh = false;   // fresh: k_coefficients left norms / degrees / poison rule to k_wide_degrees
  cplx* d_gdense = nullptr;   // [J][2n][2n] dense generators (for exponentials())
  cplx* d_coef_const = nullptr;  // [J]
  double* d_gnorm = nullptr;  // [J][n] column sums of |G_u| + |G_w|
  // cost: host copies of all kets [n_vec][npad]; device copies only of the "live" kets (start kets
  // annihilated by every generator never move: their overlap is the constant d_const)
  std::vector<cplx> h_starts, h_targets;
  std::vector<int> h_perm;    // optional symmetry hint (qocvl_set_symmetry)
  cplx* d_starts = nullptr;   // [n_live][npad]
  cplx* d_targets = nullptr;
  int n_live = 0, adiab_live = 0;
  cplx d_const = {0.0, 0.0};
  // noise
  double* d_mul = nullptr;    // [S][J]
  cplx* d_add = nullptr;      // [S][J]
  double* d_mulmax = nullptr; // [J]
  double* d_addmax = nullptr; // [J]
  // filter
  double* d_taps = nullptr;   // [M]
  // per-call work buffers
  double* d_raw = nullptr;    // [nt][C]
  double* d_reg = nullptr;    // [nt][C]
  double* d_wave = nullptr;   // [M][C]
  cplx* d_coef = nullptr;     // [M][J]
  cplx* d_dcoef = nullptr;    // [M][J][C]
  int* d_qdeg = nullptr;      // [M]
  int* d_flag = nullptr;      // [1] sticky error flag
  cplx* d_ckpt = nullptr;     // checkpoints
  size_t ckpt_bytes = 0;
  double* d_sample_out = nullptr;  // [S][3]
  double* d_gw_partial = nullptr;  // [blocks][M][C]
  size_t gw_partial_elems = 0;
  double* d_gwave = nullptr;  // [M][C]
  double* d_greg = nullptr;   // [nt][C]
  double* d_graw = nullptr;   // [nt][C]
  double* d_out3 = nullptr;   // [3]
  cplx *d_gblk_u = nullptr, *d_gblk_w = nullptr;   // [J][n][n] dense generator blocks (dense_blocks.cu, n > 16)
  cplx* d_blk_work = nullptr;                      // its per-batch workspace
  cplx* d_dense_a = nullptr;  // [M][2n][2n] ping
  cplx* d_dense_b = nullptr;  // [M/2+1][2n][2n] pong
  // cached CUDA graphs of whole evaluations (api.cu): a caller that repeats qocvl_cost / qocvl_cost_grad with the same
  // buffers -- an optimiser loop -- gets the ~15 launches of an evaluation replayed as one graph.  cfg_epoch counts the
  // (re)configurations and every setter that changes what a launch would do; a graph is valid for one epoch only.
  struct GraphSlot {
    cudaGraphExec_t exec = nullptr;
    int kind = 0;
    const void* in = nullptr;
    void* out = nullptr;
    long long epoch = -1, launches = 0;
  };
  GraphSlot graphs[2][2];                 // [cost / cost + gradient][two buffer pairs: a caller whose output tensors ping-pong]
  const void* recent_in[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // buffer pairs of the last two calls of each kind
  void* recent_out[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
  int recent_pos[2] = {0, 0}, graph_victim[2] = {0, 0};
  cudaStream_t cap_stream = nullptr;
  long long cfg_epoch = 0, evals_in_epoch = 0, evals_epoch = -1;
  bool graph_failed = false;
  // optional kernel timing
  bool timing = false, timed_once = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_stage[4] = {nullptr, nullptr, nullptr, nullptr};   // forward sweep start / end, reverse sweep start / end
  bool timing_now = false;   // this launch is being timed (not inside a graph capture)
  bool lean_now = false;     // this evaluation feeds the pulse-map VJP: gradients of dead channels are not needed
  bool stage_timed[2] = {false, false};
  // chain launch geometry
  int spb = 1, groups = 0, threads = 0, blocks = 0;
  bool tlocal = true;      // Taylor terms of the reverse sweep in thread-local (L1) instead of shared memory
  bool use_big = false;    // rows-strided-over-a-CTA kernel (n > 16, or forced with QOCVL_FORCE_BIG=1)
  // time-parallel single-sample path (chain_small.cu): chunk propagators -> combine -> per-chunk sweeps
  bool use_tp = false;
  int tp_chunk = 0, tp_nchunks = 0, tp_ncol_pad = 0, tp_threads = 0, tp_blocks_a = 0, tp_blocks_c = 0;
  size_t tp_smem_a = 0, tp_smem_c = 0;
  cplx *d_tp_uc = nullptr, *d_tp_wc = nullptr, *d_tp_zc = nullptr, *d_tp_lc = nullptr;
  size_t n_parts = 0;      // number of per-warp / per-CTA partial gradients in d_gw_partial
  cplx* d_tring = nullptr; // big kernel: Taylor-term ring [S][qcap][nvp][n] in global memory (L2-resident)
  size_t tring_bytes = 0;
  size_t chain_smem = 0;
  // reduced ensemble path (chain_aug.cu): every propagated vector restricted to the basis states it can reach,
  // symmetric sectors folded onto orbit representatives, all of it stacked into ONE sparse system of aug_n rows
  bool use_wide = false;   // big-n ensembles: samples-minor CTA kernel (chain_wide.cu)
  WideLayout lwide;
  WideLayout lwide_a;      // 8 basis columns per CTA: column sweeps of the time-parallel single-sample path
  bool use_wide_tp = false;
  bool use_aug = false;
  bool use_duo = false;    // ... on the two-lanes-per-sector kernels (chain_duo.cu) instead of the lane-group kernel
  DuoState duo;
  std::vector<std::complex<double>> h_aug_gen;   // [J][aug_n][aug_n] stacked generators (host copy for chain_duo.cu)
  std::vector<int> h_aug_cls;                    // [J] multiplier class of every generator
  std::vector<char> h_aug_hasadd;                // [J] generator carries an additive per-sample coefficient
  std::vector<double> h_aug_mcls, h_aug_wq;      // [S][ncls] class multipliers, [2][L] adiabaticity weights
  std::vector<cplx> h_aug_st, h_aug_tg;          // [L] stacked start / target kets
  std::vector<int> h_aug_pair;   // host copy of the partner rows of the stacked vector (chain_duo.cu)
  int aug_p0 = 0;                                // first stacked row that belongs to p = W psi (rows before: U x_v)
  std::vector<std::vector<int>> h_aug_members;   // [aug_n] basis states a stacked row stands for (its orbit)
  int c64 = 0;             // qocvl_set_precision: 1 = the ensemble kernel runs in complex64 (reduced stacked system only)
  // ("taylor_tol_exp" = e: remainder tolerance 2^-e instead of the precision's default)
  double taylor_tol() const {
    if (has_opt("taylor_tol_exp")) return ldexp(1.0, -std::max(8, std::min(70, opt("taylor_tol_exp"))));
    return c64 ? QOCVL_TAYLOR_TOL_C64 : QOCVL_TAYLOR_TOL_C128;
  }
  int aug_n = 0, aug_lpg = 0, aug_nnz = 0, aug_wmax = 0, aug_parts = 0, aug_qs = 0;
  bool aug_store = false;         // Taylor terms streamed through HBM instead of recomputed in the reverse sweep
  cplx* d_ypart = nullptr;        // [warps][M][column contributions][aug_lpg] per-warp sums of m * Xbar (c64: float2)
  size_t ypart_bytes = 0;
  AugLayout la;            // lane-per-row tables of the stacked system
  std::vector<double> h_mul;              // [S][J] host copy of the noise tables (generator classes)
  std::vector<cplx> h_add;
  std::vector<cplx> h_live_s, h_live_t;   // [n_live][n] live start / (weighted) target kets, unpadded
  cplx *d_aug_starts = nullptr, *d_aug_targets = nullptr;   // [aug_lpg]
  int* d_aug_pair = nullptr;      // [aug_lpg] row of the partner (q row <-> p row) in the adiabaticity term, or -1
  double* d_aug_wq = nullptr;     // [aug_lpg] orbit size of the row (weight of its term in q^dag p)
};

// pulse.cu
int qocvl_launch_pulse_forward(qocvl_plan* p, const double* abc, cudaStream_t st);
int qocvl_launch_coefficients(qocvl_plan* p, const double* wave, cudaStream_t st);
int qocvl_launch_pulse_vjp(qocvl_plan* p, const double* abc, const double* gwave, double* out_grads,
                           cudaStream_t st);
int qocvl_launch_aux(qocvl_plan* p, double* aux, cudaStream_t st);
// chain_setup.cu (common host logic + dispatch), chain_small.cu (n <= 16), chain_big.cu (any n)
int qocvl_chain_configure(qocvl_plan* p);
int qocvl_launch_chain(qocvl_plan* p, bool want_grad, double* out3, double* gwave, cudaStream_t st);
int qocvl_small_configure(qocvl_plan* p);
int qocvl_small_launch(qocvl_plan* p, bool want_grad, cudaStream_t st);
int qocvl_tp_configure(qocvl_plan* p);
int qocvl_tp_launch(qocvl_plan* p, bool want_grad, double* gwave, cudaStream_t st);
int qocvl_big_configure(qocvl_plan* p);
int qocvl_big_launch(qocvl_plan* p, bool want_grad, cudaStream_t st);
int qocvl_wide_configure(qocvl_plan* p);
int qocvl_wide_launch(qocvl_plan* p, bool want_grad, double* gwave, cudaStream_t st);
int qocvl_aug_configure(qocvl_plan* p);
int qocvl_aug_launch(qocvl_plan* p, bool want_grad, double* gwave, cudaStream_t st);
int qocvl_duo_configure(qocvl_plan* p);
int qocvl_duo_launch(qocvl_plan* p, bool want_grad, double* gwave, cudaStream_t st);
int qocvl_build_lane_layout(qocvl_plan* p, const std::vector<std::complex<double>>& blocks, int n, int npad,
                            bool diag_slot, BlockLayout* out);
int qocvl_build_csr(qocvl_plan* p, const std::vector<std::complex<double>>& blocks, CsrLayout* out);
// dense.cu
int qocvl_launch_dense_expm(qocvl_plan* p, cplx* out, cudaStream_t st);
int qocvl_launch_dense_tree(qocvl_plan* p, cplx* slices, cplx* out, cudaStream_t st);
// dense_blocks.cu (n > 16: block-triangular pairs, batched DMMA complex GEMM)
int qocvl_launch_block_expm(qocvl_plan* p, cplx* out, cudaStream_t st);
int qocvl_launch_block_propagate(qocvl_plan* p, cplx* out, cudaStream_t st);
