// chain_wide.cuh -- kernel template of chain_wide.cu (see its header); instantiated per compiled generator-block
// count in chain_wide_b13.cu / chain_wide_b24.cu so that the variants compile in parallel.
#pragma once
#include "common.cuh"

#define WIDE_THREADS 512
#define WIDE_MAXB 2     // generator blocks with off-diagonal entries
#define WIDE_MAXD 4     // generator blocks with diagonal entries
#ifndef WIDE_CHUNK
#define WIDE_CHUNK 2    // entries per list chunk: one 4-byte (2) or 8-byte (4) load of 16-bit element offsets
#endif
#ifndef WIDE_GATHER_UNROLL
#define WIDE_GATHER_UNROLL 1
#endif
constexpr int kWideGatherUnroll = WIDE_GATHER_UNROLL;   // list chunks in flight per thread
#define WIDE_NVP 2      // propagated vectors: one moving start ket q and p = W psi

struct WideArgs {
  int n, J, M, C, S, qcap, want_grad, store, ngroups, nb, nd;
  int ent_fwd, ent_rev, dreal;              // padded entries per direction; diagonals stored as real doubles
  double dt, inv_duration, w_infid, w_adiab, inv_fid_norm, inv_samples;
  cplx d_const;
  int bgen[WIDE_MAXB], bfam[WIDE_MAXB];   // generator and family (0 = diagonal blocks u, 1 = top-right block w)
  int buni[WIDE_MAXB];                    // every off-diagonal entry of the block has the same value bval
  cplx bval[WIDE_MAXB];
  int dgen[WIDE_MAXD], dfam[WIDE_MAXD];
  const int* grow;        // [ngroups * RW] row owned by (group, row slot), -1 = padding
  const int2* fwd_co;     // [nb][ngroups] {entries per row of the group, offset of the group's entry block}
  const int2* rev_co;
  const cplx* fwd_val;    // values in list order (only read for blocks whose entries differ)
  const unsigned short* fwd_col;   // pre-scaled element offsets row * NVP * SPC, lists [row slot][padded entries]
  const cplx* rev_val;    // column lists (row, G[row][col]) -- conjugated in the kernel
  const unsigned short* rev_col;
  const cplx* dval;       // [nd][ngroups][RW] diagonal of each diagonal block, in owner order
  const double* dval_re;  // the same as real doubles (dreal)
  const unsigned char* dmask;   // [ngroups][RW] bit d set: diagonal block d is non-zero on the row
  const cplx* coef;       // [M][J]
  const cplx* dcoef;      // [M][J][C]
  const int* qdeg;        // [M]
  const double* mul;      // [S][J]
  const cplx* add;        // [S][J]
  const cplx* starts;     // [1][n]
  const cplx* targets;    // [1][n]  (rotated by the accumulated phase of the shifts when `shift` is set)
  const double* shift;    // [M] phi_k: the sweeps apply X_k - i phi_k I (null: no shift)
  // time-parallel single-sample path: mode 1 = CTA propagates 8 basis columns through one chunk of slices (forward
  // only) and writes the columns of the chunk's Van Loan blocks U_c, W_c; mode 2 = CTA runs the forward + reverse
  // sweep of one chunk between the boundary states / cotangents k_wide_combine stitched together; mode 0 = ensemble.
  int mode, chunk_len, n_chunks, ctas_per_chunk;
  cplx *tp_uc, *tp_wc;    // [chunk][column][n]
  const cplx *tp_zc, *tp_lc;   // [chunk + 1][NVP][n]
  double* tp_gwave;       // [M][C], written directly (every slice belongs to one chunk)
  cplx* ring;             // [cta][(store ? slices of the CTA : 1) * qcap][NVP][n][SPC]
  cplx* ckpt;             // recompute mode: [cta][M][NVP][n][SPC]
  double* sample_out;     // [S][3]
  double* gw_partial;     // [cta][M][C]
};

__device__ __forceinline__ cplx ld_stream(const cplx* p) {
  const double2 v = __ldcs(reinterpret_cast<const double2*>(p));
  return v;
}
__device__ __forceinline__ void st_stream(cplx* p, cplx v) { __stcs(reinterpret_cast<double2*>(p), v); }
__device__ __forceinline__ cplxf ld_stream(const cplxf* p) { return __ldcs(reinterpret_cast<const float2*>(p)); }
__device__ __forceinline__ void st_stream(cplxf* p, cplxf v) { __stcs(reinterpret_cast<float2*>(p), v); }
// arithmetic type of the sweeps: CT = cplx (complex128, the parity path) or cplxf (the optional complex64 mode: vectors,
// Taylor terms, the HBM stream / checkpoints and the per-slice Ybar accumulators in fp32; coefficients, cost, terminal
// cotangents, chunk boundaries and every reduction stay fp64)
template <typename CT> struct WideReal;
template <> struct WideReal<cplx> { typedef double R; };
template <> struct WideReal<cplxf> { typedef float R; };
__device__ __forceinline__ cplx wide_mk(double re, double im, cplx*) { return cmake(re, im); }
__device__ __forceinline__ cplxf wide_mk(double re, double im, cplxf*) { return cmakef((float)re, (float)im); }
template <typename CT> __device__ __forceinline__ CT wide_from(const cplx v) { return wide_mk(v.x, v.y, (CT*)nullptr); }
__device__ __forceinline__ cplx wide_to_d(const cplx v) { return v; }
__device__ __forceinline__ cplx wide_to_d(const cplxf v) { return cmake((double)v.x, (double)v.y); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// FWD = true compiles the forward sweep only (mode 1, the column sweeps of the time-parallel path): without the
// cotangent / gradient state it needs few enough registers for two CTAs per SM when a thread owns <= 4 rows.
template <typename CT, int SPC, int ITEMS, int THREADS, int NB, int ND, bool SIMPLE, bool FWD = false>
__global__ void __launch_bounds__(THREADS, (FWD && ITEMS <= 4) ? 2 : 1) k_chain_wide(const WideArgs P) {
  typedef typename WideReal<CT>::R R;
  constexpr int NVP = WIDE_NVP, NV = NVP - 1, VA = 0;   // vectors: q (index 0, also psi of the adiabaticity term), p
  constexpr int RW = 32 / SPC, NSLOT = NB + ND;   // compiled generator-block counts (off-diagonal, diagonal)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int s = lane % SPC, rw = lane / SPC;
  const int n = P.n, J = P.J, C = P.C, M = P.M;
  const int cta = blockIdx.x, mode = P.mode;
  int chunk = 0, colbase = 0, k_begin = 0, k_end = P.M;
  if (mode == 1) { chunk = cta / P.ctas_per_chunk; colbase = (cta % P.ctas_per_chunk) * SPC; }
  if (mode == 2) chunk = cta;
  if (mode) { k_begin = chunk * P.chunk_len; k_end = min(P.M, k_begin + P.chunk_len); }
  const int nslices = mode ? P.chunk_len : P.M;          // slices a CTA owns (ring / checkpoint stride)
  // live lanes of this CTA (padding lanes repeat the last): noise samples, or basis columns in mode 1
  const int nact = mode == 0 ? min(SPC, P.S - cta * SPC) : (mode == 1 ? min(SPC, n - colbase) : 1);
  const int nr = n + 1;                                  // row n of every vector stays zero (read by padding entries)
  const int VS = NVP * nr * SPC;                         // elements of one "all vectors" buffer
  const CT czero = wide_mk(0.0, 0.0, (CT*)nullptr);
  const cplx dzero = cmake(0, 0);

  // ---- shared memory (the sweeps' own type first, then the fp64 tables)
  CT* s_buf = (CT*)smem_raw;                             // Taylor term / cotangent term, double buffered
  CT* s_cs = s_buf + 2 * VS;                             // dt * (coef_kg * mul_sg + add_sg), double buffered by slice
  cplx* carve = (cplx*)(s_cs + 2 * J * SPC);
  cplx* s_add = carve; carve += J * SPC;
  cplx* s_dco = carve; carve += J * C;
  cplx* s_scr = carve; carve += (size_t)(THREADS / 32) * NSLOT * SPC;
  cplx* s_ysum = carve; carve += NSLOT * SPC;
  cplx* s_dvc = carve; carve += P.dreal ? 0 : ND * P.ngroups * RW;     // diagonals of the diagonal blocks
  double* s_mul = (double*)carve;                        // [J][SPC]
  double* s_red = s_mul + J * SPC;                       // [nwarps][3][SPC] cost reductions
  double* s_dsum = s_red + (THREADS / 32) * 3 * SPC;   // [2][SPC]
  double* s_dvr = s_dsum + 2 * SPC;                      // diagonals as real doubles
  int2* s_fco = (int2*)(s_dvr + (P.dreal ? ND * P.ngroups * RW : 0));   // [NB][ngroups] {entries, offset}
  int2* s_rco = s_fco + NB * P.ngroups;
  unsigned short* s_fcol = (unsigned short*)(s_rco + NB * P.ngroups);   // column indices, [entry][RW] per group
  unsigned short* s_rcol = s_fcol + ((P.ent_fwd + 7) & ~7);
  unsigned char* s_dmask = (unsigned char*)(s_rcol + ((P.ent_rev + 7) & ~7));   // diagonal blocks that are non-zero on a row
  for (int i = tid; i < P.ngroups * RW; i += blockDim.x) s_dmask[i] = P.dmask[i];
  for (int i = tid; i < P.nb * P.ngroups; i += blockDim.x) { s_fco[i] = P.fwd_co[i]; s_rco[i] = P.rev_co[i]; }
  for (int i = tid; i < P.ent_fwd; i += blockDim.x) s_fcol[i] = P.fwd_col[i];
  for (int i = tid; i < P.ent_rev; i += blockDim.x) s_rcol[i] = P.rev_col[i];
  for (int i = tid; i < P.nd * P.ngroups * RW; i += blockDim.x) {
    if (P.dreal) s_dvr[i] = P.dval_re[i]; else s_dvc[i] = P.dval[i];
  }
  auto diag_value = [&](int d, int g) {
    const int idx = (d * P.ngroups + g) * RW + rw;
    return P.dreal ? wide_mk(s_dvr[idx], 0.0, (CT*)nullptr) : wide_from<CT>(s_dvc[idx]);
  };

  // ---- owned rows
  int row[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int g = i * nwarps + warp;
    row[i] = (g < P.ngroups) ? P.grow[g * RW + rw] : -1;
  }
  auto at = [&](int v, int r) { return (r * NVP + v) * SPC + s; };     // [row][vector][sample]: p is 128 B after q
  for (int i = tid; i < 2 * NVP * SPC; i += blockDim.x)
    s_buf[(i / (NVP * SPC)) * VS + n * NVP * SPC + (i % (NVP * SPC))] = czero;

  for (int i = tid; i < J * SPC; i += blockDim.x) {
    const int g = i / SPC, ss = i % SPC, sm = mode ? 0 : cta * SPC + min(ss, nact - 1);
    s_mul[i] = P.mul[(size_t)sm * J + g];
    s_add[i] = P.add[(size_t)sm * J + g];
  }
  __syncthreads();
  auto load_cs = [&](int k, int b) {
    for (int i = tid; i < J * SPC; i += blockDim.x) {
      const int g = i / SPC;
      s_cs[b * J * SPC + i] = wide_from<CT>(cscale(cadd(cscale(P.coef[(size_t)k * J + g], s_mul[i]), s_add[i]), P.dt));
    }
  };
  // coefficients of this thread's sample for the generator blocks of the slice, in registers
  CT csb[NB], csd[ND];
  CT bvl[NB];                                            // common value of every uniform-valued block
#pragma unroll
  for (int x = 0; x < NB; ++x) bvl[x] = wide_from<CT>(P.bval[x]);
  CT csh = czero;                                        // -i phi_k: diagonal shift of the slice (see k_wide_degrees)
  auto coefficients = [&](int b) {
#pragma unroll
    for (int x = 0; x < NB; ++x) csb[x] = (x < P.nb) ? s_cs[(b * J + P.bgen[x]) * SPC + s] : czero;
#pragma unroll
    for (int x = 0; x < ND; ++x) csd[x] = (x < P.nd) ? s_cs[(b * J + P.dgen[x]) * SPC + s] : czero;
  };

  CT z[ITEMS][NVP];        // forward: state; reverse: cotangent lambda
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    z[i][0] = (row[i] >= 0) ? wide_from<CT>(P.starts[row[i]]) : czero;
    z[i][NV] = czero;
    if (mode == 1) z[i][0] = (row[i] == colbase + min(s, nact - 1)) ? wide_mk(1.0, 0.0, (CT*)nullptr) : czero;   // basis column
    if (mode == 2 && row[i] >= 0) {
      z[i][0] = wide_from<CT>(P.tp_zc[((size_t)chunk * NVP + 0) * n + row[i]]);
      z[i][NV] = wide_from<CT>(P.tp_zc[((size_t)chunk * NVP + NV) * n + row[i]]);
    }
  }
  // sum of the gathered elements of one list (WIDE_CHUNK pre-scaled 16-bit indices per 4- or 8-byte load; lists are
  // padded to whole chunks: 7.45 entries per row on the 12-atom chain with chunks of 2, 8.49 with chunks of 4, 6.94 real):
  // q and p, or q only
  auto gather_sum = [&](const CT* ts, const unsigned short* list, int chunks, bool both, CT& sq, CT& sp) {
#if WIDE_CHUNK == 4
    if (both) {
#pragma unroll kWideGatherUnroll
      for (int e = 0; e < chunks; ++e) {
        const uint2 c4 = *reinterpret_cast<const uint2*>(list + 4 * e);
        const CT* p0 = ts + (c4.x & 0xffffu); const CT* p1 = ts + (c4.x >> 16);
        const CT* p2 = ts + (c4.y & 0xffffu); const CT* p3 = ts + (c4.y >> 16);
        sq = cadd(sq, cadd(cadd(p0[0], p1[0]), cadd(p2[0], p3[0])));
        sp = cadd(sp, cadd(cadd(p0[SPC], p1[SPC]), cadd(p2[SPC], p3[SPC])));
      }
    } else {
#pragma unroll kWideGatherUnroll
      for (int e = 0; e < chunks; ++e) {
        const uint2 c4 = *reinterpret_cast<const uint2*>(list + 4 * e);
        sq = cadd(sq, cadd(cadd(ts[c4.x & 0xffffu], ts[c4.x >> 16]), cadd(ts[c4.y & 0xffffu], ts[c4.y >> 16])));
      }
    }
#else
    if (both) {
#pragma unroll kWideGatherUnroll
      for (int e = 0; e < chunks; ++e) {
        const unsigned c2 = *reinterpret_cast<const unsigned*>(list + 2 * e);
        const CT* p0 = ts + (c2 & 0xffffu); const CT* p1 = ts + (c2 >> 16);
        sq = cadd(sq, cadd(p0[0], p1[0]));
        sp = cadd(sp, cadd(p0[SPC], p1[SPC]));
      }
    } else {
#pragma unroll kWideGatherUnroll
      for (int e = 0; e < chunks; ++e) {
        const unsigned c2 = *reinterpret_cast<const unsigned*>(list + 2 * e);
        sq = cadd(sq, cadd(ts[c2 & 0xffffu], ts[c2 >> 16]));
      }
    }
#endif
  };
  // one forward Taylor step of one owned row:  y = (X t)[row] / j
  auto forward_item = [&](int i, const CT* t, R invj, CT* y) {
    const int r = row[i], g = i * nwarps + warp;
    const CT* ts = t + s;                        // element (row, q) of this thread's sample; p follows SPC later
    const CT tq = t[at(VA, r)], tp = t[at(NV, r)];
    CT accq = cmul(csh, tq), accp = cmul(csh, tp);       // the shift's diagonal
    if (SIMPLE) {          // one off-diagonal block, uniform value, acting on q and p (checked on the host)
      const int2 co = s_fco[g];
      CT sq = czero, sp = czero;
      gather_sum(ts, s_fcol + co.y + rw * WIDE_CHUNK * co.x, co.x, true, sq, sp);
      const CT c = cmul(csb[0], bvl[0]);
      accq = cfma(c, sq, accq);
      accp = cfma(c, sp, accp);
    } else {
#pragma unroll
    for (int bb = 0; bb < NB; ++bb) {
      if (bb < P.nb) {
        const int2 co = s_fco[bb * P.ngroups + g];            // {chunks per row, offset of the group's lists}
        const unsigned short* list = s_fcol + co.y + rw * WIDE_CHUNK * co.x;
        const bool both = (P.bfam[bb] == 0);     // u blocks act on q and p, w blocks map q into p
        CT sq = czero, sp = czero;
        if (P.buni[bb]) {                        // one common value: plain sums, scaled once
          gather_sum(ts, list, co.x, both, sq, sp);
          sq = cmul(bvl[bb], sq);
          sp = cmul(bvl[bb], sp);
        } else {
          const cplx* ev = P.fwd_val + co.y + rw * WIDE_CHUNK * co.x;
#pragma unroll 2
          for (int e = 0; e < WIDE_CHUNK * co.x; ++e) {
            const CT val = wide_from<CT>(__ldg(&ev[e]));
            const int col = list[e];
            sq = cfma(val, ts[col], sq);
            if (both) sp = cfma(val, ts[col + SPC], sp);
          }
        }
        if (both) { accq = cfma(csb[bb], sq, accq); accp = cfma(csb[bb], sp, accp); }
        else accp = cfma(csb[bb], sq, accp);
      }
    }
    }
    const unsigned dm = s_dmask[g * RW + rw];
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      if ((dm >> d) & 1u) {
        const CT c = cmul(csd[d], diag_value(d, g));
        if (P.dfam[d] == 0) { accq = cfma(c, tq, accq); accp = cfma(c, tp, accp); }
        else accp = cfma(c, tq, accp);
      }
    }
    y[0] = cscale(accq, invj);
    y[NV] = cscale(accp, invj);
  };

  const size_t ring_slices = P.store ? (size_t)nslices : 1;
  CT* ring = (CT*)P.ring + (size_t)cta * ring_slices * P.qcap * VS;    // (elements of the sweeps' type)
  CT* ckpt = P.store ? nullptr : (CT*)P.ckpt + (size_t)cta * nslices * VS;

  // =================================== forward sweep ===================================
  load_cs(k_begin, k_begin & 1);
  __syncthreads();
  for (int k = k_begin; k < k_end; ++k) {
    const int q = P.qdeg[k], b = k & 1;
    if (k + 1 < k_end) load_cs(k + 1, b ^ 1);
    coefficients(b);
    if (P.shift) csh = wide_mk(0.0, -P.shift[k], (CT*)nullptr);
    const int kl = k - k_begin;
    CT* t0 = (P.store ? ring + (size_t)kl * P.qcap * VS : ckpt + (size_t)kl * VS);
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      if (row[i] < 0) continue;
#pragma unroll
      for (int v = 0; v < NVP; ++v) {
        s_buf[at(v, row[i])] = z[i][v];
        if (!FWD && P.want_grad) st_stream(&t0[at(v, row[i])], z[i][v]);
      }
    }
    __syncthreads();
    int cur = 0;
    for (int j = 1; j <= q; ++j) {
      const CT* t = s_buf + cur * VS;
      CT* y = s_buf + (cur ^ 1) * VS;
      const R invj = (R)(1.0 / (double)j);
      CT* rj = ring + ((size_t)kl * P.qcap + j) * VS;
      const bool keep = (j < q);
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        if (row[i] < 0) continue;
        CT yy[NVP];
        forward_item(i, t, invj, yy);
#pragma unroll
        for (int v = 0; v < NVP; ++v) {
          z[i][v] = cadd(z[i][v], yy[v]);
          if (keep) {
            y[at(v, row[i])] = yy[v];
            if (!FWD && P.want_grad && P.store) st_stream(&rj[at(v, row[i])], yy[v]);
          }
        }
      }
      __syncthreads();
      cur ^= 1;
    }
  }

  if (FWD || mode == 1) {       // columns colbase .. of the chunk's blocks: U_c e_b = q, W_c e_b = p
    if (s < nact) {
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        if (row[i] < 0) continue;
        const size_t o = ((size_t)chunk * n + colbase + s) * n + row[i];
        P.tp_uc[o] = wide_to_d(z[i][0]);
        P.tp_wc[o] = wide_to_d(z[i][NV]);
      }
    }
    return;
  }
  // =================================== cost ============================================
  if (mode == 0) {
    double red[3] = {0, 0, 0};   // Re d, Im d, Re q^dag p
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      if (row[i] < 0) continue;
      const cplx y = P.targets[row[i]], zq = wide_to_d(z[i][0]), zp = wide_to_d(z[i][NV]);
      red[0] += y.x * zq.x + y.y * zq.y;
      red[1] += y.x * zq.y - y.y * zq.x;
      red[2] += zq.x * zp.x + zq.y * zp.y;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      for (int off = SPC; off < 32; off <<= 1) red[c] += __shfl_xor_sync(0xffffffffu, red[c], off);
      if (rw == 0) s_red[(warp * 3 + c) * SPC + s] = red[c];
    }
    __syncthreads();
    if (tid < SPC) {
      double a[3] = {0, 0, 0};
      for (int w = 0; w < nwarps; ++w)
        for (int c = 0; c < 3; ++c) a[c] += s_red[(w * 3 + c) * SPC + tid];
      const double dx = a[0] + P.d_const.x, dy = a[1] + P.d_const.y;
      s_dsum[tid] = dx; s_dsum[SPC + tid] = dy;
      if (tid < nact) {
        const double infid = 1.0 - (dx * dx + dy * dy) * P.inv_fid_norm;
        const double adiab = 1.0 - a[2] * P.inv_duration;
        double* o = P.sample_out + (size_t)(cta * SPC + tid) * 3;
        o[0] = P.w_infid * infid + P.w_adiab * adiab;
        o[1] = infid;
        o[2] = adiab;
      }
    }
    __syncthreads();
  }
  if (!P.want_grad) return;

  // terminal cotangents (scaled by 1 / global sample count):  lam_q = cf * d * y + ca * p,  lam_p = ca * q
  if (mode == 2) {
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      if (row[i] < 0) continue;
      z[i][0] = wide_from<CT>(P.tp_lc[((size_t)(chunk + 1) * NVP + 0) * n + row[i]]);
      z[i][NV] = wide_from<CT>(P.tp_lc[((size_t)(chunk + 1) * NVP + NV) * n + row[i]]);
    }
  } else {
    const double ca = -0.5 * P.w_adiab * P.inv_duration * P.inv_samples;
    const double cf = -P.w_infid * P.inv_fid_norm * P.inv_samples;
    const cplx dsum = cmake(s_dsum[s], s_dsum[SPC + s]);
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      if (row[i] < 0) continue;
      const cplx zq = wide_to_d(z[i][0]), zp = wide_to_d(z[i][NV]);
      z[i][0] = wide_from<CT>(cadd(cscale(cmul(dsum, P.targets[row[i]]), cf), cscale(zp, ca)));
      z[i][NV] = wide_from<CT>(cscale(zq, ca));
    }
  }

  // =================================== reverse sweep ===================================
  load_cs(k_end - 1, (k_end - 1) & 1);
  __syncthreads();
  for (int k = k_end - 1; k >= k_begin; --k) {
    const int q = P.qdeg[k], b = k & 1, qprev = k > k_begin ? P.qdeg[k - 1] : 0, kl = k - k_begin;
    if (k > k_begin) load_cs(k - 1, b ^ 1);
    coefficients(b);
    if (P.shift) csh = wide_mk(0.0, -P.shift[k], (CT*)nullptr);
    CT* rk = ring + (P.store ? (size_t)kl * P.qcap * VS : 0);
    int cur = 0;
    if (!P.store) {
      // ---- recompute the forward terms t_0 .. t_{q-1} of this slice into the ring (own elements)
      const CT* ck = ckpt + (size_t)kl * VS;
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        if (row[i] < 0) continue;
#pragma unroll
        for (int v = 0; v < NVP; ++v) {
          const CT zz = ld_stream(&ck[at(v, row[i])]);
          s_buf[at(v, row[i])] = zz;
          st_stream(&rk[at(v, row[i])], zz);
        }
      }
      __syncthreads();
      for (int j = 1; j < q; ++j) {
        const CT* t = s_buf + cur * VS;
        CT* y = s_buf + (cur ^ 1) * VS;
        CT* rj = rk + (size_t)j * VS;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
          if (row[i] < 0) continue;
          CT yy[NVP];
          forward_item(i, t, (R)(1.0 / (double)j), yy);
#pragma unroll
          for (int v = 0; v < NVP; ++v) { y[at(v, row[i])] = yy[v]; st_stream(&rj[at(v, row[i])], yy[v]); }
        }
        __syncthreads();
        cur ^= 1;
      }
    }
    // ---- descending adjoint recurrence  tb_{j-1} = lam + X^dag tb_j / j
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      if (row[i] < 0) continue;
#pragma unroll
      for (int v = 0; v < NVP; ++v) s_buf[cur * VS + at(v, row[i])] = z[i][v];
    }
    __syncthreads();
    // (the previous slice's readers of s_dco / s_ysum are past this barrier)
    for (int i = tid; i < J * C; i += blockDim.x) s_dco[i] = P.dcoef[(size_t)k * J * C + i];
    CT yb[NB], yd[ND];     // Ybar of every generator block: sum_j <G^dag tb_j , t_{j-1}> / j
#pragma unroll
    for (int bb = 0; bb < NB; ++bb) yb[bb] = czero;
#pragma unroll
    for (int d = 0; d < ND; ++d) yd[d] = czero;
    for (int j = q; j >= 1; --j) {
      const CT* t = s_buf + cur * VS;
      CT* y = s_buf + (cur ^ 1) * VS;
      const R invj = (R)(1.0 / (double)j);
      const CT* rj = rk + (size_t)(j - 1) * VS;
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        const int c = row[i], g = i * nwarps + warp;
        if (c < 0) continue;
        const CT fq_raw = ld_stream(&rj[at(VA, c)]);               // own elements of t_{j-1}, used after the gathers
        const CT fp_raw = ld_stream(&rj[at(NV, c)]);
        if (j > 2) {                                                // pull the terms of step j-2 from HBM into L2
          prefetch_l2(&rj[at(VA, c) - 2 * VS]);
          prefetch_l2(&rj[at(NV, c) - 2 * VS]);
        } else if (P.store && k > k_begin && qprev - 3 + j >= 0) {       // ... or the top terms of the next slice to process
          const CT* nx = rk - (size_t)P.qcap * VS + (size_t)(qprev - 3 + j) * VS;
          prefetch_l2(&nx[at(VA, c)]);
          prefetch_l2(&nx[at(NV, c)]);
        }
        const CT bq = t[at(VA, c)], bp = t[at(NV, c)];             // own elements of tb_j
        CT accq = cfma_conj(csh, bq, czero), accp = cfma_conj(csh, bp, czero);   // (X - i phi)^dag: the shift's diagonal
        CT sqb[NB], spb[NB];                                        // (G^dag tb_q)[c], (G^dag tb_p)[c] per block
        if (SIMPLE) {
          const int2 co = s_rco[g];
          CT sq = czero, sp = czero;
          gather_sum(t + s, s_rcol + co.y + rw * WIDE_CHUNK * co.x, co.x, true, sq, sp);
          const CT vc = cconj(bvl[0]);
          sq = cmul(vc, sq);
          sp = cmul(vc, sp);
          accq = cfma_conj(csb[0], sq, accq);
          accp = cfma_conj(csb[0], sp, accp);
          sqb[0] = sq; spb[0] = sp;
        } else {
#pragma unroll
        for (int bb = 0; bb < NB; ++bb) {
          sqb[bb] = czero; spb[bb] = czero;
          if (bb < P.nb) {
            const int2 co = s_rco[bb * P.ngroups + g];
            const unsigned short* list = s_rcol + co.y + rw * WIDE_CHUNK * co.x;
            const bool both = (P.bfam[bb] == 0);   // w blocks: (G_w^dag tb_p)[c] feeds q
            CT sq = czero, sp = czero;
            if (P.buni[bb]) {
              if (both) gather_sum(t + s, list, co.x, true, sq, sp);
              else gather_sum(t + s + SPC, list, co.x, false, sp, sq);
              const CT vc = cconj(bvl[bb]);
              sq = cmul(vc, sq);
              sp = cmul(vc, sp);
            } else {
              const cplx* ev = P.rev_val + co.y + rw * WIDE_CHUNK * co.x;
              const CT* ts = t + s;
#pragma unroll 2
              for (int e = 0; e < WIDE_CHUNK * co.x; ++e) {
                const CT val = wide_from<CT>(__ldg(&ev[e]));
                const int a = list[e];
                sp = cfma_conj(val, ts[a + SPC], sp);
                if (both) sq = cfma_conj(val, ts[a], sq);
              }
            }
            if (both) { accq = cfma_conj(csb[bb], sq, accq); accp = cfma_conj(csb[bb], sp, accp); }
            else accq = cfma_conj(csb[bb], sp, accq);
            sqb[bb] = sq; spb[bb] = sp;
          }
        }
        }
        const CT fq = cscale(fq_raw, invj), fp = cscale(fp_raw, invj);
#pragma unroll
        for (int bb = 0; bb < NB; ++bb) {
          if (bb < P.nb) {
            if (SIMPLE || P.bfam[bb] == 0) yb[bb] = cfma_conj(spb[bb], fp, cfma_conj(sqb[bb], fq, yb[bb]));
            else yb[bb] = cfma_conj(spb[bb], fq, yb[bb]);
          }
        }
        // diagonal blocks: Ybar_d += dval * (conj(tb_q) f_q + conj(tb_p) f_p)  (u)   |   dval * conj(tb_p) f_q  (w)
        const CT wsum = cfma_conj(bp, fp, cfma_conj(bq, fq, czero));
        const CT wpq = cfma_conj(bp, fq, czero);
        const unsigned dm = s_dmask[g * RW + rw];
#pragma unroll
        for (int d = 0; d < ND; ++d) {
          if ((dm >> d) & 1u) {
            const CT dv = diag_value(d, g);
            const CT cc = cmul(csd[d], dv);
            if (P.dfam[d] == 0) {
              accq = cfma_conj(cc, bq, accq); accp = cfma_conj(cc, bp, accp);
              yd[d] = cfma(dv, wsum, yd[d]);
            } else {
              accq = cfma_conj(cc, bp, accq);
              yd[d] = cfma(dv, wpq, yd[d]);
            }
          }
        }
        const CT nq = cadd(z[i][0], cscale(accq, invj)), np = cadd(z[i][NV], cscale(accp, invj));
        if (j > 1) { y[at(VA, c)] = nq; y[at(NV, c)] = np; }
        else { z[i][0] = nq; z[i][NV] = np; }
      }
      __syncthreads();
      cur ^= 1;
    }
    // ---- d cost / d wave[k][c] = sum_s sum_g Re(2 * mul_sg * Ybar_g(s) * dcoef[k][g][c])
    auto warp_partial = [&](cplx v, int x) {
      for (int off = SPC; off < 32; off <<= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, off);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, off);
      }
      if (rw == 0) s_scr[(warp * NSLOT + x) * SPC + s] = v;
    };
#pragma unroll
    for (int bb = 0; bb < NB; ++bb) warp_partial(wide_to_d(yb[bb]), bb);
#pragma unroll
    for (int dd = 0; dd < ND; ++dd) warp_partial(wide_to_d(yd[dd]), NB + dd);
    __syncthreads();
    if (tid < NSLOT * SPC) {
      cplx a = dzero;
      for (int w = 0; w < nwarps; ++w) a = cadd(a, s_scr[w * NSLOT * SPC + tid]);
      s_ysum[tid] = a;
    }
    __syncthreads();
    if (tid < C) {
      double gw = 0.0;
      for (int x = 0; x < NSLOT; ++x) {
        int g;
        if (x < NB) { if (x >= P.nb) continue; g = P.bgen[x]; }
        else { if (x - NB >= P.nd) continue; g = P.dgen[x - NB]; }
        const cplx dc = s_dco[g * C + tid];
        for (int ss = 0; ss < nact; ++ss) {
          const cplx yv = cscale(s_ysum[x * SPC + ss], 2.0 * P.dt * s_mul[g * SPC + ss]);
          gw += yv.x * dc.x - yv.y * dc.y;
        }
      }
      if (mode == 2) P.tp_gwave[(size_t)k * C + tid] = gw;
      else P.gw_partial[((size_t)cta * M + k) * C + tid] = gw;
    }
    // s_dco / s_scr / s_ysum are rewritten only after a barrier of the next slice
  }
}

typedef void (*wide_kernel_t)(const WideArgs);
template <typename CT, int NB, int ND, bool SIMPLE>
static inline wide_kernel_t wide_pick_blocks(int spc, int items, int threads) {
  if (spc == 8 && threads == 512) {
    if (items <= 2) return k_chain_wide<CT, 8, 2, 512, NB, ND, SIMPLE>;
    if (items <= 4) return k_chain_wide<CT, 8, 4, 512, NB, ND, SIMPLE>;
    return nullptr;
  }
  if (spc == 8 && threads == 384) {
    if (items <= 8) return k_chain_wide<CT, 8, 8, 384, NB, ND, SIMPLE>;
    return nullptr;
  }
  if (spc == 8 && threads == 256) {
    if (items <= 14) return k_chain_wide<CT, 8, 14, 256, NB, ND, SIMPLE>;
    return nullptr;
  }
  if (spc == 1) {
    if (items <= 1) return k_chain_wide<CT, 1, 1, 512, NB, ND, SIMPLE>;
    if (items <= 2) return k_chain_wide<CT, 1, 2, 512, NB, ND, SIMPLE>;
  }
  return nullptr;
}
// forward-only column sweeps (8 basis columns per CTA, 384 threads)
template <typename CT, int NB, int ND, bool SIMPLE>
static inline wide_kernel_t wide_pick_blocks_fwd(int items) {
  if (items <= 4) return k_chain_wide<CT, 8, 4, 384, NB, ND, SIMPLE, true>;
  if (items <= 8) return k_chain_wide<CT, 8, 8, 384, NB, ND, SIMPLE, true>;
  return nullptr;
}

// (c64 = 1: the complex64 instantiations, compiled in chain_wide_b13f.cu / chain_wide_b24f.cu)
wide_kernel_t qocvl_wide_pick_b13_fwd(int items, int c64);
wide_kernel_t qocvl_wide_pick_b24_fwd(int items, int c64);
wide_kernel_t qocvl_wide_pick_b13(int spc, int items, int threads, int c64);   // one uniform-valued off-diagonal u block, <= 3 diagonal blocks
wide_kernel_t qocvl_wide_pick_b24(int spc, int items, int threads, int c64);   // <= 2 off-diagonal, <= 4 diagonal blocks
wide_kernel_t qocvl_wide_pick_b13f_fwd(int items);
wide_kernel_t qocvl_wide_pick_b24f_fwd(int items);
wide_kernel_t qocvl_wide_pick_b13f(int spc, int items, int threads);
wide_kernel_t qocvl_wide_pick_b24f(int spc, int items, int threads);
