// chain_duo.cuh -- the ensemble hot kernels, third generation: every coupled sector of the reduced stacked system on
// TWO lanes, one warp per scheduler, everything in registers.
//
// Same numbers as chain_aug.cu (ST:247-302 / TQ:460-567 and the reverse pass of jax.value_and_grad,
// TQ:592) on the same reduced stacked system Xaug (chain_aug.cu builds it: reachability + orbit folding; CZ: 12 rows,
// 36 non-zeros).  What changes is the mapping, driven by the profiles of the older kernels (issue slots
// ~50 %, FP64 pipe 24-45 % busy, `wait` / `short_scoreboard` the top stalls: every Taylor step paid an exchange round
// trip per ROW, and half of the issued DFMAs were lane padding):
//   * A role (connected component of the stacked pattern; CZ: {U|10>, W|10>} and {U|11>}, 6 rows each) is cut into two
//     TRIPLES of rows, one lane each: a lane keeps its 3 vector elements, its 3x3 own block and the <= 2x2 block that
//     couples it to the partner lane in REGISTERS.  The sparsity of both blocks is a compile-time mask, so a Taylor
//     step is a fully unrolled, branch-free sequence of DFMAs (CZ: 11 / 9 complex multiply-adds per lane and step
//     for 9 useful on average) and the only communication is ONE xor-1 shuffle of the two "send" rows per step.
//     The host finds the cut and the row order by brute force (<= 720 permutations per role) and picks the cheapest
//     compiled pattern that covers it; a system with a sector that does not split this way falls back to chain_aug.cu.
//   * 16 samples x 2 lanes per warp, one role per CTA (role = blockIdx % nroles, uniform for the compiler), 4 warps
//     per CTA: at 4096 samples that is 512 warps = one warp per scheduler on 128 SMs.  A lone warp per scheduler can
//     keep the FP64 pipe busy because the unrolled step carries 6 independent accumulation chains.
//   * Unscaled Taylor powers (u_j = X u_{j-1}, z' = sum_j u_j / j!) and the matching adjoint w_j = tbar_j / j!
//     (w_{j-1} = lam / (j-1)! + X^dag w_j,  Xbar = sum_j conj(w_j) u_{j-1}^T): one DFMA pair with
//     a constant operand per row and step instead of a scale and an add.
//   * HBM carries only the slice checkpoints z_k (192 B per sample and slice, written by the forward kernel, read
//     back by the reverse kernel): the older kernels streamed every Taylor term (5.2-5.9 KB per unit, 40+ GB per
//     evaluation), which capped them at the HBM roofline.  The reverse kernel recomputes u_0 .. u_{q-1} into a
//     per-warp shared-memory slab instead (one mat-vec per term).
//   * TMA staging of slice batches (north_star (3)): the per-slice entry tables (sample independent, L2 resident) and
//     the checkpoints of DUO_BT consecutive slices arrive in shared memory through 1-D bulk copies
//     (cp.async.bulk.shared::cluster.global + mbarrier complete_tx), two stages per warp, issued by one lane one batch
//     ahead; there is no barrier of any kind between warps.
//   * The per-sample Xbar entries are scaled by the entry's class multiplier and summed over the warp's 16 samples
//     with xor shuffles; two lanes per warp write the sums (<= 13 complex per lane type and slice), k_duo_reduce adds
//     the warps in a fixed order and contracts with d coef / d wave.  Deterministic: no atomics anywhere.
// Forward sweep, terminal cost / cotangent and reverse sweep are three launches (the cost couples the roles of a
// sample; the checkpoints live in HBM anyway), plus the table kernel before and the reduction after.
#pragma once
#include "common.cuh"
#include <type_traits>

#define DUO_MAXR 4      // roles per system
#define DUO_BT 4        // slices per staged batch
#define DUO_SPW 16      // samples per warp
#define DUO_NVAR 5      // compiled patterns
#ifndef DUO_UNROLL
#define DUO_UNROLL 1     // unroll factor of the Taylor loops
#endif
constexpr int kDuoUnroll = DUO_UNROLL;

struct DuoArgs {
  int M, S, C, J, N, nroles, wpc, qcap, want_grad, KG;
  // Time-parallel evaluation of small ensembles (strong scaling): the slice axis is cut into `chunks` chunks of `lc`
  // slices.  mode 0: one warp sweeps all slices of its 16 samples.  mode 1: (sample group, chunk, basis column) items
  // push a unit vector of the role's 6-dimensional space through one chunk -> chunk propagators `prop`.
  // mode 2: (sample group, chunk) items sweep one chunk between the boundary states / cotangents that k_duo_combine
  // obtained from the chunk propagators.
  int mode, chunks, lc, G;
  int cpi;                // mode 1: basis columns swept by one item (DUO_COLS on the multi-column kernel, else 1)
  int ncol[DUO_MAXR], mirror[DUO_MAXR];   // mirror: 1 + lane type whose own-block propagator is copied to the other triple
  int variant[DUO_MAXR], nu[DUO_MAXR], nx[DUO_MAXR], warps[DUO_MAXR];
  unsigned long long tab_off[DUO_MAXR], con_off[DUO_MAXR], y_off[DUO_MAXR], warp0[DUO_MAXR];
  int ent0[DUO_MAXR];   // first (role, type, slot) descriptor of the role
  int ntot;             // descriptors in total: sum_r 2 nu_r
  double dt, inv_duration, w_infid, w_adiab, inv_fid_norm, inv_samples;
  cplx d_const;
  const cplx* tab;      // role r at tab_off[r]: [M][2 lane types][nu_r]   P_u(k) = dt sum_g coef_kg val_ug
  const double* mcls;   // [S][ncls] per-sample multiplier of every generator class (class 0 = 1)
  int ncls;
  unsigned clsmask[2 * DUO_MAXR], amask[2 * DUO_MAXR];   // per (role, lane type): entries of class 1 / with an additive part
  const cplx* con_a;    //                       [2][nu_r][S]  additive-noise part of the entry
  const int* srow;      // [nroles][2][3]  stacked row of (role, lane type, local row), -1 = padding row
  const int* ent_gen;   // [ntot][KG] generators of a descriptor (-1 = none)
  const cplx* ent_val;  // [ntot][KG]
  const int* ent_x;     // [ntot] index of the descriptor among its lane type's Xbar slots, -1 = not a gradient slot
  const cplx* coef;     // [M][J]
  const cplx* dcoef;    // [M][J][C]
  const double* phi;    // [nroles][M] phase shift of every role and slice (null: none), see k_slice_degrees
  const int* ent_srole; // [ntot] role whose shift the descriptor carries (own-block diagonal entries), -1 = none
  const int* qdeg;      // [nroles][M] Taylor degree of every role and slice (qstride = M), or [M] for all roles (qstride = 0)
  int qstride;
  const cplx* starts;   // [N]
  const cplx* targets;  // [N]
  const int* pair;      // [N]
  const double* wq;     // [2][N]
  cplx* ckpt;           // [global warp][M][3][32]
  cplx* zfin;           // [S][N]
  cplx* lam;            // [S][N]
  cplx* ypart;          // role r at y_off[r]: [warps_r][M][2][nx_r]
  const int* colrow;    // [nroles][6] local row (3 lane type + i) that basis column j of the role excites
  cplx* prop;           // [S][chunks][nroles][6 columns][6 local rows]  chunk propagators (mode 1 -> k_duo_combine)
  cplx* zstart;         // [S][chunks][N]  state at the start of chunk c
  cplx* lamend;         // [S][chunks][N]  cotangent at the end of chunk c
  double* sample_out;   // [S][3]
  double* gwave;        // [M][C]
  double ifact[QOCVL_QMAX + 2];   // 1 / j!
  float ifactf[QOCVL_QMAX + 2];   // the same for the complex64 sweeps
};

__host__ __device__ constexpr int duo_popc(unsigned v) { return v == 0 ? 0 : (int)(v & 1u) + duo_popc(v >> 1); }

// Compile-time description of a lane program: TM = 9-bit mask of the 3x3 own block (bit 3 i + i'), CM = 4-bit mask of
// the 2x2 coupling block (bit 2 g + h: my send row g <- partner's send row h), S0 / S1 = local index of the send rows.
// Compact slot order of a lane's entries: own block | coupling block as row owner | coupling block as column owner.
// XM = the own-block entries whose Xbar is accumulated (entries whose coefficient depends on no live pulse channel --
// the CZ model's diagonal on the pulse-map path, where the detuning channel is a forced constant -- need none).
__host__ __device__ constexpr int duo_nth_bit(unsigned mask, int n) {   // position of the n-th set bit
  int pos = 0;
  while (pos < 32) {
    if ((mask >> pos) & 1u) { if (n == 0) return pos; --n; }
    ++pos;
  }
  return 0;
}
template <int TM_, int CM_, int S0_, int S1_, int XM_ = TM_>
struct DuoPat {
  static constexpr int TM = TM_, CM = CM_, S0 = S0_, S1 = S1_, XM = XM_;
  static constexpr int NT = duo_popc(TM_), NC = duo_popc(CM_), NU = NT + 2 * NC, NXT = duo_popc(XM_), NX = NXT + NC;
  __host__ __device__ static constexpr bool hx(int i, int ip) { return (XM_ >> (3 * i + ip)) & 1; }
  __host__ __device__ static constexpr int xt(int i, int ip) { return duo_popc(XM_ & ((1u << (3 * i + ip)) - 1u)); }
  // entry slot (multiplier) of Xbar slot x
  __host__ __device__ static constexpr int ux(int x) {
    return x < NXT ? duo_popc(TM_ & ((1u << duo_nth_bit(XM_, x)) - 1u)) : x - NXT + NT + NC;
  }
  __host__ __device__ static constexpr bool ht(int i, int ip) { return (TM_ >> (3 * i + ip)) & 1; }
  __host__ __device__ static constexpr bool hc(int g, int h) { return (CM_ >> (2 * g + h)) & 1; }
  __host__ __device__ static constexpr int t(int i, int ip) { return duo_popc(TM_ & ((1u << (3 * i + ip)) - 1u)); }
  __host__ __device__ static constexpr int c(int g, int h) { return duo_popc(CM_ & ((1u << (2 * g + h)) - 1u)); }
  __host__ __device__ static constexpr int nr(int g, int h) { return NT + c(g, h); }
  __host__ __device__ static constexpr int nc(int g, int h) { return NT + NC + c(g, h); }
  __host__ __device__ static constexpr int send(int g) { return g == 0 ? S0_ : S1_; }
};
// The compiled patterns.  0: tridiagonal own block, full coupling, send rows (0, 2)  -- {U psi, W psi} of a 3-level
// ladder (q lane / p lane; CZ role {U|10>, W|10>}, the state-transfer model).  1: tridiagonal own block, anti-diagonal
// coupling, send rows (2, 1) -- a 6-cycle-free ladder of two chains (CZ role {U|11>} on its 6 orbit rows).
// 2: dense own block, full coupling (anything that splits at all).
typedef DuoPat<0x1BB, 0xF, 0, 2> DuoPat0;   // rows: {0,1} {0,1,2} {1,2}  ->  bits 0,1 | 3,4,5 | 7,8
typedef DuoPat<0x1BB, 0x6, 2, 1> DuoPat1;
typedef DuoPat<0x1FF, 0xF, 0, 2> DuoPat2;
typedef DuoPat<0x1BB, 0xF, 0, 2, 0x0AA> DuoPat3;   // patterns 0 / 1 without Xbar on the diagonal
typedef DuoPat<0x1BB, 0x6, 2, 1, 0x0AA> DuoPat4;

// compile-time loop: every index below is a constant expression, so the entry / accumulator arrays stay in registers
template <int I, int N, class F>
__device__ __forceinline__ void duo_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    duo_for<I + 1, N>(f);
  }
}
#define DUO_I(v) decltype(v)::value

// ---- arithmetic type of the sweeps: CT = double2 (complex128, the parity path) or float2 (the optional complex64 mode:
//      state, Taylor powers, entries, checkpoints and the Xbar accumulators in fp32; tables written in fp32 by
//      k_duo_tables; cost, terminal cotangent, boundary states, the sample sums and the reduction stay fp64)
template <typename CT> struct DuoReal;
template <> struct DuoReal<cplx> { typedef double R; };
template <> struct DuoReal<cplxf> { typedef float R; };
__device__ __forceinline__ cplx duo_mk(double re, double im, cplx*) { return cmake(re, im); }
__device__ __forceinline__ cplxf duo_mk(double re, double im, cplxf*) { return cmakef((float)re, (float)im); }
template <typename CT> __device__ __forceinline__ CT duo_from(const cplx v) { return duo_mk(v.x, v.y, (CT*)nullptr); }
__device__ __forceinline__ double duo_ifact(const DuoArgs& Q, int j, double*) { return Q.ifact[j]; }
__device__ __forceinline__ float duo_ifact(const DuoArgs& Q, int j, float*) { return Q.ifactf[j]; }
__device__ __forceinline__ cplx duo_to_d(const cplx v) { return v; }
__device__ __forceinline__ cplx duo_to_d(const cplxf v) { return cmake((double)v.x, (double)v.y); }
__device__ __forceinline__ float duo_xchg(float v) { return __shfl_xor_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ cplxf duo_xchg(cplxf v) { return cmakef(duo_xchg(v.x), duo_xchg(v.y)); }

// ---- small device helpers -----------------------------------------------------------------------------------------
__device__ __forceinline__ double duo_xchg(double v) {   // partner lane's value (32-bit halves: the CUDA double overload packs with `asm volatile`, which pins the order)
  const int lo = __shfl_xor_sync(0xffffffffu, __double2loint(v), 1);
  const int hi = __shfl_xor_sync(0xffffffffu, __double2hiint(v), 1);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ cplx duo_xchg(cplx v) { return cmake(duo_xchg(v.x), duo_xchg(v.y)); }
__device__ __forceinline__ double duo_xor(double v, int m) {
  const int lo = __shfl_xor_sync(0xffffffffu, __double2loint(v), m);
  const int hi = __shfl_xor_sync(0xffffffffu, __double2hiint(v), m);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ unsigned duo_s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void duo_bar_init(unsigned bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void duo_bar_expect(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void duo_bulk(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void duo_bar_wait(unsigned bar, unsigned parity) {
  unsigned done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

#ifndef DUO_HALO_SMEM
#define DUO_HALO_SMEM 0   // 1: the two send rows travel through a double-buffered shared-memory slot instead of shuffles
#endif
// h[g] = the partner lane's element S_g.  `xs` = this warp's exchange slab [2 buffers][2 rows][32 lanes], `buf` = step parity.
template <class PT, typename CT>
__device__ __forceinline__ void duo_halo(const CT (&x)[3], CT (&h)[2], CT* xs, const int buf, const int lane) {
#if DUO_HALO_SMEM
  CT* slot = xs + buf * 64;
  slot[lane] = x[PT::S0];
  slot[32 + lane] = x[PT::S1];
  __syncwarp();
  h[0] = slot[lane ^ 1];
  h[1] = slot[32 + (lane ^ 1)];
#else
  h[0] = duo_xchg(x[PT::S0]);
  h[1] = duo_xchg(x[PT::S1]);
#endif
}

// The loops below run COLUMN by column and issue, for one vector element, first every product with its real part and
// then every product with its imaginary part: consecutive DFMAs share a source register in the same operand slot, which
// ptxas turns into operand-reuse hits.  That matters because the FP64 pipe of a scheduler is fed one 64-bit register
// operand per cycle (profiles/micro/dfma_operands.cu: 3.6 cycles per DFMA with three fresh operands and one warp per
// scheduler, 2.8-2.9 with one operand reused).
template <bool FIRST, typename R>
__device__ __forceinline__ void duo_fma(R& acc, const R a, const R b) {
  if constexpr (FIRST) acc = a * b;
  else acc = fma(a, b, acc);
}

// y = X x on my rows: own block, then the coupling block on the partner's send rows (one exchange per step)
template <class PT, typename CT>
__device__ __forceinline__ void duo_mv(const CT (&e)[PT::NU], const CT (&x)[3], CT (&y)[3], CT* xs, const int buf,
                                       const int lane) {
  CT h[2];
  duo_halo<PT>(x, h, xs, buf, lane);
#pragma unroll
  for (int i = 0; i < 3; ++i) y[i] = duo_mk(0.0, 0.0, (CT*)nullptr);
  duo_for<0, 3>([&](auto ip_) {
    constexpr int ip = DUO_I(ip_);
    duo_for<0, 3>([&](auto i_) {
      constexpr int i = DUO_I(i_);
      if constexpr (PT::ht(i, ip)) {
        constexpr bool first = (PT::TM & ((1u << (3 * i + ip)) - 1u) & (7u << (3 * i))) == 0u;
        duo_fma<first>(y[i].x, e[PT::t(i, ip)].x, x[ip].x);
        duo_fma<first>(y[i].y, e[PT::t(i, ip)].y, x[ip].x);
      }
    });
    duo_for<0, 3>([&](auto i_) {
      constexpr int i = DUO_I(i_);
      if constexpr (PT::ht(i, ip)) {
        y[i].x = fma(-e[PT::t(i, ip)].y, x[ip].y, y[i].x);
        y[i].y = fma(e[PT::t(i, ip)].x, x[ip].y, y[i].y);
      }
    });
  });
  duo_for<0, 2>([&](auto hh_) {
    constexpr int hh = DUO_I(hh_);
    duo_for<0, 2>([&](auto g_) {
      constexpr int g = DUO_I(g_);
      if constexpr (PT::hc(g, hh)) {
        y[PT::send(g)].x = fma(e[PT::nr(g, hh)].x, h[hh].x, y[PT::send(g)].x);
        y[PT::send(g)].y = fma(e[PT::nr(g, hh)].y, h[hh].x, y[PT::send(g)].y);
      }
    });
    duo_for<0, 2>([&](auto g_) {
      constexpr int g = DUO_I(g_);
      if constexpr (PT::hc(g, hh)) {
        y[PT::send(g)].x = fma(-e[PT::nr(g, hh)].y, h[hh].y, y[PT::send(g)].x);
        y[PT::send(g)].y = fma(e[PT::nr(g, hh)].x, h[hh].y, y[PT::send(g)].y);
      }
    });
  });
}

// per-warp shared-memory carve-up (bytes), shared by host and device
struct DuoSmem {
  int tab, ck, bars, xs, con_a, us, total;
};
__host__ __device__ inline DuoSmem duo_smem_layout(int nu, int qcap, bool reverse, int esz) {   // esz = sizeof(CT)
  DuoSmem s;
  int o = 0;
  s.tab = o; o += 2 * DUO_BT * 2 * nu * esz;
  s.ck = o; o += reverse ? 2 * DUO_BT * 3 * 32 * esz : 0;
  s.bars = o; o += 16;
  s.xs = o; o += 2 * 2 * 32 * esz;                          // halo exchange slab (DUO_HALO_SMEM)
  s.con_a = o; o += nu * 32 * esz;
  const int powers = qcap * 96 * esz, staging = 14 * 33 * (int)sizeof(cplx);   // powers | Xbar staging (always fp64)
  s.us = o; o += reverse ? (powers > staging ? powers : staging) : 0;
  s.total = (o + 15) / 16 * 16;
  return s;
}

// What a warp of role `role` works on: sample group g, slice range [k0, k1), basis column (mode 1)
struct DuoItem {
  int g, c, col, k0, k1;
};
__device__ __forceinline__ DuoItem duo_item(const DuoArgs& Q, const int role, const int it) {
  DuoItem I;
  I.col = 0; I.c = 0; I.g = it; I.k0 = 0; I.k1 = Q.M;
  if (Q.mode == 1) {
    const int nc = (Q.ncol[role] + Q.cpi - 1) / Q.cpi;    // column groups of the role
    I.col = (it % nc) * Q.cpi; I.c = (it / nc) % Q.chunks; I.g = it / (nc * Q.chunks);
  } else if (Q.mode == 2) {
    I.c = it % Q.chunks; I.g = it / Q.chunks;
  }
  if (Q.mode != 0) { I.k0 = I.c * Q.lc; I.k1 = min(Q.M, I.k0 + Q.lc); }
  return I;
}

// =====================================================================================================================
// Forward sweep of one warp: 16 samples x 2 lanes of role `role`.
// =====================================================================================================================
// (Tried: the next slice's table values prefetched into a second register set -- 218 registers, 2.39 instead of 2.12 ms.)
template <class PT, typename CT>
__device__ __forceinline__ void duo_forward(const DuoArgs& Q, const int role, unsigned char* smem_raw) {
  constexpr int NU = PT::NU;
  typedef typename DuoReal<CT>::R R;
  const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // uniform for the compiler
  const int type = lane & 1, sl = lane >> 1;
  const int wg = (blockIdx.x / Q.nroles) * Q.wpc + warp;          // warp of this role
  if (wg >= Q.warps[role]) return;
  const DuoItem I = duo_item(Q, role, wg);
  const int sample = I.g * DUO_SPW + sl;
  const bool live = sample < Q.S;
  const int sm = live ? sample : Q.S - 1;
  const R m1 = (R)(Q.ncls > 1 ? Q.mcls[(size_t)sm * Q.ncls + 1] : 1.0);
  const DuoSmem L = duo_smem_layout(NU, Q.qcap, false, (int)sizeof(CT));
  unsigned char* base = smem_raw + (size_t)warp * L.total;
  const CT* s_tab = (const CT*)(base + L.tab);
  CT* s_a = (CT*)(base + L.con_a) + lane;
  const unsigned cmask = Q.clsmask[role * 2 + type], amask = Q.amask[role * 2 + type];
  CT* s_x = (CT*)(base + L.xs);
  const unsigned bar0 = duo_s32(base + L.bars), tab0 = duo_s32(base + L.tab);
  const unsigned row_bytes = 2u * NU * sizeof(CT);              // both lane types of one slice
  const CT* gtab = (const CT*)Q.tab + Q.tab_off[role];
  const int k0 = I.k0, nsl = I.k1 - I.k0;                         // my slices: k0 .. k0 + nsl - 1
  const int nb = (nsl + DUO_BT - 1) / DUO_BT;
  auto issue = [&](int b) {
    const int cnt = min(DUO_BT, nsl - b * DUO_BT);
    const unsigned bar = bar0 + (b & 1) * 8, bytes = cnt * row_bytes;
    duo_bar_expect(bar, bytes);
    duo_bulk(tab0 + (b & 1) * DUO_BT * row_bytes, gtab + (size_t)(k0 + b * DUO_BT) * 2 * NU, bytes, bar);
  };
  if (lane == 0) {
    duo_bar_init(bar0);
    duo_bar_init(bar0 + 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    issue(0);
    if (nb > 1) issue(1);
  }
  // per-sample constants of my entries: class multiplier and additive-noise part
  {
    const cplx* ga = Q.con_a + Q.con_off[role] + (size_t)type * NU * Q.S + sm;
#pragma unroll
    for (int u = 0; u < NU; ++u)
      if ((amask >> u) & 1u) s_a[u * 32] = duo_from<CT>(ga[(size_t)u * Q.S]);
  }
  const CT czero = duo_mk(0.0, 0.0, (CT*)nullptr);
  int srow[3];
  CT z[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    srow[i] = Q.srow[(role * 2 + type) * 3 + i];
    if (Q.mode == 0) z[i] = srow[i] >= 0 ? duo_from<CT>(Q.starts[srow[i]]) : czero;
    else if (Q.mode == 1) z[i] = duo_mk(Q.colrow[role * 6 + I.col] == type * 3 + i ? 1.0 : 0.0, 0.0, (CT*)nullptr);
    else z[i] = srow[i] >= 0 ? duo_from<CT>(Q.zstart[((size_t)sm * Q.chunks + I.c) * Q.N + srow[i]]) : czero;
  }
  __syncwarp();
  const bool keep = Q.want_grad && Q.mode != 1;
  CT* ck = (CT*)Q.ckpt + ((size_t)(Q.warp0[role] + I.g) * Q.M) * 96 + lane;
  const int* qd = Q.qdeg + (size_t)role * Q.qstride;
  int q_next = __ldg(qd + k0);
  for (int b = 0; b < nb; ++b) {
    duo_bar_wait(bar0 + (b & 1) * 8, (b >> 1) & 1);
    const int cnt = min(DUO_BT, nsl - b * DUO_BT);
    for (int kk = 0; kk < cnt; ++kk) {
      const int k = k0 + b * DUO_BT + kk;
      const int q = q_next;
      if (k + 1 < I.k1) q_next = __ldg(qd + k + 1);
      const CT* row = s_tab + ((size_t)((b & 1) * DUO_BT + kk) * 2 + type) * NU;
      CT e[NU];                                            // entry = A_u(s) + m_u(s) P_u(k); m_u is 1 or the sample's class-1
#pragma unroll                                              // multiplier, A_u exists on few entries (additive noise)
      for (int u = 0; u < NU; ++u) {
        const CT pv = row[u];
        const R mv = ((cmask >> u) & 1u) ? m1 : (R)1;
        const CT av = ((amask >> u) & 1u) ? s_a[u * 32] : czero;
        e[u].x = fma(mv, pv.x, av.x); e[u].y = fma(mv, pv.y, av.y);
      }
      if (keep) {
#pragma unroll
        for (int i = 0; i < 3; ++i) ck[((size_t)k * 3 + i) * 32] = z[i];
      }
      CT x[3] = {z[0], z[1], z[2]};
#pragma unroll kDuoUnroll
      for (int j = 1; j <= q; ++j) {
        CT y[3];
        duo_mv<PT>(e, x, y, s_x, j & 1, lane);
        const R cj = duo_ifact(Q, j, (R*)nullptr);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          x[i] = y[i];
          z[i].x = fma(cj, y[i].x, z[i].x);
          z[i].y = fma(cj, y[i].y, z[i].y);
        }
      }
    }
    __syncwarp();                                          // every lane is done with this stage
    if (lane == 0 && b + 2 < nb) issue(b + 2);
  }
  if (live && Q.mode == 0) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
      if (srow[i] >= 0) Q.zfin[(size_t)sample * Q.N + srow[i]] = duo_to_d(z[i]);
  }
  if (live && Q.mode == 1) {                               // column I.col of the chunk propagator of my role
    cplx* pc = Q.prop + ((((size_t)sample * Q.chunks + I.c) * Q.nroles + role) * 6 + I.col) * 6 + type * 3;
#pragma unroll
    for (int i = 0; i < 3; ++i) pc[i] = duo_to_d(z[i]);
  }
}

// =====================================================================================================================
// Chunk propagators (mode 1), NC basis columns per item: the columns share the matrix entries, so the mat-vec issues NC
// consecutive DFMAs per entry component with the SAME entry register (operand-reuse hits: 2.3-2.9 instead of 3.0-3.6
// cycles per DFMA), and the per-slice entry assembly is paid once for NC columns.
// =====================================================================================================================
#define DUO_COLS 3
template <class PT, typename CT, int NC>
__device__ __forceinline__ void duo_forward_cols(const DuoArgs& Q, const int role, unsigned char* smem_raw) {
  constexpr int NU = PT::NU;
  typedef typename DuoReal<CT>::R R;
  const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int type = lane & 1, sl = lane >> 1;
  const int wg = (blockIdx.x / Q.nroles) * Q.wpc + warp;
  if (wg >= Q.warps[role]) return;
  const DuoItem I = duo_item(Q, role, wg);
  const int sample = I.g * DUO_SPW + sl;
  const bool live = sample < Q.S;
  const int sm = live ? sample : Q.S - 1;
  const R m1 = (R)(Q.ncls > 1 ? Q.mcls[(size_t)sm * Q.ncls + 1] : 1.0);
  const DuoSmem L = duo_smem_layout(NU, Q.qcap, false, (int)sizeof(CT));
  unsigned char* base = smem_raw + (size_t)warp * L.total;
  const CT* s_tab = (const CT*)(base + L.tab);
  CT* s_a = (CT*)(base + L.con_a) + lane;
  const unsigned cmask = Q.clsmask[role * 2 + type], amask = Q.amask[role * 2 + type];
  const unsigned bar0 = duo_s32(base + L.bars), tab0 = duo_s32(base + L.tab);
  const unsigned row_bytes = 2u * NU * sizeof(CT);
  const CT* gtab = (const CT*)Q.tab + Q.tab_off[role];
  const int k0 = I.k0, nsl = I.k1 - I.k0;
  const int nb = (nsl + DUO_BT - 1) / DUO_BT;
  auto issue = [&](int b) {
    const int cnt = min(DUO_BT, nsl - b * DUO_BT);
    const unsigned bar = bar0 + (b & 1) * 8, bytes = cnt * row_bytes;
    duo_bar_expect(bar, bytes);
    duo_bulk(tab0 + (b & 1) * DUO_BT * row_bytes, gtab + (size_t)(k0 + b * DUO_BT) * 2 * NU, bytes, bar);
  };
  if (lane == 0) {
    duo_bar_init(bar0);
    duo_bar_init(bar0 + 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    issue(0);
    if (nb > 1) issue(1);
  }
  {
    const cplx* ga = Q.con_a + Q.con_off[role] + (size_t)type * NU * Q.S + sm;
#pragma unroll
    for (int u = 0; u < NU; ++u)
      if ((amask >> u) & 1u) s_a[u * 32] = duo_from<CT>(ga[(size_t)u * Q.S]);
  }
  const CT czero = duo_mk(0.0, 0.0, (CT*)nullptr);
  CT z[NC][3];
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const bool one = I.col + c < Q.ncol[role] && Q.colrow[role * 6 + min(I.col + c, 5)] == type * 3 + i;
      z[c][i] = duo_mk(one ? 1.0 : 0.0, 0.0, (CT*)nullptr);
    }
  __syncwarp();
  const int* qd = Q.qdeg + (size_t)role * Q.qstride;
  int q_next = __ldg(qd + k0);
  for (int b = 0; b < nb; ++b) {
    duo_bar_wait(bar0 + (b & 1) * 8, (b >> 1) & 1);
    const int cnt = min(DUO_BT, nsl - b * DUO_BT);
    for (int kk = 0; kk < cnt; ++kk) {
      const int k = k0 + b * DUO_BT + kk;
      const int q = q_next;
      if (k + 1 < I.k1) q_next = __ldg(qd + k + 1);
      const CT* row = s_tab + ((size_t)((b & 1) * DUO_BT + kk) * 2 + type) * NU;
      CT e[NU];
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const CT pv = row[u];
        const R mv = ((cmask >> u) & 1u) ? m1 : (R)1;
        const CT av = ((amask >> u) & 1u) ? s_a[u * 32] : czero;
        e[u].x = fma(mv, pv.x, av.x); e[u].y = fma(mv, pv.y, av.y);
      }
      CT x[NC][3];
#pragma unroll
      for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int i = 0; i < 3; ++i) x[c][i] = z[c][i];
      for (int j = 1; j <= q; ++j) {
        CT h[NC][2], y[NC][3];
#pragma unroll
        for (int c = 0; c < NC; ++c) { h[c][0] = duo_xchg(x[c][PT::S0]); h[c][1] = duo_xchg(x[c][PT::S1]); }
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
          for (int i = 0; i < 3; ++i) y[c][i] = czero;
        duo_for<0, 3>([&](auto ip_) {
          constexpr int ip = DUO_I(ip_);
          duo_for<0, 3>([&](auto i_) {
            constexpr int i = DUO_I(i_);
            if constexpr (PT::ht(i, ip)) {
              constexpr bool first = (PT::TM & ((1u << (3 * i + ip)) - 1u) & (7u << (3 * i))) == 0u;
              const CT ev = e[PT::t(i, ip)];
#pragma unroll
              for (int c = 0; c < NC; ++c) duo_fma<first>(y[c][i].x, ev.x, x[c][ip].x);
#pragma unroll
              for (int c = 0; c < NC; ++c) duo_fma<first>(y[c][i].y, ev.y, x[c][ip].x);
#pragma unroll
              for (int c = 0; c < NC; ++c) y[c][i].x = fma(-ev.y, x[c][ip].y, y[c][i].x);
#pragma unroll
              for (int c = 0; c < NC; ++c) y[c][i].y = fma(ev.x, x[c][ip].y, y[c][i].y);
            }
          });
        });
        duo_for<0, 4>([&](auto gh_) {
          constexpr int g = DUO_I(gh_) >> 1, hh = DUO_I(gh_) & 1;
          if constexpr (PT::hc(g, hh)) {
            constexpr int i = PT::send(g);
            const CT ev = e[PT::nr(g, hh)];
#pragma unroll
            for (int c = 0; c < NC; ++c) y[c][i].x = fma(ev.x, h[c][hh].x, y[c][i].x);
#pragma unroll
            for (int c = 0; c < NC; ++c) y[c][i].y = fma(ev.y, h[c][hh].x, y[c][i].y);
#pragma unroll
            for (int c = 0; c < NC; ++c) y[c][i].x = fma(-ev.y, h[c][hh].y, y[c][i].x);
#pragma unroll
            for (int c = 0; c < NC; ++c) y[c][i].y = fma(ev.x, h[c][hh].y, y[c][i].y);
          }
        });
        const R cj = duo_ifact(Q, j, (R*)nullptr);
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            x[c][i] = y[c][i];
            z[c][i].x = fma(cj, y[c][i].x, z[c][i].x);
            z[c][i].y = fma(cj, y[c][i].y, z[c][i].y);
          }
      }
    }
    __syncwarp();
    if (lane == 0 && b + 2 < nb) issue(b + 2);
  }
  if (live) {
#pragma unroll
    for (int c = 0; c < NC; ++c)
      if (I.col + c < Q.ncol[role]) {
        cplx* pc = Q.prop + ((((size_t)sample * Q.chunks + I.c) * Q.nroles + role) * 6 + I.col + c) * 6 + type * 3;
#pragma unroll
        for (int i = 0; i < 3; ++i) pc[i] = duo_to_d(z[c][i]);
      }
  }
}

// =====================================================================================================================
// Reverse sweep of one warp: recompute the Taylor powers of slice k from its checkpoint, run the adjoint recurrence,
// accumulate Xbar on my entries, emit the warp's sums.
// =====================================================================================================================
template <class PT, typename CT>
__device__ __forceinline__ void duo_reverse(const DuoArgs& Q, const int role, unsigned char* smem_raw) {
  constexpr int NU = PT::NU, NX = PT::NX;
  typedef typename DuoReal<CT>::R R;
  const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int type = lane & 1, sl = lane >> 1;
  const int wg = (blockIdx.x / Q.nroles) * Q.wpc + warp;
  if (wg >= Q.warps[role]) return;
  const DuoItem I = duo_item(Q, role, wg);
  const int sample = I.g * DUO_SPW + sl;
  const bool live = sample < Q.S;
  const int sm = live ? sample : Q.S - 1;
  const R m1 = (R)(Q.ncls > 1 ? Q.mcls[(size_t)sm * Q.ncls + 1] : 1.0);
  const DuoSmem L = duo_smem_layout(NU, Q.qcap, true, (int)sizeof(CT));
  unsigned char* base = smem_raw + (size_t)warp * L.total;
  const CT* s_tab = (const CT*)(base + L.tab);
  const CT* s_ck = (const CT*)(base + L.ck) + lane;
  CT* s_a = (CT*)(base + L.con_a) + lane;
  const unsigned cmask = Q.clsmask[role * 2 + type], amask = Q.amask[role * 2 + type];
  CT* s_x = (CT*)(base + L.xs);
  CT* s_u = (CT*)(base + L.us) + lane;
  const CT czero = duo_mk(0.0, 0.0, (CT*)nullptr);
  const unsigned bar0 = duo_s32(base + L.bars), tab0 = duo_s32(base + L.tab), ck0 = duo_s32(base + L.ck);
  const unsigned row_bytes = 2u * NU * sizeof(CT), ck_bytes = 96u * sizeof(CT);
  const CT* gtab = (const CT*)Q.tab + Q.tab_off[role];
  const CT* gck = (const CT*)Q.ckpt + ((size_t)(Q.warp0[role] + I.g) * Q.M) * 96;
  const int k0 = I.k0, nsl = I.k1 - I.k0;
  const int nb = (nsl + DUO_BT - 1) / DUO_BT;
  // batch b covers slices [b BT, b BT + cnt); the sweep visits the batches downwards: visit v = nb - 1 - b
  auto issue = [&](int b) {
    const int v = nb - 1 - b, cnt = min(DUO_BT, nsl - b * DUO_BT);
    const unsigned bar = bar0 + (v & 1) * 8;
    duo_bar_expect(bar, cnt * (row_bytes + ck_bytes));
    duo_bulk(tab0 + (v & 1) * DUO_BT * row_bytes, gtab + (size_t)(k0 + b * DUO_BT) * 2 * NU, cnt * row_bytes, bar);
    duo_bulk(ck0 + (v & 1) * DUO_BT * ck_bytes, gck + (size_t)(k0 + b * DUO_BT) * 96, cnt * ck_bytes, bar);
  };
  if (lane == 0) {
    duo_bar_init(bar0);
    duo_bar_init(bar0 + 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    issue(nb - 1);
    if (nb > 1) issue(nb - 2);
  }
  {
    const cplx* ga = Q.con_a + Q.con_off[role] + (size_t)type * NU * Q.S + sm;
#pragma unroll
    for (int u = 0; u < NU; ++u)
      if ((amask >> u) & 1u) s_a[u * 32] = duo_from<CT>(ga[(size_t)u * Q.S]);
  }
  CT w[3];                                                 // running cotangent of my rows
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int sr = Q.srow[(role * 2 + type) * 3 + i];
    const cplx* lsrc = Q.mode == 2 ? Q.lamend + ((size_t)sm * Q.chunks + I.c) * Q.N : Q.lam + (size_t)sm * Q.N;
    w[i] = (sr >= 0 && live) ? duo_from<CT>(lsrc[sr]) : czero;
  }
  __syncwarp();
  cplx* yout = Q.ypart + Q.y_off[role] + ((size_t)I.g * Q.M) * 2 * NX;
  cplx* s_e = (cplx*)(base + L.us);                       // [NX][33] staging of the per-sample Xbar entries
  const int* qd = Q.qdeg + (size_t)role * Q.qstride;
  int q_next = __ldg(qd + I.k1 - 1);
  // The table values and the checkpoint of the NEXT slice to be visited are loaded right after the adjoint loop of the
  // current one -- into the entry registers, which are dead by then -- so that they arrive behind the Xbar staging
  // instead of in front of the recomputation.
  CT e[NU], xk[3];
  auto prefetch = [&](int vv, int kk2) {
    const CT* row = s_tab + ((size_t)((vv & 1) * DUO_BT + kk2) * 2 + type) * NU;
    const CT* zk = s_ck + (size_t)((vv & 1) * DUO_BT + kk2) * 96;
#pragma unroll
    for (int u = 0; u < NU; ++u) e[u] = row[u];
    xk[0] = zk[0]; xk[1] = zk[32]; xk[2] = zk[64];
  };
  duo_bar_wait(bar0, 0);
  prefetch(0, min(DUO_BT, nsl - (nb - 1) * DUO_BT) - 1);
  for (int v = 0; v < nb; ++v) {
    const int b = nb - 1 - v;
    const int cnt = min(DUO_BT, nsl - b * DUO_BT);
    for (int kk = cnt - 1; kk >= 0; --kk) {
      const int k = k0 + b * DUO_BT + kk;
      const int q = q_next;
      if (k > k0) q_next = __ldg(qd + k - 1);
#pragma unroll                                              // entry = A_u(s) + m_u(s) P_u(k); m_u is 1 or the sample's class-1
      for (int u = 0; u < NU; ++u) {                       // multiplier, A_u exists on few entries (additive noise)
        const CT pv = e[u];
        const R mv = ((cmask >> u) & 1u) ? m1 : (R)1;
        const CT av = ((amask >> u) & 1u) ? s_a[u * 32] : czero;
        e[u].x = fma(mv, pv.x, av.x); e[u].y = fma(mv, pv.y, av.y);
      }
      // ---- recompute u_0 .. u_{q-1}
      {
        CT x[3] = {xk[0], xk[1], xk[2]};
#pragma unroll
        for (int i = 0; i < 3; ++i) s_u[i * 32] = x[i];
#pragma unroll kDuoUnroll
        for (int j = 1; j < q; ++j) {
          CT y[3];
          duo_mv<PT>(e, x, y, s_x, j & 1, lane);
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            x[i] = y[i];
            s_u[(j * 3 + i) * 32] = y[i];
          }
        }
      }
      // ---- adjoint recurrence, descending; Xbar on my own block and on the coupling block I own by column
      CT lamk[3], xb[NX];
      {
        const R cq = duo_ifact(Q, q, (R*)nullptr);
#pragma unroll
        for (int i = 0; i < 3; ++i) { lamk[i] = w[i]; w[i] = cscale(w[i], cq); }
      }
#pragma unroll
      for (int x = 0; x < NX; ++x) xb[x] = czero;
      CT u[3] = {s_u[((q - 1) * 3 + 0) * 32], s_u[((q - 1) * 3 + 1) * 32], s_u[((q - 1) * 3 + 2) * 32]};
#pragma unroll kDuoUnroll
      for (int j = q; j >= 1; --j) {
        CT wh[2];
        duo_halo<PT>(w, wh, s_x, j & 1, lane);
        CT un[3];                                          // next step's powers: in flight during this step
        {
          const int jn = j >= 2 ? j - 2 : 0;
#pragma unroll
          for (int i = 0; i < 3; ++i) un[i] = s_u[(jn * 3 + i) * 32];
        }
        const R cj = duo_ifact(Q, j - 1, (R*)nullptr);
        CT wn[3];                                          // w_{j-1} = c_{j-1} lam + X^dag w_j  (conjugated entries)
#pragma unroll
        for (int i = 0; i < 3; ++i) wn[i] = cscale(lamk[i], cj);
        // own block, source element w[ip] at a time: X^dag w (column i of row ip) and Xbar (row ip: conj(w_ip) u_i)
        duo_for<0, 3>([&](auto ip_) {
          constexpr int ip = DUO_I(ip_);
          duo_for<0, 3>([&](auto i_) {           // products with Re w[ip]
            constexpr int i = DUO_I(i_);
            if constexpr (PT::ht(ip, i)) {
              wn[i].x = fma(e[PT::t(ip, i)].x, w[ip].x, wn[i].x);
              wn[i].y = fma(-e[PT::t(ip, i)].y, w[ip].x, wn[i].y);
              if constexpr (PT::hx(ip, i)) {
                xb[PT::xt(ip, i)].x = fma(w[ip].x, u[i].x, xb[PT::xt(ip, i)].x);
                xb[PT::xt(ip, i)].y = fma(w[ip].x, u[i].y, xb[PT::xt(ip, i)].y);
              }
            }
          });
          duo_for<0, 3>([&](auto i_) {           // products with Im w[ip]
            constexpr int i = DUO_I(i_);
            if constexpr (PT::ht(ip, i)) {
              wn[i].x = fma(e[PT::t(ip, i)].y, w[ip].y, wn[i].x);
              wn[i].y = fma(e[PT::t(ip, i)].x, w[ip].y, wn[i].y);
              if constexpr (PT::hx(ip, i)) {
                xb[PT::xt(ip, i)].x = fma(w[ip].y, u[i].y, xb[PT::xt(ip, i)].x);
                xb[PT::xt(ip, i)].y = fma(-w[ip].y, u[i].x, xb[PT::xt(ip, i)].y);
              }
            }
          });
        });
        // coupling block I own by column: rows = partner's send rows g (halo wh[g]), columns = my send rows hh
        duo_for<0, 2>([&](auto g_) {
          constexpr int g = DUO_I(g_);
          duo_for<0, 2>([&](auto hh_) {
            constexpr int hh = DUO_I(hh_);
            if constexpr (PT::hc(g, hh)) {
              constexpr int i = PT::send(hh);
              wn[i].x = fma(e[PT::nc(g, hh)].x, wh[g].x, wn[i].x);
              wn[i].y = fma(-e[PT::nc(g, hh)].y, wh[g].x, wn[i].y);
              xb[PT::NXT + PT::c(g, hh)].x = fma(wh[g].x, u[i].x, xb[PT::NXT + PT::c(g, hh)].x);
              xb[PT::NXT + PT::c(g, hh)].y = fma(wh[g].x, u[i].y, xb[PT::NXT + PT::c(g, hh)].y);
            }
          });
          duo_for<0, 2>([&](auto hh_) {
            constexpr int hh = DUO_I(hh_);
            if constexpr (PT::hc(g, hh)) {
              constexpr int i = PT::send(hh);
              wn[i].x = fma(e[PT::nc(g, hh)].y, wh[g].y, wn[i].x);
              wn[i].y = fma(e[PT::nc(g, hh)].x, wh[g].y, wn[i].y);
              xb[PT::NXT + PT::c(g, hh)].x = fma(wh[g].y, u[i].y, xb[PT::NXT + PT::c(g, hh)].x);
              xb[PT::NXT + PT::c(g, hh)].y = fma(-wh[g].y, u[i].x, xb[PT::NXT + PT::c(g, hh)].y);
            }
          });
        });
#pragma unroll
        for (int i = 0; i < 3; ++i) { w[i] = wn[i]; u[i] = un[i]; }
      }
      // ---- next slice's table values and checkpoint (see above)
      if (kk > 0) prefetch(v, kk - 1);
      else if (v + 1 < nb) {
        duo_bar_wait(bar0 + ((v + 1) & 1) * 8, ((v + 1) >> 1) & 1);
        prefetch(v + 1, min(DUO_BT, nsl - (b - 1) * DUO_BT) - 1);
      }
      // ---- Ybar_x = sum over the warp's samples of m_s * Xbar_x, through shared memory (the slab of the Taylor powers
      //      is free now): a shuffle tree would cost 16 SHFL per number, and an SM moves only ~0.5-0.75 SHFL per clock
      //      (profiles/micro/latency.cu) -- it was 10 % of the reverse sweep.  Fixed summation order: deterministic.
      duo_for<0, NX>([&](auto x_) {                       // (all multiplier loads before the first store: the
        constexpr int x = DUO_I(x_);                       //  compiler cannot reorder them across possibly aliasing stores)
        xb[x] = cscale(xb[x], ((cmask >> PT::ux(x)) & 1u) ? m1 : (R)1);
      });
      __syncwarp();                                        // every lane is done with its powers
#pragma unroll
      for (int x = 0; x < NX; ++x) s_e[x * 33 + lane] = duo_to_d(xb[x]);
      __syncwarp();
      if (lane < 2 * NX) {                                 // lane = 2 x + lane type: sum its 16 samples
        const cplx* src = s_e + (lane >> 1) * 33 + (lane & 1);
        cplx p0 = src[0], p1 = src[2], p2 = src[4], p3 = src[6];
#pragma unroll
        for (int sl2 = 4; sl2 < 16; sl2 += 4) {
          p0 = cadd(p0, src[2 * sl2]); p1 = cadd(p1, src[2 * sl2 + 2]);
          p2 = cadd(p2, src[2 * sl2 + 4]); p3 = cadd(p3, src[2 * sl2 + 6]);
        }
        yout[((size_t)k * 2 + (lane & 1)) * NX + (lane >> 1)] = cadd(cadd(p0, p1), cadd(p2, p3));
      }
      __syncwarp();                                        // before the next slice's powers overwrite the slab
    }
    __syncwarp();
    if (lane == 0 && v + 2 < nb) issue(nb - 1 - (v + 2));
  }
}

typedef void (*duo_fn)(const DuoArgs);
duo_fn qocvl_duo_forward_kernel();
duo_fn qocvl_duo_reverse_kernel();
