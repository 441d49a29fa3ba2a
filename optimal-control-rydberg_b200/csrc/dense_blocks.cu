// dense_blocks.cu -- the dense Van Loan propagators of the reference API for n > 16 (blockade chains, n = 144 / 377):
//   exponentials(): ST:247-256 / TQ:460-484   [M][2n][2n] = expm(dt * (drift + sum_j aux_kj G_j))
//   propagate():    ST:258-283 / TQ:489-535   pairwise product tree, later slice on the LEFT, odd leftover folded in
// (dense.cu keeps the shared-memory path for the reference's own models, 2n <= 32.)
//
// Two things make the dense regime affordable here (north_star (2)):
//   * BLOCK-TRIANGULAR ARITHMETIC.  Every matrix of the path has the form [[u, w], [0, u]] (generators, Taylor terms,
//     exponentials, products): it is stored as the pair (u, w) of n x n blocks and multiplied as
//         (u2, w2) (u1, w1) = (u2 u1, u2 w1 + w2 u1)          3 n x n products instead of the 8 of a dense 2n x 2n one.
//   * FP64 TENSOR CORES.  The n x n complex products run on DMMA (mma.sync.m8n8k4.f64): k_zgemm_dmma is a batched,
//     strided complex GEMM C = alpha A B + beta C (+ optional accumulation of the result into a second matrix), 32 x 32
//     complex output tile per CTA, 4 warps x (2 x 2) m8n8 tiles, real and imaginary parts as separate shared-memory
//     planes, four real DMMAs per complex tile product (re += ar br, re += (-ai) bi, im += ar bi, im += ai br).
// The exponential is scaling-and-squaring around a degree-16 Taylor polynomial (|X / 2^s|_1 <= 1/2, s taken as the
// maximum over the slices of a batch so that the whole batch runs the same GEMM sequence), on slice batches of
// DENSE_BATCH so that the workspace stays bounded.
#include "common.cuh"
#include <algorithm>

#define ZG_BM 32
#define ZG_BN 32
#define ZG_BK 8
#define DENSE_BATCH 256

__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, const double a, const double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// C[z] = alpha * A[z] * B[z] + beta * C[z];  if acc != nullptr additionally acc[z] += C[z] (new value).
// All matrices n x n complex, row-major, leading dimension n; batch z at element offsets z * strideX.
__global__ void __launch_bounds__(128) k_zgemm_dmma(int n, const cplx* __restrict__ A, size_t strideA,
                                                    const cplx* __restrict__ B, size_t strideB, cplx* __restrict__ Cm,
                                                    size_t strideC, double alpha, double beta, cplx* __restrict__ acc,
                                                    size_t strideAcc) {
  __shared__ double As_re[ZG_BM][ZG_BK + 1], As_im[ZG_BM][ZG_BK + 1];
  __shared__ double Bs_re[ZG_BK][ZG_BN + 1], Bs_im[ZG_BK][ZG_BN + 1];
  const int z = blockIdx.z, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = blockIdx.y * ZG_BM, col0 = blockIdx.x * ZG_BN;
  const int wm = (warp >> 1) * 16, wn = (warp & 1) * 16;        // this warp's 16 x 16 corner inside the CTA tile
  A += (size_t)z * strideA; B += (size_t)z * strideB; Cm += (size_t)z * strideC;
  double cre[2][2][2], cim[2][2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) { cre[i][j][0] = cre[i][j][1] = cim[i][j][0] = cim[i][j][1] = 0.0; }
  const int ar = lane >> 2, ac = lane & 3;                       // fragment coordinates (A: row, k; B: k, col)
  for (int k0 = 0; k0 < n; k0 += ZG_BK) {
    for (int e = tid; e < ZG_BM * ZG_BK; e += 128) {             // A tile: 32 x 8
      const int r = e / ZG_BK, c = e % ZG_BK;
      const bool ok = row0 + r < n && k0 + c < n;
      const cplx v = ok ? A[(size_t)(row0 + r) * n + k0 + c] : cmake(0.0, 0.0);
      As_re[r][c] = v.x; As_im[r][c] = v.y;
    }
    for (int e = tid; e < ZG_BK * ZG_BN; e += 128) {             // B tile: 8 x 32
      const int r = e / ZG_BN, c = e % ZG_BN;
      const bool ok = k0 + r < n && col0 + c < n;
      const cplx v = ok ? B[(size_t)(k0 + r) * n + col0 + c] : cmake(0.0, 0.0);
      Bs_re[r][c] = v.x; Bs_im[r][c] = v.y;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < ZG_BK; kk += 4) {
      double a_re[2], a_im[2], b_re[2], b_im[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) { a_re[i] = As_re[wm + 8 * i + ar][kk + ac]; a_im[i] = As_im[wm + 8 * i + ar][kk + ac]; }
#pragma unroll
      for (int j = 0; j < 2; ++j) { b_re[j] = Bs_re[kk + ac][wn + 8 * j + ar]; b_im[j] = Bs_im[kk + ac][wn + 8 * j + ar]; }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          dmma_8x8x4(cre[i][j][0], cre[i][j][1], a_re[i], b_re[j]);
          dmma_8x8x4(cre[i][j][0], cre[i][j][1], -a_im[i], b_im[j]);
          dmma_8x8x4(cim[i][j][0], cim[i][j][1], a_re[i], b_im[j]);
          dmma_8x8x4(cim[i][j][0], cim[i][j][1], a_im[i], b_re[j]);
        }
    }
    __syncthreads();
  }
  // epilogue: thread holds C[row = lane / 4][col = 2 (lane % 4) + {0, 1}] of every 8 x 8 tile
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = row0 + wm + 8 * i + (lane >> 2), c = col0 + wn + 8 * j + 2 * (lane & 3) + h;
        if (r < n && c < n) {
          const size_t o = (size_t)r * n + c;
          cplx v = cmake(alpha * cre[i][j][h], alpha * cim[i][j][h]);
          if (beta != 0.0) { const cplx old = Cm[o]; v.x = fma(beta, old.x, v.x); v.y = fma(beta, old.y, v.y); }
          Cm[o] = v;
          if (acc) { cplx* a2 = acc + (size_t)z * strideAcc + o; *a2 = cadd(*a2, v); }
        }
      }
}

// X_k blocks of a batch of slices: xu[z] = dt sum_j coef_kj Gu_j, xw[z] = dt sum_j coef_kj Gw_j (k = k0 + z), and the
// per-slice 1-norm of [[xu, xw], [0, xu]] (max column sum) into nrm[z].  One CTA per slice.
__global__ void __launch_bounds__(256) k_block_generators(int n, int J, double dt, int k0, const cplx* __restrict__ gu,
                                                          const cplx* __restrict__ gw, const cplx* __restrict__ coef,
                                                          cplx* __restrict__ xu, cplx* __restrict__ xw,
                                                          double* __restrict__ nrm) {
  __shared__ double red[256];
  const int z = blockIdx.x, k = k0 + z;
  const size_t nn = (size_t)n * n;
  for (size_t o = threadIdx.x; o < nn; o += blockDim.x) {
    cplx u = cmake(0, 0), w = cmake(0, 0);
    for (int j = 0; j < J; ++j) {
      const cplx cj = coef[(size_t)k * J + j];
      u = cfma(cj, gu[(size_t)j * nn + o], u);
      w = cfma(cj, gw[(size_t)j * nn + o], w);
    }
    xu[(size_t)z * nn + o] = cscale(u, dt);
    xw[(size_t)z * nn + o] = cscale(w, dt);
  }
  __syncthreads();
  double best = 0.0;
  for (int c = threadIdx.x; c < n; c += blockDim.x) {            // column c of the right half: |xw| + |xu| (>= left half)
    double s = 0.0;
    for (int r = 0; r < n; ++r) {
      const cplx u = xu[(size_t)z * nn + (size_t)r * n + c], w = xw[(size_t)z * nn + (size_t)r * n + c];
      s += hypot(u.x, u.y) + hypot(w.x, w.y);
    }
    best = fmax(best, s);
  }
  red[threadIdx.x] = best;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + off]);
    __syncthreads();
  }
  if (threadIdx.x == 0) nrm[z] = red[0];
}

// t = x * 2^-s ; r = I + t (u block) / t (w block)
__global__ void k_block_scale_init(size_t nn, int n, int count, double sc, cplx* __restrict__ xu, cplx* __restrict__ xw,
                                   cplx* __restrict__ tu, cplx* __restrict__ tw, cplx* __restrict__ ru,
                                   cplx* __restrict__ rw) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn * count) return;
  const size_t o = i % nn;
  const cplx u = cscale(xu[i], sc), w = cscale(xw[i], sc);
  xu[i] = u; xw[i] = w; tu[i] = u; tw[i] = w;
  ru[i] = cadd(u, (o / n == o % n) ? cmake(1, 0) : cmake(0, 0));
  rw[i] = w;
}

// dense [count][2n][2n] = [[u, w], [0, u]]
__global__ void k_block_assemble(int n, int count, const cplx* __restrict__ u, const cplx* __restrict__ w,
                                 cplx* __restrict__ out) {
  const size_t D = 2 * (size_t)n, i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D * D * count) return;
  const size_t z = i / (D * D), o = i % (D * D), r = o / D, c = o % D;
  cplx v = cmake(0, 0);
  if (r < (size_t)n && c < (size_t)n) v = u[(z * n + r) * n + c];
  else if (r < (size_t)n) v = w[(z * n + r) * n + (c - n)];
  else if (c >= (size_t)n) v = u[(z * n + (r - n)) * n + (c - n)];
  out[i] = v;
}

namespace {
struct BlockWork {
  cplx *xu, *xw, *tu, *tw, *su, *sw, *ru, *rw;   // generator blocks, Taylor term, scratch, result (one batch each)
  double* nrm;
};

void zgemm(cudaStream_t st, int n, int batch, const cplx* A, size_t sA, const cplx* B, size_t sB, cplx* C, size_t sC,
           double alpha, double beta, cplx* acc = nullptr, size_t sAcc = 0) {
  dim3 grid((n + ZG_BN - 1) / ZG_BN, (n + ZG_BM - 1) / ZG_BM, batch);
  k_zgemm_dmma<<<grid, 128, 0, st>>>(n, A, sA, B, sB, C, sC, alpha, beta, acc, sAcc);
}
}  // namespace

// (eu, ew)[z] = block form of expm([[xu, xw], [0, xu]]) for the slices k0 .. k0 + count - 1
static int block_expm_batch(qocvl_plan* p, const BlockWork& W, int k0, int count, cplx* eu, cplx* ew, cudaStream_t st) {
  const int n = p->d.n, J = p->d.n_gen;
  const size_t nn = (size_t)n * n;
  k_block_generators<<<count, 256, 0, st>>>(n, J, p->d.dt, k0, p->d_gblk_u, p->d_gblk_w, p->d_coef, W.xu, W.xw, W.nrm);
  std::vector<double> h(count);
  QCUDA(cudaMemcpyAsync(h.data(), W.nrm, count * sizeof(double), cudaMemcpyDeviceToHost, st));
  QCUDA(cudaStreamSynchronize(st));                      // (API path, not the optimiser's hot path)
  double nrm = 0.0;
  for (double v : h) nrm = std::max(nrm, v);
  if (!(nrm == nrm) || nrm > 1e6) return QOCVL_ERR_NORM;
  int sq = 0;
  while (nrm > 0.5 && sq < 60) { nrm *= 0.5; ++sq; }
  const size_t tot = nn * count;
  k_block_scale_init<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(nn, n, count, ldexp(1.0, -sq), W.xu, W.xw, W.tu, W.tw,
                                                                 W.ru, W.rw);
  cplx *tu = W.tu, *tw = W.tw, *su = W.su, *sw = W.sw;
  for (int j = 2; j <= 16; ++j) {                        // t_j = x t_{j-1} / j in block form; r += t_j
    const double inv = 1.0 / (double)j;
    zgemm(st, n, count, W.xu, nn, tw, nn, sw, nn, inv, 0.0);                       // xu tw
    zgemm(st, n, count, W.xw, nn, tu, nn, sw, nn, inv, 1.0, W.rw, nn);              // + xw tu   -> rw += .
    zgemm(st, n, count, W.xu, nn, tu, nn, su, nn, inv, 0.0, W.ru, nn);              // xu tu     -> ru += .
    std::swap(tu, su); std::swap(tw, sw);
    p->launches += 3;
  }
  cplx *ru = W.ru, *rw = W.rw;
  for (int s = 0; s < sq; ++s) {                         // (u, w)^2 = (u u, u w + w u)
    zgemm(st, n, count, ru, nn, rw, nn, sw, nn, 1.0, 0.0);
    zgemm(st, n, count, rw, nn, ru, nn, sw, nn, 1.0, 1.0);
    zgemm(st, n, count, ru, nn, ru, nn, su, nn, 1.0, 0.0);
    std::swap(ru, su); std::swap(rw, sw);
    p->launches += 3;
  }
  QCUDA(cudaMemcpyAsync(eu, ru, tot * sizeof(cplx), cudaMemcpyDeviceToDevice, st));
  QCUDA(cudaMemcpyAsync(ew, rw, tot * sizeof(cplx), cudaMemcpyDeviceToDevice, st));
  QCUDA(cudaGetLastError());
  return QOCVL_OK;
}

static int ensure_blocks(qocvl_plan* p, BlockWork* W) {
  const int n = p->d.n, J = p->d.n_gen;
  const size_t nn = (size_t)n * n;
  if (!p->d_gblk_u) {                                    // dense n x n generator blocks, uploaded on first use
    std::vector<cplx> gu(J * nn), gw(J * nn);
    for (size_t i = 0; i < J * nn; ++i) {
      gu[i] = cmake(p->h_gu[i].real(), p->h_gu[i].imag());
      gw[i] = cmake(p->h_gw[i].real(), p->h_gw[i].imag());
    }
    QCUDA(cudaMalloc((void**)&p->d_gblk_u, J * nn * sizeof(cplx)));
    QCUDA(cudaMalloc((void**)&p->d_gblk_w, J * nn * sizeof(cplx)));
    QCUDA(cudaMemcpy(p->d_gblk_u, gu.data(), J * nn * sizeof(cplx), cudaMemcpyHostToDevice));
    QCUDA(cudaMemcpy(p->d_gblk_w, gw.data(), J * nn * sizeof(cplx), cudaMemcpyHostToDevice));
    p->bytes += 2 * J * nn * sizeof(cplx);
  }
  const int batch = std::min(DENSE_BATCH, p->d.n_slices);
  if (!p->d_blk_work) {
    QCUDA(cudaMalloc((void**)&p->d_blk_work, 8 * nn * batch * sizeof(cplx) + batch * sizeof(double)));
    p->bytes += 8 * nn * batch * sizeof(cplx);
  }
  cplx* base = p->d_blk_work;
  cplx** slots[8] = {&W->xu, &W->xw, &W->tu, &W->tw, &W->su, &W->sw, &W->ru, &W->rw};
  for (int i = 0; i < 8; ++i) *slots[i] = base + (size_t)i * nn * batch;
  W->nrm = (double*)(base + 8 * nn * batch);
  return QOCVL_OK;
}

// exponentials() for n > 16: out = dense [M][2n][2n] (caller's buffer)
int qocvl_launch_block_expm(qocvl_plan* p, cplx* out, cudaStream_t st) {
  BlockWork W;
  int rc = ensure_blocks(p, &W);
  if (rc) return rc;
  const int n = p->d.n, M = p->d.n_slices, batch = std::min(DENSE_BATCH, M);
  const size_t nn = (size_t)n * n, D2 = 4 * nn;
  cplx *eu = nullptr, *ew = nullptr;
  QCUDA(cudaMalloc((void**)&eu, 2 * nn * batch * sizeof(cplx)));
  ew = eu + nn * batch;
  for (int k0 = 0; k0 < M && rc == QOCVL_OK; k0 += batch) {
    const int count = std::min(batch, M - k0);
    rc = block_expm_batch(p, W, k0, count, eu, ew, st);
    if (rc == QOCVL_OK) {
      const size_t tot = D2 * count;
      k_block_assemble<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(n, count, eu, ew, out + (size_t)k0 * D2);
      p->launches++;
    }
  }
  cudaStreamSynchronize(st);
  cudaFree(eu);
  if (rc == QOCVL_OK) QCUDA(cudaGetLastError());
  return rc;
}

// propagate() for n > 16: the reference's halving order on (u, w) pairs; out = dense [2n][2n]
int qocvl_launch_block_propagate(qocvl_plan* p, cplx* out, cudaStream_t st) {
  BlockWork W;
  int rc = ensure_blocks(p, &W);
  if (rc) return rc;
  const int n = p->d.n, M = p->d.n_slices, batch = std::min(DENSE_BATCH, M);
  const size_t nn = (size_t)n * n, pair = 2 * nn;      // a pair = u block then w block
  cplx *a = nullptr, *b = nullptr;
  QCUDA(cudaMalloc((void**)&a, (size_t)M * pair * sizeof(cplx)));
  if (cudaMalloc((void**)&b, (size_t)(M / 2 + 1) * pair * sizeof(cplx)) != cudaSuccess) { cudaFree(a); return QOCVL_ERR_CUDA; }
  cplx *eu = nullptr;
  if (cudaMalloc((void**)&eu, 2 * nn * batch * sizeof(cplx)) != cudaSuccess) { cudaFree(a); cudaFree(b); return QOCVL_ERR_CUDA; }
  cplx* ew = eu + nn * batch;
  for (int k0 = 0; k0 < M && rc == QOCVL_OK; k0 += batch) {
    const int count = std::min(batch, M - k0);
    rc = block_expm_batch(p, W, k0, count, eu, ew, st);
    if (rc == QOCVL_OK) {                               // interleave into pairs [k][u | w]
      cudaMemcpy2DAsync(a + (size_t)k0 * pair, pair * sizeof(cplx), eu, nn * sizeof(cplx), nn * sizeof(cplx), count,
                        cudaMemcpyDeviceToDevice, st);
      cudaMemcpy2DAsync(a + (size_t)k0 * pair + nn, pair * sizeof(cplx), ew, nn * sizeof(cplx), nn * sizeof(cplx), count,
                        cudaMemcpyDeviceToDevice, st);
    }
  }
  int count = M;
  cplx *src = a, *dst = b;
  while (rc == QOCVL_OK && count > 1) {                  // ST:77-87 / TQ:145-154 halving sequence
    const int fold = count & 1, np_ = count / 2;
    // out_i = in_{2i+1} in_{2i}:  u = u2 u1,  w = u2 w1 + w2 u1     (pair stride 2 * pair between consecutive i)
    zgemm(st, n, np_, src + pair, 2 * pair, src, 2 * pair, dst, pair, 1.0, 0.0);                  // u2 u1
    zgemm(st, n, np_, src + pair, 2 * pair, src + nn, 2 * pair, dst + nn, pair, 1.0, 0.0);        // u2 w1
    zgemm(st, n, np_, src + pair + nn, 2 * pair, src, 2 * pair, dst + nn, pair, 1.0, 1.0);        // + w2 u1
    p->launches += 3;
    if (fold) {                                          // last product left-multiplied by the odd leftover
      const cplx* lo = src + (size_t)(2 * np_) * pair;
      cplx* last = dst + (size_t)(np_ - 1) * pair;
      cplx* tmp = W.su;                                  // 2 nn of scratch (su, sw are contiguous slots of >= nn each)
      zgemm(st, n, 1, lo, 0, last + nn, 0, W.sw, 0, 1.0, 0.0);          // lu w
      zgemm(st, n, 1, lo + nn, 0, last, 0, W.sw, 0, 1.0, 1.0);          // + lw u
      zgemm(st, n, 1, lo, 0, last, 0, tmp, 0, 1.0, 0.0);                // lu u
      QCUDA(cudaMemcpyAsync(last, tmp, nn * sizeof(cplx), cudaMemcpyDeviceToDevice, st));
      QCUDA(cudaMemcpyAsync(last + nn, W.sw, nn * sizeof(cplx), cudaMemcpyDeviceToDevice, st));
      p->launches += 3;
    }
    std::swap(src, dst);
    count = np_;
  }
  if (rc == QOCVL_OK) {
    const size_t tot = 4 * nn;
    k_block_assemble<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(n, 1, src, src + nn, out);
    p->launches++;
  }
  cudaStreamSynchronize(st);
  cudaFree(a); cudaFree(b); cudaFree(eu);
  if (rc == QOCVL_OK) QCUDA(cudaGetLastError());
  return rc;
}
