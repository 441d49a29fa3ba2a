// dense.cu -- the dense Van Loan propagators the reference API exposes:
//   exponentials(): ST:247-256 / TQ:460-484   [M][2n][2n] = expm(dt * (drift + sum_j aux_kj G_j))
//   propagate():    ST:258-283 / TQ:489-535   pairwise product tree, later slice on the LEFT,
//                                             odd leftover folded into the tail (same order as the reference)
// One CTA per slice / per product; matrices (2n <= 32) live in shared memory.  The exponential
// is scaling-and-squaring around a degree-16 Taylor polynomial (|X/2^s|_1 <= 1/2).
#include "common.cuh"

#define DENSE_THREADS 256

// C = A * B  (D x D complex, all in shared memory); every thread owns a strided set of outputs
__device__ __forceinline__ void smem_matmul(int D, const cplx* A, const cplx* B, cplx* Cm, double scale) {
  for (int o = threadIdx.x; o < D * D; o += blockDim.x) {
    const int i = o / D, j = o % D;
    cplx acc = cmake(0, 0);
    for (int l = 0; l < D; ++l) acc = cfma(A[i * D + l], B[l * D + j], acc);
    Cm[o] = cscale(acc, scale);
  }
}

__global__ void __launch_bounds__(DENSE_THREADS) k_dense_expm(int D, int J, double dt,
                                                            const cplx* __restrict__ gdense,
                                                            const cplx* __restrict__ coef,
                                                            cplx* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* X = (cplx*)smem_raw;       // [D][D]
  cplx* T = X + D * D;             // term
  cplx* Rm = T + D * D;            // result
  cplx* W = Rm + D * D;            // scratch
  __shared__ double colsum[64];
  __shared__ int s_sq;
  const int k = blockIdx.x, tid = threadIdx.x, DD = D * D;
  for (int o = tid; o < DD; o += blockDim.x) {
    cplx acc = cmake(0, 0);
    for (int j = 0; j < J; ++j) acc = cfma(coef[(size_t)k * J + j], gdense[(size_t)j * DD + o], acc);
    X[o] = cscale(acc, dt);
  }
  __syncthreads();
  if (tid < D) {
    double s = 0.0;
    for (int i = 0; i < D; ++i) s += hypot(X[i * D + tid].x, X[i * D + tid].y);
    colsum[tid] = s;
  }
  __syncthreads();
  if (tid == 0) {
    double nrm = 0.0;
    for (int j = 0; j < D; ++j) nrm = fmax(nrm, colsum[j]);
    int s = 0;
    while (nrm > 0.5 && s < 60) { nrm *= 0.5; ++s; }
    s_sq = s;
  }
  __syncthreads();
  const int sq = s_sq;
  const double sc = ldexp(1.0, -sq);
  for (int o = tid; o < DD; o += blockDim.x) {
    X[o] = cscale(X[o], sc);
    T[o] = X[o];
    Rm[o] = cadd(X[o], ((o / D) == (o % D)) ? cmake(1, 0) : cmake(0, 0));
  }
  __syncthreads();
  for (int j = 2; j <= 16; ++j) {
    smem_matmul(D, T, X, W, 1.0 / (double)j);
    __syncthreads();
    for (int o = tid; o < DD; o += blockDim.x) { T[o] = W[o]; Rm[o] = cadd(Rm[o], W[o]); }
    __syncthreads();
  }
  for (int s = 0; s < sq; ++s) {
    smem_matmul(D, Rm, Rm, W, 1.0);
    __syncthreads();
    for (int o = tid; o < DD; o += blockDim.x) Rm[o] = W[o];
    __syncthreads();
  }
  for (int o = tid; o < DD; o += blockDim.x) out[(size_t)k * DD + o] = Rm[o];
}

// out[i] = in[2i+1] * in[2i]  for i < npairs; if fold != 0 the LAST product is additionally
// left-multiplied by in[2*npairs] (the odd leftover):  out[npairs-1] = leftover * (in[2m-1]*in[2m-2]).
__global__ void __launch_bounds__(DENSE_THREADS) k_pair_products(int D, int npairs, int fold,
                                                               const cplx* __restrict__ in,
                                                               cplx* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* A = (cplx*)smem_raw;
  cplx* B = A + D * D;
  cplx* Cm = B + D * D;
  const int i = blockIdx.x, DD = D * D;
  for (int o = threadIdx.x; o < DD; o += blockDim.x) {
    A[o] = in[(size_t)(2 * i + 1) * DD + o];
    B[o] = in[(size_t)(2 * i) * DD + o];
  }
  __syncthreads();
  smem_matmul(D, A, B, Cm, 1.0);
  __syncthreads();
  if (fold && i == npairs - 1) {
    for (int o = threadIdx.x; o < DD; o += blockDim.x) A[o] = in[(size_t)(2 * npairs) * DD + o];
    __syncthreads();
    smem_matmul(D, A, Cm, B, 1.0);
    __syncthreads();
    for (int o = threadIdx.x; o < DD; o += blockDim.x) out[(size_t)i * DD + o] = B[o];
  } else {
    for (int o = threadIdx.x; o < DD; o += blockDim.x) out[(size_t)i * DD + o] = Cm[o];
  }
}

static int ensure_dense(qocvl_plan* p) {
  const int D = 2 * p->d.n, M = p->d.n_slices;
  if (D > 32) return QOCVL_ERR_UNSUPPORTED;
  if (!p->d_dense_a) {
    QCUDA(cudaMalloc((void**)&p->d_dense_a, (size_t)M * D * D * sizeof(cplx)));
    QCUDA(cudaMalloc((void**)&p->d_dense_b, (size_t)(M / 2 + 1) * D * D * sizeof(cplx)));
    p->bytes += (size_t)(M + M / 2 + 1) * D * D * sizeof(cplx);
  }
  return QOCVL_OK;
}

int qocvl_launch_dense_expm(qocvl_plan* p, cplx* out, cudaStream_t st) {
  int rc = ensure_dense(p);
  if (rc) return rc;
  const int D = 2 * p->d.n;
  if (!out) out = p->d_dense_a;
  const size_t smem = (size_t)4 * D * D * sizeof(cplx);
  QCUDA(cudaFuncSetAttribute(k_dense_expm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_dense_expm<<<p->d.n_slices, DENSE_THREADS, smem, st>>>(D, p->d.n_gen, p->d.dt, p->d_gdense, p->d_coef, out);
  p->launches++;
  QCUDA(cudaGetLastError());
  return QOCVL_OK;
}

// slices: [M][D][D] (destroyed); out: [D][D]
int qocvl_launch_dense_tree(qocvl_plan* p, cplx* slices, cplx* out, cudaStream_t st) {
  int rc = ensure_dense(p);
  if (rc) return rc;
  const int D = 2 * p->d.n;
  const size_t smem = (size_t)3 * D * D * sizeof(cplx);
  QCUDA(cudaFuncSetAttribute(k_pair_products, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int count = p->d.n_slices;
  cplx* src = slices;
  cplx* dst = p->d_dense_b;
  while (count > 1) {                      // ST:77-87 / TQ:145-154 halving sequence
    const int fold = count & 1, npairs = count / 2;
    k_pair_products<<<npairs, DENSE_THREADS, smem, st>>>(D, npairs, fold, src, dst);
    p->launches++;
    cplx* t = src; src = dst; dst = t;
    count = npairs;
  }
  QCUDA(cudaMemcpyAsync(out, src, (size_t)D * D * sizeof(cplx), cudaMemcpyDeviceToDevice, st));
  QCUDA(cudaGetLastError());
  return QOCVL_OK;
}
