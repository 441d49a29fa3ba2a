// lindblad.cu -- master-equation propagation on the GPU: the step AFTER the hot path.
//
// Replaces what the reference delegates to QuTiP's mesolve in two_qubit/qutip_simulation.py:57-103 and
// two_qubit/qutip_five_level.py:88-142 (two 5-level atoms, 25-dimensional density matrix, 8 collapse operators):
//     d rho / dt = L(t) rho,   L(t) = L_0 + sum_j f_j(t) L_j      (vectorised rho, n = 625)
// with the Liouvillian pieces L_j passed as values on one union CSR pattern.  Time stepping: exponential midpoint
// rule, rho <- exp(L(t_mid) dt) rho per step, the exponential applied with the same truncated Taylor series as the
// chain kernels (degree from the host-computed 1-norm bound of the step, remainder < 2^-55).  The host wrapper
// chooses the sub-stepping and evaluates f_j at the midpoints.
// Mapping: ONE CTA PER TRAJECTORY (initial state x coefficient set), one matrix row per thread, the vector, the
// Taylor term and the assembled step matrix resident in shared memory, one __syncthreads per Taylor step.
#include "common.cuh"
#include <algorithm>
#include <cmath>

struct LindbladArgs {
  int n, nnz, J, n_steps, n_obs;
  const int *rowptr, *col;
  const cplx* vals;     // [J][nnz]
  const cplx* coef;     // [n_traj][n_steps][J]
  const int* qdeg;      // [n_traj][n_steps]
  const cplx* rho0;     // [n_traj][n]
  const cplx* obs;      // [n_obs][n]
  double dt;
  cplx* expect;         // [n_traj][n_steps + 1][n_obs]
  cplx* rho_out;        // [n_traj][n]
};

__global__ void __launch_bounds__(1024) k_lindblad(const LindbladArgs P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, nt = blockDim.x, traj = blockIdx.x;
  const int n = P.n, J = P.J;
  cplx* s_a = (cplx*)smem_raw;           // [nnz] assembled dt * L(t_mid)
  cplx* s_z = s_a + P.nnz;               // [n] state
  cplx* s_t0 = s_z + n;                  // [n] Taylor term (ping)
  cplx* s_t1 = s_t0 + n;                 // [n] Taylor term (pong)
  cplx* s_c = s_t1 + n;                  // [J] step coefficients
  __shared__ double s_red[2 * 32];
  const cplx czero = cmake(0, 0);
  const cplx* coef = P.coef + (size_t)traj * P.n_steps * J;
  for (int i = tid; i < n; i += nt) s_z[i] = P.rho0[(size_t)traj * n + i];
  __syncthreads();

  // expectation values <O_o> = sum_i obs[o][i] * rho[i], block-reduced in a fixed order
  auto observe = [&](int k) {
    for (int o = 0; o < P.n_obs; ++o) {
      cplx acc = czero;
      for (int i = tid; i < n; i += nt) acc = cfma(P.obs[(size_t)o * n + i], s_z[i], acc);
      for (int off = 16; off > 0; off >>= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
      }
      if ((tid & 31) == 0) { s_red[2 * (tid >> 5)] = acc.x; s_red[2 * (tid >> 5) + 1] = acc.y; }
      __syncthreads();
      if (tid == 0) {
        cplx tot = czero;
        for (int w = 0; w < (nt + 31) / 32; ++w) tot = cadd(tot, cmake(s_red[2 * w], s_red[2 * w + 1]));
        P.expect[((size_t)traj * (P.n_steps + 1) + k) * P.n_obs + o] = tot;
      }
      __syncthreads();
    }
  };
  observe(0);
  for (int k = 0; k < P.n_steps; ++k) {
    if (tid < J) s_c[tid] = cscale(coef[(size_t)k * J + tid], P.dt);
    __syncthreads();
    for (int e = tid; e < P.nnz; e += nt) {
      cplx acc = czero;
      for (int j = 0; j < J; ++j) acc = cfma(s_c[j], P.vals[(size_t)j * P.nnz + e], acc);
      s_a[e] = acc;
    }
    for (int i = tid; i < n; i += nt) s_t0[i] = s_z[i];
    __syncthreads();
    const int q = P.qdeg[(size_t)traj * P.n_steps + k];
    cplx* tin = s_t0;
    cplx* tout = s_t1;
    for (int j = 1; j <= q; ++j) {
      const double invj = 1.0 / (double)j;
      for (int r = tid; r < n; r += nt) {
        cplx acc = czero;
        for (int e = P.rowptr[r]; e < P.rowptr[r + 1]; ++e) acc = cfma(s_a[e], tin[P.col[e]], acc);
        acc = cscale(acc, invj);
        tout[r] = acc;
        s_z[r] = cadd(s_z[r], acc);
      }
      __syncthreads();
      cplx* sw = tin; tin = tout; tout = sw;
    }
    observe(k + 1);
  }
  for (int i = tid; i < n; i += nt) P.rho_out[(size_t)traj * n + i] = s_z[i];
}

// All pointers are HOST pointers: this is a one-shot validation call, not a stream-ordered hot-path entry.
extern "C" int qocvl_lindblad_propagate(int device, int n, int nnz, int n_gen, int n_steps, int n_traj, int n_obs,
                                        const int* rowptr, const int* col, const double* vals, const double* coef,
                                        const double* rho0, const double* obs, double dt, double* expect_out,
                                        double* rho_out) {
  if (n < 1 || nnz < 1 || n_gen < 1 || n_steps < 1 || n_traj < 1 || n_obs < 0 || !rowptr || !col || !vals || !coef ||
      !rho0 || !rho_out || !(dt > 0) || (n_obs > 0 && (!obs || !expect_out)))
    return QOCVL_ERR_ARG;
  QCUDA(cudaSetDevice(device));
  const size_t smem = ((size_t)nnz + 3 * (size_t)n + n_gen) * sizeof(cplx);
  if (smem > 220 * 1024) return QOCVL_ERR_UNSUPPORTED;
  // ---- per-step Taylor degree from the entrywise bound  |dt L|_1 <= dt * max_col sum_j |c_j| * colsum_j
  std::vector<double> colsum((size_t)n_gen * n, 0.0);
  for (int j = 0; j < n_gen; ++j)
    for (int r = 0; r < n; ++r)
      for (int e = rowptr[r]; e < rowptr[r + 1]; ++e) {
        if (col[e] < 0 || col[e] >= n) return QOCVL_ERR_ARG;
        colsum[(size_t)j * n + col[e]] += std::hypot(vals[2 * ((size_t)j * nnz + e)], vals[2 * ((size_t)j * nnz + e) + 1]);
      }
  std::vector<int> qdeg((size_t)n_traj * n_steps);
  for (size_t ts = 0; ts < qdeg.size(); ++ts) {
    double worst = 0.0;
    for (int c = 0; c < n; ++c) {
      double s = 0.0;
      for (int j = 0; j < n_gen; ++j)
        s += std::hypot(coef[2 * (ts * n_gen + j)], coef[2 * (ts * n_gen + j) + 1]) * colsum[(size_t)j * n + c];
      worst = std::max(worst, s);
    }
    const double nrm = worst * dt;
    if (!(nrm <= QOCVL_NORM_MAX)) return QOCVL_ERR_NORM;   // also NaN: the caller has to sub-step finer
    qdeg[ts] = qocvl_taylor_degree(nrm);
  }
  int* d_rowptr = nullptr; int* d_col = nullptr; int* d_q = nullptr;
  cplx *d_vals = nullptr, *d_coef = nullptr, *d_rho0 = nullptr, *d_obs = nullptr, *d_exp = nullptr, *d_out = nullptr;
  auto cleanup = [&]() {
    for (void* q : {(void*)d_rowptr, (void*)d_col, (void*)d_q, (void*)d_vals, (void*)d_coef, (void*)d_rho0, (void*)d_obs,
                    (void*)d_exp, (void*)d_out})
      if (q) cudaFree(q);
  };
#define LB_UP(dst, src, count, type)                                                                   \
  do {                                                                                                 \
    if (cudaMalloc((void**)&(dst), std::max<size_t>(1, (count)) * sizeof(type)) != cudaSuccess ||      \
        ((count) > 0 && cudaMemcpy((dst), (src), (count) * sizeof(type), cudaMemcpyHostToDevice) != cudaSuccess)) { \
      g_qocvl_cuda_error = "lindblad upload: " + std::string(cudaGetErrorString(cudaGetLastError()));   \
      cleanup();                                                                                       \
      return QOCVL_ERR_CUDA;                                                                           \
    }                                                                                                  \
  } while (0)
  const size_t n_exp = (size_t)n_traj * (n_steps + 1) * std::max(1, n_obs);
  LB_UP(d_rowptr, rowptr, (size_t)n + 1, int);
  LB_UP(d_col, col, (size_t)nnz, int);
  LB_UP(d_q, qdeg.data(), qdeg.size(), int);
  LB_UP(d_vals, vals, (size_t)n_gen * nnz, cplx);
  LB_UP(d_coef, coef, (size_t)n_traj * n_steps * n_gen, cplx);
  LB_UP(d_rho0, rho0, (size_t)n_traj * n, cplx);
  LB_UP(d_obs, obs, (size_t)n_obs * n, cplx);
#undef LB_UP
  if (cudaMalloc((void**)&d_exp, n_exp * sizeof(cplx)) != cudaSuccess ||
      cudaMalloc((void**)&d_out, (size_t)n_traj * n * sizeof(cplx)) != cudaSuccess) {
    g_qocvl_cuda_error = "lindblad workspace allocation failed";
    cleanup();
    return QOCVL_ERR_CUDA;
  }
  LindbladArgs a;
  a.n = n; a.nnz = nnz; a.J = n_gen; a.n_steps = n_steps; a.n_obs = n_obs;
  a.rowptr = d_rowptr; a.col = d_col; a.vals = d_vals; a.coef = d_coef; a.qdeg = d_q; a.rho0 = d_rho0; a.obs = d_obs;
  a.dt = dt; a.expect = d_exp; a.rho_out = d_out;
  const int threads = std::min(1024, (n + 31) / 32 * 32);
  cudaError_t e = cudaFuncSetAttribute(k_lindblad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) {
    k_lindblad<<<n_traj, threads, smem>>>(a);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e == cudaSuccess && n_obs > 0)
    e = cudaMemcpy(expect_out, d_exp, (size_t)n_traj * (n_steps + 1) * n_obs * sizeof(cplx), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(rho_out, d_out, (size_t)n_traj * n * sizeof(cplx), cudaMemcpyDeviceToHost);
  cleanup();
  if (e != cudaSuccess) {
    g_qocvl_cuda_error = std::string("lindblad: ") + cudaGetErrorString(e);
    return QOCVL_ERR_CUDA;
  }
  return QOCVL_OK;
}
