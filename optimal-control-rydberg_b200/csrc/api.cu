// api.cu -- hot-path entry points of the C-ABI (include/qocvl.h): each composes the kernels of
// pulse.cu, chain.cu and dense.cu on the caller's stream.  No host synchronisation here.
#include "common.cuh"

static int check_ready(qocvl_plan* p) {
  if (!p) return QOCVL_ERR_ARG;
  if (!p->have_gens || !p->have_cost) return QOCVL_ERR_STATE;
  if (p->d.pad > 0 && !p->have_filter) return QOCVL_ERR_STATE;
  if (cudaSetDevice(p->device) != cudaSuccess) return QOCVL_ERR_CUDA;
  return qocvl_chain_configure(p);
}

extern "C" int qocvl_physical_amplitudes(qocvl_plan* p, const double* abc, double* wave, void* stream) {
  int rc = check_ready(p);
  if (rc) return rc;
  if (!abc || !wave) return QOCVL_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = qocvl_launch_pulse_forward(p, abc, st))) return rc;
  QCUDA(cudaMemcpyAsync(wave, p->d_wave, (size_t)p->d.n_slices * p->d.n_chan * sizeof(double),
                        cudaMemcpyDeviceToDevice, st));
  return QOCVL_OK;
}

extern "C" int qocvl_physical_amplitudes_vjp(qocvl_plan* p, const double* abc, const double* gwave,
                                             double* grads, void* stream) {
  int rc = check_ready(p);
  if (rc) return rc;
  if (!abc || !gwave || !grads) return QOCVL_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = qocvl_launch_pulse_forward(p, abc, st))) return rc;   // refreshes raw[] for the reverse pass
  return qocvl_launch_pulse_vjp(p, abc, gwave, grads, st);
}

extern "C" int qocvl_auxiliary_amplitudes(qocvl_plan* p, const double* abc, double* aux, void* stream) {
  int rc = check_ready(p);
  if (rc) return rc;
  if (!abc || !aux) return QOCVL_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = qocvl_launch_pulse_forward(p, abc, st))) return rc;
  if ((rc = qocvl_launch_coefficients(p, p->d_wave, st))) return rc;
  return qocvl_launch_aux(p, aux, st);
}

extern "C" int qocvl_exponentials(qocvl_plan* p, const double* abc, double* out, void* stream) {
  int rc = check_ready(p);
  if (rc) return rc;
  if (!abc || !out) return QOCVL_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = qocvl_launch_pulse_forward(p, abc, st))) return rc;
  if ((rc = qocvl_launch_coefficients(p, p->d_wave, st))) return rc;
  return qocvl_launch_dense_expm(p, (cplx*)out, st);
}

extern "C" int qocvl_propagate(qocvl_plan* p, const double* abc, double* out, void* stream) {
  int rc = check_ready(p);
  if (rc) return rc;
  if (!abc || !out) return QOCVL_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = qocvl_launch_pulse_forward(p, abc, st))) return rc;
  if ((rc = qocvl_launch_coefficients(p, p->d_wave, st))) return rc;
  if ((rc = qocvl_launch_dense_expm(p, nullptr, st))) return rc;   // into the plan's slice buffer
  return qocvl_launch_dense_tree(p, p->d_dense_a, (cplx*)out, st);
}

extern "C" int qocvl_cost(qocvl_plan* p, const double* abc, double* out3, void* stream) {
  int rc = check_ready(p);
  if (rc) return rc;
  if (!abc || !out3) return QOCVL_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = qocvl_launch_pulse_forward(p, abc, st))) return rc;
  if ((rc = qocvl_launch_coefficients(p, p->d_wave, st))) return rc;
  return qocvl_launch_chain(p, false, out3, nullptr, st);
}

extern "C" int qocvl_cost_grad(qocvl_plan* p, const double* abc, double* out, void* stream) {
  int rc = check_ready(p);
  if (rc) return rc;
  if (!abc || !out) return QOCVL_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = qocvl_launch_pulse_forward(p, abc, st))) return rc;
  if ((rc = qocvl_launch_coefficients(p, p->d_wave, st))) return rc;
  if ((rc = qocvl_launch_chain(p, true, out, p->d_gwave, st))) return rc;
  return qocvl_launch_pulse_vjp(p, abc, p->d_gwave, out + 3, st);
}

extern "C" int qocvl_cost_grad_from_wave(qocvl_plan* p, const double* wave, double* out3, double* gwave,
                                         void* stream) {
  int rc = check_ready(p);
  if (rc) return rc;
  if (!wave || !out3) return QOCVL_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = qocvl_launch_coefficients(p, wave, st))) return rc;
  return qocvl_launch_chain(p, gwave != nullptr, out3, gwave ? gwave : nullptr, st);
}

// Synchronises the device and reports (then clears) the sticky norm / degree flag.
extern "C" int qocvl_check(qocvl_plan* p) {
  if (!p) return QOCVL_ERR_ARG;
  int flag = 0;
  QCUDA(cudaSetDevice(p->device));
  QCUDA(cudaDeviceSynchronize());
  QCUDA(cudaMemcpy(&flag, p->d_flag, sizeof(int), cudaMemcpyDeviceToHost));
  if (flag) {
    cudaMemset(p->d_flag, 0, sizeof(int));
    return QOCVL_ERR_NORM;
  }
  return QOCVL_OK;
}

// Executed-FP64-flop model of the last evaluation (reads the per-slice Taylor degrees; synchronises).
// Counts only the useful complex multiply-adds (8 flops each) of the chain kernel:
//   forward  : q_k * (R_u*nnz_u + nnz_w)                      R_u = n_vec + 1 vectors
//   reverse  : (q_k - 1) * (same)  recompute  +  2 * q_k * (same)  adjoint mat-vec and contraction
extern "C" double qocvl_flops_per_call(const qocvl_plan* p, int with_grad) {
  if (!p || !p->have_gens) return 0.0;
  const int M = p->d.n_slices;
  std::vector<int> q(M);
  cudaSetDevice(p->device);
  cudaDeviceSynchronize();
  if (cudaMemcpy(q.data(), p->d_qdeg, M * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return 0.0;
  const double macs = (double)(p->d.n_vec + 1) * p->lu.nnz + p->lw.nnz;
  double total = 0.0;
  for (int k = 0; k < M; ++k) {
    total += q[k] * macs;
    if (with_grad) total += ((q[k] - 1) + 2.0 * q[k]) * macs;
  }
  return total * 8.0 * (double)p->n_samples;
}
