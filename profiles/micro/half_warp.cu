// half_warp.cu -- does the FP64 pipe skip the inactive half of a warp?  DFMA stream (three register operands, 16
// independent accumulators) issued by warps in which only lanes 0-15 (or 0-7) execute it, 1 / 2 / 4 warps per scheduler.
// If passes without active threads are skipped, two half-full warps per scheduler cost what one full warp costs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o half_warp half_warp.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N_IT 4096
#define NA 16

__global__ void k(double* out, const double* in, long long* cyc, int active) {
  double a[NA], b[NA], c[NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) { a[i] = in[i + threadIdx.x]; b[i] = in[NA + i + threadIdx.x]; c[i] = in[2 * NA + i + threadIdx.x]; }
  __syncthreads();
  const long long t0 = clock64();
  if ((threadIdx.x & 31) < active) {
#pragma unroll 1
    for (int it = 0; it < N_IT; ++it) {
#pragma unroll
      for (int i = 0; i < NA; ++i) c[i] = fma(a[i], b[i], c[i]);
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NA; ++i) s += c[i] + a[i] + b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  double *out, *in;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8);
  cudaMalloc(&in, 4096 * 8);
  cudaMemset(in, 0, 4096 * 8);
  cudaMalloc(&cyc, 148 * 8);
  for (int active : {32, 16, 8}) {
    for (int wps = 1; wps <= 4; wps *= 2) {
      k<<<148, 128 * wps>>>(out, in, cyc, active);
      cudaDeviceSynchronize();
      long long h[148];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double avg = 0;
      for (int i = 0; i < 148; ++i) avg += h[i];
      avg /= 148;
      printf("active lanes %2d, warps/scheduler %d: %.2f cycles per warp-DFMA per scheduler (%.2f per warp)\n", active, wps,
             avg / ((double)N_IT * NA * wps), avg / ((double)N_IT * NA));
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
