// latency.cu -- dependent-issue latency of DFMA / DADD and latency + throughput of SHFL.BFLY on sm_100a, one warp per scheduler.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o latency latency.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N_IT 8192

template <int MODE>
__global__ void k(double* out, const double* in, long long* cyc) {
  double a = in[threadIdx.x], b = in[32 + threadIdx.x], c = in[64 + threadIdx.x];
  double v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = in[100 + 8 * i + threadIdx.x];
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < N_IT; ++it) {
    if (MODE == 0) {          // 8 dependent DFMAs
#pragma unroll
      for (int i = 0; i < 8; ++i) c = fma(a, b, c);
    } else if (MODE == 1) {   // 8 dependent DADDs
#pragma unroll
      for (int i = 0; i < 8; ++i) c = c + a;
    } else if (MODE == 2) {   // 8 dependent 64-bit xor-1 exchanges (2 SHFL each)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int lo = __shfl_xor_sync(0xffffffffu, __double2loint(c), 1), hi = __shfl_xor_sync(0xffffffffu, __double2hiint(c), 1);
        c = __hiloint2double(hi, lo);
      }
    } else if (MODE == 3) {   // 8 independent 64-bit exchanges (16 SHFL)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int lo = __shfl_xor_sync(0xffffffffu, __double2loint(v[i]), 1), hi = __shfl_xor_sync(0xffffffffu, __double2hiint(v[i]), 1);
        v[i] = __hiloint2double(hi, lo);
      }
    } else if (MODE == 4) {   // exchange -> DFMA -> exchange -> DFMA ... (the Taylor step's critical path), 4 rounds
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int lo = __shfl_xor_sync(0xffffffffu, __double2loint(c), 1), hi = __shfl_xor_sync(0xffffffffu, __double2hiint(c), 1);
        c = fma(a, __hiloint2double(hi, lo), b);
      }
    } else if (MODE == 5) {   // 16 independent SHFL + 16 independent DFMA interleaved
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int lo = __shfl_xor_sync(0xffffffffu, __double2loint(v[i]), 1), hi = __shfl_xor_sync(0xffffffffu, __double2hiint(v[i]), 1);
        v[i] = fma(a, __hiloint2double(hi, lo), b);
      }
    }
  }
  const long long t1 = clock64();
  double s = a + b + c;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_it, double* out, double* in, long long* cyc) {
  for (int wps = 1; wps <= 2; wps *= 2) {
    k<MODE><<<148, 128 * wps>>>(out, in, cyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    printf("%-60s warps/scheduler %d: %.2f cycles per item\n", name, wps, avg / ((double)N_IT * per_it));
  }
}

int main() {
  double *out, *in;
  long long* cyc;
  cudaMalloc(&out, 148 * 512 * 8);
  cudaMalloc(&in, 2048 * 8);
  cudaMemset(in, 0, 2048 * 8);
  cudaMalloc(&cyc, 148 * 8);
  run<0>("dependent DFMA (latency)", 8, out, in, cyc);
  run<1>("dependent DADD (latency)", 8, out, in, cyc);
  run<2>("dependent 64-bit xor exchange = 2 SHFL (latency)", 8, out, in, cyc);
  run<3>("independent 64-bit xor exchange = 2 SHFL (throughput)", 8, out, in, cyc);
  run<4>("exchange -> DFMA round trip", 4, out, in, cyc);
  run<5>("independent exchange + DFMA pairs", 8, out, in, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
