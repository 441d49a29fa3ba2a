// dfma_operands.cu -- how fast does ONE warp per scheduler issue DFMAs, as a function of the register operands read?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_operands dfma_operands.cu ; run: ./dfma_operands
// Each variant runs N_IT iterations of 16 independent accumulator updates; prints cycles per warp-DFMA per scheduler
// for 1, 2 and 4 warps per scheduler (CTA = 128 threads x {1,2,4}, one CTA per SM).
#include <cstdio>
#include <cuda_runtime.h>

#define N_IT 4096
#define NA 16

template <int MODE>
__global__ void k(double* out, const double* in, long long* cyc) {
  double a[NA], b[NA], c[NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) { a[i] = in[i + threadIdx.x]; b[i] = in[NA + i + threadIdx.x]; c[i] = in[2 * NA + i + threadIdx.x]; }
  const double s0 = in[100 + threadIdx.x], s1 = in[200 + threadIdx.x];
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < N_IT; ++it) {
    if (MODE == 0) {          // c_i = fma(c_i, s0, s1): one changing register operand, two loop-invariant ones
#pragma unroll
      for (int i = 0; i < NA; ++i) c[i] = fma(c[i], s0, s1);
    } else if (MODE == 1) {   // c_i = fma(a_i, b_i, c_i): three distinct register operands
#pragma unroll
      for (int i = 0; i < NA; ++i) c[i] = fma(a[i], b[i], c[i]);
    } else if (MODE == 2) {   // c_i = fma(s0, b_i, c_i): one operand shared by consecutive instructions
#pragma unroll
      for (int i = 0; i < NA; ++i) c[i] = fma(s0, b[i], c[i]);
    } else if (MODE == 3) {   // pairs sharing one operand: (a_i, b_i), (a_i, b_{i+1})  -- the complex-MAC pattern
#pragma unroll
      for (int i = 0; i < NA; i += 2) { c[i] = fma(a[i], b[i], c[i]); c[i + 1] = fma(a[i], b[i + 1], c[i + 1]); }
    } else if (MODE == 4) {   // groups of four sharing one operand
#pragma unroll
      for (int i = 0; i < NA; i += 4) {
        c[i] = fma(a[i], b[i], c[i]); c[i + 1] = fma(a[i], b[i + 1], c[i + 1]);
        c[i + 2] = fma(a[i], b[i + 2], c[i + 2]); c[i + 3] = fma(a[i], b[i + 3], c[i + 3]);
      }
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NA; ++i) s += c[i] + a[i] + b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, double* out, double* in, long long* cyc) {
  for (int wps = 1; wps <= 4; wps *= 2) {
    k<MODE><<<148, 128 * wps>>>(out, in, cyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    printf("%-44s warps/scheduler %d: %.2f cycles per DFMA per scheduler\n", name, wps, avg / ((double)N_IT * NA * wps));
  }
}

int main() {
  double *out, *in;
  long long* cyc;
  cudaMalloc(&out, 148 * 512 * 8);
  cudaMalloc(&in, 2048 * 8);
  cudaMemset(in, 0, 2048 * 8);
  cudaMalloc(&cyc, 148 * 8);
  run<0>("fma(c, inv, inv)  [1 changing operand]", out, in, cyc);
  run<1>("fma(a_i, b_i, c_i) [3 distinct]", out, in, cyc);
  run<2>("fma(s, b_i, c_i)   [1 shared by all]", out, in, cyc);
  run<3>("pairs sharing a_i", out, in, cyc);
  run<4>("fours sharing a_i", out, in, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
