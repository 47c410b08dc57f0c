// Micro-benchmark: does a DFMA cost one or two issue slots of its scheduler?  8 independent DFMA chains per thread (enough to
// saturate the FP64 pipe) interleaved with K independent integer multiply-adds per DFMA (K = 0, 1, 2, 3); 16 warps per SM.
// If the FP64 instruction only occupies its pipe for two cycles, K = 1 is free (the integer op takes the other cycle); if it
// holds the issue port, every added instruction adds time.   nvcc -arch=sm_100a -O3 -o dfma_issue dfma_issue.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int K>
__global__ void k(double *out, int *iout, int iters) {
  double a[8];
  int q[8];
  for (int c = 0; c < 8; c++) { a[c] = threadIdx.x * 1e-9 + c; q[c] = threadIdx.x + c; }
  const double b = 1.0000001, d = 1e-9;
  const int m = 1664525, n = 1013904223;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < 8; c++) {
      a[c] = fma(a[c], b, d);
#pragma unroll
      for (int j = 0; j < K; j++) q[(c + j) & 7] = q[(c + j) & 7] * m + n;
    }
  }
  double s = 0; int t = 0;
  for (int c = 0; c < 8; c++) { s += a[c]; t += q[c]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  iout[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
template <int K>
void run(double *out, int *iout) {
  const int nsm = 148, threads = 512, iters = 100000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<K><<<nsm, threads>>>(out, iout, 1000);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<K><<<nsm, threads>>>(out, iout, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  // per scheduler: 4 warps x 8 DFMA x iters warp-instructions
  double dfma_per_sched = 4.0 * 8 * iters;
  double cycles = ms * 1e-3 * 1.92e9;
  printf("K=%d integer ops per DFMA: %.3f ms, %.2f cycles per DFMA per scheduler (%.2f TFLOP/s)\n", K, ms, cycles / dfma_per_sched,
         2.0 * nsm * threads * 8.0 * iters / (ms * 1e-3) / 1e12);
}
int main() {
  double *out; int *iout;
  cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&iout, 148 * 512 * 4);
  run<0>(out, iout); run<1>(out, iout); run<2>(out, iout); run<3>(out, iout);
  return 0;
}
