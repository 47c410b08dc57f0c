// Micro-benchmark: FP64 FMA throughput as a function of resident warps per scheduler and independent chains per thread
// (how many warps x chains the DFMA pipe of one SM sub-partition needs to stay busy).  nvcc -arch=sm_100a -O3 -o dfma_latency ...
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(double *out, int iters) {
  double a[CH];
  for (int c = 0; c < CH; c++) a[c] = threadIdx.x * 1e-9 + c;
  const double b = 1.0000001, d = 1e-9;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < CH; c++) a[c] = fma(a[c], b, d);
  }
  double s = 0;
  for (int c = 0; c < CH; c++) s += a[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
void run(int warps_per_sm, double *out) {
  int nsm = 148;
  int iters = 200000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  // one block of 32*warps_per_sm threads per SM
  k<CH><<<nsm, 32 * warps_per_sm>>>(out, 1000);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<CH><<<nsm, 32 * warps_per_sm>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double fma_per_s = (double)nsm * 32 * warps_per_sm * CH * iters / (ms * 1e-3);
  // cycles per dependent FMA of one chain: ms * clock / iters
  printf("chains %d warps/SM %2d (%.2f per scheduler): %.2f TFLOP/s  %.3f ns per dependent step\n", CH, warps_per_sm,
         warps_per_sm / 4.0, 2 * fma_per_s / 1e12, ms * 1e6 / iters / 1.0);
}
int main() {
  double *out;
  cudaMalloc(&out, 148 * 1024 * 8);
  for (int w : {4, 8, 12, 16, 20, 24, 32}) run<1>(w, out);
  for (int w : {4, 8, 16, 20}) run<2>(w, out);
  for (int w : {4, 8, 16}) run<4>(w, out);
  return 0;
}
