// grtrans_b200 -- fluid sampling + comoving frame + unit conversion kernel.
//   k_fluid_ortho    thread / (ray, point)   x^mu, k^mu, fluid gather, tetrad, cgs conversion -> 128-byte point record per mdot
// THIS FILE IS COMPILED WITH -fmad=false.  The reference does its cell-index, weight, metric and LNRF arithmetic in
// real(4) (fluid_model_*.f90, kerr.f90:361-379,427-500); gfortran/x86-64 code does not fuse multiply-adds, and several of
// those expressions cancel strongly (b.b with the float32 metric, u^t normalisation), so a fused fmaf changes the result by
// far more than one ulp.  Unfused float32 arithmetic + correctly rounded float32 libm calls (r4sin, r4log, ... in
// grt_math.cuh) make this stage bit-faithful to the CPU evaluation order.
#include "grt_kernels.h"

namespace grt {

#ifndef GRT_FLUID_MINB
#define GRT_FLUID_MINB 6   // measured (nx=512 step): 3 blocks/SM 42.0 ms, 4: 35.8, 5: 33.6, 6: 32.6, 8: 34.4, 10: 38.1, 12: 41.4
#endif
// Round 2, measured and dropped (config 5, this stage per step; gpurun_out/r2_fl2.jsonl): block-wide phase barriers between
// sample / tetrad / record stages as in k_geo_points (ncu shows 9-15 % "no instruction" stalls here: the hot path of one fluid
// model is ~80 KB of code): 152.8 ms with one 768-thread block per SM, 136.8 ms with two of 384, against 130.7 ms.
#ifndef GRT_FLUID_BS
#define GRT_FLUID_BS 128
#endif
__global__ void __launch_bounds__(GRT_FLUID_BS, GRT_FLUID_MINB) k_fluid_ortho(const RayConst *__restrict__ rc, const RayInv *__restrict__ inv,
                                                     const RayMeta *__restrict__ meta, const __grid_constant__ FluidKernelArgs A, long long total,
                                                     int nrays, PointArrays pa, double *__restrict__ rec, StageDump sd) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int ray = (int)(idx / A.maxpts), slot = (int)(idx % A.maxpts);
  point_record(rc[ray], meta[ray], A, ray, slot, nrays, pa, rec, sd);
}

void launch_fluid_ortho(const RayConst *rc, const RayInv *inv, const RayMeta *meta, const FluidKernelArgs &A, long long total,
                        int nrays, PointArrays pa, double *rec, StageDump sd, cudaStream_t st) {
  const int B = GRT_FLUID_BS;
  // 4 blocks / SM (128 registers): 12.0 ms vs 12.8 (3 blocks) / 12.1 (5, 6) / 12.4 (8) at nx=256
  k_fluid_ortho<<<(unsigned)((total + B - 1) / B), B, 0, st>>>(rc, inv, meta, A, total, nrays, pa, rec, sd);
}

}  // namespace grt
