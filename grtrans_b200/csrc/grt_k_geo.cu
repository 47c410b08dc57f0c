// grtrans_b200 -- geodesic kernels (SURVEY.md 8a rows R2, R3 bookkeeping).
//   k_ray_setup      thread / ray            geo_setup()   -> RayInv  (per-ray invariants: roots, NCASE, u0/mu0 integrals)
//   k_geo_points     thread / (ray, point)   geo_point()   -> u, mu, dt, dphi, lambda, tpm, tpr
//   k_ray_finalize   thread / ray            trim / status / lambda_N, t_1
#define GRT_USE_PHASE_BARRIER  // see geo_point (grt_geokerr.cuh)
#include "grt_kernels.h"

namespace grt {

__global__ void __launch_bounds__(128) k_ray_setup(const RayConst *__restrict__ rc, GeoFlags fl, RayInv *__restrict__ inv, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  RayConst r = rc[i];
  RayInv c;
  geo_setup(r, fl, c);
  inv[i] = c;
}

template <int BS, int MINB>
__global__ void __launch_bounds__(BS, MINB) k_geo_points(const RayInv *__restrict__ inv, int nup, int maxpts, long long total, PointArrays pa) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = idx < total;  // no early exit: every thread of the block takes part in the phase barriers
  if (!in_range) idx = total - 1;
  int ray = (int)(idx / maxpts), slot = (int)(idx % maxpts);
  point_store(inv[ray], nup, slot, (size_t)idx, pa, in_range);
}

__global__ void __launch_bounds__(128) k_ray_finalize(const RayInv *__restrict__ inv, int maxpts, int usegeor, int n, PointArrays pa,
                               RayMeta *__restrict__ meta) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  RayMeta m;
  ray_finalize(inv[i].npts, (size_t)i, maxpts, usegeor, pa, m);
  meta[i] = m;
}

void launch_ray_setup(const RayConst *rc, GeoFlags fl, RayInv *inv, int n, cudaStream_t st) {
  const int B = 128;
  k_ray_setup<<<(n + B - 1) / B, B, 0, st>>>(rc, fl, inv, n);
}
void launch_geo_points(const RayInv *inv, int nup, int maxpts, long long total, PointArrays pa, cudaStream_t st) {
  // 1024-thread blocks (one per SM, 64 registers): measured best of {128x6, 256x3, 384x2, 512x1, 768x1, 1024x1} -- the phase
  // barriers of geo_point work per block, so one big block keeps all 32 warps of the SM inside the same code stage
  // (nx=256 bench: 28.3 / 22.9 / 20.5 / 23.3 / 19.5 / 18.3 ms; 59.6 ms before the code was compacted and phased)
  // Round 2, measured and dropped (config 5, ms per step for this stage; gpurun_out/r2_geo.jsonl): R_F, R_D and the R_J's of
  // a radial piece in ONE duplication loop that shares the three square roots per iteration (each integral with its own
  // sum, convergence test and closing series; identical image): 235.0 against 210.0 -- the fused loop's state (three R_J
  // arguments, sums and principal-value terms) spills at 64 registers; with 768-thread blocks / 80 registers 235.9, with
  // 512-thread blocks / 128 registers 267.1.  The stage scales with the number of resident warps, not with instructions.
  // Round 2: two 512-thread blocks per SM (same 32 warps at 64 registers; the phase barriers synchronise 16 warps instead of
  // 32): 203.9 ms per config-5 step against 211.2 with one 1024-thread block; 2 x 576 (56 registers) 204.2, 3 x 320 255.1,
  // 4 x 256 261.9 (gpurun_out/r2_geo2.jsonl).  All three phase barriers of geo_point earn their place at this block size:
  // without the one before the mu integrals 204.3 ms, without the one after the u integrals 221.2, without the first 229.4.
#ifndef GRT_GEO_BS
#define GRT_GEO_BS 512
#endif
#ifndef GRT_GEO_MINB
#define GRT_GEO_MINB 2
#endif
  constexpr int BS = GRT_GEO_BS;
  k_geo_points<BS, GRT_GEO_MINB><<<(unsigned)((total + BS - 1) / BS), BS, 0, st>>>(inv, nup, maxpts, total, pa);
}
void launch_ray_finalize(const RayInv *inv, int maxpts, int usegeor, int n, PointArrays pa, RayMeta *meta, cudaStream_t st) {
  const int B = 128;
  k_ray_finalize<<<(n + B - 1) / B, B, 0, st>>>(inv, maxpts, usegeor, n, pa, meta);
}

}  // namespace grt
