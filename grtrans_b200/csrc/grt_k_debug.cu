// grtrans_b200 -- debug = 1 (`geodebug.out`): one block over the points of the one ray of a one-ray tile (grt_debug.cuh).
#include "grt_kernels.h"
#include "grt_debug.cuh"

namespace grt {

__global__ void k_geodebug(const RayConst *__restrict__ rc, const RayMeta *__restrict__ meta, int maxpts, PointArrays pa,
                           StageDump sd, TransferParams P, const double *__restrict__ lbh_m, const double *__restrict__ rec,
                           double nu_obs, double *hist, double *rows, double *scal) {
  const RayMeta m = meta[0];
  const int npts = (m.status == 1) ? m.npts : 0;
  for (int p = threadIdx.x; p < npts; p += blockDim.x) geodebug_point(rc[0], m, maxpts, pa, sd, p, rows);
  if (threadIdx.x == 0) {
    scal[0] = 0.0;
    scal[1] = 1.0;
    if (npts > 0) {
      P.LBH = lbh_m[0];
      geodebug_ray(P, rec, npts, maxpts, nu_obs, hist, rows, scal);
    }
  }
}

void launch_geodebug(const RayConst *rc, const RayMeta *meta, int maxpts, PointArrays pa, StageDump sd, const TransferParams &P,
                     const double *lbh_m, const double *rec, double nu_obs, double *hist, double *rows, double *scal,
                     cudaStream_t st) {
  k_geodebug<<<1, 128, 0, st>>>(rc, meta, maxpts, pa, sd, P, lbh_m, rec, nu_obs, hist, rows, scal);
}

}  // namespace grt
