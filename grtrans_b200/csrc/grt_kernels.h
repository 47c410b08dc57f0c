// grtrans_b200 -- launchers of the sm_100a kernels, one translation unit per pipeline stage:
//   grt_k_geo.cu       k_ray_setup / k_geo_points / k_ray_finalize   (FP64, FMA contraction on)
//   grt_k_fluid.cu     k_fluid_ortho                                  (compiled -fmad=false: the float32 index /
//                                                                      weight / metric arithmetic of the reference's
//                                                                      real(4) code must round exactly like unfused
//                                                                      x86 code so the gathered cells are the same)
//   grt_k_transfer.cu  k_transfer + unit-surface kernels              (FP64, FMA contraction on)
// The session / C ABI (grt_engine.cu) only sees these plain functions.
#pragma once
#include <cuda_runtime.h>

#include "grt_pipeline.cuh"

namespace grt {

void launch_ray_setup(const RayConst *rc, GeoFlags fl, RayInv *inv, int n, cudaStream_t st);
void launch_geo_points(const RayInv *inv, int nup, int maxpts, long long total, PointArrays pa, cudaStream_t st);
void launch_ray_finalize(const RayInv *inv, int maxpts, int usegeor, int n, PointArrays pa, RayMeta *meta, cudaStream_t st);
void launch_fluid_ortho(const RayConst *rc, const RayInv *inv, const RayMeta *meta, const FluidKernelArgs &A, long long total,
                        int nrays, PointArrays pa, double *rec, StageDump sd, cudaStream_t st);
// image layout: [cam][ray_local][nvals] float32, cam = cam0 + m*nfreq + k
void launch_transfer(const RayMeta *meta, const TransferParams &P, const double *lbh_m, const double *freqs, int nfreq,
                     int nparams, int nrays, int maxpts, const double *rec, float *image, long long ray_off, long long nlocal,
                     int cam0, cudaStream_t st);

// extra = 1: image layout [cam][ray_local][nvals + nextra] (nextra = 19, or 7 for nup = 1); xrec [nparams][nrays][maxpts][8]; hist scratch [maxpts][nrays*nparams*nfreq]
void launch_transfer_extra(const RayMeta *meta, const TransferParams &P, const double *lbh_m, const double *freqs, int nfreq,
                           int nparams, int nrays, int maxpts, const double *rec, const double *xrec, double *hist, float *image,
                           long long ray_off, long long nlocal, int cam0, int nextra, cudaStream_t st);

// debug = 1: per-point internals of ray 0 of a one-ray tile for the geodebug.out dump (grt_debug.cuh): rows [GD_NROWS][maxpts],
// scal[2] = fac, r%npts; hist = scratch [maxpts][4]
void launch_geodebug(const RayConst *rc, const RayMeta *meta, int maxpts, PointArrays pa, StageDump sd, const TransferParams &P,
                     const double *lbh_m, const double *rec, double nu_obs, double *hist, double *rows, double *scal,
                     cudaStream_t st);

// multi-GPU gather: scatter one block [ncamb][n][nv] of rank r (local rays off..off+n-1) into rank 0's full image [ncams][npix][nv]
void launch_gather_scatter(const float *blk, float *full, int ncamb, int n, int nv, long long npix, int cam0, long long off,
                           int r, int nranks, int i1, cudaStream_t st);

// unit-level surfaces
void launch_polsynchth(int n, const double *nu, const double *nc, const double *b, const double *t, const double *th, double *e);
void launch_polsynchpl(int n, EmisTables tab, const double *nu, const double *nnth, const double *b, const double *th,
                       const double *p, const double *gmin, double gmax, double *e);
void launch_synchemis(int n, const double *nu, const double *nc, const double *b, const double *t, double *e);
void launch_rt_integrate(int iflag, int neq, int nrays, int npts, const double *s, const double *j, const double *K,
                         const double *tau, double *intensity, int *nptsout, const TransferParams &P);
void launch_dfma(double *out, int blocks, int threads, int iters);
void launch_math_probe(int which, int n, const double *x, double *out);

}  // namespace grt
