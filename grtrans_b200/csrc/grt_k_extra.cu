// grtrans_b200 -- extra = 1: Stokes values plus the 19 per-pixel diagnostics of grtrans_driver.f90:230-292.
//   k_extra   thread / (ray, mdot, nu)   the march of k_transfer with its intensity history kept in a scratch array, then
//                                        one forward pass over the ray for the depth / average integrals
// A diagnostic mode: its own translation unit so that the hot kernels of grt_k_transfer.cu are untouched.
#include "grt_kernels.h"

namespace grt {

// image layout: [cam][ray_local][nvals + nextra] float32, nextra = 19 (rays of several points) or 7 (nup = 1)
template <int ET, int MODE>
__global__ void __launch_bounds__(128) k_extra(const RayMeta *__restrict__ meta, TransferParams P, const double *__restrict__ lbh_m,
                                               const double *__restrict__ freqs, int nfreq, int nparams, int nrays, int maxpts,
                                               const double *__restrict__ rec, const double *__restrict__ xrec,
                                               double *__restrict__ hist, float *__restrict__ image, long long ray_off,
                                               long long nlocal, int cam0, int nextra) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)nrays * nparams * nfreq;
  if (idx >= total) return;
  int f = (int)(idx % nfreq);
  int mm = (int)((idx / nfreq) % nparams);
  int ray = (int)(idx / ((long long)nfreq * nparams));
  RayMeta m = meta[ray];
  constexpr int neq = (MODE == MODE_QUAD || MODE == MODE_RK1) ? 1 : 4;
  double out[4] = {0, 0, 0, 0}, xt[19];
  for (int q = 0; q < 19; q++) xt[q] = 0.0;
  if (m.status == 1 && m.npts > 0) {
    P.LBH = lbh_m[mm];
    const size_t base = (size_t)mm * nrays * maxpts + (size_t)ray * maxpts;
    // history of this thread: hist[col * total + idx] (coalesced across the warp's frequencies)
    ray_extra_t<ET, MODE>(P, rec + base * record_doubles(ET), xrec + base * 8, m.npts, freqs[f], out, xt, hist + idx, (size_t)total);
  }
  int cam = cam0 + mm * nfreq + f;
  float *dst = image + ((size_t)cam * nlocal + (ray_off + ray)) * (neq + nextra);
  for (int i = 0; i < neq; i++) dst[i] = (float)out[i];
  for (int q = 0; q < nextra; q++) dst[neq + q] = (float)xt[q];
}

template <int ET>
static void launch_extra_e(int mode, unsigned grid, cudaStream_t st, const RayMeta *meta, const TransferParams &P,
                           const double *lbh_m, const double *freqs, int nfreq, int nparams, int nrays, int maxpts,
                           const double *rec, const double *xrec, double *hist, float *image, long long ray_off, long long nlocal,
                           int cam0, int nextra) {
  const int B = 128;
#define GRT_LAUNCH(M) \
  k_extra<ET, M><<<grid, B, 0, st>>>(meta, P, lbh_m, freqs, nfreq, nparams, nrays, maxpts, rec, xrec, hist, image, ray_off, nlocal, cam0, nextra)
  switch (mode) {
    case MODE_DELO: GRT_LAUNCH(MODE_DELO); break;
    case MODE_FORMAL: GRT_LAUNCH(MODE_FORMAL); break;
    case MODE_QUAD: GRT_LAUNCH(MODE_QUAD); break;
    case MODE_RK4: GRT_LAUNCH(MODE_RK4); break;
    default: GRT_LAUNCH(MODE_RK1); break;
  }
#undef GRT_LAUNCH
}

// The 60 instantiations are split over three objects (-DGRT_EXTRA_PART=0/1/2) so that the build compiles them in parallel.
#ifndef GRT_EXTRA_PART
#error "compile with -DGRT_EXTRA_PART=0, 1 or 2"
#endif
#if GRT_EXTRA_PART == 0
#define GRT_PART_ETYPES(OP) OP(EM_POLSYNCHTH) OP(EM_SYNCHTHAV) OP(EM_BB) OP(EM_BBPOL)
#define GRT_PART_NAME launch_extra_part0
#elif GRT_EXTRA_PART == 1
#define GRT_PART_ETYPES(OP) OP(EM_POLSYNCHPL) OP(EM_SYNCHPL) OP(EM_HYBRIDPL) OP(EM_MAXJUTT)
#define GRT_PART_NAME launch_extra_part1
#else
#define GRT_PART_ETYPES(OP) OP(EM_HYBRIDTHPL) OP(EM_HYBRIDTH) OP(EM_POLSYNCHSYMTH) OP(EM_BINS)
#define GRT_PART_NAME launch_extra_part2
#endif

#define GRT_EXTRA_ARGS                                                                                                     \
  int mode, unsigned grid, cudaStream_t st, const RayMeta *meta, const TransferParams &P, const double *lbh_m,             \
      const double *freqs, int nfreq, int nparams, int nrays, int maxpts, const double *rec, const double *xrec, double *hist, \
      float *image, long long ray_off, long long nlocal, int cam0, int nextra
bool launch_extra_part0(GRT_EXTRA_ARGS);
bool launch_extra_part1(GRT_EXTRA_ARGS);
bool launch_extra_part2(GRT_EXTRA_ARGS);

bool GRT_PART_NAME(GRT_EXTRA_ARGS) {
  switch (em_slot(P.etype)) {
#define OP(E) \
  case E: launch_extra_e<E>(mode, grid, st, meta, P, lbh_m, freqs, nfreq, nparams, nrays, maxpts, rec, xrec, hist, image, ray_off, nlocal, cam0, nextra); return true;
    GRT_PART_ETYPES(OP)
#undef OP
    default: return false;
  }
}

#if GRT_EXTRA_PART == 0
void launch_transfer_extra(const RayMeta *meta, const TransferParams &P, const double *lbh_m, const double *freqs, int nfreq,
                           int nparams, int nrays, int maxpts, const double *rec, const double *xrec, double *hist, float *image,
                           long long ray_off, long long nlocal, int cam0, int nextra, cudaStream_t st) {
  const int B = 128;
  long long tt = (long long)nrays * nparams * nfreq;
  unsigned grid = (unsigned)((tt + B - 1) / B);
  const int mode = transfer_mode(P);
#define GRT_TRY(F) F(mode, grid, st, meta, P, lbh_m, freqs, nfreq, nparams, nrays, maxpts, rec, xrec, hist, image, ray_off, nlocal, cam0, nextra)
  if (!GRT_TRY(launch_extra_part0) && !GRT_TRY(launch_extra_part1)) (void)GRT_TRY(launch_extra_part2);
#undef GRT_TRY
}
#endif

}  // namespace grt
