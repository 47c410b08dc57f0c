// grtrans_b200 -- emissivity + polarised transfer kernels and the unit-surface kernels.
//   k_transfer       thread / (ray, mdot, nu)  emissivity + DELO/FORMAL/quadrature/RK march -> float32 Stokes image
#include <algorithm>
#include <cstdlib>

#include "grt_kernels.h"

namespace grt {

// image layout: [cam][ray_local][nvals] float32, cam = (mu_index*nparams + m)*nfreq + k.
// Thread = (ray, mdot, frequency), frequency fastest: the 32 lanes of a warp march the SAME ray at 32 frequencies, so the
// 128-byte point record is one broadcast load per step and the lanes never diverge on npts.
// 128 registers, 16 warps / SM (4 per scheduler), one warp per block.  Measured on the headline kernel <POLSYNCHTH, DELO>
// (config 5, ms per step): blocks of 128 threads 578.6, of 32 threads 569.9; 96 registers (5 warps per scheduler, spills)
// 664, with the coefficients of node k+1 kept in shared memory (GRT_TRANSFER_CSM=1) 623.6; 80 registers 774.
// A dependent DFMA issues every 8.7 cycles (tools/ubench/dfma_latency.cu), so four warps of pure dependency chain could
// keep the pipe 88 % busy; the kernel reaches 65-68 %.  Also measured and dropped: producer / consumer warp pairs (one warp
// evaluates the coefficients of node k into a shared-memory ring, its partner does the DELO steps; named barriers): 896 ms at
// 128 registers (8 pairs / SM), 932 ms at 96, 1123 ms at 80 -- the step time follows the number of (ray, frequency) marches
// in flight per SM, which pairing halves; straight-line reciprocals (MUFU.RCP64H + Newton, no slow-path branch) 592 ms.
#ifndef GRT_TRANSFER_MINB
#define GRT_TRANSFER_MINB 16
#endif
#ifndef GRT_TRANSFER_BS
#define GRT_TRANSFER_BS 32
#endif
#ifndef GRT_TRANSFER_CSM
#define GRT_TRANSFER_CSM 0
#endif
template <int ET, int MODE, bool STAGE = false>
__global__ void __launch_bounds__(GRT_TRANSFER_BS, GRT_TRANSFER_MINB) k_transfer(const RayMeta *__restrict__ meta, TransferParams P, const double *__restrict__ lbh_m,
                                                  const double *__restrict__ freqs, int nfreq, int nparams, int nrays, int maxpts,
                                                  const double *__restrict__ rec, float *__restrict__ image, long long ray_off,
                                                  long long nlocal, int cam0) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)nrays * nparams * nfreq;
  if (idx >= total) return;
  int f = (int)(idx % nfreq);
  int mm = (int)((idx / nfreq) % nparams);
  int ray = (int)(idx / ((long long)nfreq * nparams));
  RayMeta m = meta[ray];
  constexpr int neq = (MODE == MODE_QUAD || MODE == MODE_RK1) ? 1 : 4;
  double out[4] = {0, 0, 0, 0};
  if (m.status == 1 && m.npts > 0) {
    P.LBH = lbh_m[mm];
    const double *rr = rec + ((size_t)mm * nrays * maxpts + (size_t)ray * maxpts) * record_doubles(ET);
    if (STAGE) {  // (launched only when nfreq is a multiple of 32: the lanes of a warp share ray and mdot)
      __shared__ __align__(16) double stage[GRT_TRANSFER_BS / 32][GRT_STAGE_DEPTH * record_doubles(ET)];
      if (GRT_TRANSFER_CSM && MODE == MODE_DELO) {
        __shared__ double csm[GRT_TRANSFER_BS / 32][11 * GRT_CSM_STRIDE];
        ray_transfer_t<ET, MODE, 0, 2>(P, rr, m.npts, freqs[f], out, nullptr, 0, nullptr, stage[threadIdx.x >> 5],
                                       threadIdx.x & 31, &csm[threadIdx.x >> 5][threadIdx.x & 31]);
      } else {
        ray_transfer_t<ET, MODE, 0, 1>(P, rr, m.npts, freqs[f], out, nullptr, 0, nullptr, stage[threadIdx.x >> 5],
                                       threadIdx.x & 31);
      }
    } else {
      ray_transfer_t<ET, MODE>(P, rr, m.npts, freqs[f], out);
    }
  }
  int cam = cam0 + mm * nfreq + f;
  float *dst = image + ((size_t)cam * nlocal + (ray_off + ray)) * neq;
  if (neq == 4)
    *reinterpret_cast<float4 *>(dst) = make_float4((float)out[0], (float)out[1], (float)out[2], (float)out[3]);
  else
    dst[0] = (float)out[0];
}

// Gather: one block [ncamb][n][nv] of rank r's image (local rays off .. off+n-1) -> the full image
// [ncams][npix][nv] of rank 0.  Rank r owns the 256-ray shards r, r+nranks, ... of the index range starting at i1
// (grt_shard_rays).  One thread per float: pure HBM traffic (4 B read + 4 B write), coalesced over 256-ray runs.
__global__ void k_gather_scatter(const float *__restrict__ blk, float *__restrict__ full, int ncamb, int n, int nv, long long npix,
                                 int cam0, long long off, int r, int nranks, int i1) {
  const long long per_cam = (long long)n * nv, total = per_cam * ncamb;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int cam = (int)(idx / per_cam);
    const long long rem = idx - (long long)cam * per_cam;
    const long long q = off + rem / nv;
    const int k = (int)(rem % nv);
    const long long pix = (long long)(i1 - 1) + ((q / 256) * nranks + r) * 256 + q % 256;
    full[((size_t)(cam0 + cam) * npix + pix) * nv + k] = blk[idx];
  }
}
void launch_gather_scatter(const float *blk, float *full, int ncamb, int n, int nv, long long npix, int cam0, long long off,
                           int r, int nranks, int i1, cudaStream_t st) {
  const long long total = (long long)ncamb * n * nv;
  if (total <= 0) return;
  const int B = 256;
  const unsigned grid = (unsigned)std::min<long long>((total + B - 1) / B, 148LL * 16);  // a few blocks per SM, grid-stride
  k_gather_scatter<<<grid, B, 0, st>>>(blk, full, ncamb, n, nv, npix, cam0, off, r, nranks, i1);
}

// small stand-alone kernels for the unit-level C ABI
__global__ void k_polsynchth(int n, const double *nu, const double *nc, const double *b, const double *t, const double *th, double *e) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  polsynchth_eval(nu[i], nc[i], b[i], t[i], th[i], e + 11 * (size_t)i);
}
__global__ void k_polsynchpl(int n, EmisTables tab, const double *nu, const double *nnth, const double *b, const double *th,
                             const double *p, const double *gmin, double gmax, double *e) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  synchpl_eval<true>(tab, nu[i], nnth[i], b[i], th[i], p[i], gmin[i], gmax, e + 11 * (size_t)i);
}
__global__ void k_synchemis(int n, const double *nu, const double *nc, const double *b, const double *t, double *e) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  synchemis_eval(nu[i], nc[i], b[i], t[i], e + 11 * (size_t)i);
}

// generic-coefficient integration (radtrans_integrate f2py surface): one thread per ray
__global__ void k_rt_integrate(int iflag, int neq, int nrays, int npts, const double *s, const double *j, const double *K,
                               const double *tau, double *intensity, int *nptsout, TransferParams P) {
  int ray = blockIdx.x * blockDim.x + threadIdx.x;
  if (ray >= nrays) return;
  const int nk = 1 + neq * (neq - 1) / 2;
  const double *ss = s + (size_t)ray * npts, *jj = j + (size_t)ray * npts * neq, *KK = K + (size_t)ray * npts * nk,
               *tt = tau + (size_t)ray * npts;
  double *I = intensity + (size_t)ray * npts * neq;
  for (int i = 0; i < npts * neq; i++) I[i] = 0.0;
  auto load = [&](int k, Coef &c) {  // k 0-based node
    for (int i = 0; i < 4; i++) c.j[i] = (i < neq) ? jj[(size_t)k * neq + i] : 0.0;
    for (int i = 0; i < 7; i++) c.K[i] = (i < nk) ? KK[(size_t)k * nk + i] : 0.0;
  };
  if (iflag != 0 && neq == 4) {
    double Ip[4] = {0, 0, 0, 0}, In[4];
    Coef c0, c1;
    load(npts - 1, c1);
    for (int k = npts - 2; k >= 0; k--) {
      load(k, c0);
      double dx = ss[k] - ss[k + 1];
      if (iflag == 1)
        delo_step(c0, c1, dx, tt[k + 1] - tt[k], P.thin, Ip, In);
      else {
        double v[4];
        for (int i = 0; i < 4; i++) v[i] = c0.j[i] * dx + Ip[i];
        formal_apply(c0.K, dx, v, In);
      }
      int col = npts - 1 - k;
      for (int i = 0; i < 4; i++) {
        Ip[i] = In[i];
        I[(size_t)col * 4 + i] = In[i];
      }
      c1 = c0;
    }
    nptsout[ray] = npts;
    return;
  }
  if (iflag != 0) {
    double acc = 0.0, yprev = jj[0] * exp(-tt[0]);
    I[0] = 0.0;
    for (int k = 1; k < npts; k++) {
      double y = jj[(size_t)k * neq] * exp(-tt[k]);
      acc += (ss[k] - ss[k - 1]) * (yprev + y) / 2.0;
      I[(size_t)k * neq] = -acc;
      yprev = y;
    }
    nptsout[ray] = npts;
    return;
  }
  // lsoda replacement
  int lamdex = npts, i1 = 1, i2 = 0;
  double maxtau = tt[0];
  for (int k = 1; k < npts; k++) maxtau = dmax(maxtau, tt[k]);
  if (maxtau > P.max_tau) {
    for (int k = 0; k < npts; k++)
      if (tt[k] > P.max_tau) {
        lamdex = k + 1;
        break;
      }
  }
  if (jj[0] == 0.0)
    for (int ii = 2; ii <= npts; ii++)
      if (jj[(size_t)(ii - 1) * neq] != 0.0) {
        i1 = ii;
        break;
      }
  if (jj[(size_t)(lamdex - 1) * neq] == 0.0) {
    for (int ii = 1; ii <= lamdex - 1; ii++)
      if (jj[(size_t)(lamdex - ii - 1) * neq] != 0.0) {
        i2 = lamdex - ii;
        break;
      }
  } else
    i2 = lamdex;
  nptsout[ray] = i2;
  if (i2 <= i1) return;
  double y[4] = {0, 0, 0, 0}, h = P.hmax;
  Coef ca, cb;
  bool ok = true;
  for (int m = 1; m <= i2 - i1; m++) {
    int indx = i2 - m + 1;
    load(indx - 1, ca);
    load(((indx + 1 <= npts) ? indx + 1 : npts) - 1, cb);
    double la = ss[indx - 1], lb = ss[indx - 2];
    if (lb > la) ok = ck_interval(neq, ca, cb, la, lb, y, h, P);
    if (!ok) break;
    for (int i = 0; i < neq; i++) I[(size_t)(i1 - 1 + m) * neq + i] = y[i];
  }
  if (!ok)
    for (int i = 0; i < npts * neq; i++) I[i] = 0.0;
}

// the hot loop's own elementary functions evaluated on the device, for the accuracy tests (0: exp_fast, 1: log_ge1)
__global__ void k_math_probe(int which, int n, const double *x, double *out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = (which == 0) ? exp_fast(x[i]) : log_ge1(x[i]);
}
void launch_math_probe(int which, int n, const double *x, double *out) { k_math_probe<<<(n + 127) / 128, 128>>>(which, n, x, out); }

// FP64 FMA peak micro-benchmark: 8 independent chains per thread
__global__ void k_dfma(double *out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

template <int ET>
static void launch_transfer_e(int mode, unsigned grid, cudaStream_t st, const RayMeta *meta, const TransferParams &P,
                              const double *lbh_m, const double *freqs, int nfreq, int nparams, int nrays, int maxpts,
                              const double *rec, float *image, long long ray_off, long long nlocal, int cam0) {
  const int B = GRT_TRANSFER_BS;
#define GRT_LAUNCH(M) \
  k_transfer<ET, M><<<grid, B, 0, st>>>(meta, P, lbh_m, freqs, nfreq, nparams, nrays, maxpts, rec, image, ray_off, nlocal, cam0)
  // GRT_TRANSFER_DSMEM (experiments only): dynamic shared memory per block, to lower the number of resident warps
  static const int dsmem = getenv("GRT_TRANSFER_DSMEM") ? atoi(getenv("GRT_TRANSFER_DSMEM")) : 0;
#define GRT_LAUNCH_ST(M) \
  k_transfer<ET, M, true><<<grid, B, dsmem, st>>>(meta, P, lbh_m, freqs, nfreq, nparams, nrays, maxpts, rec, image, ray_off, nlocal, cam0)
  // warp-staged records (see stage_issue): every warp must hold 32 frequencies of one ray, and the ring holds records of at
  // most 64 doubles (GRT_TRANSFER_STAGE=0 switches it off)
  const char *se = getenv("GRT_TRANSFER_STAGE");  // (read per launch: the tests switch it inside one process)
  const int stage_env = se ? atoi(se) : 1;
  if constexpr (record_doubles(ET) <= 64) {
    if (stage_env && nfreq % 32 == 0 && (mode == MODE_DELO || mode == MODE_FORMAL)) {
      if (mode == MODE_DELO)
        GRT_LAUNCH_ST(MODE_DELO);
      else
        GRT_LAUNCH_ST(MODE_FORMAL);
      return;
    }
  }
  switch (mode) {
    case MODE_DELO: GRT_LAUNCH(MODE_DELO); break;
    case MODE_FORMAL: GRT_LAUNCH(MODE_FORMAL); break;
    case MODE_QUAD: GRT_LAUNCH(MODE_QUAD); break;
    case MODE_RK4: GRT_LAUNCH(MODE_RK4); break;
    default: GRT_LAUNCH(MODE_RK1); break;
  }
#undef GRT_LAUNCH
}

void launch_transfer(const RayMeta *meta, const TransferParams &P, const double *lbh_m, const double *freqs, int nfreq,
                     int nparams, int nrays, int maxpts, const double *rec, float *image, long long ray_off, long long nlocal,
                     int cam0, cudaStream_t st) {
  const int B = GRT_TRANSFER_BS;
  long long tt = (long long)nrays * nparams * nfreq;
  unsigned grid = (unsigned)((tt + B - 1) / B);
  const int mode = transfer_mode(P);
  switch (em_slot(P.etype)) {
#define OP(E) \
  case E: launch_transfer_e<E>(mode, grid, st, meta, P, lbh_m, freqs, nfreq, nparams, nrays, maxpts, rec, image, ray_off, nlocal, cam0); break;
    GRT_FOR_ETYPES(OP)
#undef OP
    default: break;
  }
}
void launch_polsynchth(int n, const double *nu, const double *nc, const double *b, const double *t, const double *th, double *e) {
  const int B = 128;
  k_polsynchth<<<(n + B - 1) / B, B>>>(n, nu, nc, b, t, th, e);
}
void launch_polsynchpl(int n, EmisTables tab, const double *nu, const double *nnth, const double *b, const double *th,
                       const double *p, const double *gmin, double gmax, double *e) {
  const int B = 128;
  k_polsynchpl<<<(n + B - 1) / B, B>>>(n, tab, nu, nnth, b, th, p, gmin, gmax, e);
}
void launch_synchemis(int n, const double *nu, const double *nc, const double *b, const double *t, double *e) {
  const int B = 128;
  k_synchemis<<<(n + B - 1) / B, B>>>(n, nu, nc, b, t, e);
}
void launch_rt_integrate(int iflag, int neq, int nrays, int npts, const double *s, const double *j, const double *K,
                         const double *tau, double *intensity, int *nptsout, const TransferParams &P) {
  const int B = 64;
  k_rt_integrate<<<(nrays + B - 1) / B, B>>>(iflag, neq, nrays, npts, s, j, K, tau, intensity, nptsout, P);
}
void launch_dfma(double *out, int blocks, int threads, int iters) { k_dfma<<<blocks, threads>>>(out, iters); }

}  // namespace grt
