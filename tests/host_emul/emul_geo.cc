// Debug/test harness: runs the product's __host__ __device__ geodesic code on the CPU so that kernel
// logic can be compared with the oracle without a GPU.  NOT part of the product (the product .so only
// launches CUDA kernels); built by tests/test_host_emul.py.
#include "../../grtrans_b200/csrc/grt_geokerr.cuh"
extern "C" {
// rays: arrays of n; out arrays [n][cap]
int emul_geo(int n, int cap, int nup, int kext, int next, int mufill, int usegeor, double a, double mu0, double u0,
             double uout, double offset, const double *alpha, const double *beta, const double *q2, const double *l,
             const double *su, const double *sm, const double *uf, const double *muf, const int *tpm, const int *tpr,
             int *npts, int *ncase, double *u, double *mu, double *dt, double *dphi, double *lam, int *tpmo, int *tpro) {
  grt::GeoFlags fl{nup, kext, next, mufill, usegeor, 1};
#pragma omp parallel for schedule(dynamic, 8)
  for (int i = 0; i < n; i++) {
    grt::RayConst rc;
    rc.a = a; rc.mu0 = mu0; rc.u0 = u0; rc.uout = uout; rc.alpha = alpha[i]; rc.beta = beta[i]; rc.q2 = q2[i];
    rc.l = l[i]; rc.su = su[i]; rc.sm = sm[i]; rc.uf = uf[i]; rc.muf = muf[i]; rc.offset = offset; rc.phi0 = 0;
    rc.t0 = 0; rc.tpm = tpm[i]; rc.tpr = tpr[i];
    grt::RayInv inv;
    grt::geo_setup(rc, fl, inv);
    npts[i] = inv.npts;
    ncase[i] = inv.ncase;
    for (int s = 0; s < cap; s++) {
      grt::GeoPoint p;
      bool ok = grt::geo_point(inv, nup, s, s < inv.npts, p);
      size_t q = (size_t)i * cap + s;
      if (!ok) { u[q] = mu[q] = dt[q] = dphi[q] = lam[q] = 0; tpmo[q] = tpro[q] = 0; continue; }
      u[q] = p.u; mu[q] = p.mu; dt[q] = p.dt; dphi[q] = p.dphi; lam[q] = p.lam; tpmo[q] = p.tpm; tpro[q] = p.tpr;
    }
  }
  return 0;
}
}
