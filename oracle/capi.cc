// ORACLE -- TEST INFRASTRUCTURE ONLY (see geokerr.cc header).
// Plain-C entry points so tests/ (ctypes) and bench.py's cpu_baseline leg can drive the CPU
// restatement.  Takes the same grt_args struct as the product's C ABI (interface definition only;
// nothing of the product's implementation is used here).
#include <cstring>
#include <string>

#include "../include/grtrans_b200.h"
#include "oracle.h"

using namespace orc;

static RunArgs to_run(const grt_args *a) {
  RunArgs r;
  r.standard = a->standard; r.nmu = a->nmu; r.nrotype = a->nrotype; r.nvals = a->nvals; r.nfreq = a->nfreq;
  r.nmdot = a->nmdot; r.nt = a->nt; r.nload = a->nload; r.extra = a->extra; r.i1 = a->i1; r.i2 = a->i2; r.dt = a->dt;
  r.mumin = a->mumin; r.mumax = a->mumax; r.phi0 = a->phi0; r.spin = a->spin; r.uout = a->uout; r.uin = a->uin;
  r.rcut = a->rcut;
  for (int i = 0; i < 4; i++) r.gridvals[i] = a->gridvals[i];
  for (int i = 0; i < 3; i++) r.nn[i] = a->nn[i];
  r.fname = a->fname; r.ename = a->ename; r.iname = a->iname; r.stype = a->stype;
  r.mdotmin = a->mdotmin; r.mdotmax = a->mdotmax; r.mbh = a->mbh; r.fmin = a->fmin; r.fmax = a->fmax;
  r.muval = a->muval; r.gmin = a->gmin; r.gmax = a->gmax; r.p1 = a->p1; r.p2 = a->p2; r.jetalpha = a->jetalpha;
  r.sigcut = a->sigcut; r.ftscl = a->ftscl; r.frscl = a->frscl;
  for (int i = 0; i < 7; i++) r.coefindx[i] = a->epcoefindx[i];
  r.nepotherargs = a->nepotherargs < 8 ? a->nepotherargs : 8;
  for (int i = 0; i < 8; i++) r.epotherargs[i] = a->epotherargs[i];
  r.fargs.dfile = a->fdfile; r.fargs.hfile = a->fhfile; r.fargs.gfile = a->fgfile; r.fargs.nt = a->fnt;
  r.fargs.indf = a->findf; r.fargs.mdot = a->fmdot; r.fargs.mbh = a->mbh; r.fargs.rin = a->frin; r.fargs.rout = a->frout; r.fargs.scalefac = a->fscalefac;
  return r;
}

extern "C" {

// Test-only switches (tests/test_oracle_golden.py): physical constants of an older reference commit and "legacy" code
// variants, used to pin the oracle against goldens the reference's repository still carries from older commits.
// name = "h","k","c","e","G","m","mp","sigb","msun","sigt" or a LegacyOptions field; "reset" restores everything.
int orc_set_option(const char *name, double v) {
  const std::string n(name);
  if (n == "reset") {
    pc::h = 6.626e-27; pc::k = 1.38e-16; pc::c = 2.99792458e10; pc::e = 4.8032e-10; pc::G = 6.67e-8; pc::m = 9.10938188e-28;
    pc::mp = 1.67262158e-24; pc::sigb = 5.6704e-5; pc::msun = 1.998e33; pc::sigt = 6.6523e-25;
    legacy = LegacyOptions();
  } else if (n == "h") pc::h = v;
  else if (n == "k") pc::k = v;
  else if (n == "c") pc::c = v;
  else if (n == "e") pc::e = v;
  else if (n == "G") pc::G = v;
  else if (n == "m") pc::m = v;
  else if (n == "mp") pc::mp = v;
  else if (n == "sigb") pc::sigb = v;
  else if (n == "msun") pc::msun = v;
  else if (n == "sigt") pc::sigt = v;
  else if (n == "kstarq_sign_old") legacy.kstarq_sign_old = (int)v;
  else if (n == "pl_rhoq_scale") legacy.pl_rhoq_scale = v;
  else if (n == "pl_rhov_scale") legacy.pl_rhov_scale = v;
  else if (n == "pl_jv_scale") legacy.pl_jv_scale = v;
  else if (n == "pl_av_scale") legacy.pl_av_scale = v;
  else if (n == "pl_ai_scale") legacy.pl_ai_scale = v;
  else if (n == "pl_aq_scale") legacy.pl_aq_scale = v;
  else if (n == "rt_perturb") legacy.rt_perturb = v;
  else if (n == "lambda_perturb_ulp") legacy.lambda_perturb_ulp = v;
  else if (n == "delo_inverse_lu") legacy.delo_inverse_lu = (int)v;
  else return 1;
  pc::c2 = pc::c * pc::c;
  return 0;
}

double orc_rf(double x, double y, double z) { return rf(x, y, z); }
double orc_rc(double x, double y) { return rc(x, y); }
double orc_rd(double x, double y, double z) { return rd(x, y, z); }
double orc_rj(double x, double y, double z, double p) { return rj(x, y, z, p); }
void orc_sncndn(double u, double emmc, double *sn, double *cn, double *dn) { sncndn(u, emmc, *sn, *cn, *dn); }
double orc_besselk(int n, double x) { return n == 0 ? besselk0(x) : (n == 1 ? besselk1(x) : besselkn(n, x)); }
double orc_bnu(double T, double nu) { return bnu(T, nu); }

void orc_polsynchth(int n, const double *nu, const double *nc, const double *b, const double *t, const double *th, double *e) {
  for (int i = 0; i < n; i++) polsynchth(nu[i], nc[i], b[i], t[i], th[i], e + 11 * i);
}
void orc_polsynchpl(int n, const double *nu, const double *nnth, const double *b, const double *th, const double *p,
                    const double *gmin, double gmax, double *e) {
  for (int i = 0; i < n; i++) polsynchpl(nu[i], nnth[i], b[i], th[i], p[i], gmin[i], gmax, e + 11 * i);
}
void orc_synchemis(int n, const double *nu, const double *nc, const double *b, const double *t, double *e) {
  for (int i = 0; i < n; i++) synchemis(nu[i], nc[i], b[i], t[i], e + 11 * i);
}
// synchbinemis with the bins of emis_model_bins: n[npt][nbin] electrons per unit gamma
void orc_synchbinemis(int npt, const double *nu, const double *n, int nbin, double gmin, double gmax, const double *b,
                      const double *th, double *e, double *gammas, double *dgammas) {
  std::vector<double> g, dg;
  emis_model_bins(gmin, gmax, nbin, g, dg);
  for (int i = 0; i < nbin; i++) { gammas[i] = g[i]; dgammas[i] = dg[i]; }
  for (int i = 0; i < npt; i++) synchbinemis(nu[i], n + (size_t)i * nbin, nbin, b[i], th[i], g.data(), dg.data(), e + 11 * i);
}

// one ray: intensity[npts][neq]; returns nptsout
int orc_rt_integrate(int iflag, int neq, int npts, const double *s, const double *j, const double *K, const double *tau,
                     double *intensity, double max_tau, double hmax, double oatol, double ortol, double thin,
                     int maxsteps, int as_written, int *nsteps) {
  RtOptions o;
  o.max_tau = max_tau; o.hmax = hmax; o.oatol = oatol; o.ortol = ortol; o.thin = thin; o.maxsteps = maxsteps;
  o.lsoda_neighbor_as_written = as_written != 0;
  return rt_integrate(iflag, neq, npts, s, j, K, tau, intensity, o, nsteps);
}

// camera constants
void orc_camera(int standard, double mu0, double a, double uout, double rcut, int nrotype, const double *gridvals,
                int nro, int nphi, int nup, double *u0, double *offset, double *alpha, double *beta, double *q2,
                double *l, double *su, double *sm, double *uf, double *muf, int *tpm, int *tpr) {
  GeokerrArgs g;
  initialize_pixels(g, standard, mu0, 0.0, a, uout, 1.0, rcut, nrotype, gridvals[0], gridvals[1], gridvals[2],
                    gridvals[3], nro, nphi, nup);
  int n = nro * nphi;
  *u0 = g.u0;
  *offset = g.offset;
  for (int i = 0; i < n; i++) {
    alpha[i] = g.alpha[i]; beta[i] = g.beta[i]; q2[i] = g.q2[i]; l[i] = g.l[i]; su[i] = g.su[i]; sm[i] = g.sm[i];
    uf[i] = g.uf[i]; muf[i] = g.muf[i]; tpm[i] = g.tpm[i]; tpr[i] = g.tpr[i];
  }
}

// full run.  pixvals: float32 [ncams][npix][nvals]; pixloc float32 [npix][2]; freqs[nfreq]
int orc_main(const grt_args *a, int nthreads, float *pixvals, float *pixloc, double *freqs, double *loop_seconds,
             long long *npoints) {
  RunArgs r = to_run(a);
  r.nthreads = nthreads;
  RunResult res;
  grtrans_main(r, res);
  memcpy(pixvals, res.pixvals.data(), res.pixvals.size() * sizeof(float));
  memcpy(pixloc, res.pixloc.data(), res.pixloc.size() * sizeof(float));
  for (size_t i = 0; i < res.freqs.size(); i++) freqs[i] = res.freqs[i];
  if (loop_seconds) *loop_seconds = res.loop_seconds;
  if (npoints) *npoints = res.npoints;
  return 0;
}

// stage dump for rays i1..i2 (1-based) of the first mu0: arrays [nrays][cap] (zero padded).
// e_last/K_last: emissivity arrays of the LAST frequency after rotate/invariant/scale ([.][cap][4|7]).
int orc_ray_dump(const grt_args *a, int i1, int i2, int cap, int *npts, int *status, int *ncase, double *lambda,
                 double *t, double *r, double *th, double *ph, double *kt, double *kr, double *kth, double *kph,
                 int *tpm, int *tpr, float *rho, float *p, float *bmag, double *u, double *b, double *s2xi,
                 double *c2xi, double *ang, double *rshift, double *cosne, double *ncgs, double *ncgsnth, double *bcgs,
                 double *tcgs, double *e_last, double *K_last, double *tau, double *stokes) {
  RunArgs ra = to_run(a);
  std::vector<double> freqs, mdots, mu0;
  compute_freqs(ra, freqs, mdots, mu0);
  FluidData fd;
  FluidArgs fa = ra.fargs;
  if (!load_fluid_model(fd, ra.fname, ra.spin, fa)) return 3;
  GeokerrArgs g;
  initialize_pixels(g, ra.standard, mu0[0], ra.phi0, ra.spin, ra.uout, ra.uin, ra.rcut, ra.nrotype, ra.gridvals[0],
                    ra.gridvals[1], ra.gridvals[2], ra.gridvals[3], ra.nn[0], ra.nn[1], ra.nn[2]);
  int nfreq = (int)freqs.size();
#pragma omp parallel for schedule(dynamic, 4)
  for (int i = i1; i <= i2; i++) {
    RayDump d;
    grtrans_ray_dump(ra, fd, g, i, freqs, d);
    size_t q = (size_t)(i - i1);
    npts[q] = d.npts;
    status[q] = d.status;
    ncase[q] = d.ncase;
    int n = std::min(d.npts, cap);
    if (d.status != 1) n = 0;
    auto cp = [&](double *dst, const std::vector<double> &src, int w) {
      if (!dst) return;
      for (size_t z = 0; z < (size_t)cap * w; z++) dst[q * cap * w + z] = 0.0;
      for (size_t z = 0; z < (size_t)n * w && z < src.size(); z++) dst[q * cap * w + z] = src[z];
    };
    cp(lambda, d.lambda, 1); cp(t, d.t, 1); cp(r, d.r, 1); cp(th, d.th, 1); cp(ph, d.ph, 1);
    cp(kt, d.kt, 1); cp(kr, d.kr, 1); cp(kth, d.kth, 1); cp(kph, d.kph, 1);
    cp(u, d.u, 4); cp(b, d.b, 4);
    cp(s2xi, d.s2xi, 1); cp(c2xi, d.c2xi, 1); cp(ang, d.ang, 1); cp(rshift, d.rshift, 1); cp(cosne, d.cosne, 1);
    cp(ncgs, d.ncgs, 1); cp(ncgsnth, d.ncgsnth, 1); cp(bcgs, d.bcgs, 1); cp(tcgs, d.tcgs, 1);
    cp(e_last, d.j, 4); cp(K_last, d.K, 7); cp(tau, d.tau, 1);
    for (int z = 0; z < cap; z++) {
      if (tpm) tpm[q * cap + z] = (z < n && z < (int)d.tpm.size()) ? d.tpm[z] : 0;
      if (tpr) tpr[q * cap + z] = (z < n && z < (int)d.tpr.size()) ? d.tpr[z] : 0;
      if (rho) rho[q * cap + z] = (z < n && z < (int)d.rho.size()) ? d.rho[z] : 0.f;
      if (p) p[q * cap + z] = (z < n && z < (int)d.p.size()) ? d.p[z] : 0.f;
      if (bmag) bmag[q * cap + z] = (z < n && z < (int)d.bmag.size()) ? d.bmag[z] : 0.f;
    }
    if (stokes)
      for (int z = 0; z < nfreq * ra.nvals; z++)
        stokes[q * nfreq * ra.nvals + z] = (z < (int)d.stokes.size()) ? d.stokes[z] : 0.0;
  }
  return 0;
}

// host logic of the time axis: load the dump slots, advance nsteps times by dt; out[step*5..] = indf, toffset, tstep,
// slot-weighted sums of rho_arr and u0_arr
int orc_fluid_series(const grt_args *a, int nsteps, double dt, double *out) {
  RunArgs r = to_run(a);
  FluidData fd;
  FluidArgs fa = r.fargs;
  fa.mbh = r.mbh;
  if (!load_fluid_model(fd, r.fname, r.spin, fa)) return 1;
  for (int st = 0; st <= nsteps; st++) {
    if (st > 0) {
      advance_fluid_timestep(fd, r.fname, r.spin, fa, dt);
      if (!fd.ok) return 1;
    }
    double sr = 0, su = 0;
    for (int k = 0; k < fd.nt; k++)
      for (size_t i = 0; i < fd.ncell; i++) {
        sr += (k + 1) * (double)fd.rho_arr[(size_t)k * fd.ncell + i];
        su += (k + 1) * (double)fd.u0_arr[(size_t)k * fd.ncell + i];
      }
    double *o = out + 5 * st;
    o[0] = fd.indf; o[1] = (fd.model == F_IHARM) ? fd.toffset_d : (double)fd.toffset;
    o[2] = (fd.model == F_IHARM) ? fd.tstep_d : (double)fd.tstep; o[3] = sr; o[4] = su;
  }
  return 0;
}

}  // extern "C"
