// Debug/test harness: drives the product's __host__ __device__ stage bodies (grt_pipeline.cuh) on the
// CPU, with the product's own host layer (grt_io.cpp), to compare kernel logic with the oracle without a
// GPU.  NOT part of the product; never loaded by grtrans_b200.  Built by tests/test_host_emul.py.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../grtrans_b200/csrc/grt_host.h"
#include "../../grtrans_b200/csrc/grt_pipeline.cuh"
#include "../../grtrans_b200/csrc/grt_debug.cuh"

using namespace grt;

// emul_geodebug: file to write for the (single) ray of the next emul_full call, like grt_geodebug
static const char *g_geodebug_path = nullptr;

extern "C" int emul_full(const grt_args *a, int i1, int i2, float *image /*[ncams][n][nvals]*/, int *npts_out,
                         double *stage /* optional [n][maxpts][8]: lam,r,th,ph,s2xi,c2xi,ang,rshift */) {
  int eneq = 1;
  int etype = grt_host::emis_type_id(a->ename, &eneq), iflag = grt_host::integrator_flag(a->iname);
  if (etype < 0 || iflag < 0) return 2;
  grt_host::FluidHost fh;
  int rc = grt_host::load_fluid(fh, *a);
  if (rc) return rc;
  std::vector<double> freqs, mdots, mu0;
  grt_host::make_grids(*a, freqs, mdots, mu0);
  grt_host::EmisTablesHost th;
  grt_host::build_emis_tables(th);
  FluidKernelArgs fk;
  memset(&fk, 0, sizeof(fk));
  FluidDev &F = fk.F;
  F.model = fh.model; F.nx1 = fh.nx1; F.nx2 = fh.nx2; F.nx3 = fh.nx3; F.h = fh.h;
  F.td_mbh = fh.td_mbh; F.td_mdot = fh.td_mdot; F.td_rin = fh.td_rin; F.td_rout = fh.td_rout;
  F.hslope = fh.hslope; F.mks_smooth = fh.mks_smooth; F.poly_xt = fh.poly_xt; F.poly_alpha = fh.poly_alpha;
  F.startx1 = fh.startx1; F.is_mmks = fh.is_mmks; F.eheat = fh.eheat; F.gam = fh.gam;
  F.kh = fh.kh; F.kr0 = fh.kr0; F.kaa = fh.kaa; F.kbb = fh.kbb; F.kpp = fh.kpp; F.ktype = fh.ktype;
  F.x2mintrust = fh.x2mintrust; F.x2maxtrust = fh.x2maxtrust;
  F.ax1 = fh.ax1.data(); F.axr = fh.axr.data(); F.ax2 = fh.ax2.data(); F.axth = fh.axth.data(); F.ax3 = fh.ax3.data();
  F.cells = reinterpret_cast<const Cell *>(fh.cells.data());
  F.nslots = fh.nslots; F.ncell = fh.ncell; F.tstep = fh.tstep;
  F.nth = fh.nth.data(); F.nrelbin = fh.nrelbin;
  F.uvals = fh.uvals.data(); F.vvals = fh.vvals.data(); F.uv = fh.uv.data(); F.tvals = fh.tvals.data();
  TransferParams P;
  memset(&P, 0, sizeof(P));
  EmisTables &tb = P.tab;
  tb.gx = th.gx.data(); tb.gy = th.gy.data(); tb.xv = th.xv.data(); tb.ya = th.ya.data(); tb.yp = th.yp.data();
  tb.yv = th.yv.data(); tb.yap = th.yap.data(); tb.yav = th.yav.data();
  tb.ch_mu = th.ch_mu.data(); tb.ch_I = th.ch_I.data(); tb.ch_del = th.ch_del.data();
  {
    const char *fe = getenv("GRT_FARADAY_EXACT");
    const bool exact = fe && atoi(fe) != 0;
    tb.far = exact ? nullptr : th.far.data(); tb.far_u0 = th.far_u0; tb.far_ih = th.far_ih; tb.far_n = exact ? 0 : th.far_n;
  }
  fk.tab = tb;
  int nparams = a->nmdot, nfreq = a->nfreq, nv = a->nvals;
  const int nextra = (a->extra == 1) ? (a->nn[2] > 1 ? 19 : 7) : 0, nout = nv + nextra;  // image: [cam][ray][nvals + nextra]
  fk.nparams = nparams; fk.etype = etype;
  for (int m = 0; m < nparams; m++) {
    SourceDev &sp = fk.sp[m];
    sp.mdot = mdots[m]; sp.mbh = a->mbh; sp.nfac = a->ftscl; sp.bfac = a->frscl; sp.jetalphaval = a->jetalpha;
    sp.gminval = a->gmin; sp.muval = a->muval; sp.gmax = a->gmax; sp.p1 = a->p1; sp.p2 = a->p2; sp.sigcut = a->sigcut;
    sp.stail = (std::string(a->stype) == "tail") ? 1 : 0; sp.pad = 0;
  }
  P.stail = (std::string(a->stype) == "tail") ? 1 : 0;
  fk.ncomp = 0; P.ncomp = 0; P.isel = 0;
  if (etype == EM_MAXJUTT || etype == EM_MAXCOMP) {  // as grt_session_create
    const int first = (etype == EM_MAXCOMP) ? 2 : 1, nw = (a->nepotherargs < 8 ? a->nepotherargs : 8) - first;
    if (nw < 1 || nw > GRT_MAXJUTT_NCOMP) return 3;
    const double delta = a->epotherargs[0];
    double wsum = 0.0, dwsum = 0.0;
    for (int i = 0; i < nw; i++) wsum += a->epotherargs[first + i];
    for (int i = 0; i < nw; i++) dwsum += a->epotherargs[first + i] / wsum * std::pow(delta, (double)i);
    for (int i = 0; i < nw; i++) {
      fk.mj_w[i] = a->epotherargs[first + i] / wsum;
      fk.mj_t[i] = std::pow(delta, (double)i) / dwsum;
    }
    fk.ncomp = nw; P.ncomp = nw;
    P.isel = (etype == EM_MAXCOMP) ? (int)a->epotherargs[1] : 0;
  }
  std::vector<double> bins;
  if (etype == EM_BINS || etype == EM_HYBRIDTHBINS) {  // as grt_session_create
    if (fh.nrelbin < 1 || fh.nrelbin > GRT_MAXBINS) return 3;
    grt_host::emis_model_bins((double)fh.relgammamin, (double)fh.relgammamax, fh.nrelbin, bins);
    P.nbins = fh.nrelbin; P.bins = bins.data();
  }
  P.etype = etype; P.eneq = eneq; P.rneq = nv; P.iflag = iflag; P.model = fh.model; P.npts_is_one = (a->nn[2] == 1);
  for (int i = 0; i < 7; i++) P.coefindx[i] = a->epcoefindx[i];
  P.LBH = pc::G * a->mbh * pc::msun / pc::c2;
  P.gmin = (double)(float)a->gmin; P.gmax = (double)(float)a->gmax; P.p = (double)(float)a->p1;
  P.max_tau = 10.0; P.hmax = 0.1; P.ortol = 1e-6; P.oatol = 1e-8; P.thin = 1e-2; P.maxsteps = 100000;
  int n = i2 - i1 + 1;
  for (int jm = 0; jm < a->nmu; jm++) {
    grt_host::Camera c;
    grt_host::build_camera(c, a->standard, mu0[jm], a->phi0, a->spin, a->uout, a->rcut, a->nrotype, a->gridvals, a->nn[0],
                           a->nn[1], a->nn[2]);
    int mp = c.maxpts();
    fk.maxpts = mp;
    GeoFlags fl{a->nn[2], c.kext, c.next, c.mufill, c.usegeor, 1, (fh.nslots <= 1 && !g_geodebug_path) ? 1 : 0};  // as grt_session_create / grt_geodebug
#pragma omp parallel for schedule(dynamic, 4)
    for (int q = 0; q < n; q++) {
      int ix = i1 - 1 + q;
      RayConst r;
      r.a = c.a; r.mu0 = c.mu0; r.u0 = c.u0; r.uout = c.uout; r.alpha = c.alpha[ix]; r.beta = c.beta[ix]; r.q2 = c.q2[ix];
      r.l = c.l[ix]; r.su = c.su[ix]; r.sm = c.sm[ix]; r.uf = c.uf[ix]; r.muf = c.muf[ix]; r.offset = c.offset;
      r.phi0 = c.phi0; r.t0 = 0; r.tpm = c.tpm[ix]; r.tpr = c.tpr[ix];
      RayInv inv;
      geo_setup(r, fl, inv);
      std::vector<double> bu(mp), bmu(mp), bdt(mp), bdp(mp), bl(mp), rec((size_t)nparams * mp * record_doubles(etype), 0.0);
      std::vector<double> xrec(nextra ? (size_t)nparams * mp * 8 : 0, 0.0), hist(nextra ? mp : 0, 0.0);
      FluidKernelArgs fkr = fk;  // (per-thread copy: the extra-quantity record of this ray)
      fkr.xrec = nextra ? xrec.data() : nullptr;
      std::vector<signed char> btm(mp), btr(mp);
      PointArrays pa{bu.data(), bmu.data(), bdt.data(), bdp.data(), bl.data(), btm.data(), btr.data()};
      for (int s = 0; s < mp; s++) point_store(inv, fl.nup, s, (size_t)s, pa);
      RayMeta m;
      ray_finalize(inv.npts, 0, mp, fl.usegeor, pa, m);
      StageDump sd;
      memset(&sd, 0, sizeof(sd));
      std::vector<double> sdl(mp), sdt(mp), sdr(mp), sdth(mp), sdph(mp), sk0(mp), sk1(mp), sk2(mp), sk3(mp), su(4 * mp),
          sb(4 * mp), s2(mp), c2(mp), an(mp), rs(mp), cn(mp), n1(mp), n2(mp), n3(mp), n4(mp);
      std::vector<float> fr(mp), fp(mp), fb(mp);
      const bool dbg = g_geodebug_path != nullptr && jm == a->nmu - 1 && q == n - 1;
      if ((stage && jm == 0) || dbg) {
        sd.lambda = sdl.data(); sd.t = sdt.data(); sd.r = sdr.data(); sd.th = sdth.data(); sd.ph = sdph.data();
        sd.kt = sk0.data(); sd.kr = sk1.data(); sd.kth = sk2.data(); sd.kph = sk3.data(); sd.u = su.data(); sd.b = sb.data();
        sd.s2xi = s2.data(); sd.c2xi = c2.data(); sd.ang = an.data(); sd.rshift = rs.data(); sd.cosne = cn.data();
        sd.ncgs = n1.data(); sd.ncgsnth = n2.data(); sd.bcgs = n3.data(); sd.tcgs = n4.data();
        sd.rho = fr.data(); sd.p = fp.data(); sd.bmag = fb.data();
      }
      for (int s = 0; s < mp; s++) point_record(r, m, fkr, 0, s, 1, pa, rec.data(), sd);
      if (jm == 0) npts_out[q] = (m.status == 1) ? m.npts : 0;
      if (dbg) {  // as grt_geodebug + k_geodebug
        grt_host::GeoDebugData d;
        d.npts = (m.status == 1) ? m.npts : 0;
        d.maxpts = mp;
        d.alpha = r.alpha; d.beta = r.beta; d.su = r.su; d.sm = r.sm; d.a = r.a; d.mu0 = r.mu0;
        d.mdots = mdots; d.mbh.assign(mdots.size(), a->mbh);
        d.rows.assign((size_t)GD_NROWS * mp, 0.0);
        std::vector<double> hist4((size_t)4 * mp, 0.0);
        double scal[2] = {0.0, 1.0};
        for (int p = 0; p < d.npts; p++) geodebug_point(r, m, mp, pa, sd, p, d.rows.data());
        if (d.npts > 0) geodebug_ray(P, rec.data(), d.npts, mp, freqs[0], hist4.data(), d.rows.data(), scal);
        d.fac = scal[0];
        d.lambda = sdl; d.t = sdt; d.r = sdr; d.th = sdth; d.ph = sdph; d.kt = sk0; d.kr = sk1; d.kth = sk2; d.kph = sk3;
        d.u = su; d.b = sb; d.ncgs = n1; d.tcgs = n4; d.bcgs = n3; d.rshift = rs; d.ang = an; d.s2xi = s2; d.c2xi = c2;
        d.rho = fr; d.p = fp; d.bmag = fb;
        grt_host::write_geodebug(g_geodebug_path, d);
      }
      if (stage && jm == 0)
        for (int s = 0; s < mp; s++) {
          double *o = stage + ((size_t)q * mp + s) * 8;
          bool v = m.status == 1 && s < m.npts;
          o[0] = v ? sdl[s] : 0; o[1] = v ? sdr[s] : 0; o[2] = v ? sdth[s] : 0; o[3] = v ? sdph[s] : 0; o[4] = v ? s2[s] : 0;
          o[5] = v ? c2[s] : 0; o[6] = v ? an[s] : 0; o[7] = v ? rs[s] : 0;
        }
      for (int mm = 0; mm < nparams; mm++)
        for (int f = 0; f < nfreq; f++) {
          double out[4] = {0, 0, 0, 0}, xt[19];
          for (int k = 0; k < 19; k++) xt[k] = 0.0;
          const double *rr = rec.data() + (size_t)mm * mp * record_doubles(etype);
          if (m.status == 1 && m.npts > 0) {
            if (nextra)
              ray_extra(P, rr, xrec.data() + (size_t)mm * mp * 8, m.npts, freqs[f], out, xt, hist.data(), 1);
            else
              ray_transfer(P, rr, m.npts, freqs[f], out);
          }
          int cam = jm * nparams * nfreq + mm * nfreq + f;
          for (int k = 0; k < nv; k++) image[((size_t)cam * n + q) * nout + k] = (float)out[k];
          for (int k = 0; k < nextra; k++) image[((size_t)cam * n + q) * nout + nv + k] = (float)xt[k];
        }
    }
  }
  return 0;
}

extern "C" int emul_geodebug(const grt_args *a, int pixel, const char *path) {
  grt_args aa = *a;
  aa.extra = 0;
  const int ncams = aa.nmdot * aa.nfreq * aa.nmu;
  std::vector<float> img((size_t)ncams * aa.nvals, 0.f);
  int npts = 0;
  g_geodebug_path = path;
  int rc = emul_full(&aa, pixel, pixel, img.data(), &npts, nullptr);
  g_geodebug_path = nullptr;
  return rc;
}

// Debug: the transfer coefficients node_coef() hands to the integrator, for one ray at one frequency: out[p*11 ..] =
// j[4] (invariant, not divided by fac), K[7] (times LBH / g).  Returns the number of points (mu0 index 0, mdot 0).
extern "C" int emul_ray_coefs(const grt_args *a, int ix1, double nu_obs, double *out, int cap) {
  grt_host::FluidHost fh;
  if (grt_host::load_fluid(fh, *a)) return -1;
  std::vector<double> freqs, mdots, mu0;
  grt_host::make_grids(*a, freqs, mdots, mu0);
  int eneq = 1;
  const int etype = grt_host::emis_type_id(a->ename, &eneq);
  FluidKernelArgs fk;
  memset(&fk, 0, sizeof(fk));
  FluidDev &F = fk.F;
  F.model = fh.model; F.nx1 = fh.nx1; F.nx2 = fh.nx2; F.nx3 = fh.nx3; F.h = fh.h; F.nslots = fh.nslots; F.ncell = fh.ncell;
  F.tstep = fh.tstep; F.td_mbh = fh.td_mbh; F.td_mdot = fh.td_mdot; F.td_rin = fh.td_rin; F.td_rout = fh.td_rout;
  F.ax1 = fh.ax1.data(); F.axr = fh.axr.data(); F.ax2 = fh.ax2.data(); F.axth = fh.axth.data(); F.ax3 = fh.ax3.data();
  F.uvals = fh.uvals.data(); F.vvals = fh.vvals.data(); F.uv = fh.uv.data(); F.tvals = fh.tvals.data();
  F.cells = reinterpret_cast<const Cell *>(fh.cells.data());
  grt_host::EmisTablesHost th;
  grt_host::build_emis_tables(th);
  TransferParams P;
  memset(&P, 0, sizeof(P));
  EmisTables &tb = P.tab;
  tb.gx = th.gx.data(); tb.gy = th.gy.data(); tb.xv = th.xv.data(); tb.ya = th.ya.data(); tb.yp = th.yp.data();
  tb.yv = th.yv.data(); tb.yap = th.yap.data(); tb.yav = th.yav.data();
  tb.far = nullptr; tb.far_n = 0;
  fk.tab = tb; fk.nparams = 1; fk.etype = etype;
  SourceDev &sp = fk.sp[0];
  sp.mdot = mdots[0]; sp.mbh = a->mbh; sp.nfac = a->ftscl; sp.bfac = a->frscl; sp.jetalphaval = a->jetalpha;
  sp.gminval = a->gmin; sp.muval = a->muval; sp.gmax = a->gmax; sp.p1 = a->p1; sp.p2 = a->p2; sp.sigcut = a->sigcut;
  P.etype = etype; P.eneq = eneq; P.rneq = a->nvals; P.npts_is_one = (a->nn[2] == 1);
  for (int i = 0; i < 7; i++) P.coefindx[i] = a->epcoefindx[i];
  P.LBH = pc::G * a->mbh * pc::msun / pc::c2;
  P.gmin = (double)(float)a->gmin; P.gmax = (double)(float)a->gmax; P.p = (double)(float)a->p1;
  grt_host::Camera c;
  grt_host::build_camera(c, a->standard, mu0[0], a->phi0, a->spin, a->uout, a->rcut, a->nrotype, a->gridvals, a->nn[0], a->nn[1],
                         a->nn[2]);
  const int mp = c.maxpts(), ix = ix1 - 1;
  fk.maxpts = mp;
  GeoFlags fl{a->nn[2], c.kext, c.next, c.mufill, c.usegeor, 1, 1};
  RayConst r;
  r.a = c.a; r.mu0 = c.mu0; r.u0 = c.u0; r.uout = c.uout; r.alpha = c.alpha[ix]; r.beta = c.beta[ix]; r.q2 = c.q2[ix];
  r.l = c.l[ix]; r.su = c.su[ix]; r.sm = c.sm[ix]; r.uf = c.uf[ix]; r.muf = c.muf[ix]; r.offset = c.offset; r.phi0 = c.phi0;
  r.t0 = 0; r.tpm = c.tpm[ix]; r.tpr = c.tpr[ix];
  RayInv inv;
  geo_setup(r, fl, inv);
  std::vector<double> bu(mp), bmu(mp), bdt(mp), bdp(mp), bl(mp), rec((size_t)mp * record_doubles(etype), 0.0);
  std::vector<signed char> btm(mp), btr(mp);
  PointArrays pa{bu.data(), bmu.data(), bdt.data(), bdp.data(), bl.data(), btm.data(), btr.data()};
  for (int s = 0; s < mp; s++) point_store(inv, fl.nup, s, (size_t)s, pa);
  RayMeta m;
  ray_finalize(inv.npts, 0, mp, fl.usegeor, pa, m);
  StageDump sd;
  memset(&sd, 0, sizeof(sd));
  for (int s = 0; s < mp; s++) point_record(r, m, fk, 0, s, 1, pa, rec.data(), sd);
  if (m.status != 1) return 0;
  const FreqConst Fq = freq_const(nu_obs);
  for (int p = 0; p < m.npts && p < cap; p++) {
    Coef cf;
    double jraw;
    const double *rk = rec.data() + (size_t)p * record_doubles(etype);
    if (etype == EM_POLSYNCHTH) {
      RunConst R = run_const<EM_POLSYNCHTH>(P, nu_obs);
      node_coef<EM_POLSYNCHTH>(P, R, rk, Fq, cf, jraw);
    } else {
      return -2;
    }
    for (int i = 0; i < 4; i++) out[p * 11 + i] = cf.j[i];
    for (int i = 0; i < 7; i++) out[p * 11 + 4 + i] = cf.K[i];
  }
  return m.npts;
}

// The kernel's DELO march (delo_step: closed-form 4x4 solve, merged thick / thin weights) on caller-supplied coefficients:
// s[npts] (lambda, descending), j[npts][4], K[npts][7]; intensity[npts][4] like grt_rt_integrate (column k = intensity(:,k)).
extern "C" void emul_delo(int npts, const double *s, const double *j, const double *K, double thin, double *intensity) {
  double Ip[4] = {0, 0, 0, 0}, In[4];
  for (int i = 0; i < 4 * npts; i++) intensity[i] = 0.0;
  auto load = [&](int k, Coef &c) {
    for (int i = 0; i < 4; i++) c.j[i] = j[(size_t)k * 4 + i];
    for (int i = 0; i < 7; i++) c.K[i] = K[(size_t)k * 7 + i];
  };
  Coef c0, c1;
  load(npts - 1, c1);
  for (int k = npts - 2; k >= 0; k--) {
    load(k, c0);
    const double dx = s[k] - s[k + 1];
    const double delta = dx * 0.5 * (fabs(c0.K[0]) + fabs(c1.K[0]));
    delo_step(c0, c1, dx, delta, thin, Ip, In);
    const int col = npts - 1 - k;
    for (int i = 0; i < 4; i++) {
      Ip[i] = In[i];
      intensity[(size_t)col * 4 + i] = In[i];
    }
    c1 = c0;
  }
}

// Host logic of the time axis: load the dump slots, then advance `nsteps` times by dt.  out[(step)*5 + ..] = indf, toffset,
// tstep, sum_k (k+1) sum_cells rho, sum_k (k+1) sum_cells u0 (slot-weighted, so a slot shift is visible).
extern "C" int emul_fluid_series(const grt_args *a, int nsteps, double dt, double *out) {
  grt_host::FluidHost fh;
  int rc = grt_host::load_fluid(fh, *a);
  if (rc) return rc;
  for (int st = 0; st <= nsteps; st++) {
    if (st > 0) {
      rc = grt_host::advance_fluid(fh, *a, dt);
      if (rc) return rc;
    }
    double sr = 0, su = 0;
    for (int k = 0; k < fh.nslots; k++)
      for (size_t i = 0; i < fh.ncell; i++) {
        const grt_host::HostCell &c = fh.cells[(size_t)k * fh.ncell + i];
        sr += (k + 1) * (double)c.rho;
        su += (k + 1) * (double)c.u0;
      }
    double *o = out + 5 * st;
    o[0] = fh.indf; o[1] = (fh.model == 29) ? fh.toffset_d : (double)fh.toffset;
    o[2] = (fh.model == 29) ? fh.tstep_d : (double)fh.tstep; o[3] = sr; o[4] = su;
  }
  return 0;
}

// scalar math replacements used on the per-(point, frequency) path
extern "C" void emul_math(int which, int n, const double *x, double *o) {
  for (int i = 0; i < n; i++) o[i] = (which == 0) ? exp_fast(x[i]) : log_ge1(x[i]);
}

// Faraday-fit table against the direct evaluation of the fits: out[i*4..] = jff_table, log_table, jff_direct, log_direct
extern "C" void emul_faraday(int n, const double *xarg, double *out) {
  grt_host::EmisTablesHost th;
  grt_host::build_emis_tables(th);
  for (int i = 0; i < n; i++) {
    const double x = xarg[i], u = (std::log(x) - th.far_u0) * th.far_ih;
    double *o = out + 4 * (size_t)i;
    o[0] = o[1] = NAN;
    if (u >= 0.0 && u < (double)th.far_n) {
      const int k = (int)u;
      const double w = u - k, *q = &th.far[(size_t)k * 8];
      o[0] = q[0] + w * (q[1] + w * (q[2] + w * q[3]));
      o[1] = q[4] + w * (q[5] + w * (q[6] + w * q[7]));
    }
    faraday_fits(std::log(x), true, o[2], o[3]);
  }
}
