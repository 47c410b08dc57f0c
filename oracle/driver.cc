// ORACLE -- TEST INFRASTRUCTURE ONLY (see geokerr.cc header).
// CPU restatement of the per-pixel driver and its orchestration:
//   grtrans_driver.f90:57-466 (grtrans_driver), :468-513 (convert_model, grtrans_compute_intensity),
//   emis.f90:461-571 (calc_emissivity copy-out), :786-847 (calc_opt_depth, rotate_emis,
//   invariant_emis/intensity), pgrtrans.f90:12-253 (grtrans_main: frequency/mdot/mu grids, camera
//   indexing, the OpenMP `schedule(static,1)` pixel loop and its wall-clock bracket).
#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "oracle.h"

namespace orc {

// pgrtrans.f90:88-106
void compute_freqs(const RunArgs &a, std::vector<double> &freqs, std::vector<double> &mdots,
                   std::vector<double> &mu0) {
  freqs.resize(a.nfreq);
  mdots.resize(a.nmdot);
  mu0.resize(a.nmu);
  if (a.nfreq == 1)
    freqs[0] = a.fmin;
  else
    for (int i = 0; i < a.nfreq; i++) freqs[i] = a.fmin * std::exp(i * std::log(a.fmax / a.fmin) / (a.nfreq - 1));
  if (a.nmdot == 1)
    mdots[0] = a.mdotmin;
  else
    for (int i = 0; i < a.nmdot; i++)
      mdots[i] = a.mdotmin * std::exp(i * std::log(a.mdotmax / a.mdotmin) / (a.nmdot - 1));
  if (a.nmu == 1)
    mu0[0] = a.mumin;
  else
    for (int i = 0; i < a.nmu; i++) mu0[i] = a.mumin + (a.mumax - a.mumin) / (a.nmu - 1) * i;
}

static int iflag_from_name(const std::string &iname) {  // rad_trans.f90:30-38
  if (iname == "lsoda") return 0;
  if (iname == "delo") return 1;
  if (iname == "formal") return 2;
  if (iname == "lsodasph") return 3;
  return 0;
}

namespace {
struct RayWork {
  Geo g;
  std::vector<FluidPt> f;
  std::vector<double> s2xi, c2xi, ang, rshift, cosne, ncgs, ncgsnth, bcgs, tcgs, gmin;
  std::vector<double> ej, eK, j, K, tau, intensity, absK;
};
}  // namespace

// pgrtrans.f90:107-113: 19 extra quantities per pixel for rays with more than one point, 7 for nup = 1
static int extra_count(const RunArgs &A) { return (A.extra == 1) ? (A.nn[2] > 1 ? 19 : 7) : 0; }

// grtrans_driver.f90:230-292 (EXTRA_QUANTS == 1, npts > 1).  w.ej / w.eK hold the rotated, invariant, scaled coefficients,
// w.tau the optical depth from the camera, w.intensity the integrator's output columns 1..rnpts.  rt[0..18] = r%tau(1:19).
static void extra_quantities(const RayWork &w, int rneq, int rnpts, double iscale, double *rt) {
  const Geo &g = w.g;
  const int npts = g.npts;
  for (int q = 0; q < 19; q++) rt[q] = 0.0;
  std::vector<double> s(npts), y(npts), acc(npts), dummy(npts), norm(npts);
  std::vector<std::vector<double>> tau_arr(6, std::vector<double>(npts));
  const int inds[6] = {0, 1, 2, 3, 4, 6};  // K columns 1,2,3,4,5,7: alpha_I,Q,U,V, rho_Q, rho_V
  for (int p = 0; p < npts; p++) s[p] = g.lambda[0] - g.lambda[p];
  for (int ii = 0; ii < 6; ii++) {
    for (int p = 0; p < npts; p++) y[p] = std::fabs(w.eK[(size_t)p * 7 + inds[ii]]);
    tsum(npts, s.data(), y.data(), tau_arr[ii].data());
  }
  // photosphere index: minloc(|tau-1|); (sic) "if(minval(tau).lt.1d0) taudex=g%npts" -- tau(1) = 0, so this always holds
  int taudex = 1;
  double best = std::fabs(w.tau[0] - 1.0), mintau = w.tau[0];
  for (int p = 1; p < npts; p++) {
    double d = std::fabs(w.tau[p] - 1.0);
    if (d < best) { best = d; taudex = p + 1; }
    if (w.tau[p] < mintau) mintau = w.tau[p];
  }
  if (mintau < 1.0) taudex = npts;
  if (taudex == 1) taudex = 2;
  for (int ii = 0; ii < 6; ii++) rt[ii] = tau_arr[ii][taudex - 1];
  for (int p = 0; p < npts; p++) dummy[p] = w.ej[(size_t)p * 4] * std::exp(-w.tau[p]);
  tsum(npts, g.lambda.data(), dummy.data(), norm.data());
  auto weighted = [&](auto val) {
    for (int p = 0; p < npts; p++) y[p] = dummy[p] * val(p);
    tsum(npts, g.lambda.data(), y.data(), acc.data());
    return 1.0 / norm[taudex - 1] * acc[taudex - 1];
  };
  rt[6] = weighted([&](int p) { return g.x[p].d[1]; });
  rt[7] = weighted([&](int p) { return g.x[p].d[2]; });
  rt[8] = weighted([&](int p) { return g.x[p].d[3]; });
  rt[9] = weighted([&](int p) { return w.ncgs[p]; });
  rt[10] = weighted([&](int p) { return w.tcgs[p]; });
  rt[11] = weighted([&](int p) { return w.bcgs[p]; });
  rt[12] = weighted([&](int p) { return (double)(w.f[p].p * 2.f / (w.f[p].bmag * w.f[p].bmag)); });  // plasma beta, real(4)
  rt[13] = weighted([&](int p) { return std::cos(g.x[p].d[2]) >= 0.0 ? 1.0 : -1.0; });
  // (sic) "LP" increments use Stokes I and Q: sqrt(I(1,:)**2 + I(2,:)**2), reversed against the point arrays
  {
    std::vector<double> dlp(std::max(rnpts - 1, 0), 0.0);  // (rnpts < 2: empty sums, 0/0)
    if (rneq == 4) {
      auto P = [&](int col) {  // 1-based column
        double a = w.intensity[(size_t)(col - 1) * 4] * iscale, b = w.intensity[(size_t)(col - 1) * 4 + 1] * iscale;
        return std::sqrt(a * a + b * b);
      };
      for (int kk = 1; kk <= rnpts - 1; kk++) dlp[kk - 1] = P(rnpts + 1 - kk) - P(rnpts - kk);
    }
    double sd = 0, sr = 0, sth = 0, s5 = 0, s6 = 0, ssg = 0;
    for (int kk = 0; kk < rnpts - 1; kk++) {
      sd += dlp[kk];
      sr += dlp[kk] * g.x[kk].d[1];
      sth += dlp[kk] * g.x[kk].d[2];
      s5 += dlp[kk] * tau_arr[4][kk];
      s6 += dlp[kk] * tau_arr[5][kk];
      ssg += dlp[kk] * (std::cos(g.x[kk].d[2]) >= 0.0 ? 1.0 : -1.0);
    }
    rt[14] = sr / sd; rt[15] = sth / sd; rt[16] = s5 / sd; rt[17] = s6 / sd; rt[18] = ssg / sd;
  }
}

// Process one geodesic index for all (mdot, freq).  out[(m*nfreq+k)*nvals + s] in cgs (double).
static void drive_ray(const RunArgs &A, const FluidData &fd, const GeokerrArgs &gargs, int i,
                      const std::vector<double> &freqs, const std::vector<double> &mdots, double *out,
                      long long *npoints, RayDump *dump) {
  const int nfreq = (int)freqs.size(), nparams = (int)mdots.size(), nvals = A.nvals;
  const int nextra = extra_count(A), nout = nvals + nextra;  // pixel record: Stokes values, then the extra quantities
  for (int q = 0; q < nparams * nfreq * nout; q++) out[q] = 0.0;
  RayWork w;
  Geo &g = w.g;
  int status;
  initialize_geodesic(g, gargs, i, status);
  if (dump) {
    dump->npts = g.npts;
    dump->status = status;
    dump->ncase = g.ncase;
  }
  if (!(status == 1 && g.npts > 0)) return;
  int npts = g.npts;
  if (npoints) *npoints += npts;
  if (legacy.lambda_perturb_ulp > 0.0) {
    // test-only (oracle.h): the affine parameter is the difference of two numbers of size 1/u0 (geokerr_wrapper.f LAMBDA
    // integrals; u0 <= 1e-4), so the reference's own values carry rounding noise of a few ulp(1/u0).  This adds noise of
    // that size, to measure how far the reference ALGORITHM amplifies it on a given ray.
    const double ulp = 2.220446049250313e-16 / gargs.u0 * legacy.lambda_perturb_ulp;
    unsigned long long h = 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
    for (int p = 0; p < npts; p++) {
      h ^= h << 13; h ^= h >> 7; h ^= h << 17;
      g.lambda[p] += ulp * ((double)(h >> 11) / 9007199254740992.0 * 2.0 - 1.0);
    }
  }
  int eneq;
  int etype = emis_type_from_name(A.ename, eneq);
  int enk = 1 + eneq * (eneq - 1) / 2;
  get_fluid_vars(fd, g, w.f);
  w.s2xi.assign(npts, 0); w.c2xi.assign(npts, 0); w.ang.assign(npts, 0); w.rshift.assign(npts, 0); w.cosne.assign(npts, 0);
  for (int p = 0; p < npts; p++)
    comoving_ortho(g.x[p].d[1], g.x[p].d[2], g.a, g.alpha, g.beta, g.mu0, w.f[p].u, w.f[p].b, g.k[p].d, w.s2xi[p],
                   w.c2xi[p], w.ang[p], w.rshift[p], w.cosne[p]);
  int rneq = nvals;
  int rnk = 1 + rneq * (rneq - 1) / 2;
  int iflag = iflag_from_name(A.iname);
  int rnpts = npts;
  RtOptions opt;
  if (getenv("ORC_LSODA_INTENDED")) opt.lsoda_neighbor_as_written = false;  // experiment switch
  if (getenv("ORC_RT_TIGHT")) { opt.ortol = 1e-9; opt.oatol = 1e-11; }
  w.ej.assign((size_t)npts * 4, 0.0);
  w.eK.assign((size_t)npts * 7, 0.0);
  w.j.assign((size_t)npts * rneq, 0.0);
  w.K.assign((size_t)npts * rnk, 0.0);
  w.tau.assign(npts, 0.0);
  w.absK.assign(npts, 0.0);
  w.intensity.assign((size_t)npts * rneq, 0.0);
  if (dump) {
    dump->lambda = g.lambda;
    dump->tpm = g.tpmarr;
    dump->tpr = g.tprarr;
    for (int p = 0; p < npts; p++) {
      dump->t.push_back(g.x[p].d[0]); dump->r.push_back(g.x[p].d[1]); dump->th.push_back(g.x[p].d[2]);
      dump->ph.push_back(g.x[p].d[3]);
      dump->kt.push_back(g.k[p].d[0]); dump->kr.push_back(g.k[p].d[1]); dump->kth.push_back(g.k[p].d[2]);
      dump->kph.push_back(g.k[p].d[3]);
      dump->rho.push_back(w.f[p].rho); dump->p.push_back(w.f[p].p); dump->bmag.push_back(w.f[p].bmag);
      for (int c = 0; c < 4; c++) {
        dump->u.push_back(w.f[p].u[c]);
        dump->b.push_back(w.f[p].b[c]);
      }
    }
    dump->s2xi = w.s2xi; dump->c2xi = w.c2xi; dump->ang = w.ang; dump->rshift = w.rshift; dump->cosne = w.cosne;
    dump->stokes.assign((size_t)nfreq * nvals, 0.0);
  }
  for (int m = 0; m < nparams; m++) {
    SourceParams sp;
    sp.mdot = mdots[m]; sp.mbh = A.mbh; sp.nfac = A.ftscl; sp.bfac = A.frscl; sp.jetalphaval = A.jetalpha;
    sp.gminval = A.gmin; sp.muval = A.muval; sp.gmax = A.gmax; sp.p1 = A.p1; sp.p2 = A.p2; sp.sigcut = A.sigcut;
    double MBH = sp.mbh;
    double LBH = pc::G * MBH * pc::msun / pc::c2;
    sp.type = (A.stype == "tail") ? 1 : 0;  // assign_source_params_type (fluid.f90:1614-1624)
    convert_fluid_vars(fd, w.f, sp, w.ncgs, w.ncgsnth, w.bcgs, w.tcgs);
    assign_source_params(sp, w.ncgs, w.tcgs, w.ncgsnth, w.gmin);  // fluid.f90:582
    // emis_params are single precision in the reference (emis.f90:32-40): gmin, gmax, p1 pass through real(4)
    double e_gmax = (double)(float)A.gmax, e_p = (double)(float)A.p1;
    if (dump && m == 0) {
      dump->ncgs = w.ncgs; dump->ncgsnth = w.ncgsnth; dump->bcgs = w.bcgs; dump->tcgs = w.tcgs;
    }
    std::fill(w.eK.begin(), w.eK.end(), 0.0);
    // BINS / HYBRIDTHBINS: bin centres and widths (emis_model_bins, emis.f90:884-908; the limits travel through the real(4)
    // module variables of fluid.f90:45-46).  Fluid models without bins have nrelbin = 0: the sums are empty.
    std::vector<double> bin_g, bin_dg, bin_n;
    if (etype == E_BINS || etype == E_HYBRIDTHBINS) {
      emis_model_bins((double)(float)fd.relgammamin, (double)(float)fd.relgammamax, fd.nrelbin, bin_g, bin_dg);
      bin_n.assign(std::max(1, fd.nrelbin), 0.0);
    }
    for (int k = 0; k < nfreq; k++) {
      // ---- calc_emissivity (emis.f90:461-571)
      for (int p = 0; p < npts; p++) {
        double nu = freqs[k] / w.rshift[p];
        double K11[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        const double e_gmin = (double)(float)w.gmin[p];  // ep%gmin=sp%gmin (grtrans_driver.f90:149), real(4)
        switch (etype) {
          case E_POLSYNCHTH: polsynchth(nu, w.ncgs[p], w.bcgs[p], w.tcgs[p], w.ang[p], K11); break;
          case E_POLSYNCHSYMTH: sympolemisth(nu, w.ncgs[p], w.bcgs[p], w.tcgs[p], w.ang[p], K11); break;
          case E_MAXJUTT:
          case E_MAXCOMP:
            maxjutt(nu, w.ncgs[p], w.bcgs[p], w.tcgs[p], w.ang[p], A.epotherargs, A.nepotherargs, etype == E_MAXCOMP, K11);
            break;
          case E_POLSYNCHPL: polsynchpl(nu, w.ncgsnth[p], w.bcgs[p], w.ang[p], e_p, e_gmin, e_gmax, K11); break;
          case E_HYBRIDTHPL:
          case E_HYBRIDTH:
          case E_HYBRIDPL: {  // emis.f90:469-485: thermal + power-law populations
            double Kth[11], Kpl[11];
            polsynchth(nu, w.ncgs[p], w.bcgs[p], w.tcgs[p], w.ang[p], Kth);
            polsynchpl(nu, w.ncgsnth[p], w.bcgs[p], w.ang[p], e_p, e_gmin, e_gmax, Kpl);
            for (int c = 0; c < 11; c++) {
              if (etype == E_HYBRIDTHPL) K11[c] = Kth[c] + Kpl[c];
              else if (c < 4) K11[c] = (etype == E_HYBRIDTH) ? Kth[c] : Kpl[c];   // emission of one population ...
              else K11[c] = Kth[c] + Kpl[c];                                       // ... absorbed by both
            }
          } break;
          case E_SYNCHPL: synchpl(nu, w.ncgsnth[p], w.bcgs[p], w.ang[p], e_p, e_gmin, e_gmax, K11); break;
          case E_SYNCHTHAV: synchemis(nu, w.ncgs[p], w.bcgs[p], w.tcgs[p], K11); break;
          case E_SYNCHTHAVNOABS:  // polsynchemis.f90:906-913: synchemis with the absorption columns zeroed
            synchemis(nu, w.ncgs[p], w.bcgs[p], w.tcgs[p], K11);
            for (int c = 4; c < 11; c++) K11[c] = 0.0;
            break;
          case E_BREMS: brememis(nu, w.ncgs[p], w.tcgs[p], K11); break;
          case E_RHO: K11[0] = w.ncgs[p] * w.rshift[p]; break;  // rhoemis (emis.f90:144-151): density map, no absorption
          case E_SYNCHTHBREMS: {  // emis.f90:531-534
            double Kbr[11];
            synchemis(nu, w.ncgs[p], w.bcgs[p], w.tcgs[p], K11);
            brememis(nu, w.ncgs[p], w.tcgs[p], Kbr);
            for (int c = 0; c < 11; c++) K11[c] += Kbr[c];
          } break;
          case E_BINS:
          case E_HYBRIDTHBINS: {  // emis.f90:520-526: binned non-thermal electrons (+ angle-averaged thermal synchrotron)
            for (int b = 0; b < fd.nrelbin; b++) bin_n[b] = (b < (int)w.f[p].nnth.size()) ? (double)w.f[p].nnth[b] : 0.0;
            synchbinemis(nu, bin_n.data(), fd.nrelbin, w.bcgs[p], w.ang[p], bin_g.data(), bin_dg.data(), K11);
            if (etype == E_HYBRIDTHBINS) {
              double Kth[11];
              synchemis(nu, w.ncgs[p], w.bcgs[p], w.tcgs[p], Kth);
              for (int c = 0; c < 11; c++) K11[c] += Kth[c];
            }
          } break;
          case E_BB: K11[0] = bnu(w.tcgs[p], nu); break;
          case E_BBPOL: {
            double fcol = 1.8;
            K11[0] = std::pow(fcol, -4.0) * bnu(w.tcgs[p] * fcol, nu);
            float cI, cdel;
            interp_chandra_tab24((float)w.cosne[p], cI, cdel);
            K11[0] = K11[0] * (double)cI;
            K11[1] = K11[0] * (double)cdel;
          } break;
          default: break;
        }
        double *ej = &w.ej[(size_t)p * 4], *eK = &w.eK[(size_t)p * 7];
        if (eneq == 4) {
          for (int c = 0; c < 4; c++) ej[c] = K11[c];
          for (int c = 0; c < 7; c++)
            if (A.coefindx[c] == 1) eK[c] = K11[4 + c];
        } else {
          ej[0] = K11[0];
          eK[0] = K11[4];
        }
      }
      double fac = 0.0;
      if (npts > 1) {
        for (int p = 0; p < npts; p++) fac += w.ej[(size_t)p * 4] * (g.lambda[0] - g.lambda[p]);
      } else {
        for (int p = 0; p < npts; p++) fac += w.ej[(size_t)p * 4] / npts;
      }
      double *o = out + ((size_t)m * nfreq + k) * nout;
      if (fac / 1e16 > 0.0) {
        if (eneq == 4) {  // rotate_emis emis.f90:797-829
          for (int p = 0; p < npts; p++) {
            double *ej = &w.ej[(size_t)p * 4], *eK = &w.eK[(size_t)p * 7];
            double jq = ej[1], ju = ej[2], aq = eK[1], au = eK[2], rhoq = eK[4], rhou = eK[5];
            double c2 = w.c2xi[p], s2 = w.s2xi[p];
            ej[1] = c2 * jq - s2 * ju;
            ej[2] = s2 * jq + c2 * ju;
            eK[1] = c2 * aq - s2 * au;
            eK[2] = s2 * aq + c2 * au;
            eK[4] = c2 * rhoq - s2 * rhou;
            eK[5] = s2 * rhoq + c2 * rhou;
          }
        }
        if (npts != 1) {
          for (int p = 0; p < npts; p++) {
            double gg = w.rshift[p];
            double *ej = &w.ej[(size_t)p * 4], *eK = &w.eK[(size_t)p * 7];
            for (int c = 0; c < eneq; c++) ej[c] = ej[c] * gg * gg;  // invariant_emis
            for (int c = 0; c < enk; c++) eK[c] = eK[c] / gg;
            for (int c = 0; c < eneq; c++) ej[c] = ej[c] / fac;
            for (int c = 0; c < enk; c++) eK[c] = eK[c] * LBH;
            w.absK[p] = std::fabs(eK[0]);
          }
          tsum(npts, g.lambda.data(), w.absK.data(), w.tau.data());
          for (int p = 0; p < npts; p++) w.tau[p] = -w.tau[p];
          // pack to the integrator's width (A21: nvals and the emissivity's neq are independent)
          for (int p = 0; p < npts; p++) {
            for (int c = 0; c < rneq; c++) w.j[(size_t)p * rneq + c] = (c < eneq) ? w.ej[(size_t)p * 4 + c] : 0.0;
            for (int c = 0; c < rnk; c++) w.K[(size_t)p * rnk + c] = (c < enk) ? w.eK[(size_t)p * 7 + c] : 0.0;
          }
          if (legacy.rt_perturb > 0.0) {
            // test-only (oracle.h): relative noise of amplitude rt_perturb on the coefficients, to measure how much the
            // reference ALGORITHM amplifies rounding-level differences on this ray (conditioning of the transfer problem)
            unsigned long long h = 0x9E3779B97F4A7C15ull * (unsigned long long)(k + 1) + (unsigned long long)npts;
            auto nz = [&]() {
              h ^= h << 13; h ^= h >> 7; h ^= h << 17;
              return 1.0 + legacy.rt_perturb * ((double)(h >> 11) / 9007199254740992.0 * 2.0 - 1.0);
            };
            for (size_t q = 0; q < (size_t)npts * rneq; q++) w.j[q] *= nz();
            for (size_t q = 0; q < (size_t)npts * rnk; q++) w.K[q] *= nz();
          }
          rnpts = rt_integrate(iflag, rneq, npts, g.lambda.data(), w.j.data(), w.K.data(), w.tau.data(),
                               w.intensity.data(), opt);
          if (rnpts >= 1)
            for (int c = 0; c < rneq; c++) o[c] = w.intensity[(size_t)(rnpts - 1) * rneq + c] * fac * LBH;
          if (nextra == 19) extra_quantities(w, rneq, rnpts, fac * LBH, o + nvals);
        } else {
          // invariant_intensity(e,rshift,3); grtrans_compute_intensity
          double g3 = std::pow(w.rshift[0], 3);
          for (int c = 0; c < std::min(rneq, eneq); c++) o[c] = w.ej[c] * g3;
          if (nextra == 7) {  // grtrans_driver.f90:298-312: geometry of the single (disk) point
            double s2psi, c2psi, cosne2;
            calc_polar_psi(g.x[0].d[1], std::cos(g.x[0].d[2]), g.q2, g.a, g.alpha, g.beta, w.rshift[0], g.mu0, g.k[0].d, c2psi,
                           s2psi, cosne2);
            double *rt = o + nvals;
            rt[0] = w.s2xi[0]; rt[1] = s2psi; rt[2] = cosne2; rt[3] = w.ang[0]; rt[4] = w.cosne[0]; rt[5] = w.c2xi[0];
            rt[6] = c2psi;
          }
        }
      }
      if (dump && m == 0) {
        for (int c = 0; c < nvals; c++) dump->stokes[(size_t)k * nvals + c] = o[c];
        if (k == nfreq - 1) {
          dump->j = w.ej;
          dump->K = w.eK;
          dump->tau = w.tau;
        }
      }
    }
  }
}

void grtrans_ray_dump(const RunArgs &args, const FluidData &fd, const GeokerrArgs &gargs, int i,
                      const std::vector<double> &freqs, RayDump &d) {
  std::vector<double> mdots(1, args.mdotmin);
  std::vector<double> out((size_t)freqs.size() * (args.nvals + extra_count(args)));
  d = RayDump();
  drive_ray(args, fd, gargs, i, freqs, mdots, out.data(), nullptr, &d);
}

// pgrtrans.f90:12-253
void grtrans_main(const RunArgs &A, RunResult &res) {
  std::vector<double> freqs, mdots, mu0;
  compute_freqs(A, freqs, mdots, mu0);
  int nro = A.nn[0], nphi = A.nn[1], nup = A.nn[2];
  int npix = nro * nphi;
  int nparams = A.nmdot, nfreq = A.nfreq;
  int ncams = A.nmdot * A.nfreq * A.nmu * A.nt;
  res.nvals = A.nvals + extra_count(A);
  res.npix = npix;
  res.ncams = ncams;
  res.freqs = freqs;
  const int nout = A.nvals + extra_count(A);
  res.pixvals.assign((size_t)ncams * npix * nout, 0.f);
  res.pixloc.assign((size_t)npix * 2, 0.f);
  res.loop_seconds = 0;
  res.npoints = 0;
  FluidData fd;
  FluidArgs fa = A.fargs;
  fa.mbh = A.mbh;
  load_fluid_model(fd, A.fname, A.spin, fa);
  int i1 = A.i1, i2 = (A.i2 < 0) ? npix : A.i2;
  int nth = A.nthreads > 0 ? A.nthreads : omp_get_max_threads();
  for (int jm = 0; jm < A.nmu; jm++) {
    GeokerrArgs gargs;
    initialize_pixels(gargs, A.standard, mu0[jm], A.phi0, A.spin, A.uout, A.uin, A.rcut, A.nrotype, A.gridvals[0],
                      A.gridvals[1], A.gridvals[2], A.gridvals[3], nro, nphi, nup);
    if (A.nload > 1) {
      // pgrtrans.f90:172-191: one-point geodesics give each pixel's arrival-time offset t0.  (sic) initialize_geo_tabs
      // (geodesics.f90:113-128) zeroes kext / next / mufill in the shared gargs and only nup is restored afterwards, so the
      // main pass of an nload > 1 run has no mu-fill points.
      gargs.nup = 1;
      gargs.kext = 0; gargs.next = 0; gargs.mufill = 0;
#pragma omp parallel for num_threads(nth) schedule(static, 1)
      for (int i = i1; i <= i2; i++) {
        Geo g;
        int status;
        initialize_geodesic(g, gargs, i, status);
        gargs.t0[i - 1] = (g.npts == 0) ? 0.0 : g.x[0].d[0];
      }
      double mx = -1e300, mn = 1e300;
      for (double v : gargs.t0) mx = std::max(mx, v);
      for (double &v : gargs.t0) v = (v > 0.) ? v - mx : 0.;
      for (double v : gargs.t0) mn = std::min(mn, v);
      for (double &v : gargs.t0) v -= mn;
      gargs.nup = nup;
    }
    auto t0 = std::chrono::steady_clock::now();
    long long npoints = 0;
    for (int lt = 0; lt < A.nt; lt++) {
      int l = jm * A.nt + lt;  // 0-based (j-1)*nt+l-1
#pragma omp parallel num_threads(nth) reduction(+ : npoints)
      {
        std::vector<double> out((size_t)nparams * nfreq * nout);
#pragma omp for schedule(static, 1)
        for (int i = i1; i <= i2; i++) {
          drive_ray(A, fd, gargs, i, freqs, mdots, out.data(), &npoints, nullptr);
          for (int m = 0; m < nparams; m++)
            for (int k = 0; k < nfreq; k++) {
              size_t cam = (size_t)l * nfreq * nparams + (size_t)m * nfreq + k;
              for (int c = 0; c < nout; c++)
                res.pixvals[(cam * npix + (i - 1)) * nout + c] = (float)out[((size_t)m * nfreq + k) * nout + c];
            }
          if (l == 0) {
            res.pixloc[(size_t)(i - 1) * 2] = (float)gargs.alpha[i - 1];
            res.pixloc[(size_t)(i - 1) * 2 + 1] = (float)gargs.beta[i - 1];
          }
        }
      }
      if (lt < A.nt - 1) advance_fluid_timestep(fd, A.fname, A.spin, fa, A.dt);  // pgrtrans.f90:211
    }
    res.loop_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    res.npoints += npoints;
  }
}

}  // namespace orc
