// grtrans_b200 -- host-side input/output layer (no arithmetic of the hot path lives here).
//
//   build_camera      geokerr_wrapper.f:1-354 (INITIALIZE_CAMERA_GEOKERR) + geodesics.f90:64-111
//   load_fluid        fluid.f90:163-243 dispatch; readers fluid_model_ffjet.f90:181-210 (Fortran
//                     unformatted), fluid_model_harm.f90:275-408 (ASCII header + dump, MKS->KS->BL,
//                     LNRF velocities :468-485), harm3d container, SPHACC tables, thin-disk params
//   read_inputs_file  read_inputs.f90:47-100 (six namelists) ; read_files_in grtrans_program.f90
//   write_output      camera.f90:156-344 (FITS without cfitsio, or Fortran sequential unformatted)
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>

#include "grt_fluid.cuh"
#include "grt_host.h"
#include "grt_tables.h"

namespace grt_host {

static const double PI = 3.141592653589793238462643383279502884;

// ------------------------------------------------------------------------------------------ camera
void build_camera(Camera &c, int standard, double mu0, double phi0, double a, double uout, double rcut, int nrotype,
                  const double gv[4], int nro, int nphi, int nup) {
  const int n = nro * nphi;
  c.standard = standard; c.nro = nro; c.nphi = nphi; c.nup = nup;
  c.a = a; c.mu0 = mu0; c.phi0 = phi0; c.uout = uout;
  c.alpha.assign(n, 0); c.beta.assign(n, 0); c.q2.assign(n, 0); c.l.assign(n, 0); c.su.assign(n, 1.0);
  c.sm.assign(n, 0); c.uf.assign(n, 0); c.muf.assign(n, 0); c.tpm.assign(n, 0); c.tpr.assign(n, 0);
  if (standard == 1) {
    c.usegeor = 0;
    c.mufill = nup > 1 ? 1 : 0;
    c.kext = nup > 1 ? (int)std::ceil(3.0 * nup / 400.0) : 0;
    c.next = nup > 1 ? (int)std::ceil(60.0 * nup / 400.0) : 0;
  } else {
    c.usegeor = 1;
    c.mufill = c.kext = c.next = 0;
  }
  const double a1 = gv[0], a2 = gv[1], b1 = gv[2], b2 = gv[3];
  double abmax;
  std::vector<double> ro(nro + 1, 0.0);
  if (nrotype == 1) {  // circular grid (geokerr_wrapper.f:139-149)
    abmax = rcut * rcut;
    for (int i = 1; i <= nro; i++) ro[i] = a1 * std::pow(rcut / a1, (double)((float)i / (float)nro));
  } else {
    double ma = std::max(a1 * a1, a2 * a2), mb = std::max(b1 * b1, b2 * b2);
    abmax = ma * ma + mb * mb;
  }
  const double uplus = 1.0 / (1.0 + std::sqrt(1.0 - a * a));
  c.u0 = std::min(1.e-4, 1.0 / (100.0 * abmax));
  c.offset = 0.5;
  if (nup == 1) c.offset = (standard == 2) ? 0.0 : 1.e-8;
  double aa = a;  // the reference snaps tiny |a| to zero inside the loop, after the first q2
  for (int ii = 0; ii < n; ii++) {
    int i = ii / nphi, j = ii % nphi;
    double alpha, beta;
    if (nrotype > 1) {
      alpha = a1 + (a2 - a1) * ((double)i + 0.5) / (double)nro;
      beta = b1 + (b2 - b1) * ((double)j + 0.5) / (double)nphi;
    } else {
      double phi = (nphi != 1) ? 2.0 * PI * ((double)j + 0.5) / (double)nphi : 0.0;
      alpha = ro[i + 1] * std::cos(phi);
      beta = ro[i + 1] * std::sin(phi);
    }
    double l = -alpha * std::sqrt(1.0 - mu0 * mu0);
    double q2 = beta * beta - (aa * aa - alpha * alpha) * mu0 * mu0;
    if (std::fabs(q2) < 1e-10) q2 = 0.0;
    if (std::fabs(aa) < 1e-5) aa = 0.0;
    if (std::fabs(l) < 1e-5) l = 0.0;
    c.alpha[ii] = alpha; c.beta[ii] = beta; c.q2[ii] = q2; c.l[ii] = l; c.su[ii] = 1.0;
    if (standard == 2) {
      c.sm[ii] = (beta > 0.0 && mu0 < 1.0) ? 1.0 : -1.0;
      c.tpm[ii] = (int)(((std::signbit(mu0) ? -1.0 : 1.0) * c.sm[ii] + 1.0) / 2.0);
      c.muf[ii] = 0.0;
    } else {
      c.sm[ii] = (mu0 < 1.0 && beta >= 0.0) ? 1.0 : -1.0;
      c.uf[ii] = uplus;
      c.tpr[ii] = 1;
    }
  }
}

// ------------------------------------------------------------------------------------------ names
int fluid_model_id(const std::string &f) {
  if (f == "SPHACC") return 1;
  if (f == "THINDISK") return 2;
  if (f == "HARM") return 12;
  if (f == "FFJET") return 13;
  if (f == "HARM3D") return 19;
  if (f == "IHARM") return 29;
  if (f == "KORAL" || f == "KORALNTH") return 22;  // KORALNTH: f%model = KORAL + the non-thermal bins (fluid.f90:343-349)
  if (f == "KORAL3D") return 25;
  if (f == "KORAL3D_DISK") return 26;
  return -1;
}
// emis_model_bins (emis.f90:884-908): logarithmic bin edges between the real(4) limits of the fluid module; centres =
// geometric means of the edges (built by repeated multiplication, as written), widths = edge differences, the last one
// closing exactly on relgammamax
void emis_model_bins(double gmn, double gmx, int nb, std::vector<double> &bins) {
  bins.assign(2 * (size_t)std::max(nb, 0), 0.0);
  if (nb < 1) return;
  std::vector<double> edges(nb + 1);
  const double logbinspace = (std::log(gmx) - std::log(gmn)) / (double)(float)nb, binspace = std::exp(logbinspace);
  edges[0] = gmn;
  bins[0] = std::exp(std::log(gmn) + 0.5 * logbinspace);
  for (int i = 1; i < nb; i++) {
    edges[i] = edges[i - 1] * binspace;
    bins[i] = bins[i - 1] * binspace;
    bins[nb + i - 1] = edges[i] - edges[i - 1];
  }
  bins[2 * nb - 1] = gmx - edges[nb - 1];
}
int emis_type_id(const std::string &e, int *neq) {
  struct {
    const char *n;
    int t, q;
  } tab[] = {{"POLSYNCHTH", 2, 4}, {"SYNCHPL", 3, 1}, {"POLSYNCHPL", 4, 4}, {"BB", 8, 1}, {"BBPOL", 9, 4}, {"SYNCHTHAV", 13, 1},
             {"HYBRIDTHPL", 20, 4}, {"HYBRIDTH", 21, 4}, {"HYBRIDPL", 22, 4},
             {"BREMS", 5, 1}, {"SYNCHTHAVNOABS", 14, 1}, {"SYNCHTHBREMS", 27, 1}, {"POLSYNCHSYMTH", 15, 4}, {"MAXJUTT", 23, 4}, {"MAXCOMP", 24, 4}, {"RHO", 12, 1},
             {"BINS", 25, 1}, {"HYBRIDTHBINS", 26, 1}};
  for (auto &x : tab)
    if (e == x.n) {
      *neq = x.q;
      return x.t;
    }
  *neq = 1;
  return -1;
}
int integrator_flag(const std::string &i) {
  if (i == "lsoda") return 0;
  if (i == "delo") return 1;
  if (i == "formal") return 2;
  return -1;
}

void make_grids(const grt_args &a, std::vector<double> &freqs, std::vector<double> &mdots, std::vector<double> &mu0) {
  freqs.resize(a.nfreq); mdots.resize(a.nmdot); mu0.resize(a.nmu);
  for (int i = 0; i < a.nfreq; i++)
    freqs[i] = (a.nfreq == 1) ? a.fmin : a.fmin * std::exp(i * std::log(a.fmax / a.fmin) / (a.nfreq - 1));
  for (int i = 0; i < a.nmdot; i++)
    mdots[i] = (a.nmdot == 1) ? a.mdotmin : a.mdotmin * std::exp(i * std::log(a.mdotmax / a.mdotmin) / (a.nmdot - 1));
  for (int i = 0; i < a.nmu; i++) mu0[i] = (a.nmu == 1) ? a.mumin : a.mumin + (a.mumax - a.mumin) / (a.nmu - 1) * i;
}

// ------------------------------------------------------------------------------------------ tables
void build_emis_tables(EmisTablesHost &t) {
  auto lg = [](const double *src, int n, std::vector<double> &dst) {
    dst.resize(n);
    for (int i = 0; i < n; i++) dst[i] = std::log(src[i]);
  };
  lg(GRT_PL_GXVALS_RAW, 181, t.gx); lg(GRT_PL_GYVALS_RAW, 543, t.gy); lg(GRT_PL_XVALS_RAW, 81, t.xv);
  lg(GRT_PL_YAVALS_RAW, 243, t.ya); lg(GRT_PL_YPVALS_RAW, 243, t.yp); lg(GRT_PL_YVVALS_RAW, 243, t.yv);
  lg(GRT_PL_YAPVALS_RAW, 243, t.yap); lg(GRT_PL_YAVVALS_RAW, 243, t.yav);
  t.ch_mu.resize(21); t.ch_I.resize(21); t.ch_del.resize(21);
  for (int i = 0; i < 21; i++) {
    t.ch_mu[i] = (float)GRT_CH24_MU[i]; t.ch_I[i] = (float)GRT_CH24_I[i]; t.ch_del[i] = (float)GRT_CH24_DELTA[i];
  }
  // Faraday fits of polsynchth (jffunc with its high-frequency correction, the logarithm of shgmfunc;
  // polsynchemis.f90:737-761) depend on xarg alone and cost 3 exp + cos + log per (point, frequency).  They are tabulated
  // on 1e-4 <= xarg <= 1e5 as cubic Hermite pieces in u = log(xarg), node derivatives by the 4th-order central difference:
  // max |error| 4e-10 for jffunc (an O(1) function), 2e-11 for the logarithm (tests/test_host_emul.py checks it); outside
  // the range the kernels evaluate the fits directly.
  auto jff = [](double x) {
    const double e1 = .011 * std::exp(-x / 47.2);
    double j = 2.011 * std::exp(-std::pow(x, 1.035) / 4.7) - std::cos(0.5 * x) * std::exp(-std::pow(x, 1.2) / 2.73) - e1;
    const double tt = x / 120.0;
    if (tt > 0.1353352832366127) {
      double sw = 1.0;
      if (tt < 54.598150033144236) {
        const double t20 = std::pow(tt, 20.0);
        sw = t20 / (1.0 + t20);
      }
      const double c83 = 0.7937005259840998 / 67.44733739010397 * PI * 1e4;
      j += (e1 - c83 * std::pow(x, -8.0 / 3.0)) * sw;
    }
    return j;
  };
  auto lgt = [](double x) { return std::log(1.0 + (double)0.00185777f * std::pow(x, (double)1.50316886f)); };
  const int N = 2048;
  const double u0 = std::log(1e-4), u1 = std::log(1e5), h = (u1 - u0) / N;
  t.far_n = N; t.far_u0 = u0; t.far_ih = 1.0 / h;
  t.far.assign((size_t)N * 8, 0.0);
  for (int f = 0; f < 2; f++) {
    auto fn = [&](double u) { return f == 0 ? jff(std::exp(u)) : lgt(std::exp(u)); };
    std::vector<double> v(N + 1), d(N + 1);
    for (int k = 0; k <= N; k++) {
      const double u = u0 + k * h;
      v[k] = fn(u);
      d[k] = (-fn(u + 2 * h) + 8 * fn(u + h) - 8 * fn(u - h) + fn(u - 2 * h)) / 12.0;  // h f'(u)
    }
    for (int k = 0; k < N; k++) {
      double *q = &t.far[(size_t)k * 8 + 4 * f];
      q[0] = v[k];
      q[1] = d[k];
      q[2] = 3.0 * (v[k + 1] - v[k]) - (2.0 * d[k] + d[k + 1]);
      q[3] = 2.0 * (v[k] - v[k + 1]) + (d[k] + d[k + 1]);
    }
  }
}

// ------------------------------------------------------------------------------------------ fluid loaders
namespace {

// BL components from the dump's modified-KS components, then LNRF 3-velocity (all as the reader does:
// umksh2uks :73-98, uks2ubl kerr.f90:131-144, lnrf_frame_real kerr.f90:427-449)
void ingest_cell(FluidHost &f, size_t i, float r, float x2, float th, double rho, double uu, const double *uc,
                 const double *bc, bool three_d) {
  float dthdx2 = three_d ? (float)PI : (float)(PI * (1. + (double)(1.f - f.h) * std::cos(2. * PI * (double)x2)));
  double rr = (double)r, as = (double)f.asim;
  double delta = rr * rr - 2.0 * rr + as * as;
  double u[4], b[4];
  for (int c = 0; c < 2; c++) {
    const double *s = c ? bc : uc;
    double v1 = rr * s[1], v2 = s[2] * (double)dthdx2;
    double *d = c ? b : u;
    d[0] = s[0] - 2.0 * rr / delta * v1;
    d[1] = v1;
    d[2] = v2;
    d[3] = s[3] - as / delta * v1;
  }
  float vr = (float)(u[1] / u[0]), vt = (float)(u[2] / u[0]), om = (float)(u[3] / u[0]);
  float a = f.asim, mu = std::cos(th);
  float D = r * r - 2.f * r + a * a;
  float AR = (r * r + a * a) * (r * r + a * a) - a * a * D * (1.f - mu * mu);
  float RHO = r * r + a * a * mu * mu;
  float ENU = std::sqrt(D * RHO / AR), EMU1 = std::sqrt(RHO / D), EMU2 = std::sqrt(RHO);
  float EPSI = std::sqrt(1.f - mu * mu) * std::sqrt(AR / RHO), OM = 2.f * a * r / AR;
  float vrl = EMU1 / ENU * vr, vtl = EMU2 / ENU * vt, vpl = EPSI / ENU * (om - OM);
  if (!((double)D > 0.0)) vrl = vtl = vpl = 0.f;
  HostCell &c = f.cells[i];
  c.rho = (float)rho;
  c.p = (float)(uu * (double)(f.gam - 1.f));
  c.vrl = vrl; c.vtl = vtl; c.vpl = vpl;
  c.b0 = (float)b[0]; c.br = (float)b[1]; c.bth = (float)b[2]; c.bph = (float)b[3];
  c.u0 = (float)u[0];
  c.kela = c.be = 0.f;
}

int load_ffjet(FluidHost &f, const std::string &path) {
  FILE *fp = fopen(path.c_str(), "rb");
  if (!fp) return GRT_ERR_IO;
  std::vector<std::vector<char>> recs;
  for (;;) {
    int n1 = 0, n2 = 0;
    if (fread(&n1, 4, 1, fp) != 1) break;
    std::vector<char> b(n1);
    if (n1 && fread(b.data(), 1, n1, fp) != (size_t)n1) break;
    if (fread(&n2, 4, 1, fp) != 1 || n1 != n2) break;
    recs.push_back(std::move(b));
  }
  fclose(fp);
  if (recs.size() < 4 || recs[0].size() < 8) return GRT_ERR_IO;
  int nx;
  memcpy(&f.asim, recs[0].data(), 4);
  memcpy(&nx, recs[0].data() + 4, 4);
  int n1 = (int)std::sqrt((float)nx);
  if ((size_t)nx * 4 * 3 > recs[1].size() || (size_t)nx * 4 * 5 > recs[2].size() || (size_t)nx * 4 * 4 > recs[3].size())
    return GRT_ERR_IO;
  auto col = [&](const std::vector<char> &r, int k) { return reinterpret_cast<const float *>(r.data()) + (size_t)k * nx; };
  const float *rc = col(recs[1], 0), *thc = col(recs[1], 1), *rho = col(recs[1], 2);
  const float *b0 = col(recs[2], 1), *br = col(recs[2], 2), *bth = col(recs[2], 3), *bph = col(recs[2], 4);
  const float *u0 = col(recs[3], 0), *vr = col(recs[3], 1), *vth = col(recs[3], 2), *vph = col(recs[3], 3);
  f.nx1 = f.nx2 = n1;
  f.nx3 = 1;
  f.axr.resize(n1); f.ax1.resize(n1); f.ax2.resize(n1); f.axth.resize(n1);
  for (int i = 0; i < n1; i++) {
    f.axr[i] = rc[i];
    f.ax1[i] = std::log(rc[i]);
    f.ax2[i] = thc[(size_t)i * n1];
    f.axth[i] = f.ax2[i];
  }
  f.cells.resize(nx);
  for (int i = 0; i < nx; i++) {
    HostCell &c = f.cells[i];
    c.rho = rho[i]; c.p = 0.f;  // p_arr is allocated but never filled/read by the reference
    c.vrl = vr[i]; c.vtl = vth[i]; c.vpl = vph[i];
    c.b0 = b0[i]; c.br = br[i]; c.bth = bth[i]; c.bph = bph[i]; c.u0 = u0[i]; c.kela = c.be = 0.f;
  }
  return GRT_OK;
}

int load_harm2d(FluidHost &f, const grt_args &a) {
  std::ifstream hf(a.fhfile);
  if (!hf) return GRT_ERR_IO;
  double hd[26];
  for (double &v : hd)
    if (!(hf >> v)) return GRT_ERR_IO;
  f.nx1 = (int)hd[1]; f.nx2 = (int)hd[2]; f.nx3 = 1;
  f.asim = (float)hd[9]; f.gam = (float)hd[10]; f.h = (float)hd[24];
  if (std::fabs(f.asim - (float)a.spin) > 1e-4f) {
    fprintf(stderr, "ERROR -- Different simulation and grtrans spin values!\n");
    return GRT_ERR_BAD_ARG;
  }
  char app[16];
  snprintf(app, sizeof app, "%03d", a.findf);
  std::ifstream df(std::string(a.fdfile) + app);
  if (!df) return GRT_ERR_IO;
  for (double &v : hd)
    if (!(df >> v)) return GRT_ERR_IO;
  f.tcur = hd[0];  // fluid_model_harm.f90:328
  const size_t n = (size_t)f.nx1 * f.nx2;
  f.cells.resize(n);
  std::vector<float> x1(n), x2(n);
  for (size_t i = 0; i < n; i++) {
    double row[34];
    for (double &v : row)
      if (!(df >> v)) return GRT_ERR_IO;
    x1[i] = (float)row[0];
    x2[i] = (float)row[1];
    ingest_cell(f, i, (float)row[2], x2[i], (float)row[3], row[4], row[5], row + 13, row + 21, false);
  }
  f.ax1.resize(f.nx1); f.axr.resize(f.nx1); f.ax2.resize(f.nx2); f.axth.resize(f.nx2);
  for (int i = 0; i < f.nx2; i++) {
    f.ax2[i] = x2[i];
    f.axth[i] = (float)(PI * (double)f.ax2[i] + (double)((1.f - f.h) / 2.f) * std::sin(2. * PI * (double)f.ax2[i]));
  }
  for (int i = 0; i < f.nx1; i++) {
    f.ax1[i] = x1[(size_t)i * f.nx2];
    f.axr[i] = std::exp(f.ax1[i]);
  }
  return GRT_OK;
}

// synthetic 3-D container (see oracle/fluid.cc load_harm3d for the layout; written by tools/make_snapshot.py)
void harm3d_axes(FluidHost &f, const std::vector<float> &x1, const std::vector<float> &x2, const std::vector<float> &x3) {
  f.ax1.resize(f.nx1); f.axr.resize(f.nx1); f.ax2.resize(f.nx2); f.axth.resize(f.nx2); f.ax3.resize(f.nx3);
  for (int i = 0; i < f.nx3; i++) f.ax3[i] = x3[i];
  for (int i = 0; i < f.nx2; i++) {
    f.ax2[i] = x2[(size_t)i * f.nx3];
    f.axth[i] = (float)(PI * (double)f.ax2[i]);
  }
  for (int i = 0; i < f.nx1; i++) {
    f.ax1[i] = x1[(size_t)i * f.nx3 * f.nx2];
    f.axr[i] = std::exp(f.ax1[i]);
  }
}

// the reference's own harm3d input: header file of 15 numbers (fluid_model_harm3d.f90:337-366) + `<dfile>%03d.bin` =
// one text header line, then float32 data(35, n) (:516-584): x1,x2,x3 columns 4-6, rho 10, u 11, u^mu 19-22, b^mu 27-30
int load_harm3d_native(FluidHost &f, const grt_args &a, FILE *fp) {
  std::ifstream hf(a.fhfile);
  double header[15];
  bool ok = (bool)hf;
  for (int i = 0; ok && i < 15; i++) ok = (bool)(hf >> header[i]);
  if (!ok) {
    fclose(fp);
    return GRT_ERR_IO;
  }
  f.nx1 = (int)header[1]; f.nx2 = (int)header[2]; f.nx3 = (int)header[3];
  f.asim = (float)header[10]; f.gam = (float)header[11]; f.h = (float)header[13];
  if (std::fabs(f.asim - (float)a.spin) > 1e-4f) {
    fclose(fp);
    return GRT_ERR_BAD_ARG;
  }
  // dump time = first number of the text header line (what the reference's ASCII branch reads, :407; its binary branch
  // leaves tcur unassigned, so tstep is undefined there when nt > 1)
  int ch;
  std::string line;
  while ((ch = fgetc(fp)) != EOF && ch != 10) line.push_back((char)ch);
  f.tcur = atof(line.c_str());
  const size_t n = (size_t)f.nx1 * f.nx2 * f.nx3;
  std::vector<float> data(35 * n);
  ok = fread(data.data(), 4, data.size(), fp) == data.size();
  fclose(fp);
  if (!ok) return GRT_ERR_IO;
  std::vector<float> x1(n), x2(n), x3(n);
  f.cells.resize(n);
  for (size_t i = 0; i < n; i++) {
    const float *d = data.data() + 35 * i;
    x1[i] = d[3]; x2[i] = d[4]; x3[i] = d[5];
    double uc[4] = {d[18], d[19], d[20], d[21]}, bc[4] = {d[26], d[27], d[28], d[29]};
    double pp = d[10];
    ingest_cell(f, i, std::exp(x1[i]), x2[i], (float)((double)x2[i] * PI), (double)d[9], (pp > (double)1e-18f) ? pp : (double)1e-18f, uc, bc, true);
  }
  harm3d_axes(f, x1, x2, x3);
  return GRT_OK;
}

int load_harm3d(FluidHost &f, const grt_args &a) {
  {  // the reference's format if `<dfile>%03d.bin` exists, else the synthetic container (tools/make_snapshot.py)
    char append[16];
    snprintf(append, sizeof append, "%03d", a.findf);
    FILE *fn = fopen((std::string(a.fdfile) + append + ".bin").c_str(), "rb");
    if (fn) return load_harm3d_native(f, a, fn);
  }
  FILE *fp = fopen(a.fdfile, "rb");
  if (!fp) return GRT_ERR_IO;
  int dims[3];
  float par[3];
  if (fread(dims, 4, 3, fp) != 3 || fread(par, 4, 3, fp) != 3) {
    fclose(fp);
    return GRT_ERR_IO;
  }
  f.nx1 = dims[0]; f.nx2 = dims[1]; f.nx3 = dims[2];
  f.asim = par[0]; f.gam = par[1]; f.h = par[2];
  f.tcur = std::nan("");  // the container is a single dump without a time stamp
  if (std::fabs(f.asim - (float)a.spin) > 1e-4f) {
    fclose(fp);
    return GRT_ERR_BAD_ARG;
  }
  const size_t n = (size_t)f.nx1 * f.nx2 * f.nx3;
  std::vector<float> g[6];
  for (auto &v : g) {
    v.resize(n);
    if (fread(v.data(), 4, n, fp) != n) {
      fclose(fp);
      return GRT_ERR_IO;
    }
  }
  std::vector<double> rho(n), uu(n), uc(4 * n), bc(4 * n);
  bool ok = fread(rho.data(), 8, n, fp) == n && fread(uu.data(), 8, n, fp) == n && fread(uc.data(), 8, 4 * n, fp) == 4 * n &&
            fread(bc.data(), 8, 4 * n, fp) == 4 * n;
  fclose(fp);
  if (!ok) return GRT_ERR_IO;
  f.cells.resize(n);
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < (long long)n; i++)
    ingest_cell(f, (size_t)i, g[3][i], g[1][i], g[4][i], rho[i], uu[i], &uc[4 * i], &bc[4 * i], true);
  harm3d_axes(f, g[0], g[1], g[2]);
  return GRT_OK;
}

// iharm: header text file (read_iharm_data_header, fluid_model_iharm.f90:427-470) + `<dfile><indf %03d>.bin`, a stream of
// float32 data(13+eHEAT, n) = x1,x2,x3, rho, u, u^mu(4), b^mu(4) [, kel] per cell in (M)MKS components
// (read_iharm_data_file :508-651; load_iharm_data :684-747).  nt = 1.
int load_iharm(FluidHost &f, const grt_args &a) {
  std::ifstream hf(a.fhfile);
  if (!hf) return GRT_ERR_IO;
  std::vector<double> header(200, -1e20);
  for (int i = 0; i < 200; i++)
    if (!(hf >> header[i])) break;
  f.nx1 = (int)header[1]; f.nx2 = (int)header[2]; f.nx3 = (int)header[3];
  f.asim = (float)header[4];
  f.hslope = header[5];
  f.gam = (float)header[6];
  f.mks_smooth = header[7]; f.poly_xt = header[8]; f.poly_alpha = header[9]; f.startx1 = header[10];
  f.is_mmks = (int)header[11];
  f.eheat = (int)header[12];
  if (f.nx1 < 2 || f.nx2 < 2 || f.nx3 < 2) return GRT_ERR_IO;
  if (std::fabs(f.asim - (float)a.spin) > 1e-4f) {
    fprintf(stderr, "grtrans_b200: ERROR -- Different simulation and grtrans spin values! %g %g\n", f.asim, a.spin);
    return GRT_ERR_BAD_ARG;
  }
  f.tcur = header[0];  // :441 (the only assignment of tcur: every dump slot would carry the header file's time)
  char append[16];
  snprintf(append, sizeof append, "%03d", a.findf);
  std::string path = std::string(a.fdfile) + append + ".bin";
  FILE *fp = fopen(path.c_str(), "rb");
  if (!fp) return GRT_ERR_IO;
  const size_t n = (size_t)f.nx1 * f.nx2 * f.nx3;
  const int dlen = 13 + f.eheat;
  std::vector<float> data((size_t)dlen * n);
  bool ok = fread(data.data(), 4, data.size(), fp) == data.size();
  fclose(fp);
  if (!ok) return GRT_ERR_IO;
  grt::IharmGrid G{f.hslope, f.mks_smooth, f.poly_xt, f.poly_alpha, f.startx1};
  f.cells.resize(n);
  f.ax1.resize(f.nx1); f.axr.resize(f.nx1); f.ax2.resize(f.nx2); f.axth.resize(f.nx2); f.ax3.resize(f.nx3);
  for (size_t i = 0; i < n; i++) {
    const float *d = data.data() + (size_t)dlen * i;
    float x1 = d[0], x2 = d[1];
    float r = std::exp(x1);
    float th = (f.is_mmks == 1) ? grt::iharm_th_mmks(G, x2, x1) : (float)grt::iharm_th_mks(G, (double)x2);
    double um[4] = {d[5], d[6], d[7], d[8]}, bm[4] = {d[9], d[10], d[11], d[12]};
    double uks[4], bks[4];
    if (f.is_mmks == 1) {  // ummks2uks :165-184: central differences of the real(4) theta function
      double x1d = (double)x1, x2d = (double)x2;
      double dx2 = 1e-6 * x2d, dx1 = 1e-4 * x1d;
      double drdx1 = std::exp(x1d);
      double dthdx1 = (double)(grt::iharm_th_mmks(G, (float)x2d, (float)(x1d + .5 * dx1)) -
                               grt::iharm_th_mmks(G, (float)x2d, (float)(x1d - .5 * dx1))) / dx1;
      double dthdx2 = (double)(grt::iharm_th_mmks(G, (float)(x2d + .5 * dx2), (float)x1d) -
                               grt::iharm_th_mmks(G, (float)(x2d - .5 * dx2), (float)x1d)) / dx2;
      for (int c = 0; c < 2; c++) {
        const double *v = c ? bm : um;
        double *o = c ? bks : uks;
        o[0] = v[0];
        o[1] = drdx1 * v[1];
        o[2] = dthdx1 * v[1] + dthdx2 * v[2];
        o[3] = v[3];
      }
    } else {  // umksh2uks :186-220
      float dthdx2 = (float)(PI * (1. + (1. - f.hslope) * std::cos(2 * PI * (double)x2)));
      for (int c = 0; c < 2; c++) {
        const double *v = c ? bm : um;
        double *o = c ? bks : uks;
        o[0] = v[0];
        o[1] = (double)r * v[1];
        o[2] = v[2] * (double)dthdx2;
        o[3] = v[3];
      }
    }
    double rr = (double)r, as = (double)f.asim;
    double delta = rr * rr - 2.0 * rr + as * as;
    double u[4], b[4];
    for (int c = 0; c < 2; c++) {  // uks2ubl (kerr.f90:131-144)
      const double *v = c ? bks : uks;
      double *o = c ? b : u;
      o[0] = v[0] - 2.0 * rr / delta * v[1];
      o[1] = v[1];
      o[2] = v[2];
      o[3] = v[3] - as / delta * v[1];
    }
    float vr = (float)(u[1] / u[0]), vt = (float)(u[2] / u[0]), om = (float)(u[3] / u[0]);
    float aa = f.asim, mu = std::cos(th);  // lnrf_frame_real (kerr.f90:427-449)
    float D = r * r - 2.f * r + aa * aa;
    float AR = (r * r + aa * aa) * (r * r + aa * aa) - aa * aa * D * (1.f - mu * mu);
    float RHO = r * r + aa * aa * mu * mu;
    float ENU = std::sqrt(D * RHO / AR), EMU1 = std::sqrt(RHO / D), EMU2 = std::sqrt(RHO);
    float EPSI = std::sqrt(1.f - mu * mu) * std::sqrt(AR / RHO), OM = 2.f * aa * r / AR;
    float vrl = EMU1 / ENU * vr, vtl = EMU2 / ENU * vt, vpl = EPSI / ENU * (om - OM);
    if (!((double)D > 0.0)) vrl = vtl = vpl = 0.f;
    HostCell &c = f.cells[i];
    c.rho = d[3];
    c.p = d[4];  // internal energy u (the (gamma-1) factor is applied in the unit conversion, fluid.f90:1008)
    c.vrl = vrl; c.vtl = vtl; c.vpl = vpl;
    c.b0 = (float)b[0]; c.br = (float)b[1]; c.bth = (float)b[2]; c.bph = (float)b[3];
    c.u0 = (float)u[0];
    c.kela = (f.eheat == 1) ? d[13] : 0.f;
    c.be = 0.f;
  }
  for (int i = 0; i < f.nx3; i++) f.ax3[i] = data[(size_t)dlen * i + 2];
  for (int i = 0; i < f.nx2; i++) {
    f.ax2[i] = data[(size_t)dlen * ((size_t)i * f.nx3) + 1];
    f.axth[i] = 0.f;  // (uniqth is only built for BL == 1, never set by the reader)
  }
  for (int i = 0; i < f.nx1; i++) {
    f.ax1[i] = data[(size_t)dlen * ((size_t)i * f.nx3 * f.nx2)];
    f.axr[i] = std::exp(f.ax1[i]);
  }
  return GRT_OK;
}

// koral (2-D, fluid_model_koral.f90:296-488) and koral3d "shortfile" (fluid_model_koral3d.f90:359-632) inputs: a header
// file of 10 / 11 numbers and an ASCII dump `<dfile>%03d` / `<dfile>%04d` (its own header line, then 38 / 23 columns per
// cell, x2 fastest, then x1, then x3).  u^mu, b^mu are Boyer-Lindquist components already; rho, b are scaled by scalefac.
int load_koral(FluidHost &f, const grt_args &a) {
  const bool d3 = f.model != 22;
  std::ifstream hf(a.fhfile);
  if (!hf) return GRT_ERR_IO;
  const int nhead = d3 ? 11 : 10;
  std::vector<double> header(nhead);
  for (int i = 0; i < nhead; i++)
    if (!(hf >> header[i])) return GRT_ERR_IO;
  const int o = d3 ? 1 : 0;
  f.nx1 = (int)header[1]; f.nx2 = (int)header[2]; f.nx3 = d3 ? (int)header[3] : 1;
  f.asim = (float)header[3 + o];
  f.kr0 = (float)header[5 + o]; f.kh = (float)header[6 + o]; f.kaa = (float)header[7 + o]; f.kbb = (float)header[8 + o];
  f.kpp = (float)header[9 + o];
  f.ktype = (f.model == 26) ? 1 : 0;
  const bool doingnth = std::string(a.fname) == "KORALNTH";
  if (doingnth) {  // second header line: nrelbin, relgammamin, relgammamax (fluid_model_koral.f90:318-326)
    double h2[3];
    for (int i = 0; i < 3; i++)
      if (!(hf >> h2[i])) return GRT_ERR_IO;
    f.nrelbin = (int)h2[0]; f.relgammamin = (float)h2[1]; f.relgammamax = (float)h2[2];
    if (f.nrelbin < 1) return GRT_ERR_IO;
  }
  if (f.nx1 < 2 || f.nx2 < 8 || (d3 && f.nx3 < 2)) return GRT_ERR_IO;
  if (std::fabs(f.asim - (float)a.spin) > 1e-4f) return GRT_ERR_BAD_ARG;
  char append[16];
  snprintf(append, sizeof append, d3 ? "%04d" : "%03d", a.findf);
  std::ifstream df(std::string(a.fdfile) + append);
  if (!df) return GRT_ERR_IO;
  double tmp;
  for (int i = 0; i < nhead + (doingnth ? 3 : 0); i++) {
    if (!(df >> tmp)) return GRT_ERR_IO;
    if (i == 0) f.tcur = tmp;  // tcur=header(1)
  }
  const size_t n = (size_t)f.nx1 * f.nx2 * f.nx3;
  // (sic) KORALNTH rows hold 41 + nrelbin numbers and bin ie is read from column nthpos + ie - 1 with nthpos = 43
  // (fluid_model_koral.f90:345-355, 400): the last bin lies one element past the row, which in the column-major
  // data(dlen, nx2) block is the first column of the next row of the block (past the allocation for the block's last row:
  // undefined in the reference, 0 here)
  const int nthpos = 43;
  if (doingnth) f.nth.assign(n * (size_t)f.nrelbin, 0.f);
  const int dlen = d3 ? 23 : (doingnth ? 41 + f.nrelbin : 38), rhopos = 10, vpos = 12, bpos = d3 ? 16 : 25, ppos = d3 ? 21 : 33, gridpos = 4, bepos = 22;
  const double scalefac = a.fscalefac, sq = std::sqrt(scalefac);
  std::vector<double> row(dlen);
  std::vector<float> x1(n), x2(n), x3(n);
  f.cells.resize(n);
  for (size_t i = 0; i < n; i++) {
    for (int c = 0; c < dlen; c++)
      if (!(df >> row[c])) return GRT_ERR_IO;
    x1[i] = (float)row[gridpos - 1]; x2[i] = (float)row[gridpos];
    float r, th;
    if (d3) {
      x3[i] = (float)row[gridpos + 1]; r = (float)row[gridpos + 2]; th = (float)row[gridpos + 3];
    } else {
      r = (float)row[gridpos + 1]; th = (float)row[gridpos + 2];
    }
    const double *u = &row[vpos - 1], *b = &row[bpos - 1];
    float vr = (float)(u[1] / u[0]), vt = (float)(u[2] / u[0]), om = (float)(u[3] / u[0]);
    float aa = f.asim, mu = std::cos(th);  // lnrf_frame_real (kerr.f90:427-449)
    float D = r * r - 2.f * r + aa * aa;
    float AR = (r * r + aa * aa) * (r * r + aa * aa) - aa * aa * D * (1.f - mu * mu);
    float RHO = r * r + aa * aa * mu * mu;
    float ENU = std::sqrt(D * RHO / AR), EMU1 = std::sqrt(RHO / D), EMU2 = std::sqrt(RHO);
    float EPSI = std::sqrt(1.f - mu * mu) * std::sqrt(AR / RHO), OM = 2.f * aa * r / AR;
    float vrl = EMU1 / ENU * vr, vtl = EMU2 / ENU * vt, vpl = EPSI / ENU * (om - OM);
    if (!((double)D > 0.0)) vrl = vtl = vpl = 0.f;
    HostCell &c = f.cells[i];
    c.rho = (float)(row[rhopos - 1] * scalefac);
    c.p = (float)row[ppos - 1];
    c.be = d3 ? (float)row[bepos - 1] : 0.f;
    c.b0 = (float)(b[0] * sq); c.br = (float)(b[1] * sq); c.bth = (float)(b[2] * sq); c.bph = (float)(b[3] * sq);
    c.u0 = (float)u[0];
    c.vrl = vrl; c.vtl = vtl; c.vpl = vpl;
    c.kela = 0.f;
    if (doingnth) {
      for (int ie = 0; ie < f.nrelbin; ie++) {
        const int col = nthpos + ie;  // 1-based
        if (col <= dlen) f.nth[i * (size_t)f.nrelbin + ie] = (float)(row[col - 1] * scalefac);
      }
      if (i > 0 && (i % (size_t)f.nx2) != 0) f.nth[(i - 1) * (size_t)f.nrelbin + (f.nrelbin - 1)] = (float)(row[0] * scalefac);
    }
  }
  f.ax1.resize(f.nx1); f.axr.resize(f.nx1); f.ax2.resize(f.nx2); f.axth.assign(f.nx2, 0.f); f.ax3.assign(std::max(f.nx3, 2), 0.f);
  for (int i = 0; i < f.nx2; i++) f.ax2[i] = x2[i];
  for (int i = 0; i < f.nx1; i++) {
    f.ax1[i] = x1[(size_t)i * f.nx2];
    f.axr[i] = f.kr0 + std::exp(f.ax1[i]);
  }
  if (d3)
    for (int i = 0; i < f.nx3; i++) f.ax3[i] = x3[(size_t)i * f.nx2 * f.nx1];
  const int minpolecell = 4;
  f.x2mintrust = f.ax2[minpolecell - 1];
  f.x2maxtrust = f.ax2[f.nx2 - (minpolecell - 1) - 1];
  return GRT_OK;
}

}  // namespace

namespace {
int load_one_dump(FluidHost &f, const grt_args &a);

bool has_time_axis(int model) { return model == 12 || model == 19 || model == 22 || model == 25 || model == 26 || model == 29; }

// load_*_data(count): dump files indf, indf-1, ... into slots 0..count-1; tstep = t(1)-t(2) when more than one is read
int read_dump_slots(FluidHost &f, const grt_args &a, int count) {
  for (int k = 0; k < count; k++) {
    grt_args b = a;
    b.findf = f.indf - k;
    FluidHost one;
    int rc = load_one_dump(one, b);
    if (rc != GRT_OK) return rc;
    if (one.cells.size() != f.ncell) return GRT_ERR_IO;
    if (f.nslots > 1 && std::isnan(one.tcur)) return GRT_ERR_BAD_ARG;
    std::copy(one.cells.begin(), one.cells.end(), f.cells.begin() + (size_t)k * f.ncell);
    f.tdump[k] = one.tcur;
  }
  if (count > 1) {
    f.tstep_d = f.tdump[0] - f.tdump[1];
    f.tstep = (float)f.tstep_d;
  }
  return GRT_OK;
}
}  // namespace

int load_fluid(FluidHost &f, const grt_args &a) {
  int rc = load_one_dump(f, a);
  if (rc != GRT_OK) return rc;
  f.indf = a.findf;
  f.ncell = f.cells.size();
  f.nslots = 1;
  f.tdump.assign(1, f.tcur);
  if (has_time_axis(f.model) && a.fnt > 1) {
    // iharm: t(k) is the header file's time for every slot, so tstep = 0 and the slot index is undefined in the reference
    if (f.model == 29) return GRT_ERR_BAD_ARG;
    if (f.nrelbin > 0) return GRT_ERR_BAD_ARG;  // KORALNTH: one dump resident
    f.nslots = a.fnt;
    f.tdump.assign(a.fnt, 0.0);
    f.cells.resize(f.ncell * (size_t)a.fnt);
    rc = read_dump_slots(f, a, a.fnt);
  }
  return rc;
}

// advance_*_timestep + update_*_data (fluid_model_harm.f90:511-546; the same code in the harm3d, iharm, koral3d modules).
// Mixed real(4)/real(8) arithmetic as declared there.  (sic) the reference shifts every field of the older slots except
// u0_arr, so after an update by 0 < nupdate <= nt dumps those slots keep the u^t they held before.
int advance_fluid(FluidHost &f, const grt_args &a, double dtobs) {
  if (!has_time_axis(f.model)) return GRT_OK;
  int nupdate;
  if (f.model == 29) {
    nupdate = (int)std::floor(dtobs / f.tstep_d);
    f.toffset_d = f.toffset_d + dtobs - f.tstep_d * nupdate;
    nupdate += (int)std::floor(f.toffset_d / f.tstep_d);
    f.toffset_d -= std::floor(f.toffset_d / f.tstep_d) * f.tstep_d;
  } else {
    nupdate = (int)std::floor(dtobs / (double)f.tstep);
    f.toffset = (float)((double)f.toffset + dtobs - (double)(f.tstep * (float)nupdate));
    nupdate += (int)std::floor(f.toffset / f.tstep);
    f.toffset = f.toffset - (float)(int)std::floor(f.toffset / f.tstep) * f.tstep;
  }
  f.indf += nupdate;
  if (nupdate > f.nslots) return read_dump_slots(f, a, f.nslots);
  if (nupdate <= 0) return GRT_OK;
  const size_t keep = (size_t)(f.nslots - nupdate) * f.ncell, sh = (size_t)nupdate * f.ncell;
  std::vector<float> u0(keep);
  for (size_t i = 0; i < keep; i++) u0[i] = f.cells[sh + i].u0;
  std::copy_backward(f.cells.begin(), f.cells.begin() + keep, f.cells.end());
  for (size_t i = 0; i < keep; i++) f.cells[sh + i].u0 = u0[i];
  return read_dump_slots(f, a, nupdate);
}

namespace {
int load_one_dump(FluidHost &f, const grt_args &a) {
  f = FluidHost();
  f.model = fluid_model_id(a.fname);
  switch (f.model) {
    case 1: {  // SPHACC (fluid_model_sphacc.f90:12-443)
      f.uvals.resize(498); f.vvals.resize(498); f.uv.resize(461); f.tvals.resize(461);
      for (int i = 0; i < 498; i++) {
        f.uvals[i] = (float)i * .001f + 1.f / 399.f;
        f.vvals[i] = (float)GRT_SPHACC_VVALS[i];
      }
      for (int i = 0; i < 461; i++) {
        f.uv[i] = (float)(1.0 / (double)(float)GRT_SPHACC_RVALS[i]);
        f.tvals[i] = (float)GRT_SPHACC_TVALS[i];
      }
      return GRT_OK;
    }
    case 2:  // THINDISK: init_thindisk(real(a),ifile,real(mdot),real(mbh),real(rin),real(rout)) fluid.f90:226-227
      f.td_mdot = (float)a.fmdot; f.td_mbh = (float)a.mbh; f.td_rin = (float)a.frin; f.td_rout = (float)a.frout;
      return GRT_OK;
    case 12: return load_harm2d(f, a);
    case 13: return load_ffjet(f, a.fdfile);
    case 19: return load_harm3d(f, a);
    case 29: return load_iharm(f, a);
    case 22:
    case 25:
    case 26: return load_koral(f, a);
    default: return GRT_ERR_BAD_ARG;
  }
}
}  // namespace

// ------------------------------------------------------------------------------------------ namelists
namespace {
typedef std::map<std::string, std::vector<std::string>> NmlGroup;

std::string lower(std::string s) {
  for (auto &c : s) c = (char)tolower((unsigned char)c);
  return s;
}

// parse "&name key=v1,v2, key2='str', /" groups
bool parse_namelists(const std::string &text, std::map<std::string, NmlGroup> &out) {
  size_t i = 0, n = text.size();
  while (i < n) {
    while (i < n && text[i] != '&') i++;
    if (i >= n) break;
    size_t j = ++i;
    while (j < n && (isalnum((unsigned char)text[j]) || text[j] == '_')) j++;
    std::string gname = lower(text.substr(i, j - i));
    i = j;
    NmlGroup grp;
    std::string key;
    while (i < n) {
      while (i < n && (isspace((unsigned char)text[i]) || text[i] == ',')) i++;
      if (i >= n) break;
      if (text[i] == '/') {
        i++;
        break;
      }
      if (text[i] == '!') {
        while (i < n && text[i] != '\n') i++;
        continue;
      }
      // token: quoted string or bare
      std::string tok;
      if (text[i] == '\'' || text[i] == '"') {
        char q = text[i++];
        while (i < n && text[i] != q) tok += text[i++];
        if (i < n) i++;
        if (!key.empty()) grp[key].push_back(tok);
        continue;
      }
      size_t s = i;
      while (i < n && !isspace((unsigned char)text[i]) && text[i] != ',' && text[i] != '=' && text[i] != '/') i++;
      tok = text.substr(s, i - s);
      size_t k = i;
      while (k < n && isspace((unsigned char)text[k])) k++;
      if (k < n && text[k] == '=') {
        key = lower(tok);
        size_t par = key.find('(');
        if (par != std::string::npos) key = key.substr(0, par);
        grp[key];
        i = k + 1;
      } else if (!key.empty()) {
        grp[key].push_back(tok);
      }
    }
    out[gname] = grp;
  }
  return !out.empty();
}

double to_d(std::string s) {
  for (auto &c : s)
    if (c == 'd' || c == 'D') c = 'e';
  return atof(s.c_str());
}
void get(const NmlGroup &g, const char *k, double &v) {
  auto it = g.find(k);
  if (it != g.end() && !it->second.empty()) v = to_d(it->second[0]);
}
void get(const NmlGroup &g, const char *k, int &v) {
  auto it = g.find(k);
  if (it != g.end() && !it->second.empty()) v = (int)to_d(it->second[0]);
}
void getl(const NmlGroup &g, const char *k, int &v) {
  auto it = g.find(k);
  if (it != g.end() && !it->second.empty()) {
    std::string s = lower(it->second[0]);
    v = (s.find('t') != std::string::npos) ? 1 : 0;
  }
}
template <size_t N>
void gets(const NmlGroup &g, const char *k, char (&dst)[N]) {
  auto it = g.find(k);
  if (it != g.end() && !it->second.empty()) {
    std::string s = it->second[0];
    while (!s.empty() && s.back() == ' ') s.pop_back();
    snprintf(dst, N, "%s", s.c_str());
  }
}
}  // namespace

int read_inputs_file(const char *ifile, grt_args *a) {
  std::ifstream f(ifile);
  if (!f) return GRT_ERR_IO;
  std::stringstream ss;
  ss << f.rdbuf();
  std::map<std::string, NmlGroup> nl;
  if (!parse_namelists(ss.str(), nl)) return GRT_ERR_IO;
  grt_args_defaults(a);
  const NmlGroup &g = nl["geodata"], &fl = nl["fluiddata"], &e = nl["emisdata"], &ge = nl["general"], &h = nl["harm"],
                 &an = nl["analytic"];
  get(g, "standard", a->standard); get(g, "mumin", a->mumin); get(g, "mumax", a->mumax); get(g, "nmu", a->nmu);
  get(g, "phi0", a->phi0); get(g, "spin", a->spin); get(g, "uout", a->uout); get(g, "uin", a->uin);
  get(g, "rcut", a->rcut); get(g, "nrotype", a->nrotype); get(g, "i1", a->i1); get(g, "i2", a->i2);
  get(g, "extra", a->extra); get(g, "debug", a->debug);
  {
    auto it = g.find("gridvals");
    if (it != g.end())
      for (size_t k = 0; k < 4 && k < it->second.size(); k++) a->gridvals[k] = to_d(it->second[k]);
    it = g.find("nn");
    if (it != g.end())
      for (size_t k = 0; k < 3 && k < it->second.size(); k++) a->nn[k] = (int)to_d(it->second[k]);
  }
  gets(fl, "fname", a->fname); get(fl, "dt", a->dt); get(fl, "nt", a->nt); get(fl, "nload", a->nload);
  get(fl, "nmdot", a->nmdot); get(fl, "mdotmin", a->mdotmin); get(fl, "mdotmax", a->mdotmax); get(fl, "sigcut", a->sigcut);
  gets(e, "ename", a->ename); get(e, "mbh", a->mbh); get(e, "nfreq", a->nfreq); get(e, "fmin", a->fmin);
  get(e, "fmax", a->fmax); get(e, "muval", a->muval); get(e, "gmin", a->gmin); get(e, "gmax", a->gmax);
  get(e, "p1", a->p1); get(e, "p2", a->p2); get(e, "jetalpha", a->jetalpha); gets(e, "stype", a->stype);
  double delta = 1.0;
  int nweights = 1;
  get(e, "delta", delta); get(e, "nweights", nweights);
  {
    auto it = e.find("coefindx");
    if (it != e.end())
      for (size_t k = 0; k < 7 && k < it->second.size(); k++) a->epcoefindx[k] = (int)to_d(it->second[k]);
  }
  getl(ge, "use_geokerr", a->use_geokerr); get(ge, "nvals", a->nvals); gets(ge, "iname", a->iname);
  get(ge, "cflag", a->cflag);
  gets(h, "fdfile", a->fdfile); gets(h, "fgfile", a->fgfile); gets(h, "fhfile", a->fhfile); get(h, "fnt", a->fnt);
  get(h, "fnfiles", a->fnfiles); get(h, "findf", a->findf); get(h, "fjonfix", a->fjonfix); get(h, "foffset", a->foffset);
  gets(h, "fsim", a->fsim); get(h, "fdindf", a->fdindf); get(h, "fmagcrit", a->fmagcrit); get(h, "fscalefac", a->fscalefac);
  get(an, "fnw", a->fnw); get(an, "fwmin", a->fwmin); get(an, "fwmax", a->fwmax); get(an, "fnfreq_tab", a->fnfreq_tab);
  get(an, "ffmin", a->ffmin); get(an, "ffmax", a->ffmax); get(an, "frmax", a->frmax); get(an, "fnr", a->fnr);
  get(an, "fsigt", a->fsigt); get(an, "ffcol", a->ffcol); get(an, "frspot", a->frspot); get(an, "fr0spot", a->fr0spot);
  get(an, "fn0spot", a->fn0spot); get(an, "ftscl", a->ftscl); get(an, "frscl", a->frscl); get(an, "fmdot", a->fmdot);
  get(an, "fnscl", a->fnscl); get(an, "fnnthscl", a->fnnthscl); get(an, "fnnthp", a->fnnthp); get(an, "fbeta", a->fbeta);
  get(an, "fbl06", a->fbl06); get(an, "fnp", a->fnp); get(an, "ftp", a->ftp); get(an, "frin", a->frin);
  get(an, "frout", a->frout); get(an, "fthin", a->fthin); get(an, "fthout", a->fthout); get(an, "fphiin", a->fphiin);
  get(an, "fphiout", a->fphiout);
  // epotherargs (read_inputs.f90:86-98)
  if (nweights == 1) {
    a->nepotherargs = 2;
    a->epotherargs[0] = delta;
    a->epotherargs[1] = 1.;
  } else {
    const float w[6] = {0.10714295f, 0.4459917f, 1.34219997f, 2.17450422f, 1.36515333f, 0.52892829f};
    a->nepotherargs = 7;
    a->epotherargs[0] = delta;
    for (int k = 0; k < 6; k++) a->epotherargs[1 + k] = w[k];
  }
  return GRT_OK;
}

int read_files_in(const char *path, std::string &ifile, std::string &ofile) {
  std::ifstream f(path);
  if (!f) return GRT_ERR_IO;
  std::stringstream ss;
  ss << f.rdbuf();
  std::map<std::string, NmlGroup> nl;
  if (!parse_namelists(ss.str(), nl) || !nl.count("files")) return GRT_ERR_IO;
  const NmlGroup &g = nl["files"];
  auto a = g.find("ifile"), b = g.find("ofile");
  if (a == g.end() || b == g.end() || a->second.empty() || b->second.empty()) return GRT_ERR_IO;
  ifile = a->second[0];
  ofile = b->second[0];
  return GRT_OK;
}

// ------------------------------------------------------------------------------------------ output
namespace {
void card(std::string &hdr, const std::string &s) {
  std::string c = s;
  c.resize(80, ' ');
  hdr += c;
}
std::string kw(const std::string &name) {
  std::string n = name;
  while (!n.empty() && n[0] == ' ') n.erase(0, 1);
  for (auto &c : n) c = (char)toupper((unsigned char)c);
  if (n.size() <= 8 && n.find('(') == std::string::npos && n.find('_') == std::string::npos) {
    n.resize(8, ' ');
    return n + "= ";
  }
  return "HIERARCH " + n + " = ";
}
void card_int(std::string &h, const std::string &name, long v, const std::string &com) {
  char b[96];
  std::string k = kw(name);
  if (k.size() == 10)
    snprintf(b, sizeof b, "%s%20ld / %s", k.c_str(), v, com.c_str());
  else
    snprintf(b, sizeof b, "%s%ld / %s", k.c_str(), v, com.c_str());
  card(h, b);
}
void card_dbl(std::string &h, const std::string &name, double v, int dec, bool dexp, const std::string &com) {
  char num[48], b[120];
  snprintf(num, sizeof num, "%.*E", dec, v);
  if (dexp)
    for (char *p = num; *p; p++)
      if (*p == 'E') *p = 'D';
  std::string k = kw(name);
  if (k.size() == 10)
    snprintf(b, sizeof b, "%s%20s / %s", k.c_str(), num, com.c_str());
  else
    snprintf(b, sizeof b, "%s%s / %s", k.c_str(), num, com.c_str());
  card(h, b);
}
void card_str(std::string &h, const std::string &name, const std::string &v, const std::string &com) {
  std::string val = v;
  if (val.size() < 8) val.resize(8, ' ');
  std::string s = kw(name) + "'" + val + "' / " + com;
  card(h, s);
}
void card_log(std::string &h, const std::string &name, bool v, const std::string &com) {
  char b[96];
  snprintf(b, sizeof b, "%s%20s / %s", kw(name).c_str(), v ? "T" : "F", com.c_str());
  card(h, b);
}
void pad2880(std::string &s, char fill) {
  size_t r = s.size() % 2880;
  if (r) s.append(2880 - r, fill);
}
void put_be_floats(std::string &out, const float *v, size_t n) {
  size_t off = out.size();
  out.resize(off + 4 * n);
  for (size_t i = 0; i < n; i++) {
    unsigned int u;
    memcpy(&u, v + i, 4);
    out[off + 4 * i] = (char)(u >> 24);
    out[off + 4 * i + 1] = (char)(u >> 16);
    out[off + 4 * i + 2] = (char)(u >> 8);
    out[off + 4 * i + 3] = (char)u;
  }
}
// checked output: any short write marks the writer as failed (a full disk must not produce a truncated file + GRT_OK)
struct OutFile {
  FILE *fp = nullptr;
  bool ok = true;
  void put(const void *p, size_t n) {
    if (ok && n && fwrite(p, 1, n, fp) != n) ok = false;
  }
};
// one Fortran sequential-unformatted record.  gfortran's 4-byte markers cannot describe 2 GiB or more: such a record is
// written as subrecords of at most 2147483639 bytes whose leading marker is negative when another subrecord follows and
// whose trailing marker is negative when one precedes (libgfortran io/unix.c, the format gfortran reads back).
void frec(OutFile &f, const void *p, size_t nbytes) {
  const size_t MAXSUB = 2147483639u;
  const char *c = (const char *)p;
  size_t done = 0;
  bool first = true;
  do {
    const size_t n = std::min(MAXSUB, nbytes - done);
    const bool more = done + n < nbytes;
    const int head = more ? -(int)n : (int)n, tail = first ? (int)n : -(int)n;
    f.put(&head, 4);
    f.put(c + done, n);
    f.put(&tail, 4);
    done += n;
    first = false;
  } while (done < nbytes);
}
// real(8) image values (exact copies of the real(4) pixel values) -> real(4), in chunks
template <class F>
void for_float_chunks(const double *v, size_t n, F &&fn) {
  const size_t CH = 1u << 20;
  std::vector<float> buf(std::min(n, CH));
  for (size_t off = 0; off < n; off += CH) {
    const size_t m = std::min(CH, n - off);
    for (size_t i = 0; i < m; i++) buf[i] = (float)v[off + i];
    fn(buf.data(), m);
  }
}
}  // namespace

int write_output(const grt_args &a, const char *outfile, int nx, int ny, int nv, int ncams, const float *pixloc,
                 const double *ivals, const std::vector<double> &freqs) {
  const size_t npix = (size_t)nx * ny;
  OutFile f;
  f.fp = fopen(outfile, "wb");
  if (!f.fp) return GRT_ERR_IO;
  if (a.cflag == 1) {
    std::string hdu;
    // primary: FTPHPR(simple=T, bitpix=-32, naxis=1, naxes=[2*npix], pcount=0, gcount=1, extend=T)
    card_log(hdu, "SIMPLE", true, "file does conform to FITS standard");
    card_int(hdu, "BITPIX", -32, "number of bits per data pixel");
    card_int(hdu, "NAXIS", 1, "number of data axes");
    card_int(hdu, "NAXIS1", (long)(2 * npix), "length of data axis 1");
    card_log(hdu, "EXTEND", true, "FITS dataset may contain extensions");
    card(hdu, "END");
    pad2880(hdu, ' ');
    std::string data;
    put_be_floats(data, pixloc, 2 * npix);
    pad2880(data, '\0');
    f.put(hdu.data(), hdu.size());
    f.put(data.data(), data.size());
    for (int m = 0; m < ncams; m++) {
      std::string h;
      card_str(h, "XTENSION", "IMAGE", "IMAGE extension");
      card_int(h, "BITPIX", -32, "number of bits per data pixel");
      card_int(h, "NAXIS", 1, "number of data axes");
      card_int(h, "NAXIS1", (long)(npix * nv), "length of data axis 1");
      card_int(h, "PCOUNT", 0, "required keyword; must = 0");
      card_int(h, "GCOUNT", 1, "required keyword; must = 1");
      // cards 6,7,8: nx, ny, nu as float32 with 6 decimals (ftpkye)
      card_dbl(h, "NX", (double)(float)nx, 6, false, "# x pixels");
      card_dbl(h, "NY", (double)(float)ny, 6, false, "# y pixels");
      card_dbl(h, "NU", (double)(float)freqs[m % a.nfreq], 6, false, "Frequency (Hz)");
      card_int(h, "standard", a.standard, "standard");
      card_dbl(h, "mumin", a.mumin, 6, true, "mumin"); card_dbl(h, "mumax", a.mumax, 6, true, "mumax");
      card_int(h, "nmu", a.nmu, "nmu");
      card_dbl(h, "phi0", a.phi0, 6, true, "phi0"); card_dbl(h, "spin", a.spin, 6, true, "spin");
      card_dbl(h, "uout", a.uout, 6, true, "uout"); card_dbl(h, "uin", a.uin, 6, true, "uin");
      card_dbl(h, "rcut", a.rcut, 6, true, "rcut");
      card_int(h, "nrotype", a.nrotype, "nrotype");
      for (int k = 0; k < 4; k++) {
        char nm[24];
        snprintf(nm, sizeof nm, "gridvals(%d)", k + 1);
        card_dbl(h, nm, a.gridvals[k], 6, true, nm);
      }
      for (int k = 0; k < 3; k++) {
        char nm[24];
        snprintf(nm, sizeof nm, "nn(%d)", k + 1);
        card_int(h, nm, a.nn[k], nm);
      }
      card_str(h, "fname", a.fname, "fname");
      card_dbl(h, "dt", a.dt, 6, true, "dt"); card_int(h, "nt", a.nt, "nt"); card_int(h, "nload", a.nload, "nload");
      card_int(h, "nmdot", a.nmdot, "nmdot");
      card_dbl(h, "mdotmin", a.mdotmin, 6, true, "mdotmin"); card_dbl(h, "mdotmax", a.mdotmax, 6, true, "mdotmax");
      card_str(h, "ename", a.ename, "ename");
      card_dbl(h, "mbh", a.mbh, 6, true, "mbh"); card_int(h, "nfreq", a.nfreq, "nfreq");
      card_dbl(h, "fmin", a.fmin, 6, true, "fmin"); card_dbl(h, "fmax", a.fmax, 6, true, "fmax");
      card_dbl(h, "muval", a.muval, 6, true, "muval"); card_dbl(h, "gmin", a.gmin, 6, true, "gmin");
      card_dbl(h, "gmax", a.gmax, 6, true, "gmax"); card_dbl(h, "p1", a.p1, 6, true, "p1");
      card_dbl(h, "p2", a.p2, 6, true, "p2"); card_dbl(h, "jetalpha", a.jetalpha, 6, true, "jetalpha");
      card_str(h, "stype", a.stype, "stype");
      card_str(h, "use_geokerr", a.use_geokerr ? "T" : "F", "use_geokerr");
      card_int(h, "nvals", a.nvals, "nvals");
      card_str(h, "iname", a.iname, "iname");
      card_int(h, "cflag", a.cflag, "cflag"); card_int(h, "extra", a.extra, "extra");
      card(h, "END");
      pad2880(h, ' ');
      f.put(h.data(), h.size());
      // big-endian float32 data unit, streamed in chunks, zero padded to a multiple of 2880 bytes
      for_float_chunks(ivals + (size_t)m * npix * nv, npix * nv, [&](const float *v, size_t n) {
        std::string d;
        put_be_floats(d, v, n);
        f.put(d.data(), d.size());
      });
      const size_t rem = (npix * nv * 4) % 2880;
      if (rem) {
        const std::string z(2880 - rem, '\0');
        f.put(z.data(), z.size());
      }
    }
  } else {
    // Fortran sequential unformatted, 4-byte record markers (camera.f90:322-341)
    for (int m = 0; m < ncams; m++) {
      int dims[3] = {nx, ny, nv};
      frec(f, dims, 12);
      int nkey = 3;
      frec(f, &nkey, 4);
      float kv[3] = {(float)nx, (float)ny, (float)freqs[m % a.nfreq]};
      for (int k = 0; k < 3; k++) frec(f, &kv[k], 4);
      frec(f, pixloc, 8 * npix);
      std::vector<float> pv(npix * nv);
      for (size_t i = 0; i < pv.size(); i++) pv[i] = (float)ivals[(size_t)m * npix * nv + i];
      frec(f, pv.data(), 4 * npix * nv);
    }
  }
  if (fclose(f.fp) != 0) f.ok = false;
  return f.ok ? GRT_OK : GRT_ERR_IO;
}

}  // namespace grt_host

extern "C" void grt_args_defaults(grt_args *a) {
  memset(a, 0, sizeof(*a));
  a->standard = 2; a->mumin = .1; a->mumax = 1.; a->nmu = 10; a->phi0 = -0.5; a->spin = 0.998; a->uout = 0.04;
  a->uin = 1.; a->rcut = 1.; a->nrotype = 2;
  a->gridvals[0] = -25.; a->gridvals[1] = 25.; a->gridvals[2] = -25.; a->gridvals[3] = 25.;
  a->nn[0] = 100; a->nn[1] = 100; a->nn[2] = 1;
  a->i1 = 1; a->i2 = 10000;
  snprintf(a->fname, GRT_STRLEN, "THINDISK");
  a->dt = 5.; a->nt = 1; a->nload = 1; a->nmdot = 1; a->mdotmin = 1.5e15; a->mdotmax = 1.5e15;
  snprintf(a->ename, GRT_STRLEN, "BB");
  a->mbh = 10.; a->nfreq = 1; a->fmin = 1.e17; a->fmax = 3.e19; a->muval = 0.25; a->gmin = 100; a->gmax = 1e5;
  a->p1 = 3.5; a->p2 = 3.5; a->jetalpha = 0.02;
  snprintf(a->stype, GRT_STRLEN, "const");
  a->use_geokerr = 1; a->nvals = 1;
  snprintf(a->iname, GRT_STRLEN, "lsoda");
  a->cflag = 1;
  snprintf(a->fdfile, GRT_PATHLEN, "fieldline"); snprintf(a->fgfile, GRT_PATHLEN, "dump0000rr2.bin");
  snprintf(a->fhfile, GRT_PATHLEN, "dump040"); snprintf(a->fsim, GRT_STRLEN, "thickdiskrr2");
  a->fnt = 1; a->findf = 5206; a->fnfiles = 1; a->fjonfix = 1; a->fnw = 500; a->fnfreq_tab = 100; a->fnr = 500;
  a->fdindf = 1; a->frspot = 1.5; a->fr0spot = 6.0; a->fn0spot = 1e4; a->ftscl = 1.; a->frscl = 6.; a->fwmin = 1e-4;
  a->fwmax = 1e4; a->ffmin = 1.e17 / 3.; a->ffmax = 9e19; a->frmax = 1e4; a->fsigt = 0.4; a->ffcol = 1.7; a->fmdot = 0.1;
  a->fnscl = 3e7; a->fnnthscl = 8e4; a->fnnthp = 2.9; a->fbeta = 10.; a->frin = -1.; a->frout = 1e8; a->fthin = -10.;
  a->fthout = 10.; a->fphiout = 1e4; a->fscalefac = 1.; a->sigcut = 1e10;
  for (int i = 0; i < 7; i++) a->epcoefindx[i] = 1;
  a->nepotherargs = 2; a->epotherargs[0] = 1.; a->epotherargs[1] = 1.;
}

// ------------------------------------------------------------------------------------------ geodebug.out
namespace grt_host {
// One list-directed WRITE of the reference per line (gfortran does not wrap list-directed records written to a file), in
// the order of grtrans_driver.f90:93-110 (geodesic) and :342-418 (fluid, emission, intensities, tetrad); the reader
// (read_geodebug_file.py) splits every line on white space.  Row indices of `rows` are those of grt_debug.cuh.
int write_geodebug(const char *path, const GeoDebugData &d) {
  FILE *f = fopen(path, "w");
  if (!f) return GRT_ERR_IO;
  const int n = d.npts, mp = d.maxpts;
  auto dline = [&](const double *v, int stride = 1) {
    for (int i = 0; i < n; i++) fprintf(f, " %24.16E", v[(size_t)i * stride]);
    fputc('\n', f);
  };
  auto fline = [&](const float *v) {
    for (int i = 0; i < n; i++) fprintf(f, " %15.8E", (double)v[i]);
    fputc('\n', f);
  };
  auto iline = [&](const double *v) {
    for (int i = 0; i < n; i++) fprintf(f, " %11d", (int)v[i]);
    fputc('\n', f);
  };
  auto row = [&](int r) { return d.rows.data() + (size_t)r * mp; };
  // header: npts, nitems, alpha, beta, su, sm, mdot(:), mbh(:), fac, a, mu0   ((sic) fac is whatever the variable held
  // before this pixel's frequency loop in the reference; 0 here)
  fprintf(f, " %11d %11d %24.16E %24.16E %24.16E %24.16E", n, 18, d.alpha, d.beta, d.su, d.sm);
  for (double m : d.mdots) fprintf(f, " %24.16E", m);
  for (double m : d.mbh) fprintf(f, " %24.16E", m);
  fprintf(f, " %24.16E %24.16E %24.16E\n", 0.0, d.a, d.mu0);
  dline(d.lambda.data()); dline(d.t.data()); dline(d.r.data()); dline(d.th.data()); dline(d.ph.data());
  dline(d.kt.data()); dline(d.kr.data()); dline(d.kth.data()); dline(d.kph.data());
  iline(row(35)); iline(row(36));                                     // tprarr, tpmarr
  if (n > 0) {
    fline(d.rho.data()); fline(d.p.data()); fline(d.bmag.data());
    for (int c = 0; c < 4; c++) dline(d.u.data() + c, 4);
    for (int c = 0; c < 4; c++) dline(d.b.data() + c, 4);
    dline(d.ncgs.data()); dline(d.tcgs.data()); dline(d.bcgs.data()); dline(d.rshift.data()); dline(d.ang.data());
    dline(row(0)); dline(row(4));                                     // j_I, alpha_I
    dline(row(1)); dline(row(3)); dline(row(5)); dline(row(7));       // j_Q, j_V, alpha_Q, alpha_V
    dline(row(8)); dline(row(10));                                    // rho_Q, rho_V
    for (int c = 0; c < 4; c++) dline(row(12 + c));                   // I, Q, U, V / fac / LBH along the ray
    dline(row(11));                                                   // tau
    dline(row(2)); dline(row(6)); dline(row(9));                      // j_U, alpha_U, rho_U
    dline(d.s2xi.data()); dline(d.c2xi.data());
    dline(row(16)); dline(row(17));                                   // calc_kb_ang, calcg
    dline(row(33)); dline(row(34));                                   // s2xi, c2xi of comoving_ortho_debug
    for (int c = 0; c < 3; c++)                                       // aahat(npts, 3), column-major, in ONE record
      for (int i = 0; i < n; i++) fprintf(f, " %24.16E", row(18 + c)[i]);
    fputc('\n', f);
    for (int c = 0; c < 4; c++) dline(row(21 + c));                   // aat, aar, aath, aaph
    for (int c = 0; c < 4; c++) dline(row(25 + c));                   // kht, khr, khth, khph
    for (int c = 0; c < 4; c++) dline(row(29 + c));                   // bht, bhr, bhth, bhph
    fprintf(f, " %24.16E\n", d.fac);
  }
  const bool bad = ferror(f) != 0;
  if (fclose(f) != 0 || bad) return GRT_ERR_IO;
  return GRT_OK;
}
}  // namespace grt_host
