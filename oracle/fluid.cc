// ORACLE -- TEST INFRASTRUCTURE ONLY (see geokerr.cc header).
// CPU restatement of the fluid sampling stage of the hot path:
//   fluid.f90 (get_*_fluidvars, convert_fluidvars_*, scale_sim_units, monika_e, nonthermale_b2,
//   andrew_sigcut), fluid_model_{thindisk,sphacc,ffjet,harm,harm3d}.f90, interpolate.f90, hunt.f,
//   math.f90:zbrent.  Single/double precision placement follows the reference declarations
//   (real vs real(kind=8)) and Fortran's mixed-mode promotion rules.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

#include "oracle.h"
#include "ref_tables.h"

namespace orc {

using pc::pi;

// ------------------------------------------------------------------ hunt.f (float)
// returns 1-based jlo; jlo may be 0 or n when x is out of range
static void hunt(const float *xx /*1-based*/, int n, float x, int &jlo) {
  bool ascnd = (xx[n] >= xx[1]);
  int jhi, inc;
  if (jlo <= 0 || jlo > n) {
    jlo = 0;
    jhi = n + 1;
  } else {
    inc = 1;
    if ((x >= xx[jlo]) == ascnd) {
      for (;;) {
        jhi = jlo + inc;
        if (jhi > n) {
          jhi = n + 1;
          break;
        } else {
          if ((x < xx[jhi]) == ascnd) break;
          jlo = jhi;
          inc = inc + inc;
        }
      }
    } else {
      jhi = jlo;
      for (;;) {
        jlo = jhi - inc;
        if (jlo < 1) {
          jlo = 0;
          break;
        } else {
          if ((x >= xx[jlo]) == ascnd) break;
          jhi = jlo;
          inc = inc + inc;
        }
      }
    }
  }
  for (;;) {
    if (jhi - jlo <= 1) {
      if (x == xx[n]) jlo = n - 1;
      if (x == xx[1]) jlo = 1;
      break;
    } else {
      int jm = (jhi + jlo) / 2;
      if ((x >= xx[jm]) == ascnd)
        jlo = jm;
      else
        jhi = jm;
    }
  }
}

// interpolate.f90:76-84
static void get_weight(const float *xx, int n, float x, int &jlo, float &weight) {
  hunt(xx, n, x, jlo);
  int j = std::min(std::max(jlo, 1), n - 1);  // guard (the reference reads out of bounds)
  jlo = j;
  weight = (x - xx[j]) / (xx[j + 1] - xx[j]);
}

// ------------------------------------------------------------------ math.f90:56-135
template <class F>
static double zbrent(F func, double x1, double x2, double tol) {
  const int itmax = 100;
  const double eps = 2.220446049250313e-16;
  double a = x1, b = x2, c, d = 0, e = 0, fa, fb, fc, p, q, r, s, tol1, xm;
  fa = func(a);
  fb = func(b);
  if ((fa > 0.0 && fb > 0.0) || (fa < 0.0 && fb < 0.0)) {
    // 'ROOT MUST BE BRACKETED FOR ZBRENT' (the reference STOPs)
    return std::nan("");
  }
  c = b;
  fc = fb;
  for (int iter = 1; iter <= itmax; iter++) {
    if ((fb > 0.0 && fc > 0.0) || (fb < 0.0 && fc < 0.0)) {
      c = a;
      fc = fa;
      d = b - a;
      e = d;
    }
    if (std::fabs(fc) < std::fabs(fb)) {
      a = b;
      b = c;
      c = a;
      fa = fb;
      fb = fc;
      fc = fa;
    }
    tol1 = 2.0 * eps * std::fabs(b) + 0.5 * tol;
    xm = 0.5 * (c - b);
    if (std::fabs(xm) <= tol1 || fb == 0.0) return b;
    if (std::fabs(e) >= tol1 && std::fabs(fa) > std::fabs(fb)) {
      s = fb / fa;
      if (a == c) {
        p = 2.0 * xm * s;
        q = 1.0 - s;
      } else {
        q = fa / fc;
        r = fb / fc;
        p = s * (2.0 * xm * q * (q - r) - (b - a) * (r - 1.0));
        q = (q - 1.0) * (r - 1.0) * (s - 1.0);
      }
      if (p > 0.0) q = -q;
      p = std::fabs(p);
      if (2.0 * p < std::min(3.0 * xm * q - std::fabs(tol1 * q), std::fabs(e * q))) {
        e = d;
        d = p / q;
      } else {
        d = xm;
        e = d;
      }
    } else {
      d = xm;
      e = d;
    }
    a = b;
    fa = fb;
    b = b + ((std::fabs(d) > tol1) ? d : (std::signbit(xm) ? -std::fabs(tol1) : std::fabs(tol1)));
    fb = func(b);
  }
  return b;
}

// ------------------------------------------------------------------ kerr.f90:147-162 bl2ks
static double bl2ks_time(double r, double x, double a) {
  return x + std::log(r * r - 2 * r + a * a) +
         1. / (2.0 * std::sqrt(1. - a * a)) * std::log((r - 1. - std::sqrt(1. - a * a)) / (r - 1 + std::sqrt(1. - a * a)));
}
static double bl2ks_phi(double r, double x, double a) {
  return x + a / (2.0 * std::sqrt(1. - a * a)) * std::log((r - 1. - std::sqrt(1. - a * a)) / (r - 1. + std::sqrt(1. - a * a)));
}

// Dump slot of one geodesic point (fluid_model_harm.f90:133-183; the same lines in the harm3d :147-234, iharm :258-327,
// koral :118-190 and koral3d :123-218 modules): tt = KS look-back time of the point relative to the ray's first point
// (real(4)); tindx = floor(tt/tstep) limited to nt-1.  `eps`: harm / koral add 1d-8 before flooring.  The time weight is
// forced to zero (rttd=0.), so the sample is slot tindx alone; slot tindx+1 only enters multiplied by zero.
static size_t time_slot_offset(const FluidData &fd, const Geo &g, int i, float a, bool eps) {
  if (fd.nt <= 1) return 0;  // (nt=1: tindx=maxtindx+1=0 whenever tt>=0)
  float zr = (float)g.x[i].d[1], zt = (float)g.x[i].d[0], zr1 = (float)g.x[0].d[1];
  float tt = (float)(bl2ks_time((double)zr, (double)zt, (double)a) - bl2ks_time((double)zr1, 0.0, (double)a));
  tt = eps ? (float)((double)(-tt) + 1e-8) : -tt;
  int maxtindx = fd.nt - 2;
  double q = (fd.model == F_IHARM) ? std::floor((double)tt / fd.tstep_d) : (double)std::floor(tt / fd.tstep);
  int tindx = (q <= (double)maxtindx) ? (int)q : maxtindx + 1;
  if (tindx < 0) tindx = 0;  // guard (the reference would index before the arrays)
  return fd.ncell * (size_t)tindx;
}

// ------------------------------------------------------------------ THINDISK
// kerr.f90:102-129 krolikc (real(4) variables, mixed-mode arithmetic as written)
static float krolikc(float r, float a) {
  float pif = std::acos(-1.f);
  float rms = calc_rms(a);
  float y = std::sqrt(r);
  float yms = std::sqrt(rms);
  float y1 = (float)(2.0 * (double)std::cos(1.f / 3.f * (std::acos(a) - pif)));
  float y2 = (float)(2.0 * (double)std::cos(1.f / 3.f * (std::acos(a) + pif)));
  float y3 = (float)(-2.0 * (double)std::cos(1.f / 3.f * std::acos(a)));
  float arg1 = (float)((double)(3.f * a) / (2.0 * (double)y));
  float arg2 = (float)((double)3.f * std::pow((double)(y1 - a), 2.0) / (double)(y * y1 * (y1 - y2) * (y1 - y3)));
  float arg3 = (float)((double)3.f * std::pow((double)(y2 - a), 2.0) / (double)(y * y2 * (y2 - y1) * (y2 - y3)));
  float arg4 = (float)((double)3.f * std::pow((double)(y3 - a), 2.0) / (double)(y * y3 * (y3 - y1) * (y3 - y2)));
  float kc = 1.f - yms / y - arg1 * std::log(y / yms) - arg2 * std::log((y - y1) / (yms - y1)) -
             arg3 * std::log((y - y2) / (yms - y2)) - arg4 * std::log((y - y3) / (yms - y3));
  return kc;
}

// fluid_model_thindisk.f90:51-84
static void thindisk_vals(const FluidData &fd, float &rin_state, float r, float th, float a, float &T, float &omega) {
  float MBH = fd.td_mbh, Mdot = fd.td_mdot;
  float lbh = (float)((double)MBH * pc::msun * pc::G / pc::c2);
  float b = 1.f - 3.f / r + 2.f * a / std::pow(r, 3.f / 2.f);
  float rms = calc_rms(a);
  rin_state = std::max(rms, rin_state);
  float kc = krolikc(r, a);
  float d = r * r - 2.f * r + a * a;
  float lc = (rms * rms - 2.f * a * std::sqrt(rms) + a * a) / (std::pow(rms, 1.5f) - 2.f * std::sqrt(rms) + a);
  float hc = (2.f * r - a * lc) / d;
  float ar = std::pow(r * r + a * a, 2.f) - a * a * d * std::pow(std::sin(th), 2.f);
  float om = 2.f * a * r / ar;
  // ledd(m)=4.*pi*G*m*msun*mp*c/sigt  (double result), Mdotedd=ledd(MBH)/c2 -> real(4)
  double leddval = 4. * pi * pc::G * (double)MBH * pc::msun * pc::mp * pc::c / pc::sigt;
  float Mdotedd = (float)(leddval / pc::c2);
  float T0 = (float)std::pow(3.0 / 8.0 / pi * pc::G * (double)MBH * pc::msun * (double)Mdot * (double)Mdotedd /
                                 (double)lbh / (double)lbh / (double)lbh / pc::sigb,
                             (double).25f);
  if (r > rms)
    omega = std::max(1.f / (std::pow(r, 3.f / 2.f) + a), om);
  else
    omega = std::max((lc + a * hc) / (r * r + 2.f * r * (1.f + hc)), om);
  if (r > rin_state && r < fd.td_rout)
    T = T0 * std::pow(kc / b / std::pow(r, 3.f), 1.f / 4.f);
  else
    T = (float)((double)T0 / 1e5);
}

// fluid.f90:586-620
static void get_thindisk_fluidvars(const FluidData &fd, const Geo &g, std::vector<FluidPt> &f) {
  float a = (float)g.a;
  float rin_state = fd.td_rin;
  for (int i = 0; i < g.npts; i++) {
    float T, omega;
    float r = (float)g.x[i].d[1], th = (float)g.x[i].d[2];
    thindisk_vals(fd, rin_state, r, th, a, T, omega);
    f[i].rho = T;
    float metric[10];
    blmetric_cov_real(r, th, a, metric);
    // f%u%data(1)=sqrt(-1./(metric(:,1)+2d0*metric(:,4)*omega+metric(:,10)*omega*omega))
    double den = (double)metric[0] + 2.0 * (double)metric[3] * (double)omega + (double)(metric[9] * omega * omega);
    f[i].u[0] = std::sqrt(-1. / den);
    f[i].u[1] = 0.;
    f[i].u[2] = 0.;
    f[i].u[3] = (double)omega * f[i].u[0];
    calc_polvec(g.x[i].d[1], std::cos(g.x[i].d[2]), g.k[i].d, (double)a, std::asin(1.0), f[i].b);
  }
}

// ------------------------------------------------------------------ SPHACC
static void init_sphacc(FluidData &fd) {
  fd.uvals.assign(499, 0.f);
  fd.vvals.assign(499, 0.f);
  fd.rvals.assign(462, 0.f);
  fd.tvals.assign(462, 0.f);
  for (int I = 1; I <= 498; I++) fd.uvals[I] = (float)(I - 1) * .001f + 1.f / 399.f;
  for (int I = 1; I <= 461; I++) {
    fd.tvals[I] = (float)GRT_SPHACC_TVALS[I - 1];
    fd.rvals[I] = (float)GRT_SPHACC_RVALS[I - 1];
  }
  for (int I = 1; I <= 498; I++) fd.vvals[I] = (float)GRT_SPHACC_VVALS[I - 1];
}

// fluid_model_sphacc.f90:450-497 + fluid.f90:1215-1247
static void get_sphacc_fluidvars(const FluidData &fd, const Geo &g, std::vector<FluidPt> &f) {
  float gamma = (float)(5.0 / 3.0), alpha = (float)(1.0 / 4.0), ninf = 1.f, Tinf = (float).917e-9, us = (float).94;
  static thread_local std::vector<float> uv;
  if (uv.size() != 462) {
    uv.assign(462, 0.f);
    for (int i = 1; i <= 461; i++) uv[i] = (float)(1.0 / (double)fd.rvals[i]);
  }
  int lindx_v = 1, lindx_t = 1;
  for (int i = 0; i < g.npts; i++) {
    float u = (float)(1. / g.x[i].d[1]);
    float n = (float)((double)(ninf * alpha / 4.f / us) * std::pow((double)(2 * gamma * Tinf), -1.5) *
                      std::pow(2.0 * (double)u, 1.5));
    float w;
    get_weight(uv.data(), 461, u, lindx_t, w);
    float T = (float)((1.0 - (double)w) * (double)fd.tvals[lindx_t] + (double)(w * fd.tvals[lindx_t + 1]));
    float B = (float)std::sqrt(8. * pi * (double)n * pc::mp / 2. * pc::c2 * (double)u);
    get_weight(fd.uvals.data(), 498, u, lindx_v, w);
    float ur = (float)((1.0 - (double)w) * (double)fd.vvals[lindx_v] + (double)(w * fd.vvals[lindx_v + 1]));
    float g00 = -(1.f - 2.f * u);
    float grr = (float)(-1.0 / (double)g00);
    FluidPt &p = f[i];
    p.u[1] = -(double)ur;
    p.u[2] = 0.0;
    p.u[3] = 0.0;
    p.u[0] = std::sqrt((-(double)grr * p.u[1] * p.u[1] - 1) / (double)g00);
    p.b[2] = 0.0;
    p.b[3] = 0.0;
    double B2 = (double)(B * B);
    p.b[0] = std::sqrt(p.u[1] * p.u[1] * (double)grr * B2 /
                       (p.u[1] * p.u[1] * (double)g00 * (double)grr + p.u[0] * p.u[0] * (double)g00 * (double)g00));
    p.b[1] = -std::sqrt((double)((B * B) / grr) - p.b[0] * p.b[0] * (double)g00 / (double)grr);
    p.rho = n;
    p.p = T;
    p.bmag = B;
  }
}

// ------------------------------------------------------------------ gridded models: shared tail
// b.b with float32 metric, LNRF^-1 velocities, u^i = u^t v^i  (fluid_model_ffjet.f90:133-162 etc.)
static void finish_grid_point(FluidPt &p, float zr, float th, float a, float vrl0, float vtl0, float vpl0,
                              double bfac) {
  float metric[10];
  blmetric_cov_real(zr, th, a, metric);
  double m[10];
  for (int i = 0; i < 10; i++) m[i] = (double)metric[i];
  float bmag = (float)dot4(m, p.b, p.b);
  bmag = ((double)bmag >= 0.0) ? bmag : 0.f;
  bmag = std::sqrt(bmag);
  p.bmag = (float)((double)bmag * bfac);
  float vr0, vth0, vph0;
  lnrf_frame_inv_real(vrl0, vtl0, vpl0, zr, a, th, vr0, vth0, vph0);
  p.u[1] = p.u[0] * (double)vr0;
  p.u[2] = p.u[0] * (double)vth0;
  p.u[3] = p.u[0] * (double)vph0;
}

static inline float bilin(const float *v, float xd, float yd) {  // interpolate.f90:171-182
  float w1 = v[0] * (1.f - yd) + v[1] * yd;
  float w2 = v[2] * (1.f - yd) + v[3] * yd;
  return w1 * (1.f - xd) + w2 * xd;
}
static inline float trilin(const float *v, float xd, float yd, float zd) {  // interpolate.f90:184-201
  float i1 = v[0] * (1.f - zd) + v[1] * zd;
  float i2 = v[2] * (1.f - zd) + v[3] * zd;
  float j1 = v[4] * (1.f - zd) + v[5] * zd;
  float j2 = v[6] * (1.f - zd) + v[7] * zd;
  float w1 = i1 * (1.f - yd) + i2 * yd;
  float w2 = j1 * (1.f - yd) + j2 * yd;
  return w1 * (1.f - xd) + w2 * xd;
}
static inline float quadlin(const float *v, float td, float xd, float yd, float zd) {  // interpolate.f90:203-232
  float i1 = v[0] * (1.f - zd) + v[1] * zd, i2 = v[2] * (1.f - zd) + v[3] * zd;
  float j1 = v[4] * (1.f - zd) + v[5] * zd, j2 = v[6] * (1.f - zd) + v[7] * zd;
  float k1 = v[8] * (1.f - zd) + v[9] * zd, k2 = v[10] * (1.f - zd) + v[11] * zd;
  float l1 = v[12] * (1.f - zd) + v[13] * zd, l2 = v[14] * (1.f - zd) + v[15] * zd;
  float m1 = i1 * (1.f - yd) + i2 * yd, m2 = j1 * (1.f - yd) + j2 * yd;
  float n1 = k1 * (1.f - yd) + k2 * yd, n2 = l1 * (1.f - yd) + l2 * yd;
  float w1 = m1 * (1.f - xd) + m2 * xd, w2 = n1 * (1.f - xd) + n2 * xd;
  return w1 * (1.f - td) + w2 * td;
}

// ------------------------------------------------------------------ FFJET
// fluid_model_ffjet.f90:181-210 (Fortran sequential unformatted, 4-byte record markers)
static bool load_ffjet(FluidData &fd, const std::string &dfile) {
  FILE *fp = fopen(dfile.c_str(), "rb");
  if (!fp) return false;
  auto rec = [&](std::vector<char> &buf) -> bool {
    int n1 = 0, n2 = 0;
    if (fread(&n1, 4, 1, fp) != 1) return false;
    buf.resize(n1);
    if (n1 && fread(buf.data(), 1, n1, fp) != (size_t)n1) return false;
    if (fread(&n2, 4, 1, fp) != 1) return false;
    return n1 == n2;
  };
  std::vector<char> b;
  if (!rec(b)) { fclose(fp); return false; }
  float aa;
  int nx;
  memcpy(&aa, b.data(), 4);
  memcpy(&nx, b.data() + 4, 4);
  fd.asim = aa;
  int n1d = (int)std::sqrt((float)nx);
  fd.nx1 = n1d;
  fd.nx2 = n1d;
  auto take = [&](const char *&p, std::vector<float> &v) {
    v.resize(nx);
    memcpy(v.data(), p, 4 * (size_t)nx);
    p += 4 * (size_t)nx;
  };
  std::vector<float> dummy;
  const char *p;
  if (!rec(b)) { fclose(fp); return false; }
  p = b.data();
  take(p, fd.r_arr); take(p, fd.th_arr); take(p, fd.rho_arr);
  if (!rec(b)) { fclose(fp); return false; }
  p = b.data();
  take(p, dummy); take(p, fd.b0_arr); take(p, fd.br_arr); take(p, fd.bth_arr); take(p, fd.bph_arr);
  if (!rec(b)) { fclose(fp); return false; }
  p = b.data();
  take(p, fd.u0_arr); take(p, fd.vrl_arr); take(p, fd.vtl_arr); take(p, fd.vpl_arr);
  fclose(fp);
  fd.p_arr.assign(nx, 0.f);  // allocated but never read by the reference (garbage there)
  return true;
}

// fluid_model_ffjet.f90:42-171
static void ffjet_vals(const FluidData &fd, const Geo &g, std::vector<FluidPt> &f) {
  float a = (float)g.a;
  int nx1 = fd.nx1, nx2 = nx1;
  std::vector<float> uniqr(nx1 + 1), uniqx1(nx1 + 1), uniqx2(nx1 + 1);
  for (int i = 1; i <= nx1; i++) {
    uniqr[i] = fd.r_arr[i - 1];
    uniqx2[i] = fd.th_arr[(i - 1) * nx1];
    uniqx1[i] = std::log(uniqr[i]);
  }
  const float *uniqth = uniqx2.data();
  for (int i = 0; i < g.npts; i++) {
    float theta = (float)g.x[i].d[2];
    float zm = std::cos(theta);
    float x2 = std::acos(std::fabs(zm));
    theta = x2;
    float zr = (float)g.x[i].d[1];
    float x1 = std::log(zr);
    int lx1 = (int)((x1 - uniqx1[1]) / (uniqx1[nx1] - uniqx1[1]) * (float)(nx1 - 1)) + 1;
    int ux1 = lx1 + 1;
    int lx2 = (int)((x2 - uniqx2[1]) / (uniqx2[nx2] - uniqx2[1]) * (float)(nx2 - 1)) + 1;
    lx2 = (lx2 <= (nx2 - 1)) ? ((lx2 >= 1) ? lx2 : 1) : (nx2 - 1);
    int ux2 = lx2 + 1;
    // guard the radial indices (the reference reads out of bounds and discards via merge)
    int lx1s = std::min(std::max(lx1, 1), nx1 - 1), ux1s = lx1s + 1;
    float td = (theta - uniqth[lx2]) / (uniqth[ux2] - uniqth[lx2]);
    float rd = (zr - uniqr[lx1s]) / (uniqr[ux1s] - uniqr[lx1s]);
    (void)ux1;
    int x1l = lx1s - 1, x1u = ux1s - 1, x2l = (lx2 - 1) * nx1, x2u = (ux2 - 1) * nx1;
    int idx[4] = {x1l + x2l, x1l + x2u, x1u + x2l, x1u + x2u};  // 0-based
    auto gat = [&](const std::vector<float> &arr) {
      float v[4] = {arr[idx[0]], arr[idx[1]], arr[idx[2]], arr[idx[3]]};
      return bilin(v, rd, td);
    };
    bool in = x1 > uniqx1[1];
    FluidPt &p = f[i];
    p.rho = in ? gat(fd.rho_arr) : 0.f;
    p.p = in ? gat(fd.p_arr) : 1.f;
    float vrl0 = in ? gat(fd.vrl_arr) : 0.f;
    float vtl0 = in ? gat(fd.vtl_arr) : 1.f;
    p.b[0] = in ? gat(fd.b0_arr) : 1.f;
    p.b[1] = in ? gat(fd.br_arr) : 1.f;
    p.b[2] = in ? gat(fd.bth_arr) : 1.f;
    p.b[3] = in ? gat(fd.bph_arr) : 1.f;
    p.u[0] = in ? (double)gat(fd.u0_arr) : 1.0;
    float vpl0 = in ? gat(fd.vpl_arr) : 0.f;
    finish_grid_point(p, zr, (float)g.x[i].d[2], a, vrl0, vtl0, vpl0, 1.0);
  }
}

// ------------------------------------------------------------------ HARM (2-D)
// fluid_model_harm.f90:100-265
static void harm_vals(const FluidData &fd, const Geo &g, std::vector<FluidPt> &f) {
  float a = (float)g.a;
  int nx1 = fd.nx1, nx2 = fd.nx2;
  float h = fd.h;
  std::vector<float> uniqx1(nx1 + 1), uniqr(nx1 + 1), uniqx2(nx2 + 1), uniqth(nx2 + 1);
  for (int i = 1; i <= nx2; i++) {
    uniqx2[i] = fd.x2_arr[i - 1];
    uniqth[i] = (float)(pi * (double)uniqx2[i] + (double)((1.f - h) / 2.f) * std::sin(2. * pi * (double)uniqx2[i]));
  }
  for (int i = 1; i <= nx1; i++) {
    uniqx1[i] = fd.x1_arr[(i - 1) * nx2];
    uniqr[i] = std::exp(uniqx1[i]);
  }
  float rh = 1.f + std::sqrt(1.f - std::pow(a, 2.f));
  for (int i = 0; i < g.npts; i++) {
    float theta = (float)g.x[i].d[2];
    float zr = (float)g.x[i].d[1];
    float x1 = std::log(zr);
    double thd = (double)theta, hd = (double)h;
    float x2 = (float)zbrent([&](double x) { return thd - pi * x - .5 * (1 - hd) * std::sin(2. * pi * x); }, 0.0, 1.0, 1e-6);
    int lx1 = (int)std::floor((x1 - uniqx1[1]) / (uniqx1[nx1] - uniqx1[1]) * (float)(nx1 - 1)) + 1;
    int ux1 = lx1 + 1;
    int lx2 = (int)std::floor((x2 - uniqx2[1]) / (uniqx2[nx2] - uniqx2[1]) * (float)(nx2 - 1)) + 1;
    int ux2 = lx2 + 1;
    lx2 = (lx2 >= 1) ? lx2 : 1;
    ux2 = (ux2 <= nx2) ? ux2 : nx2;
    float dth = (ux2 != lx2) ? (uniqth[ux2] - uniqth[lx2]) : uniqth[1];
    float td = std::fabs(theta - uniqth[lx2]) / dth;
    int lx1s = std::min(std::max(lx1, 1), nx1 - 1), ux1s = lx1s + 1;  // guard
    (void)ux1;
    float rd = (zr - uniqr[lx1s]) / (uniqr[ux1s] - uniqr[lx1s]);
    if (uniqr[lx1s] <= rh) rd = 1.f;
    int x2l = lx2 - 1, x2u = ux2 - 1, x1l = (lx1s - 1) * nx2, x1u = (ux1s - 1) * nx2;
    // rttd=0: the "second time slice" enters with weight zero (and wraps onto the last slot, indx-n)
    size_t ts = time_slot_offset(fd, g, i, a, true);
    size_t idx[8] = {ts + x1l + x2l, ts + x1l + x2u, ts + x1u + x2l, ts + x1u + x2u, ts + x1l + x2l, ts + x1l + x2u,
                     ts + x1u + x2l, ts + x1u + x2u};
    auto gat = [&](const std::vector<float> &arr) {
      float v[8];
      for (int q = 0; q < 8; q++) v[q] = arr[idx[q]];
      return trilin(v, 0.f, rd, td);
    };
    bool in = x1 > uniqx1[1];
    FluidPt &p = f[i];
    p.rho = in ? gat(fd.rho_arr) : 0.f;
    p.p = in ? gat(fd.p_arr) : 1.f;
    float vrl0 = in ? gat(fd.vrl_arr) : 0.f;
    float vtl0 = in ? gat(fd.vtl_arr) : 1.f;
    p.b[0] = in ? gat(fd.b0_arr) : 1.f;
    p.b[1] = in ? gat(fd.br_arr) : 1.f;
    p.b[2] = in ? gat(fd.bth_arr) : 1.f;
    p.b[3] = in ? gat(fd.bph_arr) : 1.f;
    p.u[0] = in ? (double)gat(fd.u0_arr) : 1.0;
    float vpl0 = in ? gat(fd.vpl_arr) : 0.f;
    finish_grid_point(p, zr, (float)g.x[i].d[2], a, vrl0, vtl0, vpl0, 1.0);
  }
}

// ------------------------------------------------------------------ HARM3D
// fluid_model_harm3d.f90:107-327
static void harm3d_vals(const FluidData &fd, const Geo &g, std::vector<FluidPt> &f) {
  float a = (float)g.a;
  int nx1 = fd.nx1, nx2 = fd.nx2, nx3 = fd.nx3;
  std::vector<float> uniqx1(nx1 + 1), uniqr(nx1 + 1), uniqx2(nx2 + 1), uniqth(nx2 + 1), uniqx3(nx3 + 1);
  for (int i = 1; i <= nx3; i++) uniqx3[i] = fd.x3_arr[i - 1];
  for (int i = 1; i <= nx2; i++) {
    uniqx2[i] = fd.x2_arr[(i - 1) * nx3];
    uniqth[i] = (float)(pi * (double)uniqx2[i]);
  }
  for (int i = 1; i <= nx1; i++) {
    uniqx1[i] = fd.x1_arr[(size_t)(i - 1) * nx3 * nx2];
    uniqr[i] = std::exp(uniqx1[i]);
  }
  const float *uniqph = uniqx3.data();
  float rh = 1.f + std::sqrt(1.f - std::pow(a, 2.f));
  for (int i = 0; i < g.npts; i++) {
    float theta = (float)g.x[i].d[2];
    float zr = (float)g.x[i].d[1];
    float zpp = (float)g.x[i].d[3];
    float zphi = (float)bl2ks_phi((double)zr, (double)zpp, (double)a);
    zphi = (float)std::fmod((double)zphi, 2. * pi);
    if (zphi < 0.f) zphi = (float)((double)zphi + 2. * pi);
    float x1 = std::log(zr);
    float x3 = zphi;
    float x2 = (float)((double)theta / pi);
    int lx1 = (int)std::floor((x1 - uniqx1[1]) / (uniqx1[nx1] - uniqx1[1]) * (float)(nx1 - 1)) + 1;
    int lx2 = (int)std::floor((x2 - uniqx2[1]) / (uniqx2[nx2] - uniqx2[1]) * (float)(nx2 - 1)) + 1;
    int ux2 = lx2 + 1;
    int lx3 = (int)std::floor((x3 - uniqx3[1]) / (uniqx3[nx3] - uniqx3[1]) * (float)(nx3 - 1)) + 1;
    int ux3 = lx3 + 1;
    lx2 = (lx2 <= (nx2 - 1)) ? ((lx2 >= 1) ? lx2 : 1) : (nx2 - 1);
    lx1 = (lx1 >= 1) ? lx1 : 1;
    int ux1 = lx1 + 1;
    lx2 = (lx2 >= 1) ? lx2 : 1;
    lx2 = (lx2 <= nx2) ? lx2 : nx2;
    ux2 = (ux2 >= 1) ? ux2 : 1;
    ux2 = (ux2 <= nx2) ? ux2 : nx2;
    float dth = (ux2 != lx2) ? (uniqth[ux2] - uniqth[lx2]) : (uniqth[1] * 2);
    // periodic in phi
    int lx3s = std::min(std::max(lx3, 1), nx3);  // guard for the read below when lx3<1
    float minph = uniqph[lx3s];
    if (ux3 > nx3) ux3 = 1;
    if (lx3 < 1) {
      lx3 = nx3;
      minph = (float)((double)uniqph[lx3] - 2. * pi);
    }
    float pd = (zphi - minph) / (uniqph[2] - uniqph[1]);
    float td = std::fabs(theta - uniqth[lx2]) / dth;
    int ux1s = std::min(ux1, nx1);  // guard (reference reads out of bounds beyond the outer edge)
    int lx1s = std::min(lx1, nx1);
    float rd = (zr - uniqr[lx1s]) / (uniqr[ux1s] - uniqr[lx1s]);
    double pfac, nfac, bfac;
    if (uniqr[lx1s] <= std::max(uniqr[1], rh)) {
      rd = 1.f;
      pfac = 1e-6; nfac = 1e-6; bfac = 1e-6;
    } else {
      pfac = 1.0; nfac = 1.0; bfac = 1.0;
    }
    size_t x3l = lx3 - 1, x3u = ux3 - 1, x2l = (size_t)(lx2 - 1) * nx3, x2u = (size_t)(ux2 - 1) * nx3;
    size_t x1l = (size_t)(lx1s - 1) * nx2 * nx3, x1u = (size_t)(ux1s - 1) * nx2 * nx3;
    size_t idx[16] = {x1l + x2l + x3l, x1l + x2u + x3l, x1u + x2l + x3l, x1u + x2u + x3l,
                      x1l + x2l + x3u, x1l + x2u + x3u, x1u + x2l + x3u, x1u + x2u + x3u};
    {
      size_t ts = time_slot_offset(fd, g, i, a, false);
      for (int q = 0; q < 8; q++) idx[q] += ts;
    }
    for (int q = 0; q < 8; q++) idx[8 + q] = idx[q];  // second slice: weight zero (rttd=0.)
    auto gat = [&](const std::vector<float> &arr) {
      float v[16];
      for (int q = 0; q < 16; q++) v[q] = arr[idx[q]];
      return quadlin(v, 0.f, pd, rd, td);
    };
    bool in = x1 > uniqx1[1];
    FluidPt &p = f[i];
    p.rho = (float)((double)gat(fd.rho_arr) * nfac);
    p.p = (float)((double)gat(fd.p_arr) * pfac);
    float vrl0 = in ? gat(fd.vrl_arr) : 0.f;
    float vtl0 = in ? gat(fd.vtl_arr) : 0.f;
    p.b[0] = in ? gat(fd.b0_arr) : 1.f;
    p.b[1] = in ? gat(fd.br_arr) : 1.f;
    p.b[2] = in ? gat(fd.bth_arr) : 1.f;
    p.b[3] = in ? gat(fd.bph_arr) : 1.f;
    p.u[0] = in ? (double)gat(fd.u0_arr) : 1.0;
    float vpl0 = in ? gat(fd.vpl_arr) : 0.f;
    finish_grid_point(p, zr, (float)g.x[i].d[2], a, vrl0, vtl0, vpl0, bfac);
  }
}


// ------------------------------------------------------------------ IHARM
// fluid_model_iharm.f90.  theta(x2) of the MKS grid (calcthmksh :82-87, real(8)) and theta(x2,x1) of the MMKS/FMKS grid
// (calcthmmks :107-118, a real(4) function whose sub-expressions are promoted by the real(8) module variables pi, hslope,
// startx1 exactly as written).
static double calcthmksh(const FluidData &fd, double x2) {
  return pi * x2 + (1. - fd.hslope) / 2. * std::sin(2. * pi * x2);
}
static float calcthmmks(const FluidData &fd, float x2, float x1) {
  float A = (float)fd.mks_smooth, B = (float)fd.poly_xt, C = (float)fd.poly_alpha;
  float D = (float)(pi / (double)(2.f + 2.f / (std::pow(B, C) * (1 + C))));
  float thetag = (float)(pi * (double)x2 + (1. - fd.hslope) / 2. * std::sin(2. * pi * (double)x2));
  float thetaj = (float)((double)(D * (2.f * x2 - 1.f) * (1.f + std::pow((2.f * x2 - 1.f) / B, C) / (1.f + C))) + pi / 2.);
  return (float)((double)thetag + std::exp(-(double)A * ((double)x1 - fd.startx1)) * (double)(thetaj - thetag));
}

// iharm_vals :222-424: nearest-neighbour lookup (rd = td = pd = 1 selects the upper corner of the cell in all three
// directions, time weight 0).  nt = 1 only (tindx = 0).
static void iharm_vals(const FluidData &fd, const Geo &g, std::vector<FluidPt> &f) {
  float a = (float)g.a;
  int nx1 = fd.nx1, nx2 = fd.nx2, nx3 = fd.nx3;
  std::vector<float> uniqx1(nx1 + 1), uniqr(nx1 + 1), uniqx2(nx2 + 1), uniqx3(nx3 + 1);
  for (int i = 1; i <= nx3; i++) uniqx3[i] = fd.x3_arr[i - 1];
  for (int i = 1; i <= nx2; i++) uniqx2[i] = fd.x2_arr[(size_t)(i - 1) * nx3];
  for (int i = 1; i <= nx1; i++) {
    uniqx1[i] = fd.x1_arr[(size_t)(i - 1) * nx3 * nx2];
    uniqr[i] = std::exp(uniqx1[i]);
  }
  float rh = 1.f + std::sqrt(1.f - std::pow(a, 2.f));
  for (int i = 0; i < g.npts; i++) {
    float theta = (float)g.x[i].d[2];
    float zr = (float)g.x[i].d[1];
    float zpp = (float)g.x[i].d[3];
    float zphi = (float)bl2ks_phi((double)zr, (double)zpp, (double)a);
    zphi = (float)std::fmod((double)zphi, 2. * pi);
    if (zphi < 0.f) zphi = (float)((double)zphi + 2. * pi);
    float x1 = std::log(zr), x2, x3 = zphi;
    if (fd.is_mmks == 1) {  // transformbl2mmks :132-146
      double x1d = (double)x1, thd = (double)theta;
      x2 = (float)zbrent([&](double x) { return thd - (double)calcthmmks(fd, (float)x, (float)x1d); }, -1.0, 1.0, 1e-6);
    } else {  // transformbl2mksh :148-163
      double thd = (double)theta;
      x2 = (float)zbrent([&](double x) { return thd - calcthmksh(fd, x); }, 0.0, 1.0, 1e-6);
    }
    int lx1 = (int)std::floor((x1 - uniqx1[1]) / (uniqx1[nx1] - uniqx1[1]) * (float)(nx1 - 1)) + 1;
    int ux1 = lx1 + 1;
    int lx2 = (int)std::floor((x2 - uniqx2[1]) / (uniqx2[nx2] - uniqx2[1]) * (float)(nx2 - 1)) + 1;
    int ux2 = lx2 + 1;
    int lx3 = (int)std::floor((x3 - uniqx3[1]) / (uniqx3[nx3] - uniqx3[1]) * (float)(nx3 - 1)) + 1;
    int ux3 = lx3 + 1;
    lx2 = (lx2 >= 1) ? lx2 : 1;
    lx2 = (lx2 <= nx2) ? lx2 : nx2;
    ux2 = (ux2 >= 1) ? ux2 : 1;
    ux2 = (ux2 <= nx2) ? ux2 : nx2;
    if (ux3 > nx3) ux3 = 1;
    if (lx3 < 1) lx3 = nx3;
    // guards: the reference indexes uniqr(lx1) / the data arrays out of bounds outside the radial range
    int lx1s = std::min(std::max(lx1, 1), nx1), ux1s = std::min(std::max(ux1, 1), nx1);
    double pfac = 1.0, nfac = 1.0;
    if (uniqr[lx1s] <= rh) {
      pfac = 1e-3;
      nfac = 1e-3;
    }
    size_t idx = (size_t)(ux1s - 1) * nx2 * nx3 + (size_t)(ux2 - 1) * nx3 + (size_t)(ux3 - 1);
    idx += time_slot_offset(fd, g, i, a, false);
    bool in = x1 > uniqx1[1];
    FluidPt &p = f[i];
    p.rho = (float)((double)(in ? fd.rho_arr[idx] : 0.f) * nfac);
    p.p = (float)((double)(in ? fd.p_arr[idx] : 1.f) * pfac);
    if (fd.eheat == 1) p.kela = (float)((double)(in ? fd.kela_arr[idx] : 1.f) * pfac);
    float vrl0 = in ? fd.vrl_arr[idx] : 0.f;
    float vtl0 = in ? fd.vtl_arr[idx] : 0.f;
    p.b[0] = in ? fd.b0_arr[idx] : 1.f;
    p.b[1] = in ? fd.br_arr[idx] : 1.f;
    p.b[2] = in ? fd.bth_arr[idx] : 1.f;
    p.b[3] = in ? fd.bph_arr[idx] : 1.f;
    p.u[0] = in ? (double)fd.u0_arr[idx] : 1.0;
    float vpl0 = in ? fd.vpl_arr[idx] : 0.f;
    finish_grid_point(p, zr, (float)g.x[i].d[2], a, vrl0, vtl0, vpl0, 1.0);
  }
}

// read_iharm_data_header :427-470, read_iharm_data_file :508-651, load_iharm_data :684-747 (nt = 1)
static bool load_iharm(FluidData &fd, const FluidArgs &fa, double a) {
  std::ifstream hf(fa.hfile);
  if (!hf) return false;
  std::vector<double> header(200, -1e20);
  for (int i = 0; i < 200; i++)
    if (!(hf >> header[i])) break;
  fd.tcur = header[0];  // (the only place tcur is set: every dump slot gets the header file's time, :441,745)
  fd.nx1 = (int)header[1]; fd.nx2 = (int)header[2]; fd.nx3 = (int)header[3];
  fd.asim = (float)header[4];
  fd.hslope = header[5];
  fd.gam = (float)header[6];  // (gam is real(8) in the module; f%gamma is real)
  fd.mks_smooth = header[7]; fd.poly_xt = header[8]; fd.poly_alpha = header[9]; fd.startx1 = header[10];
  fd.is_mmks = (int)header[11];
  fd.eheat = (int)header[12];
  if (std::fabs(fd.asim - (float)a) > 1e-4f) {
    fprintf(stderr, "ERROR -- Different simulation and grtrans spin values!\n");
    return false;
  }
  char append[16];
  snprintf(append, sizeof append, "%03d", fa.indf);
  FILE *fp = fopen((fa.dfile + append + ".bin").c_str(), "rb");
  if (!fp) return false;
  size_t n = (size_t)fd.nx1 * fd.nx2 * fd.nx3;
  int dlen = 13 + fd.eheat;
  std::vector<float> data((size_t)dlen * n);
  bool ok = fread(data.data(), 4, data.size(), fp) == data.size();
  fclose(fp);
  if (!ok) return false;
  fd.x1_arr.resize(n); fd.x2_arr.resize(n); fd.x3_arr.resize(n); fd.r_arr.resize(n); fd.th_arr.resize(n); fd.ph_arr.resize(n);
  fd.rho_arr.resize(n); fd.p_arr.resize(n); fd.u0_arr.resize(n); fd.vrl_arr.resize(n); fd.vtl_arr.resize(n);
  fd.vpl_arr.resize(n); fd.b0_arr.resize(n); fd.br_arr.resize(n); fd.bth_arr.resize(n); fd.bph_arr.resize(n);
  if (fd.eheat == 1) fd.kela_arr.resize(n);
  for (size_t i = 0; i < n; i++) {
    const float *d = data.data() + (size_t)dlen * i;  // data(1:dlen, i)
    // grid=transpose(data(1:3,:)) is real(8); x?_arr are real
    float x1 = d[0], x2 = d[1], x3 = d[2];
    fd.x1_arr[i] = x1; fd.x2_arr[i] = x2; fd.x3_arr[i] = x3;
    float r = std::exp(x1);
    fd.r_arr[i] = r;
    fd.ph_arr[i] = x3;
    float th = (fd.is_mmks == 1) ? calcthmmks(fd, x2, x1) : (float)calcthmksh(fd, (double)x2);
    fd.th_arr[i] = th;
    double um[4] = {d[5], d[6], d[7], d[8]}, bm[4] = {d[9], d[10], d[11], d[12]};
    double uks[4], bks[4];
    if (fd.is_mmks == 1) {  // ummks2uks :165-184 (central differences of the real(4) theta function)
      double x1d = (double)x1, x2d = (double)x2;
      double dx2 = 1e-6 * x2d, dx1 = 1e-4 * x1d;
      double drdx1 = std::exp(x1d);
      double dthdx1 = (double)(calcthmmks(fd, (float)x2d, (float)(x1d + .5 * dx1)) - calcthmmks(fd, (float)x2d, (float)(x1d - .5 * dx1))) / dx1;
      double dthdx2 = (double)(calcthmmks(fd, (float)(x2d + .5 * dx2), (float)x1d) - calcthmmks(fd, (float)(x2d - .5 * dx2), (float)x1d)) / dx2;
      for (int c = 0; c < 2; c++) {
        const double *v = c ? bm : um;
        double *o = c ? bks : uks;
        o[1] = drdx1 * v[1];
        o[2] = dthdx1 * v[1] + dthdx2 * v[2];
        o[3] = v[3];
        o[0] = v[0];
      }
    } else {  // umksh2uks :186-220
      float dthdx2 = (float)(pi * (1. + (1. - fd.hslope) * std::cos(2 * pi * (double)x2)));
      for (int c = 0; c < 2; c++) {
        const double *v = c ? bm : um;
        double *o = c ? bks : uks;
        o[0] = v[0];
        o[1] = (double)r * v[1];
        o[2] = v[2] * (double)dthdx2;
        o[3] = v[3];
      }
    }
    // uks2ubl (kerr.f90:131-144)
    double rr = (double)r, asim = (double)fd.asim;
    double delta = rr * rr - 2.0 * rr + asim * asim;
    double u[4], b[4];
    for (int c = 0; c < 2; c++) {
      const double *v = c ? bks : uks;
      double *o = c ? b : u;
      o[0] = v[0] - 2.0 * rr / delta * v[1];
      o[1] = v[1];
      o[2] = v[2];
      o[3] = v[3] - asim / delta * v[1];
    }
    float vrl, vtl, vpl;
    lnrf_frame_real((float)(u[1] / u[0]), (float)(u[2] / u[0]), (float)(u[3] / u[0]), r, fd.asim, th, vrl, vtl, vpl);
    fd.rho_arr[i] = (float)(double)d[3];
    fd.p_arr[i] = (float)(double)d[4];  // (internal energy u: the (gamma-1) factor is applied in convert_fluidvars_iharm)
    fd.b0_arr[i] = (float)b[0]; fd.br_arr[i] = (float)b[1]; fd.bth_arr[i] = (float)b[2]; fd.bph_arr[i] = (float)b[3];
    fd.u0_arr[i] = (float)u[0];
    fd.vrl_arr[i] = vrl; fd.vtl_arr[i] = vtl; fd.vpl_arr[i] = vpl;
    if (fd.eheat == 1) fd.kela_arr[i] = (float)(double)d[13];
  }
  return true;
}


// ------------------------------------------------------------------ KORAL (2-D) and KORAL3D
// MKS3 coordinate maps (fluid_model_koral.f90:62-80; real(4) arguments, promoted by the real(8) pi as written)
static float koral_x2_of_theta(float th, float r, float h, float a, float b, float p) {
  float rp = std::pow(r, p);
  double num = std::atan(std::tan((double)(0.5f * h) * pi) * (1. - (double)(2.f * th) / pi));
  float den = (b - a) * std::pow(2.f, p) + (a - 0.5f) * rp;
  return (float)(0.5 * (1. + ((double)rp / ((double)h * pi)) * (num / (double)den)));
}
static float koral_theta_of_x2(float x2, float r, float h, float a, float b, float p) {
  float inner = -0.5f + x2 + (1.f - 2.f * x2) * (a + std::pow(2.f, p) * (b - a) / std::pow(r, p));
  return (float)(0.5 * pi * (1. + std::tan((double)h * pi * (double)inner) / std::tan((double)(0.5f * h) * pi)));
}

// koral_vals (fluid_model_koral.f90:84-286) / koral3d_vals (fluid_model_koral3d.f90:83-346); nt = 1
static void koral_vals(const FluidData &fd, const Geo &g, std::vector<FluidPt> &f) {
  const bool d3 = fd.model != F_KORAL;
  float a = (float)g.a;
  int nx1 = fd.nx1, nx2 = fd.nx2, nx3 = d3 ? fd.nx3 : 1;
  std::vector<float> uniqx1(nx1 + 1), uniqr(nx1 + 1), uniqx2(nx2 + 1), uniqx3(nx3 + 1);
  for (int i = 1; i <= nx2; i++) uniqx2[i] = fd.x2_arr[i - 1];
  for (int i = 1; i <= nx1; i++) {
    uniqx1[i] = fd.x1_arr[(size_t)(i - 1) * nx2];
    uniqr[i] = fd.kr0 + std::exp(uniqx1[i]);
  }
  if (d3)
    for (int i = 1; i <= nx3; i++) uniqx3[i] = fd.x3_arr[(size_t)(i - 1) * nx2 * nx1];
  float rh = 1.f + std::sqrt(1.f - std::pow(a, 2.f));
  float x2mintrust = uniqx2[fd.minpolecell], x2maxtrust = uniqx2[nx2 - (fd.minpolecell - 1)];
  float thmintrust = 0.f, thmaxtrust = (float)pi;
  if (fd.ktype == 1) {
    thmintrust = 0.25f;
    thmaxtrust = (float)(pi - 0.25);
  }
  for (int i = 0; i < g.npts; i++) {
    float theta = (float)g.x[i].d[2];
    float zr = (float)g.x[i].d[1];
    float x1 = (float)std::log((double)zr - (double)fd.kr0);
    float x2 = koral_x2_of_theta(theta, zr, fd.kh, fd.kaa, fd.kbb, fd.kpp);
    int lx1 = (int)std::floor((x1 - uniqx1[1]) / (uniqx1[nx1] - uniqx1[1]) * (float)(nx1 - 1)) + 1;
    int ux1 = lx1 + 1;
    lx1 = (lx1 >= 1) ? lx1 : 1;
    ux1 = (ux1 <= nx1) ? ux1 : nx1;
    lx1 = std::min(lx1, nx1);  // guard (the reference reads uniqr(lx1) out of bounds beyond the outer edge)
    ux1 = std::max(ux1, 1);
    int lx2 = (int)std::floor((x2 - uniqx2[1]) / (uniqx2[nx2] - uniqx2[1]) * (float)(nx2 - 1)) + 1;
    int ux2 = lx2 + 1;
    lx2 = (lx2 >= 1) ? lx2 : 1;
    ux2 = (ux2 <= nx2) ? ux2 : nx2;
    lx2 = std::min(lx2, nx2);  // guard
    ux2 = std::max(ux2, 1);
    float pd = 0.f;
    int lx3 = 1, ux3 = 1;
    if (d3) {
      float zphi = (float)g.x[i].d[3];
      zphi = (float)std::fmod((double)zphi, 2. * pi);
      if (zphi < 0.f) zphi = (float)((double)zphi + 2. * pi);
      if ((double)zphi > pi) zphi = (float)((double)zphi - 2. * pi);
      float x3 = zphi;
      lx3 = (int)std::floor((x3 - uniqx3[1]) / (uniqx3[nx3] - uniqx3[1]) * (float)(nx3 - 1)) + 1;
      ux3 = lx3 + 1;
      int lx3s = std::min(std::max(lx3, 1), nx3);
      float minph = uniqx3[lx3s];
      ux3 = (ux3 <= nx3) ? ux3 : 1;
      if (lx3 < 1) {
        lx3 = nx3;
        minph = (float)((double)uniqx3[lx3] - 2. * pi);
      }
      lx3 = std::min(lx3, nx3);  // guard
      pd = (zphi - minph) / (uniqx3[2] - uniqx3[1]);
    }
    float rd = (zr - uniqr[lx1]) / (uniqr[ux1] - uniqr[lx1]);
    double pfac = 1.0, nfac = 1.0;
    if (uniqr[lx1] <= rh || lx1 == 1) {
      rd = 1.f;
      if (d3) {
        pfac = 1e-3;
        nfac = 1e-3;
      }
    }
    float thl = koral_theta_of_x2(uniqx2[lx2], uniqr[lx1], fd.kh, fd.kaa, fd.kbb, fd.kpp);
    float thu = koral_theta_of_x2(uniqx2[ux2], uniqr[lx1], fd.kh, fd.kaa, fd.kbb, fd.kpp);
    float thone = koral_theta_of_x2(uniqx2[1], uniqr[lx1], fd.kh, fd.kaa, fd.kbb, fd.kpp);
    float td = (ux2 != lx2) ? std::fabs((theta - thl) / (thu - thl)) : std::fabs((theta - thl) / thone);
    // x2 fastest, then x1, then x3
    size_t x2l = lx2 - 1, x2u = ux2 - 1, x1l = (size_t)(lx1 - 1) * nx2, x1u = (size_t)(ux1 - 1) * nx2;
    size_t x3l = (size_t)(lx3 - 1) * nx2 * nx1, x3u = (size_t)(ux3 - 1) * nx2 * nx1;
    size_t idx[16] = {x1l + x2l + x3l, x1l + x2u + x3l, x1u + x2l + x3l, x1u + x2u + x3l,
                      x1l + x2l + x3u, x1l + x2u + x3u, x1u + x2l + x3u, x1u + x2u + x3u};
    {
      size_t ts = time_slot_offset(fd, g, i, a, true);
      for (int q = 0; q < 8; q++) idx[q] += ts;
    }
    for (int q = 0; q < 8; q++) idx[8 + q] = idx[q];  // second slice: weight zero (rttd=0.)
    auto gat = [&](const std::vector<float> &arr) {
      float v[16];
      if (d3) {
        for (int q = 0; q < 16; q++) v[q] = arr[idx[q]];
        return quadlin(v, 0.f, pd, rd, td);
      }
      for (int q = 0; q < 4; q++) v[q] = v[4 + q] = arr[idx[q]];
      return trilin(v, 0.f, rd, td);
    };
    bool in = x1 > uniqx1[1];
    bool trust = (x2 > x2mintrust) && (x2 < x2maxtrust);
    if (d3 && fd.ktype == 1) trust = trust && (theta > thmintrust) && (theta < thmaxtrust);
    bool ok = in && trust;
    FluidPt &p = f[i];
    // (the reference applies nfac/pfac between the two merges; a masked value is 0 (rho) or pfac (p) either way)
    p.rho = (float)((double)(in ? gat(fd.rho_arr) : 0.f) * nfac);
    p.p = (float)((double)(in ? gat(fd.p_arr) : 1.f) * pfac);
    if (!trust) {
      p.rho = 0.f;
      p.p = 1.f;
    }
    p.be = (d3 && ok) ? gat(fd.be_arr) : 0.f;
    float vrl0 = ok ? gat(fd.vrl_arr) : 0.f;
    float vtl0 = ok ? gat(fd.vtl_arr) : 1.f;
    p.b[0] = ok ? gat(fd.b0_arr) : 1.f;
    p.b[1] = ok ? gat(fd.br_arr) : 1.f;
    p.b[2] = ok ? gat(fd.bth_arr) : 1.f;
    p.b[3] = ok ? gat(fd.bph_arr) : 1.f;
    p.u[0] = ok ? (double)gat(fd.u0_arr) : 1.0;
    float vpl0 = ok ? gat(fd.vpl_arr) : 0.f;
    if (fd.doingnth) {  // fluid_model_koral.f90:262-270: masked below the inner edge and (sic) at the first pole only
      p.nnth.assign(fd.nrelbin, 0.f);
      for (int ie = 0; ie < fd.nrelbin; ie++) {
        float v[8];
        for (int q = 0; q < 4; q++) v[q] = v[4 + q] = fd.nth_arr[idx[q] * (size_t)fd.nrelbin + ie];
        float nt_ = trilin(v, 0.f, rd, td);
        nt_ = in ? nt_ : 0.f;
        p.nnth[ie] = (x2 > x2mintrust) ? nt_ : 0.f;
      }
    }
    finish_grid_point(p, zr, (float)g.x[i].d[2], a, vrl0, vtl0, vpl0, 1.0);
  }
}

// read_koral_data_header / read_koral_data_file / load_koral_data (fluid_model_koral.f90:296-488) and the koral3d
// "shortfile" variants (fluid_model_koral3d.f90:359-632): ASCII, list-directed
static bool load_koral(FluidData &fd, const FluidArgs &fa, double a) {
  const bool d3 = fd.model != F_KORAL;
  std::ifstream hf(fa.hfile);
  if (!hf) return false;
  int nhead = d3 ? 11 : 10;
  std::vector<double> header(nhead);
  for (int i = 0; i < nhead; i++)
    if (!(hf >> header[i])) return false;
  int o = d3 ? 1 : 0;
  fd.nx1 = (int)header[1]; fd.nx2 = (int)header[2]; fd.nx3 = d3 ? (int)header[3] : 1;
  fd.asim = (float)header[3 + o];
  fd.kr0 = (float)header[5 + o]; fd.kh = (float)header[6 + o]; fd.kaa = (float)header[7 + o]; fd.kbb = (float)header[8 + o];
  fd.kpp = (float)header[9 + o];
  fd.scalefac = fa.scalefac;
  if (fd.doingnth) {  // second header line: nrelbin, relgammamin, relgammamax (fluid_model_koral.f90:318-326)
    double h2[3];
    for (int i = 0; i < 3; i++)
      if (!(hf >> h2[i])) return false;
    fd.nrelbin = (int)h2[0]; fd.relgammamin = h2[1]; fd.relgammamax = h2[2];
    if (fd.nrelbin < 1) return false;
  }
  if (std::fabs(fd.asim - (float)a) > 1e-4f) return false;
  char append[16];
  snprintf(append, sizeof append, d3 ? "%04d" : "%03d", fa.indf);
  std::ifstream df(fa.dfile + append);
  if (!df) return false;
  double tmp;
  for (int i = 0; i < nhead + (fd.doingnth ? 3 : 0); i++) {
    if (!(df >> tmp)) return false;
    if (i == 0) fd.tcur = tmp;  // tcur=header(1)
  }
  size_t n = (size_t)fd.nx1 * fd.nx2 * fd.nx3;
  int dlen = d3 ? 23 : (fd.nrelbin > 0 ? 41 + fd.nrelbin : 38);
  // (sic) nthpos = 43 with rows of dlen = 41 + nrelbin numbers (fluid_model_koral.f90:345-355, 400): bin ie is read from
  // column 42 + ie, so bin nrelbin comes from column dlen + 1 -- in the column-major data(dlen, nx2) block that is column 1
  // of the NEXT row of the block.  For the last row of a block it lies one element past the allocation (undefined in
  // the reference); 0 is used there.
  const int nthpos = 43;
  if (fd.doingnth) fd.nth_arr.assign(n * (size_t)fd.nrelbin, 0.f);
  std::vector<double> prevrow;
  size_t prev_i = 0;
  int rhopos = 10, vpos = 12, bpos = d3 ? 16 : 25, ppos = d3 ? 21 : 33, gridpos = 4, bepos = 22;
  fd.x1_arr.resize(n); fd.x2_arr.resize(n); fd.x3_arr.resize(n); fd.r_arr.resize(n); fd.th_arr.resize(n); fd.ph_arr.resize(n);
  fd.rho_arr.resize(n); fd.p_arr.resize(n); fd.u0_arr.resize(n); fd.vrl_arr.resize(n); fd.vtl_arr.resize(n);
  fd.vpl_arr.resize(n); fd.b0_arr.resize(n); fd.br_arr.resize(n); fd.bth_arr.resize(n); fd.bph_arr.resize(n);
  fd.be_arr.assign(n, 0.f);
  std::vector<double> row(dlen);
  double sq = std::sqrt(fd.scalefac);
  for (size_t i = 0; i < n; i++) {
    for (int c = 0; c < dlen; c++)
      if (!(df >> row[c])) return false;
    fd.x1_arr[i] = (float)row[gridpos - 1]; fd.x2_arr[i] = (float)row[gridpos];
    if (d3) {
      fd.x3_arr[i] = (float)row[gridpos + 1]; fd.r_arr[i] = (float)row[gridpos + 2]; fd.th_arr[i] = (float)row[gridpos + 3];
      fd.ph_arr[i] = (float)row[gridpos + 4];
    } else {
      fd.r_arr[i] = (float)row[gridpos + 1]; fd.th_arr[i] = (float)row[gridpos + 2];
    }
    const double *u = &row[vpos - 1], *b = &row[bpos - 1];
    float vrl, vtl, vpl;
    lnrf_frame_real((float)(u[1] / u[0]), (float)(u[2] / u[0]), (float)(u[3] / u[0]), fd.r_arr[i], fd.asim, fd.th_arr[i], vrl,
                    vtl, vpl);
    fd.rho_arr[i] = (float)(row[rhopos - 1] * fd.scalefac);
    fd.p_arr[i] = (float)row[ppos - 1];  // (electron temperature straight from koral)
    if (d3) fd.be_arr[i] = (float)row[bepos - 1];
    fd.b0_arr[i] = (float)(b[0] * sq); fd.br_arr[i] = (float)(b[1] * sq); fd.bth_arr[i] = (float)(b[2] * sq);
    fd.bph_arr[i] = (float)(b[3] * sq);
    fd.u0_arr[i] = (float)u[0];
    fd.vrl_arr[i] = vrl; fd.vtl_arr[i] = vtl; fd.vpl_arr[i] = vpl;
    if (fd.doingnth) {
      for (int ie = 0; ie < fd.nrelbin; ie++) {
        const int col = nthpos + ie;  // 1-based column of the row
        if (col <= dlen) fd.nth_arr[i * (size_t)fd.nrelbin + ie] = (float)(row[col - 1] * fd.scalefac);
      }
      // the out-of-range last bin of the PREVIOUS row of the same nx2-block is this row's first column
      if (!prevrow.empty() && (i % (size_t)fd.nx2) != 0)
        fd.nth_arr[prev_i * (size_t)fd.nrelbin + (fd.nrelbin - 1)] = (float)(row[0] * fd.scalefac);
      prevrow = row;
      prev_i = i;
    }
  }
  return true;
}

// ------------------------------------------------------------------ HARM data ingestion
// umksh2uks + uks2ubl + lnrf_frame (fluid_model_harm.f90:73-98,380-386,468-485; kerr.f90:131-144).
// Inputs are per-cell MKS contravariant components as the dump stores them.  For the 3-D model
// dthdx2 = pi (fluid_model_harm3d.f90:96).
void harm_transform_and_store(FluidData &fd, const std::vector<double> &rho, const std::vector<double> &uu,
                              const std::vector<double> &ucon, const std::vector<double> &bcon) {
  size_t n = rho.size();
  fd.rho_arr.resize(n); fd.p_arr.resize(n); fd.u0_arr.resize(n); fd.vrl_arr.resize(n); fd.vtl_arr.resize(n);
  fd.vpl_arr.resize(n); fd.b0_arr.resize(n); fd.br_arr.resize(n); fd.bth_arr.resize(n); fd.bph_arr.resize(n);
  bool is3d = (fd.model == F_HARM3D);
  for (size_t i = 0; i < n; i++) {
    float r = fd.r_arr[i], x2 = fd.x2_arr[i], th = fd.th_arr[i];
    float dthdx2 = is3d ? (float)pi : (float)(pi * (1. + (double)(1.f - fd.h) * std::cos(2. * pi * (double)x2)));
    double asim = (double)fd.asim;
    double rr = (double)r;
    double delta = rr * rr - 2.0 * rr + asim * asim;
    double u[4], b[4];
    for (int c = 0; c < 2; c++) {
      const double *src = (c == 0 ? ucon.data() : bcon.data()) + 4 * i;
      double v[4] = {src[0], rr * src[1], src[2] * (double)dthdx2, src[3]};
      double *dst = c == 0 ? u : b;
      dst[0] = v[0] - 2.0 * rr / delta * v[1];
      dst[1] = v[1];
      dst[2] = v[2];
      dst[3] = v[3] - asim / delta * v[1];
    }
    float vrl, vtl, vpl;
    lnrf_frame_real((float)(u[1] / u[0]), (float)(u[2] / u[0]), (float)(u[3] / u[0]), r, fd.asim, th, vrl, vtl, vpl);
    fd.rho_arr[i] = (float)rho[i];
    fd.p_arr[i] = (float)(uu[i] * (double)(fd.gam - 1.f));
    fd.b0_arr[i] = (float)b[0]; fd.br_arr[i] = (float)b[1]; fd.bth_arr[i] = (float)b[2]; fd.bph_arr[i] = (float)b[3];
    fd.u0_arr[i] = (float)u[0];
    fd.vrl_arr[i] = vrl; fd.vtl_arr[i] = vtl; fd.vpl_arr[i] = vpl;
  }
}

// fluid_model_harm.f90:275-408  (ASCII header: 26 numbers; rows of 34 columns)
static bool load_harm(FluidData &fd, const FluidArgs &fa, double a) {
  std::ifstream hf(fa.hfile);
  if (!hf) return false;
  double header[26];
  for (int i = 0; i < 26; i++)
    if (!(hf >> header[i])) return false;
  fd.nx1 = (int)header[1];
  fd.nx2 = (int)header[2];
  fd.asim = (float)header[9];
  fd.gam = (float)header[10];
  fd.h = (float)header[24];
  fd.nx3 = 1;
  if (std::fabs(fd.asim - (float)a) > 1e-4f) {
    fprintf(stderr, "ERROR -- Different simulation and grtrans spin values!\n");
    return false;
  }
  char append[16];
  snprintf(append, sizeof append, "%03d", fa.indf);
  std::ifstream df(fa.dfile + append);
  if (!df) return false;
  for (int i = 0; i < 26; i++)
    if (!(df >> header[i])) return false;
  fd.tcur = header[0];  // :328
  size_t n = (size_t)fd.nx1 * fd.nx2;
  fd.x1_arr.resize(n); fd.x2_arr.resize(n); fd.r_arr.resize(n); fd.th_arr.resize(n);
  std::vector<double> rho(n), uu(n), ucon(4 * n), bcon(4 * n);
  for (size_t i = 0; i < n; i++) {
    double row[34];
    for (int c = 0; c < 34; c++)
      if (!(df >> row[c])) return false;
    fd.x1_arr[i] = (float)row[0]; fd.x2_arr[i] = (float)row[1]; fd.r_arr[i] = (float)row[2]; fd.th_arr[i] = (float)row[3];
    rho[i] = row[4];
    uu[i] = row[5];
    for (int c = 0; c < 4; c++) {
      ucon[4 * i + c] = row[13 + c];
      bcon[4 * i + c] = row[21 + c];
    }
  }
  harm_transform_and_store(fd, rho, uu, ucon, bcon);
  return true;
}

// Synthetic 3-D snapshot container used by config 5 (binary, little endian):
//   int32 nx1,nx2,nx3; float32 asim,gam,h; then float32[n] x1,x2,x3,r,th,ph; float64[n] rho,u;
//   float64[4n] ucon, bcon (MKS components, index (ix1*nx2+ix2)*nx3+ix3).
// The reference's own harm3d input (fluid_model_harm3d.f90:337-366 header file of 15 numbers; :516-584 data file
// `<dfile>%03d.bin` = one text header line + float32 data(35, n): x1,x2,x3 = columns 4-6, rho 10, u 11, u^mu 19-22,
// b^mu 27-30; p = max(u, 1e-18) (gam-1)).
static bool load_harm3d_native(FluidData &fd, const FluidArgs &fa, double a, FILE *fp) {
  std::ifstream hf(fa.hfile);
  if (!hf) { fclose(fp); return false; }
  double header[15];
  for (int i = 0; i < 15; i++)
    if (!(hf >> header[i])) { fclose(fp); return false; }
  fd.nx1 = (int)header[1]; fd.nx2 = (int)header[2]; fd.nx3 = (int)header[3];
  fd.asim = (float)header[10]; fd.gam = (float)header[11]; fd.h = (float)header[13];
  if (std::fabs(fd.asim - (float)a) > 1e-4f) { fclose(fp); return false; }
  // The binary branch of read_harm3d_data_file never assigns tcur (:441-446; only the ASCII branch does, :407), so t(k)
  // and tstep are undefined in the reference when nt > 1.  Here the time is the first number of the text header line,
  // which is what the ASCII branch reads (header(1)).
  int ch;
  std::string line;
  while ((ch = fgetc(fp)) != EOF && ch != 10) line.push_back((char)ch);
  fd.tcur = atof(line.c_str());
  size_t n = (size_t)fd.nx1 * fd.nx2 * fd.nx3;
  std::vector<float> data(35 * n);
  bool ok = fread(data.data(), 4, data.size(), fp) == data.size();
  fclose(fp);
  if (!ok) return false;
  fd.x1_arr.resize(n); fd.x2_arr.resize(n); fd.x3_arr.resize(n); fd.r_arr.resize(n); fd.th_arr.resize(n); fd.ph_arr.resize(n);
  std::vector<double> rho(n), uu(n), ucon(4 * n), bcon(4 * n);
  for (size_t i = 0; i < n; i++) {
    const float *d = data.data() + 35 * i;
    fd.x1_arr[i] = d[3]; fd.x2_arr[i] = d[4]; fd.x3_arr[i] = d[5];
    fd.r_arr[i] = std::exp(fd.x1_arr[i]);
    fd.th_arr[i] = (float)((double)fd.x2_arr[i] * pi);
    fd.ph_arr[i] = fd.x3_arr[i];
    rho[i] = d[9];
    double pp = d[10];
    uu[i] = (pp > (double)1e-18f) ? pp : (double)1e-18f;  // p=merge(p,pmin,p.gt.pmin); pmin=1e-18 is a single-precision literal
    for (int c = 0; c < 4; c++) {
      ucon[4 * i + c] = d[18 + c];
      bcon[4 * i + c] = d[26 + c];
    }
  }
  harm_transform_and_store(fd, rho, uu, ucon, bcon);
  return true;
}

static bool load_harm3d(FluidData &fd, const FluidArgs &fa, double a) {
  {  // the reference's format if `<dfile>%03d.bin` exists, else the synthetic container below
    char append[16];
    snprintf(append, sizeof append, "%03d", fa.indf);
    FILE *fn = fopen((fa.dfile + append + ".bin").c_str(), "rb");
    if (fn) return load_harm3d_native(fd, fa, a, fn);
  }
  FILE *fp = fopen(fa.dfile.c_str(), "rb");
  if (!fp) return false;
  int dims[3];
  float par[3];
  if (fread(dims, 4, 3, fp) != 3 || fread(par, 4, 3, fp) != 3) { fclose(fp); return false; }
  fd.nx1 = dims[0]; fd.nx2 = dims[1]; fd.nx3 = dims[2];
  fd.asim = par[0]; fd.gam = par[1]; fd.h = par[2];
  if (std::fabs(fd.asim - (float)a) > 1e-4f) { fclose(fp); return false; }
  size_t n = (size_t)fd.nx1 * fd.nx2 * fd.nx3;
  std::vector<float> *fl[6] = {&fd.x1_arr, &fd.x2_arr, &fd.x3_arr, &fd.r_arr, &fd.th_arr, &fd.ph_arr};
  for (auto v : fl) {
    v->resize(n);
    if (fread(v->data(), 4, n, fp) != n) { fclose(fp); return false; }
  }
  std::vector<double> rho(n), uu(n), ucon(4 * n), bcon(4 * n);
  bool ok = fread(rho.data(), 8, n, fp) == n && fread(uu.data(), 8, n, fp) == n &&
            fread(ucon.data(), 8, 4 * n, fp) == 4 * n && fread(bcon.data(), 8, 4 * n, fp) == 4 * n;
  fclose(fp);
  if (!ok) return false;
  harm_transform_and_store(fd, rho, uu, ucon, bcon);
  return true;
}

// fluid.f90:163-243 with one dump slot
static bool load_fluid_single(FluidData &fd, const std::string &fname, double a, const FluidArgs &fa) {
  fd = FluidData();
  if (fname == "SPHACC") {
    fd.model = F_SPHACC;
    init_sphacc(fd);
  } else if (fname == "FFJET") {
    fd.model = F_FFJET;
    fd.ok = load_ffjet(fd, fa.dfile);
  } else if (fname == "THINDISK") {
    fd.model = F_THINDISK;
    fd.td_mdot = (float)fa.mdot; fd.td_mbh = (float)fa.mbh; fd.td_rin = (float)fa.rin; fd.td_rout = (float)fa.rout;
  } else if (fname == "HARM") {
    fd.model = F_HARM;
    fd.ok = load_harm(fd, fa, a);
  } else if (fname == "HARM3D") {
    fd.model = F_HARM3D;
    fd.ok = load_harm3d(fd, fa, a);
  } else if (fname == "IHARM") {
    fd.model = F_IHARM;
    fd.ok = load_iharm(fd, fa, a);
  } else if (fname == "KORAL" || fname == "KORAL3D" || fname == "KORAL3D_DISK" || fname == "KORALNTH") {
    fd.model = (fname == "KORAL" || fname == "KORALNTH") ? F_KORAL : (fname == "KORAL3D" ? F_KORAL3D : F_KORAL3D_DISK);
    fd.doingnth = (fname == "KORALNTH");  // fluid.f90:218-220, 343-349: the 2-D koral model with non-thermal electron bins
    fd.ktype = (fname == "KORAL3D_DISK") ? 1 : 0;
    fd.ok = load_koral(fd, fa, a);
  } else {
    fd.model = F_DUMMY;
  }
  return fd.ok;
}

static bool has_time_axis(int model) {
  return model == F_HARM || model == F_HARM3D || model == F_IHARM || model == F_KORAL || model == F_KORAL3D ||
         model == F_KORAL3D_DISK;
}
static std::vector<float> *slot_fields(FluidData &fd, int q) {
  std::vector<float> *f[12] = {&fd.rho_arr, &fd.p_arr, &fd.vrl_arr, &fd.vtl_arr, &fd.vpl_arr, &fd.b0_arr, &fd.br_arr,
                               &fd.bth_arr, &fd.bph_arr, &fd.kela_arr, &fd.be_arr, &fd.u0_arr};
  return f[q];
}

// load_*_data(nload) (fluid_model_harm.f90:446-509 and its copies): dump files indf, indf-1, ... fill slots 0..nload-1 of
// the nt-slot arrays; with more than one file read, tstep = t(1)-t(2).
static bool load_slots(FluidData &fd, const std::string &fname, double a, const FluidArgs &fa, int nload) {
  for (int k = 0; k < nload; k++) {
    FluidData one;
    FluidArgs fb = fa;
    fb.indf = fd.indf - k;
    fb.nt = 1;
    if (!load_fluid_single(one, fname, a, fb)) return false;
    for (int q = 0; q < 12; q++) {
      std::vector<float> &dst = *slot_fields(fd, q), &src = *slot_fields(one, q);
      if (src.empty()) continue;
      if (dst.size() < fd.ncell * (size_t)fd.nt) dst.resize(fd.ncell * (size_t)fd.nt, 0.f);
      std::copy(src.begin(), src.end(), dst.begin() + (size_t)k * fd.ncell);
    }
    fd.tdump[k] = one.tcur;
  }
  if (nload > 1) {
    fd.tstep_d = fd.tdump[0] - fd.tdump[1];
    fd.tstep = (float)fd.tstep_d;
  }
  return true;
}

// fluid.f90:163-243
bool load_fluid_model(FluidData &fd, const std::string &fname, double a, const FluidArgs &fa) {
  if (!load_fluid_single(fd, fname, a, fa)) return false;
  fd.indf = fa.indf;
  fd.nt = 1;
  fd.ncell = fd.rho_arr.size();
  fd.tdump.assign(1, fd.tcur);
  if (has_time_axis(fd.model) && fa.nt > 1) {
    if (fd.doingnth) return fd.ok = false;  // KORALNTH: one dump resident (the bins have no slot axis here)
    fd.nt = fa.nt;
    fd.tdump.assign(fa.nt, 0.0);
    fd.ok = load_slots(fd, fname, a, fa, fa.nt);
  }
  return fd.ok;
}

// advance_*_timestep + update_*_data (fluid_model_harm.f90:511-546; identical code in the harm3d :689-723, iharm :799-841,
// koral3d :635-669 modules).  tstep/toffset are real(4) (real(8) in iharm); mixed-mode arithmetic as written.  (sic) the shift
// of the slots leaves u0_arr where it was: after an update by 0 < nupdate <= nt slots the older slots keep the u^t of the
// dump that occupied them before.
void advance_fluid_timestep(FluidData &fd, const std::string &fname, double a, const FluidArgs &fa, double dtobs) {
  if (!has_time_axis(fd.model)) return;
  int nupdate;
  if (fd.model == F_IHARM) {
    nupdate = (int)std::floor(dtobs / fd.tstep_d);
    fd.toffset_d = fd.toffset_d + dtobs - fd.tstep_d * nupdate;
    nupdate = nupdate + (int)std::floor(fd.toffset_d / fd.tstep_d);
    fd.toffset_d = fd.toffset_d - std::floor(fd.toffset_d / fd.tstep_d) * fd.tstep_d;
  } else {
    nupdate = (int)std::floor(dtobs / (double)fd.tstep);
    fd.toffset = (float)((double)fd.toffset + dtobs - (double)(fd.tstep * (float)nupdate));
    nupdate = nupdate + (int)std::floor(fd.toffset / fd.tstep);
    fd.toffset = fd.toffset - (float)(int)std::floor(fd.toffset / fd.tstep) * fd.tstep;
  }
  fd.indf += nupdate;
  if (nupdate > fd.nt) {
    fd.ok = load_slots(fd, fname, a, fa, fd.nt);
  } else if (nupdate > 0) {
    size_t sh = fd.ncell * (size_t)nupdate, tot = fd.ncell * (size_t)fd.nt;
    for (int q = 0; q < 11; q++) {  // every field but u0_arr
      std::vector<float> &v = *slot_fields(fd, q);
      if (v.size() == tot && sh < tot) std::copy_backward(v.begin(), v.begin() + (tot - sh), v.end());
    }
    fd.ok = load_slots(fd, fname, a, fa, nupdate);
  }
}

// fluid.f90:473-526
void get_fluid_vars(const FluidData &fd, const Geo &g, std::vector<FluidPt> &f) {
  f.assign(g.npts, FluidPt());
  switch (fd.model) {
    case F_SPHACC: get_sphacc_fluidvars(fd, g, f); break;
    case F_FFJET: ffjet_vals(fd, g, f); break;
    case F_THINDISK: get_thindisk_fluidvars(fd, g, f); break;
    case F_HARM: harm_vals(fd, g, f); break;
    case F_HARM3D: harm3d_vals(fd, g, f); break;
    case F_IHARM: iharm_vals(fd, g, f); break;
    case F_KORAL:
    case F_KORAL3D:
    case F_KORAL3D_DISK: koral_vals(fd, g, f); break;
    default: break;
  }
}

// fluid.f90:765-790
static void scale_sim_units(double mbh, double mdotcgs, double mdot, float rho, float p, float bmag, double &ncgs,
                            double &bcgs, double &tempcgs) {
  double lcgs = pc::G * mbh * pc::msun / (pc::c * pc::c), tcgs = lcgs / pc::c;
  double rhocgs = mdotcgs / mdot / (lcgs * lcgs * lcgs) * tcgs * (double)rho;
  ncgs = rhocgs / pc::mp;
  double pcgs = (double)p * rhocgs / (double)rho * std::pow(pc::c, 2.0);
  tempcgs = pcgs / ncgs / pc::k;
  bcgs = (double)bmag * std::sqrt(rhocgs / (double)rho) * pc::c;
  bcgs = bcgs * std::sqrt(4.0 * pi);
}

// fluid.f90:870-892
static double monika_e(float rho, float p, float b, double rlow, double rhigh) {
  (void)rho;
  double beta = (double)(p / (b * b)) / 0.5;
  double b2 = beta * beta;
  if ((double)b > 0.0) return rhigh * b2 / (1.0 + b2) + rlow / (1.0 + b2);
  return rhigh;
}

// fluid.f90:528-584 and the per-model converters (:957-993, :1164-1172, :1190-1196, :1249-1259)
void convert_fluid_vars(const FluidData &fd, const std::vector<FluidPt> &f, const SourceParams &sp,
                        std::vector<double> &ncgs, std::vector<double> &ncgsnth, std::vector<double> &bcgs,
                        std::vector<double> &tcgs) {
  size_t n = f.size();
  ncgs.assign(n, 0.0); ncgsnth.assign(n, 0.0); bcgs.assign(n, 0.0); tcgs.assign(n, 0.0);
  for (size_t i = 0; i < n; i++) {
    const FluidPt &p = f[i];
    switch (fd.model) {
      case F_SPHACC:
        ncgs[i] = p.rho; bcgs[i] = p.bmag; tcgs[i] = p.p; ncgsnth[i] = 0.;
        break;
      case F_FFJET:
        ncgsnth[i] = (double)p.rho * sp.nfac; bcgs[i] = (double)p.bmag * sp.bfac; ncgs[i] = 0.; tcgs[i] = 0.;
        break;
      case F_THINDISK:
        tcgs[i] = p.rho; ncgs[i] = 1.;
        break;
      case F_HARM: {
        double mdot = .003;
        scale_sim_units(sp.mbh, sp.mdot, mdot, p.rho, p.p, p.bmag, ncgs[i], bcgs[i], tcgs[i]);
        double trat = monika_e(p.rho, p.p, p.bmag, sp.muval, sp.muval / sp.gminval);
        tcgs[i] = tcgs[i] * trat;
        ncgsnth[i] = ncgs[i];
      } break;
      case F_HARM3D: {
        double mdot = pc::G * sp.mbh * pc::msun / (pc::c * pc::c * pc::c);
        scale_sim_units(sp.mbh, sp.mdot, mdot, p.rho, p.p, p.bmag, ncgs[i], bcgs[i], tcgs[i]);
        double trat = monika_e(p.rho, p.p, p.bmag, 1.0 / sp.muval - 1.0, sp.gminval * (1.0 / sp.muval - 1.0));
        tcgs[i] = tcgs[i] / (1.0 + trat);
        // nonthermale_b2 (fluid.f90:914-923)
        double bmagrho = std::pow((double)p.bmag, 2.0) / (double)p.rho;
        if (bmagrho > 1.0)
          ncgsnth[i] = sp.jetalphaval * std::pow(bcgs[i], 2.0) / 8.0 / pi / sp.gminval * (sp.p1 - 2.0) / (sp.p1 - 1.0) / 8.2e-7;
        else
          ncgsnth[i] = 0.0;
        // andrew_sigcut (fluid.f90:792-810)
        double rhocgs = ncgs[i] * pc::mp;
        double sigmacgs = (bcgs[i] * bcgs[i]) / (rhocgs * 8.988e20 * 4.0 * pi);
        if (sigmacgs >= sp.sigcut) {
          ncgs[i] = 0.0;
          tcgs[i] = 1e9;
        }
      } break;
      case F_KORAL:
      case F_KORAL3D:
      case F_KORAL3D_DISK: {  // convert_fluidvars_koral (fluid.f90:1075-1162); quantities are already cgs
        double rhocgs = (double)p.rho;
        ncgs[i] = rhocgs / pc::mp;
        bcgs[i] = (double)p.bmag * std::sqrt(4 * pi);
        bcgs[i] = bcgs[i] * std::sqrt(sp.nfac);
        ncgs[i] = ncgs[i] * sp.nfac;
        rhocgs = rhocgs * sp.nfac;
        if (sp.gminval >= 1.0) {  // charles_e(rho, p+Be, 2p+Be, bmag, 1, 1, gmin): p = T_e, Be = T_p here (fluid.f90:845-868)
          float pp = p.p + p.be, uu = 2.f * p.p + p.be;
          double beta = 2.0 * (double)p.rho * pc::k * (double)pp / pc::mp / (double)(p.bmag * p.bmag);
          double b2 = beta * beta;
          double trat = ((double)p.bmag > 0.0) ? sp.gminval * b2 / (1.0 + b2) + 1.0 / (1.0 + b2) : sp.gminval;
          tcgs[i] = (double)uu / (2.0 + trat);
        } else if (sp.gminval < 0.0) {
          tcgs[i] = (double)p.p;
        }
        double sigmacgs = (bcgs[i] * bcgs[i]) / (rhocgs * 8.988e20 * 4.0 * pi);
        if (sigmacgs >= sp.sigcut) {
          ncgs[i] = 0.0;
          tcgs[i] = 1e9;
        }
        if (fd.ktype == 1 && p.be >= 0.05f) {  // zero out the jet (Bernoulli cut)
          tcgs[i] = 10.;
          bcgs[i] = 0.;
        }
        ncgsnth[i] = 0.;
      } break;
      case F_IHARM: {  // convert_fluidvars_iharm (fluid.f90:995-1026)
        double mdot = pc::G * sp.mbh * pc::msun / (pc::c * pc::c * pc::c);
        scale_sim_units(sp.mbh, sp.mdot, mdot, p.rho, p.p, p.bmag, ncgs[i], bcgs[i], tcgs[i]);
        double rhocgs = ncgs[i] * pc::mp;
        if (sp.gminval >= 1.0) {
          // f%p holds the internal energy u: monika_e gets p = u (gamma-1) (real(4) product)
          double trat = monika_e(p.rho, p.p * (fd.gam - 1.f), p.bmag, 1.0 / sp.muval - 1.0, sp.gminval * (1.0 / sp.muval - 1.0));
          tcgs[i] = 2.0 * tcgs[i] / 3.0 / (2.0 + trat);
        } else if (sp.gminval < 0.0) {
          if (sp.gminval == -1.0) {  // ressler_e (fluid.f90:894-903): electron entropy -> T_e
            double gamma = 4.0 / 3.0;
            double thetae = pc::mp / pc::m * (double)p.kela * std::pow((double)p.rho, gamma - 1.);
            tcgs[i] = thetae * pc::m * pc::c2 / pc::k;
          }
        }
        double bmagrho = std::pow((double)p.bmag, 2.0) / (double)p.rho;
        if (bmagrho > 1.0)
          ncgsnth[i] = sp.jetalphaval * std::pow(bcgs[i], 2.0) / 8.0 / pi / sp.gminval * (sp.p1 - 2.0) / (sp.p1 - 1.0) / 8.2e-7;
        else
          ncgsnth[i] = 0.0;
        double sigmacgs = (bcgs[i] * bcgs[i]) / (rhocgs * 8.988e20 * 4.0 * pi);
        if (sigmacgs >= sp.sigcut) {
          ncgs[i] = 0.0;
          tcgs[i] = 1e9;
        }
      } break;
      default: break;
    }
  }
}

// calcgmin.f90:6-78: minimum Lorentz factor and non-thermal fraction of a power-law tail matched to a thermal distribution
// of dimensionless temperature theta, carrying the fraction eta of its energy.  (sic) the p = 3.5 branch assigns the
// gmin_lin_sin_* constants twice (the second set wins) and never assigns gmin_inv_sin_*: uninitialised in the reference,
// taken as 0 here.
static void calc_gmin(double p, double theta, double eta, double &gmin, double &nfrac) {
  const double acenter = 0.5668090982352612, anormal = (double)0.52624783f, azero = 3.0 / std::sqrt(2.0),
               astwo = std::log(std::sqrt(2.0));
  double lin_cons, lin_coeff, lin_power, inv_cons, inv_coeff, inv_power, inv_sin_coeff, inv_sin_freq, inv_sin_delay,
      lin_sin_coeff, lin_sin_freq, lin_sin_delay;
  if (p == 3.5) {
    lin_power = 0.276589155355; lin_coeff = -13.5593749125; lin_cons = 16.0797900684;
    inv_power = 6.53997654139; inv_coeff = 151.597731214; inv_cons = 0.722506578136;
    lin_sin_coeff = lin_sin_freq = lin_sin_delay = 0.121815691108;
    inv_sin_coeff = inv_sin_freq = inv_sin_delay = 0.0;
  } else {
    lin_cons = 21.38307186; lin_coeff = -16.7811712; lin_power = 0.15128533;
    inv_cons = 0.74798712; inv_coeff = 0.62609462; inv_power = 0.81567379;
    inv_sin_coeff = .00638946501; inv_sin_freq = -16.8034428; inv_sin_delay = 3.72208398;
    lin_sin_coeff = lin_sin_freq = lin_sin_delay = 0.0;
  }
  // (the constants above are default-real literals in the reference: single precision promoted to double)
  auto f4 = [](double v) { return (double)(float)v; };
  double linc = f4(lin_cons) + f4(lin_coeff) * std::pow(eta, f4(lin_power)) +
                f4(lin_sin_coeff) * std::sin(eta * f4(lin_sin_freq) + f4(lin_sin_delay));
  double invc = f4(inv_cons) + f4(inv_coeff) * std::pow(eta, f4(inv_power)) +
                f4(inv_sin_coeff) * std::sin(eta * f4(inv_sin_freq) + f4(inv_sin_delay));
  double raw = theta * linc + invc;
  gmin = (raw >= 1.0) ? raw : 1.0;
  double atheta = theta * azero * std::exp(astwo * std::tanh(anormal * std::log(theta / f4(acenter))));
  nfrac = eta * atheta * ((p - 2.) / (p - 1.)) * std::pow(gmin, p - 2.);
}

// fluid.f90:1640-1677 assign_source_params: gmin per point ('const': gminval) and, for stype = 'tail', the non-thermal
// density of the matched tail, which replaces the fluid model's own
void assign_source_params(const SourceParams &sp, const std::vector<double> &ncgs, const std::vector<double> &tcgs,
                          std::vector<double> &ncgsnth, std::vector<double> &gmin) {
  size_t n = ncgs.size();
  gmin.assign(n, sp.gminval);
  if (sp.type != 1) return;
  for (size_t i = 0; i < n; i++) {
    double g, x;
    calc_gmin(sp.p2, sp.muval * pc::k * tcgs[i] / pc::m / pc::c / pc::c, sp.jetalphaval, g, x);
    double gmax = sp.gmax;
    gmin[i] = (g <= gmax) ? g : gmax / 2.0;
    double factor = (g <= gmax) ? 1.0 : std::pow(gmax / 2.0 / g, sp.p2 - 2.);
    ncgsnth[i] = factor * ((x > 0.0) ? x * ncgs[i] * std::pow(gmin[i], 1. - sp.p2) : 0.0);
  }
}

}  // namespace orc
