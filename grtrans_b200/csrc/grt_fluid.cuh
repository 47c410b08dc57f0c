// grtrans_b200 -- fluid sampling device functions (one geodesic point per thread).
//
// Replaces fluid.f90:473-526 (get_fluid_vars_arr) and the *_vals routines of
// fluid_model_thindisk.f90:51-84, fluid_model_sphacc.f90:450-497, fluid_model_ffjet.f90:42-171,
// fluid_model_harm.f90:100-265, fluid_model_harm3d.f90:107-327 with interpolate.f90 / hunt.f /
// math.f90:zbrent, plus the unit conversions fluid.f90:765-810, 870-892, 914-993, 1164-1259.
//
// Snapshot layout in HBM: the reference keeps ten float32 SoA arrays (rho,p,u0,vrl,vtl,vpl,b0,br,bth,
// bph).  Every sample needs all ten fields at the same 4/8 corners, so the snapshot is re-packed as
// an array of 48-byte cells {rho,p,vrl,vtl | vpl,b0,br,bth | bph,u0,-,-}: one corner = three
// 128-bit read-only loads instead of ten scattered 32-bit loads.  Index and weight arithmetic stays in
// float32 exactly where the reference has `real` variables so the gathered cells are the same.
#pragma once
#include "grt_kerr.cuh"

namespace grt {

enum { FL_DUMMY = 0, FL_SPHACC = 1, FL_THINDISK = 2, FL_HARM = 12, FL_FFJET = 13, FL_HARM3D = 19, FL_KORAL = 22, FL_KORAL3D = 25, FL_KORAL3D_DISK = 26, FL_IHARM = 29 };

struct Cell {  // 48 bytes, 16-byte aligned
  float rho, p, vrl, vtl, vpl, b0, br, bth, bph, u0, kela /* iharm electron entropy */, be /* koral3d Bernoulli parameter */;
};

struct FluidDev {
  int model, nx1, nx2, nx3;
  float h, rh_dummy;
  float td_mbh, td_mdot, td_rin, td_rout;
  const float *ax1;   // uniqx1 (log r for HARM/FFJET)          [nx1]
  const float *axr;   // uniqr                                   [nx1]
  const float *ax2;   // uniqx2                                  [nx2]
  const float *axth;  // uniqth                                  [nx2]
  const float *ax3;   // uniqx3 / uniqph                         [nx3]
  const Cell *cells;  // [nslots][ncell] packed snapshot(s): slot k = dump file indf-k (load_harm_data, fluid_model_harm.f90:459-486)
  int nslots;                  // nt of the fluid module (dumps resident for slow light)
  unsigned long long ncell;    // cells per dump
  float tstep;                 // dump spacing, real(4) module variable (t(1)-t(2), default 20.)
  // iharm grid (fluid_model_iharm.f90:13-24, real(8) module variables read from the header file)
  double hslope, mks_smooth, poly_xt, poly_alpha, startx1;
  int is_mmks, eheat;
  float gam;
  // koral / koral3d MKS3 grid (fluid_model_koral.f90:11-29) and trust region in x2 (minpolecell cells at each pole)
  float kh, kr0, kaa, kbb, kpp, x2mintrust, x2maxtrust;
  int ktype;
  // KORALNTH (fluid_model_koral.f90:16-30): dn/dgamma of the non-thermal electrons, [ncell][nrelbin] real(4) (else nullptr / 0)
  const float *nth;
  int nrelbin;
  // SPHACC tables
  const float *uvals, *vvals, *uv, *tvals;  // 498, 498, 461 (=1/rvals), 461
};

struct FluidSample {
  float rho, p, bmag, kela, be;
  double u[4], b[4];
};

struct Field10 {
  float rho, p, vrl, vtl, vpl, b0, br, bth, bph, u0, kela, be;
};

GRT_HD Field10 load_cell(const Cell *cells, size_t idx) {
  Field10 f;
#if defined(__CUDA_ARCH__)
  const float4 *q = reinterpret_cast<const float4 *>(cells + idx);
  float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
  f.rho = a.x; f.p = a.y; f.vrl = a.z; f.vtl = a.w; f.vpl = b.x; f.b0 = b.y; f.br = b.z; f.bth = b.w; f.bph = c.x; f.u0 = c.y;
  f.kela = c.z; f.be = c.w;
#else
  const Cell &c = cells[idx];
  f.rho = c.rho; f.p = c.p; f.vrl = c.vrl; f.vtl = c.vtl; f.vpl = c.vpl; f.b0 = c.b0; f.br = c.br; f.bth = c.bth;
  f.bph = c.bph; f.u0 = c.u0; f.kela = c.kela; f.be = c.be;
#endif
  return f;
}

#define GRT_FOR_FIELDS(OP) OP(rho) OP(p) OP(vrl) OP(vtl) OP(vpl) OP(b0) OP(br) OP(bth) OP(bph) OP(u0)
#define GRT_FOR_FIELDS_BE(OP) GRT_FOR_FIELDS(OP) OP(be)

// bilinear blend of four corners (interpolate.f90:171-182), all float32
GRT_HD Field10 blend4(const Field10 &c0, const Field10 &c1, const Field10 &c2, const Field10 &c3, float xd, float yd) {
  Field10 o;
#define OP(f) { float w1 = c0.f * (1.f - yd) + c1.f * yd; float w2 = c2.f * (1.f - yd) + c3.f * yd; o.f = w1 * (1.f - xd) + w2 * xd; }
  GRT_FOR_FIELDS(OP)
#undef OP
  return o;
}

// finish a gridded sample: b.b with the float32 metric, LNRF^-1 velocity, u^i = u^t v^i
GRT_HD void finish_grid(FluidSample &s, const Field10 &v, bool in, float fill_vtl, float zr, float th, float a,
                        double bfac) {
  float vrl0 = in ? v.vrl : 0.f, vtl0 = in ? v.vtl : fill_vtl, vpl0 = in ? v.vpl : 0.f;
  s.b[0] = in ? v.b0 : 1.f;
  s.b[1] = in ? v.br : 1.f;
  s.b[2] = in ? v.bth : 1.f;
  s.b[3] = in ? v.bph : 1.f;
  s.u[0] = in ? (double)v.u0 : 1.0;
  MetricF mf = bl_metric_f(zr, th, a);
  Metric m;
  m.tt = mf.tt; m.tp = mf.tp; m.rr = mf.rr; m.hh = mf.hh; m.pp = mf.pp;
  float bm = (float)mdot(m, s.b, s.b);
  bm = ((double)bm >= 0.0) ? bm : 0.f;
  bm = sqrtf(bm);
  s.bmag = (float)((double)bm * bfac);
  float vr0, vt0, vp0;
  lnrf_inv_f(vrl0, vtl0, vpl0, zr, a, th, vr0, vt0, vp0);
  s.u[1] = s.u[0] * (double)vr0;
  s.u[2] = s.u[0] * (double)vt0;
  s.u[3] = s.u[0] * (double)vp0;
}

// ---- Brent root finder specialised to the HARM theta(x2) map (math.f90:56-135, tol=1e-6) ---------
GRT_HD double harm_x2_of_theta(double theta, double h) {
  const double pi = GRT_PI, eps = 2.220446049250313e-16, tol = 1e-6;
#define GRT_F(x) (theta - pi * (x) - .5 * (1 - h) * sin(2. * pi * (x)))
  double a = 0.0, b = 1.0, c, d = 0, e = 0, fa = GRT_F(a), fb = GRT_F(b), fc, p, q, r, s, tol1, xm;
  if ((fa > 0.0 && fb > 0.0) || (fa < 0.0 && fb < 0.0)) return nan("");
  c = b;
  fc = fb;
  for (int iter = 0; iter < 100; iter++) {
    if ((fb > 0.0 && fc > 0.0) || (fb < 0.0 && fc < 0.0)) {
      c = a; fc = fa; d = b - a; e = d;
    }
    if (fabs(fc) < fabs(fb)) {
      a = b; b = c; c = a; fa = fb; fb = fc; fc = fa;
    }
    tol1 = 2.0 * eps * fabs(b) + 0.5 * tol;
    xm = 0.5 * (c - b);
    if (fabs(xm) <= tol1 || fb == 0.0) return b;
    if (fabs(e) >= tol1 && fabs(fa) > fabs(fb)) {
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
      p = fabs(p);
      if (2.0 * p < dmin(3.0 * xm * q - fabs(tol1 * q), fabs(e * q))) {
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
    b = b + ((fabs(d) > tol1) ? d : (signbit(xm) ? -fabs(tol1) : fabs(tol1)));
    fb = GRT_F(b);
  }
#undef GRT_F
  return b;
}

// ---- generic Brent root finder (math.f90:56-135 zbrent, scalar) ----------------------------------------
template <class Fn>
GRT_HD double zbrent_hd(const Fn &func, double x1, double x2, double tol) {
  const double eps = 2.220446049250313e-16;
  double a = x1, b = x2, c, d = 0, e = 0, fa = func(a), fb = func(b), fc, p, q, r, s, tol1, xm;
  if ((fa > 0.0 && fb > 0.0) || (fa < 0.0 && fb < 0.0)) return nan("");
  c = b;
  fc = fb;
  for (int iter = 0; iter < 100; iter++) {
    if ((fb > 0.0 && fc > 0.0) || (fb < 0.0 && fc < 0.0)) {
      c = a; fc = fa; d = b - a; e = d;
    }
    if (fabs(fc) < fabs(fb)) {
      a = b; b = c; c = a; fa = fb; fb = fc; fc = fa;
    }
    tol1 = 2.0 * eps * fabs(b) + 0.5 * tol;
    xm = 0.5 * (c - b);
    if (fabs(xm) <= tol1 || fb == 0.0) return b;
    if (fabs(e) >= tol1 && fabs(fa) > fabs(fb)) {
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
      p = fabs(p);
      if (2.0 * p < dmin(3.0 * xm * q - fabs(tol1 * q), fabs(e * q))) {
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
    b = b + ((fabs(d) > tol1) ? d : (signbit(xm) ? -fabs(tol1) : fabs(tol1)));
    fb = func(b);
  }
  return b;
}

// ---- iharm coordinate maps (fluid_model_iharm.f90:82-118) ------------------------------------------------
// theta(x2) of the MKS grid (real(8)) and theta(x2, x1) of the MMKS / FMKS grid (a real(4) function whose
// sub-expressions are promoted by the real(8) module variables pi, hslope, startx1 exactly as written)
struct IharmGrid {
  double hslope, mks_smooth, poly_xt, poly_alpha, startx1;
};
GRT_HD double iharm_th_mks(const IharmGrid &G, double x2) {
  return GRT_PI * x2 + (1. - G.hslope) / 2. * sin(2. * GRT_PI * x2);
}
GRT_HD float iharm_th_mmks(const IharmGrid &G, float x2, float x1) {
  float A = (float)G.mks_smooth, B = (float)G.poly_xt, C = (float)G.poly_alpha;
  float D = (float)(GRT_PI / (double)(2.f + 2.f / (r4pow(B, C) * (1 + C))));
  float thetag = (float)(GRT_PI * (double)x2 + (1. - G.hslope) / 2. * sin(2. * GRT_PI * (double)x2));
  float thetaj = (float)((double)(D * (2.f * x2 - 1.f) * (1.f + r4pow((2.f * x2 - 1.f) / B, C) / (1.f + C))) + GRT_PI / 2.);
  return (float)((double)thetag + exp(-(double)A * ((double)x1 - G.startx1)) * (double)(thetaj - thetag));
}
struct IharmThetaMmks {
  IharmGrid G;
  double x1, th;
  GRT_HD double operator()(double x) const { return th - (double)iharm_th_mmks(G, (float)x, (float)x1); }
};
struct IharmThetaMks {
  IharmGrid G;
  double th;
  GRT_HD double operator()(double x) const { return th - iharm_th_mks(G, x); }
};

// ---- koral MKS3 coordinate maps (fluid_model_koral.f90:62-80; real(4) arguments promoted by the real(8) pi) -------
GRT_HD float koral_x2_of_theta(float th, float r, float h, float a, float b, float p) {
  float rp = r4pow(r, p);
  double num = atan(tan((double)(0.5f * h) * GRT_PI) * (1. - (double)(2.f * th) / GRT_PI));
  float den = (b - a) * r4pow(2.f, p) + (a - 0.5f) * rp;
  return (float)(0.5 * (1. + ((double)rp / ((double)h * GRT_PI)) * (num / (double)den)));
}
GRT_HD float koral_theta_of_x2(float x2, float r, float h, float a, float b, float p) {
  float inner = -0.5f + x2 + (1.f - 2.f * x2) * (a + r4pow(2.f, p) * (b - a) / r4pow(r, p));
  return (float)(0.5 * GRT_PI * (1. + tan((double)h * GRT_PI * (double)inner) / tan((double)(0.5f * h) * GRT_PI)));
}

// bracket search equivalent to NR hunt/locate for an in-range x (1-based result clamped to [1,n-1])
GRT_HD int bracket(const float *xx /*0-based*/, int n, float x) {
  bool ascnd = xx[n - 1] >= xx[0];
  int jl = 0, ju = n + 1;
  while (ju - jl > 1) {
    int jm = (ju + jl) >> 1;
    if (ascnd == (x >= xx[jm - 1]))
      jl = jm;
    else
      ju = jm;
  }
  if (x == xx[0]) jl = 1;
  if (x == xx[n - 1]) jl = n - 1;
  if (jl < 1) jl = 1;
  if (jl > n - 1) jl = n - 1;
  return jl;
}

GRT_HD float krolik_c(float r, float a) {
  float pif = r4acos(-1.f);
  float rms = isco_radius_f(a);
  float y = sqrtf(r), yms = sqrtf(rms);
  float ac = r4acos(a);
  float y1 = (float)(2.0 * (double)r4cos(1.f / 3.f * (ac - pif)));
  float y2 = (float)(2.0 * (double)r4cos(1.f / 3.f * (ac + pif)));
  float y3 = (float)(-2.0 * (double)r4cos(1.f / 3.f * ac));
  float arg1 = (float)((double)(3.f * a) / (2.0 * (double)y));
  float arg2 = (float)(3.0 * sq((double)(y1 - a)) / (double)(y * y1 * (y1 - y2) * (y1 - y3)));
  float arg3 = (float)(3.0 * sq((double)(y2 - a)) / (double)(y * y2 * (y2 - y1) * (y2 - y3)));
  float arg4 = (float)(3.0 * sq((double)(y3 - a)) / (double)(y * y3 * (y3 - y1) * (y3 - y2)));
  return 1.f - yms / y - arg1 * r4log(y / yms) - arg2 * r4log((y - y1) / (yms - y1)) - arg3 * r4log((y - y2) / (yms - y2)) -
         arg4 * r4log((y - y3) / (yms - y3));
}

// ---- the sampler ---------------------------------------------------------------------------------
// Dump slot of a geodesic point (fluid_model_harm.f90:133-183; same lines in the harm3d / koral / koral3d modules): tt = KS
// look-back time of the point relative to the first point of the ray, real(4); tindx = floor(tt/tstep), at most nt-1.  The
// time weight is zero (rttd=0.), so the sample is slot tindx alone.  harm and koral add 1d-8 to tt before the floor.
GRT_HD double bl2ks_time(double r, double x, double a) {  // kerr.f90:147-154
  double sq1 = sqrt(1. - a * a);
  return x + log(r * r - 2 * r + a * a) + 1. / (2.0 * sq1) * log((r - 1. - sq1) / (r - 1 + sq1));
}
GRT_HD const Cell *slot_cells(const FluidDev &F, float a, double rr, double tpt, double r1) {
  if (F.nslots <= 1) return F.cells;
  float zr = (float)rr, zt = (float)tpt, zr1 = (float)r1;
  float tt = (float)(bl2ks_time((double)zr, (double)zt, (double)a) - bl2ks_time((double)zr1, 0.0, (double)a));
  tt = (F.model == FL_HARM3D) ? -tt : (float)((double)(-tt) + 1e-8);
  float q = floorf(tt / F.tstep);
  int tindx = (q <= (float)(F.nslots - 2)) ? (int)q : F.nslots - 1;
  if (tindx < 0) tindx = 0;  // guard
  return F.cells + F.ncell * (unsigned long long)tindx;
}

GRT_HDN void fluid_sample(const FluidDev &F, double ad, double rr, double thh, double phh, double tpt, double r1,
                          const double *k4, FluidSample &s) {
  const double pi = GRT_PI;
  // reference constants (phys_constants.f90)
  const double pcG = 6.67e-8, pcc = 2.99792458e10, pcc2 = pcc * pcc, pcmsun = 1.998e33, pcmp = 1.67262158e-24,
               pcsigt = 6.6523e-25, pcsigb = 5.6704e-5;
  float a = (float)ad;
  s.rho = s.p = s.bmag = s.kela = s.be = 0.f;
  for (int i = 0; i < 4; i++) s.u[i] = s.b[i] = 0.0;
  const Cell *const cells = slot_cells(F, a, rr, tpt, r1);
  switch (F.model) {
    case FL_THINDISK: {
      float r = (float)rr, th = (float)thh;
      float MBH = F.td_mbh, Mdot = F.td_mdot;
      float lbh = (float)((double)MBH * pcmsun * pcG / pcc2);
      float b = 1.f - 3.f / r + 2.f * a / r4pow(r, 1.5f);
      float rms = isco_radius_f(a);
      float rin = fmaxf(rms, F.td_rin);
      float kc = krolik_c(r, a);
      float d = r * r - 2.f * r + a * a;
      float lc = (rms * rms - 2.f * a * sqrtf(rms) + a * a) / (r4pow(rms, 1.5f) - 2.f * sqrtf(rms) + a);
      float hc = (2.f * r - a * lc) / d;
      float sn = r4sin(th);
      float ar = (r * r + a * a) * (r * r + a * a) - a * a * d * (sn * sn);
      float om = 2.f * a * r / ar;
      float Mdotedd = (float)(4. * pi * pcG * (double)MBH * pcmsun * pcmp * pcc / pcsigt / pcc2);
      float T0 = (float)pow(3.0 / 8.0 / pi * pcG * (double)MBH * pcmsun * (double)Mdot * (double)Mdotedd / (double)lbh /
                                (double)lbh / (double)lbh / pcsigb,
                            0.25);
      float omega;
      if (r > rms)
        omega = fmaxf(1.f / (r4pow(r, 1.5f) + a), om);
      else
        omega = fmaxf((lc + a * hc) / (r * r + 2.f * r * (1.f + hc)), om);
      float T;
      if (r > rin && r < F.td_rout)
        T = T0 * r4pow(kc / b / r4pow(r, 3.f), 0.25f);
      else
        T = (float)((double)T0 / 1e5);
      s.rho = T;
      MetricF m = bl_metric_f(r, th, a);
      double den = (double)m.tt + 2.0 * (double)m.tp * (double)omega + (double)(m.pp * omega * omega);
      s.u[0] = sqrt(-1. / den);
      s.u[3] = (double)omega * s.u[0];
      disk_polvec(rr, cos(thh), k4, (double)a, asin(1.0), s.b);
    } break;
    case FL_SPHACC: {
      float gamma = (float)(5.0 / 3.0), alpha = 0.25f, ninf = 1.f, Tinf = (float).917e-9, us = (float).94;
      float u = (float)(1. / rr);
      float n = (float)((double)(ninf * alpha / 4.f / us) * pow((double)(2 * gamma * Tinf), -1.5) * pow(2.0 * (double)u, 1.5));
      int j = bracket(F.uv, 461, u);
      float w = (u - F.uv[j - 1]) / (F.uv[j] - F.uv[j - 1]);
      float T = (float)((1.0 - (double)w) * (double)F.tvals[j - 1] + (double)(w * F.tvals[j]));
      float B = (float)sqrt(8. * pi * (double)n * pcmp / 2. * pcc2 * (double)u);
      j = bracket(F.uvals, 498, u);
      w = (u - F.uvals[j - 1]) / (F.uvals[j] - F.uvals[j - 1]);
      float ur = (float)((1.0 - (double)w) * (double)F.vvals[j - 1] + (double)(w * F.vvals[j]));
      float g00 = -(1.f - 2.f * u);
      float grr = (float)(-1.0 / (double)g00);
      s.u[1] = -(double)ur;
      s.u[0] = sqrt((-(double)grr * s.u[1] * s.u[1] - 1) / (double)g00);
      double B2 = (double)(B * B);
      s.b[0] = sqrt(s.u[1] * s.u[1] * (double)grr * B2 /
                    (s.u[1] * s.u[1] * (double)g00 * (double)grr + s.u[0] * s.u[0] * (double)g00 * (double)g00));
      s.b[1] = -sqrt((double)((B * B) / grr) - s.b[0] * s.b[0] * (double)g00 / (double)grr);
      s.rho = n;
      s.p = T;
      s.bmag = B;
    } break;
    case FL_FFJET: {
      const int nx1 = F.nx1, nx2 = F.nx2;
      float theta = (float)thh;
      float zm = r4cos(theta);
      float x2 = r4acos(fabsf(zm));
      theta = x2;
      float zr = (float)rr;
      float x1 = r4log(zr);
      int lx1 = (int)((x1 - F.ax1[0]) / (F.ax1[nx1 - 1] - F.ax1[0]) * (float)(nx1 - 1)) + 1;
      int lx2 = (int)((x2 - F.ax2[0]) / (F.ax2[nx2 - 1] - F.ax2[0]) * (float)(nx2 - 1)) + 1;
      lx2 = (lx2 <= (nx2 - 1)) ? ((lx2 >= 1) ? lx2 : 1) : (nx2 - 1);
      int ux2 = lx2 + 1;
      lx1 = lx1 < 1 ? 1 : (lx1 > nx1 - 1 ? nx1 - 1 : lx1);  // guard: reference discards these through merge()
      int ux1 = lx1 + 1;
      float td = (theta - F.axth[lx2 - 1]) / (F.axth[ux2 - 1] - F.axth[lx2 - 1]);
      float rd = (zr - F.axr[lx1 - 1]) / (F.axr[ux1 - 1] - F.axr[lx1 - 1]);
      int x1l = lx1 - 1, x1u = ux1 - 1, x2l = (lx2 - 1) * nx1, x2u = (ux2 - 1) * nx1;  // r fastest
      Field10 v = blend4(load_cell(cells, x1l + x2l), load_cell(cells, x1l + x2u), load_cell(cells, x1u + x2l),
                         load_cell(cells, x1u + x2u), rd, td);
      bool in = x1 > F.ax1[0];
      s.rho = in ? v.rho : 0.f;
      s.p = in ? v.p : 1.f;
      finish_grid(s, v, in, 1.f, zr, (float)thh, a, 1.0);
    } break;
    case FL_HARM: {
      const int nx1 = F.nx1, nx2 = F.nx2;
      float theta = (float)thh, zr = (float)rr;
      float x1 = r4log(zr);
      float x2 = (float)harm_x2_of_theta((double)theta, (double)F.h);
      int lx1 = (int)floorf((x1 - F.ax1[0]) / (F.ax1[nx1 - 1] - F.ax1[0]) * (float)(nx1 - 1)) + 1;
      int lx2 = (int)floorf((x2 - F.ax2[0]) / (F.ax2[nx2 - 1] - F.ax2[0]) * (float)(nx2 - 1)) + 1;
      int ux2 = lx2 + 1;
      lx2 = (lx2 >= 1) ? lx2 : 1;
      ux2 = (ux2 <= nx2) ? ux2 : nx2;
      float dth = (ux2 != lx2) ? (F.axth[ux2 - 1] - F.axth[lx2 - 1]) : F.axth[0];
      float td = fabsf(theta - F.axth[lx2 - 1]) / dth;
      lx1 = lx1 < 1 ? 1 : (lx1 > nx1 - 1 ? nx1 - 1 : lx1);  // guard
      int ux1 = lx1 + 1;
      float rd = (zr - F.axr[lx1 - 1]) / (F.axr[ux1 - 1] - F.axr[lx1 - 1]);
      float rh = 1.f + sqrtf(1.f - a * a);
      if (F.axr[lx1 - 1] <= rh) rd = 1.f;
      int x2l = lx2 - 1, x2u = ux2 - 1, x1l = (lx1 - 1) * nx2, x1u = (ux1 - 1) * nx2;  // theta fastest
      // trilinear with time weight 0 (rttd=0.) reduces to the first time slice: y=w1*(1-0)+w2*0
      Field10 v = blend4(load_cell(cells, x1l + x2l), load_cell(cells, x1l + x2u), load_cell(cells, x1u + x2l),
                         load_cell(cells, x1u + x2u), rd, td);
#define OP(f) v.f = v.f * (1.f - 0.f) + v.f * 0.f;
      GRT_FOR_FIELDS(OP)
#undef OP
      bool in = x1 > F.ax1[0];
      s.rho = in ? v.rho : 0.f;
      s.p = in ? v.p : 1.f;
      finish_grid(s, v, in, 1.f, zr, (float)thh, a, 1.0);
    } break;
    case FL_HARM3D: {
      const int nx1 = F.nx1, nx2 = F.nx2, nx3 = F.nx3;
      float theta = (float)thh, zr = (float)rr, zpp = (float)phh;
      double sq1 = sqrt(1. - (double)a * (double)a);
      float zphi = (float)((double)zpp + (double)a / (2.0 * sq1) * log(((double)zr - 1. - sq1) / ((double)zr - 1. + sq1)));
      zphi = (float)fmod((double)zphi, 2. * pi);
      if (zphi < 0.f) zphi = (float)((double)zphi + 2. * pi);
      float x1 = r4log(zr), x3 = zphi;
      float x2 = (float)((double)theta / pi);
      int lx1 = (int)floorf((x1 - F.ax1[0]) / (F.ax1[nx1 - 1] - F.ax1[0]) * (float)(nx1 - 1)) + 1;
      int lx2 = (int)floorf((x2 - F.ax2[0]) / (F.ax2[nx2 - 1] - F.ax2[0]) * (float)(nx2 - 1)) + 1;
      int ux2 = lx2 + 1;
      int lx3 = (int)floorf((x3 - F.ax3[0]) / (F.ax3[nx3 - 1] - F.ax3[0]) * (float)(nx3 - 1)) + 1;
      int ux3 = lx3 + 1;
      lx2 = (lx2 <= (nx2 - 1)) ? ((lx2 >= 1) ? lx2 : 1) : (nx2 - 1);
      lx1 = (lx1 >= 1) ? lx1 : 1;
      int ux1 = lx1 + 1;
      ux2 = (ux2 >= 1) ? ux2 : 1;
      ux2 = (ux2 <= nx2) ? ux2 : nx2;
      float dth = (ux2 != lx2) ? (F.axth[ux2 - 1] - F.axth[lx2 - 1]) : (F.axth[0] * 2);
      int lx3s = lx3 < 1 ? 1 : (lx3 > nx3 ? nx3 : lx3);
      float minph = F.ax3[lx3s - 1];
      if (ux3 > nx3) ux3 = 1;
      if (lx3 < 1) {
        lx3 = nx3;
        minph = (float)((double)F.ax3[lx3 - 1] - 2. * pi);
      }
      float pd = (zphi - minph) / (F.ax3[1] - F.ax3[0]);
      float td = fabsf(theta - F.axth[lx2 - 1]) / dth;
      if (ux1 > nx1) ux1 = nx1;  // guard beyond the outer edge
      if (lx1 > nx1) lx1 = nx1;
      float rd = (zr - F.axr[lx1 - 1]) / (F.axr[ux1 - 1] - F.axr[lx1 - 1]);
      float rh = 1.f + sqrtf(1.f - a * a);
      double fac = 1.0;
      if (F.axr[lx1 - 1] <= fmaxf(F.axr[0], rh)) {
        rd = 1.f;
        fac = 1e-6;
      }
      size_t x3l = lx3 - 1, x3u = ux3 - 1, x2l = (size_t)(lx2 - 1) * nx3, x2u = (size_t)(ux2 - 1) * nx3;
      size_t x1l = (size_t)(lx1 - 1) * nx2 * nx3, x1u = (size_t)(ux1 - 1) * nx2 * nx3;
      // quadlininterp(vv, td=0, xd=pd, yd=rd, zd=td): theta, then r, then phi
      Field10 lo = blend4(load_cell(cells, x1l + x2l + x3l), load_cell(cells, x1l + x2u + x3l),
                          load_cell(cells, x1u + x2l + x3l), load_cell(cells, x1u + x2u + x3l), rd, td);
      Field10 hi = blend4(load_cell(cells, x1l + x2l + x3u), load_cell(cells, x1l + x2u + x3u),
                          load_cell(cells, x1u + x2l + x3u), load_cell(cells, x1u + x2u + x3u), rd, td);
      Field10 v;
#define OP(f) { float w1 = lo.f * (1.f - pd) + hi.f * pd; v.f = w1 * (1.f - 0.f) + w1 * 0.f; }
      GRT_FOR_FIELDS(OP)
#undef OP
      bool in = x1 > F.ax1[0];
      s.rho = (float)((double)v.rho * fac);
      s.p = (float)((double)v.p * fac);
      finish_grid(s, v, in, 0.f, zr, (float)thh, a, fac);
    } break;
    case FL_KORAL:
    case FL_KORAL3D:
    case FL_KORAL3D_DISK: {
      // koral_vals (fluid_model_koral.f90:84-286) / koral3d_vals (fluid_model_koral3d.f90:83-346): x2 fastest, then x1, then
      // x3; closed-form MKS3 maps; cells next to the poles (and, for KORAL3D_DISK, a polar cone) are not trusted
      const bool d3 = F.model != FL_KORAL;
      const int nx1 = F.nx1, nx2 = F.nx2, nx3 = d3 ? F.nx3 : 1;
      float theta = (float)thh, zr = (float)rr;
      float x1 = (float)log((double)zr - (double)F.kr0);
      float x2 = koral_x2_of_theta(theta, zr, F.kh, F.kaa, F.kbb, F.kpp);
      int lx1 = (int)floorf((x1 - F.ax1[0]) / (F.ax1[nx1 - 1] - F.ax1[0]) * (float)(nx1 - 1)) + 1;
      int ux1 = lx1 + 1;
      lx1 = (lx1 >= 1) ? lx1 : 1;
      ux1 = (ux1 <= nx1) ? ux1 : nx1;
      lx1 = lx1 > nx1 ? nx1 : lx1;  // guards (the reference reads out of bounds outside the radial range)
      ux1 = ux1 < 1 ? 1 : ux1;
      int lx2 = (int)floorf((x2 - F.ax2[0]) / (F.ax2[nx2 - 1] - F.ax2[0]) * (float)(nx2 - 1)) + 1;
      int ux2 = lx2 + 1;
      lx2 = (lx2 >= 1) ? lx2 : 1;
      ux2 = (ux2 <= nx2) ? ux2 : nx2;
      lx2 = lx2 > nx2 ? nx2 : lx2;
      ux2 = ux2 < 1 ? 1 : ux2;
      float pd = 0.f;
      int lx3 = 1, ux3 = 1;
      if (d3) {
        float zphi = (float)phh;
        zphi = (float)fmod((double)zphi, 2. * pi);
        if (zphi < 0.f) zphi = (float)((double)zphi + 2. * pi);
        if ((double)zphi > pi) zphi = (float)((double)zphi - 2. * pi);
        lx3 = (int)floorf((zphi - F.ax3[0]) / (F.ax3[nx3 - 1] - F.ax3[0]) * (float)(nx3 - 1)) + 1;
        ux3 = lx3 + 1;
        int lx3s = lx3 < 1 ? 1 : (lx3 > nx3 ? nx3 : lx3);
        float minph = F.ax3[lx3s - 1];
        ux3 = (ux3 <= nx3) ? ux3 : 1;
        if (lx3 < 1) {
          lx3 = nx3;
          minph = (float)((double)F.ax3[lx3 - 1] - 2. * pi);
        }
        lx3 = lx3 > nx3 ? nx3 : lx3;
        pd = (zphi - minph) / (F.ax3[1] - F.ax3[0]);
      }
      float rd = (zr - F.axr[lx1 - 1]) / (F.axr[ux1 - 1] - F.axr[lx1 - 1]);
      float rh = 1.f + sqrtf(1.f - a * a);
      double fac = 1.0;
      if (F.axr[lx1 - 1] <= rh || lx1 == 1) {
        rd = 1.f;
        if (d3) fac = 1e-3;
      }
      float thl = koral_theta_of_x2(F.ax2[lx2 - 1], F.axr[lx1 - 1], F.kh, F.kaa, F.kbb, F.kpp);
      float thu = koral_theta_of_x2(F.ax2[ux2 - 1], F.axr[lx1 - 1], F.kh, F.kaa, F.kbb, F.kpp);
      float thone = koral_theta_of_x2(F.ax2[0], F.axr[lx1 - 1], F.kh, F.kaa, F.kbb, F.kpp);
      float td = (ux2 != lx2) ? fabsf((theta - thl) / (thu - thl)) : fabsf((theta - thl) / thone);
      size_t x2l = lx2 - 1, x2u = ux2 - 1, x1l = (size_t)(lx1 - 1) * nx2, x1u = (size_t)(ux1 - 1) * nx2;
      size_t x3l = (size_t)(lx3 - 1) * nx2 * nx1, x3u = (size_t)(ux3 - 1) * nx2 * nx1;
      Field10 v;
      if (d3) {
        Field10 c0 = load_cell(cells, x1l + x2l + x3l), c1 = load_cell(cells, x1l + x2u + x3l),
                c2 = load_cell(cells, x1u + x2l + x3l), c3 = load_cell(cells, x1u + x2u + x3l),
                c4 = load_cell(cells, x1l + x2l + x3u), c5 = load_cell(cells, x1l + x2u + x3u),
                c6 = load_cell(cells, x1u + x2l + x3u), c7 = load_cell(cells, x1u + x2u + x3u);
        // quadlininterp(vv, td=0, xd=pd, yd=rd, zd=td) (interpolate.f90:203-232): theta, then r, then phi, then time
#define OP(f) { float i1 = c0.f * (1.f - td) + c1.f * td, i2 = c2.f * (1.f - td) + c3.f * td;                 float j1 = c4.f * (1.f - td) + c5.f * td, j2 = c6.f * (1.f - td) + c7.f * td;                 float m1 = i1 * (1.f - rd) + i2 * rd, m2 = j1 * (1.f - rd) + j2 * rd;                       float w1 = m1 * (1.f - pd) + m2 * pd; v.f = w1 * (1.f - 0.f) + w1 * 0.f; }
        GRT_FOR_FIELDS_BE(OP)
#undef OP
      } else {
        Field10 c0 = load_cell(cells, x1l + x2l), c1 = load_cell(cells, x1l + x2u), c2 = load_cell(cells, x1u + x2l),
                c3 = load_cell(cells, x1u + x2u);
#define OP(f) { float i1 = c0.f * (1.f - td) + c1.f * td, i2 = c2.f * (1.f - td) + c3.f * td;                 float w1 = i1 * (1.f - rd) + i2 * rd; v.f = w1 * (1.f - 0.f) + w1 * 0.f; }
        GRT_FOR_FIELDS_BE(OP)
#undef OP
      }
      bool in = x1 > F.ax1[0];
      bool trust = (x2 > F.x2mintrust) && (x2 < F.x2maxtrust);
      if (d3 && F.ktype == 1) trust = trust && (theta > 0.25f) && (theta < (float)(pi - 0.25));
      bool ok = in && trust;
      s.rho = (float)((double)(in ? v.rho : 0.f) * fac);
      s.p = (float)((double)(in ? v.p : 1.f) * fac);
      if (!trust) {
        s.rho = 0.f;
        s.p = 1.f;
      }
      s.be = (d3 && ok) ? v.be : 0.f;
      finish_grid(s, v, ok, 1.f, zr, (float)thh, a, 1.0);
    } break;
    case FL_IHARM: {
      // iharm_vals (fluid_model_iharm.f90:222-424): nearest-neighbour lookup -- rd = td = pd = 1 puts all the weight on
      // the upper corner of the cell in r, theta and phi; time weight 0 (nt = 1)
      const int nx1 = F.nx1, nx2 = F.nx2, nx3 = F.nx3;
      float theta = (float)thh, zr = (float)rr, zpp = (float)phh;
      double sq1 = sqrt(1. - (double)a * (double)a);
      float zphi = (float)((double)zpp + (double)a / (2.0 * sq1) * log(((double)zr - 1. - sq1) / ((double)zr - 1. + sq1)));
      zphi = (float)fmod((double)zphi, 2. * pi);
      if (zphi < 0.f) zphi = (float)((double)zphi + 2. * pi);
      float x1 = r4log(zr), x3 = zphi, x2;
      IharmGrid G{F.hslope, F.mks_smooth, F.poly_xt, F.poly_alpha, F.startx1};
      if (F.is_mmks == 1) {
        IharmThetaMmks fn{G, (double)x1, (double)theta};
        x2 = (float)zbrent_hd(fn, -1.0, 1.0, 1e-6);
      } else {
        IharmThetaMks fn{G, (double)theta};
        x2 = (float)zbrent_hd(fn, 0.0, 1.0, 1e-6);
      }
      int lx1 = (int)floorf((x1 - F.ax1[0]) / (F.ax1[nx1 - 1] - F.ax1[0]) * (float)(nx1 - 1)) + 1;
      int ux1 = lx1 + 1;
      int lx2 = (int)floorf((x2 - F.ax2[0]) / (F.ax2[nx2 - 1] - F.ax2[0]) * (float)(nx2 - 1)) + 1;
      int ux2 = lx2 + 1;
      int lx3 = (int)floorf((x3 - F.ax3[0]) / (F.ax3[nx3 - 1] - F.ax3[0]) * (float)(nx3 - 1)) + 1;
      int ux3 = lx3 + 1;
      lx2 = (lx2 >= 1) ? lx2 : 1;
      lx2 = (lx2 <= nx2) ? lx2 : nx2;
      ux2 = (ux2 >= 1) ? ux2 : 1;
      ux2 = (ux2 <= nx2) ? ux2 : nx2;
      if (ux3 > nx3) ux3 = 1;
      if (lx3 < 1) lx3 = nx3;
      // guards: the reference indexes uniqr(lx1) / the data arrays out of bounds outside the radial range
      int lx1s = lx1 < 1 ? 1 : (lx1 > nx1 ? nx1 : lx1), ux1s = ux1 < 1 ? 1 : (ux1 > nx1 ? nx1 : ux1);
      float rh = 1.f + sqrtf(1.f - a * a);
      double fac = (F.axr[lx1s - 1] <= rh) ? 1e-3 : 1.0;
      Field10 v = load_cell(cells, (size_t)(ux1s - 1) * nx2 * nx3 + (size_t)(ux2 - 1) * nx3 + (size_t)(ux3 - 1));
      bool in = x1 > F.ax1[0];
      s.rho = (float)((double)(in ? v.rho : 0.f) * fac);
      s.p = (float)((double)(in ? v.p : 1.f) * fac);
      s.kela = (F.eheat == 1) ? (float)((double)(in ? v.kela : 1.f) * fac) : 0.f;
      finish_grid(s, v, in, 0.f, zr, (float)thh, a, 1.0);
    } break;
    default:
      break;
  }
}

// ---- code units -> cgs (convert_fluid_vars_arr and the per-model converters) -----------------------
struct SourceDev {
  double nfac, bfac, mbh, mdot, p1, p2, gmax, gminval, jetalphaval, muval, sigcut;
  int stail, pad;  // stype = 'tail' (assign_source_params_type, fluid.f90:1614-1624)
};
struct CgsVars {
  double ncgs, ncgsnth, bcgs, tcgs, gmin;  // gmin: minimum Lorentz factor of the power law at this point (sp%gmin)
};

// calcgmin.f90:6-78: minimum Lorentz factor and non-thermal fraction of a power-law tail matched to a thermal distribution
// of dimensionless temperature theta, carrying the fraction eta of its energy.  The fit constants are default-real literals
// in the reference (single precision promoted to double).  (sic) its p = 3.5 branch assigns gmin_lin_sin_* twice and never
// assigns gmin_inv_sin_*: uninitialised there, 0 here.
GRT_HD void tail_gmin(double p, double theta, double eta, double &gmin, double &nfrac) {
  const double azero = 3.0 / sqrt(2.0), astwo = log(sqrt(2.0));
  double linc, invc;
  if (p == 3.5) {
    linc = (double)16.0797900684f + (double)-13.5593749125f * pow(eta, (double)0.276589155355f) +
           (double)0.121815691108f * sin(eta * (double)0.121815691108f + (double)0.121815691108f);
    invc = (double)0.722506578136f + (double)151.597731214f * pow(eta, (double)6.53997654139f);
  } else {
    linc = (double)21.38307186f + (double)-16.7811712f * pow(eta, (double)0.15128533f);
    invc = (double)0.74798712f + (double)0.62609462f * pow(eta, (double)0.81567379f) +
           (double).00638946501f * sin(eta * (double)-16.8034428f + (double)3.72208398f);
  }
  const double raw = theta * linc + invc;
  gmin = (raw >= 1.0) ? raw : 1.0;
  const double atheta = theta * azero * exp(astwo * tanh((double)0.52624783f * log(theta / (double)0.5668090982352612f)));
  nfrac = eta * atheta * ((p - 2.) / (p - 1.)) * pow(gmin, p - 2.);
}
// assign_source_params (fluid.f90:1640-1677), applied after every model's unit conversion (fluid.f90:582)
GRT_HDN void tail_source(const SourceDev &sp, CgsVars &o) {
  const double pcc = 2.99792458e10, pck = 1.38e-16, pcm = 9.10938188e-28;
  double g, x;
  tail_gmin(sp.p2, sp.muval * pck * o.tcgs / pcm / pcc / pcc, sp.jetalphaval, g, x);
  const bool in = g <= sp.gmax;
  o.gmin = in ? g : sp.gmax / 2.0;
  const double factor = in ? 1.0 : pow(sp.gmax / 2.0 / g, sp.p2 - 2.);
  o.ncgsnth = factor * ((x > 0.0) ? x * o.ncgs * pow(o.gmin, 1. - sp.p2) : 0.0);
}
GRT_HD void source_params(const SourceDev &sp, CgsVars &o) {
  o.gmin = sp.gminval;
  if (sp.stail) tail_source(sp, o);  // (out of line: rarely used, keeps the common path's registers)
}

GRT_HD void to_cgs(int model, const SourceDev &sp, float rho, float p, float bmag, float kela, float be, float gam,
                   CgsVars &o) {
  const double pi = GRT_PI;
  const double pcG = 6.67e-8, pcc = 2.99792458e10, pcmsun = 1.998e33, pcmp = 1.67262158e-24, pck = 1.38e-16;
  o.ncgs = o.ncgsnth = o.bcgs = o.tcgs = o.gmin = 0.0;
  switch (model) {
    case FL_SPHACC:
      o.ncgs = rho; o.bcgs = bmag; o.tcgs = p;
      break;
    case FL_FFJET:
      o.ncgsnth = (double)rho * sp.nfac; o.bcgs = (double)bmag * sp.bfac;
      break;
    case FL_THINDISK:
      o.tcgs = rho; o.ncgs = 1.;
      break;
    case FL_HARM:
    case FL_HARM3D: {
      double mdot = (model == FL_HARM) ? .003 : pcG * sp.mbh * pcmsun / (pcc * pcc * pcc);
      double lcgs = pcG * sp.mbh * pcmsun / (pcc * pcc), tcg = lcgs / pcc;
      double rhocgs = sp.mdot / mdot / (lcgs * lcgs * lcgs) * tcg * (double)rho;
      o.ncgs = rhocgs / pcmp;
      double pcgs = (double)p * rhocgs / (double)rho * (pcc * pcc);
      o.tcgs = pcgs / o.ncgs / pck;
      o.bcgs = (double)bmag * sqrt(rhocgs / (double)rho) * pcc;
      o.bcgs = o.bcgs * sqrt(4.0 * pi);
      double beta = (double)(p / (bmag * bmag)) / 0.5;
      double b2 = beta * beta;
      if (model == FL_HARM) {
        double rlow = sp.muval, rhigh = sp.muval / sp.gminval;
        double trat = ((double)bmag > 0.0) ? rhigh * b2 / (1.0 + b2) + rlow / (1.0 + b2) : rhigh;
        o.tcgs = o.tcgs * trat;
        o.ncgsnth = o.ncgs;
      } else {
        double rlow = 1.0 / sp.muval - 1.0, rhigh = sp.gminval * (1.0 / sp.muval - 1.0);
        double trat = ((double)bmag > 0.0) ? rhigh * b2 / (1.0 + b2) + rlow / (1.0 + b2) : rhigh;
        o.tcgs = o.tcgs / (1.0 + trat);
        double bmagrho = sq((double)bmag) / (double)rho;
        o.ncgsnth = (bmagrho > 1.0)
                        ? sp.jetalphaval * sq(o.bcgs) / 8.0 / pi / sp.gminval * (sp.p1 - 2.0) / (sp.p1 - 1.0) / 8.2e-7
                        : 0.0;
        double rc2 = o.ncgs * pcmp;
        double sigmacgs = (o.bcgs * o.bcgs) / (rc2 * 8.988e20 * 4.0 * pi);
        if (sigmacgs >= sp.sigcut) {
          o.ncgs = 0.0;
          o.tcgs = 1e9;
        }
      }
    } break;
    case FL_KORAL:
    case FL_KORAL3D:
    case FL_KORAL3D_DISK: {  // convert_fluidvars_koral (fluid.f90:1075-1162): koral quantities are already cgs
      double rhocgs = (double)rho;
      o.ncgs = rhocgs / pcmp;
      o.bcgs = (double)bmag * sqrt(4 * pi);
      o.bcgs = o.bcgs * sqrt(sp.nfac);
      o.ncgs = o.ncgs * sp.nfac;
      rhocgs = rhocgs * sp.nfac;
      if (sp.gminval >= 1.0) {  // charles_e(rho, p+Be, 2p+Be, bmag, 1, 1, gmin) (fluid.f90:845-868): p = T_e, Be = T_p
        float pp = p + be, uu = 2.f * p + be;
        double beta = 2.0 * (double)rho * pck * (double)pp / pcmp / (double)(bmag * bmag);
        double b2 = beta * beta;
        double trat = ((double)bmag > 0.0) ? sp.gminval * b2 / (1.0 + b2) + 1.0 / (1.0 + b2) : sp.gminval;
        o.tcgs = (double)uu / (2.0 + trat);
      } else if (sp.gminval < 0.0) {
        o.tcgs = (double)p;
      }
      double sigmacgs = (o.bcgs * o.bcgs) / (rhocgs * 8.988e20 * 4.0 * pi);
      if (sigmacgs >= sp.sigcut) {
        o.ncgs = 0.0;
        o.tcgs = 1e9;
      }
      if (model == FL_KORAL3D_DISK && be >= 0.05f) {  // zero out the jet (Bernoulli cut)
        o.tcgs = 10.;
        o.bcgs = 0.;
      }
      o.ncgsnth = 0.;
    } break;
    case FL_IHARM: {  // convert_fluidvars_iharm (fluid.f90:995-1026)
      double mdot = pcG * sp.mbh * pcmsun / (pcc * pcc * pcc);
      double lcgs = pcG * sp.mbh * pcmsun / (pcc * pcc), tcg = lcgs / pcc;
      double rhocgs = sp.mdot / mdot / (lcgs * lcgs * lcgs) * tcg * (double)rho;
      o.ncgs = rhocgs / pcmp;
      double pcgs = (double)p * rhocgs / (double)rho * (pcc * pcc);
      o.tcgs = pcgs / o.ncgs / pck;
      o.bcgs = (double)bmag * sqrt(rhocgs / (double)rho) * pcc;
      o.bcgs = o.bcgs * sqrt(4.0 * pi);
      double rc2 = o.ncgs * pcmp;
      if (sp.gminval >= 1.0) {  // monika_e with p = u (gamma-1) (f%p holds the internal energy)
        float pg = p * (gam - 1.f);
        double beta = (double)(pg / (bmag * bmag)) / 0.5;
        double b2 = beta * beta;
        double rlow = 1.0 / sp.muval - 1.0, rhigh = sp.gminval * (1.0 / sp.muval - 1.0);
        double trat = ((double)bmag > 0.0) ? rhigh * b2 / (1.0 + b2) + rlow / (1.0 + b2) : rhigh;
        o.tcgs = 2.0 * o.tcgs / 3.0 / (2.0 + trat);
      } else if (sp.gminval < 0.0) {
        if (sp.gminval == -1.0) {  // ressler_e: electron entropy -> T_e
          const double pcm = 9.10938188e-28;
          double thetae = pcmp / pcm * (double)kela * pow((double)rho, 4.0 / 3.0 - 1.);
          o.tcgs = thetae * pcm * (pcc * pcc) / pck;
        }
      }
      double bmagrho = sq((double)bmag) / (double)rho;
      o.ncgsnth = (bmagrho > 1.0)
                      ? sp.jetalphaval * sq(o.bcgs) / 8.0 / pi / sp.gminval * (sp.p1 - 2.0) / (sp.p1 - 1.0) / 8.2e-7
                      : 0.0;
      double sigmacgs = (o.bcgs * o.bcgs) / (rc2 * 8.988e20 * 4.0 * pi);
      if (sigmacgs >= sp.sigcut) {
        o.ncgs = 0.0;
        o.tcgs = 1e9;
      }
    } break;
    default:
      break;
  }
  source_params(sp, o);
}

}  // namespace grt
