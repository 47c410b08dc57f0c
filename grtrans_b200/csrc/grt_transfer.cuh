// grtrans_b200 -- per-(ray, mdot, frequency) polarised transfer: one thread marches one ray at one
// frequency, evaluating emissivities on the fly from the per-point records written by the
// fluid/geometry kernel.
//
// Replaces grtrans_driver.f90:157-325 (frequency loop: fac, rotate_emis, invariant_emis, scaling,
// optical depth, integrate, save pixel), emis.f90:786-847 and radtrans_integrate.f90:
//   DELO    :746-842  -- same E,F,G / thin-branch algebra per step; the step is applied as
//                        I_new = M^-1 [ (F-G) S_k + G S_k+1 + (E - G K'_k+1) I_prev ] (one 4x4 cofactor
//                        inverse + two mat-vecs) instead of forming P and Q = M^-1 (..) separately
//   FORMAL  :844-876 + calc_O :615-683 -- O(dx) applied to a vector without building M1..M4
//   quadrature :878-882
//   LSODA   :200-307 + radtrans_aux :453-465 -- same tau cut, zero-emissivity trim and (as written)
//           neighbour-index coefficient model; ODEPACK is replaced by an embedded Cash-Karp 5(4)
//           Runge-Kutta with the reference tolerances (rtol 1e-6, atol 1e-8 on fac-normalised
//           intensities, hmax 0.1) that never steps across a tabulated node.
// Because the transfer equation is linear and I_0 = 0, the reference's normalisation j/fac ... I*fac
// cancels for DELO / FORMAL / quadrature; fac is still accumulated because "fac <= 0 (or NaN) => zero
// pixel" is reference behaviour (grtrans_driver.f90:195,314-317).
#pragma once
#include "grt_emis.cuh"
#include "grt_fluid.cuh"

namespace grt {

#define GRT_NREC 16  // doubles per point record (128 B)

struct TransferParams {
  int etype, eneq, rneq, iflag, model, npts_is_one;
  int coefindx[7];
  double LBH, gmin, gmax, p;  // emis_params pass through float32 in the reference
  double max_tau, hmax, ortol, oatol, thin;
  int maxsteps;
  EmisTables tab;
};

struct Coef {
  double j[4], K[7];
};

// record -> invariant, LBH-scaled coefficients at observed frequency nu_obs.  jI_raw returns the
// comoving un-boosted j_I used for fac.
GRT_HD void node_coef(const TransferParams &P, const double *rec, double nu_obs, Coef &c, double &jI_raw) {
  const double g = rec[1];
  double nu = nu_obs / g;
  double e[11];
  switch (P.etype) {
    case EM_POLSYNCHTH: {
      ThermalPoint T;
      T.n = rec[4]; T.b = rec[5]; T.t = rec[6]; T.sinth = rec[7]; T.costh = rec[8]; T.k0 = rec[9]; T.k1ok2 = rec[10];
      T.ik2 = rec[11]; T.step = rec[12];
      T.tanth = rec[13];
      T.thetae = pc::k * T.t / pc::m / pc::c / pc::c + 1e-10;
      T.nucfac = 3.0 * pc::e * T.b * T.sinth / 4.0 / GRT_PI / pc::m / pc::c * T.thetae * T.thetae + 1.0;
      polsynchth_eval(T, nu, e);
    } break;
    case EM_POLSYNCHPL:
      synchpl_eval<true>(P.tab, nu, rec[4], rec[5], rec[6], P.p, P.gmin, P.gmax, e);
      break;
    case EM_SYNCHPL:
      synchpl_eval<false>(P.tab, nu, rec[4], rec[5], rec[6], P.p, P.gmin, P.gmax, e);
      break;
    case EM_SYNCHTHAV:
      synchemis_eval(nu, rec[4], rec[5], rec[6], e);
      break;
    case EM_BB:
      for (int i = 0; i < 11; i++) e[i] = 0.0;
      e[0] = planck_bnu(rec[6], nu);
      break;
    case EM_BBPOL: {
      for (int i = 0; i < 11; i++) e[i] = 0.0;
      const double fcol = 1.8;
      e[0] = pow(fcol, -4.0) * planck_bnu(rec[6] * fcol, nu);
      e[0] = e[0] * rec[7];
      e[1] = e[0] * rec[8];
    } break;
    default:
      for (int i = 0; i < 11; i++) e[i] = 0.0;
  }
  jI_raw = e[0];
  for (int i = 0; i < 4; i++) c.j[i] = 0.0;
  for (int i = 0; i < 7; i++) c.K[i] = 0.0;
  if (P.eneq == 4) {
    for (int i = 0; i < 4; i++) c.j[i] = e[i];
    for (int i = 0; i < 7; i++)
      if (P.coefindx[i] == 1) c.K[i] = e[4 + i];
    const double s2 = rec[2], c2 = rec[3];
    double jq = c.j[1], ju = c.j[2], aq = c.K[1], au = c.K[2], rq = c.K[4], ru = c.K[5];
    c.j[1] = c2 * jq - s2 * ju;
    c.j[2] = s2 * jq + c2 * ju;
    c.K[1] = c2 * aq - s2 * au;
    c.K[2] = s2 * aq + c2 * au;
    c.K[4] = c2 * rq - s2 * ru;
    c.K[5] = s2 * rq + c2 * ru;
  } else {
    c.j[0] = e[0];
    c.K[0] = e[4];
  }
  if (!P.npts_is_one) {
    for (int i = 0; i < 4; i++) c.j[i] = c.j[i] * g * g;
    for (int i = 0; i < 7; i++) c.K[i] = c.K[i] / g * P.LBH;
  } else {
    double g3 = g * g * g;
    for (int i = 0; i < 4; i++) c.j[i] = c.j[i] * g3;
  }
  if (P.rneq == 1) {  // Stokes-I-only transfer of a polarised model (A21): keep j_I, alpha_I
    c.j[1] = c.j[2] = c.j[3] = 0.0;
    for (int i = 1; i < 7; i++) c.K[i] = 0.0;
  }
}

// ---- 4x4 helpers ------------------------------------------------------------------------------------
GRT_HD void kmat_vec(const double *K, const double *x, double *y) {  // opacity_matrix(K) * x
  y[0] = K[0] * x[0] + K[1] * x[1] + K[2] * x[2] + K[3] * x[3];
  y[1] = K[1] * x[0] + K[0] * x[1] + K[6] * x[2] - K[5] * x[3];
  y[2] = K[2] * x[0] - K[6] * x[1] + K[0] * x[2] + K[4] * x[3];
  y[3] = K[3] * x[0] + K[5] * x[1] - K[4] * x[2] + K[0] * x[3];
}

// y = (c1*Id + c2*Kmat)^-1 x through the cofactor inverse of radtrans_integrate.f90:685-733
GRT_HD void solve_delo(double c1, double c2, const double *K, const double *x, double *y) {
  const double d = c1 + c2 * K[0];
  const double a12 = c2 * K[1], a13 = c2 * K[2], a14 = c2 * K[3];
  const double a23 = c2 * K[6], a24 = -c2 * K[5], a34 = c2 * K[4];
  const double a11 = d, a22 = d, a33 = d, a44 = d, a21 = a12, a31 = a13, a41 = a14, a32 = -a23, a42 = -a24, a43 = -a34;
  double b11 = a22 * a33 * a44 + a23 * a34 * a42 + a24 * a32 * a43 - a22 * a34 * a43 - a23 * a32 * a44 - a24 * a33 * a42;
  double b12 = a12 * a34 * a43 + a13 * a32 * a44 + a14 * a33 * a42 - a12 * a33 * a44 - a13 * a34 * a42 - a14 * a32 * a43;
  double b13 = a12 * a23 * a44 + a13 * a24 * a42 + a14 * a22 * a43 - a12 * a24 * a43 - a13 * a22 * a44 - a14 * a23 * a42;
  double b14 = a12 * a24 * a33 + a13 * a22 * a34 + a14 * a23 * a32 - a12 * a23 * a34 - a13 * a24 * a32 - a14 * a22 * a33;
  double b21 = a21 * a34 * a43 + a23 * a31 * a44 + a24 * a33 * a41 - a21 * a33 * a44 - a23 * a34 * a41 - a24 * a31 * a43;
  double b22 = a11 * a33 * a44 + a13 * a34 * a41 + a14 * a31 * a43 - a11 * a34 * a43 - a13 * a31 * a44 - a14 * a33 * a41;
  double b23 = a11 * a24 * a43 + a13 * a21 * a44 + a14 * a23 * a41 - a11 * a23 * a44 - a13 * a24 * a41 - a14 * a21 * a43;
  double b24 = a11 * a23 * a34 + a13 * a24 * a31 + a14 * a21 * a33 - a11 * a24 * a33 - a13 * a21 * a34 - a14 * a23 * a31;
  double b31 = a21 * a32 * a44 + a22 * a34 * a41 + a24 * a31 * a42 - a21 * a34 * a42 - a22 * a31 * a44 - a24 * a32 * a41;
  double b32 = a11 * a34 * a42 + a12 * a31 * a44 + a14 * a32 * a41 - a11 * a32 * a44 - a12 * a34 * a41 - a14 * a31 * a42;
  double b33 = a11 * a22 * a44 + a12 * a24 * a41 + a14 * a21 * a42 - a11 * a24 * a42 - a12 * a21 * a44 - a14 * a22 * a41;
  double b34 = a11 * a24 * a32 + a12 * a21 * a34 + a14 * a22 * a31 - a11 * a22 * a34 - a12 * a24 * a31 - a14 * a21 * a32;
  double b41 = a21 * a33 * a42 + a22 * a31 * a43 + a23 * a32 * a41 - a21 * a32 * a43 - a22 * a33 * a41 - a23 * a31 * a42;
  double b42 = a11 * a32 * a43 + a12 * a33 * a41 + a13 * a31 * a42 - a11 * a33 * a42 - a12 * a31 * a43 - a13 * a32 * a41;
  double b43 = a11 * a23 * a42 + a12 * a21 * a43 + a13 * a22 * a41 - a11 * a22 * a43 - a12 * a23 * a41 - a13 * a21 * a42;
  double b44 = a11 * a22 * a33 + a12 * a23 * a31 + a13 * a21 * a32 - a11 * a23 * a32 - a12 * a21 * a33 - a13 * a22 * a31;
  double det = a11 * b11 + a21 * b12 + a31 * b13 + a41 * b14;
  double id = 1.0 / det;
  y[0] = (b11 * x[0] + b12 * x[1] + b13 * x[2] + b14 * x[3]) * id;
  y[1] = (b21 * x[0] + b22 * x[1] + b23 * x[2] + b24 * x[3]) * id;
  y[2] = (b31 * x[0] + b32 * x[1] + b33 * x[2] + b34 * x[3]) * id;
  y[3] = (b41 * x[0] + b42 * x[1] + b43 * x[2] + b44 * x[3]) * id;
}

// one DELO step from node k+1 (c1) to node k (c0): returns I at node k
GRT_HD void delo_step(const Coef &c0, const Coef &c1, double dx, double delta, double thin, const double *iprev,
                      double *inew) {
  double rhs[4], t[4];
  if (delta > thin) {
    double E = exp(-delta), F = 1.0 - E, G = (1.0 - (1.0 + delta) * E) / delta;
    // (E*Id - G*Kp1) iprev, Kp1 = K1/aI1 - Id
    kmat_vec(c1.K, iprev, t);
    for (int i = 0; i < 4; i++)
      rhs[i] = (F - G) * (c0.j[i] / c0.K[0]) + G * (c1.j[i] / c1.K[0]) + (E * iprev[i] - G * (t[i] / c1.K[0] - iprev[i]));
    // matrix = Id + (F-G)*(K0/aI0 - Id) = (1-(F-G)) Id + ((F-G)/aI0) K0
    solve_delo(1.0 - (F - G), (F - G) / c0.K[0], c0.K, rhs, inew);
  } else {
    double ki = c0.K[0];
    kmat_vec(c1.K, iprev, t);
    double qa = 1.0 - 0.5 * dx * ki + 1.0 / 6.0 * dx * dx * ki * ki, qb = 0.5 * dx - 1.0 / 3.0 * dx * dx;
    for (int i = 0; i < 4; i++)
      rhs[i] = (0.5 * dx * c0.j[i] - 1.0 / 6.0 * dx * dx * ki * c0.j[i]) + (0.5 * c1.j[i] * dx - 1.0 / 3.0 * dx * dx * ki * c1.j[i]) +
               (qa * iprev[i] - qb * t[i]);
    solve_delo(1.0 - delta / 2.0 + delta * delta / 6.0, 0.5 * dx - 1.0 / 6.0 * dx * dx * ki, c0.K, rhs, inew);
  }
}

// inew = O(dx; a, rho) * v  (Landi Degl'Innocenti & Landi Degl'Innocenti 1985 evolution operator)
GRT_HD void formal_apply(const double *K, double dx, const double *v, double *inew) {
  const double aq = K[1], au = K[2], av = K[3], rq = K[4], ru = K[5], rv = K[6];
  double onopol = exp(-K[0] * dx);
  double a2 = aq * aq + au * au + av * av, p2 = rq * rq + ru * ru + rv * rv;
  if (a2 == 0.0 && p2 == 0.0) {
    for (int i = 0; i < 4; i++) inew[i] = onopol * v[i];
    return;
  }
  double ap = aq * rq + au * ru + av * rv;
  double d = a2 - p2;
  double root = sqrt(d * d / 4.0 + ap * ap);
  double lam1 = sqrt(root + d / 2.0), lam2 = sqrt(root - d / 2.0);
  double theta = lam1 * lam1 + lam2 * lam2;
  double sig = signbit(ap) ? -1.0 : 1.0;
  // M2 v, M3 v, M4 v from the structured entries (calc_O :633-668)
  double pa = lam2 * aq - sig * lam1 * rq, pb = lam2 * au - sig * lam1 * ru, pcv = lam2 * av - sig * lam1 * rv;
  double qa = sig * lam1 * aq + lam2 * rq, qb = sig * lam1 * au + lam2 * ru, qc = sig * lam1 * av + lam2 * rv;
  double m2[4] = {pa * v[1] + pb * v[2] + pcv * v[3], pa * v[0] + qc * v[2] - qb * v[3], pb * v[0] - qc * v[1] + qa * v[3],
                  pcv * v[0] + qb * v[1] - qa * v[2]};
  double ra = lam1 * aq + sig * lam2 * rq, rb = lam1 * au + sig * lam2 * ru, rcv = lam1 * av + sig * lam2 * rv;
  double sa = sig * lam2 * aq - lam1 * rq, sb = sig * lam2 * au - lam1 * ru, sc = sig * lam2 * av - lam1 * rv;
  double m3[4] = {ra * v[1] + rb * v[2] + rcv * v[3], ra * v[0] - sc * v[2] + sb * v[3], rb * v[0] + sc * v[1] - sa * v[3],
                  rcv * v[0] - sb * v[1] + sa * v[2]};
  double hp = (a2 + p2) / 2.0;
  double x1 = au * rv - av * ru, x2 = av * rq - aq * rv, x3 = aq * ru - au * rq;
  double m4[4] = {hp * v[0] - x1 * v[1] - x2 * v[2] - x3 * v[3],
                  x1 * v[0] + (aq * aq + rq * rq - hp) * v[1] + (aq * au + rq * ru) * v[2] + (av * aq + rv * rq) * v[3],
                  x2 * v[0] + (aq * au + rq * ru) * v[1] + (au * au + ru * ru - hp) * v[2] + (au * av + ru * rv) * v[3],
                  x3 * v[0] + (av * aq + rv * rq) * v[1] + (au * av + ru * rv) * v[2] + (av * av + rv * rv - hp) * v[3]};
  double ch = cosh(lam1 * dx), cs = cos(lam2 * dx), sh = sinh(lam1 * dx), sn = sin(lam2 * dx);
  for (int i = 0; i < 4; i++)
    inew[i] = onopol * (0.5 * (ch + cs) * v[i] - sn * (m2[i] / theta) - sh * (m3[i] / theta) + 0.5 * (ch - cs) * (2.0 / theta * m4[i]));
}

GRT_HD void rt_rhs(int neq, const double *j, const double *K, const double *I, double *d) {
  if (neq == 4) {
    double t[4];
    kmat_vec(K, I, t);
    for (int i = 0; i < 4; i++) d[i] = j[i] - t[i];
  } else {
    d[0] = j[0] - K[0] * I[0];
  }
}

// Cash-Karp 5(4) across one tabulated interval [la, lb] with coefficients linear between the node
// values (ca at w=0, cb at w=1).  Returns false on step overflow / NaN.
GRT_HD bool ck_interval(int neq, const Coef &ca, const Coef &cb, double la, double lb, double *y, double &h,
                        const TransferParams &P) {
  const double b21 = 0.2, b31 = 3.0 / 40.0, b32 = 9.0 / 40.0, b41 = 0.3, b42 = -0.9, b43 = 1.2, b51 = -11.0 / 54.0,
               b52 = 2.5, b53 = -70.0 / 27.0, b54 = 35.0 / 27.0, b61 = 1631.0 / 55296.0, b62 = 175.0 / 512.0,
               b63 = 575.0 / 13824.0, b64 = 44275.0 / 110592.0, b65 = 253.0 / 4096.0;
  const double c1 = 37.0 / 378.0, c3 = 250.0 / 621.0, c4 = 125.0 / 594.0, c6 = 512.0 / 1771.0;
  const double d1 = c1 - 2825.0 / 27648.0, d3 = c3 - 18575.0 / 48384.0, d4 = c4 - 13525.0 / 55296.0, d5 = -277.0 / 14336.0,
               d6 = c6 - 0.25;
  const double a2 = 0.2, a3 = 0.3, a4 = 0.6, a5 = 1.0, a6 = 0.875;
  const double span = lb - la;
  double t = la;
  int steps = 0;
  auto coef = [&](double tt, double *j, double *K) {
    double w = (tt - la) / span;
    for (int i = 0; i < neq; i++) j[i] = (1.0 - w) * ca.j[i] + w * cb.j[i];
    int nk = (neq == 4) ? 7 : 1;
    for (int i = 0; i < nk; i++) K[i] = (1.0 - w) * ca.K[i] + w * cb.K[i];
  };
  while (t < lb) {
    if (++steps > P.maxsteps) return false;
    double hh = dmin(dmin(h, P.hmax), lb - t);
    bool last = hh >= lb - t;
    double j[4], K[7], k1[4], k2[4], k3[4], k4[4], k5[4], k6[4], yt[4], yn[4];
    coef(t, j, K);
    rt_rhs(neq, j, K, y, k1);
    for (int i = 0; i < neq; i++) yt[i] = y[i] + hh * b21 * k1[i];
    coef(t + a2 * hh, j, K);
    rt_rhs(neq, j, K, yt, k2);
    for (int i = 0; i < neq; i++) yt[i] = y[i] + hh * (b31 * k1[i] + b32 * k2[i]);
    coef(t + a3 * hh, j, K);
    rt_rhs(neq, j, K, yt, k3);
    for (int i = 0; i < neq; i++) yt[i] = y[i] + hh * (b41 * k1[i] + b42 * k2[i] + b43 * k3[i]);
    coef(t + a4 * hh, j, K);
    rt_rhs(neq, j, K, yt, k4);
    for (int i = 0; i < neq; i++) yt[i] = y[i] + hh * (b51 * k1[i] + b52 * k2[i] + b53 * k3[i] + b54 * k4[i]);
    coef(t + a5 * hh, j, K);
    rt_rhs(neq, j, K, yt, k5);
    for (int i = 0; i < neq; i++)
      yt[i] = y[i] + hh * (b61 * k1[i] + b62 * k2[i] + b63 * k3[i] + b64 * k4[i] + b65 * k5[i]);
    coef(t + a6 * hh, j, K);
    rt_rhs(neq, j, K, yt, k6);
    double err = 0.0;
    for (int i = 0; i < neq; i++) {
      yn[i] = y[i] + hh * (c1 * k1[i] + c3 * k3[i] + c4 * k4[i] + c6 * k6[i]);
      double ei = hh * (d1 * k1[i] + d3 * k3[i] + d4 * k4[i] + d5 * k5[i] + d6 * k6[i]);
      double scv = P.oatol + P.ortol * dmax(fabs(y[i]), fabs(yn[i]));
      err += (ei / scv) * (ei / scv);
    }
    err = sqrt(err / neq);
    if (!(err <= 1.0)) {
      if (!(err == err)) return false;
      h = hh * dmax(0.2, 0.9 * pow(err, -0.25));
      continue;
    }
    for (int i = 0; i < neq; i++) y[i] = yn[i];
    t = last ? lb : t + hh;
    double fac = (err > 0.0) ? dmin(5.0, dmax(0.2, 0.9 * pow(err, -0.2))) : 5.0;
    if (!last || fac < 1.0) h = hh * fac;
  }
  return true;
}

// ---- the per-(ray, frequency) march ------------------------------------------------------------------
// rec: [npts][GRT_NREC] records of this ray.  out: nvals Stokes values in cgs.
GRT_HDN void ray_transfer(const TransferParams &P, const double *rec, int npts, double nu_obs, double *out) {
  const int neq = P.rneq;
  for (int i = 0; i < neq; i++) out[i] = 0.0;
  if (npts <= 0) return;
  double jraw;
  if (npts == 1) {
    Coef c;
    node_coef(P, rec, nu_obs, c, jraw);
    double fac = jraw / 1.0;
    if (fac / 1e16 > 0.0)
      for (int i = 0; i < neq; i++) out[i] = c.j[i];
    return;
  }
  const double lam0 = rec[0];
  double fac = 0.0;
  if (P.iflag != 0 && neq == 4) {
    // DELO / FORMAL: march from the deepest point to the camera
    double I[4] = {0, 0, 0, 0}, In[4];
    Coef c1, c0;
    node_coef(P, rec + (size_t)(npts - 1) * GRT_NREC, nu_obs, c1, jraw);
    fac += jraw * (lam0 - rec[(size_t)(npts - 1) * GRT_NREC]);
    for (int k = npts - 2; k >= 0; k--) {
      const double *rk = rec + (size_t)k * GRT_NREC;
      node_coef(P, rk, nu_obs, c0, jraw);
      fac += jraw * (lam0 - rk[0]);
      double dx = rk[0] - rk[GRT_NREC];
      if (P.iflag == 1) {
        double delta = dx * 0.5 * (fabs(c0.K[0]) + fabs(c1.K[0]));
        delo_step(c0, c1, dx, delta, P.thin, I, In);
      } else {
        double v[4];
        for (int i = 0; i < 4; i++) v[i] = c0.j[i] * dx + I[i];
        formal_apply(c0.K, dx, v, In);
      }
      for (int i = 0; i < 4; i++) I[i] = In[i];
      c1 = c0;
    }
    if (fac / 1e16 > 0.0)
      for (int i = 0; i < 4; i++) out[i] = I[i] * P.LBH;
    return;
  }
  if (P.iflag != 0) {
    // neq == 1 with delo/formal: quadrature  I = -tsum(s, j exp(-tau)), tau from the camera inwards
    Coef c;
    node_coef(P, rec, nu_obs, c, jraw);
    double tau = 0.0, acc = 0.0, yprev = c.j[0], kprev = fabs(c.K[0]);
    for (int k = 1; k < npts; k++) {
      const double *rk = rec + (size_t)k * GRT_NREC;
      node_coef(P, rk, nu_obs, c, jraw);
      fac += jraw * (lam0 - rk[0]);
      double ds = rk[0] - rk[-GRT_NREC];
      tau -= ds * (kprev + fabs(c.K[0])) / 2.0;
      double y = c.j[0] * exp(-tau);
      acc += ds * (yprev + y) / 2.0;
      yprev = y;
      kprev = fabs(c.K[0]);
    }
    if (fac / 1e16 > 0.0) out[0] = -acc * P.LBH;
    return;
  }
  // ---- LSODA replacement: pass 1 (fac, tau cut, zero trims), pass 2 (adaptive RK deep -> camera)
  int lamdex = npts, i1 = 1, i2;
  {
    Coef c;
    double tau = 0.0, kprev = 0.0;
    int firstnz = 0, lastnz_cut = 0;
    bool cut = false;
    for (int k = 0; k < npts; k++) {
      const double *rk = rec + (size_t)k * GRT_NREC;
      node_coef(P, rk, nu_obs, c, jraw);
      fac += jraw * (lam0 - rk[0]);
      if (k > 0) tau -= (rk[0] - rk[-GRT_NREC]) * (kprev + fabs(c.K[0])) / 2.0;
      kprev = fabs(c.K[0]);
      if (!cut && tau > P.max_tau) {  // locate(tau,MAX_TAU)+1 == first index with tau > MAX_TAU (1-based k+1)
        lamdex = k + 1;
        cut = true;
      }
      bool nz = c.j[0] != 0.0;
      if (nz && firstnz == 0) firstnz = k + 1;
      if (nz && (!cut || k + 1 <= lamdex)) lastnz_cut = k + 1;
    }
    // first/last non-zero emissivity (radtrans_integrate.f90:226-249)
    i1 = (firstnz > 1) ? firstnz : 1;
    i2 = lastnz_cut;  // equals lamdex when j(lamdex) != 0
  }
  if (!(fac / 1e16 > 0.0) || i2 <= i1) return;
  {
    double y[4] = {0, 0, 0, 0};
    double h = P.hmax;
    Coef cb, ca, ctmp;  // interval [s0(indx), s0(indx-1)] uses nodes (indx, indx+1) as written
    auto norm = [&](Coef &c) {
      for (int i = 0; i < 4; i++) c.j[i] /= fac;
    };
    int indx = i2;
    node_coef(P, rec + (size_t)(indx - 1) * GRT_NREC, nu_obs, ca, jraw);
    norm(ca);
    int up = (indx + 1 <= npts) ? indx + 1 : npts;
    node_coef(P, rec + (size_t)(up - 1) * GRT_NREC, nu_obs, cb, jraw);
    norm(cb);
    bool ok = true;
    for (; indx >= i1 + 1; indx--) {
      double la = rec[(size_t)(indx - 1) * GRT_NREC], lb = rec[(size_t)(indx - 2) * GRT_NREC];
      if (lb > la) ok = ck_interval(neq, ca, cb, la, lb, y, h, P);
      if (!ok) break;
      // slide: next interval indx-1 uses nodes (indx-1, indx)
      cb = ca;
      if (indx - 1 >= 1) {
        node_coef(P, rec + (size_t)(indx - 2) * GRT_NREC, nu_obs, ctmp, jraw);
        norm(ctmp);
        ca = ctmp;
      }
    }
    if (ok)
      for (int i = 0; i < neq; i++) out[i] = y[i] * fac * P.LBH;
  }
}

}  // namespace grt
