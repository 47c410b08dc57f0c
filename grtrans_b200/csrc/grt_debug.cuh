// grtrans_b200 -- debug = 1: the per-point internals of one ray that the reference dumps to `geodebug.out`
// (grtrans_driver.f90:86-110, 341-427; read back by read_geodebug_file.py).  The reference writes the file from inside
// its pixel loop, for the first (mdot, frequency) of every pixel, each pixel overwriting the last: what is left is the ray
// of the last pixel of the run.  Here that ray is re-run on the device (one-ray session) with the stage dump switched on,
// and the two bodies below add what the stage dump does not hold: the tetrad components of comoving_ortho_debug
// (kerr.f90:562-584), the two independent checks calc_kb_ang (:220-252) and CALCG (:181-218), the transfer coefficients
// as the integrator sees them (rotated, invariant, j / fac, K LBH), the optical depth and the intensity history.
#pragma once
#include "grt_pipeline.cuh"

namespace grt {

// rows of the debug buffer, [row][maxpts]
enum {
  GD_J = 0,        // 4: e%j(:, 1:4) = j g^2 (rotated) / fac
  GD_K = 4,        // 7: e%K(:, 1:7) = K LBH / g (rotated)
  GD_TAU = 11,     // optical depth from the camera
  GD_I = 12,       // 4: r%I(1:4, :) / fac / LBH
  GD_KBANG = 16,   // calc_kb_ang
  GD_CALCG = 17,   // CALCG
  GD_AAHAT = 18,   // 3
  GD_AA = 21,      // 4: aat, aar, aath, aaph
  GD_KHAT = 25,    // 4
  GD_BHAT = 29,    // 4
  GD_S2XI = 33,    // s2xi, c2xi of the second comoving_ortho call
  GD_C2XI = 34,
  GD_TPR = 35,     // g%tprarr, g%tpmarr
  GD_TPM = 36,
  GD_NROWS = 37
};

// per point p < npts of ray 0 of the one-ray tile: everything that does not depend on the frequency
GRT_HD void geodebug_point(const RayConst &r0, const RayMeta &m, int maxpts, PointArrays pa, StageDump sd, int p, double *rows) {
  const double r = sd.r[p], th = sd.th[p];
  const double k4[4] = {sd.kt[p], sd.kr[p], sd.kth[p], sd.kph[p]};
  const double *u = sd.u + 4 * (size_t)p, *b = sd.b + 4 * (size_t)p;
  OrthoOut oo;
  OrthoDebug od;
  comoving_ortho(r, th, r0.a, r0.alpha, r0.beta, r0.mu0, u, b, k4, oo, &od);
  const size_t src = (m.npts != 1) ? (size_t)(m.i1 - 1) + p : 0;
  const int tpm = pa.tpm[src], tpr = pa.tpr[src];
  auto put = [&](int row, double v) { rows[(size_t)row * maxpts + p] = v; };
  put(GD_KBANG, kb_angle(r, th, r0.a, k4, b, u));
  put(GD_CALCG, calcg_lnrf(r, cos(th), r0.q2, r0.l, r0.a, tpm, tpr, r0.su, r0.sm, u[1] / u[0], u[2] / u[0], u[3] / u[0]));
  for (int i = 0; i < 3; i++) put(GD_AAHAT + i, od.aahat[i]);
  for (int i = 0; i < 4; i++) {
    put(GD_AA + i, od.aa[i]);
    put(GD_KHAT + i, od.khat[i]);
    put(GD_BHAT + i, od.bhat[i]);
  }
  put(GD_S2XI, oo.s2xi);
  put(GD_C2XI, oo.c2xi);
  put(GD_TPR, (double)tpr);
  put(GD_TPM, (double)tpm);
}

// runtime dispatch over (emissivity, mode) of the march with the full Stokes history, and of node_coef
GRT_HDN void ray_transfer_hist4(const TransferParams &P, const double *rec, int npts, double nu_obs, double *out, double *hist,
                                MarchInfo *mi) {
  const int mode = transfer_mode(P);
  switch (em_slot(P.etype)) {
#define OP(E)                                                                                          \
  case E:                                                                                              \
    switch (mode) {                                                                                    \
      case MODE_DELO: ray_transfer_t<E, MODE_DELO, 4>(P, rec, npts, nu_obs, out, hist, 1, mi); break;   \
      case MODE_FORMAL: ray_transfer_t<E, MODE_FORMAL, 4>(P, rec, npts, nu_obs, out, hist, 1, mi); break; \
      case MODE_QUAD: ray_transfer_t<E, MODE_QUAD, 4>(P, rec, npts, nu_obs, out, hist, 1, mi); break;   \
      case MODE_RK4: ray_transfer_t<E, MODE_RK4, 4>(P, rec, npts, nu_obs, out, hist, 1, mi); break;     \
      default: ray_transfer_t<E, MODE_RK1, 4>(P, rec, npts, nu_obs, out, hist, 1, mi); break;           \
    }                                                                                                  \
    break;
    GRT_FOR_ETYPES(OP)
#undef OP
    default: break;
  }
}
GRT_HDN void node_coef_any(const TransferParams &P, const double *rec, double nu_obs, Coef &c, double &jraw) {
  const FreqConst F = freq_const(nu_obs);
  switch (em_slot(P.etype)) {
#define OP(E)                                         \
  case E: {                                           \
    const RunConst R = run_const<E>(P, nu_obs);       \
    node_coef<E>(P, R, rec, F, c, jraw);              \
  } break;
    GRT_FOR_ETYPES(OP)
#undef OP
    default: break;
  }
}

// the frequency-dependent part for (mdot 1, frequency 1): scal[0] = fac, scal[1] = r%npts.  hist: scratch [maxpts][4].
GRT_HD void geodebug_ray(const TransferParams &P, const double *rec, int npts, int maxpts, double nu_obs, double *hist,
                         double *rows, double *scal) {
  const int nrec = record_doubles(P.etype);
  for (int i = 0; i < 4 * maxpts; i++) hist[i] = 0.0;
  double out[4] = {0, 0, 0, 0};
  MarchInfo mi;
  mi.rnpts = npts;
  mi.fac = 0.0;
  ray_transfer_hist4(P, rec, npts, nu_obs, out, hist, &mi);
  const double fac = mi.fac;
  scal[0] = fac;
  scal[1] = (double)((fac / 1e16 > 0.0) ? mi.rnpts : 1);
  const bool rk = (P.iflag == 0);  // the Runge-Kutta march works on fac-normalised intensities already
  const bool lit = fac / 1e16 > 0.0;
  double tau = 0.0, kprev = 0.0, lprev = 0.0;
  for (int p = 0; p < npts; p++) {
    Coef c;
    double jraw;
    const double *rk_ = rec + (size_t)p * nrec;
    node_coef_any(P, rk_, nu_obs, c, jraw);
    // (sic) an unlit ray keeps the comoving coefficients: rotate / invariant / scaling only run when fac > 0; the
    // invariant ones are listed here either way, the normalisation by fac only when it was applied
    for (int i = 0; i < 4; i++) rows[(size_t)(GD_J + i) * maxpts + p] = lit ? c.j[i] / fac : c.j[i];
    for (int i = 0; i < 7; i++) rows[(size_t)(GD_K + i) * maxpts + p] = c.K[i];
    if (p > 0) tau -= (rk_[0] - lprev) * (kprev + fabs(c.K[0])) / 2.0;
    kprev = fabs(c.K[0]);
    lprev = rk_[0];
    rows[(size_t)GD_TAU * maxpts + p] = lit ? tau : 0.0;
    for (int i = 0; i < 4; i++) {
      const double v = hist[(size_t)p * 4 + i];
      rows[(size_t)(GD_I + i) * maxpts + p] = lit ? (rk ? v : v / fac) : 0.0;
    }
  }
}

}  // namespace grt
