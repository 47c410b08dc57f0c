// ORACLE -- TEST INFRASTRUCTURE ONLY (see geokerr.cc header).
// CPU restatement of the transfer integrators of the hot path, radtrans_integrate.f90:
//   integrate :139-198, DELO :746-842 (+opacity_matrix :735, imatrix_4 :685), FORMAL :844-876
//   (+calc_O :615-683), quadrature :878-882, and the LSODA path :200-307 with its coefficient model
//   radtrans_aux :453-465.  ODEPACK's DLSODA itself is NOT restated (north_star replaces it with an
//   embedded adaptive Runge-Kutta at the same tolerances): here the same ODE, same ray cut/trim,
//   same linear-in-lambda coefficient lookup and same hmax are integrated with Dormand-Prince 5(4).
//   tests/test_oracle_integrators.py checks this against scipy.integrate.odeint (= ODEPACK LSODA).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "oracle.h"

namespace orc {

// math.f90:30-54  cumulative trapezoid
void tsum(int n, const double *x, const double *y, double *out) {
  double acc = 0.0;
  out[0] = 0.0;
  for (int i = 1; i < n; i++) {
    double xdif = x[i] - x[i - 1];
    double yavg = (y[i - 1] + y[i]) / 2.0;
    acc += xdif * yavg;
    out[i] = acc;
  }
}

typedef double M4[4][4];

// radtrans_integrate.f90:685-733 (cofactor inverse)
// test-only alternative (oracle.h: legacy.delo_inverse_lu): the same inverse by Gauss-Jordan elimination with partial
// pivoting.  Algebraically identical to the cofactor formula; where the two differ in the image, the difference is the
// rounding-level ambiguity of the reference ALGORITHM (near-singular DELO matrices in cells where the polsynchth fits
// give |alpha_V| > alpha_I), which no implementation can be held to.
static void imatrix_4_lu(const M4 a, M4 b) {
  double w[4][8];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      w[i][j] = a[i][j];
      w[i][4 + j] = (i == j) ? 1.0 : 0.0;
    }
  for (int c = 0; c < 4; c++) {
    int piv = c;
    for (int r = c + 1; r < 4; r++)
      if (std::fabs(w[r][c]) > std::fabs(w[piv][c])) piv = r;
    if (piv != c)
      for (int j = 0; j < 8; j++) std::swap(w[c][j], w[piv][j]);
    const double d = w[c][c];
    for (int j = 0; j < 8; j++) w[c][j] /= d;
    for (int r = 0; r < 4; r++)
      if (r != c) {
        const double f = w[r][c];
        for (int j = 0; j < 8; j++) w[r][j] -= f * w[c][j];
      }
  }
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) b[i][j] = w[i][4 + j];
}

static void imatrix_4(const M4 a, M4 b) {
  if (legacy.delo_inverse_lu) {
    imatrix_4_lu(a, b);
    return;
  }
  double a11 = a[0][0], a12 = a[0][1], a13 = a[0][2], a14 = a[0][3];
  double a21 = a[1][0], a22 = a[1][1], a23 = a[1][2], a24 = a[1][3];
  double a31 = a[2][0], a32 = a[2][1], a33 = a[2][2], a34 = a[2][3];
  double a41 = a[3][0], a42 = a[3][1], a43 = a[3][2], a44 = a[3][3];
  b[0][0] = a22 * a33 * a44 + a23 * a34 * a42 + a24 * a32 * a43 - a22 * a34 * a43 - a23 * a32 * a44 - a24 * a33 * a42;
  b[0][1] = a12 * a34 * a43 + a13 * a32 * a44 + a14 * a33 * a42 - a12 * a33 * a44 - a13 * a34 * a42 - a14 * a32 * a43;
  b[0][2] = a12 * a23 * a44 + a13 * a24 * a42 + a14 * a22 * a43 - a12 * a24 * a43 - a13 * a22 * a44 - a14 * a23 * a42;
  b[0][3] = a12 * a24 * a33 + a13 * a22 * a34 + a14 * a23 * a32 - a12 * a23 * a34 - a13 * a24 * a32 - a14 * a22 * a33;
  b[1][0] = a21 * a34 * a43 + a23 * a31 * a44 + a24 * a33 * a41 - a21 * a33 * a44 - a23 * a34 * a41 - a24 * a31 * a43;
  b[1][1] = a11 * a33 * a44 + a13 * a34 * a41 + a14 * a31 * a43 - a11 * a34 * a43 - a13 * a31 * a44 - a14 * a33 * a41;
  b[1][2] = a11 * a24 * a43 + a13 * a21 * a44 + a14 * a23 * a41 - a11 * a23 * a44 - a13 * a24 * a41 - a14 * a21 * a43;
  b[1][3] = a11 * a23 * a34 + a13 * a24 * a31 + a14 * a21 * a33 - a11 * a24 * a33 - a13 * a21 * a34 - a14 * a23 * a31;
  b[2][0] = a21 * a32 * a44 + a22 * a34 * a41 + a24 * a31 * a42 - a21 * a34 * a42 - a22 * a31 * a44 - a24 * a32 * a41;
  b[2][1] = a11 * a34 * a42 + a12 * a31 * a44 + a14 * a32 * a41 - a11 * a32 * a44 - a12 * a34 * a41 - a14 * a31 * a42;
  b[2][2] = a11 * a22 * a44 + a12 * a24 * a41 + a14 * a21 * a42 - a11 * a24 * a42 - a12 * a21 * a44 - a14 * a22 * a41;
  b[2][3] = a11 * a24 * a32 + a12 * a21 * a34 + a14 * a22 * a31 - a11 * a22 * a34 - a12 * a24 * a31 - a14 * a21 * a32;
  b[3][0] = a21 * a33 * a42 + a22 * a31 * a43 + a23 * a32 * a41 - a21 * a32 * a43 - a22 * a33 * a41 - a23 * a31 * a42;
  b[3][1] = a11 * a32 * a43 + a12 * a33 * a41 + a13 * a31 * a42 - a11 * a33 * a42 - a12 * a31 * a43 - a13 * a32 * a41;
  b[3][2] = a11 * a23 * a42 + a12 * a21 * a43 + a13 * a22 * a41 - a11 * a22 * a43 - a12 * a23 * a41 - a13 * a21 * a42;
  b[3][3] = a11 * a22 * a33 + a12 * a23 * a31 + a13 * a21 * a32 - a11 * a23 * a32 - a12 * a21 * a33 - a13 * a22 * a31;
  double detA = a[0][0] * b[0][0] + a[1][0] * b[0][1] + a[2][0] * b[0][2] + a[3][0] * b[0][3];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) b[i][j] = b[i][j] / detA;
}

// radtrans_integrate.f90:735-744   a = (aI,aQ,aU,aV), p = (rhoQ,rhoU,rhoV)
static void opacity_matrix(const double *a, const double *p, M4 K) {
  K[0][0] = a[0]; K[1][1] = a[0]; K[2][2] = a[0]; K[3][3] = a[0];
  K[1][0] = a[1]; K[2][0] = a[2]; K[3][0] = a[3];
  K[0][1] = a[1]; K[0][2] = a[2]; K[0][3] = a[3];
  K[2][1] = -p[2]; K[3][1] = p[1]; K[1][2] = p[2]; K[1][3] = -p[1];
  K[3][2] = -p[0]; K[2][3] = p[0];
}

static void matvec(const M4 A, const double *x, double *y) {
  for (int i = 0; i < 4; i++) y[i] = A[i][0] * x[0] + A[i][1] * x[1] + A[i][2] * x[2] + A[i][3] * x[3];
}
static void matmul(const M4 A, const M4 B, M4 C) {
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) C[i][j] = A[i][0] * B[0][j] + A[i][1] * B[1][j] + A[i][2] * B[2][j] + A[i][3] * B[3][j];
}

// radtrans_integrate.f90:795-842.  j:[npts][4], K:[npts][7] (aI,aQ,aU,aV,rhoQ,rhoU,rhoV)
static void integrate_delo(int npts, const double *x, const double *tau, const double *j, const double *K,
                           double thin, double *intensity) {
  double iprev[4] = {0, 0, 0, 0};
  for (int i = 0; i < 4; i++) intensity[i] = 0.0;
  for (int k = npts - 1; k >= 1; k--) {  // 1-based k as in the reference
    const double *ak = K + 7 * (k - 1), *ak1 = K + 7 * k;
    const double *jk = j + 4 * (k - 1), *jk1 = j + 4 * k;
    double delta = tau[k] - tau[k - 1];
    double dx = x[k - 1] - x[k];
    M4 K0, K1, matrix, imatrix, qt, tmp;
    opacity_matrix(ak, ak + 4, K0);
    opacity_matrix(ak1, ak1 + 4, K1);
    double pt[4];
    if (delta > thin) {
      double E = std::exp(-delta);
      double F = 1.0 - E;
      double G = (1.0 - (1.0 + delta) * E) / delta;
      double Sp[4], Sp1[4];
      for (int i = 0; i < 4; i++) {
        Sp[i] = jk[i] / ak[0];
        Sp1[i] = jk1[i] / ak1[0];
      }
      M4 Kp, Kp1;
      for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) {
          double id = (r == c) ? 1.0 : 0.0;
          Kp[r][c] = K0[r][c] / ak[0] - id;
          Kp1[r][c] = K1[r][c] / ak1[0] - id;
          matrix[r][c] = id + (F - G) * Kp[r][c];
        }
      imatrix_4(matrix, imatrix);
      double v[4];
      for (int i = 0; i < 4; i++) v[i] = (F - G) * Sp[i] + G * Sp1[i];
      matvec(imatrix, v, pt);
      for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) tmp[r][c] = ((r == c) ? 1.0 : 0.0) * E - G * Kp1[r][c];
      matmul(imatrix, tmp, qt);
    } else {
      double ki = ak[0];
      for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) {
          double id = (r == c) ? 1.0 : 0.0;
          matrix[r][c] = (1.0 - delta / 2.0 + delta * delta / 6.0) * id + (0.5 * dx - 1.0 / 6.0 * dx * dx * ki) * K0[r][c];
        }
      imatrix_4(matrix, imatrix);
      double v[4];
      for (int i = 0; i < 4; i++)
        v[i] = (0.5 * dx * jk[i] - 1.0 / 6.0 * dx * dx * ki * jk[i]) + (0.5 * jk1[i] * dx - 1.0 / 3.0 * dx * dx * ki * jk1[i]);
      matvec(imatrix, v, pt);
      for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) {
          double id = (r == c) ? 1.0 : 0.0;
          tmp[r][c] = id * (1.0 - 0.5 * dx * ki + 1.0 / 6.0 * dx * dx * ki * ki) - (0.5 * dx - 1.0 / 3.0 * dx * dx) * K1[r][c];
        }
      matmul(imatrix, tmp, qt);
    }
    double q[4];
    matvec(qt, iprev, q);
    int col = npts - k;  // 0-based column of intensity(:,npts-k+1)
    for (int i = 0; i < 4; i++) {
      intensity[4 * col + i] = pt[i] + q[i];
      iprev[i] = intensity[4 * col + i];
    }
  }
}

// radtrans_integrate.f90:615-683
static void calc_O(const double *a, const double *rho, double dx, M4 O) {
  double onopol = std::exp(-a[0] * dx);
  double aq = a[1], au = a[2], av = a[3];
  double rhoq = rho[0], rhou = rho[1], rhov = rho[2];
  double a2 = aq * aq + au * au + av * av;
  double p2 = rhoq * rhoq + rhou * rhou + rhov * rhov;
  if (a2 == 0.0 && p2 == 0.0) {
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) O[r][c] = ((r == c) ? 1.0 : 0.0) * onopol;
    return;
  }
  double ap = aq * rhoq + au * rhou + av * rhov;
  double d = (a2 - p2);
  double lam1 = std::sqrt(std::sqrt(d * d / 4.0 + ap * ap) + d / 2.0);
  double lam2 = std::sqrt(std::sqrt(d * d / 4.0 + ap * ap) - d / 2.0);
  double theta = lam1 * lam1 + lam2 * lam2;
  double sig = std::signbit(ap) ? -1.0 : 1.0;
  M4 M2, M3, M4_;
  // M(r,c) with the reference's (row,col) = (first,second) index
  M2[0][0] = 0.0; M2[1][0] = lam2 * aq - sig * lam1 * rhoq; M2[2][0] = lam2 * au - sig * lam1 * rhou; M2[3][0] = lam2 * av - sig * lam1 * rhov;
  M2[0][1] = lam2 * aq - sig * lam1 * rhoq; M2[1][1] = 0.0; M2[2][1] = -sig * lam1 * av - lam2 * rhov; M2[3][1] = sig * lam1 * au + lam2 * rhou;
  M2[0][2] = lam2 * au - sig * lam1 * rhou; M2[1][2] = sig * lam1 * av + lam2 * rhov; M2[2][2] = 0.0; M2[3][2] = -sig * lam1 * aq - lam2 * rhoq;
  M2[0][3] = lam2 * av - sig * lam1 * rhov; M2[1][3] = -sig * lam1 * au - lam2 * rhou; M2[2][3] = sig * lam1 * aq + lam2 * rhoq; M2[3][3] = 0.0;
  M3[0][0] = 0.0; M3[1][0] = lam1 * aq + sig * lam2 * rhoq; M3[2][0] = lam1 * au + sig * lam2 * rhou; M3[3][0] = lam1 * av + sig * lam2 * rhov;
  M3[0][1] = lam1 * aq + sig * lam2 * rhoq; M3[1][1] = 0.0; M3[2][1] = sig * lam2 * av - lam1 * rhov; M3[3][1] = -sig * lam2 * au + lam1 * rhou;
  M3[0][2] = lam1 * au + sig * lam2 * rhou; M3[1][2] = -sig * lam2 * av + lam1 * rhov; M3[2][2] = 0.0; M3[3][2] = sig * lam2 * aq - lam1 * rhoq;
  M3[0][3] = lam1 * av + sig * lam2 * rhov; M3[1][3] = sig * lam2 * au - lam1 * rhou; M3[2][3] = -sig * lam2 * aq + lam1 * rhoq; M3[3][3] = 0.0;
  M4_[0][0] = (a2 + p2) / 2.0; M4_[1][0] = au * rhov - av * rhou; M4_[2][0] = av * rhoq - aq * rhov; M4_[3][0] = aq * rhou - au * rhoq;
  M4_[0][1] = av * rhou - au * rhov; M4_[1][1] = aq * aq + rhoq * rhoq - (a2 + p2) / 2.0; M4_[2][1] = aq * au + rhoq * rhou; M4_[3][1] = av * aq + rhov * rhoq;
  M4_[0][2] = aq * rhov - av * rhoq; M4_[1][2] = aq * au + rhoq * rhou; M4_[2][2] = au * au + rhou * rhou - (a2 + p2) / 2.0; M4_[3][2] = au * av + rhou * rhov;
  M4_[0][3] = au * rhoq - aq * rhou; M4_[1][3] = av * aq + rhov * rhoq; M4_[2][3] = au * av + rhou * rhov; M4_[3][3] = av * av + rhov * rhov - (a2 + p2) / 2.0;
  double ch = std::cosh(lam1 * dx), cs = std::cos(lam2 * dx), sh = std::sinh(lam1 * dx), sn = std::sin(lam2 * dx);
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) {
      double m1 = (r == c) ? 1.0 : 0.0;
      double m2 = 1.0 / theta * M2[r][c], m3 = 1.0 / theta * M3[r][c], m4 = 2.0 / theta * M4_[r][c];
      O[r][c] = onopol * (1.0 / 2.0 * (ch + cs) * m1 - sn * m2 - sh * m3 + 1.0 / 2.0 * (ch - cs) * m4);
    }
}

// radtrans_integrate.f90:844-876
static void integrate_formal(int npts, const double *x, const double *j, const double *K, double *intensity) {
  double iprev[4] = {0, 0, 0, 0};
  for (int i = 0; i < 4; i++) intensity[i] = 0.0;
  for (int k = npts - 1; k >= 1; k--) {
    int l = npts - k;  // 0-based column
    const double *at = K + 7 * (k - 1), *jt = j + 4 * (k - 1);
    double dx = x[k - 1] - x[k];
    M4 Ot;
    calc_O(at, at + 4, dx, Ot);
    double v[4], inew[4];
    for (int i = 0; i < 4; i++) v[i] = jt[i] * dx + iprev[i];
    matvec(Ot, v, inew);
    for (int i = 0; i < 4; i++) {
      iprev[i] = inew[i];
      intensity[4 * l + i] = inew[i];
    }
  }
}

// ------------------------------------------------------------------ LSODA path
struct CoefModel {  // radtrans_aux :453-465 on the reversed grid ss
  int npts, neq, nk;
  const double *s0, *jj, *KK;
  bool as_written;
  // interval `cell` (1-based original index indx): lam in [s0(indx), s0(indx-1))
  void eval(int indx, double lam, double *j, double *K) const {
    double w = (lam - s0[indx - 1]) / (s0[indx - 2] - s0[indx - 1]);
    int uindx = as_written ? std::min(indx + 1, npts) : (indx - 1);
    for (int i = 0; i < neq; i++) j[i] = (1.0 - w) * jj[(indx - 1) * neq + i] + w * jj[(uindx - 1) * neq + i];
    for (int i = 0; i < nk; i++) K[i] = (1.0 - w) * KK[(indx - 1) * nk + i] + w * KK[(uindx - 1) * nk + i];
  }
};

static inline void rhs(int neq, const double *j, const double *K, const double *I, double *d) {
  if (neq == 4) {  // radtrans_rhs_form :349-362
    d[0] = j[0] - (K[0] * I[0] + K[1] * I[1] + K[2] * I[2] + K[3] * I[3]);
    d[1] = j[1] - (K[1] * I[0] + K[0] * I[1] + K[6] * I[2] - K[5] * I[3]);
    d[2] = j[2] - (K[2] * I[0] - K[6] * I[1] + K[0] * I[2] + K[4] * I[3]);
    d[3] = j[3] - (K[3] * I[0] + K[5] * I[1] - K[4] * I[2] + K[0] * I[3]);
  } else {
    d[0] = j[0] - K[0] * I[0];
  }
}

// Dormand-Prince 5(4) over one tabulated interval; returns false on step overflow
static bool dopri_interval(const CoefModel &cm, int indx, double la, double lb, double *y, double &h, double rtol,
                           double atol, double hmax, int maxsteps, int &nsteps) {
  static const double c2 = 1. / 5, c3 = 3. / 10, c4 = 4. / 5, c5 = 8. / 9;
  static const double a21 = 1. / 5, a31 = 3. / 40, a32 = 9. / 40, a41 = 44. / 45, a42 = -56. / 15, a43 = 32. / 9,
                      a51 = 19372. / 6561, a52 = -25360. / 2187, a53 = 64448. / 6561, a54 = -212. / 729,
                      a61 = 9017. / 3168, a62 = -355. / 33, a63 = 46732. / 5247, a64 = 49. / 176, a65 = -5103. / 18656,
                      a71 = 35. / 384, a73 = 500. / 1113, a74 = 125. / 192, a75 = -2187. / 6784, a76 = 11. / 84;
  static const double e1 = 71. / 57600, e3 = -71. / 16695, e4 = 71. / 1920, e5 = -17253. / 339200, e6 = 22. / 525,
                      e7 = -1. / 40;
  int neq = cm.neq;
  double t = la;
  int local = 0;
  while (t < lb) {
    if (++local > maxsteps) return false;
    double hh = std::min(std::min(h, hmax), lb - t);
    bool last = (hh >= lb - t);
    double j[4], K[7], k1[4], k2[4], k3[4], k4[4], k5[4], k6[4], k7[4], yt[4], ynew[4];
    cm.eval(indx, t, j, K);
    rhs(neq, j, K, y, k1);
    for (int i = 0; i < neq; i++) yt[i] = y[i] + hh * a21 * k1[i];
    cm.eval(indx, t + c2 * hh, j, K);
    rhs(neq, j, K, yt, k2);
    for (int i = 0; i < neq; i++) yt[i] = y[i] + hh * (a31 * k1[i] + a32 * k2[i]);
    cm.eval(indx, t + c3 * hh, j, K);
    rhs(neq, j, K, yt, k3);
    for (int i = 0; i < neq; i++) yt[i] = y[i] + hh * (a41 * k1[i] + a42 * k2[i] + a43 * k3[i]);
    cm.eval(indx, t + c4 * hh, j, K);
    rhs(neq, j, K, yt, k4);
    for (int i = 0; i < neq; i++) yt[i] = y[i] + hh * (a51 * k1[i] + a52 * k2[i] + a53 * k3[i] + a54 * k4[i]);
    cm.eval(indx, t + c5 * hh, j, K);
    rhs(neq, j, K, yt, k5);
    for (int i = 0; i < neq; i++)
      yt[i] = y[i] + hh * (a61 * k1[i] + a62 * k2[i] + a63 * k3[i] + a64 * k4[i] + a65 * k5[i]);
    cm.eval(indx, t + hh, j, K);
    rhs(neq, j, K, yt, k6);
    for (int i = 0; i < neq; i++)
      ynew[i] = y[i] + hh * (a71 * k1[i] + a73 * k3[i] + a74 * k4[i] + a75 * k5[i] + a76 * k6[i]);
    rhs(neq, j, K, ynew, k7);
    double err = 0.0;
    for (int i = 0; i < neq; i++) {
      double ei = hh * (e1 * k1[i] + e3 * k3[i] + e4 * k4[i] + e5 * k5[i] + e6 * k6[i] + e7 * k7[i]);
      double sc = atol + rtol * std::max(std::fabs(y[i]), std::fabs(ynew[i]));
      err += (ei / sc) * (ei / sc);
    }
    err = std::sqrt(err / neq);
    if (!(err <= 1.0)) {
      if (!(err == err)) return false;  // NaN
      h = hh * std::max(0.2, 0.9 * std::pow(err, -0.2));
      continue;
    }
    nsteps++;
    for (int i = 0; i < neq; i++) y[i] = ynew[i];
    t = last ? lb : t + hh;
    double fac = (err > 0.0) ? std::min(5.0, std::max(0.2, 0.9 * std::pow(err, -0.2))) : 5.0;
    if (!last || fac < 1.0) h = hh * fac;
  }
  return true;
}

// locate.f (double)
static int dlocate(const double *xx /*0-based*/, int n, double x) {
  bool ascnd = (xx[n - 1] >= xx[0]);
  int jl = 0, ju = n + 1;
  while (ju - jl > 1) {
    int jm = (ju + jl) / 2;
    if (ascnd == (x >= xx[jm - 1]))
      jl = jm;
    else
      ju = jm;
  }
  if (x == xx[0]) return 1;
  if (x == xx[n - 1]) return n - 1;
  return jl;
}

// radtrans_integrate.f90:200-307
static int integrate_lsoda(int neq, int npts, const double *s0, const double *jj, const double *KK, const double *tau,
                           double *intensity, const RtOptions &opt, int *nsteps_out) {
  int nk = 1 + neq * (neq - 1) / 2;
  int i1 = 1, i2 = npts, lamdex;
  double maxtau = tau[0];
  for (int i = 1; i < npts; i++) maxtau = std::max(maxtau, tau[i]);
  if (maxtau <= opt.max_tau) {
    lamdex = npts;
  } else {
    lamdex = dlocate(tau, npts, opt.max_tau);
    lamdex = lamdex + 1;
  }
  if (jj[0] == 0.0) {
    for (int ii = 2; ii <= npts; ii++)
      if (jj[(ii - 1) * neq] != 0.0) {
        i1 = ii;
        break;
      }
  } else {
    i1 = 1;
  }
  if (jj[(lamdex - 1) * neq] == 0.0) {
    i2 = 0;  // (reference leaves i2 = uninitialised nptstot if nothing is found)
    for (int ii = 1; ii <= lamdex - 1; ii++)
      if (jj[(lamdex - ii - 1) * neq] != 0.0) {
        i2 = lamdex - ii;
        break;
      }
  } else {
    i2 = lamdex;
  }
  int nptsout = i2;
  for (int i = 0; i < neq * npts; i++) intensity[i] = 0.0;
  int nsteps = 0;
  if (i2 > i1) {
    CoefModel cm{npts, neq, nk, s0, jj, KK, opt.lsoda_neighbor_as_written};
    double y[4] = {0, 0, 0, 0};
    // outputs: intensity(:,i1) = I0 at s0(i2); intensity(:,i1+m) at s0(i2-m)
    for (int i = 0; i < neq; i++) intensity[(i1 - 1) * neq + i] = y[i];
    double h = opt.hmax;
    bool ok = true;
    for (int m = 1; m <= i2 - i1; m++) {
      int indx = i2 - m + 1;  // interval [s0(indx), s0(indx-1)]
      double la = s0[indx - 1], lb = s0[indx - 2];
      if (lb > la) ok = dopri_interval(cm, indx, la, lb, y, h, opt.ortol, opt.oatol, opt.hmax, opt.maxsteps, nsteps);
      if (!ok) break;
      for (int i = 0; i < neq; i++) intensity[(i1 - 1 + m) * neq + i] = y[i];
    }
    if (!ok)
      for (int i = 0; i < neq * npts; i++) intensity[i] = 0.0;  // odepack.f90:91-96
  }
  if (nsteps_out) *nsteps_out = nsteps;
  return nptsout;
}

// radtrans_integrate.f90:139-198
int rt_integrate(int iflag, int neq, int npts, const double *s, const double *j, const double *K, const double *tau,
                 double *intensity, const RtOptions &opt, int *nsteps_out) {
  for (int i = 0; i < neq * npts; i++) intensity[i] = 0.0;
  if (nsteps_out) *nsteps_out = 0;
  if (iflag == 0) return integrate_lsoda(neq, npts, s, j, K, tau, intensity, opt, nsteps_out);
  if (neq == 4) {
    if (iflag == 1)
      integrate_delo(npts, s, tau, j, K, opt.thin, intensity);
    else
      integrate_formal(npts, s, j, K, intensity);
  } else {
    // quadrature :878-882  intensity(1,:)=-tsum(s,j*exp(-tau))
    std::vector<double> y(npts), o(npts);
    for (int i = 0; i < npts; i++) y[i] = j[i * neq] * std::exp(-tau[i]);
    tsum(npts, s, y.data(), o.data());
    for (int i = 0; i < npts; i++) intensity[i * neq] = -o[i];
  }
  return npts;
}

}  // namespace orc
