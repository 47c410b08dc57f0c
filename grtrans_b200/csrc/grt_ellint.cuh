// grtrans_b200 -- Carlson-form reductions of the Kerr radial integrals, organised for the GPU.
//
// The reference evaluates every integral of Dexter & Agol (2009) Table 1 through five generic
// routines (geokerr_wrapper.f: ellcubicreal :1188, ellcubiccomplex :1264, ellquarticreal :1362,
// ellquarticcomplex :1444, elldoublecomplex :1531), recomputing the same square roots and symmetric
// arguments for each of the 5 integrals a geodesic point needs (I_u and the four t/phi/lambda
// pieces).  Here each family is a small struct: `prep()` computes the endpoint-dependent quantities
// once, `base()` gives the R_F piece (the I_u integral) and `m2()/m4()` give the integrals with one
// extra factor (a5+b5 t)^-1 / ^-2 re-using everything already in registers.  The algebra is
// Carlson (1988, 1989, 1991, 1992) exactly as the reference cites it, so values agree to round-off.
#pragma once
#include "grt_math.cuh"

namespace grt {

// \int_y^x dt [(a1+b1 t)(a2+b2 t)(a3+b3 t)]^{-1/2} (a4+b4 t)^{p4/2},  all real roots (Table 1 rows 1,2)
struct CubicReal {
  double a1, b1, a2, b2, a3, b3, x, y;
  double X1, X2, X3, Y1, Y2, Y3, U1c, U12, U22, U32, d12, d13;
  GRT_HD void prep(double a1_, double b1_, double a2_, double b2_, double a3_, double b3_, double y_, double x_) {
    a1 = a1_; b1 = b1_; a2 = a2_; b2 = b2_; a3 = a3_; b3 = b3_; x = x_; y = y_;
    d12 = a1 * b2 - a2 * b1;
    d13 = a1 * b3 - a3 * b1;
    X1 = sqrt(a1 + b1 * x); X2 = sqrt(a2 + b2 * x); X3 = sqrt(a3 + b3 * x);
    Y1 = sqrt(a1 + b1 * y); Y2 = sqrt(a2 + b2 * y); Y3 = sqrt(a3 + b3 * y);
    double inv = 1.0 / (x - y);
    U1c = (X1 * Y2 * Y3 + Y1 * X2 * X3) * inv;
    U12 = U1c * U1c;
    U22 = sq((X2 * Y1 * Y3 + Y2 * X1 * X3) * inv);
    U32 = sq((X3 * Y1 * Y2 + Y3 * X1 * X2) * inv);
  }
  GRT_HD double rff() const { return rf(U32, U22, U12); }
  GRT_HD double base(double r) const { return 2.0 * r; }
  GRT_HD double m2(double a4, double b4, double r) const {  // p4 = -2
    double d14 = a1 * b4 - a4 * b1, d24 = a2 * b4 - a4 * b2, d34 = a3 * b4 - a4 * b3;
    double X42 = a4 + b4 * x, Y42 = a4 + b4 * y;
    double W22 = U12 - b4 * d12 * d13 / d14;
    double Q22 = X42 * Y42 / sq(X1 * Y1) * W22;
    double P22 = Q22 + b4 * d24 * d34 / d14;
    double I3c = 2.0 * rc(P22, Q22) - 2.0 * b1 * d12 * d13 / 3.0 / d14 * rj(U32, U22, U12, W22);
    return (b4 * I3c - b1 * 2.0 * r) / d14;
  }
  GRT_HD double m4(double a4, double b4, double r) const {  // p4 = -4
    double d14 = a1 * b4 - a4 * b1, d24 = a2 * b4 - a4 * b2, d34 = a3 * b4 - a4 * b3;
    double X42 = a4 + b4 * x, Y42 = a4 + b4 * y;
    double r12 = a1 / b1 - a2 / b2, r13 = a1 / b1 - a3 / b3;
    double r24i = b2 * b4 / d24, r34i = b3 * b4 / d34;
    double I2c = 2.0 / 3.0 * d12 * d13 * rd(U32, U22, U12) + 2.0 * X1 * Y1 / U1c;
    double K2c = b2 * b3 * I2c - 2.0 * b4 * (X1 * X2 * X3 / X42 - Y1 * Y2 * Y3 / Y42);
    return 0.5 * b4 / d14 / d24 / d34 * K2c + sq(b1 / d14) * (1.0 - 0.5 * r12 * r13 * r24i * r34i) * 2.0 * r;
  }
};

// one real root + complex pair:  [(a1+b1 t)(f+g t+h t^2)]^{-1/2} (a4+b4 t)^{p4/2}  (Table 1 row 3)
struct CubicComplex {
  double a1, b1, f, g, h, x, y;
  double X1, Y1, a11, xi, eta, M2, Lp2, Lm2, beta1, xmy;
  GRT_HD void prep(double a1_, double b1_, double f_, double g_, double h_, double y_, double x_) {
    a1 = a1_; b1 = b1_; f = f_; g = g_; h = h_; x = x_; y = y_;
    X1 = sqrt(a1 + b1 * x);
    Y1 = sqrt(a1 + b1 * y);
    beta1 = g * b1 - 2.0 * h * a1;
    a11 = sqrt(2.0 * f * b1 * b1 - 2.0 * g * a1 * b1 + 2.0 * h * a1 * a1);
    xi = sqrt(f + g * x + h * x * x);
    eta = sqrt(f + g * y + h * y * y);
    xmy = x - y;
    M2 = sq((X1 + Y1) * sqrt(sq(xi + eta) - h * xmy * xmy) / xmy);
    double s2h = sqrt(2.0 * h);
    Lp2 = M2 - beta1 + s2h * a11;
    Lm2 = M2 - beta1 - s2h * a11;
  }
  GRT_HD double rff() const { return rf(M2, Lm2, Lp2); }
  GRT_HD double base(double r) const { return 4.0 * r; }
  GRT_HD double m2(double a4, double b4, double r) const {
    double d14 = a1 * b4 - a4 * b1;
    double X42 = a4 + b4 * x, Y42 = a4 + b4 * y;
    double c44 = sqrt(2.0 * f * b4 * b4 - 2.0 * g * a4 * b4 + 2.0 * h * a4 * a4);
    double a142 = 2.0 * f * b1 * b4 - g * (a1 * b4 + a4 * b1) + 2.0 * h * a1 * a4;
    double U = (X1 * eta + Y1 * xi) / xmy, U2 = U * U;
    double Wp2 = M2 - b1 * (a142 + a11 * c44) / d14;
    double W2 = U2 - a11 * a11 * b4 / 2.0 / d14;
    double Q2 = X42 * Y42 / sq(X1 * Y1) * W2;
    double P2 = Q2 + c44 * c44 * b4 / 2.0 / d14;
    double I3c = (2.0 * a11 / 3.0 / c44) * ((-4.0 * b1 / d14) * (a142 + a11 * c44) * rj(M2, Lm2, Lp2, Wp2) - 6.0 * r +
                                            3.0 * rc(U2, W2)) +
                 2.0 * rc(P2, Q2);
    return (b4 * I3c - b1 * 4.0 * r) / d14;
  }
  GRT_HD double m4(double a4, double b4, double r) const {
    double d14 = a1 * b4 - a4 * b1;
    double X42 = a4 + b4 * x, Y42 = a4 + b4 * y;
    double c442 = 2.0 * f * b4 * b4 - 2.0 * g * a4 * b4 + 2.0 * h * a4 * a4;
    double U = (X1 * eta + Y1 * xi) / xmy;
    double rho = sqrt(2.0 * h) * a11 - beta1;
    double r24Xr34 = 0.5 * c442 / h / (b4 * b4);
    double r12Xr13 = 0.5 * a11 * a11 / h / (b1 * b1);
    double N2c = 2.0 / 3.0 * sqrt(2.0 * h) / a11 * (4.0 * rho * rd(M2, Lm2, Lp2) - 6.0 * r + 3.0 / U) + 2.0 / X1 / Y1 / U;
    double K2c = 0.5 * a11 * a11 * N2c - 2.0 * d14 * (xi / X1 / X42 - eta / Y1 / Y42);
    return 0.5 / d14 / (h * b4 * r24Xr34) * K2c + sq(b1 / d14) * (1.0 - 0.5 * r12Xr13 / r24Xr34) * 4.0 * r;
  }
};

// four real roots: [prod_{i=1..4}(a_i+b_i t)]^{-1/2} (a5+b5 t)^{p5/2}   (Table 1 rows 7,8)
struct QuarticReal {
  double a1, b1, a2, b2, a3, b3, a4, b4, x, y;
  double X1, X2, X3, X4, Y1, Y2, Y3, Y4, U122, U132, U142, d12, d13, d14, d24, d34;
  GRT_HD void prep(double a1_, double b1_, double a2_, double b2_, double a3_, double b3_, double a4_, double b4_,
                   double y_, double x_) {
    a1 = a1_; b1 = b1_; a2 = a2_; b2 = b2_; a3 = a3_; b3 = b3_; a4 = a4_; b4 = b4_; x = x_; y = y_;
    d12 = a1 * b2 - a2 * b1; d13 = a1 * b3 - a3 * b1; d14 = a1 * b4 - a4 * b1;
    d24 = a2 * b4 - a4 * b2; d34 = a3 * b4 - a4 * b3;
    X1 = sqrt(a1 + b1 * x); X2 = sqrt(a2 + b2 * x); X3 = sqrt(a3 + b3 * x); X4 = sqrt(a4 + b4 * x);
    Y1 = sqrt(a1 + b1 * y); Y2 = sqrt(a2 + b2 * y); Y3 = sqrt(a3 + b3 * y); Y4 = sqrt(a4 + b4 * y);
    double inv = 1.0 / (x - y);
    U122 = sq((X1 * X2 * Y3 * Y4 + Y1 * Y2 * X3 * X4) * inv);
    U132 = sq((X1 * X3 * Y2 * Y4 + Y1 * Y3 * X2 * X4) * inv);
    U142 = sq((X1 * X4 * Y2 * Y3 + Y1 * Y4 * X2 * X3) * inv);
  }
  GRT_HD double rff() const { return rf(U122, U132, U142); }
  GRT_HD double base(double r) const { return 2.0 * r; }
  GRT_HD double m2(double a5, double b5, double r) const {
    double d15 = a1 * b5 - a5 * b1, d25 = a2 * b5 - a5 * b2, d35 = a3 * b5 - a5 * b3, d45 = a4 * b5 - a5 * b4;
    double X52 = a5 + b5 * x, Y52 = a5 + b5 * y;
    double W22 = U122 - d13 * d14 * d25 / d15;
    double Q22 = X52 * Y52 / sq(X1 * Y1) * W22;
    double P22 = Q22 + d25 * d35 * d45 / d15;
    double I3 = 2.0 * d12 * d13 * d14 / 3.0 / d15 * rj(U122, U132, U142, W22) + 2.0 * rc(P22, Q22);
    return (b5 * I3 - b1 * 2.0 * r) / d15;
  }
  GRT_HD double m4(double a5, double b5, double r) const {
    double d15 = a1 * b5 - a5 * b1, d25 = a2 * b5 - a5 * b2, d35 = a3 * b5 - a5 * b3, d45 = a4 * b5 - a5 * b4;
    double X52 = a5 + b5 * x, Y52 = a5 + b5 * y;
    double I2 = 2.0 / 3.0 * d12 * d13 * rd(U122, U132, U142) + 2.0 * X1 * Y1 / X4 / Y4 / sqrt(U142);
    double r12 = a1 / b1 - a2 / b2, r13 = a1 / b1 - a3 / b3;
    double r25i = b2 * b5 / d25, r35i = b3 * b5 / d35;
    double A111m1m2 = X1 * X2 * X3 / X4 / X52 - Y1 * Y2 * Y3 / Y4 / Y52;
    return 0.5 * b5 * b5 * d24 * d34 / d15 / d25 / d35 / d45 * I2 +
           sq(b1 / d15) * (1.0 - 0.5 * r12 * r13 * r25i * r35i) * 2.0 * r - b5 * b5 / d15 / d25 / d35 * A111m1m2;
  }
};

// two real roots + complex pair: [(a1+b1 t)(a4+b4 t)(f+g t+h t^2)]^{-1/2} (a5+b5 t)^{p5/2}  (row 5)
struct QuarticComplex {
  double a1, b1, a4, b4, f, g, h, x, y;
  double X1, X4, Y1, Y4, a11, c44, a142, xi, eta, M2, Lp2, Lm2, xmy;
  GRT_HD void prep(double a1_, double b1_, double a4_, double b4_, double f_, double g_, double h_, double y_,
                   double x_) {
    a1 = a1_; b1 = b1_; a4 = a4_; b4 = b4_; f = f_; g = g_; h = h_; x = x_; y = y_;
    X1 = sqrt(a1 + b1 * x); X4 = sqrt(a4 + b4 * x);
    Y1 = sqrt(a1 + b1 * y); Y4 = sqrt(a4 + b4 * y);
    a11 = sqrt(2.0 * f * b1 * b1 - 2.0 * g * a1 * b1 + 2.0 * h * a1 * a1);
    c44 = sqrt(2.0 * f * b4 * b4 - 2.0 * g * a4 * b4 + 2.0 * h * a4 * a4);
    a142 = 2.0 * f * b1 * b4 - g * (a1 * b4 + a4 * b1) + 2.0 * h * a1 * a4;
    xi = sqrt(f + g * x + h * x * x);
    eta = sqrt(f + g * y + h * y * y);
    xmy = x - y;
    M2 = sq((X1 * Y4 + Y1 * X4) * sqrt(sq(xi + eta) - h * xmy * xmy) / xmy);
    Lp2 = M2 + a142 + a11 * c44;
    Lm2 = dmax(M2 + a142 - a11 * c44, 0.0);
  }
  GRT_HD double rff() const { return rf(M2, Lm2, Lp2); }
  GRT_HD double base(double r) const { return 4.0 * r; }
  GRT_HD double m2(double a5, double b5, double r) const {
    double d14 = a1 * b4 - a4 * b1, d15 = a1 * b5 - a5 * b1, d45 = a4 * b5 - a5 * b4;
    double X52 = a5 + b5 * x, Y52 = a5 + b5 * y;
    double c552 = 2.0 * f * b5 * b5 - 2.0 * g * a5 * b5 + 2.0 * h * a5 * a5;
    double c55 = sqrt(c552);
    double a152 = 2.0 * f * b1 * b5 - g * (a1 * b5 + a5 * b1) + 2.0 * h * a5 * a1;
    double U = (X1 * X4 * eta + Y1 * Y4 * xi) / xmy, U2 = U * U;
    double Wp2 = M2 + d14 * (a152 + a11 * c55) / d15;
    double W2 = U2 - a11 * a11 * d45 / 2.0 / d15;
    double Q2 = X52 * Y52 / sq(X1 * Y1) * W2;
    double P2 = Q2 + c552 * d45 / 2.0 / d15;
    double I3 = (2.0 * a11 / 3.0 / c55) *
                    ((4.0 * d14 / d15) * (a152 + a11 * c55) * rj(M2, Lm2, Lp2, Wp2) - 6.0 * r + 3.0 * rc(U2, W2)) +
                2.0 * rc(P2, Q2);
    return (b5 * I3 - b1 * 4.0 * r) / d15;
  }
  GRT_HD double m4(double a5, double b5, double r) const {
    double d15 = a1 * b5 - a5 * b1, d45 = a4 * b5 - a5 * b4;
    double X52 = a5 + b5 * x, Y52 = a5 + b5 * y;
    double c552 = 2.0 * f * b5 * b5 - 2.0 * g * a5 * b5 + 2.0 * h * a5 * a5;
    double U = (X1 * X4 * eta + Y1 * Y4 * xi) / xmy;
    double I2 = 2.0 * a11 / 3.0 / c44 * (4.0 * (a142 + a11 * c44) * rd(M2, Lm2, Lp2) - 6.0 * r) + 2.0 * a11 / c44 / U +
                2.0 * X1 * Y1 / X4 / Y4 / U;
    double A111m1m2 = X1 * xi / X4 / X52 - Y1 * eta / Y4 / Y52;
    return b5 * b5 / (2.0 * d15 * d45) * c44 * c44 / c552 * I2 +
           b1 * b1 / (d15 * d15) * (1.0 - 0.5 * b5 * b5 * a11 * a11 / (b1 * b1) / c552) * 4.0 * r -
           2.0 * b5 * b5 / d15 / c552 * A111m1m2;
  }
};

// two complex pairs: [(f1+g1 t+h1 t^2)(f2+g2 t+h2 t^2)]^{-1/2} (a5+b5 t)^{p5/2}   (Table 1 row 6)
struct DoubleComplex {
  double f1, g1, h1, f2, g2, h2, x, y;
  double xi1, xi2, eta1, eta2, theta1, theta2, M2, d122, d112, d222, Delta, Deltap, Lp2, Lm2, xmy;
  GRT_HD void prep(double f1_, double g1_, double h1_, double f2_, double g2_, double h2_, double y_, double x_) {
    f1 = f1_; g1 = g1_; h1 = h1_; f2 = f2_; g2 = g2_; h2 = h2_; x = x_; y = y_;
    xi1 = sqrt(f1 + g1 * x + h1 * x * x); xi2 = sqrt(f2 + g2 * x + h2 * x * x);
    eta1 = sqrt(f1 + g1 * y + h1 * y * y); eta2 = sqrt(f2 + g2 * y + h2 * y * y);
    theta1 = 2.0 * f1 + g1 * (x + y) + 2.0 * h1 * x * y;
    theta2 = 2.0 * f2 + g2 * (x + y) + 2.0 * h2 * x * y;
    double zeta1 = sqrt(2 * xi1 * eta1 + theta1), zeta2 = sqrt(2 * xi2 * eta2 + theta2);
    xmy = x - y;
    M2 = sq(zeta1 * zeta2 / xmy);
    d122 = 2.0 * f1 * h2 + 2.0 * f2 * h1 - g1 * g2;
    d112 = 4.0 * f1 * h1 - g1 * g1;
    d222 = 4.0 * f2 * h2 - g2 * g2;
    Delta = sqrt(d122 * d122 - d112 * d222);
    Deltap = d122 + Delta;
    Lp2 = M2 + Deltap;
    Lm2 = M2 + (d122 - Delta);
  }
  GRT_HD double rff() const { return rf(M2, Lm2, Lp2); }
  GRT_HD double base(double r) const { return 4.0 * r; }
  // shared H term (Carlson 1992 eq. 2.22)
  GRT_HD double Hterm(double a5, double b5, double &gamma1, double &gamma2, double &beta15, double &beta25) const {
    double U = (xi1 * eta2 + eta1 * xi2) / xmy, U2 = U * U;
    double alpha15 = 2.0 * f1 * b5 - g1 * a5, alpha25 = 2.0 * f2 * b5 - g2 * a5;
    beta15 = g1 * b5 - 2.0 * h1 * a5;
    beta25 = g2 * b5 - 2.0 * h2 * a5;
    gamma1 = 0.5 * (alpha15 * b5 - beta15 * a5);
    gamma2 = 0.5 * (alpha25 * b5 - beta25 * a5);
    double Lambda = d112 * gamma2 / gamma1;
    double Omega2 = M2 + Lambda;
    double psi = 0.5 * (alpha15 * beta25 - alpha25 * beta15);
    double xi5 = a5 + b5 * x, eta5 = a5 + b5 * y;
    double Am111m1 = 1.0 / xi1 * xi2 - 1.0 / eta1 * eta2;
    double A1111m4 = xi1 * xi2 / (xi5 * xi5) - eta1 * eta2 / (eta5 * eta5);
    double XX = xi5 * eta5 * (theta1 * 0.5 * Am111m1 - xi5 * eta5 * A1111m4) / (xmy * xmy);
    double S = 0.5 * (M2 + d122) - U2, S2 = S * S;
    double mu = gamma1 * xi5 * eta5 / xi1 / eta1;
    double T = mu * S + 2.0 * gamma1 * gamma2;
    double V2 = mu * mu * (S2 + Lambda * U2);
    double b2 = Omega2 * Omega2 * (S2 / U2 + Lambda);
    double a2 = b2 + Lambda * Lambda * psi * psi / gamma1 / gamma2;
    return d112 * psi * (rj(M2, Lm2, Lp2, Omega2) / 3.0 + 0.5 * rc(a2, b2)) / (gamma1 * gamma1) - XX * rc(T * T, V2);
  }
  GRT_HD double m2(double a5, double b5, double r) const {
    double g1_, g2_, b15, b25;
    double H = Hterm(a5, b5, g1_, g2_, b15, b25);
    return -2.0 * (b5 * H + b15 * r / g1_);
  }
  GRT_HD double m4(double a5, double b5, double r) const {
    double gm1, gm2, b15, b25;
    double H = Hterm(a5, b5, gm1, gm2, b15, b25);
    double U = (xi1 * eta2 + eta1 * xi2) / xmy;
    double xi5 = a5 + b5 * x, eta5 = a5 + b5 * y;
    double A1111m2 = xi1 * xi2 / xi5 - eta1 * eta2 / eta5;
    double xi1p = 0.5 * (g1 + 2.0 * h1 * x) / xi1, eta1p = 0.5 * (g1 + 2.0 * h1 * y) / eta1;
    double B = xi1p * xi2 - eta1p * eta2;
    double G = 2.0 / 3.0 * Delta * Deltap * rd(M2, Lm2, Lp2) + 0.5 * Delta / U +
               (d122 * theta1 - d112 * theta2) / 4.0 / xi1 / eta1 / U;
    double Sigma = G - Deltap * r + B;
    return b5 * (b15 / gm1 + b25 / gm2) * H + b15 * b15 * r / (gm1 * gm1) + b5 * b5 * (Sigma - b5 * A1111m2) / gm1 / gm2;
  }
};

}  // namespace grt
