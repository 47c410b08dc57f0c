// grtrans_b200 -- device math primitives for the Kerr geodesic kernels (sm_100a, FP64).
//
// Carlson symmetric elliptic integrals R_F, R_C, R_D, R_J by the duplication theorem, Jacobi
// sn/cn/dn by the descending Landen/AGM recurrence and a small complex type for Laguerre root
// polishing.  Tolerances and series coefficients are the ones the reference uses
// (geokerr_wrapper.f:3277-3346 SNCNDN CA=3e-8, :3444 rf ERRTOL=.0025, :3608 rc .0012,
// :3648 rd .0015, :3697 rj .0015) so iteration counts -- and therefore results -- agree to
// round-off with the Fortran.
//
// Everything here is __host__ __device__ so the same kernel code can be driven on the CPU by
// tests/host_emul (a debugging harness; the product only ever launches the CUDA kernels).
#pragma once
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define GRT_HD __host__ __device__ __forceinline__
#define GRT_HDN inline __host__ __device__ __noinline__
#else
#define GRT_HD inline
#define GRT_HDN inline
#endif

#define GRT_PI 3.141592653589793238462643383279502884

namespace grt {

GRT_HD double dmax(double a, double b) { return a > b ? a : b; }
GRT_HD double dmin(double a, double b) { return a < b ? a : b; }
GRT_HD double sgn(double b) { return signbit(b) ? -1.0 : 1.0; }      // Fortran SIGN(1d0,b)
GRT_HD double pm1(int n) { return (n & 1) ? -1.0 : 1.0; }            // (-1)**n
GRT_HD double sq(double x) { return x * x; }

// ---- real(4) libm calls ---------------------------------------------------------------------------
// The reference's real(4) code calls glibc's float functions, which are correctly rounded in all but
// vanishingly rare cases.  CUDA's sinf/cosf/logf/powf/... are 1-2 ulp functions, and a 1-ulp difference
// is amplified by the cancellations downstream (u^t normalisation, b.b), so on the device the float32
// calls are evaluated in FP64 and rounded once -- the same value the CPU produces.
#if defined(__CUDA_ARCH__)
GRT_HD float r4sin(float x) { return (float)sin((double)x); }
GRT_HD float r4cos(float x) { return (float)cos((double)x); }
GRT_HD float r4log(float x) { return (float)log((double)x); }
GRT_HD float r4exp(float x) { return (float)exp((double)x); }
GRT_HD float r4acos(float x) { return (float)acos((double)x); }
GRT_HD float r4pow(float x, float y) { return (float)pow((double)x, (double)y); }
#else
GRT_HD float r4sin(float x) { return sinf(x); }
GRT_HD float r4cos(float x) { return cosf(x); }
GRT_HD float r4log(float x) { return logf(x); }
GRT_HD float r4exp(float x) { return expf(x); }
GRT_HD float r4acos(float x) { return acosf(x); }
GRT_HD float r4pow(float x, float y) { return powf(x, y); }
#endif

GRT_HD double nan_val() {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double(0x7ff8000000000000LL);
#else
  return __builtin_nan("");
#endif
}

// ---- exp() for the transfer hot loop -------------------------------------------------------------
// The emissivity + DELO loop calls exp() three times per (point, frequency).  On this ISA every 64-bit constant of a
// polynomial costs an issue slot of its own (one LDC, or two UMOV for an immediate), so a long polynomial pays twice: ncu
// attributed 6 % of k_transfer's issued instructions to the 11 Horner steps of the previous version, half of them constant
// loads.  This one is table driven:  e^x = 2^k 2^(j/64) e^r,  n = round(x 64/ln2) = 64 k + j,  |r| <= ln2/128, with
// 2^(j/64) = hi + lo from a 64-entry table (one 16-byte L1 load, issued right after n is known and needed last) and
// e^r - 1 = r + r^2 (1/2 + r (1/6 + r (1/24 + r/120))) (truncation < 3.5e-17 relative): 6 dependent FP64 operations and 4
// constants instead of 11 and 12.  No branch; the scaling by 2^k is done in two halves so that results in the subnormal
// range round once, like the reference's (the DELO step has to tell "underflowed to a subnormal" from 0).
// Domain: x <= 709 (callers guarantee it); x < -800 -> 0; NaN -> NaN.  Error <= 1 ulp (tests/test_host_emul.py).
#define GRT_EXP_TAB \
  1.0, 0.0, 1.0108892860517005, -1.5234778603368577e-17, 1.0218971486541166, 5.109225028973444e-17, \
  1.0330248790212284, 7.600838874027088e-18, 1.0442737824274138, 8.551889705537965e-17, 1.0556451783605572, \
  1.759325738772092e-18, 1.0671404006768237, -7.899853966841582e-17, 1.0787607977571199, \
  -6.656660436056593e-17, 1.0905077326652577, -3.046782079812471e-17, 1.102382583307841, \
  5.2660368715706944e-17, 1.1143867425958924, 1.0410278456845571e-16, 1.1265216186082418, \
  5.165856758795457e-17, 1.1387886347566916, 8.912812676025408e-17, 1.1511892299529827, 3.250710218863827e-17, \
  1.1637248587775775, 3.8292048369240935e-17, 1.1763969916502812, 5.554203254218079e-17, 1.189207115002721, \
  3.982015231465646e-17, 1.202156731452703, 6.644981499252301e-17, 1.215247359980469, -7.712630692681488e-17, \
  1.22848053610687, -1.89878163130253e-17, 1.241857812073484, 4.658027591836937e-17, 1.255380757024691, \
  -6.7113898212968784e-18, 1.2690509571917332, 2.667932131342186e-18, 1.2828700160787783, \
  1.713594918243561e-17, 1.2968395546510096, 2.5382502794888315e-17, 1.3109612115247644, \
  -7.181536135519454e-17, 1.3252366431597413, -2.8587312100388614e-17, 1.339667524053303, \
  8.927282594831732e-17, 1.3542555469368927, 7.70094837980299e-17, 1.3690024229745905, 9.593797919118849e-17, \
  1.383909881963832, -6.770511658794786e-17, 1.3989796725383112, -9.614213209051323e-17, 1.4142135623730951, \
  -9.667293313452913e-17, 1.42961333839197, -1.2031642489053655e-17, 1.4451808069770467, \
  -3.0237581349939873e-17, 1.460917794180647, -5.600377186075216e-17, 1.4768261459394993, \
  -3.483994556892796e-17, 1.4929077282912648, 1.4192920154284036e-17, 1.5091644275934228, \
  -1.016455327754295e-16, 1.5255981507445384, -1.1024941712342561e-16, 1.5422108254079407, \
  7.949834809697621e-17, 1.559004400237837, 3.7812070533575275e-17, 1.5759808451078865, \
  -1.0136916471278304e-17, 1.593142151342267, -1.0094406542311964e-16, 1.6104903319492543, \
  2.4707192569797888e-17, 1.6280274218573478, -6.712955084707084e-17, 1.645755478153965, \
  -1.0125679913674773e-16, 1.6636765803267364, 5.8909926967131e-17, 1.681792830507429, 8.199010020581497e-17, \
  1.7001063537185235, -8.0237193703977e-18, 1.718619298122478, -1.851380418263111e-17, 1.7373338352737062, \
  3.164389299292957e-17, 1.7562521603732995, 2.960140695448873e-17, 1.7753764925265212, 6.429731796556572e-17, \
  1.7947090750031072, 1.8227458427912087e-17, 1.8142521755003989, -9.969531538920349e-17, 1.8340080864093424, \
  3.283107224245627e-17, 1.8539791250833855, 9.761887490727594e-17, 1.8741676341103, -6.122763413004143e-17, \
  1.8945759815869656, 3.4034035352165297e-17, 1.9152065613971474, -1.0619946056195963e-16, 1.9360617934922943, \
  1.0332385960676326e-16, 1.9571441241754002, 8.960767791036668e-17, 1.978456026387951, 4.0388753109278167e-17
#if defined(__CUDACC__)
static __device__ __align__(16) const double grt_exp_tab_dev[128] = {GRT_EXP_TAB};  // 2^(j/64) as hi, lo pairs
#endif
GRT_HD double exp_fast(double x) {
#if !defined(__CUDA_ARCH__)
  static const double grt_exp_tab_host[128] = {GRT_EXP_TAB};
#endif
  x = (x < -800.0) ? -800.0 : x;  // (keeps NaN)
  double t = fma(x, 92.33248261689366, 6755399441055744.0);  // round(x 64/ln2) in the low mantissa bits
  int n;
#if defined(__CUDA_ARCH__)
  n = __double2loint(t);
  const double2 T2 = __ldg(reinterpret_cast<const double2 *>(grt_exp_tab_dev) + (n & 63));
  const double T = T2.x, Tl = T2.y;
#else
  {
    long long b;
    memcpy(&b, &t, 8);
    n = (int)(unsigned)(b & 0xffffffffLL);
  }
  const double T = grt_exp_tab_host[2 * (n & 63)], Tl = grt_exp_tab_host[2 * (n & 63) + 1];
#endif
  const double nf = t - 6755399441055744.0;
  double r = fma(nf, -0.010830424696249145, x);   // ln2/64 in two parts
  r = fma(nf, -3.623510646634843e-19, r);
  double p = fma(r, 1.0 / 120.0, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  const double q = fma(p, r * r, r);
  const double y = T + fma(T, q, Tl);
  const int k = n >> 6, k1 = k >> 1, k2 = k - k1;
  double s1, s2;
#if defined(__CUDA_ARCH__)
  s1 = __hiloint2double((k1 + 1023) << 20, 0);
  s2 = __hiloint2double((k2 + 1023) << 20, 0);
#else
  {
    long long b1 = (long long)(k1 + 1023) << 52, b2 = (long long)(k2 + 1023) << 52;
    memcpy(&s1, &b1, 8);
    memcpy(&s2, &b2, 8);
  }
#endif
  return (y * s1) * s2;
}

// log(x) for finite x >= 1 (the Faraday-rotation fit takes log(1 + positive) once per (point, frequency)): the argument
// reduction and degree-7 minimax polynomial of the classic s = f/(2+f) scheme (x = 2^k (1+f), sqrt(1/2) < 1+f < sqrt 2;
// log(1+f) = f - (f^2/2 - s (f^2/2 + R(s^2)))), without libm's branches for zero / negative / subnormal / infinite arguments.
// Max relative error 0.6 ulp (checked against mpmath over 1 .. 1e12, tests/test_oracle_math.py).
GRT_HD double log_ge1(double x) {
  if (!(x <= 8.9e307)) return log(x);  // inf / NaN
  long long b;
#if defined(__CUDA_ARCH__)
  b = __double_as_longlong(x);
#else
  memcpy(&b, &x, 8);
#endif
  int hx = (int)(b >> 32);
  int k = (hx >> 20) - 1023;
  hx &= 0x000fffff;
  const int i = (hx + 0x95f64) & 0x100000;  // mantissa above sqrt 2: halve it
  k += i >> 20;
  b = ((long long)(hx | (i ^ 0x3ff00000)) << 32) | (b & 0xffffffffLL);
  double m;
#if defined(__CUDA_ARCH__)
  m = __longlong_as_double(b);
#else
  memcpy(&m, &b, 8);
#endif
  const double f = m - 1.0;
#if defined(__CUDA_ARCH__)
  const double s = f * __drcp_rn(2.0 + f);
#else
  const double s = f / (2.0 + f);
#endif
  const double z = s * s, w = z * z;
  const double t1 = w * (3.999999999940941908e-01 + w * (2.222219843214978396e-01 + w * 1.531383769920937332e-01));
  const double t2 = z * (6.666666666666735130e-01 +
                         w * (2.857142874366239149e-01 + w * (1.818357216161805012e-01 + w * 1.479819860511658591e-01)));
  const double R = t2 + t1, hfsq = 0.5 * f * f, dk = (double)k;
  return dk * 6.93147180369123816490e-01 - ((hfsq - (s * (hfsq + R) + dk * 1.90821492927058770002e-10)) - f);
}

// ---- Carlson R_F ------------------------------------------------------------------------------
GRT_HDN double rf(double x, double y, double z) {
  double ave, dx, dy, dz;
  do {
    double sx = sqrt(x), sy = sqrt(y), sz = sqrt(z);
    double lam = sx * (sy + sz) + sy * sz;
    x = 0.25 * (x + lam);
    y = 0.25 * (y + lam);
    z = 0.25 * (z + lam);
    ave = (1.0 / 3.0) * (x + y + z);
    if (ave == 0.0) {
      dx = dy = dz = 0.0;
    } else {  // one reciprocal instead of three divisions (the deviations only feed the 1e-3-sized series correction)
      const double iave = 1.0 / ave;
      dx = (ave - x) * iave;
      dy = (ave - y) * iave;
      dz = (ave - z) * iave;
    }
  } while (dmax(dmax(fabs(dx), fabs(dy)), fabs(dz)) > 0.0025);
  double e2 = dx * dy - dz * dz, e3 = dx * dy * dz;
  return (1.0 + ((1.0 / 24.0) * e2 - 0.1 - (3.0 / 44.0) * e3) * e2 + (1.0 / 14.0) * e3) / sqrt(ave);
}

// ---- Carlson R_C ------------------------------------------------------------------------------
// The reference iterates the duplication theorem (geokerr_wrapper.f:3608, ERRTOL=.0012, ~5 iterations with two square
// roots each).  R_C is elementary, R_C(x,y) = atan(t)/(t sqrt x) for y > x, atanh(t)/(t sqrt x) for 0 < y < x with
// t = sqrt(|y-x|/x) (DLMF 19.2.17-18), and it is called ~27 times per geodesic point (once per R_J iteration), so the closed
// form is used: same value to round-off (the series is truncated at ~1e-17), about half the instructions.
// R_C(x, y) for y > 0 with sx = sqrt(x) supplied by the caller (R_J knows it without a square root), scaled by w
GRT_HDN double rc_pos(double x, double sx, double y, double w = 1.0) {
  const double d = y - x;
  if (d == 0.0) return w / sx;
  if (x == 0.0) return w * (0.5 * GRT_PI) / sqrt(y);
  const double sd = sqrt(fabs(d));
  const double t = sd / sx;
  if (t < 0.03) {  // atan(t)/t, atanh(t)/t by their series: exact to < 1e-17 and free of cancellation
    const double q = (d > 0.0) ? -t * t : t * t;
    const double f = 1.0 + q * (1.0 / 3.0 + q * (0.2 + q * (1.0 / 7.0 + q * (1.0 / 9.0 + q * (1.0 / 11.0)))));
    return w * f / sx;
  }
  if (d > 0.0) return w * atan(t) / sd;            // atan(t) / (t sqrt x)
  return w * log((sx + sd) / sqrt(y)) / sd;        // atanh(t) = log((sqrt x + sqrt(x-y)) / sqrt y): no cancellation for y << x
}
GRT_HD double rc(double x, double y) {
  double w = 1.0;
  if (!(y > 0.0)) {  // Cauchy principal value for y < 0
    double xt = x - y;
    w = sqrt(x) / sqrt(xt);
    x = xt;
    y = -y;
  }
  return rc_pos(x, sqrt(x), y, w);
}

// ---- Carlson R_D ------------------------------------------------------------------------------
GRT_HDN double rd(double x, double y, double z) {
  const double C1 = 3.0 / 14.0, C2 = 1.0 / 6.0, C3 = 9.0 / 22.0, C4 = 3.0 / 26.0, C5 = 0.25 * C3, C6 = 1.5 * C4;
  double sum = 0.0, fac = 1.0, ave, dx, dy, dz;
  do {
    double sx = sqrt(x), sy = sqrt(y), sz = sqrt(z);
    double lam = sx * (sy + sz) + sy * sz;
    sum += fac / (sz * (z + lam));
    fac *= 0.25;
    x = 0.25 * (x + lam);
    y = 0.25 * (y + lam);
    z = 0.25 * (z + lam);
    ave = 0.2 * (x + y + 3.0 * z);
    const double iave = 1.0 / ave;
    dx = (ave - x) * iave;
    dy = (ave - y) * iave;
    dz = (ave - z) * iave;
  } while (dmax(dmax(fabs(dx), fabs(dy)), fabs(dz)) > 0.0015);
  double ea = dx * dy, eb = dz * dz, ec = ea - eb, ed = ea - 6.0 * eb, ee = ed + ec + ec;
  return 3.0 * sum +
         fac * (1.0 + ed * (-C1 + C5 * ed - C6 * dz * ee) + dz * (C2 * ee + dz * (-C3 * ec + dz * C4 * ea))) /
             (ave * sqrt(ave));
}

// ---- Carlson R_J (p<=0 handled through the Cauchy principal value transformation) ---------------
GRT_HDN double rj(double x, double y, double z, double p) {
  const double C1 = 3.0 / 14.0, C2 = 1.0 / 3.0, C3 = 3.0 / 22.0, C4 = 3.0 / 26.0, C5 = 0.75 * C3, C6 = 1.5 * C4,
               C7 = 0.5 * C2, C8 = C3 + C3;
  double a = 0.0, b = 0.0, rcx = 0.0;
  const bool neg = !(p > 0.0);
  if (neg) {
    double lo = dmin(dmin(x, y), z), hi = dmax(dmax(x, y), z);
    double mid = x + y + z - lo - hi;
    a = 1.0 / (mid - p);
    b = a * (hi - mid) * (mid - lo);
    double pt = mid + b;
    rcx = rc(lo * hi / mid, p * pt / mid);
    x = lo;
    y = mid;
    z = hi;
    p = pt;
  }
  double sum = 0.0, fac = 1.0, ave, dx, dy, dz, dp;
  do {
    double sx = sqrt(x), sy = sqrt(y), sz = sqrt(z);
    double lam = sx * (sy + sz) + sy * sz;
    double sal = p * (sx + sy + sz) + sx * sy * sz;  // alpha = sal^2: its square root is known (sal > 0 since p > 0)
    double be = p * sq(p + lam);
    sum += fac * rc_pos(sal * sal, sal, be);
    fac *= 0.25;
    x = 0.25 * (x + lam);
    y = 0.25 * (y + lam);
    z = 0.25 * (z + lam);
    p = 0.25 * (p + lam);
    ave = 0.2 * (x + y + z + p + p);
    const double iave = 1.0 / ave;
    dx = (ave - x) * iave;
    dy = (ave - y) * iave;
    dz = (ave - z) * iave;
    dp = (ave - p) * iave;
  } while (dmax(dmax(fabs(dx), fabs(dy)), dmax(fabs(dz), fabs(dp))) > 0.0015);
  double ea = dx * (dy + dz) + dy * dz, eb = dx * dy * dz, ec = dp * dp, ed = ea - 3.0 * ec,
         ee = eb + 2.0 * dp * (ea - ec);
  double r = 3.0 * sum + fac *
                             (1.0 + ed * (-C1 + C5 * ed - C6 * ee) + eb * (C7 + dp * (-C8 + dp * C4)) +
                              dp * ea * (C2 - dp * C3) - C2 * dp * ec) /
                             (ave * sqrt(ave));
  if (neg) r = a * (b * r + 3.0 * (rcx - rf(x, y, z)));
  return r;
}

// ---- Jacobi elliptic functions sn, cn, dn for parameter complement emmc = 1-k^2 ------------------
GRT_HDN void sncndn(double u, double emc, double &sn, double &cn, double &dn) {
  if (emc == 0.0) {
    cn = 1.0 / cosh(u);
    dn = cn;
    sn = tanh(u);
    return;
  }
  double em[13], en[13];
  const bool bo = emc < 0.0;
  double d = 1.0;
  if (bo) {
    d = 1.0 - emc;
    emc = -emc / d;
    d = sqrt(d);
    u = d * u;
  }
  double a = 1.0, c = 0.0;
  dn = 1.0;
  int l = 0;
  for (int i = 0; i < 13; i++) {
    l = i;
    em[i] = a;
    emc = sqrt(emc);
    en[i] = emc;
    c = 0.5 * (a + emc);
    if (fabs(a - emc) <= 3.e-8 * a) break;
    emc = a * emc;
    a = c;
  }
  u = c * u;
  sn = sin(u);
  cn = cos(u);
  if (sn != 0.0) {
    a = cn / sn;
    c = a * c;
    for (int i = l; i >= 0; i--) {
      double b = em[i];
      a = c * a;
      c = dn * c;
      dn = (en[i] + a) / (b + a);
      a = c / b;
    }
    a = 1.0 / sqrt(c * c + 1.0);
    sn = (sn < 0.0) ? -a : a;
    cn = c * sn;
  }
  if (bo) {
    a = dn;
    dn = cn;
    cn = a;
    sn = sn / d;
  }
}

// ---- minimal complex arithmetic (naive multiply, Smith divide: Fortran rules) --------------------
struct cpx {
  double re, im;
};
GRT_HD cpx cmk(double r, double i) {
  cpx z;
  z.re = r;
  z.im = i;
  return z;
}
GRT_HD cpx cadd(cpx a, cpx b) { return cmk(a.re + b.re, a.im + b.im); }
GRT_HD cpx csub(cpx a, cpx b) { return cmk(a.re - b.re, a.im - b.im); }
GRT_HD cpx cmul(cpx a, cpx b) { return cmk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
GRT_HD cpx cscale(cpx a, double s) { return cmk(a.re * s, a.im * s); }
GRT_HD double cabs_(cpx a) { return hypot(a.re, a.im); }
GRT_HD cpx cdiv(cpx a, cpx b) {
  if (fabs(b.re) >= fabs(b.im)) {
    double r = b.im / b.re, den = b.re + r * b.im;
    return cmk((a.re + a.im * r) / den, (a.im - a.re * r) / den);
  }
  double r = b.re / b.im, den = b.im + r * b.re;
  return cmk((a.re * r + a.im) / den, (a.im * r - a.re) / den);
}
GRT_HD cpx csqrt_(cpx z) {  // principal branch
  double m = cabs_(z);
  if (m == 0.0) return cmk(0.0, 0.0);
  double t = sqrt(0.5 * (m + fabs(z.re)));
  if (z.re >= 0.0) return cmk(t, z.im / (2.0 * t));
  return cmk(fabs(z.im) / (2.0 * t), (signbit(z.im) ? -t : t));
}

// Laguerre iteration on a polynomial with coefficients A[0..M] (A[0] constant term)
// (Numerical-Recipes LAGUER as used by the reference, geokerr_wrapper.f:3533-3606)
GRT_HDN void laguer(const cpx *A, int M, cpx &X) {
  const double frac[9] = {0., 0.5, 0.25, 0.75, 0.13, 0.38, 0.62, 0.88, 1.0};
  for (int it = 1; it <= 80; it++) {
    cpx B = A[M], D = cmk(0., 0.), F = cmk(0., 0.);
    double err = cabs_(B), abx = cabs_(X);
    for (int j = M - 1; j >= 0; j--) {
      F = cadd(cmul(X, F), D);
      D = cadd(cmul(X, D), B);
      B = cadd(cmul(X, B), A[j]);
      err = cabs_(B) + abx * err;
    }
    err *= 1.e-15;
    if (cabs_(B) <= err) return;
    cpx G = cdiv(D, B);
    cpx G2 = cmul(G, G);
    cpx H = csub(G2, cscale(cdiv(F, B), 2.0));
    cpx SQ = csqrt_(cscale(csub(cscale(H, (double)M), G2), (double)(M - 1)));
    cpx GP = cadd(G, SQ), GM = csub(G, SQ);
    double abp = cabs_(GP), abm = cabs_(GM);
    if (abp < abm) GP = GM;
    cpx DX;
    if (dmax(abp, abm) > 0.0) {
      DX = cdiv(cmk((double)M, 0.0), GP);
    } else {
      float mag = r4exp((float)log(1.0 + abx));
      DX = cmk((double)(mag * r4cos((float)it)), (double)(mag * r4sin((float)it)));
    }
    cpx X1 = csub(X, DX);
    if (X.re == X1.re && X.im == X1.im) return;
    if (it % 10 != 0)
      X = X1;
    else
      X = csub(X, cscale(DX, frac[it / 10]));
  }
}

// All roots of a real-coefficient polynomial of degree M<=6, sorted by real part
// (ZROOTS with polishing, geokerr_wrapper.f:3348-3441, including its start-point heuristic).
GRT_HDN void zroots(const double *coef, int M, cpx *roots) {
  cpx A[7], AD[7];
  for (int j = 0; j <= M; j++) {
    A[j] = cmk(coef[j], 0.0);
    AD[j] = A[j];
  }
  cpx xstart = cmk(0., 0.);
  int ncc = 0;
  for (int j = M; j >= 1; j--) {  // j = current degree; root index j-1
    cpx X = xstart;
    if (j < M) {
      if (roots[j].im != 0.0 && ncc == 0) {
        xstart = cmk(roots[j].re, -roots[j].re);
        ncc = 1;
      } else {
        xstart = cmk(0., 0.);
        ncc = 0;
      }
    } else {
      xstart = cmk(0., 0.);
      ncc = 0;
    }
    laguer(AD, j, X);
    if (fabs(X.im) <= 2.0 * 1e-6 * 1e-6 * fabs(X.re)) X.im = 0.0;
    roots[j - 1] = X;
    cpx B = AD[j];
    for (int jj = j - 1; jj >= 0; jj--) {
      cpx C = AD[jj];
      AD[jj] = B;
      B = cadd(cmul(X, B), C);
    }
  }
  for (int j = 0; j < M; j++) laguer(A, M, roots[j]);
  for (int j = 1; j < M; j++) {  // insertion sort on the real part
    cpx X = roots[j];
    int i = j - 1;
    while (i >= 0 && roots[i].re > X.re) {
      roots[i + 1] = roots[i];
      i--;
    }
    roots[i + 1] = X;
  }
}

}  // namespace grt
