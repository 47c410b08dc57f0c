// ORACLE -- TEST INFRASTRUCTURE ONLY (see geokerr.cc header).  CPU restatement of the grtrans
// per-pixel hot path used as the parity checker for the CUDA product path.
#pragma once
#include <complex>
#include <string>
#include <vector>

namespace orc {

// physical constants exactly as the reference defines them (phys_constants.f90:3-6; NOT CODATA).  They are variables
// (defined in driver.cc with the current phys_constants.f90 values) so that tests can re-run the oracle with the
// constant set of an OLDER reference commit when pinning against goldens that predate phys_constants.f90
// (orc_set_option, capi.cc): the product path always uses the current values.
namespace pc {
extern double h, k, c, e, G, m, mp, c2, sigb, msun, sigt;
extern const double pi;  // acos(-1d0)
}  // namespace pc

// switches for bisecting goldens written by older reference commits (all 0 = the reference as it is today)
struct LegacyOptions {
  int kstarq_sign_old = 0;   // polsynchpl: rho_Q with the sign it had before the "JAD changing sign here" edit (polsynchemis.f90:583-585)
  double pl_rhoq_scale = 1.0, pl_rhov_scale = 1.0, pl_jv_scale = 1.0, pl_av_scale = 1.0, pl_ai_scale = 1.0, pl_aq_scale = 1.0;  // bisecting only
  int delo_inverse_lu = 0;  // DELO: 4x4 inverse by pivoted elimination instead of cofactors (rounding-ambiguity measurements)
  double lambda_perturb_ulp = 0.0;  // noise on the affine parameter in units of ulp(1/u0) (conditioning measurements)
  double rt_perturb = 0.0;  // relative noise on the transfer coefficients (conditioning measurements)
};
extern LegacyOptions legacy;

// per-ray cached integrals: the by-reference arguments threaded through GEOMU/GEOR/GEOPHITIME
// (geokerr_wrapper.f:405-409)
struct GeoCache {
  int ncase = 0;
  double h1 = 0, u1 = 0, u2 = 0, u3 = 0, u4 = 0;
  double rffu0 = 0, rffu1 = 0, rffmu1 = 0, rffmu2 = 0, rffmu3 = 0, iu0 = 0, i1mu = 0, i3mu = 0;
  double rdc = 0, rjc = 0, tu01 = 0, tu02 = 0, tu03 = 0, tu04 = 0, tmu1 = 0, tmu3 = 0, phimu1 = 0, phimu3 = 0;
};

double rf(double x, double y, double z);
double rc(double x, double y);
double rd(double x, double y, double z);
double rj(double x, double y, double z, double p);
void sncndn(double uu, double emmc, double &sn, double &cn, double &dn);
void zroots(const std::complex<double> *A, int M, std::complex<double> *ROOTS, bool POLISH);

void geomu(double &u0, double &uf, double mu0, double &muf, double a, double l, double &l2, double q2, double &iu,
           int &tpm, int &tpr, double su, double sm, GeoCache &c, bool pht, bool firstpt);
void geor(double &u0, double &uf, double mu0, double muf, double a, double l, double l2, double q2, double &imu,
          int &tpm, int &tpr, double su, double sm, GeoCache &c, bool pht, bool firstpt);
void geophitime(double u0, double uf, double mu0, double muf, double a, double l, double l2, double q2, int tpm, int tpr,
                double su, double sm, double iu, double &phimu, double &tmu, double &phiu, double &tu, double &lambda,
                GeoCache &c, bool firstpt);
void geokerr(double u0, double uf, double uout, double mu0, double muf, double a, double l, double q2, double alpha,
             double beta, int tpm, int tpr, double su, double sm, int &nup, double offset, bool phit, bool usegeor,
             bool mufill, int &ncase_out, int kext, int npts, double *ufi, double *mufi, double *dti, double *dphi,
             int *tpmi, int *tpri, double *lambdai);
void initialize_camera_geokerr(int standard, double a1, double a2, double b1, double b2, double rcut, int nrotype,
                               int nro, int nphi, int nup, double &u0, double uout, double mu0, double &a,
                               double &offset, double *ufarr, double *mufarr, double *alarr, double *bearr,
                               double *q2arr, double *larr, double *smarr, double *suarr, int *tpmarr, int *tprarr);

// ---------------------------------------------------------------------------------------------
// geodesics.f90: geokerr_args (per-camera SoA) and geo (one ray)
struct GeokerrArgs {
  double u0 = 0, uout = 0, uin = 0, offset = 0, mu0 = 0, phi0 = 0, a = 0;
  int usegeor = 0, phit = 0, mufill = 0, kext = 0, next = 0, nup = 0;
  std::vector<double> uf, muf, su, sm, alpha, beta, l, q2, t0;
  std::vector<int> tpm, tpr;
};

struct Vec4 {
  double d[4];
};

struct Geo {
  int npts = 0;
  int ncase = 0;
  double alpha = 0, beta = 0, a = 0, mu0 = 0, q2 = 0, l = 0, su = 0, sm = 0;
  std::vector<double> lambda;
  std::vector<int> tpmarr, tprarr;
  std::vector<Vec4> x, k;  // x = (t, r, theta, phi); k = contravariant wave vector
};

void initialize_pixels(GeokerrArgs &g, int standard, double mu0, double phi0, double a, double uout, double uin,
                       double rcut, int nrotype, double a1, double a2, double b1, double b2, int nro, int nphi,
                       int nup);
// i is the 1-based geodesic index.  status: 1 ok, -1 rejected (geodesics.f90:206-257)
void initialize_geodesic(Geo &g, const GeokerrArgs &gargs, int i, int &status);

// kerr.f90
void blmetric_cov(double r, double th, double a, double *g10);
void blmetric_cov_real(float r, float th, float a, float *g10);
double dot4(const double *metric10, const double *a4, const double *b4);
void calc_polar_psi(double r, double muf, double q2, double a, double alpha, double beta, double rshift, double mus,
                    const double *p, double &c2psi, double &s2psi, double &cosne);
void calc_nullp(double q2, double l, double a, double r, double mu, double su, double smu, double *k4);
void lnrf_frame_inv_real(float vr, float vt, float omega, float r, float a, float th, float &vrl, float &vtl,
                         float &vpl);
void lnrf_frame_real(float vr, float vt, float omega, float r, float a, float th, float &vrl, float &vtl, float &vpl);
void calc_polvec(double r, double mu, const double *p4, double a, double psi, double *f4);
float calc_rms(float a);
void comoving_ortho(double r, double th, double a, double alpha, double beta, double mus, const double *u4,
                    const double *b4, const double *k4, double &s2xi, double &c2xi, double &ang, double &g,
                    double &cosne);

// ---------------------------------------------------------------------------------------------
// fluid models
enum FluidModel { F_DUMMY = 0, F_SPHACC = 1, F_THINDISK = 2, F_HARM = 12, F_FFJET = 13, F_HARM3D = 19, F_KORAL = 22, F_KORAL3D = 25, F_KORAL3D_DISK = 26, F_IHARM = 29 };

struct FluidArgs {  // subset of fluid.f90:fluid_args that the supported models read
  std::string dfile, hfile, gfile;
  int nt = 1, indf = 0;
  double mdot = 0.1, mbh = 10., rin = -1., rout = 1e8, scalefac = 1.0;
};

struct FluidData {  // module-level state of fluid_model_{thindisk,sphacc,ffjet,harm,harm3d}.f90
  int model = F_DUMMY;
  // thindisk
  float td_mbh = 0, td_mdot = 0, td_rin = 0, td_rout = 0;
  // sphacc
  std::vector<float> uvals, vvals, rvals, tvals;
  // gridded models
  int nx1 = 0, nx2 = 0, nx3 = 1, nt = 1;
  float h = 0, asim = 0, gam = 0;
  std::vector<float> x1_arr, x2_arr, x3_arr, r_arr, th_arr, ph_arr;
  std::vector<float> rho_arr, p_arr, u0_arr, vrl_arr, vpl_arr, vtl_arr, b0_arr, br_arr, bth_arr, bph_arr;
  // iharm (fluid_model_iharm.f90:13-24): real(8) module variables + flags from the header file
  std::vector<float> kela_arr;
  double hslope = 0, mks_smooth = 0, poly_xt = 0, poly_alpha = 0, startx1 = 0;
  int is_mmks = 0, eheat = 0;
  // koral / koral3d (fluid_model_koral.f90:11-29): MKS3 grid parameters, Bernoulli parameter, density scale factor
  float kh = 0, kr0 = 0, kaa = 0, kbb = 0, kpp = 0;
  double scalefac = 1.0;
  int ktype = 0, minpolecell = 4;
  std::vector<float> be_arr;
  // KORALNTH (fluid_model_koral.f90:16-30, 318-326): non-thermal electron bins nth_arr(cell, bin), real(4)
  bool doingnth = false;
  int nrelbin = 0;
  double relgammamin = 1., relgammamax = 1.;
  std::vector<float> nth_arr;  // [nt][ncell][nrelbin]
  // time axis (fluid_model_harm.f90:13-16,446-546 and the same code in the harm3d / iharm / koral modules): `nt` dump slots of
  // `ncell` cells each, slot k (0-based) = dump file indf-k; tstep / toffset are real(4) module variables (real(8) in iharm)
  size_t ncell = 0;
  int indf = 0;
  double tcur = 0;              // time stamp of the dump file read last
  std::vector<double> tdump;    // t(1:nt)
  float tstep = 20.f, toffset = 0.f;
  double tstep_d = 20., toffset_d = 0.;
  bool ok = true;
};

struct FluidPt {  // one entry of fluid.f90:type fluid
  float rho = 0, p = 0, bmag = 0, kela = 0, be = 0;
  double u[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
  std::vector<float> nnth;  // f%nnth(point, 1:nrelbin) (KORALNTH only)
};

struct SourceParams {  // fluid.f90:69-75
  double nfac = 1, bfac = 1, mbh = 1, mdot = 1, p1 = 0, p2 = 0, gmax = 0, gminval = 0, jetalphaval = 0, muval = 0,
         sigcut = 0;
  int type = 0;
};

bool load_fluid_model(FluidData &fd, const std::string &fname, double a, const FluidArgs &fa);
void advance_fluid_timestep(FluidData &fd, const std::string &fname, double a, const FluidArgs &fa, double dtobs);
void get_fluid_vars(const FluidData &fd, const Geo &g, std::vector<FluidPt> &f);
void assign_source_params(const SourceParams &sp, const std::vector<double> &ncgs, const std::vector<double> &tcgs,
                          std::vector<double> &ncgsnth, std::vector<double> &gmin);
void convert_fluid_vars(const FluidData &fd, const std::vector<FluidPt> &f, const SourceParams &sp,
                        std::vector<double> &ncgs, std::vector<double> &ncgsnth, std::vector<double> &bcgs,
                        std::vector<double> &tcgs);
// synthetic snapshot writers for tests/bench (HARM ASCII layout, harm3d arrays)
void harm_transform_and_store(FluidData &fd, const std::vector<double> &rho, const std::vector<double> &uu,
                              const std::vector<double> &ucon /*4n*/, const std::vector<double> &bcon /*4n*/);

// ---------------------------------------------------------------------------------------------
// emissivities (emis.f90 / polsynchemis.f90)
enum EmisType { E_SYNCHPL = 3, E_POLSYNCHTH = 2, E_POLSYNCHPL = 4, E_BB = 8, E_BBPOL = 9, E_SYNCHTHAV = 13,
                E_HYBRIDTHPL = 20, E_HYBRIDTH = 21, E_HYBRIDPL = 22, E_BREMS = 5, E_SYNCHTHAVNOABS = 14, E_SYNCHTHBREMS = 27, E_POLSYNCHSYMTH = 15, E_MAXJUTT = 23, E_MAXCOMP = 24, E_RHO = 12,
                E_BINS = 25, E_HYBRIDTHBINS = 26 };
int emis_type_from_name(const std::string &ename, int &neq);
void polsynchth(double nu, double n, double b, double t, double theta, double *e11);
void polsynchpl(double nu, double nnth, double b, double th, double p, double gmin, double gmax, double *e11);
void synchpl(double nu, double nnth, double b, double th, double p, double gmin, double gmax, double *e11);
void brememis(double nu, double ne, double T, double *e11);
void sympolemisth(double nu, double n, double b, double t, double theta, double *e11);
void maxjutt(double nu, double n, double b, double t, double theta, const double *args, int nargs, bool comp, double *e11);
void synchemis(double nu, double n, double B, double T, double *e11);
// synchbinemis (polsynchemis.f90:1036-1150): n[nbin] electrons per bin, bin centres / widths from emis_model_bins
void synchbinemis(double nu, const double *n, int nbin, double B, double theta, const double *gammas, const double *dgammas,
                  double *e11);
void emis_model_bins(double relgammamin, double relgammamax, int nrelbin, std::vector<double> &gammas, std::vector<double> &dgammas);
double bnu(double T, double nu);
void interp_chandra_tab24(float mu, float &I, float &del);
double besselk0(double x);
double besselk1(double x);
double besselkn(int n, double x);

// ---------------------------------------------------------------------------------------------
// transfer integration (radtrans_integrate.f90).  Arrays are row-major [npts][neq] / [npts][nk].
struct RtOptions {
  double max_tau = 10., hmax = 0.1, ortol = 1e-6, oatol = 1e-8, thin = 1e-2;
  int maxsteps = 100000;
  bool lsoda_neighbor_as_written = true;  // radtrans_aux uses index indx+1
};
// iflag: 0 lsoda(->adaptive RK), 1 delo, 2 formal.  intensity is [neq][npts] column order like the
// reference (intensity(:,k)); returns nptsout (1-based column that holds the observer value).
int rt_integrate(int iflag, int neq, int npts, const double *s, const double *j, const double *K, const double *tau,
                 double *intensity, const RtOptions &opt, int *nsteps_out = nullptr);
void tsum(int n, const double *x, const double *y, double *out);

// ---------------------------------------------------------------------------------------------
// driver (grtrans_driver.f90 + pgrtrans.f90)
struct RunArgs {
  int standard = 1, nmu = 1, nrotype = 2, nvals = 1, nfreq = 1, nmdot = 1, nt = 1, nload = 1, extra = 0, i1 = 1,
      i2 = -1;
  double dt = 5.;  // observer-time spacing of the nt frames (pgrtrans.f90:211 advance_fluid_timestep(fname,dt))
  double mumin = 0, mumax = 0, phi0 = -0.5, spin = 0, uout = 0.04, uin = 1, rcut = 1;
  double gridvals[4] = {-25, 25, -25, 25};
  int nn[3] = {100, 100, 1};
  std::string fname = "THINDISK", ename = "BB", iname = "lsoda", stype = "const";
  double mdotmin = 1, mdotmax = 1, mbh = 10, fmin = 1e17, fmax = 1e17, muval = 0.25, gmin = 100, gmax = 1e5, p1 = 3.5,
         p2 = 3.5, jetalpha = 0.02, sigcut = 1e10;
  double ftscl = 30, frscl = 6;  // nfac, bfac (pgrtrans.f90:129)
  int coefindx[7] = {1, 1, 1, 1, 1, 1, 1};
  int nepotherargs = 2;                 // ep%otherargs (pgrtrans.f90:119-120): MAXJUTT / MAXCOMP component weights
  double epotherargs[8] = {1, 1, 0, 0, 0, 0, 0, 0};
  FluidArgs fargs;
  int nthreads = 0;  // OpenMP threads for the pixel loop (0 = default)
};

struct RunResult {
  int nvals = 0, npix = 0, ncams = 0;
  std::vector<float> pixvals;  // [cam][pix][nvals]  (camera.f90 pixvals, float32)
  std::vector<float> pixloc;   // [pix][2]
  std::vector<double> freqs;
  double loop_seconds = 0;  // wall clock around the pixel loop (pgrtrans.f90:195-213)
  long long npoints = 0;    // sum over rays of npts
};

// per-ray stage dump for kernel-level parity tests
struct RayDump {
  int npts = 0, status = 0, ncase = 0;
  std::vector<double> lambda, t, r, th, ph, kt, kr, kth, kph;
  std::vector<int> tpm, tpr;
  std::vector<float> rho, p, bmag;
  std::vector<double> u, b;  // 4 per point
  std::vector<double> s2xi, c2xi, ang, rshift, cosne, ncgs, ncgsnth, bcgs, tcgs;
  std::vector<double> j, K;   // last frequency processed: [npts][4], [npts][7] after rotate/invariant/scale
  std::vector<double> tau;    // optical depth used by integrate
  std::vector<double> stokes; // [nfreq][nvals] in cgs (before float32 cast)
};

void grtrans_main(const RunArgs &args, RunResult &res);
// run the per-pixel driver for one geodesic index and keep all stage arrays (first mdot only)
void grtrans_ray_dump(const RunArgs &args, const FluidData &fd, const GeokerrArgs &gargs, int i,
                      const std::vector<double> &freqs, RayDump &d);
void compute_freqs(const RunArgs &args, std::vector<double> &freqs, std::vector<double> &mdots,
                   std::vector<double> &mu0);

}  // namespace orc
