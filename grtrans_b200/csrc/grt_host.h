// grtrans_b200 -- host-side declarations shared by the engine (CUDA) and the I/O layer.
#pragma once
#include <string>
#include <vector>

#include "../../include/grtrans_b200.h"

namespace grt_host {

// ---- camera (R1): geokerr_wrapper.f:1-354 INITIALIZE_CAMERA_GEOKERR + geodesics.f90:64-111 -----------
struct Camera {
  int standard = 1, nro = 0, nphi = 0, nup = 0;
  int usegeor = 0, phit = 1, mufill = 0, kext = 0, next = 0;
  double a = 0, mu0 = 0, phi0 = 0, uout = 0, u0 = 0, offset = 0.5;
  std::vector<double> alpha, beta, q2, l, su, sm, uf, muf;
  std::vector<int> tpm, tpr;
  int maxpts() const { return nup + next - 2 * kext; }
};
void build_camera(Camera &c, int standard, double mu0, double phi0, double a, double uout, double rcut, int nrotype,
                  const double gridvals[4], int nro, int nphi, int nup);

// ---- fluid snapshot on the host (reference module state of fluid_model_*.f90) --------------------------
struct HostCell {
  float rho, p, vrl, vtl, vpl, b0, br, bth, bph, u0, kela, be;
};
struct FluidHost {
  int model = 0, nx1 = 0, nx2 = 0, nx3 = 1;
  float h = 0, asim = 0, gam = 0;
  float td_mbh = 0, td_mdot = 0, td_rin = 0, td_rout = 0;
  double hslope = 0, mks_smooth = 0, poly_xt = 0, poly_alpha = 0, startx1 = 0;  // iharm header (fluid_model_iharm.f90:427-470)
  int is_mmks = 0, eheat = 0;
  float kh = 0, kr0 = 0, kaa = 0, kbb = 0, kpp = 0, x2mintrust = 0, x2maxtrust = 0;  // koral MKS3 grid (fluid_model_koral.f90)
  int ktype = 0;
  // KORALNTH (fluid_model_koral.f90:16-30, 318-326): non-thermal electron bins nth_arr(cell, bin), real(4)
  int nrelbin = 0;
  float relgammamin = 1.f, relgammamax = 1.f;
  std::vector<float> nth;  // [ncell][nrelbin]
  std::vector<float> ax1, axr, ax2, axth, ax3;
  std::vector<HostCell> cells;  // [nslots][ncell]
  std::vector<float> uvals, vvals, uv, tvals;
  // time axis of the gridded models (fluid_model_harm.f90:13-16, 446-546 and its copies in the harm3d / iharm / koral
  // modules): nslots = fnt dumps resident, slot k = dump file indf-k; tstep / toffset are real(4) there (real(8) in iharm)
  int nslots = 1, indf = 0;
  size_t ncell = 0;
  double tcur = 0;            // time stamp of the dump read last (NaN: the format carries none)
  std::vector<double> tdump;  // t(1:nt)
  float tstep = 20.f, toffset = 0.f;
  double tstep_d = 20., toffset_d = 0.;
};
// returns GRT_OK / GRT_ERR_IO / GRT_ERR_BAD_ARG
int load_fluid(FluidHost &f, const grt_args &a);
// advance_fluid_timestep (fluid.f90:245-281): moves the resident dumps on by the observer time dt (no-op for analytic models)
int advance_fluid(FluidHost &f, const grt_args &a, double dt);

// emissivity tables (log of the tabulated G(x) integrals; Chandrasekhar table)
struct EmisTablesHost {
  std::vector<double> gx, gy, xv, ya, yp, yv, yap, yav;
  std::vector<float> ch_mu, ch_I, ch_del;
  // Faraday-fit table: cubic Hermite pieces of jffunc(xarg) and log(1 + 0.00185777 xarg^1.50316886) on a uniform grid in
  // log(xarg) (see EmisTables in grt_emis.cuh); far_n pieces of 8 doubles
  std::vector<double> far;
  double far_u0 = 0, far_ih = 0;
  int far_n = 0;
};
void build_emis_tables(EmisTablesHost &t);
// emis_model_bins (emis.f90:884-908): bins[0..nb) = centres, bins[nb..2nb) = widths of nb logarithmic Lorentz-factor bins
void emis_model_bins(double relgammamin, double relgammamax, int nb, std::vector<double> &bins);

int fluid_model_id(const std::string &fname);                 // fluid.f90:283-377 (supported subset)
int emis_type_id(const std::string &ename, int *neq);         // emis.f90:294-367
int integrator_flag(const std::string &iname);                // rad_trans.f90:30-38

void make_grids(const grt_args &a, std::vector<double> &freqs, std::vector<double> &mdots, std::vector<double> &mu0);

// ---- namelists / output -----------------------------------------------------------------------------
int read_inputs_file(const char *ifile, grt_args *a);                      // read_inputs.f90:47-100
int read_files_in(const char *path, std::string &ifile, std::string &ofile);  // grtrans_program.f90
// camera.f90:156-344 kwrite_raytrace_camera for all cameras; ivals [ncams][npix][nv] = dble() of the real(4) pixel values
int write_output(const grt_args &a, const char *outfile, int nx, int ny, int nv, int ncams, const float *pixloc,
                 const double *ivals, const std::vector<double> &freqs);

// debug = 1: the text dump `geodebug.out` of one ray (grtrans_driver.f90:86-110, 341-427), one record per line in the
// order read_geodebug_file.py expects.  rows = the device debug buffer [nrows][maxpts] (grt_debug.cuh).
struct GeoDebugData {
  int npts = 0, maxpts = 0;
  double alpha = 0, beta = 0, su = 0, sm = 0, a = 0, mu0 = 0, fac = 0;
  std::vector<double> mdots, mbh;
  std::vector<double> lambda, t, r, th, ph, kt, kr, kth, kph, u, b, ncgs, tcgs, bcgs, rshift, ang, s2xi, c2xi, rows;
  std::vector<float> rho, p, bmag;
};
int write_geodebug(const char *path, const GeoDebugData &d);

}  // namespace grt_host
