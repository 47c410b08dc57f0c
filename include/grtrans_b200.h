/* grtrans_b200 -- C ABI of the B200-native grtrans per-pixel hot path.
 *
 * This is the drop-in boundary.  Every entry point replaces one interface of the
 * reference (jadexter/grtrans; file:line relative to the reference tree) and is what the reference's
 * own FFI would bind: plain pointers and sizes, no C++/torch types.  INTEGRATION.md shows the
 * Fortran `bind(C)` interface module and the Python ctypes stub that bind these symbols.
 *
 * All compute entry points run on the GPU (sm_100a).  There is no CPU fallback: when no CUDA device
 * is usable they return GRT_ERR_NO_DEVICE and write nothing.
 */
#ifndef GRTRANS_B200_H
#define GRTRANS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define GRT_OK 0
#define GRT_ERR_NO_DEVICE 1   /* no usable CUDA device / kernel launch failure               */
#define GRT_ERR_BAD_ARG 2     /* unsupported fname / ename / iname / argument combination     */
#define GRT_ERR_IO 3          /* input file missing or malformed, output not writable         */
#define GRT_ERR_STATE 4       /* grt_main called while data exist (pgrtrans "data already exist") */
#define GRT_ERR_CUDA 5

#define GRT_STRLEN 100
#define GRT_PATHLEN 300

/* Argument list of  pgrtrans.grtrans_main  (pgrtrans.f90:12-58), same names, same meaning, same
 * order.  Strings are NUL-terminated (the Fortran shim trims the blank padding). */
typedef struct grt_args {
  int standard;
  double mumin, mumax;
  int nmu;
  double phi0, spin, uout, uin, rcut;
  int nrotype;
  double gridvals[4];
  int nn[3];
  int i1, i2;
  char fname[GRT_STRLEN];
  double dt;
  int nt, nload, nmdot;
  double mdotmin, mdotmax;
  char ename[GRT_STRLEN];
  double mbh;
  int nfreq;
  double fmin, fmax, muval, gmin, gmax, p1, p2, jetalpha;
  char stype[GRT_STRLEN];
  int use_geokerr;
  int nvals;
  char iname[GRT_STRLEN];
  int cflag, extra, debug;
  char outfile[GRT_PATHLEN];
  char fdfile[GRT_PATHLEN], fhfile[GRT_PATHLEN], fgfile[GRT_PATHLEN], fsim[GRT_STRLEN];
  int fnt, findf, fnfiles, fjonfix, fnw, fnfreq_tab, fnr, foffset, fdindf, fmagcrit;
  double frspot, fr0spot, fn0spot, ftscl, frscl, fwmin, fwmax, ffmin, ffmax, frmax, fsigt, ffcol, fmdot, fnscl,
      fnnthscl, fnnthp, fbeta;
  int fbl06;
  double fnp, ftp, frin, frout, fthin, fthout, fphiin, fphiout, fscalefac, sigcut;
  int epcoefindx[7];
  int nepotherargs;
  double epotherargs[8];
} grt_args;

/* Fill *a with the defaults of grtrans_batch.py:grtrans_inputs (grtrans_batch.py:24-150). */
void grt_args_defaults(grt_args *a);

/* ---- pgrtrans module surface (pgrtrans.f90:6-8, 12-58, 255-258) -------------------------------- */
/* grtrans_main: trace i1..i2, all mu0 / mdot / freq cameras.  Results stay owned by the library
 * until grt_del_data() (like the module variables ivals/ab/freqs).
 * Multi-GPU: the geodesic index range is sharded over every selected device (grt_set_devices, else the
 * environment GRT_DEVICES="0,2,3" / GRT_NGPU=4, else all visible devices), one host thread per device in
 * place of the reference's OpenMP pixel loop (pgrtrans.f90:198-209); every finished tile streams into the
 * single host image.  The result does not depend on the number of devices (bit-identical). */
int grt_main(const grt_args *a);
/* devices grt_main uses (n = 0: back to the default) */
int grt_set_devices(const int *devices, int n);
/* after grt_main: devices used, host->device / device->host bytes of that call, kernels launched */
int grt_main_info(int *ndevices, long long *h2d_bytes, long long *d2h_bytes, int *launches);
/* grt_main pools its device sessions, pinned buffers and the ivals array between calls with identical
 * arguments (grt_del_data only marks them free, like deallocate); this gives the memory back.
 * GRT_CACHE=0 in the environment disables the pool. */
void grt_release_cache(void);
/* ivals(nvals+nextra, npix, ncams) column-major exactly like the Fortran module array.
 * dims = {nvals+nextra, npix, ncams}.  Returns GRT_ERR_STATE when nothing is allocated. */
int grt_get_ivals(const double **ivals, int dims[3]);
int grt_get_ab(const double **ab, int dims[2]);     /* ab(2, npix)      */
int grt_get_freqs(const double **freqs, int *nfreq); /* freqs(nfreq), Hz */
int grt_data_allocated(void);                        /* allocated(ivals) */
void grt_del_data(void);                             /* del_pgrtrans_data */

/* ---- CLI / file surface (grtrans.f90:2-52, grtrans_program.f90, read_inputs.f90:47-100) -------- */
/* grtrans_run(ifile, outfile): read the six namelists from ifile, run, write outfile
 * (cflag=1 FITS, else Fortran sequential unformatted; camera.f90:156-344), free the data. */
int grt_run(const char *ifile, const char *outfile);
/* program call_grtrans: read &files ifile, ofile / from files_in then grt_run. */
int grt_program(const char *files_in);
/* read_inputs: parse a namelist file into a grt_args (no compute). */
int grt_read_inputs(const char *ifile, grt_args *a);
/* kwrite_raytrace_camera for data already held by the library. */
int grt_write_output(const grt_args *a, const char *outfile);

/* ---- kernel-level surfaces used by the reference's own unit boundaries ------------------------- */
/* class_geokerr.f90:69-97,128-144 / geodesics.f90:64-111: camera pixel -> constants of motion.
 * All output arrays have nro*nphi entries. */
int grt_geokerr_camera(int standard, double mu0, double a, double uout, double rcut, int nrotype, const double gridvals[4],
                       int nro, int nphi, int nup, double *u0, double *offset, double *alpha, double *beta,
                       double *q2, double *l, double *su, double *sm, double *uf, double *muf, int *tpm, int *tpr);
/* geodesics.f90:206-377 on the GPU for geodesic indices i1..i2 (1-based, inclusive) of the camera above.
 * maxpts = nup + next - 2*kext.  Outputs are [nrays][maxpts] row-major (ray-major), rays with fewer
 * points are zero padded; npts[nrays] holds the per-ray count (0 = rejected ray). */
int grt_geokerr_rays(int standard, double mu0, double phi0, double a, double uout, double rcut, int nrotype,
                     const double gridvals[4], int nro, int nphi, int nup, int i1, int i2, int *maxpts_out,
                     int *npts, double *lambda, double *t, double *r, double *th, double *ph, double *kt,
                     double *kr, double *kth, double *kph, int *tpm, int *tpr, int cap_pts);
/* debug = 1 (grtrans_driver.f90:86-110, 341-427): writes the text file the reference leaves behind as `geodebug.out` --
 * the per-point internals of ONE ray (geodesic, fluid sample, tetrad, emission / absorption / rotation coefficients of
 * the first (mdot, frequency), optical depth, intensities along the ray), one record per line in the order
 * read_geodebug_file.py reads them.  pixel = 1-based geodesic index.  grt_main / grt_run call it with the last pixel of
 * the run and path "geodebug.out" when args.debug == 1 (the reference's pixel loop overwrites the file pixel by pixel). */
int grt_geodebug(const grt_args *a, int pixel, const char *path);
/* Stage dump of the frequency-independent per-point quantities for rays i1..i2 of a full setup:
 * fluid sample (fluid.f90:473-526), comoving_ortho (kerr.f90:586-730) and unit conversion
 * (fluid.f90:528-584).  Arrays [nrays][maxpts]; u and b are [nrays][maxpts][4]. */
int grt_ray_stage_dump(const grt_args *a, int i1, int i2, int cap_pts, int *maxpts_out, int *npts, double *lambda,
                       double *r, double *th, double *ph, float *rho, float *p, float *bmag, double *u, double *b,
                       double *s2xi, double *c2xi, double *ang, double *rshift, double *cosne, double *ncgs,
                       double *ncgsnth, double *bcgs, double *tcgs);
/* polsynchemis f2py surface (Makefile:198-199; polsynchemis.f90:527,633,700,865): n points, e is
 * [n][11] row-major = (jI,jQ,jU,jV,aI,aQ,aU,aV,rhoQ,rhoU,rhoV). */
int grt_polsynchth(int n, const double *nu, const double *ncgs, const double *b, const double *t, const double *theta,
                   double *e);
int grt_polsynchpl(int n, const double *nu, const double *nnth, const double *b, const double *theta, const double *p,
                   const double *gmin, double gmax, double *e);
int grt_synchemis(int n, const double *nu, const double *ncgs, const double *b, const double *t, double *e);
/* radtrans_integrate f2py surface (radtrans_integrate.f90:68,139,884; unit_tests_integration.py:52-55).
 * nrays independent rays of npts points each: s[nrays][npts], j[nrays][npts][neq], K[nrays][npts][nk],
 * tau[nrays][npts]; out intensity[nrays][npts][neq] (column k = intensity(:,k)), nptsout[nrays]. */
int grt_rt_integrate(int iflag, int neq, int nrays, int npts, const double *s, const double *j, const double *K,
                     const double *tau, double *intensity, int *nptsout, double max_tau, double hmax, double oatol,
                     double ortol, double thin, int maxsteps);

/* ---- resident-session surface (benchmarking / multi-GPU) ---------------------------------------- */
/* A session keeps camera constants, fluid snapshot and image on the device so that the per-pixel path
 * can be re-run with nothing but kernel launches (BASELINE metric: rays*freq / s).  rank/nranks shard
 * the geodesic index range i1..i2 in interleaved tiles; each rank holds its own rays. */
typedef struct grt_session grt_session;
int grt_session_create(const grt_args *a, int device, int rank, int nranks, grt_session **out);
void grt_session_destroy(grt_session *s);
/* launch the whole per-pixel path for this rank's rays on the session stream (asynchronous) */
int grt_session_launch(grt_session *s);
int grt_session_sync(grt_session *s);
/* launch + stream each finished tile of the image into a session-owned PINNED host buffer on a second stream (the
 * device->host read-back overlaps the computation of the following tiles).  *host_image: float32
 * [ncams][nrays_local][nvals], complete after grt_session_sync(), valid until the next call or grt_session_destroy;
 * *bytes = its size */
int grt_session_launch_fetch(grt_session *s, const float **host_image, long long *bytes);
/* host->device half of an end-to-end step: re-copy camera constants, fluid snapshot and frequency grid from the
 * session's host buffers; *bytes = bytes copied */
int grt_session_upload(grt_session *s, long long *bytes);
/* nsteps back-to-back passes bracketed by CUDA events on the session stream; *ms = elapsed device time */
int grt_session_run_timed(grt_session *s, int nsteps, float *ms);
/* one launch bracketed by CUDA events on the session stream; ms for the whole path and per stage
 * (stage_ms[0..3] = ray setup, geodesic points, fluid+ortho, emissivity+transfer) */
int grt_session_time(grt_session *s, float *total_ms, float stage_ms[4]);
/* number of kernel launches one grt_session_launch() issues */
int grt_session_launch_count(const grt_session *s);
/* kernels launched by the last pass (a gather pass adds rank 0's scatter kernels); rays per tile */
int grt_session_last_launches(const grt_session *s);
int grt_session_tile(const grt_session *s);
/* local rays of this rank and their total number of geodesic points (after a launch+sync) */
int grt_session_counts(grt_session *s, long long *nrays_local, long long *npoints_local);
/* ---- multi-GPU with one process per GPU (torchrun / mpirun): NCCL inside the library ------------ */
/* geodesic indices (1-based) of rank's rays: the 256-ray shards rank, rank+nranks, ... of i1..i2
 * (replaces the schedule(static,1) of pgrtrans.f90:198).  Returns the count; ray_index may be NULL.
 * Pure host logic (no device needed). */
long long grt_shard_rays(int i1, int i2, int rank, int nranks, int *ray_index);
/* rank 0 makes the NCCL unique id (128 bytes); the launcher hands it to every rank */
int grt_nccl_unique_id(char id[128]);
/* collective: ncclCommInitRank over the nranks sessions created with the same arguments */
int grt_session_comm_init(grt_session *s, const char id[128]);
/* collective: launch the whole path; every finished tile goes to rank 0 with ncclSend/ncclRecv on a
 * second stream, behind the kernels of the following tiles, and is scattered into rank 0's full
 * device image [ncams][npix][nvals] (zero outside i1..i2).  nranks = 1 needs no communicator. */
int grt_session_launch_gather(grt_session *s);
/* grt_session_run_timed with gather passes; the closing event waits for this rank's share of the gather */
int grt_session_run_timed_gather(grt_session *s, int nsteps, float *ms);
/* rank 0: the gathered device image (complete after grt_session_sync on every rank) */
int grt_session_full_image(grt_session *s, const float **d_full, long long *nelem);
/* rank 0: synchronous copy of the gathered image to host memory, float32 [ncams][npix][nvals] */
int grt_session_fetch_full(grt_session *s, float *image);
/* copy the local image to host (float32 [ncams][nrays_local][nvals]) and the ray indices */
int grt_session_fetch(grt_session *s, float *image, int *ray_index);
int grt_session_ncams(const grt_session *s);
/* values per pixel of the image: nvals, or nvals + 19 when extra = 1 (pgrtrans.f90:107-113, ivals(nvals+nextra,:,:) :221) */
int grt_session_nvals(const grt_session *s);

/* misc */
const char *grt_version(void);
int grt_device_count(void);
/* FP64 FMA micro-benchmark on the current device: returns achieved DFMA TFLOP/s (2 flop per FMA) */
int grt_dfma_peak(int device, double *tflops);
/* Test hook: the transfer loop's own elementary functions as the device evaluates them (which = 0: exp, 1: log for
 * x >= 1), n values. */
int grt_math_probe(int which, int n, const double *x, double *out);

#ifdef __cplusplus
}
#endif
#endif
