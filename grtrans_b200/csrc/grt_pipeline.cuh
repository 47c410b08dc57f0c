// grtrans_b200 -- per-ray / per-point stage bodies shared by the CUDA kernels (grt_engine.cu) and the
// CPU debugging harness (tests/host_emul).  See grt_engine.cu for the pipeline description.
#pragma once
#include "grt_geokerr.cuh"
#include "grt_transfer.cuh"

namespace grt {

struct PointArrays {  // raw geokerr outputs, [ray][maxpts]
  double *u, *mu, *dt, *dphi, *lam;
  signed char *tpm, *tpr;
};

struct RayMeta {  // per ray after finalize (explicit padding: the struct is copied to the host as raw bytes)
  int i1, npts, status, pad;
  double lamN, dt1;
};

struct StageDump {  // optional per-point outputs for parity tests (device pointers or null)
  double *lambda, *t, *r, *th, *ph, *kt, *kr, *kth, *kph, *u, *b, *s2xi, *c2xi, *ang, *rshift, *cosne, *ncgs, *ncgsnth,
      *bcgs, *tcgs;
  float *rho, *p, *bmag;
};

struct FluidKernelArgs {
  FluidDev F;
  SourceDev sp[8];
  int nparams, etype, maxpts;
  int ncomp;                                              // MAXJUTT / MAXCOMP: number of thermal components
  double mj_w[GRT_MAXJUTT_NCOMP], mj_t[GRT_MAXJUTT_NCOMP];  // their density weights w_i and temperature factors delta^i / sum(w delta^i)
  EmisTables tab;
  double *xrec;  // extra = 1: [nparams][nrays][maxpts][8] = r, theta, phi, n, T, B, plasma beta, unused (else nullptr)
};

// in_range = false for the padding threads of the last block (they only take part in the phase barriers)
GRT_HD void point_store(const RayInv &c, int nup, int slot, size_t idx, PointArrays pa, bool in_range = true) {
  GeoPoint p;
  bool ok = geo_point(c, nup, slot, in_range && slot < c.npts, p);
  if (!in_range) return;
  if (!ok) {
    p.u = 0.0; p.mu = 0.0; p.dt = 0.0; p.dphi = 0.0; p.lam = 0.0; p.tpm = 0; p.tpr = 0;
  }
  pa.u[idx] = p.u; pa.mu[idx] = p.mu; pa.dt[idx] = p.dt; pa.dphi[idx] = p.dphi; pa.lam[idx] = p.lam;
  pa.tpm[idx] = (signed char)p.tpm; pa.tpr[idx] = (signed char)p.tpr;
}

// geodesics.f90:294-323 (trim u<0 ends, standard=2 only), :255 (status), :343-347 (reference values)
GRT_HD void ray_finalize(int npts, size_t i, int maxpts, int usegeor, PointArrays pa, RayMeta &m) {
  const double *u = pa.u + i * maxpts;
  int i1 = 1, i2 = npts;
  if (usegeor && npts > 1) {
    if (u[0] < 0.0)
      for (int ii = 2; ii <= npts; ii++)
        if (u[ii - 1] >= 0.0) {
          i1 = ii;
          break;
        }
    if (u[npts - 1] < 0.0)
      for (int ii = 1; ii <= npts - 1; ii++)
        if (u[npts - ii - 1] >= 0.0) {
          i2 = npts - ii;
          break;
        }
    npts = i2 - i1 + 1;
  }
  double sumr = 0.0;
  for (int k = 0; k < npts; k++) sumr += 1.0 / u[i1 - 1 + k];
  m.i1 = i1;
  m.pad = 0;
  m.npts = npts;
  m.status = (sumr < 0.0) ? -1 : 1;
  // (sic) LAMBDAI(g%npts): the trimmed COUNT indexes the untrimmed array
  m.lamN = (npts >= 1) ? pa.lam[i * maxpts + (npts - 1)] : 0.0;
  m.dt1 = pa.dt[i * maxpts];
}

// MAXJUTT / MAXCOMP (calc_maxjutt.f90:8-65): one ThermalPoint per component, density w_i n and temperature T delta^i/sum(w delta^i)
GRT_HDN void maxjutt_record(double *q, const FluidKernelArgs &A, int mm, const CgsVars &cg, const OrthoOut &oo, double lam) {
  q[0] = lam; q[1] = oo.g; q[2] = oo.s2xi; q[3] = oo.c2xi;
  for (int cc = 0; cc < A.ncomp; cc++) {
    ThermalPoint T;
    thermal_point(A.mj_w[cc] * cg.ncgs, cg.bcgs, cg.tcgs * A.mj_t[cc], oo.ang, oo.g, pc::G * A.sp[mm].mbh * pc::msun / pc::c2, T);
    double *qc = q + 4 + 10 * cc;
    qc[0] = T.xp13; qc[1] = T.AI; qc[2] = T.AV; qc[3] = T.hkT_g; qc[4] = T.lAx; qc[5] = T.RQ; qc[6] = T.RV0; qc[7] = T.RV1;
    qc[8] = T.ixp13; qc[9] = T.xpm32;
  }
}

// BINS / HYBRIDTHBINS on KORALNTH (fluid_model_koral.f90:262-270, fluid.f90:1149-1151): the non-thermal electron spectrum of
// the point, dn/dgamma per bin, bilinear in (theta, r) between the four cells koral_vals uses for every other field.  Kept
// out of line and written in place (the record is larger than the staging array of the common models); it repeats the
// cell location of fluid_sample's FL_KORAL branch (2-D only) instead of carrying indices and weights through the common path.
GRT_HDN void bins_record(double *q, const FluidKernelArgs &A, const CgsVars &cg, const OrthoOut &oo, double lam, double a_,
                         double rr, double thh) {
  const FluidDev &F = A.F;
  q[0] = lam; q[1] = oo.g; q[2] = oo.s2xi; q[3] = oo.c2xi;
  q[4] = cg.ncgs; q[5] = cg.bcgs; q[6] = cg.tcgs; q[7] = oo.ang;
  for (int i = 8; i < 12; i++) q[i] = 0.0;
  const int nx1 = F.nx1, nx2 = F.nx2;
  const float a = (float)a_;
  float theta = (float)thh, zr = (float)rr;
  float x1 = (float)log((double)zr - (double)F.kr0);
  float x2 = koral_x2_of_theta(theta, zr, F.kh, F.kaa, F.kbb, F.kpp);
  int lx1 = (int)floorf((x1 - F.ax1[0]) / (F.ax1[nx1 - 1] - F.ax1[0]) * (float)(nx1 - 1)) + 1;
  int ux1 = lx1 + 1;
  lx1 = (lx1 >= 1) ? lx1 : 1;
  ux1 = (ux1 <= nx1) ? ux1 : nx1;
  lx1 = lx1 > nx1 ? nx1 : lx1;
  ux1 = ux1 < 1 ? 1 : ux1;
  int lx2 = (int)floorf((x2 - F.ax2[0]) / (F.ax2[nx2 - 1] - F.ax2[0]) * (float)(nx2 - 1)) + 1;
  int ux2 = lx2 + 1;
  lx2 = (lx2 >= 1) ? lx2 : 1;
  ux2 = (ux2 <= nx2) ? ux2 : nx2;
  lx2 = lx2 > nx2 ? nx2 : lx2;
  ux2 = ux2 < 1 ? 1 : ux2;
  float rd = (zr - F.axr[lx1 - 1]) / (F.axr[ux1 - 1] - F.axr[lx1 - 1]);
  float rh = 1.f + sqrtf(1.f - a * a);
  if (F.axr[lx1 - 1] <= rh || lx1 == 1) rd = 1.f;
  float thl = koral_theta_of_x2(F.ax2[lx2 - 1], F.axr[lx1 - 1], F.kh, F.kaa, F.kbb, F.kpp);
  float thu = koral_theta_of_x2(F.ax2[ux2 - 1], F.axr[lx1 - 1], F.kh, F.kaa, F.kbb, F.kpp);
  float thone = koral_theta_of_x2(F.ax2[0], F.axr[lx1 - 1], F.kh, F.kaa, F.kbb, F.kpp);
  float td = (ux2 != lx2) ? fabsf((theta - thl) / (thu - thl)) : fabsf((theta - thl) / thone);
  const size_t nb = (size_t)F.nrelbin;
  const float *c0 = F.nth + ((size_t)(lx1 - 1) * nx2 + (lx2 - 1)) * nb, *c1 = F.nth + ((size_t)(lx1 - 1) * nx2 + (ux2 - 1)) * nb,
              *c2 = F.nth + ((size_t)(ux1 - 1) * nx2 + (lx2 - 1)) * nb, *c3 = F.nth + ((size_t)(ux1 - 1) * nx2 + (ux2 - 1)) * nb;
  // masked below the inner edge and (sic) below the first trusted polar cell only (x2 > x2mintrust, no upper bound)
  const bool keep = (x1 > F.ax1[0]) && (x2 > F.x2mintrust);
  for (int ie = 0; ie < F.nrelbin; ie++) {
    float i1 = c0[ie] * (1.f - td) + c1[ie] * td, i2 = c2[ie] * (1.f - td) + c3[ie] * td;
    float w1 = i1 * (1.f - rd) + i2 * rd;
    float v = w1 * (1.f - 0.f) + w1 * 0.f;
    q[12 + ie] = keep ? (double)v : 0.0;
  }
}

// extra = 1: the second per-point record (kept out of line: the common path must not pay registers for it)
GRT_HDN void point_xrec(double *x, int npts, const RayConst &r0, double r, double th, double ph, const double *k4,
                        const FluidSample &fs, const OrthoOut &oo, const CgsVars &cg) {
  if (npts != 1) {  // quantities averaged along the ray by the extra = 1 diagnostics (grtrans_driver.f90:262-283)
    x[0] = r; x[1] = th; x[2] = ph; x[3] = cg.ncgs; x[4] = cg.tcgs; x[5] = cg.bcgs;
    x[6] = (double)(fs.p * 2.f / (fs.bmag * fs.bmag));  // f%p*2./f%bmag**2. in real(4)
    x[7] = 0.0;
  } else {  // single (disk) point: the 7 geometric quantities of grtrans_driver.f90:298-312
    double c2psi, s2psi, cosne2;
    polar_psi(r, cos(th), r0.q2, r0.a, r0.alpha, r0.beta, oo.g, r0.mu0, k4, c2psi, s2psi, cosne2);
    x[0] = oo.s2xi; x[1] = s2psi; x[2] = cosne2; x[3] = oo.ang; x[4] = oo.cosne; x[5] = oo.c2xi; x[6] = c2psi;
    x[7] = 0.0;
  }
}

// x^mu, k^mu, fluid sample, comoving frame, unit conversion and the 128-byte point record(s)
GRT_HD void point_record(const RayConst &r0, const RayMeta &m, const FluidKernelArgs &A, int ray, int slot, int nrays,
                         PointArrays pa, double *rec, StageDump sd) {
  const int maxpts = A.maxpts;
  if (m.status != 1 || slot >= m.npts) return;
  size_t src = (size_t)ray * maxpts + (m.i1 - 1) + slot;
  double u = pa.u[src], mu = pa.mu[src];
  double r = 1.0 / u, th = acos(mu);
  // g%gk%phi0 = acos(-1.)*phi0 : pi in SINGLE precision (geodesics.f90:218)
  double ph = (double)r4acos(-1.f) * r0.phi0 - pa.dphi[src];
  if (fabs(r0.mu0) == 1.0) ph += sgn(r0.mu0) * atan2(r0.beta, r0.alpha);
  double lam, t;
  int tpm, tpr;
  if (m.npts != 1) {
    lam = m.lamN - pa.lam[src];
    t = m.dt1 - pa.dt[src] - r0.t0;
    tpm = pa.tpm[src];
    tpr = pa.tpr[src];
  } else {
    size_t s1 = (size_t)ray * maxpts;
    lam = pa.lam[s1];
    t = pa.dt[s1];
    tpm = pa.tpm[s1];
    tpr = pa.tpr[s1];
  }
  double k4[4];
  null_momentum(r0.q2, r0.l, r0.a, r, cos(th), r0.su * pm1(tpr), r0.sm * pm1(tpm), k4);
  FluidSample fs;
  // radius of the ray's first point: reference of the look-back time that picks the dump slot (nslots > 1)
  const double r1 = (A.F.nslots > 1) ? 1.0 / pa.u[(size_t)ray * maxpts + (m.i1 - 1)] : r;
  fluid_sample(A.F, r0.a, r, th, ph, t, r1, k4, fs);
  OrthoOut oo;
  comoving_ortho(r, th, r0.a, r0.alpha, r0.beta, r0.mu0, fs.u, fs.b, k4, oo);
  size_t dst = (size_t)ray * maxpts + slot;
  for (int mm = 0; mm < A.nparams; mm++) {
    CgsVars cg;
    to_cgs(A.F.model, A.sp[mm], fs.rho, fs.p, fs.bmag, fs.kela, fs.be, A.F.gam, cg);
    const int nrec = record_doubles(A.etype);
    double *q = rec + ((size_t)mm * nrays * maxpts + dst) * nrec;
    if (em_slot(A.etype) == EM_MAXJUTT || em_slot(A.etype) == EM_BINS) {  // (out of line, written in place: the record is larger than the staging array)
      if (em_slot(A.etype) == EM_MAXJUTT)
        maxjutt_record(q, A, mm, cg, oo, lam);
      else
        bins_record(q, A, cg, oo, lam, r0.a, r, th);
      if (A.xrec) point_xrec(A.xrec + ((size_t)mm * nrays * maxpts + dst) * 8, m.npts, r0, r, th, ph, k4, fs, oo, cg);
      if (mm == 0 && sd.ncgs) {
        sd.ncgs[dst] = cg.ncgs; sd.ncgsnth[dst] = cg.ncgsnth; sd.bcgs[dst] = cg.bcgs; sd.tcgs[dst] = cg.tcgs;
      }
      continue;
    }
    double o[GRT_NREC_MAX];
    for (int i = 0; i < GRT_NREC_MAX; i++) o[i] = 0.0;
    o[0] = lam; o[1] = oo.g; o[2] = oo.s2xi; o[3] = oo.c2xi;
    switch (em_slot(A.etype)) {
      case EM_HYBRIDTHPL:
      case EM_HYBRIDTH:
      case EM_HYBRIDPL: {  // thermal data as below + the power-law population's point data
        PlPoint Q;
        pl_point<true>(cg.ncgsnth, cg.bcgs, oo.ang, oo.g, Q);
        o[14] = Q.nnth; o[15] = Q.nubperp; o[16] = Q.L; o[17] = Q.itanth;
        if (A.sp[mm].stail) {  // gmin varies from point to point: its share of pl_const travels with the record
          const PlConst C = pl_const((double)(float)A.sp[mm].p1, (double)(float)cg.gmin, (double)(float)A.sp[mm].gmax);
          o[18] = C.l2gmin; o[19] = C.kq; o[20] = C.kv; o[21] = C.cA;
        }
      }  // fall through
      case EM_POLSYNCHTH: {
        ThermalPoint T;
        thermal_point(cg.ncgs, cg.bcgs, cg.tcgs, oo.ang, oo.g, pc::G * A.sp[mm].mbh * pc::msun / pc::c2, T);
        o[4] = T.xp13; o[5] = T.AI; o[6] = T.AV; o[7] = T.hkT_g; o[8] = T.lAx; o[9] = T.RQ; o[10] = T.RV0; o[11] = T.RV1;
        o[12] = T.ixp13; o[13] = T.xpm32;
      } break;
      case EM_POLSYNCHPL:
      case EM_SYNCHPL: {
        PlPoint Q;
        if (A.etype == EM_POLSYNCHPL)
          pl_point<true>(cg.ncgsnth, cg.bcgs, oo.ang, oo.g, Q);
        else
          pl_point<false>(cg.ncgsnth, cg.bcgs, oo.ang, oo.g, Q);
        o[4] = Q.nnth; o[5] = Q.nubperp; o[6] = Q.L; o[7] = Q.itanth;
        if (A.sp[mm].stail) {
          const PlConst C = pl_const((double)(float)A.sp[mm].p1, (double)(float)cg.gmin, (double)(float)A.sp[mm].gmax);
          o[8] = C.l2gmin; o[9] = C.kq; o[10] = C.kv; o[11] = C.cA;
        }
      } break;
      case EM_SYNCHTHAV:
        o[4] = cg.ncgs; o[5] = cg.bcgs; o[6] = cg.tcgs;
        break;
      case EM_POLSYNCHSYMTH: {
        SymPoint Q;
        sym_point(cg.ncgs, cg.bcgs, cg.tcgs, oo.ang, Q);
        o[4] = Q.Xs; o[5] = Q.A; o[6] = Q.S; o[7] = Q.cq; o[8] = Q.v0; o[9] = Q.v1; o[10] = Q.T; o[11] = Q.lAx;
        o[12] = Q.RQ; o[13] = Q.RV0; o[14] = Q.RV1;
      } break;
      case EM_BB:
        o[6] = cg.tcgs;
        break;
      case EM_BBPOL: {
        float cI, cdel;
        chandra_interp(A.tab, (float)oo.cosne, cI, cdel);
        o[6] = cg.tcgs; o[7] = (double)cI; o[8] = (double)cdel;
      } break;
    }
#if defined(__CUDA_ARCH__)
    double2 *q2 = reinterpret_cast<double2 *>(q);
#pragma unroll
    for (int i = 0; i < GRT_NREC_MAX / 2; i++)
      if (2 * i < nrec) q2[i] = make_double2(o[2 * i], o[2 * i + 1]);
#else
    for (int i = 0; i < nrec; i++) q[i] = o[i];
#endif
    if (A.xrec) point_xrec(A.xrec + ((size_t)mm * nrays * maxpts + dst) * 8, m.npts, r0, r, th, ph, k4, fs, oo, cg);
    if (mm == 0 && sd.ncgs) {
      sd.ncgs[dst] = cg.ncgs; sd.ncgsnth[dst] = cg.ncgsnth; sd.bcgs[dst] = cg.bcgs; sd.tcgs[dst] = cg.tcgs;
    }
  }
  if (sd.lambda) {
    sd.lambda[dst] = lam; sd.t[dst] = t; sd.r[dst] = r; sd.th[dst] = th; sd.ph[dst] = ph;
    sd.kt[dst] = k4[0]; sd.kr[dst] = k4[1]; sd.kth[dst] = k4[2]; sd.kph[dst] = k4[3];
    for (int i = 0; i < 4; i++) {
      sd.u[dst * 4 + i] = fs.u[i];
      sd.b[dst * 4 + i] = fs.b[i];
    }
    sd.s2xi[dst] = oo.s2xi; sd.c2xi[dst] = oo.c2xi; sd.ang[dst] = oo.ang; sd.rshift[dst] = oo.g; sd.cosne[dst] = oo.cosne;
    sd.rho[dst] = fs.rho; sd.p[dst] = fs.p; sd.bmag[dst] = fs.bmag;
  }
}

}  // namespace grt
