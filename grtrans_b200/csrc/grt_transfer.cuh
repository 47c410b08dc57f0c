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
//           neighbour-index coefficient model; ODEPACK is replaced by an embedded Dormand-Prince 5(4)
//           Runge-Kutta with the reference tolerances (rtol 1e-6, atol 1e-8 on fac-normalised
//           intensities, hmax 0.1) that never steps across a tabulated node.
// Because the transfer equation is linear and I_0 = 0, the reference's normalisation j/fac ... I*fac
// cancels for DELO / FORMAL / quadrature; fac is still accumulated because "fac <= 0 (or NaN) => zero
// pixel" is reference behaviour (grtrans_driver.f90:195,314-317).
#pragma once
#include "grt_emis.cuh"
#include "grt_fluid.cuh"

namespace grt {

#define GRT_NREC 16      // doubles per point record (128 B) ...
#define GRT_NREC_MAX 24  // ... 24 for the hybrid thermal + power-law emissivities (thermal and power-law point data)
GRT_HD constexpr bool em_is_hybrid(int et) { return et == EM_HYBRIDTHPL || et == EM_HYBRIDTH || et == EM_HYBRIDPL; }
GRT_HD constexpr int record_doubles(int etype) {  // (same for every em_slot member)
  return em_is_hybrid(etype) ? GRT_NREC_MAX
                             : ((etype == EM_MAXJUTT || etype == EM_MAXCOMP) ? 4 + 10 * GRT_MAXJUTT_NCOMP + 2
                                : ((etype == EM_BINS || etype == EM_HYBRIDTHBINS) ? 12 + GRT_MAXBINS : GRT_NREC));
}
// record layout: [0] lambda, [1] g (redshift), [2] sin 2xi, [3] cos 2xi, [4..] emissivity-model inputs:
//   POLSYNCHTH        [4..13] ThermalPoint (xp13, AI, AV, hkT_g, log Ax, RQ, RV0, RV1, 1/xp13, xp13^-1.5)
//   POLSYNCHPL/SYNCHPL [4] n_nth, [5] nu_B,perp, [6] log(g) + log(1.5 nu_B,perp), [7] 1/tan(angle)  (PlPoint);
//                     stype = 'tail': [8..11] the gmin-dependent constants of pl_const (2 log gmin, kq, kv, cA)
//   HYBRIDTHPL/TH/PL  [4..13] ThermalPoint, [14..17] PlPoint, [18..21] tail constants (24-double record)
//   SYNCHTHAV         [4] n, [5] B, [6] T          BB [6] T          BBPOL [6] T, [7] I(mu), [8] delta(mu)
//   POLSYNCHSYMTH     [4..14] SymPoint (Xs, A, S, cq, v0, v1, T, log Ax, RQ, RV0, RV1)
//   MAXJUTT/MAXCOMP   [4 + 10 c ..] ThermalPoint of component c < ncomp <= 7 (76-double record)
//   BINS/HYBRIDTHBINS [4] n, [5] B, [6] T, [7] field angle, [12 + b] dn/dgamma of bin b < nbins <= 64 (76-double record)

// integration modes (rad_trans.f90:30-38 iflag x neq)
enum { MODE_RK4 = 0, MODE_DELO = 1, MODE_FORMAL = 2, MODE_QUAD = 3, MODE_RK1 = 4 };

struct TransferParams {
  int etype, eneq, rneq, iflag, model, npts_is_one;
  int coefindx[7];
  double LBH, gmin, gmax, p;  // emis_params pass through float32 in the reference
  double max_tau, hmax, ortol, oatol, thin;
  int maxsteps;
  int stail;  // stype = 'tail': the power law's gmin-dependent constants come from the point record
  int ncomp, isel;  // MAXJUTT / MAXCOMP: thermal components in the record; MAXCOMP: component whose emission is kept (0: all)
  int nbins;           // BINS / HYBRIDTHBINS: Lorentz-factor bins in the record
  const double *bins;  // [2 nbins] bin centres, bin widths (emis_model_bins, emis.f90:884-908)
  EmisTables tab;
};

GRT_HD int transfer_mode(const TransferParams &P) {
  if (P.iflag != 0) return (P.rneq == 4) ? ((P.iflag == 1) ? MODE_DELO : MODE_FORMAL) : MODE_QUAD;
  return (P.rneq == 4) ? MODE_RK4 : MODE_RK1;
}

struct Coef {
  double j[4], K[7];
};

// the next point record (one 128-byte line) is pulled into L1 while the current node is evaluated: ncu showed 6.5% of the
// stall samples on the first use of the record (long scoreboard) when it was only loaded at the top of the iteration
GRT_HD void prefetch_record(const double *p) {
#if defined(__CUDA_ARCH__)
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}

// Record staging (device only): the 32 lanes of a warp march one ray, so the warp streams that ray's records -- 128 / 192
// contiguous bytes per step -- through a small ring in shared memory with cp.async, GRT_STAGE_DEPTH - 1 steps ahead of
// their use.  ncu (capture r2_kt_nf1) showed 18 % of the kernel's stall samples on the first uses of the record's 64-bit
// loads (about 550 cycles per step: the L1 prefetch hint above does not bring the line in time); staged records are there
// when the step starts and are read with broadcast shared-memory loads.
#define GRT_STAGE_DEPTH 4
#ifndef GRT_MARCH_UNROLL2
#define GRT_MARCH_UNROLL2 0
#endif
GRT_HD void stage_issue(double *slot, const double *src, int ndbl, int lane) {  // lanes 0 .. ndbl/2-1 copy 16 bytes each
#if defined(__CUDA_ARCH__)
  if (src != nullptr && lane < ndbl / 2) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(slot + 2 * lane);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + 2 * lane) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
#else
  (void)slot; (void)src; (void)ndbl; (void)lane;
#endif
}
GRT_HD void stage_wait() {  // the oldest of the GRT_STAGE_DEPTH groups in flight has landed, for every lane of the warp
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.wait_group %0;" ::"n"(GRT_STAGE_DEPTH - 1) : "memory");
  __syncwarp();
#endif
}
GRT_HD void stage_release() {
#if defined(__CUDA_ARCH__)
  __syncwarp();
#endif
}

// record -> invariant, LBH-scaled coefficients at observed frequency F.nu.  jI_raw returns the comoving un-boosted j_I used
// for the driver's `fac` (grtrans_driver.f90:183-187).  Replaces calc_emissivity + the e%j/e%K copy (emis.f90:461-571),
// rotate_emis (:797-829), invariant_emis (:831-838) and the K*LBH scaling (grtrans_driver.f90:217).
struct RunConst {  // per-thread constants of the march (emissivity-model constants that do not depend on the point)
  PlConst pl;
  double lnu;    // log(nu_obs)
  bool allcoef;  // every absorption / rotation coefficient is switched on (ep%coefindx all 1: the usual case)
};
template <int ET>
GRT_HD RunConst run_const(const TransferParams &P, double nu_obs) {
  RunConst R;
  R.lnu = 0.0;
  R.allcoef = true;
  for (int i = 0; i < 7; i++) R.allcoef = R.allcoef && (P.coefindx[i] == 1);
  if (ET == EM_POLSYNCHPL || ET == EM_SYNCHPL || em_is_hybrid(ET)) {
    R.pl = pl_const(P.p, P.gmin, P.gmax);
    R.lnu = log(nu_obs);
  }
  return R;
}

template <int ET>
GRT_HD void node_coef(const TransferParams &P, const RunConst &R, const double *rec, const FreqConst &F, Coef &c, double &jI_raw,
                      const bool emis_width = false) {  // emis_width: keep all columns of e%K (the extra = 1 depth integrals)
  const double g = rec[1];
  double e[11];
  bool invariant = false;  // e already holds j g^2 and K LBH / g
  if (ET == EM_MAXJUTT) {  // calc_maxjutt.f90 / calc_maxcomp.f90: sum of thermal components [4 + 10 c ...]
    for (int i = 0; i < 11; i++) e[i] = 0.0;
    jI_raw = 0.0;
    double esel[4] = {0, 0, 0, 0}, jsel = 0.0;
#pragma unroll 1
    for (int cc = 0; cc < P.ncomp; cc++) {
      const double *rc_ = rec + 4 + 10 * cc;
      ThermalPoint T;
      T.xp13 = rc_[0]; T.AI = rc_[1]; T.AV = rc_[2]; T.hkT_g = rc_[3]; T.lAx = rc_[4]; T.RQ = rc_[5]; T.RV0 = rc_[6];
      T.RV1 = rc_[7]; T.ixp13 = rc_[8]; T.xpm32 = rc_[9]; T.g = g;
      double ec[11], jr;
      thermal_eval(T, F, pc::c2 * P.LBH / (2.0 * pc::h), ec, jr, P.tab);
      for (int i = 0; i < 11; i++) e[i] += ec[i];
      jI_raw += jr;
      if (cc + 1 == P.isel) {
        for (int i = 0; i < 4; i++) esel[i] = ec[i];
        jsel = jr;
      }
    }
    if (P.isel > 0 && P.isel <= P.ncomp) {  // MAXCOMP: emission of the selected component only
      for (int i = 0; i < 4; i++) e[i] = esel[i];
      jI_raw = jsel;
    }
    invariant = true;
  } else if (ET == EM_POLSYNCHTH || em_is_hybrid(ET)) {
    ThermalPoint T;
    T.xp13 = rec[4]; T.AI = rec[5]; T.AV = rec[6]; T.hkT_g = rec[7]; T.lAx = rec[8]; T.RQ = rec[9]; T.RV0 = rec[10];
    T.RV1 = rec[11]; T.ixp13 = rec[12]; T.xpm32 = rec[13]; T.g = g;
    thermal_eval(T, F, pc::c2 * P.LBH / (2.0 * pc::h), e, jI_raw, P.tab);
    invariant = true;
    if (em_is_hybrid(ET)) {  // emis.f90:469-485: add the power-law population (brought to the same invariant form)
      PlPoint Q;
      Q.nnth = rec[14]; Q.nubperp = rec[15]; Q.L = rec[16]; Q.itanth = rec[17];
      double ep[11];
      if (P.stail) {
        PlConst C = R.pl;
        C.l2gmin = rec[18]; C.kq = rec[19]; C.kv = rec[20]; C.cA = rec[21];
        pl_eval<true>(P.tab, C, Q, F.nu / g, R.lnu, ep);
      } else {
        pl_eval<true>(P.tab, R.pl, Q, F.nu / g, R.lnu, ep);
      }
      const double g2 = g * g, kg = P.LBH / g;
      if (ET == EM_HYBRIDTHPL) {
        jI_raw += ep[0];
        for (int i = 0; i < 4; i++) e[i] += ep[i] * g2;
      } else if (ET == EM_HYBRIDPL) {
        jI_raw = ep[0];
        for (int i = 0; i < 4; i++) e[i] = ep[i] * g2;
      }
      for (int i = 4; i < 11; i++) e[i] += ep[i] * kg;
    }
  } else {
    const double nu = F.nu / g;
    if (ET == EM_POLSYNCHPL || ET == EM_SYNCHPL) {
      PlPoint Q;
      Q.nnth = rec[4]; Q.nubperp = rec[5]; Q.L = rec[6]; Q.itanth = rec[7];
      if (P.stail) {  // per-point gmin (assign_source_params, fluid.f90:1656-1667)
        PlConst C = R.pl;
        C.l2gmin = rec[8]; C.kq = rec[9]; C.kv = rec[10]; C.cA = rec[11];
        pl_eval<ET == EM_POLSYNCHPL>(P.tab, C, Q, nu, R.lnu, e);
      } else {
        pl_eval<ET == EM_POLSYNCHPL>(P.tab, R.pl, Q, nu, R.lnu, e);
      }
    }
    else if (ET == EM_POLSYNCHSYMTH) {
      SymPoint Q;
      Q.Xs = rec[4]; Q.A = rec[5]; Q.S = rec[6]; Q.cq = rec[7]; Q.v0 = rec[8]; Q.v1 = rec[9]; Q.T = rec[10]; Q.lAx = rec[11];
      Q.RQ = rec[12]; Q.RV0 = rec[13]; Q.RV1 = rec[14];
      sym_eval(Q, nu, P.tab, e);
    }
    else if (ET == EM_BINS) {  // emis.f90:520-526
      bins_eval(nu, rec[5], rec[7], rec + 12, P.nbins, P.bins, e);
      if (P.etype == EM_HYBRIDTHBINS) {
        double et[11];
        synchemis_eval(nu, rec[4], rec[5], rec[6], et);
        for (int i = 0; i < 11; i++) e[i] = et[i] + e[i];
      }
    }
    else if (ET == EM_SYNCHTHAV) {  // the (n, B, T) scalar models: emis.f90:525-540
      if (P.etype == EM_BREMS) {
        brems_eval(nu, rec[4], rec[6], e);
      } else if (P.etype == EM_RHO) {  // rhoemis (emis.f90:144-151): j = n g, no absorption (a column-density map)
        for (int i = 0; i < 11; i++) e[i] = 0.0;
        e[0] = rec[4] * g;
      } else {
        synchemis_eval(nu, rec[4], rec[5], rec[6], e);
        if (P.etype == EM_SYNCHTHAVNOABS) e[4] = 0.0;
        if (P.etype == EM_SYNCHTHBREMS) {
          double eb[11];
          brems_eval(nu, rec[4], rec[6], eb);
          e[0] += eb[0];
          e[4] += eb[4];
        }
      }
    }
    else {
      for (int i = 0; i < 11; i++) e[i] = 0.0;
      if (ET == EM_BB) e[0] = planck_bnu(rec[6], nu);
      if (ET == EM_BBPOL) {
        const double fcol = 1.8;
        e[0] = pow(fcol, -4.0) * planck_bnu(rec[6] * fcol, nu);
        e[0] = e[0] * rec[7];
        e[1] = e[0] * rec[8];
      }
    }
    jI_raw = e[0];
  }
  for (int i = 0; i < 4; i++) c.j[i] = 0.0;
  for (int i = 0; i < 7; i++) c.K[i] = 0.0;
  constexpr bool POL = (ET == EM_POLSYNCHTH || ET == EM_POLSYNCHPL || ET == EM_BBPOL || ET == EM_HYBRIDTHPL ||
                        ET == EM_HYBRIDTH || ET == EM_HYBRIDPL || ET == EM_POLSYNCHSYMTH || ET == EM_MAXJUTT);
  if (POL) {
    for (int i = 0; i < 4; i++) c.j[i] = e[i];
    if (R.allcoef) {  // (uniform)
      for (int i = 0; i < 7; i++) c.K[i] = e[4 + i];
    } else {
      for (int i = 0; i < 7; i++)
        if (P.coefindx[i] == 1) c.K[i] = e[4 + i];
    }
    const double s2 = rec[2], c2 = rec[3];
    double jq = c.j[1], ju = c.j[2], aq = c.K[1], au = c.K[2], rq = c.K[4], ru = c.K[5];
    // the thermal fits have no U components in the field frame (e[2] = e[6] = e[9] = 0): their products with s2 / c2 are
    // left out (an IEEE multiply by zero cannot be folded by the compiler)
    constexpr bool UZERO = (ET == EM_POLSYNCHTH || ET == EM_MAXJUTT);
    if (UZERO) {
      c.j[1] = c2 * jq; c.j[2] = s2 * jq;
      c.K[1] = c2 * aq; c.K[2] = s2 * aq;
      c.K[4] = c2 * rq; c.K[5] = s2 * rq;
    } else {
      c.j[1] = c2 * jq - s2 * ju;
      c.j[2] = s2 * jq + c2 * ju;
      c.K[1] = c2 * aq - s2 * au;
      c.K[2] = s2 * aq + c2 * au;
      c.K[4] = c2 * rq - s2 * ru;
      c.K[5] = s2 * rq + c2 * ru;
    }
  } else {
    c.j[0] = e[0];
    c.K[0] = e[4];
  }
  if (!P.npts_is_one) {
    if (!invariant) {
      for (int i = 0; i < 4; i++) c.j[i] = c.j[i] * g * g;
      for (int i = 0; i < 7; i++) c.K[i] = c.K[i] / g * P.LBH;
    }
  } else {  // single-point rays (thin disk): I = j g^3, no LBH (grtrans_driver.f90:293-297)
    double s = invariant ? g : g * g * g;
    for (int i = 0; i < 4; i++) c.j[i] = c.j[i] * s;
  }
  if (P.rneq == 1 && !emis_width) {  // Stokes-I-only transfer of a polarised model (A21): keep j_I, alpha_I
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

// y = (c1*Id + c2*Kmat)^-1 x.  The reference inverts this 4x4 by cofactors (imatrix_4, radtrans_integrate.f90:685-733,
// ~290 flops).  Kmat = [[aI, a^T], [a, aI - [r]x]] with a = (aQ,aU,aV) and r = (rQ,rU,rV), so with d = c1 + c2 aI,
// a -> c2 a, r -> c2 r and B w = d w + w x r :
//   B^-1 w = (d^2 w + d (r x w) + r (r.w)) / (d (d^2 + r^2)),   x0 = (y0 - a.B^-1 yv) / (d - a.B^-1 a),   xv = B^-1 (yv - a x0)
// which is the same exact inverse (det = d^2 (d^2 + r^2 - a^2) - (a.r)^2) in ~70 flops and two reciprocals.
GRT_HD void solve_delo(double c1, double c2, const double *K, const double *x, double *y) {
  const double d = c1 + c2 * K[0];
  const double a1 = c2 * K[1], a2 = c2 * K[2], a3 = c2 * K[3];
  const double r1 = c2 * K[4], r2 = c2 * K[5], r3 = c2 * K[6];
  const double d2 = d * d;
  const double rr = r1 * r1 + r2 * r2 + r3 * r3, aa = a1 * a1 + a2 * a2 + a3 * a3, ra = r1 * a1 + r2 * a2 + r3 * a3;
  const double ry = r1 * x[1] + r2 * x[2] + r3 * x[3];
  // p = d^2 yv + d (r x yv) + r (r.yv),  s = d^2 a + d (r x a) + r (r.a)
  const double p1 = d2 * x[1] + d * (r2 * x[3] - r3 * x[2]) + r1 * ry;
  const double p2 = d2 * x[2] + d * (r3 * x[1] - r1 * x[3]) + r2 * ry;
  const double p3 = d2 * x[3] + d * (r1 * x[2] - r2 * x[1]) + r3 * ry;
  const double s1 = d2 * a1 + d * (r2 * a3 - r3 * a2) + r1 * ra;
  const double s2 = d2 * a2 + d * (r3 * a1 - r1 * a3) + r2 * ra;
  const double s3 = d2 * a3 + d * (r1 * a2 - r2 * a1) + r3 * ra;
  const double dn = d * (d2 + rr);
  const double det = d2 * (d2 + rr - aa) - ra * ra;
  // one reciprocal for both 1/det and 1/dn (det ~ d^4, dn ~ d^3 with d = O(1): the product cannot over/underflow)
  const double inv = 1.0 / (det * dn);
  const double x0 = (x[0] * dn - (a1 * p1 + a2 * p2 + a3 * p3)) * (inv * dn);
  const double q = inv * det;
  y[0] = x0;
  y[1] = (p1 - x0 * s1) * q;
  y[2] = (p2 - x0 * s2) * q;
  y[3] = (p3 - x0 * s3) * q;
}

// one DELO step from node k+1 (c1) to node k (c0): returns I at node k
// carry (optional): 1 / alpha_I of node k+1 as the previous step left it (NaN: not available); on return 1 / alpha_I of
// node k if this step formed it.  Saves one of the two reciprocals of consecutive optically thick steps.
// The coefficients of node k+1 are read through an accessor: from registers (CoefRegs), or one by one from the shared-memory
// column where the previous step of this thread left them (CoefSmem, k_transfer's staged kernels: the 22 registers of a
// second coefficient set are what keeps the kernel from running a fifth warp per scheduler).
struct CoefRegs {
  const Coef &c;
  GRT_HD double j(int i) const { return c.j[i]; }
  GRT_HD double K(int i) const { return c.K[i]; }
};
#define GRT_CSM_STRIDE 32  // one column of 11 doubles per lane, element i of lane l at [i * 32 + l]: conflict-free
struct CoefSmem {
  const double *p;
  GRT_HD double j(int i) const { return p[i * GRT_CSM_STRIDE]; }
  GRT_HD double K(int i) const { return p[(4 + i) * GRT_CSM_STRIDE]; }
};
GRT_HD void coef_to_smem(const Coef &c, double *p) {
  for (int i = 0; i < 4; i++) p[i * GRT_CSM_STRIDE] = c.j[i];
  for (int i = 0; i < 7; i++) p[(4 + i) * GRT_CSM_STRIDE] = c.K[i];
}
template <class C1>
GRT_HD void delo_step_t(const Coef &c0, const C1 &c1, double dx, double delta, double thin, const double *iprev, double *inew,
                        double *carry = nullptr) {
  double rhs[4], t[4];
  {  // opacity_matrix(K_k+1) * iprev
    const double k0 = c1.K(0), k1 = c1.K(1), k2 = c1.K(2), k3 = c1.K(3), k4 = c1.K(4), k5 = c1.K(5), k6 = c1.K(6);
    t[0] = k0 * iprev[0] + k1 * iprev[1] + k2 * iprev[2] + k3 * iprev[3];
    t[1] = k1 * iprev[0] + k0 * iprev[1] + k6 * iprev[2] - k5 * iprev[3];
    t[2] = k2 * iprev[0] - k6 * iprev[1] + k0 * iprev[2] + k4 * iprev[3];
    t[3] = k3 * iprev[0] + k5 * iprev[1] - k4 * iprev[2] + k0 * iprev[3];
  }
  // The optically thick and thin forms (radtrans_integrate.f90 delo) differ only in five scalar weights: the lanes of a warp
  // (32 frequencies of one ray) often straddle delta = thin, so the weights are chosen in the branch and the right-hand
  // side and the 4x4 solve are issued once, after it, for all lanes together.
  double wj0, wj1, qi, qt, sc;
  if (delta > thin) {
    const double E = exp_fast(-delta), F = 1.0 - E, G = (1.0 - (1.0 + delta) * E) / delta;
    const double i0 = 1.0 / c0.K[0];
    double i1;
    if (carry) {
      i1 = *carry;
      if (i1 != i1) i1 = 1.0 / c1.K(0);
      *carry = i0;
    } else {
      i1 = 1.0 / c1.K(0);
    }
    // (F-G) S_k + G S_k+1 + (E*Id - G*Kp1) iprev,  S = j/aI,  Kp1 = K1/aI1 - Id
    // matrix = Id + (F-G)*(K0/aI0 - Id) = (1-(F-G)) Id + ((F-G)/aI0) K0
    wj0 = (F - G) * i0; wj1 = G * i1; qi = E + G; qt = wj1; sc = 1.0 - (F - G);
  } else {
    const double ki = c0.K[0];
    qi = 1.0 - 0.5 * dx * ki + 1.0 / 6.0 * dx * dx * ki * ki; qt = 0.5 * dx - 1.0 / 3.0 * dx * dx;
    wj0 = 0.5 * dx - 1.0 / 6.0 * dx * dx * ki; wj1 = 0.5 * dx - 1.0 / 3.0 * dx * dx * ki;
    sc = 1.0 - 0.5 * delta + delta * delta * (1.0 / 6.0);  // (delta^2/6 as a product: an FP64 division costs ~25 instructions)
    if (carry) *carry = nan_val();
  }
  for (int i = 0; i < 4; i++) rhs[i] = wj0 * c0.j[i] + wj1 * c1.j(i) + (qi * iprev[i] - qt * t[i]);
  solve_delo(sc, wj0, c0.K, rhs, inew);
}
GRT_HD void delo_step(const Coef &c0, const Coef &c1, double dx, double delta, double thin, const double *iprev,
                      double *inew, double *carry = nullptr) {
  delo_step_t(c0, CoefRegs{c1}, dx, delta, thin, iprev, inew, carry);
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

// Dormand-Prince 5(4) across one tabulated interval [la, lb] with coefficients linear between the node values (ca at
// w=0, cb at w=1) -- the integrator north_star puts in LSODA's place, with the reference's tolerances (rtol 1e-6, atol
// 1e-8 on fac-normalised intensities, hmax 0.1, maxsteps per output interval like mxstep).  Same tableau, error norm and
// step controller as the oracle's restatement (oracle/radtrans.cc), which is the one pinned against ODEPACK's LSODA
// (scipy.integrate.odeint, tests/test_oracle_integrators.py).  Returns false on step overflow / NaN (pixel zeroed, like
// istate < 0 in odepack.f90:91-96).
GRT_HD bool ck_interval(int neq, const Coef &ca, const Coef &cb, double la, double lb, double *y, double &h,
                        const TransferParams &P) {
  const double c2 = 1. / 5, c3 = 3. / 10, c4 = 4. / 5, c5 = 8. / 9;
  const double a21 = 1. / 5, a31 = 3. / 40, a32 = 9. / 40, a41 = 44. / 45, a42 = -56. / 15, a43 = 32. / 9,
               a51 = 19372. / 6561, a52 = -25360. / 2187, a53 = 64448. / 6561, a54 = -212. / 729, a61 = 9017. / 3168,
               a62 = -355. / 33, a63 = 46732. / 5247, a64 = 49. / 176, a65 = -5103. / 18656, a71 = 35. / 384,
               a73 = 500. / 1113, a74 = 125. / 192, a75 = -2187. / 6784, a76 = 11. / 84;
  const double e1 = 71. / 57600, e3 = -71. / 16695, e4 = 71. / 1920, e5 = -17253. / 339200, e6 = 22. / 525, e7 = -1. / 40;
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
    double j[4], K[7], k1[4], k2[4], k3[4], k4[4], k5[4], k6[4], k7[4], yt[4], yn[4];
    coef(t, j, K);
    rt_rhs(neq, j, K, y, k1);
    for (int i = 0; i < neq; i++) yt[i] = y[i] + hh * a21 * k1[i];
    coef(t + c2 * hh, j, K);
    rt_rhs(neq, j, K, yt, k2);
    for (int i = 0; i < neq; i++) yt[i] = y[i] + hh * (a31 * k1[i] + a32 * k2[i]);
    coef(t + c3 * hh, j, K);
    rt_rhs(neq, j, K, yt, k3);
    for (int i = 0; i < neq; i++) yt[i] = y[i] + hh * (a41 * k1[i] + a42 * k2[i] + a43 * k3[i]);
    coef(t + c4 * hh, j, K);
    rt_rhs(neq, j, K, yt, k4);
    for (int i = 0; i < neq; i++) yt[i] = y[i] + hh * (a51 * k1[i] + a52 * k2[i] + a53 * k3[i] + a54 * k4[i]);
    coef(t + c5 * hh, j, K);
    rt_rhs(neq, j, K, yt, k5);
    for (int i = 0; i < neq; i++)
      yt[i] = y[i] + hh * (a61 * k1[i] + a62 * k2[i] + a63 * k3[i] + a64 * k4[i] + a65 * k5[i]);
    coef(t + hh, j, K);
    rt_rhs(neq, j, K, yt, k6);
    for (int i = 0; i < neq; i++)
      yn[i] = y[i] + hh * (a71 * k1[i] + a73 * k3[i] + a74 * k4[i] + a75 * k5[i] + a76 * k6[i]);
    rt_rhs(neq, j, K, yn, k7);
    double err = 0.0;
    for (int i = 0; i < neq; i++) {
      double ei = hh * (e1 * k1[i] + e3 * k3[i] + e4 * k4[i] + e5 * k5[i] + e6 * k6[i] + e7 * k7[i]);
      double scv = P.oatol + P.ortol * dmax(fabs(y[i]), fabs(yn[i]));
      err += (ei / scv) * (ei / scv);
    }
    err = sqrt(err / neq);
    if (!(err <= 1.0)) {
      if (!(err == err)) return false;
      h = hh * dmax(0.2, 0.9 * pow(err, -0.2));
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
// rec: [npts][GRT_NREC] records of this ray.  out: nvals Stokes values in cgs.  One instance per (emissivity, mode) so that
// the hot loop of each kernel is a few thousand instructions (it has to live in the 32 KB L1.5 instruction cache).
// HIST = 1 (extra = 1): the march also records sqrt(I^2 + Q^2) of every output column of the reference's intensity array
// (hist[col * hs], col 0-based; the caller zero-fills it), the number of columns r%npts and the emission scale fac.
// HIST = 4 (debug dump): all neq Stokes components of every column, hist[(col * 4 + i) * hs], as the integrator holds them
// (DELO / formal: I before the division by fac; RK: the fac-normalised y).
struct MarchInfo {
  int rnpts;
  double fac;
};
// STAGE (device, DELO / FORMAL only): `stage` is this warp's ring of GRT_STAGE_DEPTH records in shared memory; every lane of
// the warp must march the same ray.
// STAGE = 2 (DELO): additionally csm = this lane's shared-memory column for the coefficients of node k+1 (CoefSmem).
template <int ET, int MODE, int HIST = 0, int STAGE = 0>
GRT_HD void ray_transfer_t(const TransferParams &P, const double *rec, int npts, double nu_obs, double *out,
                           double *hist = nullptr, size_t hs = 0, MarchInfo *info = nullptr, double *stage = nullptr,
                           int lane = 0, double *csm = nullptr) {
  constexpr int neq = (MODE == MODE_QUAD || MODE == MODE_RK1) ? 1 : 4;
  if (HIST) {
    info->rnpts = npts;
    info->fac = 0.0;
  }
  constexpr int NREC = record_doubles(ET);  // record stride of this emissivity model
  for (int i = 0; i < neq; i++) out[i] = 0.0;
  if (npts <= 0) return;
  const FreqConst F = freq_const(nu_obs);
  const RunConst R = run_const<ET>(P, nu_obs);
  double jraw;
  if (npts == 1) {
    Coef c;
    node_coef<ET>(P, R, rec, F, c, jraw);
    double fac = jraw / 1.0;
    if (HIST) info->fac = fac;
    if (fac / 1e16 > 0.0)
      for (int i = 0; i < neq; i++) out[i] = c.j[i];
    return;
  }
  const double lam0 = rec[0];
  double fac = 0.0;
  if (MODE == MODE_DELO || MODE == MODE_FORMAL) {
    // march from the deepest point to the camera, I = 0 behind the last point
    double I[4] = {0, 0, 0, 0}, In[4];
    Coef c1, c0;
    double lam1 = 0.0, icarry = nan_val();
    if (STAGE)  // prologue: records npts-1 ... npts-DEPTH are on their way
      for (int j = 0; j < GRT_STAGE_DEPTH; j++)
        stage_issue(stage + j * NREC, (npts - 1 - j >= 0) ? rec + (size_t)(npts - 1 - j) * NREC : nullptr, NREC, lane);
    int slot = 0;
    // one step: coefficients of node k into cn, DELO / formal step from node k+1 (co) to node k
    auto step = [&](int k, Coef &cn, const Coef &co) {
      const double *rk = rec + (size_t)k * NREC;
      if (STAGE) {
        stage_wait();
        rk = stage + slot * NREC;
      } else if (k > 0) {
        prefetch_record(rk - NREC);
      }
      const double lamk = rk[0];
      node_coef<ET>(P, R, rk, F, cn, jraw);
      fac += jraw * (lam0 - lamk);
      if (k < npts - 1) {
        const double dx = lamk - lam1;
        if (MODE == MODE_DELO) {
          if (STAGE == 2) {
            const CoefSmem cs{csm};
            double delta = dx * 0.5 * (fabs(cn.K[0]) + fabs(cs.K(0)));
            delo_step_t(cn, cs, dx, delta, P.thin, I, In, &icarry);
          } else {
            double delta = dx * 0.5 * (fabs(cn.K[0]) + fabs(co.K[0]));
            delo_step(cn, co, dx, delta, P.thin, I, In, &icarry);
          }
        } else {
          double v[4];
          for (int i = 0; i < 4; i++) v[i] = cn.j[i] * dx + I[i];
          formal_apply(cn.K, dx, v, In);
        }
        for (int i = 0; i < 4; i++) I[i] = In[i];
        if (HIST == 1) hist[(size_t)(npts - k - 1) * hs] = sqrt(In[0] * In[0] + In[1] * In[1]);  // intensity(:, npts-k) (1-based)
        if (HIST == 4)
          for (int i = 0; i < 4; i++) hist[((size_t)(npts - k - 1) * 4 + i) * hs] = In[i];
      }
      lam1 = lamk;
      if (STAGE == 2 && MODE == MODE_DELO) coef_to_smem(cn, csm);
      if (STAGE) {  // this step's slot is free: refill it with the record GRT_STAGE_DEPTH steps further on
        stage_release();
        stage_issue(stage + slot * NREC, (k - GRT_STAGE_DEPTH >= 0) ? rec + (size_t)(k - GRT_STAGE_DEPTH) * NREC : nullptr, NREC,
                    lane);
        slot = (slot + 1 == GRT_STAGE_DEPTH) ? 0 : slot + 1;
      }
    };
    // GRT_MARCH_UNROLL2: two steps per trip with the roles of the two coefficient sets exchanged, which removes the
    // 22-register copy c1 = c0 of every step (at twice the loop body)
    constexpr bool U2 = (GRT_MARCH_UNROLL2 != 0) && !HIST;
    if (U2) {
#pragma unroll 1
      for (int k = npts - 1; k >= 0; k -= 2) {
        step(k, c0, c1);
        if (k - 1 >= 0) step(k - 1, c1, c0);
      }
    } else {
#pragma unroll 1
      for (int k = npts - 1; k >= 0; k--) {
        step(k, c0, c1);
        if (!(STAGE == 2 && MODE == MODE_DELO)) c1 = c0;
      }
    }
    if (HIST) info->fac = fac;
    if (fac / 1e16 > 0.0)
      for (int i = 0; i < 4; i++) out[i] = I[i] * P.LBH;
    return;
  }
  if (MODE == MODE_QUAD) {
    // neq == 1 with delo/formal: quadrature  I = -tsum(s, j exp(-tau)), tau from the camera inwards
    Coef c;
    double tau = 0.0, acc = 0.0, yprev = 0.0, kprev = 0.0, lprev = 0.0;
#pragma unroll 1
    for (int k = 0; k < npts; k++) {
      const double *rk = rec + (size_t)k * NREC;
      node_coef<ET>(P, R, rk, F, c, jraw);
      if (k > 0) {
        fac += jraw * (lam0 - rk[0]);
        double ds = rk[0] - lprev;
        tau -= ds * (kprev + fabs(c.K[0])) / 2.0;
        double y = c.j[0] * exp(-tau);
        acc += ds * (yprev + y) / 2.0;
        yprev = y;
      } else {
        yprev = c.j[0];
      }
      kprev = fabs(c.K[0]);
      lprev = rk[0];
    }
    if (HIST) info->fac = fac;
    if (fac / 1e16 > 0.0) out[0] = -acc * P.LBH;
    return;
  }
  // ---- LSODA replacement: pass 1 (fac, tau cut, zero trims), pass 2 (adaptive RK deep -> camera)
  int lamdex = npts, i1 = 1, i2;
  {
    Coef c;
    double tau = 0.0, kprev = 0.0, lprev = 0.0;
    int firstnz = 0, lastnz_cut = 0;
    bool cut = false;
#pragma unroll 1
    for (int k = 0; k < npts; k++) {
      const double *rk = rec + (size_t)k * NREC;
      node_coef<ET>(P, R, rk, F, c, jraw);
      fac += jraw * (lam0 - rk[0]);
      if (k > 0) tau -= (rk[0] - lprev) * (kprev + fabs(c.K[0])) / 2.0;
      kprev = fabs(c.K[0]);
      lprev = rk[0];
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
  if (HIST) {
    info->fac = fac;
    info->rnpts = i2;
  }
  if (!(fac / 1e16 > 0.0) || i2 <= i1) return;
  {
    double y[4] = {0, 0, 0, 0};
    double h = P.hmax;
    // interval [s0(indx), s0(indx-1)] uses nodes (indx, indx+1) as written (radtrans_aux :453-465 index map); the nodes
    // are produced in the order up, i2, i2-1, ..., i1+1 by a single node_coef call site
    Coef cb, ca;
    for (int i = 0; i < 4; i++) ca.j[i] = 0.0;
    for (int i = 0; i < 7; i++) ca.K[i] = 0.0;
    bool ok = true;
    const int up = (i2 + 1 <= npts) ? i2 + 1 : npts;
#pragma unroll 1
    for (int s_ = i2 + 1; s_ >= i1 + 1 && ok; s_--) {
      const int src = (s_ == i2 + 1) ? up : s_;
      cb = ca;
      node_coef<ET>(P, R, rec + (size_t)(src - 1) * NREC, F, ca, jraw);
      for (int i = 0; i < 4; i++) ca.j[i] /= fac;
      if (s_ <= i2) {
        double la = rec[(size_t)(s_ - 1) * NREC], lb = rec[(size_t)(s_ - 2) * NREC];
        if (lb > la) ok = ck_interval(neq, ca, cb, la, lb, y, h, P);
        // intensity(:, i1 + m), m = i2 - s_ + 1 (1-based column)
        if (HIST == 1 && neq == 4 && ok) hist[(size_t)(i1 + i2 - s_) * hs] = sqrt(y[0] * y[0] + y[1] * y[1]);
        if (HIST == 4 && ok)
          for (int i = 0; i < neq; i++) hist[((size_t)(i1 + i2 - s_) * 4 + i) * hs] = y[i];
      }
    }
    if (ok)
      for (int i = 0; i < neq; i++) out[i] = y[i] * fac * P.LBH;
    else if (HIST)
      for (int c = 0; c < npts * (HIST == 4 ? 4 : 1); c++) hist[(size_t)c * hs] = 0.0;  // the failed integration returns zeros (odepack.f90:91-96)
  }
}

// ---- extra = 1: the 19 diagnostics of grtrans_driver.f90:230-292 (rays with more than one point) ------------------------
// xrec: [npts][8] = r, theta, phi, n, T, B (cgs), plasma beta, unused.  xt[0..18] = r%tau(1:19):
//   0-5   optical / Faraday depths int |alpha_I,Q,U,V|, |rho_Q|, |rho_V| ds at the photosphere index
//   6-13  emission-weighted (j_I e^-tau) ray averages of r, theta, phi, n, T, B, beta, sign(cos theta)
//   14-18 averages of r, theta, the two Faraday depths and sign(cos theta) weighted with the increments of sqrt(I^2+Q^2)
// (sic) the photosphere index is npts whenever min(tau) < 1, i.e. always (tau(1) = 0); (sic) I and Q, not Q and U.
template <int ET, int MODE>
GRT_HD void ray_extra_t(const TransferParams &P, const double *rec, const double *xrec, int npts, double nu_obs, double *out,
                        double *xt, double *hist, size_t hs) {
  constexpr int neq = (MODE == MODE_QUAD || MODE == MODE_RK1) ? 1 : 4;
  constexpr int NREC = record_doubles(ET);
  for (int q = 0; q < 19; q++) xt[q] = 0.0;
  for (int c = 0; c < npts; c++) hist[(size_t)c * hs] = 0.0;
  MarchInfo mi;
  ray_transfer_t<ET, MODE, 1>(P, rec, npts, nu_obs, out, hist, hs, &mi);
  if (!(mi.fac / 1e16 > 0.0)) return;  // r%tau(:) = 0
  if (npts == 1) {  // single (disk) point: the record holds the 7 geometric quantities (grtrans_driver.f90:298-312)
    for (int q = 0; q < 7; q++) xt[q] = xrec[q];
    return;
  }
  const int rnpts = mi.rnpts;
  const FreqConst F = freq_const(nu_obs);
  const RunConst R = run_const<ET>(P, nu_obs);
  const int inds[6] = {0, 1, 2, 3, 4, 6};
  const double lam0 = rec[0];
  double ta[6] = {0, 0, 0, 0, 0, 0}, kprev[6] = {0, 0, 0, 0, 0, 0}, tau = 0.0;
  double D = 0.0, X[8] = {0, 0, 0, 0, 0, 0, 0, 0}, xprev[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dprev = 0.0, lprev = 0.0, sprev = 0.0;
  double Sd = 0.0, Sr = 0.0, Sth = 0.0, S5 = 0.0, S6 = 0.0, Ssg = 0.0;
  double best = 0.0, mintau = 0.0, snap[15];
  int taudex = 1;
  for (int q = 0; q < 15; q++) snap[q] = 0.0;
  Coef c;
  double jraw;
#pragma unroll 1
  for (int p = 0; p < npts; p++) {
    const double *rk = rec + (size_t)p * NREC, *xr = xrec + (size_t)p * 8;
    node_coef<ET>(P, R, rk, F, c, jraw, true);
    const double lam = rk[0], sp_ = lam0 - lam;
    double kabs[6];
    for (int ii = 0; ii < 6; ii++) kabs[ii] = fabs(c.K[inds[ii]]);
    if (p > 0) {
      const double ds = sp_ - sprev;
      for (int ii = 0; ii < 6; ii++) ta[ii] += ds * ((kprev[ii] + kabs[ii]) / 2.0);
      tau -= (lam - lprev) * ((kprev[0] + kabs[0]) / 2.0);
    }
    const double d = c.j[0] * exp(-tau);
    double x[8];
    for (int q = 0; q < 7; q++) x[q] = xr[q];
    x[7] = (cos(xr[1]) >= 0.0) ? 1.0 : -1.0;
    if (p > 0) {
      const double dl = lam - lprev;
      D += dl * ((dprev + d) / 2.0);
      for (int q = 0; q < 8; q++) X[q] += dl * ((dprev * xprev[q] + d * x[q]) / 2.0);
    }
    const double dist = fabs(tau - 1.0);
    if (p == 0) {
      best = dist;
      mintau = tau;
    }
    if ((p > 0 && dist < best) || p == 1) {  // candidate photosphere index (index 1 is moved to 2)
      if (dist < best || taudex == 1) {
        if (dist < best) best = dist;
        taudex = p + 1;
        for (int ii = 0; ii < 6; ii++) snap[ii] = ta[ii];
        snap[6] = D;
        for (int q = 0; q < 8; q++) snap[7 + q] = X[q];
      }
    }
    if (tau < mintau) mintau = tau;
    if (neq == 4 && p + 1 <= rnpts - 1) {
      const double dlp = hist[(size_t)(rnpts - p - 1) * hs] - hist[(size_t)(rnpts - p - 2) * hs];
      Sd += dlp; Sr += dlp * x[0]; Sth += dlp * x[1]; S5 += dlp * ta[4]; S6 += dlp * ta[5]; Ssg += dlp * x[7];
    }
    for (int ii = 0; ii < 6; ii++) kprev[ii] = kabs[ii];
    for (int q = 0; q < 8; q++) xprev[q] = x[q];
    dprev = d; lprev = lam; sprev = sp_;
  }
  if (mintau < 1.0) {  // taudex = npts
    for (int ii = 0; ii < 6; ii++) snap[ii] = ta[ii];
    snap[6] = D;
    for (int q = 0; q < 8; q++) snap[7 + q] = X[q];
  }
  for (int ii = 0; ii < 6; ii++) xt[ii] = snap[ii];
  for (int q = 0; q < 8; q++) xt[6 + q] = 1.0 / snap[6] * snap[7 + q];
  xt[14] = Sr / Sd; xt[15] = Sth / Sd; xt[16] = S5 / Sd; xt[17] = S6 / Sd; xt[18] = Ssg / Sd;
}

// runtime dispatcher (host emulator / generic callers); the CUDA kernels are instantiated per (ET, MODE) directly
#define GRT_FOR_ETYPES(OP) \
  OP(EM_POLSYNCHTH) OP(EM_POLSYNCHPL) OP(EM_SYNCHPL) OP(EM_SYNCHTHAV) OP(EM_BB) OP(EM_BBPOL) OP(EM_HYBRIDTHPL) OP(EM_HYBRIDTH) OP(EM_HYBRIDPL) \
  OP(EM_POLSYNCHSYMTH) OP(EM_MAXJUTT) OP(EM_BINS)
template <int ET>
GRT_HD void ray_transfer_e(const TransferParams &P, const double *rec, int npts, double nu_obs, double *out) {
  switch (transfer_mode(P)) {
    case MODE_DELO: ray_transfer_t<ET, MODE_DELO>(P, rec, npts, nu_obs, out); break;
    case MODE_FORMAL: ray_transfer_t<ET, MODE_FORMAL>(P, rec, npts, nu_obs, out); break;
    case MODE_QUAD: ray_transfer_t<ET, MODE_QUAD>(P, rec, npts, nu_obs, out); break;
    case MODE_RK4: ray_transfer_t<ET, MODE_RK4>(P, rec, npts, nu_obs, out); break;
    default: ray_transfer_t<ET, MODE_RK1>(P, rec, npts, nu_obs, out); break;
  }
}
template <int ET>
GRT_HD void ray_extra_e(const TransferParams &P, const double *rec, const double *xrec, int npts, double nu_obs, double *out,
                        double *xt, double *hist, size_t hs) {
  switch (transfer_mode(P)) {
    case MODE_DELO: ray_extra_t<ET, MODE_DELO>(P, rec, xrec, npts, nu_obs, out, xt, hist, hs); break;
    case MODE_FORMAL: ray_extra_t<ET, MODE_FORMAL>(P, rec, xrec, npts, nu_obs, out, xt, hist, hs); break;
    case MODE_QUAD: ray_extra_t<ET, MODE_QUAD>(P, rec, xrec, npts, nu_obs, out, xt, hist, hs); break;
    case MODE_RK4: ray_extra_t<ET, MODE_RK4>(P, rec, xrec, npts, nu_obs, out, xt, hist, hs); break;
    default: ray_extra_t<ET, MODE_RK1>(P, rec, xrec, npts, nu_obs, out, xt, hist, hs); break;
  }
}
#if !defined(__CUDA_ARCH__)
// host emulator only (the CUDA kernels are instantiated per (ET, MODE) in grt_k_extra.cu)
inline void ray_extra(const TransferParams &P, const double *rec, const double *xrec, int npts, double nu_obs, double *out,
                      double *xt, double *hist, size_t hs) {
  switch (em_slot(P.etype)) {
#define OP(E) case E: ray_extra_e<E>(P, rec, xrec, npts, nu_obs, out, xt, hist, hs); break;
    GRT_FOR_ETYPES(OP)
#undef OP
    default: break;
  }
}
#endif
GRT_HDN void ray_transfer(const TransferParams &P, const double *rec, int npts, double nu_obs, double *out) {
  switch (em_slot(P.etype)) {
#define OP(E) case E: ray_transfer_e<E>(P, rec, npts, nu_obs, out); break;
    GRT_FOR_ETYPES(OP)
#undef OP
    default:
      for (int i = 0; i < P.rneq; i++) out[i] = 0.0;
  }
}

}  // namespace grt
