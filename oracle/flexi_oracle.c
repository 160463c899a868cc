/* flexi_oracle.c -- CPU restatement of FLEXI's DGSEM right-hand side, LSERK stepping and time-step reduction.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker the CUDA path is compared with; it is never part of the product
 * path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * It is a loop-for-loop restatement (own code, same arithmetic and array layout) of the reference routines named in each
 * function header; file:line refer to /root/reference/src.  Arrays are stored exactly like the Fortran module arrays
 * (column-major, first index fastest), e.g. U(var,i,j,k,elem) -> U[(((e*np+k)*np+j)*np+i)*5+var].
 *
 * Parity status: ProlongToFace and SurfInt are pinned against the reference's unit-test fixtures (tests/golden/
 * reference_unittests, see tests/test_oracle_golden.py).  The whole Navier-Stokes path (weak-form VolInt on Gauss nodes, BR1
 * lifting with its boundary fluxes, viscous fluxes, Riemann_Side + RoeEntropyFix, GetBoundaryFlux types 2 and 4, CalcTimeStep,
 * the LSERK update) is pinned end to end on the reference's own golden state of regressioncheck/checks/parabolic/cavity_3D:
 * 353 time steps from rest reproduce its DG_Solution to 2e-13 absolute (tests/test_reference_cavity3d.py); the TGV regression
 * CSV rows of tgv/split, tgv/splitGauss and tgv/oInt pin the split forms, the overintegration and N = 11
 * (tests/test_oracle_tgv_regression.py); CalcTimeStep and the metrics agree with ElemData 'dt' of the reference's NACA0012
 * state file to round-off (tests/test_hopr_mesh.py).  The other Riemann solvers and split fluxes, BR2 and BC types 3 / 9 have
 * no reference-held datum ("parity unpinned" at function level); they are validated through identities
 * (consistency, free-stream preservation, conservation, weak==split(SD) on LGL, partition invariance, convergence order).
 *
 * MPI is emulated inside one process: a "world" holds one oracle per rank; the non-blocking exchanges of mpi/mpi.f90:244-445
 * become slice copies between ranks at the point where the reference calls FinishExchangeMPIData.  With OpenMP the ranks of
 * a world advance in parallel (one rank per thread == the reference's one MPI rank per core).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/flexi_gpu.h"

#define NVAR 5
#define NPRIM 6

/* conservative / primitive variable positions, idealgas/eos.h (0-based here) */
enum { DENS = 0, MOM1 = 1, MOM2 = 2, MOM3 = 3, ENER = 4 };
enum { VEL1 = 1, VEL2 = 2, VEL3 = 3, PRES = 4, TEMP = 5 };
/* extended vector U(1:11): cons, 1/rho, v1,v2,v3, p, T   (eos.h:91-106) */
enum { EXT_DENS = 0, EXT_MOM1 = 1, EXT_MOM2 = 2, EXT_MOM3 = 3, EXT_ENER = 4, EXT_SRHO = 5, EXT_VEL1 = 6, EXT_VEL2 = 7,
       EXT_VEL3 = 8, EXT_PRES = 9, EXT_TEMP = 10 };
/* local sides, flexi.h:90-95 */
enum { ZETA_MINUS = 1, ETA_MINUS = 2, XI_PLUS = 3, ETA_PLUS = 4, XI_MINUS = 5, ZETA_PLUS = 6 };

typedef struct oracle {
  flexi_gpu_desc d; /* scalar fields valid; pointer fields re-pointed to owned copies */
  int np, nvl;      /* N+1, PP_nVarLifting */
  int lift_map[5];  /* PRIM_LIFT (0-based prim index of each lifting variable), eos.h:118,127 */
  int lv_vel1, lv_temp;
  /* derived constants, eos.f90:136-163 */
  double kappaM1, sKappaM1, cp, cv, lambda_fac, KappasPr;
  /* state */
  double *U, *Ut, *UPrim, *Ut_tmp;
  double *U_master, *U_slave, *UPrim_master, *UPrim_slave, *Flux_master, *Flux_slave;
  double *gradUx, *gradUy, *gradUz;
  double *gradUx_master, *gradUy_master, *gradUz_master, *gradUx_slave, *gradUy_slave, *gradUz_slave;
  double *FluxX, *FluxY, *FluxZ; /* BR2 */
  double *MetricsAdv, *MetricsVisc; /* InitCalctimestep */
  /* split form on Gauss nodes (PP_EntropyVars=1, src/CMakeLists.txt:64-71): entropy variables and face metrics */
  double *V, *V_master, *V_slave; /* dg_vars.f90:59 */
  double* FilterMat;              /* (0:N,0:N) Fortran order, NULL = FilterType none; set by oracle_set_filter */
  double* OverintegrationMat;     /* (0:N,0:N) Fortran order, NULL = OverintegrationType 0; set by oracle_set_overintegration */
  /* OverintegrationType 2 (conservative cut-off, overintegration.f90:118-135): NUnder, Vdm_N_NUnder(0:NUnder,0:N),
     Vdm_NUnder_N(0:N,0:NUnder), sJNUnder(0:NUnder,0:NUnder,0:NUnder,nElems); NUnder < 0 = off */
  int NUnder;
  double *Vdm_N_NUnder, *Vdm_NUnder_N, *sJNUnder;
  double *Ja_Face, *Ja_slave;     /* (3,3,0:N,0:N,nSides), mesh_vars.f90:70-78; set by oracle_set_face_metrics */
  /* sponge (sponge/sponge_vars.f90): SpongeMap(nSpongeElems) 1-based element ids, SpongeMat(0:N,0:N,0:N,nSpongeElems) pre-multiplied
     by the Jacobian, SpBaseFlow(5,0:N,0:N,0:N,nElems); set by oracle_set_sponge */
  /* FilterType 3, LAF (filter.f90:149-186,295-448): per-element state + IntegrationWeight(0:N,0:N,0:N,nElems), normMod, LAF_alpha;
     laf != 0 switches Filter_Pointer to Filter_LAF (with FilterMat = the cut-off matrix of NFilter) */
  int laf;
  double *laf_lim, *laf_eRatio, *laf_r, *laf_ekin_old, *laf_fluc_old, *laf_iw;
  double laf_normMod, laf_alpha;
  /* PP_VISC (eos.h:73-87): 0 constant, 1 Sutherland, 2 power law; MOD_EOS_Vars values after InitEos; set by oracle_set_viscosity */
  int visc;
  double Tref, Ts, ExpoSuth, cSuth;
  int nSpongeElems;
  int* SpongeMap;
  double *SpongeMat, *SpBaseFlow;
  /* owned copies of descriptor arrays */
  void* owned[64];
  int n_owned;
  /* side ranges (1-based inclusive), mesh.f90:257-269 */
  int firstBCSide, lastBCSide, firstInnerSide, lastInnerSide, firstMPISide_MINE, lastMPISide_MINE, firstMPISide_YOUR,
      lastMPISide_YOUR;
} oracle;

/* ------------------------------------------------------------------------------------------------------------------ */
/* helpers                                                                                                            */
/* ------------------------------------------------------------------------------------------------------------------ */
static void* xcalloc(size_t n, size_t s) {
  void* p = calloc(n ? n : 1, s);
  if (!p) {
    fprintf(stderr, "oracle: out of memory\n");
    abort();
  }
  return p;
}

static void* own_copy(oracle* o, const void* src, size_t bytes) {
  if (!src) return NULL;
  void* p = xcalloc(bytes ? bytes : 1, 1);
  memcpy(p, src, bytes);
  o->owned[o->n_owned++] = p;
  return p;
}

#define VOL(a, nv, v, i, j, k, e) (a)[((((size_t)(e)*np + (k)) * np + (j)) * np + (i)) * (nv) + (v)]
#define SID(a, nv, v, p, q, s) (a)[(((size_t)(s)*np + (q)) * np + (p)) * (nv) + (v)]
#define MAT(a, r, c) (a)[(size_t)(c)*np + (r)] /* Fortran (r,c) of an (0:N,0:N) matrix */
/* S2V2(c,p,q,flip,locSide) with locSide 1..6 */
#define S2V2(c, p, q, f, ls) o->d.S2V2[((((size_t)((ls)-1) * 5 + (f)) * np + (q)) * np + (p)) * 2 + (c)]
#define S2E(c, s) o->d.SideToElem[(size_t)(s)*5 + (c)] /* c: 0 ELEM_ID,1 NB_ELEM_ID,2 LOC_SIDE_ID,3 NB_LOC_SIDE_ID,4 FLIP */

/* eos.f90:203-231 ConsToPrim */
static inline void cons_to_prim(const oracle* o, double* prim, const double* cons) {
  double sRho = 1. / cons[DENS];
  prim[DENS] = cons[DENS];
  prim[VEL1] = cons[MOM1] * sRho;
  prim[VEL2] = cons[MOM2] * sRho;
  prim[VEL3] = cons[MOM3] * sRho;
  prim[PRES] = o->kappaM1 * (cons[ENER] - 0.5 * (cons[MOM1] * prim[VEL1] + cons[MOM2] * prim[VEL2] + cons[MOM3] * prim[VEL3]));
  prim[TEMP] = prim[PRES] * sRho / o->d.R;
}

/* eos.f90:308-330 PrimToCons */
static inline void prim_to_cons(const oracle* o, const double* prim, double* cons) {
  cons[DENS] = prim[DENS];
  cons[MOM1] = prim[VEL1] * prim[DENS];
  cons[MOM2] = prim[VEL2] * prim[DENS];
  cons[MOM3] = prim[VEL3] * prim[DENS];
  cons[ENER] = o->sKappaM1 * prim[PRES] + 0.5 * (cons[MOM1] * prim[VEL1] + cons[MOM2] * prim[VEL2] + cons[MOM3] * prim[VEL3]);
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* fluxes: flux.f90                                                                                                   */
/* ------------------------------------------------------------------------------------------------------------------ */
/* flux.f90:66-119 EvalFlux3D_Point */
static inline void eval_flux3d(const double* U, const double* P, double* f, double* g, double* h) {
  double Ep = U[ENER] + P[PRES];
  f[DENS] = U[MOM1];
  f[MOM1] = U[MOM1] * P[VEL1] + P[PRES];
  f[MOM2] = U[MOM1] * P[VEL2];
  f[MOM3] = U[MOM1] * P[VEL3];
  f[ENER] = Ep * P[VEL1];
  g[DENS] = U[MOM2];
  g[MOM1] = f[MOM2];
  g[MOM2] = U[MOM2] * P[VEL2] + P[PRES];
  g[MOM3] = U[MOM2] * P[VEL3];
  g[ENER] = Ep * P[VEL2];
  h[DENS] = U[MOM3];
  h[MOM1] = f[MOM3];
  h[MOM2] = g[MOM3];
  h[MOM3] = U[MOM3] * P[VEL3] + P[PRES];
  h[ENER] = Ep * P[VEL3];
}

/* VISCOSITY_PRIM / VISCOSITY_TEMPERATURE (eos.h:73-87): mu0, muSuth(T) (viscosity.f90:57-90) or mu0*T**ExpoSuth */
static inline double viscosity_T(const oracle* o, double T) {
  if (o->visc == 1) {
    const double TnoDim = T * o->Tref; /* Tref = 1/Tref ! */
    if (TnoDim >= o->Ts) return o->d.mu0 * pow(TnoDim, o->ExpoSuth) * (1 + o->Ts) / (TnoDim + o->Ts); /* Ts = Ts/Tref ! */
    return o->d.mu0 * TnoDim * o->cSuth;
  }
  if (o->visc == 2) return o->d.mu0 * pow(T, o->ExpoSuth);
  return o->d.mu0;
}

/* flux.f90:149-241 EvalDiffFlux3D_Point (mu = VISCOSITY_PRIM(UPrim); lambda = mu*cp/Pr, eos.h:46) */
static inline void eval_diff_flux3d(const oracle* o, const double* P, const double* gx, const double* gy, const double* gz,
                                    double* f, double* g, double* h) {
  const double s23 = 2. / 3., s43 = 4. / 3.;
  const int V1 = o->lv_vel1, V2 = V1 + 1, V3 = V1 + 2, T = o->lv_temp;
  double muS = viscosity_T(o, P[TEMP]);
  double lambda = muS * o->cp / o->d.Pr;
  double v1 = P[VEL1], v2 = P[VEL2], v3 = P[VEL3];
  double tau_xx = muS * (s43 * gx[V1] - s23 * gy[V2] - s23 * gz[V3]);
  double tau_yy = muS * (-s23 * gx[V1] + s43 * gy[V2] - s23 * gz[V3]);
  double tau_zz = muS * (-s23 * gx[V1] - s23 * gy[V2] + s43 * gz[V3]);
  double tau_xy = muS * (gy[V1] + gx[V2]);
  double tau_xz = muS * (gz[V1] + gx[V3]);
  double tau_yz = muS * (gz[V2] + gy[V3]);
  f[DENS] = 0.;
  f[MOM1] = -tau_xx;
  f[MOM2] = -tau_xy;
  f[MOM3] = -tau_xz;
  f[ENER] = -tau_xx * v1 - tau_xy * v2 - tau_xz * v3 - lambda * gx[T];
  g[DENS] = 0.;
  g[MOM1] = -tau_xy;
  g[MOM2] = -tau_yy;
  g[MOM3] = -tau_yz;
  g[ENER] = -tau_xy * v1 - tau_yy * v2 - tau_yz * v3 - lambda * gy[T];
  h[DENS] = 0.;
  h[MOM1] = -tau_xz;
  h[MOM2] = -tau_yz;
  h[MOM3] = -tau_zz;
  h[ENER] = -tau_xz * v1 - tau_yz * v2 - tau_zz * v3 - lambda * gz[T];
}

/* flux.f90:396-417 EvalEulerFlux1D_fast */
static inline void eval_euler_flux1d_fast(const double* U, double* F) {
  F[DENS] = U[EXT_MOM1];
  F[MOM1] = U[EXT_MOM1] * U[EXT_VEL1] + U[EXT_PRES];
  F[MOM2] = U[EXT_MOM1] * U[EXT_VEL2];
  F[MOM3] = U[EXT_MOM1] * U[EXT_VEL3];
  F[ENER] = (U[EXT_ENER] + U[EXT_PRES]) * U[EXT_VEL1];
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* split fluxes: splitflux.f90                                                                                        */
/* ------------------------------------------------------------------------------------------------------------------ */
/* splitflux.f90:994-1020 GetLogMean (Ismail & Roe): (U_L+U_R)/g with the series of ln(chi)/f around chi = 1 */
static inline double get_log_mean(double U_L, double U_R) {
  const double sqrt_epsilon = 0.1;
  double chi = U_L / U_R, f = (chi - 1.) / (chi + 1.), g;
  if (fabs(f) < sqrt_epsilon) {
    double u = f * f;
    g = 2. + 2. / 3. * u + 0.4 * u * u + 2. / 7. * u * u * u;
  } else {
    g = log(chi) / f;
  }
  return (U_L + U_R) / g;
}

/* physical-space two-point fluxes fT,gT,hT of the chosen split form (factor 2 folded in, splitflux.f90:21-22) */
static void split_volume_flux(const oracle* o, int type, const double* UR, const double* PR, const double* U, const double* P, const double* MR,
                              const double* M, double* Flux) {
  double fT[5], gT[5], hT[5];
  switch (type) {
    case FLEXI_SPLITDG_SD: { /* splitflux.f90:160-222 */
      double rhoEpRef = UR[ENER] + PR[PRES], rhoEp = U[ENER] + P[PRES];
      fT[DENS] = (UR[MOM1] + U[MOM1]);
      fT[MOM1] = (UR[MOM1] * PR[VEL1] + PR[PRES] + U[MOM1] * P[VEL1] + P[PRES]);
      fT[MOM2] = (UR[MOM1] * PR[VEL2] + U[MOM1] * P[VEL2]);
      fT[MOM3] = (UR[MOM1] * PR[VEL3] + U[MOM1] * P[VEL3]);
      fT[ENER] = (rhoEpRef * PR[VEL1] + rhoEp * P[VEL1]);
      gT[DENS] = (UR[MOM2] + U[MOM2]);
      gT[MOM1] = (UR[MOM1] * PR[VEL2] + U[MOM1] * P[VEL2]);
      gT[MOM2] = (UR[MOM2] * PR[VEL2] + PR[PRES] + U[MOM2] * P[VEL2] + P[PRES]);
      gT[MOM3] = (UR[MOM2] * PR[VEL3] + U[MOM2] * P[VEL3]);
      gT[ENER] = (rhoEpRef * PR[VEL2] + rhoEp * P[VEL2]);
      hT[DENS] = (UR[MOM3] + U[MOM3]);
      hT[MOM1] = (UR[MOM1] * PR[VEL3] + U[MOM1] * P[VEL3]);
      hT[MOM2] = (UR[MOM2] * PR[VEL3] + U[MOM2] * P[VEL3]);
      hT[MOM3] = (UR[MOM3] * PR[VEL3] + PR[PRES] + U[MOM3] * P[VEL3] + P[PRES]);
      hT[ENER] = (rhoEpRef * PR[VEL3] + rhoEp * P[VEL3]);
    } break;
    case FLEXI_SPLITDG_DU: { /* splitflux.f90:262-320 */
      double su = PR[VEL1] + P[VEL1], sv = PR[VEL2] + P[VEL2], sw = PR[VEL3] + P[VEL3], sp = PR[PRES] + P[PRES];
      double sE = UR[ENER] + U[ENER] + PR[PRES] + P[PRES];
      fT[DENS] = 0.5 * (UR[DENS] + U[DENS]) * su;
      fT[MOM1] = 0.5 * (UR[MOM1] + U[MOM1]) * su + sp;
      fT[MOM2] = 0.5 * (UR[MOM2] + U[MOM2]) * su;
      fT[MOM3] = 0.5 * (UR[MOM3] + U[MOM3]) * su;
      fT[ENER] = 0.5 * sE * su;
      gT[DENS] = 0.5 * (UR[DENS] + U[DENS]) * sv;
      gT[MOM1] = 0.5 * (UR[MOM1] + U[MOM1]) * sv;
      gT[MOM2] = 0.5 * (UR[MOM2] + U[MOM2]) * sv + sp;
      gT[MOM3] = 0.5 * (UR[MOM3] + U[MOM3]) * sv;
      gT[ENER] = 0.5 * sE * sv;
      hT[DENS] = 0.5 * (UR[DENS] + U[DENS]) * sw;
      hT[MOM1] = 0.5 * (UR[MOM1] + U[MOM1]) * sw;
      hT[MOM2] = 0.5 * (UR[MOM2] + U[MOM2]) * sw;
      hT[MOM3] = 0.5 * (UR[MOM3] + U[MOM3]) * sw + sp;
      hT[ENER] = 0.5 * sE * sw;
    } break;
    case FLEXI_SPLITDG_KG:   /* splitflux.f90:359-425 */
    case FLEXI_SPLITDG_PI: { /* splitflux.f90:603-666 */
      double sr = UR[DENS] + U[DENS], su = PR[VEL1] + P[VEL1], sv = PR[VEL2] + P[VEL2], sw = PR[VEL3] + P[VEL3];
      double sp = PR[PRES] + P[PRES];
      fT[DENS] = 0.5 * sr * su;
      fT[MOM1] = 0.25 * sr * (su * su) + sp;
      fT[MOM2] = 0.25 * sr * su * sv;
      fT[MOM3] = 0.25 * sr * su * sw;
      gT[DENS] = 0.5 * sr * sv;
      gT[MOM1] = fT[MOM2];
      gT[MOM2] = 0.25 * sr * (sv * sv) + sp;
      gT[MOM3] = 0.25 * sr * sv * sw;
      hT[DENS] = 0.5 * sr * sw;
      hT[MOM1] = fT[MOM3];
      hT[MOM2] = gT[MOM3];
      hT[MOM3] = 0.25 * sr * (sw * sw) + sp;
      if (type == FLEXI_SPLITDG_KG) {
        double ERef = UR[ENER] / UR[DENS], E = U[ENER] / U[DENS];
        fT[ENER] = 0.25 * sr * su * (ERef + E) + 0.5 * sp * su;
        gT[ENER] = 0.25 * sr * sv * (ERef + E) + 0.5 * sp * sv;
        hT[ENER] = 0.25 * sr * sw * (ERef + E) + 0.5 * sp * sw;
      } else {
        double HRef = (UR[ENER] + PR[PRES]) / UR[DENS], H = (U[ENER] + P[PRES]) / U[DENS];
        fT[ENER] = 0.25 * sr * su * (HRef + H);
        gT[ENER] = 0.25 * sr * sv * (HRef + H);
        hT[ENER] = 0.25 * sr * sw * (HRef + H);
      }
    } break;
    case FLEXI_SPLITDG_MO: { /* splitflux.f90:471-552 (Morinishi) */
      double rhoepRef = UR[ENER] - 0.5 * UR[DENS] * (PR[VEL1] * PR[VEL1] + PR[VEL2] * PR[VEL2] + PR[VEL3] * PR[VEL3]) + PR[PRES];
      double rhoep = U[ENER] - 0.5 * U[DENS] * (P[VEL1] * P[VEL1] + P[VEL2] * P[VEL2] + P[VEL3] * P[VEL3]) + P[PRES];
      double* T[3] = {fT, gT, hT};
      for (int d = 0; d < 3; ++d) { /* direction d: mass flux rho*v_d */
        double mR = UR[MOM1 + d], m = U[MOM1 + d];
        double* t = T[d];
        t[DENS] = (mR + m);
        for (int c = 0; c < 3; ++c) t[MOM1 + c] = 0.5 * (mR + m) * (PR[VEL1 + c] + P[VEL1 + c]);
        t[MOM1 + d] = t[MOM1 + d] + (PR[PRES] + P[PRES]);
        t[ENER] = (rhoepRef * PR[VEL1 + d] + rhoep * P[VEL1 + d]) +
                  0.5 * (mR * PR[VEL1] + m * P[VEL1]) * (PR[VEL1] + P[VEL1]) +
                  0.5 * (mR * PR[VEL2] + m * P[VEL2]) * (PR[VEL2] + P[VEL2]) +
                  0.5 * (mR * PR[VEL3] + m * P[VEL3]) * (PR[VEL3] + P[VEL3]) -
                  0.5 * (mR * PR[VEL1] * PR[VEL1] + m * P[VEL1] * P[VEL1]) -
                  0.5 * (mR * PR[VEL2] * PR[VEL2] + m * P[VEL2] * P[VEL2]) -
                  0.5 * (mR * PR[VEL3] * PR[VEL3] + m * P[VEL3] * P[VEL3]);
      }
    } break;
    case FLEXI_SPLITDG_CH: { /* splitflux.f90:711-795 (Chandrashekar): entropy conserving + kinetic energy preserving */
      double uMean = 0.5 * (PR[VEL1] + P[VEL1]), vMean = 0.5 * (PR[VEL2] + P[VEL2]), wMean = 0.5 * (PR[VEL3] + P[VEL3]);
      double betaRef = 0.5 * UR[DENS] / PR[PRES], beta = 0.5 * U[DENS] / P[PRES];
      double rhoLogMean = get_log_mean(UR[DENS], U[DENS]);
      double betaLogMean = get_log_mean(betaRef, beta);
      double pHatMean = 0.5 * (UR[DENS] + U[DENS]) / (betaRef + beta);
      double HMean = 0.5 * o->sKappaM1 / betaLogMean + pHatMean / rhoLogMean +
                     0.5 * (PR[VEL1] * P[VEL1] + PR[VEL2] * P[VEL2] + PR[VEL3] * P[VEL3]);
      fT[DENS] = rhoLogMean * uMean;
      fT[MOM1] = rhoLogMean * uMean * uMean + pHatMean;
      fT[MOM2] = rhoLogMean * uMean * vMean;
      fT[MOM3] = rhoLogMean * uMean * wMean;
      fT[ENER] = rhoLogMean * HMean * uMean;
      gT[DENS] = rhoLogMean * vMean;
      gT[MOM1] = rhoLogMean * vMean * uMean;
      gT[MOM2] = rhoLogMean * vMean * vMean + pHatMean;
      gT[MOM3] = rhoLogMean * vMean * wMean;
      gT[ENER] = rhoLogMean * HMean * vMean;
      hT[DENS] = rhoLogMean * wMean;
      hT[MOM1] = rhoLogMean * wMean * uMean;
      hT[MOM2] = rhoLogMean * wMean * vMean;
      hT[MOM3] = rhoLogMean * wMean * wMean + pHatMean;
      hT[ENER] = rhoLogMean * HMean * wMean;
      for (int v = 0; v < 5; ++v) /* :794 the factor 2 is applied after the metric contraction */
        Flux[v] = (0.5 * (MR[0] + M[0]) * fT[v] + 0.5 * (MR[2] + M[2]) * hT[v] + 0.5 * (MR[1] + M[1]) * gT[v]) * 2.;
      return;
    }
    case FLEXI_SPLITDG_IR: { /* splitflux.f90:857-927 (Ismail & Roe).  As in the reference: NO factor 2 in this volume flux (every
                                other formulation carries it, :21-22), and 'ir' is not a selectable SplitDG string (:96-102) -- it is
                                reachable through InitSplitDG(SplitDG_in) only.  Restated as it stands. */
      const double kappa = o->d.kappa, kappaM1 = o->kappaM1, kappaP1 = kappa + 1.;
      double z[5], zRef[5];
      z[0] = sqrt(U[DENS] / P[PRES]);
      z[1] = z[0] * P[VEL1];
      z[2] = z[0] * P[VEL2];
      z[3] = z[0] * P[VEL3];
      z[4] = sqrt(U[DENS] * P[PRES]);
      zRef[0] = sqrt(UR[DENS] / PR[PRES]);
      zRef[1] = zRef[0] * PR[VEL1];
      zRef[2] = zRef[0] * PR[VEL2];
      zRef[3] = zRef[0] * PR[VEL3];
      zRef[4] = sqrt(UR[DENS] * PR[PRES]);
      double z1LogMean = get_log_mean(z[0], zRef[0]);
      double z5LogMean = get_log_mean(z[4], zRef[4]);
      double z1Mean = 0.5 * (z[0] + zRef[0]);
      double rhoHat = z1Mean * z5LogMean;
      double pHat = 0.5 * (z[4] + zRef[4]) / z1Mean;
      double p2Hat = 1. / (2. * kappa) * (kappaP1 * z5LogMean / z1LogMean + kappaM1 * pHat);
      double uHat = 0.5 * (z[1] + zRef[1]) / z1Mean;
      double vHat = 0.5 * (z[2] + zRef[2]) / z1Mean;
      double wHat = 0.5 * (z[3] + zRef[3]) / z1Mean;
      double hHat = kappa * p2Hat / (rhoHat * kappaM1) + 0.5 * (uHat * uHat + vHat * vHat + wHat * wHat);
      double Meff[3] = {0.5 * (MR[0] + M[0]), 0.5 * (MR[1] + M[1]), 0.5 * (MR[2] + M[2])};
      double qeff = rhoHat * (uHat * Meff[0] + vHat * Meff[1] + wHat * Meff[2]);
      Flux[0] = qeff;
      Flux[1] = qeff * uHat + pHat * Meff[0];
      Flux[2] = qeff * vHat + pHat * Meff[1];
      Flux[3] = qeff * wHat + pHat * Meff[2];
      Flux[4] = qeff * hHat;
      return;
    }
    default:
      fprintf(stderr, "oracle: SplitDG formulation %d not restated\n", type);
      abort();
  }
  for (int v = 0; v < 5; ++v)
    Flux[v] = 0.5 * (MR[0] + M[0]) * fT[v] + 0.5 * (MR[2] + M[2]) * hT[v] + 0.5 * (MR[1] + M[1]) * gT[v];
}

static void split_surface_flux(const oracle* o, int type, const double* L, const double* R, double* F) {
  switch (type) {
    case FLEXI_SPLITDG_SD: /* splitflux.f90:228-252 */
      F[DENS] = 0.5 * (L[EXT_MOM1] + R[EXT_MOM1]);
      F[MOM1] = 0.5 * (L[EXT_MOM1] * L[EXT_VEL1] + L[EXT_PRES] + R[EXT_MOM1] * R[EXT_VEL1] + R[EXT_PRES]);
      F[MOM2] = 0.5 * (L[EXT_MOM1] * L[EXT_VEL2] + R[EXT_MOM1] * R[EXT_VEL2]);
      F[MOM3] = 0.5 * (L[EXT_MOM1] * L[EXT_VEL3] + R[EXT_MOM1] * R[EXT_VEL3]);
      F[ENER] = 0.5 * ((L[EXT_ENER] + L[EXT_PRES]) * L[EXT_VEL1] + (R[EXT_ENER] + R[EXT_PRES]) * R[EXT_VEL1]);
      break;
    case FLEXI_SPLITDG_DU: /* splitflux.f90:326-349 */
      F[DENS] = 0.25 * (L[EXT_DENS] + R[EXT_DENS]) * (L[EXT_VEL1] + R[EXT_VEL1]);
      F[MOM1] = 0.25 * (L[EXT_MOM1] + R[EXT_MOM1]) * (L[EXT_VEL1] + R[EXT_VEL1]) + 0.5 * (L[EXT_PRES] + R[EXT_PRES]);
      F[MOM2] = 0.25 * (L[EXT_MOM2] + R[EXT_MOM2]) * (L[EXT_VEL1] + R[EXT_VEL1]);
      F[MOM3] = 0.25 * (L[EXT_MOM3] + R[EXT_MOM3]) * (L[EXT_VEL1] + R[EXT_VEL1]);
      F[ENER] = 0.25 * (L[EXT_ENER] + R[EXT_ENER] + L[EXT_PRES] + R[EXT_PRES]) * (L[EXT_VEL1] + R[EXT_VEL1]);
      break;
    case FLEXI_SPLITDG_KG:   /* splitflux.f90:431-460 */
    case FLEXI_SPLITDG_PI: { /* splitflux.f90:672-700 */
      double sr = L[EXT_DENS] + R[EXT_DENS], su = L[EXT_VEL1] + R[EXT_VEL1];
      F[DENS] = 0.25 * sr * su;
      F[MOM1] = 0.125 * sr * (su * su) + 0.5 * (L[EXT_PRES] + R[EXT_PRES]);
      F[MOM2] = 0.125 * sr * su * (L[EXT_VEL2] + R[EXT_VEL2]);
      F[MOM3] = 0.125 * sr * su * (L[EXT_VEL3] + R[EXT_VEL3]);
      if (type == FLEXI_SPLITDG_KG) {
        double E_LL = L[EXT_ENER] / L[EXT_DENS], E_RR = R[EXT_ENER] / R[EXT_DENS];
        F[ENER] = 0.125 * sr * (E_LL + E_RR) * su + 0.25 * (L[EXT_PRES] + R[EXT_PRES]) * su;
      } else {
        double H_LL = (L[EXT_ENER] + L[EXT_PRES]) / L[EXT_DENS], H_RR = (R[EXT_ENER] + R[EXT_PRES]) / R[EXT_DENS];
        F[ENER] = 0.125 * sr * (H_LL + H_RR) * su;
      }
    } break;
    case FLEXI_SPLITDG_MO: { /* splitflux.f90:558-593 */
      double rhoep_LL = L[EXT_ENER] - 0.5 * L[EXT_DENS] * (L[EXT_VEL1] * L[EXT_VEL1] + L[EXT_VEL2] * L[EXT_VEL2] + L[EXT_VEL3] * L[EXT_VEL3]) + L[EXT_PRES];
      double rhoep_RR = R[EXT_ENER] - 0.5 * R[EXT_DENS] * (R[EXT_VEL1] * R[EXT_VEL1] + R[EXT_VEL2] * R[EXT_VEL2] + R[EXT_VEL3] * R[EXT_VEL3]) + R[EXT_PRES];
      F[DENS] = 0.5 * (L[EXT_MOM1] + R[EXT_MOM1]);
      F[MOM1] = 0.25 * (L[EXT_MOM1] + R[EXT_MOM1]) * (L[EXT_VEL1] + R[EXT_VEL1]) + 0.5 * (L[EXT_PRES] + R[EXT_PRES]);
      F[MOM2] = 0.25 * (L[EXT_MOM1] + R[EXT_MOM1]) * (L[EXT_VEL2] + R[EXT_VEL2]);
      F[MOM3] = 0.25 * (L[EXT_MOM1] + R[EXT_MOM1]) * (L[EXT_VEL3] + R[EXT_VEL3]);
      F[ENER] = 0.5 * (rhoep_LL * L[EXT_VEL1] + rhoep_RR * R[EXT_VEL1]) +
                0.25 * (L[EXT_MOM1] * L[EXT_VEL1] + R[EXT_MOM1] * R[EXT_VEL1]) * (L[EXT_VEL1] + R[EXT_VEL1]) +
                0.25 * (L[EXT_MOM1] * L[EXT_VEL2] + R[EXT_MOM1] * R[EXT_VEL2]) * (L[EXT_VEL2] + R[EXT_VEL2]) +
                0.25 * (L[EXT_MOM1] * L[EXT_VEL3] + R[EXT_MOM1] * R[EXT_VEL3]) * (L[EXT_VEL3] + R[EXT_VEL3]) -
                0.25 * (L[EXT_MOM1] * L[EXT_VEL1] * L[EXT_VEL1] + R[EXT_MOM1] * R[EXT_VEL1] * R[EXT_VEL1]) -
                0.25 * (L[EXT_MOM1] * L[EXT_VEL2] * L[EXT_VEL2] + R[EXT_MOM1] * R[EXT_VEL2] * R[EXT_VEL2]) -
                0.25 * (L[EXT_MOM1] * L[EXT_VEL3] * L[EXT_VEL3] + R[EXT_MOM1] * R[EXT_VEL3] * R[EXT_VEL3]);
    } break;
    case FLEXI_SPLITDG_CH: { /* splitflux.f90:803-852 */
      double uMean = 0.5 * (L[EXT_VEL1] + R[EXT_VEL1]), vMean = 0.5 * (L[EXT_VEL2] + R[EXT_VEL2]), wMean = 0.5 * (L[EXT_VEL3] + R[EXT_VEL3]);
      double beta_LL = 0.5 * L[EXT_DENS] / L[EXT_PRES], beta_RR = 0.5 * R[EXT_DENS] / R[EXT_PRES];
      double rhoLogMean = get_log_mean(L[EXT_DENS], R[EXT_DENS]);
      double betaLogMean = get_log_mean(beta_LL, beta_RR);
      double pHatMean = 0.5 * (L[EXT_DENS] + R[EXT_DENS]) / (beta_LL + beta_RR);
      double HMean = 0.5 * o->sKappaM1 / betaLogMean + pHatMean / rhoLogMean +
                     0.5 * (L[EXT_VEL1] * R[EXT_VEL1] + L[EXT_VEL2] * R[EXT_VEL2] + L[EXT_VEL3] * R[EXT_VEL3]);
      F[DENS] = rhoLogMean * uMean;
      F[MOM1] = F[DENS] * uMean + pHatMean;
      F[MOM2] = F[DENS] * vMean;
      F[MOM3] = F[DENS] * wMean;
      F[ENER] = F[DENS] * HMean;
    } break;
    case FLEXI_SPLITDG_IR: { /* splitflux.f90:932-989 */
      const double kappa = o->d.kappa, kappaM1 = o->kappaM1, kappaP1 = kappa + 1.;
      double z_LL[5], z_RR[5];
      z_LL[0] = sqrt(L[EXT_DENS] / L[EXT_PRES]);
      z_LL[1] = z_LL[0] * L[EXT_VEL1];
      z_LL[2] = z_LL[0] * L[EXT_VEL2];
      z_LL[3] = z_LL[0] * L[EXT_VEL3];
      z_LL[4] = sqrt(L[EXT_DENS] * L[EXT_PRES]);
      z_RR[0] = sqrt(R[EXT_DENS] / R[EXT_PRES]);
      z_RR[1] = z_RR[0] * R[EXT_VEL1];
      z_RR[2] = z_RR[0] * R[EXT_VEL2];
      z_RR[3] = z_RR[0] * R[EXT_VEL3];
      z_RR[4] = sqrt(R[EXT_DENS] * R[EXT_PRES]);
      double z1LogMean = get_log_mean(z_LL[0], z_RR[0]);
      double z5LogMean = get_log_mean(z_LL[4], z_RR[4]);
      double z1Mean = 0.5 * (z_LL[0] + z_RR[0]);
      double rhoHat = z1Mean * z5LogMean;
      double pHat = 0.5 * (z_LL[4] + z_RR[4]) / z1Mean;
      double p2Hat = 1. / (2. * kappa) * (kappaP1 * z5LogMean / z1LogMean + kappaM1 * pHat);
      double uHat = 0.5 * (z_LL[1] + z_RR[1]) / z1Mean;
      double vHat = 0.5 * (z_LL[2] + z_RR[2]) / z1Mean;
      double wHat = 0.5 * (z_LL[3] + z_RR[3]) / z1Mean;
      double hHat = kappa * p2Hat / (rhoHat * kappaM1) + 0.5 * (uHat * uHat + vHat * vHat + wHat * wHat);
      F[0] = rhoHat * uHat;
      F[1] = F[0] * uHat + pHat;
      F[2] = F[0] * vHat;
      F[3] = F[0] * wHat;
      F[4] = F[0] * hHat;
    } break;
    default:
      fprintf(stderr, "oracle: SplitDG formulation %d not restated\n", type);
      abort();
  }
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* Riemann solvers: riemann.f90                                                                                       */
/* ------------------------------------------------------------------------------------------------------------------ */
typedef struct {
  double RoeVel[3], RoeH, Roec, absVel, H_L, H_R;
} roe_mean;

static inline roe_mean roe_average(const oracle* o, const double* L, const double* R) {
  roe_mean m;
  m.H_L = (L[EXT_ENER] + L[EXT_PRES]) * L[EXT_SRHO]; /* TOTALENTHALPY_HE */
  m.H_R = (R[EXT_ENER] + R[EXT_PRES]) * R[EXT_SRHO];
  double SqrtRho_L = sqrt(L[EXT_DENS]), SqrtRho_R = sqrt(R[EXT_DENS]);
  double sSqrtRho = 1. / (SqrtRho_L + SqrtRho_R);
  for (int c = 0; c < 3; ++c) m.RoeVel[c] = (SqrtRho_R * R[EXT_VEL1 + c] + SqrtRho_L * L[EXT_VEL1 + c]) * sSqrtRho;
  m.RoeH = (SqrtRho_R * m.H_R + SqrtRho_L * m.H_L) * sSqrtRho;
  m.absVel = m.RoeVel[0] * m.RoeVel[0] + m.RoeVel[1] * m.RoeVel[1] + m.RoeVel[2] * m.RoeVel[2];
  m.Roec = sqrt(o->kappaM1 * (m.RoeH - 0.5 * m.absVel)); /* ROEC_RIEMANN_H, eos.h:50 */
  return m;
}

static inline double speed_of_sound_he(const oracle* o, const double* UE) { return sqrt(o->d.kappa * UE[EXT_PRES] * UE[EXT_SRHO]); }

static void roe_eigenvectors(const roe_mean* m, double* a, double* r1, double* r2, double* r3, double* r4, double* r5) {
  const double* v = m->RoeVel;
  a[0] = v[0] - m->Roec; a[1] = v[0]; a[2] = v[0]; a[3] = v[0]; a[4] = v[0] + m->Roec;
  r1[0] = 1.; r1[1] = a[0]; r1[2] = v[1]; r1[3] = v[2]; r1[4] = m->RoeH - v[0] * m->Roec;
  r2[0] = 1.; r2[1] = v[0]; r2[2] = v[1]; r2[3] = v[2]; r2[4] = 0.5 * m->absVel;
  r3[0] = 0.; r3[1] = 0.; r3[2] = 1.; r3[3] = 0.; r3[4] = v[1];
  r4[0] = 0.; r4[1] = 0.; r4[2] = 0.; r4[3] = 1.; r4[4] = v[2];
  r5[0] = 1.; r5[1] = a[4]; r5[2] = v[1]; r5[3] = v[2]; r5[4] = m->RoeH + v[0] * m->Roec;
}

/* The solver proper. split != 0 selects the SPLIT_DG variants (central part = two-point surface flux). */
static void riemann_solver(const oracle* o, int solver, const double* F_L, const double* F_R, const double* L, const double* R,
                           double* F) {
  const int split = o->d.splitDG;
  const int stype = o->d.splitDGType;
  switch (solver) {
    case FLEXI_RIEMANN_LF: { /* riemann.f90:558-589 */
      double LambdaMax = fmax(fabs(R[EXT_VEL1]), fabs(L[EXT_VEL1])) + fmax(speed_of_sound_he(o, L), speed_of_sound_he(o, R));
      if (!split) {
        for (int v = 0; v < 5; ++v) F[v] = 0.5 * ((F_L[v] + F_R[v]) - LambdaMax * (R[v] - L[v]));
      } else {
        split_surface_flux(o, stype, L, R, F);
        for (int v = 0; v < 5; ++v) F[v] = F[v] - 0.5 * LambdaMax * (R[v] - L[v]);
      }
    } break;
    case FLEXI_RIEMANN_HLLC: { /* riemann.f90:594-663 (Version C wave speeds) */
      roe_mean m = roe_average(o, L, R);
      double Ssl = m.RoeVel[0] - m.Roec, Ssr = m.RoeVel[0] + m.Roec;
      if (Ssl >= 0.) {
        for (int v = 0; v < 5; ++v) F[v] = F_L[v];
      } else if (Ssr <= 0.) {
        for (int v = 0; v < 5; ++v) F[v] = F_R[v];
      } else {
        double sMu_L = Ssl - L[EXT_VEL1], sMu_R = Ssr - R[EXT_VEL1];
        double SStar = (R[EXT_PRES] - L[EXT_PRES] + L[EXT_MOM1] * sMu_L - R[EXT_MOM1] * sMu_R) /
                       (L[EXT_DENS] * sMu_L - R[EXT_DENS] * sMu_R);
        double U_Star[5];
        if ((Ssl <= 0.) && (SStar >= 0.)) {
          double EStar = L[EXT_ENER] * L[EXT_SRHO] + (SStar - L[EXT_VEL1]) * (SStar + L[EXT_PRES] * L[EXT_SRHO] / sMu_L);
          double fac = L[EXT_DENS] * sMu_L / (Ssl - SStar);
          U_Star[0] = fac * 1.; U_Star[1] = fac * SStar; U_Star[2] = fac * L[EXT_VEL2]; U_Star[3] = fac * L[EXT_VEL3];
          U_Star[4] = fac * EStar;
          for (int v = 0; v < 5; ++v) F[v] = F_L[v] + Ssl * (U_Star[v] - L[v]);
        } else {
          double EStar = R[EXT_ENER] * R[EXT_SRHO] + (SStar - R[EXT_VEL1]) * (SStar + R[EXT_PRES] * R[EXT_SRHO] / sMu_R);
          double fac = R[EXT_DENS] * sMu_R / (Ssr - SStar);
          U_Star[0] = fac * 1.; U_Star[1] = fac * SStar; U_Star[2] = fac * R[EXT_VEL2]; U_Star[3] = fac * R[EXT_VEL3];
          U_Star[4] = fac * EStar;
          for (int v = 0; v < 5; ++v) F[v] = F_R[v] + Ssr * (U_Star[v] - R[v]);
        }
      }
    } break;
    case FLEXI_RIEMANN_ROE:     /* riemann.f90:668-740 */
    case FLEXI_RIEMANN_ROEL2: { /* riemann.f90:983-1061 */
      roe_mean m = roe_average(o, L, R);
      double a[5], r1[5], r2[5], r3[5], r4[5], r5[5], dU[6];
      roe_eigenvectors(&m, a, r1, r2, r3, r4, r5);
      for (int v = 0; v < 5; ++v) dU[v] = R[v] - L[v];
      dU[5] = dU[4] - (dU[2] - m.RoeVel[1] * dU[0]) * m.RoeVel[1] - (dU[3] - m.RoeVel[2] * dU[0]) * m.RoeVel[2];
      if (solver == FLEXI_RIEMANN_ROEL2) {
        double Ma_loc = sqrt(m.absVel) / (m.Roec * sqrt(o->d.kappa));
        dU[1] *= Ma_loc; dU[2] *= Ma_loc; dU[3] *= Ma_loc;
      }
      double Alpha3 = dU[2] - m.RoeVel[1] * dU[0];
      double Alpha4 = dU[3] - m.RoeVel[2] * dU[0];
      /* ALPHA2_RIEMANN_H, eos.h:51 */
      double Alpha2 = o->kappaM1 / (m.Roec * m.Roec) * (dU[0] * (m.RoeH - m.RoeVel[0] * m.RoeVel[0]) - dU[5] + m.RoeVel[0] * dU[1]);
      double Alpha1 = 0.5 / m.Roec * (dU[0] * (m.RoeVel[0] + m.Roec) - dU[1] - m.Roec * Alpha2);
      double Alpha5 = dU[0] - Alpha1 - Alpha2;
      if (!split) {
        for (int v = 0; v < 5; ++v)
          F[v] = 0.5 * ((F_L[v] + F_R[v]) - Alpha1 * fabs(a[0]) * r1[v] - Alpha2 * fabs(a[1]) * r2[v] - Alpha3 * fabs(a[2]) * r3[v] -
                        Alpha4 * fabs(a[3]) * r4[v] - Alpha5 * fabs(a[4]) * r5[v]);
      } else {
        split_surface_flux(o, stype, L, R, F);
        for (int v = 0; v < 5; ++v)
          F[v] = F[v] - 0.5 * (Alpha1 * fabs(a[0]) * r1[v] + Alpha2 * fabs(a[1]) * r2[v] + Alpha3 * fabs(a[2]) * r3[v] +
                               Alpha4 * fabs(a[3]) * r4[v] + Alpha5 * fabs(a[4]) * r5[v]);
      }
    } break;
    case FLEXI_RIEMANN_ROEENTROPYFIX: { /* riemann.f90:748-855 */
      double c_L = speed_of_sound_he(o, L), c_R = speed_of_sound_he(o, R);
      roe_mean m = roe_average(o, L, R);
      double RoeDens = sqrt(L[EXT_DENS] * R[EXT_DENS]);
      double dU[5], a[5], r1[5], r2[5], r3[5], r4[5], r5[5], Alpha[5], al[5], ar[5];
      dU[0] = R[EXT_DENS] - L[EXT_DENS];
      dU[1] = R[EXT_VEL1] - L[EXT_VEL1];
      dU[2] = R[EXT_VEL2] - L[EXT_VEL2];
      dU[3] = R[EXT_VEL3] - L[EXT_VEL3];
      dU[4] = R[EXT_PRES] - L[EXT_PRES];
      roe_eigenvectors(&m, a, r1, r2, r3, r4, r5);
      double tmp = 0.5 / (m.Roec * m.Roec);
      Alpha[0] = tmp * (dU[4] - RoeDens * m.Roec * dU[1]);
      Alpha[1] = dU[0] - dU[4] * 2. * tmp;
      Alpha[2] = RoeDens * dU[2];
      Alpha[3] = RoeDens * dU[3];
      Alpha[4] = tmp * (dU[4] + RoeDens * m.Roec * dU[1]);
      al[0] = L[EXT_VEL1] - c_L; al[1] = al[2] = al[3] = L[EXT_VEL1]; al[4] = L[EXT_VEL1] + c_L;
      ar[0] = R[EXT_VEL1] - c_R; ar[1] = ar[2] = ar[3] = R[EXT_VEL1]; ar[4] = R[EXT_VEL1] + c_R;
      for (int iVar = 0; iVar < 5; ++iVar) { /* Harten-Hyman 2 */
        double da = fmax(0., fmax(a[iVar] - al[iVar], ar[iVar] - a[iVar]));
        if (fabs(a[iVar]) < da)
          a[iVar] = 0.5 * (a[iVar] * a[iVar] / da + da);
        else
          a[iVar] = fabs(a[iVar]);
      }
      if (!split) {
        for (int v = 0; v < 5; ++v)
          F[v] = 0.5 * ((F_L[v] + F_R[v]) - Alpha[0] * a[0] * r1[v] - Alpha[1] * a[1] * r2[v] - Alpha[2] * a[2] * r3[v] -
                        Alpha[3] * a[3] * r4[v] - Alpha[4] * a[4] * r5[v]);
      } else {
        split_surface_flux(o, stype, L, R, F);
        for (int v = 0; v < 5; ++v)
          F[v] = F[v] - 0.5 * (Alpha[0] * a[0] * r1[v] + Alpha[1] * a[1] * r2[v] + Alpha[2] * a[2] * r3[v] + Alpha[3] * a[3] * r4[v] +
                               Alpha[4] * a[4] * r5[v]);
      }
    } break;
    case FLEXI_RIEMANN_HLL:     /* riemann.f90:1067-1113 */
    case FLEXI_RIEMANN_HLLE:    /* riemann.f90:1119-1164 */
    case FLEXI_RIEMANN_HLLEM: { /* riemann.f90:1170-1228 */
      roe_mean m = roe_average(o, L, R);
      double Ssl, Ssr;
      if (solver == FLEXI_RIEMANN_HLL) {
        Ssl = m.RoeVel[0] - m.Roec;
        Ssr = m.RoeVel[0] + m.Roec;
      } else {
        double beta = sqrt(0.5 * o->kappaM1 / o->d.kappa); /* BETA_RIEMANN_H, eos.h:49 */
        Ssl = fmin(fmin(m.RoeVel[0] - m.Roec, L[EXT_VEL1] - beta * speed_of_sound_he(o, L)), 0.);
        Ssr = fmax(fmax(m.RoeVel[0] + m.Roec, R[EXT_VEL1] + beta * speed_of_sound_he(o, R)), 0.);
      }
      if (Ssl >= 0.) {
        for (int v = 0; v < 5; ++v) F[v] = F_L[v];
      } else if (Ssr <= 0.) {
        for (int v = 0; v < 5; ++v) F[v] = F_R[v];
      } else if (solver != FLEXI_RIEMANN_HLLEM) {
        for (int v = 0; v < 5; ++v) F[v] = (Ssr * F_L[v] - Ssl * F_R[v] + Ssl * Ssr * (R[v] - L[v])) / (Ssr - Ssl);
      } else {
        double RoeDens = sqrt(L[EXT_DENS] * R[EXT_DENS]);
        double delta = m.Roec / (m.Roec + fabs(0.5 * (Ssl + Ssr)));
        double A2 = (R[EXT_DENS] - L[EXT_DENS]) - (R[EXT_PRES] - L[EXT_PRES]) / (m.Roec * m.Roec);
        double A3 = RoeDens * (R[EXT_VEL2] - L[EXT_VEL2]), A4 = RoeDens * (R[EXT_VEL3] - L[EXT_VEL3]);
        double r2[5] = {1., m.RoeVel[0], m.RoeVel[1], m.RoeVel[2], 0.5 * m.absVel};
        double r3[5] = {0., 0., 1., 0., m.RoeVel[1]};
        double r4[5] = {0., 0., 0., 1., m.RoeVel[2]};
        for (int v = 0; v < 5; ++v)
          F[v] = (Ssr * F_L[v] - Ssl * F_R[v] + Ssl * Ssr * (R[v] - L[v] - delta * (r2[v] * A2 + r3[v] * A3 + r4[v] * A4))) / (Ssr - Ssl);
      }
    } break;
    case FLEXI_RIEMANN_AVERAGE: /* riemann.f90:1236-1253 Riemann_FluxAverage (SPLIT_DG only): the bare two-point surface flux */
      if (!split) goto unknown;
      split_surface_flux(o, stype, L, R, F);
      break;
    case FLEXI_RIEMANN_CH: { /* riemann.f90:1261-1310 (SPLIT_DG only): kinetic-energy consistent scalar dissipation */
      if (!split) goto unknown;
      double LambdaMax = fmax(fabs(R[EXT_VEL1]), fabs(L[EXT_VEL1])) + fmax(speed_of_sound_he(o, L), speed_of_sound_he(o, R));
      double rhoMean = 0.5 * (L[EXT_DENS] + R[EXT_DENS]);
      double uMean = 0.5 * (L[EXT_VEL1] + R[EXT_VEL1]), vMean = 0.5 * (L[EXT_VEL2] + R[EXT_VEL2]), wMean = 0.5 * (L[EXT_VEL3] + R[EXT_VEL3]);
      double beta_LL = 0.5 * L[EXT_DENS] / L[EXT_PRES], beta_RR = 0.5 * R[EXT_DENS] / R[EXT_PRES];
      double betaLogMean = get_log_mean(beta_LL, beta_RR);
      split_surface_flux(o, stype, L, R, F);
      for (int v = 0; v < 4; ++v) F[v] = F[v] - 0.5 * LambdaMax * (R[v] - L[v]);
      F[ENER] = F[ENER] - 0.5 * LambdaMax * ((R[EXT_DENS] - L[EXT_DENS]) * (0.5 * o->sKappaM1 / betaLogMean +
                                                 0.5 * (R[EXT_VEL1] * L[EXT_VEL1] + R[EXT_VEL2] * L[EXT_VEL2] + R[EXT_VEL3] * L[EXT_VEL3])) +
                                             rhoMean * uMean * (R[EXT_VEL1] - L[EXT_VEL1]) + rhoMean * vMean * (R[EXT_VEL2] - L[EXT_VEL2]) +
                                             rhoMean * wMean * (R[EXT_VEL3] - L[EXT_VEL3]) +
                                             0.5 * rhoMean * o->sKappaM1 * (1. / beta_RR - 1. / beta_LL));
    } break;
    case FLEXI_RIEMANN_IR: { /* riemann.f90:1315-1402 (SPLIT_DG only): entropy-stable matrix dissipation R |Lambda| T R^T [v] */
      if (!split) goto unknown;
      const double kappa = o->d.kappa, kappaM1 = o->kappaM1, kappaP1 = kappa + 1.;
      double sR = log(R[EXT_PRES]) - kappa * log(R[EXT_DENS]), sL = log(L[EXT_PRES]) - kappa * log(L[EXT_DENS]);
      double rho_pR = R[EXT_DENS] / R[EXT_PRES], rho_pL = L[EXT_DENS] / L[EXT_PRES];
      double vR[5], vL[5];
      vR[0] = (kappa - sR) / kappaM1 - 0.5 * rho_pR * (R[EXT_VEL1] * R[EXT_VEL1] + R[EXT_VEL2] * R[EXT_VEL2] + R[EXT_VEL3] * R[EXT_VEL3]);
      vL[0] = (kappa - sL) / kappaM1 - 0.5 * rho_pL * (L[EXT_VEL1] * L[EXT_VEL1] + L[EXT_VEL2] * L[EXT_VEL2] + L[EXT_VEL3] * L[EXT_VEL3]);
      vR[1] = rho_pR * R[EXT_VEL1]; vL[1] = rho_pL * L[EXT_VEL1];
      vR[2] = rho_pR * R[EXT_VEL2]; vL[2] = rho_pL * L[EXT_VEL2];
      vR[3] = rho_pR * R[EXT_VEL3]; vL[3] = rho_pL * L[EXT_VEL3];
      vR[4] = -rho_pR;              vL[4] = -rho_pL;
      double z_LL[5], z_RR[5];
      z_LL[0] = sqrt(L[EXT_DENS] / L[EXT_PRES]);
      z_LL[1] = z_LL[0] * L[EXT_VEL1];
      z_LL[2] = z_LL[0] * L[EXT_VEL2];
      z_LL[3] = z_LL[0] * L[EXT_VEL3];
      z_LL[4] = sqrt(L[EXT_DENS] * L[EXT_PRES]);
      z_RR[0] = sqrt(R[EXT_DENS] / R[EXT_PRES]);
      z_RR[1] = z_RR[0] * R[EXT_VEL1];
      z_RR[2] = z_RR[0] * R[EXT_VEL2];
      z_RR[3] = z_RR[0] * R[EXT_VEL3];
      z_RR[4] = sqrt(R[EXT_DENS] * R[EXT_PRES]);
      double z1LogMean = get_log_mean(z_LL[0], z_RR[0]);
      double z5LogMean = get_log_mean(z_LL[4], z_RR[4]);
      double z1Mean = 0.5 * (z_LL[0] + z_RR[0]);
      double rhoHat = z1Mean * z5LogMean;
      double uHat = 0.5 * (z_LL[1] + z_RR[1]) / z1Mean;
      double vHat = 0.5 * (z_LL[2] + z_RR[2]) / z1Mean;
      double wHat = 0.5 * (z_LL[3] + z_RR[3]) / z1Mean;
      double pHat = 0.5 * (z_LL[4] + z_RR[4]) / z1Mean;
      double p2Hat = 1. / (2. * kappa) * (kappaP1 * z5LogMean / z1LogMean + kappaM1 * pHat);
      double aHat = sqrt(kappa * p2Hat / rhoHat);
      double hHat = kappa * p2Hat / (rhoHat * kappaM1) + 0.5 * (uHat * uHat + vHat * vHat + wHat * wHat);
      double RHat[5][5] = {{1., 1., 0., 0., 1.},
                           {uHat - aHat, uHat, 0., 0., uHat + aHat},
                           {vHat, vHat, 1., 0., vHat},
                           {wHat, wHat, 0., 1., wHat},
                           {hHat - uHat * aHat, 0.5 * (uHat * uHat + vHat * vHat + wHat * wHat), vHat, wHat, hHat + uHat * aHat}};
      double DHat[5];
      DHat[0] = 0.5 * fabs(uHat - aHat) * rhoHat / kappa;
      DHat[1] = fabs(uHat) * (kappaM1 / kappa) * rhoHat;
      DHat[2] = fabs(uHat) * pHat;
      DHat[3] = DHat[2];
      DHat[4] = 0.5 * fabs(uHat + aHat) * rhoHat / kappa;
      double vJump[5], diss[5], out[5];
      for (int v = 0; v < 5; ++v) vJump[v] = vR[v] - vL[v];
      for (int c = 0; c < 5; ++c) /* R^T [v]: RHat(r,:)*vJump(r) summed over rows r */
        diss[c] = RHat[0][c] * vJump[0] + RHat[1][c] * vJump[1] + RHat[2][c] * vJump[2] + RHat[3][c] * vJump[3] + RHat[4][c] * vJump[4];
      for (int c = 0; c < 5; ++c) diss[c] = DHat[c] * diss[c];
      for (int r = 0; r < 5; ++r)
        out[r] = RHat[r][0] * diss[0] + RHat[r][1] * diss[1] + RHat[r][2] * diss[2] + RHat[r][3] * diss[3] + RHat[r][4] * diss[4];
      split_surface_flux(o, stype, L, R, F);
      for (int v = 0; v < 5; ++v) F[v] = F[v] - 0.5 * out[v];
    } break;
    default:
    unknown:
      fprintf(stderr, "oracle: Riemann solver %d not restated\n", solver);
      abort();
  }
}

/* riemann.f90:360-440 Riemann_Point (rotation, solver, back rotation) */
static void riemann_point(const oracle* o, int solver, double* FOut, const double* U_L, const double* U_R, const double* UPrim_L,
                          const double* UPrim_R, const double* nv, const double* t1, const double* t2) {
  double U_LL[11], U_RR[11], F_L[5] = {0, 0, 0, 0, 0}, F_R[5] = {0, 0, 0, 0, 0}, F[5];
  U_LL[EXT_DENS] = U_L[DENS];
  U_LL[EXT_SRHO] = 1. / U_LL[EXT_DENS];
  U_LL[EXT_ENER] = U_L[ENER];
  U_LL[EXT_PRES] = UPrim_L[PRES];
  U_LL[EXT_VEL1] = UPrim_L[VEL1] * nv[0] + UPrim_L[VEL2] * nv[1] + UPrim_L[VEL3] * nv[2];
  U_LL[EXT_VEL2] = UPrim_L[VEL1] * t1[0] + UPrim_L[VEL2] * t1[1] + UPrim_L[VEL3] * t1[2];
  U_LL[EXT_MOM1] = U_LL[EXT_DENS] * U_LL[EXT_VEL1];
  U_LL[EXT_MOM2] = U_LL[EXT_DENS] * U_LL[EXT_VEL2];
  U_LL[EXT_VEL3] = UPrim_L[VEL1] * t2[0] + UPrim_L[VEL2] * t2[1] + UPrim_L[VEL3] * t2[2];
  U_LL[EXT_MOM3] = U_LL[EXT_DENS] * U_LL[EXT_VEL3];
  U_LL[EXT_TEMP] = 0.;
  U_RR[EXT_DENS] = U_R[DENS];
  U_RR[EXT_SRHO] = 1. / U_RR[EXT_DENS];
  U_RR[EXT_ENER] = U_R[ENER];
  U_RR[EXT_PRES] = UPrim_R[PRES];
  U_RR[EXT_VEL1] = UPrim_R[VEL1] * nv[0] + UPrim_R[VEL2] * nv[1] + UPrim_R[VEL3] * nv[2];
  U_RR[EXT_VEL2] = UPrim_R[VEL1] * t1[0] + UPrim_R[VEL2] * t1[1] + UPrim_R[VEL3] * t1[2];
  U_RR[EXT_MOM1] = U_RR[EXT_DENS] * U_RR[EXT_VEL1];
  U_RR[EXT_MOM2] = U_RR[EXT_DENS] * U_RR[EXT_VEL2];
  U_RR[EXT_VEL3] = UPrim_R[VEL1] * t2[0] + UPrim_R[VEL2] * t2[1] + UPrim_R[VEL3] * t2[2];
  U_RR[EXT_MOM3] = U_RR[EXT_DENS] * U_RR[EXT_VEL3];
  U_RR[EXT_TEMP] = 0.;
  if (!o->d.splitDG) { /* riemann.f90:333-336: F_L/F_R are only formed without SPLIT_DG */
    eval_euler_flux1d_fast(U_LL, F_L);
    eval_euler_flux1d_fast(U_RR, F_R);
  }
  riemann_solver(o, solver, F_L, F_R, U_LL, U_RR, F);
  FOut[DENS] = F[DENS];
  for (int c = 0; c < 3; ++c) FOut[MOM1 + c] = nv[c] * F[MOM1] + t1[c] * F[MOM2] + t2[c] * F[MOM3];
  FOut[ENER] = F[ENER];
}

/* riemann.f90:506-551 ViscousFlux_Point */
static void viscous_flux_point(const oracle* o, double* F, const double* UPrim_L, const double* UPrim_R, const double* gx_L,
                               const double* gy_L, const double* gz_L, const double* gx_R, const double* gy_R, const double* gz_R,
                               const double* nv) {
  double fL[5], gL[5], hL[5], fR[5], gR[5], hR[5];
  eval_diff_flux3d(o, UPrim_L, gx_L, gy_L, gz_L, fL, gL, hL);
  eval_diff_flux3d(o, UPrim_R, gx_R, gy_R, gz_R, fR, gR, hR);
  for (int v = 0; v < 5; ++v) F[v] = 0.5 * (nv[0] * (fL[v] + fR[v]) + nv[1] * (gL[v] + gR[v]) + nv[2] * (hL[v] + hR[v]));
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* ProlongToFace: interpolation/prolongtoface.t90:21-303                                                              */
/* ------------------------------------------------------------------------------------------------------------------ */
/* EvalElemFaceG (:211-264) / EvalElemFaceGL (:270-303): Uface(v,a,b) of one element and local side */
static void eval_elem_face(const oracle* o, int nv, const double* Uvol_e, double* Uface, int locSide) {
  const int np = o->np, N = np - 1;
  const double* Lm = o->d.L_Minus;
  const double* Lp = o->d.L_Plus;
#define UE(v, i, j, k) Uvol_e[((((size_t)(k)) * np + (j)) * np + (i)) * nv + (v)]
#define UF(v, a, b) Uface[(((size_t)(b)) * np + (a)) * nv + (v)]
  if (o->d.nodeType == FLEXI_NODETYPE_GAUSS) {
    const double* L = (locSide == XI_MINUS || locSide == ETA_MINUS || locSide == ZETA_MINUS) ? Lm : Lp;
    for (int b = 0; b < np; ++b)
      for (int a = 0; a < np; ++a)
        for (int v = 0; v < nv; ++v) {
          double s;
          switch (locSide) {
            case XI_MINUS: case XI_PLUS:
              s = UE(v, 0, a, b) * L[0];
              for (int l = 1; l <= N; ++l) s = s + UE(v, l, a, b) * L[l];
              break;
            case ETA_MINUS: case ETA_PLUS:
              s = UE(v, a, 0, b) * L[0];
              for (int l = 1; l <= N; ++l) s = s + UE(v, a, l, b) * L[l];
              break;
            default:
              s = UE(v, a, b, 0) * L[0];
              for (int l = 1; l <= N; ++l) s = s + UE(v, a, b, l) * L[l];
          }
          UF(v, a, b) = s;
        }
  } else {
    for (int b = 0; b < np; ++b)
      for (int a = 0; a < np; ++a)
        for (int v = 0; v < nv; ++v) {
          switch (locSide) {
            case XI_MINUS: UF(v, a, b) = UE(v, 0, a, b); break;
            case XI_PLUS: UF(v, a, b) = UE(v, N, a, b); break;
            case ETA_MINUS: UF(v, a, b) = UE(v, a, 0, b); break;
            case ETA_PLUS: UF(v, a, b) = UE(v, a, N, b); break;
            case ZETA_MINUS: UF(v, a, b) = UE(v, a, b, 0); break;
            default: UF(v, a, b) = UE(v, a, b, N);
          }
        }
  }
#undef UE
}

/* ProlongToFace (:21-144) */
static void prolong_to_face(const oracle* o, int nv, const double* Uvol, double* Uface_master, double* Uface_slave, int doMPISides) {
  const int np = o->np;
  int firstSideID, lastSideID;
  if (doMPISides) {
    firstSideID = o->firstMPISide_YOUR;
    lastSideID = o->d.nSides;
  } else {
    firstSideID = 1;
    lastSideID = o->lastMPISide_MINE;
  }
  double* Uface = (double*)xcalloc((size_t)np * np * nv, sizeof(double));
  const size_t esz = (size_t)np * np * np * nv;
  for (int SideID = firstSideID; SideID <= lastSideID; ++SideID) {
    const int s = SideID - 1;
    int ElemID = S2E(0, s), nbElemID = S2E(1, s);
    if (ElemID > 0) {
      int locSide = S2E(2, s);
      eval_elem_face(o, nv, Uvol + (size_t)(ElemID - 1) * esz, Uface, locSide);
      for (int q = 0; q < np; ++q)
        for (int p = 0; p < np; ++p)
          for (int v = 0; v < nv; ++v) SID(Uface_master, nv, v, p, q, s) = UF(v, S2V2(0, p, q, 0, locSide), S2V2(1, p, q, 0, locSide));
    }
    if (nbElemID > 0) {
      int nblocSide = S2E(3, s), flip = S2E(4, s);
      eval_elem_face(o, nv, Uvol + (size_t)(nbElemID - 1) * esz, Uface, nblocSide);
      for (int q = 0; q < np; ++q)
        for (int p = 0; p < np; ++p)
          for (int v = 0; v < nv; ++v)
            SID(Uface_slave, nv, v, p, q, s) = UF(v, S2V2(0, p, q, flip, nblocSide), S2V2(1, p, q, flip, nblocSide));
    }
  }
  free(Uface);
}
#undef UF

/* ------------------------------------------------------------------------------------------------------------------ */
/* SurfInt: dg/surfint.t90 (non-split Gauss :21-239, Gauss-Lobatto :475-616, DoSurfInt :683-766)                      */
/* ------------------------------------------------------------------------------------------------------------------ */
/* DoSurfInt: Flux in the volume system of the element; Ut_e = Ut(:,:,:,:,ElemID) */
static void do_surf_int(const oracle* o, int nv, const double* Flux, const double* L_HatMinus, const double* L_HatPlus, int locSide,
                        double* Ut_e, int full_line) {
  const int np = o->np, N = np - 1;
#define UTE(v, i, j, k) Ut_e[((((size_t)(k)) * np + (j)) * np + (i)) * nv + (v)]
#define FL(v, a, b) Flux[(((size_t)(b)) * np + (a)) * nv + (v)]
  if (full_line) { /* PP_NodeType==1 (or EXACT_MM): all nodes on the line, surfint.t90:713-737 */
    const double* LH = (locSide == XI_MINUS || locSide == ETA_MINUS || locSide == ZETA_MINUS) ? L_HatMinus : L_HatPlus;
    switch (locSide) {
      case XI_MINUS: case XI_PLUS:
        for (int k = 0; k < np; ++k) for (int j = 0; j < np; ++j) for (int l = 0; l < np; ++l)
          for (int v = 0; v < nv; ++v) UTE(v, l, j, k) = UTE(v, l, j, k) + FL(v, j, k) * LH[l];
        break;
      case ETA_MINUS: case ETA_PLUS:
        for (int k = 0; k < np; ++k) for (int l = 0; l < np; ++l) for (int i = 0; i < np; ++i)
          for (int v = 0; v < nv; ++v) UTE(v, i, l, k) = UTE(v, i, l, k) + FL(v, i, k) * LH[l];
        break;
      default:
        for (int l = 0; l < np; ++l) for (int j = 0; j < np; ++j) for (int i = 0; i < np; ++i)
          for (int v = 0; v < nv; ++v) UTE(v, i, j, l) = UTE(v, i, j, l) + FL(v, i, j) * LH[l];
    }
  } else { /* PP_NodeType==2 without EXACT_MM: boundary node only, surfint.t90:738-762 */
    const double lm = L_HatMinus[0], lp = L_HatPlus[N];
    switch (locSide) {
      case XI_MINUS:
        for (int k = 0; k < np; ++k) for (int j = 0; j < np; ++j) for (int v = 0; v < nv; ++v) UTE(v, 0, j, k) = UTE(v, 0, j, k) + FL(v, j, k) * lm;
        break;
      case ETA_MINUS:
        for (int k = 0; k < np; ++k) for (int i = 0; i < np; ++i) for (int v = 0; v < nv; ++v) UTE(v, i, 0, k) = UTE(v, i, 0, k) + FL(v, i, k) * lm;
        break;
      case ZETA_MINUS:
        for (int j = 0; j < np; ++j) for (int i = 0; i < np; ++i) for (int v = 0; v < nv; ++v) UTE(v, i, j, 0) = UTE(v, i, j, 0) + FL(v, i, j) * lm;
        break;
      case XI_PLUS:
        for (int k = 0; k < np; ++k) for (int j = 0; j < np; ++j) for (int v = 0; v < nv; ++v) UTE(v, N, j, k) = UTE(v, N, j, k) + FL(v, j, k) * lp;
        break;
      case ETA_PLUS:
        for (int k = 0; k < np; ++k) for (int i = 0; i < np; ++i) for (int v = 0; v < nv; ++v) UTE(v, i, N, k) = UTE(v, i, N, k) + FL(v, i, k) * lp;
        break;
      default:
        for (int j = 0; j < np; ++j) for (int i = 0; i < np; ++i) for (int v = 0; v < nv; ++v) UTE(v, i, j, N) = UTE(v, i, j, N) + FL(v, i, j) * lp;
    }
  }
#undef UTE
#undef FL
}

/* SurfInt for nv variables. slave_sign = -1 for the DG operator (surfint.t90:583), +1 for strong-form BR1 lifting
 * (lifting_br1.t90:270-281). L_Hat* are passed in because the GL unit-test fixture feeds Gauss L_Hat. */
static void surf_int(const oracle* o, int nv, const double* Flux_master, const double* Flux_slave, double* Ut, int doMPISides,
                     const double* L_HatMinus, const double* L_HatPlus, double slave_sign) {
  const int np = o->np;
  int firstSideID, lastSideID;
  if (doMPISides) {
    firstSideID = o->firstMPISide_YOUR;
    lastSideID = o->d.nSides;
  } else {
    firstSideID = 1;
    lastSideID = o->lastMPISide_MINE;
  }
  const int full_line = (o->d.nodeType == FLEXI_NODETYPE_GAUSS);
  double* FluxTmp = (double*)xcalloc((size_t)np * np * nv, sizeof(double));
  const size_t esz = (size_t)np * np * np * nv;
  for (int SideID = firstSideID; SideID <= lastSideID; ++SideID) {
    const int s = SideID - 1;
    int ElemID = S2E(0, s), nbElemID = S2E(1, s);
    if (ElemID > 0) {
      int locSideID = S2E(2, s);
      for (int q = 0; q < np; ++q)
        for (int p = 0; p < np; ++p)
          for (int v = 0; v < nv; ++v)
            FluxTmp[(((size_t)S2V2(1, p, q, 0, locSideID)) * np + S2V2(0, p, q, 0, locSideID)) * nv + v] = SID(Flux_master, nv, v, p, q, s);
      do_surf_int(o, nv, FluxTmp, L_HatMinus, L_HatPlus, locSideID, Ut + (size_t)(ElemID - 1) * esz, full_line);
    }
    if (nbElemID > 0) {
      int nblocSideID = S2E(3, s), flip = S2E(4, s);
      for (int q = 0; q < np; ++q)
        for (int p = 0; p < np; ++p)
          for (int v = 0; v < nv; ++v)
            FluxTmp[(((size_t)S2V2(1, p, q, flip, nblocSideID)) * np + S2V2(0, p, q, flip, nblocSideID)) * nv + v] =
                slave_sign * SID(Flux_slave, nv, v, p, q, s);
      do_surf_int(o, nv, FluxTmp, L_HatMinus, L_HatPlus, nblocSideID, Ut + (size_t)(nbElemID - 1) * esz, full_line);
    }
  }
  free(FluxTmp);
}

/* SurfInt with SPLIT_DG on Gauss nodes, dg/surfint.t90:21-239 (Mateo-Gabin et al., JCP 2023, eq. 5): besides the numerical flux
 * every interior node of the line behind a face node receives the two-point fluxes between the line's nodes and the face state,
 *   Ut(line node l) += (+-Flux - 1/2 (sum_k Leg(k) FluxB(k) - FluxB(l)) * NormalSign) * LegHat(l).
 * S2V(:,l,p,q,flip,locSide) is rebuilt from S2V2: the line behind volume-face node (a,b) runs along the side's normal direction. */
static void surf_int_split_gauss(const oracle* o, const double* Flux_master, const double* Flux_slave, double* Ut, int doMPISides) {
  const int np = o->np, st = o->d.splitDGType;
  static const int NormalDirs[6] = {3, 2, 1, 2, 1, 3};
  static const double NormalSigns[6] = {-1., -1., 1., 1., -1., 1.};
  int firstSideID, lastSideID;
  if (doMPISides) {
    firstSideID = o->firstMPISide_YOUR;
    lastSideID = o->d.nSides;
  } else {
    firstSideID = 1;
    lastSideID = o->lastMPISide_MINE;
  }
  double* FluxB = (double*)xcalloc((size_t)np * NVAR, sizeof(double));
  for (int SideID = firstSideID; SideID <= lastSideID; ++SideID) {
    const int s = SideID - 1;
    for (int pass = 0; pass < 2; ++pass) { /* 0: master element, 1: slave element */
      const int elem = S2E(pass, s);
      if (elem <= 0) continue;
      const int e = elem - 1;
      const int locSide = S2E(2 + pass, s), flip = pass ? S2E(4, s) : 0;
      const int idx = NormalDirs[locSide - 1]; /* 1 xi, 2 eta, 3 zeta */
      const int minus = (locSide == 5 || locSide == 2 || locSide == 1);
      const double* Leg = minus ? o->d.L_Minus : o->d.L_Plus;
      const double* LegHat = minus ? o->d.L_HatMinus : o->d.L_HatPlus;
      const double* Met = idx == 1 ? o->d.Metrics_fTilde : idx == 2 ? o->d.Metrics_gTilde : o->d.Metrics_hTilde;
      const double* Uf = pass ? o->U_slave : o->U_master;
      const double* Pf = pass ? o->UPrim_slave : o->UPrim_master;
      const double* Ja = pass ? o->Ja_slave : o->Ja_Face;
      const double* Fl = pass ? Flux_slave : Flux_master;
      const double fsgn = pass ? -1. : 1.;
      for (int q = 0; q < np; ++q)
        for (int p = 0; p < np; ++p) {
          const int a = S2V2(0, p, q, flip, locSide), b = S2V2(1, p, q, flip, locSide);
          const double* JaRow = Ja + ((((size_t)s * np + q) * np + p) * 3) * 3; /* Ja(1:3 dd fastest? see below) */
          double Jv[3];
          for (int c = 0; c < 3; ++c) Jv[c] = JaRow[c * 3 + (idx - 1)]; /* Fortran Ja(dd,c,p,q,side): dd fastest */
          double FluxB_sum[NVAR] = {0., 0., 0., 0., 0.};
          for (int l = 0; l < np; ++l) {
            const int i = idx == 1 ? l : a, j = idx == 1 ? a : (idx == 2 ? l : b), k = idx == 3 ? l : b;
            split_volume_flux(o, st, &VOL(o->U, NVAR, 0, i, j, k, e), &VOL(o->UPrim, NPRIM, 0, i, j, k, e), &SID(Uf, NVAR, 0, p, q, s),
                              &SID(Pf, NPRIM, 0, p, q, s), &VOL(Met, 3, 0, i, j, k, e), Jv, FluxB + (size_t)l * NVAR);
            for (int v = 0; v < NVAR; ++v) FluxB_sum[v] = FluxB_sum[v] + FluxB[(size_t)l * NVAR + v] * Leg[l];
          }
          for (int l = 0; l < np; ++l) {
            const int i = idx == 1 ? l : a, j = idx == 1 ? a : (idx == 2 ? l : b), k = idx == 3 ? l : b;
            for (int v = 0; v < NVAR; ++v)
              VOL(Ut, NVAR, v, i, j, k, e) = VOL(Ut, NVAR, v, i, j, k, e) +
                                           (fsgn * SID(Fl, NVAR, v, p, q, s) - 1. / 2. * (FluxB_sum[v] - FluxB[(size_t)l * NVAR + v]) * NormalSigns[locSide - 1]) * LegHat[l];
          }
        }
    }
  }
  free(FluxB);
}

/* eos.f90:401-420 PrimToEntropy */
static void prim_to_entropy(const oracle* o, double* entropy, const double* prim) {
  const double kappa = o->d.kappa;
  const double s = log(prim[4]) - kappa * log(prim[0]);
  const double rho_p = prim[0] / prim[4];
  entropy[0] = (kappa - s) * o->sKappaM1 - rho_p * 0.5 * (prim[1] * prim[1] + prim[2] * prim[2] + prim[3] * prim[3]);
  entropy[1] = rho_p * prim[1];
  entropy[2] = rho_p * prim[2];
  entropy[3] = rho_p * prim[3];
  entropy[4] = -rho_p;
}
/* eos.f90:499-531 EntropyToCons */
static void entropy_to_cons(const oracle* o, const double* entropy, double* cons) {
  const double kappa = o->d.kappa, kappaM1 = o->kappaM1;
  double e2[NVAR];
  for (int v = 0; v < NVAR; ++v) e2[v] = entropy[v] * kappaM1;
  const double ekin = 0.5 * (e2[1] * e2[1] + e2[2] * e2[2] + e2[3] * e2[3]) / e2[4];
  const double s = kappa - e2[0] + ekin;
  const double rhoe = pow(kappaM1, o->sKappaM1) * pow(-e2[4], -(kappa / kappaM1)) * exp(-s * o->sKappaM1);
  cons[0] = -rhoe * e2[4];
  cons[1] = rhoe * e2[1];
  cons[2] = rhoe * e2[2];
  cons[3] = rhoe * e2[3];
  cons[4] = rhoe * (1. - ekin);
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* volume primitive state and face primitive state                                                                    */
/* ------------------------------------------------------------------------------------------------------------------ */
/* eos.f90:280-300 ConsToPrim_Volume */
static void cons_to_prim_volume(oracle* o) {
  const size_t n = (size_t)o->d.nElems * o->np * o->np * o->np;
  for (size_t i = 0; i < n; ++i) cons_to_prim(o, o->UPrim + i * NPRIM, o->U + i * NVAR);
}

/* equation.f90:189-216 GetPrimitiveStateSurface */
static void get_primitive_state_surface(oracle* o) {
  const int np = o->np;
  const size_t fs = (size_t)np * np;
  for (int iSide = 1; iSide <= o->d.nSides; ++iSide) {
    if (iSide >= o->firstMPISide_YOUR && iSide <= o->lastMPISide_YOUR) continue;
    for (size_t i = 0; i < fs; ++i)
      cons_to_prim(o, o->UPrim_master + ((size_t)(iSide - 1) * fs + i) * NPRIM, o->U_master + ((size_t)(iSide - 1) * fs + i) * NVAR);
  }
  for (int iSide = o->firstInnerSide; iSide <= o->lastMPISide_YOUR; ++iSide)
    for (size_t i = 0; i < fs; ++i)
      cons_to_prim(o, o->UPrim_slave + ((size_t)(iSide - 1) * fs + i) * NPRIM, o->U_slave + ((size_t)(iSide - 1) * fs + i) * NVAR);
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* Lifting: dg/lifting/ (.t90)                                                                                            */
/* ------------------------------------------------------------------------------------------------------------------ */
/* lifting_volint.t90:26-118 Lifting_VolInt_Nonconservative */
static void lifting_volint_nonconservative(oracle* o) {
  const int np = o->np, N = np - 1, nvl = o->nvl;
  const double* D_T = o->d.D_T;
  for (int e = 0; e < o->d.nElems; ++e)
    for (int k = 0; k < np; ++k)
      for (int j = 0; j < np; ++j)
        for (int i = 0; i < np; ++i) {
          double gxi[5], geta[5], gzeta[5];
          for (int v = 0; v < nvl; ++v) {
            const int pv = o->lift_map[v];
            gxi[v] = MAT(D_T, 0, i) * VOL(o->UPrim, NPRIM, pv, 0, j, k, e);
            geta[v] = MAT(D_T, 0, j) * VOL(o->UPrim, NPRIM, pv, i, 0, k, e);
            gzeta[v] = MAT(D_T, 0, k) * VOL(o->UPrim, NPRIM, pv, i, j, 0, e);
            for (int l = 1; l <= N; ++l) {
              gxi[v] = gxi[v] + MAT(D_T, l, i) * VOL(o->UPrim, NPRIM, pv, l, j, k, e);
              geta[v] = geta[v] + MAT(D_T, l, j) * VOL(o->UPrim, NPRIM, pv, i, l, k, e);
              gzeta[v] = gzeta[v] + MAT(D_T, l, k) * VOL(o->UPrim, NPRIM, pv, i, j, l, e);
            }
          }
          const double* Mf = &VOL(o->d.Metrics_fTilde, 3, 0, i, j, k, e);
          const double* Mg = &VOL(o->d.Metrics_gTilde, 3, 0, i, j, k, e);
          const double* Mh = &VOL(o->d.Metrics_hTilde, 3, 0, i, j, k, e);
          for (int v = 0; v < nvl; ++v) {
            VOL(o->gradUx, nvl, v, i, j, k, e) = Mf[0] * gxi[v] + Mg[0] * geta[v] + Mh[0] * gzeta[v];
            VOL(o->gradUy, nvl, v, i, j, k, e) = Mf[1] * gxi[v] + Mg[1] * geta[v] + Mh[1] * gzeta[v];
            VOL(o->gradUz, nvl, v, i, j, k, e) = Mf[2] * gxi[v] + Mg[2] * geta[v] + Mh[2] * gzeta[v];
          }
        }
}

/* lifting_fillflux.t90:40-107 Lifting_FillFlux (dir=0: BR1; dir=1..3: BR2), strong form (sig=-1) */
static void lifting_fill_flux(oracle* o, int dir, double* Flux, int doMPISides) {
  const int np = o->np, nvl = o->nvl;
  int firstSideID = doMPISides ? o->firstMPISide_MINE : o->firstInnerSide;
  int lastSideID = doMPISides ? o->lastMPISide_MINE : o->lastInnerSide;
  const double sig = -1.;
  for (int SideID = firstSideID; SideID <= lastSideID; ++SideID) {
    const int s = SideID - 1;
    for (int q = 0; q < np; ++q)
      for (int p = 0; p < np; ++p)
        for (int v = 0; v < nvl; ++v) {
          const int pv = o->lift_map[v];
          double fl = sig * SID(o->UPrim_master, NPRIM, pv, p, q, s) + SID(o->UPrim_slave, NPRIM, pv, p, q, s);
          double surf = o->d.SurfElem[((size_t)s * np + q) * np + p];
          if (dir == 0)
            SID(Flux, nvl, v, p, q, s) = 0.5 * fl * surf;
          else
            SID(Flux, nvl, v, p, q, s) = 0.5 * fl * SID(o->d.NormVec, 3, dir - 1, p, q, s) * surf;
        }
  }
}

/* lifting_fillflux.t90:189-239 Lifting_FillFlux_NormVec */
static void lifting_fill_flux_normvec(oracle* o, const double* Flux, double* FluxX, double* FluxY, double* FluxZ, int doMPISides) {
  const int np = o->np, nvl = o->nvl;
  int firstSideID = doMPISides ? o->firstMPISide_YOUR : 1;
  int lastSideID = doMPISides ? o->lastMPISide_YOUR : o->lastMPISide_MINE;
  for (int SideID = firstSideID; SideID <= lastSideID; ++SideID) {
    const int s = SideID - 1;
    for (int q = 0; q < np; ++q)
      for (int p = 0; p < np; ++p)
        for (int v = 0; v < nvl; ++v) {
          double f = SID(Flux, nvl, v, p, q, s);
          SID(FluxX, nvl, v, p, q, s) = f * SID(o->d.NormVec, 3, 0, p, q, s);
          SID(FluxY, nvl, v, p, q, s) = f * SID(o->d.NormVec, 3, 1, p, q, s);
          SID(FluxZ, nvl, v, p, q, s) = f * SID(o->d.NormVec, 3, 2, p, q, s);
        }
  }
}

/* applyjacobian.t90:62-112 ApplyJacobian (toPhysical) */
static void apply_jacobian(const oracle* o, int nv, double* A) {
  const size_t n = (size_t)o->d.nElems * o->np * o->np * o->np;
  for (size_t i = 0; i < n; ++i)
    for (int v = 0; v < nv; ++v) A[i * nv + v] = A[i * nv + v] * o->d.sJ[i];
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* VolInt: dg/volint.f90                                                                                              */
/* ------------------------------------------------------------------------------------------------------------------ */
/* EvalFlux3D + EvalDiffFlux3D + VolInt_Metrics (volint.f90:378-422) of one element into f,g,h(5,i,j,k) */
static void volint_fluxes(const oracle* o, int e, int advective, int viscous, double* f, double* g, double* h) {
  const int np = o->np, nvl = o->nvl;
  for (int k = 0; k < np; ++k)
    for (int j = 0; j < np; ++j)
      for (int i = 0; i < np; ++i) {
        double fa[5] = {0, 0, 0, 0, 0}, ga[5] = {0, 0, 0, 0, 0}, ha[5] = {0, 0, 0, 0, 0};
        const double* U = &VOL(o->U, NVAR, 0, i, j, k, e);
        const double* P = &VOL(o->UPrim, NPRIM, 0, i, j, k, e);
        if (advective) eval_flux3d(U, P, fa, ga, ha);
        if (viscous) {
          double fv[5], gv[5], hv[5];
          eval_diff_flux3d(o, P, &VOL(o->gradUx, nvl, 0, i, j, k, e), &VOL(o->gradUy, nvl, 0, i, j, k, e),
                           &VOL(o->gradUz, nvl, 0, i, j, k, e), fv, gv, hv);
          for (int v = 0; v < 5; ++v) {
            fa[v] = fa[v] + fv[v];
            ga[v] = ga[v] + gv[v];
            ha[v] = ha[v] + hv[v];
          }
        }
        const double* Mf = &VOL(o->d.Metrics_fTilde, 3, 0, i, j, k, e);
        const double* Mg = &VOL(o->d.Metrics_gTilde, 3, 0, i, j, k, e);
        const double* Mh = &VOL(o->d.Metrics_hTilde, 3, 0, i, j, k, e);
        const size_t n = (((size_t)k * np + j) * np + i) * 5;
        for (int v = 0; v < 5; ++v) {
          f[n + v] = fa[v] * Mf[0] + ga[v] * Mf[1] + ha[v] * Mf[2];
          g[n + v] = fa[v] * Mg[0] + ga[v] * Mg[1] + ha[v] * Mg[2];
          h[n + v] = fa[v] * Mh[0] + ga[v] * Mh[1] + ha[v] * Mh[2];
        }
      }
}

/* volint.f90:117-187 VolInt_weakForm */
static void volint_weak_form(oracle* o) {
  const int np = o->np, N = np - 1;
  const double* DH = o->d.D_Hat_T;
  const size_t esz = (size_t)np * np * np * 5;
  double* f = (double*)xcalloc(3 * esz, sizeof(double));
  double *g = f + esz, *h = g + esz;
#define F3(a, v, i, j, k) a[((((size_t)(k)) * np + (j)) * np + (i)) * 5 + (v)]
  for (int e = 0; e < o->d.nElems; ++e) {
    volint_fluxes(o, e, 1, o->d.parabolic, f, g, h);
    for (int k = 0; k < np; ++k)
      for (int j = 0; j < np; ++j)
        for (int i = 0; i < np; ++i)
          for (int v = 0; v < 5; ++v) {
            double ut = MAT(DH, 0, i) * F3(f, v, 0, j, k) + MAT(DH, 0, k) * F3(h, v, i, j, 0) + MAT(DH, 0, j) * F3(g, v, i, 0, k);
            for (int l = 1; l <= N; ++l)
              ut = ut + MAT(DH, l, i) * F3(f, v, l, j, k) + MAT(DH, l, k) * F3(h, v, i, j, l) + MAT(DH, l, j) * F3(g, v, i, l, k);
            VOL(o->Ut, NVAR, v, i, j, k, e) = ut;
          }
  }
  free(f);
}

/* volint.f90:211-371 VolInt_splitForm */
static void volint_split_form(oracle* o) {
  const int np = o->np, N = np - 1;
  const double* DH = o->d.D_Hat_T;
  const double* DV = o->d.DVolSurf;
  const int st = o->d.splitDGType;
  const size_t esz = (size_t)np * np * np * 5;
  double* fv = (double*)xcalloc(3 * esz, sizeof(double));
  double *gv = fv + esz, *hv = gv + esz;
  for (int e = 0; e < o->d.nElems; ++e) {
    if (o->d.parabolic) {
      volint_fluxes(o, e, 0, 1, fv, gv, hv);
      for (int k = 0; k < np; ++k)
        for (int j = 0; j < np; ++j)
          for (int i = 0; i < np; ++i)
            for (int v = 0; v < 5; ++v)
              VOL(o->Ut, NVAR, v, i, j, k, e) = MAT(DH, i, i) * F3(fv, v, i, j, k) + MAT(DH, j, j) * F3(gv, v, i, j, k) + MAT(DH, k, k) * F3(hv, v, i, j, k);
    } else {
      memset(&VOL(o->Ut, NVAR, 0, 0, 0, 0, e), 0, esz * sizeof(double));
    }
    for (int k = 0; k < np; ++k)
      for (int j = 0; j < np; ++j)
        for (int i = 0; i < np; ++i) {
          double Flux[5];
          for (int l = i + 1; l <= N; ++l) {
            split_volume_flux(o, st, &VOL(o->U, NVAR, 0, i, j, k, e), &VOL(o->UPrim, NPRIM, 0, i, j, k, e), &VOL(o->U, NVAR, 0, l, j, k, e),
                              &VOL(o->UPrim, NPRIM, 0, l, j, k, e), &VOL(o->d.Metrics_fTilde, 3, 0, i, j, k, e),
                              &VOL(o->d.Metrics_fTilde, 3, 0, l, j, k, e), Flux);
            for (int v = 0; v < 5; ++v) {
              VOL(o->Ut, NVAR, v, i, j, k, e) = VOL(o->Ut, NVAR, v, i, j, k, e) + MAT(DV, l, i) * Flux[v];
              VOL(o->Ut, NVAR, v, l, j, k, e) = VOL(o->Ut, NVAR, v, l, j, k, e) + MAT(DV, i, l) * Flux[v];
            }
            if (o->d.parabolic)
              for (int v = 0; v < 5; ++v) {
                VOL(o->Ut, NVAR, v, i, j, k, e) = VOL(o->Ut, NVAR, v, i, j, k, e) + MAT(DH, l, i) * F3(fv, v, l, j, k);
                VOL(o->Ut, NVAR, v, l, j, k, e) = VOL(o->Ut, NVAR, v, l, j, k, e) + MAT(DH, i, l) * F3(fv, v, i, j, k);
              }
          }
          for (int l = j + 1; l <= N; ++l) {
            split_volume_flux(o, st, &VOL(o->U, NVAR, 0, i, j, k, e), &VOL(o->UPrim, NPRIM, 0, i, j, k, e), &VOL(o->U, NVAR, 0, i, l, k, e),
                              &VOL(o->UPrim, NPRIM, 0, i, l, k, e), &VOL(o->d.Metrics_gTilde, 3, 0, i, j, k, e),
                              &VOL(o->d.Metrics_gTilde, 3, 0, i, l, k, e), Flux);
            for (int v = 0; v < 5; ++v) {
              VOL(o->Ut, NVAR, v, i, j, k, e) = VOL(o->Ut, NVAR, v, i, j, k, e) + MAT(DV, l, j) * Flux[v];
              VOL(o->Ut, NVAR, v, i, l, k, e) = VOL(o->Ut, NVAR, v, i, l, k, e) + MAT(DV, j, l) * Flux[v];
            }
            if (o->d.parabolic)
              for (int v = 0; v < 5; ++v) {
                VOL(o->Ut, NVAR, v, i, j, k, e) = VOL(o->Ut, NVAR, v, i, j, k, e) + MAT(DH, l, j) * F3(gv, v, i, l, k);
                VOL(o->Ut, NVAR, v, i, l, k, e) = VOL(o->Ut, NVAR, v, i, l, k, e) + MAT(DH, j, l) * F3(gv, v, i, j, k);
              }
          }
          for (int l = k + 1; l <= N; ++l) {
            split_volume_flux(o, st, &VOL(o->U, NVAR, 0, i, j, k, e), &VOL(o->UPrim, NPRIM, 0, i, j, k, e), &VOL(o->U, NVAR, 0, i, j, l, e),
                              &VOL(o->UPrim, NPRIM, 0, i, j, l, e), &VOL(o->d.Metrics_hTilde, 3, 0, i, j, k, e),
                              &VOL(o->d.Metrics_hTilde, 3, 0, i, j, l, e), Flux);
            for (int v = 0; v < 5; ++v) {
              VOL(o->Ut, NVAR, v, i, j, k, e) = VOL(o->Ut, NVAR, v, i, j, k, e) + MAT(DV, l, k) * Flux[v];
              VOL(o->Ut, NVAR, v, i, j, l, e) = VOL(o->Ut, NVAR, v, i, j, l, e) + MAT(DV, k, l) * Flux[v];
            }
            if (o->d.parabolic)
              for (int v = 0; v < 5; ++v) {
                VOL(o->Ut, NVAR, v, i, j, k, e) = VOL(o->Ut, NVAR, v, i, j, k, e) + MAT(DH, l, k) * F3(hv, v, i, j, l);
                VOL(o->Ut, NVAR, v, i, j, l, e) = VOL(o->Ut, NVAR, v, i, j, l, e) + MAT(DH, k, l) * F3(hv, v, i, j, k);
              }
          }
        }
  }
  free(fv);
}
#undef F3


/* ------------------------------------------------------------------------------------------------------------------ */
/* exact functions, boundary states and boundary fluxes: idealgas/exactfunc.f90, idealgas/getboundaryflux.f90         */
/* ------------------------------------------------------------------------------------------------------------------ */
#define BTYPE(ibc, c) o->d.BoundaryType[(size_t)(c)*o->d.nBCs + ((ibc)-1)] /* Fortran BoundaryType(ibc, c+1): c=0 BC_TYPE, 1 BC_STATE */

/* exactfunc.f90:225-362 ExactFunc; cases on the hot path: 1 (constant RefState), 4 (oblique sine wave) */
static void exact_func(const oracle* o, int func, double t, const double* x, double* cons) {
  switch (func) {
    case 1: prim_to_cons(o, o->d.iniRefStatePrim, cons); break;
    case 4: {
      const double Omega = M_PI * o->d.iniFrequency, a = o->d.advVel[0] * 2. * M_PI;
      const double r = 2. + o->d.iniAmplitude * sin(Omega * (x[0] + x[1] + x[2]) - a * t);
      cons[DENS] = cons[MOM1] = cons[MOM2] = cons[MOM3] = r;
      cons[ENER] = cons[DENS] * cons[DENS];
    } break;
    default:
      fprintf(stderr, "oracle: ExactFunc %d not restated\n", func);
      abort();
  }
}

/* eos.f90:561-583 PRESSURE_RIEMANN (state in the face-normal system) */
static double pressure_riemann(const oracle* o, const double* P) {
  const double kappa = o->d.kappa, kappaFac = 2. * kappa * o->sKappaM1, sKappaP1 = 1. / (kappa + 1.);
  if (P[VEL1] <= 0.) return P[PRES] * pow(fmax(0.0001, (1. + 0.5 * o->kappaM1 * P[VEL1] / sqrt(kappa * P[PRES] / P[DENS]))), kappaFac);
  const double ar = 2. * sKappaP1 / P[DENS], br = o->kappaM1 * sKappaP1 * P[PRES];
  return P[PRES] + P[VEL1] / ar * 0.5 * (P[VEL1] + sqrt(P[VEL1] * P[VEL1] + 4. * ar * (P[PRES] + br)));
}

/* getboundaryflux.f90:224-470 GetBoundaryState for one face node; types 2 (Dirichlet), 3/4 (no-slip walls), 9 (slip wall) */
static void get_boundary_state(const oracle* o, int SideID, double t, const double* Pm, const double* nv, const double* t1,
                               const double* t2, const double* xf, double* Pb) {
  const int ibc = o->d.BC[SideID - 1], BCType = BTYPE(ibc, 0), BCState = BTYPE(ibc, 1);
  switch (BCType) {
    case 2:
      if (BCState == 0) {
        double cons[5];
        exact_func(o, o->d.iniExactFunc, t, xf, cons);
        cons_to_prim(o, Pb, cons);
      } else {
        for (int v = 0; v < 6; ++v) Pb[v] = o->d.RefStatePrim[(size_t)(BCState - 1) * 6 + v];
      }
      break;
    case 3: case 4: case 9: {
      Pb[DENS] = Pm[DENS];
      Pb[VEL1] = Pm[VEL1] * nv[0] + Pm[VEL2] * nv[1] + Pm[VEL3] * nv[2];
      Pb[VEL2] = Pm[VEL1] * t1[0] + Pm[VEL2] * t1[1] + Pm[VEL3] * t1[2];
      Pb[VEL3] = Pm[VEL1] * t2[0] + Pm[VEL2] * t2[1] + Pm[VEL3] * t2[2];
      Pb[PRES] = Pm[PRES];
      Pb[TEMP] = Pm[TEMP];
      if (BCType == 3) {
        Pb[PRES] = pressure_riemann(o, Pb);
        Pb[VEL1] = Pb[VEL2] = Pb[VEL3] = 0.;
        Pb[TEMP] = Pm[TEMP];
        Pb[DENS] = Pb[PRES] / (Pb[TEMP] * o->d.R);
      } else if (BCType == 4) {
        Pb[PRES] = pressure_riemann(o, Pb);
        Pb[VEL1] = Pb[VEL2] = Pb[VEL3] = 0.;
        Pb[TEMP] = o->d.RefStatePrim[(size_t)(BCState - 1) * 6 + TEMP];
        Pb[DENS] = Pb[PRES] / (Pb[TEMP] * o->d.R);
      } else {
        Pb[PRES] = pressure_riemann(o, Pb);
        Pb[VEL1] = 0.;
        Pb[DENS] = Pm[DENS];
        Pb[TEMP] = Pb[PRES] / (Pb[DENS] * o->d.R);
      }
      const double vn = Pb[VEL1], vt1 = Pb[VEL2], vt2 = Pb[VEL3];
      for (int c = 0; c < 3; ++c) Pb[VEL1 + c] = vn * nv[c] + vt1 * t1[c] + vt2 * t2[c];
    } break;
    default:
      fprintf(stderr, "oracle: boundary type %d not restated\n", BCType);
      abort();
  }
}

/* getboundaryflux.f90:480-790 GetBoundaryFlux for all BC sides (before the SurfElem scaling of fillflux.f90:179-186) */
static void fill_flux_bc(oracle* o, double t) {
  const int np = o->np, nvl = o->nvl;
  const int solver = (o->d.riemannBC == FLEXI_RIEMANN_SAME) ? o->d.riemann : o->d.riemannBC;
  for (int SideID = 1; SideID <= o->d.nBCSides; ++SideID) {
    const int s = SideID - 1;
    const int BCType = BTYPE(o->d.BC[s], 0);
    for (int q = 0; q < np; ++q)
      for (int p = 0; p < np; ++p) {
        const double* Pm = &SID(o->UPrim_master, NPRIM, 0, p, q, s);
        const double* nv = &SID(o->d.NormVec, 3, 0, p, q, s);
        const double* t1 = &SID(o->d.TangVec1, 3, 0, p, q, s);
        const double* t2 = &SID(o->d.TangVec2, 3, 0, p, q, s);
        const double* xf = o->d.Face_xGP ? &SID(o->d.Face_xGP, 3, 0, p, q, s) : NULL;
        double Pb[6], F[5];
        get_boundary_state(o, SideID, t, Pm, nv, t1, t2, xf, Pb);
        if (BCType == 2) {
          double Ucm[5], Ucb[5];
          prim_to_cons(o, Pm, Ucm);
          prim_to_cons(o, Pb, Ucb);
          riemann_point(o, solver, F, Ucm, Ucb, Pm, Pb, nv, t1, t2);
          if (o->d.parabolic) {
            double Fv[5];
            const double *gx = &SID(o->gradUx_master, nvl, 0, p, q, s), *gy = &SID(o->gradUy_master, nvl, 0, p, q, s),
                         *gz = &SID(o->gradUz_master, nvl, 0, p, q, s);
            viscous_flux_point(o, Fv, Pm, Pb, gx, gy, gz, gx, gy, gz, nv);
            for (int v = 0; v < 5; ++v) F[v] = F[v] + Fv[v];
          }
        } else { /* walls 3,4,9 */
          F[DENS] = 0.;
          for (int c = 0; c < 3; ++c) F[MOM1 + c] = Pb[PRES] * nv[c];
          F[ENER] = 0.;
          if (o->d.parabolic) {
            double f[5], g[5], h[5], gxl[5], gyl[5], gzl[5];
            const double *gx = &SID(o->gradUx_master, nvl, 0, p, q, s), *gy = &SID(o->gradUy_master, nvl, 0, p, q, s),
                         *gz = &SID(o->gradUz_master, nvl, 0, p, q, s);
            if (BCType == 9) { /* BCGradMat = I - n n^T */
              const double B11 = 1. - nv[0] * nv[0], B22 = 1. - nv[1] * nv[1], B33 = 1. - nv[2] * nv[2];
              const double B12 = -nv[0] * nv[1], B13 = -nv[0] * nv[2], B32 = -nv[2] * nv[1];
              for (int v = 0; v < nvl; ++v) {
                gxl[v] = B11 * gx[v] + B12 * gy[v] + B13 * gz[v];
                gyl[v] = B12 * gx[v] + B22 * gy[v] + B32 * gz[v];
                gzl[v] = B13 * gx[v] + B32 * gy[v] + B33 * gz[v];
              }
              gx = gxl; gy = gyl; gz = gzl;
            }
            eval_diff_flux3d(o, Pb, gx, gy, gz, f, g, h);
            if (BCType == 3) f[ENER] = g[ENER] = h[ENER] = 0.;
            for (int v = 1; v < 5; ++v) F[v] = F[v] + nv[0] * f[v] + nv[1] * g[v] + nv[2] * h[v];
          }
        }
        for (int v = 0; v < 5; ++v) SID(o->Flux_master, NVAR, v, p, q, s) = F[v];
      }
  }
}

/* getboundaryflux.f90:855-940 Lifting_GetBoundaryFlux (strong form) into Flux (dir=0) or FluxX/Y/Z (lifting_fillflux.t90:137-186) */
static void lifting_fill_flux_bc(oracle* o, double t, double* Flux, double* FluxX, double* FluxY, double* FluxZ) {
  const int np = o->np, nvl = o->nvl;
  for (int SideID = 1; SideID <= o->d.nBCSides; ++SideID) {
    const int s = SideID - 1;
    const int BCType = BTYPE(o->d.BC[s], 0);
    for (int q = 0; q < np; ++q)
      for (int p = 0; p < np; ++p) {
        const double* Pm = &SID(o->UPrim_master, NPRIM, 0, p, q, s);
        const double* nv = &SID(o->d.NormVec, 3, 0, p, q, s);
        const double* xf = o->d.Face_xGP ? &SID(o->d.Face_xGP, 3, 0, p, q, s) : NULL;
        double Pb[6], Fl[5];
        get_boundary_state(o, SideID, t, Pm, nv, &SID(o->d.TangVec1, 3, 0, p, q, s), &SID(o->d.TangVec2, 3, 0, p, q, s), xf, Pb);
        for (int v = 0; v < nvl; ++v) {
          const int pv = o->lift_map[v];
          if (BCType == 2)
            Fl[v] = 0.5 * (Pm[pv] + Pb[pv]);
          else if (BCType == 3 || BCType == 4)
            Fl[v] = (pv == DENS) ? Pb[DENS] : (pv == TEMP) ? Pb[TEMP] : 0.;
          else
            Fl[v] = (pv == DENS) ? Pm[DENS] : (pv == TEMP) ? Pm[TEMP] : Pb[pv];
          Fl[v] = Fl[v] - Pm[pv];
          Fl[v] = Fl[v] * o->d.SurfElem[((size_t)s * np + q) * np + p];
          if (Flux) SID(Flux, nvl, v, p, q, s) = Fl[v];
          if (FluxX) {
            SID(FluxX, nvl, v, p, q, s) = Fl[v] * nv[0];
            SID(FluxY, nvl, v, p, q, s) = Fl[v] * nv[1];
            SID(FluxZ, nvl, v, p, q, s) = Fl[v] * nv[2];
          }
        }
      }
  }
}

/* exactfunc.f90:727-815 CalcSource, case 4 (manufactured solution of the oblique sine wave): Ut += src/sJ */
static void calc_source(oracle* o, double t) {
  if (o->d.iniExactFunc != 4) {
    fprintf(stderr, "oracle: CalcSource for IniExactFunc %d not restated\n", o->d.iniExactFunc);
    abort();
  }
  const double kappa = o->d.kappa, Amplitude = o->d.iniAmplitude, Omega = M_PI * o->d.iniFrequency, a = o->d.advVel[0] * 2. * M_PI;
  double tmp[6];
  tmp[0] = -a + 3. * Omega;
  tmp[1] = -a + 0.5 * Omega * (1. + kappa * 5.);
  tmp[2] = Amplitude * Omega * o->kappaM1;
  tmp[3] = 0.5 * ((9. + kappa * 15.) * Omega - 8. * a);
  tmp[4] = Amplitude * (3. * Omega * kappa - a);
  tmp[5] = o->d.parabolic ? 3. * o->d.mu0 * kappa * Omega * Omega / o->d.Pr : 0.;
  for (int i = 0; i < 6; ++i) tmp[i] = tmp[i] * Amplitude;
  const double at = a * t;
  const size_t n = (size_t)o->d.nElems * o->np * o->np * o->np;
  for (size_t i = 0; i < n; ++i) {
    const double* x = o->d.Elem_xGP + 3 * i;
    const double arg = Omega * (x[0] + x[1] + x[2]) - at;
    const double cosX = cos(arg), sinX = sin(arg), sinX2 = 2. * sinX * cosX;
    double src[5];
    src[DENS] = tmp[0] * cosX;
    src[MOM1] = src[MOM2] = src[MOM3] = tmp[1] * cosX + tmp[2] * sinX2;
    src[ENER] = tmp[3] * cosX + tmp[4] * sinX2 + tmp[5] * sinX;
    for (int v = 0; v < 5; ++v) o->Ut[i * 5 + v] = o->Ut[i * 5 + v] + src[v] / o->d.sJ[i];
  }
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* FillFlux: dg/fillflux.f90:40-215 (inner and MPI_MINE sides; BC sides handled in fill_flux_bc)                      */
/* ------------------------------------------------------------------------------------------------------------------ */

static void fill_flux(oracle* o, double t, int doMPISides) {
  const int np = o->np, nvl = o->nvl;
  int firstSideID_wo_BC, firstSideID, lastSideID;
  if (doMPISides) {
    firstSideID_wo_BC = o->firstMPISide_MINE;
    firstSideID = o->firstMPISide_MINE;
    lastSideID = o->lastMPISide_MINE;
  } else {
    firstSideID_wo_BC = o->firstInnerSide;
    firstSideID = o->firstBCSide;
    lastSideID = o->lastInnerSide;
  }
  for (int SideID = firstSideID_wo_BC; SideID <= lastSideID; ++SideID) {
    const int s = SideID - 1;
    for (int q = 0; q < np; ++q)
      for (int p = 0; p < np; ++p) {
        double F[5];
        const double* nv = &SID(o->d.NormVec, 3, 0, p, q, s);
        riemann_point(o, o->d.riemann, F, &SID(o->U_master, NVAR, 0, p, q, s), &SID(o->U_slave, NVAR, 0, p, q, s),
                      &SID(o->UPrim_master, NPRIM, 0, p, q, s), &SID(o->UPrim_slave, NPRIM, 0, p, q, s), nv,
                      &SID(o->d.TangVec1, 3, 0, p, q, s), &SID(o->d.TangVec2, 3, 0, p, q, s));
        if (o->d.parabolic) {
          double Fv[5];
          viscous_flux_point(o, Fv, &SID(o->UPrim_master, NPRIM, 0, p, q, s), &SID(o->UPrim_slave, NPRIM, 0, p, q, s),
                             &SID(o->gradUx_master, nvl, 0, p, q, s), &SID(o->gradUy_master, nvl, 0, p, q, s),
                             &SID(o->gradUz_master, nvl, 0, p, q, s), &SID(o->gradUx_slave, nvl, 0, p, q, s),
                             &SID(o->gradUy_slave, nvl, 0, p, q, s), &SID(o->gradUz_slave, nvl, 0, p, q, s), nv);
          for (int v = 0; v < 5; ++v) F[v] = F[v] + Fv[v];
        }
        for (int v = 0; v < 5; ++v) SID(o->Flux_master, NVAR, v, p, q, s) = F[v];
      }
  }
  if (!doMPISides && o->d.nBCSides > 0) fill_flux_bc(o, t);
  for (int SideID = firstSideID; SideID <= lastSideID; ++SideID) {
    const int s = SideID - 1;
    for (int q = 0; q < np; ++q)
      for (int p = 0; p < np; ++p) {
        double surf = o->d.SurfElem[((size_t)s * np + q) * np + p];
        for (int v = 0; v < 5; ++v) {
          SID(o->Flux_master, NVAR, v, p, q, s) = SID(o->Flux_master, NVAR, v, p, q, s) * surf;
          SID(o->Flux_slave, NVAR, v, p, q, s) = SID(o->Flux_master, NVAR, v, p, q, s);
        }
      }
  }
}


/* ------------------------------------------------------------------------------------------------------------------ */
/* source term / sign / Jacobian: dg.f90:656-676                                                                      */
/* ------------------------------------------------------------------------------------------------------------------ */
/* filter.f90:264-288 Filter -> changeBasis.t90 ChangeBasis3D(nVar,N,N,FilterMat,U): the 1-D matrix along xi, eta, zeta, every sum
   accumulated from zero in ascending input index */
static void filter_volume(const oracle* o, const double* Mat, double* U) {
  const int np = o->np;
  const size_t ne = (size_t)np * np * np * NVAR;
  double* b1 = (double*)xcalloc(ne, 8);
  double* b2 = (double*)xcalloc(ne, 8);
#define E5(A, v, i, j, k) (A)[(v) + NVAR * ((i) + (size_t)np * ((j) + (size_t)np * (k)))]
#define FM(io, ii) Mat[(io) + (size_t)np * (ii)]
  for (int e = 0; e < o->d.nElems; ++e) {
    double* Ue = U + (size_t)e * ne;
    memset(b1, 0, ne * 8);
    for (int k = 0; k < np; ++k)
      for (int j = 0; j < np; ++j)
        for (int ii = 0; ii < np; ++ii)
          for (int io = 0; io < np; ++io)
            for (int v = 0; v < NVAR; ++v) E5(b1, v, io, j, k) = E5(b1, v, io, j, k) + FM(io, ii) * E5(Ue, v, ii, j, k);
    memset(b2, 0, ne * 8);
    for (int k = 0; k < np; ++k)
      for (int ji = 0; ji < np; ++ji)
        for (int jo = 0; jo < np; ++jo)
          for (int i = 0; i < np; ++i)
            for (int v = 0; v < NVAR; ++v) E5(b2, v, i, jo, k) = E5(b2, v, i, jo, k) + FM(jo, ji) * E5(b1, v, i, ji, k);
    memset(Ue, 0, ne * 8);
    for (int ki = 0; ki < np; ++ki)
      for (int ko = 0; ko < np; ++ko)
        for (int j = 0; j < np; ++j)
          for (int i = 0; i < np; ++i)
            for (int v = 0; v < NVAR; ++v) E5(Ue, v, i, j, ko) = E5(Ue, v, i, j, ko) + FM(ko, ki) * E5(b2, v, i, j, ki);
  }
#undef E5
#undef FM
  free(b1);
  free(b2);
}

/* filter.f90:295-448 Filter_LAF: per element, the cut-off filtered state, the kinetic energy of the filtered-out velocity scales
   and of the resolved fluctuations, their ratio against an adaptive limit; U is replaced by its filtered version only where the
   ratio exceeds the limit.  Loop and summation order as in the Fortran. */
static void filter_laf(oracle* o) {
  const int np = o->np;
  const size_t n3 = (size_t)np * np * np, ne = n3 * NVAR;
  double* Uf = (double*)xcalloc(ne, 8);
  for (int e = 0; e < o->d.nElems; ++e) {
    double* Ue = o->U + (size_t)e * ne;
    memcpy(Uf, Ue, ne * 8);
    { /* U_filtered = Filter(U_in) of this element (same three sweeps as filter_volume) */
      oracle one = *o;
      one.d.nElems = 1;
      filter_volume(&one, o->FilterMat, Uf);
    }
    const double* iw = o->laf_iw + (size_t)e * n3;
    double Uavg[3] = {0., 0., 0.};
    for (size_t n = 0; n < n3; ++n)
      for (int c = 0; c < 3; ++c) Uavg[c] = Uavg[c] + Uf[n * NVAR + 1 + c] / Ue[n * NVAR] * iw[n];
    double ekin_avg = 0., ekin_fluc_avg = 0.;
    for (size_t n = 0; n < n3; ++n) {
      const double rho = Ue[n * NVAR];
      const double f1 = Ue[n * NVAR + 1] - Uf[n * NVAR + 1], f2 = Ue[n * NVAR + 2] - Uf[n * NVAR + 2], f3 = Ue[n * NVAR + 3] - Uf[n * NVAR + 3];
      ekin_fluc_avg = ekin_fluc_avg + (f1 * f1 / (rho * rho) + f3 * f3 / (rho * rho) + f2 * f2 / (rho * rho)) * iw[n];
      const double a1 = Uf[n * NVAR + 1] - Uavg[0], a2 = Uf[n * NVAR + 2] - Uavg[1], a3 = Uf[n * NVAR + 3] - Uavg[2];
      ekin_avg = ekin_avg + (a1 * a1 / (rho * rho) + a2 * a2 / (rho * rho) + a3 * a3 / (rho * rho)) * iw[n];
    }
    o->laf_eRatio[e] = (ekin_fluc_avg / ekin_avg) * o->laf_normMod;
    o->laf_lim[e] = fmax(o->laf_lim[e], 1.0E-8);
    const double dfluc = ekin_fluc_avg - o->laf_fluc_old[e];
    const double dedt_lim = fabs(ekin_avg - o->laf_ekin_old[e]) / dfluc;
    if (dedt_lim > 1.0 && dfluc > 0.) {
      o->laf_lim[e] = o->laf_eRatio[e];
    } else if (dfluc < 0.) {
      o->laf_lim[e] = o->laf_eRatio[e];
    } else if (o->laf_eRatio[e] > o->laf_lim[e] * o->laf_alpha) {
      memcpy(Ue, Uf, ne * 8);
      o->laf_r[e] = o->laf_r[e] + 1.;
    }
    o->laf_ekin_old[e] = ekin_avg;
    o->laf_fluc_old[e] = ekin_fluc_avg;
  }
  free(Uf);
}

/* overintegration.f90:181-330 FilterConservative: JU_t on N -> projection to NUnder (ChangeBasisVolume with Vdm_N_NUnder) -> times
   sJNUnder -> interpolation back to N (Vdm_NUnder_N).  Sums run over ascending input index like ChangeBasisVolume. */
static void filter_conservative(const oracle* o, double* U) {
  const int np = o->np, nu = o->NUnder + 1;
  const size_t ne = (size_t)np * np * np * NVAR;
  double* b1 = (double*)xcalloc((size_t)nu * np * np * NVAR, 8);
  double* b2 = (double*)xcalloc((size_t)nu * nu * np * NVAR, 8);
  double* ul = (double*)xcalloc((size_t)nu * nu * nu * NVAR, 8);
  double* b3 = (double*)xcalloc((size_t)np * nu * nu * NVAR, 8);
  double* b4 = (double*)xcalloc((size_t)np * np * nu * NVAR, 8);
#define DN(iu, i) o->Vdm_N_NUnder[(iu) + (size_t)nu * (i)]
#define UP(i, iu) o->Vdm_NUnder_N[(i) + (size_t)np * (iu)]
#define X5(A, n0, n1, v, a, b, c) (A)[(v) + NVAR * ((a) + (size_t)(n0) * ((b) + (size_t)(n1) * (c)))]
  for (int e = 0; e < o->d.nElems; ++e) {
    double* Ue = U + (size_t)e * ne;
    for (int k = 0; k < np; ++k)
      for (int j = 0; j < np; ++j)
        for (int iu = 0; iu < nu; ++iu)
          for (int v = 0; v < NVAR; ++v) {
            double s = DN(iu, 0) * X5(Ue, np, np, v, 0, j, k);
            for (int i = 1; i < np; ++i) s = s + DN(iu, i) * X5(Ue, np, np, v, i, j, k);
            X5(b1, nu, np, v, iu, j, k) = s;
          }
    for (int k = 0; k < np; ++k)
      for (int ju = 0; ju < nu; ++ju)
        for (int iu = 0; iu < nu; ++iu)
          for (int v = 0; v < NVAR; ++v) {
            double s = DN(ju, 0) * X5(b1, nu, np, v, iu, 0, k);
            for (int j = 1; j < np; ++j) s = s + DN(ju, j) * X5(b1, nu, np, v, iu, j, k);
            X5(b2, nu, nu, v, iu, ju, k) = s;
          }
    for (int ku = 0; ku < nu; ++ku)
      for (int ju = 0; ju < nu; ++ju)
        for (int iu = 0; iu < nu; ++iu)
          for (int v = 0; v < NVAR; ++v) {
            double s = 0.;
            for (int k = 0; k < np; ++k) s = s + DN(ku, k) * X5(b2, nu, nu, v, iu, ju, k);
            X5(ul, nu, nu, v, iu, ju, ku) = s * o->sJNUnder[((size_t)e * nu + ku) * nu * nu + (size_t)ju * nu + iu];
          }
    for (int ku = 0; ku < nu; ++ku)
      for (int ju = 0; ju < nu; ++ju)
        for (int i = 0; i < np; ++i)
          for (int v = 0; v < NVAR; ++v) {
            double s = UP(i, 0) * X5(ul, nu, nu, v, 0, ju, ku);
            for (int iu = 1; iu < nu; ++iu) s = s + UP(i, iu) * X5(ul, nu, nu, v, iu, ju, ku);
            X5(b3, np, nu, v, i, ju, ku) = s;
          }
    for (int ku = 0; ku < nu; ++ku)
      for (int j = 0; j < np; ++j)
        for (int i = 0; i < np; ++i)
          for (int v = 0; v < NVAR; ++v) {
            double s = UP(j, 0) * X5(b3, np, nu, v, i, 0, ku);
            for (int ju = 1; ju < nu; ++ju) s = s + UP(j, ju) * X5(b3, np, nu, v, i, ju, ku);
            X5(b4, np, np, v, i, j, ku) = s;
          }
    for (int k = 0; k < np; ++k)
      for (int j = 0; j < np; ++j)
        for (int i = 0; i < np; ++i)
          for (int v = 0; v < NVAR; ++v) {
            double s = 0.;
            for (int ku = 0; ku < nu; ++ku) s = s + UP(k, ku) * X5(b4, np, np, v, i, j, ku);
            X5(Ue, np, np, v, i, j, k) = s;
          }
  }
#undef DN
#undef UP
#undef X5
  free(b1);
  free(b2);
  free(ul);
  free(b3);
  free(b4);
}

/* sponge.f90:612-678 Sponge(Ut): Ut = Ut - SpongeMat*(U - SpBaseFlow) on the sponge elements (DG elements; SpongeMat already
   carries the Jacobian, so this happens in reference space, dg.f90:659-661) */
static void sponge(oracle* o) {
  const size_t np3 = (size_t)o->np * o->np * o->np;
  for (int is = 0; is < o->nSpongeElems; ++is) {
    const size_t e = (size_t)o->SpongeMap[is] - 1;
    for (size_t n = 0; n < np3; ++n) {
      const double sm = o->SpongeMat[(size_t)is * np3 + n];
      for (int v = 0; v < NVAR; ++v) {
        const size_t k = (e * np3 + n) * NVAR + v;
        o->Ut[k] = o->Ut[k] - sm * (o->U[k] - o->SpBaseFlow[k]);
      }
    }
  }
}

static void finish_ut(oracle* o, double t) {
  const size_t n = (size_t)o->d.nElems * o->np * o->np * o->np;
  for (size_t i = 0; i < n * NVAR; ++i) o->Ut[i] = -o->Ut[i];
  if (o->d.doCalcSource) calc_source(o, t);
  if (o->nSpongeElems > 0) sponge(o);
  /* dg.f90:663-676: OVERINTEGRATIONTYPE_CUTOFF filters Ut before the Jacobian is applied (dg/overintegration.f90:162-187) */
  if (o->NUnder >= 0) { /* OVERINTEGRATIONTYPE_CONSCUTOFF: the filter applies sJ on NUnder itself, no ApplyJacobian for DG elements */
    filter_conservative(o, o->Ut);
    return;
  }
  if (o->OverintegrationMat) filter_volume(o, o->OverintegrationMat, o->Ut);
  apply_jacobian(o, NVAR, o->Ut);
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* world = the ranks of one emulated MPI job                                                                          */
/* ------------------------------------------------------------------------------------------------------------------ */
typedef struct oracle_world {
  int n;
  oracle** r;
} oracle_world;

/* mpi.f90:244-445: SendID=1 sends MINE slices (received as YOUR), SendID=2 sends YOUR slices (received as MINE).
 * `off` is the byte offset of the exchanged array inside struct oracle so that the same array is addressed on every rank. */
static void exchange(oracle_world* w, size_t member_off, int dataSizeSide, int SendID) {
  for (int r = 0; r < w->n; ++r) {
    oracle* src = w->r[r];
    double* a_src = *(double**)((char*)src + member_off);
    for (int i = 0; i < src->d.nNbProcs; ++i) {
      int nb = src->d.NbProc[i];
      oracle* dst = w->r[nb];
      double* a_dst = *(double**)((char*)dst + member_off);
      int j = -1;
      for (int jj = 0; jj < dst->d.nNbProcs; ++jj)
        if (dst->d.NbProc[jj] == r) j = jj;
      if (j < 0) {
        fprintf(stderr, "oracle: asymmetric neighbour tables\n");
        abort();
      }
      int s0, n, d0, nd;
      if (SendID == 1) {
        s0 = src->d.offsetMPISides_MINE[i];
        n = src->d.nMPISides_MINE_Proc[i];
        d0 = dst->d.offsetMPISides_YOUR[j];
        nd = dst->d.nMPISides_YOUR_Proc[j];
      } else {
        s0 = src->d.offsetMPISides_YOUR[i];
        n = src->d.nMPISides_YOUR_Proc[i];
        d0 = dst->d.offsetMPISides_MINE[j];
        nd = dst->d.nMPISides_MINE_Proc[j];
      }
      if (n != nd) {
        fprintf(stderr, "oracle: message size mismatch %d vs %d\n", n, nd);
        abort();
      }
      memcpy(a_dst + (size_t)d0 * dataSizeSide, a_src + (size_t)s0 * dataSizeSide, (size_t)n * dataSizeSide * sizeof(double));
    }
  }
}
#define OFF(m) ((size_t)((char*)&((oracle*)0)->m - (char*)0))
#define FOR_RANKS _Pragma("omp parallel for schedule(static)") for (int r = 0; r < w->n; ++r)

/* lifting_br1.t90:48-181 Lifting_BR1 (strong, non-conservative) */
static void lifting_br1_world(oracle_world* w, double t) {
  const int np = w->r[0]->np, nvl = w->r[0]->nvl;
  const int szg = nvl * np * np;
  FOR_RANKS lifting_fill_flux(w->r[r], 0, w->r[r]->gradUz_slave, 1);
  exchange(w, OFF(gradUz_slave), szg, 1);
  FOR_RANKS {
    oracle* o = w->r[r];
    lifting_volint_nonconservative(o);
    lifting_fill_flux_bc(o, t, o->gradUz_slave, NULL, NULL, NULL);
    lifting_fill_flux(o, 0, o->gradUz_slave, 0);
    lifting_fill_flux_normvec(o, o->gradUz_slave, o->gradUx_master, o->gradUy_master, o->gradUz_master, 0);
    surf_int(o, nvl, o->gradUx_master, o->gradUx_master, o->gradUx, 0, o->d.L_HatMinus, o->d.L_HatPlus, 1.);
    surf_int(o, nvl, o->gradUy_master, o->gradUy_master, o->gradUy, 0, o->d.L_HatMinus, o->d.L_HatPlus, 1.);
    surf_int(o, nvl, o->gradUz_master, o->gradUz_master, o->gradUz, 0, o->d.L_HatMinus, o->d.L_HatPlus, 1.);
    lifting_fill_flux_normvec(o, o->gradUz_slave, o->gradUx_master, o->gradUy_master, o->gradUz_master, 1);
    surf_int(o, nvl, o->gradUx_master, o->gradUx_master, o->gradUx, 1, o->d.L_HatMinus, o->d.L_HatPlus, 1.);
    surf_int(o, nvl, o->gradUy_master, o->gradUy_master, o->gradUy, 1, o->d.L_HatMinus, o->d.L_HatPlus, 1.);
    surf_int(o, nvl, o->gradUz_master, o->gradUz_master, o->gradUz, 1, o->d.L_HatMinus, o->d.L_HatPlus, 1.);
    apply_jacobian(o, nvl, o->gradUx);
    apply_jacobian(o, nvl, o->gradUy);
    apply_jacobian(o, nvl, o->gradUz);
    prolong_to_face(o, nvl, o->gradUx, o->gradUx_master, o->gradUx_slave, 1);
    prolong_to_face(o, nvl, o->gradUy, o->gradUy_master, o->gradUy_slave, 1);
    prolong_to_face(o, nvl, o->gradUz, o->gradUz_master, o->gradUz_slave, 1);
  }
  exchange(w, OFF(gradUx_slave), szg, 2);
  exchange(w, OFF(gradUy_slave), szg, 2);
  exchange(w, OFF(gradUz_slave), szg, 2);
  FOR_RANKS {
    oracle* o = w->r[r];
    prolong_to_face(o, nvl, o->gradUx, o->gradUx_master, o->gradUx_slave, 0);
    prolong_to_face(o, nvl, o->gradUy, o->gradUy_master, o->gradUy_slave, 0);
    prolong_to_face(o, nvl, o->gradUz, o->gradUz_master, o->gradUz_slave, 0);
  }
}


/* lifting_br2.t90:194-311 Lifting_SurfInt_BR2 for one gradient direction */
static void lifting_surf_int_br2(oracle* o, const double* Flux, double* gradU, double* gradU_master, double* gradU_slave, int doMPISides) {
  const int np = o->np, N = np - 1, nvl = o->nvl;
  const int gl = (o->d.nodeType == FLEXI_NODETYPE_GAUSS_LOBATTO);
  const int lmax = gl ? 0 : N;
  for (int pass = 0; pass < 2; ++pass) { /* 0: slave sides, 1: master sides */
    int firstSideID, lastSideID;
    if (pass == 0) {
      firstSideID = doMPISides ? o->firstMPISide_YOUR : o->firstInnerSide;
      lastSideID = doMPISides ? o->lastMPISide_YOUR : o->lastInnerSide;
    } else {
      if (doMPISides) continue; /* only mortar MPI sides: none */
      firstSideID = o->firstBCSide;
      lastSideID = o->lastMPISide_MINE;
    }
    for (int SideID = firstSideID; SideID <= lastSideID; ++SideID) {
      const int s = SideID - 1;
      double eta = o->d.etaBR2;
      if (pass == 1 && SideID <= o->d.nBCSides) {
        const int bt = BTYPE(o->d.BC[s], 0);
        if (bt == 3 || bt == 4) eta = o->d.etaBR2_wall;
      }
      const int ElemID = (pass == 0) ? S2E(1, s) : S2E(0, s);
      if (ElemID <= 0) continue;
      const int locSide = (pass == 0) ? S2E(3, s) : S2E(2, s);
      const int flip = (pass == 0) ? S2E(4, s) : 0;
      double* gface = (pass == 0) ? gradU_slave : gradU_master;
      const int minus = (locSide == XI_MINUS || locSide == ETA_MINUS || locSide == ZETA_MINUS);
      for (int q = 0; q < np; ++q)
        for (int p = 0; p < np; ++p) {
          const int a = S2V2(0, p, q, flip, locSide), b = S2V2(1, p, q, flip, locSide);
          for (int l = 0; l <= lmax; ++l) {
            const int d = minus ? l : N - l; /* S2V(:,l,p,q,flip,locSide), mappings.f90:336-387 */
            int i, j, k;
            if (locSide == XI_MINUS || locSide == XI_PLUS) { i = d; j = a; k = b; }
            else if (locSide == ETA_MINUS || locSide == ETA_PLUS) { i = a; j = d; k = b; }
            else { i = a; j = b; k = d; }
            const double sJ = o->d.sJ[(((size_t)(ElemID - 1) * np + k) * np + j) * np + i];
            for (int v = 0; v < nvl; ++v) {
              const double F_loc = sJ * SID(Flux, nvl, v, p, q, s) * o->d.L_HatMinus[l];
              VOL(gradU, nvl, v, i, j, k, ElemID - 1) = VOL(gradU, nvl, v, i, j, k, ElemID - 1) + F_loc;
              SID(gface, nvl, v, p, q, s) = SID(gface, nvl, v, p, q, s) + eta * o->d.L_Minus[l] * F_loc;
            }
          }
        }
    }
  }
}

/* lifting_br2.t90:44-185 Lifting_BR2 */
static void lifting_br2_world(oracle_world* w, double t) {
  const int np = w->r[0]->np, nvl = w->r[0]->nvl;
  const int szg = nvl * np * np;
  FOR_RANKS {
    oracle* o = w->r[r];
    lifting_fill_flux(o, 1, o->FluxX, 1);
    lifting_fill_flux(o, 2, o->FluxY, 1);
    lifting_fill_flux(o, 3, o->FluxZ, 1);
  }
  exchange(w, OFF(FluxX), szg, 1);
  exchange(w, OFF(FluxY), szg, 1);
  exchange(w, OFF(FluxZ), szg, 1);
  FOR_RANKS {
    oracle* o = w->r[r];
    lifting_fill_flux_bc(o, t, NULL, o->FluxX, o->FluxY, o->FluxZ);
    lifting_fill_flux(o, 1, o->FluxX, 0);
    lifting_fill_flux(o, 2, o->FluxY, 0);
    lifting_fill_flux(o, 3, o->FluxZ, 0);
    lifting_volint_nonconservative(o);
    apply_jacobian(o, nvl, o->gradUx);
    apply_jacobian(o, nvl, o->gradUy);
    apply_jacobian(o, nvl, o->gradUz);
    for (int mpi = 1; mpi >= 0; --mpi) {
      prolong_to_face(o, nvl, o->gradUx, o->gradUx_master, o->gradUx_slave, mpi);
      prolong_to_face(o, nvl, o->gradUy, o->gradUy_master, o->gradUy_slave, mpi);
      prolong_to_face(o, nvl, o->gradUz, o->gradUz_master, o->gradUz_slave, mpi);
    }
    for (int mpi = 1; mpi >= 0; --mpi) {
      lifting_surf_int_br2(o, o->FluxX, o->gradUx, o->gradUx_master, o->gradUx_slave, mpi);
      lifting_surf_int_br2(o, o->FluxY, o->gradUy, o->gradUy_master, o->gradUy_slave, mpi);
      lifting_surf_int_br2(o, o->FluxZ, o->gradUz, o->gradUz_master, o->gradUz_slave, mpi);
    }
  }
  exchange(w, OFF(gradUx_slave), szg, 2);
  exchange(w, OFF(gradUy_slave), szg, 2);
  exchange(w, OFF(gradUz_slave), szg, 2);
}

/* PP_EntropyVars = 1 <=> SPLIT_DG on Gauss nodes (src/CMakeLists.txt:64-71) */
static int entropy_vars(const oracle* o) { return o->d.splitDG && o->d.nodeType == FLEXI_NODETYPE_GAUSS; }

/* eos.f90:423-441 ConsToPrimToEntropy_Volume (UPrim is already there) */
static void cons_to_entropy_volume(oracle* o) {
  const size_t n = (size_t)o->d.nElems * o->np * o->np * o->np;
  for (size_t i = 0; i < n; ++i) prim_to_entropy(o, o->V + i * NVAR, o->UPrim + i * NPRIM);
}

/* prolongtoface.t90:151-204 ProlongToFaceEntropy: prolong the entropy variables, convert back to conservative on the faces */
static void prolong_to_face_entropy(oracle* o, int doMPISides) {
  const int np = o->np;
  const size_t fs = (size_t)np * np;
  prolong_to_face(o, NVAR, o->V, o->V_master, o->V_slave, doMPISides);
  const int firstSideID = doMPISides ? o->firstMPISide_YOUR : 1;
  const int lastSideID = doMPISides ? o->d.nSides : o->lastMPISide_MINE;
  for (int SideID = firstSideID; SideID <= lastSideID; ++SideID) {
    const int s = SideID - 1;
    if (S2E(0, s) > 0)
      for (size_t i = 0; i < fs; ++i) entropy_to_cons(o, o->V_master + ((size_t)s * fs + i) * NVAR, o->U_master + ((size_t)s * fs + i) * NVAR);
    if (S2E(1, s) > 0)
      for (size_t i = 0; i < fs; ++i) entropy_to_cons(o, o->V_slave + ((size_t)s * fs + i) * NVAR, o->U_slave + ((size_t)s * fs + i) * NVAR);
  }
}

/* dg.f90:221-678 DGTimeDerivative_weakForm */
void oracle_world_dg_time_derivative(oracle_world* w, double t) {
  const int np = w->r[0]->np;
  const int szs = NVAR * np * np;
  FOR_RANKS {
    oracle* o = w->r[r];
    if (o->FilterMat && o->laf) filter_laf(o);               /* 1. dg.f90:329-330 with Filter_Pointer => Filter_LAF */
    else if (o->FilterMat) filter_volume(o, o->FilterMat, o->U);  /* 1. dg.f90:329-330: Filter_Pointer(U,FilterMat), in place */
    cons_to_prim_volume(o);                                  /* 2. */
    if (entropy_vars(o)) {
      cons_to_entropy_volume(o);                             /* ConsToPrimToEntropy, dg.f90:333-337 */
      prolong_to_face_entropy(o, 1);                         /* dg.f90:371-375 */
    } else {
      prolong_to_face(o, NVAR, o->U, o->U_master, o->U_slave, 1); /* 3. MPI YOUR sides first */
    }
  }
  exchange(w, OFF(U_slave), szs, 2);
  FOR_RANKS {
    oracle* o = w->r[r];
    if (entropy_vars(o))
      prolong_to_face_entropy(o, 0);                         /* dg.f90:409-413 */
    else
      prolong_to_face(o, NVAR, o->U, o->U_master, o->U_slave, 0);
    get_primitive_state_surface(o); /* 4. */
  }
  if (w->r[0]->d.parabolic) { /* 6. */
    if (w->r[0]->d.lifting == 1)
      lifting_br1_world(w, t);
    else
      lifting_br2_world(w, t);
  }
  FOR_RANKS { /* 8. */
    oracle* o = w->r[r];
    if (o->d.splitDG)
      volint_split_form(o);
    else
      volint_weak_form(o);
    fill_flux(o, t, 1); /* 11. MPI MINE */
  }
  exchange(w, OFF(Flux_slave), szs, 1);
  FOR_RANKS {
    oracle* o = w->r[r];
    fill_flux(o, t, 0);
    if (entropy_vars(o)) { /* SPLIT_DG on Gauss nodes: surface integral with the two-point correction, surfint.t90:90-137,190-209 */
      surf_int_split_gauss(o, o->Flux_master, o->Flux_slave, o->Ut, 0);
      surf_int_split_gauss(o, o->Flux_master, o->Flux_slave, o->Ut, 1);
    } else {
      surf_int(o, NVAR, o->Flux_master, o->Flux_slave, o->Ut, 0, o->d.L_HatMinus, o->d.L_HatPlus, -1.);
      surf_int(o, NVAR, o->Flux_master, o->Flux_slave, o->Ut, 1, o->d.L_HatMinus, o->d.L_HatPlus, -1.);
    }
    finish_ut(o, t); /* 12.-14. */
  }
}

/* timestep.f90:82-152 TimeStepByLSERKW2 */
void oracle_world_timestep(oracle_world* w, double t, double dt) {
  const flexi_gpu_desc* d0 = &w->r[0]->d;
  for (int iStage = 1; iStage <= d0->nRKStages; ++iStage) {
    double tStage = (iStage == 1) ? t : t + d0->RKc[iStage - 2] * dt;
    oracle_world_dg_time_derivative(w, tStage);
    FOR_RANKS {
      oracle* o = w->r[r];
      const size_t n = (size_t)o->d.nElems * o->np * o->np * o->np * NVAR;
      const double b_dt = o->d.RKb[iStage - 1] * dt;
      if (iStage == 1) {
        for (size_t i = 0; i < n; ++i) o->Ut_tmp[i] = o->Ut[i];
      } else {
        const double mA = -o->d.RKA[iStage - 2];
        for (size_t i = 0; i < n; ++i) o->Ut_tmp[i] = o->Ut_tmp[i] * mA + o->Ut[i];
      }
      for (size_t i = 0; i < n; ++i) o->U[i] = o->U[i] + o->Ut_tmp[i] * b_dt;
    }
  }
}

/* calctimestep.f90:42-83 InitCalctimestep */
static void init_calc_timestep(oracle* o) {
  const int np = o->np;
  const size_t n = (size_t)o->d.nElems * np * np * np;
  o->MetricsAdv = (double*)xcalloc(n * 3, sizeof(double));
  o->MetricsVisc = (double*)xcalloc(n * 3, sizeof(double));
  const double KappasPr_max = fmax(4. / 3., o->KappasPr);
  for (size_t i = 0; i < n; ++i) {
    const double* M[3] = {o->d.Metrics_fTilde + 3 * i, o->d.Metrics_gTilde + 3 * i, o->d.Metrics_hTilde + 3 * i};
    const double sJ = o->d.sJ[i];
    for (int dd = 0; dd < 3; ++dd) {
      o->MetricsAdv[3 * i + dd] = sJ * sqrt(M[dd][0] * M[dd][0] + M[dd][1] * M[dd][1] + M[dd][2] * M[dd][2]);
      o->MetricsVisc[3 * i + dd] =
          KappasPr_max * ((M[dd][0] * sJ) * (M[dd][0] * sJ) + (M[dd][1] * sJ) * (M[dd][1] * sJ) + (M[dd][2] * sJ) * (M[dd][2] * sJ));
    }
  }
}

/* calctimestep.f90:89-227 CalcTimeStep (per rank part) */
static void calc_timestep_rank(oracle* o, double* TimeStep3) {
  const int np = o->np;
  const size_t ne = (size_t)np * np * np;
  int errType = 0;
  double TimeStepConv = HUGE_VAL, TimeStepVisc = HUGE_VAL;
  for (int e = 0; e < o->d.nElems; ++e) {
    double Max_Lambda[3] = {0, 0, 0}, Max_Lambda_v[3] = {0, 0, 0};
    for (size_t n = 0; n < ne; ++n) {
      const size_t i = (size_t)e * ne + n;
      double UE[11];
      for (int v = 0; v < 5; ++v) UE[v] = o->U[i * 5 + v];
      UE[EXT_SRHO] = 1. / UE[EXT_DENS];
      UE[EXT_VEL1] = UE[EXT_MOM1] * UE[EXT_SRHO];
      UE[EXT_VEL2] = UE[EXT_MOM2] * UE[EXT_SRHO];
      UE[EXT_VEL3] = UE[EXT_MOM3] * UE[EXT_SRHO];
      UE[EXT_PRES] = o->kappaM1 * (UE[EXT_ENER] - 0.5 * (UE[EXT_VEL1] * UE[EXT_MOM1] + UE[EXT_VEL2] * UE[EXT_MOM2] + UE[EXT_VEL3] * UE[EXT_MOM3]));
      UE[EXT_TEMP] = UE[EXT_PRES] * UE[EXT_SRHO] / o->d.R;
      if (isnan(UE[EXT_DENS])) errType = 1;
      double c = speed_of_sound_he(o, UE);
      double sJ = o->d.sJ[i];
      double vsJ[3] = {UE[EXT_VEL1] * sJ, UE[EXT_VEL2] * sJ, UE[EXT_VEL3] * sJ};
      const double* M[3] = {o->d.Metrics_fTilde + 3 * i, o->d.Metrics_gTilde + 3 * i, o->d.Metrics_hTilde + 3 * i};
      for (int dd = 0; dd < 3; ++dd)
        Max_Lambda[dd] = fmax(Max_Lambda[dd], fabs(M[dd][0] * vsJ[0] + M[dd][1] * vsJ[1] + M[dd][2] * vsJ[2]) + c * o->MetricsAdv[3 * i + dd]);
      if (o->d.parabolic) {
        double mu = viscosity_T(o, UE[EXT_TEMP]);
        for (int dd = 0; dd < 3; ++dd) Max_Lambda_v[dd] = fmax(Max_Lambda_v[dd], mu * UE[EXT_SRHO] * o->MetricsVisc[3 * i + dd]);
      }
    }
    double dtElem = o->d.CFLScale * 2. / (Max_Lambda[0] + Max_Lambda[1] + Max_Lambda[2]);
    TimeStepConv = fmin(TimeStepConv, dtElem);
    if (isnan(TimeStepConv) || isnan(dtElem)) errType = 2;
    if (o->d.parabolic) {
      double sv = Max_Lambda_v[0] + Max_Lambda_v[1] + Max_Lambda_v[2];
      if (sv > 0.) TimeStepVisc = fmin(TimeStepVisc, o->d.DFLScale * 4. / sv);
      if (isnan(TimeStepVisc) || isnan(sv)) errType = 3;
    }
  }
  TimeStep3[0] = TimeStepConv;
  TimeStep3[1] = TimeStepVisc;
  TimeStep3[2] = -(double)errType;
}

double oracle_world_calc_timestep(oracle_world* w, double* dt_conv, double* dt_visc, int* errType) {
  double ts[3] = {HUGE_VAL, HUGE_VAL, HUGE_VAL};
  double* loc = (double*)xcalloc((size_t)3 * w->n, sizeof(double));
  FOR_RANKS calc_timestep_rank(w->r[r], loc + 3 * r);
  for (int r = 0; r < w->n; ++r) /* MPI_ALLREDUCE(MIN), calctimestep.f90:221 */
    for (int c = 0; c < 3; ++c) ts[c] = fmin(ts[c], loc[3 * r + c]);
  free(loc);
  *errType = (int)(-ts[2]);
  *dt_conv = ts[0];
  *dt_visc = ts[1];
  return fmin(ts[0], ts[1]);
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* construction / access                                                                                              */
/* ------------------------------------------------------------------------------------------------------------------ */
oracle* oracle_create(const flexi_gpu_desc* d) {
  if (d->abi_version != FLEXI_GPU_ABI_VERSION || d->struct_bytes != (int)sizeof(flexi_gpu_desc)) {
    fprintf(stderr, "oracle: descriptor ABI mismatch (%d bytes vs %d)\n", d->struct_bytes, (int)sizeof(flexi_gpu_desc));
    return NULL;
  }
  oracle* o = (oracle*)xcalloc(1, sizeof(oracle));
  o->NUnder = -1; /* no conservative cut-off */
  o->d = *d;
  const int np = d->N + 1;
  o->np = np;
  o->nvl = d->nVarLifting;
  if (o->nvl == 5) {
    int m[5] = {0, 1, 2, 3, 5};
    memcpy(o->lift_map, m, sizeof m);
    o->lv_vel1 = 1;
    o->lv_temp = 4;
  } else {
    int m[5] = {1, 2, 3, 5, 0};
    memcpy(o->lift_map, m, sizeof m);
    o->lv_vel1 = 0;
    o->lv_temp = 3;
  }
  o->kappaM1 = d->kappa - 1.;
  o->sKappaM1 = 1. / o->kappaM1;
  o->cp = d->R * d->kappa / (d->kappa - 1.);
  o->cv = d->R / (d->kappa - 1.);
  o->KappasPr = d->kappa / d->Pr;
  const size_t nvol = (size_t)d->nElems * np * np * np, nsid = (size_t)d->nSides * np * np;
  /* own the descriptor arrays */
  o->d.Metrics_fTilde = (const double*)own_copy(o, d->Metrics_fTilde, nvol * 3 * 8);
  o->d.Metrics_gTilde = (const double*)own_copy(o, d->Metrics_gTilde, nvol * 3 * 8);
  o->d.Metrics_hTilde = (const double*)own_copy(o, d->Metrics_hTilde, nvol * 3 * 8);
  o->d.sJ = (const double*)own_copy(o, d->sJ, nvol * 8);
  o->d.Elem_xGP = (const double*)own_copy(o, d->Elem_xGP, nvol * 3 * 8);
  o->d.NormVec = (const double*)own_copy(o, d->NormVec, nsid * 3 * 8);
  o->d.TangVec1 = (const double*)own_copy(o, d->TangVec1, nsid * 3 * 8);
  o->d.TangVec2 = (const double*)own_copy(o, d->TangVec2, nsid * 3 * 8);
  o->d.SurfElem = (const double*)own_copy(o, d->SurfElem, nsid * 8);
  o->d.Face_xGP = (const double*)own_copy(o, d->Face_xGP, nsid * 3 * 8);
  o->d.Ja_Face = o->d.Ja_slave = NULL; /* the oracle keeps its own (freeable) copies, see oracle_set_face_metrics */
  if (d->Ja_Face && d->Ja_slave) {
    o->Ja_Face = (double*)xcalloc(nsid * 9, 8);
    o->Ja_slave = (double*)xcalloc(nsid * 9, 8);
    memcpy(o->Ja_Face, d->Ja_Face, nsid * 9 * 8);
    memcpy(o->Ja_slave, d->Ja_slave, nsid * 9 * 8);
  }
  o->d.SideToElem = (const int32_t*)own_copy(o, d->SideToElem, (size_t)d->nSides * 5 * 4);
  o->d.ElemToSide = (const int32_t*)own_copy(o, d->ElemToSide, (size_t)d->nElems * 12 * 4);
  o->d.S2V2 = (const int32_t*)own_copy(o, d->S2V2, (size_t)2 * np * np * 5 * 6 * 4);
  o->d.BC = (const int32_t*)own_copy(o, d->BC, (size_t)d->nBCSides * 4);
  o->d.BoundaryType = (const int32_t*)own_copy(o, d->BoundaryType, (size_t)d->nBCs * 3 * 4);
  o->d.RefStatePrim = (const double*)own_copy(o, d->RefStatePrim, (size_t)d->nRefState * 6 * 8);
  o->d.xGP = (const double*)own_copy(o, d->xGP, np * 8);
  o->d.wGP = (const double*)own_copy(o, d->wGP, np * 8);
  o->d.D = (const double*)own_copy(o, d->D, (size_t)np * np * 8);
  o->d.D_T = (const double*)own_copy(o, d->D_T, (size_t)np * np * 8);
  o->d.D_Hat_T = (const double*)own_copy(o, d->D_Hat_T, (size_t)np * np * 8);
  o->d.DVolSurf = (const double*)own_copy(o, d->DVolSurf, (size_t)np * np * 8);
  o->d.L_Minus = (const double*)own_copy(o, d->L_Minus, np * 8);
  o->d.L_Plus = (const double*)own_copy(o, d->L_Plus, np * 8);
  o->d.L_HatMinus = (const double*)own_copy(o, d->L_HatMinus, np * 8);
  o->d.L_HatPlus = (const double*)own_copy(o, d->L_HatPlus, np * 8);
  o->d.RKA = (const double*)own_copy(o, d->RKA, (size_t)(d->nRKStages > 1 ? d->nRKStages - 1 : 1) * 8);
  o->d.RKb = (const double*)own_copy(o, d->RKb, (size_t)(d->nRKStages > 0 ? d->nRKStages : 1) * 8);
  o->d.RKc = (const double*)own_copy(o, d->RKc, (size_t)(d->nRKStages > 1 ? d->nRKStages - 1 : 1) * 8);
  o->d.NbProc = (const int32_t*)own_copy(o, d->NbProc, (size_t)d->nNbProcs * 4);
  o->d.nMPISides_MINE_Proc = (const int32_t*)own_copy(o, d->nMPISides_MINE_Proc, (size_t)d->nNbProcs * 4);
  o->d.nMPISides_YOUR_Proc = (const int32_t*)own_copy(o, d->nMPISides_YOUR_Proc, (size_t)d->nNbProcs * 4);
  o->d.offsetMPISides_MINE = (const int32_t*)own_copy(o, d->offsetMPISides_MINE, (size_t)(d->nNbProcs + 1) * 4);
  o->d.offsetMPISides_YOUR = (const int32_t*)own_copy(o, d->offsetMPISides_YOUR, (size_t)(d->nNbProcs + 1) * 4);
  o->d.ncclUniqueId = NULL;
  /* side ranges, mesh.f90:257-269 */
  o->firstBCSide = 1;
  o->lastBCSide = d->nBCSides;
  o->firstInnerSide = d->nBCSides + 1;
  o->lastInnerSide = d->nBCSides + d->nInnerSides;
  o->firstMPISide_MINE = o->lastInnerSide + 1;
  o->lastMPISide_MINE = o->lastInnerSide + d->nMPISides_MINE;
  o->firstMPISide_YOUR = o->lastMPISide_MINE + 1;
  o->lastMPISide_YOUR = o->lastMPISide_MINE + d->nMPISides_YOUR;
  /* state, dg.f90:73-111, lifting.f90:146-166 */
  o->U = (double*)xcalloc(nvol * NVAR, 8);
  o->Ut = (double*)xcalloc(nvol * NVAR, 8);
  o->Ut_tmp = (double*)xcalloc(nvol * NVAR, 8);
  o->UPrim = (double*)xcalloc(nvol * NPRIM, 8);
  o->U_master = (double*)xcalloc(nsid * NVAR, 8);
  o->U_slave = (double*)xcalloc(nsid * NVAR, 8);
  o->UPrim_master = (double*)xcalloc(nsid * NPRIM, 8);
  o->UPrim_slave = (double*)xcalloc(nsid * NPRIM, 8);
  o->Flux_master = (double*)xcalloc(nsid * NVAR, 8);
  o->Flux_slave = (double*)xcalloc(nsid * NVAR, 8);
  if (d->parabolic) {
    o->gradUx = (double*)xcalloc(nvol * o->nvl, 8);
    o->gradUy = (double*)xcalloc(nvol * o->nvl, 8);
    o->gradUz = (double*)xcalloc(nvol * o->nvl, 8);
    o->gradUx_master = (double*)xcalloc(nsid * o->nvl, 8);
    o->gradUy_master = (double*)xcalloc(nsid * o->nvl, 8);
    o->gradUz_master = (double*)xcalloc(nsid * o->nvl, 8);
    o->gradUx_slave = (double*)xcalloc(nsid * o->nvl, 8);
    o->gradUy_slave = (double*)xcalloc(nsid * o->nvl, 8);
    o->gradUz_slave = (double*)xcalloc(nsid * o->nvl, 8);
    o->FluxX = (double*)xcalloc(nsid * o->nvl, 8);
    o->FluxY = (double*)xcalloc(nsid * o->nvl, 8);
    o->FluxZ = (double*)xcalloc(nsid * o->nvl, 8);
  }
  if (d->splitDG && d->nodeType == FLEXI_NODETYPE_GAUSS) { /* PP_EntropyVars = 1, dg.f90:80-87 */
    o->V = (double*)xcalloc(nvol * NVAR, 8);
    o->V_master = (double*)xcalloc(nsid * NVAR, 8);
    o->V_slave = (double*)xcalloc(nsid * NVAR, 8);
  }
  init_calc_timestep(o);
  return o;
}

/* Ja_Face / Ja_slave (3,3,0:N,0:N,nSides) of MOD_Mesh_Vars (metrics.f90:681-705): only the split form on Gauss nodes reads them; they
 * come with the descriptor (ABI 2, copied in oracle_create) or, like in the GPU library, through this extra entry. */
void oracle_set_face_metrics(oracle* o, const double* Ja_Face, const double* Ja_slave) {
  const size_t n = (size_t)o->d.nSides * o->np * o->np * 9;
  free(o->Ja_Face);
  free(o->Ja_slave);
  o->Ja_Face = (double*)xcalloc(n, 8);
  o->Ja_slave = (double*)xcalloc(n, 8);
  memcpy(o->Ja_Face, Ja_Face, n * 8);
  memcpy(o->Ja_slave, Ja_slave, n * 8);
}

/* PP_VISC and the MOD_EOS_Vars values InitEos leaves (eos.f90:160-192): Sutherland: Tref = 1/Tref_ini, Ts = Ts_ini/Tref_ini,
   cSuth = Ts**ExpoSuth*(1+Ts)/(2*Ts*Ts); power law: mu0 (descriptor) already divided by Tref**ExpoSuth */
int oracle_set_viscosity(oracle* o, int type, double Ts, double Tref, double ExpoSuth, double cSuth) {
  if (type < 0 || type > 2) return -1;
  o->visc = type;
  o->Ts = Ts;
  o->Tref = Tref;
  o->ExpoSuth = ExpoSuth;
  o->cSuth = cSuth;
  return 0;
}
/* InitSponge results (sponge.f90:97-400 stay on the host): nSpongeElems = 0 switches the sponge off */
int oracle_set_sponge(oracle* o, int nSpongeElems, const int* SpongeMap, const double* SpongeMat, const double* SpBaseFlow) {
  free(o->SpongeMap);
  free(o->SpongeMat);
  free(o->SpBaseFlow);
  o->SpongeMap = NULL;
  o->SpongeMat = o->SpBaseFlow = NULL;
  o->nSpongeElems = 0;
  if (nSpongeElems <= 0) return 0;
  if (!SpongeMap || !SpongeMat || !SpBaseFlow || nSpongeElems > o->d.nElems) return -1;
  const size_t np3 = (size_t)o->np * o->np * o->np;
  o->SpongeMap = (int*)xcalloc((size_t)nSpongeElems, sizeof(int));
  o->SpongeMat = (double*)xcalloc((size_t)nSpongeElems * np3, 8);
  o->SpBaseFlow = (double*)xcalloc((size_t)o->d.nElems * np3 * NVAR, 8);
  memcpy(o->SpongeMap, SpongeMap, (size_t)nSpongeElems * sizeof(int));
  memcpy(o->SpongeMat, SpongeMat, (size_t)nSpongeElems * np3 * 8);
  memcpy(o->SpBaseFlow, SpBaseFlow, (size_t)o->d.nElems * np3 * NVAR * 8);
  o->nSpongeElems = nSpongeElems;
  return 0;
}
/* OverintegrationType (dg/overintegration.f90:40-47): 0 none, 1 cutoff (OverintegrationMat = Vdm_Leg diag(1..1,0..0) sVdm_Leg,
   :104-116).  2 (conservative cutoff) is not restated. */
int oracle_set_overintegration(oracle* o, int type, const double* Mat) {
  free(o->OverintegrationMat);
  o->OverintegrationMat = NULL;
  if (type == 0) return 0;
  if (type != 1 || !Mat) return -1;
  const size_t n = (size_t)o->np * o->np;
  o->OverintegrationMat = (double*)xcalloc(n, 8);
  memcpy(o->OverintegrationMat, Mat, n * 8);
  return 0;
}
/* OverintegrationType 2 (conservative cut-off): the arrays InitOverintegration builds (overintegration.f90:118-135), Fortran order;
   NUnder < 0 switches it off.  The host also filters the initial solution (timedisc.f90:119-122): oracle_filter_conservative. */
int oracle_set_overintegration_cons(oracle* o, int NUnder, const double* Vdm_N_NUnder, const double* Vdm_NUnder_N, const double* sJNUnder) {
  free(o->Vdm_N_NUnder);
  free(o->Vdm_NUnder_N);
  free(o->sJNUnder);
  o->Vdm_N_NUnder = o->Vdm_NUnder_N = o->sJNUnder = NULL;
  o->NUnder = -1;
  if (NUnder < 0) return 0;
  if (NUnder >= o->np - 1 || !Vdm_N_NUnder || !Vdm_NUnder_N || !sJNUnder) return -1;
  const size_t nu = (size_t)NUnder + 1, n1 = nu * o->np, n3 = nu * nu * nu * o->d.nElems;
  o->Vdm_N_NUnder = (double*)xcalloc(n1, 8);
  o->Vdm_NUnder_N = (double*)xcalloc(n1, 8);
  o->sJNUnder = (double*)xcalloc(n3, 8);
  memcpy(o->Vdm_N_NUnder, Vdm_N_NUnder, n1 * 8);
  memcpy(o->Vdm_NUnder_N, Vdm_NUnder_N, n1 * 8);
  memcpy(o->sJNUnder, sJNUnder, n3 * 8);
  o->NUnder = NUnder;
  return 0;
}
void oracle_filter_conservative(oracle* o, double* JU) { filter_conservative(o, JU); }
/* FilterType 3 (LAF): call oracle_set_filter(o, 1, cut-off FilterMat of NFilter) first, then this with IntegrationWeight
   (0:N,0:N,0:N,nElems), normMod and LAF_alpha (filter.f90:149-186); the per-element state starts as InitFilter leaves it.
   IntegrationWeight == NULL switches LAF off again. */
int oracle_set_filter_laf(oracle* o, const double* IntegrationWeight, double normMod, double LAF_alpha) {
  free(o->laf_lim); free(o->laf_eRatio); free(o->laf_r); free(o->laf_ekin_old); free(o->laf_fluc_old); free(o->laf_iw);
  o->laf_lim = o->laf_eRatio = o->laf_r = o->laf_ekin_old = o->laf_fluc_old = o->laf_iw = NULL;
  o->laf = 0;
  if (!IntegrationWeight) return 0;
  const size_t nE = (size_t)o->d.nElems, n3 = (size_t)o->np * o->np * o->np;
  o->laf_lim = (double*)xcalloc(nE, 8);
  o->laf_eRatio = (double*)xcalloc(nE, 8);
  o->laf_r = (double*)xcalloc(nE, 8);
  o->laf_ekin_old = (double*)xcalloc(nE, 8);
  o->laf_fluc_old = (double*)xcalloc(nE, 8);
  o->laf_iw = (double*)xcalloc(nE * n3, 8);
  memcpy(o->laf_iw, IntegrationWeight, nE * n3 * 8);
  for (size_t e = 0; e < nE; ++e) {
    o->laf_lim[e] = 1E-8;
    o->laf_ekin_old[e] = 1.E-16;
    o->laf_fluc_old[e] = 1.E-16;
  }
  o->laf_normMod = normMod;
  o->laf_alpha = LAF_alpha;
  o->laf = 1;
  return 0;
}
/* ElemData of the LAF filter: eRatio, lim, r (nElems each) */
void oracle_get_filter_laf(const oracle* o, double* eRatio, double* lim, double* r) {
  const size_t nE = (size_t)o->d.nElems;
  memcpy(eRatio, o->laf_eRatio, nE * 8);
  memcpy(lim, o->laf_lim, nE * 8);
  memcpy(r, o->laf_r, nE * 8);
}
/* FilterType (filter/filter.f90:67-72,135-196): 0 none; 1 cutoff and 2 modal (Hesthaven) both point Filter_Pointer at Filter with
   the host-built FilterMat (:138-190); 3 (LAF) is not restated. */
int oracle_set_filter(oracle* o, int type, const double* Mat) {
  free(o->FilterMat);
  o->FilterMat = NULL;
  if (type == 0) return 0;
  if ((type != 1 && type != 2) || !Mat) return -1;
  const size_t n = (size_t)o->np * o->np;
  o->FilterMat = (double*)xcalloc(n, 8);
  memcpy(o->FilterMat, Mat, n * 8);
  return 0;
}
/* Filter(U_in, FilterMat) on any (5,0:N,0:N,0:N,nElems) array, e.g. the initial Overintegration(U) of timedisc.f90:117-122 */
void oracle_filter(oracle* o, const double* Mat, double* U) { filter_volume(o, Mat, U); }

void oracle_destroy(oracle* o) {
  if (!o) return;
  free(o->OverintegrationMat);
  free(o->FilterMat);
  free(o->SpongeMap);
  free(o->SpongeMat);
  free(o->SpBaseFlow);
  free(o->Vdm_N_NUnder);
  free(o->Vdm_NUnder_N);
  free(o->sJNUnder);
  free(o->laf_lim); free(o->laf_eRatio); free(o->laf_r); free(o->laf_ekin_old); free(o->laf_fluc_old); free(o->laf_iw);
  for (int i = 0; i < o->n_owned; ++i) free(o->owned[i]);
  double* st[] = {o->U, o->Ut, o->Ut_tmp, o->UPrim, o->U_master, o->U_slave, o->UPrim_master, o->UPrim_slave, o->Flux_master,
                  o->Flux_slave, o->gradUx, o->gradUy, o->gradUz, o->gradUx_master, o->gradUy_master, o->gradUz_master,
                  o->gradUx_slave, o->gradUy_slave, o->gradUz_slave, o->FluxX, o->FluxY, o->FluxZ, o->MetricsAdv, o->MetricsVisc,
                  o->V, o->V_master, o->V_slave, o->Ja_Face, o->Ja_slave};
  for (size_t i = 0; i < sizeof st / sizeof st[0]; ++i) free(st[i]);
  free(o);
}

oracle_world* oracle_world_create(oracle** ranks, int n) {
  oracle_world* w = (oracle_world*)xcalloc(1, sizeof(oracle_world));
  w->n = n;
  w->r = (oracle**)xcalloc(n, sizeof(oracle*));
  for (int i = 0; i < n; ++i) w->r[i] = ranks[i];
  return w;
}
void oracle_world_destroy(oracle_world* w) {
  if (!w) return;
  free(w->r);
  free(w);
}

/* pointer to a named module array of one rank (same numbering as enum flexi_gpu_array, plus oracle-only extras >= 100) */
double* oracle_array(oracle* o, int which, int64_t* n_out) {
  const int np = o->np;
  const size_t nvol = (size_t)o->d.nElems * np * np * np, nsid = (size_t)o->d.nSides * np * np;
  double* p = NULL;
  size_t n = 0;
  switch (which) {
    case FLEXI_ARR_U: p = o->U; n = nvol * NVAR; break;
    case FLEXI_ARR_UT: p = o->Ut; n = nvol * NVAR; break;
    case FLEXI_ARR_U_MASTER: p = o->U_master; n = nsid * NVAR; break;
    case FLEXI_ARR_U_SLAVE: p = o->U_slave; n = nsid * NVAR; break;
    case FLEXI_ARR_FLUX: p = o->Flux_master; n = nsid * NVAR; break;
    case FLEXI_ARR_GRADUX: p = o->gradUx; n = nvol * o->nvl; break;
    case FLEXI_ARR_GRADUY: p = o->gradUy; n = nvol * o->nvl; break;
    case FLEXI_ARR_GRADUZ: p = o->gradUz; n = nvol * o->nvl; break;
    case FLEXI_ARR_GRADUX_MASTER: p = o->gradUx_master; n = nsid * o->nvl; break;
    case FLEXI_ARR_GRADUY_MASTER: p = o->gradUy_master; n = nsid * o->nvl; break;
    case FLEXI_ARR_GRADUZ_MASTER: p = o->gradUz_master; n = nsid * o->nvl; break;
    case FLEXI_ARR_GRADUX_SLAVE: p = o->gradUx_slave; n = nsid * o->nvl; break;
    case FLEXI_ARR_GRADUY_SLAVE: p = o->gradUy_slave; n = nsid * o->nvl; break;
    case FLEXI_ARR_GRADUZ_SLAVE: p = o->gradUz_slave; n = nsid * o->nvl; break;
    case 100: p = o->Flux_slave; n = nsid * NVAR; break;
    case 101: p = o->UPrim; n = nvol * NPRIM; break;
    case 102: p = o->Ut_tmp; n = nvol * NVAR; break;
    default: break;
  }
  if (n_out) *n_out = (int64_t)n;
  return p;
}

/* operator-level entry points used to replay the reference's unit tests (unitTests/ProlongToFace.f90, SurfInt.f90) */
void oracle_prolong_to_face(oracle* o, int nv, const double* Uvol, double* Uface_master, double* Uface_slave, int doMPISides) {
  prolong_to_face(o, nv, Uvol, Uface_master, Uface_slave, doMPISides);
}
void oracle_surf_int(oracle* o, int nv, const double* Flux_master, const double* Flux_slave, double* Ut, int doMPISides,
                     const double* L_HatMinus, const double* L_HatPlus) {
  surf_int(o, nv, Flux_master, Flux_slave, Ut, doMPISides, L_HatMinus, L_HatPlus, -1.);
}

/* point-level entry points for the identity tests (consistency / symmetry of fluxes) */
void oracle_riemann_point(oracle* o, int solver, double* FOut, const double* U_L, const double* U_R, const double* nv, const double* t1,
                          const double* t2) {
  double PL[6], PR[6];
  cons_to_prim(o, PL, U_L);
  cons_to_prim(o, PR, U_R);
  riemann_point(o, solver, FOut, U_L, U_R, PL, PR, nv, t1, t2);
}
void oracle_split_volume_flux(oracle* o, int type, const double* URef, const double* U, const double* MRef, const double* M, double* Flux) {
  double PR[6], P[6];
  cons_to_prim(o, PR, URef);
  cons_to_prim(o, P, U);
  split_volume_flux(o, type, URef, PR, U, P, MRef, M, Flux);
}
void oracle_eval_flux3d(oracle* o, const double* U, double* f, double* g, double* h) {
  double P[6];
  cons_to_prim(o, P, U);
  eval_flux3d(U, P, f, g, h);
}
/* gx,gy,gz in the oracle's lifting-variable layout (nVarLifting entries); f,g,h full 5-vectors (row DENS = 0) */
void oracle_eval_diff_flux3d(oracle* o, const double* UPrim, const double* gx, const double* gy, const double* gz, double* f, double* g,
                             double* h) {
  eval_diff_flux3d(o, UPrim, gx, gy, gz, f, g, h);
}
void oracle_prim_to_entropy(oracle* o, const double* prim, double* entropy) { prim_to_entropy(o, entropy, prim); }
void oracle_entropy_to_cons(oracle* o, const double* entropy, double* cons) { entropy_to_cons(o, entropy, cons); }
int oracle_lv_vel1(oracle* o) { return o->lv_vel1; }
int oracle_lv_temp(oracle* o) { return o->lv_temp; }
int oracle_desc_bytes(void) { return (int)sizeof(flexi_gpu_desc); }
