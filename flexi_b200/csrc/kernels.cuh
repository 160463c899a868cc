// kernels.cuh -- the CUDA kernels of the DGSEM right-hand side, templated on NP = N+1.
//
// Kernel classes (one launch each per RK stage):
//   k_volume  element kernel  : ConsToPrim + BR1/BR2 lifting (volume + gathered surface part) + ApplyJacobianLifting +
//                               ProlongToFace of the gradients + VolInt (weak or split form).   dg.f90:334,523,546
//   k_flux    side kernel     : GetPrimitiveStateSurface + Riemann + ViscousFlux + *SurfElem.   fillflux.f90:40-215
//   k_update  element kernel  : SurfInt (gathered) + sign + ApplyJacobian + LSERK stage update + ProlongToFace of the new
//                               state.                                           dg.f90:636-675, timestep.f90:126-131
//   k_prolong element kernel  : ProlongToFace of U alone (after set_state).                      prolongtoface.t90:21-144
//   k_dt      element kernel  : CalcTimeStep.                                                   calctimestep.f90:89-227
#pragma once
#include "layout.cuh"

namespace fx {

enum KernelClass { KC_VOLUME = 0, KC_FLUX = 1, KC_UPDATE = 2, KC_PROLONG = 3, KC_DT = 4, KC_LAYOUT = 5, KC_COUNT = 6 };

struct VolArgs {
  Mesh m;
  Eqn q;
  const int32_t* elems;  // element list (NULL = identity)
  int nList;
  const double* U;         // [e][5][NP3]
  const double* U_master;  // [s][5][NP2]
  const double* U_slave;
  double* Ut;  // raw volume integral (before sign / Jacobian)
  double* gm[3];  // gradU{x,y,z}_master [s][nvl][NP2]
  double* gs[3];  // gradU{x,y,z}_slave
  double* gvol[3];  // optional volume gradients [e][nvl][NP3] (NULL = not stored)
  const double* basisHost;  // HOST copy of Mesh::basis (k_volume_gl2 takes the D matrices as kernel parameters)
  // viscous-trace protocol (k_volume_gl2 only): n.F_v(U,gradU) of the element's own side, [s][4][NP2]; gradient traces are then
  // written for boundary-condition sides only, unless fullGrad asks for all of them
  double* fvm;
  double* fvs;
  int fullGrad;
  int dbg;            // development build only (FLEXI_GPU_DEV): phase-skipping bits of the k_volume_gl2 timing experiments; 0 in the release library
  int prefetchAhead;  // k_volume_gl2: CTA b prefetches the element of CTA b+prefetchAhead into L2 (0 = off)
};

struct FluxArgs {
  Mesh m;
  Eqn q;
  int firstSide, nSidesRange;  // 0-based range
  int parabolic;
  const double* U_master;
  const double* U_slave;
  const double* gm[3];
  const double* gs[3];
  double* Flux;  // [s][5][NP2]  (Flux_master == Flux_slave, fillflux.f90:201)
  const double* fvm;  // viscous-trace protocol (vt != 0): n.F_v of the master / slave element's side, [s][4][NP2]
  const double* fvs;
  int vt;
};

struct UpdArgs {
  Mesh m;
  const int32_t* elems;
  int nList;
  const double* UtVol;  // raw volume integral
  const double* Flux;
  double* U;
  double* UtTmp;  // LSERK register (Ut_tmp, timedisc_vars)
  double* UtOut;  // Ut when mode == 0
  double* U_master;
  double* U_slave;
  int mode;    // 0: write Ut only (DGTimeDerivative_weakForm); 1: first RK stage; 2: later RK stage
  double rkA;  // RKA(iStage)
  double b_dt; // RKb(iStage)*dt
  // CalcSource, exactfunc.f90:727-815 (case 4), evaluated at the stage time t
  int entropy;  // traces of the new state through entropy variables (ProlongToFaceEntropy); needs q
  Eqn q;
  int doSource;
  const double* xGP;  // Elem_xGP as [e][3][NP3]
  double t, srcOmega, srcA, srcTmp[6];
  const double* filt;  // OverintegrationMat as [out][in], NP*NP doubles; nullptr = OverintegrationType 0 (k_update<.,.,.,true> reads it)
  // OVERINTEGRATIONTYPE_CONSCUTOFF (overintegration.f90:181-330): filt = Vdm_N_NUnder and filt2 = Vdm_NUnder_N, both padded with
  // zeros to NP x NP [out][in]; sJNU = sJNUnder padded to [e][NP3] (zero outside the NUnder cube); nullptr = plain cut-off
  const double* filt2;
  const double* sJNU;
  // Sponge(Ut), sponge.f90:612-678: spIdx[e] = index of element e in the sponge list or -1 (nullptr = doSponge off),
  // spMat [iSponge][NP3] (SpongeMat, pre-multiplied by the Jacobian), spBase [e][5][NP3] (SpBaseFlow)
  const int32_t* spIdx;
  const double* spMat;
  const double* spBase;
};

struct FiltArgs {
  double* U;          // [e][5][NP3], filtered in place
  const double* mat;  // FilterMat as [out][in]
  // FilterType 3 (LAF, filter.f90:295-448): iw = IntegrationWeight [e][NP3] (nullptr = plain filter); per-element state
  const double* iw;
  double *lim, *eRatio, *r, *ekinOld, *flucOld;
  double normMod, alpha;
};

struct ProArgs {
  Mesh m;
  const double* V;  // [e][nv][NP3]
  double* master;   // [s][nv][NP2]
  double* slave;
  int nv;
  int firstSide, lastSide;  // inclusive 0-based side range that is written
  int entropy;  // ProlongToFaceEntropy (prolongtoface.t90:151-204): V holds U (nv = 5), entropy variables are prolonged
  Eqn q;
};

// k_filter: Filter(U, FilterMat) in place (filter.f90:264-288 -> ChangeBasis3D), the solution filter the reference applies at the
// start of every DG operator evaluation when FilterType > 0 (dg.f90:329-330).  One CTA per element, one thread per node, three
// 1-D sweeps through 5 shared planes; sums start from zero and run over ascending input index like changeBasis.t90.  A plain
// extra pass over U (HBM-bound: 80 B/DOF); the fused form is described in DESIGN.md section 7.
template <int NP>
__global__ void __launch_bounds__(Dim<NP>::THREADS) k_filter(FiltArgs A) {
  using D = Dim<NP>;
  extern __shared__ double sm_raw[];
  auto sm = elem_smem<NP>(sm_raw);
  const int n = elem_tid<NP>();
  const size_t e = elem_blk<NP>();
  const int i = n % NP, j = (n / NP) % NP, k = n / D::NP2;
  double w[5];
#pragma unroll
  for (int v = 0; v < 5; ++v) w[v] = A.U[(e * 5 + v) * D::NP3 + n];
#pragma unroll
  for (int dir = 0; dir < 3; ++dir) {
    const int me = (dir == 0) ? i : (dir == 1) ? j : k;
    const int stride = (dir == 0) ? 1 : (dir == 1) ? NP : D::NP2;
    const int base = n - me * stride;
#pragma unroll
    for (int v = 0; v < 5; ++v) sm[v * D::NP3 + n] = w[v];
    elem_sync<NP>();
    const double* row = A.mat + me * NP;
#pragma unroll
    for (int v = 0; v < 5; ++v) {
      double acc = 0.;
#pragma unroll
      for (int l = 0; l < NP; ++l) acc = acc + __ldg(row + l) * sm[v * D::NP3 + base + l * stride];
      w[v] = acc;
    }
    elem_sync<NP>();
  }
  if (A.iw) {
    // ---- Filter_LAF: the filtered state replaces U only where the energy of the filtered-out velocity scales, relative to the
    //      resolved fluctuations, exceeds the element's adaptive limit.  Element sums: tree reduction through the (now free)
    //      planes, in a fixed order. ----
    double u[5];
#pragma unroll
    for (int v = 0; v < 5; ++v) u[v] = A.U[(e * 5 + v) * D::NP3 + n];
    const double wgt = A.iw[e * D::NP3 + n], sr = 1. / u[0];
    auto reduce = [&](int nq) {  // sums of planes 0..nq-1 over the element end up in sm[q * NP3]
      elem_sync<NP>();
      int p2 = 1;
      while (p2 < D::NP3) p2 <<= 1;
      for (int off = p2 >> 1; off > 0; off >>= 1) {
        if (n < off && n + off < D::NP3)
          for (int qq = 0; qq < nq; ++qq) sm[qq * D::NP3 + n] = sm[qq * D::NP3 + n] + sm[qq * D::NP3 + n + off];
        elem_sync<NP>();
      }
    };
#pragma unroll
    for (int c = 0; c < 3; ++c) sm[c * D::NP3 + n] = w[1 + c] * sr * wgt;  // U_Avg(2:4)
    reduce(3);
    const double a1 = sm[0], a2 = sm[D::NP3], a3 = sm[2 * D::NP3];
    elem_sync<NP>();
    const double f1 = u[1] - w[1], f2 = u[2] - w[2], f3 = u[3] - w[3];
    const double sr2 = sr * sr;
    sm[n] = (f1 * f1 * sr2 + f3 * f3 * sr2 + f2 * f2 * sr2) * wgt;                                            // ekin_fluc_avg
    sm[D::NP3 + n] = ((w[1] - a1) * (w[1] - a1) * sr2 + (w[2] - a2) * (w[2] - a2) * sr2 + (w[3] - a3) * (w[3] - a3) * sr2) * wgt;  // ekin_avg
    reduce(2);
    if (n == 0) {
      const double fluc = sm[0], ekin = sm[D::NP3];
      const double ratio = (fluc / ekin) * A.normMod;
      double lim = fmax(A.lim[e], 1.0E-8);
      const double dfluc = fluc - A.flucOld[e];
      const double dedt = fabs(ekin - A.ekinOld[e]) / dfluc;
      double doFilter = 0.;
      if (dedt > 1.0 && dfluc > 0.) {
        lim = ratio;
      } else if (dfluc < 0.) {
        lim = ratio;
      } else if (ratio > lim * A.alpha) {
        doFilter = 1.;
        A.r[e] = A.r[e] + 1.;
      }
      A.eRatio[e] = ratio;
      A.lim[e] = lim;
      A.ekinOld[e] = ekin;
      A.flucOld[e] = fluc;
      sm[2 * D::NP3] = doFilter;
    }
    elem_sync<NP>();
    const bool keep = sm[2 * D::NP3] == 0.;
    elem_exit<NP>();
    if (keep) return;
  }
#pragma unroll
  for (int v = 0; v < 5; ++v) A.U[(e * 5 + v) * D::NP3 + n] = w[v];
}

// ---------------------------------------------------------------------------------------------------------------------------
template <int NP, class SP>
__device__ __forceinline__ void load_basis(SP sB, const double* g, int tid, int nthreads) {
  for (int t = tid; t < Dim<NP>::NBASIS; t += nthreads) sB[t] = g[t];
}

// ---------------------------------------------------------------------------------------------------------------------------
// k_prolong: ProlongToFace for nv variables, element gather.
template <int NP, bool GL>
__global__ void __launch_bounds__(Dim<NP>::THREADS) k_prolong(ProArgs A) {
  using D = Dim<NP>;
  extern __shared__ double sm_raw[];
  auto sm = elem_smem<NP>(sm_raw);
  auto sB = sm;
  auto sV = sm + D::NBASIS;  // nv planes
  const int tid = elem_tid<NP>();
  const int e = elem_blk<NP>();
  load_basis<NP>(sB, A.m.basis, tid, D::NP3);
  for (int v = 0; v < A.nv; ++v) sV[v * D::NP3 + tid] = A.V[((size_t)e * A.nv + v) * D::NP3 + tid];
  if (A.entropy) {  // own node: cons -> prim -> entropy variables, in place
    double Uc[5], P[6], w[5];
    for (int v = 0; v < 5; ++v) Uc[v] = sV[v * D::NP3 + tid];
    cons_to_prim(A.q, Uc, P);
    prim_to_entropy(A.q, P, w);
    for (int v = 0; v < 5; ++v) sV[v * D::NP3 + tid] = w[v];
  }
  elem_sync<NP>();
  for (int t = tid; t < 6 * D::NP2; t += D::NP3) {
    const int ls = t / D::NP2 + 1, pq = t % D::NP2;
    const SideRef sr = A.m.e2s[e * 6 + ls - 1];
    if (sr.side < A.firstSide || sr.side > A.lastSide) continue;
    const int ab = A.m.s2v2[(sr.flip * 6 + ls - 1) * D::NP2 + pq];
    const int a = ab % NP, b = ab / NP;
    double* dst = (sr.flip == 0) ? A.master : A.slave;
    if (A.entropy) {
      double w[5], Uc[5];
      for (int v = 0; v < 5; ++v) w[v] = face_eval<NP, GL>(sV + v * D::NP3, sB, ls, a, b);
      entropy_to_cons(A.q, w, Uc);
      for (int v = 0; v < 5; ++v) dst[((size_t)sr.side * 5 + v) * D::NP2 + pq] = Uc[v];
    } else {
      for (int v = 0; v < A.nv; ++v)
        dst[((size_t)sr.side * A.nv + v) * D::NP2 + pq] = face_eval<NP, GL>(sV + v * D::NP3, sB, ls, a, b);
    }
  }
  elem_exit<NP>();
}

// ---------------------------------------------------------------------------------------------------------------------------
// k_volume
template <int NP, bool GL, bool SPLIT, bool PARA, int LIFT = 1>
struct VolSmem {
  using D = Dim<NP>;
  static constexpr int WORK_PLANES = (PARA || !SPLIT) ? 15 : 9;
  static constexpr int FACE_DOUBLES = PARA ? 6 * 8 * D::NP2 : 0;
  // BR1: the staged face fluxes alias the work planes; BR2 needs them next to the volume-only gradients
  static constexpr int WORK = (LIFT == 2) ? WORK_PLANES * D::NP3 : ((WORK_PLANES * D::NP3 > FACE_DOUBLES) ? WORK_PLANES * D::NP3 : FACE_DOUBLES);
  static constexpr int O_FACE2 = D::NBASIS + 7 * D::NP3 + WORK;
  static constexpr int DOUBLES = O_FACE2 + ((LIFT == 2 && PARA) ? FACE_DOUBLES : 0);
  static constexpr size_t BYTES = sizeof(double) * DOUBLES;
};

// pairwise two-point volume fluxes of one node along its three lines, volint.f90:211-371 (advective part)
template <int NP, int STYPE, class SP>
__device__ __forceinline__ void split_lines(const SP sP, const SP sM, const SP sB, int i, int j, int k, int n,
                                            const double* M, double* ut, const double kappa) {
  using D = Dim<NP>;
  const auto dv = sB + D::O_DV;
  Node a;
  a.rho = sP[0 * D::NP3 + n];
  a.v1 = sP[1 * D::NP3 + n];
  a.v2 = sP[2 * D::NP3 + n];
  a.v3 = sP[3 * D::NP3 + n];
  a.p = sP[4 * D::NP3 + n];
  a.E = sP[6 * D::NP3 + n];
#pragma unroll
  for (int dir = 0; dir < 3; ++dir) {
    const int me = (dir == 0) ? i : (dir == 1) ? j : k;
    const int stride = (dir == 0) ? 1 : (dir == 1) ? NP : D::NP2;
    const int base = n - me * stride;
#pragma unroll
    for (int l = 0; l < NP; ++l) {
      const int nl = base + l * stride;
      Node b;
      b.rho = sP[0 * D::NP3 + nl];
      b.v1 = sP[1 * D::NP3 + nl];
      b.v2 = sP[2 * D::NP3 + nl];
      b.v3 = sP[3 * D::NP3 + nl];
      b.p = sP[4 * D::NP3 + nl];
      b.E = sP[6 * D::NP3 + nl];
      const double Mb[3] = {sM[(3 * dir + 0) * D::NP3 + nl], sM[(3 * dir + 1) * D::NP3 + nl], sM[(3 * dir + 2) * D::NP3 + nl]};
      double F[5];
      split_volume_flux<STYPE>(a, b, M + 3 * dir, Mb, F, kappa);
      const double d = (l == me) ? 0. : dv[l * NP + me];  // DVolSurf(l,me); self pair weighted with zero (no divergent skip)
#pragma unroll
      for (int v = 0; v < 5; ++v) ut[v] = ut[v] + d * F[v];
    }
  }
}

// prim index of lifting variable v (PRIM_LIFT, eos.h:118,127): nvl=5 -> (rho,v1,v2,v3,T), nvl=4 -> (v1,v2,v3,T)
__host__ __device__ constexpr int lift_prim(int nvl, int v) { return nvl == 5 ? (v < 4 ? v : 5) : (v < 3 ? v + 1 : 5); }

// Resident CTAs per SM the register allocation of the Navier-Stokes instantiations is sized for (FX_KV_MINB, default 1; timing
// experiments set it through `make dev DEVFLAGS=-DFX_KV_MINB=2`): N = 7 runs 512 threads per element at ~122 registers, i.e. one
// element per SM; 2 forces 64 registers (spills) for two resident elements.
#ifndef FX_KV_MINB
#define FX_KV_MINB 1
#endif
template <int NP, bool PARA>
constexpr int kv_min_blocks() { return (PARA && !Dim<NP>::PAIR && Dim<NP>::NP3 * FX_KV_MINB <= 1024) ? FX_KV_MINB : 1; }

template <int NP, bool GL, bool SPLIT, bool PARA, int NVL, int LIFT>
__global__ void __launch_bounds__(Dim<NP>::THREADS, kv_min_blocks<NP, PARA>()) k_volume(VolArgs A) {
  using D = Dim<NP>;
  extern __shared__ double sm_raw[];
  auto sm = elem_smem<NP>(sm_raw);
  auto sB = sm;
  auto sP = sm + D::NBASIS;  // rho,v1,v2,v3,p,T,rhoE
  auto sW = sP + 7 * D::NP3;
  auto sLF = (LIFT == 2) ? sm + VolSmem<NP, GL, SPLIT, PARA, LIFT>::O_FACE2 : sW;  // staged lifting fluxes of the faces
  const int tid = elem_tid<NP>();
  const int e = A.elems ? A.elems[elem_blk<NP>()] : elem_blk<NP>();
  const int i = tid % NP, j = (tid / NP) % NP, k = tid / D::NP2;
  const int n = tid;
  const Eqn& q = A.q;

  load_basis<NP>(sB, A.m.basis, tid, D::NP3);
  double U[5], P[6], M[9];
#pragma unroll
  for (int v = 0; v < 5; ++v) U[v] = A.U[((size_t)e * 5 + v) * D::NP3 + n];
  cons_to_prim(q, U, P);
#pragma unroll
  for (int v = 0; v < 6; ++v) sP[v * D::NP3 + n] = P[v];
  sP[6 * D::NP3 + n] = U[ENER];
#pragma unroll
  for (int c = 0; c < 9; ++c) M[c] = A.m.Met[((size_t)e * 9 + c) * D::NP3 + n];
  const double sJ = A.m.sJ[(size_t)e * D::NP3 + n];
  elem_sync<NP>();

  double ut[5] = {0., 0., 0., 0., 0.};
  const auto dh = sB + D::O_DH;

  if (PARA) {
    constexpr int nvl = NVL;
    constexpr int LV = (NVL == 5) ? 1 : 0;  // LIFT_VEL1
    // ---- 1. lifting surface flux of the six sides, lifting_fillflux.t90:81-107 (strong form: 0.5*(U_slave-U_master)*SurfElem)
    for (int t = tid; t < 6 * D::NP2; t += D::NP3) {
      const int ls = t / D::NP2 + 1, pq = t % D::NP2;
      const SideRef sr = A.m.e2s[e * 6 + ls - 1];
      const int ab = A.m.s2v2[(sr.flip * 6 + ls - 1) * D::NP2 + pq];
      const double* nbTrace = (sr.flip == 0) ? A.U_slave : A.U_master;
      const double* myTrace = (sr.flip == 0) ? A.U_master : A.U_slave;
      double Un[5], Um[5], Pn[6], Pm[6];
#pragma unroll
      for (int v = 0; v < 5; ++v) Un[v] = nbTrace[((size_t)sr.side * 5 + v) * D::NP2 + pq];
#pragma unroll
      for (int v = 0; v < 5; ++v) Um[v] = myTrace[((size_t)sr.side * 5 + v) * D::NP2 + pq];
      cons_to_prim(q, Un, Pn);
      cons_to_prim(q, Um, Pm);
      const double* g = A.m.Geo + (size_t)sr.side * 10 * D::NP2 + pq;
      const double surf = g[9 * D::NP2];
      auto o = sLF + (ls - 1) * 8 * D::NP2 + ab;
      const double sgn = (sr.flip == 0) ? 1. : -1.;  // (slave - master) seen from this element
#pragma unroll
      for (int v = 0; v < nvl; ++v) {
            const int pv = lift_prim(nvl, v);
        o[v * D::NP2] = 0.5 * (sgn * (Pn[pv] - Pm[pv])) * surf;
      }
      o[5 * D::NP2] = g[0];
      o[6 * D::NP2] = g[D::NP2];
      o[7 * D::NP2] = g[2 * D::NP2];
    }
    // ---- 2. volume part, lifting_volint.t90:26-118 (non-conservative)
    double gx[5], gy[5], gz[5];
    {
      const auto dt = sB + D::O_DT;
#pragma unroll
      for (int v = 0; v < nvl; ++v) {
        const auto pl = sP + lift_prim(nvl, v) * D::NP3;
        const int bx = n - i, by = n - j * NP, bz = n - k * D::NP2;
        double gxi = dt[i] * pl[bx], geta = dt[j] * pl[by], gzeta = dt[k] * pl[bz];
#pragma unroll
        for (int l = 1; l < NP; ++l) {
          gxi = gxi + dt[l * NP + i] * pl[bx + l];
          geta = geta + dt[l * NP + j] * pl[by + l * NP];
          gzeta = gzeta + dt[l * NP + k] * pl[bz + l * D::NP2];
        }
        gx[v] = M[0] * gxi + M[3] * geta + M[6] * gzeta;
        gy[v] = M[1] * gxi + M[4] * geta + M[7] * gzeta;
        gz[v] = M[2] * gxi + M[5] * geta + M[8] * gzeta;
      }
    }
    elem_sync<NP>();
    if (LIFT == 2) {
      // BR2 (lifting_br2.t90:111-131): Jacobian first, faces are prolonged from the volume-only gradients
#pragma unroll
      for (int v = 0; v < nvl; ++v) {
        gx[v] = gx[v] * sJ;
        gy[v] = gy[v] * sJ;
        gz[v] = gz[v] * sJ;
        sW[(0 * nvl + v) * D::NP3 + n] = gx[v];
        sW[(1 * nvl + v) * D::NP3 + n] = gy[v];
        sW[(2 * nvl + v) * D::NP3 + n] = gz[v];
      }
      elem_sync<NP>();
      // face gradients: prolonged value + eta * sum_l L_Minus(l) * sJ_l * Flux * L_HatMinus(l)   (lifting_br2.t90:253-307)
      for (int t = tid; t < 6 * D::NP2; t += D::NP3) {
        const int ls = t / D::NP2 + 1, pq = t % D::NP2;
        const SideRef sr = A.m.e2s[e * 6 + ls - 1];
        const int ab = A.m.s2v2[(sr.flip * 6 + ls - 1) * D::NP2 + pq];
        const int a = ab % NP, b = ab / NP;
        int base, stride;
        face_line<NP>(ls, a, b, base, stride);
        const bool minus = is_minus_side(ls);
        double eta = q.etaBR2;
        if (sr.side < A.m.nBCSides) {
          const int bt = A.m.bcType[sr.side];
          if (bt == 3 || bt == 4) eta = q.etaBR2_wall;
        }
        double wsum = 0.;  // sum_l L_Minus(l) * sJ_l * L_HatMinus(l)
        {
          const auto lm = sB + D::O_LM;
          const auto lhm = sB + D::O_LHM;
          const int lmax = GL ? 0 : NP - 1;
          for (int l = 0; l <= lmax; ++l) {
            const int d = minus ? l : NP - 1 - l;
            wsum = wsum + lm[l] * (A.m.sJ[(size_t)e * D::NP3 + base + d * stride] * lhm[l]);
          }
        }
        const auto o = sLF + (ls - 1) * 8 * D::NP2 + ab;
        for (int d = 0; d < 3; ++d) {
          double* dst = (sr.flip == 0) ? A.gm[d] : A.gs[d];
          const double nd = o[(5 + d) * D::NP2];
          for (int v = 0; v < nvl; ++v)
            dst[((size_t)sr.side * nvl + v) * D::NP2 + pq] =
                face_eval<NP, GL>(sW + (d * nvl + v) * D::NP3, sB, ls, a, b) + eta * ((o[v * D::NP2] * nd) * wsum);
        }
      }
      // volume gradients: + sJ * Flux * L_HatMinus(depth)
      {
        const auto lhm = sB + D::O_LHM;
#pragma unroll
        for (int ls = 1; ls <= 6; ++ls) {
          int ab, me;
          if (ls == XI_MINUS || ls == XI_PLUS) {
            ab = j + NP * k;
            me = i;
          } else if (ls == ETA_MINUS || ls == ETA_PLUS) {
            ab = i + NP * k;
            me = j;
          } else {
            ab = i + NP * j;
            me = k;
          }
          const int depth = is_minus_side(ls) ? me : NP - 1 - me;
          if (GL && depth != 0) continue;
          const double lh = lhm[depth];
          const auto o = sLF + (ls - 1) * 8 * D::NP2 + ab;
          const double n0 = o[5 * D::NP2], n1 = o[6 * D::NP2], n2 = o[7 * D::NP2];
#pragma unroll
          for (int v = 0; v < nvl; ++v) {
            const double f = o[v * D::NP2];
            gx[v] = gx[v] + sJ * (f * n0) * lh;
            gy[v] = gy[v] + sJ * (f * n1) * lh;
            gz[v] = gz[v] + sJ * (f * n2) * lh;
          }
        }
      }
      if (A.gvol[0]) {
#pragma unroll
        for (int v = 0; v < nvl; ++v) {
          A.gvol[0][((size_t)e * nvl + v) * D::NP3 + n] = gx[v];
          A.gvol[1][((size_t)e * nvl + v) * D::NP3 + n] = gy[v];
          A.gvol[2][((size_t)e * nvl + v) * D::NP3 + n] = gz[v];
        }
      }
    } else {
    // ---- 3. surface part, lifting_br1.t90:195-292 (same sign for master and slave element), then *sJ (applyjacobian.t90)
    {
      const auto lhm = sB + D::O_LHM;
      const auto lhp = sB + D::O_LHP;
#pragma unroll
      for (int ls = 1; ls <= 6; ++ls) {
        int ab, me;
        if (ls == XI_MINUS || ls == XI_PLUS) {
          ab = j + NP * k;
          me = i;
        } else if (ls == ETA_MINUS || ls == ETA_PLUS) {
          ab = i + NP * k;
          me = j;
        } else {
          ab = i + NP * j;
          me = k;
        }
        const bool minus = is_minus_side(ls);
        if (GL && me != (minus ? 0 : NP - 1)) continue;
        const double lh = minus ? lhm[GL ? 0 : me] : lhp[GL ? NP - 1 : me];
        const auto o = sW + (ls - 1) * 8 * D::NP2 + ab;
        const double n0 = o[5 * D::NP2], n1 = o[6 * D::NP2], n2 = o[7 * D::NP2];
#pragma unroll
        for (int v = 0; v < nvl; ++v) {
          const double f = o[v * D::NP2];
          gx[v] = gx[v] + (f * n0) * lh;
          gy[v] = gy[v] + (f * n1) * lh;
          gz[v] = gz[v] + (f * n2) * lh;
        }
      }
#pragma unroll
      for (int v = 0; v < nvl; ++v) {
        gx[v] = gx[v] * sJ;
        gy[v] = gy[v] * sJ;
        gz[v] = gz[v] * sJ;
      }
    }
    elem_sync<NP>();  // everyone is done with the face staging area
#pragma unroll
    for (int v = 0; v < nvl; ++v) {
      sW[(0 * nvl + v) * D::NP3 + n] = gx[v];
      sW[(1 * nvl + v) * D::NP3 + n] = gy[v];
      sW[(2 * nvl + v) * D::NP3 + n] = gz[v];
    }
    if (A.gvol[0]) {
#pragma unroll
      for (int v = 0; v < nvl; ++v) {
        A.gvol[0][((size_t)e * nvl + v) * D::NP3 + n] = gx[v];
        A.gvol[1][((size_t)e * nvl + v) * D::NP3 + n] = gy[v];
        A.gvol[2][((size_t)e * nvl + v) * D::NP3 + n] = gz[v];
      }
    }
    elem_sync<NP>();
    // ---- 4. ProlongToFaceLifting, lifting_br1.t90:150-174
    for (int t = tid; t < 6 * D::NP2; t += D::NP3) {
      const int ls = t / D::NP2 + 1, pq = t % D::NP2;
      const SideRef sr = A.m.e2s[e * 6 + ls - 1];
      const int ab = A.m.s2v2[(sr.flip * 6 + ls - 1) * D::NP2 + pq];
      const int a = ab % NP, b = ab / NP;
      for (int d = 0; d < 3; ++d) {
        double* dst = (sr.flip == 0) ? A.gm[d] : A.gs[d];
        for (int v = 0; v < nvl; ++v)
          dst[((size_t)sr.side * nvl + v) * D::NP2 + pq] = face_eval<NP, GL>(sW + (d * nvl + v) * D::NP3, sB, ls, a, b);
      }
    }
    }  // BR1
    // ---- 5. viscous (+ advective in weak form) fluxes of this node in the reference element, volint.f90:378-422
    double F[5], G[5], H[5];
    {
      double fv[4], gv[4], hv[4];
      eval_diff_flux3d(q, P[TEMP], P[VEL1], P[VEL2], P[VEL3], gx + LV, gy + LV, gz + LV, fv, gv, hv);
      // eval_diff_flux3d reads g?[0..2] = velocity gradients and g?[3] = temperature gradient: lifting sets keep T right
      // behind the velocities (eos.h:118,127), so the same pointer offset serves both.
      double fa[5] = {0., fv[0], fv[1], fv[2], fv[3]}, ga[5] = {0., gv[0], gv[1], gv[2], gv[3]},
             ha[5] = {0., hv[0], hv[1], hv[2], hv[3]};
      if (!SPLIT) {
        double fe[5], ge[5], he[5];
        eval_flux3d(U, P, fe, ge, he);
#pragma unroll
        for (int v = 0; v < 5; ++v) {
          fa[v] = fe[v] + fa[v];
          ga[v] = ge[v] + ga[v];
          ha[v] = he[v] + ha[v];
        }
      }
#pragma unroll
      for (int v = 0; v < 5; ++v) {
        F[v] = fa[v] * M[0] + ga[v] * M[1] + ha[v] * M[2];
        G[v] = fa[v] * M[3] + ga[v] * M[4] + ha[v] * M[5];
        H[v] = fa[v] * M[6] + ga[v] * M[7] + ha[v] * M[8];
      }
    }
    elem_sync<NP>();  // face threads are done reading the gradients
#pragma unroll
    for (int v = 0; v < 5; ++v) {
      sW[(0 + v) * D::NP3 + n] = F[v];
      sW[(5 + v) * D::NP3 + n] = G[v];
      sW[(10 + v) * D::NP3 + n] = H[v];
    }
    elem_sync<NP>();
  } else if (!SPLIT) {
    double fe[5], ge[5], he[5];
    eval_flux3d(U, P, fe, ge, he);
#pragma unroll
    for (int v = 0; v < 5; ++v) {
      sW[(0 + v) * D::NP3 + n] = fe[v] * M[0] + ge[v] * M[1] + he[v] * M[2];
      sW[(5 + v) * D::NP3 + n] = fe[v] * M[3] + ge[v] * M[4] + he[v] * M[5];
      sW[(10 + v) * D::NP3 + n] = fe[v] * M[6] + ge[v] * M[7] + he[v] * M[8];
    }
    elem_sync<NP>();
  }

  if (PARA || !SPLIT) {
    // ---- 6. weak-form derivative, volint.f90:172-184 (term order xi, zeta, eta)
    const int bx = n - i, by = n - j * NP, bz = n - k * D::NP2;
    constexpr int V0 = (SPLIT ? 1 : 0);  // the viscous density flux is identically zero
#pragma unroll
    for (int v = V0; v < 5; ++v) {
      const auto f = sW + (0 + v) * D::NP3;
      const auto g = sW + (5 + v) * D::NP3;
      const auto h = sW + (10 + v) * D::NP3;
      double s = dh[i] * f[bx] + dh[k] * h[bz] + dh[j] * g[by];
#pragma unroll
      for (int l = 1; l < NP; ++l) s = s + dh[l * NP + i] * f[bx + l] + dh[l * NP + k] * h[bz + l * D::NP2] + dh[l * NP + j] * g[by + l * NP];
      ut[v] = s;
    }
  }

  if (SPLIT) {
    if (PARA) elem_sync<NP>();  // weak-form reads of the work area are done
#pragma unroll
    for (int c = 0; c < 9; ++c) sW[c * D::NP3 + n] = M[c];
    elem_sync<NP>();
    switch (q.splitType) {
      case FLEXI_SPLITDG_SD: split_lines<NP, FLEXI_SPLITDG_SD>(sP, sW, sB, i, j, k, n, M, ut, q.kappa); break;
      case FLEXI_SPLITDG_MO: split_lines<NP, FLEXI_SPLITDG_MO>(sP, sW, sB, i, j, k, n, M, ut, q.kappa); break;
      case FLEXI_SPLITDG_DU: split_lines<NP, FLEXI_SPLITDG_DU>(sP, sW, sB, i, j, k, n, M, ut, q.kappa); break;
      case FLEXI_SPLITDG_KG: split_lines<NP, FLEXI_SPLITDG_KG>(sP, sW, sB, i, j, k, n, M, ut, q.kappa); break;
      case FLEXI_SPLITDG_CH: split_lines<NP, FLEXI_SPLITDG_CH>(sP, sW, sB, i, j, k, n, M, ut, q.kappa); break;
      case FLEXI_SPLITDG_IR: split_lines<NP, FLEXI_SPLITDG_IR>(sP, sW, sB, i, j, k, n, M, ut, q.kappa); break;
      default: split_lines<NP, FLEXI_SPLITDG_PI>(sP, sW, sB, i, j, k, n, M, ut, q.kappa);
    }
  }
#pragma unroll
  for (int v = 0; v < 5; ++v) A.Ut[((size_t)e * 5 + v) * D::NP3 + n] = ut[v];
  elem_exit<NP>();
}

// ---------------------------------------------------------------------------------------------------------------------------
// k_flux: one thread per face node of the side range.  fillflux.f90:40-215, riemann.f90:267-352,448-498
// SOLVER != 0 fixes the Riemann solver at compile time (the switch over all solvers costs registers, i.e. occupancy of a
// memory-bound kernel); SOLVER == 0 reads it from Eqn::riemann.
template <int NP, int SOLVER>
__global__ void __launch_bounds__(128) k_flux(FluxArgs A) {
  using D = Dim<NP>;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)A.nSidesRange * D::NP2) return;
  const int s = A.firstSide + (int)(idx / D::NP2);
  const int pq = (int)(idx % D::NP2);
  const Eqn& q = A.q;
  double UL[5], UR[5], PL[6], PR[6];
#pragma unroll
  for (int v = 0; v < 5; ++v) {
    UL[v] = A.U_master[((size_t)s * 5 + v) * D::NP2 + pq];
    UR[v] = A.U_slave[((size_t)s * 5 + v) * D::NP2 + pq];
  }
  cons_to_prim(q, UL, PL);  // equation.f90:189-216: face primitives come from the conservative traces
  cons_to_prim(q, UR, PR);
  const double* g = A.m.Geo + (size_t)s * 10 * D::NP2 + pq;
  const double nv[3] = {g[0], g[D::NP2], g[2 * D::NP2]};
  const double t1[3] = {g[3 * D::NP2], g[4 * D::NP2], g[5 * D::NP2]};
  const double t2[3] = {g[6 * D::NP2], g[7 * D::NP2], g[8 * D::NP2]};
  const double surf = g[9 * D::NP2];
  double F[5];
  riemann_point(q, SOLVER ? SOLVER : q.riemann, UL, UR, PL, PR, nv, t1, t2, F);
  if (A.parabolic && A.vt) {
    // ViscousFlux_Side (riemann.f90:448-498) with the normal contraction already done by the two element kernels
#pragma unroll
    for (int c = 0; c < 4; ++c)
      F[c + 1] = F[c + 1] + 0.5 * (A.fvm[((size_t)s * 4 + c) * D::NP2 + pq] + A.fvs[((size_t)s * 4 + c) * D::NP2 + pq]);
  } else if (A.parabolic) {
    const int nvl = q.nvl, o = q.lv_vel1;
    double gL[3][4], gR[3][4];
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        gL[d][c] = A.gm[d][((size_t)s * nvl + o + c) * D::NP2 + pq];
        gR[d][c] = A.gs[d][((size_t)s * nvl + o + c) * D::NP2 + pq];
      }
    double fL[4], gLf[4], hL[4], fR[4], gRf[4], hR[4];
    eval_diff_flux3d(q, PL[TEMP], PL[VEL1], PL[VEL2], PL[VEL3], gL[0], gL[1], gL[2], fL, gLf, hL);
    eval_diff_flux3d(q, PR[TEMP], PR[VEL1], PR[VEL2], PR[VEL3], gR[0], gR[1], gR[2], fR, gRf, hR);
#pragma unroll
    for (int c = 0; c < 4; ++c) F[c + 1] = F[c + 1] + 0.5 * (nv[0] * (fL[c] + fR[c]) + nv[1] * (gLf[c] + gRf[c]) + nv[2] * (hL[c] + hR[c]));
  }
#pragma unroll
  for (int v = 0; v < 5; ++v) A.Flux[((size_t)s * 5 + v) * D::NP2 + pq] = F[v] * surf;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Boundary sides.  k_bc_pre turns the lifting boundary flux of Lifting_GetBoundaryFlux (getboundaryflux.f90:855-940, strong form)
// into a virtual slave trace: the element kernels evaluate 0.5*(UPrim(U_slave)-UPrim_master)*SurfElem on every side, so U_slave of
// a BC side is set to the conservative state whose primitives are 2*F_bc - UPrim_master (for Riemann-type BCs that is the
// boundary state itself).  k_flux_bc is GetBoundaryFlux (getboundaryflux.f90:480-790) * SurfElem (fillflux.f90:179-186).
struct BCArgs {
  Mesh m;
  Eqn q;
  BCData bc;
  double t;
  int parabolic;
  const double* U_master;
  double* U_slave;
  const double* gm[3];
  double* Flux;
};

template <int NP>
__device__ __forceinline__ void bc_load(const BCArgs& A, int s, int pq, double* Um, double* Pm, double* nv, double* t1, double* t2,
                                        double* xf, double& surf) {
  using D = Dim<NP>;
#pragma unroll
  for (int v = 0; v < 5; ++v) Um[v] = A.U_master[((size_t)s * 5 + v) * D::NP2 + pq];
  cons_to_prim(A.q, Um, Pm);
  const double* g = A.m.Geo + (size_t)s * 10 * D::NP2 + pq;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    nv[c] = g[c * D::NP2];
    t1[c] = g[(3 + c) * D::NP2];
    t2[c] = g[(6 + c) * D::NP2];
    xf[c] = A.bc.faceX ? A.bc.faceX[((size_t)s * 3 + c) * D::NP2 + pq] : 0.;
  }
  surf = g[9 * D::NP2];
}

template <int NP>
__global__ void __launch_bounds__(128) k_bc_pre(BCArgs A) {
  using D = Dim<NP>;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)A.bc.nBCSides * D::NP2) return;
  const int s = (int)(idx / D::NP2), pq = (int)(idx % D::NP2);
  const int type = A.bc.bcType[s], state = A.bc.bcState[s];
  double Um[5], Pm[6], Pb[6], nv[3], t1[3], t2[3], xf[3], surf;
  bc_load<NP>(A, s, pq, Um, Pm, nv, t1, t2, xf, surf);
  boundary_state(A.q, A.bc, type, state, A.t, Pm, nv, t1, t2, xf, Pb);
  double V[6];
  if (type == 2) {
#pragma unroll
    for (int v = 0; v < 6; ++v) V[v] = Pb[v];
  } else if (type == 3 || type == 4) {  // F = (rho_b, 0, T_b)
    V[DENS] = 2. * Pb[DENS] - Pm[DENS];
    V[VEL1] = -Pm[VEL1];
    V[VEL2] = -Pm[VEL2];
    V[VEL3] = -Pm[VEL3];
    V[TEMP] = 2. * Pb[TEMP] - Pm[TEMP];
    V[PRES] = V[DENS] * A.q.R * V[TEMP];
  } else {  // 9: F = (rho_m, v_b, T_m)
    V[DENS] = Pm[DENS];
    V[VEL1] = 2. * Pb[VEL1] - Pm[VEL1];
    V[VEL2] = 2. * Pb[VEL2] - Pm[VEL2];
    V[VEL3] = 2. * Pb[VEL3] - Pm[VEL3];
    V[TEMP] = Pm[TEMP];
    V[PRES] = V[DENS] * A.q.R * V[TEMP];
  }
  double Uv[5];
  prim_to_cons(A.q, V, Uv);
#pragma unroll
  for (int v = 0; v < 5; ++v) A.U_slave[((size_t)s * 5 + v) * D::NP2 + pq] = Uv[v];
}

template <int NP>
__global__ void __launch_bounds__(128) k_flux_bc(BCArgs A) {
  using D = Dim<NP>;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)A.bc.nBCSides * D::NP2) return;
  const int s = (int)(idx / D::NP2), pq = (int)(idx % D::NP2);
  const int type = A.bc.bcType[s], state = A.bc.bcState[s];
  const Eqn& q = A.q;
  double Um[5], Pm[6], Pb[6], nv[3], t1[3], t2[3], xf[3], surf;
  bc_load<NP>(A, s, pq, Um, Pm, nv, t1, t2, xf, surf);
  boundary_state(q, A.bc, type, state, A.t, Pm, nv, t1, t2, xf, Pb);
  double F[5];
  double gx[4], gy[4], gz[4];
  if (A.parabolic) {
    const int nvl = q.nvl, o = q.lv_vel1;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      gx[c] = A.gm[0][((size_t)s * nvl + o + c) * D::NP2 + pq];
      gy[c] = A.gm[1][((size_t)s * nvl + o + c) * D::NP2 + pq];
      gz[c] = A.gm[2][((size_t)s * nvl + o + c) * D::NP2 + pq];
    }
  }
  if (type == 2) {  // Riemann-type BC: getboundaryflux.f90:561-578
    double Ucm[5], Ucb[5];
    prim_to_cons(q, Pm, Ucm);
    prim_to_cons(q, Pb, Ucb);
    riemann_point(q, A.bc.riemannBC, Ucm, Ucb, Pm, Pb, nv, t1, t2, F);
    if (A.parabolic) {
      double fL[4], gL[4], hL[4], fR[4], gR[4], hR[4];
      eval_diff_flux3d(q, Pm[TEMP], Pm[VEL1], Pm[VEL2], Pm[VEL3], gx, gy, gz, fL, gL, hL);
      eval_diff_flux3d(q, Pb[TEMP], Pb[VEL1], Pb[VEL2], Pb[VEL3], gx, gy, gz, fR, gR, hR);
#pragma unroll
      for (int c = 0; c < 4; ++c) F[c + 1] = F[c + 1] + 0.5 * (nv[0] * (fL[c] + fR[c]) + nv[1] * (gL[c] + gR[c]) + nv[2] * (hL[c] + hR[c]));
    }
  } else {  // walls: getboundaryflux.f90:580-780
    F[0] = 0.;
    F[1] = Pb[PRES] * nv[0];
    F[2] = Pb[PRES] * nv[1];
    F[3] = Pb[PRES] * nv[2];
    F[4] = 0.;
    if (A.parabolic) {
      if (type == 9) {  // BCGradMat = I - n n^T: no normal derivatives
        const double B11 = 1. - nv[0] * nv[0], B22 = 1. - nv[1] * nv[1], B33 = 1. - nv[2] * nv[2];
        const double B12 = -nv[0] * nv[1], B13 = -nv[0] * nv[2], B32 = -nv[2] * nv[1];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const double x = gx[c], y = gy[c], z = gz[c];
          gx[c] = B11 * x + B12 * y + B13 * z;
          gy[c] = B12 * x + B22 * y + B32 * z;
          gz[c] = B13 * x + B32 * y + B33 * z;
        }
      }
      double f[4], g[4], h[4];
      eval_diff_flux3d(q, Pb[TEMP], Pb[VEL1], Pb[VEL2], Pb[VEL3], gx, gy, gz, f, g, h);
      if (type == 3) f[3] = g[3] = h[3] = 0.;  // adiabatic wall: no energy flux
#pragma unroll
      for (int c = 0; c < 4; ++c) F[c + 1] = F[c + 1] + nv[0] * f[c] + nv[1] * g[c] + nv[2] * h[c];
    }
  }
#pragma unroll
  for (int v = 0; v < 5; ++v) A.Flux[((size_t)s * 5 + v) * D::NP2 + pq] = F[v] * surf;
}

// ---------------------------------------------------------------------------------------------------------------------------
// k_update: SurfInt gather + Ut=-Ut + ApplyJacobian + LSERK + ProlongToFace(new U) through a shared-memory copy of the new
// state, so that the traces leave the SM as whole 512-byte planes.  (A barrier-free Gauss-Lobatto variant in which the owner of a
// boundary node wrote its trace entries directly was measured 19 % slower at 48^3, N=7: 3.41 vs 2.87 ms; the scattered 8-byte
// stores cost more than the staging.)
// ENT: the traces of the new state are prolonged in entropy variables (split form on Gauss nodes, ProlongToFaceEntropy)
// FILT: OVERINTEGRATIONTYPE_CUTOFF (dg.f90:663-676, dg/overintegration.f90:162-187): Ut is filtered with OverintegrationMat along
// xi, eta, zeta (filter.f90:264-288 -> ChangeBasis3D) between the source term and ApplyJacobian, in the shared planes that
// later stage the new state.  A separate instantiation: the unfiltered kernels are not touched by it.
// SPONGE: Sponge(Ut) between the source term and ApplyJacobian (sponge.f90:612-678).  A separate instantiation as well: with
// the sponge code in the plain kernel its register count crossed 64 and the kernel lost half its resident warps (+32 % time).
template <int NP, bool GL, bool ENT = false, bool FILT = false, bool SPONGE = false>
__global__ void __launch_bounds__(Dim<NP>::THREADS) k_update(UpdArgs A) {
  using D = Dim<NP>;
  extern __shared__ double sm_raw[];
  auto sm = elem_smem<NP>(sm_raw);
  auto sB = sm;
  auto sU = sm + D::NBASIS;  // 5 planes
  const int tid = elem_tid<NP>();
  const int e = A.elems ? A.elems[elem_blk<NP>()] : elem_blk<NP>();
  const int i = tid % NP, j = (tid / NP) % NP, k = tid / D::NP2;
  const int n = tid;
  load_basis<NP>(sB, A.m.basis, tid, D::NP3);
  double ut[5];
#pragma unroll
  for (int v = 0; v < 5; ++v) ut[v] = A.UtVol[((size_t)e * 5 + v) * D::NP3 + n];
  const double sJ = A.m.sJ[(size_t)e * D::NP3 + n];
  // the RK register and the state are read here, together with everything else, and not next to their stores below: the
  // stores may alias the loads as far as the compiler knows, which would serialise five load/store round trips per thread
  double uOld[5], rkOld[5];
#pragma unroll
  for (int v = 0; v < 5; ++v) {
    const size_t g = ((size_t)e * 5 + v) * D::NP3 + n;
    uOld[v] = (A.mode != 0) ? A.U[g] : 0.;
    rkOld[v] = (A.mode == 2) ? A.UtTmp[g] : 0.;
  }
  // Sponge(Ut) (sponge.f90:612-678, called between CalcSource and ApplyJacobian, dg.f90:659-661): SpongeMat*(U - SpBaseFlow)
  double sp[5] = {0., 0., 0., 0., 0.};
  const int iSp = SPONGE ? A.spIdx[e] : -1;
  if (SPONGE && iSp >= 0) {
    const double smat = A.spMat[(size_t)iSp * D::NP3 + n];
#pragma unroll
    for (int v = 0; v < 5; ++v) {
      const size_t g = ((size_t)e * 5 + v) * D::NP3 + n;
      sp[v] = smat * (((A.mode != 0) ? uOld[v] : A.U[g]) - A.spBase[g]);
    }
  }
  elem_sync<NP>();
  const auto lhm = sB + D::O_LHM;
  const auto lhp = sB + D::O_LHP;
#pragma unroll
  for (int ls = 1; ls <= 6; ++ls) {  // surfint.t90:475-616 / :21-239, DoSurfInt :683-766
    int ab, me;
    if (ls == XI_MINUS || ls == XI_PLUS) {
      ab = j + NP * k;
      me = i;
    } else if (ls == ETA_MINUS || ls == ETA_PLUS) {
      ab = i + NP * k;
      me = j;
    } else {
      ab = i + NP * j;
      me = k;
    }
    const bool minus = is_minus_side(ls);
    if (GL && me != (minus ? 0 : NP - 1)) continue;
    const double lh = minus ? lhm[GL ? 0 : me] : lhp[GL ? NP - 1 : me];
    const SideRef sr = A.m.e2s[e * 6 + ls - 1];
    const int pq = A.m.v2s[(sr.flip * 6 + ls - 1) * D::NP2 + ab];
    const double sgn = (sr.flip == 0) ? 1. : -1.;  // surfint.t90:583: slave element sees -Flux_slave
#pragma unroll
    for (int v = 0; v < 5; ++v) ut[v] = ut[v] + (sgn * A.Flux[((size_t)sr.side * 5 + v) * D::NP2 + pq]) * lh;
  }
  if (FILT) {
    double w[5];
#pragma unroll
    for (int v = 0; v < 5; ++v) w[v] = -ut[v];  // dg.f90:656
    if (A.doSource) {                           // exactfunc.f90:809-813: Ut += src/sJ
      const double* x = A.xGP + (size_t)e * 3 * D::NP3 + n;
      const double arg = A.srcOmega * (x[0] + x[D::NP3] + x[2 * D::NP3]) - A.srcA * A.t;
      const double cosX = cos(arg), sinX = sin(arg), sinX2 = 2. * sinX * cosX;
      double src[5];
      src[0] = A.srcTmp[0] * cosX;
      src[1] = src[2] = src[3] = A.srcTmp[1] * cosX + A.srcTmp[2] * sinX2;
      src[4] = A.srcTmp[3] * cosX + A.srcTmp[4] * sinX2 + A.srcTmp[5] * sinX;
#pragma unroll
      for (int v = 0; v < 5; ++v) w[v] = w[v] + src[v] / sJ;
    }
    if (SPONGE && iSp >= 0) {
#pragma unroll
      for (int v = 0; v < 5; ++v) w[v] = w[v] - sp[v];
    }
    // three 1-D sweeps through the 5 shared planes; sums start from zero and run over ascending input index (changeBasis.t90)
    auto sweep3 = [&](const double* mat) {
#pragma unroll
      for (int dir = 0; dir < 3; ++dir) {
        const int me = (dir == 0) ? i : (dir == 1) ? j : k;
        const int stride = (dir == 0) ? 1 : (dir == 1) ? NP : D::NP2;
        const int base = n - me * stride;
#pragma unroll
        for (int v = 0; v < 5; ++v) sU[v * D::NP3 + n] = w[v];
        elem_sync<NP>();
        const double* row = mat + me * NP;
#pragma unroll
        for (int v = 0; v < 5; ++v) {
          double acc = 0.;
#pragma unroll
          for (int l = 0; l < NP; ++l) acc = acc + __ldg(row + l) * sU[v * D::NP3 + base + l * stride];
          w[v] = acc;
        }
        elem_sync<NP>();  // every read of this sweep is done before the planes are overwritten (next sweep / new state below)
      }
    };
    sweep3(A.filt);
    if (A.filt2) {
      // conservative cut-off: JU_t projected to NUnder (the padded rows give zero), times sJNUnder, interpolated back to N; the
      // Jacobian has been applied on NUnder, ApplyJacobianCons is skipped (dg.f90:664-668)
      const double sj = A.sJNU[(size_t)e * D::NP3 + n];
#pragma unroll
      for (int v = 0; v < 5; ++v) w[v] = w[v] * sj;
      sweep3(A.filt2);
#pragma unroll
      for (int v = 0; v < 5; ++v) ut[v] = w[v];
    } else {
#pragma unroll
      for (int v = 0; v < 5; ++v) ut[v] = w[v] * sJ;  // ApplyJacobianCons, dg.f90:672
    }
  } else if (A.doSource) {  // Ut = -Ut; Ut += src/sJ; Ut *= sJ  (dg.f90:656-675, exactfunc.f90:809-813)
    const double* x = A.xGP + (size_t)e * 3 * D::NP3 + n;
    const double arg = A.srcOmega * (x[0] + x[D::NP3] + x[2 * D::NP3]) - A.srcA * A.t;
    const double cosX = cos(arg), sinX = sin(arg), sinX2 = 2. * sinX * cosX;
    double src[5];
    src[0] = A.srcTmp[0] * cosX;
    src[1] = src[2] = src[3] = A.srcTmp[1] * cosX + A.srcTmp[2] * sinX2;
    src[4] = A.srcTmp[3] * cosX + A.srcTmp[4] * sinX2 + A.srcTmp[5] * sinX;
#pragma unroll
    for (int v = 0; v < 5; ++v) ut[v] = SPONGE ? (((-ut[v]) + src[v] / sJ) - sp[v]) * sJ : ((-ut[v]) + src[v] / sJ) * sJ;
  } else if (SPONGE && iSp >= 0) {
#pragma unroll
    for (int v = 0; v < 5; ++v) ut[v] = ((-ut[v]) - sp[v]) * sJ;
  } else {
#pragma unroll
    for (int v = 0; v < 5; ++v) ut[v] = (-ut[v]) * sJ;  // dg.f90:656,675
  }
  if (A.mode == 0) {
#pragma unroll
    for (int v = 0; v < 5; ++v) A.UtOut[((size_t)e * 5 + v) * D::NP3 + n] = ut[v];
    elem_exit<NP>();
    return;
  }
  // timestep.f90:126-131
#pragma unroll
  for (int v = 0; v < 5; ++v) {
    const size_t g = ((size_t)e * 5 + v) * D::NP3 + n;
    double tmp = ut[v];
    if (A.mode == 2) tmp = rkOld[v] * (-A.rkA) + ut[v];
    A.UtTmp[g] = tmp;
    const double u = uOld[v] + tmp * A.b_dt;
    A.U[g] = u;
    sU[v * D::NP3 + n] = u;
  }
  if (ENT) {  // own node: cons -> prim -> entropy variables, in place (only this thread touches its node before the barrier)
    double Uc[5], P[6], w[5];
#pragma unroll
    for (int v = 0; v < 5; ++v) Uc[v] = sU[v * D::NP3 + n];
    cons_to_prim(A.q, Uc, P);
    prim_to_entropy(A.q, P, w);
#pragma unroll
    for (int v = 0; v < 5; ++v) sU[v * D::NP3 + n] = w[v];
  }
  elem_sync<NP>();
  for (int t = tid; t < 6 * D::NP2; t += D::NP3) {
    const int ls = t / D::NP2 + 1, pq = t % D::NP2;
    const SideRef sr = A.m.e2s[e * 6 + ls - 1];
    const int ab = A.m.s2v2[(sr.flip * 6 + ls - 1) * D::NP2 + pq];
    const int a = ab % NP, b = ab / NP;
    double* dst = (sr.flip == 0) ? A.U_master : A.U_slave;
    if (ENT) {
      double w[5], Uc[5];
#pragma unroll
      for (int v = 0; v < 5; ++v) w[v] = face_eval<NP, GL>(sU + v * D::NP3, sB, ls, a, b);
      entropy_to_cons(A.q, w, Uc);
#pragma unroll
      for (int v = 0; v < 5; ++v) dst[((size_t)sr.side * 5 + v) * D::NP2 + pq] = Uc[v];
    } else {
#pragma unroll
      for (int v = 0; v < 5; ++v) dst[((size_t)sr.side * 5 + v) * D::NP2 + pq] = face_eval<NP, GL>(sU + v * D::NP3, sB, ls, a, b);
    }
  }
  elem_exit<NP>();
}

// ---------------------------------------------------------------------------------------------------------------------------
// k_surfcorr: the two-point surface correction of the split form on Gauss nodes, dg/surfint.t90:90-137 (master) and :190-209
// (slave), as an element gather added to the raw volume integral (k_update then adds the numerical flux like on any Gauss
// element):  Ut(l) -= 1/2 (sum_k Leg(k) FluxB(k) - FluxB(l)) * NormalSign * LegHat(l)  for every line behind a face node, with
// FluxB(k) the two-point volume flux between line node k and the face state (metrics of the node and Ja_Face / Ja_slave).
struct SurfCorrArgs {
  Mesh m;
  Eqn q;
  const int32_t* elems;
  int nList;
  const double* U;         // [e][5][NP3]
  const double* U_master;  // [s][5][NP2]
  const double* U_slave;
  const double* JaF;       // Ja_Face  [s][9][NP2], plane dd + 3*c = Ja(dd,c)
  const double* JaS;       // Ja_slave
  double* Ut;              // raw volume integral, updated in place
};

template <int NP>
__global__ void __launch_bounds__(Dim<NP>::THREADS) k_surfcorr(SurfCorrArgs A) {
  using D = Dim<NP>;
  extern __shared__ double sm_raw[];
  auto sm = elem_smem<NP>(sm_raw);
  auto sB = sm;
  auto sFB = sm + D::NBASIS;  // Leg(l)*FluxB(l) of the current local side, [5][NP3]
  const int tid = elem_tid<NP>(), n = tid;
  const int e = A.elems ? A.elems[elem_blk<NP>()] : elem_blk<NP>();
  const int i = tid % NP, j = (tid / NP) % NP, k = tid / D::NP2;
  const Eqn& q = A.q;
  load_basis<NP>(sB, A.m.basis, tid, D::NP3);
  double Uc[5], P[6], M[9];
#pragma unroll
  for (int v = 0; v < 5; ++v) Uc[v] = A.U[((size_t)e * 5 + v) * D::NP3 + n];
  cons_to_prim(q, Uc, P);
#pragma unroll
  for (int c = 0; c < 9; ++c) M[c] = A.m.Met[((size_t)e * 9 + c) * D::NP3 + n];
  const Node nd{P[DENS], P[VEL1], P[VEL2], P[VEL3], P[PRES], Uc[ENER]};
  double corr[5] = {0., 0., 0., 0., 0.};
  elem_sync<NP>();
  for (int ls = 1; ls <= 6; ++ls) {
    // line through this node along the side's normal direction; (a,b) = volume-face coordinates, l = own position on the line
    const int idx = (ls == XI_MINUS || ls == XI_PLUS) ? 0 : (ls == ETA_MINUS || ls == ETA_PLUS) ? 1 : 2;
    const int ab = idx == 0 ? j + NP * k : idx == 1 ? i + NP * k : i + NP * j;
    const int l = idx == 0 ? i : idx == 1 ? j : k;
    const int stride = idx == 0 ? 1 : idx == 1 ? NP : D::NP2;
    const bool minus = is_minus_side(ls);
    const auto Leg = sB + (minus ? D::O_LM : D::O_LP);
    const auto LegHat = sB + (minus ? D::O_LHM : D::O_LHP);
    const double nsign = minus ? -1. : 1.;  // NormalSigns (mesh_vars.f90:105-111)
    const SideRef sr = A.m.e2s[e * 6 + ls - 1];
    const int pq = A.m.v2s[(sr.flip * 6 + ls - 1) * D::NP2 + ab];
    const double* Uf = ((sr.flip == 0) ? A.U_master : A.U_slave) + (size_t)sr.side * 5 * D::NP2 + pq;
    const double* Ja = ((sr.flip == 0) ? A.JaF : A.JaS) + (size_t)sr.side * 9 * D::NP2 + pq;
    double Ufc[5], Pf[6];
#pragma unroll
    for (int v = 0; v < 5; ++v) Ufc[v] = Uf[v * D::NP2];
    cons_to_prim(q, Ufc, Pf);  // UPrim_master / UPrim_slave, equation.f90:189-216
    const Node nf{Pf[DENS], Pf[VEL1], Pf[VEL2], Pf[VEL3], Pf[PRES], Ufc[ENER]};
    const double Jv[3] = {Ja[(idx + 0) * D::NP2], Ja[(idx + 3) * D::NP2], Ja[(idx + 6) * D::NP2]};
    double FB[5];
    switch (q.splitType) {  // SplitDGVolume_pointer(U, UPrim, U_face, UPrim_face, Metrics, Ja_face), surfint.t90:118-121
      case FLEXI_SPLITDG_SD: split_volume_flux<FLEXI_SPLITDG_SD>(nd, nf, M + 3 * idx, Jv, FB, q.kappa); break;
      case FLEXI_SPLITDG_MO: split_volume_flux<FLEXI_SPLITDG_MO>(nd, nf, M + 3 * idx, Jv, FB, q.kappa); break;
      case FLEXI_SPLITDG_DU: split_volume_flux<FLEXI_SPLITDG_DU>(nd, nf, M + 3 * idx, Jv, FB, q.kappa); break;
      case FLEXI_SPLITDG_KG: split_volume_flux<FLEXI_SPLITDG_KG>(nd, nf, M + 3 * idx, Jv, FB, q.kappa); break;
      case FLEXI_SPLITDG_CH: split_volume_flux<FLEXI_SPLITDG_CH>(nd, nf, M + 3 * idx, Jv, FB, q.kappa); break;
      case FLEXI_SPLITDG_IR: split_volume_flux<FLEXI_SPLITDG_IR>(nd, nf, M + 3 * idx, Jv, FB, q.kappa); break;
      default: split_volume_flux<FLEXI_SPLITDG_PI>(nd, nf, M + 3 * idx, Jv, FB, q.kappa);
    }
#pragma unroll
    for (int v = 0; v < 5; ++v) sFB[v * D::NP3 + n] = FB[v] * Leg[l];
    elem_sync<NP>();
    const int base = n - l * stride;
#pragma unroll
    for (int v = 0; v < 5; ++v) {
      double sum = 0.;
      for (int kk = 0; kk < NP; ++kk) sum = sum + sFB[v * D::NP3 + base + kk * stride];
      corr[v] = corr[v] + (-1. / 2. * (sum - FB[v]) * nsign) * LegHat[l];
    }
    elem_sync<NP>();
  }
#pragma unroll
  for (int v = 0; v < 5; ++v) {
    const size_t g = ((size_t)e * 5 + v) * D::NP3 + n;
    A.Ut[g] = A.Ut[g] + corr[v];
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// operator-level SurfInt (what unitTests/SurfInt.f90 drives): Ut += surface integral of the given side fluxes
struct SurfArgs {
  Mesh m;
  const double* Flux_master;
  const double* Flux_slave;
  double* Ut;
  int nv;
  int firstSide, lastSide;
};
template <int NP, bool GL>
__global__ void __launch_bounds__(Dim<NP>::THREADS) k_surfint(SurfArgs A) {
  using D = Dim<NP>;
  extern __shared__ double sm_raw[];
  auto sm = elem_smem<NP>(sm_raw);
  auto sB = sm;
  const int tid = elem_tid<NP>(), e = elem_blk<NP>();
  const int i = tid % NP, j = (tid / NP) % NP, k = tid / D::NP2;
  load_basis<NP>(sB, A.m.basis, tid, D::NP3);
  elem_sync<NP>();
  const auto lhm = sB + D::O_LHM;
  const auto lhp = sB + D::O_LHP;
  for (int v = 0; v < A.nv; ++v) {
    double ut = A.Ut[((size_t)e * A.nv + v) * D::NP3 + tid];
    for (int ls = 1; ls <= 6; ++ls) {
      int ab, me;
      if (ls == XI_MINUS || ls == XI_PLUS) {
        ab = j + NP * k;
        me = i;
      } else if (ls == ETA_MINUS || ls == ETA_PLUS) {
        ab = i + NP * k;
        me = j;
      } else {
        ab = i + NP * j;
        me = k;
      }
      const bool minus = is_minus_side(ls);
      if (GL && me != (minus ? 0 : NP - 1)) continue;
      const SideRef sr = A.m.e2s[e * 6 + ls - 1];
      if (sr.side < A.firstSide || sr.side > A.lastSide) continue;
      const double lh = minus ? lhm[GL ? 0 : me] : lhp[GL ? NP - 1 : me];
      const int pq = A.m.v2s[(sr.flip * 6 + ls - 1) * D::NP2 + ab];
      const double f = (sr.flip == 0) ? A.Flux_master[((size_t)sr.side * A.nv + v) * D::NP2 + pq]
                                      : -A.Flux_slave[((size_t)sr.side * A.nv + v) * D::NP2 + pq];
      ut = ut + f * lh;
    }
    A.Ut[((size_t)e * A.nv + v) * D::NP3 + tid] = ut;
  }
  elem_exit<NP>();
}

// ---------------------------------------------------------------------------------------------------------------------------
// k_dt: CalcTimeStep, calctimestep.f90:89-227.  One block per element; writes per-element (dtConv, dtVisc, errFlag).
// MetricsAdv / MetricsVisc of InitCalctimestep (calctimestep.f90:42-83) are formed on the fly from the metric terms the kernel
// reads anyway (same expressions, same values) instead of being read from two more 3-plane tables: 120 instead of 168 bytes per
// DOF.  The seven block reductions share one shuffle + one shared-memory stage.
struct DtArgs {
  Mesh m;
  Eqn q;
  const double* U;
  double KappasPr_max;  // MAX(4/3, kappa/Pr), calctimestep.f90:66
  double CFLScale, DFLScale;
  int parabolic;
  double* out;  // [e][3]
};

// The block is NP^3 rounded up to whole warps: the lanes beyond the last node carry neutral values (0) into the warp shuffles
// (a shuffle from a lane that does not exist returns an undefined value).
template <int NP>
__host__ __device__ constexpr int dt_threads() { return Dim<NP>::PAIR ? 1024 : (Dim<NP>::NP3 + 31) / 32 * 32; }
template <int NP>
__global__ void __launch_bounds__(dt_threads<NP>()) k_dt(DtArgs A) {
  using D = Dim<NP>;
  constexpr int NT = dt_threads<NP>(), NW = NT / 32;
  __shared__ double red[7][NW];
  const int tid = threadIdx.x, e = blockIdx.x;
  const Eqn& q = A.q;
  double r[7];  // max over the element of: lambda_adv(3), lambda_visc(3), NaN flag
#pragma unroll
  for (int s = 0; s < 7; ++s) r[s] = 0.;
  // one node per thread; the large degrees (Dim::PAIR) walk their nodes with the block as stride and keep the running maxima
  for (int n = tid; n < D::NP3; n += NT) {
    double rp[7];
#pragma unroll
    for (int s = 0; s < 7; ++s) rp[s] = r[s];
    double U[5], P[6], M[9];
#pragma unroll
    for (int v = 0; v < 5; ++v) U[v] = A.U[((size_t)e * 5 + v) * D::NP3 + n];
    cons_to_prim(q, U, P);
    const double srho = 1. / U[0];
    const double c = sqrt(q.kappa * P[PRES] * srho);
    const double sJ = A.m.sJ[(size_t)e * D::NP3 + n];
#pragma unroll
    for (int cI = 0; cI < 9; ++cI) M[cI] = A.m.Met[((size_t)e * 9 + cI) * D::NP3 + n];
    const double vs[3] = {P[VEL1] * sJ, P[VEL2] * sJ, P[VEL3] * sJ};
    // NaN in the density flags errType 1 (calctimestep.f90:157-160); NaN propagates through the fmax-free sums below
    r[6] = (U[0] != U[0]) ? 1. : 0.;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double m0 = M[3 * d], m1 = M[3 * d + 1], m2 = M[3 * d + 2];
      const double metAdv = sJ * sqrt(m0 * m0 + m1 * m1 + m2 * m2);                                                   // MetricsAdv
      const double metVisc = A.KappasPr_max * ((m0 * sJ) * (m0 * sJ) + (m1 * sJ) * (m1 * sJ) + (m2 * sJ) * (m2 * sJ));  // MetricsVisc
      r[d] = fabs(m0 * vs[0] + m1 * vs[1] + m2 * vs[2]) + c * metAdv;
      r[3 + d] = A.parabolic ? viscosity(q, P[TEMP]) * srho * metVisc : 0.;
      // Fortran MAX may or may not propagate a NaN operand; NaN is flagged separately
      if (r[d] != r[d] || r[3 + d] != r[3 + d]) r[6] = fmax(r[6], 2.);
    }
    if (D::PAIR) {
#pragma unroll
      for (int s = 0; s < 7; ++s) r[s] = fmax(r[s], rp[s]);
    }
  }
#pragma unroll
  for (int s = 0; s < 7; ++s) {
    double v = r[s];
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((tid & 31) == 0) red[s][tid >> 5] = v;
  }
  __syncthreads();
  if (tid == 0) {
    double mx[7];
    for (int s = 0; s < 7; ++s) {
      mx[s] = red[s][0];
      for (int w = 1; w < NW; ++w) mx[s] = fmax(mx[s], red[s][w]);
    }
    const double dtc = A.CFLScale * 2. / (mx[0] + mx[1] + mx[2]);
    double dtv = HUGE_VAL;
    const double sv = mx[3] + mx[4] + mx[5];
    if (A.parabolic && sv > 0.) dtv = A.DFLScale * 4. / sv;
    A.out[(size_t)e * 3 + 0] = dtc;
    A.out[(size_t)e * 3 + 1] = dtv;
    A.out[(size_t)e * 3 + 2] = mx[6];
  }
}

}  // namespace fx
