// volume_gl2.cuh -- line-resident Gauss-Lobatto split-form element kernel (the hot kernel of BASELINE configs[1]).
//
// Replaces, for PP_NodeType = GAUSS-LOBATTO with SPLIT_DG (file:line under /root/reference/src):
//   eos.f90:280-300 (ConsToPrim_Volume), dg/lifting/lifting_br1.t90:48-181 / lifting_br2.t90:44-311, lifting_volint.t90:26-118,
//   lifting_fillflux.t90:40-239, interpolation/applyjacobian.t90:62-112, prolongtoface.t90:270-303,
//   dg/volint.f90:211-371 (split form incl. its weak-form viscous part), equations/navierstokes/splitflux.f90:160-700,
//   flux.f90:149-241.
//
// Why a second GL kernel: the ncu profile of k_volume_gl (profiles/r01b_*) shows it bound by shared-memory wavefronts (78 % of
// L1 throughput, FP64 pipe 40 %): every FP64 operand of the two-point fluxes and of the line derivatives came from a broadcast
// LDS.64, and an LDS.64 costs two wavefronts no matter how few distinct words it reads.  Here a thread owns HALF = NP/2 nodes
// of ONE LINE and the work is organised in direction phases (xi-lines, eta-lines, zeta-lines): the own half line lives in
// registers, the other half is read once (every lane a distinct word), pairs inside the own half are formed once and applied
// to both nodes (the symmetric update of volint.f90:318-333), and D-matrix entries are immediate operands from the kernel
// parameter bank (both halves use the same entries, see gl2_ap).  Partial results change owner between phases through shared
// memory.  Shared-memory loads drop from ~400 to ~110 per DOF, FP64 instructions from ~1000 to
// ~800 per DOF.
//
// Shared-memory plane layout (NP = 8): node (i,j,k) lives at (i ^ s(j,k)) + 8*(j ^ (k&1)) + 64*k with
// s = (j>>1) | ((k&1)<<2), which makes the three access patterns (16 lanes = 16 xi-lines reading one column, 16 lanes = two
// rows of an eta- or zeta-line sweep) bank-conflict free.  Metric terms never enter shared memory: each phase reads the three
// planes of its direction for the whole line straight from L2 (they were fetched from HBM once by the lifting phase).
#pragma once
#include "volume_gl.cuh"

namespace fx {

__device__ __forceinline__ void cp_async_wait_group1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// Phase-skipping switches for timing experiments (tools/phase_timing.py): they produce wrong results and exist only in the
// development build (make dev: -DFLEXI_GPU_DEV -> libflexi_gpu_dev.so); the release library has no such branch.
#ifdef FLEXI_GPU_DEV
#define FX_DBG(A, bit) (((A).dbg & (bit)) != 0)
#else
#define FX_DBG(A, bit) false
#endif
// FX_DEV_MINB=<n> (development build) overrides the resident-CTA target of __launch_bounds__.
template <int NP>
struct GL2Basis {
  double dt[NP * NP];   // D_T(l,i)      at [l*NP+i]  (lifting derivative, lifting_volint.t90:67-83)
  double dh[NP * NP];   // D_Hat_T(l,i)  at [l*NP+i]  (weak-form viscous derivative, volint.f90:310-315)
  double dvs[NP * NP];  // DVolSurf(l,i) at [l*NP+i], times 1/8 for the Pirozzoli flux (constants folded, exact scaling)
  double lhm0, lhpN, lm0;  // L_HatMinus(0), L_HatPlus(N), L_Minus(0)
};
template <int NP>
struct VolArgs2 {
  VolArgs a;
  GL2Basis<NP> b;
};

template <int NP>
__device__ __forceinline__ int gl2_phys(int i, int j, int k) {
  if (NP == 8) return (i ^ (((j >> 1) & 3) | ((k & 1) << 2))) + 8 * (j ^ (k & 1)) + 64 * k;
  return i + NP * (j + NP * k);
}

// Shared-memory budget (NP = 8, Navier-Stokes, 4 lifted variables): 18 planes of 4 KB + the face table = 73.5 KB, so that THREE
// elements are resident per SM (12 warps; the 25-plane layout of round 1 allowed two, and the kernel is latency-bound):
//   P  7 planes  rho,u,v,w,p,T,X5        T is dead after the lifting phase: its plane becomes the hand-over plane of Ut(5)
//   W  2*NVC     G1|G2 (xi-, eta-derivatives of the lifting variables), overwritten in place by f~|g~ (4+4 planes)
//   F  3 planes  lifting flux * L_Hat of the six faces [6][NVC][NP2]; dead after the node phase: hand-over planes of Ut(2..4)
// What round 1 kept in shared memory on top of that and where it went:
//   * neighbour traces + face geometry (cp.async staging, 9 values per face node): the face thread now holds the raw trace in
//     registers across the lifting sweeps and stores only the NVC lifting fluxes; the normal vector of a boundary node is read
//     from L2 (prefetched one wave ahead) by the node's owner;
//   * h~ (4 planes): the node phase runs in zeta ownership, so the zeta derivative of h~ is formed from registers: a thread
//     adds the contributions of its own half line to ALL eight nodes of the line -- its own four stay with it, the other four go
//     to the partner through the hand-over planes (same number of DFMAs as before, 16 fewer shared-memory loads per thread);
//   * the hand-over plane of Ut(1): the density has no viscous part and is handed from sweep to sweep through the (L2-resident)
//     output array itself.
template <int NP, bool PARA, int NVC>
struct VolGL2Cfg {
  using D = Dim<NP>;
  static_assert(NP % 2 == 0, "line halves need an even number of nodes");
  static constexpr int HALF = NP / 2;
  static constexpr int NT = 2 * D::NP2;
  static constexpr int PL = D::NP3;
  static constexpr int P_PLANES = PARA ? 7 : 6;      // rho,u,v,w,p,[T],X5
  static constexpr int W_PLANES = PARA ? 2 * NVC : 0;
  static constexpr int F_FACE = PARA ? 6 * NVC * D::NP2 : 0;
  // hand-over planes of Ut(2..4) alias the face fluxes; Ut(5) goes to the T plane (Navier-Stokes) or a fourth plane (Euler)
  static constexpr int F_UT = PARA ? 3 * PL : 4 * PL;
  static constexpr int F_DOUBLES = ((F_FACE > F_UT ? F_FACE : F_UT) + 1) & ~1;
  static constexpr int NFD = PARA ? ((3 * D::NP2 + 1) & ~1) : 0;  // int table [6][NP2]
  static constexpr int O_FDST = 0;
  static constexpr int O_ETA = O_FDST + NFD;
  static constexpr int O_P = O_ETA + (PARA ? 6 : 0);
  static constexpr int O_W = O_P + P_PLANES * PL;
  static constexpr int O_F = O_W + W_PLANES * PL;
  static constexpr int O_UT5 = PARA ? O_P + 5 * PL : O_F + 3 * PL;
  static constexpr int DOUBLES = O_F + F_DOUBLES;
  static constexpr size_t BYTES = sizeof(double) * DOUBLES;
#if defined(FLEXI_GPU_DEV) && defined(FX_DEV_MINB)
  static constexpr int MINB = FX_DEV_MINB;
#else
  // resident CTAs per SM the register allocation aims at: (227 KB - 1 KB reserved per CTA) / 3 = 76458 bytes
  static constexpr int MINB = (BYTES <= 76000 && NT <= 256) ? 3 : (BYTES <= 113000) ? 2 : 1;
#endif
};

// Canonical line positions: a thread of half 0 owns positions 0..HALF-1 and meets HALF..N as partners; a thread of half 1 walks
// the line BACKWARDS (canonical position q is the actual position N-q).  The Gauss-Lobatto operators are centro-antisymmetric,
// M(N-l,N-i) = -M(l,i) for D_T, D_Hat_T and the off-diagonal DVolSurf (symmetric nodes and weights), so both halves run the
// SAME instruction stream with the same immediate coefficients and half 1 multiplies its sums by -1.  (One instruction stream
// for all warps matters: two per-half copies of this straight-line code made the kernel instruction-fetch bound, 36 % no_inst
// stalls in profiles/r01d.)  The launch wrapper verifies the antisymmetry of the tables it is given.
template <int NP>
__device__ __forceinline__ int gl2_ap(int hb, int q) {
  return hb ? (NP - 1 - q) : q;
}

// shared-memory (swizzled) address / natural node index of position p on the line of direction dir through sheet position (r0,r1):
// xi-lines (j,k) = (r0,r1), eta-lines (i,k) = (r0,r1), zeta-lines (i,j) = (r0,r1).  dir is a run-time value: the three direction
// sweeps share ONE copy of the (long, straight-line) sweep code, which keeps the instruction footprint of the kernel small.
template <int NP>
__device__ __forceinline__ int gl2_sa(int dir, int p, int r0, int r1) {
  const int i = (dir == 0) ? p : r0, j = (dir == 0) ? r0 : (dir == 1) ? p : r1, k = (dir == 2) ? p : r1;
  return gl2_phys<NP>(i, j, k);
}
template <int NP>
__device__ __forceinline__ int gl2_gn(int dir, int p, int r0, int r1) {
  const int i = (dir == 0) ? p : r0, j = (dir == 0) ? r0 : (dir == 1) ? p : r1, k = (dir == 2) ? p : r1;
  return i + NP * (j + NP * k);
}

// One direction sweep of the advective two-point fluxes (+ the weak-form derivative of the viscous flux of that direction).
//   sa[q] / gn[q]  shared-memory address / natural node index of canonical line position q
//   vp[v]          plane of the contravariant viscous flux of this direction, variable v+1
//   ut[c][v]       result for the own nodes c (HALF of them), WITHOUT the half sign (the caller applies it)
template <int NP, int STYPE, bool PARA>
__device__ __forceinline__ void gl2_pair_sweep(const GL2Basis<NP>& B, const double* __restrict__ sP, const double* const (&vp)[4],
                                               const double* __restrict__ gMd, const int (&sa)[NP], const int (&gn)[NP],
                                               double (&ut)[NP / 2][5], const double kappa) {
  constexpr int HALF = NP / 2, PL = NP * NP * NP;
  constexpr bool PI = (STYPE == FLEXI_SPLITDG_PI);
  constexpr int PX = PARA ? 6 : 5;
  // metric terms of this direction along the whole line, straight from L2
  double Mo[HALF][3], Mp[HALF][3];
#pragma unroll
  for (int c = 0; c < HALF; ++c)
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      Mo[c][m] = __ldg(gMd + m * PL + gn[c]);
      Mp[c][m] = __ldg(gMd + m * PL + gn[HALF + c]);
    }
#pragma unroll
  for (int c = 0; c < HALF; ++c)
#pragma unroll
    for (int v = 0; v < 5; ++v) ut[c][v] = 0.;
  if (PARA) {  // volint.f90:310-315: Ut += D_Hat_T(l,i) * f~visc(l)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      double L[NP];
#pragma unroll
      for (int p = 0; p < NP; ++p) L[p] = vp[v][sa[p]];
#pragma unroll
      for (int c = 0; c < HALF; ++c) {  // two partial sums per node: twice the independent FP64 chains
        double s0 = B.dh[c] * L[0], s1 = B.dh[HALF * NP + c] * L[HALF];
#pragma unroll
        for (int p = 1; p < HALF; ++p) {
          s0 = s0 + B.dh[p * NP + c] * L[p];
          s1 = s1 + B.dh[(HALF + p) * NP + c] * L[HALF + p];
        }
        ut[c][v + 1] = s0 + s1;
      }
    }
  }
  // own half line
  double a[HALF][6];
#pragma unroll
  for (int c = 0; c < HALF; ++c) {
    const int s = sa[c];
#pragma unroll
    for (int v = 0; v < 5; ++v) a[c][v] = sP[v * PL + s];
    a[c][5] = sP[PX * PL + s];
  }
  // pairs inside the own half: formed once, applied to both nodes
#pragma unroll
  for (int c = 0; c < HALF; ++c)
#pragma unroll
    for (int c2 = c + 1; c2 < HALF; ++c2) {
      const double d12 = B.dvs[c2 * NP + c];  // DVolSurf(l=c2, c) -> node c
      const double d21 = B.dvs[c * NP + c2];  // DVolSurf(l=c, c2) -> node c2
      if (PI) {
        const double su = a[c][1] + a[c2][1], sv = a[c][2] + a[c2][2], sw = a[c][3] + a[c2][3], sp4 = a[c][4] + a[c2][4],
                     sH = a[c][5] + a[c2][5];
        const double m0 = Mo[c][0] + Mo[c2][0], mm1 = Mo[c][1] + Mo[c2][1], mm2 = Mo[c][2] + Mo[c2][2];
        const double rq = (a[c][0] + a[c2][0]) * (m0 * su + mm1 * sv + mm2 * sw);
        {
          const double w = d12 * rq, dp = d12 * sp4;
          ut[c][0] += 2. * w;
          ut[c][1] += w * su;
          ut[c][2] += w * sv;
          ut[c][3] += w * sw;
          ut[c][4] += w * sH;
          ut[c][1] += dp * m0;
          ut[c][2] += dp * mm1;
          ut[c][3] += dp * mm2;
        }
        {
          const double w = d21 * rq, dp = d21 * sp4;
          ut[c2][0] += 2. * w;
          ut[c2][1] += w * su;
          ut[c2][2] += w * sv;
          ut[c2][3] += w * sw;
          ut[c2][4] += w * sH;
          ut[c2][1] += dp * m0;
          ut[c2][2] += dp * mm1;
          ut[c2][3] += dp * mm2;
        }
      } else {
        double Fl[5];
        split_pair_generic<STYPE>(a[c], Mo[c], a[c2][0], a[c2][1], a[c2][2], a[c2][3], a[c2][4], a[c2][5], Mo[c2][0], Mo[c2][1], Mo[c2][2], Fl, kappa);
#pragma unroll
        for (int v = 0; v < 5; ++v) {
          ut[c][v] = ut[c][v] + d12 * Fl[v];
          ut[c2][v] = ut[c2][v] + d21 * Fl[v];
        }
      }
    }
  // pairs with the other half: partner read once, paired with all own nodes
#pragma unroll
  for (int c2 = 0; c2 < HALF; ++c2) {
    const int l = HALF + c2;
    const int s = sa[l];
    const double br = sP[s], bu = sP[PL + s], bv = sP[2 * PL + s], bw = sP[3 * PL + s], bp = sP[4 * PL + s], bx = sP[PX * PL + s];
#pragma unroll
    for (int c = 0; c < HALF; ++c) {
      const double dl = B.dvs[l * NP + c];
      if (PI) {
        const double su = a[c][1] + bu, sv = a[c][2] + bv, sw = a[c][3] + bw, sp4 = a[c][4] + bp, sH = a[c][5] + bx;
        const double m0 = Mo[c][0] + Mp[c2][0], mm1 = Mo[c][1] + Mp[c2][1], mm2 = Mo[c][2] + Mp[c2][2];
        const double rq = (a[c][0] + br) * (m0 * su + mm1 * sv + mm2 * sw);
        const double w = dl * rq, dp = dl * sp4;
        ut[c][0] += 2. * w;
        ut[c][1] += w * su;
        ut[c][2] += w * sv;
        ut[c][3] += w * sw;
        ut[c][4] += w * sH;
        ut[c][1] += dp * m0;
        ut[c][2] += dp * mm1;
        ut[c][3] += dp * mm2;
      } else {
        double Fl[5];
        split_pair_generic<STYPE>(a[c], Mo[c], br, bu, bv, bw, bp, bx, Mp[c2][0], Mp[c2][1], Mp[c2][2], Fl, kappa);
#pragma unroll
        for (int v = 0; v < 5; ++v) ut[c][v] = ut[c][v] + dl * Fl[v];
      }
    }
  }
}

// Template parameters: NVL = nVarLifting of the gradient arrays (their layout), NVC = number of lifting variables actually
// lifted: NVL, or 4 = (v1,v2,v3,T) on the time-stepping path, where nothing reads the density gradient (EvalDiffFlux3D does
// not use it and no gradient array is handed out).
template <int NP, int STYPE, bool PARA, int NVL, int NVC, int LIFT>
__global__ void __launch_bounds__(VolGL2Cfg<NP, PARA, NVC>::NT, VolGL2Cfg<NP, PARA, NVC>::MINB)
    k_volume_gl2(const __grid_constant__ VolArgs2<NP> AA) {
  using D = Dim<NP>;
  using C = VolGL2Cfg<NP, PARA, NVC>;
  constexpr int HALF = C::HALF, NT = C::NT, PL = C::PL, N = NP - 1;
  constexpr bool PI = (STYPE == FLEXI_SPLITDG_PI);
  constexpr int PX = PARA ? 6 : 5;
  constexpr int VOFF = NVL - NVC;  // slot of the first lifted variable inside the nVarLifting layout
  const VolArgs& A = AA.a;
  const GL2Basis<NP>& B = AA.b;
  extern __shared__ __align__(16) double sm[];
  int* sFD = reinterpret_cast<int*>(sm + C::O_FDST);
  double* sEta = sm + C::O_ETA;
  double* sP = sm + C::O_P;
  double* sW = sm + C::O_W;   // G1 | G2 lifting planes, later f~ | g~
  double* sF = sm + C::O_F;   // lifting fluxes of the six faces, later the hand-over planes of Ut(2..4)
  // hand-over plane of Ut(v+1), v = 1..4 (the density goes through the output array)
  auto ut_plane = [&](int v) -> double* { return v < 4 ? sF + (v - 1) * PL : sm + C::O_UT5; };
  const int tid = threadIdx.x;
  if (FX_DBG(A, 8)) return;  // timing experiment: the cost of scheduling one CTA per element
  const int e = A.elems ? A.elems[blockIdx.x] : blockIdx.x;
  const int hb = tid / D::NP2, r = tid % D::NP2;  // half index (warp-uniform for NP = 8) and position inside the NP x NP sheet
  const int r0 = r % NP, r1 = r / NP;
  const double sg = hb ? -1. : 1.;
  const Eqn& q = A.q;
  const double* gU = A.U + (size_t)e * 5 * D::NP3;
  const double* gM = A.m.Met + (size_t)e * 9 * D::NP3;
  double* gUt = A.Ut + (size_t)e * 5 * D::NP3;
  constexpr int FACE_IT = (6 * D::NP2 + NT - 1) / NT;

  // ---- phase A (zeta ownership): own conservative state on its way first, then the neighbour traces of the face nodes this
  //      thread converts (registers), then ConsToPrim of the own nodes ----
  double Uc[HALF][5];
#pragma unroll
  for (int c = 0; c < HALF; ++c)
#pragma unroll
    for (int v = 0; v < 5; ++v) Uc[c][v] = __ldg(gU + v * D::NP3 + gl2_gn<NP>(2, gl2_ap<NP>(hb, c), r0, r1));
  // metric terms + Jacobian of the first own node of the lifting phase: a first touch of HBM, so it starts here
  double Mn[9], sJn;
  if (PARA) {
    const int n = gl2_gn<NP>(2, gl2_ap<NP>(hb, 0), r0, r1);
#pragma unroll
    for (int m = 0; m < 9; ++m) Mn[m] = __ldg(gM + m * D::NP3 + n);
    sJn = __ldg(A.m.sJ + (size_t)e * D::NP3 + n);
  }
  SideRef nsr{-1, 0};  // side of face tid/21 of that later element (its traces + geometry are prefetched after S1)
  if (A.prefetchAhead > 0 && (int)blockIdx.x + A.prefetchAhead < A.nList && !FX_DBG(A, 32)) {
    // pull the element that starts a wave of CTAs later (U, metrics, Jacobian: 60 KB) from HBM into L2 now
    const int en = A.elems ? A.elems[blockIdx.x + A.prefetchAhead] : (int)blockIdx.x + A.prefetchAhead;
    if (PARA && tid < 126) nsr = A.m.e2s[en * 6 + tid / 21];
    const char* pu = reinterpret_cast<const char*>(A.U + (size_t)en * 5 * D::NP3);
    const char* pm = reinterpret_cast<const char*>(A.m.Met + (size_t)en * 9 * D::NP3);
    const char* pj = reinterpret_cast<const char*>(A.m.sJ + (size_t)en * D::NP3);
    constexpr int LU = 5 * D::NP3 * 8 / 128, LM = 9 * D::NP3 * 8 / 128, LJ = D::NP3 * 8 / 128;
    for (int l = tid; l < LU + LM + LJ; l += NT) {
      const char* p = l < LU ? pu + (size_t)l * 128 : l < LU + LM ? pm + (size_t)(l - LU) * 128 : pj + (size_t)(l - LU - LM) * 128;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    }
  }
  int faceSlot[FACE_IT];
  bool faceLh[FACE_IT];  // plus side: L_HatPlus(N) (BR1)
  double faceSgn[FACE_IT];
  double fU[FACE_IT][5], fSurf[FACE_IT];  // raw neighbour trace + SurfElem of the face node, held until the lifting sweeps are done
  if (PARA && !FX_DBG(A, 16)) {
#pragma unroll
    for (int it = 0; it < FACE_IT; ++it) {
      const int t = tid + it * NT;
      faceSlot[it] = -1;
      faceLh[it] = false;
      faceSgn[it] = 0.;
      fSurf[it] = 0.;
#pragma unroll
      for (int v = 0; v < 5; ++v) fU[it][v] = 1.;
      if (t < 6 * D::NP2) {
        const int ls = t / D::NP2 + 1, pq = t % D::NP2;
        const SideRef sr = A.m.e2s[e * 6 + ls - 1];
        const int ab = A.m.s2v2[(sr.flip * 6 + ls - 1) * D::NP2 + pq];
        const double* nbTrace = ((sr.flip == 0) ? A.U_slave : A.U_master) + (size_t)sr.side * 5 * D::NP2 + pq;
#pragma unroll
        for (int v = 0; v < 5; ++v) fU[it][v] = __ldg(nbTrace + v * D::NP2);
        fSurf[it] = __ldg(A.m.Geo + (size_t)sr.side * 10 * D::NP2 + 9 * D::NP2 + pq);
        // destination of this face node's traces: side*NP2 + pq, bit 31 = slave arrays, bit 30 = boundary-condition side
        sFD[(ls - 1) * D::NP2 + ab] =
            (sr.side * D::NP2 + pq) | (sr.flip == 0 ? 0 : (int)0x80000000) | (sr.side < A.m.nBCSides ? 0x40000000 : 0);
        if (LIFT == 2 && pq == 0) {  // lifting_br2.t90:277-288: etaBR2_wall on no-slip wall sides
          double eta = q.etaBR2;
          if (sr.side < A.m.nBCSides) {
            const int bt = A.m.bcType[sr.side];
            if (bt == 3 || bt == 4) eta = q.etaBR2_wall;
          }
          sEta[ls - 1] = eta;
        }
        // volume node of this face node (prolongtoface.t90:270-303), in the swizzled plane layout
        const int fa = ab % NP, fb = ab / NP, fd = is_minus_side(ls) ? 0 : N;
        int vn;
        if (ls == XI_MINUS || ls == XI_PLUS)
          vn = gl2_phys<NP>(fd, fa, fb);
        else if (ls == ETA_MINUS || ls == ETA_PLUS)
          vn = gl2_phys<NP>(fa, fd, fb);
        else
          vn = gl2_phys<NP>(fa, fb, fd);
        faceSlot[it] = ((ls - 1) * NVC * D::NP2 + ab) | (vn << 16);
        faceSgn[it] = (sr.flip == 0) ? 1. : -1.;
        faceLh[it] = !is_minus_side(ls);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < HALF; ++c) {
    double Pn[6];
    const double* const Ucc = Uc[c];
    const double sRho = cons_to_prim_srho(q, Ucc, Pn);
    const int s = gl2_sa<NP>(2, gl2_ap<NP>(hb, c), r0, r1);
#pragma unroll
    for (int v = 0; v < 4; ++v) sP[v * PL + s] = Pn[v];
    sP[4 * PL + s] = PI ? 4. * Pn[PRES] : Pn[PRES];
    if (PARA) sP[5 * PL + s] = Pn[TEMP];
    // 6th two-point quantity: H = (rhoE+p)/rho (PI), E = rhoE/rho (KG), rhoE (SD, DU)
    sP[PX * PL + s] = PI ? (Ucc[ENER] + Pn[PRES]) * sRho : (STYPE == FLEXI_SPLITDG_KG) ? Ucc[ENER] * sRho : Ucc[ENER];
  }
  __syncthreads();  // S1: primitive planes complete

  if (PARA && nsr.side >= 0) {
    // L2 prefetch of the six faces of the element one wave later (neighbour trace 20 lines, normal vector 12, SurfElem 4 per
    // face; 21 threads per face): without it the start of every element waits on the chain ElemToSide -> S2V2 -> HBM
    const char* tr = reinterpret_cast<const char*>(((nsr.flip == 0) ? A.U_slave : A.U_master) + (size_t)nsr.side * 5 * D::NP2);
    const char* ge = reinterpret_cast<const char*>(A.m.Geo + (size_t)nsr.side * 10 * D::NP2);
    constexpr int LT = 5 * D::NP2 * 8 / 128, LG = 3 * D::NP2 * 8 / 128, LS = D::NP2 * 8 / 128;
    for (int l = tid % 21; l < LT + LG + LS; l += 21) {
      const char* p = l < LT ? tr + (size_t)l * 128 : l < LT + LG ? ge + (size_t)(l - LT) * 128 : ge + (size_t)9 * D::NP2 * 8 + (size_t)(l - LT - LG) * 128;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    }
  }
  if (PARA) {
    // ---- reference-space derivatives of the lifting variables (lifting_volint.t90:67-83), one line sweep per direction:
    //      xi and eta results change owner through the planes G1, G2; the zeta result stays with this thread ----
    double gzt[HALF][NVC];
#pragma unroll 1
    for (int dir = FX_DBG(A, 2) ? 3 : 0; dir < 3; ++dir) {
      int sa[NP];
#pragma unroll
      for (int p = 0; p < NP; ++p) sa[p] = gl2_sa<NP>(dir, gl2_ap<NP>(hb, p), r0, r1);
#pragma unroll
      for (int v = 0; v < NVC; ++v) {
        const double* pl = sP + lift_prim(NVC, v) * PL;
        double L[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) L[p] = pl[sa[p]];
#pragma unroll
        for (int c = 0; c < HALF; ++c) {  // two partial sums per node: twice the independent FP64 chains
          double s0 = B.dt[c] * L[0], s1 = B.dt[HALF * NP + c] * L[HALF];
#pragma unroll
          for (int p = 1; p < HALF; ++p) {
            s0 = s0 + B.dt[p * NP + c] * L[p];
            s1 = s1 + B.dt[(HALF + p) * NP + c] * L[HALF + p];
          }
          gzt[c][v] = sg * (s0 + s1);
        }
      }
      if (dir < 2) {
#pragma unroll
        for (int c = 0; c < HALF; ++c)
#pragma unroll
          for (int v = 0; v < NVC; ++v) sW[(dir * NVC + v) * PL + sa[c]] = gzt[c][v];
      }
    }
    // ---- lifting flux of the six faces (lifting_fillflux.t90:81-107, strong form):
    //      0.5*(UPrim_slave-UPrim_master)*SurfElem per lifting variable, already times L_Hat of the face.  The neighbour trace
    //      had the line sweeps above to arrive. ----
#pragma unroll
    for (int it = 0; it < FACE_IT; ++it) {
      if (faceSlot[it] >= 0 && !FX_DBG(A, 16)) {
        double* o = sF + (faceSlot[it] & 0xffff);
        const int nF = faceSlot[it] >> 16;
        double Pb[6];
        const double lh = (LIFT == 1 && faceLh[it]) ? B.lhpN : B.lhm0;
        cons_to_prim(q, fU[it], Pb);
#pragma unroll
        for (int v = 0; v < NVC; ++v) {
          const int pv = lift_prim(NVC, v);
          o[v * D::NP2] = (0.5 * (faceSgn[it] * (Pb[pv] - sP[pv * PL + nF])) * fSurf[it]) * lh;
        }
      }
    }
    __syncthreads();  // S2: G1, G2 and the face lifting fluxes complete

    // ---- zeta ownership: metric combination, surface part, Jacobian, traces for the side kernel, viscous flux ----
    double hvc[HALF][4];  // h~ of the own nodes: its zeta derivative is formed from registers after S3
#pragma unroll
    for (int c = 0; c < HALF; ++c)
#pragma unroll
      for (int v = 0; v < 4; ++v) hvc[c][v] = 0.;
    if (!FX_DBG(A, 4)) {
      constexpr int LV = (NVC == 5) ? 1 : 0;
      const int i = r0, j = r1;
#pragma unroll
      for (int c = 0; c < HALF; ++c) {
        const int k = gl2_ap<NP>(hb, c);
        const int s = gl2_sa<NP>(2, k, r0, r1);
        const int n = i + NP * (j + NP * k);
        double Mc[9];
#pragma unroll
        for (int m = 0; m < 9; ++m) Mc[m] = Mn[m];
        const double sJc = sJn;
        if (c + 1 < HALF) {  // metric terms + Jacobian straight from HBM, one node ahead of the arithmetic
          const int n1 = gl2_gn<NP>(2, gl2_ap<NP>(hb, c + 1), r0, r1);
#pragma unroll
          for (int m = 0; m < 9; ++m) Mn[m] = __ldg(gM + m * D::NP3 + n1);
          sJn = __ldg(A.m.sJ + (size_t)e * D::NP3 + n1);
        }
        // ---- boundary nodes.  One (partially active) block per sheet direction instead of one per local side: the lanes of a
        //      warp that sit on xi- and xi+ run the same code with their own side index.  The zeta faces belong to the line
        //      end c == 0 of every thread (k = 0 for half 0, k = N for half 1), a warp-uniform condition. ----
        const bool onI = (i == 0) | (i == N), onJ = (j == 0) | (j == N), onK = (c == 0);
        const int lsI = (i == 0) ? XI_MINUS : XI_PLUS, lsJ = (j == 0) ? ETA_MINUS : ETA_PLUS, lsK = hb ? ZETA_PLUS : ZETA_MINUS;
        const int abI = j + NP * k, abJ = i + NP * k, abK = i + NP * j;
        // trace destination + normal vector of the (up to three) faces of this node; the vector comes from L2 (prefetched)
        int fdI = 0, fdJ = 0, fdK = 0;
        double nI[3] = {0., 0., 0.}, nJ[3] = {0., 0., 0.}, nK[3] = {0., 0., 0.};
        auto face_geo = [&](int ls, int ab, int& fd, double (&nv)[3]) {
          fd = sFD[(ls - 1) * D::NP2 + ab];
          const int fl = fd & 0x3fffffff;
          const double* g = A.m.Geo + (size_t)(fl / D::NP2) * 10 * D::NP2 + fl % D::NP2;
          nv[0] = __ldg(g);
          nv[1] = __ldg(g + D::NP2);
          nv[2] = __ldg(g + 2 * D::NP2);
        };
        if (onI) face_geo(lsI, abI, fdI, nI);
        if (onJ) face_geo(lsJ, abJ, fdJ, nJ);
        if (onK) face_geo(lsK, abK, fdK, nK);
        double gx[NVC], gy[NVC], gz[NVC];
#pragma unroll
        for (int v = 0; v < NVC; ++v) {
          const double gxi = sW[v * PL + s], geta = sW[(NVC + v) * PL + s], gzeta = gzt[c][v];
          gx[v] = Mc[0] * gxi + Mc[3] * geta + Mc[6] * gzeta;
          gy[v] = Mc[1] * gxi + Mc[4] * geta + Mc[7] * gzeta;
          gz[v] = Mc[2] * gxi + Mc[5] * geta + Mc[8] * gzeta;
        }
        double gxo[NVC], gyo[NVC], gzo[NVC];  // BR2: gradients without the surface part (what the faces are prolonged from)
        // lifting, surface part (BR1 lifting_br1.t90:195-292; BR2 lifting_br2.t90:194-311)
        auto surf_in = [&](int ls, int ab, const double (&nv)[3]) {
          const double* o = sF + (ls - 1) * NVC * D::NP2 + ab;
#pragma unroll
          for (int v = 0; v < NVC; ++v) {
            const double f = o[v * D::NP2];  // Flux * L_Hat
            if (LIFT == 1) {
              gx[v] = gx[v] + f * nv[0];
              gy[v] = gy[v] + f * nv[1];
              gz[v] = gz[v] + f * nv[2];
            } else {
              gx[v] = gx[v] + sJc * (f * nv[0]);
              gy[v] = gy[v] + sJc * (f * nv[1]);
              gz[v] = gz[v] + sJc * (f * nv[2]);
            }
          }
        };
        if (LIFT == 1) {  // BR1: surface part, then ApplyJacobianLifting
          if (onI) surf_in(lsI, abI, nI);
          if (onJ) surf_in(lsJ, abJ, nJ);
          if (onK) surf_in(lsK, abK, nK);
#pragma unroll
          for (int v = 0; v < NVC; ++v) {
            gx[v] = gx[v] * sJc;
            gy[v] = gy[v] * sJc;
            gz[v] = gz[v] * sJc;
          }
        } else {  // BR2: Jacobian first, then sJ*Flux*L_HatMinus(0) on the boundary node
#pragma unroll
          for (int v = 0; v < NVC; ++v) {
            gx[v] = gx[v] * sJc;
            gy[v] = gy[v] * sJc;
            gz[v] = gz[v] * sJc;
            gxo[v] = gx[v];
            gyo[v] = gy[v];
            gzo[v] = gz[v];
          }
          if (onI) surf_in(lsI, abI, nI);
          if (onJ) surf_in(lsJ, abJ, nJ);
          if (onK) surf_in(lsK, abK, nK);
        }
        if (A.gvol[0]) {
#pragma unroll
          for (int v = 0; v < NVC; ++v) {
            A.gvol[0][((size_t)e * NVL + v + VOFF) * D::NP3 + n] = gx[v];
            A.gvol[1][((size_t)e * NVL + v + VOFF) * D::NP3 + n] = gy[v];
            A.gvol[2][((size_t)e * NVL + v + VOFF) * D::NP3 + n] = gz[v];
          }
        }
        // ---- viscous flux of this node (flux.f90:149-241) ----
        const double v1 = sP[VEL1 * PL + s], v2 = sP[VEL2 * PL + s], v3 = sP[VEL3 * PL + s];
        double fv[4], gv[4], hv[4];
        eval_diff_flux3d(q, v1, v2, v3, gx + LV, gy + LV, gz + LV, fv, gv, hv);
        // ---- what the side kernel needs from this element.  ViscousFlux_Side (riemann.f90:448-498) forms
        //      0.5*n.(F_v(U_L,gradU_L) + F_v(U_R,gradU_R)); each half depends on one element only, and on LGL nodes the face
        //      state and (BR1) the face gradient are copies of the boundary node's (prolongtoface.t90:270-303), so the element
        //      writes n.F_v of its own side (4 doubles per face node) instead of the 3*nVarLifting gradient traces.  The
        //      gradient traces themselves are still written for boundary-condition sides (GetBoundaryFlux needs them) and, when
        //      the host asked for a full DG operator evaluation, for every side. ----
        auto face_out = [&](int ls, int ab, int fd, const double (&nv)[3]) {
          const bool master = fd >= 0;
          const int fl = fd & 0x3fffffff, side = fl / D::NP2, pq = fl % D::NP2;
          double gfx[NVC], gfy[NVC], gfz[NVC];
          if (LIFT == 2) {  // BR2 face gradient: volume-only value + eta*L_Minus(0)*sJ*Flux*L_HatMinus(0) of the side itself
            const double* of = sF + (ls - 1) * NVC * D::NP2 + ab;
            const double w = sEta[ls - 1] * B.lm0;
#pragma unroll
            for (int v = 0; v < NVC; ++v) {
              const double f = of[v * D::NP2];
              gfx[v] = gxo[v] + w * (sJc * (f * nv[0]));
              gfy[v] = gyo[v] + w * (sJc * (f * nv[1]));
              gfz[v] = gzo[v] + w * (sJc * (f * nv[2]));
            }
          }
          if (A.fullGrad || (fd & 0x40000000)) {
            double* const d0 = master ? A.gm[0] : A.gs[0];
            double* const d1 = master ? A.gm[1] : A.gs[1];
            double* const d2 = master ? A.gm[2] : A.gs[2];
            const size_t o0 = (size_t)side * NVL * D::NP2 + pq;
#pragma unroll
            for (int v = 0; v < NVC; ++v) {
              const size_t o = o0 + (size_t)(v + VOFF) * D::NP2;
              d0[o] = (LIFT == 2) ? gfx[v] : gx[v];
              d1[o] = (LIFT == 2) ? gfy[v] : gy[v];
              d2[o] = (LIFT == 2) ? gfz[v] : gz[v];
            }
          }
          double* const dv = (master ? A.fvm : A.fvs) + (size_t)side * 4 * D::NP2 + pq;
          if (LIFT == 2) {
            double ff[4], gf[4], hf[4];
            eval_diff_flux3d(q, v1, v2, v3, gfx + LV, gfy + LV, gfz + LV, ff, gf, hf);
#pragma unroll
            for (int v = 0; v < 4; ++v) dv[v * D::NP2] = nv[0] * ff[v] + nv[1] * gf[v] + nv[2] * hf[v];
          } else {
#pragma unroll
            for (int v = 0; v < 4; ++v) dv[v * D::NP2] = nv[0] * fv[v] + nv[1] * gv[v] + nv[2] * hv[v];
          }
        };
        if (onI) face_out(lsI, abI, fdI, nI);
        if (onJ) face_out(lsJ, abJ, fdJ, nJ);
        if (onK) face_out(lsK, abK, fdK, nK);
        // ---- contravariant components (volint.f90:378-422): f~, g~ over the node's own G1/G2 slots (only this thread reads
        //      them), h~ stays in registers ----
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          sW[(0 + v) * PL + s] = fv[v] * Mc[0] + gv[v] * Mc[1] + hv[v] * Mc[2];
          sW[(4 + v) * PL + s] = fv[v] * Mc[3] + gv[v] * Mc[4] + hv[v] * Mc[5];
          hvc[c][v] = fv[v] * Mc[6] + gv[v] * Mc[7] + hv[v] * Mc[8];
        }
      }
    }
    __syncthreads();  // S3: f~, g~ complete; the face fluxes and the T plane are free from here on

    // ---- h~ of the own nodes goes to the (now free) hand-over planes: the zeta sweep, which comes first, reads it from there
    //      like the other two sweeps read f~ and g~ ----
#pragma unroll
    for (int c = 0; c < HALF; ++c) {
      const int s = gl2_sa<NP>(2, gl2_ap<NP>(hb, c), r0, r1);
#pragma unroll
      for (int v = 0; v < 4; ++v) ut_plane(v + 1)[s] = hvc[c][v];
    }
    __syncthreads();  // S3b
  }

  // ---- three direction sweeps (zeta, xi, eta) of ONE copy of the sweep code.  The partial sums change owner twice: after the
  //      zeta sweep through the output array itself (it stays in L2; the hand-over planes still hold h~ while partners read it),
  //      after the xi sweep through the hand-over planes (density: output array again); the eta sweep sends the raw volume
  //      integral to HBM ----
#pragma unroll 1
  for (int it = FX_DBG(A, 1) ? 3 : 0; it < 3; ++it) {
    const int dir = (it == 0) ? 2 : it - 1;
    int sa[NP], gn[NP];  // indexed by canonical position
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      sa[p] = gl2_sa<NP>(dir, gl2_ap<NP>(hb, p), r0, r1);
      gn[p] = gl2_gn<NP>(dir, gl2_ap<NP>(hb, p), r0, r1);
    }
    const double* vp[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) vp[v] = (dir < 2) ? sW + (dir * 4 + v) * PL : ut_plane(v + 1);
    double ut[HALF][5];
    gl2_pair_sweep<NP, STYPE, PARA>(B, sP, vp, gM + dir * 3 * D::NP3, sa, gn, ut, A.q.kappa);
#pragma unroll
    for (int c = 0; c < HALF; ++c) {
      if (it == 0) {
#pragma unroll
        for (int v = 0; v < 5; ++v) gUt[v * D::NP3 + gn[c]] = sg * ut[c][v];
      } else if (it == 1) {
        gUt[gn[c]] = __ldcg(gUt + gn[c]) + sg * ut[c][0];
#pragma unroll
        for (int v = 1; v < 5; ++v) ut_plane(v)[sa[c]] = __ldcg(gUt + v * D::NP3 + gn[c]) + sg * ut[c][v];
      } else {
        gUt[gn[c]] = __ldcg(gUt + gn[c]) + sg * ut[c][0];
#pragma unroll
        for (int v = 1; v < 5; ++v) gUt[v * D::NP3 + gn[c]] = ut_plane(v)[sa[c]] + sg * ut[c][v];
      }
    }
    if (it < 2) __syncthreads();  // S4, S5
  }
}

}  // namespace fx
