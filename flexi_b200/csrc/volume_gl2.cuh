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
// x * (+-1) as an integer operation on the sign bit (sb = 0 or 0x80000000): same bits as the product, no FP64 issue slot
__device__ __forceinline__ double sgn_apply(double x, int sb) { return __hiloint2double(__double2hiint(x) ^ sb, __double2loint(x)); }
// store under a predicate, without a branch
__device__ __forceinline__ void st_global_if(bool on, double* p, double v) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q st.global.f64 [%0], %1;\n\t}" ::"l"(p), "d"(v), "r"((int)on) : "memory");
}

// Phase-skipping switches for timing experiments (tools/phase_timing.py): they produce wrong results and exist only in the
// development build (make dev: -DFLEXI_GPU_DEV -> libflexi_gpu_dev.so); the release library has no such branch.
#if defined(FLEXI_GPU_DEV) && defined(FX_DEV_DBG)
#define FX_DBG(A, bit) (((A).dbg & (bit)) != 0)
#else
#define FX_DBG(A, bit) false
#endif
// FX_STAMP (development build): per-warp clock64() deltas of the kernel phases, summed over all warps (tools/phase_stamps.py).
#if defined(FLEXI_GPU_DEV) && defined(FX_DEV_STAMPS)
static __device__ unsigned long long g_gl2_stamps[32];
#define FX_STAMP_INIT long long fx_t0_ = clock64();
#define FX_STAMP(i)                                                                                          \
  {                                                                                                          \
    const long long t_ = clock64();                                                                          \
    if ((threadIdx.x & 31) == 0) atomicAdd(&g_gl2_stamps[i], (unsigned long long)(t_ - fx_t0_));             \
    fx_t0_ = t_;                                                                                             \
  }
#else
#define FX_STAMP_INIT
#define FX_STAMP(i)
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

// Shared-memory budget (NP = 8, Navier-Stokes, 4 lifted variables): 28 planes of 4 KB = 112 KB, two elements resident per SM.
// Every operand of the node phase and of the three two-point sweeps comes from shared memory; the only global loads of the
// kernel are issued in phase A (round 1 re-read the metric terms from L2 in the node phase and in every sweep: seven exposed L2
// round trips per element with two warps per scheduler to hide them):
//   R0 1 plane   ElemToSide/S2V2 face table (int [6][NP2]) + etaBR2 until the node phase is done, then hand-over plane of Ut(1)
//   P  7 planes  rho,u,v,w,p,T,X5        T is dead after the lifting phase: its plane becomes the hand-over plane of Ut(5)
//   W  2*NVC     G1|G2 (xi-, eta-derivatives of the lifting variables), overwritten in place by f~|g~ (4+4 planes)
//   F  3 planes  lifting flux * L_Hat of the six faces [6][NVC][NP2]; dead after the node phase: hand-over planes of Ut(2..4)
//   M  9 planes  metric terms f~,g~,h~Tilde in the swizzled layout, cp.async (LDGSTS) issued first thing in phase A
// What round 1 kept in shared memory on top of that and where it went:
//   * neighbour traces + face geometry (cp.async staging, 9 values per face node): the face thread holds the raw trace in
//     registers across the lifting sweeps and stores only the NVC lifting fluxes; the normal vectors of a thread's boundary
//     nodes are requested right after S1 and have the lifting sweeps to arrive;
//   * h~ (4 planes): the node phase runs in zeta ownership, so the zeta derivative of h~ is formed from registers: a thread
//     adds the contributions of its own half line to ALL eight nodes of the line -- its own four stay with it, the other four go
//     to the partner through the hand-over planes (same number of DFMAs as before, 16 fewer shared-memory loads per thread).
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
  static constexpr int NFD = PARA ? ((3 * D::NP2 + 1) & ~1) : 0;  // int table [6][NP2], in doubles
#if defined(FLEXI_GPU_DEV) && defined(FX_DEV_MGLOBAL)
  static constexpr bool WANT_MSM = false;  // timing experiment: 18-plane layout, metric terms from L2, Ut(1) handed over through L2
#else
  static constexpr bool WANT_MSM = true;
#endif
  static constexpr int R0_DOUBLES = (((WANT_MSM && PL > NFD + 6) ? PL : NFD + 6) + 1) & ~1;
  static constexpr int O_FDST = 0;
  static constexpr int O_ETA = O_FDST + NFD;
  static constexpr int O_UT1 = 0;
  static constexpr int O_P = R0_DOUBLES;
  static constexpr int O_W = O_P + P_PLANES * PL;
  static constexpr int O_F = O_W + W_PLANES * PL;
  static constexpr int O_UT5 = PARA ? O_P + 5 * PL : O_F + 3 * PL;
  static constexpr int O_M = O_F + F_DOUBLES;
  // metric planes in shared memory unless the element does not fit (NP = 10 with all five lifting variables): then they are
  // read from global memory as in round 1 (natural layout == plane layout for NP != 8)
  static constexpr bool MSM = WANT_MSM && sizeof(double) * (O_M + 9 * PL) <= 232448;
  static constexpr bool UT1S = R0_DOUBLES >= PL;  // hand-over plane of Ut(1) in shared memory (else through the output array in L2)
  static constexpr int DOUBLES = O_M + (MSM ? 9 * PL : 0);
  static constexpr size_t BYTES = sizeof(double) * DOUBLES;
#if defined(FLEXI_GPU_DEV) && defined(FX_DEV_MINB)
  static constexpr int MINB = FX_DEV_MINB;
#else
  // resident CTAs per SM the register allocation aims at: 228 KB per SM, 1 KB reserved per CTA
  static constexpr int MINB = (BYTES <= 76000 && NT <= 256) ? 3 : (BYTES <= 115712) ? 2 : 1;
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

// Swizzled address of CANONICAL line position p (actual position gl2_ap(hb, p)) of the line of direction dir through (r0, r1).
// For NP = 8 the swizzle is XOR-linear in the line position: address = base(dir; hb, r0, r1) ^ off(dir, p) with a per-thread base
// and compile-time offsets -- one LOP3 per address instead of the ~6 integer instructions of gl2_phys (the lifting sweeps spent as
// many integer as FP64 instructions).  Derivation with h = hb ? 7 : 0 (N - p = p ^ 7):
//   xi   lines (i = p^h, j = r0, k = r1): (p ^ h ^ s(r0,r1)) | (r0 ^ (r1&1)) << 3 | r1 << 6
//   eta  lines (j = p^h, i = r0, k = r1): (r0 ^ (r1&1)<<2 ^ h>>1 ^ p>>1) | (h ^ (r1&1) ^ p) << 3 | r1 << 6
//   zeta lines (k = p^h, i = r0, j = r1): (r0 ^ r1>>1 ^ (h&1)<<2 ^ (p&1)<<2) | (r1 ^ (h&1) ^ (p&1)) << 3 | (h ^ p) << 6
template <int NP>
__device__ __forceinline__ int gl2_sac(int dir, int p, int hb, int r0, int r1) {
  if (NP == 8) {
    const int h = hb ? 7 : 0;
    const int b0 = (h ^ (((r0 >> 1) & 3) | ((r1 & 1) << 2))) | ((r0 ^ (r1 & 1)) << 3) | (r1 << 6);
    const int b1 = (r0 ^ ((r1 & 1) << 2) ^ ((h >> 1) & 3)) | ((h ^ (r1 & 1)) << 3) | (r1 << 6);
    const int b2 = (r0 ^ ((r1 >> 1) & 3) ^ ((h & 1) << 2)) | ((r1 ^ (h & 1)) << 3) | (h << 6);
    const int base = (dir == 0) ? b0 : (dir == 1) ? b1 : b2;
    const int off = (dir == 0) ? p : (dir == 1) ? ((p >> 1) | (p << 3)) : (((p & 1) * 12) | (p << 6));
    return base ^ off;
  }
  return gl2_sa<NP>(dir, gl2_ap<NP>(hb, p), r0, r1);
}

// One direction sweep of the advective two-point fluxes (+ the weak-form derivative of the viscous flux of that direction).
//   sa[q]          shared-memory address of canonical line position q (state planes and metric planes share the layout)
//   sMd, ma[q]     the three metric planes of this direction and the address of position q in them (ma == sa when the planes
//                  live in shared memory, the natural node index when they are read from global memory)
//   vp[v]          plane of the contravariant viscous flux of this direction, variable v+1 (read when lineVisc)
//   ut[c][v]       in/out: result for the own nodes c (HALF of them), WITHOUT the half sign (the caller applies it); the caller
//                  presets it (zero, or the zeta derivative of h~ that the node phase formed from registers)
template <int NP, int STYPE, bool PARA>
__device__ __forceinline__ void gl2_pair_sweep(const GL2Basis<NP>& B, const double* __restrict__ sP, const double* const (&vp)[4],
                                               const bool lineVisc, const double* __restrict__ sMd, const int (&sa)[NP],
                                               const int (&ma)[NP], double (&ut)[NP / 2][5], const double kappa) {
  constexpr int HALF = NP / 2, PL = NP * NP * NP;
  constexpr bool PI = (STYPE == FLEXI_SPLITDG_PI);
  constexpr int PX = PARA ? 6 : 5;
  if (PARA && lineVisc) {  // volint.f90:310-315: Ut += D_Hat_T(l,i) * f~visc(l)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      double L[NP];
#pragma unroll
      for (int p = 0; p < NP; ++p) L[p] = vp[v][sa[p]];
#pragma unroll
      for (int c = 0; c < HALF; ++c) {
        double s0 = B.dh[c] * L[0];
#pragma unroll
        for (int p = 1; p < NP; ++p) s0 = s0 + B.dh[p * NP + c] * L[p];
        ut[c][v + 1] = s0;
      }
    }
  }
  // metric terms of this direction along the whole line
  double Mo[HALF][3];
#pragma unroll
  for (int c = 0; c < HALF; ++c)
#pragma unroll
    for (int m = 0; m < 3; ++m) Mo[c][m] = sMd[m * PL + ma[c]];
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
    const double pm0 = sMd[ma[l]], pm1 = sMd[PL + ma[l]], pm2 = sMd[2 * PL + ma[l]];
#pragma unroll
    for (int c = 0; c < HALF; ++c) {
      const double dl = B.dvs[l * NP + c];
      if (PI) {
        const double su = a[c][1] + bu, sv = a[c][2] + bv, sw = a[c][3] + bw, sp4 = a[c][4] + bp, sH = a[c][5] + bx;
        const double m0 = Mo[c][0] + pm0, mm1 = Mo[c][1] + pm1, mm2 = Mo[c][2] + pm2;
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
        split_pair_generic<STYPE>(a[c], Mo[c], br, bu, bv, bw, bp, bx, pm0, pm1, pm2, Fl, kappa);
#pragma unroll
        for (int v = 0; v < 5; ++v) ut[c][v] = ut[c][v] + dl * Fl[v];
      }
    }
  }
}

// Template parameters: NVL = nVarLifting of the gradient arrays (their layout), NVC = number of lifting variables actually
// lifted: NVL, or 4 = (v1,v2,v3,T) on the time-stepping path, where nothing reads the density gradient (EvalDiffFlux3D does
// not use it and no gradient array is handed out).
// VVISC: PP_VISC != 0 (Sutherland / power law) -- a separate instantiation, so that the constant-viscosity kernel keeps mu0 and
// lambda as uniform constant-bank operands (as selects they cost per-thread registers and spills: 5.39 -> 5.70 ms).
template <int NP, int STYPE, bool PARA, int NVL, int NVC, int LIFT, bool VVISC = false>
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
  double* sM = sm + C::O_M;   // metric terms, 9 planes
  // hand-over plane of Ut(v+1), v = 0..4
  auto ut_plane = [&](int v) -> double* { return v == 0 ? sm + C::O_UT1 : v < 4 ? sF + (v - 1) * PL : sm + C::O_UT5; };
  const int tid = threadIdx.x;
  FX_STAMP_INIT
  if (FX_DBG(A, 8)) return;  // timing experiment: the cost of scheduling one CTA per element
  const int e = A.elems ? A.elems[blockIdx.x] : blockIdx.x;
  const int hb = tid / D::NP2, r = tid % D::NP2;  // half index (warp-uniform for NP = 8) and position inside the NP x NP sheet
  const int r0 = r % NP, r1 = r / NP;
  const double sg = hb ? -1. : 1.;
  const int sb = hb ? (int)0x80000000 : 0;  // sign bit of sg
  const Eqn& q = A.q;
  const double* gU = A.U + (size_t)e * 5 * D::NP3;
  const double* gM = A.m.Met + (size_t)e * 9 * D::NP3;
  double* gUt = A.Ut + (size_t)e * 5 * D::NP3;
  constexpr int FACE_IT = (6 * D::NP2 + NT - 1) / NT;
  constexpr bool ALLF = (FACE_IT * NT == 6 * D::NP2);  // every thread converts exactly FACE_IT face nodes (NP = 8: 3)

  // ---- phase A (zeta ownership): own conservative state on its way first, then the metric terms (asynchronous copies into
  //      their swizzled planes: no register, no wait until S2), then the neighbour traces of the face nodes this thread
  //      converts (registers), then ConsToPrim of the own nodes ----
  double Uc[HALF][5];
#pragma unroll
  for (int c = 0; c < HALF; ++c)
#pragma unroll
    for (int v = 0; v < 5; ++v) Uc[c][v] = __ldg(gU + v * D::NP3 + gl2_gn<NP>(2, gl2_ap<NP>(hb, c), r0, r1));
  if (C::MSM) {
#pragma unroll
    for (int c = 0; c < HALF; ++c) {
      const int n = gl2_gn<NP>(2, gl2_ap<NP>(hb, c), r0, r1), s = gl2_sac<NP>(2, c, hb, r0, r1);
#pragma unroll
      for (int m = 0; m < 9; ++m) cp_async8(sM + m * PL + s, gM + m * D::NP3 + n);
    }
    cp_async_commit();
  }
  double sJv[HALF];
  if (PARA) {
#pragma unroll
    for (int c = 0; c < HALF; ++c) sJv[c] = __ldg(A.m.sJ + (size_t)e * D::NP3 + gl2_gn<NP>(2, gl2_ap<NP>(hb, c), r0, r1));
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
      if (ALLF || t < 6 * D::NP2) {
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
    const double sRho = cons_to_prim_fast(q, Ucc, Pn);
    const int s = gl2_sac<NP>(2, c, hb, r0, r1);
#pragma unroll
    for (int v = 0; v < 4; ++v) sP[v * PL + s] = Pn[v];
    sP[4 * PL + s] = PI ? 4. * Pn[PRES] : Pn[PRES];
    if (PARA) sP[5 * PL + s] = Pn[TEMP];
    // 6th two-point quantity: H = (rhoE+p)/rho (PI), E = rhoE/rho (KG), rhoE (SD, DU)
    sP[PX * PL + s] = PI ? (Ucc[ENER] + Pn[PRES]) * sRho : (STYPE == FLEXI_SPLITDG_KG) ? Ucc[ENER] * sRho : Ucc[ENER];
  }
  if (!PARA && C::MSM) cp_async_wait_all();
  FX_STAMP(0)
  __syncthreads();  // S1: primitive planes (Euler: and metric planes) and the face table complete
  FX_STAMP(1)

  // the zeta derivative of h~ as this thread's contributions to the eight nodes of its line (own half | partner's half)
  double cown[HALF][4], cpart[HALF][4];
  if (PARA) {
    // ---- normal vectors of the (up to 2*HALF+1) face nodes among this thread's own nodes: requested now, used after S2.
    //      A line sits on the xi faces if i is 0 or N (all its nodes), on the eta faces if j is 0 or N; its end c == 0 sits on
    //      a zeta face (k = 0 for half 0, k = N for half 1). ----
    // Interior lanes (no xi / eta face) run the boundary code of the minus side with a zero normal vector: every statement of
    // the node phase below is unconditional (adding f*0 is exact), which leaves the four nodes of a thread in ONE basic block
    // that the compiler can interleave.  The divergent version had ~7 short blocks per node, each with its own exposed
    // shared-memory latency, and cost the same FP64 issue slots (a warp instruction occupies the pipe whatever its mask).
    const bool onI = (r0 == 0) | (r0 == N), onJ = (r1 == 0) | (r1 == N);
    const int lsI = (r0 == N) ? XI_PLUS : XI_MINUS, lsJ = (r1 == N) ? ETA_PLUS : ETA_MINUS, lsK = hb ? ZETA_PLUS : ZETA_MINUS;
    int fdI[HALF], fdJ[HALF], fdK = 0;
    double nI[HALF][3], nJ[HALF][3], nK[3];
    auto face_geo = [&](bool on, int ls, int ab, int& fd, double (&nv)[3]) {
      const int f = sFD[(ls - 1) * D::NP2 + ab];  // every slot of the table is valid: interior lanes load what their row's boundary lane loads
      const int fl = f & 0x3fffffff;
      const double* g = A.m.Geo + (size_t)(fl / D::NP2) * 10 * D::NP2 + fl % D::NP2;
      const double n0 = __ldg(g), n1 = __ldg(g + D::NP2), n2 = __ldg(g + 2 * D::NP2);
      fd = on ? f : 0;
      nv[0] = on ? n0 : 0.;
      nv[1] = on ? n1 : 0.;
      nv[2] = on ? n2 : 0.;
    };
#pragma unroll
    for (int c = 0; c < HALF; ++c) {
      const int k = gl2_ap<NP>(hb, c);
      face_geo(onI && !FX_DBG(A, 16), lsI, r1 + NP * k, fdI[c], nI[c]);
      face_geo(onJ && !FX_DBG(A, 16), lsJ, r0 + NP * k, fdJ[c], nJ[c]);
    }
    face_geo(!FX_DBG(A, 16), lsK, r0 + NP * r1, fdK, nK);

    if (nsr.side >= 0) {
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
    // ---- reference-space derivatives of the lifting variables (lifting_volint.t90:67-83), one line sweep per direction:
    //      xi and eta results change owner through the planes G1, G2; the zeta result stays with this thread ----
    double gzt[HALF][NVC];
#pragma unroll
    for (int dir = 0; dir < 3; ++dir) {
      if (FX_DBG(A, 2)) break;
      int sa[NP];
#pragma unroll
      for (int p = 0; p < NP; ++p) sa[p] = gl2_sac<NP>(dir, p, hb, r0, r1);
#pragma unroll
      for (int v = 0; v < NVC; ++v) {
        const double* pl = sP + lift_prim(NVC, v) * PL;
        double L[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) L[p] = pl[sa[p]];
#pragma unroll
        for (int c = 0; c < HALF; ++c) {
          double s0 = B.dt[c] * L[0];
#pragma unroll
          for (int p = 1; p < NP; ++p) s0 = s0 + B.dt[p * NP + c] * L[p];
          gzt[c][v] = sgn_apply(s0, sb);
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
      if ((ALLF || faceSlot[it] >= 0) && !FX_DBG(A, 16)) {
        double* o = sF + (faceSlot[it] & 0xffff);
        const int nF = faceSlot[it] >> 16;
        double Pb[6];
        const double lh = (LIFT == 1 && faceLh[it]) ? B.lhpN : B.lhm0;
        cons_to_prim_fast(q, fU[it], Pb);
#pragma unroll
        for (int v = 0; v < NVC; ++v) {
          const int pv = lift_prim(NVC, v);
          o[v * D::NP2] = (0.5 * (faceSgn[it] * (Pb[pv] - sP[pv * PL + nF])) * fSurf[it]) * lh;
        }
      }
    }
    FX_STAMP(2)
    if (C::MSM) cp_async_wait_all();
    // gradient output (volume gradients, gradient traces of BC sides / of all sides) is the rare case: one block-uniform flag,
    // reduced by the barrier itself, so that the node phase tests a single condition per node
    int bcBits = fdK;
#pragma unroll
    for (int c = 0; c < HALF; ++c) bcBits |= fdI[c] | fdJ[c];
    // S2: G1, G2, the face lifting fluxes and the metric planes complete
    const bool rareAny = __syncthreads_or(A.fullGrad || (bcBits & 0x40000000) || A.gvol[0] != nullptr);
    FX_STAMP(3)

    // ---- zeta ownership: metric combination, surface part, Jacobian, traces for the side kernel, viscous flux ----
#pragma unroll
    for (int c = 0; c < HALF; ++c)
#pragma unroll
      for (int v = 0; v < 4; ++v) cown[c][v] = cpart[c][v] = 0.;
    // PP_VISC != 0: the T plane is dead after S2 (until it hands over Ut(5) after S3), so every thread turns the temperature of
    // its own nodes into their viscosity in place, in one cold block here -- inside the four unrolled node bodies the inlined
    // exp/log cost the hot path registers and spills
    if (VVISC) {
#pragma unroll 1
      for (int c = 0; c < HALF; ++c) {
        double* const pT = sP + TEMP * PL + gl2_sac<NP>(2, c, hb, r0, r1);
        *pT = viscosity_var(q, *pT);
      }
    }
    if (!FX_DBG(A, 4)) {
      constexpr int LV = (NVC == 5) ? 1 : 0;
      const int i = r0, j = r1;
#pragma unroll
      for (int c = 0; c < HALF; ++c) {
        const int k = gl2_ap<NP>(hb, c);
        const int s = gl2_sac<NP>(2, c, hb, r0, r1);
        const int n = i + NP * (j + NP * k);
        double Mc[9];
#pragma unroll
        for (int m = 0; m < 9; ++m) Mc[m] = C::MSM ? sM[m * PL + s] : __ldg(gM + m * D::NP3 + n);
        const double sJc = sJv[c];
        // ---- boundary nodes: unconditional code, see the remark at the normal vectors.  The zeta faces belong to the line
        //      end c == 0 of every thread, a compile-time condition. ----
        const bool onK = (c == 0);  // compile-time after unrolling
        const int abI = j + NP * k, abJ = i + NP * k, abK = i + NP * j;
        const double* const oI = sF + (lsI - 1) * NVC * D::NP2 + abI;
        const double* const oJ = sF + (lsJ - 1) * NVC * D::NP2 + abJ;
        const double* const oK = sF + (lsK - 1) * NVC * D::NP2 + abK;
        double gx[NVC], gy[NVC], gz[NVC];
#pragma unroll
        for (int v = 0; v < NVC; ++v) {
          const double gxi = sW[v * PL + s], geta = sW[(NVC + v) * PL + s], gzeta = gzt[c][v];
          gx[v] = Mc[0] * gxi + Mc[3] * geta + Mc[6] * gzeta;
          gy[v] = Mc[1] * gxi + Mc[4] * geta + Mc[7] * gzeta;
          gz[v] = Mc[2] * gxi + Mc[5] * geta + Mc[8] * gzeta;
        }
        double gxo[NVC], gyo[NVC], gzo[NVC];  // BR2: gradients without the surface part (what the faces are prolonged from)
        double fI[NVC], fJ[NVC], fK[NVC];     // Flux * L_Hat of the node's faces
#pragma unroll
        for (int v = 0; v < NVC; ++v) {
          fI[v] = oI[v * D::NP2];
          fJ[v] = oJ[v * D::NP2];
          fK[v] = onK ? oK[v * D::NP2] : 0.;
        }
        // lifting, surface part (BR1 lifting_br1.t90:195-292; BR2 lifting_br2.t90:194-311)
        auto surf_in = [&](const double (&f)[NVC], const double (&nv)[3]) {
#pragma unroll
          for (int v = 0; v < NVC; ++v) {
            if (LIFT == 1) {
              gx[v] = gx[v] + f[v] * nv[0];
              gy[v] = gy[v] + f[v] * nv[1];
              gz[v] = gz[v] + f[v] * nv[2];
            } else {
              gx[v] = gx[v] + sJc * (f[v] * nv[0]);
              gy[v] = gy[v] + sJc * (f[v] * nv[1]);
              gz[v] = gz[v] + sJc * (f[v] * nv[2]);
            }
          }
        };
        if (LIFT == 1) {  // BR1: surface part, then ApplyJacobianLifting
          surf_in(fI, nI[c]);
          surf_in(fJ, nJ[c]);
          if (onK) surf_in(fK, nK);
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
          surf_in(fI, nI[c]);
          surf_in(fJ, nJ[c]);
          if (onK) surf_in(fK, nK);
        }
        // ---- viscous flux of this node (flux.f90:149-241) ----
        const double v1 = sP[VEL1 * PL + s], v2 = sP[VEL2 * PL + s], v3 = sP[VEL3 * PL + s];
        const double muC = VVISC ? sP[TEMP * PL + s] : q.mu0, lamC = VVISC ? muC * q.cpPr : q.lambda;
        double fv[4], gv[4], hv[4];
        eval_diff_flux3d_mu(muC, lamC, v1, v2, v3, gx + LV, gy + LV, gz + LV, fv, gv, hv);
        // ---- what the side kernel needs from this element.  ViscousFlux_Side (riemann.f90:448-498) forms
        //      0.5*n.(F_v(U_L,gradU_L) + F_v(U_R,gradU_R)); each half depends on one element only, and on LGL nodes the face
        //      state and (BR1) the face gradient are copies of the boundary node's (prolongtoface.t90:270-303), so the element
        //      writes n.F_v of its own side (4 doubles per face node) instead of the 3*nVarLifting gradient traces.  The
        //      gradient traces themselves are still written for boundary-condition sides (GetBoundaryFlux needs them) and, when
        //      the host asked for a full DG operator evaluation, for every side (the one rare branch of the node phase). ----
        auto face_out = [&](bool on, int ls, int fd, const double (&f)[NVC], const double (&nv)[3]) {
          const bool master = fd >= 0;
          const int fl = fd & 0x3fffffff, side = fl / D::NP2, pq = fl % D::NP2;
          double gfx[NVC], gfy[NVC], gfz[NVC];
          double nf[4];
          if (LIFT == 2) {  // BR2 face gradient: volume-only value + eta*L_Minus(0)*sJ*Flux*L_HatMinus(0) of the side itself
            const double w = sEta[ls - 1] * B.lm0;
#pragma unroll
            for (int v = 0; v < NVC; ++v) {
              gfx[v] = gxo[v] + w * (sJc * (f[v] * nv[0]));
              gfy[v] = gyo[v] + w * (sJc * (f[v] * nv[1]));
              gfz[v] = gzo[v] + w * (sJc * (f[v] * nv[2]));
            }
            double ff[4], gf[4], hf[4];
            eval_diff_flux3d_mu(muC, lamC, v1, v2, v3, gfx + LV, gfy + LV, gfz + LV, ff, gf, hf);
#pragma unroll
            for (int v = 0; v < 4; ++v) nf[v] = nv[0] * ff[v] + nv[1] * gf[v] + nv[2] * hf[v];
          } else {
#pragma unroll
            for (int v = 0; v < 4; ++v) nf[v] = nv[0] * fv[v] + nv[1] * gv[v] + nv[2] * hv[v];
          }
          // predicated stores (no branch: see the remark at the normal vectors)
          double* const dv = (master ? A.fvm : A.fvs) + (size_t)side * 4 * D::NP2 + pq;
#pragma unroll
          for (int v = 0; v < 4; ++v) st_global_if(on, dv + v * D::NP2, nf[v]);
        };
        face_out(onI, lsI, fdI[c], fI, nI[c]);
        face_out(onJ, lsJ, fdJ[c], fJ, nJ[c]);
        if (onK) face_out(true, lsK, fdK, fK, nK);
        if (__builtin_expect(rareAny, 0)) {  // gradient output: volume gradients on request, gradient traces of BC sides / all sides
          if (A.gvol[0]) {
#pragma unroll
            for (int v = 0; v < NVC; ++v) {
              A.gvol[0][((size_t)e * NVL + v + VOFF) * D::NP3 + n] = gx[v];
              A.gvol[1][((size_t)e * NVL + v + VOFF) * D::NP3 + n] = gy[v];
              A.gvol[2][((size_t)e * NVL + v + VOFF) * D::NP3 + n] = gz[v];
            }
          }
          auto grad_out = [&](bool on, int ls, int fd, const double (&f)[NVC], const double (&nv)[3]) {
            if (!(on && (A.fullGrad || (fd & 0x40000000)))) return;
            const bool master = fd >= 0;
            const int fl = fd & 0x3fffffff, side = fl / D::NP2, pq = fl % D::NP2;
            double* const d0 = master ? A.gm[0] : A.gs[0];
            double* const d1 = master ? A.gm[1] : A.gs[1];
            double* const d2 = master ? A.gm[2] : A.gs[2];
            const size_t o0 = (size_t)side * NVL * D::NP2 + pq;
            const double w = (LIFT == 2) ? sEta[ls - 1] * B.lm0 : 0.;
#pragma unroll
            for (int v = 0; v < NVC; ++v) {
              const size_t o = o0 + (size_t)(v + VOFF) * D::NP2;
              d0[o] = (LIFT == 2) ? gxo[v] + w * (sJc * (f[v] * nv[0])) : gx[v];
              d1[o] = (LIFT == 2) ? gyo[v] + w * (sJc * (f[v] * nv[1])) : gy[v];
              d2[o] = (LIFT == 2) ? gzo[v] + w * (sJc * (f[v] * nv[2])) : gz[v];
            }
          };
          grad_out(onI, lsI, fdI[c], fI, nI[c]);
          grad_out(onJ, lsJ, fdJ[c], fJ, nJ[c]);
          if (onK) grad_out(true, lsK, fdK, fK, nK);
        }
        // ---- contravariant components (volint.f90:378-422): f~, g~ over the node's own G1/G2 slots (only this thread reads
        //      them); h~ is spent right away on its zeta derivative (volint.f90:310-315 for the zeta direction):
        //      D_Hat_T(l = c, i = t) * h~(c) goes to node t of the line, t < HALF own, t >= HALF the partner's ----
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          sW[(0 + v) * PL + s] = fv[v] * Mc[0] + gv[v] * Mc[1] + hv[v] * Mc[2];
          sW[(4 + v) * PL + s] = fv[v] * Mc[3] + gv[v] * Mc[4] + hv[v] * Mc[5];
          const double hvc = fv[v] * Mc[6] + gv[v] * Mc[7] + hv[v] * Mc[8];
#pragma unroll
          for (int t = 0; t < HALF; ++t) {
            cown[t][v] = cown[t][v] + B.dh[c * NP + t] * hvc;
            cpart[t][v] = cpart[t][v] + B.dh[c * NP + HALF + t] * hvc;
          }
        }
      }
    }
    FX_STAMP(4)
    __syncthreads();  // S3: f~, g~ complete; the face table, the face fluxes and the T plane are free from here on
    FX_STAMP(5)

    // ---- the partner's share of the zeta derivative of h~ goes to the partner's slots of the hand-over planes (true sign) ----
#pragma unroll
    for (int t = 0; t < HALF; ++t) {
      const int s = gl2_sac<NP>(2, HALF + t, hb, r0, r1);
#pragma unroll
      for (int v = 0; v < 4; ++v) ut_plane(v + 1)[s] = sgn_apply(cpart[t][v], sb);
    }
    __syncthreads();  // S3b
    FX_STAMP(6)
  }

  // ---- three direction sweeps (zeta, xi, eta) of ONE copy of the sweep code.  The partial sums change owner twice through the
  //      hand-over planes; the eta sweep sends the raw volume integral to HBM ----
#pragma unroll 1
  for (int it = FX_DBG(A, 1) ? 3 : 0; it < 3; ++it) {
    const int dir = (it == 0) ? 2 : it - 1;
    int sa[NP];  // indexed by canonical position
#pragma unroll
    for (int p = 0; p < NP; ++p) sa[p] = gl2_sac<NP>(dir, p, hb, r0, r1);
    const double* vp[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) vp[v] = sW + ((dir & 1) * 4 + v) * PL;
    double ut[HALF][5];
#pragma unroll
    for (int c = 0; c < HALF; ++c) {
      ut[c][0] = 0.;
#pragma unroll
      for (int v = 1; v < 5; ++v) ut[c][v] = (PARA && it == 0) ? cown[c][v - 1] + sg * ut_plane(v)[sa[c]] : 0.;
    }
    if constexpr (C::MSM) {
      gl2_pair_sweep<NP, STYPE, PARA>(B, sP, vp, dir < 2, sM + dir * 3 * PL, sa, sa, ut, A.q.kappa);
    } else {
      int gn[NP];
#pragma unroll
      for (int p = 0; p < NP; ++p) gn[p] = gl2_gn<NP>(dir, gl2_ap<NP>(hb, p), r0, r1);
      gl2_pair_sweep<NP, STYPE, PARA>(B, sP, vp, dir < 2, gM + dir * 3 * D::NP3, sa, gn, ut, A.q.kappa);
    }
    // hand-over: v = 0 through the plane R0 or (when R0 is too small for a plane) through the output array, which stays in L2
    if (it == 0) {
#pragma unroll
      for (int c = 0; c < HALF; ++c) {
        if (C::UT1S)
          ut_plane(0)[sa[c]] = sgn_apply(ut[c][0], sb);
        else
          gUt[gl2_gn<NP>(dir, gl2_ap<NP>(hb, c), r0, r1)] = sgn_apply(ut[c][0], sb);
#pragma unroll
        for (int v = 1; v < 5; ++v) ut_plane(v)[sa[c]] = sgn_apply(ut[c][v], sb);
      }
    } else if (it == 1) {
#pragma unroll
      for (int c = 0; c < HALF; ++c) {
        if (C::UT1S) {
          ut_plane(0)[sa[c]] = ut_plane(0)[sa[c]] + sg * ut[c][0];
        } else {
          double* const g = gUt + gl2_gn<NP>(dir, gl2_ap<NP>(hb, c), r0, r1);
          *g = __ldcg(g) + sg * ut[c][0];
        }
#pragma unroll
        for (int v = 1; v < 5; ++v) ut_plane(v)[sa[c]] = ut_plane(v)[sa[c]] + sg * ut[c][v];
      }
    } else {
#pragma unroll
      for (int c = 0; c < HALF; ++c) {
        const int n = gl2_gn<NP>(dir, gl2_ap<NP>(hb, c), r0, r1);
        gUt[n] = (C::UT1S ? ut_plane(0)[sa[c]] : __ldcg(gUt + n)) + sg * ut[c][0];
#pragma unroll
        for (int v = 1; v < 5; ++v) gUt[v * D::NP3 + n] = ut_plane(v)[sa[c]] + sg * ut[c][v];
      }
    }
    FX_STAMP(7 + it)
    if (it < 2) __syncthreads();  // S4, S5
    FX_STAMP(10 + it)
  }
}

}  // namespace fx
