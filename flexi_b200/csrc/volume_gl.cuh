// volume_gl.cuh -- the first Gauss-Lobatto element kernel; since volume_gl2.cuh took over the split form on even N+1 (the headline
// configuration) it serves the GL weak form and the split form on odd N+1: ConsToPrim + BR1/BR2 lifting +
// ApplyJacobianLifting + ProlongToFaceLifting + VolInt (split or weak form) of one element per CTA.
//
// Replaces, for PP_NodeType = GAUSS-LOBATTO (file:line under /root/reference/src):
//   eos.f90:280-300 (ConsToPrim_Volume), dg/lifting/lifting_br1.t90:48-181, lifting_volint.t90:26-118,
//   lifting_fillflux.t90:40-239, interpolation/applyjacobian.t90:62-112, prolongtoface.t90:270-303,
//   dg/volint.f90:117-187 (weak) and :211-371 (split), equations/navierstokes/splitflux.f90:160-700, flux.f90:66-241.
//
// Thread mapping: NP*NP*(NP/KPT) threads per element; thread (i,j,kb) owns the KPT nodes (i,j,kb*KPT+c).  A warp is a
// (8 i) x (4 j) patch of one zeta-plane, so xi/eta line reads from shared memory are broadcasts (1 wavefront) and the
// zeta direction is served from the thread's own registers plus the other k-blocks.  The element's conservative state and
// (split form) its 9 metric planes are staged with one cp.async.bulk (TMA bulk copy) each, completion on an mbarrier.
// On Gauss-Lobatto nodes a face value is a volume node, so face work is done by the thread that owns the boundary node:
// it reads the neighbour trace, forms the lifting flux, and writes the gradient traces -- no face staging, and the
// volume gradients never leave the SM.
#pragma once
#include <stdint.h>
#include "kernels.cuh"

namespace fx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

template <int NP, int KPT, bool SPLIT, bool PARA>
struct VolGLCfg {
  using D = Dim<NP>;
  static_assert(NP % KPT == 0, "KPT must divide NP");
  static constexpr int KB = NP / KPT;
  static constexpr int NT = D::NP2 * KB;
  static constexpr bool BULK = (NP % 2 == 0);  // plane bytes must be a multiple of 16 for cp.async.bulk
  static constexpr int P_PLANES = 6;
  // region X: first the six faces' neighbour traces + face geometry (9 values per face node, prefetched with cp.async and
  // turned into lifting fluxes in place), later the 9 metric planes the two-point fluxes read from their line partners
  static constexpr int X_FACE = PARA ? 6 * 9 * D::NP2 : 0;
  static constexpr int X_MET = SPLIT ? 9 * D::NP3 : 0;
  static constexpr int X_DOUBLES = ((X_FACE > X_MET ? X_FACE : X_MET) + 1) & ~1;
  static constexpr int W_PLANES = PARA ? (SPLIT ? 12 : 15) : (SPLIT ? 0 : 15);
  static constexpr int O_BASIS = 2;  // [0] = mbarrier
  static constexpr int NB = (D::NBASIS + 1) & ~1;
  static constexpr int O_FDST = O_BASIS + NB;  // int table [6][NP2]: where a boundary node's gradient trace goes
  static constexpr int NFD = PARA ? ((3 * D::NP2 + 1) & ~1) : 0;
  static constexpr int O_ETA = O_FDST + NFD;  // BR2 penalty eta of the six sides
  static constexpr int O_P = O_ETA + (PARA ? 6 : 0);
  static constexpr int O_X = O_P + P_PLANES * D::NP3;
  static constexpr int O_W = O_X + X_DOUBLES;
  static constexpr int DOUBLES = O_W + W_PLANES * D::NP3;
  static constexpr size_t BYTES = sizeof(double) * DOUBLES;
  static constexpr int MINB = (NT <= 256 && BYTES <= 115000) ? 2 : 1;  // two elements per SM when registers and smem allow
};

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// raw two-point quantities of the Pirozzoli flux between nodes a and b with summed metrics m = Ma + Mb:
//   contravariant flux (splitflux.f90:603-666 contracted with 0.5*(Ma+Mb)):
//   F0 = 1/4 rq, F1 = 1/8 rq su + 1/2 m0 sp, F2 = 1/8 rq sv + 1/2 m1 sp, F3 = 1/8 rq sw + 1/2 m2 sp, F4 = 1/8 rq sH
//   with rq = (ra+rb) * (m . (ua+ub)).  Constants are applied once per node after the pair loop.
struct PiAcc {
  double A0, A1, A2, A3, A4, B0, B1, B2;
  __device__ __forceinline__ void zero() { A0 = A1 = A2 = A3 = A4 = B0 = B1 = B2 = 0.; }
  __device__ __forceinline__ void flush(double* ut) {
    ut[0] = ut[0] + 0.25 * A0;
    ut[1] = ut[1] + (0.125 * A1 + 0.5 * B0);
    ut[2] = ut[2] + (0.125 * A2 + 0.5 * B1);
    ut[3] = ut[3] + (0.125 * A3 + 0.5 * B2);
    ut[4] = ut[4] + 0.125 * A4;
  }
};
struct PiPair {
  double rq, su, sv, sw, sH, sp, m0, m1, m2;
  __device__ __forceinline__ void form(const double* a, const double* Ma, double br, double bu, double bv, double bw, double bp, double bH,
                                       double b0, double b1, double b2) {
    su = a[1] + bu;
    sv = a[2] + bv;
    sw = a[3] + bw;
    sp = a[4] + bp;
    sH = a[5] + bH;
    m0 = Ma[0] + b0;
    m1 = Ma[1] + b1;
    m2 = Ma[2] + b2;
    rq = (a[0] + br) * (m0 * su + m1 * sv + m2 * sw);
  }
  __device__ __forceinline__ void add(PiAcc& acc, double dl) const {
    const double w8 = dl * rq, dp = dl * sp;
    acc.A0 += w8;
    acc.A1 += w8 * su;
    acc.A2 += w8 * sv;
    acc.A3 += w8 * sw;
    acc.A4 += w8 * sH;
    acc.B0 += dp * m0;
    acc.B1 += dp * m1;
    acc.B2 += dp * m2;
  }
};

// generic two-point flux for SD / MO / DU / KG / CH / IR in reference order; a[] = (rho,u,v,w,p,x5) with x5 = rhoE or E (KG)
template <int STYPE>
__device__ __forceinline__ void split_pair_generic(const double* a, const double* Ma, double br, double bu, double bv, double bw, double bp,
                                                   double bx, double b0, double b1, double b2, double* F, const double kappa) {
  Node na{a[0], a[1], a[2], a[3], a[4], a[5]}, nb{br, bu, bv, bw, bp, bx};
  if (STYPE == FLEXI_SPLITDG_KG) {  // Node.E is read as rho*E by split_volume_flux: rebuild it (x5 holds E = rhoE/rho)
    na.E = a[5] * a[0];
    nb.E = bx * br;
  }
  const double Mb[3] = {b0, b1, b2};
  split_volume_flux<STYPE>(na, nb, Ma, Mb, F, kappa);
}

template <int NP, int KPT, bool SPLIT, int STYPE, bool PARA, int NVL, int LIFT>
__global__ void __launch_bounds__(VolGLCfg<NP, KPT, SPLIT, PARA>::NT, VolGLCfg<NP, KPT, SPLIT, PARA>::MINB)
    k_volume_gl(VolArgs A) {
  using D = Dim<NP>;
  using C = VolGLCfg<NP, KPT, SPLIT, PARA>;
  constexpr int KB = C::KB, NT = C::NT;
  constexpr int N = NP - 1;
  extern __shared__ __align__(16) double sm[];
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sm);
  double* sB = sm + C::O_BASIS;
  double* sP = sm + C::O_P;  // rho,u,v,w,p, T (later H / rhoE / E for the split fluxes)
  int* sFD = reinterpret_cast<int*>(sm + C::O_FDST);
  double* sEta = sm + C::O_ETA;
  double* sX = sm + C::O_X;  // lifting fluxes of the faces, then the metric planes
  double* sM = sX;
  double* sW = sm + C::O_W;
  const int tid = threadIdx.x;
  const int e = A.elems ? A.elems[blockIdx.x] : blockIdx.x;
  const int ij = tid % D::NP2, i = ij % NP, j = ij / NP, kb = tid / D::NP2, k0 = kb * KPT;
  const Eqn& q = A.q;
  const double* dt = sB + D::O_DT;
  const double* dh = sB + D::O_DH;
  const double* dv = sB + D::O_DV;
  const double* gU = A.U + (size_t)e * 5 * D::NP3;
  const double* gM = A.m.Met + (size_t)e * 9 * D::NP3;
  constexpr uint32_t bU = 5 * D::NP3 * 8, bM = 9 * D::NP3 * 8;
  constexpr bool MET_EARLY = SPLIT && !PARA;  // without lifting region X is free from the start
  // own-node metrics of the lifting / flux loop: staged into the first 9 work planes; each node's slot is consumed by its
  // owner right before the owner overwrites it with the node's contravariant fluxes
  constexpr bool MET_WORK = (PARA || !SPLIT) && C::BULK;
  constexpr int FACE_IT = (6 * D::NP2 + NT - 1) / NT;

  // ---- stage the element: one TMA bulk copy for the 5 conservative planes (+ the 9 metric planes) ------------------------
  if (C::BULK) {
    if (tid == 0) mbar_init(mbar, 1);
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(mbar, bU + ((MET_EARLY || MET_WORK) ? bM : 0));
      bulk_g2s(sP, gU, bU, mbar);
      if (MET_EARLY) bulk_g2s(sM, gM, bM, mbar);
      if (MET_WORK) bulk_g2s(sW, gM, bM, mbar);
    }
  } else {
    for (int t = tid; t < 5 * D::NP3; t += NT) sP[t] = gU[t];
    if (MET_EARLY)
      for (int t = tid; t < 9 * D::NP3; t += NT) sM[t] = gM[t];
  }
  load_basis<NP>(sB, A.m.basis, tid, NT);
  // neighbour traces + face geometry of the six sides: cp.async into region X at the face node's volume-face slot, in flight
  // while the element itself is being staged
  int faceSlot[FACE_IT];
  double faceSgn[FACE_IT];
  if (PARA) {
#pragma unroll
    for (int it = 0; it < FACE_IT; ++it) {
      const int t = tid + it * NT;
      faceSlot[it] = -1;
      faceSgn[it] = 0.;
      if (t < 6 * D::NP2) {
        const int ls = t / D::NP2 + 1, pq = t % D::NP2;
        const SideRef sr = A.m.e2s[e * 6 + ls - 1];
        const int ab = A.m.s2v2[(sr.flip * 6 + ls - 1) * D::NP2 + pq];
        const double* nbTrace = ((sr.flip == 0) ? A.U_slave : A.U_master) + (size_t)sr.side * 5 * D::NP2 + pq;
        const double* g = A.m.Geo + (size_t)sr.side * 10 * D::NP2 + pq;
        double* o = sX + (ls - 1) * 9 * D::NP2 + ab;
#pragma unroll
        for (int v = 0; v < 5; ++v) cp_async8(o + v * D::NP2, nbTrace + v * D::NP2);
        cp_async8(o + 5 * D::NP2, g);
        cp_async8(o + 6 * D::NP2, g + D::NP2);
        cp_async8(o + 7 * D::NP2, g + 2 * D::NP2);
        cp_async8(o + 8 * D::NP2, g + 9 * D::NP2);
        faceSlot[it] = (ls - 1) * 9 * D::NP2 + ab;
        // destination of this face node's gradient trace: (side*NVL*NP2 + pq), sign bit set for the slave arrays
        sFD[(ls - 1) * D::NP2 + ab] = (sr.side * NVL * D::NP2 + pq) | (sr.flip == 0 ? 0 : (int)0x80000000);
        if (LIFT == 2 && pq == 0) {  // lifting_br2.t90:277-288: etaBR2_wall on no-slip wall sides
          double eta = q.etaBR2;
          if (sr.side < A.m.nBCSides) {
            const int bt = A.m.bcType[sr.side];
            if (bt == 3 || bt == 4) eta = q.etaBR2_wall;
          }
          sEta[ls - 1] = eta;
        }
        int base, stride;
        face_line<NP>(ls, ab % NP, ab / NP, base, stride);
        faceSgn[it] = (sr.flip == 0) ? 1. : -1.;
        faceSlot[it] |= (base + (is_minus_side(ls) ? 0 : N * stride)) << 16;  // volume node of this face node
      }
    }
    cp_async_commit();
  }
  double sJ[KPT];
#pragma unroll
  for (int c = 0; c < KPT; ++c) sJ[c] = A.m.sJ[(size_t)e * D::NP3 + ij + D::NP2 * (k0 + c)];
  if (C::BULK)
    mbar_wait(mbar, 0);
  else
    __syncthreads();

  // ---- ConsToPrim of the own nodes, in place ------------------------------------------------------------------------------
  double Pn[KPT][6], Uc[KPT][5], sRho[KPT];
#pragma unroll
  for (int c = 0; c < KPT; ++c) {
    const int n = ij + D::NP2 * (k0 + c);
#pragma unroll
    for (int v = 0; v < 5; ++v) Uc[c][v] = sP[v * D::NP3 + n];
    sRho[c] = cons_to_prim_srho(q, Uc[c], Pn[c]);
#pragma unroll
    for (int v = 1; v < 6; ++v) sP[v * D::NP3 + n] = Pn[c][v];  // only the owner touches its nodes
  }
  __syncthreads();  // primitive planes and basis visible to everyone

  double ut[KPT][5];
#pragma unroll
  for (int c = 0; c < KPT; ++c)
#pragma unroll
    for (int v = 0; v < 5; ++v) ut[c][v] = 0.;

  if (PARA) {
    // ---- lifting flux of the six faces, all lanes busy (lifting_fillflux.t90:81-107, strong form):
    //      0.5*(UPrim_slave-UPrim_master)*SurfElem per lifting variable, in place of the prefetched neighbour trace ----
    cp_async_wait_all();
#pragma unroll
    for (int it = 0; it < FACE_IT; ++it) {
      if (faceSlot[it] >= 0) {
        double* o = sX + (faceSlot[it] & 0xffff);
        const int nF = faceSlot[it] >> 16;
        double Un[5], Pb[6];
#pragma unroll
        for (int v = 0; v < 5; ++v) Un[v] = o[v * D::NP2];
        const double surf = o[8 * D::NP2];
        cons_to_prim(q, Un, Pb);
#pragma unroll
        for (int v = 0; v < NVL; ++v) {
          const int pv = lift_prim(NVL, v);
          o[v * D::NP2] = 0.5 * (faceSgn[it] * (Pb[pv] - sP[pv * D::NP3 + nF])) * surf;
        }
      }
    }
    __syncthreads();
  }

  if (PARA || !SPLIT) {
    constexpr int LV = (NVL == 5) ? 1 : 0;
#pragma unroll
    for (int c = 0; c < KPT; ++c) {
      const int k = k0 + c, n = ij + D::NP2 * k;
      double M[9];
#pragma unroll
      for (int m = 0; m < 9; ++m) M[m] = MET_EARLY ? sM[m * D::NP3 + n] : MET_WORK ? sW[m * D::NP3 + n] : gM[m * D::NP3 + n];
      double F[5] = {0., 0., 0., 0., 0.}, G[5] = {0., 0., 0., 0., 0.}, H[5] = {0., 0., 0., 0., 0.};
      if (PARA) {
        // ---- lifting, volume part (lifting_volint.t90:26-118) ----
        double gx[NVL], gy[NVL], gz[NVL];
        const int bx = n - i, by = n - j * NP;
#pragma unroll
        for (int v = 0; v < NVL; ++v) {
          const int pv = lift_prim(NVL, v);
          const double* pl = sP + pv * D::NP3;
          double gxi = dt[i] * pl[bx], geta = dt[j] * pl[by];
#pragma unroll
          for (int l = 1; l < NP; ++l) {
            gxi = gxi + dt[l * NP + i] * pl[bx + l];
            geta = geta + dt[l * NP + j] * pl[by + l * NP];
          }
          // zeta line: own k-block from registers, the other blocks from shared memory
          double gzeta = 0.;
#pragma unroll
          for (int c2 = 0; c2 < KPT; ++c2) gzeta = gzeta + dt[(k0 + c2) * NP + k] * Pn[c2][pv];
#pragma unroll
          for (int ob = 1; ob < KB; ++ob) {
            const int l0 = ((kb + ob) % KB) * KPT;
#pragma unroll
            for (int c2 = 0; c2 < KPT; ++c2) gzeta = gzeta + dt[(l0 + c2) * NP + k] * pl[ij + D::NP2 * (l0 + c2)];
          }
          gx[v] = M[0] * gxi + M[3] * geta + M[6] * gzeta;
          gy[v] = M[1] * gxi + M[4] * geta + M[7] * gzeta;
          gz[v] = M[2] * gxi + M[5] * geta + M[8] * gzeta;
        }
        // ---- lifting, surface part: on LGL nodes only the boundary nodes receive it ----
        const bool onFace[6] = {k == 0, j == 0, i == N, j == N, i == 0, k == N};  // local sides 1..6 (flexi.h:90-95)
        const int abFace[6] = {i + NP * j, i + NP * k, j + NP * k, i + NP * k, j + NP * k, i + NP * j};
        double gxo[NVL], gyo[NVL], gzo[NVL];  // BR2: gradients without the surface part (what the faces are prolonged from)
        if (LIFT == 1) {  // BR1 (lifting_br1.t90:195-292), then ApplyJacobianLifting
#pragma unroll
          for (int ls = 1; ls <= 6; ++ls) {
            if (onFace[ls - 1]) {
              const double* o = sX + (ls - 1) * 9 * D::NP2 + abFace[ls - 1];
              const double n0 = o[5 * D::NP2], n1 = o[6 * D::NP2], n2 = o[7 * D::NP2];
              const double lh = is_minus_side(ls) ? sB[D::O_LHM] : sB[D::O_LHP + N];
#pragma unroll
              for (int v = 0; v < NVL; ++v) {
                const double f = o[v * D::NP2];
                gx[v] = gx[v] + (f * n0) * lh;
                gy[v] = gy[v] + (f * n1) * lh;
                gz[v] = gz[v] + (f * n2) * lh;
              }
            }
          }
#pragma unroll
          for (int v = 0; v < NVL; ++v) {
            gx[v] = gx[v] * sJ[c];
            gy[v] = gy[v] * sJ[c];
            gz[v] = gz[v] * sJ[c];
          }
        } else {  // BR2 (lifting_br2.t90:111-131,194-311): Jacobian first, then sJ*Flux*L_HatMinus(0) on the boundary node
#pragma unroll
          for (int v = 0; v < NVL; ++v) {
            gx[v] = gx[v] * sJ[c];
            gy[v] = gy[v] * sJ[c];
            gz[v] = gz[v] * sJ[c];
            gxo[v] = gx[v];
            gyo[v] = gy[v];
            gzo[v] = gz[v];
          }
#pragma unroll
          for (int ls = 1; ls <= 6; ++ls) {
            if (onFace[ls - 1]) {
              const double* o = sX + (ls - 1) * 9 * D::NP2 + abFace[ls - 1];
              const double n0 = o[5 * D::NP2], n1 = o[6 * D::NP2], n2 = o[7 * D::NP2];
              const double lh = sB[D::O_LHM];
#pragma unroll
              for (int v = 0; v < NVL; ++v) {
                const double f = o[v * D::NP2];
                gx[v] = gx[v] + sJ[c] * (f * n0) * lh;
                gy[v] = gy[v] + sJ[c] * (f * n1) * lh;
                gz[v] = gz[v] + sJ[c] * (f * n2) * lh;
              }
            }
          }
        }
        if (A.gvol[0]) {
#pragma unroll
          for (int v = 0; v < NVL; ++v) {
            A.gvol[0][((size_t)e * NVL + v) * D::NP3 + n] = gx[v];
            A.gvol[1][((size_t)e * NVL + v) * D::NP3 + n] = gy[v];
            A.gvol[2][((size_t)e * NVL + v) * D::NP3 + n] = gz[v];
          }
        }
        // ---- ProlongToFaceLifting on LGL nodes = copy of the boundary node (prolongtoface.t90:270-303);
        //      BR2 adds eta*L_Minus(0)*sJ*Flux*L_HatMinus(0) of the side itself to the volume-only value ----
#pragma unroll
        for (int ls = 1; ls <= 6; ++ls) {
          if (onFace[ls - 1]) {
            const int fd = sFD[(ls - 1) * D::NP2 + abFace[ls - 1]];
            const bool master = fd >= 0;
            double* const d0 = master ? A.gm[0] : A.gs[0];
            double* const d1 = master ? A.gm[1] : A.gs[1];
            double* const d2 = master ? A.gm[2] : A.gs[2];
            const size_t o0 = (size_t)(fd & 0x7fffffff);
            if (LIFT == 1) {
#pragma unroll
              for (int v = 0; v < NVL; ++v) {
                const size_t o = o0 + (size_t)v * D::NP2;
                d0[o] = gx[v];
                d1[o] = gy[v];
                d2[o] = gz[v];
              }
            } else {
              const double* of = sX + (ls - 1) * 9 * D::NP2 + abFace[ls - 1];
              const double n0 = of[5 * D::NP2], n1 = of[6 * D::NP2], n2 = of[7 * D::NP2];
              const double w = sEta[ls - 1] * sB[D::O_LM];
              const double lh = sB[D::O_LHM];
#pragma unroll
              for (int v = 0; v < NVL; ++v) {
                const size_t o = o0 + (size_t)v * D::NP2;
                const double f = of[v * D::NP2];
                d0[o] = gxo[v] + w * (sJ[c] * (f * n0) * lh);
                d1[o] = gyo[v] + w * (sJ[c] * (f * n1) * lh);
                d2[o] = gzo[v] + w * (sJ[c] * (f * n2) * lh);
              }
            }
          }
        }
        // ---- viscous flux of this node (flux.f90:149-241) ----
        double fv[4], gv[4], hv[4];
        eval_diff_flux3d(q, Pn[c][TEMP], Pn[c][VEL1], Pn[c][VEL2], Pn[c][VEL3], gx + LV, gy + LV, gz + LV, fv, gv, hv);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          F[v + 1] = fv[v];
          G[v + 1] = gv[v];
          H[v + 1] = hv[v];
        }
      }
      if (!SPLIT) {  // weak form: advective flux goes through the same derivative (volint.f90:117-187)
        double fe[5], ge[5], he[5];
        eval_flux3d(Uc[c], Pn[c], fe, ge, he);
#pragma unroll
        for (int v = 0; v < 5; ++v) {
          F[v] = fe[v] + F[v];
          G[v] = ge[v] + G[v];
          H[v] = he[v] + H[v];
        }
      }
      // contravariant fluxes (volint.f90:378-422) -> work planes
      constexpr int V0 = SPLIT ? 1 : 0;  // the viscous density flux is identically zero
      constexpr int NV = 5 - V0;
#pragma unroll
      for (int v = V0; v < 5; ++v) {
        sW[(0 * NV + v - V0) * D::NP3 + n] = F[v] * M[0] + G[v] * M[1] + H[v] * M[2];
        sW[(1 * NV + v - V0) * D::NP3 + n] = F[v] * M[3] + G[v] * M[4] + H[v] * M[5];
        sW[(2 * NV + v - V0) * D::NP3 + n] = F[v] * M[6] + G[v] * M[7] + H[v] * M[8];
      }
    }
    __syncthreads();  // work planes complete; region X (face fluxes) and the T plane are free from here on
    if (SPLIT && PARA) {
      // the metric planes of the line partners arrive while the weak-form derivative below runs
      if (C::BULK) {
        if (tid == 0) {
          fence_proxy_async();
          mbar_expect_tx(mbar, bM);
          bulk_g2s(sM, gM, bM, mbar);
        }
      } else {
        for (int t = tid; t < 9 * D::NP3; t += NT) sM[t] = gM[t];
      }
    }
    // ---- weak-form derivative with D_Hat_T (volint.f90:172-184) ----
    {
      constexpr int V0 = SPLIT ? 1 : 0;
      constexpr int NV = 5 - V0;
#pragma unroll
      for (int c = 0; c < KPT; ++c) {
        const int k = k0 + c, n = ij + D::NP2 * k;
        const int bx = n - i, by = n - j * NP;
#pragma unroll
        for (int v = V0; v < 5; ++v) {
          const double* f = sW + (0 * NV + v - V0) * D::NP3;
          const double* g = sW + (1 * NV + v - V0) * D::NP3;
          const double* h = sW + (2 * NV + v - V0) * D::NP3;
          double s = 0.;
#pragma unroll
          for (int l = 0; l < NP; ++l) s = s + dh[l * NP + i] * f[bx + l] + dh[l * NP + j] * g[by + l * NP] + dh[l * NP + k] * h[ij + D::NP2 * l];
          ut[c][v] = s;
        }
      }
    }
  }

  if (SPLIT) {
    // ---- advective two-point fluxes (volint.f90:211-371) ----
    // plane 5 switches from T to the 6th two-point quantity: H = (rhoE+p)/rho (PI), E = rhoE/rho (KG), rhoE (SD, DU)
    double a[KPT][6];
#pragma unroll
    for (int c = 0; c < KPT; ++c) {
#pragma unroll
      for (int v = 0; v < 5; ++v) a[c][v] = Pn[c][v];
      a[c][5] = (STYPE == FLEXI_SPLITDG_PI) ? (Uc[c][ENER] + Pn[c][PRES]) * sRho[c] : (STYPE == FLEXI_SPLITDG_KG) ? Uc[c][ENER] * sRho[c] : Uc[c][ENER];
    }
    // (with lifting, the barrier after the flux loop already separated the last read of the T plane from this write)
#pragma unroll
    for (int c = 0; c < KPT; ++c) sP[5 * D::NP3 + ij + D::NP2 * (k0 + c)] = a[c][5];
    __syncthreads();
    if (PARA && C::BULK) mbar_wait(mbar, 1);

    constexpr bool PI = (STYPE == FLEXI_SPLITDG_PI);
    PiAcc acc[KPT];
#pragma unroll
    for (int c = 0; c < KPT; ++c) acc[c].zero();
    // xi and eta lines: partner from shared memory
#pragma unroll
    for (int c = 0; c < KPT; ++c) {
      const int n = ij + D::NP2 * (k0 + c);
#pragma unroll
      for (int dir = 0; dir < 2; ++dir) {
        const int me = dir == 0 ? i : j;
        const int stride = dir == 0 ? 1 : NP;
        const int base = n - me * stride;
        const double Ma[3] = {sM[(3 * dir + 0) * D::NP3 + n], sM[(3 * dir + 1) * D::NP3 + n], sM[(3 * dir + 2) * D::NP3 + n]};
#pragma unroll
        for (int l = 0; l < NP; ++l) {
          // the self pair is computed and weighted with zero instead of being skipped: a per-lane `continue` would split
          // the warp into NP paths that never reconverge inside the unrolled loop
          const int nl = base + l * stride;
          const double dl = (l == me) ? 0. : dv[l * NP + me];
          if (PI) {
            PiPair pr;
            pr.form(a[c], Ma, sP[nl], sP[D::NP3 + nl], sP[2 * D::NP3 + nl], sP[3 * D::NP3 + nl], sP[4 * D::NP3 + nl], sP[5 * D::NP3 + nl],
                    sM[(3 * dir + 0) * D::NP3 + nl], sM[(3 * dir + 1) * D::NP3 + nl], sM[(3 * dir + 2) * D::NP3 + nl]);
            pr.add(acc[c], dl);
          } else {
            double Fl[5];
            split_pair_generic<STYPE>(a[c], Ma, sP[nl], sP[D::NP3 + nl], sP[2 * D::NP3 + nl], sP[3 * D::NP3 + nl], sP[4 * D::NP3 + nl],
                                      sP[5 * D::NP3 + nl], sM[(3 * dir + 0) * D::NP3 + nl], sM[(3 * dir + 1) * D::NP3 + nl],
                                      sM[(3 * dir + 2) * D::NP3 + nl], Fl, q.kappa);
#pragma unroll
            for (int v = 0; v < 5; ++v) ut[c][v] = ut[c][v] + dl * Fl[v];
          }
        }
      }
    }
    // zeta line: pairs inside the own k-block are formed once and applied to both nodes (the symmetric update of
    // volint.f90:318-333); the other blocks' nodes are loaded once and paired with all own nodes
    double Mz[KPT][3];
#pragma unroll
    for (int c = 0; c < KPT; ++c)
#pragma unroll
      for (int m = 0; m < 3; ++m) Mz[c][m] = sM[(6 + m) * D::NP3 + ij + D::NP2 * (k0 + c)];
#pragma unroll
    for (int c = 0; c < KPT; ++c)
#pragma unroll
      for (int c2 = c + 1; c2 < KPT; ++c2) {
        const double d12 = dv[(k0 + c2) * NP + (k0 + c)];  // DVolSurf(l=k2, k1) -> node c
        const double d21 = dv[(k0 + c) * NP + (k0 + c2)];  // DVolSurf(l=k1, k2) -> node c2
        if (PI) {
          PiPair pr;
          pr.form(a[c], Mz[c], a[c2][0], a[c2][1], a[c2][2], a[c2][3], a[c2][4], a[c2][5], Mz[c2][0], Mz[c2][1], Mz[c2][2]);
          pr.add(acc[c], d12);
          pr.add(acc[c2], d21);
        } else {
          double Fl[5];
          split_pair_generic<STYPE>(a[c], Mz[c], a[c2][0], a[c2][1], a[c2][2], a[c2][3], a[c2][4], a[c2][5], Mz[c2][0], Mz[c2][1], Mz[c2][2], Fl, q.kappa);
#pragma unroll
          for (int v = 0; v < 5; ++v) {
            ut[c][v] = ut[c][v] + d12 * Fl[v];
            ut[c2][v] = ut[c2][v] + d21 * Fl[v];
          }
        }
      }
#pragma unroll
    for (int ob = 1; ob < KB; ++ob) {
      const int l0 = ((kb + ob) % KB) * KPT;
#pragma unroll
      for (int c2 = 0; c2 < KPT; ++c2) {
        const int l = l0 + c2, nl = ij + D::NP2 * l;
        const double br = sP[nl], bu = sP[D::NP3 + nl], bv = sP[2 * D::NP3 + nl], bw = sP[3 * D::NP3 + nl], bp = sP[4 * D::NP3 + nl],
                     bx = sP[5 * D::NP3 + nl];
        const double b0 = sM[6 * D::NP3 + nl], b1 = sM[7 * D::NP3 + nl], b2 = sM[8 * D::NP3 + nl];
#pragma unroll
        for (int c = 0; c < KPT; ++c) {
          const double dl = dv[l * NP + (k0 + c)];
          if (PI) {
            PiPair pr;
            pr.form(a[c], Mz[c], br, bu, bv, bw, bp, bx, b0, b1, b2);
            pr.add(acc[c], dl);
          } else {
            double Fl[5];
            split_pair_generic<STYPE>(a[c], Mz[c], br, bu, bv, bw, bp, bx, b0, b1, b2, Fl, q.kappa);
#pragma unroll
            for (int v = 0; v < 5; ++v) ut[c][v] = ut[c][v] + dl * Fl[v];
          }
        }
      }
    }
    if (PI) {
#pragma unroll
      for (int c = 0; c < KPT; ++c) acc[c].flush(ut[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < KPT; ++c)
#pragma unroll
    for (int v = 0; v < 5; ++v) A.Ut[((size_t)e * 5 + v) * D::NP3 + ij + D::NP2 * (k0 + c)] = ut[c][v];
}

}  // namespace fx
