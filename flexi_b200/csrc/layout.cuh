// layout.cuh -- device data model of the DGSEM hot path.
//
// The Fortran host stores everything "variable fastest" (AoS): U(var,i,j,k,elem), U_master(var,p,q,side)
// (dg_vars.f90:31-102, mesh_vars.f90:62-153).  On the device every array is element-major SoA:
//
//   volume   A[elem][var][n]     n  = i + NP*(j + NP*k)      one contiguous NP^3-double plane per variable and element
//   side     S[side][var][pq]    pq = p + NP*q               one contiguous NP^2-double plane per variable and side
//   metrics  Met[elem][9][n]     (f1,f2,f3, g1,g2,g3, h1,h2,h3) = Metrics_{f,g,h}Tilde(1:3), sJ[elem][n]
//   facegeo  Geo[side][10][pq]   (NormVec 1:3, TangVec1 1:3, TangVec2 1:3, SurfElem)
//
// so that a warp reading consecutive nodes of one variable issues fully coalesced 256-byte requests.  Conversions happen
// only at the C-ABI boundary (set_state / get_state / get_array / init).
#pragma once
#include <cooperative_groups.h>
#include <stdint.h>

#include "physics.cuh"

namespace fx {

// local sides, flexi.h:90-95
enum { ZETA_MINUS = 1, ETA_MINUS = 2, XI_PLUS = 3, ETA_PLUS = 4, XI_MINUS = 5, ZETA_PLUS = 6 };

// per-element side table entry: ElemToSide(E2S_SIDE_ID,locSide,elem) (0-based here) and ElemToSide(E2S_FLIP,...)
struct SideRef {
  int32_t side;  // 0-based SideID
  int32_t flip;  // 0 = this element is the master of the side
};

// read-only mesh + basis tables shared by all kernels
struct Mesh {
  int nElems, nSides;
  const double* Met;   // [e][9][NP3]
  const double* sJ;    // [e][NP3]
  const double* Geo;   // [s][10][NP2]
  const SideRef* e2s;  // [e][6]
  const int32_t* s2v2; // [(flip*6+ls-1)*NP2 + pq] -> a + NP*b (volume-face coordinates), mappings.f90:389-397
  const int32_t* v2s;  // inverse: [(flip*6+ls-1)*NP2 + ab] -> pq
  const double* basis; // [dt | dh | dv | L_Minus | L_Plus | L_HatMinus | L_HatPlus], matrices as M[l*NP+i] = M(l,i)
  int nBCSides;           // BC sides are sides 0..nBCSides-1 (mesh.f90:257-269)
  const int32_t* bcType;  // [nBCSides] BC_TYPE of each BC side (needed by BR2: etaBR2_wall on types 3/4)
};

template <int NP>
struct Dim {
  static constexpr int NP2 = NP * NP;
  static constexpr int NP3 = NP * NP * NP;
  static constexpr int NBASIS = 3 * NP2 + 4 * NP;
  static constexpr int O_DT = 0, O_DH = NP2, O_DV = 2 * NP2, O_LM = 3 * NP2, O_LP = 3 * NP2 + NP, O_LHM = 3 * NP2 + 2 * NP,
                       O_LHP = 3 * NP2 + 3 * NP;
  // Large degrees (N = 11: 1728 nodes): the one-thread-per-node element kernels run on a thread-block CLUSTER of two CTAs, each
  // with half the nodes; the work planes of the element span the shared memory of both (distributed shared memory), see PairPtr.
  static constexpr bool PAIR = NP3 > 1024;
  static constexpr int THREADS = PAIR ? NP3 / 2 : NP3;
  static constexpr int PAIR_DOUBLES = 14 * NP3;  // doubles of the element's work space held by each CTA of the pair (N = 11: 189 KB)
  static_assert(!PAIR || (NP3 % 2 == 0 && NP3 / 2 <= 1024), "the CTA pair splits the nodes of an element in two halves");
};

// ---- element context: one CTA per element (N <= 9), or a cluster of two CTAs per element (Dim::PAIR) ----
// PairPtr is a pointer into the element's VIRTUAL work space [0, 2*PAIR_DOUBLES): the lower half lives in the shared memory of
// CTA rank 0, the upper half in that of rank 1; either CTA reaches both through generic addresses (DSMEM).  It offers exactly what
// the kernels use of a double* (p + n, p[n]), so the kernel bodies are the same source for both organisations.
template <int NP>
struct PairPtr {
  static constexpr int PER = Dim<NP>::PAIR_DOUBLES;
  double* lo;  // shared memory of CTA rank 0
  double* hi;  // shared memory of CTA rank 1, minus PER doubles
  int off;
  __device__ __forceinline__ PairPtr operator+(int d) const { return PairPtr{lo, hi, off + d}; }
  __device__ __forceinline__ double& operator[](int i) const {
    const int q = off + i;
    return (q < PER ? lo : hi)[q];
  }
};
template <int NP>
__device__ __forceinline__ auto elem_smem(double* raw) {
  if constexpr (Dim<NP>::PAIR) {
    namespace cg = cooperative_groups;
    cg::cluster_group c = cg::this_cluster();
    c.sync();  // the partner CTA has started: its shared memory may be addressed from here on
    return PairPtr<NP>{c.map_shared_rank(raw, 0), c.map_shared_rank(raw, 1) - Dim<NP>::PAIR_DOUBLES, 0};
  } else {
    return raw;
  }
}
// node id of this thread within its element, element slot of this CTA (pair), barrier over all threads of the element
template <int NP>
__device__ __forceinline__ int elem_tid() {
  if constexpr (Dim<NP>::PAIR)
    return (int)(cooperative_groups::this_cluster().block_rank() * Dim<NP>::THREADS + threadIdx.x);
  else
    return (int)threadIdx.x;
}
template <int NP>
__device__ __forceinline__ int elem_blk() {
  if constexpr (Dim<NP>::PAIR)
    return (int)(blockIdx.x >> 1);
  else
    return (int)blockIdx.x;
}
template <int NP>
__device__ __forceinline__ void elem_sync() {
  if constexpr (Dim<NP>::PAIR)
    cooperative_groups::this_cluster().sync();
  else
    __syncthreads();
}
// before a CTA of a pair exits, its partner must be done reading its shared memory
template <int NP>
__device__ __forceinline__ void elem_exit() {
  if constexpr (Dim<NP>::PAIR) cooperative_groups::this_cluster().sync();
}

// volume node index of the face node (a,b) at depth l below local side ls (prolongtoface.t90:232-263):
// xi faces carry (j,k), eta faces (i,k), zeta faces (i,j).
template <int NP>
__device__ __forceinline__ void face_line(int ls, int a, int b, int& base, int& stride) {
  if (ls == XI_MINUS || ls == XI_PLUS) {
    base = NP * (a + NP * b);
    stride = 1;
  } else if (ls == ETA_MINUS || ls == ETA_PLUS) {
    base = a + NP * NP * b;
    stride = NP;
  } else {
    base = a + NP * b;
    stride = NP * NP;
  }
}
__device__ __forceinline__ bool is_minus_side(int ls) { return ls == XI_MINUS || ls == ETA_MINUS || ls == ZETA_MINUS; }

// EvalElemFaceG / EvalElemFaceGL (prolongtoface.t90:211-303) of one variable plane held in shared memory
template <int NP, bool GL, class SP>
__device__ __forceinline__ double face_eval(const SP plane, const SP sB, int ls, int a, int b) {
  int base, stride;
  face_line<NP>(ls, a, b, base, stride);
  if (GL) return plane[base + (is_minus_side(ls) ? 0 : (NP - 1) * stride)];
  const auto L = sB + (is_minus_side(ls) ? Dim<NP>::O_LM : Dim<NP>::O_LP);
  double s = plane[base] * L[0];
#pragma unroll
  for (int l = 1; l < NP; ++l) s = s + plane[base + l * stride] * L[l];
  return s;
}

}  // namespace fx
