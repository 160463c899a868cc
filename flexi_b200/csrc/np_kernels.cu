// np_kernels.cu -- instantiates kernels.cuh for one polynomial degree (compile with -DFLEXI_NP=<N+1>).
#include <math.h>
#include <stdlib.h>

#include "np_ops.h"
#if FLEXI_NP <= 10
#include "volume_gl.cuh"
#include "volume_gl2.cuh"
#endif

#ifndef FLEXI_NP
#error "compile with -DFLEXI_NP=<N+1>"
#endif

namespace fx {
namespace {
constexpr int NP = FLEXI_NP;
using D = Dim<NP>;

// ---- launch of an element kernel: one CTA of NP^3 threads per element, or (Dim::PAIR, N = 11) a cluster of two CTAs of NP^3/2
// threads whose shared memories together hold the element's work planes (layout.cuh: PairPtr) ----
template <class K>
cudaError_t elem_smem_attr(K kern, size_t bytes) {
  if (D::PAIR && bytes > 2 * sizeof(double) * (size_t)D::PAIR_DOUBLES) return cudaErrorInvalidValue;
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, D::PAIR ? (int)(sizeof(double) * D::PAIR_DOUBLES) : (int)bytes);
}
template <class K, class A>
cudaError_t elem_launch(K kern, int nBlocks, size_t bytes, cudaStream_t s, const A& a) {
  if (D::PAIR) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2u * (unsigned)nBlocks);
    cfg.blockDim = dim3(D::THREADS);
    cfg.dynamicSmemBytes = sizeof(double) * D::PAIR_DOUBLES;
    cfg.stream = s;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = 2;
    at.val.clusterDim.y = 1;
    at.val.clusterDim.z = 1;
    cfg.attrs = &at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, a);
  }
  kern<<<nBlocks, D::NP3, bytes, s>>>(a);
  return cudaGetLastError();
}

// k-block size of the Gauss-Lobatto element kernel: half a zeta column per thread for even NP, one node otherwise
constexpr int KPT = (NP % 2 == 0) ? NP / 2 : 1;
static_assert(KPT >= 1, "");

template <bool PARA, int NVL, int LIFT, bool GLN = false>
cudaError_t volume_setup1() {
  return elem_smem_attr(k_volume<NP, GLN, false, PARA, NVL, LIFT>, VolSmem<NP, GLN, false, PARA, LIFT>::BYTES);
}
template <bool PARA, int NVL, bool GLN = false>
cudaError_t volume_setup() {
  cudaError_t e = volume_setup1<PARA, NVL, 1, GLN>();
  if (e == cudaSuccess && PARA) e = volume_setup1<PARA, NVL, 2, GLN>();
  return e;
}
// split form on Gauss nodes: the generic element kernel with the two-point volume fluxes (DVolSurf = (D - 1/2 M^-1 V^T B V)^T)
template <bool PARA, int NVL, int LIFT, bool GLN = false>
cudaError_t volume_gs_setup1() {
  return elem_smem_attr(k_volume<NP, GLN, true, PARA, NVL, LIFT>, VolSmem<NP, GLN, true, PARA, LIFT>::BYTES);
}
template <bool PARA, int NVL, bool GLN = false>
cudaError_t volume_gs_launch(int lift, const VolArgs& a, int nBlocks, cudaStream_t s) {
  if (PARA && lift == 2) return elem_launch(k_volume<NP, GLN, true, PARA, NVL, 2>, nBlocks, VolSmem<NP, GLN, true, PARA, 2>::BYTES, s, a);
  return elem_launch(k_volume<NP, GLN, true, PARA, NVL, 1>, nBlocks, VolSmem<NP, GLN, true, PARA, 1>::BYTES, s, a);
}
template <bool PARA, int NVL, bool GLN = false>
cudaError_t volume_launch(int lift, const VolArgs& a, int nBlocks, cudaStream_t s) {
  if (PARA && lift == 2) return elem_launch(k_volume<NP, GLN, false, PARA, NVL, 2>, nBlocks, VolSmem<NP, GLN, false, PARA, 2>::BYTES, s, a);
  return elem_launch(k_volume<NP, GLN, false, PARA, NVL, 1>, nBlocks, VolSmem<NP, GLN, false, PARA, 1>::BYTES, s, a);
}
#if FLEXI_NP <= 10
template <bool SPLIT, int STYPE, bool PARA, int NVL>
cudaError_t volume_gl_setup() {
  using C = VolGLCfg<NP, KPT, SPLIT, PARA>;
  cudaError_t e = cudaFuncSetAttribute(k_volume_gl<NP, KPT, SPLIT, STYPE, PARA, NVL, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::BYTES);
  if (e == cudaSuccess && PARA)
    e = cudaFuncSetAttribute(k_volume_gl<NP, KPT, SPLIT, STYPE, PARA, NVL, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::BYTES);
  return e;
}
template <bool SPLIT, int STYPE, bool PARA, int NVL>
cudaError_t volume_gl_launch(int lift, const VolArgs& a, int nBlocks, cudaStream_t s) {
  using C = VolGLCfg<NP, KPT, SPLIT, PARA>;
  if (PARA && lift == 2)
    k_volume_gl<NP, KPT, SPLIT, STYPE, PARA, NVL, 2><<<nBlocks, C::NT, C::BYTES, s>>>(a);
  else
    k_volume_gl<NP, KPT, SPLIT, STYPE, PARA, NVL, 1><<<nBlocks, C::NT, C::BYTES, s>>>(a);
  return cudaGetLastError();
}

// line-resident split-form GL kernel (volume_gl2.cuh) for even NP; odd NP keeps k_volume_gl
constexpr bool HAS_GL2 = (NP % 2 == 0);
bool g_use_gl2 = HAS_GL2;
// measurement hook (development build only): FLEXI_GPU_SMEM_PAD=<bytes> of extra dynamic shared memory per CTA of k_volume_gl2 (forces 1 CTA/SM to
// separate what the second resident element buys from what contention costs; profiles/README.md)
#ifdef FLEXI_GPU_DEV
int gl2_smem_pad() {
  static int pad = -1;
  if (pad < 0) {
    const char* v = getenv("FLEXI_GPU_SMEM_PAD");
    pad = v ? atoi(v) : 0;
  }
  return pad;
}
#else
constexpr int gl2_smem_pad() { return 0; }
#endif
template <int STYPE, bool PARA, int NVL, int NVC, int LIFT, bool VVISC = false>
struct GL2 {
  static constexpr int NPE = HAS_GL2 ? NP : 2;  // odd NP never instantiates the real kernel
  using C = VolGL2Cfg<NPE, PARA, NVC>;
  static cudaError_t setup() {
    return cudaFuncSetAttribute(k_volume_gl2<NPE, STYPE, PARA, NVL, NVC, LIFT, VVISC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::BYTES + gl2_smem_pad());
  }
  static cudaError_t launch(const VolArgs& a, int nBlocks, cudaStream_t s) {
    VolArgs2<NPE> aa;
    aa.a = a;
    const double* b = a.basisHost;
    const double scale = (STYPE == FLEXI_SPLITDG_PI) ? 0.125 : 1.;
    for (int t = 0; t < NPE * NPE; ++t) {
      aa.b.dt[t] = b[Dim<NPE>::O_DT + t];
      aa.b.dh[t] = b[Dim<NPE>::O_DH + t];
      aa.b.dvs[t] = scale * b[Dim<NPE>::O_DV + t];
    }
    aa.b.lhm0 = b[Dim<NPE>::O_LHM];
    aa.b.lhpN = b[Dim<NPE>::O_LHP + NPE - 1];
    aa.b.lm0 = b[Dim<NPE>::O_LM];
    k_volume_gl2<NPE, STYPE, PARA, NVL, NVC, LIFT, VVISC><<<nBlocks, C::NT, C::BYTES + gl2_smem_pad(), s>>>(aa);
    return cudaGetLastError();
  }
};
// M(N-l,N-i) == -M(l,i) for the three operators (off-diagonal only for DVolSurf): what gl2_ap relies on
bool gl2_antisymmetric(const double* b) {
  const int offs[3] = {D::O_DT, D::O_DH, D::O_DV};
  for (int t = 0; t < 3; ++t) {
    double mx = 0.;
    for (int x = 0; x < NP * NP; ++x) mx = fmax(mx, fabs(b[offs[t] + x]));
    for (int l = 0; l < NP; ++l)
      for (int i = 0; i < NP; ++i) {
        if (t == 2 && l == i) continue;
        if (fabs(b[offs[t] + l * NP + i] + b[offs[t] + (NP - 1 - l) * NP + (NP - 1 - i)]) > 1e-14 * mx) return false;
      }
  }
  return true;
}
// all nVarLifting variables are lifted when gradients are handed out; the time-stepping path lifts (v1,v2,v3,T) only.
// PP_VISC != 0 has its own instantiations (Navier-Stokes only).
template <int STYPE, bool PARA, int NVL, bool VV>
cudaError_t gl2_setup_v() {
  cudaError_t e = GL2<STYPE, PARA, NVL, NVL, 1, VV>::setup();
  if (e == cudaSuccess && PARA) e = GL2<STYPE, PARA, NVL, NVL, 2, VV>::setup();
  if (e == cudaSuccess && PARA && NVL == 5) e = GL2<STYPE, PARA, NVL, 4, 1, VV>::setup();
  if (e == cudaSuccess && PARA && NVL == 5) e = GL2<STYPE, PARA, NVL, 4, 2, VV>::setup();
  return e;
}
template <int STYPE, bool PARA, int NVL>
cudaError_t gl2_setup() {
  cudaError_t e = gl2_setup_v<STYPE, PARA, NVL, false>();
  if (e == cudaSuccess && PARA) e = gl2_setup_v<STYPE, PARA, NVL, true>();
  return e;
}
template <int STYPE, bool PARA, int NVL, bool VV>
cudaError_t gl2_launch_v(int lift, const VolArgs& a, int nBlocks, cudaStream_t s) {
  const bool all = !PARA || NVL == 4 || a.fullGrad || a.gvol[0];
  if (PARA && lift == 2) return all ? GL2<STYPE, PARA, NVL, NVL, 2, VV>::launch(a, nBlocks, s) : GL2<STYPE, PARA, NVL, 4, 2, VV>::launch(a, nBlocks, s);
  return all ? GL2<STYPE, PARA, NVL, NVL, 1, VV>::launch(a, nBlocks, s) : GL2<STYPE, PARA, NVL, 4, 1, VV>::launch(a, nBlocks, s);
}
template <int STYPE, bool PARA, int NVL>
cudaError_t gl2_launch(int lift, const VolArgs& a, int nBlocks, cudaStream_t s) {
  if (PARA && a.q.visc != 0) return gl2_launch_v<STYPE, PARA, NVL, PARA>(lift, a, nBlocks, s);
  return gl2_launch_v<STYPE, PARA, NVL, false>(lift, a, nBlocks, s);
}

#endif  // FLEXI_NP <= 10

constexpr size_t upd_smem = sizeof(double) * (D::NBASIS + 5 * D::NP3);

#if defined(FLEXI_GPU_DEV) && defined(FX_DEV_ONLY_PI)  // development build: compile the Pirozzoli kernels only (minutes -> seconds)
#define FX_SPLIT_TYPES(X) X(FLEXI_SPLITDG_PI)
#else
#define FX_SPLIT_TYPES(X) \
  X(FLEXI_SPLITDG_SD) X(FLEXI_SPLITDG_MO) X(FLEXI_SPLITDG_DU) X(FLEXI_SPLITDG_KG) X(FLEXI_SPLITDG_PI) X(FLEXI_SPLITDG_CH) X(FLEXI_SPLITDG_IR)
#endif

cudaError_t setup() {
  cudaError_t e;
  if ((e = volume_setup<false, 5>()) != cudaSuccess) return e;
  if ((e = volume_setup<true, 5>()) != cudaSuccess) return e;
  if ((e = volume_setup<true, 4>()) != cudaSuccess) return e;
  if ((e = volume_gs_setup1<false, 5, 1>()) != cudaSuccess) return e;
  if ((e = volume_gs_setup1<true, 5, 1>()) != cudaSuccess) return e;
  if ((e = volume_gs_setup1<true, 5, 2>()) != cudaSuccess) return e;
  if ((e = volume_gs_setup1<true, 4, 1>()) != cudaSuccess) return e;
  if ((e = volume_gs_setup1<true, 4, 2>()) != cudaSuccess) return e;
#if FLEXI_NP > 10  // large degrees: Gauss-Lobatto nodes run the generic element kernel as well
  if ((e = volume_setup<false, 5, true>()) != cudaSuccess) return e;
  if ((e = volume_setup<true, 5, true>()) != cudaSuccess) return e;
  if ((e = volume_setup<true, 4, true>()) != cudaSuccess) return e;
  if ((e = volume_gs_setup1<false, 5, 1, true>()) != cudaSuccess) return e;
  if ((e = volume_gs_setup1<true, 5, 1, true>()) != cudaSuccess) return e;
  if ((e = volume_gs_setup1<true, 5, 2, true>()) != cudaSuccess) return e;
  if ((e = volume_gs_setup1<true, 4, 1, true>()) != cudaSuccess) return e;
  if ((e = volume_gs_setup1<true, 4, 2, true>()) != cudaSuccess) return e;
#else
  if ((e = volume_gl_setup<false, 0, false, 5>()) != cudaSuccess) return e;
  if ((e = volume_gl_setup<false, 0, true, 5>()) != cudaSuccess) return e;
  if ((e = volume_gl_setup<false, 0, true, 4>()) != cudaSuccess) return e;
#if FLEXI_NP % 2 == 1  // split form on odd NP: k_volume_gl; even NP: k_volume_gl2 below
#define X(T)                                                                     \
  if ((e = volume_gl_setup<true, T, false, 5>()) != cudaSuccess) return e;       \
  if ((e = volume_gl_setup<true, T, true, 5>()) != cudaSuccess) return e;        \
  if ((e = volume_gl_setup<true, T, true, 4>()) != cudaSuccess) return e;
  FX_SPLIT_TYPES(X)
#undef X
#endif
  if (HAS_GL2) {
#define X(T)                                                           \
  if ((e = gl2_setup<T, false, 5>()) != cudaSuccess) return e;         \
  if ((e = gl2_setup<T, true, 5>()) != cudaSuccess) return e;          \
  if ((e = gl2_setup<T, true, 4>()) != cudaSuccess) return e;
    FX_SPLIT_TYPES(X)
#undef X
  }
#endif  // FLEXI_NP > 10
#define X(GLv, ENTv, FILTv, SPv) \
  if ((e = elem_smem_attr(k_update<NP, GLv, ENTv, FILTv, SPv>, upd_smem)) != cudaSuccess) return e;
  X(true, false, false, false) X(false, false, false, false) X(false, true, false, false)
  X(true, false, true, false) X(false, false, true, false) X(false, true, true, false)
  X(true, false, false, true) X(false, false, false, true) X(false, true, false, true)
  X(true, false, true, true) X(false, false, true, true) X(false, true, true, true)
#undef X
  if ((e = elem_smem_attr(k_surfcorr<NP>, sizeof(double) * (D::NBASIS + 5 * D::NP3))) != cudaSuccess) return e;
  const size_t pro = sizeof(double) * (D::NBASIS + 6 * D::NP3);
  if ((e = elem_smem_attr(k_prolong<NP, true>, pro)) != cudaSuccess) return e;
  if ((e = elem_smem_attr(k_prolong<NP, false>, pro)) != cudaSuccess) return e;
  if (D::PAIR) {
    if ((e = elem_smem_attr(k_surfint<NP, true>, sizeof(double) * D::NBASIS)) != cudaSuccess) return e;
    if ((e = elem_smem_attr(k_surfint<NP, false>, sizeof(double) * D::NBASIS)) != cudaSuccess) return e;
    if ((e = elem_smem_attr(k_filter<NP>, sizeof(double) * 5 * D::NP3)) != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t prolong(bool gl, const ProArgs& a, int nElems, cudaStream_t s) {
  if (a.nv > 6) return cudaErrorInvalidValue;
  const size_t sm = sizeof(double) * (D::NBASIS + (size_t)a.nv * D::NP3);
  if (gl) return elem_launch(k_prolong<NP, true>, nElems, sm, s, a);
  return elem_launch(k_prolong<NP, false>, nElems, sm, s, a);
}

cudaError_t volume(bool gl, bool split, int stype, bool para, int nvl, const VolArgs& a, int nBlocks, cudaStream_t s) {
  if (nBlocks <= 0) return cudaSuccess;
  const int lift = a.q.lifting;
  if (split && !gl) {
    if (!para) return volume_gs_launch<false, 5>(lift, a, nBlocks, s);
    return nvl == 5 ? volume_gs_launch<true, 5>(lift, a, nBlocks, s) : volume_gs_launch<true, 4>(lift, a, nBlocks, s);
  }
  if (!gl) {
    if (!para) return volume_launch<false, 5>(lift, a, nBlocks, s);
    return nvl == 5 ? volume_launch<true, 5>(lift, a, nBlocks, s) : volume_launch<true, 4>(lift, a, nBlocks, s);
  }
#if FLEXI_NP > 10
  if (!split) {
    if (!para) return volume_launch<false, 5, true>(lift, a, nBlocks, s);
    return nvl == 5 ? volume_launch<true, 5, true>(lift, a, nBlocks, s) : volume_launch<true, 4, true>(lift, a, nBlocks, s);
  }
  if (!para) return volume_gs_launch<false, 5, true>(lift, a, nBlocks, s);
  return nvl == 5 ? volume_gs_launch<true, 5, true>(lift, a, nBlocks, s) : volume_gs_launch<true, 4, true>(lift, a, nBlocks, s);
#else
  if (!split) {
    if (!para) return volume_gl_launch<false, 0, false, 5>(lift, a, nBlocks, s);
    return nvl == 5 ? volume_gl_launch<false, 0, true, 5>(lift, a, nBlocks, s) : volume_gl_launch<false, 0, true, 4>(lift, a, nBlocks, s);
  }
  if (HAS_GL2 && g_use_gl2 && gl2_antisymmetric(a.basisHost)) {
    switch (stype) {
#define X(T)                                                                                   \
  case T:                                                                                      \
    if (!para) return gl2_launch<T, false, 5>(lift, a, nBlocks, s);                            \
    return nvl == 5 ? gl2_launch<T, true, 5>(lift, a, nBlocks, s) : gl2_launch<T, true, 4>(lift, a, nBlocks, s);
      FX_SPLIT_TYPES(X)
#undef X
      default: return cudaErrorNotSupported;
    }
  }
#if FLEXI_NP % 2 == 1
  switch (stype) {
#define X(T)                                                                                                     \
  case T:                                                                                                        \
    if (!para) return volume_gl_launch<true, T, false, 5>(lift, a, nBlocks, s);                                  \
    return nvl == 5 ? volume_gl_launch<true, T, true, 5>(lift, a, nBlocks, s) : volume_gl_launch<true, T, true, 4>(lift, a, nBlocks, s);
    FX_SPLIT_TYPES(X)
#undef X
    default: return cudaErrorNotSupported;
  }
#else
  return cudaErrorNotSupported;  // even NP and the D matrices are not centro-antisymmetric: not Gauss-Lobatto operators
#endif
#endif  // FLEXI_NP > 10
}

cudaError_t bc_pre(const BCArgs& a, cudaStream_t s) {
  const long n = (long)a.bc.nBCSides * D::NP2;
  if (n <= 0) return cudaSuccess;
  k_bc_pre<NP><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t flux_bc(const BCArgs& a, cudaStream_t s) {
  const long n = (long)a.bc.nBCSides * D::NP2;
  if (n <= 0) return cudaSuccess;
  k_flux_bc<NP><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t flux(const FluxArgs& a, cudaStream_t s) {
  if (a.nSidesRange <= 0) return cudaSuccess;
  const long n = (long)a.nSidesRange * D::NP2;
  if (a.q.riemann == FLEXI_RIEMANN_LF)
    k_flux<NP, FLEXI_RIEMANN_LF><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(a);
  else
    k_flux<NP, 0><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(a);
  return cudaGetLastError();
}

template <bool GL, bool ENT, bool FILT, bool SPONGE>
cudaError_t update_launch(const UpdArgs& a, int nBlocks, cudaStream_t s) {
  return elem_launch(k_update<NP, GL, ENT, FILT, SPONGE>, nBlocks, upd_smem, s, a);
}
template <bool FILT, bool SPONGE>
cudaError_t update_nodes(bool gl, const UpdArgs& a, int nBlocks, cudaStream_t s) {
  if (gl) return update_launch<true, false, FILT, SPONGE>(a, nBlocks, s);
  if (a.entropy) return update_launch<false, true, FILT, SPONGE>(a, nBlocks, s);
  return update_launch<false, false, FILT, SPONGE>(a, nBlocks, s);
}
cudaError_t update(bool gl, const UpdArgs& a, int nBlocks, cudaStream_t s) {
  if (nBlocks <= 0) return cudaSuccess;
  const bool filt = a.filt != nullptr, sponge = a.spIdx != nullptr;  // OverintegrationType = cutoff; doSponge
  if (filt) return sponge ? update_nodes<true, true>(gl, a, nBlocks, s) : update_nodes<true, false>(gl, a, nBlocks, s);
  return sponge ? update_nodes<false, true>(gl, a, nBlocks, s) : update_nodes<false, false>(gl, a, nBlocks, s);
}

cudaError_t surfcorr(const SurfCorrArgs& a, int nBlocks, cudaStream_t s) {
  if (nBlocks <= 0) return cudaSuccess;
  return elem_launch(k_surfcorr<NP>, nBlocks, sizeof(double) * (D::NBASIS + 5 * D::NP3), s, a);
}

cudaError_t surfint(bool gl, const SurfArgs& a, int nElems, cudaStream_t s) {
  const size_t sm = sizeof(double) * D::NBASIS;
  if (gl) return elem_launch(k_surfint<NP, true>, nElems, sm, s, a);
  return elem_launch(k_surfint<NP, false>, nElems, sm, s, a);
}

cudaError_t dt(const DtArgs& a, int nElems, cudaStream_t s) {
  k_dt<NP><<<nElems, dt_threads<NP>(), 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t filter(const FiltArgs& a, int nElems, cudaStream_t s) {
  if (nElems <= 0) return cudaSuccess;
  return elem_launch(k_filter<NP>, nElems, sizeof(double) * 5 * D::NP3, s, a);  // <= 40 KB at NP = 10: no opt-in needed
}

const NpOps ops = {NP, setup, prolong, volume, flux, update, surfcorr, surfint, dt, bc_pre, flux_bc, filter};
}  // namespace
}  // namespace fx

#if defined(FLEXI_GPU_DEV) && defined(FX_DEV_STAMPS) && FLEXI_NP == 8
// development build: read (and reset) the per-phase clock sums of k_volume_gl2 (tools/phase_stamps.py)
extern "C" int flexi_gpu_dev_stamps(unsigned long long* out) {
  cudaDeviceSynchronize();
  unsigned long long z[32] = {0};
  if (cudaMemcpyFromSymbol(out, fx::g_gl2_stamps, sizeof(z)) != cudaSuccess) return -1;
  return cudaMemcpyToSymbol(fx::g_gl2_stamps, z, sizeof(z)) == cudaSuccess ? 0 : -1;
}
#endif

#define FX_CAT2(a, b) a##b
#define FX_CAT(a, b) FX_CAT2(a, b)
extern "C" const fx::NpOps* FX_CAT(fx_ops_np, FLEXI_NP)() { return &fx::ops; }
