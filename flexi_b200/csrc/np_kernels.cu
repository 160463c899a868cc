// np_kernels.cu -- instantiates kernels.cuh for one polynomial degree (compile with -DFLEXI_NP=<N+1>).
#include "np_ops.h"

#ifndef FLEXI_NP
#error "compile with -DFLEXI_NP=<N+1>"
#endif

namespace fx {
namespace {
constexpr int NP = FLEXI_NP;
using D = Dim<NP>;

template <bool GL, bool SPLIT, bool PARA, int NVL>
cudaError_t volume_setup() {
  return cudaFuncSetAttribute(k_volume<NP, GL, SPLIT, PARA, NVL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)VolSmem<NP, GL, SPLIT, PARA>::BYTES);
}
template <bool GL, bool SPLIT, bool PARA, int NVL>
cudaError_t volume_launch(const VolArgs& a, int nBlocks, cudaStream_t s) {
  k_volume<NP, GL, SPLIT, PARA, NVL><<<nBlocks, D::NP3, VolSmem<NP, GL, SPLIT, PARA>::BYTES, s>>>(a);
  return cudaGetLastError();
}

constexpr size_t upd_smem = sizeof(double) * (D::NBASIS + 5 * D::NP3);

cudaError_t setup() {
  cudaError_t e;
#define S(GL, SP, PA, NV) \
  if ((e = volume_setup<GL, SP, PA, NV>()) != cudaSuccess) return e;
  S(false, false, false, 5) S(false, false, true, 5) S(false, false, true, 4)
  S(true, false, false, 5) S(true, false, true, 5) S(true, false, true, 4)
  S(true, true, false, 5) S(true, true, true, 5) S(true, true, true, 4)
#undef S
  if ((e = cudaFuncSetAttribute(k_update<NP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)upd_smem)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(k_update<NP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)upd_smem)) != cudaSuccess) return e;
  const int pro = (int)(sizeof(double) * (D::NBASIS + 6 * D::NP3));
  if ((e = cudaFuncSetAttribute(k_prolong<NP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pro)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(k_prolong<NP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pro)) != cudaSuccess) return e;
  return cudaSuccess;
}

cudaError_t prolong(bool gl, const ProArgs& a, int nElems, cudaStream_t s) {
  if (a.nv > 6) return cudaErrorInvalidValue;
  const size_t sm = sizeof(double) * (D::NBASIS + (size_t)a.nv * D::NP3);
  if (gl)
    k_prolong<NP, true><<<nElems, D::NP3, sm, s>>>(a);
  else
    k_prolong<NP, false><<<nElems, D::NP3, sm, s>>>(a);
  return cudaGetLastError();
}

cudaError_t volume(bool gl, bool split, bool para, int nvl, const VolArgs& a, int nBlocks, cudaStream_t s) {
  if (nBlocks <= 0) return cudaSuccess;
  if (split && !gl) return cudaErrorNotSupported;
#define L(GL, SP, PA, NV) return volume_launch<GL, SP, PA, NV>(a, nBlocks, s)
  if (!para) {
    if (!gl) L(false, false, false, 5);
    if (!split) L(true, false, false, 5);
    L(true, true, false, 5);
  }
  if (nvl == 5) {
    if (!gl) L(false, false, true, 5);
    if (!split) L(true, false, true, 5);
    L(true, true, true, 5);
  }
  if (!gl) L(false, false, true, 4);
  if (!split) L(true, false, true, 4);
  L(true, true, true, 4);
#undef L
}

cudaError_t flux(const FluxArgs& a, cudaStream_t s) {
  if (a.nSidesRange <= 0) return cudaSuccess;
  const long n = (long)a.nSidesRange * D::NP2;
  k_flux<NP><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t update(bool gl, const UpdArgs& a, int nBlocks, cudaStream_t s) {
  if (nBlocks <= 0) return cudaSuccess;
  if (gl)
    k_update<NP, true><<<nBlocks, D::NP3, upd_smem, s>>>(a);
  else
    k_update<NP, false><<<nBlocks, D::NP3, upd_smem, s>>>(a);
  return cudaGetLastError();
}

cudaError_t surfint(bool gl, const SurfArgs& a, int nElems, cudaStream_t s) {
  const size_t sm = sizeof(double) * D::NBASIS;
  if (gl)
    k_surfint<NP, true><<<nElems, D::NP3, sm, s>>>(a);
  else
    k_surfint<NP, false><<<nElems, D::NP3, sm, s>>>(a);
  return cudaGetLastError();
}

cudaError_t dt(const DtArgs& a, int nElems, cudaStream_t s) {
  k_dt<NP><<<nElems, D::NP3, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t dt_init(const DtInitArgs& a, cudaStream_t s) {
  k_dt_init<NP><<<a.nElems, D::NP3, 0, s>>>(a);
  return cudaGetLastError();
}

const NpOps ops = {NP, setup, prolong, volume, flux, update, surfint, dt, dt_init};
}  // namespace
}  // namespace fx

#define FX_CAT2(a, b) a##b
#define FX_CAT(a, b) FX_CAT2(a, b)
extern "C" const fx::NpOps* FX_CAT(fx_ops_np, FLEXI_NP)() { return &fx::ops; }
