// np_ops.h -- per-polynomial-degree launch table.  kernels.cuh is instantiated once per NP = N+1 in its own translation
// unit (np_kernels.cu compiled with -DFLEXI_NP=<NP>), so the degrees build in parallel; the host picks the table at init
// (the reference fixes PP_N at compile time, src/CMakeLists.txt:23-60 -- here it is a run-time choice among N = 1..9 and 11).
#pragma once
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace fx {

struct NpOps {
  int np;
  cudaError_t (*setup)();
  cudaError_t (*prolong)(bool gl, const ProArgs& a, int nElems, cudaStream_t s);
  cudaError_t (*volume)(bool gl, bool split, int stype, bool para, int nvl, const VolArgs& a, int nBlocks, cudaStream_t s);
  cudaError_t (*flux)(const FluxArgs& a, cudaStream_t s);
  cudaError_t (*update)(bool gl, const UpdArgs& a, int nBlocks, cudaStream_t s);
  cudaError_t (*surfcorr)(const SurfCorrArgs& a, int nBlocks, cudaStream_t s);
  cudaError_t (*surfint)(bool gl, const SurfArgs& a, int nElems, cudaStream_t s);
  cudaError_t (*dt)(const DtArgs& a, int nElems, cudaStream_t s);
  cudaError_t (*bc_pre)(const BCArgs& a, cudaStream_t s);
  cudaError_t (*flux_bc)(const BCArgs& a, cudaStream_t s);
  cudaError_t (*filter)(const FiltArgs& a, int nElems, cudaStream_t s);
};

}  // namespace fx

#define FX_DECLARE_OPS(NP) extern "C" const fx::NpOps* fx_ops_np##NP();
FX_DECLARE_OPS(2)
FX_DECLARE_OPS(3)
FX_DECLARE_OPS(4)
FX_DECLARE_OPS(5)
FX_DECLARE_OPS(6)
FX_DECLARE_OPS(7)
FX_DECLARE_OPS(8)
FX_DECLARE_OPS(9)
FX_DECLARE_OPS(10)
FX_DECLARE_OPS(12)
