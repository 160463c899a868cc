// common_kernels.cuh -- degree-independent kernels (included by flexi_gpu.cu only).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace fx {

// reduction of the per-element time steps: min(dtConv), min(dtVisc), errType (calctimestep.f90:187-222), in two levels: each
// block of the first launch reduces a slice of the elements to one (dtConv, dtVisc, err) triple, a single block finishes
// (one block over 110 592 elements took 0.19 ms per time step)
__device__ __forceinline__ void dt_block_reduce(double a, double b, double c, double* out3, bool negErr) {
  __shared__ double s0[256], s1[256], s2[256];
  const int tid = threadIdx.x;
  s0[tid] = a;
  s1[tid] = b;
  s2[tid] = c;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) {
      s0[tid] = fmin(s0[tid], s0[tid + o]);
      s1[tid] = fmin(s1[tid], s1[tid + o]);
      s2[tid] = fmax(s2[tid], s2[tid + o]);
    }
    __syncthreads();
  }
  if (tid == 0) {
    out3[0] = s0[0];
    out3[1] = s1[0];
    out3[2] = negErr ? -s2[0] : s2[0];  // the reference reduces -errType with MIN (calctimestep.f90:217-222)
  }
}
// level 1: per-element triples (dtConv, dtVisc, flag) -> per-block triples (min, min, errType)
__global__ void k_dt_reduce1(const double* in, int nElems, double* part) {
  double a = HUGE_VAL, b = HUGE_VAL, c = 0.;
  for (int e = blockIdx.x * 256 + threadIdx.x; e < nElems; e += gridDim.x * 256) {
    const double x = in[(size_t)e * 3], y = in[(size_t)e * 3 + 1], z = in[(size_t)e * 3 + 2];
    if (x != x) c = fmax(c, 2.);
    if (y != y) c = fmax(c, 3.);
    if (z == 1.) c = fmax(c, 1.);
    else if (z > 0.) c = fmax(c, 2.);
    a = fmin(a, x);
    b = fmin(b, y);
  }
  dt_block_reduce(a, b, c, part + 3 * blockIdx.x, false);
}
// level 2: per-block triples -> (dtConv, dtVisc, -errType)
__global__ void k_dt_reduce2(const double* part, int nPart, double* out3) {
  double a = HUGE_VAL, b = HUGE_VAL, c = 0.;
  for (int i = threadIdx.x; i < nPart; i += 256) {
    a = fmin(a, part[3 * i]);
    b = fmin(b, part[3 * i + 1]);
    c = fmax(c, part[3 * i + 2]);
  }
  dt_block_reduce(a, b, c, out3, true);
}

// ---------------------------------------------------------------------------------------------------------------------------
// layout conversion at the ABI boundary: host AoS block [item][n][v] <-> device SoA [item][v][n]
__global__ void k_aos_to_soa(const double* __restrict__ src, double* __restrict__ dst, int nv, int npts, long nItems) {
  const long total = nItems * nv * npts;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const long item = idx / ((long)nv * npts);
    const int r = (int)(idx % ((long)nv * npts));
    const int v = r / npts, n = r % npts;
    dst[idx] = src[(item * npts + n) * nv + v];
  }
}
__global__ void k_soa_to_aos(const double* __restrict__ src, double* __restrict__ dst, int nv, int npts, long nItems) {
  const long total = nItems * nv * npts;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const long item = idx / ((long)nv * npts);
    const int r = (int)(idx % ((long)nv * npts));
    const int n = r / nv, v = r % nv;
    dst[idx] = src[(item * nv + v) * npts + n];
  }
}

}  // namespace fx
