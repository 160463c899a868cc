// fp64_pipes.cu -- B200 micro-benchmark: do DFMA (FP64 vector pipe) and DMMA m8n8k4 (FP64 tensor path) share an issue pipe?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu ; run on one GPU, prints instr/clk/SM.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NF, int NM>  // NF DFMA and NM DMMA per loop iteration, all independent chains
__global__ void __launch_bounds__(256) k(double* out, int iters, double x) {
  double f[8], c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { f[i] = threadIdx.x * 1e-3 + i; c[i][0] = i; c[i][1] = -i; }
  const double a = x * 1.000001, b = x * 0.999999;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (u < NF) f[u] = fma(f[u], a, b);
      if (u < NM) dmma(c[u][0], c[u][1], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += f[i] + c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NF, int NM>
void run(const char* name, double* d, int nsm) {
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int warps = 4; warps <= 16; warps *= 2) {  // warps per SM: 256-thread CTAs x (warps/8) .. use CTAs of 128 threads
    const int ctas = nsm * warps / 4;
    k<NF, NM><<<ctas, 128>>>(d, 100, 1.0);
    cudaEventRecord(e0);
    k<NF, NM><<<ctas, 128>>>(d, iters, 1.0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cyc = ms * 1e-3 * clk * 1e3;
    const double winst = (double)warps * iters;  // per SM, per kind-unit
    printf("%-18s warps/SM %2d  %.3f ms  DFMA warp-inst/clk/SM %.3f  DMMA warp-inst/clk/SM %.3f  (FMA-equivalents/clk/SM %.1f)\n", name, warps, ms,
           NF * winst / cyc, NM * winst / cyc, (NF * 32 + NM * 256) * winst / cyc);
  }
}

int main() {
  int nsm;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  double* d;
  cudaMalloc(&d, (size_t)nsm * 16 * 128 * 8 * 8);
  run<8, 0>("DFMA x8", d, nsm);
  run<0, 8>("DMMA x8", d, nsm);
  run<0, 4>("DMMA x4", d, nsm);
  run<8, 8>("DFMA x8 + DMMA x8", d, nsm);
  run<8, 2>("DFMA x8 + DMMA x2", d, nsm);
  run<4, 4>("DFMA x4 + DMMA x4", d, nsm);
  printf("cuda error: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
