// analyze.cuh -- AnalyzeTestcase of the Taylor-Green vortex test case on the device, steps 1-4 of
// /root/reference/src/testcase/taylorgreenvortex/testcase.f90:241-420 (nullify, interpolate U / sJ / GradVel to the analyze
// mesh with Vdm_GaussN_NAnalyze (ChangeBasis3D), ConsToPrim + gradient-based quantities, integrate with wGPVolAnalyze/sJ).
// Steps 5-7 (MPI_REDUCE, normalisation with Vol and rho0, file output) stay in the host.  SURVEY.md section 8(f) rank 1: the
// analysis runs every nAnalyzeTestCase (10) steps and otherwise forces U and three gradient arrays (4.5 GB at config 2) over
// PCIe each time; here 12 doubles come back.
#pragma once
#include "layout.cuh"

namespace fx {

enum { TGV_T_MEAN = 0, TGV_ENTROPY, TGV_EKIN, TGV_EKIN_COMP, TGV_ENSTR, TGV_DR_U, TGV_DR_S, TGV_DR_SD, TGV_DR_P, TGV_ED_S, TGV_ED_D,
       TGV_VORT_MAX, TGV_NSUMS };

struct TgvArgs {
  Eqn q;
  int np, na1, nvl, parabolic;  // N+1, NAnalyze+1, nVarLifting
  const double* U;         // [e][5][NP3]
  const double* sJ;        // [e][NP3]
  const double* gvol[3];   // gradU{x,y,z} [e][nvl][NP3]
  const double* vdm;       // Vdm_GaussN_NAnalyze as [iOut][iIn]
  const double* wVol;      // wGPVolAnalyze [k][j][i]
  double* out;             // [e][TGV_NSUMS]
};

// ChangeBasis3D (changeBasis.t90) of nf fields of one element to the analyze mesh, one zeta slab of the analyze mesh at a time:
// zeta pass for slab k, xi pass, then the eta pass delivers the thread's analyze point (iO,jO,k) for every field.  The caller
// must __syncthreads() before the next slab (sT1 / sT2 are reused).
__device__ __forceinline__ void interp_slab(int nf, int np, int na1, int k, const double* sX, const double* sV, double* sT1,
                                            double* sT2, double* val) {
  const int np2 = np * np, np3 = np2 * np, tid = threadIdx.x, nt = blockDim.x;
  const int iO = tid % na1, jO = tid / na1;
  for (int t = tid; t < nf * np2; t += nt) {  // zeta pass
    const int f = t / np2, ij = t % np2;
    double s = 0.;
    for (int l = 0; l < np; ++l) s = s + sV[k * np + l] * sX[f * np3 + ij + np2 * l];
    sT1[t] = s;
  }
  __syncthreads();
  for (int t = tid; t < nf * na1 * np; t += nt) {  // xi pass
    const int f = t / (na1 * np), r = t % (na1 * np), io = r % na1, j = r / na1;
    double s = 0.;
    for (int l = 0; l < np; ++l) s = s + sV[io * np + l] * sT1[f * np2 + l + np * j];
    sT2[t] = s;
  }
  __syncthreads();
  for (int f = 0; f < nf; ++f) {  // eta pass: the thread's analyze point
    double s = 0.;
    for (int l = 0; l < np; ++l) s = s + sV[jO * np + l] * sT2[f * na1 * np + iO + na1 * l];
    val[f] = s;
  }
}

// one CTA per element, one thread per (i,j) of the analyze mesh; the analyze mesh is walked slab by slab in zeta so that a thread
// holds one analyze point (15 interpolated fields) at a time
__global__ void k_tgv_analyze(TgvArgs A) {
  extern __shared__ double sm[];
  const int np = A.np, na1 = A.na1, np2 = np * np, np3 = np2 * np;
  const int nf = A.parabolic ? 15 : 6;  // sJ, U(5), GradVel(3x3)
  double* sX = sm;                      // [nf][np3]
  double* sV = sX + nf * np3;           // [na1][np]
  double* sT1 = sV + na1 * np;          // [nf][np2]      zeta pass of the current slab
  double* sT2 = sT1 + nf * np2;         // [nf][na1*np]   + xi pass
  double* sRed = sT2 + nf * na1 * np;   // [TGV_NSUMS][32]
  const int tid = threadIdx.x, nt = blockDim.x, e = blockIdx.x;
  const int v0 = (A.nvl == 5) ? 1 : 0;  // LIFT_VELV
  for (int t = tid; t < np3; t += nt) {
    sX[t] = A.sJ[(size_t)e * np3 + t];
    for (int v = 0; v < 5; ++v) sX[(1 + v) * np3 + t] = A.U[((size_t)e * 5 + v) * np3 + t];
    if (A.parabolic)
      for (int d = 0; d < 3; ++d)
        for (int c = 0; c < 3; ++c) sX[(6 + 3 * d + c) * np3 + t] = A.gvol[d][((size_t)e * A.nvl + v0 + c) * np3 + t];
  }
  for (int t = tid; t < na1 * np; t += nt) sV[t] = A.vdm[t];
  __syncthreads();
  const int iO = tid % na1, jO = tid / na1;
  const Eqn& q = A.q;
  double acc[TGV_NSUMS];
  for (int s = 0; s < TGV_NSUMS; ++s) acc[s] = 0.;
  for (int k = 0; k < na1; ++k) {
    double val[15];
    interp_slab(nf, np, na1, k, sX, sV, sT1, sT2, val);
    double P[6];
    cons_to_prim(q, val + 1, P);
    const double w = A.wVol[iO + na1 * (jO + na1 * k)] / val[0];  // IntFactor, testcase.f90:381
    const double v2 = P[VEL1] * P[VEL1] + P[VEL2] * P[VEL2] + P[VEL3] * P[VEL3];
    acc[TGV_T_MEAN] += w * P[TEMP];
    acc[TGV_ENTROPY] += w * (-q.sKappaM1) * P[DENS] * (log(P[PRES]) - q.kappa * log(P[DENS]));
    acc[TGV_EKIN] += w * 0.5 * v2;
    acc[TGV_EKIN_COMP] += w * 0.5 * P[DENS] * v2;
    if (A.parabolic) {
      // GradVel(c,d) = d v_c / d x_d = val[6 + 3*d + c]   (testcase.f90:352-354)
      double G[3][3];
      for (int d = 0; d < 3; ++d)
        for (int c = 0; c < 3; ++c) G[c][d] = val[6 + 3 * d + c];
      const double divU = G[0][0] + G[1][1] + G[2][2];
      double sumG = 0., sumS = 0., sumSd = 0.;
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
          const double S = 0.5 * (G[a][b] + G[b][a]);
          const double Sd = (a == b) ? S - 1. / 3. * divU : S;
          sumG += G[a][b] * G[a][b];
          sumS += S * S;
          sumSd += Sd * Sd;
        }
      const double o1 = G[2][1] - G[1][2], o2 = G[0][2] - G[2][0], o3 = G[1][0] - G[0][1];
      const double oo = o1 * o1 + o2 * o2 + o3 * o3;
      acc[TGV_VORT_MAX] = fmax(acc[TGV_VORT_MAX], sqrt(oo));
      acc[TGV_ENSTR] += w * 0.5 * P[DENS] * oo;
      acc[TGV_DR_U] += w * sumG;
      acc[TGV_DR_S] += w * sumS;
      acc[TGV_DR_SD] += w * sumSd;
      acc[TGV_DR_P] += w * P[PRES] * divU;
      const double muT = viscosity(q, P[TEMP]);  // VISCOSITY_TEMPERATURE(UPrim(TEMP)), testcase.f90:406-408
      acc[TGV_ED_S] += w * muT * oo;
      acc[TGV_ED_D] += w * muT * divU * divU;
    }
    __syncthreads();  // the next slab overwrites sT1 / sT2
  }
  // block reduction: sums in a fixed order (deterministic), the vorticity as a max
  const int lane = tid & 31, wid = tid >> 5, nw = (nt + 31) >> 5;
  for (int s = 0; s < TGV_NSUMS; ++s) {
    double v = acc[s];
    for (int o = 16; o > 0; o >>= 1) {
      const double u = __shfl_down_sync(0xffffffffu, v, o);
      v = (s == TGV_VORT_MAX) ? fmax(v, u) : v + u;
    }
    if (lane == 0) sRed[s * 32 + wid] = v;
  }
  __syncthreads();
  if (tid < TGV_NSUMS) {
    double v = sRed[tid * 32];
    for (int w2 = 1; w2 < nw; ++w2) v = (tid == TGV_VORT_MAX) ? fmax(v, sRed[tid * 32 + w2]) : v + sRed[tid * 32 + w2];
    A.out[(size_t)e * TGV_NSUMS + tid] = v;
  }
}

// CalcErrorNorms, analyze/analyze.f90:391-485, per-element part: interpolate U, Elem_xGP and J = 1/sJ to the analyze mesh,
// evaluate ExactFunc there, accumulate (U-U_exact)^2 * wGPVolAnalyze * J and max|U-U_exact| per variable.
struct ErrArgs {
  Eqn q;
  BCData bc;   // carries IniExactFunc and its parameters
  double t;
  int np, na1;
  const double* U;    // [e][5][NP3]
  const double* sJ;   // [e][NP3]
  const double* xGP;  // [e][3][NP3]
  const double* vdm;
  const double* wVol;
  double* out;        // [e][10]: L2 sums (5), Linf (5)
};

__global__ void k_error_norms(ErrArgs A) {
  extern __shared__ double sm[];
  const int np = A.np, na1 = A.na1, np2 = np * np, np3 = np2 * np;
  const int nf = 9;  // J, U(5), x(3)
  double* sX = sm;
  double* sV = sX + nf * np3;
  double* sT1 = sV + na1 * np;
  double* sT2 = sT1 + nf * np2;
  double* sRed = sT2 + nf * na1 * np;  // [10][32]
  const int tid = threadIdx.x, nt = blockDim.x, e = blockIdx.x;
  for (int t = tid; t < np3; t += nt) {
    sX[t] = 1. / A.sJ[(size_t)e * np3 + t];  // J_N, analyze.f90:437
    for (int v = 0; v < 5; ++v) sX[(1 + v) * np3 + t] = A.U[((size_t)e * 5 + v) * np3 + t];
    for (int c = 0; c < 3; ++c) sX[(6 + c) * np3 + t] = A.xGP[((size_t)e * 3 + c) * np3 + t];
  }
  for (int t = tid; t < na1 * np; t += nt) sV[t] = A.vdm[t];
  __syncthreads();
  const int iO = tid % na1, jO = tid / na1;
  double l2[5] = {0., 0., 0., 0., 0.}, li[5] = {0., 0., 0., 0., 0.};
  for (int k = 0; k < na1; ++k) {
    double val[15];
    interp_slab(nf, np, na1, k, sX, sV, sT1, sT2, val);
    double Uex[5];
    exact_func(A.q, A.bc, A.t, val + 6, Uex);
    const double w = A.wVol[iO + na1 * (jO + na1 * k)] * val[0];
    for (int v = 0; v < 5; ++v) {
      const double d = val[1 + v] - Uex[v];
      li[v] = fmax(li[v], fabs(d));
      l2[v] += d * d * w;
    }
    __syncthreads();
  }
  const int lane = tid & 31, wid = tid >> 5, nw = (nt + 31) >> 5;
  for (int s = 0; s < 10; ++s) {
    double v = s < 5 ? l2[s] : li[s - 5];
    for (int o = 16; o > 0; o >>= 1) {
      const double u = __shfl_down_sync(0xffffffffu, v, o);
      v = s < 5 ? v + u : fmax(v, u);
    }
    if (lane == 0) sRed[s * 32 + wid] = v;
  }
  __syncthreads();
  if (tid < 10) {
    double v = sRed[tid * 32];
    for (int w2 = 1; w2 < nw; ++w2) v = tid < 5 ? v + sRed[tid * 32 + w2] : fmax(v, sRed[tid * 32 + w2]);
    A.out[(size_t)e * 10 + tid] = v;
  }
}

// sum (first nSum entries) / max (the rest) over the elements of the rank, fixed order
__global__ void k_norm_reduce(const double* in, int nElems, int nVal, int nSum, double* out) {
  __shared__ double s[256];
  const int tid = threadIdx.x;
  for (int q = 0; q < nVal; ++q) {
    double v = 0.;
    for (int e = tid; e < nElems; e += 256) v = (q >= nSum) ? fmax(v, in[(size_t)e * nVal + q]) : v + in[(size_t)e * nVal + q];
    s[tid] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (tid < o) s[tid] = (q >= nSum) ? fmax(s[tid], s[tid + o]) : s[tid] + s[tid + o];
      __syncthreads();
    }
    if (tid == 0) out[q] = s[0];
    __syncthreads();
  }
}

// CalcBodyForces, equations/navierstokes/calcbodyforces.f90:46-250, per BC side: pressure force / moment (CalcPressureForce
// :117-151) and viscous force / moment (CalcViscousForce :158-240) integrated over the side with wGPSurf*SurfElem.  One CTA per BC
// side; the host adds the sides of each wall BC in SideID order like the reference loop (:75-100).
struct BodyForceArgs {
  Eqn q;
  int np, nvl, parabolic;
  const double* U_master;  // [s][5][NP2]
  const double* gm[3];     // gradU{x,y,z}_master [s][nvl][NP2]
  const double* Geo;       // [s][10][NP2]
  const double* faceX;     // Face_xGP of the BC sides [s][3][NP2] (NULL: moments are not computed)
  const double* wSurf;     // wGPSurf [NP2]
  double momOrigin[3];
  double* out;             // [s][12]: Fp(3), Fv(3), Mp(3), Mv(3)
};

__global__ void k_body_forces(BodyForceArgs A) {
  __shared__ double sRed[12][32];
  const int np2 = A.np * A.np, tid = threadIdx.x, s = blockIdx.x;
  double acc[12];
  for (int c = 0; c < 12; ++c) acc[c] = 0.;
  if (tid < np2) {
    const int pq = tid;
    double Uc[5], P[6];
    for (int v = 0; v < 5; ++v) Uc[v] = A.U_master[((size_t)s * 5 + v) * np2 + pq];
    cons_to_prim(A.q, Uc, P);  // UPrim_master, equation.f90:189-216
    const double* g = A.Geo + (size_t)s * 10 * np2 + pq;
    const double nv[3] = {g[0], g[np2], g[2 * np2]};
    const double dA = A.wSurf[pq] * g[9 * np2];
    double r[3] = {0., 0., 0.};
    if (A.faceX)
      for (int c = 0; c < 3; ++c) r[c] = A.faceX[((size_t)s * 3 + c) * np2 + pq] - A.momOrigin[c];
    double df[3];
    for (int c = 0; c < 3; ++c) {
      df[c] = P[PRES] * nv[c] * dA;
      acc[c] = df[c];
    }
    acc[6] = r[1] * df[2] - r[2] * df[1];
    acc[7] = r[2] * df[0] - r[0] * df[2];
    acc[8] = r[0] * df[1] - r[1] * df[0];
    if (A.parabolic) {
      const int v0 = (A.nvl == 5) ? 1 : 0;  // LIFT_VELV
      double G[3][3];                        // GradV(c,d) = d v_c / d x_d
      for (int d = 0; d < 3; ++d)
        for (int c = 0; c < 3; ++c) G[c][d] = A.gm[d][((size_t)s * A.nvl + v0 + c) * np2 + pq];
      const double muS = viscosity(A.q, P[TEMP]), divV = G[0][0] + G[1][1] + G[2][2];
      double tau[3][3];
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) tau[a][b] = muS * (G[a][b] + G[b][a]);
      for (int a = 0; a < 3; ++a) tau[a][a] = tau[a][a] - 2. / 3. * muS * divV;
      double dfv[3];
      for (int a = 0; a < 3; ++a) dfv[a] = (tau[a][0] * nv[0] + tau[a][1] * nv[1] + tau[a][2] * nv[2]) * A.wSurf[pq] * g[9 * np2];
      // Fv = -Fv, Mv = -Mv: the force acting on the wall (calcbodyforces.f90:236-237)
      for (int c = 0; c < 3; ++c) acc[3 + c] = -dfv[c];
      acc[9] = -(r[1] * dfv[2] - r[2] * dfv[1]);
      acc[10] = -(r[2] * dfv[0] - r[0] * dfv[2]);
      acc[11] = -(r[0] * dfv[1] - r[1] * dfv[0]);
    }
  }
  const int lane = tid & 31, wid = tid >> 5, nw = (blockDim.x + 31) >> 5;
  for (int c = 0; c < 12; ++c) {
    double v = acc[c];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sRed[c][wid] = v;
  }
  __syncthreads();
  if (tid < 12) {
    double v = sRed[tid][0];
    for (int w = 1; w < nw; ++w) v += sRed[tid][w];
    A.out[(size_t)s * 12 + tid] = v;
  }
}

}  // namespace fx
