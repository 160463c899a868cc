// Test harness (never shipped): compiles the product's point functions of flexi_b200/csrc/physics.cuh -- which are
// __host__ __device__ -- as HOST code, so that the CPU test-suite can compare the very source the CUDA kernels inline
// (two-point fluxes, Riemann solvers) against the oracle without a GPU.  Built by tests/test_host_physics.py with nvcc.
#include "../flexi_b200/csrc/physics.cuh"

namespace {
fx::Eqn make_eqn(double kappa, double R, int splitDG, int splitType, int riemann) {
  fx::Eqn q{};
  q.kappa = kappa;
  q.kappaM1 = kappa - 1.;
  q.sKappaM1 = 1. / q.kappaM1;
  q.R = R;
  q.sR = 1. / R;
  q.splitDG = splitDG;
  q.splitType = splitType;
  q.riemann = riemann;
  return q;
}
template <int T>
void vol(const fx::Eqn& q, const double* URef, const double* U, const double* MRef, const double* M, double* Flux) {
  double PR[6], P[6];
  fx::cons_to_prim(q, URef, PR);
  fx::cons_to_prim(q, U, P);
  const fx::Node a{PR[0], PR[1], PR[2], PR[3], PR[4], URef[4]}, b{P[0], P[1], P[2], P[3], P[4], U[4]};
  fx::split_volume_flux<T>(a, b, MRef, M, Flux, q.kappa);
}
}  // namespace

extern "C" {
int hp_split_volume_flux(int type, double kappa, double R, const double* URef, const double* U, const double* MRef, const double* M,
                         double* Flux) {
  const fx::Eqn q = make_eqn(kappa, R, 1, type, FLEXI_RIEMANN_LF);
  switch (type) {
    case FLEXI_SPLITDG_SD: vol<FLEXI_SPLITDG_SD>(q, URef, U, MRef, M, Flux); return 0;
    case FLEXI_SPLITDG_MO: vol<FLEXI_SPLITDG_MO>(q, URef, U, MRef, M, Flux); return 0;
    case FLEXI_SPLITDG_DU: vol<FLEXI_SPLITDG_DU>(q, URef, U, MRef, M, Flux); return 0;
    case FLEXI_SPLITDG_KG: vol<FLEXI_SPLITDG_KG>(q, URef, U, MRef, M, Flux); return 0;
    case FLEXI_SPLITDG_PI: vol<FLEXI_SPLITDG_PI>(q, URef, U, MRef, M, Flux); return 0;
    case FLEXI_SPLITDG_CH: vol<FLEXI_SPLITDG_CH>(q, URef, U, MRef, M, Flux); return 0;
    case FLEXI_SPLITDG_IR: vol<FLEXI_SPLITDG_IR>(q, URef, U, MRef, M, Flux); return 0;
  }
  return -1;
}
int hp_riemann_point(double kappa, double R, int splitDG, int splitType, int solver, const double* U_L, const double* U_R, const double* nv,
                     const double* t1, const double* t2, double* F) {
  const fx::Eqn q = make_eqn(kappa, R, splitDG, splitType, solver);
  double PL[6], PR[6];
  fx::cons_to_prim(q, U_L, PL);
  fx::cons_to_prim(q, U_R, PR);
  fx::riemann_point(q, solver, U_L, U_R, PL, PR, nv, t1, t2, F);
  return 0;
}
}
