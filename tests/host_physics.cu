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
// gv = (d/dx, d/dy, d/dz) of (v1, v2, v3, T), 4 entries each; f,g,h = rows MOM1..ENER of the viscous flux
int hp_eval_diff_flux3d(double kappa, double R, double mu0, double Pr, const double* prim, const double* gx, const double* gy, const double* gz,
                        double* f, double* g, double* h) {
  fx::Eqn q = make_eqn(kappa, R, 0, 0, FLEXI_RIEMANN_LF);
  q.mu0 = mu0;
  const double cp = R * kappa / (kappa - 1.);  // eos.f90:160-163
  q.lambda = mu0 * cp / Pr;
  fx::eval_diff_flux3d(q, prim[5], prim[1], prim[2], prim[3], gx, gy, gz, f, g, h);
  return 0;
}
int hp_prim_to_entropy(double kappa, double R, const double* prim, double* w) {
  const fx::Eqn q = make_eqn(kappa, R, 1, 0, FLEXI_RIEMANN_LF);
  fx::prim_to_entropy(q, prim, w);
  return 0;
}
int hp_entropy_to_cons(double kappa, double R, const double* w, double* cons) {
  const fx::Eqn q = make_eqn(kappa, R, 1, 0, FLEXI_RIEMANN_LF);
  fx::entropy_to_cons(q, w, cons);
  return 0;
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
