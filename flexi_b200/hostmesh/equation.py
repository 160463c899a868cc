"""Equation-system / time-disc parameters of the synthetic FLEXI host (stand-in for InitEquation, InitTimeDisc, FillIni).

Reference (under /root/reference/src):
  equations/navierstokes/idealgas/eos.f90:110-190          InitEos (kappa, R, Pr, mu0; cp, cv)
  timedisc/timedisc_vars.f90:109-214                        SetTimeDiscCoefs (carpenterrk4-5 and friends)
  timedisc/timedisc_func.f90:463-527                        FillCFL_DFL
  testcase/taylorgreenvortex/testcase.f90:102-107,176-219   InitTestcase / ExactFuncTestcase
  equations/navierstokes/idealgas/exactfunc.f90:332-362     ExactFunc case 4 (oblique sine wave)
"""
import numpy as np

# timedisc_vars.f90:175-214
RK_TABLES = {
    "carpenterrk4-5": dict(
        RKA=[567301805773.0 / 1357537059087.0, 2404267990393.0 / 2016746695238.0, 3550918686646.0 / 2091501179385.0,
             1275806237668.0 / 842570457699.0],
        RKb=[1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0, 1720146321549.0 / 2090206949498.0,
             3134564353537.0 / 4481467310338.0, 2277821191437.0 / 14882151754819.0],
        RKc=[1432997174477.0 / 9575080441755.0, 2526269341429.0 / 6820363962896.0, 2006345519317.0 / 3224310063776.0,
             2802321613138.0 / 2924317926251.0],
        CFLScaleAlpha_G=[2.0351, 1.7595, 1.5401, 1.3702, 1.2375, 1.1318, 1.0440, 0.9709, 0.9079, 0.8539, 0.8066, 0.7650,
                         0.7290, 0.6952, 0.6660],
        CFLScaleAlpha_GL=[4.7497, 3.4144, 2.8451, 2.4739, 2.2027, 1.9912, 1.8225, 1.6830, 1.5682, 1.4692, 1.3849, 1.3106,
                          1.2454, 1.1880, 1.1362],
        RelativeDFL=1.853,
    ),
    # timedisc_vars.f90:137-168 Williamson RK3-3
    "standardrk3-3": dict(
        RKA=[5.0 / 9.0, 153.0 / 128.0],
        RKb=[1.0 / 3.0, 15.0 / 16.0, 8.0 / 15.0],
        RKc=[1.0 / 3.0, 0.75],
        CFLScaleAlpha_G=[1.2285, 1.0485, 0.9101, 0.8066, 0.7268, 0.6626, 0.6109, 0.5670, 0.5299, 0.4973, 0.4703, 0.4455,
                         0.4230, 0.4039, 0.3859],
        CFLScaleAlpha_GL=[3.1871, 2.2444, 1.7797, 1.5075, 1.3230, 1.1857, 1.0800, 0.9945, 0.9247, 0.8651, 0.8134, 0.7695,
                          0.7301, 0.6952, 0.6649],
        RelativeDFL=1.0,
    ),
}
# timedisc_vars.f90:93-98
DFL_SCALE_ALPHA_G = [1.12, 0.76, 0.55, 0.41, 0.32, 0.25, 0.20, 0.17, 0.14, 0.12]
DFL_SCALE_ALPHA_GL = [4.50, 1.95, 1.18, 0.79, 0.56, 0.42, 0.32, 0.25, 0.20, 0.17]


class Equation:
    """The parsed ini-surface of the hot path (what `prms` + the CMake cache hold in the reference)."""

    def __init__(self, N, node_type, *, parabolic=True, split_dg=False, split_dg_type="pi", riemann="roeentropyfix",
                 riemann_bc="same", lifting=1, n_var_lifting=5, kappa=1.4, R=287.058, Pr=0.72, mu0=0.0,
                 time_disc="carpenterrk4-5", cfl_scale=0.9, dfl_scale=0.9, eta_br2=2.0, eta_br2_wall=-1.0,
                 boundary_type=(), ref_state_prim=(), ini_ref_state=1, ini_exact_func=0, do_calc_source=False,
                 ini_amplitude=0.1, ini_frequency=1.0, adv_vel=(0.0, 0.0, 0.0), overintegration="none", n_under=None):
        self.N = N
        self.node_type = node_type.lower()
        self.parabolic = parabolic
        self.split_dg = split_dg
        self.split_dg_type = split_dg_type.lower()
        self.riemann = riemann.lower()
        self.riemann_bc = riemann_bc.lower()
        self.lifting = lifting
        self.n_var_lifting = n_var_lifting
        self.kappa, self.R, self.Pr, self.mu0 = float(kappa), float(R), float(Pr), float(mu0)
        self.eta_br2, self.eta_br2_wall = eta_br2, eta_br2_wall
        self.boundary_type = [tuple(b) for b in boundary_type]   # rows (BC_TYPE, BC_STATE, BC_ALPHA)
        # RefState rows are given like in the ini file as (rho,u,v,w,p); T follows from the ideal gas law (equation.f90:121-133)
        self.ref_state_prim = [tuple(r[:5]) + (r[4] / (r[0] * self.R),) for r in ref_state_prim]
        self.do_calc_source = bool(do_calc_source)
        self.ini_exact_func = int(ini_exact_func)
        self.ini_amplitude, self.ini_frequency = float(ini_amplitude), float(ini_frequency)
        self.adv_vel = tuple(adv_vel)
        self.ini_ref_state_prim = self.ref_state_prim[ini_ref_state - 1] if self.ref_state_prim else (1.0, 0.0, 0.0, 0.0, 1.0, 1.0 / self.R)
        tab = RK_TABLES[time_disc]
        self.RKA, self.RKb, self.RKc = np.array(tab["RKA"]), np.array(tab["RKb"]), np.array(tab["RKc"])
        # FillCFL_DFL, timedisc_func.f90:491-519
        gl = self.node_type == "gauss-lobatto"
        # OverintegrationType (dg/overintegration.f90:40-47,97-116): modal cut-off of Ut to NUnder <= N (the cut-off branch keeps
        # the type even for NUnder = N; only the conservative variant disables itself, :139-141)
        self.n_under = N if n_under is None else int(n_under)
        assert 0 <= self.n_under <= N
        self.overintegration = {"none": 0, "cutoff": 1}[overintegration.lower()]
        # NEff = MIN(PP_N,NFilter,NUnder) enters the advective scaling only (timedisc_func.f90:142-144,481); overintegration on
        # Gauss-Lobatto nodes is a projection DG scheme and takes the Gauss table (timedisc_vars.f90:153-160)
        n_eff = min(N, self.n_under) if self.overintegration else N
        gl_cfl = gl and not self.overintegration
        alpha = (tab["CFLScaleAlpha_GL"] if gl_cfl else tab["CFLScaleAlpha_G"])[min(15, n_eff) - 1]
        self.CFLScale = cfl_scale * alpha / (2.0 * n_eff + 1.0)
        alpha_d = (DFL_SCALE_ALPHA_GL if gl else DFL_SCALE_ALPHA_G)[min(10, N) - 1] * tab["RelativeDFL"]
        self.DFLScale = dfl_scale * alpha_d / (2.0 * N + 1.0) ** 2

    # ---- EOS helpers (eos.f90:203-231, 308-330) on arrays [..., var] -------------------------------------------------
    def prim_to_cons(self, prim):
        cons = np.empty(prim.shape[:-1] + (5,))
        cons[..., 0] = prim[..., 0]
        cons[..., 1:4] = prim[..., 1:4] * prim[..., 0:1]
        cons[..., 4] = prim[..., 4] / (self.kappa - 1.0) + 0.5 * np.sum(cons[..., 1:4] * prim[..., 1:4], axis=-1)
        return cons

    def cons_to_prim(self, cons):
        prim = np.empty(cons.shape[:-1] + (6,))
        srho = 1.0 / cons[..., 0]
        prim[..., 0] = cons[..., 0]
        prim[..., 1:4] = cons[..., 1:4] * srho[..., None]
        prim[..., 4] = (self.kappa - 1.0) * (cons[..., 4] - 0.5 * np.sum(cons[..., 1:4] * prim[..., 1:4], axis=-1))
        prim[..., 5] = prim[..., 4] * srho / self.R
        return prim

    # ---- entropy variables (groundwork for the split form on Gauss nodes) ------------------------------
    def prim_to_entropy(self, prim):
        """PrimToEntropy (eos.f90:401-420): w = ((kappa-s)/(kappa-1) - rho|v|^2/(2p), rho v/p, -rho/p), s = ln p - kappa ln rho."""
        k = self.kappa
        s = np.log(prim[..., 4]) - k * np.log(prim[..., 0])
        rho_p = prim[..., 0] / prim[..., 4]
        w = np.empty(prim.shape[:-1] + (5,))
        w[..., 0] = (k - s) / (k - 1.0) - rho_p * 0.5 * np.sum(prim[..., 1:4] ** 2, axis=-1)
        w[..., 1:4] = rho_p[..., None] * prim[..., 1:4]
        w[..., 4] = -rho_p
        return w

    def entropy_to_cons(self, w):
        """EntropyToCons (eos.f90:499-531), the inverse map used after prolonging entropy variables to the faces
        (ProlongToFaceEntropy, prolongtoface.t90:151-204)."""
        k = self.kappa
        km1 = k - 1.0
        w2 = w * km1
        ekin = 0.5 * np.sum(w2[..., 1:4] ** 2, axis=-1) / w2[..., 4]
        s = k - w2[..., 0] + ekin
        rhoe = km1 ** (1.0 / km1) * (-w2[..., 4]) ** (-k / km1) * np.exp(-s / km1)
        cons = np.empty_like(w)
        cons[..., 0] = -rhoe * w2[..., 4]
        cons[..., 1:4] = rhoe[..., None] * w2[..., 1:4]
        cons[..., 4] = rhoe * (1.0 - ekin)
        return cons

    # ---- initial states ------------------------------------------------------------------------------------------------
    def tgv_state(self, x, mach=0.1, ini_const_dens=True):
        """ExactFuncTestcase of the Taylor-Green vortex (testcase.f90:176-219), U0 = rho0 = 1."""
        U0, rho0 = 1.0, 1.0
        p0 = (U0 / mach) ** 2 / self.kappa * rho0
        T0 = p0 / (rho0 * self.R)
        prim = np.zeros(x.shape[:-1] + (6,))
        X, Y, Z = x[..., 0], x[..., 1], x[..., 2]
        prim[..., 1] = U0 * np.sin(X) * np.cos(Y) * np.cos(Z)
        prim[..., 2] = -U0 * np.cos(X) * np.sin(Y) * np.cos(Z)
        prim[..., 3] = 0.0
        prim[..., 4] = p0 + (rho0 * U0 ** 2) / 16.0 * (np.cos(2 * X) * np.cos(2.0 * Z) + 2.0 * np.cos(2.0 * Y)
                                                          + 2.0 * np.cos(2.0 * X) + np.cos(2 * Y) * np.cos(2.0 * Z))
        if ini_const_dens:
            prim[..., 0] = rho0
            prim[..., 5] = prim[..., 4] / (prim[..., 0] * self.R)
        else:
            prim[..., 5] = T0
            prim[..., 0] = prim[..., 4] / (prim[..., 5] * self.R)
        return self.prim_to_cons(prim)

    def exact_func4(self, x, t=0.0):
        """ExactFunc case 4, oblique sine wave (exactfunc.f90:332-362)."""
        omega, a = np.pi * self.ini_frequency, self.adv_vel[0] * 2.0 * np.pi
        r = 2.0 + self.ini_amplitude * np.sin(omega * np.sum(x, axis=-1) - a * t)
        cons = np.empty(x.shape[:-1] + (5,))
        cons[..., 0:4] = r[..., None]
        cons[..., 4] = r * r
        return cons

    def const_state(self, shape, prim):
        p = np.zeros(tuple(shape) + (6,))
        p[...] = tuple(prim[:5]) + (prim[4] / (prim[0] * self.R),)
        return self.prim_to_cons(p)

    def random_state(self, shape, seed=1234):
        """Random-field stress state: rho,p in [0.5,1.5], v in [-0.5,0.5]^3."""
        rng = np.random.default_rng(seed)
        prim = np.zeros(tuple(shape) + (6,))
        prim[..., 0] = rng.uniform(0.5, 1.5, shape)
        prim[..., 1:4] = rng.uniform(-0.5, 0.5, tuple(shape) + (3,))
        prim[..., 4] = rng.uniform(0.5, 1.5, shape)
        prim[..., 5] = prim[..., 4] / (prim[..., 0] * self.R)
        return self.prim_to_cons(prim)

    def smooth_state(self, x, length=2 * np.pi, amp=0.2):
        """Smooth periodic non-trivial state for parity tests on (curved) periodic boxes."""
        k = 2 * np.pi / length
        X, Y, Z = k * x[..., 0], k * x[..., 1], k * x[..., 2]
        prim = np.zeros(x.shape[:-1] + (6,))
        prim[..., 0] = 1.0 + amp * np.sin(X + 0.3) * np.cos(Y) * np.cos(2 * Z + 0.1)
        prim[..., 1] = 0.3 + amp * np.sin(X) * np.cos(Y + 0.2) * np.cos(Z)
        prim[..., 2] = -0.2 + amp * np.cos(2 * X) * np.sin(Y) * np.cos(Z + 0.5)
        prim[..., 3] = 0.1 + amp * np.cos(X + 0.7) * np.cos(Y) * np.sin(Z)
        prim[..., 4] = 1.0 + amp * np.cos(X) * np.sin(2 * Y + 0.4) * np.cos(Z)
        prim[..., 5] = prim[..., 4] / (prim[..., 0] * self.R)
        return self.prim_to_cons(prim)


def sutherland_params(Ts_ini=110.4, Tref_ini=273.15, expo=1.5):
    """InitEos for PP_VISC = 1 (eos.f90:170-178): returns the module values (Ts, Tref, ExpoSuth, cSuth) with Tref = 1/Tref_ini and
    Ts = Ts_ini/Tref_ini, as flexi_gpu_set_viscosity / oracle.set_viscosity expect them."""
    Tref = 1.0 / Tref_ini
    Ts = Ts_ini * Tref
    cSuth = Ts ** expo * (1 + Ts) / (2 * Ts * Ts)
    return dict(Ts=Ts, Tref=Tref, ExpoSuth=expo, cSuth=cSuth)
