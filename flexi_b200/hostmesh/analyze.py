"""Host-side analysis of the synthetic FLEXI host (stand-in for MOD_Analyze / the TGV AnalyzeTestcase).

The Fortran host keeps these routines; the GPU library only has to hand U (and the lifted gradients) back at analyze steps.
Reference: src/analyze/analyze.f90:224-278 (InitAnalyzeBasis), :391-485 (CalcErrorNorms), :150-158 (Vol);
           src/testcase/taylorgreenvortex/testcase.f90:241-480 (AnalyzeTestcase).
"""
import numpy as np

from . import basis as B


class AnalyzeBasis:
    def __init__(self, basis, n_analyze=None):
        N = basis.N
        self.NAnalyze = 2 * (N + 1) if n_analyze is None else n_analyze
        xa, wa, _ = B.get_nodes_and_weights(self.NAnalyze, B.NODE_GAUSS_LOBATTO)
        self.Vdm = B.initialize_vandermonde(basis.xGP, basis.wBary, xa)
        self.w3 = np.einsum("k,j,i->kji", wa, wa, wa)
        self.wGP3 = np.einsum("k,j,i->kji", basis.wGP, basis.wGP, basis.wGP)

    def volume(self, sJ):
        return float(np.sum(self.wGP3[None] / sJ))


def tgv_analyze(ab, eqn, mesh, U, gradUx, gradUy, gradUz, vol, rho0=1.0):
    """AnalyzeTestcase for one rank; returns the un-normalised sums and the normalised CSV columns (single rank)."""
    nvl = gradUx.shape[-1]
    v0 = 1 if nvl == 5 else 0                                   # LIFT_VELV
    sJ_A = B.change_basis_3d(ab.Vdm, mesh.sJ[..., None])[..., 0]
    U_A = B.change_basis_3d(ab.Vdm, U)
    gvx = B.change_basis_3d(ab.Vdm, gradUx[..., v0:v0 + 3])     # [..., c] = d v_c / dx
    gvy = B.change_basis_3d(ab.Vdm, gradUy[..., v0:v0 + 3])
    gvz = B.change_basis_3d(ab.Vdm, gradUz[..., v0:v0 + 3])
    P = eqn.cons_to_prim(U_A)
    GradVel = np.stack([gvx, gvy, gvz], axis=-1)                # GradVel[..., c, d] = d v_c / d x_d
    divU = GradVel[..., 0, 0] + GradVel[..., 1, 1] + GradVel[..., 2, 2]
    S = 0.5 * (GradVel + np.swapaxes(GradVel, -1, -2))
    Sd = S.copy()
    for i in range(3):
        Sd[..., i, i] -= divU / 3.0
    vort = np.stack([GradVel[..., 2, 1] - GradVel[..., 1, 2], GradVel[..., 0, 2] - GradVel[..., 2, 0],
                     GradVel[..., 1, 0] - GradVel[..., 0, 1]], axis=-1)
    w = ab.w3[None] / sJ_A
    v2 = np.sum(P[..., 1:4] ** 2, axis=-1)
    o2 = np.sum(vort ** 2, axis=-1)
    mu = eqn.mu0
    out = dict(
        T_mean=np.sum(w * P[..., 5]) / vol,
        Entropy=np.sum(w * (-1.0 / (eqn.kappa - 1.0)) * P[..., 0] * (np.log(P[..., 4]) - eqn.kappa * np.log(P[..., 0]))) / vol,
        Ekin=np.sum(w * 0.5 * v2) / vol,
        Ekin_comp=np.sum(w * 0.5 * P[..., 0] * v2) / vol / rho0,
        Enstr=np.sum(w * 0.5 * P[..., 0] * o2) / (rho0 * vol),
        DR_u=np.sum(w * np.sum(GradVel ** 2, axis=(-1, -2))) * mu / vol,
        DR_S=np.sum(w * np.sum(S ** 2, axis=(-1, -2))) * 2.0 * mu / (rho0 * vol),
        DR_Sd=np.sum(w * np.sum(Sd ** 2, axis=(-1, -2))) * 2.0 * mu / (rho0 * vol),
        DR_p=-np.sum(w * P[..., 4] * divU) / (rho0 * vol),
        Vorticity_max=np.sqrt(np.max(o2)),
        ED_S=np.sum(w * mu * o2) / vol,
        ED_D=np.sum(w * mu * divU ** 2) * 4.0 / 3.0 / vol,
    )
    out["uPrime"] = np.sqrt(2.0 / 3.0 * out["Ekin"])
    return out


def calc_error_norms(ab, mesh, U, exact_fn, vol):
    """CalcErrorNorms (analyze.f90:391-485) for one rank: returns (sum of squared L2 contributions, Linf) per variable."""
    X_A = B.change_basis_3d(ab.Vdm, mesh.Elem_xGP)
    J_A = B.change_basis_3d(ab.Vdm, (1.0 / mesh.sJ)[..., None])[..., 0]
    U_A = B.change_basis_3d(ab.Vdm, U)
    err = U_A - exact_fn(X_A)
    linf = np.max(np.abs(err).reshape(-1, err.shape[-1]), axis=0)
    l2sq = np.einsum("ekjiv,ekji->v", err * err, ab.w3[None] * J_A)
    return l2sq, linf


def calc_body_forces(basis, eqn, mesh, boundary_type, U_master, gradUx_master, gradUy_master, gradUz_master, mom_origin=(0.0, 0.0, 0.0)):
    """CalcBodyForces (equations/navierstokes/calcbodyforces.f90:46-250) for one rank: (Fp, Fv, Mp, Mv), each [nBCs, 3].
    Wall BCs are the boundary types 3, 4, 9 (analyze_equation.f90:104-112).  Arrays as [side, q, p, var]."""
    nBCs = len(boundary_type)
    Fp, Fv, Mp, Mv = (np.zeros((nBCs, 3)) for _ in range(4))
    wS = np.einsum("q,p->qp", basis.wGP, basis.wGP)
    nvl = gradUx_master.shape[-1]
    v0 = 1 if nvl == 5 else 0
    mo = np.asarray(mom_origin, dtype=float)
    for s in range(mesh.nBCSides):
        b = int(mesh.BC[s]) - 1
        if boundary_type[b][0] not in (3, 4, 9):
            continue
        P = eqn.cons_to_prim(U_master[s])
        dA = wS * mesh.SurfElem[s]
        nv = mesh.NormVec[s]
        r = mesh.Face_xGP[s] - mo
        df = P[..., 4, None] * nv * dA[..., None]
        Fp[b] += df.sum(axis=(0, 1))
        Mp[b] += np.cross(r, df).sum(axis=(0, 1))
        G = np.stack([gradUx_master[s][..., v0:v0 + 3], gradUy_master[s][..., v0:v0 + 3], gradUz_master[s][..., v0:v0 + 3]], axis=-1)
        div = G[..., 0, 0] + G[..., 1, 1] + G[..., 2, 2]
        tau = eqn.mu0 * (G + np.swapaxes(G, -1, -2))
        for a in range(3):
            tau[..., a, a] -= 2.0 / 3.0 * eqn.mu0 * div
        dfv = np.einsum("qpab,qpb->qpa", tau, nv) * dA[..., None]
        Fv[b] -= dfv.sum(axis=(0, 1))
        Mv[b] -= np.cross(r, dfv).sum(axis=(0, 1))
    return Fp, Fv, Mp, Mv


def calc_timestep_elem(eqn, mesh, U):
    """dtElem(nElems) of CalcTimeStep (equations/navierstokes/calctimestep.f90:140-210; MetricsAdv / MetricsVisc :42-83), the
    ElemData array 'dt' FLEXI writes into its state files.  Constant viscosity; U as [e,k,j,i,5]."""
    kappa = eqn.kappa
    rho = U[..., 0]
    v = U[..., 1:4] / rho[..., None]
    p = (kappa - 1.0) * (U[..., 4] - 0.5 * rho * np.sum(v * v, axis=-1))
    c = np.sqrt(kappa * p / rho)
    sJ = mesh.sJ
    lam, lam_v = [], []
    kappas_pr = max(4.0 / 3.0, kappa / eqn.Pr)
    for M in (mesh.Metrics_fTilde, mesh.Metrics_gTilde, mesh.Metrics_hTilde):
        adv = sJ * np.sqrt(np.sum(M * M, axis=-1))
        lam.append(np.max(np.abs(np.sum(M * (v * sJ[..., None]), axis=-1)) + c * adv, axis=(1, 2, 3)))
        visc = kappas_pr * np.sum((M * sJ[..., None]) ** 2, axis=-1)
        lam_v.append(np.max(eqn.mu0 / rho * visc, axis=(1, 2, 3)))
    dt = eqn.CFLScale * 2.0 / (lam[0] + lam[1] + lam[2])
    if eqn.parabolic:
        sv = lam_v[0] + lam_v[1] + lam_v[2]
        dt = np.where(sv > 0.0, np.minimum(dt, eqn.DFLScale * 4.0 / np.where(sv > 0.0, sv, 1.0)), dt)
    return dt
