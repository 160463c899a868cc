"""Host restatement of the non-conforming (mortar) interface operators of the reference: InitMortar / MortarBasis_BigToSmall /
MortarBasis_SmallToBig (mortar/mortar.f90:40-245) and the face-level parts of U_Mortar / Flux_Mortar (mortar/fillmortar.t90:35-377).

First step of the mortar plan: the 1-D operators and their tensor-product application to one big
side, in the orientation of the big side (the placement into U_master / U_slave of the small SideIDs with their flips is mesh
bookkeeping, `MortarInfo`, not restated yet).  Not wired into the operator: the mesh reader still refuses mortar meshes and so
does the GPU library.  Matrices are stored like the reference stores them -- "already transposed": M[l, p] multiplies input
node l for output node p."""
import numpy as np

from . import basis as B


def mortar_basis(n, node_type):
    """(M_0_1, M_0_2, M_1_0, M_2_0) of InitMortar for DG elements (FVE = 0), each (0:N,0:N) indexed [l, p] = [in, out]."""
    xi, _, wbary = B.get_nodes_and_weights(n, node_type)
    # big -> small: interpolation to the two half intervals (mortar.f90:121-126); InitializeVandermonde returns Vdm[out, in] and
    # the reference's call fills M_0_x(in, out) (the "transposed" storage U_Mortar indexes as M1(l,q))
    m01 = B.initialize_vandermonde(xi, wbary, 0.5 * (xi - 1.0)).T
    m02 = B.initialize_vandermonde(xi, wbary, 0.5 * (xi + 1.0)).T
    # small -> big: L2 projection through Gauss quadrature on each half and the Legendre basis (mortar.f90:201-224,240-244)
    xg, wg, _ = B.get_nodes_and_weights(n, B.NODE_GAUSS)
    vgp = np.diag(wg) @ B.get_vandermonde(n, node_type, n, B.NODE_GAUSS)
    leg = lambda x: np.array([[B.legendre_polynomial_and_derivative(j, xx)[0] for j in range(n + 1)] for xx in x])  # noqa: E731
    vdm_leg, vphi1, vphi2 = leg(xi), leg(0.5 * (xg - 1.0)), leg(0.5 * (xg + 1.0))
    m10 = (vdm_leg @ (vphi1.T @ vgp)).T
    m20 = (vdm_leg @ (vphi2.T @ vgp)).T
    return m01, m02, m10, m20


def init_mortar_self_check(n):
    """The reference's own check in InitMortar (mortar.f90:77-88, Gauss nodes): projecting the constants 0.5 and 1.5 from the two
    halves and integrating gives |0.25 * sum((M_1_0^T 0.5 + M_2_0^T 1.5) w) - 1| < 100 * PP_RealTolerance."""
    _, _, m10, m20 = mortar_basis(n, B.NODE_GAUSS)
    _, wg, _ = B.get_nodes_and_weights(n, B.NODE_GAUSS)
    t1, t2 = np.full(n + 1, 0.5), np.full(n + 1, 1.5)
    # MATMUL(TRANSPOSE(M_1_0), test1): the stored matrices are transposed, so this is the plain [out, in] application
    return abs(0.25 * np.sum(((m10.T @ t1) + (m20.T @ t2)) * wg) - 1.0)


def u_mortar(u_big, mortar_type, m01, m02):
    """U_Mortar, fillmortar.t90:103-178: big-side data u_big[q, p, var] -> small sides [iMortar, q, p, var] (4 for type 1: first
    in eta, then in xi, numbered 1..4 = (xi-,eta-), (xi+,eta-), (xi-,eta+), (xi+,eta+); 2 for type 2 (split in eta) and 3 (in xi)).
    Sums run over ascending l like the Fortran loops."""
    def along_eta(u, m):   # out[q,p] = sum_l m[l,q] * u[l,p]
        out = m[0][:, None, None] * u[0][None]
        for l in range(1, u.shape[0]):
            out = out + m[l][:, None, None] * u[l][None]
        return out

    def along_xi(u, m):    # out[q,p] = sum_l m[l,p] * u[q,l]
        out = m[0][None, :, None] * u[:, 0][:, None]
        for l in range(1, u.shape[1]):
            out = out + m[l][None, :, None] * u[:, l][:, None]
        return out

    if mortar_type == 1:
        t1, t2 = along_eta(u_big, m01), along_eta(u_big, m02)
        return np.stack([along_xi(t1, m01), along_xi(t1, m02), along_xi(t2, m01), along_xi(t2, m02)])
    if mortar_type == 2:
        return np.stack([along_eta(u_big, m01), along_eta(u_big, m02)])
    if mortar_type == 3:
        return np.stack([along_xi(u_big, m01), along_xi(u_big, m02)])
    raise ValueError(mortar_type)


def flux_mortar(flux_small, mortar_type, m10, m20):
    """Flux_Mortar, fillmortar.t90:283-372: small-side fluxes [iMortar, q, p, var] (already in the big side's orientation and sign)
    -> the big side's flux [q, p, var]: type 1 first in xi, then in eta."""
    def xi_pair(fa, fb):   # out[q,p] = sum_l m10[l,p] fa[q,l] + m20[l,p] fb[q,l]
        out = m10[0][None, :, None] * fa[:, 0][:, None] + m20[0][None, :, None] * fb[:, 0][:, None]
        for l in range(1, fa.shape[1]):
            out = out + m10[l][None, :, None] * fa[:, l][:, None] + m20[l][None, :, None] * fb[:, l][:, None]
        return out

    def eta_pair(fa, fb):  # out[q,p] = sum_l m10[l,q] fa[l,p] + m20[l,q] fb[l,p]
        out = m10[0][:, None, None] * fa[0][None] + m20[0][:, None, None] * fb[0][None]
        for l in range(1, fa.shape[0]):
            out = out + m10[l][:, None, None] * fa[l][None] + m20[l][:, None, None] * fb[l][None]
        return out

    if mortar_type == 1:
        return eta_pair(xi_pair(flux_small[0], flux_small[1]), xi_pair(flux_small[2], flux_small[3]))
    if mortar_type == 2:
        return eta_pair(flux_small[0], flux_small[1])
    if mortar_type == 3:
        return xi_pair(flux_small[0], flux_small[1])
    raise ValueError(mortar_type)


def mortar_surf_metrics(face_ja, face_xgp, mortar_type, m01, m02):
    """Mortar_CalcSurfMetrics (mortar/mortar_metrics.f90:55-162): surface metrics and face points of the small virtual sides from
    those of the big side -- the same tensor-product interpolation as U_Mortar, the metrics scaled by 1/2 per split direction
    (M_0_12_h; the small side's reference square covers half the big one's in that direction).  face_ja [q, p, 3, 3] or any
    trailing shape, face_xgp [q, p, 3]; returns (Mortar_Ja [iMortar, q, p, ...], Mortar_xGP [iMortar, q, p, 3])."""
    sh = face_ja.shape
    ja = u_mortar(face_ja.reshape(sh[0], sh[1], -1), mortar_type, m01, m02)
    ja = ja.reshape((ja.shape[0],) + sh) * (0.25 if mortar_type == 1 else 0.5)
    return ja, u_mortar(face_xgp, mortar_type, m01, m02)
