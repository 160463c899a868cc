"""Assemble a complete synthetic FLEXI host state (mesh + basis + equation + descriptor) for one or several ranks."""
import numpy as np

from . import basis as B
from . import mappings as MP
from .equation import Equation
from .mesh import BoxSpec, build_mesh
from ..desc import make_desc


class HostRank:
    """What one FLEXI rank holds after InitFlexi: the module arrays + the descriptor handed to the GPU library."""

    def __init__(self, mesh, basis, eqn, S2V2, nccl_id=None, device=0):
        self.mesh, self.basis, self.eqn, self.S2V2 = mesh, basis, eqn, S2V2
        self.desc, self.keep = make_desc(mesh, basis, eqn, S2V2, nccl_id=nccl_id, device=device)

    def remake_desc(self, nccl_id=None, device=0):
        self.desc, self.keep = make_desc(self.mesh, self.basis, self.eqn, self.S2V2, nccl_id=nccl_id, device=device)


def build_case(spec, N, node_type, eqn_kwargs=None, nranks=1, partition="linear", blocks=None, ranks=None, with_ja_face=False):
    """Build HostRank objects for the requested ranks (default: all).  with_ja_face also fills mesh.Ja_Face / Ja_slave (only the
    split form on Gauss nodes reads them)."""
    eqn_kwargs = dict(eqn_kwargs or {})
    basis = B.DGBasis(N, node_type, split_dg=eqn_kwargs.get("split_dg", False))
    _, _, S2V2 = MP.build_mappings(N)
    out = []
    for r in (range(nranks) if ranks is None else ranks):
        mesh = build_mesh(spec, N, node_type, rank=r, nranks=nranks, partition=partition, blocks=blocks, with_ja_face=with_ja_face)
        eqn = Equation(N, node_type, **eqn_kwargs)
        out.append(HostRank(mesh, basis, eqn, S2V2))
    return out


def tgv_box(nel, **kw):
    """[0,2pi]^3 fully periodic Cartesian box (tutorials/taylorgreenvortex/parameter_hopr.ini)."""
    if isinstance(nel, int):
        nel = (nel, nel, nel)
    return BoxSpec(nel, x0=(0.0, 0.0, 0.0), length=(2 * np.pi,) * 3, periodic=(True, True, True), **kw)


TGV_EQN = dict(kappa=1.4, R=71.42857, Pr=0.72, mu0=6.25e-4)  # regressioncheck/checks/tgv/split/parameter.ini


# ---- BASELINE configs[3]: NACA0012-style synthetic curved mesh ----------------------------------------------------------------
def naca_mapping(chord=1.0, thickness=0.12, r_far=8.0, stretch=3.0, span=1.0):
    """Unit cube -> O-grid around a NACA0012-style section: xhat0 runs once around the section (clockwise, so that the frame
    (around, radial, span) is right-handed), xhat1 from the wall (0) to the far-field circle (1), xhat2 along the span.
    The section is the smooth teardrop x/c = (1 + cos(theta))/2, y/c = A sin(theta) (1 - 0.55 cos(theta)): 12 % thick with the
    maximum at 31 % chord like the NACA 0012, but with a rounded trailing edge -- an O-grid around a sharp trailing edge has a
    degenerate corner there (zero Jacobian on the Gauss-Lobatto node that sits on it)."""
    c0 = (1.0 - np.sqrt(1.0 + 2.0 * 1.1 * 1.1)) / 2.2          # cos(theta) of the thickest station
    amp = 0.5 * thickness / (np.sqrt(1.0 - c0 * c0) * (1.0 - 0.55 * c0))

    def f(X):
        th = -2.0 * np.pi * X[..., 0]
        bx = chord * 0.5 * (1.0 + np.cos(th))
        by = chord * amp * np.sin(th) * (1.0 - 0.55 * np.cos(th))
        ox, oy = 0.5 * chord + r_far * np.cos(th), r_far * np.sin(th)
        s = np.expm1(stretch * X[..., 1]) / np.expm1(stretch)
        out = np.empty_like(X)
        out[..., 0] = (1.0 - s) * bx + s * ox
        out[..., 1] = (1.0 - s) * by + s * oy
        out[..., 2] = span * X[..., 2]
        return out
    return f


NACA_BC = {(1, -1): 1, (1, 1): 2}  # wall at xhat1 = 0, far field at xhat1 = 1; around the section and along the span: periodic
# HLLC + BR2 (configs[3]); BC 1 = isothermal no-slip wall at the free-stream temperature, BC 2 = Dirichlet far field (RefState 1)
# CFL/DFL scale 0.4: with 0.8 the BR2 penalty on the strongly anisotropic wall elements makes LSERK4-5 unstable (round-off grows
# twelvefold per step -- in the oracle's two builds alike), so that no state comparison beyond 3 steps is meaningful
NACA_EQN = dict(parabolic=True, lifting=2, riemann="hllc", mu0=2e-3, eta_br2=2.0, eta_br2_wall=3.0, cfl_scale=0.4, dfl_scale=0.4,
                boundary_type=[(4, 1, 0), (2, 1, 0)], ref_state_prim=[(1.0, 0.3, 0.0, 0.0, 0.71428571)])


def naca_spec(nel, ngeo=4, orient_seed=12, span_per_elem=0.1, **kw):
    """O-mesh of nel = (around, radial, span) elements; every element in a random local frame (orient_seed) so that the side
    connectivity carries all flips and local-side pairings of an unstructured HOPR mesh (mappings.f90:188-450)."""
    return BoxSpec(nel, periodic=(True, False, True), bc_of_face=NACA_BC, ngeo=ngeo, orient_seed=orient_seed,
                   mapping=naca_mapping(span=span_per_elem * nel[2], **kw))


def naca_state(eqn, x, chord=1.0):
    """Free stream that is ramped down towards the section (no impulsive start at the no-slip wall)."""
    rs = eqn.ref_state_prim[0]
    dx = np.maximum(np.maximum(x[..., 0] - chord, -x[..., 0]), 0.0)
    d = np.sqrt(dx ** 2 + x[..., 1] ** 2)
    ramp = np.tanh(4.0 * d / chord)
    prim = np.zeros(x.shape[:-1] + (6,))
    prim[..., 0] = rs[0]
    prim[..., 1] = rs[1] * ramp
    prim[..., 2] = rs[2] * ramp
    prim[..., 3] = rs[3] * ramp
    prim[..., 4] = rs[4]
    prim[..., 5] = prim[..., 4] / (prim[..., 0] * eqn.R)
    return eqn.prim_to_cons(prim)


def synthetic_sponge(r, frac=0.4, seed=3):
    """A sponge zone as InitSponge would hand it over (sponge.f90:97-400): the elements whose barycentre lies in the last `frac` of
    the x range, a smooth ramp times the Jacobian as SpongeMat, a perturbed smooth state as SpBaseFlow.
    Returns SpongeMap (1-based), SpongeMat [nSp,k,j,i], SpBaseFlow [nElems,k,j,i,5]."""
    m = r.mesh
    xb = m.Elem_xGP[..., 0].reshape(m.nElems, -1).mean(axis=1)
    x0, x1 = m.Elem_xGP[..., 0].min(), m.Elem_xGP[..., 0].max()
    xs = x1 - frac * (x1 - x0)
    smap = np.nonzero(xb > xs)[0].astype(np.int32) + 1
    ramp = np.clip((m.Elem_xGP[smap - 1, ..., 0] - xs) / (x1 - xs), 0.0, 1.0)
    smat = 5.0 * (6 * ramp ** 5 - 15 * ramp ** 4 + 10 * ramp ** 3) / m.sJ[smap - 1]  # damping * x_star, times the Jacobian 1/sJ
    rng = np.random.default_rng(seed)
    base = r.eqn.smooth_state(m.Elem_xGP, length=2.0, amp=0.05) * (1.0 + 0.01 * rng.standard_normal(m.Elem_xGP.shape[:-1] + (5,)))
    return smap, smat, base
