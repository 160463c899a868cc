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
