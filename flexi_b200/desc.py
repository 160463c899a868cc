"""ctypes mirror of `flexi_gpu_desc` (include/flexi_gpu.h) and the code that fills it from the synthetic host's module arrays.

This is the Python twin of what the Fortran shim (flexi_b200/shim/mod_gpu.f90) does with C_LOC of the MOD_*_Vars arrays.
"""
import ctypes as C

import numpy as np

ABI_VERSION = 2
NODETYPE = {"gauss": 1, "gauss-lobatto": 2}
RIEMANN = {"same": -1, "lf": 1, "hllc": 2, "roe": 3, "roel2": 32, "roeentropyfix": 33, "hll": 4, "hlle": 5, "hllem": 6,
           "average": 0, "ch": 7, "ir": 8}  # the last three: SPLIT_DG builds only (riemann.f90:54-58)
SPLITDG = {"sd": 0, "mo": 1, "du": 2, "kg": 3, "pi": 4, "ch": 5, "ir": 6}

_PD = C.POINTER(C.c_double)
_PI = C.POINTER(C.c_int32)


class FlexiGpuDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("struct_bytes", C.c_int32),
        ("N", C.c_int32), ("nodeType", C.c_int32), ("splitDG", C.c_int32), ("parabolic", C.c_int32),
        ("lifting", C.c_int32), ("nVarLifting", C.c_int32),
        ("splitDGType", C.c_int32), ("riemann", C.c_int32), ("riemannBC", C.c_int32),
        ("doWeakLifting", C.c_int32), ("doConservativeLifting", C.c_int32),
        ("etaBR2", C.c_double), ("etaBR2_wall", C.c_double),
        ("nElems", C.c_int32), ("nSides", C.c_int32), ("nBCSides", C.c_int32), ("nInnerSides", C.c_int32),
        ("nMPISides_MINE", C.c_int32), ("nMPISides_YOUR", C.c_int32),
        ("Metrics_fTilde", _PD), ("Metrics_gTilde", _PD), ("Metrics_hTilde", _PD), ("sJ", _PD), ("Elem_xGP", _PD),
        ("NormVec", _PD), ("TangVec1", _PD), ("TangVec2", _PD), ("SurfElem", _PD), ("Face_xGP", _PD),
        ("Ja_Face", _PD), ("Ja_slave", _PD),
        ("SideToElem", _PI), ("ElemToSide", _PI), ("S2V2", _PI), ("BC", _PI),
        ("nBCs", C.c_int32), ("BoundaryType", _PI), ("nRefState", C.c_int32), ("RefStatePrim", _PD),
        ("xGP", _PD), ("wGP", _PD), ("D", _PD), ("D_T", _PD), ("D_Hat_T", _PD), ("DVolSurf", _PD),
        ("L_Minus", _PD), ("L_Plus", _PD), ("L_HatMinus", _PD), ("L_HatPlus", _PD),
        ("kappa", C.c_double), ("R", C.c_double), ("Pr", C.c_double), ("mu0", C.c_double),
        ("doCalcSource", C.c_int32), ("iniExactFunc", C.c_int32),
        ("iniAmplitude", C.c_double), ("iniFrequency", C.c_double), ("advVel", C.c_double * 3),
        ("iniRefStatePrim", C.c_double * 6),
        ("nRKStages", C.c_int32), ("RKA", _PD), ("RKb", _PD), ("RKc", _PD),
        ("CFLScale", C.c_double), ("DFLScale", C.c_double),
        ("myRank", C.c_int32), ("nProcs", C.c_int32), ("nNbProcs", C.c_int32),
        ("NbProc", _PI), ("nMPISides_MINE_Proc", _PI), ("nMPISides_YOUR_Proc", _PI),
        ("offsetMPISides_MINE", _PI), ("offsetMPISides_YOUR", _PI),
        ("ncclUniqueId", C.c_void_p), ("device", C.c_int32), ("reserved", C.c_int32),
    ]


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _fortran_matrix(a):
    """numpy M[r, c] == Fortran M(r, c)  ->  column-major memory."""
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).T)


def make_desc(mesh, basis, eqn, S2V2, nccl_id=None, device=0):
    """Fill a FlexiGpuDesc. Returns (desc, keepalive) -- keepalive holds the numpy buffers the pointers refer to."""
    d = FlexiGpuDesc()
    keep = []

    def pd(a):
        if a is None:
            return None
        a = _f64(a)
        keep.append(a)
        return a.ctypes.data_as(_PD)

    def pi(a):
        if a is None:
            return None
        a = _i32(a)
        keep.append(a)
        return a.ctypes.data_as(_PI)

    d.abi_version = ABI_VERSION
    d.struct_bytes = C.sizeof(FlexiGpuDesc)
    d.N = mesh.N
    d.nodeType = NODETYPE[basis.node_type]
    d.splitDG = int(eqn.split_dg)
    d.parabolic = int(eqn.parabolic)
    d.lifting = int(eqn.lifting)
    d.nVarLifting = int(eqn.n_var_lifting)
    d.splitDGType = SPLITDG[eqn.split_dg_type]
    d.riemann = RIEMANN[eqn.riemann]
    d.riemannBC = RIEMANN[eqn.riemann_bc]
    d.doWeakLifting = 0
    d.doConservativeLifting = 0
    d.etaBR2 = eqn.eta_br2
    d.etaBR2_wall = eqn.eta_br2_wall if eqn.eta_br2_wall != -1.0 else eqn.eta_br2
    d.nElems, d.nSides, d.nBCSides, d.nInnerSides = mesh.nElems, mesh.nSides, mesh.nBCSides, mesh.nInnerSides
    d.nMPISides_MINE, d.nMPISides_YOUR = mesh.nMPISides_MINE, mesh.nMPISides_YOUR
    d.Metrics_fTilde = pd(mesh.Metrics_fTilde)
    d.Metrics_gTilde = pd(mesh.Metrics_gTilde)
    d.Metrics_hTilde = pd(mesh.Metrics_hTilde)
    d.sJ = pd(mesh.sJ)
    d.Elem_xGP = pd(mesh.Elem_xGP)
    d.NormVec = pd(mesh.NormVec)
    d.TangVec1 = pd(mesh.TangVec1)
    d.TangVec2 = pd(mesh.TangVec2)
    d.SurfElem = pd(mesh.SurfElem)
    d.Face_xGP = pd(mesh.Face_xGP)
    # only the split form on Gauss nodes reads the face metrics; the host mesh builds them on request (with_ja_face)
    ja = hasattr(mesh, "Ja_Face") and hasattr(mesh, "Ja_slave")  # numpy [side, q, p, c, dd] == Fortran (dd,c,p,q,side)
    d.Ja_Face = pd(mesh.Ja_Face) if ja else None
    d.Ja_slave = pd(mesh.Ja_slave) if ja else None
    d.SideToElem = pi(mesh.SideToElem)
    d.ElemToSide = pi(mesh.ElemToSide)
    # numpy S2V2[c,p,q,f,s] -> Fortran memory (c fastest)
    d.S2V2 = pi(np.transpose(S2V2, (4, 3, 2, 1, 0)))
    d.BC = pi(mesh.BC)
    bt = np.asarray(eqn.boundary_type, dtype=np.int32).reshape(-1, 3) if len(eqn.boundary_type) else np.zeros((0, 3), np.int32)
    d.nBCs = bt.shape[0]
    d.BoundaryType = pi(np.ascontiguousarray(bt.T)) if bt.shape[0] else None   # Fortran (nBCs,3): BC index fastest
    rs = np.asarray(eqn.ref_state_prim, dtype=np.float64).reshape(-1, 6) if len(eqn.ref_state_prim) else np.zeros((0, 6))
    d.nRefState = rs.shape[0]
    d.RefStatePrim = pd(rs) if rs.shape[0] else None
    d.xGP, d.wGP = pd(basis.xGP), pd(basis.wGP)
    d.D, d.D_T = pd(_fortran_matrix(basis.D)), pd(_fortran_matrix(basis.D_T))
    d.D_Hat_T = pd(_fortran_matrix(basis.D_Hat_T))
    d.DVolSurf = pd(_fortran_matrix(basis.DVolSurf))
    d.L_Minus, d.L_Plus = pd(basis.L_Minus), pd(basis.L_Plus)
    d.L_HatMinus, d.L_HatPlus = pd(basis.L_HatMinus), pd(basis.L_HatPlus)
    d.kappa, d.R, d.Pr, d.mu0 = eqn.kappa, eqn.R, eqn.Pr, eqn.mu0
    d.doCalcSource = int(eqn.do_calc_source)
    d.iniExactFunc = int(eqn.ini_exact_func)
    d.iniAmplitude, d.iniFrequency = eqn.ini_amplitude, eqn.ini_frequency
    for i in range(3):
        d.advVel[i] = eqn.adv_vel[i]
    for i in range(6):
        d.iniRefStatePrim[i] = eqn.ini_ref_state_prim[i]
    d.nRKStages = len(eqn.RKb)
    d.RKA, d.RKb, d.RKc = pd(eqn.RKA), pd(eqn.RKb), pd(eqn.RKc)
    d.CFLScale, d.DFLScale = eqn.CFLScale, eqn.DFLScale
    d.myRank, d.nProcs, d.nNbProcs = mesh.rank, mesh.nranks, mesh.nNbProcs
    if mesh.nNbProcs:
        d.NbProc = pi(mesh.NbProc)
        d.nMPISides_MINE_Proc = pi(mesh.nMPISides_MINE_Proc)
        d.nMPISides_YOUR_Proc = pi(mesh.nMPISides_YOUR_Proc)
    d.offsetMPISides_MINE = pi(mesh.offsetMPISides_MINE)
    d.offsetMPISides_YOUR = pi(mesh.offsetMPISides_YOUR)
    if nccl_id is not None:
        buf = C.create_string_buffer(bytes(nccl_id), 128)
        keep.append(buf)
        d.ncclUniqueId = C.cast(buf, C.c_void_p)
    d.device = device
    return d, keep
