"""ctypes front-end of the CPU oracle (oracle/flexi_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def build(force=False):
    # make decides by time stamps: a stale library after an edit of flexi_oracle.c is rebuilt
    subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))


def _load(fast=False):
    name = "liboracle_fast.so" if fast else "liboracle.so"
    if name in _LIBS:
        return _LIBS[name]
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    P = C.c_void_p
    PD = C.POINTER(C.c_double)
    lib.oracle_create.restype = P
    lib.oracle_create.argtypes = [P]
    lib.oracle_destroy.argtypes = [P]
    lib.oracle_world_create.restype = P
    lib.oracle_world_create.argtypes = [C.POINTER(P), C.c_int]
    lib.oracle_world_destroy.argtypes = [P]
    lib.oracle_world_dg_time_derivative.argtypes = [P, C.c_double]
    lib.oracle_world_timestep.argtypes = [P, C.c_double, C.c_double]
    lib.oracle_world_calc_timestep.restype = C.c_double
    lib.oracle_world_calc_timestep.argtypes = [P, PD, PD, C.POINTER(C.c_int)]
    lib.oracle_array.restype = PD
    lib.oracle_array.argtypes = [P, C.c_int, C.POINTER(C.c_int64)]
    lib.oracle_prolong_to_face.argtypes = [P, C.c_int, PD, PD, PD, C.c_int]
    lib.oracle_surf_int.argtypes = [P, C.c_int, PD, PD, PD, C.c_int, PD, PD]
    lib.oracle_riemann_point.argtypes = [P, C.c_int, PD, PD, PD, PD, PD, PD]
    lib.oracle_split_volume_flux.argtypes = [P, C.c_int, PD, PD, PD, PD, PD]
    lib.oracle_eval_flux3d.argtypes = [P, PD, PD, PD, PD]
    lib.oracle_eval_diff_flux3d.argtypes = [P, PD, PD, PD, PD, PD, PD, PD]
    lib.oracle_prim_to_entropy.argtypes = [P, PD, PD]
    lib.oracle_entropy_to_cons.argtypes = [P, PD, PD]
    lib.oracle_lv_vel1.argtypes = [P]
    lib.oracle_lv_temp.argtypes = [P]
    lib.oracle_set_face_metrics.argtypes = [P, PD, PD]
    lib.oracle_set_overintegration.argtypes = [P, C.c_int, PD]
    lib.oracle_filter.argtypes = [P, PD, PD]
    lib.oracle_set_filter.argtypes = [P, C.c_int, PD]
    lib.oracle_set_overintegration_cons.argtypes = [P, C.c_int, PD, PD, PD]
    lib.oracle_set_overintegration_cons.restype = C.c_int
    lib.oracle_filter_conservative.argtypes = [P, PD]
    lib.oracle_set_filter_laf.argtypes = [P, PD, C.c_double, C.c_double]
    lib.oracle_set_filter_laf.restype = C.c_int
    lib.oracle_get_filter_laf.argtypes = [P, PD, PD, PD]
    lib.oracle_set_viscosity.argtypes = [P, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]
    lib.oracle_set_viscosity.restype = C.c_int
    lib.oracle_set_sponge.argtypes = [P, C.c_int, C.POINTER(C.c_int32), PD, PD]
    lib.oracle_set_sponge.restype = C.c_int
    lib.oracle_desc_bytes.restype = C.c_int
    _LIBS[name] = lib
    return lib


def _pd(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


ARR = dict(U=0, Ut=1, U_master=2, U_slave=3, Flux=4, gradUx=5, gradUy=6, gradUz=7, gradUx_master=8, gradUy_master=9,
           gradUz_master=10, gradUx_slave=11, gradUy_slave=12, gradUz_slave=13, Flux_slave=100, UPrim=101, Ut_tmp=102)


class OracleRank:
    def __init__(self, lib, desc, keep):
        self.lib = lib
        self.keep = keep
        self.desc = desc
        self.h = lib.oracle_create(C.cast(C.byref(desc), C.c_void_p))
        if not self.h:
            raise RuntimeError("oracle_create failed")
        self.np = desc.N + 1

    def set_face_metrics(self, Ja_Face, Ja_slave):
        """Ja_Face / Ja_slave of the host mesh ([side, q, p, c, dd] = Fortran (dd,c,p,q,side)) for the split form on Gauss nodes."""
        a = np.ascontiguousarray(Ja_Face, dtype=np.float64)
        b = np.ascontiguousarray(Ja_slave, dtype=np.float64)
        assert a.shape == b.shape == (self.desc.nSides, self.np, self.np, 3, 3), a.shape
        self.lib.oracle_set_face_metrics(self.h, _pd(a), _pd(b))

    def set_overintegration(self, otype, mat=None):
        """OverintegrationType 0/1 and OverintegrationMat[out, in] (dg/overintegration.f90:104-116)."""
        m = None if mat is None else np.asfortranarray(mat, dtype=np.float64)
        rc = self.lib.oracle_set_overintegration(self.h, int(otype), None if m is None else _pd(m))
        assert rc == 0, "OverintegrationType not restated"

    def set_overintegration_cons(self, n_under=-1, vdm_n_nunder=None, vdm_nunder_n=None, sj_nunder=None):
        """OverintegrationType 2 (conservative cut-off, overintegration.f90:118-135,181-330): Vdm_N_NUnder [NUnder+1, N+1],
        Vdm_NUnder_N [N+1, NUnder+1], sJNUnder [nElems, kU, jU, iU]; n_under < 0 = off."""
        if n_under < 0:
            assert self.lib.oracle_set_overintegration_cons(self.h, -1, None, None, None) == 0
            return
        a = np.asfortranarray(vdm_n_nunder, dtype=np.float64)
        b = np.asfortranarray(vdm_nunder_n, dtype=np.float64)
        c = np.ascontiguousarray(sj_nunder, dtype=np.float64)
        assert self.lib.oracle_set_overintegration_cons(self.h, int(n_under), _pd(a), _pd(b), _pd(c)) == 0

    def filter_conservative(self, JU):
        """FilterConservative on a (nElems,k,j,i,5) array holding J*U (timedisc.f90:119-122 filters the initial solution this way)."""
        out = np.ascontiguousarray(JU, dtype=np.float64).copy()
        self.lib.oracle_filter_conservative(self.h, _pd(out))
        return out

    def set_filter_laf(self, integration_weight=None, norm_mod=0.0, laf_alpha=1.0):
        """FilterType 3 (LAF, filter.f90:149-186,295-448): call set_filter(1, cut-off FilterMat of NFilter) first; IntegrationWeight
        [nElems,k,j,i] (normalised per element), normMod, LAF_alpha; None = off."""
        if integration_weight is None:
            assert self.lib.oracle_set_filter_laf(self.h, None, 0.0, 1.0) == 0
            return
        iw = np.ascontiguousarray(integration_weight, dtype=np.float64)
        assert self.lib.oracle_set_filter_laf(self.h, _pd(iw), float(norm_mod), float(laf_alpha)) == 0

    def get_filter_laf(self):
        """(eRatio, lim, r) per element: the ElemData arrays LAF_eRatio / LAF_lim / LAF_r."""
        n = self.desc.nElems
        a, b, c = np.zeros(n), np.zeros(n), np.zeros(n)
        self.lib.oracle_get_filter_laf(self.h, _pd(a), _pd(b), _pd(c))
        return a, b, c

    def set_viscosity(self, visc_type=0, Ts=0.0, Tref=1.0, ExpoSuth=0.0, cSuth=0.0):
        """PP_VISC 0/1/2 with the MOD_EOS_Vars values after InitEos (eos.f90:160-192)."""
        assert self.lib.oracle_set_viscosity(self.h, int(visc_type), float(Ts), float(Tref), float(ExpoSuth), float(cSuth)) == 0

    def set_sponge(self, sponge_map=None, sponge_mat=None, base_flow=None):
        """Sponge(Ut) (sponge/sponge.f90:612-678): SpongeMap 1-based element ids [nSp], SpongeMat [nSp, k, j, i] (times the Jacobian),
        SpBaseFlow [nElems, k, j, i, 5]; no arguments = off."""
        if sponge_map is None:
            assert self.lib.oracle_set_sponge(self.h, 0, None, None, None) == 0
            return
        mp = np.ascontiguousarray(sponge_map, dtype=np.int32)
        sm = np.ascontiguousarray(sponge_mat, dtype=np.float64)
        bf = np.ascontiguousarray(base_flow, dtype=np.float64)
        rc = self.lib.oracle_set_sponge(self.h, len(mp), mp.ctypes.data_as(C.POINTER(C.c_int32)), _pd(sm), _pd(bf))
        assert rc == 0, "sponge arrays rejected"

    def set_filter(self, ftype, mat=None):
        """FilterType 0 (none) / 1 (cutoff) / 2 (modal) and FilterMat[out, in]: U is filtered in place at the start of every
        DG operator evaluation (dg.f90:329-330)."""
        m = None if mat is None else np.asfortranarray(mat, dtype=np.float64)
        rc = self.lib.oracle_set_filter(self.h, int(ftype), None if m is None else _pd(m))
        assert rc == 0, "FilterType not restated"

    def filter(self, mat, U):
        """Filter(U, FilterMat) (filter.f90:264-288); U in the module-array layout [e, k, j, i, v]; returns a new array."""
        m = np.asfortranarray(mat, dtype=np.float64)
        out = np.array(U, dtype=np.float64, order="C", copy=True)
        self.lib.oracle_filter(self.h, _pd(m), _pd(out))
        return out

    def array(self, name):
        """numpy VIEW of a module array (Fortran memory; shape is the reversed Fortran shape)."""
        n = C.c_int64(0)
        p = self.lib.oracle_array(self.h, ARR[name], C.byref(n))
        if not p:
            raise KeyError(name)
        a = np.ctypeslib.as_array(p, shape=(n.value,))
        np_ = self.np
        d = self.desc
        nvl = d.nVarLifting
        if name in ("U", "Ut", "Ut_tmp"):
            return a.reshape(d.nElems, np_, np_, np_, 5)
        if name == "UPrim":
            return a.reshape(d.nElems, np_, np_, np_, 6)
        if name in ("gradUx", "gradUy", "gradUz"):
            return a.reshape(d.nElems, np_, np_, np_, nvl)
        if name in ("U_master", "U_slave", "Flux", "Flux_slave"):
            return a.reshape(d.nSides, np_, np_, 5)
        return a.reshape(d.nSides, np_, np_, nvl)

    def close(self):
        if self.h:
            self.lib.oracle_destroy(self.h)
            self.h = None


class OracleWorld:
    """The ranks of one emulated MPI job (one rank per OpenMP thread in the fast build)."""

    def __init__(self, descs_and_keeps, fast=False):
        self.lib = _load(fast)
        assert self.lib.oracle_desc_bytes() == C.sizeof(type(descs_and_keeps[0][0]))
        self.ranks = [OracleRank(self.lib, d, k) for d, k in descs_and_keeps]
        arr = (C.c_void_p * len(self.ranks))(*[r.h for r in self.ranks])
        self.w = self.lib.oracle_world_create(arr, len(self.ranks))

    def set_state(self, Us):
        for r, U in zip(self.ranks, Us):
            r.array("U")[...] = U

    def dg_time_derivative(self, t=0.0):
        self.lib.oracle_world_dg_time_derivative(self.w, float(t))

    def timestep(self, t, dt):
        self.lib.oracle_world_timestep(self.w, float(t), float(dt))

    def calc_timestep(self):
        a, b, e = C.c_double(0), C.c_double(0), C.c_int(0)
        dt = self.lib.oracle_world_calc_timestep(self.w, C.byref(a), C.byref(b), C.byref(e))
        return dt, a.value, b.value, e.value

    def close(self):
        if self.w:
            self.lib.oracle_world_destroy(self.w)
            self.w = None
        for r in self.ranks:
            r.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def riemann_point(rank, solver, U_L, U_R, nv, t1, t2):
    F = np.zeros(5)
    args = [np.ascontiguousarray(x, dtype=np.float64) for x in (U_L, U_R, nv, t1, t2)]
    rank.lib.oracle_riemann_point(rank.h, int(solver), _pd(F), *[_pd(a) for a in args])
    return F


def split_volume_flux(rank, stype, URef, U, MRef, M):
    F = np.zeros(5)
    args = [np.ascontiguousarray(x, dtype=np.float64) for x in (URef, U, MRef, M)]
    rank.lib.oracle_split_volume_flux(rank.h, int(stype), *[_pd(a) for a in args], _pd(F))
    return F


def eval_flux3d(rank, U):
    f, g, h = np.zeros(5), np.zeros(5), np.zeros(5)
    U = np.ascontiguousarray(U, dtype=np.float64)
    rank.lib.oracle_eval_flux3d(rank.h, _pd(U), _pd(f), _pd(g), _pd(h))
    return f, g, h


def prolong_to_face(rank, nv, Uvol, doMPISides=False):
    d = rank.desc
    np_ = rank.np
    Uvol = np.ascontiguousarray(Uvol, dtype=np.float64)
    um = np.zeros((d.nSides, np_, np_, nv))
    us = np.zeros((d.nSides, np_, np_, nv))
    rank.lib.oracle_prolong_to_face(rank.h, nv, _pd(Uvol), _pd(um), _pd(us), int(doMPISides))
    return um, us


def surf_int(rank, nv, Flux_master, Flux_slave, Ut, L_HatMinus, L_HatPlus, doMPISides=False):
    fm = np.ascontiguousarray(Flux_master, dtype=np.float64)
    fs = np.ascontiguousarray(Flux_slave, dtype=np.float64)
    lm = np.ascontiguousarray(L_HatMinus, dtype=np.float64)
    lp = np.ascontiguousarray(L_HatPlus, dtype=np.float64)
    rank.lib.oracle_surf_int(rank.h, nv, _pd(fm), _pd(fs), _pd(Ut), int(doMPISides), _pd(lm), _pd(lp))
    return Ut


def ut_noise_floor(descs_and_keeps, Us, t=0.0):
    """Per-variable max|Ut_strict - Ut_fast| of ONE DG operator evaluation between the two oracle builds (strict IEEE order
    vs. -O3 with FMA contraction, i.e. the flags the reference itself is built with, cmake/SetCompiler.cmake:83).

    This is the reference's own compiler-flag noise: on ill-conditioned states (TGV at Ma=0.1: p0 ~ 71 against rho*u^2 ~ 1, the
    pressure gradient cancels to 1e-5 of its terms) two faithful builds of the same Fortran differ by ~1e-11 relative, so a
    parity gate cannot be tighter than that there.  Returns (noise[5], Ut_strict per rank)."""
    out = []
    for fast in (False, True):
        w = OracleWorld(descs_and_keeps, fast=fast)
        w.set_state(Us)
        w.dg_time_derivative(t)
        out.append([r.array("Ut").copy() for r in w.ranks])
        w.close()
    noise = np.zeros(5)
    for a, b in zip(*out):
        noise = np.maximum(noise, np.max(np.abs(a - b).reshape(-1, 5), axis=0))
    return noise, out[0]
