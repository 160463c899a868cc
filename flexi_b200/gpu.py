"""Host-side mirror of the reference's operator interface for the hot path, on top of the C ABI (include/flexi_gpu.h).

Method names follow the reference routines they replace (file:line under /root/reference/src):

    FlexiGpu.DGTimeDerivative_weakForm(t)   dg/dg.f90:221
    FlexiGpu.TimeStep(t, dt)                timedisc/timestep.f90:82   (TimeStepByLSERKW2)
    FlexiGpu.CalcTimeStep()                 equations/navierstokes/calctimestep.f90:89
    FlexiGpu.ProlongToFace / SurfInt        interpolation/prolongtoface.t90:21, dg/surfint.t90:21

Arrays cross this boundary as numpy arrays with the memory layout of the Fortran module arrays (numpy shape = reversed
Fortran shape, e.g. U[iElem, k, j, i, var]).  There is no CPU fallback: a missing library or a missing GPU raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .desc import FlexiGpuDesc

_HERE = os.path.dirname(os.path.abspath(__file__))
# FLEXI_GPU_LIB: the development library of the timing tools (tools/*.py); there is no fallback either way
LIB_PATH = os.environ.get("FLEXI_GPU_LIB") or os.path.join(_HERE, "lib", "libflexi_gpu.so")
_LIB = None

ARRAYS = dict(U=0, Ut=1, U_master=2, U_slave=3, Flux=4, gradUx=5, gradUy=6, gradUz=7, gradUx_master=8, gradUy_master=9,
              gradUz_master=10, gradUx_slave=11, gradUy_slave=12, gradUz_slave=13)
KERNEL_CLASSES = ("volume", "flux", "update", "prolong", "calc_timestep", "layout", "wait_trace", "wait_grad", "wait_flux")

# every symbol include/flexi_gpu.h declares
EXPORTS = ("flexi_gpu_last_error", "flexi_gpu_abi_version", "flexi_gpu_nccl_unique_id", "flexi_gpu_init", "flexi_gpu_finalize",
           "flexi_gpu_set_state", "flexi_gpu_get_state", "flexi_gpu_dg_time_derivative", "flexi_gpu_timestep",
           "flexi_gpu_timestep_n", "flexi_gpu_calc_timestep", "flexi_gpu_get_array", "flexi_gpu_keep_gradients",
           "flexi_gpu_prolong_to_face", "flexi_gpu_surf_int", "flexi_gpu_analyze_tgv", "flexi_gpu_calc_error_norms", "flexi_gpu_calc_body_forces", "flexi_gpu_set_face_metrics", "flexi_gpu_set_overintegration", "flexi_gpu_set_overintegration_cons", "flexi_gpu_set_filter", "flexi_gpu_set_filter_laf", "flexi_gpu_get_filter_laf", "flexi_gpu_set_sponge", "flexi_gpu_set_viscosity", "flexi_gpu_pin_host_array", "flexi_gpu_unpin_host_array", "flexi_gpu_counters", "flexi_gpu_kernel_times",
           "flexi_gpu_set_profiling", "flexi_gpu_synchronize", "flexi_gpu_stream")


class FlexiGpuError(RuntimeError):
    pass


def build(force=False):
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j", str(os.cpu_count() or 4), "-s"]
    if force:
        cmd.append("-B")
    subprocess.check_call(cmd)


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise FlexiGpuError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    P, PD, I32 = C.c_void_p, C.POINTER(C.c_double), C.c_int32
    L.flexi_gpu_last_error.restype = C.c_char_p
    L.flexi_gpu_abi_version.restype = C.c_int
    L.flexi_gpu_nccl_unique_id.argtypes = [P]
    L.flexi_gpu_init.argtypes = [C.POINTER(FlexiGpuDesc), C.POINTER(P)]
    L.flexi_gpu_finalize.argtypes = [P]
    L.flexi_gpu_set_state.argtypes = [P, PD]
    L.flexi_gpu_get_state.argtypes = [P, PD]
    L.flexi_gpu_dg_time_derivative.argtypes = [P, C.c_double, PD]
    L.flexi_gpu_timestep.argtypes = [P, PD, C.c_double]
    L.flexi_gpu_timestep_n.argtypes = [P, PD, C.c_double, I32]
    L.flexi_gpu_calc_timestep.argtypes = [P, PD, PD, PD, C.POINTER(I32)]
    L.flexi_gpu_get_array.argtypes = [P, I32, PD]
    L.flexi_gpu_keep_gradients.argtypes = [P, I32]
    L.flexi_gpu_prolong_to_face.argtypes = [P, PD, PD, PD, I32]
    L.flexi_gpu_surf_int.argtypes = [P, PD, PD, PD, I32]
    L.flexi_gpu_analyze_tgv.argtypes = [P, I32, PD, PD, PD]
    L.flexi_gpu_calc_error_norms.argtypes = [P, C.c_double, I32, I32, PD, PD, PD, PD, PD]
    L.flexi_gpu_calc_body_forces.argtypes = [P, PD, PD, PD, PD, PD, PD, PD]
    L.flexi_gpu_set_face_metrics.argtypes = [P, PD, PD]
    L.flexi_gpu_set_overintegration.argtypes = [P, C.c_int, PD]
    L.flexi_gpu_set_filter.argtypes = [P, C.c_int, PD]
    L.flexi_gpu_set_overintegration_cons.argtypes = [P, I32, PD, PD, PD]
    L.flexi_gpu_set_filter_laf.argtypes = [P, PD, PD, C.c_double, C.c_double]
    L.flexi_gpu_get_filter_laf.argtypes = [P, PD, PD, PD]
    L.flexi_gpu_set_sponge.argtypes = [P, I32, C.POINTER(C.c_int32), PD, PD]
    L.flexi_gpu_set_viscosity.argtypes = [P, I32, C.c_double, C.c_double, C.c_double, C.c_double]
    L.flexi_gpu_pin_host_array.argtypes = [P, C.c_int64]
    L.flexi_gpu_unpin_host_array.argtypes = [P]
    L.flexi_gpu_counters.argtypes = [P, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.flexi_gpu_kernel_times.argtypes = [P, PD, C.POINTER(C.c_int64), I32, I32]
    L.flexi_gpu_set_profiling.argtypes = [P, I32]
    L.flexi_gpu_synchronize.argtypes = [P]
    L.flexi_gpu_stream.argtypes = [P, C.POINTER(P)]
    _LIB = L
    return L


def _pd(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def nccl_unique_id():
    """Rank 0 creates the NCCL bootstrap id the host broadcasts (stands in for MPI_COMM_FLEXI, mpi/mpi.f90:89-109)."""
    buf = C.create_string_buffer(128)
    L = lib()
    if L.flexi_gpu_nccl_unique_id(buf) != 0:
        raise FlexiGpuError(L.flexi_gpu_last_error().decode())
    return bytes(buf.raw)


class FlexiGpu:
    """One FLEXI rank's device-side DG operator + time stepper (InitDG/InitLifting/InitTimeDisc equivalent at construction)."""

    def __init__(self, host_rank):
        self.L = lib()
        self.host = host_rank
        d = host_rank.desc
        self.np = d.N + 1
        self.nElems, self.nSides, self.nvl = d.nElems, d.nSides, d.nVarLifting
        self.h = C.c_void_p()
        self._ck(self.L.flexi_gpu_init(C.byref(d), C.byref(self.h)))

    def _ck(self, rc):
        if rc != 0:
            raise FlexiGpuError(f"flexi_gpu error {rc}: {self.L.flexi_gpu_last_error().decode()}")

    # ---- shapes of the module arrays (reversed Fortran shapes) ----
    def _shape(self, name):
        n = self.np
        if name in ("U", "Ut"):
            return (self.nElems, n, n, n, 5)
        if name in ("gradUx", "gradUy", "gradUz"):
            return (self.nElems, n, n, n, self.nvl)
        if name in ("U_master", "U_slave", "Flux"):
            return (self.nSides, n, n, 5)
        return (self.nSides, n, n, self.nvl)

    # ---- state ----
    def set_state(self, U):
        U = np.ascontiguousarray(U, dtype=np.float64)
        assert U.shape == self._shape("U"), (U.shape, self._shape("U"))
        self._ck(self.L.flexi_gpu_set_state(self.h, _pd(U)))

    def get_state(self, out=None):
        out = np.empty(self._shape("U")) if out is None else out
        self._ck(self.L.flexi_gpu_get_state(self.h, _pd(out)))
        return out

    def get_array(self, name):
        out = np.empty(self._shape(name))
        self._ck(self.L.flexi_gpu_get_array(self.h, ARRAYS[name], _pd(out)))
        return out

    TGV_SUMS = ("T_mean", "Entropy", "Ekin", "Ekin_comp", "Enstr", "DR_u", "DR_S", "DR_Sd", "DR_p", "ED_S", "ED_D", "Vorticity_max")

    def AnalyzeTestcase_sums(self, Vdm_GaussN_NAnalyze, wGPVolAnalyze):
        """Steps 1-4 of the TGV AnalyzeTestcase (testcase.f90:241-420) on the device: the un-normalised per-rank integrals.
        Vdm_GaussN_NAnalyze: numpy [NAnalyze+1, N+1] (row = analyze point); wGPVolAnalyze: numpy [k, j, i]."""
        na1 = Vdm_GaussN_NAnalyze.shape[0]
        vdm_f = np.ascontiguousarray(Vdm_GaussN_NAnalyze.T, dtype=np.float64)  # Fortran (0:NAnalyze,0:N): first index fastest
        w = np.ascontiguousarray(wGPVolAnalyze, dtype=np.float64)
        assert vdm_f.shape == (self.np, na1) and w.shape == (na1, na1, na1)
        out = np.zeros(len(self.TGV_SUMS))
        self._ck(self.L.flexi_gpu_analyze_tgv(self.h, na1 - 1, _pd(vdm_f), _pd(w), _pd(out)))
        return dict(zip(self.TGV_SUMS, out))

    def CalcErrorNorms_sums(self, Time, AnalyzeExactFunc, Vdm_GaussN_NAnalyze, wGPVolAnalyze, Elem_xGP):
        """CalcErrorNorms (analyze.f90:391-485) per rank on the device: (sum of squared L2 contributions, Linf) per variable;
        the host reduces over ranks and takes sqrt(./Vol)."""
        na1 = Vdm_GaussN_NAnalyze.shape[0]
        vdm_f = np.ascontiguousarray(Vdm_GaussN_NAnalyze.T, dtype=np.float64)
        w = np.ascontiguousarray(wGPVolAnalyze, dtype=np.float64)
        x = np.ascontiguousarray(Elem_xGP, dtype=np.float64)
        l2, li = np.zeros(5), np.zeros(5)
        self._ck(self.L.flexi_gpu_calc_error_norms(self.h, float(Time), int(AnalyzeExactFunc), na1 - 1, _pd(vdm_f), _pd(w), _pd(x), _pd(l2),
                                                   _pd(li)))
        return l2, li

    def CalcBodyForces(self, wGPSurf, Face_xGP_bc, nBCs, MomOrigin=(0.0, 0.0, 0.0)):
        """CalcBodyForces (calcbodyforces.f90:46-250) per rank on the device: (Fp, Fv, Mp, Mv), each [nBCs, 3]."""
        w = np.ascontiguousarray(wGPSurf, dtype=np.float64)
        x = np.ascontiguousarray(Face_xGP_bc, dtype=np.float64)
        mo = np.ascontiguousarray(MomOrigin, dtype=np.float64)
        out = [np.zeros((nBCs, 3)) for _ in range(4)]
        self._ck(self.L.flexi_gpu_calc_body_forces(self.h, _pd(w), _pd(x), _pd(mo), *[_pd(o) for o in out]))
        return out

    @staticmethod
    def tgv_normalise(sums_per_rank, vol, mu0, rho0=1.0):
        """Steps 5-6 of AnalyzeTestcase (testcase.f90:422-474), kept by the host: reduce over ranks, normalise."""
        tot = {k: (max(s[k] for s in sums_per_rank) if k == "Vorticity_max" else sum(s[k] for s in sums_per_rank)) for k in sums_per_rank[0]}
        out = dict(T_mean=tot["T_mean"] / vol, Entropy=tot["Entropy"] / vol, Ekin=tot["Ekin"] / vol, Ekin_comp=tot["Ekin_comp"] / (vol * rho0),
                   Enstr=tot["Enstr"] / (rho0 * vol), DR_u=tot["DR_u"] * mu0 / vol, DR_S=tot["DR_S"] * 2.0 * mu0 / (rho0 * vol),
                   DR_Sd=tot["DR_Sd"] * 2.0 * mu0 / (rho0 * vol), DR_p=-tot["DR_p"] / (rho0 * vol), Vorticity_max=tot["Vorticity_max"],
                   ED_S=tot["ED_S"] / vol, ED_D=tot["ED_D"] * 4.0 / 3.0 / vol)
        out["uPrime"] = np.sqrt(2.0 / 3.0 * out["Ekin"])
        return out

    def set_face_metrics(self, Ja_Face, Ja_slave):
        """Ja_Face / Ja_slave ([side, q, p, c, dd] = Fortran (dd,c,p,q,side)) for the split form on Gauss nodes."""
        a = np.ascontiguousarray(Ja_Face, dtype=np.float64)
        b = np.ascontiguousarray(Ja_slave, dtype=np.float64)
        self._ck(self.L.flexi_gpu_set_face_metrics(self.h, _pd(a), _pd(b)))

    def set_overintegration(self, otype, mat=None):
        """OverintegrationType (0 none, 1 cutoff) and OverintegrationMat[out, in] (dg/overintegration.f90:104-116)."""
        m = None if mat is None else np.asfortranarray(mat, dtype=np.float64)
        self._ck(self.L.flexi_gpu_set_overintegration(self.h, int(otype), None if m is None else _pd(m)))

    def set_filter(self, ftype, mat=None):
        """FilterType (0 none, 1 cutoff, 2 modal) and FilterMat[out, in]; U is filtered at the start of every DG operator call."""
        m = None if mat is None else np.asfortranarray(mat, dtype=np.float64)
        self._ck(self.L.flexi_gpu_set_filter(self.h, int(ftype), None if m is None else _pd(m)))

    def set_overintegration_cons(self, n_under=-1, vdm_n_nunder=None, vdm_nunder_n=None, sj_nunder=None):
        """OverintegrationType 2 (conservative cut-off): Vdm_N_NUnder [NUnder+1, N+1], Vdm_NUnder_N [N+1, NUnder+1],
        sJNUnder [nElems, kU, jU, iU]; n_under < 0 = off."""
        if n_under < 0:
            self._ck(self.L.flexi_gpu_set_overintegration_cons(self.h, -1, None, None, None))
            return
        a = np.asfortranarray(vdm_n_nunder, dtype=np.float64)
        b = np.asfortranarray(vdm_nunder_n, dtype=np.float64)
        c = np.ascontiguousarray(sj_nunder, dtype=np.float64)
        self._ck(self.L.flexi_gpu_set_overintegration_cons(self.h, int(n_under), _pd(a), _pd(b), _pd(c)))

    def set_filter_laf(self, mat, integration_weight, norm_mod, laf_alpha=1.0):
        """FilterType 3 (LAF): cut-off FilterMat[out, in] of NFilter, IntegrationWeight [nElems,k,j,i], normMod, LAF_alpha."""
        m = np.asfortranarray(mat, dtype=np.float64)
        iw = np.ascontiguousarray(integration_weight, dtype=np.float64)
        self._ck(self.L.flexi_gpu_set_filter_laf(self.h, _pd(m), _pd(iw), float(norm_mod), float(laf_alpha)))

    def get_filter_laf(self):
        """(eRatio, lim, r) per element: the ElemData arrays LAF_eRatio / LAF_lim / LAF_r."""
        n = self.nElems
        a, b, c = np.zeros(n), np.zeros(n), np.zeros(n)
        self._ck(self.L.flexi_gpu_get_filter_laf(self.h, _pd(a), _pd(b), _pd(c)))
        return a, b, c

    def set_viscosity(self, visc_type=0, Ts=0.0, Tref=1.0, ExpoSuth=0.0, cSuth=0.0):
        """PP_VISC (0 constant, 1 Sutherland, 2 power law) with the MOD_EOS_Vars values after InitEos (see sutherland_params)."""
        self._ck(self.L.flexi_gpu_set_viscosity(self.h, int(visc_type), float(Ts), float(Tref), float(ExpoSuth), float(cSuth)))

    def set_sponge(self, sponge_map=None, sponge_mat=None, base_flow=None):
        """Sponge(Ut) (sponge.f90:612-678): SpongeMap 1-based element ids [nSp], SpongeMat [nSp, k, j, i] (times the Jacobian),
        SpBaseFlow [nElems, k, j, i, 5]; no arguments = off."""
        if sponge_map is None:
            self._ck(self.L.flexi_gpu_set_sponge(self.h, 0, None, None, None))
            return
        mp = np.ascontiguousarray(sponge_map, dtype=np.int32)
        sm = np.ascontiguousarray(sponge_mat, dtype=np.float64)
        bf = np.ascontiguousarray(base_flow, dtype=np.float64)
        self._ck(self.L.flexi_gpu_set_sponge(self.h, len(mp), mp.ctypes.data_as(C.POINTER(C.c_int32)), _pd(sm), _pd(bf)))

    def keep_gradients(self, on=True):
        self._ck(self.L.flexi_gpu_keep_gradients(self.h, int(on)))

    # ---- the operator ----
    def DGTimeDerivative_weakForm(self, t=0.0, want_ut=True):
        if want_ut:
            out = np.empty(self._shape("Ut"))
            self._ck(self.L.flexi_gpu_dg_time_derivative(self.h, float(t), _pd(out)))
            return out
        self._ck(self.L.flexi_gpu_dg_time_derivative(self.h, float(t), None))
        return None

    def TimeStep(self, t, dt, n_steps=1):
        tt = C.c_double(float(t))
        self._ck(self.L.flexi_gpu_timestep_n(self.h, C.byref(tt), float(dt), int(n_steps)))

    def CalcTimeStep(self):
        a, b, c, e = C.c_double(0), C.c_double(0), C.c_double(0), C.c_int32(0)
        self._ck(self.L.flexi_gpu_calc_timestep(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(e)))
        return a.value, b.value, c.value, e.value

    def ProlongToFace(self, Uvol, Uface_master, Uface_slave, doMPISides=False):
        Uvol = np.ascontiguousarray(Uvol, dtype=np.float64)
        self._ck(self.L.flexi_gpu_prolong_to_face(self.h, _pd(Uvol), _pd(Uface_master), _pd(Uface_slave), int(doMPISides)))
        return Uface_master, Uface_slave

    def SurfInt(self, Flux_master, Flux_slave, Ut, doMPISides=False):
        fm = np.ascontiguousarray(Flux_master, dtype=np.float64)
        fs = np.ascontiguousarray(Flux_slave, dtype=np.float64)
        self._ck(self.L.flexi_gpu_surf_int(self.h, _pd(fm), _pd(fs), _pd(Ut), int(doMPISides)))
        return Ut

    # ---- accounting ----
    def counters(self):
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        self._ck(self.L.flexi_gpu_counters(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(kernel_launches=a.value, h2d_bytes=b.value, d2h_bytes=c.value)

    def set_profiling(self, on):
        self._ck(self.L.flexi_gpu_set_profiling(self.h, int(on)))

    def kernel_times(self, reset=False):
        n = len(KERNEL_CLASSES)
        ms = (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        self._ck(self.L.flexi_gpu_kernel_times(self.h, ms, cnt, n, int(reset)))
        return {k: (ms[i], cnt[i]) for i, k in enumerate(KERNEL_CLASSES)}

    def synchronize(self):
        self._ck(self.L.flexi_gpu_synchronize(self.h))

    def stream(self):
        s = C.c_void_p()
        self._ck(self.L.flexi_gpu_stream(self.h, C.byref(s)))
        return s.value

    def close(self):
        if self.h:
            self.L.flexi_gpu_finalize(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
