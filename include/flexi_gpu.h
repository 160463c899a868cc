/* flexi_gpu.h -- C ABI of the B200-native DGSEM right-hand side / LSERK time stepper for FLEXI.
 *
 * Drop-in boundary for ONE path of flexi-framework/flexi (file:line under /root/reference/src):
 *
 *   flexi_gpu_dg_time_derivative  replaces  DGTimeDerivative_weakForm(t)           dg/dg.f90:221-678
 *   flexi_gpu_timestep            replaces  TimeStepByLSERKW2(t)                   timedisc/timestep.f90:82-152
 *   flexi_gpu_calc_timestep       replaces  CalcTimeStep(errType)                  equations/navierstokes/calctimestep.f90:89-227
 *   flexi_gpu_prolong_to_face     replaces  ProlongToFaceCons                      interpolation/prolongtoface.t90:21-144
 *   flexi_gpu_surf_int            replaces  SurfIntCons                            dg/surfint.t90:21-239,475-616
 *   flexi_gpu_analyze_tgv         replaces  AnalyzeTestcase steps 1-4 (TGV)        testcase/taylorgreenvortex/testcase.f90:241-420
 *   flexi_gpu_calc_error_norms    replaces  CalcErrorNorms (per-rank part)         analyze/analyze.f90:391-485
 *   flexi_gpu_calc_body_forces    replaces  CalcBodyForces (per-rank part)         equations/navierstokes/calcbodyforces.f90:46-250
 *   halo exchange inside          replaces  Start{Receive,Send}MPIData/Finish...   mpi/mpi.f90:244-445  (NCCL send/recv)
 *
 * The Fortran host keeps everything else (ini parsing, mesh/metrics, HDF5 I/O) and hands its module arrays over ONCE through
 * flexi_gpu_desc (INTEGRATION.md shows the ISO_C_BINDING shim).  All arrays are the host's contiguous ALLOCATABLEs, passed as
 * base address + extents in Fortran (column-major) order with the shapes quoted below; INTEGER is 4 bytes, REAL is 8 bytes
 * (cmake/SetCompiler.cmake:79).  The library COPIES everything at init, owns all device memory and keeps no host pointer.
 *
 * Every entry point returns 0 on success and a negative code on failure (flexi_gpu_last_error() gives the text; the shim
 * turns it into CALL Abort(__STAMP__,...), globals/globals.f90:177-224).  No exception crosses this boundary.
 * One rank <-> one OS thread <-> one GPU; calls are sequential per rank.
 */
#ifndef FLEXI_GPU_H
#define FLEXI_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLEXI_GPU_ABI_VERSION 2

/* PP_NodeType (src/CMakeLists.txt) */
#define FLEXI_NODETYPE_GAUSS 1
#define FLEXI_NODETYPE_GAUSS_LOBATTO 2

/* PRM_RIEMANN_* (equations/navierstokes/riemann.f90:50-62) */
#define FLEXI_RIEMANN_SAME (-1)
#define FLEXI_RIEMANN_LF 1
#define FLEXI_RIEMANN_HLLC 2
#define FLEXI_RIEMANN_ROE 3
#define FLEXI_RIEMANN_ROEL2 32
#define FLEXI_RIEMANN_ROEENTROPYFIX 33
#define FLEXI_RIEMANN_HLL 4
#define FLEXI_RIEMANN_HLLE 5
#define FLEXI_RIEMANN_HLLEM 6
/* SPLIT_DG only (riemann.f90:54-58): bare two-point surface flux, Chandrashekar and Ismail-Roe dissipation */
#define FLEXI_RIEMANN_AVERAGE 0
#define FLEXI_RIEMANN_CH 7
#define FLEXI_RIEMANN_IR 8

/* PRM_SPLITDG_* (equations/navierstokes/splitflux.f90:63-69) */
#define FLEXI_SPLITDG_SD 0
#define FLEXI_SPLITDG_MO 1
#define FLEXI_SPLITDG_DU 2
#define FLEXI_SPLITDG_KG 3
#define FLEXI_SPLITDG_PI 4
#define FLEXI_SPLITDG_CH 5
#define FLEXI_SPLITDG_IR 6 /* not a selectable SplitDG string in the reference (splitflux.f90:96-102), InitSplitDG(SplitDG_in) only */

/* error codes */
#define FLEXI_GPU_OK 0
#define FLEXI_GPU_ERR_ARG (-1)
#define FLEXI_GPU_ERR_CUDA (-2)
#define FLEXI_GPU_ERR_NCCL (-3)
#define FLEXI_GPU_ERR_UNSUPPORTED (-4)

#define FLEXI_GPU_NCCL_ID_BYTES 128

typedef struct flexi_gpu_handle flexi_gpu_handle;

/* Everything the path reads from MOD_*_Vars.  Shapes are Fortran shapes; np = N+1. */
typedef struct flexi_gpu_desc {
  int32_t abi_version; /* FLEXI_GPU_ABI_VERSION */
  int32_t struct_bytes; /* sizeof(flexi_gpu_desc) as seen by the caller */

  /* ---- build options of the reference, run-time here (src/CMakeLists.txt:23-230) ---- */
  int32_t N;           /* PP_N, 1..9 or 11 */
  int32_t nodeType;    /* PP_NodeType */
  int32_t splitDG;     /* SPLIT_DG: 0 = weak form (VolInt_weakForm), 1 = split form (VolInt_splitForm) */
  int32_t parabolic;   /* PARABOLIC: 0 = Euler, 1 = Navier-Stokes */
  int32_t lifting;     /* PP_Lifting: 1 = BR1, 2 = BR2 */
  int32_t nVarLifting; /* PP_nVarLifting: 5 (PP_OPTLIFT=0) or 4 (PP_OPTLIFT=1), idealgas/eos.h:108-127 */

  /* ---- run-time options ---- */
  int32_t splitDGType;           /* SplitDG, splitflux.f90:109-153 */
  int32_t riemann;               /* Riemann, riemann.f90:140-259 */
  int32_t riemannBC;             /* RiemannBC (FLEXI_RIEMANN_SAME = same as inner) */
  int32_t doWeakLifting;         /* lifting.f90:122-125 (must be 0) */
  int32_t doConservativeLifting; /* lifting.f90:122-129 (must be 0) */
  double etaBR2, etaBR2_wall;    /* lifting.f90:130-133 */

  /* ---- mesh extents and side ranges (mesh.f90:257-269); mortar counts must be 0 ---- */
  int32_t nElems, nSides, nBCSides, nInnerSides, nMPISides_MINE, nMPISides_YOUR;

  /* ---- volume geometry, mesh_vars.f90:62-78 ---- */
  const double* Metrics_fTilde; /* (3,0:N,0:N,0:N,nElems) */
  const double* Metrics_gTilde;
  const double* Metrics_hTilde;
  const double* sJ;       /* (0:N,0:N,0:N,nElems) */
  const double* Elem_xGP; /* (3,0:N,0:N,0:N,nElems), only read when a source term / exact function needs it; may be NULL */
  /* ---- surface geometry ---- */
  const double* NormVec;  /* (3,0:N,0:N,nSides) */
  const double* TangVec1;
  const double* TangVec2;
  const double* SurfElem; /* (0:N,0:N,nSides) */
  const double* Face_xGP; /* (3,0:N,0:N,nSides); may be NULL without BCs */
  /* ABI 2: Ja_Face / Ja_slave (3,3,0:N,0:N,nSides) of MOD_Mesh_Vars (mesh/metrics.f90:681-705), read by the split form on Gauss
   * nodes only (surfint.t90:90-137); NULL otherwise (or hand them over later with flexi_gpu_set_face_metrics) */
  const double* Ja_Face;
  const double* Ja_slave;
  /* ---- connectivity, mesh_vars.f90:85-119 ---- */
  const int32_t* SideToElem; /* (5,nSides), 1-based ids, -1 = absent, flexi.h:98-102 */
  const int32_t* ElemToSide; /* (2,6,nElems), flexi.h:105-106 */
  const int32_t* S2V2;       /* (2,0:N,0:N,0:4,1:6), mappings.f90:389-397 */
  const int32_t* BC;         /* (nBCSides) index into BoundaryType */
  int32_t nBCs;
  const int32_t* BoundaryType; /* (3,nBCs): BC_TYPE, BC_STATE, BC_ALPHA */
  int32_t nRefState;
  const double* RefStatePrim; /* (6,nRefState) */

  /* ---- basis, interpolation_vars.f90 + dg_vars.f90:31-102 ---- */
  const double* xGP;        /* (0:N) */
  const double* wGP;        /* (0:N) */
  const double* D;          /* (0:N,0:N) */
  const double* D_T;        /* (0:N,0:N) */
  const double* D_Hat_T;    /* (0:N,0:N) */
  const double* DVolSurf;   /* (0:N,0:N), only read if splitDG */
  const double* L_Minus;    /* (0:N) */
  const double* L_Plus;     /* (0:N) */
  const double* L_HatMinus; /* (0:N) */
  const double* L_HatPlus;  /* (0:N) */

  /* ---- equation of state, eos.f90:110-190 ---- */
  double kappa, R, Pr, mu0;

  /* ---- source term: IniExactFunc / doCalcSource, exactfunc.f90:727-815 (only case 4 is on the path) ---- */
  int32_t doCalcSource;
  int32_t iniExactFunc;
  double iniAmplitude, iniFrequency;
  double advVel[3];
  double iniRefStatePrim[6]; /* RefStatePrim(:,IniRefState) */

  /* ---- time discretisation, timedisc_vars.f90:109-214, timedisc_func.f90:463-527 ---- */
  int32_t nRKStages;
  const double* RKA; /* (2:nRKStages) -> nRKStages-1 values */
  const double* RKb; /* (1:nRKStages) */
  const double* RKc; /* (2:nRKStages) */
  double CFLScale;   /* CFLScale(0) after FillCFL_DFL */
  double DFLScale;   /* DFLScale(0) after FillCFL_DFL */

  /* ---- partition, mpi_vars.f90:53-81 ---- */
  int32_t myRank, nProcs, nNbProcs;
  const int32_t* NbProc;              /* (nNbProcs) */
  const int32_t* nMPISides_MINE_Proc; /* (nNbProcs) */
  const int32_t* nMPISides_YOUR_Proc; /* (nNbProcs) */
  const int32_t* offsetMPISides_MINE; /* (0:nNbProcs) */
  const int32_t* offsetMPISides_YOUR; /* (0:nNbProcs) */
  const void* ncclUniqueId;           /* FLEXI_GPU_NCCL_ID_BYTES from flexi_gpu_nccl_unique_id on rank 0, broadcast by the host
                                         (MPI_Bcast in the shim); NULL when nProcs == 1 */
  int32_t device;                     /* CUDA device ordinal for this rank (local rank) */
  int32_t reserved;
} flexi_gpu_desc;

/* which array flexi_gpu_get_array returns (host receives the Fortran layout of the named module array) */
enum flexi_gpu_array {
  FLEXI_ARR_U = 0,            /* U(5,0:N,0:N,0:N,nElems)               dg_vars.f90 */
  FLEXI_ARR_UT = 1,           /* Ut(5,...)                                          */
  FLEXI_ARR_U_MASTER = 2,     /* U_master(5,0:N,0:N,nSides)                         */
  FLEXI_ARR_U_SLAVE = 3,      /* U_slave                                            */
  FLEXI_ARR_FLUX = 4,         /* Flux_master(5,0:N,0:N,nSides) (== Flux_slave, fillflux.f90:201) */
  FLEXI_ARR_GRADUX = 5,       /* gradUx(nVarLifting,0:N,0:N,0:N,nElems) lifting_vars.f90 */
  FLEXI_ARR_GRADUY = 6,
  FLEXI_ARR_GRADUZ = 7,
  FLEXI_ARR_GRADUX_MASTER = 8, /* gradUx_master(nVarLifting,0:N,0:N,nSides) */
  FLEXI_ARR_GRADUY_MASTER = 9,
  FLEXI_ARR_GRADUZ_MASTER = 10,
  FLEXI_ARR_GRADUX_SLAVE = 11,
  FLEXI_ARR_GRADUY_SLAVE = 12,
  FLEXI_ARR_GRADUZ_SLAVE = 13
};

const char* flexi_gpu_last_error(void);
int flexi_gpu_abi_version(void);

/* rank 0: create the NCCL bootstrap id the host broadcasts (replaces MPI_COMM_FLEXI setup, mpi/mpi.f90:89-109) */
int flexi_gpu_nccl_unique_id(void* out_id_bytes);

/* InitDG/InitLifting/InitTimeDisc/InitMPIvars equivalent for the device side: call right after InitTimeDisc (flexilib.f90:216-245) */
int flexi_gpu_init(const flexi_gpu_desc* desc, flexi_gpu_handle** out);
int flexi_gpu_finalize(flexi_gpu_handle* h);

/* U in Fortran order (5,0:N,0:N,0:N,nElems); U is authoritative on the device between set_state and get_state */
int flexi_gpu_set_state(flexi_gpu_handle* h, const double* U);
int flexi_gpu_get_state(flexi_gpu_handle* h, double* U);

/* DGTimeDerivative_weakForm(t): U -> Ut (and all face / gradient arrays). If Ut != NULL the result is copied to the host. */
int flexi_gpu_dg_time_derivative(flexi_gpu_handle* h, double t, double* Ut);

/* TimeStepByLSERKW2(t): all nRKStages stages with the stage update fused into the last RHS kernel; t is advanced by the
 * caller exactly like in timedisc.f90:190-212 (t is INOUT only to mirror the TimeIntegrator interface, timestep.f90:27-37). */
int flexi_gpu_timestep(flexi_gpu_handle* h, double* t, double dt);
/* n consecutive steps with the same dt without returning to the host in between (the host does this when nCalcTimeStepMax>1) */
int flexi_gpu_timestep_n(flexi_gpu_handle* h, double* t, double dt, int32_t nSteps);

/* CalcTimeStep: min over elements and ranks; errType as calctimestep.f90:157-222 */
int flexi_gpu_calc_timestep(flexi_gpu_handle* h, double* dt_min, double* dt_conv, double* dt_visc, int32_t* errType);

/* analysis access (TGV analysis reads GradUx/y/z, testcase.f90:249; Ut is needed at analyze steps, timedisc_func.f90:323-326) */
int flexi_gpu_get_array(flexi_gpu_handle* h, int32_t which, double* out);

/* Volume gradients normally never leave the SM; with on != 0 every later evaluation also stores gradUx/y/z in device memory so
 * that FLEXI_ARR_GRADU{X,Y,Z} return "the gradients left by the last DG operator call" like the module arrays do
 * (testcase/taylorgreenvortex/testcase.f90:249).  Without it the first FLEXI_ARR_GRADU* request recomputes them from the current U.
 * The gradient TRACES (FLEXI_ARR_GRADU?_MASTER/SLAVE) follow the same rule: flexi_gpu_dg_time_derivative leaves all of them, the RK
 * stages of flexi_gpu_timestep write them on boundary-condition sides only (the split-form Gauss-Lobatto kernels hand n.F_visc of
 * each side to the side kernel instead), and a request after a time step recomputes them from the current U at the last stage time.
 *
 * COLLECTIVE on multi-rank runs: a recomputation is a full DG operator evaluation including the NCCL side exchange, so
 * flexi_gpu_get_array(FLEXI_ARR_GRADU*), flexi_gpu_analyze_tgv and flexi_gpu_dg_time_derivative must be called by ALL ranks of
 * the communicator in the same order (as the reference's analysis routines are); a rank-0-only call deadlocks in ncclSend/Recv.
 * Call flexi_gpu_keep_gradients(h, 1) on every rank to make the getters local.  A recomputation never changes U: the in-place
 * solution filter (flexi_gpu_set_filter) is skipped, boundary states and the source term are evaluated at the last stage time. */
int flexi_gpu_keep_gradients(flexi_gpu_handle* h, int32_t on);

/* AnalyzeTestcase of the Taylor-Green vortex test case, steps 1-4 (testcase/taylorgreenvortex/testcase.f90:241-420): the per-rank
 * integrals over the analyze mesh, computed on the device from U and the lifted gradients left by the last DG operator call
 * (flexi_gpu_keep_gradients; recomputed from the current U if they are not kept).  The host keeps steps 5-7 (MPI_REDUCE,
 * normalisation with Vol / rho0, file output).  Vdm_GaussN_NAnalyze(0:NAnalyze,0:N) and wGPVolAnalyze(0:NAnalyze,0:NAnalyze,
 * 0:NAnalyze) are MOD_Analyze_Vars arrays in Fortran order (analyze.f90:224-278).  sums[FLEXI_TGV_NSUMS] in the order below,
 * un-normalised; FLEXI_TGV_VORT_MAX is the per-rank maximum. */
enum {
  FLEXI_TGV_T_MEAN = 0, FLEXI_TGV_ENTROPY = 1, FLEXI_TGV_EKIN = 2, FLEXI_TGV_EKIN_COMP = 3, FLEXI_TGV_ENSTR = 4, FLEXI_TGV_DR_U = 5,
  FLEXI_TGV_DR_S = 6, FLEXI_TGV_DR_SD = 7, FLEXI_TGV_DR_P = 8, FLEXI_TGV_ED_S = 9, FLEXI_TGV_ED_D = 10, FLEXI_TGV_VORT_MAX = 11,
  FLEXI_TGV_NSUMS = 12
};
int flexi_gpu_analyze_tgv(flexi_gpu_handle* h, int32_t NAnalyze, const double* Vdm_GaussN_NAnalyze, const double* wGPVolAnalyze,
                          double* sums);

/* CalcErrorNorms (analyze/analyze.f90:391-485), per-rank part: U, Elem_xGP and 1/sJ are interpolated to the analyze mesh on the
 * device, ExactFunc(AnalyzeExactFunc, Time, x) is evaluated there (implemented: 1 = RefState, 4 = oblique sine wave, exactfunc.f90:
 * 225-362).  L_2_sum(PP_nVar) = sum (U-U_exact)^2 * wGPVolAnalyze * J (the host reduces over ranks and takes SQRT(./Vol), :478-483),
 * L_Inf(PP_nVar) = per-rank maximum.  Elem_xGP (3,0:N,0:N,0:N,nElems) is uploaded on the first call unless a source term already
 * keeps it on the device (then it may be NULL). */
int flexi_gpu_calc_error_norms(flexi_gpu_handle* h, double Time, int32_t AnalyzeExactFunc, int32_t NAnalyze,
                               const double* Vdm_GaussN_NAnalyze, const double* wGPVolAnalyze, const double* Elem_xGP,
                               double* L_2_sum, double* L_Inf);

/* CalcBodyForces (equations/navierstokes/calcbodyforces.f90:46-250), per-rank part: pressure and friction force / moment on every
 * wall BC (BoundaryType 3, 4, 9; analyze_equation.f90:104-112) from U_master and gradU?_master of the BC sides as left by the last
 * DG operator call.  wGPSurf(0:N,0:N) (MOD_Analyze_Vars), Face_xGP(3,0:N,0:N,1:nBCSides) (may be NULL: no moments; uploaded on the
 * first call unless exact-function BCs already keep it), MomOrigin(3) (NULL = origin).  Outputs (3,nBCs) in Fortran order; the host
 * keeps the MPI_REDUCE and BodyForce = Fv + Fp (:102-114).  Call it where the reference does: after flexi_gpu_dg_time_derivative at
 * an analyze step (timedisc_func.f90:323-326) -- between RK steps U_master/U_slave on the device already hold the traces of the
 * UPDATED state (the stage update prolongs them for the next stage), one stage ahead of the reference's module arrays. */
int flexi_gpu_calc_body_forces(flexi_gpu_handle* h, const double* wGPSurf, const double* Face_xGP, const double* MomOrigin, double* Fp,
                               double* Fv, double* Mp, double* Mv);

/* Operator-level entry points (what the reference's own unit tests drive, unitTests/ProlongToFace.f90, SurfInt.f90).
 * Host arrays in Fortran order; the handle supplies mesh + basis.  doMPISides selects the side range like the reference. */
int flexi_gpu_prolong_to_face(flexi_gpu_handle* h, const double* Uvol, double* Uface_master, double* Uface_slave,
                              int32_t doMPISides);
int flexi_gpu_surf_int(flexi_gpu_handle* h, const double* Flux_master, const double* Flux_slave, double* Ut,
                       int32_t doMPISides);

/* Split form on Gauss nodes (SPLIT_DG with PP_NodeType = GAUSS, i.e. PP_EntropyVars = 1: entropy-projected traces
 * prolongtoface.t90:151-204, two-point surface correction surfint.t90:90-137,190-209): hand over MOD_Mesh_Vars Ja_Face and Ja_slave,
 * (3,3,0:N,0:N,nSides) each, after flexi_gpu_init and before the first DG operator call (the descriptor does not carry them). */
int flexi_gpu_set_face_metrics(flexi_gpu_handle* h, const double* Ja_Face, const double* Ja_slave);

/* Overintegration by modal cut-off (MOD_Overintegration, dg/overintegration.f90:40-47,104-116,162-187; dg.f90:663-676):
 * overintegrationType 0 = none, 1 = OVERINTEGRATIONTYPE_CUTOFF -- Ut is filtered with OverintegrationMat (0:N,0:N) (Fortran order,
 * the host's MOD_Overintegration_Vars array) between the source term and ApplyJacobian of every DG operator evaluation, fused into
 * the update kernel.  2 (conservative cut-off, needs sJNUnder) returns FLEXI_GPU_ERR_UNSUPPORTED.  The initial Overintegration(U)
 * of timedisc.f90:117-122 and the CFL/DFL scaling for NUnder (timedisc_func.f90:142) stay with the host. */
int flexi_gpu_set_overintegration(flexi_gpu_handle* h, int overintegrationType, const double* OverintegrationMat);
/* OverintegrationType = 2, conservative cut-off (dg/overintegration.f90:118-135,181-330; dg.f90:664-668): JU_t is projected from N to
 * NUnder, multiplied by sJNUnder and interpolated back inside the update kernel; ApplyJacobian is skipped.  Pass what
 * InitOverintegration built: Vdm_N_NUnder(0:NUnder,0:N), Vdm_NUnder_N(0:N,0:NUnder), sJNUnder(0:NUnder,0:NUnder,0:NUnder,nElems).
 * NUnder < 0 switches it off.  The initial ApplyJacobianCons + Overintegration(U) of timedisc.f90:119-122 stays in the host. */
int flexi_gpu_set_overintegration_cons(flexi_gpu_handle* h, int32_t NUnder, const double* Vdm_N_NUnder, const double* Vdm_NUnder_N,
                                       const double* sJNUnder);

/* Solution filter (MOD_Filter, filter/filter.f90:67-72,135-196,264-288; dg.f90:329-330): filterType 0 = none, 1 = cutoff,
 * 2 = modal (Hesthaven) -- both apply the host's FilterMat (0:N,0:N) (Fortran order) to U in place at the start of every DG
 * operator evaluation, as the reference does.  3 (LAF) has its own entry below.  On multi-rank runs the traces of
 * the filtered state are prolonged and exchanged at the start of the evaluation (the stages then skip their own trace exchange). */
int flexi_gpu_set_filter(flexi_gpu_handle* h, int filterType, const double* FilterMat);
/* FilterType = 3, LAF (filter/filter.f90:149-186,295-448; Flad et al., JCP 2016): per element, the cut-off filtered state of NFilter
 * replaces U only where the kinetic energy of the filtered-out velocity scales, relative to that of the resolved fluctuations and
 * times normMod, exceeds the element's adaptive limit (lim, updated from the energy transfer between two calls).  Pass the cut-off
 * FilterMat(0:N,0:N) of NFilter, IntegrationWeight(0:N,0:N,0:N,nElems) (normalised per element, filter.f90:166-181), normMod (:184)
 * and LAF_alpha; the per-element state starts as InitFilter leaves it (lim = 1e-8, r = 0, old energies 1e-16).
 * flexi_gpu_get_filter_laf returns the ElemData arrays LAF_eRatio, LAF_lim, LAF_r (nElems each; NULL = skip). */
int flexi_gpu_set_filter_laf(flexi_gpu_handle* h, const double* FilterMat, const double* IntegrationWeight, double normMod,
                             double LAF_alpha);
int flexi_gpu_get_filter_laf(flexi_gpu_handle* h, double* eRatio, double* lim, double* r);

/* Viscosity law, the reference's build option FLEXI_VISCOSITY / PP_VISC (eos.h:73-87, viscosity.f90:57-90): 0 = constant mu0
 * (default), 1 = Sutherland, 2 = power law.  Pass the MOD_EOS_Vars values as InitEos leaves them (eos.f90:160-192): for Sutherland
 * Tref = 1/Tref_ini, Ts = Ts_ini/Tref_ini, cSuth = Ts**ExpoSuth*(1+Ts)/(2*Ts*Ts); for the power law mu0 of the descriptor is already
 * mu0_ini/Tref**ExpoSuth and only ExpoSuth is read.  Used by EvalDiffFlux3D (volume, sides, boundary sides), CalcTimeStep, the TGV
 * analysis and CalcBodyForces.  Call once after flexi_gpu_init. */
int flexi_gpu_set_viscosity(flexi_gpu_handle* h, int32_t viscType, double Ts, double Tref, double ExpoSuth, double cSuth);

/* Sponge(Ut) (sponge/sponge.f90:612-678; dg.f90:659-661): Ut = Ut - SpongeMat*(U - SpBaseFlow) on the nSpongeElems elements of
 * SpongeMap (1-based element ids), between CalcSource and ApplyJacobian of every DG operator evaluation.  The host keeps InitSponge
 * (ramp shapes, base-flow types, sponge.f90:97-400) and passes its results: SpongeMat(0:N,0:N,0:N,nSpongeElems) (already times the
 * Jacobian), SpBaseFlow(PP_nVar,0:N,0:N,0:N,nElems).  nSpongeElems = 0 switches the sponge off.  Call again after the host changed
 * the base flow (pressure-time-average base flow, sponge.f90:560-600). */
int flexi_gpu_set_sponge(flexi_gpu_handle* h, int32_t nSpongeElems, const int32_t* SpongeMap, const double* SpongeMat,
                         const double* SpBaseFlow);

/* Page-lock a host array (cudaHostRegister) so that flexi_gpu_set_state / get_state / get_array move it at full PCIe speed and
 * asynchronously; the Fortran host calls this once for MOD_DG_Vars:U after allocation (dg.f90:94) and unpins before DEALLOCATE. */
int flexi_gpu_pin_host_array(void* ptr, int64_t nbytes);
int flexi_gpu_unpin_host_array(void* ptr);

/* counters for bench.py: kernels launched by this library since init, bytes moved by host<->device copies */
int flexi_gpu_counters(flexi_gpu_handle* h, int64_t* kernel_launches, int64_t* h2d_bytes, int64_t* d2h_bytes);
/* device time [ms] accumulated by CUDA events around each hot-path kernel class since the last reset (index = kernel class) */
int flexi_gpu_kernel_times(flexi_gpu_handle* h, double* ms_out, int64_t* launches_out, int32_t n_classes, int32_t reset);
int flexi_gpu_set_profiling(flexi_gpu_handle* h, int32_t on);
int flexi_gpu_synchronize(flexi_gpu_handle* h);
/* the CUDA stream (cudaStream_t) all hot-path kernels of this handle are launched on, for event timing by the caller */
int flexi_gpu_stream(flexi_gpu_handle* h, void** stream_out);

#ifdef __cplusplus
}
#endif
#endif /* FLEXI_GPU_H */
