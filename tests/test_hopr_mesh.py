"""The reference's own mesh and state files as input (HOPR ingestion; BASELINE configs[3]: curved unstructured
NACA0012 mesh): regressioncheck/checks/naca/3D/NACA0012_652_Ng2_mesh.h5 (652 curved hexahedra, NGeo = 2, wall + far-field BCs,
periodic in z over one element layer, flips 0/1/2/4) and ..._Referenz_0000010.000000000.h5 (DG_Solution at t = 10, N = 3, Gauss
nodes, plus the per-element time step the reference itself computed), byte-identical copies under tests/golden/reference_naca3d/.

What these files pin (the reference cannot be run here -- no Fortran compiler, no libhdf5):
  * the HDF5 subset reader against the file's own attributes,
  * mesh_readin / prepare_mesh restated in hostmesh (side count = the file's nUniqueSides, watertight flips),
  * metrics.f90 restated in hostmesh + the oracle's operator: free-stream preservation on the curved mesh,
  * metrics + EOS + CalcTimeStep against ElemData 'dt' of the state file: equal to round-off (1e-13) in every element whose
    state did not change over the reference's last time step, within that step's change (1e-3) everywhere."""
import os

import numpy as np
import pytest

from flexi_b200.hostmesh import basis as B
from flexi_b200.hostmesh import mappings as MP
from flexi_b200.hostmesh.analyze import calc_timestep_elem
from flexi_b200.hostmesh.case import build_case
from flexi_b200.hostmesh.hdf5_min import File
from flexi_b200.hostmesh.hopr import HoprMesh, read_state
from flexi_b200.hostmesh.mesh import build_mesh
from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_naca3d")
MESH = os.path.join(GOLD, "NACA0012_652_Ng2_mesh.h5")
STATE = os.path.join(GOLD, "NACA0012_Re5000_AoA8_3D_Referenz_0000010.000000000.h5")
# regressioncheck/checks/naca/3D/parameter.ini: N = 3, Gauss nodes, BR1, Riemann = 33 (RoeEntropyFix), CFL / DFL 0.9
REF_PRIM = (1.0, 0.990268069, 0.139173101, 0.0, 4.4642857)
NACA_EQN = dict(parabolic=True, riemann="roeentropyfix", kappa=1.4, R=2.857142857, Pr=0.72, mu0=0.0002, cfl_scale=0.9,
                dfl_scale=0.9, ref_state_prim=[REF_PRIM])
# BCType of the file: wall = 3 (adiabatic), inflow / outflow = 2 with state 0 (Dirichlet from IniExactFunc = 1, the constant
# RefState 1): stated here as (2, RefState 1), which is the same boundary state
FARFIELD = {"inflow": (2, 1), "outflow": (2, 1)}


def naca_rank(boundary_type=FARFIELD, N=3, node=B.NODE_GAUSS, nranks=1, **eqn):
    hm = HoprMesh(MESH, boundary_type=boundary_type)
    return build_case(hm, N, node, dict(NACA_EQN, boundary_type=hm.boundary_type, **eqn), nranks=nranks)


def test_hdf5_reader_on_the_reference_files():
    f = File(MESH)
    assert int(f.attrs["nElems"][0]) == 652 and int(f.attrs["Ngeo"][0]) == 2 and int(f.attrs["nBCs"][0]) == 5
    assert f["ElemInfo"].shape == (652, 6) and f["SideInfo"].shape == (int(f.attrs["nSides"][0]), 5)
    assert f["NodeCoords"].shape == (int(f.attrs["nNodes"][0]), 3)
    assert [b.decode().strip() for b in f["BCNames"].read()] == ["BC_wall", "BC_inflow", "BC_outflow", "BC_zminus", "BC_zplus"]
    s = File(STATE)  # behind a userblock (the parameter file)
    assert s.userblock.startswith(b"{[( START USERBLOCK )]}") and float(s.attrs["Time"][0]) == 10.0
    assert s.attrs["NodeType"][0] == b"GAUSS" and int(s.attrs["N"][0]) == 3
    assert s["DG_Solution"].shape == (652, 4, 4, 4, 5)
    with pytest.raises(KeyError):
        s["no_such_dataset"]


def test_mesh_connectivity_is_watertight():
    """prepare_mesh restated: the side count is the file's nUniqueSides; every inner side has one master and one slave whose
    face points coincide through S2V2 of the file's flip (flips 1, 2 and 4 occur); normals point out of the master."""
    N = 3
    m = build_mesh(HoprMesh(MESH, boundary_type=FARFIELD), N, B.NODE_GAUSS_LOBATTO)
    assert m.nSides == int(File(MESH).attrs["nUniqueSides"][0]) == 1996
    assert m.nBCSides == 80 and m.nInnerSides == 1916
    flips = np.bincount(m.ElemToSide[:, :, 1].ravel(), minlength=5)
    assert flips[1] > 0 and flips[2] > 0 and flips[4] > 0
    _, _, S2V2 = MP.build_mappings(N)
    from test_host_mesh import face_of

    span = np.ptp(m.Elem_xGP[..., 2])
    for sid in range(m.nBCSides, m.nSides):
        me, se, ms, ss, fl = m.SideToElem[sid]
        xm, xs = face_of(m.Elem_xGP[me - 1], ms, N), face_of(m.Elem_xGP[se - 1], ss, N)
        pq = np.arange(N + 1)
        P, Q = np.meshgrid(pq, pq, indexing="ij")
        a, c = S2V2[:, P, Q, 0, ms - 1]
        a2, c2 = S2V2[:, P, Q, fl, ss - 1]
        d = xm[c, a] - xs[c2, a2]
        d[..., 2] -= span * np.round(d[..., 2] / span)   # the z-periodic pair (an element is its own neighbour there)
        assert np.max(np.abs(d)) < 1e-12, sid
        ctr = m.Elem_xGP[me - 1].reshape(-1, 3).mean(axis=0)
        assert np.dot(m.Face_xGP[sid].reshape(-1, 3).mean(axis=0) - ctr, m.NormVec[sid].reshape(-1, 3).mean(axis=0)) > 0
    # BC sides: the wall is the airfoil (|y| < 0.1, 0 <= x <= 1), the far field is far away
    wall = m.BC[:m.nBCSides] == 1
    xw = m.Face_xGP[:m.nBCSides][wall]
    assert wall.sum() > 0 and np.all(np.abs(xw[..., 1]) < 0.1) and xw[..., 0].min() > -1e-3 and xw[..., 0].max() < 1.0 + 1e-3
    assert np.min(np.hypot(*np.moveaxis(m.Face_xGP[:m.nBCSides][~wall][..., :2], -1, 0))) > 4.9


@pytest.mark.parametrize("node", [B.NODE_GAUSS, B.NODE_GAUSS_LOBATTO])
def test_freestream_preservation_on_the_reference_mesh(node):
    """Uniform flow, every boundary a Dirichlet state: the operator vanishes to round-off on the curved NGeo = 2 mesh (curl-form
    metrics of metrics.f90 interpolated to N = 3 >= NGeo, surface metrics from the same polynomial, every flip's S2V2)."""
    r = naca_rank(dict(FARFIELD, wall=(2, 1)), node=node)[0]
    U0 = np.broadcast_to(r.eqn.prim_to_cons(np.array(REF_PRIM)), r.mesh.Elem_xGP.shape[:-1] + (5,)).copy()
    w = O.OracleWorld([(r.desc, r.keep)])
    w.set_state([U0])
    w.dg_time_derivative(0.0)
    ut = np.array(w.ranks[0].array("Ut"))
    w.close()
    # scale: sJ reaches 6e4 next to the wall, the flux differences that cancel are O(sJ * |F|)
    assert np.max(np.abs(ut)) < 1e-9, np.max(np.abs(ut))


def test_time_step_of_the_reference_state_matches_the_reference():
    """ElemData 'dt' of the state file is dtElem of the reference's last CalcTimeStep call (calctimestep.f90:186-210), i.e. of
    the state one time step before the one in the file."""
    r = naca_rank()[0]
    U, attrs, ed = read_state(STATE)
    dt = calc_timestep_elem(r.eqn, r.mesh, U)
    rel = np.abs(dt / ed["dt"] - 1.0)
    assert rel.max() < 1e-3                       # everywhere: within the change over one step
    assert np.sum(rel < 1e-12) >= 8               # quiescent elements: identical to round-off
    assert np.sum(rel < 1e-9) >= 40
    # and through the oracle's CalcTimeStep (global minimum)
    w = O.OracleWorld([(r.desc, r.keep)])
    w.set_state([U])
    dt_o = w.calc_timestep()
    w.close()
    assert dt_o[3] == 0 and abs(dt_o[0] - dt.min()) <= 1e-13 * dt.min()
    assert abs(dt_o[0] / ed["dt"].min() - 1.0) < 1e-5


def test_partitioned_file_mesh_gives_the_same_operator():
    """3 ranks (linear split of the element list like mesh_readin.f90:787-797) against 1 rank on the reference state: the MPI side
    lists, MINE / YOUR flips and remote-master geometry of a file mesh."""
    U, _, _ = read_state(STATE)
    r1 = naca_rank()[0]
    w1 = O.OracleWorld([(r1.desc, r1.keep)])
    w1.set_state([U])
    w1.dg_time_derivative(10.0)
    ref = np.array(w1.ranks[0].array("Ut"))
    w1.close()
    ranks = naca_rank(nranks=3)
    w = O.OracleWorld([(r.desc, r.keep) for r in ranks])
    w.set_state([U[r.mesh.elem_gid] for r in ranks])
    w.dg_time_derivative(10.0)
    scale = np.max(np.abs(ref).reshape(-1, 5), axis=0)
    scale[1:4] = scale[1:4].max()
    for r, o in zip(ranks, w.ranks):
        d = np.max(np.abs(np.array(o.array("Ut")) - ref[r.mesh.elem_gid]).reshape(-1, 5), axis=0)
        # a side whose master role moves to the other element is evaluated from the other side (F(U_L,U_R,n) vs -F(U_R,U_L,-n)):
        # round-off, amplified by sJ ~ 6e4 next to the wall (the free-stream test above shows the same 5e-11 absolute noise)
        assert np.all(d <= 1e-11 * scale), (d, scale)
    w.close()


def test_mortar_meshes_are_refused():
    """tutorials/convtest/CART_HEX_PERIODIC_MORTAR_002_mesh.h5 (non-conforming): refused like the GPU library refuses mortar sides."""
    with pytest.raises(NotImplementedError, match="mortar"):
        HoprMesh(os.path.join(GOLD, "CART_HEX_PERIODIC_MORTAR_002_mesh.h5"))


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference tree is only present in the build container")
def test_hdf5_reader_on_every_file_of_the_reference_tree():
    """Every .h5 file the reference ships (meshes incl. mortar and 2-D ones, states, base flows, record-point sets, wall distances):
    all datasets parse and have the shape their dataspace declares."""
    import glob

    files = sorted(glob.glob("/root/reference/**/*.h5", recursive=True))
    assert len(files) > 100
    for p in files:
        f = File(p)
        n = 0
        for k in f.keys():
            try:
                d = f[k]
            except KeyError:
                continue  # a group
            assert d.read().shape == d.shape, (p, k)
            n += 1
        assert n > 0, p


def test_state_writer_round_trip_and_layout(tmp_path):
    """write_state (WriteState, hdf5_output.f90:785-839): the reference's cavity state rewritten -- read back bit-identically, the
    userblock (FLEXI's parameter file) preserved, and the dataspace / datatype messages of the datasets and the header attributes
    encoded byte for byte like in the reference's own file (the image has no libhdf5 to open it with)."""
    from flexi_b200.hostmesh.hopr import write_state

    ref = os.path.join(os.path.dirname(GOLD), "reference_cavity3d", "reggie_cavity_Re100_State_0000001.000000000.h5")
    U, attrs, ed = read_state(ref)
    fr = File(ref)
    p = str(tmp_path / "cavity_Re100_State_0000001.000000000.h5")
    write_state(p, U, 1.0, 2, "GAUSS", "cavity4x4x4_mesh.h5", project_name="cavity_Re100", elem_data=ed, userblock=fr.userblock.rstrip(b"\0"))
    U2, a2, ed2 = read_state(p)
    assert np.array_equal(U, U2) and list(ed) == list(ed2) and all(np.array_equal(ed[k], ed2[k]) for k in ed)
    fw = File(p)
    assert fw.userblock == fr.userblock
    for k in ("Time", "N", "Dimension", "MeshFile", "NodeType", "VarNames", "VarNamesAdd", "Program", "File_Type", "Project_Name", "File_Version"):
        assert np.array_equal(fr.attrs[k], fw.attrs[k]) and fr.attrs[k].dtype == fw.attrs[k].dtype, k
    for name in ("DG_Solution", "ElemData"):
        mr = {t: bytes(d) for t, d in fr._messages(fr.items[name]) if t in (1, 3)}
        mw = {t: bytes(d) for t, d in fw._messages(fw.items[name]) if t in (1, 3)}
        assert mr == mw, name
    # the superblock differs only in the addresses it holds
    br, bw = fr.buf[fr.base:fr.base + 24], fw.buf[fw.base:fw.base + 24]
    assert br == bw
