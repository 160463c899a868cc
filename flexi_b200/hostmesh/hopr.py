"""HOPR mesh files (the reference's mesh format) and FLEXI state files as input of the host stand-in.

`HoprMesh` plays the role of `BoxSpec` for `build_mesh`: it delivers what ReadMesh (mesh/mesh_readin.f90:197-470) takes from the
file -- per element and local side the neighbour element, its local side, the side's global index, master/slave role and flip
(`SideInfo`: SIDE_Type, SIDE_ID, SIDE_nbElemID, SIDE_Flip = 10*nbLocSide + flip, SIDE_BCID), the boundary conditions
(`BCNames`, `BCType`; readBCs, mesh_readin.f90:73-190) and the curved element geometry (`NodeCoords(3,0:NGeo,0:NGeo,0:NGeo,nElems)`
on equidistant points) -- and build_mesh then does what prepare_mesh.f90 / metrics.f90 do with it.  Conforming hexahedral meshes
only: mortar sides (SIDE_nbElemID < 0 / SIDE_Type < 0) are refused, as the GPU library refuses them.

`read_state` loads `DG_Solution` (and `ElemData`) of a FLEXI state file (io_hdf5/hdf5_output.f90:785-839 writes them);
`write_state` writes one.

Host-side harness code; the product library never sees a file."""
import numpy as np

from .hdf5_min import File


class HoprMesh:
    def __init__(self, path, boundary_type=None):
        """boundary_type: optional {BC name: (BC_TYPE, BC_STATE)} overriding the file's types like BoundaryName / BoundaryType of the
        parameter file do (mesh_readin.f90:104-140)."""
        f = File(path)
        self.path = path
        self.ngeo = int(f.attrs["Ngeo"][0])
        ei = f["ElemInfo"].read().astype(np.int64)   # (nElems, 6): Type, Zone, offsetIndSIDE, lastIndSIDE, offsetIndNODE, lastIndNODE
        si = f["SideInfo"].read().astype(np.int64)   # (nSides, 5)
        self.n_global_elems = ei.shape[0]
        if np.any(ei[:, 3] - ei[:, 2] != 6) or np.any(si[:, 2] < 0) or np.any(si[:, 0] < 0):
            raise NotImplementedError(f"{path}: mortar sides (non-conforming mesh) are not supported")
        if np.any(ei[:, 0] % 100 != 8):
            raise NotImplementedError(f"{path}: only hexahedra (element types x08) are supported")
        npe = (self.ngeo + 1) ** 3
        if np.any(ei[:, 5] - ei[:, 4] != npe):
            raise NotImplementedError(f"{path}: elements without the full (NGeo+1)^3 node set")
        self._nodes = f["NodeCoords"].read()
        self._node_off = ei[:, 4]
        self.side_info = si.reshape(self.n_global_elems, 6, 5)
        self.bc_names = [b.decode().strip() for b in f["BCNames"].read()]
        bt = f["BCType"].read().astype(np.int64)     # (nBCs, 4): BCType, curveIndex, BCState, alpha
        self.bc_type = bt
        rows = [(int(r[0]), int(r[2]), int(r[3])) for r in bt]
        for name, (t, s) in (boundary_type or {}).items():
            hit = [i for i, n in enumerate(self.bc_names) if name.lower() in n.lower()]
            if not hit:
                raise KeyError(f"BoundaryName {name} not in {self.bc_names}")
            for i in hit:
                rows[i] = (int(t), int(s), rows[i][2])
        self.boundary_type = rows                    # rows (BC_TYPE, BC_STATE, BC_ALPHA), indexed by SIDE_BCID - 1

    def node_coords(self, gids):
        """NodeCoords of the given elements as [e,k,j,i,3] (i fastest in the file, mesh_readin.f90:480-560)."""
        gids = np.asarray(gids)
        n1 = self.ngeo + 1
        idx = self._node_off[gids][:, None] + np.arange(n1 ** 3)[None, :]
        return self._nodes[idx].reshape(len(gids), n1, n1, n1, 3)

    def connectivity(self, gids):
        """Neighbour tables of the given elements over (elem, locSide): neighbour global element id (0-based, -1 = boundary),
        neighbour local side (1-based), global side index, BC index (1-based, 0 = none), flip of the pair, master flag."""
        s = self.side_info[np.asarray(gids)]
        bcid = s[..., 4]
        bct = np.array([r[0] for r in self.boundary_type] + [0], dtype=np.int64)[np.where(bcid > 0, bcid - 1, len(self.boundary_type))]
        is_bc = (bcid > 0) & (bct != 0) & (bct != 1) & (bct != 100)          # mesh_readin.f90:389-399
        nb = np.where(is_bc, -1, s[..., 2] - 1)
        if np.any((nb < 0) & ~is_bc):
            raise ValueError("side without neighbour and without boundary condition")
        nb_side = np.where(is_bc, 0, s[..., 3] // 10)
        face = np.abs(s[..., 1])
        pair_flip = np.where(is_bc, 0, s[..., 3] % 10)
        master = is_bc | (s[..., 1] > 0)
        return nb, nb_side, face, np.where(is_bc, bcid, 0), pair_flip, master


def read_state(path):
    """(U [nElems,k,j,i,5], attributes, ElemData dict or None) of a FLEXI state file."""
    f = File(path)
    U = f["DG_Solution"].read()
    elem_data = None
    if "ElemData" in f:
        names = [n.decode().strip() for n in f.attrs.get("VarNamesAdd", [])]
        ed = f["ElemData"].read()
        elem_data = {n: ed[:, i] for i, n in enumerate(names)}
    return U, f.attrs, elem_data


def write_state(path, U, time, N, node_type, mesh_file, project_name="flexi_b200", elem_data=None, userblock=b""):
    """A FLEXI state file (WriteState, io_hdf5/hdf5_output.f90:785-839, and its header attributes :1060-1130): `DG_Solution`
    [nElems,k,j,i,5] plus optional `ElemData` {name: array[nElems]}; readable by `read_state` and laid out like the reference's own
    files (same superblock / object-header / datatype encodings; not validated against libhdf5, which this image lacks)."""
    from .hdf5_min import write_file

    def s255(x):
        return np.array([x.ljust(255).encode() for x in ([x] if isinstance(x, str) else x)], dtype="S255")

    data = {"DG_Solution": np.ascontiguousarray(U, dtype=np.float64)}
    attrs = {"VarNames": s255(["Density", "MomentumX", "MomentumY", "MomentumZ", "EnergyStagnationDensity"]),
             "Program": np.array([b"Flexi"], dtype="S5"), "File_Type": np.array([b"State"], dtype="S5"),
             "Project_Name": np.array([project_name.encode()], dtype=f"S{len(project_name)}"), "File_Version": np.array([0.1]),
             "N": np.array([int(N)], dtype=np.int32), "Dimension": np.array([3], dtype=np.int32), "Time": np.array([float(time)]),
             "MeshFile": np.array([mesh_file.encode()], dtype=f"S{len(mesh_file)}"), "NodeType": s255(node_type.upper())}
    if elem_data:
        names = list(elem_data)
        data["ElemData"] = np.ascontiguousarray(np.stack([np.asarray(elem_data[n], dtype=np.float64) for n in names], axis=1))
        attrs["VarNamesAdd"] = s255(names)
    write_file(path, data, attrs, userblock=userblock)


def mortar_topology(path):
    """The non-conforming interfaces of a HOPR mesh as ReadMesh sees them (mesh_readin.f90:333-376): a big side is a SideInfo row with
    SIDE_nbElemID = -MortarType (1: 4 small sides, 2 / 3: 2), followed by the rows of its small virtual sides, whose SIDE_nbElemID
    are the small elements; in the small element the matching row has SIDE_Type < 0 and SIDE_Flip = flip (its local side is found by
    the side index).  Returns (node coordinates [e,k,j,i,3], NGeo, list of dicts {elem, loc_side, type, small: [(elem, loc_side,
    flip), ...]}), elements 0-based, local sides 1-based.  Bookkeeping only: build_mesh does not take mortar meshes yet."""
    f = File(path)
    ngeo = int(f.attrs["Ngeo"][0])
    ei = f["ElemInfo"].read().astype(np.int64)
    si = f["SideInfo"].read().astype(np.int64)
    n1 = ngeo + 1
    nodes = f["NodeCoords"].read()
    X = np.stack([nodes[o:o + n1 ** 3].reshape(n1, n1, n1, 3) for o in ei[:, 4]])
    # local side of every (element, side index) for the small-side lookup
    loc_of = {}
    rows = []
    for e in range(len(ei)):
        r, loc = ei[e, 2], 0
        while r < ei[e, 3]:
            loc += 1
            nb = si[r, 2]
            loc_of[(e, abs(si[r, 1]))] = (loc, si[r, 3] % 10, si[r, 0])
            if nb < 0:
                mtype = int(-nb)
                nsmall = 4 if mtype == 1 else 2
                rows.append((e, loc, mtype, [(int(si[r + 1 + m, 2]) - 1, int(abs(si[r + 1 + m, 1]))) for m in range(nsmall)]))
                r += nsmall
            r += 1
        assert loc == 6, "six local sides per hexahedron"
    out = []
    for e, loc, mtype, small in rows:
        sm = []
        for se, ind in small:
            sloc, flip, stype = loc_of[(se, ind)]
            assert stype < 0, "the small element marks the side as part of a mortar (SIDE_Type < 0)"
            sm.append((se, sloc, int(flip)))
        out.append(dict(elem=e, loc_side=loc, type=mtype, small=sm))
    return X, ngeo, out
