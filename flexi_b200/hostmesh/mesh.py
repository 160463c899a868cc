"""Synthetic HOPR-style hexahedral meshes + metrics of the synthetic FLEXI host (stand-in for InitMesh/CalcMetrics).

What the Fortran host would do by reading a HOPR HDF5 file is done here for structured boxes that are partitioned and
numbered with the reference's rules, so that the init descriptor handed to the GPU library looks like the one a real
FLEXI rank would hand over:

  * element list = x-fastest ijk ordering (HOPR `doSortIJK`); ranks own contiguous chunks (mesh_readin.f90:787-797) or
    Cartesian blocks (weak-scaling config 5),
  * side list order [BC | inner | MPI_MINE (by neighbour rank) | MPI_YOUR (by neighbour rank)] (mesh.f90:245-276),
    BC/inner IDs in first-seen order of the (iElem, iLocSide=1..6) loop (prepare_mesh.f90:184-229), MPI sides of one
    neighbour sorted by (negated) global side index, lower rank owns the first floor(n/2) as MINE (prepare_mesh.f90:236-337),
  * on MPI_MINE sides this rank holds the master (flip=0) element, on MPI_YOUR sides the slave (after `exchangeFlip`),
  * SideToElem(5,nSides), ElemToSide(2,6,nElems) with the entry positions of flexi.h:98-106,
  * metrics in curl form on Chebyshev-Lobatto points, interpolated to the solution nodes (metrics.f90:161-542),
    surface metrics from the master element (metrics.f90:601-787); the master->slave geometry exchange for MPI_YOUR
    sides (metrics.f90:558-572) is replaced by building the remote master element's metrics locally (the mesh is analytic).

All arrays are numpy C-order arrays whose memory is identical to the Fortran arrays named in mesh_vars.f90:62-153, i.e. the
numpy shape is the reversed Fortran shape: Metrics_fTilde[iElem, k, j, i, 0:3] <-> Metrics_fTilde(1:3,i,j,k,iElem,0).
Element and side ids stored *inside* SideToElem/ElemToSide are 1-based like in Fortran (-1 = absent).
"""
import numpy as np

from . import basis as B
from . import mappings as MP

# neighbour direction per local side (locSide-1): (axis, sign)
_SIDE_AXIS = {MP.ZETA_MINUS: (2, -1), MP.ETA_MINUS: (1, -1), MP.XI_PLUS: (0, +1),
              MP.ETA_PLUS: (1, +1), MP.XI_MINUS: (0, -1), MP.ZETA_PLUS: (2, +1)}
_OPPOSITE = {MP.ZETA_MINUS: MP.ZETA_PLUS, MP.ZETA_PLUS: MP.ZETA_MINUS, MP.ETA_MINUS: MP.ETA_PLUS,
             MP.ETA_PLUS: MP.ETA_MINUS, MP.XI_MINUS: MP.XI_PLUS, MP.XI_PLUS: MP.XI_MINUS}


def conforming_flip(master_loc_side, slave_loc_side):
    """Flip of the slave side of two index-aligned neighbours: S2V2(:,p,q,flip,slave) == S2V2(:,p,q,0,master)."""
    n = 3
    for f in range(5):
        ok = True
        for p in range(n + 1):
            for q in range(n + 1):
                if MP.side_to_vol2(n, p, q, f, slave_loc_side) != MP.side_to_vol2(n, p, q, 0, master_loc_side):
                    ok = False
        if ok:
            return f
    raise RuntimeError("no flip found")


_FLIP_TABLE = {(s, _OPPOSITE[s]): conforming_flip(s, _OPPOSITE[s]) for s in range(1, 7)}


# ---- element orientations ("unstructured" connectivity on a structured block) --------------------------------------------------
# A HOPR mesh from an unstructured generator gives every hexahedron its own local frame: neighbouring elements meet with any
# pair of local sides and any of the flips 0..4 (mappings.f90:188-215).  An orientation is one of the 24 proper rotations of
# the cube: local axis a (xi, eta, zeta) runs along block axis perm[a] in direction sign[a] (det = +1: positive Jacobians).
def _rotations():
    import itertools

    out = []
    for perm in itertools.permutations(range(3)):
        for sign in itertools.product((1, -1), repeat=3):
            R = np.zeros((3, 3))
            for a in range(3):
                R[perm[a], a] = sign[a]
            if np.linalg.det(R) > 0:
                out.append((perm, sign))
    out.sort(key=lambda o: (o != ((0, 1, 2), (1, 1, 1))))  # orientation 0 = the block-aligned frame
    return out


ORIENTATIONS = _rotations()
N_ORIENT = len(ORIENTATIONS)  # 24
# block direction (axis, sign) of local side s of an element with orientation o, and its inverse
_ORI_GAXIS = np.zeros((N_ORIENT, 6), dtype=np.int64)
_ORI_GSIGN = np.zeros((N_ORIENT, 6), dtype=np.int64)
_ORI_LOCSIDE = np.zeros((N_ORIENT, 3, 2), dtype=np.int64)  # [o, block axis, (sign+1)//2] -> local side (1-based)
for _o, (_perm, _sign) in enumerate(ORIENTATIONS):
    for _s in range(1, 7):
        _ax, _sg = _SIDE_AXIS[_s]
        _ORI_GAXIS[_o, _s - 1] = _perm[_ax]
        _ORI_GSIGN[_o, _s - 1] = _sg * _sign[_ax]
        _ORI_LOCSIDE[_o, _perm[_ax], (_sg * _sign[_ax] + 1) // 2] = _s


def _local_to_block_index(o, n, i, j, k):
    """Block-frame in-cell node index (a0,a1,a2) of local node (i,j,k) of an element with orientation o (n = N)."""
    perm, sign = ORIENTATIONS[o]
    loc = (i, j, k)
    out = [0, 0, 0]
    for a in range(3):
        out[perm[a]] = loc[a] if sign[a] > 0 else n - loc[a]
    return tuple(out)


_ORI_FLIP_CACHE = {}


def oriented_flip(o_m, ls_m, o_s, ls_s):
    """Flip of the slave side: the f with S2V2(p,q,f,ls_s) of the slave == S2V2(p,q,0,ls_m) of the master AS PHYSICAL POINTS
    (mappings.f90:389-397), found by comparing block-frame face-node indices of the two elements (n = 3 suffices)."""
    key = (o_m, ls_m, o_s, ls_s)
    if key in _ORI_FLIP_CACHE:
        return _ORI_FLIP_CACHE[key]
    n = 3
    gax = _ORI_GAXIS[o_m, ls_m - 1]
    assert gax == _ORI_GAXIS[o_s, ls_s - 1] and _ORI_GSIGN[o_m, ls_m - 1] == -_ORI_GSIGN[o_s, ls_s - 1], "sides do not face each other"
    keep = [a for a in range(3) if a != gax]

    def face_pts(o, ls, f):
        pts = {}
        for p in range(n + 1):
            for q in range(n + 1):
                a, b = MP.side_to_vol2(n, p, q, f, ls)
                if ls in (MP.XI_MINUS, MP.XI_PLUS):
                    ijk = (0 if ls == MP.XI_MINUS else n, a, b)
                elif ls in (MP.ETA_MINUS, MP.ETA_PLUS):
                    ijk = (a, 0 if ls == MP.ETA_MINUS else n, b)
                else:
                    ijk = (a, b, 0 if ls == MP.ZETA_MINUS else n)
                blk = _local_to_block_index(o, n, *ijk)
                pts[(p, q)] = (blk[keep[0]], blk[keep[1]])
        return pts

    ref = face_pts(o_m, ls_m, 0)
    for f in range(5):
        if face_pts(o_s, ls_s, f) == ref:
            _ORI_FLIP_CACHE[key] = f
            return f
    raise RuntimeError("no flip found for %r" % (key,))


class BoxSpec:
    """Global description of a structured hexahedral box mesh (what the HOPR file would contain)."""

    def __init__(self, nel, x0=(0.0, 0.0, 0.0), length=(2 * np.pi,) * 3, periodic=(True, True, True), ngeo=1,
                 deform=0.0, bc_of_face=None, mapping=None, orient_seed=None):
        self.nel = tuple(int(v) for v in nel)
        self.x0 = np.asarray(x0, dtype=float)
        self.length = np.asarray(length, dtype=float)
        self.periodic = tuple(bool(p) for p in periodic)
        self.ngeo = int(ngeo)
        self.deform = float(deform)
        # bc_of_face: dict {(axis, sign): bc_index (1-based into BoundaryType)} for non-periodic directions
        self.bc_of_face = bc_of_face or {}
        # optional callable X(xhat) mapping the unit cube [0,1]^3 (array [...,3]) to physical space (curved meshes)
        self.mapping = mapping
        # orient_seed: every element gets a random one of the 24 local frames (None = all block-aligned, the HOPR box mesh)
        n_glob = self.nel[0] * self.nel[1] * self.nel[2]
        if orient_seed is None:
            self.orient = np.zeros(n_glob, dtype=np.int64)
        else:
            self.orient = np.random.default_rng(int(orient_seed)).integers(0, N_ORIENT, size=n_glob)
        for a in range(3):
            if self.periodic[a] and self.nel[a] < 2:
                raise ValueError("periodic directions need at least 2 elements")

    @property
    def n_global_elems(self):
        return self.nel[0] * self.nel[1] * self.nel[2]

    def gid(self, ix, iy, iz):
        return ix + self.nel[0] * (iy + self.nel[1] * iz)

    def ijk(self, gid):
        gid = np.asarray(gid)
        ix = gid % self.nel[0]
        iy = (gid // self.nel[0]) % self.nel[1]
        iz = gid // (self.nel[0] * self.nel[1])
        return ix, iy, iz

    def node_coords(self, gids):
        """NodeCoords(3,0:NGeo,0:NGeo,0:NGeo,nElems) on equidistant (VISU) points -> numpy [e,k,j,i,3].

        Curved variant = the reference's own `meshdeform` (mesh.f90:212-223) x + 0.1 sin(pi x)sin(pi y)sin(pi z),
        applied in box-normalised coordinates xhat in [0,2]^3 and scaled by `deform`/0.1 * length/2.
        """
        ng = self.ngeo
        ix, iy, iz = self.ijk(np.asarray(gids))
        t = np.arange(ng + 1) / float(ng)
        h = 1.0 / np.asarray(self.nel, dtype=float)
        X = np.zeros((len(gids), ng + 1, ng + 1, ng + 1, 3))
        X[..., 0] = ((ix[:, None] + t[None, :]) * h[0])[:, None, None, :]
        X[..., 1] = ((iy[:, None] + t[None, :]) * h[1])[:, None, :, None]
        X[..., 2] = ((iz[:, None] + t[None, :]) * h[2])[:, :, None, None]
        if self.mapping is not None:
            X = self.mapping(X)
        else:
            if self.deform != 0.0:
                xh = 2.0 * X
                d = self.deform * np.sin(np.pi * xh[..., 0]) * np.sin(np.pi * xh[..., 1]) * np.sin(np.pi * xh[..., 2])
                X = X + 0.5 * d[..., None]
            X = self.x0 + X * self.length
        return self._to_local_frames(X, self.orient[np.asarray(gids)])

    @staticmethod
    def _to_local_frames(X, orient):
        """Block-frame node array X[e,k,j,i,3] -> the same nodes indexed in each element's own (xi,eta,zeta) frame."""
        if not np.any(orient):
            return X
        out = np.empty_like(X)
        for o in np.unique(orient):
            sel = np.nonzero(orient == o)[0]
            perm, sign = ORIENTATIONS[o]
            Y = X[sel]  # axes (e, a2, a1, a0, c): block axis b sits at array axis 3 - b
            # local axis a runs along block axis perm[a]: array axis (3 - a) of the result takes array axis (3 - perm[a]) of Y
            Y = np.transpose(Y, (0, 3 - perm[2], 3 - perm[1], 3 - perm[0], 4))
            for a in range(3):
                if sign[a] < 0:
                    Y = np.flip(Y, axis=3 - a)
            out[sel] = Y
        return out


def partition_owner(spec, nranks, kind="linear", blocks=None):
    """Rank owning each global element.

    'linear': contiguous equal chunks of the element list, first ranks get the remainder (mesh_readin.f90:787-797).
    'block' : px*py*pz Cartesian blocks (config 5: what an SFC split of an ijk-sorted mesh approximates).
    """
    ng = spec.n_global_elems
    if nranks == 1:
        return np.zeros(ng, dtype=np.int64)
    if kind == "linear":
        base, rem = divmod(ng, nranks)
        counts = np.array([base + (1 if r < rem else 0) for r in range(nranks)])
        return np.repeat(np.arange(nranks), counts)
    if kind == "block":
        px, py, pz = blocks
        assert px * py * pz == nranks
        ix, iy, iz = spec.ijk(np.arange(ng))
        bx = ix * px // spec.nel[0]
        by = iy * py // spec.nel[1]
        bz = iz * pz // spec.nel[2]
        return bx + px * (by + py * bz)
    raise ValueError(kind)


class HostMesh:
    """MOD_Mesh_Vars + MOD_MPI_Vars of one rank after InitMesh(meshMode=2)."""


def build_mesh(spec, N, node_type, rank=0, nranks=1, partition="linear", blocks=None, chunk=2048, with_ja_face=False):
    m = HostMesh()
    m.spec = spec
    m.N = N
    m.node_type = node_type
    m.rank, m.nranks = rank, nranks
    owner = partition_owner(spec, nranks, partition, blocks)
    gids = np.nonzero(owner == rank)[0]
    nE = len(gids)
    m.nElems = nE
    m.nGlobalElems = spec.n_global_elems
    m.elem_gid = gids
    g2l = -np.ones(spec.n_global_elems, dtype=np.int64)
    g2l[gids] = np.arange(nE)

    # ---- neighbour tables over (elem, locSide) -------------------------------------------------------------------------
    unstructured = hasattr(spec, "connectivity")   # a mesh file (hopr.HoprMesh): the tables come from SideInfo
    if unstructured:
        nb_gid, nb_side, face_gid, bc_idx, pair_flip, file_master = spec.connectivity(gids)
        ori = None
    else:
        ix, iy, iz = spec.ijk(gids)
        idx = np.stack([ix, iy, iz], axis=1)
        ori = spec.orient[gids]
        nb_gid = np.zeros((nE, 6), dtype=np.int64)
        nb_side = np.zeros((nE, 6), dtype=np.int64)       # local side (1-based) of the neighbour that faces this side
        face_gid = np.zeros((nE, 6), dtype=np.int64)
        bc_idx = np.zeros((nE, 6), dtype=np.int64)
        nel = np.asarray(spec.nel)
        npl = nel + 1
        per = np.asarray(spec.periodic)
        bc_lut = np.zeros((3, 2), dtype=np.int64)
        for (ax_, sg_), b_ in spec.bc_of_face.items():
            bc_lut[ax_, (sg_ + 1) // 2] = b_
        ar = np.arange(nE)
        for s in range(1, 7):
            ax, sg = _ORI_GAXIS[ori, s - 1], _ORI_GSIGN[ori, s - 1]   # block direction of this local side, per element
            j = idx.copy()
            j[ar, ax] += sg
            plane = idx[ar, ax] + (sg > 0)                             # face plane index along ax, 0..nel
            outside = (j[ar, ax] < 0) | (j[ar, ax] >= nel[ax])
            wrap = per[ax]
            j[ar, ax] = np.where(wrap, j[ar, ax] % nel[ax], j[ar, ax])
            plane = np.where(wrap, plane % nel[ax], plane)
            is_out = outside & ~wrap
            jc = np.clip(j, 0, nel - 1)
            nb = np.where(is_out, -1, spec.gid(jc[:, 0], jc[:, 1], jc[:, 2]))
            if np.any(is_out):
                bc_idx[is_out, s - 1] = bc_lut[ax[is_out], (sg[is_out] + 1) // 2]
                assert np.all(bc_idx[is_out, s - 1] > 0), "non-periodic face without a bc_of_face entry"
            f = idx.copy()
            f[ar, ax] = plane
            face_gid[:, s - 1] = 1 + ax + 3 * (f[:, 0] + npl[0] * (f[:, 1] + npl[1] * f[:, 2]))   # 1-based global side index ("ind")
            nb_gid[:, s - 1] = nb
            nb_side[:, s - 1] = _ORI_LOCSIDE[spec.orient[np.clip(nb, 0, None)], ax, (1 - sg) // 2]
    nb_rank = np.where(nb_gid >= 0, owner[np.clip(nb_gid, 0, None)], -1)
    is_bc = nb_gid < 0
    is_inner = (~is_bc) & (nb_rank == rank)
    is_mpi = (~is_bc) & (nb_rank != rank)

    side_id = np.zeros((nE, 6), dtype=np.int64)      # 1-based SideID per (elem, locSide)
    flip = np.zeros((nE, 6), dtype=np.int64)
    visit = (np.arange(nE)[:, None] * 6 + np.arange(6)[None, :])

    # BC sides: first-seen order
    nBC = int(is_bc.sum())
    side_id[is_bc] = 1 + np.arange(nBC)               # boolean indexing walks in C order == visit order
    # inner sides: each face is visited twice; number by first visit
    e_in, s_in = np.nonzero(is_inner)
    fg = face_gid[e_in, s_in]
    v = visit[e_in, s_in]
    order = np.lexsort((v, fg))
    fg_s, v_s = fg[order], v[order]
    assert len(fg_s) % 2 == 0 and np.all(fg_s[0::2] == fg_s[1::2]), "inner faces must come in pairs"
    first_v = v_s[0::2]
    nInner = len(first_v)
    rank_of_face = np.empty(nInner, dtype=np.int64)
    rank_of_face[np.argsort(first_v, kind="stable")] = np.arange(nInner)
    sid_pair = 1 + nBC + rank_of_face
    ea, sa = e_in[order][0::2], s_in[order][0::2]
    eb, sb = e_in[order][1::2], s_in[order][1::2]
    side_id[ea, sa] = sid_pair
    side_id[eb, sb] = sid_pair
    # master of an inner side = element with the lower global id
    if unstructured:   # the file says which side is the master (SIDE_ID > 0, mesh_readin.f90:352-360) and the flip of the slave
        a_master = file_master[ea, sa]
        assert np.all(file_master[ea, sa] != file_master[eb, sb]), "every inner side has exactly one master"
    else:
        a_master = gids[ea] < gids[eb]
    me, ms = np.where(a_master, ea, eb), np.where(a_master, sa, sb)
    se, ss = np.where(a_master, eb, ea), np.where(a_master, sb, sa)
    flip[me, ms] = 0
    flip[se, ss] = pair_flip[se, ss] if unstructured else _flips(ori[me], ms + 1, ori[se], ss + 1)

    # MPI sides
    nb_procs = sorted(set(nb_rank[is_mpi].tolist()))
    m.nNbProcs = len(nb_procs)
    m.NbProc = np.array(nb_procs, dtype=np.int32)
    nMINE_proc, nYOUR_proc = [], []
    mpi_lists = []
    for nbp in nb_procs:
        e_m, s_m = np.nonzero(is_mpi & (nb_rank == nbp))
        key = -face_gid[e_m, s_m]
        o = np.argsort(key, kind="stable")
        e_m, s_m = e_m[o], s_m[o]
        n = len(e_m)
        nmine = n // 2 if rank < nbp else n - n // 2
        nMINE_proc.append(nmine)
        nYOUR_proc.append(n - nmine)
        mpi_lists.append((e_m, s_m))
    off_mine = [nBC + nInner]
    for c in nMINE_proc:
        off_mine.append(off_mine[-1] + c)
    off_your = [off_mine[-1]]
    for c in nYOUR_proc:
        off_your.append(off_your[-1] + c)
    is_mine = np.zeros((nE, 6), dtype=bool)
    for inb, nbp in enumerate(nb_procs):
        e_m, s_m = mpi_lists[inb]
        n = len(e_m)
        pos = 1 + np.arange(n)
        if rank < nbp:
            mine = pos <= nMINE_proc[inb]
            sid = np.where(mine, pos + off_mine[inb], pos - nMINE_proc[inb] + off_your[inb])
        else:
            mine = pos > nYOUR_proc[inb]
            sid = np.where(mine, pos - nYOUR_proc[inb] + off_mine[inb], pos + off_your[inb])
        side_id[e_m, s_m] = sid
        is_mine[e_m, s_m] = mine
        if unstructured:   # a YOUR side keeps the pair's flip whichever side the file called master (the flips are involutions)
            flip[e_m, s_m] = np.where(mine, 0, pair_flip[e_m, s_m])
        else:
            rem_o = spec.orient[nb_gid[e_m, s_m]]
            flip[e_m, s_m] = np.where(mine, 0, _flips(rem_o, nb_side[e_m, s_m], ori[e_m], s_m + 1))
    nSides = off_your[-1]
    m.nSides, m.nBCSides, m.nInnerSides = nSides, nBC, nInner
    m.nMPISides_MINE, m.nMPISides_YOUR = sum(nMINE_proc), sum(nYOUR_proc)
    m.nMPISides_MINE_Proc = np.array(nMINE_proc, dtype=np.int32)
    m.nMPISides_YOUR_Proc = np.array(nYOUR_proc, dtype=np.int32)
    m.offsetMPISides_MINE = np.array(off_mine, dtype=np.int32)
    m.offsetMPISides_YOUR = np.array(off_your, dtype=np.int32)
    # side ranges, mesh.f90:257-269 (1-based, inclusive)
    m.firstBCSide, m.lastBCSide = 1, nBC
    m.firstInnerSide, m.lastInnerSide = nBC + 1, nBC + nInner
    m.firstMPISide_MINE, m.lastMPISide_MINE = nBC + nInner + 1, nBC + nInner + m.nMPISides_MINE
    m.firstMPISide_YOUR, m.lastMPISide_YOUR = m.lastMPISide_MINE + 1, nSides

    # ---- ElemToSide / SideToElem ---------------------------------------------------------------------------------------
    E2S = np.zeros((nE, 6, 2), dtype=np.int32)
    E2S[:, :, 0] = side_id
    E2S[:, :, 1] = flip
    S2E = -np.ones((nSides, 5), dtype=np.int32)
    e_all, s_all = np.nonzero(np.ones((nE, 6), dtype=bool))
    mast = flip[e_all, s_all] == 0
    sid0 = side_id[e_all, s_all] - 1
    S2E[sid0[mast], 0] = e_all[mast] + 1
    S2E[sid0[mast], 2] = s_all[mast] + 1
    slv = ~mast
    S2E[sid0[slv], 1] = e_all[slv] + 1
    S2E[sid0[slv], 3] = s_all[slv] + 1
    S2E[sid0[slv], 4] = flip[e_all[slv], s_all[slv]]
    # master rows: flip entry is 0 when there is no local slave (BC / MPI_MINE) -- prepare_mesh.f90:725-735
    only_master = (S2E[:, 1] < 0)
    S2E[only_master, 4] = 0
    # MPI_YOUR sides: flip stays that of the (local) slave element; master entries stay -1
    m.ElemToSide = E2S
    m.SideToElem = S2E
    m.BC = np.zeros(max(nBC, 1), dtype=np.int32)
    if nBC:
        m.BC[side_id[is_bc] - 1] = bc_idx[is_bc]

    # ---- volume geometry ------------------------------------------------------------------------------------------------
    ops = _metric_operators(N, node_type, spec.ngeo)
    np1 = N + 1
    m.Elem_xGP = np.zeros((nE, np1, np1, np1, 3))
    m.Metrics_fTilde = np.zeros((nE, np1, np1, np1, 3))
    m.Metrics_gTilde = np.zeros((nE, np1, np1, np1, 3))
    m.Metrics_hTilde = np.zeros((nE, np1, np1, np1, 3))
    m.sJ = np.zeros((nE, np1, np1, np1))
    m.NormVec = np.zeros((nSides, np1, np1, 3))
    m.TangVec1 = np.zeros((nSides, np1, np1, 3))
    m.TangVec2 = np.zeros((nSides, np1, np1, 3))
    m.SurfElem = np.zeros((nSides, np1, np1))
    m.Face_xGP = np.zeros((nSides, np1, np1, 3))
    if with_ja_face:
        m.Ja_Face = np.zeros((nSides, np1, np1, 3, 3))
        m.Ja_slave = np.zeros((nSides, np1, np1, 3, 3))
    _, _, S2V2 = MP.build_mappings(N)
    for c0 in range(0, nE, chunk):
        c1 = min(nE, c0 + chunk)
        nodes = spec.node_coords(gids[c0:c1])
        geo = _element_metrics(nodes, ops)
        m.Elem_xGP[c0:c1] = geo["xGP"]
        m.Metrics_fTilde[c0:c1] = geo["Mf"]
        m.Metrics_gTilde[c0:c1] = geo["Mg"]
        m.Metrics_hTilde[c0:c1] = geo["Mh"]
        m.sJ[c0:c1] = geo["sJ"]
        _surface_metrics(m, geo, side_id[c0:c1], flip[c0:c1], np.ones((c1 - c0, 6), dtype=bool), ops, S2V2, with_ja_face)
    # MPI_YOUR sides: geometry of the remote master element (replaces the Geo exchange, metrics.f90:558-572)
    e_y, s_y = np.nonzero(is_mpi & ~is_mine)
    if len(e_y):
        for c0 in range(0, len(e_y), chunk):
            sl = slice(c0, min(len(e_y), c0 + chunk))
            rem_gid = nb_gid[e_y[sl], s_y[sl]]
            nodes = spec.node_coords(rem_gid)
            # shift periodic images so that the remote element touches the local one geometrically (only Face_xGP cares)
            geo = _element_metrics(nodes, ops)
            k = len(rem_gid)
            sid_rem = np.zeros((k, 6), dtype=np.int64)
            sel = np.zeros((k, 6), dtype=bool)
            rem_side = nb_side[e_y[sl], s_y[sl]] - 1
            sid_rem[np.arange(k), rem_side] = side_id[e_y[sl], s_y[sl]]
            sel[np.arange(k), rem_side] = True
            _surface_metrics(m, geo, sid_rem, np.zeros((k, 6), dtype=np.int64), sel, ops, S2V2, with_ja_face)
    m.nb_gid, m.nb_rank, m.face_gid = nb_gid, nb_rank, face_gid
    return m


def _flips(o_m, ls_m, o_s, ls_s):
    """oriented_flip over arrays (orientation of the master element, its local side, the same for the slave)."""
    o_m, ls_m, o_s, ls_s = np.broadcast_arrays(o_m, ls_m, o_s, ls_s)
    key = ((o_m * 7 + ls_m) * N_ORIENT + o_s) * 7 + ls_s
    out = np.zeros(key.shape, dtype=np.int64)
    for k in np.unique(key):
        ls_s_, r = k % 7, k // 7
        o_s_, r = r % N_ORIENT, r // N_ORIENT
        ls_m_, o_m_ = r % 7, r // 7
        out[key == k] = oriented_flip(int(o_m_), int(ls_m_), int(o_s_), int(ls_s_))
    return out


def _opp_arr(s):
    lut = np.zeros(7, dtype=np.int64)
    for a, b in _OPPOSITE.items():
        lut[a] = b
    return lut[s]


def _metric_operators(N, node_type, ngeo):
    """Vandermonde / derivative matrices prepared at the top of CalcMetrics (metrics.f90:233-262) and BuildCoords (:113-114)."""
    ngeo_ref = 3 * ngeo  # mesh.f90:336
    o = {"N": N, "NGeo": ngeo, "NGeoRef": ngeo_ref}
    o["Vdm_EQNGeo_CLNGeo"] = B.get_vandermonde(ngeo, B.NODE_VISU, ngeo, B.NODE_CL)
    o["Vdm_CLNGeo_CLN"] = B.get_vandermonde(ngeo, B.NODE_CL, N, B.NODE_CL)
    xcl_g, _, _ = B.get_nodes_and_weights(ngeo, B.NODE_CL)
    o["DCL_NGeo"] = B.polynomial_derivative_matrix(xcl_g)
    o["Vdm_CLNGeo_NGeoRef"] = B.get_vandermonde(ngeo, B.NODE_CL, ngeo_ref, node_type)
    o["Vdm_NGeoRef_N"] = B.get_vandermonde(ngeo_ref, node_type, N, node_type, modal=True)
    xcl_n, _, _ = B.get_nodes_and_weights(N, B.NODE_CL)
    o["DCL_N"] = B.polynomial_derivative_matrix(xcl_n)
    o["Vdm_CLN_N"] = B.get_vandermonde(N, B.NODE_CL, N, node_type)
    # BuildCoords: EQ NGeo -> CL N -> solution nodes N
    v1 = B.get_vandermonde(ngeo, B.NODE_VISU, N, B.NODE_CL)
    o["Vdm_EQNGeo_N"] = o["Vdm_CLN_N"] @ v1
    return o


def _deriv3(D, X):
    """d/dxi_dd of X[e,k,j,i,c] for dd = 1,2,3 with D[i,l] -> list of three arrays."""
    d1 = np.einsum("il,ekjlc->ekjic", D, X)
    d2 = np.einsum("jl,eklic->ekjic", D, X)
    d3 = np.einsum("kl,eljic->ekjic", D, X)
    return d1, d2, d3


def _element_metrics(nodes, o):
    """metrics.f90:296-542 for a batch of elements. nodes = NodeCoords[e,k,j,i,3] (equidistant NGeo)."""
    N, ngeo = o["N"], o["NGeo"]
    out = {}
    out["xGP"] = B.change_basis_3d(o["Vdm_EQNGeo_N"], nodes)
    XCL_NGeo = B.change_basis_3d(o["Vdm_EQNGeo_CLNGeo"], nodes)               # 1.a
    XCL_N = B.change_basis_3d(o["Vdm_CLNGeo_CLN"], XCL_NGeo)                  # 1.b
    dG = _deriv3(o["DCL_NGeo"], XCL_NGeo)                                     # 1.c  dG[dd][...,nn]
    dRef = [B.change_basis_3d(o["Vdm_CLNGeo_NGeoRef"], d) for d in dG]        # 1.d
    a, b, c = dRef  # a[...,nn] = dX_nn/dxi_1 ...
    detJ = (a[..., 0] * (b[..., 1] * c[..., 2] - c[..., 1] * b[..., 2])
            + b[..., 0] * (c[..., 1] * a[..., 2] - a[..., 1] * c[..., 2])
            + c[..., 0] * (a[..., 1] * b[..., 2] - b[..., 1] * a[..., 2]))     # 1.e
    detJ_N = B.change_basis_3d(o["Vdm_NGeoRef_N"], detJ[..., None])[..., 0]   # 1.f
    if np.any(detJ_N <= 0.0):
        raise RuntimeError("Negative Jacobian found on Gauss point.")
    out["sJ"] = 1.0 / detJ_N
    if N >= ngeo:                                                             # 2.a
        dN = [B.change_basis_3d(o["Vdm_CLNGeo_CLN"], d) for d in dG]
    else:
        dN = _deriv3(o["DCL_N"], XCL_N)
    # 2.b curl metrics: R(dd, nn) = 1/2 (X_{nn+2} dX_{nn+1}/dxi_dd - X_{nn+1} dX_{nn+2}/dxi_dd)
    D = o["DCL_N"]
    R = []  # R[dd][..., nn]
    for dd in range(3):
        r = np.empty_like(XCL_N)
        r[..., 0] = 0.5 * (XCL_N[..., 2] * dN[dd][..., 1] - XCL_N[..., 1] * dN[dd][..., 2])
        r[..., 1] = 0.5 * (XCL_N[..., 0] * dN[dd][..., 2] - XCL_N[..., 2] * dN[dd][..., 0])
        r[..., 2] = 0.5 * (XCL_N[..., 1] * dN[dd][..., 0] - XCL_N[..., 0] * dN[dd][..., 1])
        R.append(r)
    d_xi = lambda X: np.einsum("il,ekjlc->ekjic", D, X)
    d_eta = lambda X: np.einsum("jl,eklic->ekjic", D, X)
    d_zeta = lambda X: np.einsum("kl,eljic->ekjic", D, X)
    Ja1 = -d_eta(R[2]) + d_zeta(R[1])
    Ja2 = -d_zeta(R[0]) + d_xi(R[2])
    Ja3 = -d_xi(R[1]) + d_eta(R[0])
    out["JaCL"] = (Ja1, Ja2, Ja3)
    out["XCL_N"] = XCL_N
    out["Mf"] = B.change_basis_3d(o["Vdm_CLN_N"], Ja1)                        # 2.c
    out["Mg"] = B.change_basis_3d(o["Vdm_CLN_N"], Ja2)
    out["Mh"] = B.change_basis_3d(o["Vdm_CLN_N"], Ja3)
    return out


def _face_slice(A, loc_side, N):
    """Volume array A[e,k,j,i,c] -> face array [e, b, a, c] with (a,b) = (j,k) / (i,k) / (i,j) (prolongtoface.t90:232-263)."""
    if loc_side == MP.XI_MINUS:
        return A[:, :, :, 0, :]
    if loc_side == MP.XI_PLUS:
        return A[:, :, :, N, :]
    if loc_side == MP.ETA_MINUS:
        return A[:, :, 0, :, :]
    if loc_side == MP.ETA_PLUS:
        return A[:, :, N, :, :]
    if loc_side == MP.ZETA_MINUS:
        return A[:, 0, :, :, :]
    return A[:, N, :, :, :]


def _surface_metrics(m, geo, side_id, flip, select, o, S2V2, with_ja):
    """CalcSurfMetrics + SurfMetricsFromJa (metrics.f90:601-787) for the master (flip=0) sides of a batch of elements."""
    N = o["N"]
    V = o["Vdm_CLN_N"]
    for s in range(1, 7):
        use = select[:, s - 1] & (flip[:, s - 1] == 0)
        slave_use = select[:, s - 1] & (flip[:, s - 1] != 0)
        if with_ja and np.any(slave_use):
            es = np.nonzero(slave_use)[0]
            for f in range(1, 5):
                ef = es[flip[es, s - 1] == f]
                if len(ef) == 0:
                    continue
                pa, pb = S2V2[0, :, :, f, s - 1], S2V2[1, :, :, f, s - 1]
                for dd in range(3):
                    t2 = B.change_basis_2d(V, _face_slice(geo["JaCL"][dd][ef], s, N))
                    m.Ja_slave[side_id[ef, s - 1] - 1, :, :, :, dd] = np.transpose(t2[:, pb, pa, :], (0, 2, 1, 3))
        if not np.any(use):
            continue
        e = np.nonzero(use)[0]
        sid = side_id[e, s - 1] - 1
        pa, pb = S2V2[0, :, :, 0, s - 1], S2V2[1, :, :, 0, s - 1]     # [p,q] -> volume face indices (a,b)
        ja = []
        for dd in range(3):
            t2 = B.change_basis_2d(V, _face_slice(geo["JaCL"][dd][e], s, N))   # [e,b,a,c]
            ja.append(np.transpose(t2[:, pb, pa, :], (0, 2, 1, 3)))             # -> [e,q,p,c]
        xf = B.change_basis_2d(V, _face_slice(geo["XCL_N"][e], s, N))
        m.Face_xGP[sid] = np.transpose(xf[:, pb, pa, :], (0, 2, 1, 3))
        if with_ja:
            for dd in range(3):
                m.Ja_Face[sid, :, :, :, dd] = ja[dd]
        nd, td, sg = MP.NORMAL_DIRS[s - 1] - 1, MP.TANG_DIRS[s - 1] - 1, MP.NORMAL_SIGNS[s - 1]
        surf = np.sqrt(np.sum(ja[nd] ** 2, axis=-1))
        nv = sg * ja[nd] / surf[..., None]
        t1 = ja[td] - np.sum(ja[td] * nv, axis=-1)[..., None] * nv
        t1 = t1 / np.sqrt(np.sum(t1 ** 2, axis=-1))[..., None]
        t2v = np.cross(nv, t1)
        m.SurfElem[sid] = surf
        m.NormVec[sid] = nv
        m.TangVec1[sid] = t1
        m.TangVec2[sid] = t2v
