"""Side <-> volume index mappings of the synthetic FLEXI host (stand-in for MOD_Mappings).

Reference: /root/reference/src/mesh/mappings.f90:188-450 (Flip_S2M, CGNS_VolToSide, CGNS_SideToVol, CGNS_SideToVol2,
VolToSide, SideToVol, SideToVol2) and buildMappings :40-180.  Local side numbering flexi.h:90-95:
ZETA_MINUS=1, ETA_MINUS=2, XI_PLUS=3, ETA_PLUS=4, XI_MINUS=5, ZETA_PLUS=6.  Constant tables mesh_vars.f90:105-111.

Arrays returned by build_mappings are indexed like the Fortran ones with 0-based flips and **0-based local sides**
(locSide-1) in numpy: S2V2[c, p, q, flip, locSide-1], S2V[c, l, p, q, flip, locSide-1]; the values are 0-based volume
indices exactly as in Fortran (whose volume indices are 0-based already).
"""
import numpy as np

ZETA_MINUS, ETA_MINUS, XI_PLUS, ETA_PLUS, XI_MINUS, ZETA_PLUS = 1, 2, 3, 4, 5, 6

# mesh_vars.f90:105-111 (index locSide-1)
NORMAL_DIRS = (3, 2, 1, 2, 1, 3)
TANG_DIRS = (1, 3, 2, 3, 2, 1)
NORMAL_SIGNS = (-1.0, -1.0, 1.0, 1.0, -1.0, 1.0)


def flip_s2m(n, p, q, flip):
    """mappings.f90:188-215 (3-D branch). Flip_M2S is the same function (:221-229)."""
    if flip == 0:
        return p, q
    if flip == 1:
        return q, p
    if flip == 2:
        return n - p, q
    if flip == 3:
        return n - q, n - p
    if flip == 4:
        return p, n - q
    raise ValueError(flip)


def cgns_vol_to_side(n, i, j, k, loc_side):
    """mappings.f90:237-268."""
    if loc_side == XI_MINUS:
        return k, j, i
    if loc_side == XI_PLUS:
        return j, k, n - i
    if loc_side == ETA_MINUS:
        return i, k, j
    if loc_side == ETA_PLUS:
        return n - i, k, n - j
    if loc_side == ZETA_MINUS:
        return j, i, k
    if loc_side == ZETA_PLUS:
        return i, j, n - k
    raise ValueError(loc_side)


def cgns_side_to_vol(n, l, p, q, loc_side):
    """mappings.f90:277-308."""
    if loc_side == XI_MINUS:
        return l, q, p
    if loc_side == XI_PLUS:
        return n - l, p, q
    if loc_side == ETA_MINUS:
        return p, l, q
    if loc_side == ETA_PLUS:
        return n - p, n - l, q
    if loc_side == ZETA_MINUS:
        return q, p, l
    if loc_side == ZETA_PLUS:
        return p, q, n - l
    raise ValueError(loc_side)


def cgns_side_to_vol2(n, p, q, loc_side):
    """mappings.f90:316-347."""
    if loc_side == XI_MINUS:
        return q, p
    if loc_side == XI_PLUS:
        return p, q
    if loc_side == ETA_MINUS:
        return p, q
    if loc_side == ETA_PLUS:
        return n - p, q
    if loc_side == ZETA_MINUS:
        return q, p
    if loc_side == ZETA_PLUS:
        return p, q
    raise ValueError(loc_side)


def vol_to_side(n, i, j, k, flip, loc_side):
    """mappings.f90:355-364."""
    p, q, l = cgns_vol_to_side(n, i, j, k, loc_side)
    p, q = flip_s2m(n, p, q, flip)
    return p, q, l


def side_to_vol(n, l, p, q, flip, loc_side):
    """mappings.f90:373-381."""
    p2, q2 = flip_s2m(n, p, q, flip)
    return cgns_side_to_vol(n, l, p2, q2, loc_side)


def side_to_vol2(n, p, q, flip, loc_side):
    """mappings.f90:389-397."""
    p2, q2 = flip_s2m(n, p, q, flip)
    return cgns_side_to_vol2(n, p2, q2, loc_side)


def build_mappings(n):
    """mappings.f90:40-180. Returns (V2S, S2V, S2V2) as int32 arrays."""
    np1 = n + 1
    V2S = np.zeros((3, np1, np1, np1, 5, 6), dtype=np.int32)
    S2V = np.zeros((3, np1, np1, np1, 5, 6), dtype=np.int32)
    S2V2 = np.zeros((2, np1, np1, 5, 6), dtype=np.int32)
    for f in range(5):
        for s in range(1, 7):
            for a in range(np1):
                for b in range(np1):
                    S2V2[:, a, b, f, s - 1] = side_to_vol2(n, a, b, f, s)
                    for c in range(np1):
                        V2S[:, a, b, c, f, s - 1] = vol_to_side(n, a, b, c, f, s)
                        S2V[:, a, b, c, f, s - 1] = side_to_vol(n, a, b, c, f, s)
    return V2S, S2V, S2V2
