"""Reader for the reference's unit-test fixtures (gfortran unformatted sequential files).

Layout: [int32 nbytes][payload][int32 nbytes] per record, little endian, INTEGER*4 / REAL*8
(the reference is compiled with -fdefault-real-8, cmake/SetCompiler.cmake:79).
"""
import os
import struct

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_unittests")


def read_records(name):
    path = os.path.join(GOLDEN_DIR, name)
    recs = []
    with open(path, "rb") as f:
        data = f.read()
    pos = 0
    while pos < len(data):
        (n,) = struct.unpack_from("<i", data, pos)
        pos += 4
        recs.append(data[pos : pos + n])
        pos += n
        (n2,) = struct.unpack_from("<i", data, pos)
        assert n2 == n, "corrupt record marker in %s" % name
        pos += 4
    return recs


def f64(buf, shape=None):
    a = np.frombuffer(buf, dtype="<f8")
    if shape is not None:
        a = a.reshape(shape, order="F")
    return a


class UnittestElementData3D:
    """unitTests/unittest.f90:58-101 (ReadInReferenceElementData), written by :29-53.

    One record: nElems, SideToElem(5,6), firstMPISide_YOUR, lastMPISide_MINE, nSides, S2V2(2,0:9,0:9,0:4,1:6),
    S2V(3,0:9,0:9,0:9,0:4,1:6), L_Minus, L_Plus, L_HatPlus, L_HatMinus (0:9 each), sJ(0:9,0:9,0:9,0:1,1).
    """

    NRef = 9

    def __init__(self):
        (buf,) = read_records("UnittestElementData3D.bin")
        pos = 0

        def ints(n):
            nonlocal pos
            a = np.frombuffer(buf, dtype="<i4", count=n, offset=pos)
            pos += 4 * n
            return a

        def reals(n):
            nonlocal pos
            a = np.frombuffer(buf, dtype="<f8", count=n, offset=pos)
            pos += 8 * n
            return a

        n = self.NRef + 1
        self.nElems = int(ints(1)[0])
        self.SideToElem = ints(5 * 6).reshape((5, 6), order="F")
        self.firstMPISide_YOUR = int(ints(1)[0])
        self.lastMPISide_MINE = int(ints(1)[0])
        self.nSides = int(ints(1)[0])
        self.S2V2 = ints(2 * n * n * 5 * 6).reshape((2, n, n, 5, 6), order="F")
        self.S2V = ints(3 * n * n * n * 5 * 6).reshape((3, n, n, n, 5, 6), order="F")
        self.L_Minus = reals(n)
        self.L_Plus = reals(n)
        self.L_HatPlus = reals(n)
        self.L_HatMinus = reals(n)
        self.sJ = reals(n * n * n * 2).reshape((n, n, n, 2, 1), order="F")
        assert pos == len(buf), (pos, len(buf))
