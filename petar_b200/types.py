"""numpy mirrors of the PeTar/FDPS layouts that cross the soft-force boundary.

Field-for-field equal to ``include/petar_b200_types.h`` (which cites the reference layouts:
``src/soft_ptcl.hpp:4-24, 271-278, 311-326`` and FDPS ``SPJQuadrupoleInAndOut``).
"""
import numpy as np

F64VEC = [("x", "<f8"), ("y", "<f8"), ("z", "<f8")]

EPISoft = np.dtype(
    [("id", "<i8"), ("pos", "<f8", (3,)), ("r_search", "<f8"), ("rank_org", "<i4"), ("type", "<i4")],
    align=True,
)
EPJSoft = np.dtype(
    [
        ("id", "<i8"),
        ("mass", "<f8"),
        ("pos", "<f8", (3,)),
        ("vel", "<f8", (3,)),
        ("r_in", "<f8"),
        ("r_out", "<f8"),
        ("r_search", "<f8"),
        ("r_scale_next", "<f8"),
        ("group_data", "<i8", (2,)),
        ("rank_org", "<i4"),
        ("adr_org", "<i4"),
    ],
    align=True,
)
SPJQuad = np.dtype([("mass", "<f8"), ("pos", "<f8", (3,)), ("quad", "<f8", (6,))], align=True)  # xx,yy,zz,xy,xz,yz
ForceSoft = np.dtype([("acc", "<f8", (3,)), ("pot", "<f8"), ("n_ngb", "<i8")], align=True)

# device-side list building (include/petar_b200.h: pb_tree_cell, pb_tree_group)
TreeCell = np.dtype(
    [("cm", "<f8", (3,)), ("len", "<f8"), ("in_lo", "<f8", (3,)), ("in_hi", "<f8", (3,)), ("out_lo", "<f8", (3,)), ("out_hi", "<f8", (3,)),
     ("child", "<i4", (8,)), ("first", "<i4"), ("n", "<i4"), ("leaf", "<i4"), ("n_let_sp", "<i4")],
    align=True,
)
TreeGroup = np.dtype(
    [("first", "<i4"), ("n", "<i4"), ("in_lo", "<f8", (3,)), ("in_hi", "<f8", (3,)), ("out_lo", "<f8", (3,)), ("out_hi", "<f8", (3,))],
    align=True,
)
assert TreeCell.itemsize == 176 and TreeGroup.itemsize == 104

# the fields of FPSoft / EPJSoft the changeover correction reads and updates (include/petar_b200_types.h: pb_PtclCorr)
PtclCorr = np.dtype(
    [("id", "<i8"), ("mass", "<f8"), ("pos", "<f8", (3,)), ("r_in", "<f8"), ("r_out", "<f8"),
     ("mass_backup", "<f8"), ("status", "<f8"), ("acc", "<f8", (3,)), ("pot_tot", "<f8"), ("pot_soft", "<f8")],
    align=True,
)
assert PtclCorr.itemsize == 112
LARGE_FLOAT = float(np.finfo(np.float32).max) * 0.0625   # FDPS PS::LARGE_FLOAT

assert EPISoft.itemsize == 48
assert EPJSoft.itemsize == 120
assert SPJQuad.itemsize == 80
assert ForceSoft.itemsize == 40
