"""Grids, options and fixtures shared by the tests (values from the reference's test drivers).

GRIB2 templates: /root/reference/tests/interp_mod_grib2.F90:77-116 (scalar), :421-458 (vector; grid 203
uses scan mode 72 = wind points); GRIB1 KGDS: /root/reference/tests/interp_mod_grib1.F90:67-93;
input grids: /root/reference/tests/input_data_mod_grib2.F90:24-31, input_data_mod_grib1.F90:28-34.
BASELINE.json grids: SURVEY.md §8.
"""
import lzma
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
MISSING32 = 2**31 - 1


def _m(kind):
    return 2**63 - 1 if kind == "8" else MISSING32


def g2_templates(kind="4", vector=False):
    m = _m(kind)
    t = {
        "3": (0, [6, 255, m, 255, m, 255, m, 360, 181, 0, m, 90000000, 0, 56, -90000000, 359000000, 1000000, 1000000, 0]),
        "8": (10, [6, 255, m, 255, m, 255, m, 116, 44, -48670000, 3104000, 56, 22500000, 61050000, 0, 64, 0, 318830000, 318830000]),
        "127": (40, [6, 255, m, 255, m, 255, m, 768, 384, 0, m, 89642000, 0, 48, -89642000, 359531000, 469000, 192, 0]),
        "203": (1, [6, 255, m, 255, m, 255, m, 669, 1165, 0, m, -45000000, 300000000, 56, 45000000, 60000000, 179641, 77320,
                    72 if vector else 68, -36000000, 254000000, 0]),
        "205": (1, [6, 255, m, 255, m, 255, m, 954, 835, 0, m, -45036000, 299961000, 56, 45036000, 60039000, 126000, 108000, 64,
                    -36000000, 254000000, 0]),
        "212": (20, [6, 255, m, 255, m, 255, m, 512, 512, -20826000, 145000000, 56, 60000000, 280000000, 47625000, 47625000, 0, 0]),
        "218": (30, [6, 255, m, 255, m, 255, m, 614, 428, 12190000, 226541000, 56, 25000000, 265000000, 12191000, 12191000, 0, 64,
                     25000000, 25000000, -90000000, 0]),
    }
    return {k: ("grib2", v[0], v[1]) for k, v in t.items()}


def g2_input(kind="4", vector=False):
    m = _m(kind)
    if vector:
        return ("grib2", 0, [6, 255, m, 255, m, 255, m, 360, 181, 0, m, 90000000, 0, 48, -90000000, 359000000, 1000000, 1000000, 0])
    return ("grib2", 0, [6, 255, m, 255, m, 255, m, 360, 180, 0, m, -89500000, -180000000, 48, 89500000, 179000000, 1000000, 1000000, 64])


def _k(head):
    return head + [0] * (200 - len(head))


def g1_templates():
    t = {
        "3": [0, 360, 181, 90000, 0, 128, -90000, -1000, 1000, 1000, 0, 0, 0, 0, 0, 0, 0, 0, 0, 255],
        "8": [1, 116, 44, -48670, 3104, 128, 61050, 0, 22500, 0, 64, 318830, 318830, 0, 0, 0, 0, 0, 0, 255],
        "127": [4, 768, 384, 89642, 0, 128, -89642, -469, 469, 192, 0, 0, 255, 0, 0, 0, 0, 0, 0, 255],
        "203": [203, 669, 1165, -7450, -144140, 136, 54000, -106000, 90, 77, 64, 0, 0, 0, 0, 0, 0, 0, 0, 255],
        "205": [205, 954, 835, -7491, -144134, 136, 54000, -106000, 126, 90, 64, 44540, 14800, 0, 0, 0, 0, 0, 0, 255],
        "212": [5, 512, 512, -20826, 145000, 8, -80000, 47625, 47625, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 255],
        "218": [3, 614, 428, 12190, -133459, 8, -95000, 12191, 12191, 0, 64, 25000, 25000, 0, 0, 0, 0, 0, 0, 255],
    }
    return {k: ("grib1", _k(v)) for k, v in t.items()}


def g1_input(vector=False):
    if vector:
        head = [0, 360, 181, 90000, 0, 128, -90000, -1000, 1000, 1000, 0, 0] + [-1] * 6 + [0, 255]
    else:
        head = [0, 360, 180, -89500, -180000, 128, 89500, 179000, 1000, 1000, 64, 0] + [-1] * 6 + [0, 255]
    return ("grib1", head + [-1] * 180)


def station_grid(grib="grib2"):
    if grib == "grib2":
        return ("grib2", -1, [0] * 200)
    return ("grib1", _k([-1]))


def npts_of(desc):
    if desc[0] == "grib2":
        return desc[2][7] * desc[2][8]
    return desc[1][1] * desc[1][2]


def ipopt_for(ip):
    """The tests' option choice (interp_mod_grib2.F90:188-224): 0 for ip 0/1/2, all -1 for ip 3/6."""
    return [-1] * 20 if ip in (3, 6) else [0] * 20


# golden (grid, ip) pairs present in the reference checkout
SCALAR_CASES = [("3", 0), ("8", 1), ("127", 2), ("203", 3), ("212", 6), ("218", 0)]
VECTOR_CASES = [("3", 0), ("8", 1), ("127", 2), ("212", 6), ("218", 0)]

# station-point known answers: interp_mod_grib2.F90:268-282 (scalar), :652-672 (vector)
STATION_SCALAR = {0: [77.5, 71.0, 69.5, 65.0], 1: [77.6875, 71.0625, 69.8750, 69.9375], 2: [78.0, 71.0, 69.0, 61.0],
                  3: [77.32, 71.0599, 69.32, 61.82], 6: [77.59999, 71.00000, 69.40000, 64.20000]}
STATION_SCALAR_LATLON = ([45.0, 35.0, 40.0, 35.0], [-100.0, -100.0, -90.0, -120.0])
STATION_VECTOR = {
    0: ([3.55, 6.47, 8.25, 0.68], [26.38, 17.14, 8.32, -7.28]),
    1: ([3.55, 6.47, 8.25, 0.68], [26.38, 17.14, 8.32, -7.28]),
    2: ([3.55, 6.47, 8.25, 0.68], [26.38, 17.14, 8.32, -7.28]),
    3: ([3.18230, 6.74294, 7.89213, 0.70410], [26.20624, 17.97046, 8.75334, -7.37988]),
    6: ([3.54999, 6.46999, 8.25, 0.68], [26.37999, 17.13999, 8.31999, -7.28]),
}
STATION_VECTOR_LATLON = ([45.0, 35.0, 40.0, 90.0], [-100.0, -100.0, -90.0, 10.0])

_inputs = None
_strided = None


def inputs():
    global _inputs
    if _inputs is None:
        z = np.load(os.path.join(GOLD, "inputs.npz"))
        uv = z["uv"].reshape(2, -1)
        _inputs = {"snoalb": z["snoalb"].copy(), "u": uv[0].copy(), "v": uv[1].copy()}
    return _inputs


def golden(grib, kind, sv, grid, ip):
    """-> (values[records, npts_sel], stride).  kind: '4' or '8' (the 8-byte files pin _d and _8)."""
    global _strided
    stem = f"{grib}_{kind}_{sv}_grid{grid}.opt{ip}"
    full = os.path.join(GOLD, "full", stem + ".bin.xz")
    nrec = 2 if sv == "v" else 1
    if os.path.exists(full):
        with lzma.open(full, "rb") as f:
            a = np.frombuffer(f.read(), dtype="<f4")
        return a.reshape(nrec, -1), 1
    if _strided is None:
        _strided = np.load(os.path.join(GOLD, "strided.npz"))
    return _strided[stem], int(_strided["stride"])


# ---- BASELINE.json grids ----------------------------------------------------
def baseline_grids(kind="4"):
    m = _m(kind)
    return {
        "S": ("grib2", 0, [6, 255, m, 255, m, 255, m, 1440, 721, 0, m, 90000000, 0, 48, -90000000, 359750000, 250000, 250000, 0]),
        "L": ("grib2", 30, [6, 255, m, 255, m, 255, m, 1799, 1059, 21138123, 237280472, 8, 38500000, 262500000, 3000000, 3000000,
                           0, 64, 38500000, 38500000, -90000000, 0]),
        "P": ("grib2", 20, [6, 255, m, 255, m, 255, m, 512, 512, -20826000, 145000000, 56, 60000000, 280000000, 47625000, 47625000, 0, 0]),
        "E": ("grib2", 1, [6, 255, m, 255, m, 255, m, 669, 1165, 0, m, -45000000, 300000000, 56, 45000000, 60000000, 179641, 77320,
                          68, -36000000, 254000000, 0]),
    }
