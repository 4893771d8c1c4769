"""Round-2 parity cases shared by the CPU (oracle) and GPU (CUDA vs oracle) tests.

* every projection as the INPUT grid (its inverse is what ipolates uses on the source side:
  ip_lambert_conf_grid_mod.F90:343-350, ip_polar_stereo_grid_mod.F90:407-443, ip_mercator_grid_mod.F90:290-294,
  ip_gaussian_grid_mod.F90:330-344, ip_rot_equid_cylind_grid_mod.F90:397+, ip_rot_equid_cylind_egrid_mod.F90:407-455),
  including the elliptical polar-stereographic branch (earth shape 5, ip_polar_stereo_grid_mod.F90:366-393,433-443);
* gdswzd with iopt = 0 / +1 / -1 and all seven optional outputs on every projection, plus the forward/inverse round
  trip of tests/reg_tests/gdswzd/sorc/gdswzd_driver.f90:296-330;
* every GRIB2 earth-shape code (earth_radius_mod.F90:40-95) through a projection that uses the radius.
"""
import numpy as np

import ipcases as ic


def conus_latlon(kind="4", di=500000):
    """Regional 0.5-degree lat-lon grid over North America, S->N scan (template 3.0)."""
    m = ic._m(kind)
    im = 60000000 // di + 1
    jm = 35000000 // di + 1
    return ("grib2", 0, [6, 255, m, 255, m, 255, m, im, jm, 0, m, 20000000, 235000000, 48, 55000000, 295000000, di, di, 64])


def conus_latlon_g1():
    """The same region as a GRIB1 KGDS (millidegrees)."""
    head = [0, 121, 71, 20000, -125000, 128, 55000, -65000, 500, 500, 64, 0] + [0] * 7 + [255]
    return ("grib1", ic._k(head))


def polar_ellipsoid(kind="4", shape=5):
    """Grid 212's polar stereographic template on an ellipsoidal earth (shape 5 = WGS84): the elliptical branch."""
    d = ic.g2_templates(kind)["212"]
    t = list(d[2])
    t[0] = shape
    return ("grib2", d[1], t)


def source_grids(kind="4", vector=False):
    """name -> descriptor of every projection usable as an input grid."""
    g = dict(ic.g2_templates(kind, vector))
    g["212e"] = polar_ellipsoid(kind)
    return g


def shaped(desc, code, kind="4"):
    """desc with GRIB2 shape-of-earth code `code` (octets 15-30 of section 3: template words 1-7)."""
    t = list(desc[2])
    t[0] = code
    if code == 1:
        t[1], t[2] = 2, 637000000          # scale factor 2, scaled radius -> 6 370 000 m
    elif code in (3, 7):
        t[3], t[4], t[5], t[6] = 3, 6378160, 3, 6356775   # major / minor axis in km (3) or m (7): scale 3 -> value / 1000
        if code == 7:
            t[3], t[4], t[5], t[6] = 0, 6378160, 0, 6356775
    return (desc[0], desc[1], t)


def fields_for(mi, km, seed, kind, im=None):
    """km smooth-plus-noise fields of mi points and blob bitmaps (~70 % true)."""
    rng = np.random.default_rng(seed)
    R = np.float32 if kind == "4" else np.float64
    n = np.arange(mi)
    im = im or max(int(np.sqrt(mi)), 1)
    jj, ii = np.divmod(n, im)
    g = np.stack([50.0 + 30.0 * np.sin(ii * 0.05 * (1 + k % 3)) * np.cos(jj * 0.04) + rng.standard_normal(mi) + k for k in range(km)]).astype(R)
    v = np.stack([10.0 * np.cos(ii * 0.03) + rng.standard_normal(mi) * 3 - k for k in range(km)]).astype(R)
    li = np.zeros((km, mi), np.uint8)
    for k in range(km):
        h = ((ii >> 3) * 73856093 ^ (jj >> 3) * 19349663 ^ (k * 83492791)) % 10
        li[k] = (h < 7).astype(np.uint8)
    return g, v, li


def dims_of(desc):
    if desc[0] == "grib2":
        return desc[2][7], desc[2][8]
    return desc[1][1], desc[1][2]


# ---- BASELINE.json configs 2-5 on the real grids, small field counts ----------------------
def synth_s(k, kind, second=False, im=1440, jm=721):
    """The synthetic source field of SURVEY.md section 8d on the 0.25-degree grid."""
    i = np.arange(im, dtype=np.float64)[None, :]
    j = np.arange(jm, dtype=np.float64)[:, None]
    if second:
        f = 10.0 + 8.0 * np.cos(2 * np.pi * i / im * (1 + k % 5)) * np.sin(np.pi * (j - 360) / jm) - k * 1e-3
    else:
        f = 50.0 + 40.0 * np.sin(2 * np.pi * i / im * (1 + k % 7)) * np.cos(np.pi * (j - 360) / jm) + k * 1e-3
    return f.reshape(-1).astype(np.float32 if kind == "4" else np.float64)


def synth_mask(k, im=1440, jm=721):
    i = np.arange(im)[None, :]
    j = np.arange(jm)[:, None]
    h = ((i >> 3) * 73856093 ^ (j >> 3) * 19349663 ^ (k * 83492791)) % 10
    return (h < 7).astype(np.uint8).reshape(-1)
