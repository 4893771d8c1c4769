"""Cases for the grid expanders (ipxwafs*, ipxetas), shared by the CPU (oracle) and GPU (CUDA vs oracle) tests.
Values from /root/reference/tests/test_ipxwafs.F90:24-68."""
import numpy as np

NPWAFS = [73, 73, 73, 73, 73, 73, 73, 73, 72, 72, 72, 71, 71, 71, 70, 70, 69, 69, 68, 67, 67, 66, 65, 65, 64, 63, 62, 61, 60, 60, 59, 58, 57, 56, 55,
          54, 52, 51, 50, 49, 48, 47, 45, 44, 43, 42, 40, 39, 38, 36, 35, 33, 32, 30, 29, 28, 26, 25, 23, 22, 20, 19, 17, 16, 14, 12, 11, 9, 8, 6, 5, 3, 2]
WAFS_REF12 = [0.68969796602520717, 0.68991151726138666, 0.69012506849756627, 0.69033861973374588, 0.69055217096992549, 0.69076572220610510,
              0.69097927344228471, 0.69119282467846443, 0.69140637591464404, 0.69161992715082354]
WAFS_REF3 = [0.68958514650420655, 0.68987525384392223, 0.69016536118363792, 0.69045546852335360, 0.69045546852335360, 0.69074557586306928,
             0.69103568320278508, 0.69132579054250076, 0.69132579054250076, 0.69161589788221645]
NTHIN, NFULL = 3447, 5329


def wafs_thin_field(R):
    return (np.arange(1, NTHIN + 1, dtype=R) / R(NTHIN)).astype(R)


def wafs_round_trip(lib, variant, km=1, fields=None, ib_thin=None, bitmap_thin=None):
    """expand thin -> full, then contract full -> thin, as the reference's test does."""
    tt = [0] * 19
    tt[8], tt[17], tt[18] = 73, 1250000, 64
    thin = np.stack([wafs_thin_field(lib.R)] * km) if fields is None else fields
    r = lib.ipxwafs(variant, 1, km, NPWAFS, tt, thin, [0] * 19, np.zeros((km, NFULL)), ib_thin=ib_thin, bitmap_thin=bitmap_thin)
    tf = list(r["tmpl_full"])
    tf[7], tf[8], tf[16] = 73, 73, 1250000
    back = lib.ipxwafs(variant, -1, km, NPWAFS, tt, np.zeros((km, NTHIN)), tf, r["data_full"],
                       ib_full=r.get("ib_full"), bitmap_full=r.get("bitmap_full"))
    return r, back


def egrid_case(scan, idir, im=89, jm=71):
    """A small rotated lat-lon E grid (template 3.1) in the given scan mode, and the point counts of the ipxetas call."""
    full = scan == 64
    imx = 2 * im - 1 if full else im
    dlon = 179641 // 2 if full else 179641
    t = [6, 255, -1, 255, -1, 255, -1, imx, jm, 0, -1, -3000000, 357000000, 56, 3000000, 3000000, dlon, 77320, scan, -36000000, 254000000, 0]
    npi = imx * jm
    npo = (2 * im - 1) * jm if not full else im * jm
    return t, npi, npo
