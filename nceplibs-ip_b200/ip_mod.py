"""Python mirror of the reference's `ip_mod` facade (src/ip_mod.F90:7-20) over the C ABI.

`IpLib` binds a shared library that exports the reference's bind(c) argument lists
(src/ipolates.F90:158,293,587,808; src/ipolatev.F90:382,565,680,832; src/gdswzd_c.F90:173,259)
with the build-kind suffix _4 / _d / _8 on every symbol.  `load(kind)` returns the binder for the
product library (libipb200.so, CUDA); there is no CPU path — a missing library or device is an
error.  The tests bind the CPU oracle through the same class with prefix "oracle_".

Arrays follow the Fortran layout seen from C: field k of gi starts at k*mi, so a numpy array of
shape [km, mi] in C order is GI(MI,KM).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

BILINEAR_INTERP_ID, BICUBIC_INTERP_ID, NEIGHBOR_INTERP_ID, BUDGET_INTERP_ID = 0, 1, 2, 3
SPECTRAL_INTERP_ID, NEIGHBOR_BUDGET_INTERP_ID = 4, 6  # src/ip_interpolators_mod.F90:17-27

KINDS = {"4": (np.float32, np.int32), "d": (np.float64, np.int32), "8": (np.float64, np.int64)}


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class IpLib:
    def __init__(self, path, prefix, kind):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        self.kind = kind
        self.R, self.I = KINDS[kind]
        self.cI = C.c_int32 if self.I == np.int32 else C.c_int64
        self.cR = C.c_float if self.R == np.float32 else C.c_double

    def fn(self, name):
        f = getattr(self.lib, f"{self.prefix}{name}_{self.kind}")
        f.restype = None
        return f

    def _ia(self, v, n=None):
        a = np.ascontiguousarray(np.asarray(v, dtype=self.I).ravel())
        if n is not None:
            assert a.size == n, (a.size, n)
        return a

    def _grid_args(self, desc):
        if desc[0] == "grib2":
            t = self._ia(desc[2])
            return [_p(self._ia([desc[1]])), _p(t), _p(self._ia([t.size]))], [t]
        k = self._ia(desc[1], 200)
        return [_p(k)], [k]

    def pinned(self, shape, dtype):
        """numpy array over page-locked host memory from the library's allocator (ipgpu_host_alloc); kept for the process."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.lib.ipgpu_host_alloc.restype = C.c_void_p
        self.lib.ipgpu_host_alloc.argtypes = [C.c_longlong]
        p = self.lib.ipgpu_host_alloc(C.c_longlong(max(n, 1)))
        if not p:
            raise MemoryError("ipgpu_host_alloc failed")
        return np.frombuffer((C.c_char * max(n, 1)).from_address(p), dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def ipolates(self, ip, ipopt, gin, gout, mi, mo, km, ibi, li, gi, no=0, rlat=None, rlon=None, out_init=None, single=False, pinned=False):
        """Returns dict(no, rlat, rlon, ibo, lo, go, iret).  gi: [km, mi] C-order (== Fortran GI(MI,KM)).
        single=True calls the *_single_field entry (scalar ibi/ibo, rank-1 fields; src/ipolates.F90:158,808).
        pinned=True puts the field arrays into page-locked memory (the library then copies them without staging)."""
        assert gin[0] == gout[0]
        assert not single or km == 1
        R = self.R
        gi = np.ascontiguousarray(gi, dtype=R).reshape(km, mi)
        li = np.ascontiguousarray(li, dtype=np.uint8).reshape(km, mi)
        rlat = np.zeros(mo, R) if rlat is None else np.ascontiguousarray(rlat, dtype=R).copy()
        rlon = np.zeros(mo, R) if rlon is None else np.ascontiguousarray(rlon, dtype=R).copy()
        fillv = -7777.0 if out_init is None else out_init
        go = np.full((km, mo), fillv, R)
        lo = np.full((km, mo), 7, np.uint8)
        if pinned:
            gi_, li_, go_, lo_ = self.pinned(gi.shape, R), self.pinned(li.shape, np.uint8), self.pinned(go.shape, R), self.pinned(lo.shape, np.uint8)
            gi_[:], li_[:], go_[:], lo_[:] = gi, li, go, lo
            gi, li, go, lo = gi_, li_, go_, lo_
        ibo = self._ia([-7] * km)
        no_a, iret = self._ia([no]), self._ia([-7])
        a_in, keep1 = self._grid_args(gin)
        a_out, keep2 = self._grid_args(gout)
        sc = [self._ia([ip]), self._ia(ipopt, 20), self._ia([mi]), self._ia([mo]), self._ia([km]), self._ia(ibi, km)]
        name = ("ipolates_grib2" if gin[0] == "grib2" else "ipolates_grib1") + ("_single_field" if single else "")
        self.fn(name)(_p(sc[0]), _p(sc[1]), *a_in, *a_out, _p(sc[2]), _p(sc[3]), _p(sc[4]), _p(sc[5]), _p(li), _p(gi),
                      _p(no_a), _p(rlat), _p(rlon), _p(ibo), _p(lo), _p(go), _p(iret))
        return dict(no=int(no_a[0]), rlat=rlat, rlon=rlon, ibo=ibo, lo=lo, go=go, iret=int(iret[0]))

    def ipolatev(self, ip, ipopt, gin, gout, mi, mo, km, ibi, li, ui, vi, no=0, rlat=None, rlon=None, crot=None, srot=None,
                 single=False):
        """single=True calls the *_single_field entry (src/ipolatev.F90:680,832)."""
        assert gin[0] == gout[0]
        assert not single or km == 1
        R = self.R
        ui = np.ascontiguousarray(ui, dtype=R).reshape(km, mi)
        vi = np.ascontiguousarray(vi, dtype=R).reshape(km, mi)
        li = np.ascontiguousarray(li, dtype=np.uint8).reshape(km, mi)
        mk = lambda a: np.zeros(mo, R) if a is None else np.ascontiguousarray(a, dtype=R).copy()
        rlat, rlon, crot, srot = mk(rlat), mk(rlon), mk(crot), mk(srot)
        uo = np.full((km, mo), -7777.0, R)
        vo = np.full((km, mo), -7777.0, R)
        lo = np.full((km, mo), 7, np.uint8)
        ibo = self._ia([-7] * km)
        no_a, iret = self._ia([no]), self._ia([-7])
        a_in, keep1 = self._grid_args(gin)
        a_out, keep2 = self._grid_args(gout)
        sc = [self._ia([ip]), self._ia(ipopt, 20), self._ia([mi]), self._ia([mo]), self._ia([km]), self._ia(ibi, km)]
        name = ("ipolatev_grib2" if gin[0] == "grib2" else "ipolatev_grib1") + ("_single_field" if single else "")
        self.fn(name)(_p(sc[0]), _p(sc[1]), *a_in, *a_out, _p(sc[2]), _p(sc[3]), _p(sc[4]), _p(sc[5]), _p(li), _p(ui), _p(vi),
                      _p(no_a), _p(rlat), _p(rlon), _p(crot), _p(srot), _p(ibo), _p(lo), _p(uo), _p(vo), _p(iret))
        return dict(no=int(no_a[0]), rlat=rlat, rlon=rlon, crot=crot, srot=srot, ibo=ibo, lo=lo, uo=uo, vo=vo, iret=int(iret[0]))

    def gdswzd(self, desc, iopt, npts, fill=-9999.0, xpts=None, ypts=None, rlon=None, rlat=None):
        """The C entry (scalars by value).  Returns dict of all arrays + nret."""
        R = self.R
        mk = lambda a: np.zeros(npts, R) if a is None else np.ascontiguousarray(a, dtype=R).copy()
        x, y, lon, lat = mk(xpts), mk(ypts), mk(rlon), mk(rlat)
        ex = {k: np.full(npts, 12345.0, R) for k in ("crot", "srot", "xlon", "xlat", "ylon", "ylat", "area")}
        nret = self._ia([-7])
        tail = [_p(x), _p(y), _p(lon), _p(lat), _p(nret)] + [_p(ex[k]) for k in ("crot", "srot", "xlon", "xlat", "ylon", "ylat", "area")]
        if desc[0] == "grib2":
            t = self._ia(desc[2])
            self.fn("gdswzd")(self.cI(desc[1]), _p(t), self.cI(t.size), self.cI(iopt), self.cI(npts), self.cR(fill), *tail)
        else:
            k = self._ia(desc[1], 200)
            self.fn("gdswzd_grib1")(_p(k), self.cI(iopt), self.cI(npts), self.cR(fill), *tail)
        return dict(xpts=x, ypts=y, rlon=lon, rlat=lat, nret=int(nret[0]), **ex)


    # ---- grid expanders (src/ipxwafs.F90:81, ipxwafs2.F90:92, ipxwafs3.F90:95, ipxetas.F90:90): every argument by reference ----
    def ipxwafs(self, variant, idir, km, opt_pts, tmpl_thin, data_thin, tmpl_full, data_full, ib_thin=None, bitmap_thin=None, ib_full=None,
                bitmap_full=None, nthin=3447, nfull=5329):
        """variant 1 / 2 / 3 = ipxwafs / ipxwafs2 / ipxwafs3.  Arrays are copied; returns dict of every in/out argument."""
        R = self.R
        opt = self._ia(opt_pts)
        tt, tf = self._ia(tmpl_thin), self._ia(tmpl_full)
        dt = np.ascontiguousarray(data_thin, dtype=R).reshape(km, nthin).copy()
        df = np.ascontiguousarray(data_full, dtype=R).reshape(km, nfull).copy()
        sc = [self._ia([idir]), self._ia([nthin]), self._ia([nfull]), self._ia([km]), self._ia([opt.size]), self._ia([tt.size])]
        iret = self._ia([-7])
        if variant == 1:
            self.fn("ipxwafs")(_p(sc[0]), _p(sc[1]), _p(sc[2]), _p(sc[3]), _p(sc[4]), _p(opt), _p(sc[5]), _p(tt), _p(dt), _p(tf), _p(df), _p(iret))
            return dict(opt_pts=opt, tmpl_thin=tt, tmpl_full=tf, data_thin=dt, data_full=df, iret=int(iret[0]))
        ibt = self._ia([0] * km if ib_thin is None else ib_thin, km)
        ibf = self._ia([0] * km if ib_full is None else ib_full, km)
        bt = np.ones((km, nthin), np.uint8) if bitmap_thin is None else np.ascontiguousarray(bitmap_thin, dtype=np.uint8).reshape(km, nthin).copy()
        bf = np.ones((km, nfull), np.uint8) if bitmap_full is None else np.ascontiguousarray(bitmap_full, dtype=np.uint8).reshape(km, nfull).copy()
        self.fn("ipxwafs2" if variant == 2 else "ipxwafs3")(_p(sc[0]), _p(sc[1]), _p(sc[2]), _p(sc[3]), _p(sc[4]), _p(opt), _p(sc[5]), _p(tt), _p(dt),
                                                            _p(ibt), _p(bt), _p(tf), _p(df), _p(ibf), _p(bf), _p(iret))
        return dict(opt_pts=opt, tmpl_thin=tt, tmpl_full=tf, data_thin=dt, data_full=df, ib_thin=ibt, ib_full=ibf, bitmap_thin=bt, bitmap_full=bf,
                    iret=int(iret[0]))

    def ipxetas(self, idir, igdtnumi, tmpl_in, bitmap_in, data_in, npts_out):
        R = self.R
        ti = self._ia(tmpl_in)
        di = np.ascontiguousarray(data_in, dtype=R).ravel()
        bi = np.ascontiguousarray(bitmap_in, dtype=np.uint8).ravel()
        to = self._ia([0] * ti.size)
        do = np.full(npts_out, -7777.0, R)
        bo = np.full(npts_out, 7, np.uint8)
        sc = [self._ia([idir]), self._ia([igdtnumi]), self._ia([ti.size]), self._ia([di.size]), self._ia([-7]), self._ia([npts_out]), self._ia([-7])]
        self.fn("ipxetas")(_p(sc[0]), _p(sc[1]), _p(sc[2]), _p(ti), _p(sc[3]), _p(bi), _p(di), _p(sc[4]), _p(to), _p(sc[5]), _p(bo), _p(do), _p(sc[6]))
        return dict(igdtnumo=int(sc[4][0]), tmpl_out=to, bitmap_out=bo, data_out=do, iret=int(sc[6][0]))


PRODUCT_SO = os.path.join(HERE, "libipb200.so")

_libs = {}


def load(kind="4"):
    """Binder for the CUDA library in build kind '4', 'd' or '8'."""
    if kind not in _libs:
        if not os.path.exists(PRODUCT_SO):
            raise RuntimeError(f"{PRODUCT_SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        _libs[kind] = IpLib(PRODUCT_SO, "", kind)
    return _libs[kind]
