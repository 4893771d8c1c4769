"""Comparison helpers for the CUDA-vs-oracle parity tests.

Tolerances are BASELINE.json's: lo / ibo / iret / no and neighbour values bit-exact; interpolated values within
4 ulp in the _4 kind and 1e-12 relative in _d/_8.  A point whose source coordinate sits within a rounding error
of a stencil boundary ("projection tie point") could pick a different stencil if the two sides' libm differed;
both sides use correctly rounded binary64 elementary functions, so there are none: the bar is ZERO lo mismatches
and ZERO out-of-tolerance values.  Every comparison is logged in RECORDS, any offending point is enumerated
(index, both values, both lo) and tests/conftest.py dumps the log to gpurun_out/tie_points.json at session end
(committed copy: profiles/tie_points.json).
"""
import numpy as np

MAX_TIE_FRACTION = 0.0  # no tie points are forgiven; any that appear are enumerated in RECORDS and fail the case
RECORDS = []            # one entry per compare_fields call of the session


def ulp_diff32(a, b):
    ai = a.view(np.int32).astype(np.int64)
    bi = b.view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(ai & 0x7FFFFFFF), ai)
    bi = np.where(bi < 0, -(bi & 0x7FFFFFFF), bi)
    return np.abs(ai - bi)


def compare_fields(kind, got, ref, lo_g, lo_r, name="go", exact=False, label=""):
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
    rec = dict(case=str(label), kind=kind, **stats, points=[])
    if ties:
        where = np.argwhere(bad | ~both)
        for idx in where[:64]:
            t = tuple(int(i) for i in idx)
            rec["points"].append(dict(index=t, got=float(got[t]), ref=float(ref[t]), lo_got=int(lo_g[t]), lo_ref=int(lo_r[t])))
    RECORDS.append(rec)
    assert ties <= int(MAX_TIE_FRACTION * n), rec
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
            out.append(compare_fields(kind, a[nm], b[nm], a["lo"], b["lo"], nm, exact, label))
        for nm in ("rlat", "rlon") + (("crot", "srot") if vector else ()):
            # geometry comes from the same correctly rounded arithmetic on both sides: bit-identical
            nd = int((a[nm] != b[nm]).sum())
            RECORDS.append(dict(case=str(label), kind=kind, name=nm, n=int(a[nm].size), lo_mismatch=0, out_of_tol=nd, worst=0,
                                exact_frac=1.0 - nd / max(a[nm].size, 1),
                                points=[dict(index=(int(i),), got=float(a[nm][i]), ref=float(b[nm][i]), lo_got=1, lo_ref=1)
                                        for i in np.flatnonzero(a[nm] != b[nm])[:64]]))
            assert nd == 0, (label, nm, nd)
    return out
