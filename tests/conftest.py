import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200 box)")


def pytest_sessionfinish(session, exitstatus):
    """Dump the parity log (tests/parity.py: RECORDS) so that tie points, if any ever appear, are enumerated."""
    import json
    try:
        import parity
    except Exception:
        return
    if not parity.RECORDS:
        return
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.environ.get("IPB_TIE_FILE", os.path.join(root, "gpurun_out", "tie_points.json"))
    try:
        os.makedirs(os.path.dirname(out), exist_ok=True)
        recs = parity.RECORDS
        doc = dict(cases=len(recs), points_compared=int(sum(r["n"] for r in recs)), lo_mismatches=int(sum(r["lo_mismatch"] for r in recs)),
                   out_of_tolerance=int(sum(r["out_of_tol"] for r in recs)),
                   tie_points=[dict(case=r["case"], kind=r["kind"], array=r["name"], **p) for r in recs for p in r["points"]],
                   per_case=[{k: r[k] for k in ("case", "kind", "name", "n", "lo_mismatch", "out_of_tol", "worst", "exact_frac")} for r in recs])
        with open(out, "w") as f:
            json.dump(doc, f, indent=1)
    except OSError:
        pass
