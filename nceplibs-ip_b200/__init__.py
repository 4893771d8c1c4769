"""ipb200 — B200-native ipolates / ipolatev / gdswzd (NCEPLIBS-ip hot path) behind the reference's API.

The package name contains a hyphen, so import it with
    ipb = importlib.import_module("nceplibs-ip_b200")
"""
from .ip_mod import (BICUBIC_INTERP_ID, BILINEAR_INTERP_ID, BUDGET_INTERP_ID, NEIGHBOR_BUDGET_INTERP_ID,  # noqa: F401
                     NEIGHBOR_INTERP_ID, SPECTRAL_INTERP_ID, IpLib, PRODUCT_SO, load)
from .build import build_library  # noqa: F401
from .sharding import field_slice, weak_slice  # noqa: F401
