"""Binds the CPU oracle (oracle/_build/libip_oracle.so) through the package's generic ctypes binder."""
import importlib
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "libip_oracle.so")              # correctly rounded libm (parity contract)
ORACLE_GLIBC_SO = os.path.join(ROOT, "oracle", "_build", "libip_oracle_glibc.so")  # plain glibc libm (what gfortran would link)
_pkg = importlib.import_module("nceplibs-ip_b200")
_oracles = {}


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    return ORACLE_SO


def oracle(kind, flavour="cr"):
    """flavour 'cr': every libm call correctly rounded (the contract the CUDA library is held to);
    'glibc': the host's binary64 libm, as a gfortran build of the reference would use."""
    so = ORACLE_SO if flavour == "cr" else ORACLE_GLIBC_SO
    if (kind, flavour) not in _oracles:
        if not os.path.exists(so):
            build_oracle()
        _oracles[(kind, flavour)] = _pkg.IpLib(so, "oracle_", kind)
    return _oracles[(kind, flavour)]


def product(kind):
    return _pkg.load(kind)
