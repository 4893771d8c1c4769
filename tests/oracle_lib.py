"""Binds the CPU oracle (oracle/_build/libip_oracle.so) through the package's generic ctypes binder."""
import importlib
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "libip_oracle.so")
_pkg = importlib.import_module("nceplibs-ip_b200")
_oracles = {}


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    return ORACLE_SO


def oracle(kind):
    if kind not in _oracles:
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        _oracles[kind] = _pkg.IpLib(ORACLE_SO, "oracle_", kind)
    return _oracles[kind]


def product(kind):
    return _pkg.load(kind)
