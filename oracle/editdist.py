"""ORACLE (test infrastructure): ctypes access to oracle/editdist.c.

`align(a, b)` has the shape of `edlib.align` as used at SVIM_clustering.py:45
(default NW mode, distance only)."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")


def build():
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "editdist.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def _get():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        for fn in (lib.oracle_editdist_dp, lib.oracle_editdist_myers):
            fn.restype = ctypes.c_long
            fn.argtypes = [ctypes.c_char_p, ctypes.c_long, ctypes.c_char_p, ctypes.c_long]
        _lib = lib
    return _lib


def _b(s):
    return s.encode("latin-1") if isinstance(s, str) else bytes(s)


def edit_distance_dp(a, b) -> int:
    a, b = _b(a), _b(b)
    return int(_get().oracle_editdist_dp(a, len(a), b, len(b)))


def edit_distance(a, b) -> int:
    a, b = _b(a), _b(b)
    return int(_get().oracle_editdist_myers(a, len(a), b, len(b)))


def align(a, b, **_kw):
    return {"editDistance": edit_distance(a, b)}
