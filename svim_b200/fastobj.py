"""Loader / in-tree build of _svimfastobj (csrc_host/fastobj.c): the C loop that turns svim_sig / svim_cluster records into the
Python objects of svim_b200/SVSignature.py.  Host-side marshalling only — nothing is computed here."""
import importlib.util
import os
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc_host", "fastobj.c")
SO = os.path.join(HERE, "_svimfastobj.so")
_mod = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        from .build import run_atomic
        run_atomic(["gcc", "-O2", "-fPIC", "-shared", "-Wall", "-I" + sysconfig.get_paths()["include"], "-o", "@OUT@", SRC, "-lm"], SO)
    return SO


def module():
    global _mod
    if _mod is None:
        spec = importlib.util.spec_from_file_location("_svimfastobj", build())
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod
