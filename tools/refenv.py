"""Put the read-only reference (/root/reference/src) and the third-party shims on
sys.path.  Only for tools/ scripts run in the build container."""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/src"


def activate():
    if not os.path.isdir(REF):
        raise RuntimeError("reference tree not present (expected in the build container only)")
    for p in (os.path.join(HERE, "ref_shim"), REF, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
