"""edlib stand-in for running the unmodified reference here: exact global
Levenshtein distance from oracle/editdist.c (see that file's header)."""
import os, sys
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from oracle.editdist import align  # noqa: E402,F401
