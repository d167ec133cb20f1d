"""A/B of the batch-1 / batch-4096 search against an older build of the library (bring-up aid):
    python tools/ab_old_lib.py <path to liblatentknn_*.so | current> [--rows N] [--batch B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from latent_rag_b200 import _native  # noqa: E402

lib = sys.argv[1]
if lib != "current":
    import ctypes

    _native.LIB_PATH = os.path.abspath(lib)
    probe = ctypes.CDLL(_native.LIB_PATH)
    _native.SYMBOLS = [s for s in _native.SYMBOLS if hasattr(probe, s[0])]
sys.argv = [sys.argv[0]] + sys.argv[2:]
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "prof_case.py")).read())
