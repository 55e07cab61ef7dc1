"""A/B timing of two library builds (dev aid): OSRL_B200_LIBNAME selects the .so; missing newer symbols are skipped."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from osrl_b200 import _lib
lib = ctypes.CDLL(_lib.LIB_PATH)
_lib.SYMBOLS = [s for s in _lib.SYMBOLS if hasattr(lib, s[0])]
sys.argv = [sys.argv[0]] + sys.argv[1:]
exec(open(os.path.join(os.path.dirname(__file__), "quick_bench.py")).read())
