"""cu2rec_b200 -- B200-native (sm_100a) drop-in for the SGD matrix-factorisation training path
of nickgreenquist/cu2rec. The product is cu2rec_b200/lib/libcu2b.so (C ABI in include/cu2b.h)
plus the bin/mf CLI; this package is the ctypes harness used by the tests and bench."""
from . import _lib  # noqa: F401
from .api import *  # noqa: F401,F403
