"""Alias of oracle/ref_shim.py (the shim moved there in round 2 so that bench.py and the tests share it)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.ref_shim import available, install, reference_root  # noqa: E402,F401
