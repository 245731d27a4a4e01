"""Tuning aid: bench.py against an alternative build of the library.
    PCP_LIB=path/to/libpcp_variant.so python tools/bench_with_lib.py [bench.py arguments]"""
import os, sys, runpy
sys.path.insert(0, os.getcwd())
from pcp_b200 import _lib
_lib.LIB_PATH = os.path.abspath(os.environ["PCP_LIB"])
sys.argv = ["bench.py"] + sys.argv[1:]
runpy.run_path("bench.py", run_name="__main__")
