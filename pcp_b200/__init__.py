"""Import alias: the package directory is ``practical-collab-perception_b200`` (not an identifier),
``import pcp_b200`` resolves to the same package object."""
import importlib
import sys

_real = importlib.import_module("practical-collab-perception_b200")
sys.modules[__name__] = _real
