"""Import alias: the package directory is ``practical-collab-perception_b200`` (not a Python identifier).
``import pcp_b200`` and ``import pcp_b200.<sub>`` resolve to the SAME module objects as the real package
(no second copy of any submodule is ever created)."""
import importlib
import importlib.abc
import importlib.util
import sys

_REAL = "practical-collab-perception_b200"
_ALIAS = __name__


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname == _ALIAS or fullname.startswith(_ALIAS + "."):
            return importlib.util.spec_from_loader(fullname, self)
        return None

    def create_module(self, spec):
        return importlib.import_module(_REAL + spec.name[len(_ALIAS):])

    def exec_module(self, module):
        pass


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
sys.modules[_ALIAS] = importlib.import_module(_REAL)
