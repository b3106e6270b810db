"""Import the reference's modules from the bytecode files oracle/compile_pyref.py wrote
(TEST INFRASTRUCTURE).  `install(directory)` puts a finder in front of sys.meta_path that serves
`name` from `directory/name.bc` and packages from `directory/name/__init__.bc`; `__file__` is
set to where `name.py` would sit, so code that looks for files beside itself
(clustering/neighbors.py:97-98 loads cneighbors.so that way) keeps working."""
import importlib.abc
import importlib.util
import marshal
import os
import sys


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self, root):
        self.root = root

    def _path(self, fullname):
        parts = fullname.split('.')
        base = os.path.join(self.root, *parts)
        if os.path.exists(os.path.join(base, '__init__.bc')):
            return os.path.join(base, '__init__.bc'), True
        if os.path.exists(base + '.bc'):
            return base + '.bc', False
        return None, False

    def find_spec(self, fullname, path=None, target=None):
        bc, is_pkg = self._path(fullname)
        if bc is None:
            return None
        spec = importlib.util.spec_from_loader(fullname, self, origin=bc[:-3] + '.py', is_package=is_pkg)
        if is_pkg:
            spec.submodule_search_locations = [os.path.dirname(bc)]
        return spec

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        bc, _ = self._path(module.__name__)
        with open(bc, 'rb') as f:
            data = f.read()
        code = marshal.loads(data[16:])          # PEP 552 header: magic, flags, mtime/hash, size
        module.__file__ = bc[:-3] + '.py'
        exec(code, module.__dict__)


def install(directory):
    finder = _Finder(directory)
    sys.meta_path.insert(0, finder)
    return finder


def uninstall(finder):
    if finder in sys.meta_path:
        sys.meta_path.remove(finder)
