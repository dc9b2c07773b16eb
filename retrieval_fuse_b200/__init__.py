"""Importable alias of the `retrieval-fuse_b200/` package directory (a hyphen
is not a valid Python identifier).  `import retrieval_fuse_b200.ops` resolves
to `retrieval-fuse_b200/ops.py`."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "retrieval-fuse_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
