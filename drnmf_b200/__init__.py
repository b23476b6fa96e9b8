"""Importable alias of the `dr-nmf_b200/` package directory (a hyphen cannot appear in a Python module name).

`import drnmf_b200` executes dr-nmf_b200/__init__.py with this package's __path__ pointing at that directory, so
`drnmf_b200.custom_layers`, `drnmf_b200.enhance`, ... resolve to the files under dr-nmf_b200/.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "dr-nmf_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
