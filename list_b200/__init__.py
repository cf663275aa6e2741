"""Importable alias for the package directory
``learning-implicitly-from-spatial-transformers-network_b200/``.

The product package lives in that directory (the name the build contract
asks for), but a hyphenated name is not a Python identifier, so this shim
points ``list_b200.__path__`` at it and runs its ``__init__``.  Everything is
then reachable as ``list_b200.<module>``, e.g. the reference's plugin lookup
``utils.get_class('list_b200.network.models.LIST')`` (reference
``utils.py:20-26``) resolves here.
"""
import os as _os

_real = _os.path.join(
    _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
    "learning-implicitly-from-spatial-transformers-network_b200",
)
__path__ = [_real]
_init = _os.path.join(_real, "__init__.py")
with open(_init) as _f:
    exec(compile(_f.read(), _init, "exec"))
del _f, _init
