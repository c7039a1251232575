"""Importable alias of the product package, which lives in ``dfa-nerf_b200/`` (a directory
name Python cannot import directly).  Everything is re-exported from there."""
import os as _os

_impl = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'dfa-nerf_b200')
__path__.insert(0, _impl)
with open(_os.path.join(_impl, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_impl, '__init__.py'), 'exec'))
del _f
