"""Shim package in front of the reference's ``models/`` (a directory without ``__init__.py``):
``models.unet_model`` is shadowed by the B200 drop-in; any other module of the reference's
``models`` directory further down ``sys.path`` stays importable (see ``utils/__init__.py``)."""
import os as _os
import sys as _sys

_here = _os.path.abspath(_os.path.dirname(__file__))
for _entry in list(_sys.path):
    _cand = _os.path.abspath(_os.path.join(_entry or ".", "models"))
    if _cand != _here and _os.path.isdir(_cand) and _cand not in __path__:
        __path__.append(_cand)
