"""Shim package in front of the reference's ``utils/`` (a directory without ``__init__.py``).

This package shadows only the hot-path modules (``losses``, ``metrics``); every other module the
reference's scripts import from ``utils`` (``utils.dataset``, ``utils.augmentations``,
``utils.transforms`` — scripts/train.py:17) must still come from the reference tree. So the
same-named directories found further down ``sys.path`` (scripts/train.py:11-14 inserts the
reference root itself) are appended to this package's search path, after our own directory.
"""
import os as _os
import sys as _sys

_here = _os.path.abspath(_os.path.dirname(__file__))
for _entry in list(_sys.path):
    _cand = _os.path.abspath(_os.path.join(_entry or ".", "utils"))
    if _cand != _here and _os.path.isdir(_cand) and _cand not in __path__:
        __path__.append(_cand)
