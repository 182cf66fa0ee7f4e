"""Importable alias of the package directory ``retrieval-augmented-visual-question-answering_b200``
(hyphens cannot appear in a Python module name)."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "retrieval-augmented-visual-question-answering_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f, _real
