"""Import shim: the product package lives in the directory ``kaldi-lstm_b200/`` (the name the
build contract fixes); a hyphen cannot appear in a Python module name, so this importable alias
points its package path there and executes that package's __init__."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "kaldi-lstm_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
