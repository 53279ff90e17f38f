"""Importable alias of the `unet-rir_b200/` package directory (a hyphen cannot be imported).

`import unet_rir_b200` executes unet-rir_b200/__init__.py with this module's __path__ pointing
at that directory, so `unet_rir_b200.engine`, `unet_rir_b200.dl_models.u_net`, ... resolve there.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "unet-rir_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
