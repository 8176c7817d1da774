"""Importable alias of the package directory `ceigm-unet_b200/` (a hyphen cannot appear in a Python module name).
`import ceigm_unet_b200` executes ceigm-unet_b200/__init__.py with this module as the package."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "ceigm-unet_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
