"""Importable alias for the package directory `model-based-rl_b200/` (a hyphen is not a valid
Python identifier).  All code lives there; this module only redirects the import system."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      'model-based-rl_b200')
__path__ = [_real]
__file__ = _os.path.join(_real, '__init__.py')
with open(__file__) as _f:
  exec(compile(_f.read(), __file__, 'exec'))
