"""Importable alias of the package directory `video-dqn_b200/` (a hyphen is not a valid
Python identifier).  All code lives there; this module only redirects the search path."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                                 "video-dqn_b200"))
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
