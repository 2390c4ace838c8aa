"""Import alias for the hyphenated package directory ``craniofacialsd-vae_b200/``.

The product package lives in ``craniofacialsd-vae_b200/`` (the name the build
contract fixes).  A hyphen is not importable, so this stub re-points its own
``__path__`` at that directory: ``import sdvae_b200.model`` loads
``craniofacialsd-vae_b200/model.py``.
"""
import os as _os

_pkg_dir = _os.path.join(
    _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
    "craniofacialsd-vae_b200")
__path__ = [_pkg_dir]
__file__ = _os.path.join(_pkg_dir, "__init__.py")
with open(__file__, "r") as _f:
    exec(compile(_f.read(), __file__, "exec"))
