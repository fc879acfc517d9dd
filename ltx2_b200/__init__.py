"""Importable alias for the product package, which lives in ``ltx-2-mlx_b200/``.

The contract names the package directory ``ltx-2-mlx_b200`` -- not a valid Python
identifier -- so this stub makes ``import ltx2_b200`` resolve to that directory.
"""
import os as _os

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "ltx-2-mlx_b200")
__path__ = [_pkg_dir]
__file__ = _os.path.join(_pkg_dir, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
