"""Importable name of the package that lives in ``humanoid-vision-system_b200/``.

The directory name required by the build contract contains hyphens and cannot be imported
directly; this shim points ``hvs_b200``'s module search path at it and runs its __init__.
"""
import os as _os

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "humanoid-vision-system_b200")
__path__ = [_pkg_dir]
__file__ = _os.path.join(_pkg_dir, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
