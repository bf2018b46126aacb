"""Importable alias of the product package.

The package directory the build contract names is `instance-segmentation-road-project_b200/`
(hyphens, not a Python identifier); this shim maps the importable name `masklab_b200` onto
that directory so that `import masklab_b200.layers` loads
`instance-segmentation-road-project_b200/layers/__init__.py`.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "instance-segmentation-road-project_b200")
__path__.insert(0, _PKG_DIR)

with open(_os.path.join(_PKG_DIR, "_package_init.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, "_package_init.py"), "exec"))
