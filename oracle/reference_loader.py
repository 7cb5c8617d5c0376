"""Import the UNMODIFIED reference (loganbvh/superscreen 0.13.0) from /root/reference.

TEST INFRASTRUCTURE ONLY.  Works only in the build container, where /root/reference
is mounted; nothing that runs on the GPU box may call this.  It is used by
``oracle/make_golden.py`` to generate the committed fixtures under ``tests/golden/``
and by ``tests/test_oracle_vs_reference.py`` (skipped when the reference is absent)
to pin the numpy restatement in ``oracle/port.py`` to the live reference code.

The reference is pure Python but imports matplotlib / shapely / meshpy / pint / h5py /
IPython at module import time; none of those is installed here.  The numeric hot
path (distance.py, fem.py, device/mesh.py, device/utils.py, solver/solve_film.py,
solver/solve.py:biot_savart_film_to_film, sources/current.py) never calls into them,
so they are replaced by ``MagicMock`` modules.
"""
from __future__ import annotations

import os
import sys
import types
from unittest import mock

REFERENCE_ROOT = os.environ.get("SUPERSCREEN_REFERENCE", "/root/reference")

_STUBS = [
    "matplotlib", "matplotlib.path", "matplotlib.pyplot", "matplotlib.tri",
    "matplotlib.patches", "matplotlib.colors", "matplotlib.cm", "matplotlib.ticker",
    "mpl_toolkits", "mpl_toolkits.axes_grid1", "mpl_toolkits.axes_grid1.axes_divider",
    "h5py", "shapely", "shapely.geometry", "shapely.geometry.polygon", "shapely.ops",
    "shapely.affinity", "shapely.validation", "meshpy", "meshpy.triangle", "pint",
    "IPython", "IPython.display",
]


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "superscreen"))


def load_reference() -> types.ModuleType:
    """Returns the imported reference package ``superscreen`` (stubbed third parties)."""
    if "superscreen" in sys.modules and getattr(
        sys.modules["superscreen"], "__file__", ""
    ).startswith(REFERENCE_ROOT):
        return sys.modules["superscreen"]
    if not reference_available():
        raise ImportError(f"reference not found under {REFERENCE_ROOT}")
    for name in _STUBS:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = mock.MagicMock(name=name)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import superscreen  # noqa: F401
    finally:
        sys.path.remove(REFERENCE_ROOT)
    return sys.modules["superscreen"]
