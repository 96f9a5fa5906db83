"""Import the UNMODIFIED reference (nerf-dbr ``src/``) for baselines and integration tests.  TEST INFRASTRUCTURE.

Where it comes from, in order: ``$NERF_DBR_REFERENCE`` (a checkout), ``/root/reference`` (the build container),
``oracle/_ref`` (the git-ignored copy ``tools/vendor_reference.sh`` makes so that the sources travel to the GPU box
with the gpurun snapshot).  ``matplotlib`` is not installed in this image and the reference imports it at module
import (benchmark_suite.py:8, trainer.py:13) without using it on the timed path, so an empty stand-in is registered
first.  Nothing under ``nerf_dbr_b200/`` imports this module.
"""
from __future__ import annotations

import os
import sys
import types
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
VENDORED = os.path.join(_HERE, "_ref")


def reference_root() -> Optional[str]:
    """Directory holding the reference's ``src`` package, or None."""
    for root in (os.environ.get("NERF_DBR_REFERENCE"), "/root/reference", VENDORED):
        if root and os.path.isfile(os.path.join(root, "src", "benchmark", "pytorch_renderers.py")):
            return root
    return None


def _stub_matplotlib() -> None:
    try:
        import matplotlib.pyplot  # noqa: F401
        return
    except Exception:
        pass

    class _Anything(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return lambda *a, **k: None

    mpl, plt = _Anything("matplotlib"), _Anything("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)


def import_reference():
    """Returns the reference's ``src`` package (``src.benchmark``, ``src.models`` ... importable afterwards), or
    raises ImportError naming what to run."""
    root = reference_root()
    if root is None:
        raise ImportError("the nerf-dbr reference is not available: run tools/vendor_reference.sh in the build "
                          "container (copies /root/reference/src into the git-ignored oracle/_ref/)")
    _stub_matplotlib()
    if root not in sys.path:
        sys.path.insert(0, root)
    import src.benchmark  # noqa: F401
    import src.models.nerf  # noqa: F401
    import src  # noqa: F401
    return sys.modules["src"]
