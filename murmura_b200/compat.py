"""Drop-in import alias: make ``import murmura`` resolve to this package.

    import murmura_b200.compat; murmura_b200.compat.install_alias()
    from murmura import Network, Config            # existing user code keeps working
    from murmura.aggregation import KrumAggregator

Not installed by default so the unmodified reference package (``baseline/_ref``) can live in the same interpreter.
"""
from __future__ import annotations

import importlib
import pkgutil
import sys


def install_alias(name: str = "murmura") -> None:
    import murmura_b200
    sys.modules[name] = murmura_b200
    for mod in pkgutil.walk_packages(murmura_b200.__path__, prefix="murmura_b200."):
        if ".ops" in mod.name or ".parallel" in mod.name or mod.name.endswith("__main__"):
            continue
        try:
            sys.modules[name + mod.name[len("murmura_b200"):]] = importlib.import_module(mod.name)
        except Exception:        # optional submodules (e.g. missing third-party deps) are skipped
            pass
