"""Lazy re-exports for package ``__init__`` modules (PEP 562).

``import murmura_b200`` used to pull in the whole tree (torch data loaders, ZeroMQ, every aggregator); with an export table the
sub-module that defines a name is imported on first attribute access, ``from pkg import Name`` and ``from pkg import *`` keep
working, and ``dir(pkg)`` lists the public names.
"""
from __future__ import annotations

import importlib
from typing import Callable, Dict, Iterable, List, Tuple


def lazy_exports(package: str, table: Dict[str, Iterable[str]]) -> Tuple[Callable[[str], object], Callable[[], List[str]], List[str]]:
    """``table`` maps a sub-module (relative to ``package``; dotted paths allowed) to the names it contributes.
    Returns ``(__getattr__, __dir__, __all__)`` for the package module."""
    owner = {name: sub for sub, names in table.items() for name in names}
    exported = list(owner)

    def __getattr__(name: str):
        sub = owner.get(name)
        if sub is None:
            raise AttributeError(f"module {package!r} has no attribute {name!r}")
        value = getattr(importlib.import_module(f"{package}.{sub}"), name)
        import sys
        setattr(sys.modules[package], name, value)          # cache: next access is a plain attribute lookup
        return value

    def __dir__() -> List[str]:
        import sys
        return sorted(set(exported) | set(vars(sys.modules[package])))

    return __getattr__, __dir__, exported
