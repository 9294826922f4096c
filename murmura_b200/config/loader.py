"""YAML / JSON config I/O (parity: reference ``murmura/config/loader.py:11-66``)."""
from __future__ import annotations

import json
from pathlib import Path
from typing import Union

import yaml

from murmura_b200.config.schema import Config

_YAML = {".yaml", ".yml"}


def _kind(path: Path) -> str:
    if path.suffix in _YAML:
        return "yaml"
    if path.suffix == ".json":
        return "json"
    raise ValueError(f"Unsupported config format: {path.suffix}. Use .yaml, .yml, or .json")


def load_config(config_path: Union[str, Path]) -> Config:
    path = Path(config_path)
    if not path.exists():
        raise FileNotFoundError(f"Config file not found: {path}")
    kind = _kind(path)
    with open(path, "r") as fh:
        data = yaml.safe_load(fh) if kind == "yaml" else json.load(fh)
    return Config(**data)


def save_config(config: Config, output_path: Union[str, Path]) -> None:
    path = Path(output_path)
    kind = _kind(path)
    data = config.model_dump()
    with open(path, "w") as fh:
        if kind == "yaml":
            yaml.safe_dump(data, fh, default_flow_style=False, sort_keys=False)
        else:
            json.dump(data, fh, indent=2)
