"""Config → object wiring shared by the CLI, the ZMQ node processes and the B200 engine.

Parity: reference ``murmura/utils/factories.py:16-190``.  String dispatch:
``leaf.X`` / ``wearables.X`` / ``synthetic.X`` adapters, ``examples.leaf.X`` /
``examples.wearables.X`` / ``models.X`` factories, otherwise a dotted ``module.attr`` path
(``murmura.`` prefixes are rewritten to this package so reference YAMLs keep working).
"""
from __future__ import annotations

import importlib
from typing import Any, Callable, Optional, Tuple

import torch
import torch.nn as nn

from murmura_b200.config.schema import Config


def _import_attr(path: str) -> Any:
    module_path, attr = path.rsplit(".", 1)
    if module_path == "murmura" or module_path.startswith("murmura."):
        module_path = "murmura_b200" + module_path[len("murmura"):]
    return getattr(importlib.import_module(module_path), attr)


def build_dataset_adapter(config: Config) -> Any:
    name = config.data.adapter
    n, seed = config.topology.num_nodes, config.experiment.seed
    if name.startswith("leaf."):
        from murmura_b200.examples.leaf import load_leaf_adapter
        return load_leaf_adapter(name.split(".")[1], num_nodes=n, seed=seed, **config.data.params)
    if name.startswith("wearables."):
        from murmura_b200.examples.wearables import load_wearable_adapter
        return load_wearable_adapter(dataset_type=name.split(".")[1], num_nodes=n, seed=seed, **config.data.params)
    if name.startswith("synthetic."):
        from murmura_b200.data.synthetic import load_synthetic_adapter
        return load_synthetic_adapter(name.split(".")[1], num_nodes=n, seed=seed, **config.data.params)
    return _import_attr(name)(**config.data.params)


def build_model_factory(config: Config) -> Callable[[], nn.Module]:
    path = config.model.factory
    if path.startswith("examples.leaf."):
        from murmura_b200.examples.leaf import get_leaf_model_factory
        return get_leaf_model_factory(path.split(".")[-1], **config.model.params)
    if path.startswith("examples.wearables."):
        from murmura_b200.examples.wearables import get_wearable_model_factory
        return get_wearable_model_factory(path.split(".")[-1], **config.model.params)
    if path.startswith("models."):
        path = "murmura_b200." + path
    fn = _import_attr(path)
    params = dict(config.model.params)
    return lambda: fn(**params)


def build_aggregator_factory(config: Config, model_factory: Callable[[], nn.Module],
                             device: Optional[torch.device] = None) -> Callable[[int], Any]:
    from murmura_b200 import aggregation as agg
    from murmura_b200.aggregation.base import calculate_model_dimension

    table = {"fedavg": agg.FedAvgAggregator, "krum": agg.KrumAggregator, "balance": agg.BALANCEAggregator,
             "sketchguard": agg.SketchguardAggregator, "ubar": agg.UBARAggregator,
             "evidential_trust": agg.EvidentialTrustAggregator}
    kind = config.aggregation.algorithm.lower()
    if kind not in table:
        raise ValueError(f"Unknown aggregation algorithm: {kind}")
    params = dict(config.aggregation.params)
    if kind == "sketchguard":
        params["model_dim"] = calculate_model_dimension(model_factory())
    if kind in ("sketchguard", "balance", "ubar", "evidential_trust"):
        params["total_rounds"] = config.experiment.rounds
    cls = table[kind]
    return lambda node_id: cls(**params)


def build_criterion(config: Config) -> Tuple[Optional[nn.Module], bool]:
    """Evidential iff the model factory is one of the wearable (Dirichlet-output) models."""
    evidential = config.model.factory.startswith("examples.wearables.")
    if not evidential:
        return None, False
    from murmura_b200.examples.wearables import get_evidential_loss
    return get_evidential_loss(num_classes=config.model.params.get("num_classes", 6),
                               annealing_epochs=config.experiment.rounds // 2, lambda_weight=0.1), True


def _model_attack(kind: Optional[str], config: Config):
    from murmura_b200.attacks import DirectedDeviationAttack, GaussianAttack
    n, pct, seed = config.topology.num_nodes, config.attack.percentage, config.experiment.seed
    if kind == "gaussian":
        return GaussianAttack(num_nodes=n, attack_percentage=pct,
                              noise_std=config.attack.params.get("noise_std", 10.0), seed=seed)
    if kind == "directed_deviation":
        return DirectedDeviationAttack(num_nodes=n, attack_percentage=pct,
                                       lambda_param=config.attack.params.get("lambda_param", -5.0), seed=seed)
    return None


def build_attack(config: Config) -> Optional[Any]:
    if not config.attack.enabled or not config.attack.type:
        return None
    if config.attack.type in ("gaussian", "directed_deviation"):
        return _model_attack(config.attack.type, config)
    if config.attack.type == "topology_liar":
        from murmura_b200.attacks import TopologyLiarAttack
        return TopologyLiarAttack(num_nodes=config.topology.num_nodes, attack_percentage=config.attack.percentage,
                                  seed=config.experiment.seed,
                                  model_attack=_model_attack(config.attack.params.get("model_attack_type"), config))
    return None


def build_mobility_model(config: Config):
    if config.mobility is None:
        return None
    from murmura_b200.topology.dynamic import MobilityModel
    m = config.mobility
    return MobilityModel(num_nodes=config.topology.num_nodes, area_size=m.area_size, comm_range=m.comm_range,
                         max_speed=m.max_speed, seed=m.seed, ensure_connected=m.ensure_connected)
