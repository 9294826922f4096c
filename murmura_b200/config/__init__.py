"""Experiment configuration (schema + loader)."""
from murmura_b200._lazy import lazy_exports

__getattr__, __dir__, __all__ = lazy_exports(__name__, {
    "schema": ["Config", "ExperimentConfig", "TopologyConfig", "AggregationConfig", "AttackConfig", "TrainingConfig", "DataConfig", "ModelConfig", "DistributedConfig", "MobilityConfig", "DMTTConfig", "B200Config"],
    "loader": ["load_config", "save_config"],
})
