"""Experiment configuration (schema + loader)."""
from murmura_b200.config.schema import (Config, ExperimentConfig, TopologyConfig, AggregationConfig,
                                        AttackConfig, TrainingConfig, DataConfig, ModelConfig,
                                        DistributedConfig, MobilityConfig, DMTTConfig, B200Config)
from murmura_b200.config.loader import load_config, save_config

__all__ = ["Config", "ExperimentConfig", "TopologyConfig", "AggregationConfig", "AttackConfig",
           "TrainingConfig", "DataConfig", "ModelConfig", "DistributedConfig", "MobilityConfig",
           "DMTTConfig", "B200Config", "load_config", "save_config"]
