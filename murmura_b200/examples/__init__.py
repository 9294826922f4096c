"""Bundled workloads (LEAF, wearables, programmatic example, YAML configs)."""
