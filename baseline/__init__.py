"""Reference-arm harnesses (unmodified reference install under baseline/_ref, NCCL+PyTorch baseline)."""
