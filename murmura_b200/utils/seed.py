"""Global seeding.  Behaviour parity with reference ``murmura/utils/seed.py:8-21``: Python, NumPy and torch (CPU + every CUDA
device) generators are seeded with the same value and cuDNN is put into deterministic mode on GPU machines."""
import random

import numpy as np
import torch

_SEEDERS = (random.seed, np.random.seed, torch.manual_seed, torch.cuda.manual_seed_all)


def set_seed(seed: int) -> None:
    for seeder in _SEEDERS:
        seeder(seed)
    if torch.cuda.is_available():
        cudnn = torch.backends.cudnn
        cudnn.deterministic, cudnn.benchmark = True, False
