"""CNN families: LEAF FEMNIST / CelebA CNNs + size variants, and a CIFAR-10 CNN.

Architectures follow reference ``murmura/examples/leaf/datasets.py:204-297`` and
``murmura/examples/leaf/models.py:12-216`` (layer names kept so state-dict keys line up):
FEMNIST baseline conv5(1→32)-pool-conv5(32→64)-pool-fc(3136→2048)-fc(2048→62) = 6,603,710
params; variants tiny/small/large/xlarge = 220,318 / 848,382 / 26,154,814 / 60,271,678.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class _TwoConvNet(nn.Module):
    """conv-relu-pool ×2 → fc1-relu → fc2 (the LEAF template)."""

    def __init__(self, in_ch: int, c1: int, c2: int, kernel: int, spatial_after: int, hidden: int,
                 num_classes: int):
        super().__init__()
        pad = kernel // 2
        self.num_classes = num_classes
        self.conv1 = nn.Conv2d(in_ch, c1, kernel_size=kernel, padding=pad)
        self.pool1 = nn.MaxPool2d(2, 2)
        self.conv2 = nn.Conv2d(c1, c2, kernel_size=kernel, padding=pad)
        self.pool2 = nn.MaxPool2d(2, 2)
        self.fc1 = nn.Linear(spatial_after * spatial_after * c2, hidden)
        self.fc2 = nn.Linear(hidden, num_classes)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self.pool1(F.relu(self.conv1(x)))
        x = self.pool2(F.relu(self.conv2(x)))
        return self.fc2(F.relu(self.fc1(x.flatten(1))))

    def parameter_count(self) -> int:
        return sum(p.numel() for p in self.parameters())


class LEAFFEMNISTModel(_TwoConvNet):
    def __init__(self, num_classes: int = 62):
        super().__init__(1, 32, 64, 5, 7, 2048, num_classes)


class FEMNISTBaseline(LEAFFEMNISTModel):
    pass


class FEMNISTTiny(_TwoConvNet):
    def __init__(self, num_classes: int = 62):
        super().__init__(1, 8, 16, 5, 7, 256, num_classes)


class FEMNISTSmall(_TwoConvNet):
    def __init__(self, num_classes: int = 62):
        super().__init__(1, 16, 32, 5, 7, 512, num_classes)


class FEMNISTLarge(_TwoConvNet):
    def __init__(self, num_classes: int = 62):
        super().__init__(1, 64, 128, 5, 7, 4096, num_classes)


class FEMNISTXLarge(nn.Module):
    """conv3(1→64)-conv3(64→128)-pool-conv3(128→256)-pool-fc4096-fc2048-fc(classes), dropout 0.5."""

    def __init__(self, num_classes: int = 62):
        super().__init__()
        self.num_classes = num_classes
        self.conv1 = nn.Conv2d(1, 64, kernel_size=3, padding=1)
        self.conv2 = nn.Conv2d(64, 128, kernel_size=3, padding=1)
        self.pool1 = nn.MaxPool2d(2, 2)
        self.conv3 = nn.Conv2d(128, 256, kernel_size=3, padding=1)
        self.pool2 = nn.MaxPool2d(2, 2)
        self.fc1 = nn.Linear(7 * 7 * 256, 4096)
        self.fc2 = nn.Linear(4096, 2048)
        self.fc3 = nn.Linear(2048, num_classes)
        self.dropout = nn.Dropout(0.5)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = F.relu(self.conv1(x))
        x = self.pool1(F.relu(self.conv2(x)))
        x = self.pool2(F.relu(self.conv3(x)))
        x = self.dropout(F.relu(self.fc1(x.flatten(1))))
        x = self.dropout(F.relu(self.fc2(x)))
        return self.fc3(x)

    def parameter_count(self) -> int:
        return sum(p.numel() for p in self.parameters())


class LEAFCelebAModel(_TwoConvNet):
    """84×84 RGB LeNet-style net, 2,219,692 params, Kaiming/normal(0.01) init."""

    def __init__(self, num_classes: int = 2, image_size: int = 84):
        super().__init__(3, 30, 50, 3, image_size // 4, 100, num_classes)
        self.image_size = image_size
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, 0, 0.01)
                nn.init.zeros_(m.bias)


class CIFARCNN(nn.Module):
    """Small CIFAR-10 CNN for BASELINE config 3: conv3(3→32)-conv3(32→64)-pool-conv3(64→128)-pool-fc256-fc10."""

    def __init__(self, num_classes: int = 10):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 32, 3, padding=1)
        self.conv2 = nn.Conv2d(32, 64, 3, padding=1)
        self.conv3 = nn.Conv2d(64, 128, 3, padding=1)
        self.fc1 = nn.Linear(128 * 8 * 8, 256)
        self.fc2 = nn.Linear(256, num_classes)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = F.relu(self.conv1(x))
        x = F.max_pool2d(F.relu(self.conv2(x)), 2)
        x = F.max_pool2d(F.relu(self.conv3(x)), 2)
        return self.fc2(F.relu(self.fc1(x.flatten(1))))


_VARIANTS = {"tiny": FEMNISTTiny, "small": FEMNISTSmall, "baseline": FEMNISTBaseline,
             "large": FEMNISTLarge, "xlarge": FEMNISTXLarge}


def get_model_variant(variant_name: str, num_classes: int = 62) -> nn.Module:
    try:
        return _VARIANTS[variant_name.lower()](num_classes=num_classes)
    except KeyError:
        raise ValueError(f"Unknown model variant: {variant_name}. Choose from: {list(_VARIANTS)}") from None
