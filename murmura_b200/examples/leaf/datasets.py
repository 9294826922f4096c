"""LEAF FEMNIST / CelebA loaders and the by-writer client partitioner.

Capability parity with reference ``murmura/examples/leaf/datasets.py:23-199,300-415``: reads the
JSON shards produced by LEAF's ``preprocess.sh`` (``{users, num_samples, user_data:{u:{x,y}}}``),
keeps a ``user → indices`` map, and assigns writers to nodes round-robin after sorting by
(-sample count, id) and a ``RandomState(seed)`` shuffle.  FEMNIST pixels are decoded once into a
dense ``uint8→float`` tensor (quantised through 8 bits exactly like the reference's
PIL round-trip) so shards can be placed on the GPU wholesale; CelebA decodes lazily.
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Tuple

import numpy as np
import torch
from torch.utils.data import Dataset

from murmura_b200.models.cnn import LEAFCelebAModel, LEAFFEMNISTModel

_IMAGENET_MEAN = torch.tensor([0.485, 0.456, 0.406]).view(3, 1, 1)
_IMAGENET_STD = torch.tensor([0.229, 0.224, 0.225]).view(3, 1, 1)


def _read_leaf_split(split_dir: str, label: str):
    if not os.path.exists(split_dir):
        raise FileNotFoundError(f"LEAF {label} directory not found at {split_dir}")
    files = sorted(f for f in os.listdir(split_dir) if f.endswith(".json"))
    if not files:
        raise FileNotFoundError(f"No JSON files found in {split_dir}")
    print(f"Loading {len(files)} LEAF {label} files...")
    users: List[str] = []
    user_data: Dict[str, dict] = {}
    num_samples: List[int] = []
    for name in files:
        with open(os.path.join(split_dir, name), "r") as fh:
            blob = json.load(fh)
        users.extend(blob["users"])
        user_data.update(blob["user_data"])
        num_samples.extend(blob["num_samples"])
    return users, user_data, num_samples


class _LEAFBase(Dataset):
    def _index_users(self, users, user_data) -> Tuple[list, list]:
        xs, ys = [], []
        self.user_indices: Dict[str, List[int]] = {}
        for u in users:
            start = len(xs)
            xs.extend(user_data[u]["x"])
            ys.extend(user_data[u]["y"])
            self.user_indices[u] = list(range(start, len(xs)))
        return xs, ys

    def get_user_data(self, user: str) -> List[int]:
        return self.user_indices.get(user, [])

    def __len__(self) -> int:
        return len(self.all_targets)


class LEAFFEMNISTDataset(_LEAFBase):
    def __init__(self, data_path: str, split: str = "train", transform=None):
        self.split, self.transform = split, transform
        self.users, self.user_data, self.num_samples = _read_leaf_split(
            os.path.join(data_path, split), f"FEMNIST {split}")
        self.all_data, self.all_targets = self._index_users(self.users, self.user_data)
        pixels = np.asarray(self.all_data, dtype=np.float32).reshape(-1, 1, 28, 28)
        self._images = torch.from_numpy((pixels * 255).astype(np.uint8)).float().div_(255.0)
        self._labels = torch.as_tensor(self.all_targets, dtype=torch.long)
        print(f"LEAF FEMNIST {split}: {len(self.users)} users, {len(self.all_data)} samples")

    @property
    def tensors(self):
        return self._images, self._labels

    def __getitem__(self, idx):
        return self._images[idx], int(self._labels[idx])


class LEAFCelebADataset(_LEAFBase):
    def __init__(self, data_path: str, split: str = "train", image_size: int = 84, transform=None,
                 normalize: bool = True):
        self.split, self.image_size, self.transform, self.normalize = split, image_size, transform, normalize
        self.images_dir = os.path.join(data_path, "raw", "img_align_celeba")
        self.users, self.user_data, self.num_samples = _read_leaf_split(
            os.path.join(data_path, split), f"CelebA {split}")
        self.all_data, self.all_targets = self._index_users(self.users, self.user_data)
        print(f"LEAF CelebA {split}: {len(self.users)} users (celebrities), {len(self.all_data)} samples")

    def _resolve(self, name: str) -> str:
        raw = os.path.dirname(self.images_dir)
        for cand in (os.path.join(self.images_dir, name), os.path.join(raw, name),
                     os.path.join(os.path.dirname(raw), "raw", name)):
            if os.path.exists(cand):
                return cand
        raise FileNotFoundError(f"Could not find image {name} under {self.images_dir}")

    def __getitem__(self, idx):
        from PIL import Image
        img = Image.open(self._resolve(self.all_data[idx])).resize((self.image_size, self.image_size)).convert("RGB")
        target = int(self.all_targets[idx])
        if self.transform is not None:
            return self.transform(img), target
        t = torch.from_numpy(np.asarray(img, dtype=np.float32) / 255.0).permute(2, 0, 1)
        if self.normalize:
            t = (t - _IMAGENET_MEAN) / _IMAGENET_STD
        return t, target


def create_leaf_client_partitions(train_dataset, test_dataset, num_nodes: int,
                                  seed: int = 42) -> Tuple[List[List[int]], List[List[int]]]:
    """Writers common to both splits → nodes, round-robin after a seeded shuffle."""
    if not hasattr(train_dataset, "user_indices") or not hasattr(test_dataset, "user_indices"):
        raise ValueError("Both datasets must have user_indices for LEAF partitioning")
    common = sorted(set(train_dataset.user_indices) & set(test_dataset.user_indices))
    ranked = sorted(common, key=lambda u: (-len(train_dataset.user_indices[u]), u))
    np.random.RandomState(seed).shuffle(ranked)
    train_parts: List[List[int]] = [[] for _ in range(num_nodes)]
    test_parts: List[List[int]] = [[] for _ in range(num_nodes)]
    for pos, user in enumerate(ranked):
        train_parts[pos % num_nodes].extend(train_dataset.user_indices[user])
        test_parts[pos % num_nodes].extend(test_dataset.user_indices[user])
    print(f"Distributed ALL {len(ranked)} users across {num_nodes} clients")
    print(f"Train partition sizes: {[len(p) for p in train_parts]}")
    print(f"Test partition sizes: {[len(p) for p in test_parts]}")
    return train_parts, test_parts


def load_leaf_dataset(dataset_name: str, data_path: str):
    """→ ``(train_ds, test_ds, model, num_classes, input_size)``."""
    name = dataset_name.lower()
    if name == "femnist":
        return (LEAFFEMNISTDataset(data_path, "train"), LEAFFEMNISTDataset(data_path, "test"),
                LEAFFEMNISTModel(num_classes=62), 62, 28)
    if name == "celeba":
        return (LEAFCelebADataset(data_path, "train", image_size=84), LEAFCelebADataset(data_path, "test", image_size=84),
                LEAFCelebAModel(num_classes=2, image_size=84), 2, 84)
    raise ValueError(f"Dataset {dataset_name} not supported. Use 'femnist' or 'celeba'")
