"""Wearable-sensor datasets: UCI-HAR, PAMAP2, PPG-DaLiA.

Capability parity with reference ``murmura/examples/wearables/datasets.py:12-531`` (same
on-disk formats, constructor arguments, feature layout: HAR 561-d; PAMAP2 sliding windows of
``window_size×40`` → 4000-d with NaN→column-mean and z-normalisation; PPG-DaLiA wrist signals
resampled to 4 Hz, ``window_size×6`` → 192-d).  All three keep their samples as dense tensors
(``.tensors``) so the B200 engine can place a client's shard on the GPU in one copy.
Windowing is vectorised with ``sliding_window_view`` instead of Python loops.
"""
from __future__ import annotations

import pickle
from pathlib import Path
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
from torch.utils.data import Dataset


def majority_windows(features: np.ndarray, activities: np.ndarray, size: int, stride: int,
                     label_index: Dict[int, int]) -> Tuple[np.ndarray, np.ndarray]:
    """Flattened sliding windows + majority label (smallest id wins ties, like ``np.unique``)."""
    n = len(features)
    if n < size:
        return np.empty((0, size * features.shape[1]), features.dtype), np.empty((0,), np.int64)
    starts = np.arange(0, n - size + 1, stride)
    win = np.lib.stride_tricks.sliding_window_view(features, size, axis=0)[starts]   # [W, F, size]
    flat = np.ascontiguousarray(win.transpose(0, 2, 1)).reshape(len(starts), -1)
    act = np.lib.stride_tricks.sliding_window_view(activities, size)[starts]         # [W, size]
    ids = np.unique(activities)
    counts = (act[:, :, None] == ids[None, None, :]).sum(axis=1)                      # [W, |ids|]
    major = ids[np.argmax(counts, axis=1)]
    keep = np.isin(major, list(label_index))
    labels = np.array([label_index[a] for a in major[keep]], dtype=np.int64)
    return flat[keep], labels


class _TensorBacked(Dataset):
    features: torch.Tensor
    labels: torch.Tensor

    def __len__(self) -> int:
        return len(self.labels)

    def __getitem__(self, idx):
        return self.features[idx], self.labels[idx]

    @property
    def tensors(self) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.features, self.labels

    def get_labels(self) -> np.ndarray:
        return self.labels.numpy()

    @property
    def num_features(self) -> int:
        return int(self.features.shape[1])


class UCIHARDataset(_TensorBacked):
    """30 subjects, 6 activities, 561 engineered features (``X_/y_/subject_<split>.txt``)."""
    ACTIVITIES = {1: "WALKING", 2: "WALKING_UPSTAIRS", 3: "WALKING_DOWNSTAIRS", 4: "SITTING",
                  5: "STANDING", 6: "LAYING"}
    NUM_FEATURES = 561
    NUM_CLASSES = 6

    def __init__(self, root: str, split: str = "train", normalize: bool = True):
        self.root, self.split, self.normalize = Path(root), split, normalize
        d = self.root / split
        self.features = torch.tensor(np.loadtxt(d / f"X_{split}.txt"), dtype=torch.float32)
        self.labels = torch.tensor(np.loadtxt(d / f"y_{split}.txt", dtype=np.int64) - 1, dtype=torch.long)
        self.subjects = np.loadtxt(d / f"subject_{split}.txt", dtype=np.int64)

    def get_subjects(self) -> np.ndarray:
        return self.subjects

    @property
    def num_classes(self) -> int:
        return self.NUM_CLASSES


class _WindowedDataset(_TensorBacked):
    activities: List[int]
    subject_ids: np.ndarray

    def _finish(self, feats: List[np.ndarray], labels: List[np.ndarray], subjects: List[np.ndarray],
                normalize: bool) -> None:
        x = np.vstack(feats)
        if normalize:
            self.mean = x.mean(axis=0)
            self.std = x.std(axis=0)
            self.std[self.std == 0] = 1
            x = (x - self.mean) / self.std
        self.features = torch.tensor(x, dtype=torch.float32)
        self.labels = torch.tensor(np.concatenate(labels), dtype=torch.long)
        self.subject_ids = np.concatenate(subjects)

    def get_subjects(self) -> np.ndarray:
        return self.subject_ids

    @property
    def num_classes(self) -> int:
        return len(self.activities)


class PAMAP2Dataset(_WindowedDataset):
    """9 subjects × 3 IMUs (+heart rate), 100 Hz; ``Protocol/subject10X.dat``."""
    ACTIVITY_LABELS = [1, 2, 3, 4, 5, 6, 7, 12, 13, 16, 17, 24]
    ACTIVITY_NAMES = {1: "lying", 2: "sitting", 3: "standing", 4: "walking", 5: "running", 6: "cycling",
                      7: "nordic_walking", 12: "ascending_stairs", 13: "descending_stairs",
                      16: "vacuum_cleaning", 17: "ironing", 24: "rope_jumping"}
    ACTIVITY_COL, HEART_RATE_COL = 1, 2
    IMU_STARTS = (3, 20, 37)
    FEATURE_COLS_PER_IMU = 13

    def __init__(self, root: str, subjects: Optional[List[int]] = None, activities: Optional[List[int]] = None,
                 window_size: int = 100, window_stride: int = 50, normalize: bool = True,
                 include_heart_rate: bool = True):
        self.root = Path(root)
        self.subjects = subjects or list(range(101, 110))
        self.activities = activities or self.ACTIVITY_LABELS
        self.window_size, self.window_stride = window_size, window_stride
        self.normalize, self.include_heart_rate = normalize, include_heart_rate
        self.activity_to_idx = {a: i for i, a in enumerate(self.activities)}
        cols = ([self.HEART_RATE_COL] if include_heart_rate else []) + [
            c for s in self.IMU_STARTS for c in range(s, s + self.FEATURE_COLS_PER_IMU)]
        feats, labels, subj = [], [], []
        for sid in self.subjects:
            path = self.root / "Protocol" / f"subject{sid}.dat"
            if not path.exists():
                continue
            raw = np.loadtxt(path)
            act = raw[:, self.ACTIVITY_COL].astype(int)
            keep = np.isin(act, self.activities)
            x = raw[keep][:, cols]
            col_mean = np.nan_to_num(np.nanmean(np.where(np.isnan(x), np.nan, x), axis=0), nan=0.0) \
                if len(x) else np.zeros(len(cols))
            x = np.where(np.isnan(x), col_mean[None, :], x)
            w, l = majority_windows(x, act[keep], window_size, window_stride, self.activity_to_idx)
            if len(w):
                feats.append(w); labels.append(l); subj.append(np.full(len(l), sid))
        if not feats:
            raise ValueError(f"No valid data found in {self.root}")
        self._finish(feats, labels, subj, normalize)


class PPGDaLiADataset(_WindowedDataset):
    """15 subjects, Empatica E4 wrist signals; ``S<k>/S<k>.pkl`` (EDA, TEMP, ACC×3, BVP at 4 Hz)."""
    ACTIVITY_LABELS = [1, 2, 3, 4, 5, 6, 7]
    ACTIVITY_NAMES = {0: "transient", 1: "sitting", 2: "ascending_stairs", 3: "descending_stairs",
                      4: "walking", 5: "cycling", 6: "driving", 7: "table_soccer"}

    def __init__(self, root: str, subjects: Optional[List[int]] = None, activities: Optional[List[int]] = None,
                 window_size: int = 32, window_stride: int = 16, normalize: bool = True,
                 use_wrist_only: bool = True):
        self.root = Path(root)
        self.subjects = subjects or list(range(1, 16))
        self.activities = activities or self.ACTIVITY_LABELS
        self.window_size, self.window_stride = window_size, window_stride
        self.normalize, self.use_wrist_only = normalize, use_wrist_only
        self.activity_to_idx = {a: i for i, a in enumerate(self.activities)}
        feats, labels, subj = [], [], []
        for sid in self.subjects:
            path = self.root / f"S{sid}" / f"S{sid}.pkl"
            if not path.exists():
                continue
            with open(path, "rb") as fh:
                rec = pickle.load(fh, encoding="latin1")
            x = self._resample(rec)
            act = rec["activity"].flatten().astype(int)
            n = min(len(x), len(act))
            x, act = x[:n], act[:n]
            keep = np.isin(act, self.activities)
            if not keep.any():
                continue
            w, l = majority_windows(x[keep], act[keep], window_size, window_stride, self.activity_to_idx)
            if len(w):
                feats.append(w); labels.append(l); subj.append(np.full(len(l), sid))
        if not feats:
            raise ValueError(f"No valid data found in {self.root}")
        self._finish(feats, labels, subj, normalize)

    @staticmethod
    def _resample(rec: Dict) -> np.ndarray:
        wrist = rec["signal"]["wrist"]
        eda, temp = wrist["EDA"].flatten(), wrist["TEMP"].flatten()
        acc = wrist["ACC"][::8, :]          # 32 Hz → 4 Hz
        bvp = wrist["BVP"].flatten()[::16]  # 64 Hz → 4 Hz
        n = min(len(eda), len(temp), len(acc), len(bvp))
        x = np.column_stack([eda[:n], temp[:n], acc[:n], bvp[:n]])
        return np.nan_to_num(x, nan=0.0).astype(np.float32)
