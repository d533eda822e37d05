"""Host-side handling of study/patient ids.

The reference receives ``patient_ids`` as a numpy array of unicode strings
("p<subject>_s<study>", modules/dataloaders_v0401.py:83,115) and only ever tests them for
equality (models/model_pretrain_finetune_v0520.py:489, :422).  The kernels work on int32 keys,
so opaque keys are factorised ONCE on the host into dense codes with the same equality
structure; nothing of size N x N is ever built on the host.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch


def factorize(ids) -> np.ndarray:
    """numpy array of any dtype (incl. <U strings) -> dense int32 codes, equal iff keys equal."""
    arr = np.asarray(ids).reshape(-1)
    if arr.dtype.kind in "iu" and arr.size and arr.min() >= 0 and arr.max() < 2**31:
        return arr.astype(np.int32, copy=False)
    _, inv = np.unique(arr, return_inverse=True)
    return inv.reshape(-1).astype(np.int32)


def combine_keys(ids_a, ids_b) -> np.ndarray:
    """Conjunction key (a AND b), e.g. (patient, study): factorise the pairs jointly."""
    a = factorize(ids_a).astype(np.int64)
    b = factorize(ids_b).astype(np.int64)
    if a.shape != b.shape:
        raise ValueError(f"id arrays differ in length: {a.shape} vs {b.shape}")
    return factorize(a * (int(b.max()) + 1 if b.size else 1) + b)


@dataclass
class DeviceIds:
    """int32 key arrays on the compute device (key2 is the optional second component)."""
    key: torch.Tensor
    key2: Optional[torch.Tensor] = None

    def __len__(self) -> int:
        return int(self.key.shape[0])

    def slice(self, n: int) -> "DeviceIds":
        return DeviceIds(self.key[:n], None if self.key2 is None else self.key2[:n])

    def index(self, idx: torch.Tensor) -> "DeviceIds":
        return DeviceIds(self.key[idx].contiguous(), None if self.key2 is None else self.key2[idx].contiguous())


def _tensor_to_keys(t: torch.Tensor, device: torch.device) -> DeviceIds:
    if t.dim() != 1:
        t = t.reshape(-1)
    if t.dtype in (torch.int32, torch.int16, torch.int8, torch.uint8):
        return DeviceIds(t.to(device=device, dtype=torch.int32).contiguous())
    if t.dtype == torch.int64:
        # exact without a sync or a sort: compare (low word, high word) as a two-component key
        t = t.to(device)
        lo = (t & 0xFFFFFFFF).to(torch.int32)          # wraps modulo 2^32: still injective per word
        hi = (t >> 32).to(torch.int32)
        return DeviceIds(lo.contiguous(), hi.contiguous())
    raise TypeError(f"ids tensor must be an integer tensor, got {t.dtype}")


def to_device_ids(ids, device: torch.device, n: Optional[int] = None) -> tuple[DeviceIds, Optional[np.ndarray]]:
    """Accepts what the reference passes (numpy array of strings/ints), a (patient, study)
    tuple of such arrays, or integer torch tensors (already on the device for the benchmark
    path).  ``n`` truncates to the first n ids (:488).  Returns the device keys and, when the
    ids came from the host, the host int32 codes (used for the MPC row filter without a sync)."""
    if isinstance(ids, DeviceIds):
        return (ids if n is None else ids.slice(n)), None
    if isinstance(ids, tuple) and len(ids) == 2 and not isinstance(ids[0], (int, np.integer, str)):
        a, b = ids
        if isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor):
            ka, kb = _tensor_to_keys(a, device), _tensor_to_keys(b, device)
            if ka.key2 is not None or kb.key2 is not None:
                raise TypeError("(patient, study) tensor pairs must be int32")
            out = DeviceIds(ka.key, kb.key)
            return (out if n is None else out.slice(n)), None
        codes = combine_keys(a, b)
    elif isinstance(ids, torch.Tensor):
        out = _tensor_to_keys(ids, device)
        return (out if n is None else out.slice(n)), None
    else:
        codes = factorize(ids)
    if n is not None:
        codes = codes[:n]
    codes = np.ascontiguousarray(codes)
    dev = torch.from_numpy(codes).to(device, non_blocking=False)
    return DeviceIds(dev), codes


def multi_view_rows(codes: np.ndarray) -> np.ndarray:
    """Host version of the MPC row filter (:424-426): rows whose key occurs at least twice."""
    _, inv, cnt = np.unique(codes, return_inverse=True, return_counts=True)
    return np.nonzero(cnt[inv.reshape(-1)] > 1)[0].astype(np.int32)
