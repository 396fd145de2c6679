"""Shared helpers for the parity tests."""

import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def sort_dump(dump):
    """Order a block dump (keys [nb,3] + per-block arrays) by key so two dumps can be compared."""
    k = dump["keys"]
    order = np.lexsort((k[:, 2], k[:, 1], k[:, 0]))
    return {name: np.asarray(arr)[order] for name, arr in dump.items()}


def sorted_keys(keys):
    k = np.asarray(keys).reshape(-1, 3)
    return k[np.lexsort((k[:, 2], k[:, 1], k[:, 0]))]


def blocks_checksum(dump_sorted):
    """Order-independent integer checksum of a sorted dump's float planes (bit patterns)."""
    bits = np.ascontiguousarray(dump_sorted["vox"]).view(np.uint32).astype(np.uint64)
    per_block = bits.reshape(bits.shape[0], -1).sum(axis=1, dtype=np.uint64)
    return per_block


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
