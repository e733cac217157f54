"""Packing helpers with the reference's names (code/sphops/utils/__init__.py:31-38)."""
import torch


def pack(*xx):
    """Concatenate per-section tensors; returns (packed, [section lengths])."""
    return torch.cat(xx, dim=0), [int(t.shape[0]) for t in xx]


def unpack(packed_x, sections):
    """Inverse of `pack`: views of `packed_x`, one per section."""
    return torch.split(packed_x, list(sections))
