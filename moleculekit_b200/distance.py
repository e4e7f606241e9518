"""Drop-in for the hot-path functions of ``moleculekit.distance``: cdist, pdist, squareform, calculate_contacts
(moleculekit/distance.py:221-412).  `find_clashes` (kd-tree + bond logic) is outside the accelerated path."""
from __future__ import annotations

import numpy as np

from . import distance_utils as _du
from .projections.util import _NO_BOX, _box_for, digitize_chains


def cdist(coords1: np.ndarray, coords2: np.ndarray) -> np.ndarray:
    """(N, M) float32 Euclidean distances between two (N, D) / (M, D) point sets."""
    assert coords1.ndim == 2, "cdist only supports 2D arrays"
    assert coords2.ndim == 2, "cdist only supports 2D arrays"
    assert coords1.shape[1] == coords2.shape[1], "Second dimension of input arguments must match"
    a = coords1 if coords1.dtype == np.float32 else coords1.astype(np.float32)
    b = coords2 if coords2.dtype == np.float32 else coords2.astype(np.float32)
    results = np.zeros((a.shape[0], b.shape[0]), dtype=np.float32)
    _du.cdist(a, b, results)
    return results


def pdist(coords: np.ndarray) -> np.ndarray:
    """Condensed (N(N-1)/2,) float32 pairwise distances."""
    assert coords.ndim == 2, "pdist only supports 2D arrays"
    a = coords if coords.dtype == np.float32 else coords.astype(np.float32)
    n = a.shape[0]
    results = np.zeros(int(n * (n - 1) / 2), dtype=np.float32)
    _du.pdist(a, results)
    return results


def squareform(distances: np.ndarray) -> np.ndarray:
    """Condensed vector -> symmetric (N, N) matrix with zero diagonal."""
    return np.array(_du.squareform(distances.astype(np.float32)))


def calculate_contacts(mol, sel1: np.ndarray, sel2: np.ndarray, periodic, threshold: float = 4, device=None) -> list:
    """Per frame, the (n, 2) uint32 atom-index pairs (one atom from each boolean mask) within `threshold`,
    in the reference's order (sel1 ascending, then sel2 ascending)."""
    assert isinstance(sel1, np.ndarray) and sel1.dtype == bool
    assert isinstance(sel2, np.ndarray) and sel2.dtype == bool
    selfdist = np.array_equal(sel1, sel2)
    sel1 = np.where(sel1)[0].astype(np.uint32)
    sel2 = np.where(sel2)[0].astype(np.uint32)
    coords, box = _box_for(mol, periodic, _NO_BOX)
    chains = digitize_chains(mol, periodic, sel2)
    off, pairs = _du.contacts_trajectory_arrays(coords, box, sel1, sel2, chains, selfdist, periodic is not None,
                                                threshold, device=device)
    return [pairs[off[f]:off[f + 1]].reshape(-1, 2) for f in range(len(off) - 1)]
