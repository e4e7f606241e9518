"""Host prologues of the distance kernels: mirror of moleculekit/projections/util.py:12-223.

Same argument meaning, output shapes / dtypes and RuntimeError messages; the O(F*P) numpy post-passes of the
reference (truncate, `<= threshold`) are fused into the CUDA store instead of re-reading the matrix."""
from __future__ import annotations

import logging

import numpy as np

from .. import distance_utils as _du
from ..elements import masses_of

logger = logging.getLogger(__name__)

_NO_BOX = ("No periodic box dimensions given in the molecule/trajectory. "
           "If you want to calculate distance without wrapping, set the periodic option to None")
_NO_BOX_RED = ("No periodic box dimensions given in the molecule/trajectory. "
               "If you want to calculate distance without wrapping, set the `periodic` option to None")
_FRAME_MISMATCH = ("Different number of frames in mol.coords and mol.box. "
                   "Please ensure they both have the same number of frames")
_BAD_METRIC = "The metric you asked for is not supported. Check spelling and documentation"


def _box_for(mol, periodic, msg):
    coords, box = mol.coords, mol.box
    if periodic is not None:
        if box is None or np.sum(box) == 0:
            raise RuntimeError(msg)
    else:
        box = np.zeros((3, coords.shape[2]), dtype=np.float32)
    if box.shape[1] != coords.shape[2]:
        raise RuntimeError(_FRAME_MISMATCH)
    return np.ascontiguousarray(coords, dtype=np.float32), np.ascontiguousarray(box, dtype=np.float32)


def digitize_chains(mol, periodic, sel2_atoms):
    """uint32 per-atom chain ids deciding which pairs get the minimum-image wrap (util.py:45-56):
    None -> all 0; "chains" -> index of mol.chain; "selections" -> 1 everywhere, 2 on sel2's atoms."""
    if periodic is None:
        return np.zeros(mol.numAtoms, dtype=np.uint32)
    if periodic == "chains":
        return np.unique(mol.chain, return_inverse=True)[1].astype(np.uint32)
    if periodic == "selections":
        ch = np.ones(mol.numAtoms, dtype=np.uint32)
        ch[sel2_atoms] = 2
        return ch
    raise RuntimeError(f"Invalid periodic option {periodic}")


def pp_calcDistances(mol, sel1, sel2, periodic, metric: str = "distances", threshold: float = 8, gap=1,
                     truncate=None, device=None, exact: bool = True):
    """(F, n1*n2 | n1(n2-1)/2) float32 distances or bool contacts between two boolean atom masks."""
    selfdist = np.array_equal(sel1, sel2)
    sel1 = np.where(sel1)[0].astype(np.uint32)
    sel2 = np.where(sel2)[0].astype(np.uint32)
    coords, box = _box_for(mol, periodic, _NO_BOX)
    chains = digitize_chains(mol, periodic, sel2)
    if metric not in ("contacts", "distances"):
        raise RuntimeError(_BAD_METRIC)
    shape = (mol.numFrames, _du.n_columns(len(sel1), len(sel2), selfdist))
    results = np.zeros(shape, dtype=np.float32)
    res = _du.dist_trajectory(coords, box, sel1, sel2, chains, selfdist, periodic is not None, results,
                              device=device, metric=metric, truncate=truncate, threshold=threshold, exact=exact)
    return res


def get_reduced_distances(mol, sel1, sel2, periodic, metric: str = "distances", threshold: float = 8, truncate=None,
                          reduction1: str = "closest", reduction2: str = "closest", pairs: bool = False, device=None):
    """Group-wise minimum / centre-of-mass distances (util.py:88-223).  sel: 1-D mask (each atom its own group)
    or (G, N) boolean group masks."""
    # the one-hot expansion of 1-D selections only matters for `selfdist` and the "selections" chain ids
    s1 = np.asarray(sel1)
    s2 = np.asarray(sel2)
    if s1.ndim != 2:
        idx = np.where(s1)[0]
        s1 = np.zeros((len(idx), len(s1)), dtype=bool)
        s1[np.arange(len(idx)), idx] = True
    if s2.ndim != 2:
        idx = np.where(s2)[0]
        s2 = np.zeros((len(idx), len(s2)), dtype=bool)
        s2[np.arange(len(idx)), idx] = True

    coords, box = _box_for(mol, periodic, _NO_BOX_RED)
    selfdist = np.array_equal(s1, s2)
    chains = None
    if periodic is None:
        chains = np.zeros(mol.numAtoms, dtype=np.uint32)
    elif periodic == "chains":
        chains = np.unique(mol.chain, return_inverse=True)[1].astype(np.uint32)
    elif periodic == "selections":
        chains = np.ones(mol.numAtoms, dtype=np.uint32)
        chains[np.any(s2, axis=0)] = 2
    groups1 = [np.where(row)[0].tolist() for row in s1]
    groups2 = [np.where(row)[0].tolist() for row in s2]
    if pairs and len(groups1) != len(groups2):
        raise RuntimeError("If `pairs=True` mode is used, the number of groups in sel1 should match the number of "
                           "groups in sel2.")
    if selfdist:
        ncol = int((len(groups1) * (len(groups2) - 1)) / 2)
    else:
        ncol = len(groups1) if pairs else len(groups1) * len(groups2)
    mindist = np.zeros((mol.numFrames, ncol), dtype=np.float32)
    reduction_map = {"closest": 0, "com": 1}
    gch1 = np.array([chains[g[0]] for g in groups1], dtype=np.uint32)  # chain of each group's FIRST atom (util.py:174)
    gch2 = np.array([chains[g[0]] for g in groups2], dtype=np.uint32)
    masses = masses_of(mol.element)
    if metric not in ("contacts", "distances"):
        raise RuntimeError(_BAD_METRIC)
    r1, r2 = reduction_map[reduction1.lower()], reduction_map[reduction2.lower()]
    kw = dict(device=device, metric=metric, truncate=truncate, threshold=threshold)
    if not pairs:
        return _du.dist_trajectory_reduction(coords, box, groups1, groups2, gch1, gch2, selfdist, periodic is not None,
                                             masses, r1, r2, mindist, **kw)
    return _du.dist_trajectory_reduction_pairs(coords, box, groups1, groups2, gch1, gch2, periodic is not None, masses,
                                               r1, r2, mindist, **kw)


def pp_calcMinDistances(mol, sel1, sel2, periodic, metric: str = "distances", threshold: float = 8, truncate=None):
    return get_reduced_distances(mol=mol, sel1=sel1, sel2=sel2, periodic=periodic, metric=metric,
                                 threshold=threshold, truncate=truncate)
