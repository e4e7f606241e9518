"""A minimal stand-in for ``moleculekit.molecule.Molecule`` carrying exactly what the hot path consumes.

The accelerated entry points are duck-typed: a real moleculekit ``Molecule`` works unchanged.  ``MolLite`` exists so
tests, the benchmark and users without moleculekit can feed arrays: coordinates are float32 (natoms, 3, nframes)
frame-minor and the box float32 (3, nframes), as in moleculekit/molecule.py:144-146.

String atom selections are the reference's VMD-like language (moleculekit/atomselect, out of scope); MolLite only
resolves strings registered in ``named_selections`` and otherwise asks for index / boolean arrays.
"""
from __future__ import annotations

import copy

import numpy as np


class MolLite:
    def __init__(self, coords, box=None, element=None, name=None, resname=None, resid=None, chain=None,
                 segid=None, named_selections=None, frame: int = 0, bonds=None, boxangles=None):
        coords = np.asarray(coords, dtype=np.float32)
        if coords.ndim == 2:
            coords = coords[:, :, None]
        self.coords = np.ascontiguousarray(coords)
        n, _, f = self.coords.shape
        self.box = np.zeros((3, f), dtype=np.float32) if box is None else np.ascontiguousarray(box, dtype=np.float32)

        def col(v, default, dtype=object):
            if v is None:
                return np.array([default] * n, dtype=dtype)
            return np.asarray(v, dtype=dtype)

        self.element = col(element, "")
        self.name = col(name, "")
        self.resname = col(resname, "")
        self.chain = col(chain, "")
        self.segid = col(segid, "")
        self.resid = np.zeros(n, dtype=np.int64) if resid is None else np.asarray(resid, dtype=np.int64)
        self.named_selections = dict(named_selections or {})
        self.frame = frame
        self.bonds = np.zeros((0, 2), dtype=np.uint32) if bonds is None else \
            np.asarray(bonds, dtype=np.uint32).reshape(-1, 2)
        self.boxangles = np.full((3, f), 90.0, dtype=np.float32) if boxangles is None else \
            np.asarray(boxangles, dtype=np.float32)

    @property
    def numAtoms(self) -> int:
        return self.coords.shape[0]

    @property
    def numFrames(self) -> int:
        return self.coords.shape[2]

    def copy(self):
        return copy.deepcopy(self)

    def atomselect(self, sel, indexes: bool = False):
        """bool mask (or indexes) for a selection given as bool mask, integer index array or a registered string."""
        if isinstance(sel, str):
            if sel == "all":
                mask = np.ones(self.numAtoms, dtype=bool)
            elif sel in self.named_selections:
                mask = np.asarray(self.named_selections[sel], dtype=bool)
            else:
                raise NotImplementedError(
                    f"MolLite cannot parse the selection string {sel!r}: the VMD selection language lives in "
                    "moleculekit.atomselect (outside the accelerated path). Pass a boolean mask / index array, "
                    "register the mask in named_selections, or hand a moleculekit Molecule to the same API.")
        else:
            sel = np.asarray(sel)
            if sel.dtype == bool:
                mask = sel.copy()
            else:
                mask = np.zeros(self.numAtoms, dtype=bool)
                mask[sel] = True
        return np.where(mask)[0] if indexes else mask

    def get(self, field: str, sel=None):
        idx = slice(None) if sel is None else self.atomselect(sel)
        if field == "coords":
            return self.coords[idx, :, self.frame]
        return getattr(self, field)[idx]
