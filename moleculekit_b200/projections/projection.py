"""Base class of the trajectory projections (mirror of moleculekit/projections/projection.py:13-92): a per-object
cache of molecule-derived properties (here: resolved selections) plus the abstract project / getMapping pair."""
from __future__ import annotations

import abc
from copy import deepcopy


class Projection(abc.ABC):
    def __init__(self):
        self._cache = {}

    @abc.abstractmethod
    def project(self, mol):
        """Project a molecule: one output row per frame."""

    @abc.abstractmethod
    def getMapping(self, mol):
        """DataFrame describing every projected dimension."""

    @abc.abstractmethod
    def _calculateMolProp(self, mol, props="all"):
        ...

    def _setCache(self, mol):
        """Resolve the molecule-dependent properties once; later project() calls reuse them (projection.py:67-69)."""
        self._cache.update(self._calculateMolProp(mol))

    def _getMolProp(self, mol, prop):
        if prop in self._cache:
            found = self._cache
        else:
            found = self._calculateMolProp(mol, "all" if prop == "all" else [prop])
        return found if prop == "all" else found[prop]

    def copy(self):
        return deepcopy(self)
