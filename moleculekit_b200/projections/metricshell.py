"""Drop-in for ``moleculekit.projections.metricshell.MetricShell``.

Density of `sel2` atoms in concentric shells around every `sel1` atom.  The reference materialises the whole
(frames, pairs) distance matrix through MetricDistance and histograms it with numpy column masks
(moleculekit/projections/metricshell.py:68-80,131-133,183-202); here the histogram is fused onto the distance
evaluation (K8, csrc/distance.cu) -- the matrix never exists and only (frames, centres, shells) counts come back.
Output values are identical: integer counts divided by the same float64 shell volumes.
"""
from __future__ import annotations

import logging

import numpy as np

from .metricdistance import MetricDistance
from .projection import Projection

logger = logging.getLogger(__name__)


class MetricShell(Projection):
    """Same constructor as the reference (metricshell.py:50-80): sel1, sel2, periodic, numshells=4, shellwidth=3,
    pbc (deprecated), gap (unused), truncate."""

    def __init__(self, sel1, sel2, periodic, numshells: int = 4, shellwidth: int = 3, pbc: bool | None = None,
                 gap: int | None = None, truncate: float | None = None):
        super().__init__()
        if pbc is not None:
            raise DeprecationWarning(
                "The `pbc` option is deprecated please use the `periodic` option as described in MetricDistance."
            )
        if isinstance(sel1, str) or isinstance(sel2, str):
            self.symmetrical = isinstance(sel1, str) and isinstance(sel2, str) and sel1 == sel2
        else:  # the reference's `sel1 == sel2` is only meaningful for strings; arrays compare by content here
            self.symmetrical = bool(np.array_equal(np.asarray(sel1), np.asarray(sel2)))
        self.metricdistance = MetricDistance(sel1=sel1, sel2=sel2, periodic=periodic, groupsel1=None, groupsel2=None,
                                             metric="distances", threshold=8, truncate=truncate)
        self.numshells = numshells
        self.shellwidth = shellwidth
        self.truncate = truncate
        self.description = None
        self.shellcenters = None
        self.device = None

    def _calculateMolProp(self, mol, props="all"):
        props = ("sel1", "sel2", "shellcenters", "shelledges", "shellvol") if props == "all" else props
        res = {}
        need_sel = any(p in props for p in ("sel1", "sel2", "shellcenters"))
        sel1 = self.metricdistance._getMolProp(mol, "sel1") if need_sel else None
        sel2 = self.metricdistance._getMolProp(mol, "sel2") if need_sel else None
        if "sel1" in props:
            res["sel1"] = sel1
        if "sel2" in props:
            res["sel2"] = sel2
        if "shellcenters" in props:
            # unique first atoms of the pair map (metricshell.py:92-95): sel1's atoms; when sel1 == sel2 every atom of
            # the selection appears in some pair, which is the same set
            res["shellcenters"] = np.where(sel1)[0]
            if not self.symmetrical and np.array_equal(sel1, sel2):
                # two DIFFERENT selection strings that pick the same atoms: the reference's distance matrix is condensed
                # (j > i pairs only) but its shell loop is keyed on string equality, so the centres are the first atoms
                # of the pairs -- every selected atom but the last (metricshell.py:92-95,183-186)
                res["shellcenters"] = res["shellcenters"][:-1]
        edges = np.arange(self.shellwidth * (self.numshells + 1), step=self.shellwidth)
        if "shelledges" in props:
            res["shelledges"] = edges
        if "shellvol" in props:
            res["shellvol"] = 4 / 3 * np.pi * (edges[1:] ** 3 - edges[:-1] ** 3)
        return res

    def project(self, mol) -> np.ndarray:
        """(numFrames, ncenters * numshells) float64 densities, centre-major like the reference."""
        from .. import distance_utils as du
        from .util import _NO_BOX, _box_for, digitize_chains

        props = self._getMolProp(mol, "all")
        sel1, sel2 = props["sel1"], props["sel2"]
        periodic = self.metricdistance.periodic
        if periodic == "chains":
            self.metricdistance._checkChains(mol, sel1, sel2)
        selfdist = np.array_equal(sel1, sel2)
        if selfdist and not self.symmetrical:
            return self._project_condensed_quirk(mol, props)
        i1 = np.where(sel1)[0].astype(np.uint32)
        i2 = np.where(sel2)[0].astype(np.uint32)
        coords, box = _box_for(mol, periodic, _NO_BOX)
        chains = digitize_chains(mol, periodic, i2)
        counts = du.shell_counts(coords, box, i1, i2, chains, selfdist, periodic is not None, props["shelledges"],
                                 truncate=self.truncate, device=self.device)
        dens = counts / props["shellvol"][None, None, :]
        return dens.reshape(mol.numFrames, len(i1) * self.numshells)

    def _project_condensed_quirk(self, mol, props):
        """sel1 and sel2 are different strings selecting the SAME atoms.  The reference then counts, for centre i, only the
        partners that come after it in the selection and has no row for the last atom (its `_shells` is keyed on string
        equality while the distance matrix is condensed, metricshell.py:183-202).  Reproduced literally from the condensed
        distances of K3: a rare configuration, correctness over speed."""
        distances = self.metricdistance.project(mol)
        if distances.ndim == 1:
            distances = distances[np.newaxis, :]
        idx = np.where(props["sel1"])[0]
        first = np.repeat(idx[:-1], np.arange(len(idx) - 1, 0, -1))  # first atom of every condensed pair, row-major
        centers, edges, vol = props["shellcenters"], props["shelledges"], props["shellvol"]
        out = np.ones((distances.shape[0], len(centers) * self.numshells)) * -1
        for i, c in enumerate(centers):
            cols = first == c
            for e in range(len(edges) - 1):
                inshell = (distances[:, cols] > edges[e]) & (distances[:, cols] <= edges[e + 1])
                out[:, i * self.numshells + e] = np.sum(inshell, axis=1) / vol[e]
        return out

    def getMapping(self, mol):
        from pandas import DataFrame

        centers = self._getMolProp(mol, "shellcenters")
        types, indexes, description = [], [], []
        for i in centers:
            for n in range(self.numshells):
                types += ["shell"]
                indexes += [i]
                description += [
                    "Density of sel2 atoms in shell {}-{} A centered on atom {} {} {}".format(
                        n * self.shellwidth, (n + 1) * self.shellwidth, mol.resname[i], mol.resid[i], mol.name[i])
                ]
        return DataFrame({"type": types, "atomIndexes": indexes, "description": description})
