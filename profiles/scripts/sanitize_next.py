"""Small calls of the round-2 additions for compute-sanitizer: triclinic / compact wrapping (K9b, small and long groups),
hydrogen bonds (K12), ring detectors (K13).  Usage: compute-sanitizer --tool memcheck python profiles/scripts/sanitize_next.py"""
import sys

import numpy as np

sys.path.insert(0, ".")
from moleculekit_b200 import hbonds, ringpairs  # noqa: E402
from moleculekit_b200.wrapping import wrap_compact_unitcell, wrap_triclinic_unitcell  # noqa: E402

rng = np.random.default_rng(2)
sizes = [700, 1, 2, 3, 4, 5, 37] + [3] * 200
groups = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint32)
N, F = int(groups[-1]), 37
xyz = rng.normal(0, 60, size=(N, 3, F)).astype(np.float32)
L = 30.0
bv = np.repeat(np.array([[L, 0, 0], [0, L, 0], [L / 2, L / 2, L * 2 ** 0.5 / 2]])[:, :, None], F, axis=2)
cs = np.arange(0, 600, 5, dtype=np.uint32)
for mode in (None, 0, 1):
    c = xyz.copy()
    if mode is None:
        wrap_triclinic_unitcell(groups, c, bv, cs, np.zeros(3, np.float32))
    else:
        wrap_compact_unitcell(groups[1:-1].copy(), c, bv, np.zeros(0, np.uint32), np.ones(3, np.float32), mode)
hx = (rng.uniform(0, 1, size=(300, 3, 5)) * 12).astype(np.float32)
don = np.stack([np.arange(0, 300, 3), np.arange(1, 300, 3)], 1).astype(np.uint32)
hx[don[:, 1]] = hx[don[:, 0]] + (rng.normal(size=(100, 3, 5)) * 0.55).astype(np.float32)
acc = np.arange(2, 300, 3, dtype=np.uint32)
ones = np.ones(300, np.uint32)
box = np.full((3, 5), 12.0, np.float32)
nb = sum(len(x) for x in hbonds.calculate(don, acc, hx, box, ones, ones, dist_threshold=3.0, angle_threshold=100.0, intra=True))
nb += sum(len(x) for x in hbonds.calculate(don[:, :1].copy(), acc, hx, box, ones, ones, dist_threshold=3.0, ignore_hs=True))
ra = np.arange(60, dtype=np.uint32); st = np.arange(0, 61, 6, dtype=np.uint32)
cat = np.arange(60, 90, dtype=np.uint32); hal = np.stack([cat, cat + 100], 1).astype(np.uint32)
nr = 0
for res in (ringpairs.pipi_calculate(ra, st, st, hx, box, 8.0, 60.0, 9.0, 30.0),
            ringpairs.cationpi_calculate(ra, st, cat, hx, box, 8.0, 5.0),
            ringpairs.sigmahole_calculate(ra, st, hal, hx, box, 8.0, 5.0)):
    nr += sum(len(x) for x in res[0])
print("sanitize_next done", nb, nr)
