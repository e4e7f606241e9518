"""Small occupancy calls for compute-sanitizer (memcheck / racecheck): uniform and ragged batches, compact output, cxyz,
multi-sigma channels, a 0.5 A grid.  Usage: compute-sanitizer --tool memcheck python profiles/scripts/sanitize_occ.py"""
import sys

import numpy as np

sys.path.insert(0, ".")
from moleculekit_b200 import workloads  # noqa: E402
from moleculekit_b200.tools import voxeldescriptors as vd  # noqa: E402

w = workloads.protein_pockets(B=3, n_atoms=400, box=29.0, radius=9.0, seed=3)
for kw in (dict(boxsize=[29.0, 22.0, 31.0], centers=w["centers"], voxelsize=1.0), dict(buffer=2.5, voxelsize=0.5)):
    a, _ = vd.getVoxelDescriptorsBatch(w["coords"], w["sigmas"], dtype=np.float32, transfer="dense", **kw)
    b, _ = vd.getVoxelDescriptorsBatch(w["coords"], w["sigmas"], dtype=np.float32, transfer="compact", **kw)
    c, _ = vd.getVoxelDescriptorsBatch(w["coords"], w["sigmas"], dtype=np.float32, layout="cxyz", **kw)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
rng = np.random.default_rng(1)
sg = rng.choice([0.0, 1.2, 1.7, 2.1], size=(400, 8))
f, _, _ = vd.getVoxelDescriptors(None, boxsize=[21, 19, 26], center=list(w["centers"][0]), voxelsize=1.0, userchannels=sg,
                                 usercoords=w["coords"][0])
print("sanitize_occ done", float(np.sum(f)))
