"""Two-rank NCCL check of the sharded drivers (runs only where >= 2 GPUs are visible; the world_size-2 logic itself is
covered on CPU by tests/test_sharding_gloo.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["MKB_ROOT"])
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from moleculekit_b200 import workloads
from moleculekit_b200.sharding import voxelize_sharded, project_sharded
from moleculekit_b200.tools import voxeldescriptors as vd
from moleculekit_b200.molecule_lite import MolLite
from moleculekit_b200.projections.metricdistance import MetricDistance
w = workloads.protein_pockets(B=5, n_atoms=300, box=20.0, radius=7.0, seed=9)
full, offs, rng = voxelize_sharded(w["coords"], w["sigmas"], boxsize=w["boxsize"], centers=w["centers"], gather=True)
single, dims, o2 = vd.getVoxelDescriptorsBatch(w["coords"], w["sigmas"], boxsize=w["boxsize"], centers=w["centers"], return_tensor=True)
assert torch.equal(full, single) and np.array_equal(offs, o2), "gathered shards differ from the single-GPU batch"
loc, loffs, (b, e) = voxelize_sharded(w["coords"], w["sigmas"], boxsize=w["boxsize"], centers=w["centers"])
assert torch.equal(loc, single[offs[b]:offs[e]])
g = np.random.default_rng(1)
mol = MolLite((g.normal(size=(40, 3, 9)) * 6).astype(np.float32), box=np.full((3, 9), 11.0, np.float32))
m = MetricDistance(np.arange(0, 15), np.arange(15, 40), periodic="selections")
assert np.array_equal(project_sharded(m, mol, gather=True), m.project(mol))
# frame-sharded wrapping (orthorhombic K9 and triclinic K9b) and hydrogen bonds (K12) against the single-GPU call
from moleculekit_b200.sharding import wrap_sharded, hbonds_sharded
from moleculekit_b200 import wrapping as wr
from moleculekit_b200.interactions import hbonds_calculate
N, F = 90, 9
xyz = (g.normal(size=(N, 3, F)) * 25).astype(np.float32)
bonds = np.array([[3 * k, 3 * k + 1] for k in range(30)] + [[3 * k, 3 * k + 2] for k in range(30)], dtype=np.uint32)
sel = np.zeros(N, bool); sel[:12] = True
for angles in (None, np.repeat(np.array([[60.0], [60.0], [90.0]], np.float32), F, axis=1)):
    for cell in ("rectangular", "compact", "triclinic"):
        a = MolLite(xyz.copy(), box=np.full((3, F), 14.0, np.float32), bonds=bonds, boxangles=angles)
        b = MolLite(xyz.copy(), box=np.full((3, F), 14.0, np.float32), bonds=bonds, boxangles=angles)
        wr.wrap(a, wrapsel=sel, unitcell=cell)
        assert wrap_sharded(b, wrapsel=sel, unitcell=cell) == (0, F)
        assert np.array_equal(a.coords.view(np.uint32), b.coords.view(np.uint32)), (cell, angles is None)
hx = (g.uniform(0, 1, size=(N, 3, F)) * 9).astype(np.float32)
don = np.stack([np.arange(0, N, 3), np.arange(1, N, 3)], 1).astype(np.uint32)
hx[don[:, 1]] = hx[don[:, 0]] + (g.normal(size=(30, 3, F)) * 0.55).astype(np.float32)
hm = MolLite(hx, box=np.full((3, F), 9.0, np.float32))
acc = np.arange(2, N, 3, dtype=np.uint32)
one = hbonds_calculate(hm, don, acc, np.ones(N, bool), dist_threshold=3.0, angle_threshold=100.0)
two = hbonds_sharded(hm, don, acc, np.ones(N, bool), dist_threshold=3.0, angle_threshold=100.0)
assert len(two) == F and all(np.array_equal(x, y) for x, y in zip(one, two)) and sum(len(x) for x in one) > 0
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_nccl(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MKB_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout
