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
