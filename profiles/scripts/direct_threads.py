"""'direct' end-to-end step vs the number of host threads zero-filling the empty blocks (C3, packed pinned inputs)."""
import time, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from moleculekit_b200 import sharding as bench
print(bench.bind_to_gpu_numa(0))
from moleculekit_b200 import workloads
from moleculekit_b200.tools import voxeldescriptors as vd
dev = torch.device("cuda:0")
w = workloads.protein_pockets()
batch = vd.VoxelBatch(w["coords"], w["sigmas"], boxsize=w["boxsize"], centers=w["centers"], voxelsize=w["voxelsize"])
h_coords = vd.pinned_array(batch.coords.shape, np.float32); h_coords[:] = batch.coords
h_sig = vd.pinned_array(batch.sigmas.shape, np.float64); h_sig[:] = batch.sigmas
out = vd.pinned_array((batch.total_voxels, batch.C), np.float32)
kw = dict(boxsize=w["boxsize"], centers=w["centers"], voxelsize=w["voxelsize"], atom_offsets=batch.atom_offsets, device=dev)
for nt in (1, 2, 4, 6, 8, 12, 16, 32):
    os.environ["MKB_HOST_THREADS"] = str(nt)
    for _ in range(2): vd.getVoxelDescriptorsBatch(h_coords, h_sig, out=out, **kw)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(6): vd.getVoxelDescriptorsBatch(h_coords, h_sig, out=out, **kw)
    torch.cuda.synchronize(); print("threads", nt, (time.perf_counter() - t0) / 6 * 1e3)
