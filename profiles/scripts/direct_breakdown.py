"""Where the 'direct' end-to-end step goes (C3, 256 pockets, packed pinned inputs as in bench.py): kernel-side PCIe
stores vs host zero-fill vs upload + prologue."""
import time, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from moleculekit_b200 import sharding as bench
print(bench.bind_to_gpu_numa(0))
from moleculekit_b200 import workloads, occupancy_utils as occ
from moleculekit_b200.tools import voxeldescriptors as vd

dev = torch.device("cuda:0")
w = workloads.protein_pockets()
batch = vd.VoxelBatch(w["coords"], w["sigmas"], boxsize=w["boxsize"], centers=w["centers"], voxelsize=w["voxelsize"])
h_coords = vd.pinned_array(batch.coords.shape, np.float32); h_coords[:] = batch.coords
h_sig = vd.pinned_array(batch.sigmas.shape, np.float64); h_sig[:] = batch.sigmas
out = vd.pinned_array((batch.total_voxels, batch.C), np.float32)
kw = dict(boxsize=w["boxsize"], centers=w["centers"], voxelsize=w["voxelsize"], atom_offsets=batch.atom_offsets, device=dev)
call = lambda **e: vd.getVoxelDescriptorsBatch(h_coords, h_sig, **kw, **e)
def t(fn, n=5):
    fn(); fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
print("dense e2e", t(lambda: call(out=out, transfer="dense")))
print("compact e2e", t(lambda: call(out=out, transfer="compact")))
print("direct e2e", t(lambda: call(out=out)))
real = occ.expand_compact_host
occ.expand_compact_host = lambda *a, **k: None
print("direct, no host zero-fill", t(lambda: call(out=out)))
h_rank = None; descs = None
def grab(*a, **k):
    global h_rank, descs
    descs = a[0].copy(); h_rank = a[3].copy(); return real(*a, **k)
occ.expand_compact_host = grab
call(out=out)
occ.expand_compact_host = real
for nt in (4, 8, 16, 32):
    t0 = time.perf_counter()
    for _ in range(5): real(descs, 0, 256, h_rank, None, 0, out, n_threads=nt)
    print("zero-fill alone, threads", nt, (time.perf_counter() - t0) / 5 * 1e3)
print("h2d + kernels, device out", t(lambda: call(return_tensor=True)))
