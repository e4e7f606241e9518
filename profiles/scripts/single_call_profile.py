import sys, time, cProfile, pstats, io
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from moleculekit_b200 import workloads
from moleculekit_b200.tools import voxeldescriptors as vd
from moleculekit_b200.molecule_lite import MolLite
w1 = workloads.protein_pockets(B=1, n_atoms=1639, box=60.0, radius=17.0, seed=5)
mol1 = MolLite(w1["coords"][0])
for _ in range(5): vd.getVoxelDescriptors(mol1, userchannels=w1["sigmas"][0], buffer=8, voxelsize=1)
torch.cuda.synchronize()
t0=time.perf_counter()
for _ in range(50): r=vd.getVoxelDescriptors(mol1, userchannels=w1["sigmas"][0], buffer=8, voxelsize=1)
print("ms per call", (time.perf_counter()-t0)/50*1e3, r[0].shape, r[0].dtype)
pr=cProfile.Profile(); pr.enable()
for _ in range(50): vd.getVoxelDescriptors(mol1, userchannels=w1["sigmas"][0], buffer=8, voxelsize=1)
pr.disable(); s=io.StringIO(); pstats.Stats(pr,stream=s).sort_stats("cumulative").print_stats(22); print(s.getvalue()[:3800])
