"""K9 (wrap_box) on the bench's C6 workload, a few calls -- the command profiled by ncu for profiles/r01_k9_*."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from moleculekit_b200 import wrapping as wr

dev = torch.device("cuda:0")
n_prot, n_wat, F = 5000, 18000, 512
N = n_prot + 3 * n_wat
g = torch.Generator(device=dev).manual_seed(5)
L = 82.0
d_orig = torch.empty((N, 3, F), dtype=torch.float32, device=dev)
d_orig[:n_prot] = torch.randn((n_prot, 3, 1), generator=g, device=dev) * 9 + torch.randn((1, 3, F), generator=g, device=dev) * 25
wat_c = (torch.rand((n_wat, 1, 3, F), generator=g, device=dev) - 0.5) * (5 * L)
d_orig[n_prot:] = (wat_c + torch.randn((n_wat, 3, 3, 1), generator=g, device=dev) * 0.6).reshape(3 * n_wat, 3, F)
d_b = torch.full((3, F), L, dtype=torch.float32, device=dev)
groups = torch.cat([torch.zeros(1, dtype=torch.int32, device=dev), n_prot + 3 * torch.arange(n_wat + 1, dtype=torch.int32, device=dev)])
csel = torch.arange(0, n_prot, dtype=torch.int32, device=dev)
d_c = torch.empty_like(d_orig)
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    d_c.copy_(d_orig)
    wr.wrap_box_device(d_c, d_b, groups, csel)
torch.cuda.synchronize()
print("ok")
