"""One C4a distance step (10 000 frames x 256 x 1024 periodic pairs) for an ncu capture of dist_kernel."""
import sys
import numpy as np
import torch

sys.path.insert(0, "/root/repo")
from moleculekit_b200 import distance_utils as du

dev = torch.device("cuda:0")
F, N = 10000, 1280
g = torch.Generator(device=dev).manual_seed(7)
L = 36.84
c = torch.rand((N, 3, 1), generator=g, device=dev) * L + torch.cumsum(torch.randn((N, 3, F), generator=g, device=dev) * 0.3, dim=2)
box = torch.full((3, F), L, device=dev) * (1 + 0.002 * torch.randn((1, F), generator=g, device=dev))
s1 = torch.arange(0, 256, dtype=torch.int32, device=dev); s2 = torch.arange(256, 1280, dtype=torch.int32, device=dev)
ch = torch.zeros(N, dtype=torch.int32, device=dev); ch[256:] = 1
for metric in ("distances", "contacts"):
    for _ in range(2):
        out = du.dist_trajectory_device(c.contiguous(), box.contiguous(), s1, s2, ch, False, True, metric=metric, threshold=12.0)
torch.cuda.synchronize()
print("ok", out.shape)
