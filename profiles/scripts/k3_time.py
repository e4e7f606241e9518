import torch, numpy as np, sys
sys.path.insert(0, ".")
from moleculekit_b200 import distance_utils as du, _lib
dev = torch.device("cuda", 0)
rng = np.random.default_rng(7)
n1, n2, F = 256, 1024, 10000
L = 36.84
start = rng.uniform(0, L, size=(n1 + n2, 3, 1)).astype(np.float32)
coords = start + np.cumsum(rng.normal(0, 0.3, size=(n1 + n2, 3, F)).astype(np.float32), axis=2)
box = np.repeat((L * (1 + 0.002 * rng.normal(size=F))).astype(np.float32)[None, :], 3, axis=0)
d_c = torch.from_numpy(np.ascontiguousarray(coords)).to(dev); d_b = torch.from_numpy(np.ascontiguousarray(box)).to(dev)
s1 = torch.arange(0, n1, dtype=torch.int32, device=dev); s2 = torch.arange(n1, n1 + n2, dtype=torch.int32, device=dev)
ch = torch.ones(n1 + n2, dtype=torch.int32, device=dev); ch[n1:] = 2
_lib.set_timing(True, 0)
for metric in ("distances", "contacts"):
    o = torch.empty((F, n1 * n2), dtype=torch.float32 if metric == "distances" else torch.uint8, device=dev)
    for _ in range(3): du.dist_trajectory_device(d_c, d_b, s1, s2, ch, False, True, metric=metric, threshold=12.0, out=o)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): du.dist_trajectory_device(d_c, d_b, s1, s2, ch, False, True, metric=metric, threshold=12.0, out=o)
    e1.record(); torch.cuda.synchronize()
    print(metric, e0.elapsed_time(e1) / 5, _lib.get_timing(0))
