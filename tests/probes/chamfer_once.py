"""One chamfer forward at the flow-loss size (for ncu captures)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lidar4d_b200.chamfer import chamfer_3DDist
g = torch.Generator().manual_seed(0)
a = (torch.rand(1, 100000, 3, generator=g) * 2 - 1).cuda()
b = (torch.rand(1, 100000, 3, generator=g) * 2 - 1).cuda()
f = chamfer_3DDist()
for _ in range(2):
    d1, d2, _, _ = f(a, b)
torch.cuda.synchronize()
print(float(d1.mean()), float(d2.mean()))
