"""GPU diagnostic: forward / backward launch-group durations (events around the C calls, no per-kernel marks) for the two
bench losses - does the step time depend on the VALUES of the upstream gradient?"""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from lidar4d_b200 import LiDAR4D
from lidar4d_b200.rays import synthetic_sweep
from lidar4d_b200.losses import lidar_main_loss
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = LiDAR4D(**bench.model_kwargs(16)).to(dev)
bench.randomize(m, 0)
m.materialize_weights = False
ro, rd, t = synthetic_sweep(3)
ro_d, rd_d = torch.from_numpy(ro[:16384]).to(dev), torch.from_numpy(rd[:16384]).to(dev)
gt = torch.tensor([1.0, 0.5, 0.3], device=dev).repeat(16384, 1)
def run(kind, scale_img):
    m._engine.timing = {"fwd": [], "bwd": []}
    for it in range(6):
        m.zero_grad(set_to_none=True)
        out = m.render(ro_d[None], rd_d[None], float(t), num_steps=768, perturb=True)
        if kind == "unit":
            loss = ((out["depth_lidar"] - 0.3).abs().sum() + scale_img * ((out["image_lidar"] - 0.5) ** 2).sum()) / 65536
        else:
            loss = lidar_main_loss(out["depth_lidar"], out["image_lidar"], gt[None], 1.0, 0.01, 0.1, 0.2) * (1.0 / 65536)
        loss.backward()
    torch.cuda.synchronize()
    f = [a.elapsed_time(b) for a, b in m._engine.timing["fwd"]][2:]
    b = [a.elapsed_time(b) for a, b in m._engine.timing["bwd"]][2:]
    print(f"{kind:5s} img x{scale_img:<6g} fwd {np.mean(f):7.2f} ms  bwd {np.mean(b):7.2f} ms  (16,384 rays)", flush=True)
for _ in range(2):
    run("unit", 1.0); run("main", 1.0); run("unit", 0.01); run("unit", 100.0); run("unit", 0.0)
