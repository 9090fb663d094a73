"""BASELINE.json configs[4] second half: one step of the ray-drop refinement loop (model/runner.py:868-909: random box
masking, U-Net forward, BCE against the GT ray-drop mask, Adam) on the rendered panoramas of a 50-frame sequence
[50,3,66,1030], timed with CUDA events in fp32 and under bf16 autocast.  Library convolutions (lidar4d_b200/raydrop_unet.py)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lidar4d_b200.raydrop_unet import RayDropUNet

dev = torch.device("cuda:0")
torch.manual_seed(0)
B, H, W = 50, 66, 1030
x = torch.rand(B, 3, H, W, device=dev)
gt = (torch.rand(B, 1, H, W, device=dev) > 0.2).float()
res = {}
for name, dt in (("fp32", None), ("bf16", torch.bfloat16)):
    net = RayDropUNet(3, 32, 1).to(dev).train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    bce = torch.nn.BCELoss()
    rng = np.random.default_rng(0)

    def step():
        opt.zero_grad()
        mask = torch.ones_like(x)
        for _ in range(rng.integers(32)):
            by, bx = rng.integers(1, int(0.1 * H)), rng.integers(1, int(0.1 * W))
            yi, xi = rng.integers(H - by), rng.integers(W - bx)
            mask[:, :, yi:yi + by, xi:xi + bx] = 0.0
        with torch.autocast("cuda", dtype=dt, enabled=dt is not None):
            out = net(x * mask)
        loss = bce(out.float(), gt)
        loss.backward()
        opt.step()
        return loss
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        l = step()
    e1.record()
    torch.cuda.synchronize()
    res[name] = {"ms_per_step": e0.elapsed_time(e1) / 5, "loss": float(l), "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
    print(f"[refine] {name}: {res[name]['ms_per_step']:.1f} ms per step of the {B}-frame batch, {res[name]['peak_mem_gb']:.1f} GB", flush=True)
    del net, opt
    torch.cuda.empty_cache()
print(json.dumps({"probe": "ray-drop refinement step (runner.py:868-909), RayDropUNet, batch [50,3,66,1030]", "results": res}))
