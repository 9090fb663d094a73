"""GPU diagnostic: where does the tensor-core backward lose accuracy at full size?  Variants of the upstream gradient."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import lidar4d_oracle as O
from parity_util import rel_err, grad_errors, full_oracle, cuda_model_from_oracle
out = open(os.path.join(ROOT, "gpurun_out", "diag_tc.txt"), "w")
def P(*a):
    print(*a); print(*a, file=out); out.flush()
dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "ref_full_L16_interior"
fx = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
orc = full_oracle(int(fx["levels"]), int(fx["seed"]))
S = int(fx["num_steps"])
if os.environ.get("SMOOTH"):
    # band-limited tables: level amplitude ~ base_res / res_l (a trained grid looks like this; white noise at 32769 cells
    # turns one ulp of a warped coordinate into a 1e-3 feature change and with it ReLU flips between fp32 implementations)
    with torch.no_grad():
        for k, v in orc.P.items():
            geo = orc.g_static if "hash_static" in k else (orc.g_flow if "grid_enc" in k else (orc.g_dynamic[int(k.split("/")[2])] if "hash_dynamic" in k else None))
            if geo is None: continue
            t = v.view(-1, geo.n_features)
            for l in range(geo.n_levels):
                t[int(geo.offset[l]):int(geo.offset[l + 1])] *= float(geo.resolution[0]) / float(geo.resolution[l])
    P("SMOOTH tables")
ro, rd = torch.from_numpy(fx["rays_o"]), torch.from_numpy(fx["rays_d"])
gd0, gi0 = torch.from_numpy(fx["g_depth"]), torch.from_numpy(fx["g_image"])
KEYS = ["sigma_net.params", "intensity_net.params", "raydrop_net.params", "hash_encoder.hash_static.params",
        "planes_encoder.planes.3.0", "planes_encoder.planes.0.2", "flow_net.grid_enc.params", "flow_net.mlp.0.weight", "flow_net.mlp.4.weight"]
for label, sd, si in (("depth only", 1.0, 0.0), ("image only", 0.0, 1.0), ("both", 1.0, 1.0)):
    orc.zero_grad(set_to_none=True)
    ref = orc.render(ro, rd, float(fx["time"]), num_steps=S, perturb=bool(fx["perturb"]), seed=int(fx["seed"]), return_stages=True)
    ((ref["depth_lidar"] * gd0 * sd).sum() + (ref["image_lidar"] * gi0 * si).sum()).backward()
    og = orc.ref_named_grads()
    w = ref["weights"].detach()
    P(f"#### {label}: mask fraction {float((w > 1e-4).float().mean()):.3f}, min |w-1e-4| {float((w - 1e-4).abs().min()):.2e}, sigma range {float(ref['sigma'].min()):.2e}..{float(ref['sigma'].max()):.2e}")
    for mode in ("tc", "fp32"):
        m = cuda_model_from_oracle(orc).set_mlp_fp16(mode == "tc")
        m.jitter_seed = int(fx["seed"])
        o = m.render(ro[None].to(dev), rd[None].to(dev), torch.tensor([[float(fx["time"])]]), num_steps=S, perturb=bool(fx["perturb"]))
        ((o["depth_lidar"][0] * gd0.to(dev) * sd).sum() + (o["image_lidar"][0] * gi0.to(dev) * si).sum()).backward()
        torch.cuda.synchronize()
        got = {k: p.grad for k, p in m.named_parameters()}
        nflip = int(((o["weights"].cpu() > 1e-4) != (w > 1e-4)).sum())
        P(f"  -- {mode}: mask flips vs oracle {nflip}; weights err {rel_err(o['weights'], w):.2e}")
        for k in KEYS:
            if got.get(k) is None: continue
            e = grad_errors(got[k], og[k])
            P(f"     {k:45s} max {e[0]:.2e} l2 {e[1]:.2e}")
