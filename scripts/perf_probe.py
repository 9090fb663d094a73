"""Quick GPU timing of the fused kernels (CUDA events), flushed line by line."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from lidar4d_b200 import LiDAR4D
from lidar4d_b200.rays import synthetic_sweep

dev = torch.device("cuda:0")
levels = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [8, 16]
sizes = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1024, 4096]
for L in levels:
    t0 = time.time()
    torch.manual_seed(0)
    m = LiDAR4D(**bench.model_kwargs(L)).to(dev)
    bench.randomize(m, 0)
    m.materialize_weights = False
    print(f"[probe] L={L} model built in {time.time()-t0:.1f}s", flush=True)
    ro, rd, t = synthetic_sweep(7)
    eng = m._engine
    for N in sizes:
        sel = np.linspace(0, 65535, N).astype(np.int64)
        ro_t, rd_t = torch.from_numpy(ro[sel])[None].to(dev), torch.from_numpy(rd[sel])[None].to(dev)
        for pipe, mode in (("fused", "train"), ("split", "train"), ("split-tc", "infer"), ("split-tc", "train")):
            m.pipeline = pipe.split("-")[0]
            m.set_mlp_fp16(pipe.endswith("tc"))
            eng.timing = {"fwd": [], "bwd": []}
            for it in range(3):
                if mode == "infer":
                    with torch.no_grad():
                        out = m.render(ro_t, rd_t, float(t), num_steps=768, perturb=False)
                else:
                    m.zero_grad(set_to_none=True)
                    out = m.render(ro_t, rd_t, float(t), num_steps=768, perturb=True)
                    (out["depth_lidar"].sum() + out["image_lidar"].sum()).backward()
                torch.cuda.synchronize()
            f = [a.elapsed_time(b) for a, b in eng.timing["fwd"]][1:]
            b = [a.elapsed_time(b_) for a, b_ in eng.timing["bwd"]][1:]
            eng.timing = None
            fb, bb = bench.algorithmic_bytes(m.cfg)
            msg = f"[probe] L={L} N={N} {pipe:8s} {mode}: fwd {np.mean(f):8.2f} ms ({N/np.mean(f)*1e3:9.0f} rays/s, {fb*N*768/np.mean(f)/1e6:7.0f} GB/s alg)"
            if b:
                msg += f" | bwd {np.mean(b):8.2f} ms ({N/np.mean(b)*1e3:9.0f} rays/s, {bb*N*768/np.mean(b)/1e6:7.0f} GB/s alg) | fwd+bwd {N/(np.mean(f)+np.mean(b))*1e3:9.0f} rays/s"
            print(msg, flush=True)
    del m
    torch.cuda.empty_cache()
