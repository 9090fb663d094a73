"""GPU diagnostic (not a test): per-tensor errors of the CUDA path against the oracle at the full-size configuration,
and a dump of torch.linspace on CUDA.  Writes gpurun_out/diag_fullsize.txt."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import lidar4d_oracle as O
from parity_util import rel_err, grad_errors, full_oracle, cuda_model_from_oracle
import test_gpu_fullsize_parity as T

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", "diag_fullsize.txt"), "w")
def P(*a):
    print(*a); print(*a, file=out); out.flush()
dev = torch.device("cuda:0")
lins = {}
for S in (768, 767, 1000, 129):
    a = torch.linspace(0.0, 1.0, S, device=dev).cpu().numpy(); b = O.sample_lin(S)
    lins[str(S)] = a
    bad = np.flatnonzero(a != b)
    P(f"linspace S={S}: {bad.size} differ; first {bad[:8]}; cuda {a[bad[:4]]} ours {b[bad[:4]]}")
np.savez(os.path.join(ROOT, "gpurun_out", "cuda_linspace.npz"), **lins)
names = sys.argv[1:] or T.FULL_CASES
for name in names:
    fx, orc, ref, og = T.case(name)
    for mode, pipe in T.MODES:
        m, o, got = T.cuda_run(fx, orc, dev, mode, pipe)
        P(f"== {name} {mode} {pipe}")
        for k in ("depth_lidar", "image_lidar", "weights_sum_lidar", "weights"):
            P(f"   out {k:20s} vs oracle {rel_err(o[k], ref[k]):.2e}   vs fixture {rel_err(o[k], fx['ref_' + k]) if 'ref_' + k in fx.files else -1:.2e}")
        P("   z_vals equal:", np.array_equal(o["z_vals"].cpu().numpy(), ref["z_vals"].numpy()))
        rows = []
        for k, g_ref in og.items():
            if not g_ref.numel() or got.get(k) is None:
                continue
            e = grad_errors(got[k], g_ref)
            rows.append((max(e[0], e[1]), k, e))
        for _, k, e in sorted(rows, reverse=True)[:10]:
            P(f"   grad {k:55s} max {e[0]:.2e} l2 {e[1]:.2e} mixed {e[2]:.2e}")
