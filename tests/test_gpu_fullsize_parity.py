"""GPU parity at the BENCHMARKED configuration (BASELINE.json configs[1]): default tables (2^19 static, 2^15/2^13/2^13
x 8 time slices, 2^18 flow), n_levels_hash 8 and 16 (sigma_in_dim 120 / 176), S = 768, jitter on.

Two anchors:
  * tests/golden/ref_full_*.npz - the UNMODIFIED reference modules on the tcnn shim (tests/golden/make_golden.py);
  * the oracle, run here on the host cores with the same seeded parameters, for EVERY gradient tensor in full.
Both CUDA modes (tcgen05 tensor-core kernels with fp16 working weights; fp32-FMA kernels) and both pipelines must hold
1e-4 (rel-to-max, fp32) on outputs and gradients; hash indices are bit-exact.  Also: GradScaler-sized upstream gradients
(x65536, runner.py:506-508) and an outer fp16 autocast (runner.py:497)."""
import os

import numpy as np
import pytest
import torch

from oracle import lidar4d_oracle as O
from parity_util import rel_err, grad_errors, full_oracle, cuda_model_from_oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-4
FULL_CASES = ["ref_full_L16_interior", "ref_full_L8_interior", "ref_full_L16_first", "ref_full_L8_last"]
MODES = [("tc", "split"), ("fp32", "split"), ("tc", "fused")]


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


_cache = {}


def case(name):
    """(fixture, oracle with gradients of the fixture's loss) - computed once per module on the host cores."""
    if name not in _cache:
        fx = np.load(os.path.join(GOLD, name + ".npz"))
        orc = full_oracle(int(fx["levels"]), int(fx["seed"]))
        S = int(fx["num_steps"])
        ref = orc.render(torch.from_numpy(fx["rays_o"]), torch.from_numpy(fx["rays_d"]), float(fx["time"]), num_steps=S,
                         perturb=bool(fx["perturb"]), seed=int(fx["seed"]))
        ((ref["depth_lidar"] * torch.from_numpy(fx["g_depth"])).sum() +
         (ref["image_lidar"] * torch.from_numpy(fx["g_image"])).sum()).backward()
        _cache.clear()                 # one full-size oracle resident at a time
        _cache[name] = (fx, orc, {k: v.detach() for k, v in ref.items()}, orc.ref_named_grads())
    return _cache[name]


def cuda_run(fx, orc, dev, mode, pipeline, scale=1.0, autocast=False):
    m = cuda_model_from_oracle(orc).set_mlp_fp16(mode == "tc")
    m.pipeline = pipeline
    m.jitter_seed = int(fx["seed"])
    ro, rd = torch.from_numpy(fx["rays_o"])[None].to(dev), torch.from_numpy(fx["rays_d"])[None].to(dev)
    gd, gi = torch.from_numpy(fx["g_depth"]).to(dev), torch.from_numpy(fx["g_image"]).to(dev)
    with torch.autocast("cuda", dtype=torch.float16, enabled=autocast):
        out = m.render(ro, rd, torch.tensor([[float(fx["time"])]]), num_steps=int(fx["num_steps"]),
                       perturb=bool(fx["perturb"]))
        loss = (out["depth_lidar"][0] * gd).sum() + (out["image_lidar"][0] * gi).sum()
    (loss * scale).backward()
    torch.cuda.synchronize()
    return m, out, {k: p.grad for k, p in m.named_parameters()}


@pytest.mark.parametrize("S", [768, 767, 129, 2, 1000])
def test_z_grid_matches_cuda_linspace(dev, S):
    """The kernel's / oracle's restatement of torch.linspace(0,1,S) on CUDA (renderer.py:69; the reference renders on
    the GPU) is bit-exact - the full-size goldens were generated on this grid."""
    assert np.array_equal(torch.linspace(0.0, 1.0, S, device=dev).cpu().numpy(), O.sample_lin(S))


def test_hash_indices_bit_exact_L16(dev):
    from lidar4d_b200 import LiDAR4D
    fx = np.load(os.path.join(GOLD, "hash_indices.npz"))
    m = LiDAR4D(n_levels_hash=16).to(dev)
    for gid, name in [(0, "static3d_L16"), (1, "dyn2d_xy_L16"), (3, "dyn2d_yz_L16")]:
        x = torch.from_numpy(fx[name + ":x"]).to(dev)
        for l in range(16):
            idx, w = m.hash_indices(gid, l, x)
            assert np.array_equal(idx.cpu().numpy().view(np.uint32), fx[f"{name}:idx{l}"]), (name, l)
            assert np.array_equal(w.cpu().numpy(), fx[f"{name}:w{l}"]), (name, l)


@pytest.mark.parametrize("mode,pipeline", MODES)
@pytest.mark.parametrize("name", FULL_CASES)
def test_full_size_vs_reference_golden_and_oracle(dev, name, mode, pipeline):
    fx, orc, ref, og = case(name)
    m, out, got = cuda_run(fx, orc, dev, mode, pipeline)
    # ---- forward: the reference's own numbers, then the oracle's (which has the kernel's z grid) ----
    assert rel_err(out["depth_lidar"], fx["ref_depth_lidar"]) < TOL
    assert rel_err(out["image_lidar"], fx["ref_image_lidar"]) < TOL
    assert rel_err(out["weights"], fx["ref_weights"]) < TOL
    assert np.array_equal(out["z_vals"].cpu().numpy(), ref["z_vals"].numpy())
    for k in ("depth_lidar", "image_lidar", "weights_sum_lidar", "weights"):
        assert rel_err(out[k], ref[k]) < TOL, k
    # ---- gradients against the reference fixture: full small tensors, norm / projection / sampled entries of tables ----
    for k in [k[5:] for k in fx.files if k.startswith("grad:")]:
        if fx["grad:" + k].size:
            assert rel_err(got[k], fx["grad:" + k]) < TOL, k
    for k in [k[8:] for k in fx.files if k.startswith("gradidx:")]:
        g = got[k].reshape(-1)
        n_ref = float(fx["gradnorm:" + k])
        assert abs(float(g.double().norm()) - n_ref) <= TOL * n_ref + 1e-12, k
        idx = torch.from_numpy(fx["gradidx:" + k].astype(np.int64)).to(dev)
        refv = fx["gradval:" + k]
        gv = g[idx].cpu().numpy()
        assert np.abs(gv - refv).max() <= TOL * np.abs(refv).max(), k
        assert np.all(gv[refv == 0] == 0), k                  # untouched entries stay untouched
        proj = torch.cos(torch.arange(g.numel(), device=dev, dtype=torch.float64) * 0.6180339887498949 + 0.25)
        p2 = float(fx["gradproj2:" + k])
        assert abs(float(g.double() @ proj) - p2) <= TOL * (abs(p2) + n_ref), k
    # ---- gradients against the oracle: EVERY tensor in full, three norms ----
    worst = {}
    for k, g_ref in og.items():
        if not g_ref.numel():
            continue
        if got.get(k) is None:
            assert float(g_ref.abs().max()) == 0.0, k
            continue
        e_max, e_l2, e_mix = grad_errors(got[k], g_ref)
        worst[k] = (e_max, e_l2, e_mix)
        # the support must agree exactly: an entry is touched by the kernel iff the oracle touches it
        if g_ref.numel() > 70000:
            assert torch.equal((got[k].cpu().reshape(-1) != 0), (g_ref.reshape(-1) != 0)) or e_max < 1e-6, k
    bad = {k: v for k, v in worst.items() if not (v[0] < TOL and v[1] < TOL and v[2] <= 0.0)}
    assert not bad, f"failing (max-rel, l2-rel, mixed excess): {bad}"


@pytest.mark.parametrize("mode", ["tc", "fp32"])
def test_loss_scaled_upstream_gradients(dev, mode):
    """GradScaler multiplies the loss by 65536 (torch default init_scale; runner.py:102,506-508): the gradients must
    scale exactly with it - no overflow in the fp16 delta tiles of the tensor-core backward, no flush of small ones."""
    fx, orc, ref, og = case("ref_full_L16_interior")
    m, out, got = cuda_run(fx, orc, dev, mode, "split", scale=65536.0)
    for k, g_ref in og.items():
        if not g_ref.numel() or got.get(k) is None:
            continue
        assert torch.isfinite(got[k]).all(), k
        e_max, e_l2, _ = grad_errors(got[k] / 65536.0, g_ref)
        assert e_max < TOL and e_l2 < TOL, (k, e_max, e_l2)
    # and a tiny scale (late-training loss magnitudes)
    m2, out2, got2 = cuda_run(fx, orc, dev, mode, "split", scale=2.0 ** -20)
    for k in ("sigma_net.params", "hash_encoder.hash_static.params", "flow_net.grid_enc.params", "planes_encoder.planes.3.0"):
        e_max, e_l2, _ = grad_errors(got2[k] * 2.0 ** 20, og[k])
        assert e_max < TOL and e_l2 < TOL, (k, e_max, e_l2)


def test_outer_autocast_and_grad_scaler_step(dev):
    """Trainer.train_step runs the model under torch.cuda.amp.autocast and steps through a GradScaler
    (runner.py:497,506-511): outputs stay fp32 and equal to the plain call, scaled grads unscale to the oracle's."""
    fx, orc, ref, og = case("ref_full_L16_interior")
    m, out, got = cuda_run(fx, orc, dev, "tc", "split", autocast=True)
    assert out["depth_lidar"].dtype == torch.float32 and out["image_lidar"].dtype == torch.float32
    for k in ("depth_lidar", "image_lidar"):
        assert rel_err(out[k], ref[k]) < TOL, k
    for k in ("sigma_net.params", "intensity_net.params", "hash_encoder.hash_static.params", "planes_encoder.planes.0.3"):
        e_max, e_l2, _ = grad_errors(got[k], og[k])
        assert e_max < TOL and e_l2 < TOL, (k, e_max, e_l2)
    # one real scaler step
    m.zero_grad(set_to_none=True)
    opt = torch.optim.Adam(m.get_params(1e-2), betas=(0.9, 0.99), eps=1e-15)
    scaler = torch.amp.GradScaler("cuda")
    before = m.sigma_net.params.detach().clone()
    ro, rd = torch.from_numpy(fx["rays_o"])[None].to(dev), torch.from_numpy(fx["rays_d"])[None].to(dev)
    with torch.autocast("cuda", dtype=torch.float16):
        o = m.render(ro, rd, torch.tensor([[float(fx["time"])]]), num_steps=768, perturb=True)
        loss = (o["depth_lidar"] - 0.3).abs().mean() + ((o["image_lidar"] - 0.5) ** 2).mean()
    scaler.scale(loss).backward()
    scaler.step(opt)
    scaler.update()
    assert scaler.get_scale() == 65536.0                       # no inf/nan was found in the scaled gradients
    assert not torch.equal(before, m.sigma_net.params.detach())
