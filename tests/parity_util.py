"""Shared helpers of the parity tests (test infrastructure)."""
import numpy as np
import torch

from lidar4d_b200.geometry import FieldConfig

SMALL = dict(min_resolution=8, base_resolution=16, max_resolution=512, time_resolution=8,
             n_levels_hash=4, log2_hashmap_size=12, num_frames=6, near_lidar=0.0105, far_lidar=0.851,
             hash_size_dynamic=(9, 8, 8), flow_base_resolution=8, flow_max_resolution=256,
             flow_log2_hashmap_size=10)


def small_config(**over) -> FieldConfig:
    kw = dict(SMALL)
    kw.update(over)
    return FieldConfig(**kw)


def rel_err(a, b) -> float:
    """max |a-b| / max |b| (the 1e-4 'rel fp32' bar of BASELINE.json:north_star)."""
    a = torch.as_tensor(a).detach().double().reshape(-1).cpu()
    b = torch.as_tensor(b).detach().double().reshape(-1).cpu()
    if a.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def make_surface_like(oracle, gain: float = 1.5):
    """Sharpen the density so that most samples fall below the w>1e-4 attribute
    mask (SURVEY.md 8(d) 'surface-like' regime): make the sigma row of the output
    layer positive (hidden units are >= 0) and scale it."""
    c = oracle.cfg
    with torch.no_grad():
        p = oracle.p("sigma_net.params")
        off = 64 * c.sigma_in_pad
        p[off:off + 64] = p[off:off + 64].abs() * gain
    return oracle


def cuda_model_from_oracle(oracle, device="cuda"):
    """CUDA LiDAR4D with the oracle's parameters (state_dict round trip)."""
    from lidar4d_b200 import LiDAR4D
    c = oracle.cfg
    m = LiDAR4D(min_resolution=c.min_resolution, base_resolution=c.base_resolution, max_resolution=c.max_resolution,
                time_resolution=c.time_resolution, n_levels_plane=c.n_levels_plane, n_levels_hash=c.n_levels_hash,
                log2_hashmap_size=c.log2_hashmap_size, num_frames=c.num_frames, bound=c.bound,
                near_lidar=c.near_lidar, far_lidar=c.far_lidar, density_scale=c.density_scale,
                active_sensor=c.active_sensor, hash_size_dynamic=c.hash_size_dynamic,
                flow_base_resolution=c.flow_base_resolution, flow_max_resolution=c.flow_max_resolution,
                flow_log2_hashmap_size=c.flow_log2_hashmap_size)
    res = m.load_state_dict(oracle.ref_state_dict(), strict=False)
    assert not res.missing_keys and not res.unexpected_keys, res
    # same weight precision as the oracle instance: fp32 weights unless the oracle emulates tcnn's fp16 copies
    m.set_mlp_fp16(getattr(oracle, "mlp_dtype", "fp32") == "fp16")
    return m.to(device)


def test_rays(H=3, W=8, origin=(0.1, -0.05, 0.02)):
    from lidar4d_b200.rays import lidar_rays
    pose = np.eye(4, dtype=np.float32)
    pose[:3, 3] = origin
    return lidar_rays(pose, H, W)


def relu_margin(oracle, stages) -> float:
    """Smallest |pre-activation| of the ReLU layers that the oracle exposes (flow
    MLP, sigma MLP).  Gradients are discontinuous where a pre-activation crosses
    zero, so two correct fp32 implementations can legitimately differ when this
    margin is at rounding-noise level; parity cases assert a margin first."""
    c = oracle.cfg
    with torch.no_grad():
        fin = stages["flow_in"].float()
        y1 = fin @ oracle.p("flow_net.mlp.0.weight").float().t()
        y2 = torch.relu(y1) @ oracle.p("flow_net.mlp.2.weight").float().t()
        f = stages["features"].float()
        pad = torch.ones(f.shape[0], c.sigma_in_pad - c.sigma_in_dim)
        W1 = oracle.p("sigma_net.params")[:64 * c.sigma_in_pad].float().view(64, c.sigma_in_pad)
        ys = torch.cat([f, pad], -1) @ W1.t()
        return float(min(y1.abs().min(), y2.abs().min(), ys.abs().min()))


# ---- the benchmarked (full-size) configuration ---------------------------------------------------------------------
FULL = dict(num_frames=51, near_lidar=0.0105, far_lidar=0.851)


def full_config(levels: int) -> FieldConfig:
    """BASELINE.json configs[1]: the reference's default tables (2^19 static, 2^15/2^13/2^13 x 8 slices, 2^18 flow)."""
    return FieldConfig(n_levels_hash=levels, **FULL)


def full_oracle(levels: int, seed: int, smooth: bool = False):
    """Seeded full-size oracle whose MLP masters are fp16-representable (see O.snap_mlp_weights_fp16): the fp32-FMA and
    the tensor-core kernels must both reproduce it, and tests/golden/ref_full_*.npz hold the unmodified reference's
    results for exactly these parameters."""
    from oracle import lidar4d_oracle as O
    orc = O.build_seeded(full_config(levels), seed, flow_last_std=0.02)
    O.snap_mlp_weights_fp16(orc)
    return O.band_limit_tables(orc) if smooth else orc


def grad_errors(a, b, tol: float = 1e-4):
    """(max|a-b| / max|b|,  ||a-b|| / ||b||,  fraction of entries with |a-b| > tol * max|b|)."""
    a = torch.as_tensor(a).detach().double().reshape(-1).cpu()
    b = torch.as_tensor(b).detach().double().reshape(-1).cpu()
    if a.numel() == 0:
        return 0.0, 0.0, 0.0
    d = (a - b).abs()
    bmax = float(b.abs().max()) + 1e-30
    return float(d.max() / bmax), float(d.norm() / (b.norm() + 1e-30)), float((d > tol * bmax).double().mean())


def relu_margin_all(oracle, stages, rays_d) -> float:
    """relu_margin() plus the four hidden layers of the two attribute heads (lidar4d.py:191-223) on the samples inside
    the attribute mask: the smallest |pre-activation| of EVERY ReLU of the network for this input."""
    from oracle import lidar4d_oracle as O
    c = oracle.cfg
    m = relu_margin(oracle, stages)
    with torch.no_grad():
        mask = stages["mask"].reshape(-1)
        if bool(mask.any()):
            N, S = stages["weights"].shape
            d = torch.as_tensor(rays_d).float().view(N, 1, 3).expand(N, S, 3).reshape(-1, 3)[mask]
            enc = O.frequency_encode((d + 1) / 2, c.view_degree)
            inp = torch.cat([enc, stages["geo_feat"].float()[mask]], -1)
            inp = torch.cat([inp, torch.ones(inp.shape[0], c.attr_in_pad - c.attr_in_dim)], -1)
            for net in ("intensity_net.params", "raydrop_net.params"):
                W1, W2, _ = O.mlp_layers(oracle.p(net).float(), c.attr_in_pad, 64, 2)
                y1 = inp @ W1.t()
                y2 = torch.relu(y1) @ W2.t()
                m = min(m, float(y1.abs().min()), float(y2.abs().min()))
    return m


def fill_state_dict(module, scale: float = 0.2):
    """Deterministic, storage-free parameters for fixtures of library modules: every tensor is a smooth function of its
    flat index and of its key (positive for variances), identical for any implementation with the same state_dict."""
    import zlib
    with torch.no_grad():
        for k, v in module.state_dict().items():
            if not v.dtype.is_floating_point:
                continue
            ph = (zlib.crc32(k.encode()) % 1000) * 0.01
            i = torch.arange(v.numel(), dtype=torch.float64)
            x = torch.cos(i * 0.7548776662 + ph) * scale
            if v.dim() >= 2:                                    # conv / linear weights: keep activations O(1) through deep stacks
                x = x * (1.4 / scale) / float(v[0].numel()) ** 0.5
            if k.endswith("running_var"):
                x = x.abs() + 0.5
            elif k.endswith(".weight") and v.dim() == 1:      # norm scales around 1
                x = 1.0 + x
            v.copy_(x.view(v.shape).to(v.dtype))
    return module
