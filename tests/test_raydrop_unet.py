"""the ray-drop refinement U-Net (library convolutions) against the reference's own module
(tests/golden/unet.npz, written by tests/golden/make_golden.py from /root/reference/model/unet.py)."""
import os

import numpy as np
import pytest
import torch

from parity_util import fill_state_dict, rel_err

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _net():
    from lidar4d_b200.raydrop_unet import RayDropUNet
    return fill_state_dict(RayDropUNet(in_channels=3, out_channels=1)).eval()


def test_state_dict_surface_and_function_match_reference_unet():
    fx = np.load(os.path.join(GOLD, "unet.npz"))
    net = _net()
    sd = net.state_dict()
    assert sorted(sd.keys()) == list(fx["keys"])
    assert [str(tuple(sd[k].shape)) for k in sorted(sd.keys())] == list(fx["shapes"])
    with torch.no_grad():
        y = net(torch.from_numpy(fx["x"]))
    assert y.shape == fx["y"].shape
    assert rel_err(y, fx["y"]) < 1e-5


def test_lidar4d_accepts_it_and_keeps_the_unet_keys():
    from lidar4d_b200 import LiDAR4D
    from lidar4d_b200.raydrop_unet import RayDropUNet
    m = LiDAR4D(min_resolution=8, base_resolution=16, max_resolution=64, n_levels_hash=2, log2_hashmap_size=8,
                hash_size_dynamic=(6, 6, 6), flow_base_resolution=8, flow_max_resolution=32, flow_log2_hashmap_size=8,
                unet=RayDropUNet(3, 32, 1))
    keys = [k for k in m.state_dict() if k.startswith("unet.")]
    assert "unet.inc.conv.weight" in keys and "unet.attn.proj_qkv.weight" in keys and "unet.outc.conv.2.bias" in keys
    assert not any(k.startswith("unet") for g in m.get_params(1e-2) for k in ())       # Adam groups exclude it (lidar4d.py:226-237)
    ids = {id(p) for g in m.get_params(1e-2) for p in g["params"]}
    assert not any(id(p) in ids for p in m.unet.parameters())


@pytest.mark.gpu
def test_unet_gpu_fp32_and_bf16():
    dev = torch.device("cuda:0")
    fx = np.load(os.path.join(GOLD, "unet.npz"))
    net = _net().to(dev)
    x = torch.from_numpy(fx["x"]).to(dev)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False            # the fp32 comparison is an fp32 comparison (cuDNN defaults to TF32 convs)
    try:
        with torch.no_grad():
            y = net(x)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                yb = net(x)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    e32, e16 = rel_err(y, fx["y"]), rel_err(yb.float(), fx["y"])
    assert e32 < 1e-4, e32
    assert torch.isfinite(yb).all() and e16 < 5e-2, e16      # bf16 through 23 layers
    # training mode: attention dropout mask + BN batch statistics run and give gradients to every parameter
    net.train()
    out = net(torch.rand(2, 3, 66, 130, device=dev))
    out.mean().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())
