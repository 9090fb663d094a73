"""ray generation (+ GT gather) and the main-loss epilogue, one CUDA launch each, against
  * tests/golden/lidar_rays.npz - the reference's own get_lidar_rays (data/base_dataset.py:15-102) run by make_golden.py;
  * the numpy sensor model of lidar4d_b200/rays.py (CPU, pinned to the same fixture);
  * a torch restatement of runner.py:179-213 (the expression tests/trainer_mirror.py shares with the real Trainer)."""
import os

import numpy as np
import pytest
import torch

from lidar4d_b200.rays import lidar_rays

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_numpy_sensor_model_matches_reference_get_lidar_rays():
    fx = np.load(os.path.join(GOLD, "lidar_rays.npz"))
    ro, rd = lidar_rays(fx["pose"], int(fx["H"]), int(fx["W"]), float(fx["fov_up"]), float(fx["fov"]))
    assert np.abs(ro - fx["rays_o"]).max() == 0.0
    assert np.abs(rd - fx["rays_d"]).max() < 2e-6
    assert np.array_equal(fx["inds"], np.arange(int(fx["H"]) * int(fx["W"])))


@pytest.mark.gpu
def test_get_lidar_rays_kernel():
    from lidar4d_b200.rays import get_lidar_rays
    dev = torch.device("cuda:0")
    fx = np.load(os.path.join(GOLD, "lidar_rays.npz"))
    H, W = int(fx["H"]), int(fx["W"])
    poses = torch.from_numpy(fx["pose"])[None].to(dev)
    r = get_lidar_rays(poses, [float(fx["fov_up"]), float(fx["fov"])], H, W, -1)
    assert r["rays_o"].shape == (1, H * W, 3) and r["inds"].shape == (1, H * W)
    assert torch.equal(r["rays_o"][0].cpu(), torch.from_numpy(fx["rays_o"]))
    assert float((r["rays_d"][0].cpu() - torch.from_numpy(fx["rays_d"])).abs().max()) < 2e-6
    # random pixels, patches and the GT gather: same torch.randint stream as the reference's code on this device
    g = torch.Generator().manual_seed(0)
    images = torch.rand(1, H, W, 3, generator=g).to(dev)
    for N, patch in ((37, 0), (32, 2), (40, [2, 4])):
        torch.manual_seed(11)
        r = get_lidar_rays(poses, [2.0, 26.9], H, W, N, patch_size=patch, images=images)
        torch.manual_seed(11)
        if (patch if isinstance(patch, int) else patch[0]) > 0:
            px, py = (patch, patch) if isinstance(patch, int) else patch
            n_patch = N // (px * py)
            ix = torch.randint(0, H - px, size=[n_patch], device=dev)
            iy = torch.randint(0, W, size=[n_patch], device=dev)
            pi, pj = torch.meshgrid(torch.arange(px, device=dev), torch.arange(py, device=dev), indexing="ij")
            ind2 = (torch.stack([ix, iy], -1).unsqueeze(1) + torch.stack([pi.reshape(-1), pj.reshape(-1)], -1).unsqueeze(0)).view(-1, 2)
            ind2[:, 1] = ind2[:, 1] % W
            inds = ind2[:, 0] * W + ind2[:, 1]
        else:
            inds = torch.randint(0, H * W, size=[N], device=dev)
        assert torch.equal(r["inds"][0], inds)
        # the same arithmetic as a torch op chain on this GPU (base_dataset.py:82-97)
        i, j = (inds % W).float(), (inds // W).float()
        beta = -(i - W / 2) / W * 2 * np.pi
        alpha = (2.0 - j / H * 26.9) / 180 * np.pi
        d = torch.stack([torch.cos(alpha) * torch.cos(beta), torch.cos(alpha) * torch.sin(beta), torch.sin(alpha)], -1)
        rd = d @ poses[0, :3, :3].t()
        assert float((r["rays_d"][0] - rd).abs().max()) < 1e-6
        assert torch.equal(r["gt"][0], images.view(1, H * W, 3)[0][inds])


@pytest.mark.gpu
def test_main_loss_kernel_value_and_gradient():
    from lidar4d_b200.losses import lidar_main_loss
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    N = 5000
    depth = torch.rand(1, N, generator=g).to(dev).requires_grad_(True)
    image = torch.rand(1, N, 2, generator=g).to(dev).requires_grad_(True)
    gt = torch.rand(1, N, 3, generator=g)
    gt[..., 0] = (gt[..., 0] > 0.2).float()
    gt = gt.to(dev)
    a_d, a_r, a_i, s = 1.0, 0.01, 0.1, 0.2
    # runner.py:179-213
    m = gt[:, :, 0]
    ref = (a_d * (depth * m - gt[:, :, 2] * m).abs() + a_r * (image[:, :, 0] - m.clamp(s, 1 - s)) ** 2
           + a_i * (image[:, :, 1] * m - gt[:, :, 1] * m) ** 2).sum()
    gd_ref, gi_ref = torch.autograd.grad(ref * 3.0, (depth, image))
    loss = lidar_main_loss(depth, image, gt, a_d, a_r, a_i, s)
    gd, gi = torch.autograd.grad(loss * 3.0, (depth, image))
    assert float(loss) == pytest.approx(float(ref), rel=1e-5)
    assert torch.allclose(gd, gd_ref, rtol=1e-6, atol=1e-7) and torch.allclose(gi, gi_ref, rtol=1e-6, atol=1e-7)
    with pytest.raises(RuntimeError):
        lidar_main_loss(depth.cpu(), image.cpu(), gt.cpu())
