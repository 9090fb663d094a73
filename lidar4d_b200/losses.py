"""Main LiDAR loss of the reference's train_step (model/runner.py:179-213) as ONE CUDA launch that returns the summed
loss and, in the same pass, its gradient with respect to the render outputs.

    loss = lidar_main_loss(out["depth_lidar"], out["image_lidar"], gt, alpha_d=opt.alpha_d, alpha_r=opt.alpha_r,
                           alpha_i=opt.alpha_i, smooth=opt.smooth_factor)

equals  (alpha_d*L1(depth*m, gt_d*m) + alpha_r*MSE(raydrop, clamp(m, s, 1-s)) + alpha_i*MSE(intensity*m, gt_i*m)).sum()
with m = gt[..., 0] (ray-drop mask), gt_i = gt[..., 1], gt_d = gt[..., 2] - the default criteria of main_lidar4d.py:63-66.
No CPU path."""
from __future__ import annotations

import torch

from . import _capi


class _MainLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, image, gt, alpha_d, alpha_r, alpha_i, smooth):
        if not depth.is_cuda:
            raise RuntimeError("lidar_main_loss runs on CUDA only (no CPU fallback)")
        lib = _capi.load_library()
        n = depth.numel()
        d = depth.detach().contiguous().view(-1).float()
        im = image.detach().contiguous().view(-1, 2).float()
        g = gt.detach().contiguous().view(-1, 3).float()
        if im.shape[0] != n or g.shape[0] != n:
            raise ValueError("depth [.., N], image [.., N, 2] and gt [.., N, 3] must describe the same rays")
        loss = torch.zeros((), device=d.device)
        gd = torch.empty_like(d)
        gi = torch.empty_like(im)
        with torch.cuda.device(d.device):
            rc = lib.l4d_lidar_loss(d.data_ptr(), im.data_ptr(), g.data_ptr(), n, float(alpha_d), float(alpha_r), float(alpha_i),
                                    float(smooth), loss.data_ptr(), gd.data_ptr(), gi.data_ptr(),
                                    torch.cuda.current_stream(d.device).cuda_stream)
        _capi.check(lib, rc, "l4d_lidar_loss")
        ctx.save_for_backward(gd, gi)
        ctx.shapes = (depth.shape, image.shape)
        return loss

    @staticmethod
    def backward(ctx, g):
        gd, gi = ctx.saved_tensors
        return (gd * g).view(ctx.shapes[0]), (gi * g).view(ctx.shapes[1]), None, None, None, None, None


def lidar_main_loss(depth, image, gt, alpha_d=1.0, alpha_r=0.01, alpha_i=0.1, smooth=0.2):
    return _MainLoss.apply(depth, image, gt, alpha_d, alpha_r, alpha_i, smooth)
