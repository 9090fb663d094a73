"""LiDAR ray generation with the reference's sensor model
(/root/reference/data/base_dataset.py:82-97) and the synthetic sweep of
SURVEY.md §8(d).  `lidar_rays` / `synthetic_sweep`: host-side numpy for synthetic inputs of bench and tests;
`get_lidar_rays`: the dataset-side entry point with the reference's signature, one CUDA launch."""
from __future__ import annotations

import numpy as np


def lidar_rays(pose: np.ndarray, H: int, W: int, fov_up: float = 2.0, fov: float = 26.9):
    """All H*W rays of one sweep: beta=-(i-W/2)/W*2pi, alpha=(fov_up - j/H*fov)*pi/180,
    d=(cos a cos b, cos a sin b, sin a) @ R^T, o = pose translation (float32)."""
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing="xy")
    i = i.reshape(-1)
    j = j.reshape(-1)
    beta = -(i - np.float32(W / 2)) / np.float32(W) * np.float32(2 * np.pi)
    alpha = (np.float32(fov_up) - j / np.float32(H) * np.float32(fov)) / np.float32(180) * np.float32(np.pi)
    d = np.stack([np.cos(alpha) * np.cos(beta), np.cos(alpha) * np.sin(beta), np.sin(alpha)], -1).astype(np.float32)
    R = pose[:3, :3].astype(np.float32)
    rays_d = (d @ R.T).astype(np.float32)
    rays_o = np.broadcast_to(pose[:3, 3].astype(np.float32), rays_d.shape).copy()
    return rays_o, rays_d


def synthetic_sweep(frame: int, n_frames: int = 50, H: int = 64, W: int = 1024,
                    fov_up: float = 2.0, fov: float = 26.9):
    """Frame `frame` of the synthetic straight-line trajectory of SURVEY.md §8(d):
    o_k = (-0.25 + 0.5 k/(n-1), 0, 0), identity rotation, t_k = k/n."""
    pose = np.eye(4, dtype=np.float32)
    pose[0, 3] = -0.25 + 0.5 * frame / max(n_frames - 1, 1)
    ro, rd = lidar_rays(pose, H, W, fov_up, fov)
    return ro, rd, np.float32(frame / n_frames)


def get_lidar_rays(poses, intrinsics, H, W, N=-1, patch_size=1, images=None):
    """data/base_dataset.py:15-102 `get_lidar_rays` on the GPU in ONE launch (the reference runs ~25 elementwise torch
    ops + a matmul): poses [B,4,4] cam2world on the device, intrinsics = (fov_up, fov) -> {"rays_o", "rays_d": [B,N,3],
    "inds": [B,N]}.  Pixel selection (N > 0: random pixels or patches, may duplicate) uses torch.randint exactly like the
    reference (:47-72).  Extra: images [B,H,W,C] -> "gt" [B,N,C] gathered in the same launch (kitti360_dataset.py:170-178).
    No CPU path."""
    import torch
    from . import _capi
    if not poses.is_cuda:
        raise RuntimeError("get_lidar_rays runs on CUDA only (no CPU fallback); lidar_rays() is the numpy sensor model")
    lib = _capi.load_library()
    dev = poses.device
    B = poses.shape[0]
    inds = None
    if N > 0:
        N = min(N, H * W)
        if isinstance(patch_size, int):
            px, py = patch_size, patch_size
        elif len(patch_size) == 1:
            px, py = patch_size[0], patch_size[0]
        else:
            px, py = patch_size
        if px > 0:
            num_patch = N // (px * py)
            ix = torch.randint(0, H - px, size=[num_patch], device=dev)
            iy = torch.randint(0, W, size=[num_patch], device=dev)
            pi, pj = torch.meshgrid(torch.arange(px, device=dev), torch.arange(py, device=dev), indexing="ij")
            off = torch.stack([pi.reshape(-1), pj.reshape(-1)], dim=-1)
            ind2 = (torch.stack([ix, iy], dim=-1).unsqueeze(1) + off.unsqueeze(0)).view(-1, 2)
            ind2[:, 1] = ind2[:, 1] % W
            inds = ind2[:, 0] * W + ind2[:, 1]
        else:
            inds = torch.randint(0, H * W, size=[N], device=dev)
        inds = inds.long().contiguous()
    n = int(inds.numel()) if inds is not None else H * W
    poses = poses.detach().contiguous().float()
    rays_o = torch.empty(B, n, 3, device=dev)
    rays_d = torch.empty(B, n, 3, device=dev)
    gt = None
    if images is not None:
        images = images.detach().contiguous().float().view(B, H * W, -1)
        gt = torch.empty(B, n, images.shape[-1], device=dev)
    fov_up, fov = float(intrinsics[0]), float(intrinsics[1])
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        for b in range(B):
            rc = lib.l4d_lidar_rays(poses[b].data_ptr(), fov_up, fov, H, W, inds.data_ptr() if inds is not None else None, n,
                                    images[b].data_ptr() if gt is not None else None, images.shape[-1] if gt is not None else 0,
                                    rays_o[b].data_ptr(), rays_d[b].data_ptr(), gt[b].data_ptr() if gt is not None else None, st)
            _capi.check(lib, rc, "l4d_lidar_rays")
    res = {"rays_o": rays_o, "rays_d": rays_d,
           "inds": (inds if inds is not None else torch.arange(H * W, device=dev)).expand([B, n])}
    if gt is not None:
        res["gt"] = gt
    return res
