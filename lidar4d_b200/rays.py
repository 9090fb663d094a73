"""LiDAR ray generation with the reference's sensor model
(/root/reference/data/base_dataset.py:82-97) and the synthetic sweep of
SURVEY.md §8(d).  Host-side numpy; used for synthetic inputs by bench/tests."""
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
