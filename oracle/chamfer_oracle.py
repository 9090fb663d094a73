"""CPU restatement of the reference's chamfer op -- TEST INFRASTRUCTURE ONLY (imported by tests/ and the chamfer
probe's cpu leg; the product never imports it).

Follows /root/reference/utils/chamfer3D/chamfer3D.cu:
  * NmDistanceKernel (:11-133): for every point j of cloud A the squared distance to, and the index of, its nearest
    point of cloud B.  Candidates are scanned in increasing index with a strict '<' (:33,43,...) and chunk results are
    merged with a strict '>' (:127), so the smallest index wins exact ties.  The distance is evaluated in fp32 as
    ``dx*dx + dy*dy + dz*dz`` with dx = b - a (:29-32); nvcc 12.9 contracts that expression to
    fma(dz,dz, fma(dx,dx, dy*dy)) (checked in SASS), which is emulated here through float64 (the product of two fp32
    values is exact in float64).
  * NmDistanceGradKernel (:154-174): g = 2*grad_dist; grad_a[j] += g*(a_j - b_idx); grad_b[idx] -= g*(a_j - b_idx),
    launched once per direction (:183-184).
PARITY UNPINNED for this op: the reference holds no tests or golden vectors for it and its kernels are CUDA-only
(cannot run in the build container), so the oracle is anchored on the cited lines and cross-checked against an
independent float64 ``torch.cdist`` argmin in tests/test_chamfer.py.
"""
import numpy as np


def _fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def nn_distance(a: np.ndarray, b: np.ndarray, chunk: int = 2048):
    """a [n,3], b [m,3] fp32 -> (dist [n] fp32, idx [n] int32), chamfer3D.cu:11-133."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    n = a.shape[0]
    dist = np.empty(n, np.float32)
    idx = np.empty(n, np.int32)
    for s in range(0, n, chunk):
        q = a[s:s + chunk]
        dx = (b[None, :, 0] - q[:, None, 0]).astype(np.float32)
        dy = (b[None, :, 1] - q[:, None, 1]).astype(np.float32)
        dz = (b[None, :, 2] - q[:, None, 2]).astype(np.float32)
        d = _fma32(dz, dz, _fma32(dx, dx, (dy * dy).astype(np.float32)))
        i = np.argmin(d, axis=1)              # first occurrence of the minimum == smallest index on ties
        idx[s:s + chunk] = i
        dist[s:s + chunk] = d[np.arange(q.shape[0]), i]
    return dist, idx


def chamfer_forward(xyz1: np.ndarray, xyz2: np.ndarray):
    """[B,N,3], [B,M,3] -> dist1 [B,N], dist2 [B,M], idx1, idx2 (chamfer3D.cu:136-151)."""
    B = xyz1.shape[0]
    o = [nn_distance(xyz1[i], xyz2[i]) for i in range(B)]
    r = [nn_distance(xyz2[i], xyz1[i]) for i in range(B)]
    return (np.stack([x[0] for x in o]), np.stack([x[0] for x in r]), np.stack([x[1] for x in o]), np.stack([x[1] for x in r]))


def chamfer_backward(xyz1, xyz2, g1, g2, idx1, idx2):
    """chamfer3D.cu:154-184, accumulated in float64 (the kernel's atomics have no defined order)."""
    gx1 = np.zeros(xyz1.shape, np.float64)
    gx2 = np.zeros(xyz2.shape, np.float64)
    for i in range(xyz1.shape[0]):
        for (a, b, g, idx, ga, gb) in ((xyz1[i], xyz2[i], g1[i], idx1[i], gx1[i], gx2[i]),
                                       (xyz2[i], xyz1[i], g2[i], idx2[i], gx2[i], gx1[i])):
            t = (np.float32(2) * g.astype(np.float32))[:, None] * (a - b[idx]).astype(np.float32)
            ga += t
            np.add.at(gb, idx, -t.astype(np.float64))
    return gx1.astype(np.float32), gx2.astype(np.float32)
