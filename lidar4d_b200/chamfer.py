"""Chamfer / nearest-neighbour distance op.

Drop-in for the reference's ``utils/chamfer3D/dist_chamfer_3D.py``: same class names, same call
(``chamfer_3DDist()(xyz1[B,N,3], xyz2[B,M,3]) -> dist1[B,N], dist2[B,M], idx1[B,N], idx2[B,M]``), same
autograd contract (``chamfer_3DFunction``, dist_chamfer_3D.py:31-73), backed by ``l4d_chamfer_forward`` /
``l4d_chamfer_backward`` of the C-ABI library.  Differences that SURVEY.md 5 lists as reference defects:
the kernels run on torch's *current* stream (the reference launches on the default stream,
chamfer3D.cu:141-142) and return codes are checked (dist_chamfer_3D.py:54 ignores them).
"""

import torch
from torch import nn
from torch.autograd import Function

from . import _capi


def _check_cloud(x, name):
    if x.dim() != 3 or x.shape[-1] != 3:
        raise ValueError(f"{name}: expected [B, N, 3], got {tuple(x.shape)}")     # dist_chamfer_3D.py:35-38
    if not x.is_cuda:
        raise RuntimeError("chamfer_3DDist needs CUDA tensors (no CPU fallback)")
    if x.shape[1] == 0:
        raise ValueError(f"{name}: empty point cloud")


class chamfer_3DFunction(Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        _check_cloud(xyz1, "xyz1")
        _check_cloud(xyz2, "xyz2")
        if xyz1.shape[0] != xyz2.shape[0]:
            raise ValueError("batch sizes differ")
        lib = _capi.load_library()
        ctx.in_dtypes = (xyz1.dtype, xyz2.dtype)
        xyz1 = xyz1.contiguous().float()
        xyz2 = xyz2.contiguous().float()
        b, n, _ = xyz1.shape
        m = xyz2.shape[1]
        dev = xyz1.device
        dist1 = torch.empty(b, n, device=dev)
        dist2 = torch.empty(b, m, device=dev)
        idx1 = torch.empty(b, n, dtype=torch.int32, device=dev)
        idx2 = torch.empty(b, m, dtype=torch.int32, device=dev)
        nwork = lib.l4d_chamfer_work_bytes(b, n, m)
        work = torch.empty(nwork, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = lib.l4d_chamfer_forward(xyz1.data_ptr(), xyz2.data_ptr(), b, n, m, dist1.data_ptr(), dist2.data_ptr(),
                                         idx1.data_ptr(), idx2.data_ptr(), work.data_ptr(), nwork,
                                         torch.cuda.current_stream(dev).cuda_stream)
        _capi.check(lib, rc, "l4d_chamfer_forward")
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        ctx.mark_non_differentiable(idx1, idx2)
        return dist1, dist2, idx1, idx2

    @staticmethod
    def backward(ctx, graddist1, graddist2, gradidx1, gradidx2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        lib = _capi.load_library()
        b, n, _ = xyz1.shape
        m = xyz2.shape[1]
        dev = xyz1.device
        g1 = torch.zeros(b, n, device=dev) if graddist1 is None else graddist1.contiguous().float()
        g2 = torch.zeros(b, m, device=dev) if graddist2 is None else graddist2.contiguous().float()
        gx1 = torch.zeros_like(xyz1)
        gx2 = torch.zeros_like(xyz2)
        with torch.cuda.device(dev):
            rc = lib.l4d_chamfer_backward(xyz1.data_ptr(), xyz2.data_ptr(), b, n, m, g1.data_ptr(), g2.data_ptr(),
                                          idx1.data_ptr(), idx2.data_ptr(), gx1.data_ptr(), gx2.data_ptr(),
                                          torch.cuda.current_stream(dev).cuda_stream)
        _capi.check(lib, rc, "l4d_chamfer_backward")
        return gx1.to(ctx.in_dtypes[0]), gx2.to(ctx.in_dtypes[1])      # gradients in the inputs' dtype (fp16 clouds under autocast)


class chamfer_3DDist(nn.Module):
    def forward(self, input1, input2):
        return chamfer_3DFunction.apply(input1.contiguous(), input2.contiguous())
