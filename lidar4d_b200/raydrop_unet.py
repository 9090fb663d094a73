"""Ray-drop refinement U-Net - the per-image post-process the reference applies to the rendered
(raydrop, intensity, depth) panorama (model/runner.py:413-415,838-851; network model/unet.py:139-170).

Outside the per-ray hot path and dense 2-D convolution work, so this is LIBRARY code by design (cuDNN convolutions,
torch's fused scaled-dot-product attention), channels-last and autocast-friendly (bf16 / fp16), not hand-written CUDA.
What is guaranteed is the interface: `RayDropUNet(in_channels, channels, out_channels)` has exactly the reference's
`state_dict` keys and shapes (checkpoints interchange, `LiDAR4D.state_dict()` keeps its `unet.*` entries) and computes
the same function (tests/test_raydrop_unet.py against tests/golden/unet.npz, produced by the reference module itself).

Structure restated from the reference (not copied): 1x1 stem -> 4 x [maxpool, pre-activation double conv] ->
multi-head self-attention at 1/16 resolution -> 4 x [bilinear x2 (align_corners), pad, concat skip, double conv] ->
BN-ReLU-1x1 -> sigmoid.  Two reference quirks are kept because they are part of the function: the attention output is
reinterpreted as [B,H,W,C] straight from [B,heads,HW,C/heads] memory (unet.py:101), and training-time attention dropout
is an additive -1e12 mask drawn with torch.bernoulli (unet.py:94-97).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def _pre_act_pair(c_in: int, c_out: int, c_mid: int | None = None, p: float = 0.1) -> nn.Sequential:
    """(BN, ReLU, Dropout2d, 3x3 conv) twice - indices 0..7 as in the reference's nn.Sequential."""
    c_mid = c_mid or c_out
    layers = []
    for a, b in ((c_in, c_mid), (c_mid, c_out)):
        layers += [nn.BatchNorm2d(a), nn.ReLU(inplace=True), nn.Dropout2d(p), nn.Conv2d(a, b, 3, padding=1, bias=False)]
    return nn.Sequential(*layers)


class _Holder(nn.Module):
    """Gives a child the attribute name the reference's state_dict uses (e.g. `down1.conv.double_conv.3.weight`)."""

    def __init__(self, **children):
        super().__init__()
        for k, v in children.items():
            setattr(self, k, v)


class _Attention(nn.Module):
    def __init__(self, ch: int, heads: int = 8, p: float = 0.1):
        super().__init__()
        self.proj_qkv = nn.Conv2d(ch, 3 * ch, 1, bias=False)
        self.proj = nn.Conv2d(ch, ch, 1, bias=False)
        self.norm = nn.BatchNorm2d(ch)
        self.heads, self.p = heads, p

    def forward(self, x):
        B, C, H, W = x.shape
        q, k, v = self.proj_qkv(self.norm(x)).chunk(3, dim=1)
        split = lambda t: t.reshape(B, self.heads, C // self.heads, H * W).transpose(2, 3)     # [B, heads, HW, C/heads]
        q, k, v = split(q), split(k), split(v)
        mask = None
        if self.training:            # the reference's dropout: some logits pushed to -1e12 before the softmax
            mask = torch.bernoulli(torch.full((B, self.heads, H * W, H * W), self.p, device=x.device, dtype=q.dtype)) * -1e12
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=mask)            # scale = (C/heads)^-0.5, as unet.py:91
        o = o.contiguous().view(B, H, W, C).permute(0, 3, 1, 2)                # the reference's reinterpretation (unet.py:101)
        return x + self.proj(o)


class RayDropUNet(nn.Module):
    def __init__(self, in_channels: int = 3, channels: int = 32, out_channels: int = 1):
        super().__init__()
        c = channels
        self.inc = _Holder(conv=nn.Conv2d(in_channels, c, 1))
        widths = [(c, 2 * c), (2 * c, 4 * c), (4 * c, 8 * c), (8 * c, 8 * c)]
        for i, (a, b) in enumerate(widths, 1):
            setattr(self, f"down{i}", _Holder(down=nn.MaxPool2d(2), conv=_Holder(double_conv=_pre_act_pair(a, b))))
        self.attn = _Attention(8 * c)
        for i, (a, b) in enumerate([(16 * c, 4 * c), (8 * c, 2 * c), (4 * c, c), (2 * c, c)], 1):
            setattr(self, f"up{i}", _Holder(up=nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True),
                                            conv=_Holder(double_conv=_pre_act_pair(a, b, a))))
        self.outc = _Holder(conv=nn.Sequential(nn.BatchNorm2d(c), nn.ReLU(inplace=True), nn.Conv2d(c, out_channels, 1)))
        self.sigmoid = nn.Sigmoid()

    @staticmethod
    def _merge(stage: _Holder, low, skip):
        low = stage.up(low)
        dy, dx = skip.shape[2] - low.shape[2], skip.shape[3] - low.shape[3]
        low = F.pad(low, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
        return stage.conv.double_conv(torch.cat([skip, low], dim=1))

    def forward(self, x):
        if x.is_cuda:
            x = x.contiguous(memory_format=torch.channels_last)      # NHWC kernels of cuDNN
        feats = [self.inc.conv(x)]
        for i in range(1, 5):
            st = getattr(self, f"down{i}")
            feats.append(st.conv.double_conv(st.down(feats[-1])))
        y = self.attn(feats[4])
        for i, skip in zip(range(1, 5), (feats[3], feats[2], feats[1], feats[0])):
            y = self._merge(getattr(self, f"up{i}"), y, skip)
        return self.sigmoid(self.outc.conv(y))
