"""Oracle: LPIPS (AlexNet, v0.1) perceptual distance (test infrastructure).

Call sites: latentblending/blending_engine.py:74-76 (``lpips.LPIPS(net='alex')``)
and :744-758 (``get_lpips_similarity``).  Code behind it: lpips==0.1.4
(requirements.txt:1), NOT vendored and not installed.  Restated from the
published definition: ScalingLayer, 5 AlexNet ReLU taps (64,192,384,256,256),
unit-normalise over channels (eps 1e-10), squared difference, non-negative 1x1
"lin" layers, spatial mean, sum over taps.  No pretrained weights are available
offline, so weights are seeded-random (the metric then only ranks gaps, which
is all the branch-placement logic needs).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

_SHIFT = (-0.030, -0.088, -0.188)
_SCALE = (0.458, 0.448, 0.450)
_CHANNELS = (64, 192, 384, 256, 256)


class LPIPSAlex(nn.Module):
    def __init__(self, seed=2):
        super().__init__()
        self.convs = nn.ModuleList([
            nn.Conv2d(3, 64, 11, stride=4, padding=2),
            nn.Conv2d(64, 192, 5, padding=2),
            nn.Conv2d(192, 384, 3, padding=1),
            nn.Conv2d(384, 256, 3, padding=1),
            nn.Conv2d(256, 256, 3, padding=1),
        ])
        self.lins = nn.ModuleList([nn.Conv2d(c, 1, 1, bias=False) for c in _CHANNELS])
        self.register_buffer("shift", torch.tensor(_SHIFT).view(1, 3, 1, 1))
        self.register_buffer("scale", torch.tensor(_SCALE).view(1, 3, 1, 1))
        g = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            for c in self.convs:
                fan_in = c.weight[0].numel()
                c.weight.copy_(torch.randn(c.weight.shape, generator=g) * (2.0 / fan_in) ** 0.5)
                c.bias.zero_()
            for l in self.lins:
                l.weight.copy_(torch.rand(l.weight.shape, generator=g) / l.weight.shape[1])

    def features(self, x):
        x = (x - self.shift) / self.scale
        taps = []
        for i, c in enumerate(self.convs):
            x = F.relu(c(x))
            taps.append(x)
            if i in (0, 1):
                x = F.max_pool2d(x, 3, 2)
        return taps

    def forward(self, a, b):
        total = 0.0
        for fa, fb, lin in zip(self.features(a), self.features(b), self.lins):
            na = fa / (fa.pow(2).sum(1, keepdim=True).sqrt() + 1e-10)
            nb = fb / (fb.pow(2).sum(1, keepdim=True).sqrt() + 1e-10)
            total = total + lin((na - nb) ** 2).mean(dim=(2, 3), keepdim=True)
        return total


def lpips_distance(net: LPIPSAlex, img_a, img_b):
    """blending_engine.py:750-758: uint8 HxWx3 -> [-1,1] NCHW fp32 -> float."""
    def prep(img):
        t = torch.from_numpy(np.asarray(img)).float()
        return (2 * t / 255.0 - 1).permute(2, 0, 1).unsqueeze(0)
    return float(net(prep(img_a), prep(img_b))[0, 0, 0, 0])
