"""ResNet-20 for 32x32 inputs (He et al. 2016, CIFAR variant: 3 stages x 3 basic blocks).

The reference cannot gossip BatchNorm models at all (its merge crashes on the int64
``num_batches_tracked`` buffer).  Here float buffers (running mean/var) live in the
flat parameter row and are merged like weights, while integer counters are combined with
``max`` (see :class:`gossipy_b200.engine.flat.FlatLayout`).
"""
from __future__ import annotations

import torch
from torch import nn
import torch.nn.functional as F

from ..model import TorchModel


class _BasicBlock(nn.Module):
    def __init__(self, c_in: int, c_out: int, stride: int) -> None:
        super().__init__()
        self.conv1 = nn.Conv2d(c_in, c_out, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(c_out)
        self.conv2 = nn.Conv2d(c_out, c_out, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(c_out)
        self.short = None
        if stride != 1 or c_in != c_out:
            self.short = nn.Sequential(nn.Conv2d(c_in, c_out, 1, stride, bias=False),
                                       nn.BatchNorm2d(c_out))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        out = F.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return F.relu(out + (x if self.short is None else self.short(x)))


class ResNet20(TorchModel):
    """~0.27 M parameters; ``width`` scales the channel counts (16/32/64 by default)."""

    def __init__(self, n_classes: int = 10, width: int = 16, in_channels: int = 3) -> None:
        super().__init__()
        self.stem = nn.Sequential(nn.Conv2d(in_channels, width, 3, 1, 1, bias=False),
                                  nn.BatchNorm2d(width), nn.ReLU())
        blocks, c_in = [], width
        for stage, c_out in enumerate((width, 2 * width, 4 * width)):
            for b in range(3):
                blocks.append(_BasicBlock(c_in, c_out, 2 if (stage > 0 and b == 0) else 1))
                c_in = c_out
        self.blocks = nn.Sequential(*blocks)
        self.fc = nn.Linear(c_in, n_classes)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self.blocks(self.stem(x))
        return self.fc(torch.flatten(F.adaptive_avg_pool2d(x, 1), 1))

    def init_weights(self) -> None:
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
