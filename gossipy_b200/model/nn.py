"""Model zoo of the reference (``gossipy/model/nn.py:26-198``) plus the script-local CIFAR net.

Larger architectures (ResNet-20) live in :mod:`gossipy_b200.models`.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Sequence

import torch
from torch import nn
from torch.nn.init import xavier_uniform_

from . import TorchModel

__all__ = ["TorchPerceptron", "TorchMLP", "AdaLine", "LogisticRegression", "LinearRegression",
           "CIFAR10Net"]


class TorchPerceptron(TorchModel):
    """``activation(Linear(dim, 1))`` with Xavier-uniform weight."""

    def __init__(self, dim: int, activation=nn.Sigmoid, bias: bool = True) -> None:
        super().__init__()
        self.input_dim = dim
        self.model = nn.Sequential(OrderedDict(linear=nn.Linear(dim, 1, bias=bias),
                                               sigmoid=activation()))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.model(x)

    def init_weights(self) -> None:
        xavier_uniform_(self.model.linear.weight)

    def __str__(self) -> str:
        return "TorchPerceptron(size=%d)\n%s" % (self.get_size(), str(self.model))


class TorchMLP(TorchModel):
    """Fully connected net ``Linear -> act`` per hidden layer + output ``Linear`` (no softmax)."""

    def __init__(self, input_dim: int, output_dim: int, hidden_dims: Sequence[int] = (100,),
                 activation=nn.ReLU) -> None:
        super().__init__()
        self.input_dim, self.output_dim = input_dim, output_dim
        self.hidden_dims = tuple(hidden_dims)
        self._activation = activation
        dims = [input_dim] + list(hidden_dims)
        layers = OrderedDict()
        for i in range(len(dims) - 1):
            layers["linear_%d" % (i + 1)] = nn.Linear(dims[i], dims[i + 1])
            layers["activ_%d" % (i + 1)] = activation()
        layers["linear_%d" % len(dims)] = nn.Linear(dims[-1], output_dim)
        self.model = nn.Sequential(layers)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.model(x)

    def init_weights(self) -> None:
        for m in self.model:
            if isinstance(m, nn.Linear):
                xavier_uniform_(m.weight)

    def fused_family(self):
        if len(self.hidden_dims) == 1 and self._activation is nn.ReLU:
            return ("mlp1", (self.input_dim, self.hidden_dims[0], self.output_dim))
        return None

    def __str__(self) -> str:
        return "%s(size=%d)\n%s" % (self.__class__.__name__, self.get_size(), str(self.model))


class AdaLine(TorchModel):
    """A bare weight vector ``w``; ``forward(x) = w @ x.T`` (also the Pegasos model)."""

    def __init__(self, dim: int) -> None:
        super().__init__()
        self.input_dim = dim
        self.model = nn.Parameter(torch.zeros(dim), requires_grad=False)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.model @ x.T

    def get_size(self) -> int:
        return self.input_dim

    def init_weights(self) -> None:
        pass

    def fused_family(self):
        return ("linear", (self.input_dim,))


class LogisticRegression(TorchModel):
    """``sigmoid(Linear(in, out))``; the scripts feed it to CrossEntropyLoss as is."""

    def __init__(self, input_dim: int, output_dim: int) -> None:
        super().__init__()
        self.model = nn.Linear(input_dim, output_dim)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return torch.sigmoid(self.model(x))

    def init_weights(self) -> None:
        pass

    def fused_family(self):
        return ("logreg", (self.model.in_features, self.model.out_features))

    def __str__(self) -> str:
        return "LogisticRegression(in_size=%d, out_size=%d)" % (self.model.in_features,
                                                               self.model.out_features)


class LinearRegression(TorchModel):
    """Plain ``Linear``.  FIX(B20): instantiable (the reference lacks ``init_weights``)."""

    def __init__(self, input_dim: int, output_dim: int) -> None:
        super().__init__()
        self.model = nn.Linear(input_dim, output_dim)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.model(x)

    def init_weights(self) -> None:
        pass

    def __str__(self) -> str:
        return "LinearRegression(in_size=%d, out_size=%d)" % (self.model.in_features,
                                                             self.model.out_features)


class CIFAR10Net(TorchModel):
    """The 3-conv CNN of the PENS experiment (ref ``main_onoszko_2021.py:28-56``)."""

    def __init__(self) -> None:
        super().__init__()
        self.features = nn.Sequential(
            nn.Conv2d(3, 32, 3), nn.ReLU(), nn.MaxPool2d(2, 2),
            nn.Conv2d(32, 64, 3), nn.ReLU(), nn.MaxPool2d(2, 2),
            nn.Conv2d(64, 64, 3), nn.ReLU(), nn.MaxPool2d(2, 2))
        self.head = nn.Sequential(nn.Flatten(), nn.Linear(64 * 2 * 2, 64), nn.ReLU(),
                                  nn.Linear(64, 10))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.head(self.features(x))

    def init_weights(self, *args, **kwargs) -> None:
        pass  # torch defaults, as in the reference script

    def __str__(self) -> str:
        return "CIFAR10Net(size=%d)" % self.get_size()
