"""Model families beyond the reference's zoo (BASELINE config 5 names a ResNet-20).

``from gossipy_b200.models import ResNet20`` -- everything in :mod:`gossipy_b200.model.nn` is
re-exported here as well so either import path works.
"""
from ..model import TorchModel
from ..model.nn import (AdaLine, CIFAR10Net, LinearRegression, LogisticRegression, TorchMLP,
                        TorchPerceptron)
from .resnet import ResNet20

__all__ = ["TorchModel", "TorchPerceptron", "TorchMLP", "AdaLine", "LogisticRegression",
           "LinearRegression", "CIFAR10Net", "ResNet20"]
