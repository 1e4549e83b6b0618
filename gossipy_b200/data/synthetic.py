"""Synthetic data sets with the shapes of the sets the reference scripts download.

The target boxes have no network, so benchmarks and tests use Gaussian features labelled by a
random linear (or shallow non-linear) teacher.  Shapes: spambase 4601x57x2, MNIST-like
60000/10000x784x10, CIFAR-10-like 50000/10000x3x32x32x10, MovieLens-like rating lists.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch

SHAPES = {  # name: (n_samples, n_features, n_classes)
    "spambase": (4601, 57, 2), "sonar": (208, 60, 2), "ionosphere": (351, 34, 2),
    "abalone": (4177, 8, 3), "banknote": (1372, 4, 2), "reuters": (2600, 9947, 2),
    "mnist": (70000, 784, 10),
}


def teacher_classification(n: int, d: int, c: int, seed: int = 0, noise: float = 0.5,
                           hidden: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """``X ~ N(0, I)``; labels = argmax of a random teacher (linear, or 1-hidden-layer tanh)."""
    g = torch.Generator().manual_seed(seed)
    X = torch.randn(n, d, generator=g)
    if hidden:
        H = torch.tanh(X @ torch.randn(d, hidden, generator=g) / d ** 0.5)
        logits = H @ torch.randn(hidden, c, generator=g)
    else:
        logits = X @ torch.randn(d, c, generator=g) / d ** 0.5
    logits = logits + noise * torch.randn(n, c, generator=g)
    return X, logits.argmax(dim=1)


def classification_like(name: str, as_tensor: bool = True, seed: int = 0):
    n, d, c = SHAPES[name]
    X, y = teacher_classification(n, d, c, seed=seed)
    return (X, y) if as_tensor else (X.numpy().astype("float64"), y.numpy())


def mnist_like(n_train: int = 60000, n_test: int = 10000, seed: int = 0):
    """MNIST-shaped flat vectors: ``((Xtr[60000,784], ytr), (Xte[10000,784], yte))``."""
    X, y = teacher_classification(n_train + n_test, 784, 10, seed=seed, noise=0.3)
    return (X[:n_train], y[:n_train]), (X[n_train:], y[n_train:])


def spambase_like(n_train: int = 4141, n_test: int = 460, seed: int = 0):
    """spambase-shaped (57 features, 2 classes): ``((Xtr, ytr), (Xte, yte))`` with labels in {0,1}."""
    X, y = teacher_classification(n_train + n_test, 57, 2, seed=seed, noise=0.3)
    return (X[:n_train], y[:n_train]), (X[n_train:], y[n_train:])


def images_like(name: str, as_tensor: bool = True, seed: int = 0, n_train: int = None,
                n_test: int = None):
    """CIFAR-10 (3x32x32) / Fashion-MNIST (28x28) shaped images in [0,1], ten classes.

    Every class has a smooth prototype image (a random 4x4 field per channel, bilinearly upsampled); a sample is its
    class prototype plus white noise, squashed to [0, 1].  The class signal is spatially structured, so convolutional
    nets with global pooling (ResNet-20) learn it as readily as fully connected ones -- a pixel-wise linear teacher, the
    previous generator, is invisible to them."""
    shape = {"cifar10": (3, 32, 32), "fashionmnist": (28, 28)}[name]
    ntr = n_train or {"cifar10": 50000, "fashionmnist": 60000}[name]
    nte = n_test or 10000
    g = torch.Generator().manual_seed(seed)
    chans = shape[0] if len(shape) == 3 else 1
    hw = shape[-2:]
    proto = torch.nn.functional.interpolate(torch.randn(10, chans, 4, 4, generator=g), size=hw, mode="bilinear",
                                            align_corners=False)
    y = torch.randint(0, 10, (ntr + nte,), generator=g)
    X = torch.sigmoid(1.5 * proto[y] + torch.randn(ntr + nte, chans, *hw, generator=g)).reshape(ntr + nte, *shape)
    if not as_tensor:
        return (X[:ntr].numpy(), y[:ntr].tolist()), (X[ntr:].numpy(), y[ntr:].tolist())
    return (X[:ntr], y[:ntr]), (X[ntr:], y[ntr:])


def ratings_like(name: str = "ml-100k", seed: int = 0, rank: int = 5):
    """Low-rank synthetic ratings: ``({user: [(item, r)]}, n_users, n_items)``."""
    n_users, n_items, per_user = {"ml-100k": (943, 1682, 106), "ml-1m": (6040, 3706, 165),
                                  "ml-10m": (6040, 3706, 165), "ml-20m": (6040, 3706, 165),
                                  "tiny": (64, 200, 30)}[name]
    rng = np.random.default_rng(seed)
    U = rng.normal(size=(n_users, rank)) / rank ** 0.5
    V = rng.normal(size=(n_items, rank))
    ratings: Dict[int, List[Tuple[int, float]]] = {}
    for u in range(n_users):
        items = rng.choice(n_items, size=min(per_user, n_items), replace=False)
        r = np.clip(np.rint(3.0 + 1.2 * (V[items] @ U[u]) + 0.3 * rng.normal(size=len(items))), 1, 5)
        ratings[u] = [(int(i), float(x)) for i, x in zip(items, r)]
    return ratings, n_users, n_items
