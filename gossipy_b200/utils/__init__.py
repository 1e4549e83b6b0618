"""Helpers (reference: ``gossipy/utils.py:28-189``) plus engine utilities.

``plot_evaluation`` degrades to a textual summary when matplotlib is not installed (it is not in
the B200 image); downloads raise a clear error when there is no network.
"""
from __future__ import annotations

import io
import tarfile
from json import JSONEncoder
from typing import Dict, List
from zipfile import ZipFile

import numpy as np
import torch

from .. import LOG

__all__ = ["choice_not_n", "torch_models_eq", "download_and_unzip", "download_and_untar",
           "plot_evaluation", "StringEncoder"]


def choice_not_n(mn: int, mx: int, notn: int) -> int:
    """Uniform integer in ``[mn, mx)`` different from ``notn`` (ref ``utils.py:41-64``).

    Drawn without rejection: sample from a range one shorter and skip over ``notn``.
    """
    if not (mn <= notn < mx):
        return int(np.random.randint(mn, mx))
    assert mx - mn > 1, "no admissible value"
    c = int(np.random.randint(mn, mx - 1))
    return c + 1 if c >= notn else c


def torch_models_eq(m1: torch.nn.Module, m2: torch.nn.Module) -> bool:
    """True when both modules have identical ``state_dict`` keys and values."""
    sd1, sd2 = m1.state_dict(), m2.state_dict()
    if list(sd1.keys()) != list(sd2.keys()):
        return False
    return all(torch.equal(sd1[k].cpu(), sd2[k].cpu()) for k in sd1)


def _fetch(url: str) -> bytes:
    from urllib.error import URLError
    from urllib.request import urlopen
    LOG.info("Downloading %s" % url)
    try:
        return urlopen(url, timeout=30).read()
    except URLError as exc:
        raise RuntimeError("cannot download %s (no network?) -- use the synthetic generators in "
                           "gossipy_b200.data.synthetic instead" % url) from exc


def download_and_unzip(url: str, extract_to: str = ".") -> List[str]:
    archive = ZipFile(io.BytesIO(_fetch(url)))
    archive.extractall(path=extract_to)
    return archive.namelist()


def download_and_untar(url: str, extract_to: str = ".") -> List[str]:
    archive = tarfile.open(fileobj=io.BytesIO(_fetch(url)), mode="r:gz")
    archive.extractall(path=extract_to)
    return archive.getnames()


def plot_evaluation(evals: List[List[Dict]], title: str = "Untitled plot") -> None:
    """Plot mean +/- std of every metric over repetitions (ref ``utils.py:152-183``)."""
    if not evals or not evals[0] or not evals[0][0]:
        return
    stats = {}
    for k in evals[0][0]:
        series = np.array([[d[k] for d in run] for run in evals], dtype=float)
        stats[k] = (series.mean(axis=0), series.std(axis=0))
        LOG.info("%s: %.4f" % (k, stats[k][0][-1]))
    try:
        import matplotlib.pyplot as plt
    except Exception:
        LOG.info("matplotlib not available: '%s' summarised in the log only" % title)
        return
    fig = plt.figure()
    ax = fig.add_subplot(111)
    for k, (mu, sd) in stats.items():
        xs = range(1, len(mu) + 1)
        ax.fill_between(xs, mu - sd, mu + sd, alpha=0.2)
        ax.plot(xs, mu, label=k)
    ax.set_title(title)
    ax.set_xlabel("cycle")
    ax.set_ylabel("metric value")
    ax.legend(loc="lower right")
    plt.show()


class StringEncoder(JSONEncoder):
    def default(self, o):  # noqa: D102
        return str(o)

from . import profiling  # noqa: E402,F401  (nvtx_range, DeviceTimer, PhaseProfile)
