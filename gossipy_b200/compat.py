"""Run scripts written for the reference (``from gossipy.node import GossipNode`` ...) on this framework, unmodified.

``install()`` registers this package and its sub-modules under the name ``gossipy`` in ``sys.modules`` -- the module
layout is the reference's (``gossipy/__init__.py``, ``core.py``, ``node.py``, ``simul.py``, ``flow_control.py``,
``utils.py``, ``data/``, ``model/{handler,nn,sampling}.py``), so every import of a reference script resolves here.

As a runner::

    python -m gossipy_b200.compat [--synthetic] [--max-rounds N] [--engine native] [--device cuda] \\
                                  [--native-utility K] /root/reference/main_hegedus_2021.py [script args]

``--synthetic``       data sets that cannot be downloaded are replaced by synthetic data of the same shape
``--max-rounds N``    clamps every ``start(n_rounds)`` (the reference scripts run 100 - 1 000 rounds)
``--engine native``   C++ scheduler + executor / bank instead of the Python loop (scripts never set ``sim.engine``)
``--native-utility``  tokenized runs: the utility the C++ scheduler evaluates the token accounts with (not needed when the
                      script's ``utility_fun`` is ``lambda ...: <int>`` like in the reference scripts -- that is recognised)
``--device``          ``GlobalSettings().set_device`` before the script starts
"""
from __future__ import annotations

import argparse
import importlib
import os
import runpy
import sys
from typing import List, Optional

SUBMODULES = ("core", "node", "simul", "flow_control", "utils", "data", "data.handler", "model", "model.handler",
              "model.nn", "model.sampling")


def install(name: str = "gossipy") -> None:
    """Make ``import <name>`` (and its sub-modules) resolve to this framework."""
    pkg = importlib.import_module("gossipy_b200")
    sys.modules[name] = pkg
    for sub in SUBMODULES:
        sys.modules[name + "." + sub] = importlib.import_module("gossipy_b200." + sub)


def uninstall(name: str = "gossipy") -> None:
    for key in [name] + [name + "." + sub for sub in SUBMODULES]:
        mod = sys.modules.get(key)
        if mod is not None and getattr(mod, "__name__", "").startswith("gossipy_b200"):
            del sys.modules[key]


def _clamp_rounds(limit: int) -> None:
    from . import simul

    def wrap(cls):
        orig = cls.start

        def start(self, *args, **kwargs):
            # n_rounds is the first argument of GossipSimulator.start and the second of All2AllGossipSimulator.start
            pos = 1 if cls is simul.All2AllGossipSimulator else 0
            if "n_rounds" in kwargs:
                kwargs["n_rounds"] = min(int(kwargs["n_rounds"]), limit)
            elif len(args) > pos:
                args = args[:pos] + (min(int(args[pos]), limit),) + args[pos + 1:]
            else:
                kwargs["n_rounds"] = limit
            return orig(self, *args, **kwargs)
        cls.start = start
    wrap(simul.GossipSimulator)
    wrap(simul.All2AllGossipSimulator)


def main(argv: Optional[List[str]] = None) -> None:
    ap = argparse.ArgumentParser(prog="python -m gossipy_b200.compat", description=__doc__.split("\n\n")[0])
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--max-rounds", type=int, default=None)
    ap.add_argument("--engine", default=None, choices=["python", "native"])
    ap.add_argument("--native-utility", type=int, default=None)
    ap.add_argument("--device", default=None)
    ap.add_argument("script")
    ap.add_argument("args", nargs=argparse.REMAINDER)
    a = ap.parse_args(argv)
    if a.synthetic:
        os.environ["GOSSIPY_SYNTHETIC_FALLBACK"] = "1"
    install()
    from . import GlobalSettings, simul
    if a.device:
        GlobalSettings().set_device(a.device)
    if a.engine:
        simul.GossipSimulator.engine = a.engine
    if a.native_utility is not None:
        simul.TokenizedGossipSimulator.native_utility = a.native_utility
    if a.max_rounds is not None:
        _clamp_rounds(a.max_rounds)
    sys.argv = [a.script] + list(a.args)
    runpy.run_path(a.script, run_name="__main__")


if __name__ == "__main__":
    main()
