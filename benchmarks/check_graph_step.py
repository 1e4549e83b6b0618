#!/usr/bin/env python
"""Generic (autograd) training step: CUDA-graph replay vs eager launches of the same step (model/handler.py::_graph_fwd_bwd).

Prints, for cuDNN with / without TF32 and with / without deterministic algorithms, the largest weight difference after
21 momentum-SGD steps between (a) two eager runs (run-to-run noise of the library kernels) and (b) the graph run and an
eager run -- (b) must not exceed (a) by more than the algorithm-choice noise -- and the time per step of both.
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))


def main():
    import gossipy_b200 as g
    from gossipy_b200.model.handler import TorchModelHandler
    from gossipy_b200.model.nn import TorchModel
    from test_kernels_gpu import _ConvBN

    class Net(_ConvBN, TorchModel):
        def init_weights(self):
            pass

        def __str__(self):
            return "ConvBN"

    g.GlobalSettings().set_device("cuda:0")
    gen = torch.Generator().manual_seed(5)
    X = torch.randn(200, 3, 16, 16, generator=gen)
    y = torch.randint(0, 10, (200,), generator=gen)

    def run(graphs, updates=3):
        g.GlobalSettings().cuda_graphs = graphs
        g.set_seed(11)
        torch.manual_seed(11)
        h = TorchModelHandler(Net(), torch.optim.SGD, {"lr": .05, "momentum": .9, "weight_decay": 1e-4},
                              torch.nn.CrossEntropyLoss(), local_epochs=1, batch_size=32)
        h.init()
        rows = []
        for _ in range(updates):
            h._update((X, y))
            rows.append(h.row.clone())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            h._update((X, y))
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 70 * 1e3
        return rows, ms
    for tf32 in (True, False):
        for det in (False, True):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cudnn.deterministic = det
            e1, ms_e = run(False)
            e2, _ = run(False)
            gr, ms_g = run(True)
            print(json.dumps({"cudnn_tf32": tf32, "cudnn_deterministic": det,
                              "eager_vs_eager": [float((a - b).abs().max()) for a, b in zip(e1, e2)],
                              "graph_vs_eager": [float((a - b).abs().max()) for a, b in zip(gr, e1)],
                              "ms_per_step_eager": round(ms_e, 4), "ms_per_step_graph": round(ms_g, 4)}))
    g.GlobalSettings().cuda_graphs = True
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.deterministic = False
    resnet_timing()


def resnet_timing():
    """BASELINE config 5's step (ResNet-20, batch 64, 3x32x32): ms per SGD step eager / graph, one node alone and eight
    nodes on their own streams (does the device overlap them?)."""
    import gossipy_b200 as g
    from gossipy_b200.engine import arena
    from gossipy_b200.model.handler import TorchModelHandler
    from gossipy_b200.models import ResNet20
    gen = torch.Generator().manual_seed(1)
    X = torch.randn(1024, 3, 32, 32, generator=gen)
    y = torch.randint(0, 10, (1024,), generator=gen)
    for graphs in (False, True):
        g.GlobalSettings().cuda_graphs = graphs
        hs = []
        for i in range(8):
            h = TorchModelHandler(ResNet20(10), torch.optim.SGD, {"lr": .05, "momentum": .9, "weight_decay": 1e-4},
                                  torch.nn.functional.cross_entropy, local_epochs=1, batch_size=64)
            h.owner = i
            h.init()
            hs.append(h)
        for h in hs:                       # warm-up + capture
            h._update((X, y))
        torch.cuda.synchronize()
        out = {"resnet20_batch64": True, "cuda_graphs": graphs}
        for k in (1, 8):
            t0 = time.perf_counter()
            for h in hs[:k]:
                h._update((X, y))
            t_host = time.perf_counter() - t0
            torch.cuda.synchronize()
            t = time.perf_counter() - t0
            out["ms_per_step_%d_nodes" % k] = round(t / (16 * k) * 1e3, 4)
            out["host_ms_per_step_%d_nodes" % k] = round(t_host / (16 * k) * 1e3, 4)
        print(json.dumps(out))
    g.GlobalSettings().cuda_graphs = True


if __name__ == "__main__":
    main()
